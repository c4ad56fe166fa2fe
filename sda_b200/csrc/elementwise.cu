// elementwise.cu -- HBM-bound kernels of the U-Net path: layout conversion, channel LayerNorm
// forward/backward, time-shift projection and weight packing.
//
// All kernels map one warp to one pixel and lanes to channels (c = lane + 32 j), so every global
// access of a warp is one contiguous 128 B (fp32) or 64 B (bf16) segment.
#include "common.cuh"

namespace sdab {

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kChanTile = 64;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline dim3 warp_grid(size_t n_warps) { return dim3((unsigned)((n_warps + kWarpsPerBlock - 1) / kWarpsPerBlock)); }

// ---------------------------------------------------------------------------------------------
// Planar fp32 images -> OP(Cpad): the entry of the network.  `Src::plane(n, c)` is the H x W plane that
// holds channel c of image n, or null for a zero channel:
//   NchwSrc       a contiguous (N, C, H, W) tensor, as ScoreUNet.forward hands the U-Net (sda/score.py:89-93);
//   WindowSrc     MCScoreNet.unfold + the context concat as addressing (sda/score.py:87,146-153): window i of
//                 trajectory b is the contiguous slab of frames i .. i + 2k, the context planes follow;
//   FoldAdjSrc    the adjoint of MCScoreNet.fold (sda/score.py:157-164) as addressing: slot s of window i holds
//                 the cotangent of frame j when (i, s) feeds j, and zero otherwise.
// Tile of 32 pixels along W through smem.
// ---------------------------------------------------------------------------------------------
struct NchwSrc {
  const float* x;
  int C;
  size_t HW;
  __device__ __forceinline__ const float* plane(int n, int c) const {
    return c < C ? x + ((size_t)n * C + c) * HW : nullptr;
  }
};

struct WindowSrc {
  const float* x;    // (B, L, C, H, W)
  const float* ctx;  // (Cc, H, W) or null
  WindowIO w;
  size_t HW;
  __device__ __forceinline__ const float* plane(int n, int c) const {
    const int nw = w.L - 2 * w.order, cw = (2 * w.order + 1) * w.C;
    const int wi = w.w_begin + n, b = wi / nw, i = wi % nw;
    if (c < cw) return x + (((size_t)b * w.L + i) * w.C + c) * HW;
    if (c < cw + w.Cc) return ctx + (size_t)(c - cw) * HW;
    return nullptr;
  }
};

struct FoldAdjSrc {
  const float* g;  // (B, L, C, H, W)
  WindowIO w;
  size_t HW;
  __device__ __forceinline__ const float* plane(int n, int c) const {
    const int k = w.order, nw = w.L - 2 * k;
    if (c >= (2 * k + 1) * w.C) return nullptr;
    const int wi = w.w_begin + n, b = wi / nw, i = wi % nw, slot = c / w.C, ch = c % w.C;
    int j = -1;
    if (slot == k)
      j = i + k;
    else if (i == 0 && slot < k)
      j = slot;
    else if (i == nw - 1 && slot > k)
      j = nw - 1 + slot;
    return j >= 0 ? g + (((size_t)b * w.L + j) * w.C + ch) * HW : nullptr;
  }
};

// block: 128 threads = one row segment of 32 pixels.  Phase 1: coalesced plane reads (a warp = 32 consecutive
// pixels of one channel) into smem; phase 2: one thread = 8 consecutive channels of one pixel -> one 16 B store per
// plane and halo replica (the 32 channels of a K-block are 64 contiguous bytes of the operand layout).
template <class Src>
__global__ void __launch_bounds__(128)
    pack_planes_kernel(const Src src, bf16* __restrict__ op, int N, int Cpad, int H, int W, int s2) {
  __shared__ float tile[kChanTile][33];
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int w0 = blockIdx.x * 32, h = blockIdx.y, n = blockIdx.z;
  const OpShape s{N, H, W, Cpad, s2};
  const size_t lo_off = s.lo_offset();
  for (int cb = 0; cb < Cpad; cb += kChanTile) {
    const int nc = min(kChanTile, Cpad - cb);
    __syncthreads();
    for (int c = wrp; c < nc; c += 4) {
      float v = 0.f;
      const float* pl = src.plane(n, cb + c);
      if (pl && w0 + lane < W) v = pl[(size_t)h * W + w0 + lane];
      tile[c][lane] = v;
    }
    __syncthreads();
    for (int task = tid; task < 32 * (nc >> 3); task += 128) {
      const int px = task & 31, g = task >> 5;
      const int w = w0 + px;
      if (w >= W) continue;
      const int c = g << 3;
      uint4 hi, lo;
      split_bf16x2(tile[c][px], tile[c + 1][px], hi.x, lo.x);
      split_bf16x2(tile[c + 2][px], tile[c + 3][px], hi.y, lo.y);
      split_bf16x2(tile[c + 4][px], tile[c + 5][px], hi.z, lo.z);
      split_bf16x2(tile[c + 6][px], tile[c + 7][px], hi.w, lo.w);
      const size_t blk = (size_t)((cb + c) >> 5) * s.block_stride() + ((cb + c) & 31);
      for_each_replica(h, w, H, W, [&](int hp, int wp) {
        bf16* dst = op + op_offset(s, n, hp, wp) + blk;
        *reinterpret_cast<uint4*>(dst) = hi;
        *reinterpret_cast<uint4*>(dst + lo_off) = lo;
      });
    }
  }
}

// F(Cpad) -> planar fp32 images: the exit of the network.  `Dst::plane(n, c)` is the destination plane of
// channel c of image n, or null when that channel is dropped:
//   NchwDst    a contiguous (N, Creal, H, W) tensor;
//   FoldDst    MCScoreNet.fold as addressing (sda/score.py:157-164): only the centre slot of every window and
//              the side slots of the first / last window of a trajectory are kept; the frames land either at
//              their place in the (B, L, C, H, W) score (cap == 0) or in this rank's shard of `cap` frames
//              (centre frames first, then 2k edge frames per trajectory the rank touches; sdab_frames_assemble).
struct NchwDst {
  float* x;
  int C;
  size_t HW;
  __device__ __forceinline__ float* plane(int n, int c) const { return c < C ? x + ((size_t)n * C + c) * HW : nullptr; }
};

struct FoldDst {
  float* out;
  WindowIO w;
  size_t HW;
  __device__ __forceinline__ float* plane(int n, int c) const {
    const int k = w.order, nw = w.L - 2 * k;
    if (c >= (2 * k + 1) * w.C) return nullptr;
    const int wi = w.w_begin + n, b = wi / nw, i = wi % nw, slot = c / w.C, ch = c % w.C;
    int j = -1;
    if (slot == k)
      j = i + k;
    else if (i == 0 && slot < k)
      j = slot;
    else if (i == nw - 1 && slot > k)
      j = nw - 1 + slot;
    if (j < 0) return nullptr;
    if (w.cap == 0) return out + (((size_t)b * w.L + j) * w.C + ch) * HW;
    const int pos = slot == k ? n : w.per + (b - w.w_begin / nw) * 2 * k + (slot < k ? slot : slot - 1);
    return out + ((size_t)pos * w.C + ch) * HW;
  }
};

// block: 128 threads = one row segment of 32 pixels.  Phase 1: one thread = 8 consecutive channels of one pixel
// (two 16 B loads); phase 2: coalesced plane writes (a warp = 32 consecutive pixels of one channel).
template <class Dst>
__global__ void __launch_bounds__(128)
    unpack_planes_kernel(const float* __restrict__ f, const Dst dst, int N, int Creal, int Cpad, int H, int W) {
  __shared__ float tile[kChanTile][33];
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int w0 = blockIdx.x * 32, h = blockIdx.y, n = blockIdx.z;
  for (int cb = 0; cb < Creal; cb += kChanTile) {
    const int nc = min(kChanTile, Cpad - cb);
    __syncthreads();
    for (int task = tid; task < 32 * (nc >> 3); task += 128) {
      const int px = task & 31, c = (task >> 5) << 3;
      const int w = w0 + px;
      if (w >= W) continue;
      const float4* src = reinterpret_cast<const float4*>(f + (((size_t)n * H + h) * W + w) * Cpad + cb + c);
      const float4 a = src[0], b = src[1];
      tile[c][px] = a.x, tile[c + 1][px] = a.y, tile[c + 2][px] = a.z, tile[c + 3][px] = a.w;
      tile[c + 4][px] = b.x, tile[c + 5][px] = b.y, tile[c + 6][px] = b.z, tile[c + 7][px] = b.w;
    }
    __syncthreads();
    for (int c = wrp; c < nc && cb + c < Creal; c += 4) {
      float* pl = dst.plane(n, cb + c);
      if (pl && w0 + lane < W) pl[(size_t)h * W + w0 + lane] = tile[c][lane];
    }
  }
}

// F(C) -> OP(C).  kind 0: normal; 1: S2 parity layout; 2: zero-insertion x2 (f is at H/2 x W/2).
// One thread converts 8 consecutive channels of one pixel: two 16 B loads, one 16 B store per plane
// (and per halo replica) -- the 32 channels of a K-block are 64 contiguous bytes in either layout.
__global__ void f_to_operand_kernel(const float* __restrict__ f, bf16* __restrict__ op, int N, int H, int W, int C,
                                    int kind) {
  const int groups = C >> 3;
  const size_t total = (size_t)N * H * W * groups;
  const OpShape s{N, H, W, C, kind == 1};
  const size_t lo_off = s.lo_offset();
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(t % groups);
    const size_t pix = t / groups;
    const int w = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((size_t)W * H));
    const int c = g << 3;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* src = nullptr;
    if (kind == 2) {
      if (!(h & 1) && !(w & 1)) src = f + (((size_t)n * (H / 2) + h / 2) * (W / 2) + w / 2) * C + c;
    } else {
      src = f + pix * C + c;
    }
    if (src) {
      const float4 a = reinterpret_cast<const float4*>(src)[0], b = reinterpret_cast<const float4*>(src)[1];
      v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    }
    uint4 hi, lo;
    split_bf16x2(v[0], v[1], hi.x, lo.x);
    split_bf16x2(v[2], v[3], hi.y, lo.y);
    split_bf16x2(v[4], v[5], hi.z, lo.z);
    split_bf16x2(v[6], v[7], hi.w, lo.w);
    const size_t blk = (size_t)(c >> 5) * s.block_stride() + (c & 31);
    for_each_replica(h, w, H, W, [&](int hp, int wp) {
      bf16* dst = op + op_offset(s, n, hp, wp) + blk;
      *reinterpret_cast<uint4*>(dst) = hi;
      *reinterpret_cast<uint4*>(dst + lo_off) = lo;
    });
  }
}

// ---------------------------------------------------------------------------------------------
// Operand tiles.  The LayerNorm kernels work on a tile of TP consecutive pixels of one image row,
// staged in shared memory as tile[(q * TP + p) * 32 + c] (q = plane * C/32 + K-block): in the
// operand tensor one (q, row) run of TP pixels is TP * 64 contiguous bytes, so tiles move with
// 16 B accesses that are contiguous along the pixel axis.
// ---------------------------------------------------------------------------------------------
constexpr int kLnThreads = 256;

// Writes the tile of row h, columns [w0, w0 + TP) into the operand tensor `op` (shape s, normal
// layout) with its halo replicas.  up != 0: nearest x2 upsample (s is the upsampled shape).
__device__ __forceinline__ void store_operand_tile(const bf16* tile, bf16* __restrict__ op, const OpShape& s, int n,
                                                   int h, int w0, int TP, int up, int tid) {
  const int Hd = s.H, Wd = s.W, Q = 2 * (s.C / 32);
  int rows[4], nrows = 0;
  for (int d = (up ? 2 * h : h); d <= (up ? 2 * h + 1 : h); ++d) {
    rows[nrows++] = d + 1;
    if (d == 0) rows[nrows++] = Hd + 1;
    if (d == Hd - 1) rows[nrows++] = 0;
  }
  const int DP = up ? 2 * TP : TP, wd0 = up ? 2 * w0 : w0;
  const bool has_last = wd0 + DP == Wd, has_first = wd0 == 0;
  const int slots = DP + 2;  // slot DP: left halo (column Wd-1 -> wp 0); slot DP+1: right halo (column 0 -> wp Wd+1)
  const int total = nrows * Q * slots * 4;
  const size_t bs = s.block_stride();
  for (int idx = tid; idx < total; idx += kLnThreads) {
    const int part = idx & 3;
    int t = idx >> 2;
    const int e = t % slots;
    t /= slots;
    const int q = t % Q, r = t / Q;
    int wp, src;
    if (e < DP) {
      wp = wd0 + e + 1, src = up ? e >> 1 : e;
    } else if (e == DP) {
      if (!has_last) continue;
      wp = 0, src = TP - 1;
    } else {
      if (!has_first) continue;
      wp = Wd + 1, src = 0;
    }
    const uint4 v = *reinterpret_cast<const uint4*>(tile + ((size_t)q * TP + src) * 32 + part * 8);
    *reinterpret_cast<uint4*>(op + op_offset(s, n, rows[r], wp) + q * bs + part * 8) = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Channel LayerNorm of (x + shift): zuko.nn.LayerNorm(dim=-3) as used at sda/nn.py:137,163 --
// (u - mean_C u) / sqrt(var_C,unbiased(u) + 1e-5), no affine -- applied to u = x + project(y)
// (ModResidualBlock.forward, sda/nn.py:27-28).  Output: bf16 hi/lo operand of the next conv,
// optionally nearest-upsampled x2 (the tails, sda/nn.py:161-170).
// grid: (W / TP, H, N); one warp per pixel for the statistics (lane <-> channel lane + 32 j).
// ---------------------------------------------------------------------------------------------
template <int MAXJ>
__global__ void __launch_bounds__(kLnThreads)
    ln_forward_kernel(const float* __restrict__ x, const float* __restrict__ shift, int shift_stride, int Nt,
                      bf16* __restrict__ op, float* __restrict__ rstd, int N, int H, int W, int C, int upsample,
                      int TP) {
  extern __shared__ __align__(16) bf16 ln_tile[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int w0 = blockIdx.x * TP, h = blockIdx.y, n = blockIdx.z;
  const float* sh = shift ? shift + (size_t)(Nt > 1 ? n : 0) * shift_stride : nullptr;
  const int nj = C / 32;
  for (int p = wid; p < TP; p += kLnThreads / 32) {
    const size_t pix = ((size_t)n * H + h) * W + w0 + p;
    const float* src = x + pix * C;
    float u[MAXJ];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      if (j < nj) {
        const int c = lane + 32 * j;
        u[j] = src[c] + (sh ? __ldg(sh + c) : 0.f);
        sum += u[j];
      }
    }
    const float mean = warp_sum(sum) / C;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      if (j < nj) {
        u[j] -= mean;
        sq += u[j] * u[j];
      }
    }
    const float var = warp_sum(sq) / (C - 1);
    const float r = 1.f / sqrtf(var + 1e-5f);
    if (rstd && lane == 0) rstd[pix] = r;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      if (j < nj) {
        bf16 hi, lo;
        split_bf16(u[j] * r, hi, lo);
        ln_tile[((size_t)j * TP + p) * 32 + lane] = hi;
        ln_tile[((size_t)(nj + j) * TP + p) * 32 + lane] = lo;
      }
    }
  }
  __syncthreads();
  const OpShape s{N, upsample ? 2 * H : H, upsample ? 2 * W : W, C, 0};
  store_operand_tile(ln_tile, op, s, n, h, w0, TP, upsample, threadIdx.x);
}

// Backward of the channel LayerNorm:
//   gu = (ga - mean_C(ga) - a * sum_C(ga * a) / (C - 1)) * rstd ;  gx = res + gu
// a is re-read from the saved operand (hi + lo).  pooled: ga is given at 2H x 2W and summed over
// the 2x2 block first (adjoint of the nearest upsample of the tails), and the operand holding a is
// the upsampled one (read at (2h, 2w)).
template <int MAXJ>
__global__ void __launch_bounds__(kLnThreads)
    ln_backward_kernel(const float* __restrict__ ga, const bf16* __restrict__ a_op, const float* __restrict__ rstd,
                       const float* __restrict__ res, float* __restrict__ gxF, bf16* __restrict__ gxOP, int N, int H,
                       int W, int C, int pooled, int TP) {
  extern __shared__ __align__(16) bf16 ln_tile[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int w0 = blockIdx.x * TP, h = blockIdx.y, n = blockIdx.z;
  const int nj = C / 32, Q = 2 * nj;
  // phase 0: the saved normalised activations of the tile -> shared memory
  {
    const OpShape sa{N, pooled ? 2 * H : H, pooled ? 2 * W : W, C, 0};
    const size_t bs = sa.block_stride();
    const int hp = (pooled ? 2 * h : h) + 1;
    for (int idx = threadIdx.x; idx < Q * TP * 4; idx += kLnThreads) {
      const int part = idx & 3, p = (idx >> 2) % TP, q = (idx >> 2) / TP;
      const int wp = (pooled ? 2 * (w0 + p) : w0 + p) + 1;
      *reinterpret_cast<uint4*>(ln_tile + ((size_t)q * TP + p) * 32 + part * 8) =
          *reinterpret_cast<const uint4*>(a_op + op_offset(sa, n, hp, wp) + q * bs + part * 8);
    }
  }
  __syncthreads();
  for (int p = wid; p < TP; p += kLnThreads / 32) {
    const int w = w0 + p;
    const size_t pix = ((size_t)n * H + h) * W + w;
    float g[MAXJ], a[MAXJ];
    float sg = 0.f, sga = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      if (j < nj) {
        const int c = lane + 32 * j;
        if (pooled) {
          const size_t W2 = 2 * (size_t)W;
          const float* q = ga + (((size_t)n * 2 * H + 2 * h) * W2 + 2 * w) * C + c;
          g[j] = (q[0] + q[C]) + (q[W2 * C] + q[W2 * C + C]);
        } else {
          g[j] = ga[pix * C + c];
        }
        a[j] = __bfloat162float(ln_tile[((size_t)j * TP + p) * 32 + lane]) +
               __bfloat162float(ln_tile[((size_t)(nj + j) * TP + p) * 32 + lane]);
        sg += g[j];
        sga += g[j] * a[j];
      }
    }
    sg = warp_sum(sg) / C;
    sga = warp_sum(sga) / (C - 1);
    const float r = rstd[pix];
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      if (j < nj) {
        const int c = lane + 32 * j;
        float v = (g[j] - sg - a[j] * sga) * r;
        if (res) v += res[pix * C + c];
        if (gxF) gxF[pix * C + c] = v;
        bf16 hi, lo;
        split_bf16(v, hi, lo);
        ln_tile[((size_t)j * TP + p) * 32 + lane] = hi;
        ln_tile[((size_t)(nj + j) * TP + p) * 32 + lane] = lo;
      }
    }
  }
  if (!gxOP) return;
  __syncthreads();
  const OpShape so{N, H, W, C, 0};
  store_operand_tile(ln_tile, gxOP, so, n, h, w0, TP, 0, threadIdx.x);
}

// ---------------------------------------------------------------------------------------------
// All `project` Linears of the network in one launch (sda/nn.py:132-135): out[nt][r] =
// pb[r] + sum_m pw[r][m] y[nt][m] over the concatenated rows r of every block.
// ---------------------------------------------------------------------------------------------
__global__ void time_shifts_kernel(const float* __restrict__ y, const float* __restrict__ pw,
                                   const float* __restrict__ pb, float* __restrict__ out, int Nt, int rows, int mod) {
  const size_t warp = (size_t)blockIdx.x * kWarpsPerBlock + threadIdx.y;
  if (warp >= (size_t)Nt * rows) return;
  const int r = warp % rows, nt = warp / rows;
  float acc = 0.f;
  for (int m = threadIdx.x; m < mod; m += 32) acc += pw[(size_t)r * mod + m] * y[(size_t)nt * mod + m];
  acc = warp_sum(acc);
  if (threadIdx.x == 0) out[(size_t)nt * rows + r] = acc + pb[r];
}

// ---------------------------------------------------------------------------------------------
// Weight packing: (Cout, Cin, 3, 3) fp32 ->
//   fwd[tap][chunk][plane][co][kc]  = split(W[co][32 chunk + kc][a][b]),      tap = 3a + b
//   bwd[tap][chunk][plane][ci][kc]  = split(W[32 chunk + kc][ci][2-a][2-b])   (transposed, flipped:
//        the input-gradient of a circular stride-1 correlation is the correlation of the cotangent
//        with this kernel)
// Rows and K are zero padded to multiples of 32 (so that every convolution, including the 10-channel
// final one and the 11-channel head transpose, qualifies for the CTA-pair patch kernel).
// ---------------------------------------------------------------------------------------------
// element i of the packed forward + transposed copies of one convolution
__device__ __forceinline__ void pack_conv_element(const float* __restrict__ w, bf16* __restrict__ fwd,
                                                  bf16* __restrict__ bwd, int Cout, int Cin, size_t i) {
  const int Kf = (Cin + 31) / 32 * 32, Nf = (Cout + 31) / 32 * 32;
  const int Kb = (Cout + 31) / 32 * 32, Nb = (Cin + 31) / 32 * 32;
  const size_t nf = (size_t)9 * Kf * Nf;
  {
    const bool is_b = i >= nf;
    size_t j = is_b ? i - nf : i;
    const int K = is_b ? Kb : Kf, Nn = is_b ? Nb : Nf;
    const int kc = j % 32;
    j /= 32;
    const int row = j % Nn;
    j /= Nn;
    const int chunk = j % (K / 32);
    const int tap = j / (K / 32);
    const int a = tap / 3, b = tap % 3;
    const int k = chunk * 32 + kc;
    float v = 0.f;
    if (!is_b) {
      if (row < Cout && k < Cin) v = w[(((size_t)row * Cin + k) * 3 + a) * 3 + b];
    } else {
      if (row < Cin && k < Cout) v = w[(((size_t)k * Cin + row) * 3 + (2 - a)) * 3 + (2 - b)];
    }
    bf16 hi, lo;
    split_bf16(v, hi, lo);
    bf16* dst = is_b ? bwd : fwd;
    const size_t base = (((size_t)tap * (K / 32) + chunk) * 2) * Nn * 32;
    dst[base + (size_t)row * 32 + kc] = hi;
    dst[base + (size_t)Nn * 32 + (size_t)row * 32 + kc] = lo;
  }
}

__host__ __device__ inline size_t pack_conv_elements(int Cout, int Cin) {
  const size_t Kf = (Cin + 31) / 32 * 32, Nf = (Cout + 31) / 32 * 32;
  return 2 * 9 * Kf * Nf;  // forward + transposed copy (same padded extent)
}

__global__ void pack_conv_weights_kernel(const float* __restrict__ w, bf16* __restrict__ fwd, bf16* __restrict__ bwd,
                                         int Cout, int Cin) {
  const size_t n = pack_conv_elements(Cout, Cin);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    pack_conv_element(w, fwd, bwd, Cout, Cin, i);
}

// All convolutions of a network in ONE launch (a training step repacks every weight after every optimizer step:
// 42 launches + 120 small copies were 0.8 ms of a 12 ms iteration).  The job of an element is found by binary search
// over the prefix sums of the table, which travels as a kernel parameter.
__global__ void pack_conv_weights_batched_kernel(const __grid_constant__ PackTable t) {
  const size_t total = t.start[t.n];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = t.n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (t.start[mid] <= i)
        lo = mid;
      else
        hi = mid - 1;
    }
    pack_conv_element(t.w[lo], t.fwd[lo], t.bwd[lo], t.cout[lo], t.cin[lo], i - t.start[lo]);
  }
}

// dst[j] = j < n_src ? src[j] : 0 for j < n_dst, for every entry of the table (biases with their zero padding, the
// projection weights and biases)
__global__ void copy_pad_batched_kernel(const __grid_constant__ CopyTable t) {
  const size_t total = t.start[t.n];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = t.n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (t.start[mid] <= i)
        lo = mid;
      else
        hi = mid - 1;
    }
    const size_t j = i - t.start[lo];
    t.dst[lo][j] = j < (size_t)t.n_src[lo] ? t.src[lo][j] : 0.f;
  }
}

// Combined weights of a tail (LayerNorm -> nearest x2 -> 3x3 conv, sda/nn.py:161-170):
//   tf[4 (po, pp) + 2 th + tw][chunk][plane][co][kc]: sub-pixel form -- output parity (po, pp) is a 2x2-tap
//       conv of the LOW-resolution input with taps summed over S(po, th) x S(pp, tw),
//       S(0,0) = {0}, S(0,1) = {1,2}, S(1,0) = {0,1}, S(1,1) = {2};
//   tb[4 a + b][chunk][plane][ci][kc]: its transpose -- a 4x4 stride-2 conv of the padded high-resolution
//       cotangent with taps summed over R(a) x R(b), R = {2}, {1,2}, {0,1}, {0}, channels transposed.
__global__ void pack_tail_weights_kernel(const float* __restrict__ w, bf16* __restrict__ tf, bf16* __restrict__ tb,
                                         int Cout, int Cin) {
  const size_t nf = (size_t)16 * Cin * Cout, nb = nf;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nf + nb; i += (size_t)gridDim.x * blockDim.x) {
    const bool is_b = i >= nf;
    size_t j = is_b ? i - nf : i;
    const int K = is_b ? Cout : Cin, Nn = is_b ? Cin : Cout;
    const int kc = j % 32;
    j /= 32;
    const int row = j % Nn;
    j /= Nn;
    const int chunk = j % (K / 32);
    const int tap = j / (K / 32);
    const int k = chunk * 32 + kc;
    const int co = is_b ? k : row, ci = is_b ? row : k;
    // row / column tap sets as [first, last] ranges
    int a0, a1, b0, b1;
    if (!is_b) {
      const int po = tap >> 3, pp = (tap >> 2) & 1, th = (tap >> 1) & 1, tw = tap & 1;
      a0 = po ? (th ? 2 : 0) : (th ? 1 : 0), a1 = po ? (th ? 2 : 1) : (th ? 2 : 0);
      b0 = pp ? (tw ? 2 : 0) : (tw ? 1 : 0), b1 = pp ? (tw ? 2 : 1) : (tw ? 2 : 0);
    } else {
      const int a = tap >> 2, b = tap & 3;
      a0 = a == 0 ? 2 : (a == 1 ? 1 : 0), a1 = a == 0 ? 2 : (a == 1 ? 2 : (a == 2 ? 1 : 0));
      b0 = b == 0 ? 2 : (b == 1 ? 1 : 0), b1 = b == 0 ? 2 : (b == 1 ? 2 : (b == 2 ? 1 : 0));
    }
    float v = 0.f;
    for (int a = a0; a <= a1; ++a)
      for (int b = b0; b <= b1; ++b) v += w[(((size_t)co * Cin + ci) * 3 + a) * 3 + b];
    bf16 hi, lo;
    split_bf16(v, hi, lo);
    bf16* dst = is_b ? tb : tf;
    const size_t base = (((size_t)tap * (K / 32) + chunk) * 2) * Nn * 32;
    dst[base + (size_t)row * 32 + kc] = hi;
    dst[base + (size_t)Nn * 32 + (size_t)row * 32 + kc] = lo;
  }
}

__global__ void copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

}  // namespace

int pack_nchw_to_op(const float* x, bf16* op, int N, int Creal, int Cpad, int H, int W, int s2, cudaStream_t st) {
  dim3 grid((W + 31) / 32, H, N), block(128);
  pack_planes_kernel<<<grid, block, 0, st>>>(NchwSrc{x, Creal, (size_t)H * W}, op, N, Cpad, H, W, s2);
  SDAB_LAUNCH_CHECK("pack_planes_kernel");
  return SDAB_OK;
}

int pack_windows_to_op(const float* x, const float* ctx, bf16* op, const WindowIO& w, int N, int Cpad, int H, int W,
                       cudaStream_t st) {
  dim3 grid((W + 31) / 32, H, N), block(128);
  pack_planes_kernel<<<grid, block, 0, st>>>(WindowSrc{x, ctx, w, (size_t)H * W}, op, N, Cpad, H, W, 0);
  SDAB_LAUNCH_CHECK("pack_planes_kernel");
  return SDAB_OK;
}

int pack_fold_adjoint_to_op(const float* g, bf16* op, const WindowIO& w, int N, int Cpad, int H, int W,
                            cudaStream_t st) {
  dim3 grid((W + 31) / 32, H, N), block(128);
  pack_planes_kernel<<<grid, block, 0, st>>>(FoldAdjSrc{g, w, (size_t)H * W}, op, N, Cpad, H, W, 0);
  SDAB_LAUNCH_CHECK("pack_planes_kernel");
  return SDAB_OK;
}

int unpack_f_to_nchw(const float* f, float* x, int N, int Creal, int Cpad, int H, int W, cudaStream_t st) {
  dim3 grid((W + 31) / 32, H, N), block(128);
  unpack_planes_kernel<<<grid, block, 0, st>>>(f, NchwDst{x, Creal, (size_t)H * W}, N, Creal, Cpad, H, W);
  SDAB_LAUNCH_CHECK("unpack_planes_kernel");
  return SDAB_OK;
}

int unpack_f_fold(const float* f, float* out, const WindowIO& w, int N, int Cpad, int H, int W, cudaStream_t st) {
  dim3 grid((W + 31) / 32, H, N), block(128);
  unpack_planes_kernel<<<grid, block, 0, st>>>(f, FoldDst{out, w, (size_t)H * W}, N, (2 * w.order + 1) * w.C, Cpad, H,
                                               W);
  SDAB_LAUNCH_CHECK("unpack_planes_kernel");
  return SDAB_OK;
}

int f_to_operand(const float* f, bf16* op, int N, int H, int W, int C, int kind, cudaStream_t st) {
  const size_t threads = (size_t)N * H * W * (C / 8);
  const size_t blocks = (threads + 255) / 256;
  f_to_operand_kernel<<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, st>>>(f, op, N, H, W, C, kind);
  SDAB_LAUNCH_CHECK("f_to_operand_kernel");
  return SDAB_OK;
}

int ln_forward(const float* x, const float* shift, int shift_stride, int Nt, bf16* op, float* rstd, int N, int H, int W,
               int C, int upsample, cudaStream_t st) {
  SDAB_REQUIRE(C % 32 == 0 && C <= 512, "LayerNorm channels must be a multiple of 32, at most 512");
  const int TP = W < 32 ? W : (C > 384 ? 16 : 32);
  SDAB_REQUIRE(W % TP == 0, "image width must be a multiple of the LayerNorm tile");
  const dim3 grid(W / TP, H, N);
  const size_t smem = (size_t)4 * C * TP;
  if (C <= 128)
    ln_forward_kernel<4><<<grid, kLnThreads, smem, st>>>(x, shift, shift_stride, Nt, op, rstd, N, H, W, C, upsample, TP);
  else if (C <= 256)
    ln_forward_kernel<8><<<grid, kLnThreads, smem, st>>>(x, shift, shift_stride, Nt, op, rstd, N, H, W, C, upsample, TP);
  else
    ln_forward_kernel<16><<<grid, kLnThreads, smem, st>>>(x, shift, shift_stride, Nt, op, rstd, N, H, W, C, upsample, TP);
  SDAB_LAUNCH_CHECK("ln_forward_kernel");
  return SDAB_OK;
}

int ln_backward(const float* ga, const bf16* a_op, const float* rstd, const float* res, float* gxF, bf16* gxOP, int N,
                int H, int W, int C, int pooled, cudaStream_t st) {
  SDAB_REQUIRE(C % 32 == 0 && C <= 512, "LayerNorm channels must be a multiple of 32, at most 512");
  const int TP = W < 32 ? W : (C > 384 ? 16 : 32);
  SDAB_REQUIRE(W % TP == 0, "image width must be a multiple of the LayerNorm tile");
  const dim3 grid(W / TP, H, N);
  const size_t smem = (size_t)4 * C * TP;
  if (C <= 128)
    ln_backward_kernel<4><<<grid, kLnThreads, smem, st>>>(ga, a_op, rstd, res, gxF, gxOP, N, H, W, C, pooled, TP);
  else if (C <= 256)
    ln_backward_kernel<8><<<grid, kLnThreads, smem, st>>>(ga, a_op, rstd, res, gxF, gxOP, N, H, W, C, pooled, TP);
  else
    ln_backward_kernel<16><<<grid, kLnThreads, smem, st>>>(ga, a_op, rstd, res, gxF, gxOP, N, H, W, C, pooled, TP);
  SDAB_LAUNCH_CHECK("ln_backward_kernel");
  return SDAB_OK;
}

int time_shifts(const float* y, const float* pw, const float* pb, float* out, int Nt, int rows, int mod,
                cudaStream_t st) {
  time_shifts_kernel<<<warp_grid((size_t)Nt * rows), dim3(32, kWarpsPerBlock), 0, st>>>(y, pw, pb, out, Nt, rows, mod);
  SDAB_LAUNCH_CHECK("time_shifts_kernel");
  return SDAB_OK;
}

int pack_conv_weights(const float* w, bf16* fwd, bf16* bwd, int Cout, int Cin, cudaStream_t st) {
  pack_conv_weights_kernel<<<296, 256, 0, st>>>(w, fwd, bwd, Cout, Cin);
  SDAB_LAUNCH_CHECK("pack_conv_weights_kernel");
  return SDAB_OK;
}

int pack_conv_weights_batched(PackTable& t, cudaStream_t st) {
  SDAB_REQUIRE(t.n >= 1 && t.n <= kMaxPack, "pack table out of range");
  t.start[0] = 0;
  for (int i = 0; i < t.n; ++i) t.start[i + 1] = t.start[i] + pack_conv_elements(t.cout[i], t.cin[i]);
  pack_conv_weights_batched_kernel<<<148 * 8, 256, 0, st>>>(t);
  SDAB_LAUNCH_CHECK("pack_conv_weights_batched_kernel");
  return SDAB_OK;
}

int copy_pad_batched(CopyTable& t, cudaStream_t st) {
  SDAB_REQUIRE(t.n >= 1 && t.n <= kMaxCopy, "copy table out of range");
  t.start[0] = 0;
  for (int i = 0; i < t.n; ++i) t.start[i + 1] = t.start[i] + (unsigned long long)t.n_dst[i];
  const size_t total = t.start[t.n];
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  copy_pad_batched_kernel<<<grid, 256, 0, st>>>(t);
  SDAB_LAUNCH_CHECK("copy_pad_batched_kernel");
  return SDAB_OK;
}

int pack_tail_weights(const float* w, bf16* tf, bf16* tb, int Cout, int Cin, cudaStream_t st) {
  SDAB_REQUIRE(Cout % 32 == 0 && Cin % 32 == 0, "tail channels must be multiples of 32");
  pack_tail_weights_kernel<<<296, 256, 0, st>>>(w, tf, tb, Cout, Cin);
  SDAB_LAUNCH_CHECK("pack_tail_weights_kernel");
  return SDAB_OK;
}

int copy_f32(const float* src, float* dst, size_t n, cudaStream_t st) {
  SDAB_CUDA_CHECK(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return SDAB_OK;
}

int fill_zero(void* p, size_t bytes, cudaStream_t st) {
  SDAB_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, st));
  return SDAB_OK;
}

}  // namespace sdab
