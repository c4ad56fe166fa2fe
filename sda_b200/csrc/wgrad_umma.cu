// wgrad_umma.cu -- tcgen05 weight gradient of the stride-1 3x3 circular convolutions (the 36 block
// convolutions of the U-Net, sda/nn.py:131-142): the parameter-gradient half of loss.backward() in the
// training step (sda/utils.py:136-143).
//
//   dW[tap][ci][co] = sum_{pixels} x[pixel + tap][ci] * g[pixel][co]
//
//   GEMM view    per tap, M = 128 input channels (four 32-channel chunks), N = NBLK output channels,
//                K = the pixels of an 8 x 16 (bf16) or 8 x 8 (bf16x3) tile, accumulated over every tile of the
//                CTA's pixel split.
//   operands     BOTH arrive as [pixel][32 channels] tiles straight from the operand tensors (OP layout):
//                the channels are the M / N dimension, so the descriptors are MN-major SWIZZLE_64B
//                (leading-dimension offset = chunk stride, stride offset = 8-pixel group stride; verified by
//                tools/probes/umma_mnmajor_probe.cu).  x is ONE haloed 10 x (BH + 2) patch per chunk: tap (a, b)
//                shifts the descriptor start by (10 a + b) pixel rows, as in the forward patch kernel.
//   accumulators TG taps x NBLK fp32 TMEM columns stay resident for the whole kernel; the epilogue runs once
//                and adds the CTA's partial sums into dW (cout, cin, 3, 3) with fp32 atomics.
//   precision    PLANES = 2: hi*hi + hi*lo + lo*hi like the forward path; PLANES = 1: bf16 single pass.
//   schedule     one CTA per (tap group, M block, N block, pixel split); warp 0 = TMA producer, warp 1 = MMA
//                issuer + TMEM allocator, warps 2-5 = epilogue.
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace sdab {

namespace {

constexpr int kThreads = 192;
// Pixel tile: 8 x BH pixels, BH = 16 (K = 128 per tile) or 8 (K = 64: half-size stages, so that the two
// planes of the bf16x3 mode still leave room for a second pipeline stage).
constexpr int kBW = 8, kPW = kBW + 2;
template <int BH>
struct Tile {
  static constexpr int kPH = BH + 2;                                        // haloed patch rows
  static constexpr uint32_t kPatchBytes = kPW * kPH * 64;                   // landing per (plane, chunk) of x
  static constexpr uint32_t kPatchSlot = (kPatchBytes + 1023u) & ~1023u;    // chunk stride of x in shared memory
  static constexpr uint32_t kGBytes = kBW * BH * 64;                        // per (plane, chunk) of g
};
constexpr uint32_t kSmemBudget = 227 * 1024;

struct WParams {
  int num_tiles, tiles_w, tiles_h;
  int nchunk_x, nchunk_g;  // 32-channel chunks of the x / g tensors
  int planes;
  int nblk;                // output channels per CTA (multiple of 32, <= 128)
  int tg;                  // taps per CTA
  int ntg, nmb, nnb;       // tap groups, M blocks, N blocks
  int splits;
  int stages;
  uint32_t stage_bytes, x_plane_bytes, g_plane_bytes;
  int cin, cout;
  float* dw;
  float* part;             // split partial sums [CTA][tg][128 ci][nblk co], or null: fp32 atomics straight into dw
  int ntl;                 // tap-list entries
  uint32_t tl_shift[9];    // patch-row shift of entry e (pixels)
  uint32_t tl_mask[9];     // weight taps entry e adds into
  int x_par, g_par, g_dh, g_dw;  // parity-image selection (WgradProblem)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded: a protocol bug aborts the launch instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}
// the epilogue's single wait spans the CTA's whole main loop: same idea, a bound of seconds
__device__ __forceinline__ void mbar_wait_long(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, %1;\n@px mov.s32 %0, 1;\n}" : "+r"(pred) : "r"(0xffffffffu));
  return pred != 0;
}
// MN-major SWIZZLE_64B descriptor: start address, leading-dimension offset (chunk stride), stride offset
// (8-pixel group stride), version 1, layout type 4
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}

template <int PLANES, int BH>
__global__ void __launch_bounds__(kThreads, 1)
    wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const WParams p) {
  constexpr uint32_t kPatchBytes = Tile<BH>::kPatchBytes, kPatchSlot = Tile<BH>::kPatchSlot, kGBytes = Tile<BH>::kGBytes;
  constexpr int kBH = BH;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bar_full = base, bar_empty = base + 32, bar_done = base + 64;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 96);
  const uint32_t stage0 = base + 1024;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item of this CTA
  int id = blockIdx.x;
  const int split = id % p.splits;
  id /= p.splits;
  const int nb = id % p.nnb;
  id /= p.nnb;
  const int mb = id % p.nmb;
  const int tgi = id / p.nmb;
  const int tap0 = tgi * p.tg, ntap = min(p.tg, p.ntl - tap0);  // entries of the tap list handled here
  const int xchunks = min(4, p.nchunk_x - 4 * mb), gchunks = p.nblk / 32;
  const int ntiles = split < p.num_tiles ? (p.num_tiles - split + p.splits - 1) / p.splits : 0;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + 96), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();  // nothing above touches global memory
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    const uint32_t tx = (uint32_t)PLANES * ((uint32_t)xchunks * kPatchBytes + (uint32_t)gchunks * kGBytes);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < ntiles; ++i) {
      const int tile = split + i * p.splits;
      const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, n = tile / (p.tiles_w * p.tiles_h);
      const int h0 = th * kBH, w0 = tw * kBW;
      mbar_wait(bar_empty + 8 * stage, phase ^ 1);
      if (elect_one()) {
        const uint32_t full = bar_full + 8 * stage;
        mbar_expect_tx(full, tx);
        const uint32_t sx = stage0 + stage * p.stage_bytes, sg = sx + PLANES * p.x_plane_bytes;
#pragma unroll
        for (int pl = 0; pl < PLANES; ++pl) {
          for (int c = 0; c < xchunks; ++c) {
            const int q = pl * p.nchunk_x + 4 * mb + c;
            tma_load_5d(sx + pl * p.x_plane_bytes + c * kPatchSlot, &tmX, full, 0, w0, h0, p.x_par ? 4 * q + p.x_par - 1 : q, n);
          }
          for (int c = 0; c < gchunks; ++c) {
            const int q = pl * p.nchunk_g + nb * gchunks + c;
            tma_load_5d(sg + pl * p.g_plane_bytes + c * kGBytes, &tmG, full, 0, w0 + p.g_dw, h0 + p.g_dh,
                        p.g_par ? 4 * q + p.g_par - 1 : q, n);
          }
        }
      }
      __syncwarp();
      if (++stage == p.stages) stage = 0, phase ^= 1;
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // D fp32, A / B bf16, both MN-major (bits 15, 16), M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.nblk >> 3) << 17) |
                           ((128u >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < ntiles; ++i) {
      mbar_wait(bar_full + 8 * stage, phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sx = stage0 + stage * p.stage_bytes, sg = sx + PLANES * p.x_plane_bytes;
        for (int t = 0; t < ntap; ++t) {
          const uint32_t shift = p.tl_shift[tap0 + t] * 64u;
          const uint32_t d = tmem_base + (uint32_t)(t * p.nblk);
#pragma unroll
          for (int pass = 0; pass < (PLANES == 2 ? 3 : 1); ++pass) {
            // pass 0: hi*hi, 1: x hi * g lo, 2: x lo * g hi
            const uint32_t xa = sx + (pass == 2 ? p.x_plane_bytes : 0u) + shift;
            const uint32_t ga = sg + (pass == 1 ? p.g_plane_bytes : 0u);
#pragma unroll
            for (int ks = 0; ks < BH / 2; ++ks)  // K = 16 pixels = two rows of the 8-wide tile
              umma_bf16(d, desc_mn(xa + ks * 2 * (kPW * 64), kPatchSlot, kPW * 64), desc_mn(ga + ks * 1024, kGBytes, 512),
                        idesc, (i | pass | ks) ? 1u : 0u);
          }
        }
        umma_commit(bar_empty + 8 * stage);
        if (i == ntiles - 1) umma_commit(bar_done);
      }
      __syncwarp();
      if (++stage == p.stages) stage = 0, phase ^= 1;
    }
  } else if (ntiles > 0) {
    // ===================================================================== epilogue (warps 2..5), once
    const int q = warp & 3;
    const int ci = 128 * mb + q * 32 + lane;
    mbar_wait_long(bar_done, 0);
    tc_fence_after();
    const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int t = 0; t < ntap; ++t) {
      const uint32_t mask = p.tl_mask[tap0 + t];
      for (int c0 = 0; c0 < p.nblk; c0 += 32) {
        float v[32];
        tmem_ld32(t0 + (uint32_t)(t * p.nblk + c0), v);
        if (p.part) {
          // this CTA's partial sums leave as full lines (a lane owns one input channel = one row of nblk
          // floats); wgrad_reduce_kernel sums the pixel splits in fixed order
          float4* dst = reinterpret_cast<float4*>(p.part + (((size_t)blockIdx.x * p.tg + t) * 128 + (q * 32 + lane)) * p.nblk + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else if (ci < p.cin) {
          for (int tap = 0; tap < 9; ++tap) {
            if (!(mask >> tap & 1u)) continue;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int co = nb * p.nblk + c0 + j;
              if (co < p.cout) atomicAdd(p.dw + ((size_t)co * p.cin + ci) * 9 + tap, v[j]);
            }
          }
        }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// dw[co][ci][tap] += sum over the pixel splits (fixed order: deterministic) of the partial sums the CTAs of one
// wgrad_umma_kernel launch left behind, for every weight tap the tap-list entry feeds.
// A block owns 32 output x 2 input channels and ALL entries of the tap list, one thread per (entry, ci, co): it sums
// the splits of its element (partials are read along co: full lines; thousands of threads keep the L2 latency
// covered even at 74 splits) and drops the sum into a shared-memory image of the block's piece of dw, which then
// leaves as rows of 2 x 9 = 18 consecutive floats per output channel.  (First version: each thread added straight
// into dw -- adjacent lanes were cin x 9 floats apart, one 32-byte sector per atomic: 60 us per launch, more than
// half of the main kernel's time.)
constexpr int kRedCo = 32, kRedCi = 2, kRedRow = kRedCi * 9 + 1;
__global__ void __launch_bounds__(kRedCo * kRedCi * 9) wgrad_reduce_kernel(const WParams p) {
  __shared__ float out[kRedCo * kRedRow];
  const int col = threadIdx.x % kRedCo, cil = threadIdx.x / kRedCo % kRedCi, e = threadIdx.x / (kRedCo * kRedCi);
  const int ci0 = blockIdx.x * kRedCi, co0 = blockIdx.y * kRedCo;
  const int ci = ci0 + cil, co = co0 + col;
  for (int i = threadIdx.x; i < kRedCo * kRedRow; i += blockDim.x) out[i] = 0.f;
  __syncthreads();
  if (ci < p.cin && co < p.cout) {
    const int mb = ci >> 7, nb = co / p.nblk;
    const size_t stride = (size_t)p.tg * 128 * p.nblk;
    const int tgi = e / p.tg, t = e - tgi * p.tg;
    const float* src = p.part + (((size_t)(tgi * p.nmb + mb) * p.nnb + nb) * p.splits * p.tg + t) * 128 * p.nblk +
                       (size_t)(ci & 127) * p.nblk + (co - nb * p.nblk);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int sp = 0;
    for (; sp + 4 <= p.splits; sp += 4) {
      s0 += src[(size_t)sp * stride], s1 += src[(size_t)(sp + 1) * stride];
      s2 += src[(size_t)(sp + 2) * stride], s3 += src[(size_t)(sp + 3) * stride];
    }
    for (; sp < p.splits; ++sp) s0 += src[(size_t)sp * stride];
    const float sum = (s0 + s1) + (s2 + s3);
    const uint32_t mask = p.tl_mask[e];
    // the taps of different entries are disjoint in every tap list unet.cu builds (one addend per element, so the
    // result is deterministic); the shared-memory atomics keep overlapping masks correct too
    float* mine = out + col * kRedRow + cil * 9;
    for (int tap = 0; tap < 9; ++tap)
      if (mask >> tap & 1u) atomicAdd(mine + tap, sum);
  }
  __syncthreads();
  // one thread per element and launch, launches are stream-ordered: a plain read-modify-write is exact
  const int nci = min(kRedCi, p.cin - ci0);
  for (int idx = threadIdx.x; idx < kRedCo * kRedCi * 9; idx += blockDim.x) {
    const int r = idx / (kRedCi * 9), k = idx - r * (kRedCi * 9);
    if (co0 + r < p.cout && k < nci * 9) p.dw[((size_t)(co0 + r) * p.cin + ci0) * 9 + k] += out[r * kRedRow + k];
  }
}

// db[c] += sum over the interior pixels of an operand tensor (hi + lo).  grid: (row groups, C / 32); block: 256
// threads = 64 pixels x 4 groups of 8 channels, 16-byte loads (the 32 channels of a K-block are 64 contiguous
// bytes per pixel); the block's partial sums meet in shared memory and leave as 32 atomics.
__global__ void __launch_bounds__(256)
    op_channel_sum_kernel(const bf16* __restrict__ g, float* __restrict__ db, int N, int H, int W, int C, int cout,
                          int rows_per_block) {
  __shared__ float red[8][32];
  const OpShape s{N, H, W, C, 0};
  const int tid = threadIdx.x, sub = tid & 3, px = tid >> 2, chunk = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(N * H, r0 + rows_per_block);
  const size_t lo_off = s.lo_offset();
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = r0; r < r1; ++r) {
    const int n = r / H, h = r % H;
    const bf16* row = g + op_offset(s, n, h + 1, 1) + (size_t)chunk * s.block_stride() + sub * 8;
    for (int w = px; w < W; w += 64) {
      const uint4 a = *reinterpret_cast<const uint4*>(row + (size_t)w * 32);
      const uint4 b = *reinterpret_cast<const uint4*>(row + (size_t)w * 32 + lo_off);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc[2 * k] += __uint_as_float(aw[k] << 16) + __uint_as_float(bw[k] << 16);
        acc[2 * k + 1] += __uint_as_float(aw[k] & 0xFFFF0000u) + __uint_as_float(bw[k] & 0xFFFF0000u);
      }
    }
  }
  // lanes with equal `sub` hold the same 8 channels: fold the 8 pixels of a warp, then the 8 warps
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = acc[j];
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if ((tid & 31) < 4) red[tid >> 5][sub * 8 + j] = v;
  }
  __syncthreads();
  if (tid < 32) {
    float v = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) v += red[wq][tid];
    const int c = chunk * 32 + tid;
    if (c < cout) atomicAdd(db + c, v);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int encode5(CUtensorMap* map, const void* ptr, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
  auto fn = get_encode();
  if (!fn) return fail(SDAB_ERR_DEVICE, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SDAB_ERR_DEVICE, "cuTensorMapEncodeTiled failed with code " + std::to_string(r));
  return SDAB_OK;
}

}  // namespace

// H, W: resolution of the (sub-)problem, i.e. of the pixel tiles shared by the g and x tiles
bool wgrad_umma_supported(const WgradProblem& p) {
  return p.gOP && p.xOP && p.x_kind == 0 && p.Cg % 32 == 0 && p.Cx % 32 == 0 && p.Cg >= 32 && p.Cx >= 32 &&
         p.W % kBW == 0 && p.H % 16 == 0 && p.ntl >= 0 && p.ntl <= 9;
}

int conv3x3_wgrad_umma(const WgradProblem& c, int mode, cudaStream_t stream) {
  SDAB_REQUIRE(wgrad_umma_supported(c) && c.dw, "unsupported weight-gradient problem for the tcgen05 engine");
  WParams p{};
  p.planes = mode == SDAB_MODE_BF16X3 ? 2 : 1;
  const int BH = p.planes == 2 ? 8 : 16;
  const uint32_t kPatchSlot = BH == 8 ? Tile<8>::kPatchSlot : Tile<16>::kPatchSlot;
  const uint32_t kGBytes = BH == 8 ? Tile<8>::kGBytes : Tile<16>::kGBytes;
  p.tiles_w = c.W / kBW, p.tiles_h = c.H / BH, p.num_tiles = p.tiles_w * p.tiles_h * c.N;
  p.nchunk_x = c.Cx / 32, p.nchunk_g = c.Cg / 32;
  p.nblk = c.Cg % 128 == 0 ? 128 : (c.Cg % 96 == 0 ? 96 : (c.Cg % 64 == 0 ? 64 : 32));
  if (c.ntl) {
    p.ntl = c.ntl;
    for (int e = 0; e < c.ntl; ++e) {
      SDAB_REQUIRE(c.tl_sa[e] <= 2 && c.tl_sb[e] <= 2 && c.tl_mask[e] < 512, "invalid tap list");
      p.tl_shift[e] = (uint32_t)c.tl_sa[e] * kPW + c.tl_sb[e], p.tl_mask[e] = c.tl_mask[e];
    }
  } else {
    p.ntl = 9;
    for (int e = 0; e < 9; ++e) p.tl_shift[e] = (uint32_t)(e / 3) * kPW + e % 3, p.tl_mask[e] = 1u << e;
  }
  p.x_par = c.x_par, p.g_par = c.g_par, p.g_dh = c.g_dh, p.g_dw = c.g_dw;
  SDAB_REQUIRE(p.x_par >= 0 && p.x_par <= 4 && p.g_par >= 0 && p.g_par <= 4, "invalid parity image");
  // taps per CTA: as many as fit the 512 TMEM columns, evened out over the tap groups (nine taps at nblk = 128 are
  // 3 + 3 + 3, not 4 + 4 + 1: a launch lasts as long as its longest CTAs)
  const int tg_max = 512 / p.nblk;
  p.ntg = (p.ntl + tg_max - 1) / tg_max;
  p.tg = (p.ntl + p.ntg - 1) / p.ntg;
  p.nmb = (p.nchunk_x + 3) / 4;
  p.nnb = c.Cg / p.nblk;
  // (M = 128 always reads four chunk slots; with fewer input chunks the missing slots alias the next region of the
  // stage -- rows whose results are never read -- and the stage shrinks enough for a third pipeline stage at C = 96)
  p.x_plane_bytes = (uint32_t)(p.nchunk_x < 4 ? p.nchunk_x : 4) * kPatchSlot;
  p.g_plane_bytes = (uint32_t)(p.nblk / 32) * kGBytes;
  p.stage_bytes = p.planes * (p.x_plane_bytes + p.g_plane_bytes);
  // ... except behind the LAST plane of the last stage, where the phantom slots can stick out of the stage (one
  // input chunk, single plane): the allocation carries that much slack, an MMA must not read past it
  const long over = (long)(p.planes - 1) * p.x_plane_bytes + 4L * kPatchSlot - (long)p.stage_bytes;
  const uint32_t slack = over > 0 ? (uint32_t)over : 0u;
  p.stages = (int)((kSmemBudget - 2048 - slack) / p.stage_bytes);
  if (p.stages > 3) p.stages = 3;
  SDAB_REQUIRE(p.stages >= 1, "weight-gradient tile does not fit shared memory");
  const int units = p.ntg * p.nmb * p.nnb;
  // Pixel splits: CTAs run one per SM (each allocates all 512 TMEM columns), so a launch lasts
  // waves x tiles-per-CTA = ceil(units s / 148) x ceil(num_tiles / s) tile periods; the split count minimises that
  // (smallest s on ties: every split adds a pass over the gradient to the reduction).  (The first version took
  // ceil(148 / units): 162 CTAs at units = 27, i.e. a second wave for 14 of them -- 2 x 1/6 instead of 1 x 1/5.)
  static const int wg_sms = getenv("SDAB_WGRAD_SMS") ? atoi(getenv("SDAB_WGRAD_SMS")) : 148;
  {
    long best = -1;
    const int smax = p.num_tiles < 2 * wg_sms ? p.num_tiles : 2 * wg_sms;
    for (int s = 1; s <= smax; ++s) {
      if (c.partial && (size_t)units * s * p.tg * 128 * p.nblk * sizeof(float) > c.partial_bytes) break;
      const long cost = (long)((units * s + wg_sms - 1) / wg_sms) * ((p.num_tiles + s - 1) / s);
      if (best < 0 || cost < best) best = cost, p.splits = s;
    }
  }
  p.cin = c.cin, p.cout = c.cout, p.dw = c.dw;
  // split partial sums in the caller's workspace when it is large enough, fp32 atomics otherwise
  const size_t part_bytes = (size_t)units * p.splits * p.tg * 128 * p.nblk * sizeof(float);
  p.part = (c.partial && c.partial_bytes >= part_bytes) ? c.partial : nullptr;

  CUtensorMap tmX, tmG;
  // A tensor in the S2 layout holds an image of twice the sub-problem resolution as four parity images of
  // (H + 1) x (W + 1) haloed pixels; dim 3 then indexes (plane, chunk, parity).  Boxes that stick out of a
  // parity image read zeros, which only meet pixels outside the tile.
  auto make_map = [&](CUtensorMap* map, const bf16* ptr, int nchunk, int par, bool interior, int bw, int bh) -> int {
    const cuuint64_t Q = 2 * (cuuint64_t)nchunk;
    if (par) {
      const cuuint64_t Hp = 2 * c.H + 2, Wp = 2 * c.W + 2;
      const cuuint64_t dims[5] = {32, Wp / 2, Hp / 2, 4 * Q, (cuuint64_t)c.N};
      const cuuint64_t strides[4] = {64, (Wp / 2) * 64, (Hp / 2) * (Wp / 2) * 64, Q * Hp * Wp * 64};
      const cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
      return encode5(map, ptr, dims, strides, box);
    }
    const cuuint64_t Hp = c.H + 2, Wp = c.W + 2;
    const cuuint64_t strides[4] = {64, Wp * 64, Hp * Wp * 64, Q * Hp * Wp * 64};
    const cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
    if (interior) {  // tiles of the interior pixels of a haloed operand
      const cuuint64_t dims[5] = {32, (cuuint64_t)c.W, (cuuint64_t)c.H, Q, (cuuint64_t)c.N};
      return encode5(map, ptr + ((size_t)Wp + 1) * 32, dims, strides, box);
    }
    const cuuint64_t dims[5] = {32, Wp, Hp, Q, (cuuint64_t)c.N};
    return encode5(map, ptr, dims, strides, box);
  };
  SDAB_TRY(make_map(&tmX, c.xOP, p.nchunk_x, p.x_par, false, kPW, BH + 2));
  SDAB_TRY(make_map(&tmG, c.gOP, p.nchunk_g, p.g_par, true, kBW, BH));
  const size_t smem = 2048 + (size_t)p.stages * p.stage_bytes + slack;
  static bool attr_set = false;
  if (!attr_set) {
    SDAB_CUDA_CHECK(cudaFuncSetAttribute(wgrad_umma_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    SDAB_CUDA_CHECK(cudaFuncSetAttribute(wgrad_umma_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    attr_set = true;
  }
  const int grid = units * p.splits;
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 1 : 0;
    if (p.planes == 2)
      SDAB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, wgrad_umma_kernel<2, 8>, tmX, tmG, p));
    else
      SDAB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, wgrad_umma_kernel<1, 16>, tmX, tmG, p));
  }
  SDAB_LAUNCH_CHECK("wgrad_umma_kernel");
  if (p.part) {
    wgrad_reduce_kernel<<<dim3((p.cin + kRedCi - 1) / kRedCi, (p.cout + kRedCo - 1) / kRedCo), kRedCo * kRedCi * p.ntl, 0, stream>>>(p);
    SDAB_LAUNCH_CHECK("wgrad_reduce_kernel");
  }
  if (c.db) {
    const int chunks = c.Cg / 32, rows = c.N * c.H;
    const int groups = (148 * 8 + chunks - 1) / chunks, rpb = (rows + groups - 1) / groups;
    op_channel_sum_kernel<<<dim3((rows + rpb - 1) / rpb, chunks), 256, 0, stream>>>(c.gOP, c.db, c.N, c.H, c.W, c.Cg,
                                                                                    c.cout, rpb);
    SDAB_LAUNCH_CHECK("op_channel_sum_kernel");
  }
  return SDAB_OK;
}

}  // namespace sdab
