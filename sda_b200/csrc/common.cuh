// common.cuh -- shared declarations of libsdab (sm_100a only).
//
// Internal activation formats (DESIGN.md "Data layout in HBM"):
//   F(C)   fp32, NHWC           [N][H][W][C]
//   OP(C)  bf16 hi/lo operand   [N][2][C/32][H+2][W+2][32] -- plane 0 = hi, plane 1 = lo, then the
//          32-channel K-block, then the haloed image, channels innermost.  One (plane, K-block) of
//          128 consecutive pixels is a CONTIGUOUS 8 KB run, so the TMA box that feeds one MMA
//          K-block streams full cache lines (a pixel-major layout makes every 64 B row a separate
//          strided request and was measured ~4x slower to fill).  Physically haloed: the ring
//          replicates the opposite edge (circular padding, nn.Conv2d(padding_mode='circular'),
//          sda/nn.py:125-128) so that every 3x3 tap of every tile is one in-bounds TMA box.
//   OP/S2  the same with each image de-interleaved by (row, column) parity:
//          [N][2][C/32][4][(H+2)/2][(W+2)/2][32] -- the input format of the stride-2 heads
//          (sda/nn.py:151-159), again so that each tap is one dense TMA box.
#pragma once
#include <cstdlib>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/sdab.h"

namespace sdab {

// Programmatic dependent launch (PDL).  A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization
// may become resident while its predecessor in the stream is still draining: everything up to pdl_wait() (barrier
// initialisation, TMEM allocation) overlaps the predecessor's tail, pdl_wait() returns once the predecessor has
// completed and its memory is visible.  No global memory is read or written before it.  pdl_launch_dependents()
// lets the NEXT kernel start launching as soon as every CTA of this one has issued it (they only find room as this
// grid's CTAs retire).  sdab_pdl_enabled(): SDAB_PDL=0 switches the launch attribute off (A/B).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
inline bool pdl_enabled() {
  static const int on = getenv("SDAB_PDL") ? atoi(getenv("SDAB_PDL")) : 1;
  return on != 0;
}


// ----------------------------------------------------------------------------- errors
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
void count_launch(int n = 1);
void conv_profile_before(cudaStream_t st);
void conv_profile_after(cudaStream_t st, double flops);

#define SDAB_CUDA_CHECK(expr)                                                                      \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::sdab::fail(SDAB_ERR_DEVICE, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
  } while (0)

#define SDAB_LAUNCH_CHECK(name)                                                                    \
  do {                                                                                             \
    ::sdab::count_launch();                                                                        \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess)                                                                         \
      return ::sdab::fail(SDAB_ERR_DEVICE, std::string("launch of ") + name + ": " +              \
                                               cudaGetErrorString(_e));                            \
  } while (0)

#define SDAB_REQUIRE(cond, msg)                                                                    \
  do {                                                                                             \
    if (!(cond)) return ::sdab::fail(SDAB_ERR_ARG, std::string(msg) + " [" #cond "]");             \
  } while (0)

#define SDAB_TRY(expr)                                                                             \
  do {                                                                                             \
    int _s = (expr);                                                                               \
    if (_s != SDAB_OK) return _s;                                                                  \
  } while (0)

typedef __nv_bfloat16 bf16;

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline size_t round_up_sz(size_t x, size_t m) { return (x + m - 1) / m * m; }

// ----------------------------------------------------------------------------- operand layout
struct OpShape {
  int N, H, W, C;  // logical image size, channels (multiple of 32)
  int s2;          // 1: parity de-interleaved layout
  __host__ __device__ size_t elems() const { return (size_t)N * (H + 2) * (W + 2) * 2 * C; }
  __host__ __device__ size_t bytes() const { return elems() * 2; }
  // elements between consecutive (plane, K-block) images
  __host__ __device__ size_t block_stride() const { return (size_t)(H + 2) * (W + 2) * 32; }
  // element offset from the hi plane to the lo plane of the same K-block
  __host__ __device__ size_t lo_offset() const { return (size_t)(C / 32) * block_stride(); }
};

// element offset of channel 0 of K-block 0 of the hi plane of padded pixel (hp, wp) of image n.
// Channel c of plane p lives at  + (p * C/32 + c/32) * block_stride() + c % 32.
__host__ __device__ __forceinline__ size_t op_offset(const OpShape& s, int n, int hp, int wp) {
  const int Hp = s.H + 2, Wp = s.W + 2;
  size_t pix;
  if (!s.s2) {
    pix = (size_t)hp * Wp + wp;
  } else {
    const int par = (hp & 1) * 2 + (wp & 1);
    pix = ((size_t)par * (Hp >> 1) + (hp >> 1)) * (Wp >> 1) + (wp >> 1);
  }
  return ((size_t)n * 2 * (s.C / 32) * Hp * Wp + pix) * 32;
}

// Calls f(hp, wp) for the padded positions that hold logical pixel (h, w): itself plus its
// halo replicas (up to 4).
template <class F>
__device__ __forceinline__ void for_each_replica(int h, int w, int H, int W, F&& f) {
  const int hp = h + 1, wp = w + 1;
  const int hr = (h == 0) ? H + 1 : ((h == H - 1) ? 0 : -1);
  const int wr = (w == 0) ? W + 1 : ((w == W - 1) ? 0 : -1);
  f(hp, wp);
  if (hr >= 0) f(hr, wp);
  if (wr >= 0) f(hp, wr);
  if (hr >= 0 && wr >= 0) f(hr, wr);
}

__device__ __forceinline__ void split_bf16(float v, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// Two values at once: element a in the low half-words of (hi, lo), element b in the high ones
// (one packed conversion per plane; bf16 -> fp32 is a 16-bit shift / mask).
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

// sigmoid through the flush-to-zero approximations (2 MUFU + 3 FP32 instructions; the default
// __expf / __fdividef expand to ~12 with their denormal range handling): relative error ~2^-21
__device__ __forceinline__ float sigmoid_fast(float v) {
  float t, s;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-1.4426950408889634f * v));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(1.f + t));
  return s;
}

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == 1) return v * sigmoid_fast(v);  // SiLU
  if (act == 2) return fmaxf(v, 0.f);        // ReLU
  return v;
}

// derivative of the activation at pre-activation c
__device__ __forceinline__ float act_bwd(float c, int act) {
  if (act == 1) {
    const float s = sigmoid_fast(c);
    return s * (1.f + c * (1.f - s));
  }
  if (act == 2) return c > 0.f ? 1.f : 0.f;
  return 1.f;
}

// ----------------------------------------------------------------------------- conv problem
struct ConvEpilogue {
  const float* bias;   // [Cout] or null
  const float* res;    // F(Cout) residual added before anything else, or null
  float* pre;          // F(Cout): v = acc + bias + res saved here (pre-activation), or null
  const float* dact;   // F(Cout): result multiplied by act'(dact) (backward of the activation), or null
  float* outF;         // F(Cout) or null
  bf16* outOP;         // OP(Cout) (normal layout, halo written) or null
  int act;             // forward activation applied to v: 0 none, 1 SiLU, 2 ReLU
  int dact_kind;       // which activation's derivative `dact` refers to (1 SiLU, 2 ReLU)
  // Channel LayerNorm fused into the epilogue (tcgen05 engine, C_out % 32 == 0 only; the thread that
  // owns a pixel row of the TMEM accumulator sees all C_out channels):
  //   ln == 1  forward : outOP = LN_C(f + ln_shift) instead of split(f), f = acc + bias + res (-> outF);
  //                      this is the operand of the NEXT modulated block (sda/nn.py:27-28,137)
  //   ln == 2  backward: f = res + (acc - mean_C acc - a * sum_C(acc * a) / (C-1)) * rstd  (-> outF, outOP)
  int ln;
  const float* ln_shift;   // forward: [ln_nt][ln_shift_stride] shift of the next block (or null)
  int ln_shift_stride, ln_nt;
  float* ln_rstd_out;      // forward: 1 / sqrt(var + eps) per pixel saved here (or null)
  const bf16* ln_a;        // backward: saved normalised operand OP(C_out) of the block input
  const float* ln_rstd_in; // backward: saved rstd per pixel
};

// Explicit tap list of a convolution launch (n == 0 selects the default 3x3 list).  Tap t reads the
// TMA box at row / column offsets (ca, cb) of the haloed operand image (parity image cp when the
// input is in the S2 layout) and multiplies it with tap `wtap` of the packed weights.  This covers
// the 3x3 conv, the stride-2 heads, the sub-pixel form of "nearest x2 -> conv" (4 taps per output
// parity), its transpose (a 4x4 stride-2 conv, 16 taps) and the parity classes of the head transposes.
struct ConvTaps {
  int n;
  unsigned char ca[16], cb[16], cp[16], wtap[16];
};

struct ConvProblem {
  const bf16* in;      // OP(Cin) at resolution (H*stride, W*stride); S2 layout iff stride == 2 or in_s2
  const bf16* wpk;     // packed weights [9][Cin/32][2][Cout][32]
  int N, H, W;         // OUTPUT resolution
  int Cin, Cout;       // padded: Cin % 32 == 0, Cout % 16 == 0
  int stride;          // 1 or 2
  int mode;            // SDAB_MODE_*
  int in_s2;           // input operand in the parity layout of a (2H) x (2W) image (implied by stride == 2)
  int in_strided;      // tcgen05 engine: ... but `in` is the NORMAL layout of that image, read with a TMA element stride of 2
  ConvTaps taps;       // explicit taps (tcgen05 engine only), or n == 0
  int wtaps;           // taps in the packed weight array (0 means 9)
  int os, oh0, ow0;    // output placement (tcgen05 engine only): GEMM pixel (h, w) is pixel (os h + oh0, os w + ow0)
                       // of an (os H) x (os W) output image; os == 0 means 1
  double flops;        // algorithmic FLOPs of this launch (2 * pixels * 9 * C_in,real * C_out,real), for profiling
  ConvEpilogue epi;
};

// Weight / bias gradient of one 3x3 circular convolution (csrc/wgrad.cu): dw (cout, cin, 3, 3) and db (cout)
// are ACCUMULATED into (atomics over pixel splits) -- the caller zero-fills them.
struct WgradProblem {
  const float* gF;  // cotangent of the convolution output: F(Cg) at the output resolution, or null ...
  const bf16* gOP;  // ... or as an operand tensor OP(Cg), normal layout
  const bf16* xOP;  // convolution input as operand tensor OP(Cx): x_kind 0 normal layout (stride 1), 1 parity layout
                    // of the (2H) x (2W) input (stride 2), 2 at HALF the output resolution (nearest x2 folded in)
  const float* xF;  // x_kind 3: F(Cx) at the output resolution, activation `act` applied on load
  int x_kind, act;
  int N, H, W;      // OUTPUT resolution
  int Cg, Cx;       // (padded) channel counts of the g and x tensors
  int cout, cin;    // real channel counts
  float* dw;
  float* db;        // or null
  // tcgen05 engine only -- generalised tap list for the strided / upsampled layers, which decompose into
  // stride-1 sub-problems between PARITY images (the de-interleaved S2 layout) of one operand and the other
  // operand: entry e multiplies g[pixel] with x[pixel + (tl_sa[e], tl_sb[e])] (offsets 0..2 from the patch
  // origin) and adds the product into every weight tap of the 9-bit set tl_mask[e].  ntl == 0: the nine taps.
  int ntl;
  unsigned char tl_sa[9], tl_sb[9];
  unsigned short tl_mask[9];
  float* partial;        // tcgen05 engine: workspace for the pixel-split partial sums (summed in fixed order by a
  size_t partial_bytes;  // second kernel); null or too small: fp32 atomics into dw (non-deterministic order)
  int x_par, g_par;  // 1 + parity image id ((row & 1) * 2 + (col & 1)) of an S2-layout x / g tensor, 0: normal layout
  int g_dh, g_dw;    // origin offset of the g tile inside its parity image
};
int conv3x3_wgrad(const WgradProblem& p, cudaStream_t stream);
// db[c] += sum over the pixels of an F(C) tensor, c < cout
int f_channel_sum(const float* g, float* db, size_t pixels, int C, int cout, cudaStream_t stream);
// tcgen05 engine (csrc/wgrad_umma.cu): g and x both as operand tensors, x_kind 0, tileable into 8 x 16 boxes
bool wgrad_umma_supported(const WgradProblem& p);
int conv3x3_wgrad_umma(const WgradProblem& p, int mode, cudaStream_t stream);
// dshift[(Nt > 1 ? n : 0) * stride + c] += sum_{h, w} (a - b)[n, h, w, c]   (F(C) tensors)
int shift_grad(const float* a, const float* b, float* dshift, int stride, int Nt, int N, int H, int W, int C,
               cudaStream_t stream);

int conv3x3_simt(const ConvProblem& p, cudaStream_t stream);
int conv3x3_umma(const ConvProblem& p, cudaStream_t stream);

// Applies the epilogue to 16 consecutive output channels [c0, c0+16) of one pixel.
// pix = (n*H + h)*W + w.  Shared by both engines so that they agree bit for bit in everything
// but the accumulation order.
__device__ __forceinline__ void epilogue_store16(const ConvEpilogue& e, float (&v)[16], size_t pix, int n, int h,
                                                 int w, int H, int W, int Cout, int c0) {
  const size_t off = pix * Cout + c0;
  if (e.bias) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + c0 + j));
      v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
    }
  }
  if (e.res) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 r = *reinterpret_cast<const float4*>(e.res + off + j);
      v[j] += r.x, v[j + 1] += r.y, v[j + 2] += r.z, v[j + 3] += r.w;
    }
  }
  if (e.pre) {
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      *reinterpret_cast<float4*>(e.pre + off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
  if (e.act) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = act_fwd(v[j], e.act);
  }
  if (e.dact) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 c = *reinterpret_cast<const float4*>(e.dact + off + j);
      v[j] *= act_bwd(c.x, e.dact_kind), v[j + 1] *= act_bwd(c.y, e.dact_kind);
      v[j + 2] *= act_bwd(c.z, e.dact_kind), v[j + 3] *= act_bwd(c.w, e.dact_kind);
    }
  }
  if (e.outF) {
#pragma unroll
    for (int j = 0; j < 16; j += 4)
      *reinterpret_cast<float4*>(e.outF + off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
  if (e.outOP) {
    __align__(16) bf16 hi[16];
    __align__(16) bf16 lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) split_bf16(v[j], hi[j], lo[j]);
    const OpShape s{0, H, W, Cout, 0};
    const size_t blk = (size_t)(c0 >> 5) * s.block_stride() + (c0 & 31), lo_off = s.lo_offset();
    for_each_replica(h, w, H, W, [&](int hp, int wp) {
      bf16* dst = e.outOP + op_offset(s, n, hp, wp) + blk;
      reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(hi)[0];
      reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(hi)[1];
      reinterpret_cast<uint4*>(dst + lo_off)[0] = reinterpret_cast<const uint4*>(lo)[0];
      reinterpret_cast<uint4*>(dst + lo_off)[1] = reinterpret_cast<const uint4*>(lo)[1];
    });
  }
}

// ----------------------------------------------------------------------------- elementwise launches
// Window view of a trajectory (MCScoreNet.forward, sda/score.py:134-164): the range of flattened (B, L - 2k)
// windows one call evaluates, and where the folded frames go (FoldDst in elementwise.cu).
struct WindowIO {
  int B, L, C, Cc;  // trajectory (B, L, C, H, W); context channels appended to every window
  int order;        // k: a window is 2k + 1 frames
  int w_begin;      // first flattened window of this call
  int per, cap;     // sharded output: windows per rank and frames per shard; cap == 0: write the trajectory in place
};
int pack_windows_to_op(const float* x, const float* ctx, bf16* op, const WindowIO& w, int N, int Cpad, int H, int W,
                       cudaStream_t st);
int pack_fold_adjoint_to_op(const float* g, bf16* op, const WindowIO& w, int N, int Cpad, int H, int W, cudaStream_t st);
int unpack_f_fold(const float* f, float* out, const WindowIO& w, int N, int Cpad, int H, int W, cudaStream_t st);
int pack_nchw_to_op(const float* x, bf16* op, int N, int Creal, int Cpad, int H, int W, int s2, cudaStream_t st);
int unpack_f_to_nchw(const float* f, float* x, int N, int Creal, int Cpad, int H, int W, cudaStream_t st);
// kind: 0 normal, 1 S2 (parity) layout, 2 zero-insertion x2 upsample (src at half resolution)
int f_to_operand(const float* f, bf16* op, int N, int H, int W, int C, int kind, cudaStream_t st);
// LN over channels of (x + shift) -> OP (optionally nearest x2 upsampled); rstd saved when not null.
// shift: [Nt][shift_stride] slice starting at channel 0 of this block, or null.
int ln_forward(const float* x, const float* shift, int shift_stride, int Nt, bf16* op, float* rstd, int N, int H, int W,
               int C, int upsample, cudaStream_t st);
// backward of LN: ga (optionally 2x2 sum-pooled from resolution 2H x 2W), a = hi+lo read from the
// saved operand (at (2h,2w) of the upsampled operand when upsample), gx = res + gu.
int ln_backward(const float* ga, const bf16* a_op, const float* rstd, const float* res, float* gxF, bf16* gxOP, int N,
                int H, int W, int C, int pooled, cudaStream_t st);
int time_shifts(const float* y, const float* pw, const float* pb, float* out, int Nt, int rows, int mod,
                cudaStream_t st);
int pack_conv_weights(const float* w, bf16* fwd, bf16* bwd, int Cout, int Cin, cudaStream_t st);
// the same for up to kMaxPack convolutions in one launch, and up to kMaxCopy zero-padded copies in one launch
constexpr int kMaxPack = 64, kMaxCopy = 96;
struct PackTable {
  const float* w[kMaxPack];
  bf16* fwd[kMaxPack];
  bf16* bwd[kMaxPack];
  int cout[kMaxPack], cin[kMaxPack];
  unsigned long long start[kMaxPack + 1];
  int n;
};
struct CopyTable {
  const float* src[kMaxCopy];
  float* dst[kMaxCopy];
  int n_src[kMaxCopy], n_dst[kMaxCopy];
  unsigned long long start[kMaxCopy + 1];
  int n;
};
int pack_conv_weights_batched(PackTable& t, cudaStream_t st);
int copy_pad_batched(CopyTable& t, cudaStream_t st);
int pack_tail_weights(const float* w, bf16* tf, bf16* tb, int Cout, int Cin, cudaStream_t st);
int copy_f32(const float* src, float* dst, size_t n, cudaStream_t st);
int fill_zero(void* p, size_t bytes, cudaStream_t st);

}  // namespace sdab
