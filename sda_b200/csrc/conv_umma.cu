// conv_umma.cu -- tcgen05 implicit-GEMM 3x3 circular convolution for sm_100a (product engine).
//
// Replaces nn.Conv2d(kernel_size=3, padding=1, padding_mode='circular', stride in {1,2}) as built
// by sda/nn.py:125-174, forward and (with transposed/flipped packed weights) input-gradient.
//
//   GEMM view    D[m, co] = sum_{tap, ci} A_tap[m, ci] * W[tap][co, ci]
//                M = 128 output pixels (one TileGeom box), N = C_out (<= 512 fp32 TMEM columns),
//                K = 9 taps x C_in, walked in 32-channel blocks (64 B rows, SWIZZLE_64B).
//   A operand    one TMA box per (tap, 32-channel block, hi|lo plane) of the haloed NHWC operand
//                tensor: the circular wrap is already materialised in the halo ring, so every box
//                is in bounds; stride-2 heads read the parity de-interleaved layout instead.
//   B operand    one (two for C_out > 256) TMA box per (tap, block, plane) of the packed weights.
//   precision    SDAB_MODE_BF16X3: x = hi + lo in bf16; hi*hi + hi*lo + lo*hi accumulated in the same
//                fp32 TMEM accumulator (3 MMAs per product, ~2^-16 relative error);
//                SDAB_MODE_BF16: hi*hi only.
//   schedule     persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer, warp 1 = MMA
//                issuer (one elected lane) + TMEM allocator, warps 2-5 = epilogue (TMEM -> registers ->
//                fused bias / residual / activation / activation-derivative / bf16 split -> global).
//                smem ring of `stages` K-blocks (full/empty mbarriers), TMEM accumulator double
//                buffered when C_out <= 256 so the epilogue of tile i overlaps the MMAs of tile i+1.
//                (conv_umma_patch_kernel: 384 threads -- warpgroup 0 = producer + MMA issuer, warpgroups 1, 2 = two
//                epilogue warpgroups, registers re-divided by setmaxnreg; C_out = 384 with a fused LayerNorm builds
//                its accumulator as three pieces in a ring of four TMEM slots.)
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "tile_geom.h"

namespace sdab {

namespace {

// Developer ablation of the patch kernel / LayerNorm epilogue, compiled in only by -DSDAB_ABLATE=bits (a separate
// build for SDAB_LIB, wrong results by design): 1 no MMA issue, 16 no global epilogue operands, 32 no epilogue stores,
// 64 staging written but no TMA store issued, 128 no F output (OP only)
#ifndef SDAB_ABLATE
#define SDAB_ABLATE 0
#endif
constexpr int kAblate = SDAB_ABLATE;
constexpr int kThreads = 192;
// patch kernel: warpgroup 0 = TMA producer (warp 0), MMA issuer (warp 1) and two idle warps, warpgroups 1 and 2 =
// the two epilogue warpgroups (warps 4-7 and 8-11).  The roles sit on hardware warpgroup boundaries so that the
// register file can be re-divided (setmaxnreg): 56 registers per thread for warpgroup 0, 224 for the epilogue --
// 128 x 56 + 256 x 224 = 64512, what a 384-thread CTA launched at 168 owns.
constexpr int kPatchThreads = 384;
constexpr int kPatchEpi0 = 128;  // first epilogue thread of the patch kernel (64 in conv_umma_kernel)
constexpr uint32_t kABytes = 128 * 64;  // one A box: 128 pixels x 32 channels x bf16
constexpr uint32_t kCtrlBytes = 1024;
// LayerNorm kernels: bias[512] and shift[512] copies, then the csplit statistics exchange, behind the control block
constexpr int kCstShift = 512;
constexpr uint32_t kCstBytes = 4096, kXchgBytes = 4096;
constexpr uint32_t kStageF = 128 * 128;      // epilogue staging: 128 pixels x 32 fp32 channels (SWIZZLE_128B rows)
constexpr uint32_t kStageO = 128 * 64;       // 128 pixels x 32 bf16 channels (SWIZZLE_64B rows), one per plane
constexpr uint32_t kStagingBytes = kStageF + 2 * kStageO;
constexpr uint32_t kSmemBudget = 227 * 1024;
constexpr int kMaxStages = 8;
// Patch kernel: the M tile is 8 x 16 pixels of one image, its haloed 10 x 18 input patch is ONE TMA box
// per (plane, K-block), and tap (a, b) is the same shared-memory tile addressed through a descriptor
// whose start is shifted by (10 a + b) pixel rows (64 B each) with a stride-byte-offset of one patch
// row (640 B).  SWIZZLE_64B is a function of the shared-memory address bits alone, so shifted
// starts need no base-offset correction (verified on hardware by tools/probes/umma_shift_probe.cu).
constexpr int kPatchBW = 8, kPatchBH = 16;
constexpr int kPatchW = kPatchBW + 2, kPatchH = kPatchBH + 2;
constexpr uint32_t kPatchBytes = kPatchW * kPatchH * 64;  // 11520 B landing per (plane, K-block)
constexpr uint32_t kPatchPlane = 12288;                   // padded to the 1 KB swizzle alignment
constexpr uint32_t kPatchSBO = kPatchW * 64;
constexpr int kMaxBStages = 16;  // (the control block has room for 4 patch stages)

struct UmmaParams {
  TileGeom g;
  int N, H, W, Cin, Cout, stride, nchunk;
  int planes;      // 2 in bf16x3 mode (hi, lo), 1 in bf16 mode
  int stages, acc_stages;
  int CB, nb;      // N of one MMA, number of N halves
  uint32_t b_plane_bytes, stage_bytes;
  int in_s2;       // A operand addressed by parity image (stride-2 layers)
  int in_strided;  // ... of the NORMAL layout through a TMA element stride of 2 (no parity copy exists)
  int ntaps;
  uint32_t tap[16];  // ca | cb << 8 | cp << 16 | wtap << 24
  int os, oh0, ow0, Ho, Wo;  // output placement (ConvProblem::os ...) and output image size
  int staged;      // epilogue through shared memory + TMA stores (C_out % 32 == 0)
  int sbufs;       // number of staging sets (TMA stores in flight behind the epilogue)
  int out_chunks;  // C_out / 32
  int patch;       // patch kernel: one haloed input patch per K-block serves all nine taps
  int a_stages;    // patch kernel: depth of the patch ring (`stages` is the depth of the weight ring)
  uint32_t b_stage_bytes;
  int nsplit;      // patch kernel, C_out > 256 without LayerNorm: a work item is (tile, half of the output channels)
  int item_chunks; // 32-channel blocks of one work item (out_chunks / nsplit)
  int csplit;      // patch kernel: epilogue warpgroups split channel blocks even with a double-buffered accumulator
  int epi_wgs;     // patch kernel: epilogue warpgroups in use (2; SDAB_UMMA_WG=1 leaves the second one idle)
  int pieces;      // patch kernel, C_out = 384 with a fused LayerNorm: the accumulator of a tile is built as `pieces`
                   // consecutive N = CB items in a ring of acc_stages TMEM slots (1: one item per tile)
  uint32_t ctrl_bytes;  // control block in front of the staging tiles (barriers, TMEM slot, csplit statistics exchange)
  int debug;       // SDAB_UMMA_DEBUG bits (developer ablation): 1 = no MMA issue, 2 = no TMA, 4 = no epilogue work
  ConvEpilogue epi;
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must abort the kernel (trap -> launch error), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// 2-CTA (cta_group::2) variants: the copy lands in the executing CTA's shared memory but completes
// its bytes on the LEADER CTA's barrier (peer bit of the shared::cluster address cleared).
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same offset in BOTH CTAs of the pair once the MMAs issued so far finish
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
// arrive on the leader CTA's barrier (works from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
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

// 32 consecutive fp32 columns of this thread's TMEM lane written back (the LayerNorm epilogues park values
// next to / in place of the accumulator between their two passes); tmem_st_wait before reading them again
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])),
      "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])),
      "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
      "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])),
      "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// named barrier of one epilogue warpgroup (ids 1 and 2; 0 is __syncthreads)
__device__ __forceinline__ void epi_barrier(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }
#define EPI_BARRIER() do { if (!(p.debug & 64)) epi_barrier(bar_id); } while (0)

// bias / residual / activation / activation-derivative on 32 consecutive channels of one pixel
// (same order of operations as epilogue_store16).  The residual / derivative operands are passed in
// registers: the caller issues their global loads BEFORE waiting on the accumulator.  f receives
// the value stored to the F output: the pre-activation when e.pre is set, the final value otherwise.
__device__ __forceinline__ void epilogue_math32(const ConvEpilogue& e, float (&v)[32], float (&f)[32],
                                                const float (&rr)[32], bool has_res, bool has_dact, int c0) {
  if (e.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + c0 + j));
      v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
    }
  }
  if (has_res) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += rr[j];
  }
  if (e.pre) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = v[j];
  }
  // the activation kind is uniform: branch once around straight-line loops (a per-element switch
  // compiles to every alternative plus selects)
  if (e.act == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= sigmoid_fast(v[j]);
  } else if (e.act == 2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (has_dact) {
    if (e.dact_kind == 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float s = sigmoid_fast(rr[j]);
        v[j] *= s * (1.f + rr[j] * (1.f - s));
      }
    } else if (e.dact_kind == 2) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = rr[j] > 0.f ? v[j] : 0.f;
    }
  }
  if (!e.pre) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = v[j];
  }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }


// 32 consecutive floats of one pixel with 16 B loads
__device__ __forceinline__ void load32(const float* __restrict__ src, float (&o)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 t = *reinterpret_cast<const float4*>(src + 4 * j);
    o[4 * j] = t.x, o[4 * j + 1] = t.y, o[4 * j + 2] = t.z, o[4 * j + 3] = t.w;
  }
}

// Work items of the CTA-pair kernels.  The loop index vt runs over "virtual tiles": pair-item v = vt >> 1 of CTA
// vt & 1 of the pair.  With nsplit == 1 an item is a tile (vt IS the tile index); with nsplit == 2 the pair-item
// v is half `v % 2` of the output channels of tile pair v / 2 -- twice as many, half as long items: the two
// accumulator stages overlap MMAs and epilogue even at C_out = 384 (a single 384-column accumulator cannot),
// and a short launch (a few hundred tiles on 74 pairs) wastes half as much of its last wave.
__device__ __forceinline__ void item_of(const UmmaParams& p, int vt, int& tile, int& nh) {
  if (p.nsplit == 1) {
    tile = vt, nh = 0;
  } else {
    const int v = vt >> 1, pt = v / p.nsplit;
    nh = v - pt * p.nsplit;
    tile = 2 * pt + (vt & 1);
  }
}

// ------------------------------------------------------------------------------------------ epilogue
// Epilogue role (warps 2..5 of both kernels) of the convolutions without a fused LayerNorm: TMEM -> registers ->
// fused bias / residual / activation / activation-derivative / bf16 split -> staged TMA stores.
template <bool CTA2>
__device__ __forceinline__ void epilogue_role(const UmmaParams& p, const CUtensorMap& tmF, const CUtensorMap& tmO,
                                              uint32_t bar_tfull, uint32_t bar_tempty, uint32_t tmem_base,
                                              uint32_t acc_stride, uint32_t staging_base, int tile_begin, int tile_end,
                                              int warp, int lane, int wg = 0, int nwg = 1, bool csplit = false,
                                              int epi0 = 64) {
  // `nwg` epilogue warpgroups, each with its own staging sets, named barrier and store-issuing thread (its
  // first).  They take alternate tiles (warpgroup wg always drains TMEM accumulator wg), or -- csplit, for
  // the single-accumulator case (C_out > 256) without LayerNorm -- alternate 32-channel blocks of every tile.
  const uint32_t staging0 = staging_base + wg * p.sbufs * kStagingBytes;
  const int issuer = epi0 + 128 * wg;
  const int cc0 = csplit ? wg : 0, ccstep = csplit ? nwg : 1;
  const int bar_id = wg;
  if (csplit) wg = 0, nwg = 1;  // tile walk of a single warpgroup
  const int q = warp & 3;  // TMEM lane quarter accessible to this warp
  const int m = q * 32 + lane;
  const int bw = m % p.g.BW, bh = (m / p.g.BW) % p.g.BH, bn = m / (p.g.BW * p.g.BH);
  int it = wg;
  int sbuf = 0;  // staging set of the next 32-channel block (running over tiles)
  const int v_end = ((p.g.num_tiles + 1) >> 1) * p.nsplit;  // pair-items (nsplit > 1: CTA-pair patch kernel only)
  auto in_range = [&](int vt) { return p.nsplit > 1 ? (vt >> 1) < v_end : vt < tile_end; };
  const int nchunks = p.item_chunks;
  for (int vt = tile_begin + wg * (int)gridDim.x; in_range(vt); vt += nwg * (int)gridDim.x, it += nwg) {
    const int acc = it % p.acc_stages;
    const uint32_t acc_phase = (it / p.acc_stages) & 1;
    int tile, nh;
    item_of(p, vt, tile, nh);
    const int gc0 = nh * nchunks;  // first 32-channel block of this item in the output tensor
    int n0, h0, w0;
    p.g.tile_origin(tile, n0, h0, w0);
    // (h, w): output pixel of this thread in the (Ho x Wo) output image (ConvProblem::os, oh0, ow0)
    const int n = n0 + bn, h = p.os * (h0 + bh) + p.oh0, w = p.os * (w0 + bw) + p.ow0;
    const bool valid = n < p.N;
    const size_t pix = ((size_t)n * p.Ho + h) * p.Wo + w;
    // developer ablation bits: 16 = no global operand loads, 32 = no stores
    const bool has_res = p.epi.res != nullptr && valid && !(p.debug & 16) && !(kAblate & 16),
               has_dact = p.epi.dact != nullptr && valid && !(p.debug & 16) && !(kAblate & 16);
    // the global operand (residual or saved pre-activation) of a block is requested one block ahead of its use: the
    // first one here, before the accumulator is waited for, the next ones as soon as the previous is consumed
    float rr[32];
    if (p.staged && !(p.debug & 4)) {
      if (has_res) load32(p.epi.res + pix * p.Cout + (nh * nchunks + cc0) * 32, rr);
      if (has_dact) load32(p.epi.dact + pix * p.Cout + (nh * nchunks + cc0) * 32, rr);
    }
    mbar_wait(bar_tfull + 8 * acc, acc_phase);
    tc_fence_after();
    const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * acc_stride;
    if (p.debug & 4) {
      // ablation: no epilogue work
    } else if (!p.staged) {
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        float v[16];
        tmem_ld16(t0 + c0, v);
        if (valid) epilogue_store16(p.epi, v, pix, n, h, w, p.Ho, p.Wo, p.Cout, c0);
      }
    } else {
      // Staged epilogue: every output leaves the SM as full cache lines.  Per 32-channel block the
      // 128 threads (one pixel each) write their values into swizzled staging tiles; one thread
      // then issues TMA stores (F: [pixels][C] matrix, OP: the (plane, K-block) image box).
      const bool wantF = (p.epi.outF != nullptr || p.epi.pre != nullptr) && !(kAblate & 128);
      const bool wantO = p.epi.outOP != nullptr;
      const bool edge = valid && (h == 0 || h == p.Ho - 1 || w == 0 || w == p.Wo - 1);
      const OpShape so{0, p.Ho, p.Wo, p.Cout, 0};
      {
        // pull the next item's epilogue operands of this pixel row into L2 one item ahead
        const int nvt = vt + nwg * (int)gridDim.x;
        int nt, nnh;
        item_of(p, nvt, nt, nnh);
        if (in_range(nvt) && nt < p.g.num_tiles) {
          int nn0, nh0, nw0;
          p.g.tile_origin(nt, nn0, nh0, nw0);
          const int nn = nn0 + bn, nhh = p.os * (nh0 + bh) + p.oh0, nw = p.os * (nw0 + bw) + p.ow0;
          if (nn < p.N) {
            const size_t npix = ((size_t)nn * p.Ho + nhh) * p.Wo + nw;
            const float* pf = p.epi.res ? p.epi.res : p.epi.dact;
            if (pf)
              for (int c = 0; c < nchunks; ++c) prefetch_l2(pf + npix * p.Cout + (nnh * nchunks + c) * 32);
          }
        }
      }
      for (int cc = cc0; cc < nchunks; cc += ccstep) {
        const int gc = gc0 + cc;  // block index in the output tensor; cc indexes the accumulator columns
        float v[32], f[32];
        tmem_ld32(t0 + cc * 32, v);
        if (cc + ccstep >= nchunks) {
          // all TMEM reads of this tile are done: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CTA2)
              mbar_arrive_leader(bar_tempty + 8 * acc);
            else
              mbar_arrive(bar_tempty + 8 * acc);
          }
        }
        epilogue_math32(p.epi, v, f, rr, has_res, has_dact, gc * 32);
        if (cc + ccstep < nchunks) {
          if (has_res) load32(p.epi.res + pix * p.Cout + (gc + ccstep) * 32, rr);
          if (has_dact) load32(p.epi.dact + pix * p.Cout + (gc + ccstep) * 32, rr);
        }
        if ((p.debug & 32) || (kAblate & 32)) continue;
        // staging set `sbuf` free again?  (the TMA stores issued sbufs blocks ago have read it)
        const uint32_t staging = staging0 + sbuf * kStagingBytes;
        if (threadIdx.x == issuer) {
          if (p.sbufs == 1)
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          else if (p.sbufs == 2)
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else
            asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
        }
        EPI_BARRIER();
        if (wantF) {
          const uint32_t row = staging + m * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(row + ((j ^ (m & 7)) << 4), __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                         __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
        }
        if (wantO) {
          uint32_t ph[16], pl[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], ph[j], pl[j]);
          const uint32_t rh = staging + kStageF + m * 64, rl = rh + kStageO;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t sw = (j ^ ((m >> 1) & 3)) << 4;
            st_shared_v4(rh + sw, ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
            st_shared_v4(rl + sw, pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
          }
          if (edge) {
            // halo replicas of edge pixels (the TMA box covers the interior position only)
            const size_t blk = (size_t)gc * so.block_stride(), lo_off = so.lo_offset();
            bool first = true;
            for_each_replica(h, w, p.Ho, p.Wo, [&](int hp, int wp) {
              if (first) {
                first = false;
                return;
              }
              bf16* dst = p.epi.outOP + op_offset(so, n, hp, wp) + blk;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                reinterpret_cast<uint4*>(dst)[j] = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
                reinterpret_cast<uint4*>(dst + lo_off)[j] =
                    make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
              }
            });
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        EPI_BARRIER();
        if (threadIdx.x == issuer && !(p.debug & 128) && !(kAblate & 64)) {
          // asynchronous copy-out by the TMA engine: F as rows of the [pixels][C] matrix, OP as the
          // (plane, K-block) image box; up to `sbufs` blocks are in flight behind the epilogue
          // the tensor maps already carry the output placement (stride os, offset (oh0, ow0), halo)
          if (wantF) tma_store_3d(&tmF, staging, gc * 32, w0, n0 * p.H + h0);
          if (wantO) {
            tma_store_5d(&tmO, staging + kStageF, 0, w0, h0, gc, n0);
            tma_store_5d(&tmO, staging + kStageF + kStageO, 0, w0, h0, p.out_chunks + gc, n0);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (++sbuf == p.sbufs) sbuf = 0;
      }
      continue;  // tempty already signalled
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if constexpr (CTA2)
        mbar_arrive_leader(bar_tempty + 8 * acc);
      else
        mbar_arrive(bar_tempty + 8 * acc);
    }
  }
  if (p.staged && threadIdx.x == issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// The hi / lo words of 32 channels as loaded: converting them (load32_hilo) right after the request makes the
// thread wait for the data on the spot, so a request issued ahead of its use stays raw until then.
struct RawHiLo {
  uint4 h[4], l[4];
};
__device__ __forceinline__ void load_raw_hilo(const bf16* __restrict__ hi, const bf16* __restrict__ lo, RawHiLo& r) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    r.h[j] = *reinterpret_cast<const uint4*>(hi + 8 * j);
    r.l[j] = *reinterpret_cast<const uint4*>(lo + 8 * j);
  }
}
__device__ __forceinline__ void convert_hilo(const RawHiLo& r, float (&o)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t aw[4] = {r.h[j].x, r.h[j].y, r.h[j].z, r.h[j].w}, bw[4] = {r.l[j].x, r.l[j].y, r.l[j].z, r.l[j].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[8 * j + 2 * k] = __uint_as_float(aw[k] << 16) + __uint_as_float(bw[k] << 16);
      o[8 * j + 2 * k + 1] = __uint_as_float(aw[k] & 0xFFFF0000u) + __uint_as_float(bw[k] & 0xFFFF0000u);
    }
  }
}

// ------------------------------------------------------------------------------------------ LayerNorm epilogue
// Epilogue with the channel LayerNorm fused in (ConvEpilogue::ln: 1 forward, 2 adjoint), second generation.
// The thread that owns a TMEM lane sees all C_out channels of its pixel, but not at once: the statistics need
// one sweep over the channel blocks and the normalisation a second one.  What round 1 re-read from global
// memory and recomputed in the second sweep now stays on chip:
//   forward   f = acc + bias + res is written back IN PLACE of the accumulator (tcgen05.st) by the first sweep;
//             the second sweep reads it from TMEM, normalises and stores.  One global read of `res`.
//   adjoint   the saved operand a (hi + lo) is parked in the free TMEM columns behind the accumulator when
//             there are any (C_out <= 128 double-buffered, <= 256 single), so the second sweep reads g and a
//             from TMEM and only `res` from global memory.
// Global operands are requested one channel block ahead (and the first block before the accumulator is waited
// for), so their latency overlaps the TMEM traffic and the store staging of the previous block.
// Single accumulator (C_out > 256, patch kernel): the two epilogue warpgroups split the channel blocks of every
// tile (csplit) and exchange their partial statistics through shared memory (xchg) -- both warpgroups of a
// pixel combine the two partials in the same order, so they normalise with identical numbers.
template <int LN, bool CTA2, int PC = 1>
__device__ __forceinline__ void epilogue_ln_role(const UmmaParams& p, const CUtensorMap& tmF, const CUtensorMap& tmO,
                                                 uint32_t bar_tfull, uint32_t bar_tempty, uint32_t tmem_base,
                                                 uint32_t acc_stride, uint32_t staging_base, const float* cst,
                                                 float2* xchg, int tile_begin, int tile_end, int warp, int lane, int wg, int nwg,
                                                 bool csplit, int epi0 = 64) {
  const uint32_t staging0 = staging_base + wg * p.sbufs * kStagingBytes;
  const int issuer = epi0 + 128 * wg;
  const int bar_id = wg;
  const int half = wg;  // csplit: this warpgroup takes channel blocks half, half + 2, ...
  const int cc0 = csplit ? wg : 0, ccstep = csplit ? nwg : 1;
  if (csplit) wg = 0, nwg = 1;  // tile walk of a single warpgroup
  const int q = warp & 3;
  const int m = q * 32 + lane;
  const int bw = m % p.g.BW, bh = (m / p.g.BW) % p.g.BH, bn = m / (p.g.BW * p.g.BH);
  const int C = p.Cout, nch = p.out_chunks;
  const bool wantF = p.epi.outF != nullptr && !(kAblate & 128), wantO = p.epi.outOP != nullptr;
  const OpShape so{0, p.Ho, p.Wo, C, 0};
  const size_t bs = so.block_stride(), lo_off = so.lo_offset();
  const float invC = 1.f / (float)C, invC1 = 1.f / (float)(C - 1);
  // PC > 1: the accumulator of a tile arrives as PC pieces of 128 columns (4 channel blocks) in a ring of 4 TMEM
  // slots (compile-time geometry: the slot arithmetic must not cost the other variants instructions or registers)
  const bool stash = LN == 2 && PC == 1 && (p.acc_stages == 2 ? C <= 128 : C <= 256);
  int it = wg;
  int sbuf = 0;
  for (int tile = tile_begin + wg * (int)gridDim.x; tile < tile_end; tile += nwg * (int)gridDim.x, it += nwg) {
    // Piece j of this tile is accumulator item it * npc + j: TMEM slot and barrier phase follow from that index.
    // The pieces are waited for as the first sweep reaches them and handed back as the second sweep leaves them,
    // so the MMAs of the next tile's pieces run under this tile's epilogue (4 slots of 128 columns hold the 3
    // pieces of a C_out = 384 tile plus the first piece of the next one).
    const int itb = PC == 1 ? it : it * PC;
    auto acc_of = [&](int j) { return PC == 1 ? itb % p.acc_stages : (itb + j) & 3; };
    auto phase_of = [&](int j) { return (uint32_t)(PC == 1 ? (itb / p.acc_stages) & 1 : ((itb + j) >> 2) & 1); };
    int ready = -1;
    auto wait_first = [&]() {  // one accumulator per tile: waited for once, ahead of the first sweep
      if constexpr (PC == 1) {
        mbar_wait(bar_tfull + 8 * acc_of(0), phase_of(0));
        tc_fence_after();
      }
    };
    auto wait_block = [&](int cc) {  // pieces: waited for as the first sweep reaches them
      if constexpr (PC > 1) {
        const int j = cc >> 2;
        if (j > ready) {
          while (ready < j) {
            ++ready;
            mbar_wait(bar_tfull + 8 * acc_of(ready), phase_of(ready));
          }
          tc_fence_after();
        }
      }
    };
    int n0, h0, w0;
    p.g.tile_origin(tile, n0, h0, w0);
    const int n = n0 + bn, h = p.os * (h0 + bh) + p.oh0, w = p.os * (w0 + bw) + p.ow0;
    const bool valid = n < p.N;
    const size_t pix = ((size_t)n * p.Ho + h) * p.Wo + w;
    const bool edge = valid && (h == 0 || h == p.Ho - 1 || w == 0 || w == p.Wo - 1);
    const bool has_res = p.epi.res != nullptr && valid && !(kAblate & 16);
    const float* resp = p.epi.res + pix * C;
    const bf16* a_pix = (LN == 2 && valid) ? p.epi.ln_a + op_offset(so, n, h + 1, w + 1) : nullptr;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t0 = tq + (uint32_t)acc_of(0) * acc_stride;
    auto col_of = [&](int cc) {  // TMEM address of channel block cc of this tile
      if constexpr (PC == 1) return t0 + (uint32_t)cc * 32u;
      return tq + (uint32_t)acc_of(cc >> 2) * 128u + (uint32_t)(cc & 3) * 32u;
    };
    const uint32_t ts = t0 + (uint32_t)C;  // stash columns (adjoint, PC == 1)
    // Pulls the epilogue operands of this thread's pixel of tile `pt` into L2.  Called for the NEXT tile of this
    // warpgroup between the two sweeps of the current one (and for the very first tile at its start): at the
    // 5 TB/s these kernels stream, L2 holds some 25 us of traffic, so a request issued a whole tile period
    // ahead was evicted again before its use (ncu: DRAM reads 2.9 GB above the operand sizes at C = 96).
    auto prefetch_tile = [&](int pt) {
      if (pt >= p.g.num_tiles) return;
      int nn0, nh0, nw0;
      p.g.tile_origin(pt, nn0, nh0, nw0);
      const int nn = nn0 + bn, nh = p.os * (nh0 + bh) + p.oh0, nw = p.os * (nw0 + bw) + p.ow0;
      if (nn >= p.N) return;
      const size_t npix = ((size_t)nn * p.Ho + nh) * p.Wo + nw;
      if (p.epi.res)
        for (int c = cc0; c < nch; c += ccstep) prefetch_l2(p.epi.res + npix * C + c * 32);
      if (LN == 2) {
        const bf16* pa = p.epi.ln_a + op_offset(so, nn, nh + 1, nw + 1);
        for (int c = cc0; c < nch; c += ccstep) {
          prefetch_l2(pa + (size_t)c * bs);
          prefetch_l2(pa + (size_t)c * bs + lo_off);
        }
      }
    };
    if (it == wg) prefetch_tile(tile);

    // output of one 32-channel block through the staging tiles: F <- v (fp32), then OP <- post(v) (bf16 hi / lo)
    auto stage_and_store = [&](int cc, float (&v)[32], auto&& post) {
      if constexpr ((kAblate & 32) != 0) return;
      const uint32_t staging = staging0 + sbuf * kStagingBytes;
      if (threadIdx.x == issuer) {
        if (p.sbufs == 1)
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        else if (p.sbufs == 2)
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else
          asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
      }
      EPI_BARRIER();
      if (wantF) {
        const uint32_t row = staging + m * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(row + ((j ^ (m & 7)) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                       __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
      }
      post(v);
      if (wantO) {
        uint32_t ph[16], pl[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], ph[j], pl[j]);
        const uint32_t rh = staging + kStageF + m * 64, rl = rh + kStageO;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t sw = (j ^ ((m >> 1) & 3)) << 4;
          st_shared_v4(rh + sw, ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
          st_shared_v4(rl + sw, pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
        }
        if (edge) {
          // halo replicas of edge pixels (the TMA box covers the interior position only)
          const size_t blk = (size_t)cc * bs;
          bool first = true;
          for_each_replica(h, w, p.Ho, p.Wo, [&](int hp, int wp) {
            if (first) {
              first = false;
              return;
            }
            bf16* dst = p.epi.outOP + op_offset(so, n, hp, wp) + blk;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              reinterpret_cast<uint4*>(dst)[j] = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
              reinterpret_cast<uint4*>(dst + lo_off)[j] =
                  make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
            }
          });
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      EPI_BARRIER();
      if (threadIdx.x == issuer && !(kAblate & 64)) {
        if (wantF) tma_store_3d(&tmF, staging, cc * 32, w0, n0 * p.H + h0);
        if (wantO) {
          tma_store_5d(&tmO, staging + kStageF, 0, w0, h0, cc, n0);
          tma_store_5d(&tmO, staging + kStageF + kStageO, 0, w0, h0, nch + cc, n0);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (++sbuf == p.sbufs) sbuf = 0;
    };
    // block cc was this warpgroup's last TMEM read of its piece: hand the piece's slot back to the MMA warp
    auto release_after = [&](int cc) {
      const int j = PC == 1 ? 0 : cc >> 2;
      if (cc + ccstep < nch && (PC == 1 || (cc + ccstep) >> 2 == j)) return;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTA2)
          mbar_arrive_leader(bar_tempty + 8 * acc_of(j));
        else
          mbar_arrive(bar_tempty + 8 * acc_of(j));
      }
    };
    // partial statistics of the two warpgroups of a csplit tile -> both partials, in warpgroup order
    auto exchange = [&](float a, float b, float2& p0, float2& p1) {
      float2* slot = xchg + (size_t)((it & 1) * 2) * 128;
      slot[half * 128 + m] = make_float2(a, b);
      asm volatile("bar.sync 3, 256;" ::: "memory");
      const float2 other = slot[(half ^ 1) * 128 + m];
      p0 = half == 0 ? make_float2(a, b) : other;
      p1 = half == 0 ? other : make_float2(a, b);
    };

    if constexpr (LN == 1) {
      // bias and (one shared) shift vector are read from the copies the kernel made in shared memory: as global
      // loads they cost an exposed L2 round trip per use (two epilogue warps per scheduler hide nothing)
      const float* shiftp = !p.epi.ln_shift ? nullptr
                            : p.epi.ln_nt > 1 ? p.epi.ln_shift + (size_t)(valid ? n : 0) * p.epi.ln_shift_stride
                                              : cst + kCstShift;
      const float* biasp = p.epi.bias ? cst : nullptr;
      // two residual blocks are in flight: blocks 0 and 1 are requested before the accumulator is waited for,
      // block c + 2 once block c is consumed
      float rr[2][32];
      if (has_res) {
        load32(resp + cc0 * 32, rr[0]);
        if (cc0 + ccstep < nch) load32(resp + (cc0 + ccstep) * 32, rr[1]);
      }
      wait_first();
      // ---- first sweep: f = acc + bias + res back into TMEM; shifted one-pass statistics of f + shift
      float s1 = 0.f, s2 = 0.f, K = 0.f;
      for (int cb = cc0; cb < nch; cb += 2 * ccstep) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int cc = cb + u * ccstep;
          if (cc < nch) {
            float f[32];
            wait_block(cc);
            tmem_ld32(col_of(cc), f);
            if (biasp) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = *reinterpret_cast<const float4*>(biasp + cc * 32 + j);
                f[j] += b.x, f[j + 1] += b.y, f[j + 2] += b.z, f[j + 3] += b.w;
              }
            }
            if (has_res) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] += rr[u][j];
              if (cc + 2 * ccstep < nch) load32(resp + (cc + 2 * ccstep) * 32, rr[u]);
            }
            tmem_st32(col_of(cc), f);
            if (cc == cc0) K = f[0] + (shiftp ? shiftp[cc * 32] : 0.f);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
              if (shiftp) sh = *reinterpret_cast<const float4*>(shiftp + cc * 32 + j);
              const float d0 = f[j] + sh.x - K, d1 = f[j + 1] + sh.y - K, d2 = f[j + 2] + sh.z - K,
                          d3 = f[j + 3] + sh.w - K;
              s1 += d0, s2 += d0 * d0;
              s1 += d1, s2 += d1 * d1;
              s1 += d2, s2 += d2 * d2;
              s1 += d3, s2 += d3 * d3;
            }
          }
        }
      }
      float mean, m2;
      if (!csplit) {
        mean = K + s1 * invC;
        m2 = fmaxf(s2 - s1 * s1 * invC, 0.f);
      } else {
        const float cnt = (float)(32 * ((nch - cc0 + ccstep - 1) / ccstep));
        float2 p0, p1;
        exchange(K + s1 / cnt, fmaxf(s2 - s1 * s1 / cnt, 0.f), p0, p1);
        const float c0n = (float)(32 * ((nch + 1) / 2)), c1n = (float)(32 * (nch / 2));
        const float dm = p1.x - p0.x;
        mean = (c0n * p0.x + c1n * p1.x) * invC;
        m2 = p0.y + p1.y + dm * dm * (c0n * c1n * invC);
      }
      const float rstd = 1.f / sqrtf(m2 * invC1 + 1e-5f);
      if (valid && p.epi.ln_rstd_out && (!csplit || half == 0)) p.epi.ln_rstd_out[pix] = rstd;
      tmem_st_wait();
      prefetch_tile(tile + nwg * (int)gridDim.x);
      // ---- second sweep: F <- f, OP <- (f + shift - mean) rstd
      for (int cc = cc0; cc < nch; cc += ccstep) {
        float f[32];
        tmem_ld32(col_of(cc), f);
        release_after(cc);
        stage_and_store(cc, f, [&](float (&v)[32]) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (shiftp) sh = *reinterpret_cast<const float4*>(shiftp + cc * 32 + j);
            v[j] = (v[j] + sh.x - mean) * rstd, v[j + 1] = (v[j + 1] + sh.y - mean) * rstd;
            v[j + 2] = (v[j + 2] + sh.z - mean) * rstd, v[j + 3] = (v[j + 3] + sh.w - mean) * rstd;
          }
        });
      }
    } else {
      // ---- adjoint: gx = res + (g - mean_C g - a sum_C(g a) / (C - 1)) rstd
      // Operand requests run one block ahead of their use and stay RAW until then (converting at the request makes
      // the thread wait for the data on the spot); the residual has two blocks in flight.  (This needs the 224
      // registers the epilogue warpgroups own since the roles sit on warpgroup boundaries: under the former
      // 168-register cap the second buffer spilled and the launch got 30 % slower.)
      constexpr bool kNoOperands = (kAblate & 16) != 0;
      RawHiLo raw;
      if (valid && !kNoOperands) load_raw_hilo(a_pix + (size_t)cc0 * bs, a_pix + (size_t)cc0 * bs + lo_off, raw);
      const float rstd = valid ? p.epi.ln_rstd_in[pix] : 1.f;
      wait_first();
      float sg = 0.f, sga = 0.f;
      for (int cc = cc0; cc < nch; cc += ccstep) {
        float g[32], aa[32];
        wait_block(cc);
        tmem_ld32(col_of(cc), g);
        if (valid) {
          if constexpr (kNoOperands) {
#pragma unroll
            for (int j = 0; j < 32; ++j) aa[j] = 0.f;
          } else {
            convert_hilo(raw, aa);
            if (cc + ccstep < nch) {
              const bf16* an = a_pix + (size_t)(cc + ccstep) * bs;
              load_raw_hilo(an, an + lo_off, raw);
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            sg += g[j];
            sga += g[j] * aa[j];
          }
        }
        if (stash) tmem_st32(ts + cc * 32, aa);
      }
      // operands of the second sweep: the first residual block, and (no stash) the first block of a again
      float rr[32];
      if (has_res) load32(resp + cc0 * 32, rr);
      if (!stash && valid && !kNoOperands) load_raw_hilo(a_pix + (size_t)cc0 * bs, a_pix + (size_t)cc0 * bs + lo_off, raw);
      if (csplit) {
        float2 p0, p1;
        exchange(sg, sga, p0, p1);
        sg = p0.x + p1.x, sga = p0.y + p1.y;
      }
      const float mg = sg * invC, beta = sga * invC1;
      if (stash) tmem_st_wait();
      prefetch_tile(tile + nwg * (int)gridDim.x);
      for (int cc = cc0; cc < nch; cc += ccstep) {
        {
          {
            float g[32], a[32];
            tmem_ld32(col_of(cc), g);
            if (stash) {
              tmem_ld32(ts + cc * 32, a);
            } else if (valid) {
              if constexpr (kNoOperands) {
#pragma unroll
                for (int j = 0; j < 32; ++j) a[j] = 0.f;
              } else {
                convert_hilo(raw, a);
                if (cc + ccstep < nch) {
                  const bf16* an = a_pix + (size_t)(cc + ccstep) * bs;
                  load_raw_hilo(an, an + lo_off, raw);
                }
              }
            }
            release_after(cc);
            if (valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) g[j] = (g[j] - mg - a[j] * beta) * rstd + (has_res ? rr[j] : 0.f);
            }
            // the residual of this block is consumed: request the next block's while this one is staged
            if (has_res && cc + ccstep < nch) load32(resp + (cc + ccstep) * 32, rr);
            stage_and_store(cc, g, [](float (&)[32]) {});
          }
        }
      }
    }
  }
  if (threadIdx.x == issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------ kernel
// One lane of the (converged) warp; the compiler keeps the guarded block on the uniform datapath.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, %1;\n"
      "@px mov.s32 %0, 1;\n"
      "}"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

// hi word of the K-major SWIZZLE_64B descriptor (SBO = 512 B, version 1, layout 4); the lo word is
// (smem address >> 4) | (LBO = 1) << 16 and is the only part that changes between MMAs.
constexpr uint32_t kDescHi = (512u >> 4) | (1u << 14) | (4u << 29);
__device__ __forceinline__ uint64_t desc64(uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; }

// PLANES: 2 = bf16x3 (hi and lo planes, 3 MMAs per product), 1 = bf16.  NB: number of N halves.
// CTA2: the two CTAs of a cluster pair compute two adjacent M tiles with ONE tcgen05.mma.cta_group::2
// (M = 256): each CTA stages its own A tile and only HALF of the B rows, which halves the weight
// traffic and the shared-memory reads per MMA -- the kernel is shared-memory-bandwidth bound otherwise.
template <int PLANES, int NB, int LN, bool CTA2>
__global__ void __launch_bounds__(kThreads, 1)
    conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmO,
                     const UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);

  const uint32_t bar_full = base;            // kMaxStages x 8 B
  const uint32_t bar_empty = base + 64;      // kMaxStages x 8 B
  const uint32_t bar_tfull = base + 128;     // 2 x 8 B
  const uint32_t bar_tempty = base + 144;    // 2 x 8 B
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 160);
  const uint32_t staging0 = base + p.ctrl_bytes;          // sbufs x [F 16 KB][hi 8 KB][lo 8 KB]
  const uint32_t stage0 = staging0 + p.sbufs * kStagingBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if constexpr (CTA2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool leader = rank == 0;
  // tile walk: CTA2 pairs take tiles (2 c, 2 c + 1); a pair whose second tile is past the end still
  // runs it as a ghost (out-of-range TMA boxes read zeros, the epilogue masks it)
  const int tile_begin = CTA2 ? 2 * ((int)blockIdx.x >> 1) + (int)rank : (int)blockIdx.x;
  const int tile_end = CTA2 ? p.g.num_tiles + (int)rank : p.g.num_tiles;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, CTA2 ? 8 : 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + 160), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + 160), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  pdl_wait();  // nothing above touches global memory; everything below may read what the previous kernel wrote
  if constexpr (LN != 0) {
    // bias / shift copies for the LayerNorm epilogue (epilogue_ln_role)
    float* cst = reinterpret_cast<float*>(base_ptr + kCtrlBytes);
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) {
      cst[i] = p.epi.bias ? p.epi.bias[i] : 0.f;
      cst[kCstShift + i] = (p.epi.ln_shift && p.epi.ln_nt == 1) ? p.epi.ln_shift[i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();  // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kblocks = p.ntaps * p.nchunk;
  const uint32_t acc_stride = p.acc_stages == 2 ? 256u : 0u;
  constexpr uint32_t b_off = PLANES * kABytes;

  if (warp == 0) {
    // ===================================================================== TMA producer
    // The whole warp walks the (warp-uniform) loop; one elected lane issues the copies.
    int stage = 0;
    uint32_t phase = 0;
    // bytes landing per K-block: this CTA's A planes and B rows; in CTA2 mode the leader's barrier
    // also receives the peer's bytes
    const int cbh = CTA2 ? p.CB / 2 : p.CB;  // B rows of one N half held by this CTA
    const uint32_t tx_bytes = (CTA2 ? 2u : 1u) * PLANES * (kABytes + (uint32_t)(NB * cbh) * 64u);
    const bool s1 = !p.in_s2 || p.in_strided;
    for (int tile = tile_begin; tile < tile_end; tile += gridDim.x) {
      int n0, h0, w0;
      p.g.tile_origin(tile, n0, h0, w0);
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const uint32_t tp = p.tap[tap];
        int ch = h0 + (int)(tp & 255u), cw = w0 + (int)((tp >> 8) & 255u);
        const int cp = (int)((tp >> 16) & 255u);
        // pixel (i, j) of parity image (pr, pc) is padded pixel (2 i + pr, 2 j + pc) of the normal layout: with an
        // element stride of 2 in the tensor map the same box is read from there
        if (p.in_strided) ch = 2 * ch + (cp >> 1), cw = 2 * cw + (cp & 1);
        int brow = (int)(tp >> 24) * p.nchunk * 2 * p.Cout;
        for (int chunk = 0; chunk < p.nchunk; ++chunk, brow += 2 * p.Cout) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (elect_one()) {
            const uint32_t full = bar_full + 8 * stage;
            if (p.debug & 2) {
              if (leader) mbar_arrive(full);
            } else {
              if (leader) mbar_expect_tx(full, tx_bytes);
              const uint32_t sa = stage0 + stage * p.stage_bytes;
#pragma unroll
              for (int pl = 0; pl < PLANES; ++pl) {
                const int q = pl * p.nchunk + chunk;  // (plane, K-block) image of the operand tensor
                if constexpr (CTA2)
                  tma_load_5d_2sm(sa + pl * kABytes, &tmA, full, 0, cw, ch, s1 ? q : q * 4 + cp, n0);
                else
                  tma_load_5d(sa + pl * kABytes, &tmA, full, 0, cw, ch, s1 ? q : q * 4 + cp, n0);
#pragma unroll
                for (int half = 0; half < NB; ++half) {
                  const uint32_t dst = sa + b_off + pl * p.b_plane_bytes + half * cbh * 64;
                  const int row = brow + pl * p.Cout + half * p.CB + (CTA2 ? (int)rank * cbh : 0);
                  if constexpr (CTA2)
                    tma_load_2d_2sm(dst, &tmB, full, 0, row);
                  else
                    tma_load_2d(dst, &tmB, full, 0, row);
                }
              }
            }
          }
          __syncwarp();
          if (++stage == p.stages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // Descriptor lo words are base + compile-time offsets, so a K-block is a straight run of
    // PLANES == 2 ? 6 : 2 (x NB) tcgen05.mma with no address arithmetic in between.
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.CB >> 3) << 17) |
                           (((CTA2 ? 256u : 128u) >> 4) << 24);
    const uint32_t lo0 = ((stage0 & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t stage_u = p.stage_bytes >> 4, bplane_u = p.b_plane_bytes >> 4;
    const uint32_t half_u = (uint32_t)((CTA2 ? p.CB / 2 : p.CB) * 64) >> 4;
    const uint32_t CB = (uint32_t)p.CB;
    const bool no_mma = (p.debug & 1) != 0;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    // in CTA2 mode only the leader issues (for both CTAs); the peer's MMA warp idles
    for (int tile = tile_begin; leader && tile < tile_end; tile += gridDim.x, ++it) {
      const int acc = it % p.acc_stages;
      const uint32_t acc_phase = (it / p.acc_stages) & 1;
      mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc * acc_stride;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = lo0 + stage * stage_u;
          const uint32_t b_lo = a_lo + (b_off >> 4);
          if (!no_mma) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
              for (int pass = 0; pass < (PLANES == 2 ? 3 : 1); ++pass) {
                // pass 0: hi*hi, 1: hi*lo, 2: lo*hi
                const uint32_t al = a_lo + (pass == 2 ? (kABytes >> 4) : 0u) + kk * 2;
                const uint32_t bl = b_lo + (pass == 1 ? bplane_u : 0u) + kk * 2;
#pragma unroll
                for (int half = 0; half < NB; ++half) {
                  if constexpr (CTA2)
                    umma_bf16_2sm(d0 + half * CB, desc64(al), desc64(bl + half * half_u), idesc,
                                  (kk | pass) ? 1u : (uint32_t)(kb != 0));
                  else
                    umma_bf16(d0 + half * CB, desc64(al), desc64(bl + half * half_u), idesc,
                              (kk | pass) ? 1u : (uint32_t)(kb != 0));
                }
              }
            }
          }
          if constexpr (CTA2) {
            umma_commit_2sm(bar_empty + 8 * stage);
            if (kb == kblocks - 1) umma_commit_2sm(bar_tfull + 8 * acc);
          } else {
            umma_commit(bar_empty + 8 * stage);
            if (kb == kblocks - 1) umma_commit(bar_tfull + 8 * acc);
          }
        }
        __syncwarp();
        if (++stage == p.stages) stage = 0, phase ^= 1;
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    if constexpr (LN != 0)
      epilogue_ln_role<LN, CTA2>(p, tmF, tmO, bar_tfull, bar_tempty, tmem_base, acc_stride, staging0,
                                 reinterpret_cast<const float*>(base_ptr + kCtrlBytes), nullptr, tile_begin, tile_end,
                                 warp, lane, 0, 1, false);
    else
      epilogue_role<CTA2>(p, tmF, tmO, bar_tfull, bar_tempty, tmem_base, acc_stride, staging0, tile_begin, tile_end,
                          warp, lane);
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (CTA2) cluster_sync_all();  // neither CTA frees TMEM or exits while the pair is still working
  if (warp == 1) {
    if constexpr (CTA2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ patch kernel
constexpr uint32_t kDescHiPatch = (kPatchSBO >> 4) | (1u << 14) | (4u << 29);
__device__ __forceinline__ uint64_t desc64_patch(uint32_t lo) { return ((uint64_t)kDescHiPatch << 32) | lo; }

// Same roles, epilogue and CTA-pair protocol as conv_umma_kernel, but the K loop runs K-block outer /
// tap inner over TWO rings: the patch ring (a_stages x PLANES x 12 KB, one box per K-block) and the
// weight ring (stages x one (tap, K-block) weight block).  Input traffic from L2 and into shared
// memory drops from 9 x 8 KB to 11.25 KB per (plane, K-block).
template <int PLANES, int NB, int LN, bool CTA2, int PC = 1>
__global__ void __launch_bounds__(kPatchThreads, 1)
    conv_umma_patch_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmO,
                           const UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);

  const uint32_t bar_bfull = base;            // kMaxBStages x 8 B
  const uint32_t bar_bempty = base + 128;     // kMaxBStages x 8 B
  const uint32_t bar_afull = base + 256;      // kMaxAStages x 8 B
  const uint32_t bar_aempty = base + 288;     // kMaxAStages x 8 B
  const uint32_t bar_tfull = base + 320;      // 4 x 8 B
  const uint32_t bar_tempty = base + 352;     // 4 x 8 B
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 384);
  const uint32_t staging0 = base + p.ctrl_bytes;
  const uint32_t aring0 = staging0 + 2 * p.sbufs * kStagingBytes;  // one staging set per epilogue warpgroup
  constexpr uint32_t a_stage_bytes = PLANES * kPatchPlane;
  const uint32_t bring0 = aring0 + p.a_stages * a_stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if constexpr (CTA2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool leader = rank == 0;
  const int tile_begin = CTA2 ? 2 * ((int)blockIdx.x >> 1) + (int)rank : (int)blockIdx.x;
  const int tile_end = CTA2 ? p.g.num_tiles + (int)rank : p.g.num_tiles;
  // single accumulator (C_out > 256): the two epilogue warpgroups split every tile's channel blocks (with a fused
  // LayerNorm they exchange their partial statistics, epilogue_ln_role)
  const bool csplit = (p.acc_stages == 1 || PC > 1 || (LN == 0 && p.csplit)) && p.epi_wgs == 2;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int a = 0; a < 4; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, (CTA2 ? 8 : 4) * (csplit ? 2 : 1));  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + 384), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + 384), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  pdl_wait();  // nothing above touches global memory; everything below may read what the previous kernel wrote
  if constexpr (LN != 0) {
    // bias / shift copies for the LayerNorm epilogue (epilogue_ln_role)
    float* cst = reinterpret_cast<float*>(base_ptr + kCtrlBytes);
    for (int i = threadIdx.x; i < p.Cout; i += blockDim.x) {
      cst[i] = p.epi.bias ? p.epi.bias[i] : 0.f;
      cst[kCstShift + i] = (p.epi.ln_shift && p.epi.ln_nt == 1) ? p.epi.ln_shift[i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t acc_stride = p.acc_stages == 4 ? 128u : p.acc_stages == 2 ? 256u : 0u;
  const int cbh = CTA2 ? p.CB / 2 : p.CB;  // B rows of one N half held by this CTA
  // pair-items of this launch (item_of); for nsplit == 1 "(vt >> 1) < v_end" is "tile < tile_end"
  const int v_end = ((p.g.num_tiles + 1) >> 1) * p.nsplit;

  // (each setmaxnreg sits INSIDE the branch it governs: after a join of the two, ptxas allocates everything that
  // follows under the smaller of the two limits)
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ===================================================================== TMA producer
    // Issue order: patch of K-block g + 1, then the nine weight blocks of K-block g, so that the
    // patch ring runs one K-block ahead of the weight ring.
    const uint32_t a_tx = (CTA2 ? 2u : 1u) * PLANES * kPatchBytes;
    const uint32_t b_tx = (CTA2 ? 2u : 1u) * PLANES * (uint32_t)(NB * cbh) * 64u;
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    int a_vt = tile_begin, a_piece = 0, a_chunk = 0;  // cursor of the patch ring (virtual tile = work item, item_of)
    auto issue_patch = [&]() {
      if ((a_vt >> 1) >= v_end) return;
      int a_tile, a_nh, n0, h0, w0;
      item_of(p, a_vt, a_tile, a_nh);
      p.g.tile_origin(a_tile, n0, h0, w0);
      mbar_wait(bar_aempty + 8 * as, aph ^ 1);
      if (elect_one()) {
        const uint32_t full = bar_afull + 8 * as;
        if (leader) mbar_expect_tx(full, a_tx);
#pragma unroll
        for (int pl = 0; pl < PLANES; ++pl) {
          const uint32_t dst = aring0 + as * a_stage_bytes + pl * kPatchPlane;
          if constexpr (CTA2)
            tma_load_5d_2sm(dst, &tmP, full, 0, w0, h0, pl * p.nchunk + a_chunk, n0);
          else
            tma_load_5d(dst, &tmP, full, 0, w0, h0, pl * p.nchunk + a_chunk, n0);
        }
      }
      __syncwarp();
      if (++as == p.a_stages) as = 0, aph ^= 1;
      // (every accumulator piece of a tile walks the K-blocks again: the patch is fetched once per piece)
      if (++a_chunk == p.nchunk) {
        a_chunk = 0;
        if (++a_piece == PC) a_piece = 0, a_vt += gridDim.x;
      }
    };
    issue_patch();
    for (int vt = tile_begin; (vt >> 1) < v_end; vt += gridDim.x) {
      int tile, nh;
      item_of(p, vt, tile, nh);
      for (int piece = 0; piece < PC; ++piece)
      for (int chunk = 0; chunk < p.nchunk; ++chunk) {
        issue_patch();
        for (int tap = 0; tap < 9; ++tap) {
          const int brow = (tap * p.nchunk + chunk) * 2 * p.Cout;
          mbar_wait(bar_bempty + 8 * bs, bph ^ 1);
          if (elect_one()) {
            const uint32_t full = bar_bfull + 8 * bs;
            if (leader) mbar_expect_tx(full, b_tx);
            const uint32_t sb = bring0 + bs * p.b_stage_bytes;
#pragma unroll
            for (int pl = 0; pl < PLANES; ++pl) {
#pragma unroll
              for (int half = 0; half < NB; ++half) {
                const uint32_t dst = sb + pl * p.b_plane_bytes + half * cbh * 64;
                const int row = brow + pl * p.Cout + (nh + piece + half) * p.CB + (CTA2 ? (int)rank * cbh : 0);
                if constexpr (CTA2)
                  tma_load_2d_2sm(dst, &tmB, full, 0, row);
                else
                  tma_load_2d(dst, &tmB, full, 0, row);
              }
            }
          }
          __syncwarp();
          if (++bs == p.stages) bs = 0, bph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.CB >> 3) << 17) |
                           (((CTA2 ? 256u : 128u) >> 4) << 24);
    const uint32_t a_lo0 = ((aring0 & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t b_lo0 = ((bring0 & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t bstage_u = p.b_stage_bytes >> 4, bplane_u = p.b_plane_bytes >> 4;
    const uint32_t half_u = (uint32_t)(cbh * 64) >> 4;
    const uint32_t CB = (uint32_t)p.CB;
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    int it = 0;
    for (int vi = 0, vt = tile_begin; leader && (vt >> 1) < v_end; ++vi, ++it) {
      // (vi walks the accumulator items: `pieces` consecutive ones per tile)
      if (vi == PC) {
        vi = 0, vt += gridDim.x;
        if ((vt >> 1) >= v_end) break;
      }
      const int acc = it % p.acc_stages;
      const uint32_t acc_phase = (it / p.acc_stages) & 1;
      mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc * acc_stride;
      for (int chunk = 0; chunk < p.nchunk; ++chunk) {
        mbar_wait(bar_afull + 8 * as, aph);
        const uint32_t a_stage_lo = a_lo0 + as * (a_stage_bytes >> 4);
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(bar_bfull + 8 * bs, bph);
          tc_fence_after();
          if (elect_one()) {
            // tap (a, b): start shifted by (10 a + b) rows of 64 B
            const uint32_t a_lo = a_stage_lo + (uint32_t)((tap / 3) * kPatchW + tap % 3) * 4u;
            const uint32_t b_lo = b_lo0 + bs * bstage_u;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
              for (int pass = 0; pass < (PLANES == 2 ? 3 : 1); ++pass) {
                // pass 0: hi*hi, 1: hi*lo, 2: lo*hi
                const uint32_t al = a_lo + (pass == 2 ? (kPatchPlane >> 4) : 0u) + kk * 2;
                const uint32_t bl = b_lo + (pass == 1 ? bplane_u : 0u) + kk * 2;
#pragma unroll
                for (int half = 0; half < NB; ++half) {
                  const uint32_t accum = (kk | pass) ? 1u : (uint32_t)((chunk | tap) != 0);
                  if constexpr ((kAblate & 1) != 0) continue;
                  if constexpr (CTA2)
                    umma_bf16_2sm(d0 + half * CB, desc64_patch(al), desc64(bl + half * half_u), idesc, accum);
                  else
                    umma_bf16(d0 + half * CB, desc64_patch(al), desc64(bl + half * half_u), idesc, accum);
                }
              }
            }
            const bool last_tap = tap == 8;
            if constexpr (CTA2) {
              umma_commit_2sm(bar_bempty + 8 * bs);
              if (last_tap) umma_commit_2sm(bar_aempty + 8 * as);
              if (last_tap && chunk == p.nchunk - 1) umma_commit_2sm(bar_tfull + 8 * acc);
            } else {
              umma_commit(bar_bempty + 8 * bs);
              if (last_tap) umma_commit(bar_aempty + 8 * as);
              if (last_tap && chunk == p.nchunk - 1) umma_commit(bar_tfull + 8 * acc);
            }
          }
          __syncwarp();
          if (++bs == p.stages) bs = 0, bph ^= 1;
        }
        if (++as == p.a_stages) as = 0, aph ^= 1;
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================================================================== epilogue (warps 4..7, 8..11)
    // with a double-buffered accumulator the two warpgroups take alternate tiles
    const int wg = (warp - 4) >> 2, nwg = ((p.acc_stages == 2 || csplit) && p.epi_wgs == 2) ? 2 : 1;
    if (wg < nwg) {
      if constexpr (LN != 0)
        epilogue_ln_role<LN, CTA2, PC>(p, tmF, tmO, bar_tfull, bar_tempty, tmem_base, acc_stride, staging0,
                                       reinterpret_cast<const float*>(base_ptr + kCtrlBytes),
                                       reinterpret_cast<float2*>(base_ptr + kCtrlBytes + kCstBytes), tile_begin, tile_end,
                                       warp, lane, wg, nwg, csplit, kPatchEpi0);
      else
        epilogue_role<CTA2>(p, tmF, tmO, bar_tfull, bar_tempty, tmem_base, acc_stride, staging0, tile_begin, tile_end,
                            warp, lane, wg, nwg, csplit, kPatchEpi0);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (CTA2) cluster_sync_all();
  if (warp == 1) {
    if constexpr (CTA2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
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

int encode(CUtensorMap* map, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
           const cuuint32_t* box, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
           CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_64B, const cuuint32_t* element_strides = nullptr) {
  auto fn = get_encode();
  if (!fn) return fail(SDAB_ERR_DEVICE, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (element_strides)
    for (int i = 0; i < rank; ++i) estr[i] = element_strides[i];
  CUresult r = fn(map, dtype, rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SDAB_ERR_DEVICE, "cuTensorMapEncodeTiled failed with code " + std::to_string(r));
  return SDAB_OK;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace

int conv3x3_umma(const ConvProblem& c, cudaStream_t stream) {
  UmmaParams p{};
  SDAB_TRY(make_tile_geom(c.N, c.H, c.W, p.g));
  SDAB_REQUIRE(c.Cin % 32 == 0 && c.Cin >= 32, "C_in must be padded to a multiple of 32");
  SDAB_REQUIRE(c.Cout % 16 == 0 && c.Cout >= 16 && c.Cout <= 512, "C_out must be a multiple of 16, at most 512");
  SDAB_REQUIRE(c.stride == 1 || c.stride == 2, "stride must be 1 or 2");
  p.N = c.N, p.H = c.H, p.W = c.W, p.Cin = c.Cin, p.Cout = c.Cout, p.stride = c.stride;
  p.in_s2 = c.stride == 2 || c.in_s2;
  p.in_strided = p.in_s2 && c.in_strided;
  p.os = c.os ? c.os : 1, p.oh0 = c.oh0, p.ow0 = c.ow0;
  SDAB_REQUIRE((p.os == 1 && !p.oh0 && !p.ow0) || (p.os == 2 && p.oh0 >= 0 && p.oh0 < 2 && p.ow0 >= 0 && p.ow0 < 2),
               "invalid output placement");
  p.Ho = p.os * c.H, p.Wo = p.os * c.W;
  if (c.taps.n) {
    SDAB_REQUIRE(c.taps.n >= 1 && c.taps.n <= 16, "tap list out of range");
    p.ntaps = c.taps.n;
    for (int t = 0; t < p.ntaps; ++t)
      p.tap[t] = (uint32_t)c.taps.ca[t] | ((uint32_t)c.taps.cb[t] << 8) | ((uint32_t)c.taps.cp[t] << 16) |
                 ((uint32_t)c.taps.wtap[t] << 24);
  } else {
    p.ntaps = 9;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        p.tap[3 * a + b] = p.in_s2 ? ((uint32_t)(a >> 1) | ((uint32_t)(b >> 1) << 8) |
                                      ((uint32_t)((a & 1) * 2 + (b & 1)) << 16) | ((uint32_t)(3 * a + b) << 24))
                                   : ((uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)(3 * a + b) << 24));
  }
  const int wtaps = c.wtaps ? c.wtaps : 9;
  p.nchunk = c.Cin / 32;
  p.planes = c.mode == SDAB_MODE_BF16X3 ? 2 : 1;
  static const int cta2_env = getenv("SDAB_UMMA_CTA2") ? atoi(getenv("SDAB_UMMA_CTA2")) : 1;
  const bool cta2 = cta2_env != 0 && c.Cout % 32 == 0;  // each CTA holds C_out / 2 weight rows (multiple of 16)
  static const int patch_env = getenv("SDAB_UMMA_PATCH") ? atoi(getenv("SDAB_UMMA_PATCH")) : 1;
  static const int nsplit_env = getenv("SDAB_UMMA_NSPLIT") ? atoi(getenv("SDAB_UMMA_NSPLIT")) : 1;
  const bool patch_ok = patch_env && cta2 && c.Cout % 32 == 0 && c.stride == 1 && !p.in_s2 && !c.taps.n && wtaps == 9 &&
                        p.os == 1 && c.W % kPatchBW == 0 && c.H % kPatchBH == 0;
  p.nsplit = 1, p.pieces = 1;
  static const int pieces_env = getenv("SDAB_UMMA_PIECES") ? atoi(getenv("SDAB_UMMA_PIECES")) : 1;
  if (patch_ok && pieces_env && c.epi.ln && c.Cout == 384) {
    // fused LayerNorm at C_out = 384: the statistics span all channels, so the tile stays one epilogue unit, but its
    // accumulator is built as three N = 128 items in a ring of four TMEM slots -- the MMAs of the next tile's
    // pieces run under this tile's epilogue (a single 384-column accumulator made them wait for it)
    p.CB = 128, p.nb = 1, p.acc_stages = 4, p.pieces = 3;  // (kernel template argument PC = 3)
  } else if (c.Cout <= 256) {
    p.CB = c.Cout, p.nb = 1, p.acc_stages = 2;
  } else {
    SDAB_REQUIRE(c.Cout % 32 == 0, "C_out above 256 must be a multiple of 32");
    if (patch_ok && nsplit_env && !c.epi.ln && c.Cout % 64 == 0) {
      // work items of half the output channels (item_of): two accumulator stages of C_out / 2 columns
      p.CB = c.Cout / 2, p.nb = 1, p.acc_stages = 2, p.nsplit = 2;
    } else {
      p.CB = c.Cout / 2, p.nb = 2, p.acc_stages = 1;
    }
  }
  p.b_plane_bytes = (uint32_t)round_up((cta2 ? c.Cout / 2 : c.Cout) / (p.nsplit * p.pieces) * 64, 1024);
  p.stage_bytes = p.planes * (kABytes + p.b_plane_bytes);
  p.staged = c.Cout % 32 == 0;
  p.out_chunks = c.Cout / 32;
  p.item_chunks = p.out_chunks / p.nsplit;
  SDAB_REQUIRE(!(c.epi.outF && c.epi.pre), "a convolution writes either its output or its pre-activation, not both");
  SDAB_REQUIRE(!c.epi.ln || p.staged, "the fused LayerNorm epilogue needs C_out % 32 == 0");
  SDAB_REQUIRE(c.epi.ln != 2 || (c.epi.ln_a && c.epi.ln_rstd_in && !c.epi.bias && !c.epi.act && !c.epi.dact),
               "invalid backward-LayerNorm epilogue");
  p.ctrl_bytes = kCtrlBytes + (c.epi.ln ? kCstBytes : 0);
  p.sbufs = 1;  // measured: deeper store staging does not pay for the ring stages it costs
  if (getenv("SDAB_UMMA_SBUFS")) p.sbufs = atoi(getenv("SDAB_UMMA_SBUFS"));
  SDAB_REQUIRE(p.sbufs >= 1 && p.sbufs <= 3, "staging sets out of range");
  // Patch kernel: plain stride-1 3x3 convolutions (the 36 block convolutions and their input-gradients)
  // on images that tile into 8 x 16 boxes; everything else keeps one TMA box per tap.
  static const int patch_wg = getenv("SDAB_UMMA_WG") ? atoi(getenv("SDAB_UMMA_WG")) : 2;  // epilogue warpgroups
  static const int csplit_env = getenv("SDAB_UMMA_CSPLIT") ? atoi(getenv("SDAB_UMMA_CSPLIT")) : 0;
  p.csplit = csplit_env;
  p.epi_wgs = patch_wg == 2 ? 2 : 1;
  SDAB_REQUIRE(c.epi.ln != 1 || (!c.epi.act && !c.epi.dact && !c.epi.pre),
               "the fused forward LayerNorm follows a plain (bias / residual) convolution");
  p.patch = patch_ok;
  if (p.patch) {
    p.g.BW = kPatchBW, p.g.BH = kPatchBH, p.g.BN = 1;
    p.g.tiles_w = c.W / kPatchBW, p.g.tiles_h = c.H / kPatchBH, p.g.tiles_n = c.N;
    p.g.num_tiles = p.g.tiles_w * p.g.tiles_h * p.g.tiles_n;
    p.b_stage_bytes = p.planes * p.b_plane_bytes;
    // csplit statistics exchange of the LayerNorm epilogue: 2 tile parities x 2 warpgroups x 128 pixels x float2
    if (c.epi.ln && (p.acc_stages == 1 || p.pieces > 1) && patch_wg == 2) p.ctrl_bytes += kXchgBytes;
    const uint32_t fixed = p.ctrl_bytes + 1024 + 2 * p.sbufs * kStagingBytes;
    p.a_stages = 3;
    p.stages = (int)((kSmemBudget - fixed - p.a_stages * p.planes * kPatchPlane) / p.b_stage_bytes);
    if (p.stages < 4) {
      p.a_stages = 2;
      p.stages = (int)((kSmemBudget - fixed - p.a_stages * p.planes * kPatchPlane) / p.b_stage_bytes);
    }
    if (p.stages > kMaxBStages) p.stages = kMaxBStages;
    SDAB_REQUIRE(p.stages >= 2, "convolution does not fit the shared-memory pipeline");
  } else {
    p.stages = (int)((kSmemBudget - p.ctrl_bytes - 1024 - p.sbufs * kStagingBytes) / p.stage_bytes);
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    SDAB_REQUIRE(p.stages >= 2, "convolution does not fit the shared-memory pipeline");
  }
  p.epi = c.epi;
  {
    static const int dbg = getenv("SDAB_UMMA_DEBUG") ? atoi(getenv("SDAB_UMMA_DEBUG")) : 0;
    p.debug = dbg;
  }

  // A: haloed operand tensor at the input resolution, [N][2 * C/32][Hp][Wp][32] (common.cuh)
  const int Hin = p.in_s2 ? 2 * c.H : c.H, Win = p.in_s2 ? 2 * c.W : c.W;
  const cuuint64_t Hp = Hin + 2, Wp = Win + 2, Q = 2 * (cuuint64_t)p.nchunk;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5], strides[4];
    if (!p.in_s2 || p.in_strided) {
      dims[0] = 32, dims[1] = Wp, dims[2] = Hp, dims[3] = Q, dims[4] = (cuuint64_t)c.N;
      strides[0] = 64, strides[1] = Wp * 64, strides[2] = Hp * Wp * 64, strides[3] = Q * Hp * Wp * 64;
    } else {
      // parity images: dim 3 indexes (plane, K-block, parity)
      dims[0] = 32, dims[1] = Wp / 2, dims[2] = Hp / 2, dims[3] = 4 * Q, dims[4] = (cuuint64_t)c.N;
      strides[0] = 64, strides[1] = (Wp / 2) * 64, strides[2] = (Hp / 2) * (Wp / 2) * 64,
      strides[3] = Q * Hp * Wp * 64;
    }
    // patch kernel: the haloed (BW + 2) x (BH + 2) patch of one image
    // (strided: the box spans 2 BW - 1 x 2 BH - 1 pixels of the tensor, every second one is traversed)
    const cuuint32_t sx = p.in_strided ? 2 : 1;
    const cuuint32_t box[5] = {32, sx * (cuuint32_t)(p.patch ? kPatchW : p.g.BW) - (sx - 1),
                               sx * (cuuint32_t)(p.patch ? kPatchH : p.g.BH) - (sx - 1), 1, (cuuint32_t)p.g.BN};
    const cuuint32_t estr[5] = {1, sx, sx, 1, 1};
    SDAB_TRY(encode(&tmA, c.in, 5, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B, estr));
  }
  {
    const cuuint64_t dims[2] = {32, (cuuint64_t)wtaps * p.nchunk * 2 * c.Cout};
    const cuuint64_t strides[1] = {64};
    const cuuint32_t box[2] = {32, (cuuint32_t)(cta2 ? p.CB / 2 : p.CB)};
    SDAB_TRY(encode(&tmB, c.wpk, 2, dims, strides, box));
  }

  // epilogue outputs (staged path).  F: [N * H][W][C_out] view of the fp32 output image with the
  // placement stride / offset folded into the strides / base; OP: the haloed (plane, K-block) images,
  // base at the first interior pixel.  A tile is then one box at (channel block, w0, n0 * H + h0).
  CUtensorMap tmF = tmB, tmO = tmB;
  if (p.staged && (c.epi.outF || c.epi.pre)) {
    float* basep = (c.epi.outF ? c.epi.outF : c.epi.pre) + ((size_t)p.oh0 * p.Wo + p.ow0) * c.Cout;
    const cuuint64_t dims[3] = {(cuuint64_t)c.Cout, (cuuint64_t)c.W, (cuuint64_t)c.N * c.H};
    const cuuint64_t strides[2] = {(cuuint64_t)p.os * c.Cout * 4, (cuuint64_t)p.os * p.Wo * c.Cout * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)p.g.BW, (cuuint32_t)(p.g.BH * p.g.BN)};
    SDAB_TRY(encode(&tmF, basep, 3, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  if (p.staged && c.epi.outOP) {
    const cuuint64_t Hpo = p.Ho + 2, Wpo = p.Wo + 2, Qo = 2 * (cuuint64_t)p.out_chunks;
    const bf16* basep = c.epi.outOP + ((size_t)(p.oh0 + 1) * Wpo + (p.ow0 + 1)) * 32;
    const cuuint64_t dims[5] = {32, (cuuint64_t)c.W, (cuuint64_t)c.H, Qo, (cuuint64_t)c.N};
    const cuuint64_t strides[4] = {(cuuint64_t)p.os * 64, (cuuint64_t)p.os * Wpo * 64, Hpo * Wpo * 64,
                                   Qo * Hpo * Wpo * 64};
    const cuuint32_t box[5] = {32, (cuuint32_t)p.g.BW, (cuuint32_t)p.g.BH, 1, (cuuint32_t)p.g.BN};
    SDAB_TRY(encode(&tmO, basep, 5, dims, strides, box));
  }

  const size_t smem = p.ctrl_bytes + 1024 + (size_t)(p.patch ? 2 : 1) * p.sbufs * kStagingBytes +
                      (p.patch ? (size_t)p.a_stages * p.planes * kPatchPlane + (size_t)p.stages * p.b_stage_bytes
                               : (size_t)p.stages * p.stage_bytes);
  using Kernel = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, UmmaParams);
#define SDAB_K(P, B, L) {conv_umma_kernel<P, B, L, false>, conv_umma_kernel<P, B, L, true>}
  static const Kernel kernels[2][2][3][2] = {{{SDAB_K(1, 1, 0), SDAB_K(1, 1, 1), SDAB_K(1, 1, 2)},
                                             {SDAB_K(1, 2, 0), SDAB_K(1, 2, 1), SDAB_K(1, 2, 2)}},
                                            {{SDAB_K(2, 1, 0), SDAB_K(2, 1, 1), SDAB_K(2, 1, 2)},
                                             {SDAB_K(2, 2, 0), SDAB_K(2, 2, 1), SDAB_K(2, 2, 2)}}};
#undef SDAB_K
#define SDAB_KP(P, B) {conv_umma_patch_kernel<P, B, 0, true>, conv_umma_patch_kernel<P, B, 1, true>, conv_umma_patch_kernel<P, B, 2, true>}
  static const Kernel patch_kernels[2][2][3] = {{SDAB_KP(1, 1), SDAB_KP(1, 2)}, {SDAB_KP(2, 1), SDAB_KP(2, 2)}};
#undef SDAB_KP
  // accumulator pieces (fused LayerNorm at C_out = 384): [planes - 1][ln - 1]
  static const Kernel piece_kernels[2][2] = {
      {conv_umma_patch_kernel<1, 1, 1, true, 3>, conv_umma_patch_kernel<1, 1, 2, true, 3>},
      {conv_umma_patch_kernel<2, 1, 1, true, 3>, conv_umma_patch_kernel<2, 1, 2, true, 3>}};
  static bool attr_set = false;
  if (!attr_set) {
    for (int a = 0; a < 2; ++a)
      for (int l = 0; l < 2; ++l)
        SDAB_CUDA_CHECK(cudaFuncSetAttribute(piece_kernels[a][l], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        for (int l = 0; l < 3; ++l)
          SDAB_CUDA_CHECK(
              cudaFuncSetAttribute(patch_kernels[a][b][l], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        for (int l = 0; l < 3; ++l)
          for (int t = 0; t < 2; ++t)
            SDAB_CUDA_CHECK(
                cudaFuncSetAttribute(kernels[a][b][l][t], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    attr_set = true;
  }
  SDAB_REQUIRE(c.epi.ln >= 0 && c.epi.ln <= 2, "unknown fused LayerNorm variant");
  const Kernel kernel = p.pieces > 1 ? piece_kernels[p.planes - 1][c.epi.ln - 1]
                        : p.patch    ? patch_kernels[p.planes - 1][p.nb - 1][c.epi.ln]
                                     : kernels[p.planes - 1][p.nb - 1][c.epi.ln][cta2 ? 1 : 0];
  if (cta2) {
    const int pairs = (p.g.num_tiles + 1) / 2 * p.nsplit;  // pair-items
    const int clusters = pairs < num_sms() / 2 ? pairs : num_sms() / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters), cfg.blockDim = dim3(p.patch ? kPatchThreads : kThreads), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 2 : 1;
    SDAB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmF, tmO, p));
  } else {
    const int grid = p.g.num_tiles < num_sms() ? p.g.num_tiles : num_sms();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 1 : 0;
    SDAB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmF, tmO, p));
  }
  SDAB_LAUNCH_CHECK("conv_umma_kernel");
  return SDAB_OK;
}

}  // namespace sdab
