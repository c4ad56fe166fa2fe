// umma_mnmajor_probe.cu -- descriptor semantics of MN-major (transposed) SWIZZLE_64B operands for
// tcgen05.mma kind::f16, as needed by a tensor-core weight-gradient kernel: both operands arrive as
// [pixel][32 channels] tiles (the GEMM K dimension is the pixel, the channels are M / N), i.e.
// 64 B rows along MN, 8-row groups along K.
//   D[m][n] = sum_k A[k][m] * B[k][n],  A: 4 chunks of [128 pixels][32 ch] (M = 128), B: 1 chunk (N = 32)
// Variants: which of the descriptor offsets (LBO / SBO) carries the chunk stride (8192 B) and which the
// 8-pixel group stride (512 B), and a start shifted by whole pixels (64 B rows).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int PIX = 160, KP = 128;  // pixels loaded per chunk, pixels reduced over

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Params { int lbo, sbo, shift; };

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                             float* out, Params p) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = base, bar2 = base + 8, slot = base + 16;
  const uint32_t sA = base + 1024, chunkA = PIX * 64, sB = sA + 4 * chunkA;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar2));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(5 * chunkA)) : "memory");
    for (int c = 0; c < 4; ++c)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(sA + c * chunkA), "l"(&tmA), "r"(bar), "r"(0), "r"(0), "r"(c) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(sB), "l"(&tmB), "r"(bar), "r"(0), "r"(0), "r"(0) : "memory");
  }
  {
    uint32_t ok = 0, spins = 0;
    while (!ok && ++spins < (1u << 24)) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // D fp32, A/B bf16, A and B MN-major (bits 15, 16), N = 32, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x == 0) {
    for (int ks = 0; ks < KP / 16; ++ks) {
      const uint32_t aaddr = sA + ks * 1024, baddr = sB + p.shift * 64 + ks * 1024;
      const uint64_t adesc = (uint64_t)((aaddr & 0x3FFFFu) >> 4) | ((uint64_t)(p.lbo >> 4) << 16) | ((uint64_t)(p.sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
      const uint64_t bdesc = (uint64_t)((baddr & 0x3FFFFu) >> 4) | ((uint64_t)(p.lbo >> 4) << 16) | ((uint64_t)(p.sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
      asm volatile("{.reg .pred q; setp.ne.b32 q, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;}" ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)ks) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar2) : "memory");
  }
  {
    uint32_t ok = 0, spins = 0;
    while (!ok && ++spins < (1u << 24)) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(bar2), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  // A: [4 chunks][PIX pixels][32 ch], B: [1][PIX][32]
  std::vector<__nv_bfloat16> hA(4 * PIX * 32), hB(PIX * 32);
  std::vector<float> fA(hA.size()), fB(hB.size());
  srand(2);
  for (size_t i = 0; i < hA.size(); ++i) fA[i] = (float)(rand() % 9 - 4), hA[i] = __float2bfloat16(fA[i]);
  for (size_t i = 0; i < hB.size(); ++i) fB[i] = (float)(rand() % 9 - 4), hB[i] = __float2bfloat16(fB[i]);
  __nv_bfloat16 *dA, *dB;
  float* dO;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dO, 128 * 32 * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[3] = {32, PIX, 4};
    cuuint64_t strides[2] = {64, PIX * 64};
    cuuint32_t box[3] = {32, PIX, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dA, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A failed %d\n", r); return 1; }
    cuuint64_t dimsb[3] = {32, PIX, 1};
    r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dB, dimsb, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", r); return 1; }
  }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const int chunk = PIX * 64;
  for (int variant = 0; variant < 2; ++variant)
    for (int shift = 0; shift < 4; ++shift) {
      Params p{variant == 0 ? chunk : 512, variant == 0 ? 512 : chunk, shift == 3 ? 11 : shift};
      CK(cudaMemset(dO, 0, 128 * 32 * 4));
      probe<<<1, 128, 65536>>>(tmA, tmB, dO, p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d shift %d: kernel failed: %s\n", variant, p.shift, cudaGetErrorString(e)); return 1; }
      std::vector<float> o(128 * 32);
      CK(cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
          float ref = 0.f;
          for (int k = 0; k < KP; ++k) ref += fA[((m / 32) * PIX + k) * 32 + m % 32] * fB[(k + p.shift) * 32 + n];
          if (ref != o[m * 32 + n]) ++bad;
        }
      printf("LBO=%d SBO=%d B shifted by %d pixels: %s (%d mismatches)\n", p.lbo, p.sbo, p.shift, bad ? "WRONG" : "ok", bad);
    }
  return 0;
}
