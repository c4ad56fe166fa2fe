// umma_shift_probe.cu -- does a K-major SWIZZLE_64B tcgen05 shared-memory descriptor accept a start
// address shifted by a multiple of 64 B (one pixel row of the operand) and a stride-byte-offset that
// is not a multiple of the 512 B swizzle period?  If so, ONE haloed (BH+2) x (BW+2) input patch in
// shared memory can serve all nine 3x3 taps of an implicit-GEMM convolution.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_shift_probe umma_shift_probe.cu -lcuda
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int PW = 10, PH = 18, NOUT = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Params {
  int sbo;          // stride byte offset of the A descriptor
  int row_pitch;    // rows of the patch per tile row (tap a shifts by a * row_pitch * 64 B)
  int base_mode;    // 0: base_offset = 0; 1: base_offset = (addr >> 7) & 7
};

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                             float* out, Params p) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base + 1024, sB = sA + 12288, bar = base, bar2 = base + 8, slot = base + 16;
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
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(PW * PH * 64 + NOUT * 64)) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(sA), "l"(&tmA), "r"(bar), "r"(0), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sB), "l"(&tmB), "r"(bar), "r"(0), "r"(0) : "memory");
  }
  {
    uint32_t ok = 0, spins = 0;
    while (!ok && ++spins < (1u << 24)) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NOUT >> 3) << 17) | ((128u >> 4) << 24);
  for (int tap = 0; tap < 9; ++tap) {
    const int a = tap / 3, b = tap % 3;
    if (threadIdx.x == 0) {
      for (int kk = 0; kk < 2; ++kk) {
        const uint32_t aaddr = sA + (a * p.row_pitch + b) * 64 + kk * 32;
        const uint32_t baddr = sB + kk * 32;
        const uint64_t bo = p.base_mode ? (uint64_t)((aaddr >> 7) & 7) : 0;
        const uint64_t adesc = (uint64_t)((aaddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(p.sbo >> 4) << 32) | ((uint64_t)1 << 46) | (bo << 49) | ((uint64_t)4 << 61);
        const uint64_t bdesc = (uint64_t)((baddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
        asm volatile("{.reg .pred q; setp.ne.b32 q, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;}" ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)kk) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar2) : "memory");
    }
    {
      uint32_t ok = 0, spins = 0;
      while (!ok && ++spins < (1u << 24)) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(bar2), "r"((uint32_t)(tap & 1)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(tap * 128 + warp * 32 + lane) * 16 + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  std::vector<__nv_bfloat16> hP(PH * PW * 32), hW(NOUT * 32);
  std::vector<float> fP(hP.size()), fW(hW.size());
  srand(1);
  for (size_t i = 0; i < hP.size(); ++i) fP[i] = (float)(rand() % 17 - 8), hP[i] = __float2bfloat16(fP[i]);
  for (size_t i = 0; i < hW.size(); ++i) fW[i] = (float)(rand() % 9 - 4), hW[i] = __float2bfloat16(fW[i]);
  __nv_bfloat16 *dP, *dW;
  float* dO;
  CK(cudaMalloc(&dP, hP.size() * 2));
  CK(cudaMalloc(&dW, hW.size() * 2));
  CK(cudaMalloc(&dO, 9 * 128 * 16 * 4));
  CK(cudaMemcpy(dP, hP.data(), hP.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  cuuint32_t estr[3] = {1, 1, 1};
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  // variant 0: 2-D patch, tile 8 x 16, SBO = 640; variant 1: row strip of 128 pixels (patch viewed as
  // 180 consecutive pixels), dense SBO = 512, taps shift by b pixels only (a = 0 rows of pitch 0)
  for (int variant = 0; variant < 2; ++variant) {
    CUtensorMap tmA, tmB;
    {
      cuuint64_t dims[3] = {32, PW, PH};
      cuuint64_t strides[2] = {64, PW * 64};
      cuuint32_t box[3] = {32, PW, PH};
      CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dP, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode A failed %d\n", r); return 1; }
    }
    {
      cuuint64_t dims[2] = {32, NOUT};
      cuuint64_t strides[1] = {64};
      cuuint32_t box[2] = {32, NOUT};
      CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dW, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode B failed %d\n", r); return 1; }
    }
    for (int base_mode = 0; base_mode < 2; ++base_mode) {
      Params p{variant == 0 ? 640 : 512, variant == 0 ? PW : 16, base_mode};
      CK(cudaMemset(dO, 0, 9 * 128 * 16 * 4));
      probe<<<1, 128, 32768>>>(tmA, tmB, dO, p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d base_mode %d: kernel failed: %s\n", variant, base_mode, cudaGetErrorString(e)); return 1; }
      std::vector<float> o(9 * 128 * 16);
      CK(cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost));
      for (int tap = 0; tap < 9; ++tap) {
        const int a = tap / 3, b = tap % 3;
        int bad = 0;
        for (int m = 0; m < 128; ++m) {
          // first patch pixel (flat index over the 180 pixels) of GEMM row m
          const int pix = variant == 0 ? ((m / 8 + a) * PW + (m % 8 + b)) : (m + a * 16 + b);
          if (pix >= PW * PH) continue;
          for (int n = 0; n < NOUT; ++n) {
            float ref = 0.f;
            for (int c = 0; c < 32; ++c) ref += fP[pix * 32 + c] * fW[n * 32 + c];
            if (ref != o[(tap * 128 + m) * 16 + n]) ++bad;
          }
        }
        printf("variant %d base_mode %d tap (%d,%d): %s (%d mismatches)\n", variant, base_mode, a, b, bad ? "WRONG" : "ok", bad);
      }
    }
  }
  return 0;
}
