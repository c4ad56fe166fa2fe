// conv_simt.cu -- fp32 CUDA-core implicit-GEMM 3x3 circular convolution.
//
// Cross-check engine (SDAB_ENGINE_SIMT): consumes exactly the same operands (bf16 hi/lo
// activations, packed hi/lo weights) and the same epilogue as the tcgen05 engine, but multiplies
// the reconstructed fp32 values hi+lo with FFMA.  It exists to validate the tensor-core kernel on
// the GPU and is selectable per call; it is not the product path.
#include "common.cuh"
#include "tile_geom.h"

namespace sdab {

namespace {

__global__ void __launch_bounds__(128) conv_simt_kernel(ConvProblem p, TileGeom g) {
  __shared__ float As[128][33];
  __shared__ float Bs[16][33];

  const int t = threadIdx.x;
  const int tile = blockIdx.x;
  const int co0 = blockIdx.y * 16;
  int n0, h0, w0;
  g.tile_origin(tile, n0, h0, w0);
  const int bw = t % g.BW, bh = (t / g.BW) % g.BH, bn = t / (g.BW * g.BH);
  const int n = n0 + bn, h = h0 + bh, w = w0 + bw;
  const bool valid = n < p.N;

  const OpShape sin{p.N, p.H * p.stride, p.W * p.stride, p.Cin, p.stride == 2};
  const int nchunk = p.Cin / 32;
  const bool use_lo = p.mode == SDAB_MODE_BF16X3;

  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;

  for (int tap = 0; tap < 9; ++tap) {
    const int a = tap / 3, b = tap % 3;
    const bf16* src = nullptr;
    if (valid) src = p.in + op_offset(sin, n, p.stride * h + a, p.stride * w + b);
    for (int chunk = 0; chunk < nchunk; ++chunk) {
      __syncthreads();
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        float v = 0.f;
        if (valid) {
          const bf16* sk = src + (size_t)chunk * sin.block_stride() + k;
          v = __bfloat162float(sk[0]);
          if (use_lo) v += __bfloat162float(sk[sin.lo_offset()]);
        }
        As[t][k] = v;
      }
      const bf16* wb = p.wpk + ((size_t)(tap * nchunk + chunk) * 2) * p.Cout * 32;
      for (int i = t; i < 16 * 32; i += 128) {
        const int co = i / 32, k = i % 32;
        float v = __bfloat162float(wb[(size_t)(co0 + co) * 32 + k]);
        if (use_lo) v += __bfloat162float(wb[(size_t)p.Cout * 32 + (size_t)(co0 + co) * 32 + k]);
        Bs[co][k] = v;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const float av = As[t][k];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fmaf(av, Bs[j][k], acc[j]);
      }
    }
  }

  if (valid) {
    const size_t pix = ((size_t)n * p.H + h) * p.W + w;
    epilogue_store16(p.epi, acc, pix, n, h, w, p.H, p.W, p.Cout, co0);
  }
}

}  // namespace

int conv3x3_simt(const ConvProblem& p, cudaStream_t stream) {
  TileGeom g;
  SDAB_TRY(make_tile_geom(p.N, p.H, p.W, g));
  SDAB_REQUIRE(p.Cin % 32 == 0 && p.Cout % 16 == 0, "conv channels must be padded to 32 / 16");
  dim3 grid(g.num_tiles, p.Cout / 16);
  conv_simt_kernel<<<grid, 128, 0, stream>>>(p, g);
  SDAB_LAUNCH_CHECK("conv_simt_kernel");
  return SDAB_OK;
}

}  // namespace sdab
