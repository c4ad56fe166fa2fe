// wgrad.cu -- parameter gradients of the U-Net (training path):
// what torch.autograd computes for nn.Conv2d.weight / .bias and for the time-shift Linears when
// VPSDE.loss (sda/score.py:265-276) is back-propagated through sda/nn.py:18-28,184-206.
//
//   dW[co][ci][a][b] = sum_{n,h,w} g[n,h,w,co] * x[n, s h + a - 1, s w + b - 1, ci]     (circular)
//   db[co]           = sum_{n,h,w} g[n,h,w,co]
//   dshift[n][c]     = sum_{h,w} LN^T(gA)[n,h,w,c] = sum_{h,w} (gx_out - gx_in)[n,h,w,c]
//
// First correct path: fp32 CUDA-core implicit GEMM (K = pixels) over the library's internal
// tensors -- g as F or OP, x as an operand tensor (normal / parity / half-resolution for the
// nearest-x2 tails) or as the saved pre-activation with the activation applied on load.  Partial
// sums of the pixel splits are combined with atomics into caller-zeroed fp32 gradients.
#include "common.cuh"

namespace sdab {

namespace {

constexpr int kWT = 32;        // tile edge: 32 output x 32 input channels, all nine taps
constexpr int kWThreads = 64;  // 8 x 8 threads, 4 x 4 channels each
constexpr int kWPW = 32;       // pixels of one image row per step

__device__ __forceinline__ float op_value(const bf16* op, const OpShape& s, int n, int hp, int wp, int c) {
  const bf16* ptr = op + op_offset(s, n, hp, wp) + (size_t)(c >> 5) * s.block_stride() + (c & 31);
  return __bfloat162float(ptr[0]) + __bfloat162float(ptr[s.lo_offset()]);
}

// GK: 0 = g in F layout, 1 = g as operand tensor.  XK: 0 = x operand at the input resolution (normal
// layout, stride 1), 1 = parity layout (stride 2), 2 = operand at HALF the output resolution (nearest x2
// upsample folded into the convolution), 3 = F tensor at the output resolution with the activation
// applied on load (the input of conv2 of a block is act(saved pre-activation)).
template <int GK, int XK>
__global__ void __launch_bounds__(kWThreads) wgrad_kernel(const WgradProblem p, int PW, int units_w, int total_units) {
  constexpr int S = XK == 1 ? 2 : 1;
  constexpr int XC = S * kWPW + 2;
  __shared__ __align__(16) float gs[kWPW][kWT];
  __shared__ __align__(16) float xs[3][XC][kWT];
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const int ci0 = blockIdx.x * kWT, co0 = blockIdx.y * kWT;
  const int H = p.H, W = p.W;
  const OpShape sg{p.N, H, W, p.Cg, 0};
  const OpShape sx = XK == 0   ? OpShape{p.N, H, W, p.Cx, 0}
                     : XK == 1 ? OpShape{p.N, 2 * H, 2 * W, p.Cx, 1}
                               : OpShape{p.N, H / 2, W / 2, p.Cx, 0};
  float acc[9][4][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};

  for (int u = blockIdx.z; u < total_units; u += gridDim.z) {
    const int wc = u % units_w, h = (u / units_w) % H, n = u / (units_w * H);
    const int w0 = wc * PW;
    __syncthreads();
    for (int idx = tid; idx < PW * kWT; idx += kWThreads) {
      const int px = idx / kWT, c = idx % kWT, co = co0 + c;
      float v = 0.f;
      if (co < p.Cg) v = GK == 0 ? p.gF[(((size_t)n * H + h) * W + w0 + px) * p.Cg + co] : op_value(p.gOP, sg, n, h + 1, w0 + px + 1, co);
      gs[px][c] = v;
    }
    const int xcols = S * PW + 2;
    for (int idx = tid; idx < 3 * xcols * kWT; idx += kWThreads) {
      const int c = idx % kWT, q = (idx / kWT) % xcols, a = idx / (kWT * xcols), ci = ci0 + c;
      const int hh = S * h + a, ww = S * w0 + q;  // padded coordinates at the input resolution
      float v = 0.f;
      if (ci < p.Cx) {
        if (XK == 0 || XK == 1) {
          v = op_value(p.xOP, sx, n, hh, ww, ci);
        } else if (XK == 2) {
          v = op_value(p.xOP, sx, n, ((hh - 1) >> 1) + 1, ((ww - 1) >> 1) + 1, ci);
        } else {
          const int hm = (hh - 1 + H) % H, wm = (ww - 1 + W) % W;
          v = act_fwd(p.xF[(((size_t)n * H + hm) * W + wm) * p.Cx + ci], p.act);
        }
      }
      xs[a][q][c] = v;
    }
    __syncthreads();
    for (int px = 0; px < PW; ++px) {
      const float4 g4 = *reinterpret_cast<const float4*>(&gs[px][4 * ty]);
      const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
      if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) bsum[i] += gv[i];
      }
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const float4 x4 = *reinterpret_cast<const float4*>(&xs[a][S * px + b][4 * tx]);
          const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[3 * a + b][i][j] += gv[i] * xv[j];
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + 4 * ty + i;
    if (co >= p.cout) continue;
    if (p.db && blockIdx.x == 0 && tx == 0) atomicAdd(p.db + co, bsum[i]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + 4 * tx + j;
      if (ci >= p.cin) continue;
      float* dst = p.dw + ((size_t)co * p.cin + ci) * 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) atomicAdd(dst + t, acc[t][i][j]);
    }
  }
}

// dshift[(Nt > 1 ? n : 0) * stride + c] += sum over the pixels of image n of (a - b)[.., c]
__global__ void shift_grad_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dshift,
                                  int stride, int Nt, int HW, int C, int chunks) {
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int per = (HW + chunks - 1) / chunks;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int px = p0; px < p1; ++px) {
      const size_t o = ((size_t)n * HW + px) * C + c;
      s += a[o] - b[o];
    }
    atomicAdd(dshift + (size_t)(Nt > 1 ? n : 0) * stride + c, s);
  }
}

// db[c] += sum over all pixels of an F(C) tensor
__global__ void f_channel_sum_kernel(const float* __restrict__ g, float* __restrict__ db, size_t pixels, int C, int cout,
                                     size_t per) {
  const size_t p0 = (size_t)blockIdx.x * per, p1 = p0 + per < pixels ? p0 + per : pixels;
  for (int c = threadIdx.x; c < cout; c += blockDim.x) {
    float s = 0.f;
    for (size_t px = p0; px < p1; ++px) s += g[px * C + c];
    atomicAdd(db + c, s);
  }
}

}  // namespace

int f_channel_sum(const float* g, float* db, size_t pixels, int C, int cout, cudaStream_t st) {
  const size_t blocks = pixels < 148 * 4 ? pixels : 148 * 4, per = (pixels + blocks - 1) / blocks;
  f_channel_sum_kernel<<<(unsigned)((pixels + per - 1) / per), 128, 0, st>>>(g, db, pixels, C, cout, per);
  SDAB_LAUNCH_CHECK("f_channel_sum_kernel");
  return SDAB_OK;
}

int conv3x3_wgrad(const WgradProblem& p, cudaStream_t st) {
  SDAB_REQUIRE(p.dw && (p.gF || p.gOP) && (p.xOP || p.xF), "null argument");
  SDAB_REQUIRE(p.x_kind >= 0 && p.x_kind <= 3, "unknown operand kind");
  SDAB_REQUIRE(p.x_kind != 2 || (p.H % 2 == 0 && p.W % 2 == 0), "odd resolution");
  const int PW = p.W < kWPW ? p.W : kWPW;
  SDAB_REQUIRE(p.W % PW == 0, "image width must be a multiple of 32 or below it");
  const int units_w = p.W / PW, total = p.N * p.H * units_w;
  const int tiles = ((p.cin + kWT - 1) / kWT) * ((p.cout + kWT - 1) / kWT);
  int splits = (148 * 8 + tiles - 1) / tiles;
  if (splits > total) splits = total;
  if (splits > 65535) splits = 65535;
  const dim3 grid((p.cin + kWT - 1) / kWT, (p.cout + kWT - 1) / kWT, splits);
  const int gk = p.gF ? 0 : 1;
#define SDAB_WG(G, X) wgrad_kernel<G, X><<<grid, kWThreads, 0, st>>>(p, PW, units_w, total)
  switch (gk * 4 + p.x_kind) {
    case 0: SDAB_WG(0, 0); break;
    case 1: SDAB_WG(0, 1); break;
    case 2: SDAB_WG(0, 2); break;
    case 3: SDAB_WG(0, 3); break;
    case 4: SDAB_WG(1, 0); break;
    case 5: SDAB_WG(1, 1); break;
    case 6: SDAB_WG(1, 2); break;
    default: SDAB_WG(1, 3); break;
  }
#undef SDAB_WG
  SDAB_LAUNCH_CHECK("wgrad_kernel");
  return SDAB_OK;
}

int shift_grad(const float* a, const float* b, float* dshift, int stride, int Nt, int N, int H, int W, int C,
               cudaStream_t st) {
  const int HW = H * W;
  int chunks = (148 * 4 + N - 1) / N;
  if (chunks > HW) chunks = HW;
  shift_grad_kernel<<<dim3(chunks, N), 128, 0, st>>>(a, b, dshift, stride, Nt, HW, C, chunks);
  SDAB_LAUNCH_CHECK("shift_grad_kernel");
  return SDAB_OK;
}

}  // namespace sdab
