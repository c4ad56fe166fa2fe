// score_ops.cu -- HBM-bound kernels around the U-Net: Markov-blanket window maps
// (MCScoreNet.unfold / fold, sda/score.py:146-164, and their adjoints), the predictor-corrector
// updates of VPSDE.sample (sda/score.py:250-261) and the elementwise parts of
// GaussianScore.forward (sda/score.py:387,396).  Index arithmetic is exact; every kernel is a
// coalesced grid-stride loop over the innermost (W) axis.
#include "common.cuh"

namespace sdab {

namespace {

constexpr int kBlock = 256;
constexpr int kReducePartials = 256;

inline int grid_for(size_t n) {
  size_t b = (n + kBlock - 1) / kBlock;
  const size_t cap = 148 * 16;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

// win[b, i, ch, hw]: ch < w*C -> x[b, i + ch / C, ch % C, hw]; else ctx[ch - w*C, hw]
__global__ void unfold_cat_kernel(const float* __restrict__ x, const float* __restrict__ ctx, float* __restrict__ win,
                                  int B, int L, int C, int Cc, size_t HW, int order) {
  const int wdt = 2 * order + 1, nw = L - 2 * order, CH = wdt * C + Cc;
  const size_t total = (size_t)B * nw * CH * HW;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t hw = idx % HW;
    size_t r = idx / HW;
    const int ch = r % CH;
    r /= CH;
    const int i = r % nw;
    const int b = r / nw;
    float v;
    if (ch < wdt * C) {
      const int s = ch / C, c = ch % C;
      v = x[(((size_t)b * L + i + s) * C + c) * HW + hw];
    } else {
      v = ctx[(size_t)(ch - wdt * C) * HW + hw];
    }
    win[idx] = v;
  }
}

// (window, slot) feeding output frame j -- fold map of sda/score.py:157-164
__device__ __forceinline__ void fold_src(int j, int L, int k, int& wi, int& slot) {
  const int nw = L - 2 * k;
  if (j < k) {
    wi = 0, slot = j;
  } else if (j < L - k) {
    wi = j - k, slot = k;
  } else {
    wi = nw - 1, slot = j - (nw - 1);
  }
}

__global__ void fold_kernel(const float* __restrict__ win, float* __restrict__ s, int B, int L, int C, size_t HW,
                            int order) {
  const int wdt = 2 * order + 1, nw = L - 2 * order;
  const size_t total = (size_t)B * L * C * HW;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t hw = idx % HW;
    size_t r = idx / HW;
    const int c = r % C;
    r /= C;
    const int j = r % L;
    const int b = r / L;
    int wi, slot;
    fold_src(j, L, order, wi, slot);
    s[idx] = win[((((size_t)b * nw + wi) * wdt + slot) * C + c) * HW + hw];
  }
}

// adjoint of fold: gwin[b, i, slot*C + c] = gs[b, j, c] when (i, slot) feeds frame j, else 0
__global__ void fold_transpose_kernel(const float* __restrict__ gs, float* __restrict__ gwin, int B, int L, int C,
                                      size_t HW, int order) {
  const int k = order, wdt = 2 * order + 1, nw = L - 2 * order;
  const size_t total = (size_t)B * nw * wdt * C * HW;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t hw = idx % HW;
    size_t r = idx / HW;
    const int c = r % C;
    r /= C;
    const int slot = r % wdt;
    r /= wdt;
    const int i = r % nw;
    const int b = r / nw;
    int j = -1;
    if (slot == k)
      j = i + k;
    else if (i == 0 && slot < k)
      j = slot;
    else if (i == nw - 1 && slot > k)
      j = nw - 1 + slot;
    gwin[idx] = j >= 0 ? gs[(((size_t)b * L + j) * C + c) * HW + hw] : 0.f;
  }
}

// adjoint of unfold (autograd's UnfoldBackward0): gx[b, f, c] = sum_s gwin[b, f - s, s*C + c],
// summed in increasing s -- fixed order, so the result does not depend on how windows were sharded.
__global__ void unfold_transpose_kernel(const float* __restrict__ gwin, float* __restrict__ gx, int B, int L, int C,
                                        int Cc, size_t HW, int order) {
  const int wdt = 2 * order + 1, nw = L - 2 * order, CH = wdt * C + Cc;
  const size_t total = (size_t)B * L * C * HW;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t hw = idx % HW;
    size_t r = idx / HW;
    const int c = r % C;
    r /= C;
    const int f = r % L;
    const int b = r / L;
    float acc = 0.f;
    for (int s = 0; s < wdt; ++s) {
      const int i = f - s;
      if (i >= 0 && i < nw) acc += gwin[(((size_t)b * nw + i) * CH + s * C + c) * HW + hw];
    }
    gx[idx] = acc;
  }
}

// the same sum (same order, so the same bits) four pixels per thread: one (frame, channel) plane per blockIdx.y
__global__ void unfold_transpose4_kernel(const float4* __restrict__ gwin, float4* __restrict__ gx, int L, int C, int Cc,
                                         size_t HW4, int order) {
  const int wdt = 2 * order + 1, nw = L - 2 * order, CH = wdt * C + Cc;
  const int plane = blockIdx.y, c = plane % C, f = (plane / C) % L, b = plane / (C * L);
  float4* dst = gx + (size_t)plane * HW4;
  for (size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < HW4; i4 += (size_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < wdt; ++s) {
      const int i = f - s;
      if (i >= 0 && i < nw) {
        const float4 v = gwin[(((size_t)b * nw + i) * CH + s * C + c) * HW4 + i4];
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
      }
    }
    dst[i4] = acc;
  }
}

// score frame (b, f) <- its position in the gathered shards of the window-sharded evaluation (FoldDst, elementwise.cu)
__global__ void frames_assemble_kernel(const float4* __restrict__ gathered, float4* __restrict__ s, int B, int L, int k,
                                       size_t frame4, int per, int cap) {
  const int nw = L - 2 * k;
  const int frame = blockIdx.y, b = frame / L, f = frame % L;
  const int wi = b * nw + min(max(f - k, 0), nw - 1);
  const int r = wi / per, j = wi - r * per;
  int pos;
  if (f >= k && f < L - k) {
    pos = r * cap + j;
  } else {
    const int e = f < k ? f : k + (f - (L - k));
    pos = r * cap + per + (b - (r * per) / nw) * 2 * k + e;
  }
  const float4* src = gathered + (size_t)pos * frame4;
  float4* dst = s + (size_t)frame * frame4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < frame4; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// ------------------------------------------------------------------------------- Philox4x32-10
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ static void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
  }
  __device__ __forceinline__ void operator()(uint64_t counter, uint32_t (&out)[4]) const {
    uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c, a, b);
      a += 0x9E3779B9u, b += 0xBB67AE85u;
    }
    out[0] = c[0], out[1] = c[1], out[2] = c[2], out[3] = c[3];
  }
};

// four standard normals for the group of elements [4g, 4g+4): Box-Muller on Philox(seed)[offset + g]
__device__ __forceinline__ void normal4(uint64_t seed, uint64_t group, float (&z)[4]) {
  Philox ph{(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  ph(group, r);
  const float u0 = ((float)r[0] + 0.5f) * 2.3283064365386963e-10f;  // (0, 1)
  const float u1 = ((float)r[1] + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = ((float)r[2] + 0.5f) * 2.3283064365386963e-10f;
  const float u3 = ((float)r[3] + 0.5f) * 2.3283064365386963e-10f;
  const float ra = sqrtf(-2.f * __logf(u0)), rb = sqrtf(-2.f * __logf(u2));
  float s, c;
  __sincosf(6.283185307179586f * u1, &s, &c);
  z[0] = ra * c, z[1] = ra * s;
  __sincosf(6.283185307179586f * u3, &s, &c);
  z[2] = rb * c, z[3] = rb * s;
}

__global__ void randn_kernel(float* __restrict__ out, size_t n, uint64_t seed, uint64_t offset) {
  const size_t groups = (n + 3) / 4;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    float z[4];
    normal4(seed, offset + g, z);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * g + j < n) out[4 * g + j] = z[j];
  }
}

__global__ void predict_kernel(float* __restrict__ x, const float* __restrict__ eps, float a, float b, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = a * x[i] + b * eps[i];
}
// four elements per thread (same arithmetic per element)
__global__ void predict4_kernel(float4* __restrict__ x, const float4* __restrict__ eps, float a, float b, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = x[i];
    const float4 e = eps[i];
    v.x = a * v.x + b * e.x, v.y = a * v.y + b * e.y, v.z = a * v.z + b * e.z, v.w = a * v.w + b * e.w;
    x[i] = v;
  }
}

// partial[b][p] = sum over the p-th slice of eps_b^2 (fixed slicing -> deterministic)
__global__ void sumsq_partial_kernel(const float* __restrict__ eps, float* __restrict__ partial, size_t event) {
  __shared__ float red[kBlock];
  const int b = blockIdx.y, p = blockIdx.x;
  const size_t per = (event + kReducePartials - 1) / kReducePartials;
  const size_t lo = (size_t)p * per, hi = lo + per < event ? lo + per : event;
  float acc = 0.f;
  for (size_t i = lo + threadIdx.x; i < hi; i += kBlock) {
    const float e = eps[(size_t)b * event + i];
    acc += e * e;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kBlock / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(size_t)b * kReducePartials + p] = red[0];
}

// x <- x - (delta_b eps + sqrt(2 delta_b) z) sigma,  delta_b = tau / mean(eps_b^2)   score.py:257-261
__global__ void correct_kernel(float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ zin,
                               const float* __restrict__ partial, float tau, float sigma, uint64_t seed,
                               uint64_t offset, size_t event) {
  __shared__ float s_delta;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int p = 0; p < kReducePartials; ++p) tot += partial[(size_t)b * kReducePartials + p];
    s_delta = tau / (tot / (float)event);
  }
  __syncthreads();
  const float delta = s_delta, sq = sqrtf(2.f * delta);
  const size_t groups = (event + 3) / 4;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    float z[4];
    const size_t e0 = (size_t)b * event + 4 * g;
    if (zin) {
#pragma unroll
      for (int j = 0; j < 4; ++j) z[j] = 4 * g + j < event ? zin[e0 + j] : 0.f;
    } else {
      normal4(seed, offset + (size_t)b * groups + g, z);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * g + j < event) x[e0 + j] -= (delta * eps[e0 + j] + sq * z[j]) * sigma;
  }
}

__global__ void tweedie_kernel(const float* __restrict__ x, const float* __restrict__ eps, float mu, float sigma,
                               float* __restrict__ xhat, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    xhat[i] = (x[i] - sigma * eps[i]) / mu;
}
__global__ void tweedie4_kernel(const float4* __restrict__ x, const float4* __restrict__ eps, float mu, float sigma,
                                float4* __restrict__ xhat, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = x[i], e = eps[i];
    xhat[i] = make_float4((v.x - sigma * e.x) / mu, (v.y - sigma * e.y) / mu, (v.z - sigma * e.z) / mu,
                          (v.w - sigma * e.w) / mu);
  }
}

// the same with mu and sigma read from device memory (no host synchronisation on the schedule scalars)
__global__ void tweedie_dev_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ mu,
                                   const float* __restrict__ sigma, float* __restrict__ xhat, size_t n) {
  const float m = *mu, s = *sigma;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    xhat[i] = (x[i] - s * eps[i]) / m;
}

__global__ void axpy_kernel(const float* __restrict__ a, const float* __restrict__ b, float alpha,
                            float* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = a[i] + alpha * b[i];
}

// mean over r x r blocks (KolmogorovFlow.coarsen, sda/mcs.py:340-347)
__global__ void coarsen_kernel(const float* __restrict__ x, float* __restrict__ out, size_t n_img, int H, int W, int r) {
  const int Ho = H / r, Wo = W / r;
  const size_t total = n_img * Ho * Wo;
  const float inv = 1.f / (float)(r * r);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int wo = idx % Wo, ho = (idx / Wo) % Ho;
    const size_t img = idx / ((size_t)Wo * Ho);
    const float* src = x + (img * H + (size_t)ho * r) * W + (size_t)wo * r;
    float acc = 0.f;
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < r; ++j) acc += src[(size_t)i * W + j];
    out[idx] = acc * inv;
  }
}

// central differences with circular wrap (KolmogorovFlow.vorticity, sda/mcs.py:361-375)
__global__ void vorticity_kernel(const float* __restrict__ x, float* __restrict__ out, size_t n_pair, int H, int W) {
  const size_t total = n_pair * H * W;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int w = idx % W, h = (idx / W) % H;
    const size_t pair = idx / ((size_t)W * H);
    const float* u = x + pair * 2 * H * W;
    const float* v = u + (size_t)H * W;
    const int wp = w + 1 == W ? 0 : w + 1, wm = w == 0 ? W - 1 : w - 1;
    const int hp = h + 1 == H ? 0 : h + 1, hm = h == 0 ? H - 1 : h - 1;
    const float du = (u[(size_t)h * W + wp] - u[(size_t)h * W + wm]) * 0.5f;
    const float dv = (v[(size_t)hp * W + w] - v[(size_t)hm * W + w]) * 0.5f;
    out[idx] = du - dv;
  }
}

// adjoint of coarsen: gx[h, w] = g[h / r, w / r] / r^2
__global__ void coarsen_adjoint_kernel(const float* __restrict__ g, float* __restrict__ gx, size_t n_img, int H, int W,
                                       int r) {
  const int Ho = H / r, Wo = W / r;
  const size_t total = n_img * H * W;
  const float inv = 1.f / (float)(r * r);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int w = idx % W, h = (idx / W) % H;
    const size_t img = idx / ((size_t)W * H);
    gx[idx] = g[(img * Ho + h / r) * Wo + w / r] * inv;
  }
}

// adjoint of vorticity: gu[h, w] = (g[h, w-1] - g[h, w+1]) / 2, gv[h, w] = (g[h+1, w] - g[h-1, w]) / 2 (circular)
__global__ void vorticity_adjoint_kernel(const float* __restrict__ g, float* __restrict__ gx, size_t n_pair, int H,
                                         int W) {
  const size_t total = n_pair * H * W;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int w = idx % W, h = (idx / W) % H;
    const size_t pair = idx / ((size_t)W * H);
    const float* gp = g + pair * H * W;
    const int wp = w + 1 == W ? 0 : w + 1, wm = w == 0 ? W - 1 : w - 1;
    const int hp = h + 1 == H ? 0 : h + 1, hm = h == 0 ? H - 1 : h - 1;
    float* gu = gx + pair * 2 * H * W;
    gu[(size_t)h * W + w] = (gp[(size_t)h * W + wm] - gp[(size_t)h * W + wp]) * 0.5f;
    gu[(size_t)(H + h) * W + w] = (gp[(size_t)hp * W + w] - gp[(size_t)hm * W + w]) * 0.5f;
  }
}

// KolmogorovFlow.upsample(mode='bilinear') (sda/mcs.py:349-359): circular pad 1 -> F.interpolate(scale r,
// align_corners=False) -> crop r.  Output o reads padded source coordinate s = (o + 1/2) / r + 1/2 (never
// clamped after the crop): taps floor(s) - 1 and floor(s) of the unpadded image (circular), weight frac(s).
__device__ __forceinline__ void bilinear_taps(int o, int r, int n, int& i0, int& i1, float& lam) {
  const float s = ((float)o + 0.5f) / (float)r + 0.5f;
  const int f = (int)floorf(s);
  lam = s - (float)f;
  i0 = (f - 1 + n) % n, i1 = f % n;
}

__global__ void upsample_bilinear_kernel(const float* __restrict__ x, float* __restrict__ out, size_t n_img, int H, int W,
                                         int r) {
  const int Ho = H * r, Wo = W * r;
  const size_t total = n_img * Ho * Wo;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int wo = idx % Wo, ho = (idx / Wo) % Ho;
    const size_t img = idx / ((size_t)Wo * Ho);
    int h0, h1, w0, w1;
    float lh, lw;
    bilinear_taps(ho, r, H, h0, h1, lh);
    bilinear_taps(wo, r, W, w0, w1, lw);
    const float* src = x + img * H * W;
    const float top = (1.f - lw) * src[(size_t)h0 * W + w0] + lw * src[(size_t)h0 * W + w1];
    const float bot = (1.f - lw) * src[(size_t)h1 * W + w0] + lw * src[(size_t)h1 * W + w1];
    out[idx] = (1.f - lh) * top + lh * bot;
  }
}

// adjoint of upsample_bilinear as a gather (deterministic): source pixel (h, w) collects the outputs whose taps
// include it -- output rows with floor(s) - 1 == h or floor(s) == h, i.e. o in [r (h - 1/2) - 1/2 .., r (h + 3/2) - 1/2)
__global__ void upsample_bilinear_adjoint_kernel(const float* __restrict__ g, float* __restrict__ gx, size_t n_img,
                                                 int H, int W, int r) {
  const int Ho = H * r, Wo = W * r;
  const size_t total = n_img * H * W;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int w = idx % W, h = (idx / W) % H;
    const size_t img = idx / ((size_t)W * H);
    const float* gp = g + img * Ho * Wo;
    float acc = 0.f;
    // candidate outputs: the 2 r rows / columns around the pixel (circular), weights recomputed from the taps
    for (int dh = -r; dh < 2 * r; ++dh) {
      const int ho = ((h * r + dh) % Ho + Ho) % Ho;
      int h0, h1;
      float lh;
      bilinear_taps(ho, r, H, h0, h1, lh);
      const float wh = (h0 == h ? 1.f - lh : 0.f) + (h1 == h ? lh : 0.f);
      if (wh == 0.f) continue;
      for (int dw = -r; dw < 2 * r; ++dw) {
        const int wo = ((w * r + dw) % Wo + Wo) % Wo;
        int w0, w1;
        float lw;
        bilinear_taps(wo, r, W, w0, w1, lw);
        const float ww = (w0 == w ? 1.f - lw : 0.f) + (w1 == w ? lw : 0.f);
        if (ww != 0.f) acc += wh * ww * gp[(size_t)ho * Wo + wo];
      }
    }
    gx[idx] = acc;
  }
}

}  // namespace

}  // namespace sdab

using namespace sdab;

extern "C" {

int sdab_unfold_cat(const float* x, const float* ctx, float* win, int B, int L, int C, int Cc, int H, int W, int order,
                    void* stream) {
  SDAB_REQUIRE(x && win && (Cc == 0 || ctx), "null argument");
  SDAB_REQUIRE(order >= 1 && L >= 2 * order + 1, "trajectory shorter than the window (MCScoreNet.unfold raises too)");
  SDAB_TRY(sdab_device_check());
  const size_t total = (size_t)B * (L - 2 * order) * ((2 * order + 1) * C + Cc) * H * W;
  unfold_cat_kernel<<<grid_for(total), kBlock, 0, (cudaStream_t)stream>>>(x, ctx, win, B, L, C, Cc, (size_t)H * W, order);
  SDAB_LAUNCH_CHECK("unfold_cat_kernel");
  return SDAB_OK;
}

int sdab_fold(const float* win_out, float* s, int B, int L, int C, int H, int W, int order, void* stream) {
  SDAB_REQUIRE(win_out && s, "null argument");
  SDAB_REQUIRE(order >= 1 && L >= 2 * order + 1, "trajectory shorter than the window");
  SDAB_TRY(sdab_device_check());
  const size_t total = (size_t)B * L * C * H * W;
  fold_kernel<<<grid_for(total), kBlock, 0, (cudaStream_t)stream>>>(win_out, s, B, L, C, (size_t)H * W, order);
  SDAB_LAUNCH_CHECK("fold_kernel");
  return SDAB_OK;
}

int sdab_fold_transpose(const float* gs, float* gwin, int B, int L, int C, int H, int W, int order, void* stream) {
  SDAB_REQUIRE(gs && gwin, "null argument");
  SDAB_REQUIRE(order >= 1 && L >= 2 * order + 1, "trajectory shorter than the window");
  SDAB_TRY(sdab_device_check());
  const size_t total = (size_t)B * (L - 2 * order) * (2 * order + 1) * C * H * W;
  fold_transpose_kernel<<<grid_for(total), kBlock, 0, (cudaStream_t)stream>>>(gs, gwin, B, L, C, (size_t)H * W, order);
  SDAB_LAUNCH_CHECK("fold_transpose_kernel");
  return SDAB_OK;
}

int sdab_unfold_transpose_add(const float* gwin, float* gx, int B, int L, int C, int Cc, int H, int W, int order,
                              void* stream) {
  SDAB_REQUIRE(gwin && gx, "null argument");
  SDAB_REQUIRE(order >= 1 && L >= 2 * order + 1, "trajectory shorter than the window");
  SDAB_TRY(sdab_device_check());
  const size_t HW = (size_t)H * W;
  if (HW % 4 == 0 && (size_t)B * L * C <= 65535) {
    const size_t HW4 = HW / 4;
    const int gx4 = (int)((HW4 + kBlock - 1) / kBlock < 64 ? (HW4 + kBlock - 1) / kBlock : 64);
    unfold_transpose4_kernel<<<dim3(gx4, B * L * C), kBlock, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(gwin), reinterpret_cast<float4*>(gx), L, C, Cc, HW4, order);
  } else {
    const size_t total = (size_t)B * L * C * HW;
    unfold_transpose_kernel<<<grid_for(total), kBlock, 0, (cudaStream_t)stream>>>(gwin, gx, B, L, C, Cc, HW, order);
  }
  SDAB_LAUNCH_CHECK("unfold_transpose_kernel");
  return SDAB_OK;
}

int sdab_frames_assemble(const float* gathered, float* s, int B, int L, int C, int H, int W, int order, int per, int cap,
                         void* stream) {
  SDAB_REQUIRE(gathered && s, "null argument");
  SDAB_REQUIRE(order >= 1 && L >= 2 * order + 1 && per >= 1 && cap >= per, "invalid shard geometry");
  SDAB_REQUIRE(((size_t)C * H * W) % 4 == 0, "frame size must be a multiple of 4 floats");
  SDAB_TRY(sdab_device_check());
  const size_t frame4 = (size_t)C * H * W / 4;
  const int gx = (int)((frame4 + kBlock - 1) / kBlock < 32 ? (frame4 + kBlock - 1) / kBlock : 32);
  frames_assemble_kernel<<<dim3(gx, B * L), kBlock, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(gathered), reinterpret_cast<float4*>(s), B, L, order, frame4, per, cap);
  SDAB_LAUNCH_CHECK("frames_assemble_kernel");
  return SDAB_OK;
}

int sdab_vpsde_predict(float* x, const float* eps, float a, float b, size_t n, void* stream) {
  SDAB_REQUIRE(x && eps, "null argument");
  SDAB_TRY(sdab_device_check());
  if (n % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)eps & 15) == 0)
    predict4_kernel<<<grid_for(n / 4), kBlock, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(x),
                                                                         reinterpret_cast<const float4*>(eps), a, b, n / 4);
  else
    predict_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(x, eps, a, b, n);
  SDAB_LAUNCH_CHECK("predict_kernel");
  return SDAB_OK;
}

size_t sdab_vpsde_correct_scratch_floats(int B) { return (size_t)(B > 0 ? B : 0) * kReducePartials; }

int sdab_vpsde_correct(float* x, const float* eps, const float* z, float tau, float sigma, uint64_t seed,
                       uint64_t offset, int B, size_t n, float* scratch, void* stream) {
  SDAB_REQUIRE(x && eps && scratch, "null argument");
  SDAB_REQUIRE(B >= 1 && n % B == 0 && B <= 65535, "n must be a multiple of the batch size");
  SDAB_TRY(sdab_device_check());
  const size_t event = n / B;
  sumsq_partial_kernel<<<dim3(kReducePartials, B), kBlock, 0, (cudaStream_t)stream>>>(eps, scratch, event);
  SDAB_LAUNCH_CHECK("sumsq_partial_kernel");
  int gx = grid_for((event + 3) / 4);
  correct_kernel<<<dim3(gx, B), kBlock, 0, (cudaStream_t)stream>>>(x, eps, z, scratch, tau, sigma, seed, offset, event);
  SDAB_LAUNCH_CHECK("correct_kernel");
  return SDAB_OK;
}

int sdab_randn(float* out, size_t n, uint64_t seed, uint64_t offset, void* stream) {
  SDAB_REQUIRE(out, "null argument");
  SDAB_TRY(sdab_device_check());
  randn_kernel<<<grid_for((n + 3) / 4), kBlock, 0, (cudaStream_t)stream>>>(out, n, seed, offset);
  SDAB_LAUNCH_CHECK("randn_kernel");
  return SDAB_OK;
}

int sdab_tweedie(const float* x, const float* eps, float mu, float sigma, float* xhat, size_t n, void* stream) {
  SDAB_REQUIRE(x && eps && xhat, "null argument");
  SDAB_TRY(sdab_device_check());
  if (n % 4 == 0 && (((uintptr_t)x | (uintptr_t)eps | (uintptr_t)xhat) & 15) == 0)
    tweedie4_kernel<<<grid_for(n / 4), kBlock, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(eps), mu, sigma,
        reinterpret_cast<float4*>(xhat), n / 4);
  else
    tweedie_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(x, eps, mu, sigma, xhat, n);
  SDAB_LAUNCH_CHECK("tweedie_kernel");
  return SDAB_OK;
}

int sdab_tweedie_dev(const float* x, const float* eps, const float* mu, const float* sigma, float* xhat, size_t n,
                     void* stream) {
  SDAB_REQUIRE(x && eps && xhat && mu && sigma, "null argument");
  SDAB_TRY(sdab_device_check());
  tweedie_dev_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(x, eps, mu, sigma, xhat, n);
  SDAB_LAUNCH_CHECK("tweedie_dev_kernel");
  return SDAB_OK;
}

int sdab_axpy(const float* a, const float* b, float alpha, float* out, size_t n, void* stream) {
  SDAB_REQUIRE(a && b && out, "null argument");
  SDAB_TRY(sdab_device_check());
  axpy_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(a, b, alpha, out, n);
  SDAB_LAUNCH_CHECK("axpy_kernel");
  return SDAB_OK;
}

int sdab_coarsen(const float* x, float* out, size_t n_img, int H, int W, int r, void* stream) {
  SDAB_REQUIRE(x && out, "null argument");
  SDAB_REQUIRE(r >= 1 && H % r == 0 && W % r == 0, "coarsening factor must divide the image size");
  SDAB_TRY(sdab_device_check());
  coarsen_kernel<<<grid_for(n_img * (H / r) * (W / r)), kBlock, 0, (cudaStream_t)stream>>>(x, out, n_img, H, W, r);
  SDAB_LAUNCH_CHECK("coarsen_kernel");
  return SDAB_OK;
}

int sdab_vorticity(const float* x, float* out, size_t n_pair, int H, int W, void* stream) {
  SDAB_REQUIRE(x && out, "null argument");
  SDAB_TRY(sdab_device_check());
  vorticity_kernel<<<grid_for(n_pair * H * W), kBlock, 0, (cudaStream_t)stream>>>(x, out, n_pair, H, W);
  SDAB_LAUNCH_CHECK("vorticity_kernel");
  return SDAB_OK;
}

int sdab_coarsen_adjoint(const float* g, float* gx, size_t n_img, int H, int W, int r, void* stream) {
  SDAB_REQUIRE(g && gx, "null argument");
  SDAB_REQUIRE(r >= 1 && H % r == 0 && W % r == 0, "coarsening factor must divide the image size");
  SDAB_TRY(sdab_device_check());
  coarsen_adjoint_kernel<<<grid_for(n_img * H * W), kBlock, 0, (cudaStream_t)stream>>>(g, gx, n_img, H, W, r);
  SDAB_LAUNCH_CHECK("coarsen_adjoint_kernel");
  return SDAB_OK;
}

int sdab_vorticity_adjoint(const float* g, float* gx, size_t n_pair, int H, int W, void* stream) {
  SDAB_REQUIRE(g && gx, "null argument");
  SDAB_TRY(sdab_device_check());
  vorticity_adjoint_kernel<<<grid_for(n_pair * H * W), kBlock, 0, (cudaStream_t)stream>>>(g, gx, n_pair, H, W);
  SDAB_LAUNCH_CHECK("vorticity_adjoint_kernel");
  return SDAB_OK;
}

int sdab_upsample_bilinear(const float* x, float* out, size_t n_img, int H, int W, int r, void* stream) {
  SDAB_REQUIRE(x && out, "null argument");
  SDAB_REQUIRE(r >= 1 && H >= 1 && W >= 1, "invalid upsampling factor");
  SDAB_TRY(sdab_device_check());
  upsample_bilinear_kernel<<<grid_for(n_img * H * r * W * r), kBlock, 0, (cudaStream_t)stream>>>(x, out, n_img, H, W, r);
  SDAB_LAUNCH_CHECK("upsample_bilinear_kernel");
  return SDAB_OK;
}

int sdab_upsample_bilinear_adjoint(const float* g, float* gx, size_t n_img, int H, int W, int r, void* stream) {
  SDAB_REQUIRE(g && gx, "null argument");
  SDAB_REQUIRE(r >= 1 && H >= 2 && W >= 2, "invalid upsampling factor");
  SDAB_TRY(sdab_device_check());
  upsample_bilinear_adjoint_kernel<<<grid_for(n_img * H * W), kBlock, 0, (cudaStream_t)stream>>>(g, gx, n_img, H, W, r);
  SDAB_LAUNCH_CHECK("upsample_bilinear_adjoint_kernel");
  return SDAB_OK;
}

}  // extern "C"
