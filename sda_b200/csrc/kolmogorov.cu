// kolmogorov.cu -- Kolmogorov-flow stepper: the arithmetic that sda.mcs.KolmogorovFlow
// (sda/mcs.py:244-338) delegates to jax-cfd's semi_implicit_navier_stokes -- finite-volume
// staggered-grid step (van-Leer limited Lax-Wendroff advection, 5-point diffusion, Kolmogorov
// forcing with linear drag, forward Euler) followed by an FFT-diagonalised pressure projection.
// Algorithm restated in oracle/kolmogorov_oracle.py (parity unpinned by the reference).
//
// Layout: state (E, 2, N, N) fp32 as in the reference (axis -2 = x, axis -1 = y, contiguous).
// Members are processed in PAIRS: the Poisson solve is linear with real symbol, so
// z = div_a + i div_b goes through ONE complex 2-D FFT and q_a = Re, q_b = Im -- no real-FFT
// bookkeeping and no wasted half spectrum.  The forward transforms are decimation-in-frequency
// (natural -> bit-reversed), the inverse ones decimation-in-time (bit-reversed -> natural), so no
// permutation pass exists anywhere; the spectral multiply indexes its symbol through bit reversal.
//
// Kernels per inner step (all members of the ensemble per launch):
//   explicit_step   u*, v* = v + dt F(v)             smem tile + halo 2
//   div_row_fft     z = div(u*, v*) pairs, FFT along y (rows, contiguous)
//   col_solve       FFT along x, multiply by 1 / (lambda_x + lambda_y) / N^2, inverse FFT along x
//   row_ifft        inverse FFT along y -> q (complex pair field)
//   grad_sub        v = v* - grad q
#include <cstdlib>

#include "common.cuh"

namespace sdab {

namespace {

constexpr int kTile = 32;
constexpr int kHalo = 2;
constexpr int kTS = kTile + 2 * kHalo;

__device__ __forceinline__ float fast_div(float a, float b) {
  float q;
  asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(q) : "f"(a), "f"(b));
  return q;
}

// Face value of c through a face with velocity uf (cl, c | cr, cn are the cells around the face):
// upwind - (upwind - LaxWendroff) phi(r), van Leer phi(r) = 2 r / (1 + r) for r > 0, r = a / den with
// the safe denominator den (1 when the central difference vanishes).  phi is evaluated with ONE
// division: 2 r / (1 + r) = 2 a / (den + a), and r > 0 iff a den > 0.
__device__ __forceinline__ float face_value(float cl, float c, float cr, float cn, float uf, float dt_h) {
  const float d = cr - c;
  const float den = d != 0.f ? d : 1.f;
  const bool pos = uf > 0.f;
  const float a = pos ? c - cl : cn - cr;
  const float phi = a * den > 0.f ? fast_div(2.f * a, den + a) : 0.f;
  const float upwind = pos ? c : cr;
  // Lax-Wendroff value: c + (1 - Cn) d / 2 from the left, cr - (1 + Cn) d / 2 from the right -- the
  // same number (cr = c + d), so one expression serves both signs
  const float high = c + 0.5f * (1.f - dt_h * uf) * d;
  return upwind - (upwind - high) * phi;
}

// One forward-Euler update of the explicit terms.  grid: (N/32, N/32, E), block: (32, 8).
// Each thread owns 4 consecutive rows of one column: the flux through the lower x face of a cell is
// the upper-face flux of the previous row (carried in a register) and the flux through its left
// y face comes from the neighbouring lane (warp shuffle), so every face flux is evaluated once.
__global__ void __launch_bounds__(256)
    explicit_step_kernel(const float* __restrict__ uv, float* __restrict__ uvs, int N, float dt, float h, float nu) {
  __shared__ float su[kTS][kTS + 1];
  __shared__ float sv[kTS][kTS + 1];
  __shared__ float sforce[kTile];
  const int e = blockIdx.z;
  const float* u = uv + (size_t)e * 2 * N * N;
  const float* v = u + (size_t)N * N;
  const int i0 = blockIdx.y * kTile, j0 = blockIdx.x * kTile;
  const int mask = N - 1;
  for (int idx = threadIdx.y * 32 + threadIdx.x; idx < kTS * kTS; idx += 256) {
    const int li = idx / kTS, lj = idx % kTS;
    const int gi = (i0 + li - kHalo) & mask, gj = (j0 + lj - kHalo) & mask;
    su[li][lj] = u[(size_t)gi * N + gj];
    sv[li][lj] = v[(size_t)gi * N + gj];
  }
  // Kolmogorov forcing sin(4 y) at u's offset y_{j+1/2} (constant along x)
  if (threadIdx.y == 0) sforce[threadIdx.x] = sinf(4.f * ((float)(j0 + threadIdx.x) + 0.5f) * h);
  __syncthreads();
  const float dt_h = dt / h, inv_h = 1.f / h, inv_h2 = 1.f / (h * h);
  const int lj = threadIdx.x + kHalo;
  const int lane = threadIdx.x;
  float* us = uvs + (size_t)e * 2 * N * N;
  float fxu_prev = 0.f, fxv_prev = 0.f;
#pragma unroll
  for (int k = -1; k < 4; ++k) {
    // k = -1 only produces the x-face fluxes below the first owned row
    const int r = 4 * threadIdx.y + k;
    const int li = r + kHalo;
#define U(di, dj) su[li + (di)][lj + (dj)]
#define V(di, dj) sv[li + (di)][lj + (dj)]
    // fluxes through the upper x face (between rows r and r + 1)
    const float ufu = 0.5f * (U(0, 0) + U(1, 0));  // u at (3/2, 1/2) for the u equation
    const float fxu = face_value(U(-1, 0), U(0, 0), U(1, 0), U(2, 0), ufu, dt_h) * ufu;
    const float ufv = 0.5f * (U(0, 0) + U(0, 1));  // u at (1, 1) for the v equation
    const float fxv = face_value(V(-1, 0), V(0, 0), V(1, 0), V(2, 0), ufv, dt_h) * ufv;
    if (k >= 0) {
      // fluxes through the right y face (between columns j and j + 1); the left one from lane - 1
      const float vfu = 0.5f * (V(0, 0) + V(1, 0));  // v at (1, 1) for the u equation
      const float fyu = face_value(U(0, -1), U(0, 0), U(0, 1), U(0, 2), vfu, dt_h) * vfu;
      const float vfv = 0.5f * (V(0, 0) + V(0, 1));  // v at (1/2, 3/2) for the v equation
      const float fyv = face_value(V(0, -1), V(0, 0), V(0, 1), V(0, 2), vfv, dt_h) * vfv;
      float fyu_m = __shfl_up_sync(0xffffffffu, fyu, 1), fyv_m = __shfl_up_sync(0xffffffffu, fyv, 1);
      if (lane == 0) {
        const float vfu_m = 0.5f * (V(0, -1) + V(1, -1));
        fyu_m = face_value(U(0, -2), U(0, -1), U(0, 0), U(0, 1), vfu_m, dt_h) * vfu_m;
        const float vfv_m = 0.5f * (V(0, -1) + V(0, 0));
        fyv_m = face_value(V(0, -2), V(0, -1), V(0, 0), V(0, 1), vfv_m, dt_h) * vfv_m;
      }
      const float conv_u = -((fxu - fxu_prev) + (fyu - fyu_m)) * inv_h;
      const float lap_u = (U(1, 0) + U(-1, 0) + U(0, 1) + U(0, -1) - 4.f * U(0, 0)) * inv_h2;
      const float force_u = sforce[lane] - 0.1f * U(0, 0);
      const float conv_v = -((fxv - fxv_prev) + (fyv - fyv_m)) * inv_h;
      const float lap_v = (V(1, 0) + V(-1, 0) + V(0, 1) + V(0, -1) - 4.f * V(0, 0)) * inv_h2;
      const float force_v = -0.1f * V(0, 0);
      const int gi = i0 + r, gj = j0 + lane;
      us[(size_t)gi * N + gj] = U(0, 0) + dt * (conv_u + nu * lap_u + force_u);
      us[(size_t)N * N + (size_t)gi * N + gj] = V(0, 0) + dt * (conv_v + nu * lap_v + force_v);
    }
    fxu_prev = fxu, fxv_prev = fxv;
#undef U
#undef V
  }
}

// ------------------------------------------------------------------------------------- FFT
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// tw[k] = exp(-2 pi i k / N), k < N/2
__device__ __forceinline__ void fill_twiddles(float2* tw, int N, int tid, int nthreads) {
  for (int k = tid; k < N / 2; k += nthreads) {
    float s, c;
    sincospif(-2.f * (float)k / (float)N, &s, &c);
    tw[k] = make_float2(c, s);
  }
}

// Forward DIF FFT (natural in, bit-reversed out) of one length-N line held in smem with element
// stride `st`; executed by N/2 threads (t = 0 .. N/2-1).  Caller syncs before and after.
__device__ __forceinline__ void fft_dif(float2* s, int st, int N, const float2* tw, int t) {
  for (int half = N >> 1; half >= 1; half >>= 1) {
    const int pos = t & (half - 1);
    const int i0 = ((t - pos) << 1) + pos, i1 = i0 + half;
    const float2 a = s[i0 * st], b = s[i1 * st];
    s[i0 * st] = make_float2(a.x + b.x, a.y + b.y);
    s[i1 * st] = cmul(make_float2(a.x - b.x, a.y - b.y), tw[pos * (N / (2 * half))]);
    __syncthreads();
  }
}

// Inverse DIT FFT (bit-reversed in, natural out, unnormalised).
__device__ __forceinline__ void ifft_dit(float2* s, int st, int N, const float2* tw, int t) {
  for (int half = 1; half < N; half <<= 1) {
    const int pos = t & (half - 1);
    const int i0 = ((t - pos) << 1) + pos, i1 = i0 + half;
    float2 w = tw[pos * (N / (2 * half))];
    w.y = -w.y;
    const float2 a = s[i0 * st], b = cmul(s[i1 * st], w);
    s[i0 * st] = make_float2(a.x + b.x, a.y + b.y);
    s[i1 * st] = make_float2(a.x - b.x, a.y - b.y);
    __syncthreads();
  }
}

// Rows handled per block by the row kernels.
constexpr int kRows = 4;

// z = div(u*, v*) of a member pair, then FFT along y.  grid: (N / kRows, npairs), block: (N/2, kRows)
// mode 0: divergence of uvs (projection); mode 1: z = uvs.u + i uvs.v of ONE member per "pair"
// (prior filtering; `E` members, one per blockIdx.y).
__global__ void div_row_fft_kernel(const float* __restrict__ uvs, float2* __restrict__ spec, int N, int E, float h,
                                   int mode) {
  extern __shared__ float2 smem[];
  float2* tw = smem;                 // N/2
  float2* line = smem + N / 2 + (size_t)threadIdx.y * N;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  fill_twiddles(tw, N, tid, blockDim.x * blockDim.y);
  const int i = blockIdx.x * kRows + threadIdx.y;
  const int pair = blockIdx.y;
  const int mask = N - 1;
  const float inv_h = 1.f / h;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    float2 z;
    if (mode == 0) {
      float d[2];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int e = 2 * pair + m;
        if (e < E) {
          const float* us = uvs + (size_t)e * 2 * N * N;
          const float* vs = us + (size_t)N * N;
          d[m] = ((us[(size_t)i * N + j] - us[(size_t)((i - 1) & mask) * N + j]) +
                  (vs[(size_t)i * N + j] - vs[(size_t)i * N + ((j - 1) & mask)])) *
                 inv_h;
        } else {
          d[m] = 0.f;
        }
      }
      z = make_float2(d[0], d[1]);
    } else {
      const float* us = uvs + (size_t)pair * 2 * N * N;
      z = make_float2(us[(size_t)i * N + j], us[(size_t)N * N + (size_t)i * N + j]);
    }
    line[j] = z;
  }
  __syncthreads();
  fft_dif(line, 1, N, tw, threadIdx.x);
  float2* dst = spec + ((size_t)pair * N + i) * N;
  for (int j = threadIdx.x; j < N; j += blockDim.x) dst[j] = line[j];
}

__device__ __forceinline__ int bitrev(int x, int log2n) { return (int)(__brev((unsigned)x) >> (32 - log2n)); }

// FFT along x, spectral multiply, inverse FFT along x.  grid: (N / CT, npairs), block: (CT, N/2).
// symbol 0: pseudo-inverse Laplacian 1 / (lambda[kx] + lambda[ky]) (zero mode -> 0), lambda[k] =
//           -4 sin^2(pi k / N) / h^2 (eigenvalues of the periodic 3-point second difference);
// symbol 1: sqrt of the log-normal spectral density of filtered_velocity_field (peak wavenumber 4).
__global__ void col_solve_kernel(float2* __restrict__ spec, int N, int log2n, int CT, float h, int symbol) {
  extern __shared__ float2 smem[];
  float2* tw = smem;            // N/2
  float2* tile = smem + N / 2;  // [CT][N + 1]
  const int c = threadIdx.x, t = threadIdx.y;
  const int tid = t * CT + c;
  fill_twiddles(tw, N, tid, CT * (N / 2));
  const int py = blockIdx.x * CT + c;
  float2* base = spec + (size_t)blockIdx.y * N * N;
  float2* col = tile + (size_t)c * (N + 1);
  for (int i = t; i < N; i += N / 2) col[i] = base[(size_t)i * N + py];
  __syncthreads();
  fft_dif(col, 1, N, tw, t);
  const int ky = bitrev(py, log2n);
  const float scale = 1.f / ((float)N * (float)N);
  for (int p = t; p < N; p += N / 2) {
    const int kx = bitrev(p, log2n);
    float f;
    if (symbol == 0) {
      const float sx = sinpif((float)kx / (float)N), sy = sinpif((float)ky / (float)N);
      const float lam = -4.f * (sx * sx + sy * sy) / (h * h);
      f = (kx | ky) ? scale / lam : 0.f;
    } else {
      const int mx = kx < N / 2 ? kx : kx - N, my = ky < N / 2 ? ky : ky - N;
      const float kk = sqrtf((float)(mx * mx + my * my));
      if (kk > 0.f) {
        const float variance = 0.25f, mean = logf(4.f) + variance;
        const float lk = logf(kk);
        const float dens = expf(-(mean - lk) * (mean - lk) / (2.f * variance) - lk) / sqrtf(6.283185307179586f * variance) / kk;
        f = sqrtf(dens) * scale;
      } else {
        f = 0.f;
      }
    }
    col[p].x *= f, col[p].y *= f;
  }
  __syncthreads();
  ifft_dit(col, 1, N, tw, t);
  for (int i = t; i < N; i += N / 2) base[(size_t)i * N + py] = col[i];
}

// inverse FFT along y, in place.  grid: (N / kRows, npairs), block: (N/2, kRows)
__global__ void row_ifft_kernel(float2* __restrict__ spec, int N) {
  extern __shared__ float2 smem[];
  float2* tw = smem;
  float2* line = smem + N / 2 + (size_t)threadIdx.y * N;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  fill_twiddles(tw, N, tid, blockDim.x * blockDim.y);
  const int i = blockIdx.x * kRows + threadIdx.y;
  float2* row = spec + ((size_t)blockIdx.y * N + i) * N;
  for (int j = threadIdx.x; j < N; j += blockDim.x) line[j] = row[j];
  __syncthreads();
  ifft_dit(line, 1, N, tw, threadIdx.x);
  for (int j = threadIdx.x; j < N; j += blockDim.x) row[j] = line[j];
}

// v = v* - grad q (forward differences).  q of member e is the real (even e) / imaginary (odd e)
// part of the pair field.  One thread per grid point and member.
__global__ void grad_sub_kernel(const float* __restrict__ uvs, const float2* __restrict__ qz, float* __restrict__ uv,
                                int N, int E, float h) {
  const size_t total = (size_t)E * N * N;
  const int mask = N - 1;
  const float inv_h = 1.f / h;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int j = idx & mask, i = (idx / N) & mask;
    const int e = idx / ((size_t)N * N);
    const float* q = reinterpret_cast<const float*>(qz + (size_t)(e >> 1) * N * N) + (e & 1);
    const float q0 = q[2 * ((size_t)i * N + j)];
    const float qx = q[2 * ((size_t)((i + 1) & mask) * N + j)];
    const float qy = q[2 * ((size_t)i * N + ((j + 1) & mask))];
    const size_t o = (size_t)e * 2 * N * N + (size_t)i * N + j;
    uv[o] = uvs[o] - (qx - q0) * inv_h;
    uv[o + (size_t)N * N] = uvs[o + (size_t)N * N] - (qy - q0) * inv_h;
  }
}

// prior helpers ------------------------------------------------------------------------------
__global__ void unpack_pair_kernel(const float2* __restrict__ z, float* __restrict__ uv, int N, int E) {
  const size_t total = (size_t)E * N * N;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t e = idx / ((size_t)N * N), r = idx % ((size_t)N * N);
    const float2 v = z[idx];
    uv[e * 2 * N * N + r] = v.x;
    uv[e * 2 * N * N + (size_t)N * N + r] = v.y;
  }
}

// rescale every member to maximum speed `vmax` (one block per member)
__global__ void normalize_speed_kernel(float* __restrict__ uv, int N, float vmax) {
  __shared__ float red[1024];
  float* u = uv + (size_t)blockIdx.x * 2 * N * N;
  float* v = u + (size_t)N * N;
  float m = 0.f;
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) m = fmaxf(m, u[i] * u[i] + v[i] * v[i]);
  red[threadIdx.x] = m;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  const float scale = vmax / sqrtf(red[0]);
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) u[i] *= scale, v[i] *= scale;
}


// ============================================================================================
// Fast path (N = RX^2, RX in {8, 16}): three kernels per inner step, every FFT a two-pass
// radix-RX transform held in registers (one shared-memory exchange per line instead of log2 N
// synchronised radix-2 passes), and each streaming pass fused with its neighbours:
//   fused_explicit_rowfft   u*, v* = v + dt F(v) for a slab of rows of BOTH members of a pair,
//                           z = div(u*, v*)_a + i div(u*, v*)_b, FFT along y   (was 2 kernels)
//   fused_col_solve         FFT along x, pseudo-inverse Laplacian, inverse FFT along x
//   fused_rowifft_grad      inverse FFT along y, v = v* - grad q                (was 2 kernels)
// Spectra are stored in natural frequency order (the two-pass split k = k1 + RX k2 lands there
// without a permutation).  The host walks the ensemble in chunks whose working set stays in L2.
// ============================================================================================
__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// d * exp(SIGN 2 pi i k / 16), k a compile-time constant after unrolling
template <int SIGN>
__device__ __forceinline__ float2 mul_w16(float2 d, int k) {
  constexpr float s = (float)SIGN;
  switch (k) {
    case 0: return d;
    case 4: return make_float2(-s * d.y, s * d.x);
    case 2: return make_float2(0.70710678118654752f * (d.x - s * d.y), 0.70710678118654752f * (s * d.x + d.y));
    case 6: return make_float2(-0.70710678118654752f * (d.x + s * d.y), 0.70710678118654752f * (s * d.x - d.y));
    case 1: return cmul(d, make_float2(0.92387953251128674f, s * 0.38268343236508977f));
    case 3: return cmul(d, make_float2(0.38268343236508977f, s * 0.92387953251128674f));
    case 5: return cmul(d, make_float2(-0.38268343236508977f, s * 0.92387953251128674f));
    default: return cmul(d, make_float2(-0.92387953251128674f, s * 0.38268343236508977f));
  }
}

template <int R, int HALF, int SIGN>
__device__ __forceinline__ void fft_stage(float2 (&a)[R]) {
#pragma unroll
  for (int g = 0; g < R; g += 2 * HALF) {
#pragma unroll
    for (int pos = 0; pos < HALF; ++pos) {
      const float2 x = a[g + pos], y = a[g + pos + HALF];
      a[g + pos] = x + y;
      a[g + pos + HALF] = mul_w16<SIGN>(x - y, pos * (8 / HALF));
    }
  }
}

// In-register DFT of R = 8 or 16 points, X[k] = sum_n a[n] exp(SIGN 2 pi i n k / R), natural order
// in and out (the bit reversal of the decimation-in-frequency passes is a register renaming).
template <int R, int SIGN>
__device__ __forceinline__ void fft_reg(float2 (&a)[R]) {
  static_assert(R == 8 || R == 16, "radix");
  if constexpr (R == 16) fft_stage<R, 8, SIGN>(a);
  fft_stage<R, 4, SIGN>(a);
  fft_stage<R, 2, SIGN>(a);
  fft_stage<R, 1, SIGN>(a);
  float2 b[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    int r = 0;
#pragma unroll
    for (int bit = 1, rb = R >> 1; bit < R; bit <<= 1, rb >>= 1)
      if (k & bit) r |= rb;
    b[k] = a[r];
  }
#pragma unroll
  for (int k = 0; k < R; ++k) a[k] = b[k];
}

// tw[m] = exp(-2 pi i m / N)
__device__ __forceinline__ void fill_twiddles_full(float2* tw, int N, int tid, int nthreads) {
  for (int m = tid; m < N; m += nthreads) {
    float s, c;
    sincospif(-2.f * (float)m / (float)N, &s, &c);
    tw[m] = make_float2(c, s);
  }
}

// Two-pass length-N = RX^2 transform of one line by RX cooperating threads (same warp, lanes
// t = 0 .. RX-1 of an aligned group).  `x` is a line buffer of RX (RX + 1) float2.
//   forward : in  a[n1] = x[RX n1 + t]      out a[k2] = X[t + RX k2]
//   inverse : in  a[k2] = X[t + RX k2]      out a[n1] = x[RX n1 + t]      (unnormalised)
// `grp` is the lane mask of the RX cooperating threads (the only ones that synchronise).
__device__ __forceinline__ unsigned line_group_mask(int rx) {
  const int lane = threadIdx.x & 31;
  return (rx == 32 ? 0xffffffffu : ((1u << rx) - 1u)) << (lane & ~(rx - 1));
}
template <int RX>
__device__ __forceinline__ void line_fft_forward(float2 (&a)[RX], float2* x, const float2* tw, int t, unsigned grp) {
  fft_reg<RX, -1>(a);  // over n1 -> k1
#pragma unroll
  for (int k1 = 0; k1 < RX; ++k1) x[t * (RX + 1) + k1] = cmul(a[k1], tw[t * k1]);
  __syncwarp(grp);
#pragma unroll
  for (int n2 = 0; n2 < RX; ++n2) a[n2] = x[n2 * (RX + 1) + t];
  fft_reg<RX, -1>(a);  // over n2 -> k2
}
template <int RX>
__device__ __forceinline__ void line_fft_inverse(float2 (&a)[RX], float2* x, const float2* tw, int t, unsigned grp) {
  fft_reg<RX, 1>(a);  // over k2 -> n2
#pragma unroll
  for (int n2 = 0; n2 < RX; ++n2) {
    float2 w = tw[t * n2];
    w.y = -w.y;
    x[n2 * (RX + 1) + t] = cmul(a[n2], w);
  }
  __syncwarp(grp);
#pragma unroll
  for (int k1 = 0; k1 < RX; ++k1) a[k1] = x[t * (RX + 1) + k1];
  fft_reg<RX, 1>(a);  // over k1 -> n1
}

template <int RX, int RB>
struct FusedGeom {
  static constexpr int N = RX * RX;
  static constexpr int kSlabRows = RB + 5;        // rows i0-3 .. i0+RB+1
  static constexpr int kLine = RX * (RX + 1);     // float2 per line buffer
  static constexpr int kWarps = N / 32;
  static constexpr size_t slab_bytes = (size_t)2 * kSlabRows * N * sizeof(float);
  static constexpr size_t line_bytes = (size_t)RB * kLine * sizeof(float2);
  static constexpr size_t misc_bytes = (size_t)N * sizeof(float2) + (size_t)N * sizeof(float) +
                                       (size_t)2 * kWarps * (RB + 1) * sizeof(float) + (size_t)kWarps * RB * sizeof(float);
  static constexpr size_t smem_a = slab_bytes + line_bytes + misc_bytes;
  // packed pair march: both members' slabs, pair-wide flux / v* exchange rows (three blocks per SM: <= 74.6 KB)
  static constexpr size_t smem_a2 = 2 * slab_bytes + line_bytes + misc_bytes +
                                    (size_t)2 * kWarps * (RB + 1) * sizeof(float) + (size_t)kWarps * RB * sizeof(float);
  static constexpr int kThreadsC = ((RX * (RB + 1) > N ? RX * (RB + 1) : N) + 31) / 32 * 32;
  static constexpr size_t smem_c = (size_t)(RB + 1) * kLine * sizeof(float2) + (size_t)N * sizeof(float2);
};

// grid: (N / RB, npairs), block: N threads (one per column).
template <int RX, int RB>
__global__ void __launch_bounds__(RX* RX)
    fused_explicit_rowfft_kernel(const float* __restrict__ uv, float* __restrict__ uvs, float2* __restrict__ spec,
                                 int E, float dt, float h, float nu) {
  using G = FusedGeom<RX, RB>;
  constexpr int N = G::N, mask = N - 1, SR = G::kSlabRows, LS = G::kLine, NW = G::kWarps;
  extern __shared__ __align__(16) uint8_t fsm[];
  float* su = reinterpret_cast<float*>(fsm);           // [SR][N]
  float* sv = su + SR * N;                             // [SR][N]
  float2* zl = reinterpret_cast<float2*>(sv + SR * N);  // [RB][LS]
  float2* tw = zl + RB * LS;                           // [N]
  float* sforce = reinterpret_cast<float*>(tw + N);    // [N]
  float* fyb = sforce + N;                             // [2][NW][RB + 1]
  float* vsb = fyb + 2 * NW * (RB + 1);                // [NW][RB]

  const int tid = threadIdx.x, j = tid, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.x * RB, pair = blockIdx.y;
  const float dt_h = dt / h, inv_h = 1.f / h, inv_h2 = 1.f / (h * h);

  fill_twiddles_full(tw, N, tid, N);
  sforce[j] = sinf(4.f * ((float)j + 0.5f) * h);  // Kolmogorov forcing at u's offset y_{j+1/2}

#define U(k, dj) su[((k) + 3) * N + ((j + (dj)) & mask)]
#define V(k, dj) sv[((k) + 3) * N + ((j + (dj)) & mask)]
  for (int m = 0; m < 2; ++m) {
    const int e = 2 * pair + m;
    if (e >= E) {  // odd ensemble: the missing partner contributes a zero field
      __syncthreads();
      for (int idx = tid; idx < RB * N; idx += N) zl[(idx / N) * LS + (idx % N)].y = 0.f;
      break;
    }
    const float* u = uv + (size_t)e * 2 * N * N;
    const float* v = u + (size_t)N * N;
    __syncthreads();  // previous member's march has finished reading the slab
    for (int idx = tid; idx < SR * (N / 4); idx += N) {
      const int li = idx / (N / 4), c4 = idx % (N / 4);
      const int gi = (i0 - 3 + li) & mask;
      reinterpret_cast<float4*>(su)[idx] = reinterpret_cast<const float4*>(u + (size_t)gi * N)[c4];
      reinterpret_cast<float4*>(sv)[idx] = reinterpret_cast<const float4*>(v + (size_t)gi * N)[c4];
    }
    __syncthreads();
    // y-face fluxes through the right face of the column left of every warp (lane 0's left face)
    if (tid < NW * (RB + 1)) {
      const int w = tid / (RB + 1), k = tid % (RB + 1) - 1;
      const int jb = (32 * w - 1) & mask;
#define UB(kk, dj) su[((kk) + 3) * N + ((jb + (dj)) & mask)]
#define VB(kk, dj) sv[((kk) + 3) * N + ((jb + (dj)) & mask)]
      const float vfu = 0.5f * (VB(k, 0) + VB(k + 1, 0));
      fyb[w * (RB + 1) + k + 1] = face_value(UB(k, -1), UB(k, 0), UB(k, 1), UB(k, 2), vfu, dt_h) * vfu;
      const float vfv = 0.5f * (VB(k, 0) + VB(k, 1));
      fyb[(NW + w) * (RB + 1) + k + 1] = face_value(VB(k, -1), VB(k, 0), VB(k, 1), VB(k, 2), vfv, dt_h) * vfv;
#undef UB
#undef VB
    }
    __syncthreads();
    // march down the column: the flux through the lower x face of a cell is the upper-face flux
    // of the previous row (register), the left y-face flux comes from the neighbouring lane
    float* us = uvs + (size_t)e * 2 * N * N;
    float* vs = us + (size_t)N * N;
    float fxu_prev = 0.f, fxv_prev = 0.f, ustar_prev = 0.f;
    // sliding column windows (rows k-1 .. k+2) live in registers: one new row per iteration
    const int jm1 = (j - 1) & mask, jp1 = (j + 1) & mask, jp2 = (j + 2) & mask;
    const float force = sforce[j];
    const float* pu = su + N;  // row k of the slab is pu + (k + 2) * N ...
    const float* pv = sv + N;  //   ... i.e. these point at row k = -2
    float um1 = pu[-N + j], u00 = pu[j], up1 = pu[N + j];
    float vm1 = pv[-N + j], v00 = pv[j], vp1 = pv[N + j];
#pragma unroll 2
    for (int k = -2; k < RB; ++k) {
      const float up2 = pu[2 * N + j], vp2 = pv[2 * N + j];
      const float ur1 = pu[jp1];
      const float ufu = 0.5f * (u00 + up1);
      const float fxu = face_value(um1, u00, up1, up2, ufu, dt_h) * ufu;
      const float ufv = 0.5f * (u00 + ur1);
      const float fxv = face_value(vm1, v00, vp1, vp2, ufv, dt_h) * ufv;
      if (k >= -1) {
        const float ul1 = pu[jm1], ur2 = pu[jp2];
        const float vl1 = pv[jm1], vr1 = pv[jp1], vr2 = pv[jp2];
        const float vfu = 0.5f * (v00 + vp1);
        const float fyu = face_value(ul1, u00, ur1, ur2, vfu, dt_h) * vfu;
        const float vfv = 0.5f * (v00 + vr1);
        const float fyv = face_value(vl1, v00, vr1, vr2, vfv, dt_h) * vfv;
        float fyu_m = __shfl_up_sync(0xffffffffu, fyu, 1), fyv_m = __shfl_up_sync(0xffffffffu, fyv, 1);
        if (lane == 0) fyu_m = fyb[warp * (RB + 1) + k + 1], fyv_m = fyb[(NW + warp) * (RB + 1) + k + 1];
        const float conv_u = -((fxu - fxu_prev) + (fyu - fyu_m)) * inv_h;
        const float lap_u = (up1 + um1 + ur1 + ul1 - 4.f * u00) * inv_h2;
        const float ustar = u00 + dt * (conv_u + nu * lap_u + (force - 0.1f * u00));
        const float conv_v = -((fxv - fxv_prev) + (fyv - fyv_m)) * inv_h;
        const float lap_v = (vp1 + vm1 + vr1 + vl1 - 4.f * v00) * inv_h2;
        const float vstar = v00 + dt * (conv_v + nu * lap_v - 0.1f * v00);
        const float vleft = __shfl_up_sync(0xffffffffu, vstar, 1);
        if (k >= 0) {
          const int gi = i0 + k;
          us[(size_t)gi * N + j] = ustar;
          vs[(size_t)gi * N + j] = vstar;
          // backward-difference divergence; lane 0 lacks v*(j-1), fixed up below from vsb
          const float zp = ((ustar - ustar_prev) + (lane ? vstar - vleft : vstar)) * inv_h;
          float* zc = reinterpret_cast<float*>(zl + k * LS + j) + m;
          *zc = zp;
          if (lane == 31) vsb[warp * RB + k] = vstar;
        }
        ustar_prev = ustar;
      }
      fxu_prev = fxu, fxv_prev = fxv;
      um1 = u00, u00 = up1, up1 = up2;
      vm1 = v00, v00 = vp1, vp1 = vp2;
      pu += N, pv += N;
    }
    __syncthreads();
    if (tid < NW * RB) {
      const int w = tid / RB, k = tid % RB;
      float* zc = reinterpret_cast<float*>(zl + k * LS + 32 * w) + m;
      *zc -= vsb[((w + NW - 1) % NW) * RB + k] * inv_h;
    }
  }
#undef U
#undef V
  __syncthreads();
  // FFT along y of the RB lines, RX threads per line
  for (int line = tid / RX; line < RB; line += N / RX) {
    const int t = tid % RX;
    float2* x = zl + line * LS;
    float2 a[RX];
#pragma unroll
    for (int n1 = 0; n1 < RX; ++n1) a[n1] = x[RX * n1 + t];
    const unsigned grp = line_group_mask(RX);
    __syncwarp(grp);
    line_fft_forward<RX>(a, x, tw, t, grp);
    float2* dst = spec + ((size_t)pair * N + (i0 + line)) * N;
#pragma unroll
    for (int k2 = 0; k2 < RX; ++k2) dst[t + RX * k2] = a[k2];
  }
}

// ---------------------------------------------------------------------------- packed pair march (Blackwell FFMA2)
// The two members of a pair run the SAME arithmetic on different data, and sm_100 has two-wide fp32 instructions
// (add / mul / fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2 on an aligned register pair): the march below carries
// (member a, member b) in one float2 per quantity, so every add / multiply / fma is ONE issue slot for both
// members; only compares, selects and the reciprocal stay per component.  The slab holds the members interleaved
// (one 8-byte shared-memory load per stencil point and pair).  The explicit kernel is bound by instruction issue
// ("Stepper"): this is where its time goes.
using f2 = float2;
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ f2 splat(float a) { return make_float2(a, a); }
__device__ __forceinline__ float rcp_approx(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ f2 shfl_up2(f2 v) {
  return make_float2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
}
// face_value for both members (same formula, see face_value)
__device__ __forceinline__ f2 face_value2(f2 cl, f2 c, f2 cr, f2 cn, f2 uf, f2 ndt_h) {
  const f2 d = sub2(cr, c);
  const bool px = uf.x > 0.f, py = uf.y > 0.f;
  const f2 al = sub2(c, cl), ar = sub2(cn, cr);
  const f2 a = make_float2(px ? al.x : ar.x, py ? al.y : ar.y);
  // (no safe denominator here: with d = 0 the test a d > 0 already selects phi = 0, and the limited correction
  // (high - upwind) phi vanishes with d in either formulation)
  const f2 ad = mul2(a, d), sum = add2(d, a), two = add2(a, a);
  const f2 q = mul2(two, make_float2(rcp_approx(sum.x), rcp_approx(sum.y)));
  const f2 phi = make_float2(ad.x > 0.f ? q.x : 0.f, ad.y > 0.f ? q.y : 0.f);
  const f2 upwind = make_float2(px ? c.x : cr.x, py ? c.y : cr.y);
  // high = c + (1 - Cn) d / 2
  const f2 high = fma2(mul2(fma2(uf, ndt_h, splat(1.f)), splat(0.5f)), d, c);
  return fma2(sub2(high, upwind), phi, upwind);
}

// Same work as fused_explicit_rowfft_kernel, both members of the pair marched together.  grid: (N / RB, npairs),
// block: N threads (one per column).
template <int RX, int RB>
__global__ void __launch_bounds__(RX* RX)
    fused_explicit_rowfft2_kernel(const float* __restrict__ uv, float* __restrict__ uvs, float2* __restrict__ spec,
                                  int E, float dt, float h, float nu) {
  using G = FusedGeom<RX, RB>;
  constexpr int N = G::N, mask = N - 1, SR = G::kSlabRows, LS = G::kLine, NW = G::kWarps;
  extern __shared__ __align__(16) uint8_t fsm[];
  f2* su = reinterpret_cast<f2*>(fsm);       // [SR][N] (member a, member b)
  f2* sv = su + SR * N;                      // [SR][N]
  float2* zl = sv + SR * N;                  // [RB][LS]
  float2* tw = zl + RB * LS;                 // [N]
  float* sforce = reinterpret_cast<float*>(tw + N);  // [N]
  f2* fyb = reinterpret_cast<f2*>(sforce + N);       // [2][NW][RB + 1]
  f2* vsb = fyb + 2 * NW * (RB + 1);                 // [NW][RB]

  const int tid = threadIdx.x, j = tid, lane = tid & 31, warp = tid >> 5;
  const int i0 = blockIdx.x * RB, pair = blockIdx.y;
  const float inv_h = 1.f / h;
  const f2 ndt_h = splat(-dt / h), ninv_h = splat(-inv_h), inv_h2 = splat(1.f / (h * h));

  // programmatic dependent launch: the tables below are filled while the previous kernel of the inner step drains
  pdl_launch_dependents();
  fill_twiddles_full(tw, N, tid, N);
  sforce[j] = sinf(4.f * ((float)j + 0.5f) * h);  // Kolmogorov forcing at u's offset y_{j+1/2}
  pdl_wait();  // no global access above

  const int ea = 2 * pair, eb = 2 * pair + 1;
  const bool hasb = eb < E;  // odd ensemble: the missing partner is a zero field whose results are dropped
  const float* ua = uv + (size_t)ea * 2 * N * N;
  const float* va = ua + (size_t)N * N;
  const float* ub = uv + (size_t)(hasb ? eb : ea) * 2 * N * N;
  const float* vb = ub + (size_t)N * N;
  for (int idx = tid; idx < SR * (N / 4); idx += N) {
    const int li = idx / (N / 4), c4 = idx % (N / 4);
    const int gi = (i0 - 3 + li) & mask;
    const float4 a = reinterpret_cast<const float4*>(ua + (size_t)gi * N)[c4];
    const float4 c = reinterpret_cast<const float4*>(va + (size_t)gi * N)[c4];
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f), d = b;
    if (hasb) {
      b = reinterpret_cast<const float4*>(ub + (size_t)gi * N)[c4];
      d = reinterpret_cast<const float4*>(vb + (size_t)gi * N)[c4];
    }
    float4* du = reinterpret_cast<float4*>(su + li * N + 4 * c4);
    float4* dv = reinterpret_cast<float4*>(sv + li * N + 4 * c4);
    du[0] = make_float4(a.x, b.x, a.y, b.y), du[1] = make_float4(a.z, b.z, a.w, b.w);
    dv[0] = make_float4(c.x, d.x, c.y, d.y), dv[1] = make_float4(c.z, d.z, c.w, d.w);
  }
  __syncthreads();
  // y-face fluxes through the right face of the column left of every warp (lane 0's left face)
  if (tid < NW * (RB + 1)) {
    const int w = tid / (RB + 1), k = tid % (RB + 1) - 1;
    const int jb = (32 * w - 1) & mask;
#define UB(kk, dj) su[((kk) + 3) * N + ((jb + (dj)) & mask)]
#define VB(kk, dj) sv[((kk) + 3) * N + ((jb + (dj)) & mask)]
    const f2 vfu = mul2(add2(VB(k, 0), VB(k + 1, 0)), splat(0.5f));
    fyb[w * (RB + 1) + k + 1] = mul2(face_value2(UB(k, -1), UB(k, 0), UB(k, 1), UB(k, 2), vfu, ndt_h), vfu);
    const f2 vfv = mul2(add2(VB(k, 0), VB(k, 1)), splat(0.5f));
    fyb[(NW + w) * (RB + 1) + k + 1] = mul2(face_value2(VB(k, -1), VB(k, 0), VB(k, 1), VB(k, 2), vfv, ndt_h), vfv);
#undef UB
#undef VB
  }
  __syncthreads();
  // march down the column: the flux through the lower x face of a cell is the upper-face flux of the previous
  // row (register), the left y-face flux comes from the neighbouring lane
  float* usa = uvs + (size_t)ea * 2 * N * N;
  float* vsa = usa + (size_t)N * N;
  float* usb = uvs + (size_t)eb * 2 * N * N;
  float* vsb_g = usb + (size_t)N * N;
  f2 fxu_prev = splat(0.f), fxv_prev = splat(0.f), ustar_prev = splat(0.f);
  const int jm1 = (j - 1) & mask, jp1 = (j + 1) & mask, jp2 = (j + 2) & mask;
  const f2 force = splat(sforce[j]);
  const f2 half = splat(0.5f), dt2 = splat(dt), nu2 = splat(nu), drag = splat(-0.1f), m4 = splat(-4.f);
  const f2* pu = su + N;  // row k of the slab is pu + (k + 2) * N, i.e. these point at row k = -2
  const f2* pv = sv + N;
  f2 um1 = pu[-N + j], u00 = pu[j], up1 = pu[N + j];
  f2 vm1 = pv[-N + j], v00 = pv[j], vp1 = pv[N + j];
#pragma unroll 2
  for (int k = -2; k < RB; ++k) {
    const f2 up2 = pu[2 * N + j], vp2 = pv[2 * N + j];
    const f2 ur1 = pu[jp1];
    const f2 ufu = mul2(add2(u00, up1), half);
    const f2 fxu = mul2(face_value2(um1, u00, up1, up2, ufu, ndt_h), ufu);
    const f2 ufv = mul2(add2(u00, ur1), half);
    const f2 fxv = mul2(face_value2(vm1, v00, vp1, vp2, ufv, ndt_h), ufv);
    if (k >= -1) {
      const f2 ul1 = pu[jm1], ur2 = pu[jp2];
      const f2 vl1 = pv[jm1], vr1 = pv[jp1], vr2 = pv[jp2];
      const f2 vfu = mul2(add2(v00, vp1), half);
      const f2 fyu = mul2(face_value2(ul1, u00, ur1, ur2, vfu, ndt_h), vfu);
      const f2 vfv = mul2(add2(v00, vr1), half);
      const f2 fyv = mul2(face_value2(vl1, v00, vr1, vr2, vfv, ndt_h), vfv);
      f2 fyu_m = shfl_up2(fyu), fyv_m = shfl_up2(fyv);
      if (lane == 0) fyu_m = fyb[warp * (RB + 1) + k + 1], fyv_m = fyb[(NW + warp) * (RB + 1) + k + 1];
      // conv = -((fx - fx_prev) + (fy - fy_m)) / h
      const f2 conv_u = mul2(add2(sub2(fxu, fxu_prev), sub2(fyu, fyu_m)), ninv_h);
      const f2 lap_u = mul2(fma2(u00, m4, add2(add2(up1, um1), add2(ur1, ul1))), inv_h2);
      const f2 ustar = fma2(dt2, add2(fma2(nu2, lap_u, conv_u), fma2(drag, u00, force)), u00);
      const f2 conv_v = mul2(add2(sub2(fxv, fxv_prev), sub2(fyv, fyv_m)), ninv_h);
      const f2 lap_v = mul2(fma2(v00, m4, add2(add2(vp1, vm1), add2(vr1, vl1))), inv_h2);
      const f2 vstar = fma2(dt2, fma2(drag, v00, fma2(nu2, lap_v, conv_v)), v00);
      const f2 vleft = shfl_up2(vstar);
      if (k >= 0) {
        const size_t o = (size_t)(i0 + k) * N + j;
        usa[o] = ustar.x, vsa[o] = vstar.x;
        if (hasb) usb[o] = ustar.y, vsb_g[o] = vstar.y;
        // backward-difference divergence; lane 0 lacks v*(j-1), fixed up below from vsb
        f2 zp = mul2(add2(sub2(ustar, ustar_prev), lane ? sub2(vstar, vleft) : vstar), splat(inv_h));
        if (!hasb) zp.y = 0.f;
        zl[k * LS + j] = zp;
        if (lane == 31) vsb[warp * RB + k] = vstar;
      }
      ustar_prev = ustar;
    }
    fxu_prev = fxu, fxv_prev = fxv;
    um1 = u00, u00 = up1, up1 = up2;
    vm1 = v00, v00 = vp1, vp1 = vp2;
    pu += N, pv += N;
  }
  __syncthreads();
  if (tid < NW * RB) {
    const int w = tid / RB, k = tid % RB;
    const f2 vl = vsb[((w + NW - 1) % NW) * RB + k];
    float2& z = zl[k * LS + 32 * w];
    z.x -= vl.x * inv_h;
    if (hasb) z.y -= vl.y * inv_h;
  }
  __syncthreads();
  // FFT along y of the RB lines, RX threads per line
  for (int line = tid / RX; line < RB; line += N / RX) {
    const int t = tid % RX;
    float2* x = zl + line * LS;
    float2 a[RX];
#pragma unroll
    for (int n1 = 0; n1 < RX; ++n1) a[n1] = x[RX * n1 + t];
    const unsigned grp = line_group_mask(RX);
    __syncwarp(grp);
    line_fft_forward<RX>(a, x, tw, t, grp);
    float2* dst = spec + ((size_t)pair * N + (i0 + line)) * N;
#pragma unroll
    for (int k2 = 0; k2 < RX; ++k2) dst[t + RX * k2] = a[k2];
  }
}

// FFT along x, multiply by 1 / (lambda_x + lambda_y) / N^2 (zero mode -> 0), inverse FFT along x.
// grid: (N / 16, npairs), block: 16 columns x RX threads; thread (r, c) = tid / 16, tid % 16.
template <int RX>
__global__ void __launch_bounds__(16 * RX) fused_col_solve_kernel(float2* __restrict__ spec, float h) {
  constexpr int N = RX * RX, LC = RX * (RX + 1) + 1;
  extern __shared__ __align__(16) uint8_t fsm[];
  float2* xch = reinterpret_cast<float2*>(fsm);  // [16][LC]
  float2* tw = xch + 16 * LC;                    // [N]
  float* lam = reinterpret_cast<float*>(tw + N);  // [N]
  const int tid = threadIdx.x, c = tid & 15, r = tid >> 4;
  pdl_launch_dependents();
  fill_twiddles_full(tw, N, tid, 16 * RX);
  for (int k = tid; k < N; k += 16 * RX) {
    const float s = sinpif((float)k / (float)N);
    lam[k] = -4.f * s * s / (h * h);
  }
  pdl_wait();  // no global access above
  const int py = blockIdx.x * 16 + c;  // y frequency of this column (natural order)
  float2* base = spec + (size_t)blockIdx.y * N * N + py;
  float2* x = xch + c * LC;
  float2 a[RX];
#pragma unroll
  for (int n1 = 0; n1 < RX; ++n1) a[n1] = base[(size_t)(RX * n1 + r) * N];
  __syncthreads();
  fft_reg<RX, -1>(a);
#pragma unroll
  for (int k1 = 0; k1 < RX; ++k1) x[r * (RX + 1) + k1] = cmul(a[k1], tw[r * k1]);
  __syncthreads();
#pragma unroll
  for (int n2 = 0; n2 < RX; ++n2) a[n2] = x[n2 * (RX + 1) + r];
  fft_reg<RX, -1>(a);  // a[k2] = X[kx = r + RX k2]
  const float scale = 1.f / ((float)N * (float)N);
  const float ly = lam[py];
#pragma unroll
  for (int k2 = 0; k2 < RX; ++k2) {
    const int kx = r + RX * k2;
    const float f = (kx | py) ? __fdividef(scale, lam[kx] + ly) : 0.f;
    a[k2].x *= f, a[k2].y *= f;
  }
  fft_reg<RX, 1>(a);  // over k2 -> n2
#pragma unroll
  for (int n2 = 0; n2 < RX; ++n2) {
    float2 w = tw[r * n2];
    w.y = -w.y;
    x[n2 * (RX + 1) + r] = cmul(a[n2], w);  // the slots this thread read above
  }
  __syncthreads();
#pragma unroll
  for (int k1 = 0; k1 < RX; ++k1) a[k1] = x[r * (RX + 1) + k1];
  fft_reg<RX, 1>(a);
#pragma unroll
  for (int n1 = 0; n1 < RX; ++n1) base[(size_t)(RX * n1 + r) * N] = a[n1];
}

// inverse FFT along y of rows i0 .. i0 + RB (one extra row for the x difference), then
// v = v* - grad q (forward differences).  grid: (N / RB, npairs), block: kThreadsC.
template <int RX, int RB>
__global__ void __launch_bounds__(FusedGeom<RX, RB>::kThreadsC)
    fused_rowifft_grad_kernel(const float2* __restrict__ spec, const float* __restrict__ uvs, float* __restrict__ uv,
                              int E, float h) {
  using G = FusedGeom<RX, RB>;
  constexpr int N = G::N, mask = N - 1, LS = G::kLine;
  extern __shared__ __align__(16) uint8_t fsm[];
  float2* ql = reinterpret_cast<float2*>(fsm);  // [RB + 1][LS]
  float2* tw = ql + (RB + 1) * LS;              // [N]
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * RB, pair = blockIdx.y;
  pdl_launch_dependents();
  fill_twiddles_full(tw, N, tid, blockDim.x);
  pdl_wait();  // no global access above
  __syncthreads();
  for (int line = tid / RX; line <= RB; line += blockDim.x / RX) {
    const int t = tid % RX;
    const float2* src = spec + ((size_t)pair * N + ((i0 + line) & mask)) * N;
    float2* x = ql + line * LS;
    float2 a[RX];
#pragma unroll
    for (int k2 = 0; k2 < RX; ++k2) a[k2] = src[t + RX * k2];
    const unsigned grp = line_group_mask(RX);
    line_fft_inverse<RX>(a, x, tw, t, grp);
    __syncwarp(grp);
#pragma unroll
    for (int n1 = 0; n1 < RX; ++n1) x[RX * n1 + t] = a[n1];
  }
  __syncthreads();
  // thread = 4 consecutive columns of one row group
  constexpr int kColGroups = N / 4, kRowGroups = (N >= 256 ? 256 : N) / kColGroups;
  if (tid >= kColGroups * kRowGroups) return;
  const int cg = tid % kColGroups, rg = tid / kColGroups;
  const float inv_h = 1.f / h;
  for (int k = rg; k < RB; k += kRowGroups) {
    const int gi = i0 + k;
    float2 q0[5], q1[4];
#pragma unroll
    for (int d = 0; d < 5; ++d) q0[d] = ql[k * LS + ((4 * cg + d) & mask)];
#pragma unroll
    for (int d = 0; d < 4; ++d) q1[d] = ql[(k + 1) * LS + 4 * cg + d];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const int e = 2 * pair + m;
      if (e >= E) break;
      const size_t o = (size_t)e * 2 * N * N + (size_t)gi * N + 4 * cg;
      float4 us = *reinterpret_cast<const float4*>(uvs + o);
      float4 vs = *reinterpret_cast<const float4*>(uvs + o + (size_t)N * N);
#define Q(arr, d) (m ? arr[d].y : arr[d].x)
      us.x -= (Q(q1, 0) - Q(q0, 0)) * inv_h, us.y -= (Q(q1, 1) - Q(q0, 1)) * inv_h;
      us.z -= (Q(q1, 2) - Q(q0, 2)) * inv_h, us.w -= (Q(q1, 3) - Q(q0, 3)) * inv_h;
      vs.x -= (Q(q0, 1) - Q(q0, 0)) * inv_h, vs.y -= (Q(q0, 2) - Q(q0, 1)) * inv_h;
      vs.z -= (Q(q0, 3) - Q(q0, 2)) * inv_h, vs.w -= (Q(q0, 4) - Q(q0, 3)) * inv_h;
#undef Q
      *reinterpret_cast<float4*>(uv + o) = us;
      *reinterpret_cast<float4*>(uv + o + (size_t)N * N) = vs;
    }
  }
}

}  // namespace

}  // namespace sdab

using namespace sdab;

struct sdab_kolmogorov {
  int N, log2n, steps;
  float dt_inner, h, nu;
};

namespace {

struct KWs {
  float* uvs;
  float2* spec;
  size_t total;
};

KWs kws(const sdab_kolmogorov* k, int E, void* ws) {
  const size_t n2 = (size_t)k->N * k->N;
  const size_t npairs = (E + 1) / 2;
  KWs w;
  const size_t uvs_bytes = round_up_sz((size_t)E * 2 * n2 * sizeof(float), 1024);
  // the prior filters one member per complex field: size the spectrum for E fields
  const size_t spec_bytes = round_up_sz((size_t)(E > (int)npairs ? E : npairs) * n2 * sizeof(float2), 1024);
  w.uvs = (float*)ws;
  w.spec = (float2*)((uint8_t*)ws + uvs_bytes);
  w.total = uvs_bytes + spec_bytes;
  return w;
}

// projection of `src` (u*, v*) into `dst`
int project(const sdab_kolmogorov* k, const float* src, float* dst, float2* spec, int E, cudaStream_t st) {
  const int N = k->N, npairs = (E + 1) / 2;
  const size_t row_smem = (size_t)(N / 2 + kRows * N) * sizeof(float2);
  div_row_fft_kernel<<<dim3(N / kRows, npairs), dim3(N / 2, kRows), row_smem, st>>>(src, spec, N, E, k->h, 0);
  SDAB_LAUNCH_CHECK("div_row_fft_kernel");
  const int CT = N <= 256 ? 8 : 2;
  const size_t col_smem = (size_t)(N / 2 + CT * (N + 1)) * sizeof(float2);
  col_solve_kernel<<<dim3(N / CT, npairs), dim3(CT, N / 2), col_smem, st>>>(spec, N, k->log2n, CT, k->h, 0);
  SDAB_LAUNCH_CHECK("col_solve_kernel");
  row_ifft_kernel<<<dim3(N / kRows, npairs), dim3(N / 2, kRows), row_smem, st>>>(spec, N);
  SDAB_LAUNCH_CHECK("row_ifft_kernel");
  grad_sub_kernel<<<148 * 8, 256, 0, st>>>(src, spec, dst, N, E, k->h);
  SDAB_LAUNCH_CHECK("grad_sub_kernel");
  return SDAB_OK;
}

// ---------------------------------------------------------------------------- fast path (host)
// launch with the programmatic-stream-serialization attribute (pdl_wait() in the kernel; SDAB_PDL=0: plain launch)
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

bool fast_path(int N) {
  static const int on = env_int("SDAB_KOLMO_FAST", 1);
  return on && (N == 64 || N == 256);
}

template <int RX, int RB>
int inner_step_fast(const sdab_kolmogorov* k, float* uv, float* uvs, float2* spec, int E, cudaStream_t st) {
  using G = FusedGeom<RX, RB>;
  constexpr int N = G::N;
  const int npairs = (E + 1) / 2;
  static bool attr_set = false;
  if (!attr_set) {
    SDAB_CUDA_CHECK(cudaFuncSetAttribute(fused_explicit_rowfft_kernel<RX, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)G::smem_a));
    SDAB_CUDA_CHECK(cudaFuncSetAttribute(fused_rowifft_grad_kernel<RX, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)G::smem_c));
    SDAB_CUDA_CHECK(cudaFuncSetAttribute(fused_explicit_rowfft2_kernel<RX, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)G::smem_a2));
    attr_set = true;
  }
  // SDAB_KOLMO_PACKED=0: the scalar march (one member after the other), kept as the cross-check of the packed one
  static const int packed = env_int("SDAB_KOLMO_PACKED", 1);
  if (packed) {
    SDAB_CUDA_CHECK(launch_pdl(fused_explicit_rowfft2_kernel<RX, RB>, dim3(N / RB, npairs), dim3(N), G::smem_a2, st,
                               (const float*)uv, uvs, spec, E, k->dt_inner, k->h, k->nu));
    SDAB_LAUNCH_CHECK("fused_explicit_rowfft2_kernel");
  } else {
    fused_explicit_rowfft_kernel<RX, RB><<<dim3(N / RB, npairs), N, G::smem_a, st>>>(uv, uvs, spec, E, k->dt_inner, k->h, k->nu);
    SDAB_LAUNCH_CHECK("fused_explicit_rowfft_kernel");
  }
  constexpr size_t smem_b = (size_t)16 * (RX * (RX + 1) + 1) * sizeof(float2) + (size_t)N * sizeof(float2) + (size_t)N * sizeof(float);
  SDAB_CUDA_CHECK(launch_pdl(fused_col_solve_kernel<RX>, dim3(N / 16, npairs), dim3(16 * RX), smem_b, st, spec, k->h));
  SDAB_LAUNCH_CHECK("fused_col_solve_kernel");
  SDAB_CUDA_CHECK(launch_pdl(fused_rowifft_grad_kernel<RX, RB>, dim3(N / RB, npairs), dim3(G::kThreadsC), G::smem_c, st,
                             (const float2*)spec, (const float*)uvs, uv, E, k->h));
  SDAB_LAUNCH_CHECK("fused_rowifft_grad_kernel");
  return SDAB_OK;
}

// The ensemble can be walked in chunks of (an even number of) members whose working set -- state,
// u*/v* and the pair spectra, 20 N^2 bytes per member -- stays resident in the 126 MB L2
// (SDAB_KOLMO_L2_MB); members are independent, so the order does not change results.  Measured on
// B200: the kernels are latency- rather than bandwidth-bound and the smaller grids of a chunked walk
// cost more than the L2 hits save, so the default is one chunk.
int transition_fast(const sdab_kolmogorov* k, float* uv, int E, int n_transitions, float* traj, const KWs& w,
                    cudaStream_t st) {
  const int N = k->N;
  const size_t n2 = (size_t)N * N, state = (size_t)E * 2 * n2;
  static const int chunk_mb = env_int("SDAB_KOLMO_L2_MB", 1 << 20);
  static const int rb256 = env_int("SDAB_KOLMO_RB", 8);
  int chunk = (int)((size_t)chunk_mb * 1024 * 1024 / (20 * n2));
  chunk = chunk < 2 ? 2 : chunk & ~1;
  for (int e0 = 0; e0 < E; e0 += chunk) {
    const int ne = E - e0 < chunk ? E - e0 : chunk;
    float* cuv = uv + (size_t)e0 * 2 * n2;
    float* cuvs = w.uvs + (size_t)e0 * 2 * n2;
    float2* cspec = w.spec + (size_t)(e0 / 2) * n2;
    for (int t = 0; t < n_transitions; ++t) {
      for (int s = 0; s < k->steps; ++s) {
        if (N == 64)
          SDAB_TRY((inner_step_fast<8, 8>(k, cuv, cuvs, cspec, ne, st)));
        else if (rb256 == 8)
          SDAB_TRY((inner_step_fast<16, 8>(k, cuv, cuvs, cspec, ne, st)));
        else
          SDAB_TRY((inner_step_fast<16, 16>(k, cuv, cuvs, cspec, ne, st)));
      }
      if (traj) SDAB_TRY(copy_f32(cuv, traj + (size_t)t * state + (size_t)e0 * 2 * n2, (size_t)ne * 2 * n2, st));
    }
  }
  return SDAB_OK;
}

}  // namespace

extern "C" {

int sdab_kolmogorov_create(int size, double dt, double reynolds, sdab_kolmogorov** out) {
  SDAB_REQUIRE(out, "null argument");
  SDAB_REQUIRE(size >= 32 && size <= 512 && (size & (size - 1)) == 0, "grid size must be a power of two in [32, 512]");
  SDAB_REQUIRE(dt > 0 && reynolds > 0, "dt and reynolds must be positive");
  auto* k = new sdab_kolmogorov();
  k->N = size;
  k->log2n = 0;
  while ((1 << k->log2n) < size) ++k->log2n;
  const double h = 2.0 * 3.14159265358979323846 / size;
  // cfd.equations.stable_time_step(max_velocity=5, max_courant_number=0.5): sda/mcs.py:274-284
  const double dt_min = 0.5 * h / 5.0;
  k->steps = dt_min > dt ? 1 : (int)ceil(dt / dt_min);
  k->dt_inner = (float)(dt / k->steps);
  k->h = (float)h;
  k->nu = (float)(1.0 / reynolds);
  *out = k;
  return SDAB_OK;
}

void sdab_kolmogorov_destroy(sdab_kolmogorov* k) { delete k; }

int sdab_kolmogorov_inner_steps(const sdab_kolmogorov* k) { return k ? k->steps : 0; }

size_t sdab_kolmogorov_workspace_bytes(const sdab_kolmogorov* k, int E) {
  if (!k || E < 1) return 0;
  return kws(k, E, nullptr).total;
}

int sdab_kolmogorov_transition(sdab_kolmogorov* k, float* uv, int E, int n_transitions, float* traj, void* workspace,
                               size_t workspace_bytes, void* stream) {
  SDAB_REQUIRE(k && uv && workspace, "null argument");
  SDAB_REQUIRE(E >= 1 && E <= 65535 && n_transitions >= 0, "ensemble size out of range");
  SDAB_TRY(sdab_device_check());
  const KWs w = kws(k, E, workspace);
  SDAB_REQUIRE(workspace_bytes >= w.total, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = k->N;
  const size_t state = (size_t)E * 2 * N * N;
  if (fast_path(N)) return transition_fast(k, uv, E, n_transitions, traj, w, st);
  for (int t = 0; t < n_transitions; ++t) {
    for (int s = 0; s < k->steps; ++s) {
      explicit_step_kernel<<<dim3(N / kTile, N / kTile, E), dim3(32, 8), 0, st>>>(uv, w.uvs, N, k->dt_inner, k->h,
                                                                                  k->nu);
      SDAB_LAUNCH_CHECK("explicit_step_kernel");
      SDAB_TRY(project(k, w.uvs, uv, w.spec, E, st));
    }
    if (traj) SDAB_TRY(copy_f32(uv, traj + (size_t)t * state, state, st));
  }
  return SDAB_OK;
}

int sdab_kolmogorov_prior(sdab_kolmogorov* k, float* uv, int E, uint64_t seed, void* workspace, size_t workspace_bytes,
                          void* stream) {
  SDAB_REQUIRE(k && uv && workspace, "null argument");
  SDAB_REQUIRE(E >= 1 && E <= 65535, "ensemble size out of range");
  SDAB_TRY(sdab_device_check());
  const KWs w = kws(k, E, workspace);
  SDAB_REQUIRE(workspace_bytes >= w.total, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = k->N;
  // white noise for both components, filtered as ONE complex field per member (real symbol)
  SDAB_TRY(sdab_randn(w.uvs, (size_t)E * 2 * N * N, seed, 0, stream));
  const size_t row_smem = (size_t)(N / 2 + kRows * N) * sizeof(float2);
  div_row_fft_kernel<<<dim3(N / kRows, E), dim3(N / 2, kRows), row_smem, st>>>(w.uvs, w.spec, N, E, k->h, 1);
  SDAB_LAUNCH_CHECK("div_row_fft_kernel");
  const int CT = N <= 256 ? 8 : 2;
  const size_t col_smem = (size_t)(N / 2 + CT * (N + 1)) * sizeof(float2);
  col_solve_kernel<<<dim3(N / CT, E), dim3(CT, N / 2), col_smem, st>>>(w.spec, N, k->log2n, CT, k->h, 1);
  SDAB_LAUNCH_CHECK("col_solve_kernel");
  row_ifft_kernel<<<dim3(N / kRows, E), dim3(N / 2, kRows), row_smem, st>>>(w.spec, N);
  SDAB_LAUNCH_CHECK("row_ifft_kernel");
  unpack_pair_kernel<<<148 * 8, 256, 0, st>>>(w.spec, uv, N, E);
  SDAB_LAUNCH_CHECK("unpack_pair_kernel");
  for (int it = 0; it < 3; ++it) {  // filtered_velocity_field: 3 x (project, rescale to max speed 3)
    SDAB_TRY(copy_f32(uv, w.uvs, (size_t)E * 2 * N * N, st));
    SDAB_TRY(project(k, w.uvs, uv, w.spec, E, st));
    normalize_speed_kernel<<<E, 1024, 0, st>>>(uv, N, 3.f);
    SDAB_LAUNCH_CHECK("normalize_speed_kernel");
  }
  return SDAB_OK;
}

}  // extern "C"
