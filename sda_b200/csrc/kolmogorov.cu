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
#include "common.cuh"

namespace sdab {

namespace {

constexpr int kTile = 32;
constexpr int kHalo = 2;
constexpr int kTS = kTile + 2 * kHalo;

__device__ __forceinline__ float face_value(float cl, float c, float cr, float cn, float uf, float dt_h) {
  const float d = cr - c;
  const float den = d != 0.f ? d : 1.f;
  const float r = __fdividef(uf > 0.f ? c - cl : cn - cr, den);
  const float phi = r > 0.f ? __fdividef(2.f * r, 1.f + r) : 0.f;
  const float upwind = uf > 0.f ? c : cr;
  const float courant = dt_h * uf;
  const float high = uf > 0.f ? c + 0.5f * (1.f - courant) * d : cr - 0.5f * (1.f + courant) * d;
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
