// peer.cu -- all-gather of the window-sharded score over NVLink peer memory (the one exchange
// step of the path -- per-window scores, resp. window input-gradients, back into the full-trajectory buffer).
//
// Every rank owns one buffer of the gathered size, allocated here with cudaMalloc and exported through a CUDA IPC
// handle; the other ranks of the box map it (cudaIpcOpenMemHandle, lazy peer access).  A rank's U-Net writes its
// shard into its OWN buffer; sdab_peer_allgather is then ONE kernel per rank that
//   1. pushes the shard to the same offset of every peer's buffer (16-byte stores over NVLink / NVSwitch),
//   2. fences (system scope) and, in the last block to finish, raises this rank's flag in every peer's header,
//   3. waits in that block until every peer's flag for this epoch has arrived here.
// When the kernel retires, the rank's buffer holds all shards: what follows in the stream (sdab_frames_assemble,
// sdab_unfold_transpose_add) needs no further synchronisation.  No host round trip, no proxy thread, no staging copy.
//
// Buffer layout: [header 4096 B: flags u64[world] at 0, block counter u32 at 2048][payload].  Flags carry a
// monotonically increasing epoch (never reset).  Reuse safety: the caller alternates TWO buffers per exchange; a
// rank can only be pushed to for epoch e + 2 after the pusher completed epoch e + 1, which needs this rank's push
// of e + 1, which this rank's stream orders after its consumer of epoch e.
#include <cuda.h>

#include <cmath>
#include <cstring>

#include "common.cuh"

namespace sdab {

namespace {

constexpr size_t kHeader = 4096;
constexpr int kMaxWorld = 16;

struct PeerPtrs {
  uint8_t* buf[kMaxWorld];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// announce `epoch` in every peer's header and wait for every peer's announcement (threads 0 .. world-1 of a block)
__device__ __forceinline__ void signal_and_wait(const PeerPtrs& pp, int rank, int world, unsigned long long epoch) {
  if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
    const int p = threadIdx.x;
    st_release_sys(reinterpret_cast<unsigned long long*>(pp.buf[p]) + rank, epoch);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(pp.buf[rank]) + p;
    unsigned long long spins = 0;
    while (ld_acquire_sys(mine) < epoch) {
      __nanosleep(200);
      if (++spins > (1ull << 27)) __trap();
    }
  }
}

__global__ void __launch_bounds__(256)
    peer_allgather_kernel(const PeerPtrs pp, int rank, int world, size_t shard_offset, size_t shard_vec4,
                          unsigned long long epoch) {
  const uint4* src = reinterpret_cast<const uint4*>(pp.buf[rank] + kHeader + shard_offset);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < shard_vec4; i += stride) {
    const uint4 v = src[i];
    for (int p = 0; p < world; ++p)
      if (p != rank) reinterpret_cast<uint4*>(pp.buf[p] + kHeader + shard_offset)[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(pp.buf[rank] + 2048);
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
    if (last) *counter = 0u;  // next launch on this buffer (stream-ordered) starts from zero
  }
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  // one thread per peer: announce, then wait for the peer's announcement (bounded: a dead peer aborts the launch
  // after about half a minute instead of hanging the GPU)
  signal_and_wait(pp, rank, world, epoch);
}

__global__ void peer_signal_wait_kernel(const PeerPtrs pp, int rank, int world, unsigned long long epoch) {
  __threadfence_system();
  signal_and_wait(pp, rank, world, epoch);
}

struct AdamW {
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, grad_scale;
};

// Data-parallel optimizer step as ONE kernel per rank (sda/utils.py:136-143: loss.backward(); optimizer.step() with
// torch.optim.AdamW, experiments/kolmogorov/train.py): the rank owns the slice [begin, end) of the flat parameter
// vector.  For every element of its slice it reads the gradient from ALL ranks' gradient buffers over NVLink (peer
// loads), adds them in rank order (deterministic, the same on every run), applies AdamW to its slice of the
// parameters and moments, and stores the new parameter into EVERY rank's parameter buffer (peer stores).  The last
// block to finish raises the rank's flag in the peers' parameter headers and waits for theirs: when the kernel
// retires, this rank's parameters are complete.  Gradient all-reduce, optimizer and parameter broadcast without a
// separate collective, a bucket copy or a second pass over the parameters.
__global__ void __launch_bounds__(256)
    peer_adamw_kernel(const PeerPtrs grads, const PeerPtrs params, float* __restrict__ m, float* __restrict__ v,
                      size_t begin, size_t n4, int rank, int world, AdamW o, unsigned long long epoch) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const size_t e = begin / 4 + i;  // float4 index in the flat vectors
    float4 g = reinterpret_cast<const float4*>(grads.buf[0] + kHeader)[e];
    for (int p = 1; p < world; ++p) {
      const float4 t = reinterpret_cast<const float4*>(grads.buf[p] + kHeader)[e];
      g.x += t.x, g.y += t.y, g.z += t.z, g.w += t.w;
    }
    float4 w = reinterpret_cast<const float4*>(params.buf[rank] + kHeader)[e];
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* gp = &g.x;
    float* wp = &w.x;
    float* mp = &mm.x;
    float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gp[k] * o.grad_scale;
      wp[k] *= 1.f - o.lr * o.weight_decay;
      mp[k] = o.beta1 * mp[k] + (1.f - o.beta1) * gk;
      vp[k] = o.beta2 * vp[k] + (1.f - o.beta2) * gk * gk;
      const float denom = sqrtf(vp[k]) / o.bc2_sqrt + o.eps;
      wp[k] -= (o.lr / o.bc1) * (mp[k] / denom);
    }
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    for (int p = 0; p < world; ++p) reinterpret_cast<float4*>(params.buf[p] + kHeader)[e] = w;
  }
  if (world == 1) return;
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(params.buf[rank] + 2048);
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
    if (last) *counter = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  signal_and_wait(params, rank, world, epoch);
}

}  // namespace

}  // namespace sdab

using namespace sdab;

extern "C" {

size_t sdab_peer_header_bytes(void) { return kHeader; }

int sdab_peer_alloc(size_t payload_bytes, void** ptr, void* handle64) {
  SDAB_REQUIRE(ptr && handle64, "null argument");
  void* p = nullptr;
  SDAB_CUDA_CHECK(cudaMalloc(&p, kHeader + payload_bytes));
  SDAB_CUDA_CHECK(cudaMemset(p, 0, kHeader));
  SDAB_CUDA_CHECK(cudaDeviceSynchronize());
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(SDAB_ERR_DEVICE, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return SDAB_OK;
}

int sdab_peer_open(const void* handle64, void** ptr) {
  SDAB_REQUIRE(ptr && handle64, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(SDAB_ERR_DEVICE, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  return SDAB_OK;
}

int sdab_peer_close(void* ptr) {
  if (ptr) SDAB_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
  return SDAB_OK;
}

int sdab_peer_free(void* ptr) {
  if (ptr) SDAB_CUDA_CHECK(cudaFree(ptr));
  return SDAB_OK;
}

int sdab_peer_allgather(void* const* bufs, int rank, int world, size_t shard_offset, size_t shard_bytes,
                        unsigned long long epoch, void* stream) {
  SDAB_REQUIRE(bufs && world >= 2 && world <= kMaxWorld && rank >= 0 && rank < world, "invalid peer group");
  SDAB_REQUIRE(shard_offset % 16 == 0 && shard_bytes % 16 == 0, "shards must be 16-byte aligned");
  PeerPtrs pp{};
  for (int p = 0; p < world; ++p) {
    SDAB_REQUIRE(bufs[p], "null peer buffer");
    pp.buf[p] = (uint8_t*)bufs[p];
  }
  const size_t n = shard_bytes / 16;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 4) grid = 148 * 4;
  if (grid < 1) grid = 1;
  peer_allgather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pp, rank, world, shard_offset, n, epoch);
  SDAB_LAUNCH_CHECK("peer_allgather_kernel");
  return SDAB_OK;
}


static int to_ptrs(void* const* bufs, int world, PeerPtrs& pp) {
  SDAB_REQUIRE(bufs && world >= 1 && world <= kMaxWorld, "invalid peer group");
  for (int p = 0; p < world; ++p) {
    SDAB_REQUIRE(bufs[p], "null peer buffer");
    pp.buf[p] = (uint8_t*)bufs[p];
  }
  return SDAB_OK;
}

int sdab_peer_signal_wait(void* const* bufs, int rank, int world, unsigned long long epoch, void* stream) {
  PeerPtrs pp{};
  SDAB_TRY(to_ptrs(bufs, world, pp));
  SDAB_REQUIRE(rank >= 0 && rank < world, "invalid rank");
  if (world == 1) return SDAB_OK;
  peer_signal_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pp, rank, world, epoch);
  SDAB_LAUNCH_CHECK("peer_signal_wait_kernel");
  return SDAB_OK;
}

int sdab_peer_adamw(void* const* grad_bufs, void* const* param_bufs, float* m, float* v, size_t begin, size_t end,
                    int rank, int world, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    unsigned long long epoch, void* stream) {
  PeerPtrs g{}, w{};
  SDAB_TRY(to_ptrs(grad_bufs, world, g));
  SDAB_TRY(to_ptrs(param_bufs, world, w));
  SDAB_REQUIRE(rank >= 0 && rank < world && m && v && step >= 1, "invalid argument");
  SDAB_REQUIRE(begin % 4 == 0 && end % 4 == 0 && end >= begin, "slices must be multiples of 4 floats");
  AdamW o;
  o.lr = lr, o.beta1 = beta1, o.beta2 = beta2, o.eps = eps, o.weight_decay = weight_decay;
  o.bc1 = (float)(1.0 - pow((double)beta1, step));
  o.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, step));
  o.grad_scale = 1.f / (float)world;
  const size_t n4 = (end - begin) / 4;
  int grid = (int)((n4 + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  peer_adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, w, m, v, begin, n4, rank, world, o, epoch);
  SDAB_LAUNCH_CHECK("peer_adamw_kernel");
  return SDAB_OK;
}

}  // extern "C"
