/*
 * sdab.h -- C ABI of libsdab, the sm_100a CUDA library behind sda_b200.
 *
 * This is the drop-in boundary for the SDA hot path.  The
 * reference (francois-rozet/sda) has no FFI layer: its operator API is the Python
 * classes in sda/nn.py, sda/score.py and sda/mcs.py.  Each entry point below
 * therefore cites the reference method whose arithmetic it replaces; the Python
 * classes in sda_b200/ keep the reference signatures and call these through
 * ctypes (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - tensors are dense, row-major, float32, in the reference's own layouts
 *     (NCHW for images, (B, L, C, H, W) for trajectories);
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as
 *     void*); no hidden synchronisation, no internal allocation of tensors:
 *     the caller (PyTorch) owns every buffer, including workspaces;
 *   - functions return 0 on success and a non-zero code otherwise;
 *     sdab_last_error() returns a message for the calling thread;
 *   - there is NO CPU fallback: on a machine without an sm_100 device the
 *     compute entry points fail with SDAB_ERR_DEVICE.
 */
#ifndef SDAB_H_
#define SDAB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDAB_OK 0
#define SDAB_ERR_ARG 1      /* unsupported shape / argument                  */
#define SDAB_ERR_DEVICE 2   /* no sm_100 device, or CUDA runtime failure     */
#define SDAB_ERR_STATE 3    /* call order (weights not set, no saved state)  */

#define SDAB_MAX_DEPTH 8

/* arithmetic mode of the tensor-core convolutions */
#define SDAB_MODE_BF16X3 0  /* 2-term bf16 split, 3 MMAs / product: fp32-grade parity (default) */
#define SDAB_MODE_BF16 1    /* single bf16 pass: fast mode, ~5e-3 rel. error                   */

/* convolution engine (debug / validation) */
#define SDAB_ENGINE_UMMA 0  /* tcgen05 implicit GEMM (product path)        */
#define SDAB_ENGINE_SIMT 1  /* fp32 CUDA-core implicit GEMM (cross-check)  */

#define SDAB_ACT_SILU 0
#define SDAB_ACT_RELU 1

const char* sdab_last_error(void);
int sdab_version(void);
/* 0 when an sm_100 device is current and usable */
int sdab_device_check(void);

/* ------------------------------------------------------------------------- *
 * U-Net  (reference: sda/nn.py:94-206 UNet, :18-28 ModResidualBlock)
 * ------------------------------------------------------------------------- */

typedef struct sdab_unet sdab_unet;

typedef struct sdab_unet_desc {
  int in_channels;                       /* UNet(in_channels, ...)            nn.py:96  */
  int out_channels;                      /*                                   nn.py:97  */
  int mod_features;                      /*                                   nn.py:98  */
  int depth;                             /* len(hidden_channels)                         */
  int hidden_channels[SDAB_MAX_DEPTH];   /*                                   nn.py:99  */
  int hidden_blocks[SDAB_MAX_DEPTH];     /*                                   nn.py:100 */
  int activation;                        /* SDAB_ACT_*                        nn.py:103 */
} sdab_unet_desc;

/* kernel_size 3, stride 2, spatial 2, padding_mode 'circular' only (the
 * Kolmogorov configuration, experiments/kolmogorov/utils.py:59-68). */
int sdab_unet_create(const sdab_unet_desc* desc, sdab_unet** out);
void sdab_unet_destroy(sdab_unet* h);

/* Number of convolutions / modulated blocks, in forward execution order:
 * convs : head0, descent[0][*].{conv1,conv2}, head1, ..., ascent (deepest first)
 *         blocks, each followed by its tail conv.
 * blocks: descent[0][*], descent[1][*], ..., ascent[0][*] (deepest), ...        */
int sdab_unet_num_convs(const sdab_unet* h);
int sdab_unet_num_blocks(const sdab_unet* h);
/* (C_out, C_in) of convolution i */
int sdab_unet_conv_shape(const sdab_unet* h, int i, int* c_out, int* c_in);

/* Bytes of the packed-weight buffer the caller must provide. */
size_t sdab_unet_packed_bytes(const sdab_unet* h);

/* Packs all parameters into `packed` (bf16 hi/lo, K-major per tap, forward and
 * transposed/flipped for the input-gradient).  conv_w[i]: (C_out, C_in, 3, 3),
 * conv_b[i]: (C_out); proj_w[j]: (C_j, mod), proj_b[j]: (C_j): nn.py:132-135.
 * The arrays of pointers live on the HOST; the pointed tensors on the device. */
int sdab_unet_set_weights(sdab_unet* h, const float* const* conv_w_host, const float* const* conv_b_host,
                          const float* const* proj_w_host, const float* const* proj_b_host, void* packed,
                          size_t packed_bytes, void* stream);

/* Workspace needed for N images of H x W.  save = 1 keeps what dgrad needs, save = 2 what
 * sdab_unet_backward (parameter gradients) needs as well. */
size_t sdab_unet_workspace_bytes(const sdab_unet* h, int N, int H, int W, int save);

/* UNet.forward(x, y)  nn.py:184-206.
 * x: (N, in_channels, H, W); y: (Nt, mod_features), Nt in {1, N}; out: (N, out_channels, H, W). */
int sdab_unet_forward(sdab_unet* h, const float* x, const float* y, int Nt, int N, int H, int W, float* out,
                      void* workspace, size_t workspace_bytes, int save, int mode, int engine, void* stream);

/* MCScoreNet.forward (score.py:134-144) for the windows [w_begin, w_end) of the flattened (B, L - 2k) window
 * index, with unfold (score.py:146-153), ScoreUNet.forward's context concat (score.py:87) and fold
 * (score.py:155-164) as ADDRESSING of the network's first and last layer instead of tensors:
 *   x   : the trajectory (B, L, C, H, W); window i of trajectory b is frames i .. i + 2k;
 *   ctx : (Cc, H, W) context planes appended to every window (NULL when Cc == 0);
 *   y   : (1, mod_features) modulation vector (one diffusion time per call);
 *   out : cap == 0 -- the score (B, L, C, H, W); the call writes the frames its windows feed (the centre
 *         slot of every window, the k leading / trailing slots of the first / last window of a trajectory);
 *         cap  > 0 -- this rank's shard of `cap` frames (C, H, W): the centre frame of local window n at
 *         position n, the 2k edge frames of the j-th trajectory the range touches at per + 2k j + e.
 *         Equal-size shards are what one all-gather moves; sdab_frames_assemble builds the score from them.
 * The network must have in_channels = (2k+1) C + Cc and out_channels = (2k+1) C.  Workspace as for
 * sdab_unet_forward with N = w_end - w_begin. */
int sdab_mcscore_forward(sdab_unet* h, const float* x, const float* ctx, const float* y, int B, int L, int C, int Cc,
                         int H, int W, int order, int w_begin, int w_end, float* out, int per, int cap,
                         void* workspace, size_t workspace_bytes, int save, int mode, int engine, void* stream);
/* Input-VJP of the last sdab_mcscore_forward(save != 0): gs is the cotangent of the folded score
 * (B, L, C, H, W); gwin receives the window input-gradients (w_end - w_begin, (2k+1) C, H, W) (the cotangent
 * of the unfolded windows, context channels dropped) -- sdab_unfold_transpose_add overlap-adds them. */
int sdab_mcscore_dgrad(sdab_unet* h, const float* gs, float* gwin, int B, int L, int C, int Cc, int H, int W, int order,
                       int w_begin, int w_end, void* workspace, size_t workspace_bytes, int mode, int engine,
                       void* stream);

/* Vector-Jacobian product w.r.t. x of the last forward run with save != 0 on the
 * same workspace: gx = J_x^T gout (what torch.autograd.grad does through the
 * reference UNet in GaussianScore.forward, score.py:381-394). */
int sdab_unet_dgrad(sdab_unet* h, const float* gout, float* gx, void* workspace, size_t workspace_bytes, int mode,
                    int engine, void* stream);

/* Training backward (reference: what loss.backward() computes through UNet in VPSDE.loss,
 * sda/score.py:265-276, driven by sda/utils.py:136-143): the input-gradient of sdab_unet_dgrad
 * plus the gradients of every convolution weight and bias and of the time-shift table, for
 * the last forward run with save = 2 on the same workspace.
 * conv_dw_host[i]: (C_out, C_in, 3, 3), conv_db_host[i]: (C_out) in the order of
 * sdab_unet_set_weights (host arrays of device pointers; overwritten);
 * dshift: (Nt, sdab_unet_shift_rows) = d loss / d (proj_w[j] y + proj_b[j]) for the blocks in
 * order, from which the caller derives the gradients of the projection Linears
 * (nn.py:132-135) and of y.  Weight gradients run on the fp32 CUDA cores (first path);
 * partial sums are combined with floating-point atomics. */
int sdab_unet_backward(sdab_unet* h, const float* gout, float* gx, float* const* conv_dw_host,
                       float* const* conv_db_host, float* dshift, void* workspace, size_t workspace_bytes, int mode,
                       int engine, void* stream);
/* Rows of the shift table: sum of the channels of all modulated blocks. */
int sdab_unet_shift_rows(const sdab_unet* h);

/* One 3x3 circular convolution on NCHW fp32 tensors (nn.Conv2d(kernel_size=3, padding=1,
 * padding_mode='circular', stride), sda/nn.py:125-128,151-157).  weight: (Cout, Cin, 3, 3); bias:
 * (Cout) or NULL; out: (N, Cout, H/stride, W/stride).  transpose != 0 (stride 1) computes the
 * input-gradient instead: x: (N, Cout, H, W) -> out: (N, Cin, H, W), bias must be NULL. */
size_t sdab_conv3x3_workspace_bytes(int N, int Cin, int Cout, int H, int W, int stride, int transpose);
int sdab_conv3x3(const float* x, const float* weight, const float* bias, float* out, int N, int Cin, int Cout, int H,
                 int W, int stride, int transpose, int mode, int engine, void* workspace, size_t workspace_bytes,
                 void* stream);

/* Per-launch device timing of the convolution engine (CUDA events on the launching stream):
 * enable, run, then read the summed kernel time (ms), algorithmic FLOPs and launch count. */
int sdab_conv_profile(int enable);
int sdab_conv_profile_read(double* ms, double* flops, long long* launches);

/* Number of kernels launched by this library (all threads) since the last reset. */
long long sdab_launch_count(int reset);

/* ------------------------------------------------------------------------- *
 * Window maps  (reference: sda/score.py:146-164 MCScoreNet.unfold / fold)
 * ------------------------------------------------------------------------- */

/* unfold + channel concat of a broadcast context (ScoreUNet.forward's torch.cat, score.py:87):
 * x: (B, L, C, H, W) -> win: (B, L-2k, (2k+1) C + Cc, H, W); ctx: (Cc, H, W) or NULL. */
int sdab_unfold_cat(const float* x, const float* ctx, float* win, int B, int L, int C, int Cc, int H, int W, int order,
                    void* stream);
/* fold: win_out: (B, L-2k, (2k+1) C, H, W) -> s: (B, L, C, H, W)  score.py:155-164 */
int sdab_fold(const float* win_out, float* s, int B, int L, int C, int H, int W, int order, void* stream);
/* adjoint of fold: scatter of the cotangent into zero-initialised windows */
int sdab_fold_transpose(const float* gs, float* gwin, int B, int L, int C, int H, int W, int order, void* stream);
/* Builds the score (B, L, C, H, W) from the all-gathered shards of sdab_mcscore_forward(cap > 0):
 * gathered: (world * cap, C, H, W), rank r's shard at r * cap; rank r owns windows [r per, (r+1) per). */
int sdab_frames_assemble(const float* gathered, float* s, int B, int L, int C, int H, int W, int order, int per, int cap,
                         void* stream);
/* adjoint of unfold_cat w.r.t. x: deterministic overlap-add (autograd UnfoldBackward0) */
int sdab_unfold_transpose_add(const float* gwin, float* gx, int B, int L, int C, int Cc, int H, int W, int order,
                              void* stream);

/* ------------------------------------------------------------------------- *
 * Exchange step of the window-sharded score (the reference is single-GPU, its
 * MCScoreNet.forward composes the trajectory score from independent window scores, score.py:134-144):
 * all-gather over NVLink peer memory, one kernel per rank (csrc/peer.cu).
 * ------------------------------------------------------------------------- */

/* Bytes in front of the payload of a peer buffer (flags, block counter). */
size_t sdab_peer_header_bytes(void);
/* cudaMalloc of header + payload_bytes on the current device, header zeroed; handle64 receives the 64-byte CUDA IPC
 * handle other processes of the box open with sdab_peer_open. */
int sdab_peer_alloc(size_t payload_bytes, void** ptr, void* handle64);
int sdab_peer_open(const void* handle64, void** ptr);
int sdab_peer_close(void* ptr);
int sdab_peer_free(void* ptr);
/* bufs[p]: base (header) of rank p's buffer as mapped in THIS process, p < world (host array).  Pushes
 * [shard_offset, shard_offset + shard_bytes) of the own payload to every peer, signals epoch, returns (on the stream)
 * when every peer's shard of the same epoch has arrived in the own buffer.  Epochs of one buffer must increase. */
int sdab_peer_allgather(void* const* bufs, int rank, int world, size_t shard_offset, size_t shard_bytes,
                        unsigned long long epoch, void* stream);
/* Announces `epoch` in every peer's header and waits (on the stream) for every peer's announcement: "my buffer is
 * complete" -- e.g. the gradients of loss.backward() before sdab_peer_adamw reads them over NVLink. */
int sdab_peer_signal_wait(void* const* bufs, int rank, int world, unsigned long long epoch, void* stream);
/* Data-parallel optimizer.step() of torch.optim.AdamW (the optimizer of sda/utils.py:125-143) for the slice
 * [begin, end) of the flat parameter vector owned by `rank`: gradient = mean over the ranks of grad_bufs[p] (read
 * over NVLink, added in rank order), AdamW on the slice (m, v: the slice's moments, end - begin floats), the new
 * parameters stored into EVERY rank's param_bufs[p]; returns (on the stream) when the other ranks' slices have
 * arrived in the own parameter buffer.  world == 1: a plain fused AdamW over [begin, end).  step counts from 1. */
int sdab_peer_adamw(void* const* grad_bufs, void* const* param_bufs, float* m, float* v, size_t begin, size_t end,
                    int rank, int world, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    unsigned long long epoch, void* stream);

/* ------------------------------------------------------------------------- *
 * Sampler updates  (reference: sda/score.py:250-261 VPSDE.sample loop body)
 * ------------------------------------------------------------------------- */

/* predictor: x <- a * x + b * eps */
int sdab_vpsde_predict(float* x, const float* eps, float a, float b, size_t n, void* stream);
/* corrector: per batch element i (event = n / B values):
 *   delta_i = tau / mean(eps_i^2);  x <- x - (delta_i * eps + sqrt(2 delta_i) * z) * sigma
 * z ~ N(0, 1) from a counter-based Philox keyed by (seed, global element index), or read from
 * `z` when it is not NULL (noise injection for parity tests).
 * scratch: sdab_vpsde_correct_scratch_floats(B) floats. */
size_t sdab_vpsde_correct_scratch_floats(int B);
int sdab_vpsde_correct(float* x, const float* eps, const float* z, float tau, float sigma, uint64_t seed,
                       uint64_t offset, int B, size_t n, float* scratch, void* stream);
/* standard normal draw with the same Philox stream (x(1) initialisation) */
int sdab_randn(float* out, size_t n, uint64_t seed, uint64_t offset, void* stream);

/* Tweedie estimate and final combination of GaussianScore.forward (score.py:387,396):
 *   xhat = (x - sigma * eps) / mu ;   out = eps - sigma * s                                  */
int sdab_tweedie(const float* x, const float* eps, float mu, float sigma, float* xhat, size_t n, void* stream);
/* the same with mu and sigma as device scalars: the caller does not synchronise on the schedule */
int sdab_tweedie_dev(const float* x, const float* eps, const float* mu, const float* sigma, float* xhat, size_t n,
                     void* stream);
int sdab_axpy(const float* a, const float* b, float alpha, float* out, size_t n, void* stream); /* out = a + alpha b */

/* ------------------------------------------------------------------------- *
 * Kolmogorov flow  (reference: sda/mcs.py:244-338 KolmogorovFlow, i.e. jax-cfd's
 * semi_implicit_navier_stokes finite-volume step + FFT pressure projection)
 * ------------------------------------------------------------------------- */

typedef struct sdab_kolmogorov sdab_kolmogorov;

int sdab_kolmogorov_create(int size, double dt, double reynolds, sdab_kolmogorov** out);
void sdab_kolmogorov_destroy(sdab_kolmogorov* h);
int sdab_kolmogorov_inner_steps(const sdab_kolmogorov* h);
size_t sdab_kolmogorov_workspace_bytes(const sdab_kolmogorov* h, int E);
/* uv: (E, 2, size, size), advanced in place by n_transitions transitions (mcs.py:333-338);
 * when traj != NULL every transition is also written to traj: (n_transitions, E, 2, size, size). */
int sdab_kolmogorov_transition(sdab_kolmogorov* h, float* uv, int E, int n_transitions, float* traj, void* workspace,
                               size_t workspace_bytes, void* stream);
/* prior (mcs.py:321-331): filtered random velocity field, max speed 3, peak wavenumber 4 */
int sdab_kolmogorov_prior(sdab_kolmogorov* h, float* uv, int E, uint64_t seed, void* workspace,
                          size_t workspace_bytes, void* stream);
/* observation helpers (mcs.py:340-347, :361-375) */
int sdab_coarsen(const float* x, float* out, size_t n_img, int H, int W, int r, void* stream);
int sdab_vorticity(const float* x, float* out, size_t n_pair, int H, int W, void* stream);
/* Adjoints (vector-Jacobian products) of the two operators above, and KolmogorovFlow.upsample(mode='bilinear')
 * (mcs.py:349-359: circular pad 1 -> F.interpolate(scale_factor=r) -> crop r) with its adjoint: the observation
 * operators A(x) of the guided sampler are differentiated at every score evaluation (score.py:389-394).
 *   coarsen_adjoint : g (n_img, H/r, W/r) -> gx (n_img, H, W)
 *   vorticity_adjoint: g (n_pair, H, W)   -> gx (n_pair, 2, H, W)
 *   upsample        : x (n_img, H, W)     -> out (n_img, r H, r W);  adjoint: g (n_img, r H, r W) -> gx (n_img, H, W) */
int sdab_coarsen_adjoint(const float* g, float* gx, size_t n_img, int H, int W, int r, void* stream);
int sdab_vorticity_adjoint(const float* g, float* gx, size_t n_pair, int H, int W, void* stream);
int sdab_upsample_bilinear(const float* x, float* out, size_t n_img, int H, int W, int r, void* stream);
int sdab_upsample_bilinear_adjoint(const float* g, float* gx, size_t n_img, int H, int W, int r, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* SDAB_H_ */
