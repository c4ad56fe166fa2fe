// conv_api.cu -- stand-alone 3x3 circular convolution entry point (nn.Conv2d(kernel_size=3,
// padding=1, padding_mode='circular', stride in {1, 2}), sda/nn.py:125-128,151-157) on NCHW fp32
// tensors: packs the input and the weights into the internal formats, runs one engine, unpacks.
// Used by the parity tests to pin each engine per layer shape; the U-Net keeps everything in the
// internal formats between layers instead.
#include "common.cuh"
#include "tile_geom.h"

using namespace sdab;

namespace {

struct ConvWs {
  size_t in_op, wf, wb, bias, outf, total;
};

ConvWs conv_ws(int N, int Cin, int Cout, int H, int W, int stride, int transpose) {
  // transpose: the "convolution" maps Cout -> Cin channels with the flipped kernel (stride 1 only)
  const int ci = transpose ? Cout : Cin, co = transpose ? Cin : Cout;
  const int K = round_up(ci, 32), Nn = round_up(co, 32);
  ConvWs w;
  size_t off = 0;
  auto take = [&](size_t b) {
    const size_t o = off;
    off += round_up_sz(b, 1024);
    return o;
  };
  w.in_op = take(OpShape{N, H, W, K, 0}.bytes());
  w.wf = take((size_t)9 * round_up(Cin, 32) * round_up(Cout, 32) * 2 * sizeof(bf16));
  w.wb = take((size_t)9 * round_up(Cout, 32) * round_up(Cin, 32) * 2 * sizeof(bf16));
  w.bias = take((size_t)Nn * sizeof(float));
  w.outf = take((size_t)N * (H / stride) * (W / stride) * Nn * sizeof(float));
  w.total = off;
  return w;
}

}  // namespace

extern "C" {

size_t sdab_conv3x3_workspace_bytes(int N, int Cin, int Cout, int H, int W, int stride, int transpose) {
  if (N < 1 || Cin < 1 || Cout < 1 || H < 1 || W < 1 || (stride != 1 && stride != 2)) return 0;
  return conv_ws(N, Cin, Cout, H, W, stride, transpose).total;
}

int sdab_conv3x3(const float* x, const float* weight, const float* bias, float* out, int N, int Cin, int Cout, int H,
                 int W, int stride, int transpose, int mode, int engine, void* workspace, size_t workspace_bytes,
                 void* stream) {
  SDAB_REQUIRE(x && weight && out && workspace, "null argument");
  SDAB_REQUIRE(stride == 1 || stride == 2, "stride must be 1 or 2");
  SDAB_REQUIRE(!(transpose && stride != 1), "the transposed form is stride 1 only");
  SDAB_REQUIRE(H % stride == 0 && W % stride == 0, "image size must be divisible by the stride");
  SDAB_REQUIRE(((uintptr_t)workspace & 1023) == 0, "workspace must be 1024-byte aligned");
  SDAB_TRY(sdab_device_check());
  const ConvWs w = conv_ws(N, Cin, Cout, H, W, stride, transpose);
  SDAB_REQUIRE(workspace_bytes >= w.total, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  const int ci = transpose ? Cout : Cin, co = transpose ? Cin : Cout;
  const int K = round_up(ci, 32), Nn = round_up(co, 32);
  SDAB_TRY(pack_nchw_to_op(x, (bf16*)(ws + w.in_op), N, ci, K, H, W, stride == 2, st));
  SDAB_TRY(pack_conv_weights(weight, (bf16*)(ws + w.wf), (bf16*)(ws + w.wb), Cout, Cin, st));
  SDAB_TRY(fill_zero(ws + w.bias, (size_t)Nn * sizeof(float), st));
  if (bias) SDAB_TRY(copy_f32(bias, (float*)(ws + w.bias), co, st));
  ConvProblem q{};
  q.in = (const bf16*)(ws + w.in_op);
  q.wpk = (const bf16*)(ws + (transpose ? w.wb : w.wf));
  q.N = N, q.H = H / stride, q.W = W / stride, q.Cin = K, q.Cout = Nn, q.stride = stride, q.mode = mode;
  q.epi.bias = (const float*)(ws + w.bias);
  q.epi.outF = (float*)(ws + w.outf);
  q.flops = 2.0 * 9.0 * (double)N * q.H * q.W * Cin * Cout;
  conv_profile_before(st);
  const int status = engine == SDAB_ENGINE_SIMT ? conv3x3_simt(q, st) : conv3x3_umma(q, st);
  conv_profile_after(st, q.flops);
  SDAB_TRY(status);
  return unpack_f_to_nchw((const float*)(ws + w.outf), out, N, co, Nn, q.H, q.W, st);
}

}  // extern "C"
