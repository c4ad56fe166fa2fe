// unet.cu -- the U-Net schedule: UNet.forward (sda/nn.py:184-206) and its input-gradient as a
// sequence of conv / LayerNorm launches over internal NHWC tensors (formats in common.cuh).
//
// Forward of one modulated residual block (sda/nn.py:27-28, :131-142), per block j at level d:
//     A_j   = LN_C(x + shift_j)                       ln_forward          -> OP   (saved)
//     H     = act(conv1(A_j) + b1)                    conv epilogue       -> OP,  C1_j = pre-activation (saved)
//     x'    = x + conv2(H) + b2                       conv epilogue       -> F
// Backward (input-gradient only):
//     gH    = conv2^T(gx')  ;  gC1 = gH * act'(C1_j)  conv epilogue       -> OP
//     gA    = conv1^T(gC1)                            conv                -> F
//     gx    = gx' + LN^T(gA; A_j, rstd_j)             ln_backward         -> F + OP
// Heads (stride 2) read the parity layout; their transpose is a stride-1 conv over the
// zero-upsampled cotangent.  Tails are LN -> nearest x2 -> conv; their transpose is conv^T -> 2x2
// sum-pool -> LN^T.
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "tile_geom.h"

namespace sdab {

struct ConvLayer {
  int cin, cout;        // real channels
  int kf, nf;           // forward GEMM K (ceil32 cin), N (ceil32 cout)
  int kb, nb;           // backward GEMM K (ceil32 cout), N (ceil32 cin)
  size_t off_fwd, off_bwd, off_bias;  // byte offsets in the packed buffer
  size_t off_tf = 0, off_tb = 0;      // tails (d > 0): combined sub-pixel / transposed 4x4 weights (16 taps each)
  bool is_tail = false;
};

struct Arena {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t o = off;
    off += round_up_sz(bytes, 1024);
    return o;
  }
};

}  // namespace sdab

using namespace sdab;

struct sdab_unet {
  sdab_unet_desc d;
  std::vector<ConvLayer> convs;
  std::vector<int> block_ch, block_shift_off;
  int shift_rows = 0;
  size_t off_projw = 0, off_projb = 0, packed_bytes = 0;
  uint8_t* packed = nullptr;
  bool weights_set = false;
  // indices into convs / blocks
  std::vector<int> head_conv, tail_conv;                  // per level
  std::vector<std::vector<int>> desc_c1, asc_c1;          // conv1 index per (level, block); conv2 = +1
  std::vector<std::vector<int>> desc_blk, asc_blk;        // block ids
  // saved forward
  bool saved = false;
  int saved_level = 0;  // 1: input-gradient, 2: + parameter gradients
  int sN = 0, sNt = 0, sH = 0, sW = 0;
  void* sws = nullptr;
};

namespace {

struct Plan {
  int D;
  // forward temporaries
  size_t in_op, shift, finop, outf;
  std::vector<size_t> x0, x1, skip, hop, xs2, aop_tmp;
  // saved
  std::vector<size_t> aop, c1, rstd;  // per block
  std::vector<size_t> upop, rstd_tail;  // per level (d > 0)
  // backward temporaries
  size_t gout_op, gxf;
  std::vector<size_t> gc1op, gup, gz;
  std::vector<size_t> xs2g;  // training: the tail transposes' parity operand (xs2 itself feeds the heads' wgrad)
  std::vector<size_t> hsave;  // training: act(conv1) of every block as operand tensor (input of conv2's wgrad)
  size_t wpart = 0;           // training: pixel-split partial sums of the tcgen05 weight-gradient kernel
  size_t total;
};

// one wave of at most ~150 CTAs, each leaving 512 TMEM columns x 128 lanes of fp32 partial sums
constexpr size_t kWgradPartialBytes = (size_t)300 * 512 * 128 * sizeof(float);

size_t f_bytes(int N, int H, int W, int C) { return (size_t)N * H * W * C * sizeof(float); }
size_t op_bytes(int N, int H, int W, int C) { return OpShape{N, H, W, C, 0}.bytes(); }

Plan make_plan(const sdab_unet* h, int N, int Nt, int H, int W, bool save, bool train = false) {
  Plan p;
  const int D = h->d.depth;
  p.D = D;
  Arena a;
  const int kin0 = round_up(h->d.in_channels, 32), nout = round_up(h->d.out_channels, 32);
  p.in_op = a.take(op_bytes(N, H, W, kin0));
  p.shift = a.take((size_t)Nt * h->shift_rows * sizeof(float));
  p.finop = a.take(op_bytes(N, H, W, h->d.hidden_channels[0]));
  p.outf = a.take(f_bytes(N, H, W, nout));
  p.x0.resize(D), p.x1.resize(D), p.skip.resize(D), p.hop.resize(D), p.xs2.resize(D), p.aop_tmp.resize(D);
  p.upop.assign(D, 0), p.rstd_tail.assign(D, 0), p.gc1op.assign(D, 0), p.gup.assign(D, 0), p.gz.assign(D, 0);
  p.xs2g.assign(D, 0);
  for (int d = 0; d < D; ++d) {
    const int Hd = H >> d, Wd = W >> d, C = h->d.hidden_channels[d];
    p.x0[d] = a.take(f_bytes(N, Hd, Wd, C));
    p.x1[d] = a.take(f_bytes(N, Hd, Wd, C));
    p.skip[d] = a.take(f_bytes(N, Hd, Wd, C));
    p.hop[d] = a.take(op_bytes(N, Hd, Wd, C));
    p.xs2[d] = d < D - 1 ? a.take(op_bytes(N, Hd, Wd, C)) : 0;
    p.aop_tmp[d] = save ? 0 : a.take(op_bytes(N, Hd, Wd, C));
    if (d > 0) {
      p.upop[d] = a.take(op_bytes(N, 2 * Hd, 2 * Wd, C));
      p.rstd_tail[d] = a.take((size_t)N * Hd * Wd * sizeof(float));
    }
  }
  const int nblk = (int)h->block_ch.size();
  p.aop.assign(nblk, 0), p.c1.assign(nblk, 0), p.rstd.assign(nblk, 0), p.hsave.assign(nblk, 0);
  if (save) {
    auto per_level = [&](const std::vector<std::vector<int>>& blk) {
      for (int d = 0; d < D; ++d) {
        const int Hd = H >> d, Wd = W >> d, C = h->d.hidden_channels[d];
        for (int j : blk[d]) {
          p.aop[j] = a.take(op_bytes(N, Hd, Wd, C));
          if (train) p.hsave[j] = a.take(op_bytes(N, Hd, Wd, C));
          p.c1[j] = a.take(f_bytes(N, Hd, Wd, C));
          p.rstd[j] = a.take((size_t)N * Hd * Wd * sizeof(float));
        }
      }
    };
    per_level(h->desc_blk);
    per_level(h->asc_blk);
    if (train) p.wpart = a.take(kWgradPartialBytes);
    p.gout_op = a.take(op_bytes(N, H, W, round_up(h->d.out_channels, 32)));
    p.gxf = a.take(f_bytes(N, H, W, round_up(h->d.in_channels, 32)));
    for (int d = 0; d < D; ++d) {
      const int Hd = H >> d, Wd = W >> d, C = h->d.hidden_channels[d];
      p.gc1op[d] = a.take(op_bytes(N, Hd, Wd, C));
      p.xs2g[d] = train && d < D - 1 ? a.take(op_bytes(N, Hd, Wd, C)) : p.xs2[d];
      if (d > 0) {
        p.gup[d] = a.take(f_bytes(N, 2 * Hd, 2 * Wd, C));
        p.gz[d] = a.take(op_bytes(N, 2 * Hd, 2 * Wd, C));
      }
    }
  } else {
    p.gout_op = p.gxf = 0;
  }
  p.total = a.off;
  return p;
}

int check_shape(const sdab_unet* h, int N, int H, int W) {
  SDAB_REQUIRE(N >= 1, "batch must be positive");
  const int D = h->d.depth;
  SDAB_REQUIRE((H % (1 << (D - 1))) == 0 && (W % (1 << (D - 1))) == 0, "H and W must be divisible by 2^(depth-1)");
  for (int d = 0; d < D; ++d) {
    TileGeom g;
    SDAB_TRY(make_tile_geom(N, H >> d, W >> d, g));
    SDAB_REQUIRE((H >> d) >= 2 && (W >> d) >= 2, "deepest level must be at least 2 x 2");
  }
  return SDAB_OK;
}

// cr_in, cr_out: real (unpadded) channel counts when they differ from the padded GEMM sizes;
// scale: fraction of the dense taps that are algorithmically needed (1/4 for the zero-upsampled
// transpose of a stride-2 head) -- both only feed the algorithmic FLOP count of the profiler.
int run_conv(int engine, ConvProblem p, cudaStream_t st, int cr_in = -1, int cr_out = -1, double scale = 1.0) {
  p.flops = scale * 2.0 * 9.0 * (double)p.N * p.H * p.W * (cr_in < 0 ? p.Cin : cr_in) * (cr_out < 0 ? p.Cout : cr_out);
  conv_profile_before(st);
  const int s = engine == SDAB_ENGINE_SIMT ? conv3x3_simt(p, st) : conv3x3_umma(p, st);
  conv_profile_after(st, p.flops);
  return s;
}

}  // namespace

// ================================================================================== C ABI
extern "C" {

int sdab_unet_create(const sdab_unet_desc* desc, sdab_unet** out) {
  SDAB_REQUIRE(desc && out, "null argument");
  SDAB_REQUIRE(desc->depth >= 1 && desc->depth <= SDAB_MAX_DEPTH, "depth out of range");
  SDAB_REQUIRE(desc->in_channels >= 1 && desc->in_channels <= 512, "in_channels out of range");
  SDAB_REQUIRE(desc->out_channels >= 1 && desc->out_channels <= 512, "out_channels out of range");
  SDAB_REQUIRE(desc->mod_features >= 1, "mod_features must be positive");
  SDAB_REQUIRE(desc->activation == SDAB_ACT_SILU || desc->activation == SDAB_ACT_RELU,
               "only SiLU and ReLU activations are implemented");
  for (int d = 0; d < desc->depth; ++d) {
    SDAB_REQUIRE(desc->hidden_channels[d] % 32 == 0 && desc->hidden_channels[d] >= 32 &&
                     desc->hidden_channels[d] <= 512,
                 "hidden_channels must be multiples of 32 in [32, 512]");
    SDAB_REQUIRE(desc->hidden_blocks[d] >= 0, "hidden_blocks must be non-negative");
  }
  auto* h = new sdab_unet();
  h->d = *desc;
  const int D = desc->depth;
  auto add_conv = [&](int cin, int cout) {
    ConvLayer c{};
    c.cin = cin, c.cout = cout;
    c.kf = round_up(cin, 32), c.nf = round_up(cout, 32);
    c.kb = round_up(cout, 32), c.nb = round_up(cin, 32);
    h->convs.push_back(c);
    return (int)h->convs.size() - 1;
  };
  auto add_block = [&](int C) {
    h->block_ch.push_back(C);
    h->block_shift_off.push_back(h->shift_rows);
    h->shift_rows += C;
    return (int)h->block_ch.size() - 1;
  };
  h->head_conv.resize(D), h->tail_conv.resize(D);
  h->desc_c1.resize(D), h->asc_c1.resize(D), h->desc_blk.resize(D), h->asc_blk.resize(D);
  for (int d = 0; d < D; ++d) {
    const int C = desc->hidden_channels[d];
    h->head_conv[d] = add_conv(d == 0 ? desc->in_channels : desc->hidden_channels[d - 1], C);
    for (int b = 0; b < desc->hidden_blocks[d]; ++b) {
      h->desc_c1[d].push_back(add_conv(C, C));
      add_conv(C, C);
      h->desc_blk[d].push_back(add_block(C));
    }
  }
  for (int d = D - 1; d >= 0; --d) {
    const int C = desc->hidden_channels[d];
    for (int b = 0; b < desc->hidden_blocks[d]; ++b) {
      h->asc_c1[d].push_back(add_conv(C, C));
      add_conv(C, C);
      h->asc_blk[d].push_back(add_block(C));
    }
    h->tail_conv[d] = add_conv(C, d == 0 ? desc->out_channels : desc->hidden_channels[d - 1]);
    h->convs[h->tail_conv[d]].is_tail = d > 0;
  }
  Arena a;
  for (auto& c : h->convs) {
    c.off_fwd = a.take((size_t)9 * c.kf * c.nf * 2 * sizeof(bf16));
    c.off_bwd = a.take((size_t)9 * c.kb * c.nb * 2 * sizeof(bf16));
    c.off_bias = a.take((size_t)c.nf * sizeof(float));
    if (c.is_tail) {
      c.off_tf = a.take((size_t)16 * c.cin * c.cout * 2 * sizeof(bf16));
      c.off_tb = a.take((size_t)16 * c.cin * c.cout * 2 * sizeof(bf16));
    }
  }
  h->off_projw = a.take((size_t)h->shift_rows * desc->mod_features * sizeof(float));
  h->off_projb = a.take((size_t)h->shift_rows * sizeof(float));
  h->packed_bytes = a.off;
  *out = h;
  return SDAB_OK;
}

void sdab_unet_destroy(sdab_unet* h) { delete h; }

int sdab_unet_num_convs(const sdab_unet* h) { return h ? (int)h->convs.size() : 0; }
int sdab_unet_num_blocks(const sdab_unet* h) { return h ? (int)h->block_ch.size() : 0; }

int sdab_unet_conv_shape(const sdab_unet* h, int i, int* c_out, int* c_in) {
  SDAB_REQUIRE(h && i >= 0 && i < (int)h->convs.size(), "conv index out of range");
  if (c_out) *c_out = h->convs[i].cout;
  if (c_in) *c_in = h->convs[i].cin;
  return SDAB_OK;
}

size_t sdab_unet_packed_bytes(const sdab_unet* h) { return h ? h->packed_bytes : 0; }

int sdab_unet_set_weights(sdab_unet* h, const float* const* conv_w, const float* const* conv_b,
                          const float* const* proj_w, const float* const* proj_b, void* packed, size_t packed_bytes,
                          void* stream) {
  SDAB_REQUIRE(h && conv_w && conv_b && proj_w && proj_b && packed, "null argument");
  SDAB_REQUIRE(packed_bytes >= h->packed_bytes, "packed buffer too small");
  SDAB_TRY(sdab_device_check());
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* base = (uint8_t*)packed;
  // every convolution in one pack launch per kMaxPack layers, every small copy (zero-padded biases, projection
  // weights and biases) in one copy launch per kMaxCopy entries: a training step repacks after every optimizer step
  PackTable pt{};
  CopyTable ct{};
  auto flush_pack = [&]() -> int {
    if (pt.n) SDAB_TRY(pack_conv_weights_batched(pt, st));
    pt.n = 0;
    return SDAB_OK;
  };
  auto add_copy = [&](const float* src, float* dst, size_t n_src, size_t n_dst) -> int {
    if (ct.n == kMaxCopy) {
      SDAB_TRY(copy_pad_batched(ct, st));
      ct.n = 0;
    }
    ct.src[ct.n] = src, ct.dst[ct.n] = dst, ct.n_src[ct.n] = (int)n_src, ct.n_dst[ct.n] = (int)n_dst;
    ++ct.n;
    return SDAB_OK;
  };
  for (size_t i = 0; i < h->convs.size(); ++i) {
    const ConvLayer& c = h->convs[i];
    if (pt.n == kMaxPack) SDAB_TRY(flush_pack());
    pt.w[pt.n] = conv_w[i], pt.fwd[pt.n] = (bf16*)(base + c.off_fwd), pt.bwd[pt.n] = (bf16*)(base + c.off_bwd);
    pt.cout[pt.n] = c.cout, pt.cin[pt.n] = c.cin;
    ++pt.n;
    if (c.is_tail)
      SDAB_TRY(pack_tail_weights(conv_w[i], (bf16*)(base + c.off_tf), (bf16*)(base + c.off_tb), c.cout, c.cin, st));
    SDAB_TRY(add_copy(conv_b[i], (float*)(base + c.off_bias), c.cout, c.nf));
  }
  SDAB_TRY(flush_pack());
  const int mod = h->d.mod_features;
  for (size_t j = 0; j < h->block_ch.size(); ++j) {
    const size_t r0 = h->block_shift_off[j];
    SDAB_TRY(add_copy(proj_w[j], (float*)(base + h->off_projw) + r0 * mod, (size_t)h->block_ch[j] * mod,
                      (size_t)h->block_ch[j] * mod));
    SDAB_TRY(add_copy(proj_b[j], (float*)(base + h->off_projb) + r0, h->block_ch[j], h->block_ch[j]));
  }
  if (ct.n) SDAB_TRY(copy_pad_batched(ct, st));
  h->packed = base;
  h->weights_set = true;
  h->saved = false;
  return SDAB_OK;
}

size_t sdab_unet_workspace_bytes(const sdab_unet* h, int N, int H, int W, int save) {
  if (!h || N < 1) return 0;
  // Nt <= N: size the shift table for the per-sample case
  return make_plan(h, N, N, H, W, save != 0, save == 2).total;
}

}  // extern "C"

namespace {

// wio != null: x is the trajectory (B, L, C, H, W) and ctx the context planes; the network input is the window
// batch [w_begin, w_begin + N) by addressing, and `out` receives the folded frames (WindowIO, common.cuh)
int forward_impl(sdab_unet* h, const float* x, const float* y, int Nt, int N, int H, int W, float* out,
                 void* workspace, size_t workspace_bytes, int save, int mode, int engine, void* stream,
                 const WindowIO* wio, const float* ctx) {
  SDAB_REQUIRE(h && x && y && out && workspace, "null argument");
  if (!h->weights_set) return fail(SDAB_ERR_STATE, "sdab_unet_set_weights has not been called");
  SDAB_REQUIRE(Nt == 1 || Nt == N, "the modulation batch must be 1 or N");
  SDAB_REQUIRE(mode == SDAB_MODE_BF16X3 || mode == SDAB_MODE_BF16, "unknown mode");
  SDAB_REQUIRE(engine == SDAB_ENGINE_UMMA || engine == SDAB_ENGINE_SIMT, "unknown engine");
  SDAB_TRY(check_shape(h, N, H, W));
  SDAB_TRY(sdab_device_check());
  const Plan p = make_plan(h, N, N, H, W, save != 0, save == 2);
  SDAB_REQUIRE(workspace_bytes >= p.total, "workspace too small");
  SDAB_REQUIRE(((uintptr_t)workspace & 1023) == 0, "workspace must be 1024-byte aligned");

  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  const uint8_t* pk = h->packed;
  const int D = h->d.depth;
  const int act = h->d.activation == SDAB_ACT_SILU ? 1 : 2;
  auto F = [&](size_t off) { return (float*)(ws + off); };
  auto OP = [&](size_t off) { return (bf16*)(ws + off); };
  auto wf = [&](int ci) { return (const bf16*)(pk + h->convs[ci].off_fwd); };
  auto bias = [&](int ci) { return (const float*)(pk + h->convs[ci].off_bias); };
  h->saved = false;

  const int kin0 = round_up(h->d.in_channels, 32);
  if (wio)
    SDAB_TRY(pack_windows_to_op(x, ctx, OP(p.in_op), *wio, N, kin0, H, W, st));
  else
    SDAB_TRY(pack_nchw_to_op(x, OP(p.in_op), N, h->d.in_channels, kin0, H, W, 0, st));
  SDAB_TRY(time_shifts(y, (const float*)(pk + h->off_projw), (const float*)(pk + h->off_projb), F(p.shift), Nt,
                       h->shift_rows, h->d.mod_features, st));

  // With the tcgen05 engine the channel LayerNorm of the NEXT block is fused into the epilogue of the
  // convolution that produces the residual stream (ConvEpilogue::ln == 1): `prenorm` is the block
  // whose operand has already been produced that way.
  const bool fuse = engine == SDAB_ENGINE_UMMA;
  // The stride-2 heads read their input as parity images.  Outside training the tcgen05 engine takes them straight
  // from the normal operand layout with a TMA element stride of 2 (the epilogue of the level's last convolution
  // writes that operand), so no parity copy (f_to_operand, one extra pass over the level) is made; training keeps
  // the parity copy because the heads' weight-gradient kernel consumes it.
  static const int strided_env = getenv("SDAB_STRIDED_TMA") ? atoi(getenv("SDAB_STRIDED_TMA")) : 1;
  const bool strided = fuse && save != 2 && strided_env;
  int prenorm = -1;
  auto level_of = [&](int j) {  // level of block j
    for (int d = 0; d < D; ++d) {
      for (int v : h->desc_blk[d])
        if (v == j) return d;
      for (int v : h->asc_blk[d])
        if (v == j) return d;
    }
    return -1;
  };
  auto fuse_ln = [&](ConvProblem& q, int next_j) {
    if (!fuse || next_j < 0) return;
    const int d = level_of(next_j);
    q.epi.ln = 1;
    q.epi.ln_shift = F(p.shift) + h->block_shift_off[next_j];
    q.epi.ln_shift_stride = h->shift_rows, q.epi.ln_nt = Nt;
    q.epi.ln_rstd_out = save ? F(p.rstd[next_j]) : nullptr;
    q.epi.outOP = save ? OP(p.aop[next_j]) : OP(p.aop_tmp[d]);
    prenorm = next_j;
  };

  // one modulated residual block: cur -> dst (F); dst_op: also the raw operand of dst; next_j: the
  // block that consumes dst (its LayerNorm is fused into conv2), or -1
  // tail_ln: dst feeds the tail of level d > 0 (LN -> nearest x2 -> conv, sda/nn.py:161-170): its shift-free
  // LayerNorm is fused into conv2 as well and lands in upop[d]
  auto block = [&](int d, int j, int c1, const float* cur, float* dst, bf16* dst_op, int next_j,
                   bool tail_ln = false) -> int {
    const int Hd = H >> d, Wd = W >> d, C = h->d.hidden_channels[d];
    bf16* aop = save ? OP(p.aop[j]) : OP(p.aop_tmp[d]);
    if (prenorm != j)
      SDAB_TRY(ln_forward(cur, F(p.shift) + h->block_shift_off[j], h->shift_rows, Nt, aop,
                          save ? F(p.rstd[j]) : nullptr, N, Hd, Wd, C, 0, st));
    ConvProblem q{};
    q.in = aop, q.wpk = wf(c1), q.N = N, q.H = Hd, q.W = Wd, q.Cin = C, q.Cout = C, q.stride = 1, q.mode = mode;
    bf16* hop = save == 2 ? OP(p.hsave[j]) : OP(p.hop[d]);  // training keeps act(conv1) for conv2's weight gradient
    q.epi.bias = bias(c1), q.epi.pre = save ? F(p.c1[j]) : nullptr, q.epi.act = act, q.epi.outOP = hop;
    SDAB_TRY(run_conv(engine, q, st));
    ConvProblem r{};
    r.in = hop, r.wpk = wf(c1 + 1), r.N = N, r.H = Hd, r.W = Wd, r.Cin = C, r.Cout = C, r.stride = 1,
    r.mode = mode;
    r.epi.bias = bias(c1 + 1), r.epi.res = cur, r.epi.outF = dst, r.epi.outOP = dst_op;
    if (!dst_op) fuse_ln(r, next_j);
    if (tail_ln) {
      r.epi.ln = 1, r.epi.ln_shift = nullptr, r.epi.ln_shift_stride = 0, r.epi.ln_nt = 1;
      r.epi.ln_rstd_out = save ? F(p.rstd_tail[d]) : nullptr;
      r.epi.outOP = OP(p.upop[d]);
    }
    return run_conv(engine, r, st);
  };

  const float* cur = nullptr;
  for (int d = 0; d < D; ++d) {
    const int Hd = H >> d, Wd = W >> d, C = h->d.hidden_channels[d];
    const int nb = h->d.hidden_blocks[d];
    {
      const int ci = h->head_conv[d];
      ConvProblem q{};
      q.in = d == 0 ? OP(p.in_op) : OP(p.xs2[d - 1]);
      q.wpk = wf(ci), q.N = N, q.H = Hd, q.W = Wd, q.Cin = h->convs[ci].kf, q.Cout = C, q.stride = d == 0 ? 1 : 2;
      q.in_strided = d > 0 && strided;
      q.mode = mode, q.epi.bias = bias(ci);
      q.epi.outF = nb == 0 ? F(p.skip[d]) : F(p.x0[d]);
      if (nb == 0 && d < D - 1 && strided) q.epi.outOP = OP(p.xs2[d]);  // operand of the next head
      fuse_ln(q, nb > 0 ? h->desc_blk[d][0] : -1);
      SDAB_TRY(run_conv(engine, q, st, h->convs[ci].cin));
      cur = q.epi.outF;
    }
    for (int b = 0; b < nb; ++b) {
      float* dst = b == nb - 1 ? F(p.skip[d]) : (cur == F(p.x0[d]) ? F(p.x1[d]) : F(p.x0[d]));
      const int next_j = b + 1 < nb ? h->desc_blk[d][b + 1] : (d == D - 1 ? h->asc_blk[d][0] : -1);
      // the last block of a level that feeds a head also leaves its output as that head's operand
      bf16* head_op = (b == nb - 1 && d < D - 1 && strided) ? OP(p.xs2[d]) : nullptr;
      SDAB_TRY(block(d, h->desc_blk[d][b], h->desc_c1[d][b], cur, dst, head_op, next_j));
      cur = dst;
    }
    if (d < D - 1 && !strided) SDAB_TRY(f_to_operand(cur, OP(p.xs2[d]), N, Hd, Wd, C, 1, st));
  }
  for (int d = D - 1; d >= 0; --d) {
    const int Hd = H >> d, Wd = W >> d, C = h->d.hidden_channels[d];
    const int nb = h->d.hidden_blocks[d];
    bool have_finop = false;
    for (int b = 0; b < nb; ++b) {
      float* dst = cur == F(p.x0[d]) ? F(p.x1[d]) : F(p.x0[d]);
      const bool fin = d == 0 && b == nb - 1;
      const int next_j = b + 1 < nb ? h->asc_blk[d][b + 1] : -1;
      SDAB_TRY(block(d, h->asc_blk[d][b], h->asc_c1[d][b], cur, dst, fin ? OP(p.finop) : nullptr, next_j,
                     fuse && d > 0 && b == nb - 1));
      have_finop = have_finop || fin;
      cur = dst;
    }
    const int ci = h->tail_conv[d];
    if (d > 0 && fuse) {
      // tail = LN -> nearest x2 -> conv (sda/nn.py:161-170) in sub-pixel form: the LayerNorm output stays
      // at the LOW resolution and each output parity (po, pp) is a 2x2-tap conv with summed weights
      // (4 / 9 of the FLOPs, a quarter of the operand traffic)
      if (nb == 0)  // otherwise produced by the epilogue of the level's last block (tail_ln)
        SDAB_TRY(ln_forward(cur, nullptr, 0, 1, OP(p.upop[d]), save ? F(p.rstd_tail[d]) : nullptr, N, Hd, Wd, C, 0, st));
      const int next_j = h->d.hidden_blocks[d - 1] > 0 ? h->asc_blk[d - 1][0] : -1;
      for (int cls = 0; cls < 4; ++cls) {
        const int po = cls >> 1, pp = cls & 1;
        ConvProblem q{};
        q.in = OP(p.upop[d]), q.wpk = (const bf16*)(pk + h->convs[ci].off_tf), q.wtaps = 16;
        q.N = N, q.H = Hd, q.W = Wd, q.Cin = C, q.Cout = h->d.hidden_channels[d - 1], q.stride = 1, q.mode = mode;
        q.os = 2, q.oh0 = po, q.ow0 = pp;
        q.taps.n = 4;
        for (int t = 0; t < 4; ++t) {
          q.taps.ca[t] = (unsigned char)(po + (t >> 1)), q.taps.cb[t] = (unsigned char)(pp + (t & 1));
          q.taps.cp[t] = 0, q.taps.wtap[t] = (unsigned char)(cls * 4 + t);
        }
        q.epi.bias = bias(ci), q.epi.res = F(p.skip[d - 1]), q.epi.outF = F(p.x0[d - 1]);
        fuse_ln(q, next_j);
        SDAB_TRY(run_conv(engine, q, st));  // algorithmic FLOPs: the reference's 9 taps per high-resolution pixel
      }
      cur = F(p.x0[d - 1]);
    } else if (d > 0) {
      SDAB_TRY(ln_forward(cur, nullptr, 0, 1, OP(p.upop[d]), save ? F(p.rstd_tail[d]) : nullptr, N, Hd, Wd, C, 1, st));
      ConvProblem q{};
      q.in = OP(p.upop[d]), q.wpk = wf(ci), q.N = N, q.H = 2 * Hd, q.W = 2 * Wd, q.Cin = C;
      q.Cout = h->d.hidden_channels[d - 1], q.stride = 1, q.mode = mode;
      q.epi.bias = bias(ci), q.epi.res = F(p.skip[d - 1]), q.epi.outF = F(p.x0[d - 1]);
      fuse_ln(q, h->d.hidden_blocks[d - 1] > 0 ? h->asc_blk[d - 1][0] : -1);
      SDAB_TRY(run_conv(engine, q, st));
      cur = q.epi.outF;
    } else {
      if (!have_finop) SDAB_TRY(f_to_operand(cur, OP(p.finop), N, Hd, Wd, C, 0, st));
      ConvProblem q{};
      q.in = OP(p.finop), q.wpk = wf(ci), q.N = N, q.H = Hd, q.W = Wd, q.Cin = C, q.Cout = h->convs[ci].nf;
      q.stride = 1, q.mode = mode, q.epi.bias = bias(ci), q.epi.outF = F(p.outf);
      SDAB_TRY(run_conv(engine, q, st, -1, h->d.out_channels));
      if (wio)
        SDAB_TRY(unpack_f_fold(F(p.outf), out, *wio, N, h->convs[ci].nf, Hd, Wd, st));
      else
        SDAB_TRY(unpack_f_to_nchw(F(p.outf), out, N, h->d.out_channels, h->convs[ci].nf, Hd, Wd, st));
    }
  }
  if (save) {
    h->saved = true, h->saved_level = save == 2 ? 2 : 1;
    h->sN = N, h->sNt = Nt, h->sH = H, h->sW = W, h->sws = workspace;
  }
  return SDAB_OK;
}

// parameter-gradient targets of the training backward (host arrays of device pointers)
struct WTargets {
  float* const* dw;
  float* const* db;
  float* dshift;
};

// wio != null: gout is the cotangent of the FOLDED score (B, L, C, H, W) (the adjoint of fold is addressing) and gx
// receives the window input-gradients (N, (2k+1) C, H, W) without the context channels
int backward_impl(sdab_unet* h, const float* gout, float* gx, void* workspace, size_t workspace_bytes, int mode,
                  int engine, void* stream, const WTargets* wt, const WindowIO* wio = nullptr) {
  SDAB_REQUIRE(h && gout && gx && workspace, "null argument");
  if (!h->saved || h->sws != workspace)
    return fail(SDAB_ERR_STATE, "the backward pass needs a preceding sdab_unet_forward(save != 0) on the same workspace");
  if (wt && h->saved_level != 2)
    return fail(SDAB_ERR_STATE, "parameter gradients need sdab_unet_forward(save = 2)");
  SDAB_REQUIRE(mode == SDAB_MODE_BF16X3 || mode == SDAB_MODE_BF16, "unknown mode");
  SDAB_REQUIRE(engine == SDAB_ENGINE_UMMA || engine == SDAB_ENGINE_SIMT, "unknown engine");
  const int N = h->sN, Nt = h->sNt, H = h->sH, W = h->sW;
  const Plan p = make_plan(h, N, N, H, W, true, h->saved_level == 2);
  SDAB_REQUIRE(workspace_bytes >= p.total, "workspace too small");

  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  const uint8_t* pk = h->packed;
  const int D = h->d.depth;
  const int act = h->d.activation == SDAB_ACT_SILU ? 1 : 2;
  auto F = [&](size_t off) { return (float*)(ws + off); };
  auto OP = [&](size_t off) { return (bf16*)(ws + off); };
  auto wb = [&](int ci) { return (const bf16*)(pk + h->convs[ci].off_bwd); };
  // aliases of dead forward temporaries
  auto G0 = [&](int d) { return F(p.x0[d]); };
  auto G1 = [&](int d) { return F(p.x1[d]); };
  auto GA = [&](int d) { return F(p.skip[d]); };
  auto GOP = [&](int d) { return OP(p.hop[d]); };

  if (wt) {
    for (size_t i = 0; i < h->convs.size(); ++i) {
      SDAB_TRY(fill_zero(wt->dw[i], (size_t)h->convs[i].cout * h->convs[i].cin * 9 * sizeof(float), st));
      SDAB_TRY(fill_zero(wt->db[i], (size_t)h->convs[i].cout * sizeof(float), st));
    }
    SDAB_TRY(fill_zero(wt->dshift, (size_t)Nt * h->shift_rows * sizeof(float), st));
  }
  // weight / bias gradient of convolution ci (no-op outside training)
  auto wgrad = [&](int ci, const float* gF, const bf16* gOP, int Cg, const bf16* xOP, const float* xF, int x_kind,
                   int Cx, int Ho, int Wo) -> int {
    if (!wt) return SDAB_OK;
    WgradProblem w{};
    w.gF = gF, w.gOP = gOP, w.xOP = xOP, w.xF = xF, w.x_kind = x_kind, w.act = act;
    w.N = N, w.H = Ho, w.W = Wo, w.Cg = Cg, w.Cx = Cx, w.cout = h->convs[ci].cout, w.cin = h->convs[ci].cin;
    w.dw = wt->dw[ci], w.db = wt->db[ci];
    w.partial = (float*)(ws + p.wpart), w.partial_bytes = kWgradPartialBytes;
    if (engine == SDAB_ENGINE_UMMA && wgrad_umma_supported(w)) return conv3x3_wgrad_umma(w, mode, st);
    return conv3x3_wgrad(w, st);
  };

  // Weight gradient of the stride-2 head ci (output cotangent g at Ho x Wo, input xs2 in the parity layout):
  // on the tcgen05 engine, four stride-1 sub-problems by the parity (qa, qb) of the input pixel --
  // haloed input pixel (2 o + a, 2 p + b) is pixel (o + a / 2, p + b / 2) of parity image (a & 1, b & 1).
  auto wgrad_head = [&](int ci, const float* gF, const bf16* gOP, int C, const bf16* xs2, int Cx, int Ho, int Wo) -> int {
    if (!wt) return SDAB_OK;
    WgradProblem w{};
    w.gF = gF, w.gOP = gOP, w.xOP = xs2, w.x_kind = 0, w.act = act, w.N = N, w.H = Ho, w.W = Wo, w.Cg = C, w.Cx = Cx;
    w.cout = h->convs[ci].cout, w.cin = h->convs[ci].cin, w.dw = wt->dw[ci], w.db = wt->db[ci];
    w.partial = (float*)(ws + p.wpart), w.partial_bytes = kWgradPartialBytes;
    if (engine == SDAB_ENGINE_UMMA && wgrad_umma_supported(w)) {
      for (int qa = 0; qa < 2; ++qa)
        for (int qb = 0; qb < 2; ++qb) {
          WgradProblem s = w;
          s.x_par = 1 + qa * 2 + qb, s.db = (qa | qb) ? nullptr : w.db;
          for (int sa = 0; sa < (qa ? 1 : 2); ++sa)
            for (int sb = 0; sb < (qb ? 1 : 2); ++sb) {
              s.tl_sa[s.ntl] = (unsigned char)sa, s.tl_sb[s.ntl] = (unsigned char)sb;
              s.tl_mask[s.ntl++] = (unsigned short)(1u << (3 * (2 * sa + qa) + 2 * sb + qb));
            }
          SDAB_TRY(conv3x3_wgrad_umma(s, mode, st));
        }
      return SDAB_OK;
    }
    w.x_kind = 1;
    return conv3x3_wgrad(w, st);
  };
  // Weight gradient of the tail ci = LN -> nearest x2 -> conv (output cotangent at 2 Hl x 2 Wl: gF, and in the
  // parity layout gS2; input xlo at Hl x Wl): four stride-1 sub-problems by the output parity (po, pp) --
  // output pixel (2 i + po, 2 j + pp) reads xlo[i + floor((po + a - 1) / 2), j + floor((pp + b - 1) / 2)] for
  // tap (a, b), so each low-resolution offset collects one or two taps per axis.
  auto wgrad_tail = [&](int ci, const float* gF, const bf16* gS2, int C, const bf16* xlo, int Cx, int Hl, int Wl) -> int {
    if (!wt) return SDAB_OK;
    WgradProblem w{};
    w.gF = gF, w.gOP = gS2, w.xOP = xlo, w.x_kind = 0, w.act = act, w.N = N, w.H = Hl, w.W = Wl, w.Cg = C, w.Cx = Cx;
    w.cout = h->convs[ci].cout, w.cin = h->convs[ci].cin, w.dw = wt->dw[ci], w.db = nullptr;
    w.partial = (float*)(ws + p.wpart), w.partial_bytes = kWgradPartialBytes;
    if (engine == SDAB_ENGINE_UMMA && gS2 && wgrad_umma_supported(w)) {
      SDAB_TRY(f_channel_sum(gF, wt->db[ci], (size_t)N * 4 * Hl * Wl, C, w.cout, st));
      for (int po = 0; po < 2; ++po)
        for (int pp = 0; pp < 2; ++pp) {
          WgradProblem s = w;
          s.g_par = 1 + ((po + 1) & 1) * 2 + ((pp + 1) & 1), s.g_dh = (po + 1) >> 1, s.g_dw = (pp + 1) >> 1;
          for (int da = po - 1; da <= po; ++da)
            for (int db = pp - 1; db <= pp; ++db) {
              unsigned mask = 0;
              for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) {
                  const int fa = (po + a - 1) >= 0 ? (po + a - 1) / 2 : -1, fb = (pp + b - 1) >= 0 ? (pp + b - 1) / 2 : -1;
                  if (fa == da && fb == db) mask |= 1u << (3 * a + b);
                }
              s.tl_sa[s.ntl] = (unsigned char)(da + 1), s.tl_sb[s.ntl] = (unsigned char)(db + 1);
              s.tl_mask[s.ntl++] = (unsigned short)mask;
            }
          SDAB_TRY(conv3x3_wgrad_umma(s, mode, st));
        }
      return SDAB_OK;
    }
    // CUDA-core path: the upsample is folded into the operand loader (the SIMT engine's tail input is
    // already at the high resolution)
    w.gOP = nullptr, w.H = 2 * Hl, w.W = 2 * Wl, w.db = wt->db[ci];
    w.x_kind = engine == SDAB_ENGINE_UMMA ? 2 : 0;
    return conv3x3_wgrad(w, st);
  };

  // backward of one block: cur (F, operand in GOP[d]) -> other ping-pong buffer (+ GOP[d])
  auto block_bwd = [&](int d, int j, int c1, const float* cur, float* dst) -> int {
    const int Hd = H >> d, Wd = W >> d, C = h->d.hidden_channels[d];
    // conv2: output cotangent = cur (operand copy in GOP[d]), input = the saved act(conv1)
    SDAB_TRY(wgrad(c1 + 1, cur, GOP(d), C, OP(p.hsave[j]), nullptr, 0, C, Hd, Wd));
    ConvProblem q{};
    q.in = GOP(d), q.wpk = wb(c1 + 1), q.N = N, q.H = Hd, q.W = Wd, q.Cin = C, q.Cout = C, q.stride = 1, q.mode = mode;
    q.epi.dact = F(p.c1[j]), q.epi.dact_kind = act, q.epi.outOP = OP(p.gc1op[d]);
    SDAB_TRY(run_conv(engine, q, st));
    // conv1: output cotangent = gC1 (operand tensor), input = the saved LayerNorm output
    SDAB_TRY(wgrad(c1, nullptr, OP(p.gc1op[d]), C, OP(p.aop[j]), nullptr, 0, C, Hd, Wd));
    ConvProblem r{};
    r.in = OP(p.gc1op[d]), r.wpk = wb(c1), r.N = N, r.H = Hd, r.W = Wd, r.Cin = C, r.Cout = C, r.stride = 1,
    r.mode = mode;
    if (engine == SDAB_ENGINE_UMMA) {
      // LayerNorm adjoint + residual fused into the epilogue of conv1^T (ConvEpilogue::ln == 2)
      r.epi.ln = 2, r.epi.ln_a = OP(p.aop[j]), r.epi.ln_rstd_in = F(p.rstd[j]);
      r.epi.res = cur, r.epi.outF = dst, r.epi.outOP = GOP(d);
      SDAB_TRY(run_conv(engine, r, st));
    } else {
      r.epi.outF = GA(d);
      SDAB_TRY(run_conv(engine, r, st));
      SDAB_TRY(ln_backward(GA(d), OP(p.aop[j]), F(p.rstd[j]), cur, dst, GOP(d), N, Hd, Wd, C, 0, st));
    }
    // the shift enters through the LayerNorm only: its cotangent is dst - cur summed over the pixels
    if (wt)
      SDAB_TRY(shift_grad(dst, cur, wt->dshift + h->block_shift_off[j], h->shift_rows, Nt, N, Hd, Wd, C, st));
    return SDAB_OK;
  };

  const int kout = round_up(h->d.out_channels, 32);
  if (wio)
    SDAB_TRY(pack_fold_adjoint_to_op(gout, OP(p.gout_op), *wio, N, kout, H, W, st));
  else
    SDAB_TRY(pack_nchw_to_op(gout, OP(p.gout_op), N, h->d.out_channels, kout, H, W, 0, st));
  const float* cur;
  {
    const int ci = h->tail_conv[0];
    ConvProblem q{};
    q.in = OP(p.gout_op), q.wpk = wb(ci), q.N = N, q.H = H, q.W = W, q.Cin = kout, q.Cout = h->d.hidden_channels[0];
    q.stride = 1, q.mode = mode, q.epi.outF = G0(0), q.epi.outOP = GOP(0);
    SDAB_TRY(wgrad(ci, nullptr, OP(p.gout_op), kout, OP(p.finop), nullptr, 0, h->d.hidden_channels[0], H, W));
    SDAB_TRY(run_conv(engine, q, st, h->d.out_channels));
    cur = G0(0);
  }
  std::vector<const float*> gskip(D, nullptr);
  for (int d = 0; d < D; ++d) {  // ascent, reversed
    const int Hd = H >> d, Wd = W >> d, C = h->d.hidden_channels[d];
    for (int b = h->d.hidden_blocks[d] - 1; b >= 0; --b) {
      float* dst = cur == G0(d) ? G1(d) : G0(d);
      SDAB_TRY(block_bwd(d, h->asc_blk[d][b], h->asc_c1[d][b], cur, dst));
      cur = dst;
    }
    if (d < D - 1 && engine == SDAB_ENGINE_UMMA) {
      // transpose of the sub-pixel tail: a 4x4 stride-2 conv of the cotangent (parity layout) with the
      // summed, transposed weights, directly at the low resolution, LayerNorm adjoint fused
      gskip[d] = cur;
      const int ci = h->tail_conv[d + 1];
      const int Cn = h->d.hidden_channels[d + 1];
      // tail conv: output cotangent = cur, input = nearest x2 of the saved low-resolution LayerNorm output
      // (outside training the parity images are read from cur's operand GOP(d) with a TMA element stride of 2)
      if (wt || !(getenv("SDAB_STRIDED_TMA") ? atoi(getenv("SDAB_STRIDED_TMA")) : 1))
        SDAB_TRY(f_to_operand(cur, OP(p.xs2g[d]), N, Hd, Wd, C, 1, st));
      SDAB_TRY(wgrad_tail(ci, cur, OP(p.xs2g[d]), C, OP(p.upop[d + 1]), Cn, Hd / 2, Wd / 2));
      ConvProblem q{};
      static const int strided_env = getenv("SDAB_STRIDED_TMA") ? atoi(getenv("SDAB_STRIDED_TMA")) : 1;
      const bool parity_copy = wt || !strided_env;
      q.in = parity_copy ? OP(p.xs2g[d]) : GOP(d), q.in_s2 = 1, q.in_strided = parity_copy ? 0 : 1;
      q.wpk = (const bf16*)(pk + h->convs[ci].off_tb), q.wtaps = 16;
      q.N = N, q.H = Hd / 2, q.W = Wd / 2, q.Cin = C, q.Cout = Cn, q.stride = 1, q.mode = mode;
      q.taps.n = 16;
      for (int t = 0; t < 16; ++t) {
        const int a = t >> 2, b = t & 3;
        q.taps.ca[t] = (unsigned char)(a >> 1), q.taps.cb[t] = (unsigned char)(b >> 1);
        q.taps.cp[t] = (unsigned char)((a & 1) * 2 + (b & 1)), q.taps.wtap[t] = (unsigned char)t;
      }
      q.epi.ln = 2, q.epi.ln_a = OP(p.upop[d + 1]), q.epi.ln_rstd_in = F(p.rstd_tail[d + 1]);
      q.epi.outF = G0(d + 1), q.epi.outOP = GOP(d + 1);
      SDAB_TRY(run_conv(engine, q, st, -1, -1, 4.0));  // algorithmic FLOPs are counted at the high resolution
      cur = G0(d + 1);
    } else if (d < D - 1) {
      gskip[d] = cur;
      const int ci = h->tail_conv[d + 1];
      const int Cn = h->d.hidden_channels[d + 1];
      SDAB_TRY(wgrad_tail(ci, cur, nullptr, C, OP(p.upop[d + 1]), Cn, Hd / 2, Wd / 2));
      ConvProblem q{};
      q.in = GOP(d), q.wpk = wb(ci), q.N = N, q.H = Hd, q.W = Wd, q.Cin = C, q.Cout = Cn, q.stride = 1, q.mode = mode;
      q.epi.outF = F(p.gup[d + 1]);
      SDAB_TRY(run_conv(engine, q, st));
      SDAB_TRY(ln_backward(F(p.gup[d + 1]), OP(p.upop[d + 1]), F(p.rstd_tail[d + 1]), nullptr, G0(d + 1), GOP(d + 1), N,
                           Hd / 2, Wd / 2, Cn, 1, st));
      cur = G0(d + 1);
    }
  }
  for (int d = D - 1; d >= 0; --d) {  // descent, reversed
    const int Hd = H >> d, Wd = W >> d, C = h->d.hidden_channels[d];
    for (int b = h->d.hidden_blocks[d] - 1; b >= 0; --b) {
      float* dst = cur == G0(d) ? G1(d) : G0(d);
      SDAB_TRY(block_bwd(d, h->desc_blk[d][b], h->desc_c1[d][b], cur, dst));
      cur = dst;
    }
    const int ci = h->head_conv[d];
    // head conv: output cotangent = cur; input = the level below in the parity layout (stride 2), or the
    // packed network input
    if (d > 0)
      SDAB_TRY(wgrad_head(ci, cur, GOP(d), C, OP(p.xs2[d - 1]), h->d.hidden_channels[d - 1], Hd, Wd));
    else
      SDAB_TRY(wgrad(ci, cur, GOP(0), C, OP(p.in_op), nullptr, 0, round_up(h->d.in_channels, 32), Hd, Wd));
    if (d > 0 && engine == SDAB_ENGINE_UMMA) {
      // transpose of the stride-2 head by output parity: gx[2i + po, 2j + pp] only receives the taps
      // with a = po + 1 (mod 2), b = pp + 1 (mod 2) -- 1, 2, 2 and 4 taps instead of 9 on a
      // zero-upsampled cotangent
      const int Cp = h->d.hidden_channels[d - 1];
      float* dst = gskip[d - 1] == G0(d - 1) ? G1(d - 1) : G0(d - 1);
      static const int off1[2][2] = {{1, -1}, {2, 1}}, wt1[2][2] = {{1, -1}, {2, 0}}, cnt1[2] = {1, 2};
      for (int cls = 0; cls < 4; ++cls) {
        const int po = cls >> 1, pp = cls & 1;
        ConvProblem q{};
        q.in = GOP(d), q.wpk = wb(ci), q.N = N, q.H = Hd, q.W = Wd, q.Cin = C, q.Cout = Cp, q.stride = 1, q.mode = mode;
        q.os = 2, q.oh0 = po, q.ow0 = pp;
        q.taps.n = cnt1[po] * cnt1[pp];
        for (int ta = 0, t = 0; ta < cnt1[po]; ++ta)
          for (int tb = 0; tb < cnt1[pp]; ++tb, ++t) {
            q.taps.ca[t] = (unsigned char)off1[po][ta], q.taps.cb[t] = (unsigned char)off1[pp][tb];
            q.taps.cp[t] = 0, q.taps.wtap[t] = (unsigned char)(3 * wt1[po][ta] + wt1[pp][tb]);
          }
        q.epi.res = gskip[d - 1], q.epi.outF = dst, q.epi.outOP = GOP(d - 1);
        SDAB_TRY(run_conv(engine, q, st, -1, -1, q.taps.n / 9.0));
      }
      cur = dst;
    } else if (d > 0) {
      const int Cp = h->d.hidden_channels[d - 1];
      SDAB_TRY(f_to_operand(cur, OP(p.gz[d]), N, 2 * Hd, 2 * Wd, C, 2, st));
      float* dst = gskip[d - 1] == G0(d - 1) ? G1(d - 1) : G0(d - 1);
      ConvProblem q{};
      q.in = OP(p.gz[d]), q.wpk = wb(ci), q.N = N, q.H = 2 * Hd, q.W = 2 * Wd, q.Cin = C, q.Cout = Cp, q.stride = 1;
      q.mode = mode, q.epi.res = gskip[d - 1], q.epi.outF = dst, q.epi.outOP = GOP(d - 1);
      SDAB_TRY(run_conv(engine, q, st, -1, -1, 0.25));
      cur = dst;
    } else {
      ConvProblem q{};
      q.in = GOP(0), q.wpk = wb(ci), q.N = N, q.H = H, q.W = W, q.Cin = C, q.Cout = h->convs[ci].nb, q.stride = 1;
      q.mode = mode, q.epi.outF = F(p.gxf);
      SDAB_TRY(run_conv(engine, q, st, -1, h->d.in_channels));
      SDAB_TRY(unpack_f_to_nchw(F(p.gxf), gx, N, wio ? (2 * wio->order + 1) * wio->C : h->d.in_channels,
                                h->convs[ci].nb, H, W, st));
    }
  }
  return SDAB_OK;
}

int check_windows(const sdab_unet* h, const WindowIO& w, int w_end) {
  SDAB_REQUIRE(w.B >= 1 && w.C >= 1 && w.Cc >= 0 && w.order >= 1 && w.L >= 2 * w.order + 1,
               "trajectory shorter than the window (MCScoreNet.unfold raises too)");
  SDAB_REQUIRE(h->d.in_channels == (2 * w.order + 1) * w.C + w.Cc && h->d.out_channels == (2 * w.order + 1) * w.C,
               "the network's channels do not match (2k+1) C (+ context)");
  SDAB_REQUIRE(w.w_begin >= 0 && w_end > w.w_begin && w_end <= w.B * (w.L - 2 * w.order), "window range out of bounds");
  SDAB_REQUIRE(w.cap == 0 || (w.per >= w_end - w.w_begin && w.cap >= w.per), "invalid shard geometry");
  return SDAB_OK;
}

}  // namespace

extern "C" {

int sdab_unet_forward(sdab_unet* h, const float* x, const float* y, int Nt, int N, int H, int W, float* out,
                      void* workspace, size_t workspace_bytes, int save, int mode, int engine, void* stream) {
  return forward_impl(h, x, y, Nt, N, H, W, out, workspace, workspace_bytes, save, mode, engine, stream, nullptr,
                      nullptr);
}

int sdab_mcscore_forward(sdab_unet* h, const float* x, const float* ctx, const float* y, int B, int L, int C, int Cc,
                         int H, int W, int order, int w_begin, int w_end, float* out, int per, int cap,
                         void* workspace, size_t workspace_bytes, int save, int mode, int engine, void* stream) {
  SDAB_REQUIRE(h && (Cc == 0 || ctx), "null argument");
  const WindowIO w{B, L, C, Cc, order, w_begin, per, cap};
  SDAB_TRY(check_windows(h, w, w_end));
  return forward_impl(h, x, y, 1, w_end - w_begin, H, W, out, workspace, workspace_bytes, save, mode, engine, stream, &w,
                      ctx);
}

int sdab_mcscore_dgrad(sdab_unet* h, const float* gs, float* gwin, int B, int L, int C, int Cc, int H, int W, int order,
                       int w_begin, int w_end, void* workspace, size_t workspace_bytes, int mode, int engine,
                       void* stream) {
  SDAB_REQUIRE(h != nullptr, "null argument");
  const WindowIO w{B, L, C, Cc, order, w_begin, 0, 0};
  SDAB_TRY(check_windows(h, w, w_end));
  SDAB_REQUIRE(h->sN == w_end - w_begin, "window range differs from the saved forward pass");
  return backward_impl(h, gs, gwin, workspace, workspace_bytes, mode, engine, stream, nullptr, &w);
}

int sdab_unet_dgrad(sdab_unet* h, const float* gout, float* gx, void* workspace, size_t workspace_bytes, int mode,
                    int engine, void* stream) {
  return backward_impl(h, gout, gx, workspace, workspace_bytes, mode, engine, stream, nullptr);
}

int sdab_unet_backward(sdab_unet* h, const float* gout, float* gx, float* const* conv_dw_host,
                       float* const* conv_db_host, float* dshift, void* workspace, size_t workspace_bytes, int mode,
                       int engine, void* stream) {
  SDAB_REQUIRE(conv_dw_host && conv_db_host && dshift, "null argument");
  const WTargets wt{conv_dw_host, conv_db_host, dshift};
  return backward_impl(h, gout, gx, workspace, workspace_bytes, mode, engine, stream, &wt);
}

int sdab_unet_shift_rows(const sdab_unet* h) { return h ? h->shift_rows : 0; }

}  // extern "C"
