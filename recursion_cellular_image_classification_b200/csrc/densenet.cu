// densenet.cu — the DenseNet-121 executor: one data-parallel rank's forward, loss, backward and SGD,
// enqueued natively on a CUDA stream (capturable in a CUDA graph; no host synchronisation inside).
//
// What it replaces in the reference: TwoSitesNN.forward (cell_classifier/models.py:41-57) with the north
// star's DenseNet-121 trunk (torchvision densenet121, 6-channel stem per models.py:17-27), the train step
// of ignite's create_supervised_trainer (cell_classifier/train.py:44: zero_grad / forward / CrossEntropy /
// backward / step) and torch.optim.SGD (main.py:89-93).  Parameters, gradients and momentum are flat fp32
// buffers in torchvision's named_parameters() order so a state_dict maps 1:1.
//
// Data layout in HBM (all activations bf16 NHWC):
//   * one channel-concatenated buffer X_b [B,H_b,W_b,Ctot_b] per dense block; every dense layer's 3x3
//     conv writes its 32 new channels at its channel offset (no torch.cat copies);
//   * per-channel sum / sum-of-squares of every produced channel are accumulated once by the producing
//     kernel's epilogue and shared by all later BatchNorms that read that channel;
//   * backward keeps one gradient buffer G_b of the same shape.  A consumer j with its own BatchNorm adds
//     scale_j * dy_j into G_b from the dgrad epilogue and folds the two per-channel BatchNorm-backward
//     correction terms into corrA/corrB; the exact gradient of a channel,
//       dX = G - corrA - xhat * corrB,
//     is materialised only for the 32-channel slice whose producer is being differentiated.
#include <stdlib.h>
#include <vector>
#include "conv_gemm.cuh"
#include "elementwise.cuh"

namespace rxb {

int pick_bn(int n);

struct BnLayer {
  int C = 0;
  long long gamma_off = 0, beta_off = 0, rm_off = 0, rv_off = 0;
  BnFold fold = {nullptr, nullptr, nullptr, nullptr};
  float* dsum = nullptr;
  float* dsq = nullptr;
};
struct ConvLayer {
  long long w_off = 0;
  int N = 0, K = 0;
  long long fwd_off = 0, dgrad_off = -1;
};
struct DenseLayer {
  int Cin = 0;
  BnLayer bn1, bn2;
  ConvLayer c1, c2;
  __nv_bfloat16* Y = nullptr;
  float* ysum = nullptr;
  float* ysq = nullptr;
};
struct Block {
  int H = 0, W = 0, C0 = 0, Ctot = 0;
  long long M = 0;
  std::vector<DenseLayer> layers;
  __nv_bfloat16* X = nullptr;
  __nv_bfloat16* G = nullptr;
  float *xsum = nullptr, *xsq = nullptr, *corrA = nullptr, *corrB = nullptr;
};
struct Transition {
  BnLayer bn;
  ConvLayer conv;
  __nv_bfloat16* P = nullptr;
};

}  // namespace rxb

struct rxb_dn121 {
  rxb_dn121_config cfg;
  int training = 0;
  float *params = nullptr, *grads = nullptr, *momentum = nullptr, *buffers = nullptr;
  long long n_params = 0, n_buffers = 0;
  // network
  rxb::ConvLayer conv0;
  rxb::BnLayer bn0, bn5;
  rxb::Block blocks[4];
  rxb::Transition trans[3];
  long long fc_w_off = 0, fc_b_off = 0;
  // workspace
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  int Hs = 0, Ws = 0;  // stem conv output size
  __nv_bfloat16* S0 = nullptr;
  float *s0sum = nullptr, *s0sq = nullptr;
  uint8_t* pool_idx = nullptr;
  __nv_bfloat16 *dy0 = nullptr, *dZ = nullptr, *dy2 = nullptr, *dX0 = nullptr, *dP = nullptr;
  float *feat = nullptr, *logits = nullptr, *dlogits = nullptr, *dfeat = nullptr, *loss_rows = nullptr;
  uint8_t* zero_begin = nullptr;  // region cleared at the start of every step (statistics, corrections)
  size_t zero_bytes = 0;
  __nv_bfloat16* arena = nullptr;  // bf16 GEMM operand copies of the conv weights
  long long arena_elems = 0;
  rxb::RepackJob* jobs_dev = nullptr;
  std::vector<rxb::RepackJob> jobs;
  long long max_job_elems = 0;
  bool jobs_uploaded = false;
};

namespace rxb {

static const int kBlockLayers[4] = {6, 12, 24, 16};
constexpr int kGrowth = 32, kBott = 128, kInit = 64;

// ---- bump allocator that also works as a dry run (base == nullptr)
struct Bump {
  uint8_t* base;
  size_t off = 0;
  explicit Bump(uint8_t* b) : base(b) {}
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

static void plan_bn(BnLayer& bn, int C, long long& poff, long long& boff) {
  bn.C = C;
  bn.gamma_off = poff; poff += C;
  bn.beta_off = poff; poff += C;
  bn.rm_off = boff; boff += C;
  bn.rv_off = boff; boff += C;
}

// Fills offsets (and, when ws != nullptr, workspace pointers).  Returns workspace bytes.
static size_t plan(rxb_dn121& n, uint8_t* ws) {
  const rxb_dn121_config& c = n.cfg;
  long long poff = 0, boff = 0, aoff = 0;
  n.jobs.clear();
  n.max_job_elems = 0;
  auto add_job = [&](long long src, int type, int N, int K, long long elems) {
    RepackJob j;
    j.src_off = src; j.dst_off = aoff; j.type = type; j.N = N; j.K = K; j.pad = 0;
    n.jobs.push_back(j);
    long long at = aoff;
    aoff += (elems + 63) & ~63ll;
    if (elems > n.max_job_elems) n.max_job_elems = elems;
    return at;
  };
  // parameters in torchvision named_parameters() order
  n.conv0.w_off = poff; n.conv0.N = kInit; n.conv0.K = 32; poff += 64 * 6 * 49;
  n.conv0.fwd_off = add_job(n.conv0.w_off, RP_STEM_FWD, 64, 32, 16ll * 64 * 32);
  plan_bn(n.bn0, kInit, poff, boff);
  int C = kInit;
  int H = c.H / 4, W = c.W / 4;
  for (int b = 0; b < 4; ++b) {
    Block& blk = n.blocks[b];
    blk.H = H; blk.W = W; blk.C0 = C; blk.Ctot = C + kGrowth * kBlockLayers[b];
    blk.M = (long long)c.B * H * W;
    blk.layers.assign(kBlockLayers[b], DenseLayer());
    for (int i = 0; i < kBlockLayers[b]; ++i) {
      DenseLayer& L = blk.layers[i];
      L.Cin = C + kGrowth * i;
      plan_bn(L.bn1, L.Cin, poff, boff);
      L.c1.w_off = poff; L.c1.N = kBott; L.c1.K = L.Cin; poff += (long long)kBott * L.Cin;
      L.c1.fwd_off = add_job(L.c1.w_off, RP_1x1_FWD, kBott, L.Cin, (long long)kBott * L.Cin);
      L.c1.dgrad_off = add_job(L.c1.w_off, RP_1x1_DGRAD, kBott, L.Cin, (long long)kBott * L.Cin);
      plan_bn(L.bn2, kBott, poff, boff);
      L.c2.w_off = poff; L.c2.N = kGrowth; L.c2.K = kBott; poff += (long long)kGrowth * kBott * 9;
      L.c2.fwd_off = add_job(L.c2.w_off, RP_3x3_FWD, kGrowth, kBott, 9ll * kGrowth * kBott);
      L.c2.dgrad_off = add_job(L.c2.w_off, RP_3x3_DGRAD, kGrowth, kBott, 9ll * kGrowth * kBott);
    }
    C = blk.Ctot;
    if (b < 3) {
      Transition& t = n.trans[b];
      plan_bn(t.bn, C, poff, boff);
      t.conv.w_off = poff; t.conv.N = C / 2; t.conv.K = C; poff += (long long)(C / 2) * C;
      t.conv.fwd_off = add_job(t.conv.w_off, RP_1x1_FWD, C / 2, C, (long long)(C / 2) * C);
      t.conv.dgrad_off = add_job(t.conv.w_off, RP_1x1_DGRAD, C / 2, C, (long long)(C / 2) * C);
      C /= 2; H /= 2; W /= 2;
    }
  }
  plan_bn(n.bn5, C, poff, boff);
  n.fc_w_off = poff; poff += (long long)c.num_classes * C;
  n.fc_b_off = poff; poff += c.num_classes;
  n.n_params = poff;
  n.n_buffers = boff;
  n.arena_elems = aoff;

  // ---- workspace
  Bump bp(ws);
  const int tr = n.training;
  n.Hs = c.H / 2; n.Ws = c.W / 2;
  n.arena = bp.take<__nv_bfloat16>(aoff);
  n.jobs_dev = bp.take<RepackJob>(n.jobs.size());
  // zeroed-per-step region: statistics, backward reductions, corrections
  bp.take<uint8_t>(0);
  const size_t zero_start = (bp.off + 255) & ~size_t(255);
  auto take_bn_bwd = [&](BnLayer& bn) {
    bn.dsum = bp.take<float>(bn.C);
    bn.dsq = bp.take<float>(bn.C);
  };
  n.s0sum = bp.take<float>(64);
  n.s0sq = bp.take<float>(64);
  take_bn_bwd(n.bn0);
  take_bn_bwd(n.bn5);
  for (int b = 0; b < 4; ++b) {
    Block& blk = n.blocks[b];
    blk.xsum = bp.take<float>(blk.Ctot);
    blk.xsq = bp.take<float>(blk.Ctot);
    blk.corrA = bp.take<float>(blk.Ctot);
    blk.corrB = bp.take<float>(blk.Ctot);
    for (auto& L : blk.layers) {
      L.ysum = bp.take<float>(kBott);
      L.ysq = bp.take<float>(kBott);
      take_bn_bwd(L.bn1);
      take_bn_bwd(L.bn2);
    }
    if (b < 3) take_bn_bwd(n.trans[b].bn);
  }
  bp.take<uint8_t>(0);
  const size_t zero_end = (bp.off + 255) & ~size_t(255);
  n.zero_begin = ws ? ws + zero_start : nullptr;
  n.zero_bytes = zero_end - zero_start;
  bp.off = zero_end;
  // folded BN parameters (kept from forward for backward)
  auto take_fold = [&](BnLayer& bn) {
    bn.fold.scale = bp.take<float>(bn.C + 64);
    bn.fold.shift = bp.take<float>(bn.C + 64);
    bn.fold.mean = bp.take<float>(bn.C + 64);
    bn.fold.rstd = bp.take<float>(bn.C + 64);
  };
  take_fold(n.bn0);
  take_fold(n.bn5);
  for (int b = 0; b < 4; ++b) {
    for (auto& L : n.blocks[b].layers) { take_fold(L.bn1); take_fold(L.bn2); }
    if (b < 3) take_fold(n.trans[b].bn);
  }
  // activations
  const long long stem_px = (long long)c.B * n.Hs * n.Ws;
  n.S0 = bp.take<__nv_bfloat16>(stem_px * 64);
  n.pool_idx = tr ? bp.take<uint8_t>(n.blocks[0].M * 64) : nullptr;
  long long maxM = 0, max_dx0 = 0, max_dp = 0;
  for (int b = 0; b < 4; ++b) {
    Block& blk = n.blocks[b];
    blk.X = bp.take<__nv_bfloat16>(blk.M * blk.Ctot);
    blk.G = tr ? bp.take<__nv_bfloat16>(blk.M * blk.Ctot) : nullptr;
    __nv_bfloat16* shared_Y = tr ? nullptr : bp.take<__nv_bfloat16>(blk.M * kBott);
    for (auto& L : blk.layers) L.Y = tr ? bp.take<__nv_bfloat16>(blk.M * kBott) : shared_Y;
    if (b < 3) n.trans[b].P = bp.take<__nv_bfloat16>((blk.M / 4) * blk.Ctot);
    if (blk.M > maxM) maxM = blk.M;
    if (blk.M * blk.C0 > max_dx0) max_dx0 = blk.M * blk.C0;
    if (b > 0 && blk.M * n.blocks[b - 1].Ctot > max_dp) max_dp = blk.M * n.blocks[b - 1].Ctot;
  }
  if (tr) {
    n.dy0 = bp.take<__nv_bfloat16>(stem_px * 64);
    n.dZ = bp.take<__nv_bfloat16>(maxM * kGrowth);
    n.dy2 = bp.take<__nv_bfloat16>(maxM * kBott);
    n.dX0 = bp.take<__nv_bfloat16>(max_dx0);
    n.dP = bp.take<__nv_bfloat16>(max_dp);
  }
  n.feat = bp.take<float>((long long)c.B * C);
  n.logits = bp.take<float>((long long)c.B * c.num_classes);
  n.dlogits = bp.take<float>((long long)c.B * c.num_classes);
  n.dfeat = bp.take<float>((long long)c.B * C);
  n.loss_rows = bp.take<float>(c.B);
  bp.take<uint8_t>(0);
  return ((bp.off + 255) & ~size_t(255)) + 256;
}

static int check_cfg(const rxb_dn121_config* c) {
  RXB_CHECK_ARG(c != nullptr, "dn121: null config");
  RXB_CHECK_ARG(c->B >= 1 && c->B <= 4096, "dn121: bad batch %d", c->B);
  RXB_CHECK_ARG(c->H >= 32 && c->W >= 32 && c->H % 32 == 0 && c->W % 32 == 0, "dn121: H, W must be multiples of 32");
  RXB_CHECK_ARG(c->num_classes >= 1, "dn121: bad num_classes");
  return RXB_OK;
}

// ---------------------------------------------------------------------------------------------- helpers
static int conv_store(const rxb_dn121& n, int B, int H, int W, const __nv_bfloat16* A, int ldA, int cin,
                      const ConvLayer& cv, int Cout, int taps, int pad, const BnFold* pro, __nv_bfloat16* out,
                      int ldc, int c_off, float* ssum, float* ssq, cudaStream_t st, const BnPrepArgs* fused = nullptr,
                      int reverse = 0) {
  GemmParams p = {};
  p.reverse = reverse;
  p.B = B; p.H = H; p.W = W;
  p.n_total = Cout;
  p.taps_x = p.taps_y = taps;
  p.pad_x = p.pad_y = pad;
  p.cin = cin;
  p.epi_mode = EPI_STORE;
  p.do_stats = ssum != nullptr;
  p.ch_sum = ssum ? ssum + c_off : nullptr;
  p.ch_sumsq = ssq ? ssq + c_off : nullptr;
  if (pro) { p.scale = pro->scale; p.shift = pro->shift; }
  if (fused) p.prep = *fused;
  return launch_conv_gemm(p, A, ldA, n.arena + cv.fwd_off, out, ldc, c_off, nullptr, 0, cin <= 32 ? 32 : 64,
                          pro != nullptr, st);
}

// dgrad with the fused ReLU/BatchNorm backward epilogue (reduction: bn.dsum = sum dy; sum dy*x follows from W.dW)
static int conv_dgrad_bn(const rxb_dn121& n, int B, int H, int W, const __nv_bfloat16* dOut, int ldD, int cout,
                         const ConvLayer& cv, int Nprime, int taps, int pad, const __nv_bfloat16* X, int ldx,
                         const BnLayer& bn, int out_mode, __nv_bfloat16* out, int ldc, cudaStream_t st,
                         float* fused_dW = nullptr, const BnTailArgs* tail = nullptr, const FixupArgs* fix = nullptr,
                         int reverse = 0) {
  GemmParams p = {};
  p.reverse = reverse;
  p.B = B; p.H = H; p.W = W;
  p.n_total = Nprime;
  p.taps_x = p.taps_y = taps;
  p.pad_x = p.pad_y = pad;
  p.cin = cout;
  p.epi_mode = EPI_DGRAD_BN;
  p.out_mode = out_mode;
  p.do_stats = 1;
  p.ch_sum = bn.dsum; p.ch_sumsq = bn.dsq;
  p.e_scale = bn.fold.scale; p.e_shift = bn.fold.shift;
  p.e_gamma = n.params + bn.gamma_off; p.e_beta = n.params + bn.beta_off;   // degenerate channels: direct reductions
  p.wg_dW = fused_dW;   // the conv's weight gradient accumulated by the same kernel
  if (tail) p.tail = *tail;   // ... and the BatchNorm-backward reductions that follow it (no bn_bwd_finalize launch)
  if (fix) p.fix = *fix;      // 3x3: dOut derived from the concat buffers on load (no grad_fixup launch)
  return launch_conv_gemm(p, dOut, ldD, n.arena + cv.dgrad_off, out, ldc, 0, X, ldx, cout <= 32 ? 32 : 64, false, st);
}

static int conv_wgrad_any(int B, int H, int W, const __nv_bfloat16* A, int ldA, int cin, int taps, int pad,
                          const BnFold* pro, const __nv_bfloat16* dOut, int ldD, int cout, float* dW, int w_mode,
                          cudaStream_t st) {
  const int n_tile = cout <= 256 ? cout : (cout % 256 == 0 ? 256 : (cout % 128 == 0 ? 128 : 64));
  for (int n_off = 0; n_off < cout; n_off += n_tile) {
    WgradParams p = {};
    p.t = make_tiling(B, H, W);
    p.taps_x = p.taps_y = taps;
    p.pad_x = p.pad_y = pad;
    p.cin = cin;
    p.bkc = cin <= 32 ? 32 : 64;
    p.n = n_tile;
    p.n_off = n_off;
    p.prologue = pro != nullptr;
    if (pro) { p.scale = pro->scale; p.shift = pro->shift; }
    p.dW = dW;
    p.cout_total = cout;
    p.w_mode = w_mode;
    int rc = launch_conv_wgrad(p, A, ldA, dOut, ldD, st);
    if (rc) return rc;
  }
  return RXB_OK;
}

// the same fold, computed inside the consuming conv kernel's prologue (one launch less per BatchNorm)
static BnPrepArgs fused_prep(const rxb_dn121& n, const BnLayer& bn, const float* sum, const float* sq, float count,
                             int training) {
  BnPrepArgs a = {};
  a.sum = sum; a.sumsq = sq;
  a.gamma = n.params + bn.gamma_off; a.beta = n.params + bn.beta_off;
  a.rmean = n.buffers + bn.rm_off; a.rvar = n.buffers + bn.rv_off;
  a.count = count; a.eps = n.cfg.bn_eps; a.momentum = n.cfg.bn_momentum; a.training = training;
  a.f_scale = bn.fold.scale; a.f_shift = bn.fold.shift; a.f_mean = bn.fold.mean; a.f_rstd = bn.fold.rstd;
  return a;
}

static int prep(const rxb_dn121& n, const BnLayer& bn, const float* sum, const float* sq, float count, int training,
                cudaStream_t st) {
  return bn_prep(sum, sq, count, n.params + bn.gamma_off, n.params + bn.beta_off, n.buffers + bn.rm_off,
                 n.buffers + bn.rv_off, n.cfg.bn_eps, n.cfg.bn_momentum, training, bn.C, bn.fold, st);
}

static bool snake_on() {
  static const bool off = getenv("RXB_DBG_NO_SNAKE") && atoi(getenv("RXB_DBG_NO_SNAKE")) != 0;
  return !off;
}

#define RXB_TRY(expr)          \
  do {                         \
    int rc__ = (expr);         \
    if (rc__) return rc__;     \
  } while (0)

static int sync_weights(rxb_dn121& n, cudaStream_t st) {
  if (!n.jobs_uploaded) return set_error(RXB_ERR_INVALID, "dn121: repack table missing");
  return repack_weights(n.params, n.arena, n.jobs_dev, (int)n.jobs.size(), n.max_job_elems, st);
}

static int forward(rxb_dn121& n, const void* input, int training, cudaStream_t st) {
  const rxb_dn121_config& c = n.cfg;
  RXB_CUDA(cudaMemsetAsync(n.zero_begin, 0, n.zero_bytes, st));
  const bool stats = training != 0;
  // stem: 7x7/2 conv as a 4x4 conv over the 2x2 space-to-depth input (K = 16 taps x 32)
  RXB_TRY(conv_store(n, c.B, n.Hs, n.Ws, static_cast<const __nv_bfloat16*>(input), 32, 32, n.conv0, 64, 4, 2, nullptr,
                     n.S0, 64, 0, stats ? n.s0sum : nullptr, stats ? n.s0sq : nullptr, st));
  RXB_TRY(prep(n, n.bn0, n.s0sum, n.s0sq, (float)((long long)c.B * n.Hs * n.Ws), training, st));
  Block& b0 = n.blocks[0];
  {
    // plans created for inference have no index buffer (no backward): park the indices in block 1's
    // bottleneck scratch, which is not live until the first dense layer runs.
    uint8_t* idx = n.pool_idx ? n.pool_idx : reinterpret_cast<uint8_t*>(b0.layers[0].Y);
    RXB_TRY(stem_bn_relu_maxpool(n.S0, c.B, n.Hs, n.Ws, n.bn0.fold.scale, n.bn0.fold.shift, b0.X, b0.Ctot, idx,
                                 b0.xsum, b0.xsq, st));
  }
  for (int b = 0; b < 4; ++b) {
    Block& blk = n.blocks[b];
    for (auto& L : blk.layers) {
      // both BatchNorm folds are derived inside the consuming conv kernels' prologues
      const BnPrepArgs p1 = fused_prep(n, L.bn1, blk.xsum, blk.xsq, (float)blk.M, training);
      RXB_TRY(conv_store(n, c.B, blk.H, blk.W, blk.X, blk.Ctot, L.Cin, L.c1, kBott, 1, 0, &L.bn1.fold, L.Y, kBott, 0,
                         stats ? L.ysum : nullptr, stats ? L.ysq : nullptr, st, &p1));
      const BnPrepArgs p2 = fused_prep(n, L.bn2, L.ysum, L.ysq, (float)blk.M, training);
      // (the 3x3 walks its tiles backwards: it starts on the end of Y, which the 1x1 wrote last and L2 still holds, and
      // ends on the start of the concat buffer, where the next 1x1 begins - RXB_DBG_NO_SNAKE=1 disables)
      RXB_TRY(conv_store(n, c.B, blk.H, blk.W, L.Y, kBott, kBott, L.c2, kGrowth, 3, 1, &L.bn2.fold, blk.X, blk.Ctot,
                         L.Cin, stats ? blk.xsum : nullptr, stats ? blk.xsq : nullptr, st, &p2, snake_on() ? 1 : 0));
    }
    if (b < 3) {
      Transition& t = n.trans[b];
      Block& nb = n.blocks[b + 1];
      RXB_TRY(prep(n, t.bn, blk.xsum, blk.xsq, (float)blk.M, training, st));
      // avgpool and the 1x1 conv commute: pool first, 4x fewer GEMM rows
      RXB_TRY(transition_pool_fwd(blk.X, blk.Ctot, c.B, blk.H, blk.W, blk.Ctot, t.bn.fold.scale, t.bn.fold.shift, t.P, st));
      RXB_TRY(conv_store(n, c.B, nb.H, nb.W, t.P, blk.Ctot, blk.Ctot, t.conv, nb.C0, 1, 0, nullptr, nb.X, nb.Ctot, 0,
                         stats ? nb.xsum : nullptr, stats ? nb.xsq : nullptr, st));
    }
  }
  Block& b3 = n.blocks[3];
  RXB_TRY(prep(n, n.bn5, b3.xsum, b3.xsq, (float)b3.M, training, st));
  RXB_TRY(final_bn_relu_gap(b3.X, b3.Ctot, c.B, b3.H * b3.W, b3.Ctot, n.bn5.fold.scale, n.bn5.fold.shift, n.feat, st));
  // classifier: logits = feat * Wc^T + bc  (fp32)
  RXB_TRY(sgemm_strided(c.B, c.num_classes, b3.Ctot, n.feat, b3.Ctot, 1, n.params + n.fc_w_off, 1, b3.Ctot,
                        n.params + n.fc_b_off, n.logits, c.num_classes, 1, st));
  return RXB_OK;
}

static int backward_head(rxb_dn121& n, cudaStream_t st) {
  const rxb_dn121_config& c = n.cfg;
  Block& b3 = n.blocks[3];
  const int F = b3.Ctot, NC = c.num_classes;
  // dWc[n,k] = sum_b dlogits[b,n] feat[b,k] ; dbc = column sums ; dfeat = dlogits * Wc
  RXB_TRY(sgemm_strided(NC, F, c.B, n.dlogits, 1, NC, n.feat, F, 1, nullptr, n.grads + n.fc_w_off, F, 1, st));
  RXB_TRY(column_sum(n.dlogits, c.B, NC, NC, n.grads + n.fc_b_off, st));
  RXB_TRY(sgemm_strided(c.B, F, NC, n.dlogits, NC, 1, n.params + n.fc_w_off, F, 1, nullptr, n.dfeat, F, 1, st));
  // norm5 -> relu -> global average pool backward: first writer of G_4
  RXB_TRY(bn_relu_bwd_to_G(1, n.dfeat, b3.X, b3.Ctot, c.B, b3.H, b3.W, b3.Ctot, n.bn5.fold, b3.G, n.bn5.dsum,
                           n.bn5.dsq, st));
  RXB_TRY(bn_bwd_finalize(0, nullptr, nullptr, 0, 0, n.bn5.dsum, n.bn5.dsq, n.bn5.fold, (float)b3.M, b3.Ctot,
                          n.grads + n.bn5.gamma_off, n.grads + n.bn5.beta_off, b3.corrA, b3.corrB, nullptr, nullptr, st));
  return RXB_OK;
}

static int backward_block(rxb_dn121& n, int b, const void* input, cudaStream_t st) {
  const rxb_dn121_config& c = n.cfg;
  Block& blk = n.blocks[b];
  const BnLayer& closing = b < 3 ? n.trans[b].bn : n.bn5;  // its fold covers every channel of the block
  // consecutive kernels walk their rows in alternating directions: each starts on what the previous one touched last
  int dir = 0;
  auto next_dir = [&]() { const int d = dir; if (snake_on()) dir ^= 1; return d; };
  for (int i = (int)blk.layers.size() - 1; i >= 0; --i) {
    DenseLayer& L = blk.layers[i];
    // 3x3 conv: data gradient fused with ReLU/BN2 backward reductions AND the conv's weight gradient (one kernel: the
    // full-halo dZ box and the Y tile serve both); maps too small for the 8x16 tiling keep the separate launch
    const bool fuse3 = conv_dgrad3x3_wgrad_fusable(c.B, blk.H, blk.W);
    static const bool no_tail = getenv("RXB_DBG_NO_BNTAIL") && atoi(getenv("RXB_DBG_NO_BNTAIL")) != 0;
    static const bool no_fixfold = getenv("RXB_DBG_NO_FIXFOLD") && atoi(getenv("RXB_DBG_NO_FIXFOLD")) != 0;
    // exact gradient of this layer's 32 output channels, dZ = G - corrA - xhat*corrB: derived by the fused 3x3 kernel
    // itself from the concat buffers on load (FixupArgs), or written densely by grad_fixup for the separate kernels
    const bool fixfold = fuse3 && !no_tail && !no_fixfold;
    FixupArgs fx = {};
    fx.G = blk.G; fx.X = blk.X; fx.ld = blk.Ctot; fx.c0 = L.Cin;
    fx.mean = closing.fold.mean; fx.rstd = closing.fold.rstd; fx.corrA = blk.corrA; fx.corrB = blk.corrB;
    if (!fixfold)
      RXB_TRY(grad_fixup(blk.G, blk.X, blk.Ctot, blk.M, L.Cin, kGrowth, closing.fold.mean, closing.fold.rstd, blk.corrA,
                         blk.corrB, n.dZ, st));
    if (!fuse3)
      RXB_TRY(conv_wgrad_any(c.B, blk.H, blk.W, L.Y, kBott, kBott, 3, 1, &L.bn2.fold, n.dZ, kGrowth, kGrowth,
                             n.grads + L.c2.w_off, 0, st));
    if (fuse3 && !no_tail) {
      // the kernel leaves sum(dy) and W.dW (every CTA's share) in bn2.dsum / bn2.dsq; bn_bwd_apply derives the means
      BnTailArgs t2 = {};
      t2.mode = 2;
      RXB_TRY(conv_dgrad_bn(n, c.B, blk.H, blk.W, n.dZ, kGrowth, kGrowth, L.c2, kBott, 3, 1, L.Y, kBott, L.bn2, OUT_DY,
                            n.dy2, kBott, st, n.grads + L.c2.w_off, &t2, fixfold ? &fx : nullptr, next_dir()));
      const BnRawSums raw = {n.params + L.bn2.gamma_off, n.params + L.bn2.beta_off, n.grads + L.bn2.gamma_off,
                             n.grads + L.bn2.beta_off, 1.f / (float)blk.M};
      RXB_TRY(bn_bwd_apply(n.dy2, L.Y, blk.M, kBott, L.bn2.fold, L.bn2.dsum, L.bn2.dsq, st, nullptr, &raw, next_dir()));  // dy2 := dY
    } else {
      RXB_TRY(conv_dgrad_bn(n, c.B, blk.H, blk.W, n.dZ, kGrowth, kGrowth, L.c2, kBott, 3, 1, L.Y, kBott, L.bn2, OUT_DY,
                            n.dy2, kBott, st, fuse3 ? n.grads + L.c2.w_off : nullptr));
      RXB_TRY(bn_bwd_finalize(1, n.params + L.c2.w_off, n.grads + L.c2.w_off, kGrowth, 9, L.bn2.dsum, L.bn2.dsq,
                              L.bn2.fold, (float)blk.M, kBott, n.grads + L.bn2.gamma_off,
                              n.grads + L.bn2.beta_off, nullptr, nullptr, n.params + L.bn2.gamma_off,
                              n.params + L.bn2.beta_off, st));
      RXB_TRY(bn_bwd_apply(n.dy2, L.Y, blk.M, kBott, L.bn2.fold, L.bn2.dsum, L.bn2.dsq, st));  // dy2 := dY
    }
    // 1x1 conv: ONE kernel for its data gradient (into the concat gradient) and its weight gradient - both contract
    // the same dY and X tiles (RXB_DBG_NO_WGFUSE=1: the separate weight-gradient launch, for comparison)
    static const bool no_wgfuse = getenv("RXB_DBG_NO_WGFUSE") && atoi(getenv("RXB_DBG_NO_WGFUSE")) != 0;
    if (no_wgfuse)
      RXB_TRY(conv_wgrad_any(c.B, blk.H, blk.W, blk.X, blk.Ctot, L.Cin, 1, 0, &L.bn1.fold, n.dy2, kBott, kBott,
                             n.grads + L.c1.w_off, 0, st));
    if (!no_wgfuse && !no_tail) {
      // ... and bn1's backward reductions (dgamma, dbeta, the lazy correction terms of the concat gradient): every
      // CTA adds its share in its tail
      BnTailArgs t1 = {};
      t1.mode = 1;
      t1.W = n.params + L.c1.w_off;
      t1.mean = L.bn1.fold.mean; t1.rstd = L.bn1.fold.rstd;
      t1.inv_count = 1.f / (float)blk.M;
      t1.dgamma = n.grads + L.bn1.gamma_off; t1.dbeta = n.grads + L.bn1.beta_off;
      t1.corrA = blk.corrA; t1.corrB = blk.corrB;
      RXB_TRY(conv_dgrad_bn(n, c.B, blk.H, blk.W, n.dy2, kBott, kBott, L.c1, L.Cin, 1, 0, blk.X, blk.Ctot, L.bn1,
                            OUT_G_ACCUM, blk.G, blk.Ctot, st, n.grads + L.c1.w_off, &t1, nullptr, next_dir()));
    } else {
      RXB_TRY(conv_dgrad_bn(n, c.B, blk.H, blk.W, n.dy2, kBott, kBott, L.c1, L.Cin, 1, 0, blk.X, blk.Ctot, L.bn1,
                            OUT_G_ACCUM, blk.G, blk.Ctot, st, no_wgfuse ? nullptr : n.grads + L.c1.w_off));
      RXB_TRY(bn_bwd_finalize(0, n.params + L.c1.w_off, n.grads + L.c1.w_off, kBott, 1, L.bn1.dsum, L.bn1.dsq,
                              L.bn1.fold, (float)blk.M, L.Cin, n.grads + L.bn1.gamma_off,
                              n.grads + L.bn1.beta_off, blk.corrA, blk.corrB, n.params + L.bn1.gamma_off,
                              n.params + L.bn1.beta_off, st));
    }
  }
  // exact gradient of the block's input channels
  RXB_TRY(grad_fixup(blk.G, blk.X, blk.Ctot, blk.M, 0, blk.C0, closing.fold.mean, closing.fold.rstd, blk.corrA,
                     blk.corrB, n.dX0, st));
  if (b > 0) {
    Transition& t = n.trans[b - 1];
    Block& pb = n.blocks[b - 1];
    RXB_TRY(conv_wgrad_any(c.B, blk.H, blk.W, t.P, pb.Ctot, pb.Ctot, 1, 0, nullptr, n.dX0, blk.C0, blk.C0,
                           n.grads + t.conv.w_off, 0, st));
    {
      GemmParams p = {};
      p.B = c.B; p.H = blk.H; p.W = blk.W;
      p.n_total = pb.Ctot;
      p.taps_x = p.taps_y = 1;
      p.cin = blk.C0;
      p.epi_mode = EPI_STORE;
      RXB_TRY(launch_conv_gemm(p, n.dX0, blk.C0, n.arena + t.conv.dgrad_off, n.dP, pb.Ctot, 0, nullptr, 0, 64, false, st));
    }
    RXB_TRY(bn_relu_bwd_to_G(0, n.dP, pb.X, pb.Ctot, c.B, pb.H, pb.W, pb.Ctot, t.bn.fold, pb.G, t.bn.dsum, t.bn.dsq, st));
    RXB_TRY(bn_bwd_finalize(0, nullptr, nullptr, 0, 0, t.bn.dsum, t.bn.dsq, t.bn.fold, (float)pb.M, pb.Ctot,
                            n.grads + t.bn.gamma_off,
                            n.grads + t.bn.beta_off, pb.corrA, pb.corrB, nullptr, nullptr, st));
  } else {
    RXB_TRY(stem_pool_bwd(n.dX0, n.pool_idx, n.S0, c.B, n.Hs, n.Ws, n.bn0.fold, n.dy0, n.bn0.dsum, n.bn0.dsq, st));
    const float cnt = (float)((long long)c.B * n.Hs * n.Ws);
    RXB_TRY(bn_bwd_finalize(1, nullptr, nullptr, 0, 0, n.bn0.dsum, n.bn0.dsq, n.bn0.fold, cnt, 64,
                            n.grads + n.bn0.gamma_off,
                            n.grads + n.bn0.beta_off, nullptr, nullptr, nullptr, nullptr, st));
    RXB_TRY(bn_bwd_apply(n.dy0, n.S0, (long long)c.B * n.Hs * n.Ws, 64, n.bn0.fold, n.bn0.dsum, n.bn0.dsq, st));
    RXB_TRY(conv_wgrad_any(c.B, n.Hs, n.Ws, static_cast<const __nv_bfloat16*>(input), 32, 32, 4, 2, nullptr, n.dy0, 64,
                           64, n.grads + n.conv0.w_off, 1, st));
  }
  return RXB_OK;
}

}  // namespace rxb

extern "C" {

int64_t rxb_dn121_param_count(const rxb_dn121_config* cfg) {
  if (rxb::check_cfg(cfg)) return -1;
  rxb_dn121 n;
  n.cfg = *cfg;
  rxb::plan(n, nullptr);
  return n.n_params;
}

int64_t rxb_dn121_buffer_count(const rxb_dn121_config* cfg) {
  if (rxb::check_cfg(cfg)) return -1;
  rxb_dn121 n;
  n.cfg = *cfg;
  rxb::plan(n, nullptr);
  return n.n_buffers;
}

size_t rxb_dn121_workspace_bytes(const rxb_dn121_config* cfg, int training) {
  if (rxb::check_cfg(cfg)) return 0;
  rxb_dn121 n;
  n.cfg = *cfg;
  n.training = training;
  return rxb::plan(n, nullptr);
}

int rxb_dn121_create(const rxb_dn121_config* cfg, float* params, float* grads, float* momentum, float* buffers,
                     void* workspace, size_t workspace_bytes, int training, rxb_dn121** out) {
  using namespace rxb;
  int rc = check_cfg(cfg);
  if (rc) return rc;
  RXB_CHECK_ARG(params && buffers && workspace && out, "rxb_dn121_create: null pointer");
  RXB_CHECK_ARG(!training || (grads && momentum), "rxb_dn121_create: training needs grads and momentum");
  RXB_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "rxb_dn121_create: workspace must be 256B aligned");
  rc = rxb_check_device();
  if (rc) return rc;
  rxb_dn121* n = new rxb_dn121();
  n->cfg = *cfg;
  n->training = training;
  n->params = params; n->grads = grads; n->momentum = momentum; n->buffers = buffers;
  n->ws = static_cast<uint8_t*>(workspace);
  const size_t need = plan(*n, n->ws);
  if (need > workspace_bytes) {
    delete n;
    return set_error(RXB_ERR_INVALID, "rxb_dn121_create: workspace %zu B < required %zu B", workspace_bytes, need);
  }
  n->ws_bytes = workspace_bytes;
  // the repack table is uploaded synchronously here so later calls are pure stream work (graph-capturable)
  cudaError_t ce = cudaMemcpy(n->jobs_dev, n->jobs.data(), n->jobs.size() * sizeof(RepackJob), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) {
    delete n;
    return set_error(RXB_ERR_CUDA, "rxb_dn121_create: upload of the repack table failed: %s", cudaGetErrorString(ce));
  }
  n->jobs_uploaded = true;
  *out = n;
  return RXB_OK;
}

void rxb_dn121_destroy(rxb_dn121* net) { delete net; }

int rxb_dn121_sync_weights(rxb_dn121* net, rxb_stream_t stream) {
  RXB_CHECK_ARG(net, "rxb_dn121_sync_weights: null");
  return rxb::sync_weights(*net, rxb::as_stream(stream));
}

int rxb_dn121_forward(rxb_dn121* net, const void* input_s2d, float* logits_out, int training, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(net && input_s2d, "rxb_dn121_forward: null pointer");
  cudaStream_t st = as_stream(stream);
  RXB_TRY(forward(*net, input_s2d, training, st));
  if (logits_out)
    RXB_CUDA(cudaMemcpyAsync(logits_out, net->logits, sizeof(float) * net->cfg.B * net->cfg.num_classes,
                             cudaMemcpyDeviceToDevice, st));
  return RXB_OK;
}

int rxb_dn121_num_phases(void) { return 5; }

int rxb_dn121_phase_grad_range(const rxb_dn121* net, int phase, int64_t* begin, int64_t* end) {
  RXB_CHECK_ARG(net && begin && end && phase >= 0 && phase < 5, "rxb_dn121_phase_grad_range: bad argument");
  // phase 0: head + norm5 ; phase p (1..4): dense block 5-p and the transition (or stem) in front of it
  const int64_t bounds[6] = {net->n_params,
                             net->bn5.gamma_off,
                             net->trans[2].bn.gamma_off,
                             net->trans[1].bn.gamma_off,
                             net->trans[0].bn.gamma_off,
                             0};
  *end = bounds[phase];
  *begin = bounds[phase + 1];
  return RXB_OK;
}

int rxb_dn121_train_step(rxb_dn121* net, const void* input_s2d, const int64_t* target, int global_batch,
                         float* loss_out, int phase, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(net && input_s2d && target, "rxb_dn121_train_step: null pointer");
  RXB_CHECK_ARG(net->training, "rxb_dn121_train_step: plan was created for inference");
  RXB_CHECK_ARG(phase >= -1 && phase < 5 && global_batch >= 1, "rxb_dn121_train_step: bad phase/global_batch");
  cudaStream_t st = as_stream(stream);
  rxb_dn121& n = *net;
  if (phase == -1 || phase == 0) {
    RXB_CUDA(cudaMemsetAsync(n.grads, 0, sizeof(float) * n.n_params, st));
    RXB_TRY(forward(n, input_s2d, 1, st));
    RXB_TRY(rxb_softmax_ce(n.logits, n.cfg.num_classes, target, n.cfg.B, n.cfg.num_classes, n.loss_rows, n.dlogits,
                           1.f / (float)global_batch, stream));
    if (loss_out) RXB_TRY(sum_scale(n.loss_rows, n.cfg.B, 1.f / (float)global_batch, loss_out, st));
    RXB_TRY(backward_head(n, st));
  }
  for (int ph = 1; ph <= 4; ++ph)
    if (phase == -1 || phase == ph) RXB_TRY(backward_block(n, 4 - ph, input_s2d, st));
  return RXB_OK;
}

int rxb_dn121_sgd(rxb_dn121* net, float lr, float mu, float wd, int nesterov, float grad_scale, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(net && net->training, "rxb_dn121_sgd: needs a training plan");
  RXB_TRY(rxb_sgd_step(net->params, net->grads, net->momentum, net->n_params, lr, mu, wd, nesterov, grad_scale, stream));
  return sync_weights(*net, as_stream(stream));
}

}  // extern "C"
