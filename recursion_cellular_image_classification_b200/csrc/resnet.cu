// resnet.cu — the reference's REAL model on the device: TwoSitesNN (reference cell_classifier/models.py:7-57) =
// torchvision ResNet-50 trunk with the 6-channel stem (:16-29, fc -> Identity), per-sample feature means of the
// image / negative-control / positive-control thirds concatenated (:44-53) and the
// BatchNorm1d -> Dropout -> Linear(6144,1024) -> ReLU -> BatchNorm1d -> Dropout -> Linear(1024,1108) head (:31-39).
//
// Evaluation (BatchNorm from running statistics, Dropout = identity) is what reference test.py:23-27 runs for every
// test batch, and what a checkpoint trained with the reference (models/best_model_<id>.pth, `module.`-prefixed,
// main.py:147) needs to be served on a B200.  Training (rxb_rn50_train_step / rxb_rn50_sgd) is the reference's train
// step (train.py:37,44; main.py:89-93): batch statistics per BatchNorm over this rank's images (DataParallel semantics),
// Dropout with caller-supplied masks, CrossEntropy, the whole backward, nesterov SGD on the flat buffer - optionally on
// the MLP head alone, which is the reference's two-epoch freeze of a pretrained trunk (train.py:46-67).
// Parameters and buffers are flat fp32 arrays in the reference model's own named_parameters() / named_buffers() order.
//
// Backward of a bottleneck (out = relu(bn3(c3) + idn), c3 = conv3(relu(bn2(c2))), c2 = conv2(relu(bn1(c1))), c1 =
// conv1(x)), given D = dL/d(out) accumulated from the next block:
//   du = D*[out>0] with the bn3 (and downsample-bn) reductions             (one elementwise pass, in place)
//   dc3 = bn3 backward(du)  ->  conv3 weight gradient, conv3 data gradient fused with the ReLU/bn2 backward
//   reductions (the dgrad epilogue of conv_gemm.cu; sum dy*x from W.dW)    ->  dc2 = bn2 backward  ->  conv2 likewise
//   ->  dc1  ->  conv1 weight gradient; conv1's data gradient is ADDED to du (identity blocks: the same buffer becomes
//   dL/d(x), the L2 reduce-add epilogue) or written next to the downsample branch's gradient.
//
// Every convolution is the tcgen05 implicit-GEMM kernel of conv_gemm.cu:
//   * 1x1 stride 1  : plain GEMM over NHWC pixels; the BatchNorm+ReLU in front of conv2/conv3 is the A-operand prologue
//   * 3x3 stride 1  : taps as TMA box shifts / descriptor offsets, BatchNorm+ReLU prologue
//   * 3x3 stride 2  : the stride-1 2x2-tap convolution over the 2x2 space-to-depth form of relu(bn1(.)) (one
//                     elementwise pass writes that form; weights repacked by RP_3x3S2_FWD) — K = 16*Cin for 9*Cin real
//   * 1x1 stride 2  : (torchvision's downsample branch) the 1x1 GEMM over the subsampled input
//   * the stem      : the same 7x7/2 -> 4x4-tap space-to-depth convolution + BN/ReLU/maxpool kernels as densenet.cu
// The residual join relu(bn3(c3) + identity) is one elementwise pass per bottleneck (resnet_ops.cu).
#include <algorithm>
#include <vector>
#include "conv_gemm.cuh"
#include "elementwise.cuh"
#include "resnet_ops.cuh"

namespace rxb {

struct RnBn {
  int C = 0;
  long long gamma_off = 0, beta_off = 0, rm_off = 0, rv_off = 0, fold_off = 0;
};
struct RnConv {
  long long w_off = 0, fwd_off = 0, dgrad_off = 0;
  int cin = 0, cout = 0, k = 1, stride = 1;
};
struct RnBlock {
  int cin = 0, width = 0, stride = 1, down = 0;
  int Hin = 0, Win = 0, Hout = 0, Wout = 0;
  RnConv c1, c2, c3, cd;
  RnBn b1, b2, b3, bd;
  // training: activations kept for backward
  __nv_bfloat16 *sC1 = nullptr, *sC2 = nullptr, *sC3 = nullptr, *sCD = nullptr, *sS2 = nullptr, *sXS = nullptr, *sOut = nullptr;
};

}  // namespace rxb

struct rxb_rn50 {
  rxb_rn50_config cfg;
  int training = 0;
  float *params = nullptr, *buffers = nullptr, *grads = nullptr, *momentum = nullptr;
  long long head_off = 0;   // first mlp.* parameter
  // training state: per-BatchNorm arrays indexed by RnBn::fold_off (fold_scale / fold_shift double as the fold's scale / shift)
  float *fold_mean = nullptr, *fold_rstd = nullptr, *st_sum = nullptr, *st_sq = nullptr, *d_sum = nullptr, *d_sq = nullptr;
  float *ones = nullptr, *big = nullptr, *scratch_c = nullptr, *wg_scratch = nullptr;
  long long wg_scratch_elems = 0;
  uint8_t* zero_begin = nullptr;
  size_t zero_bytes = 0;
  __nv_bfloat16 *X0 = nullptr, *dy0 = nullptr, *Da = nullptr, *Db = nullptr, *DC3 = nullptr, *DCD = nullptr, *DZ2 = nullptr,
                *DZ1 = nullptr, *DS2 = nullptr, *DXS = nullptr;
  float *y0m = nullptr, *y1 = nullptr, *y1m = nullptr, *sv_mean0 = nullptr, *sv_rstd0 = nullptr, *sv_mean1 = nullptr,
        *sv_rstd1 = nullptr, *dlogits = nullptr, *dy1 = nullptr, *dh1 = nullptr, *dy0v = nullptr, *dcat = nullptr, *dfeat = nullptr,
        *loss_rows = nullptr;
  long long n_params = 0, n_buffers = 0, n_fold = 0;
  int Bi = 0;      // images through the trunk = B * G
  int Hs = 0, Ws = 0, H1 = 0, W1 = 0;
  rxb::RnConv conv0;
  rxb::RnBn bn0, bn_m0, bn_m4;
  std::vector<rxb::RnBlock> blocks;
  long long fc1_w = 0, fc1_b = 0, fc2_w = 0, fc2_b = 0;
  int feat_dim = 2048;
  // workspace
  __nv_bfloat16 *arena = nullptr, *S0 = nullptr, *Xa = nullptr, *Xb = nullptr, *C1 = nullptr, *C2 = nullptr, *C3 = nullptr,
                *CD = nullptr, *S2 = nullptr, *XS = nullptr;
  uint8_t* pool_idx = nullptr;
  float *fold_scale = nullptr, *fold_shift = nullptr, *scratch_sum = nullptr, *feat = nullptr, *cat = nullptr, *h0 = nullptr,
        *h1 = nullptr, *h2 = nullptr, *logits = nullptr;
  rxb::RepackJob* jobs_dev = nullptr;
  rxb::BnFoldJob* fold_jobs_dev = nullptr;
  std::vector<rxb::RepackJob> jobs;
  std::vector<rxb::BnFoldJob> fold_jobs;
  long long arena_elems = 0, max_job_elems = 0;
  int max_bn_c = 0;
};

namespace rxb {

namespace {

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

const int kLayers[4] = {3, 4, 6, 3};
const int kWidths[4] = {64, 128, 256, 512};

#define RXB_TRY(expr)          \
  do {                         \
    int rc__ = (expr);         \
    if (rc__) return rc__;     \
  } while (0)

size_t plan(rxb_rn50& n, uint8_t* ws) {
  const rxb_rn50_config& c = n.cfg;
  long long poff = 0, boff = 0, aoff = 0, foff = 0;
  n.jobs.clear();
  n.fold_jobs.clear();
  n.blocks.clear();
  n.max_job_elems = 0;
  n.max_bn_c = 0;
  auto add_job = [&](long long src, int type, int N, int K, long long elems) {
    RepackJob j;
    j.src_off = src; j.dst_off = aoff; j.type = type; j.N = N; j.K = K; j.pad = 0;
    n.jobs.push_back(j);
    const long long at = aoff;
    aoff += (elems + 63) & ~63ll;
    if (elems > n.max_job_elems) n.max_job_elems = elems;
    return at;
  };
  auto plan_conv = [&](RnConv& cv, int cin, int cout, int k, int stride) {
    cv.cin = cin; cv.cout = cout; cv.k = k; cv.stride = stride;
    cv.w_off = poff;
    poff += (long long)cout * cin * k * k;
    if (k == 1) {
      cv.fwd_off = add_job(cv.w_off, RP_1x1_FWD, cout, cin, (long long)cout * cin);
      if (n.training) cv.dgrad_off = add_job(cv.w_off, RP_1x1_DGRAD, cout, cin, (long long)cout * cin);
    } else if (stride == 1) {
      cv.fwd_off = add_job(cv.w_off, RP_3x3_FWD, cout, cin, 9ll * cout * cin);
      if (n.training) cv.dgrad_off = add_job(cv.w_off, RP_3x3_DGRAD, cout, cin, 9ll * cout * cin);
    } else {
      cv.fwd_off = add_job(cv.w_off, RP_3x3S2_FWD, cout, cin, 16ll * cout * cin);
      if (n.training) cv.dgrad_off = add_job(cv.w_off, RP_3x3S2_DGRAD, cout, cin, 16ll * cout * cin);
    }
  };
  auto plan_bn = [&](RnBn& bn, int C) {
    bn.C = C;
    bn.gamma_off = poff; poff += C;
    bn.beta_off = poff; poff += C;
    if (C > n.max_bn_c) n.max_bn_c = C;
  };
  // buffers follow module registration order, which differs from the parameter order inside a bottleneck
  // (bn1, bn2, bn3, downsample.1), so they are assigned in a second pass
  auto plan_bn_buf = [&](RnBn& bn) {
    bn.rm_off = boff; boff += bn.C;
    bn.rv_off = boff; boff += bn.C;
    bn.fold_off = foff; foff += (bn.C + 63) & ~63;
    BnFoldJob j;
    j.gamma_off = bn.gamma_off; j.beta_off = bn.beta_off; j.rm_off = bn.rm_off; j.rv_off = bn.rv_off;
    j.fold_off = bn.fold_off; j.C = bn.C; j.pad = 0;
    n.fold_jobs.push_back(j);
  };
  // ---- parameters in the reference model's named_parameters() order (base_nn.*, then mlp.*)
  n.conv0.cin = 32; n.conv0.cout = 64; n.conv0.k = 4; n.conv0.stride = 1;
  n.conv0.w_off = poff; poff += 64 * 6 * 49;
  n.conv0.fwd_off = add_job(n.conv0.w_off, RP_STEM_FWD, 64, 32, 16ll * 64 * 32);
  plan_bn(n.bn0, 64);
  n.Bi = c.B * c.G;
  n.Hs = c.H / 2; n.Ws = c.W / 2;
  n.H1 = n.Hs / 2; n.W1 = n.Ws / 2;
  int cin = 64, H = n.H1, W = n.W1;
  for (int l = 0; l < 4; ++l) {
    for (int i = 0; i < kLayers[l]; ++i) {
      RnBlock b;
      b.cin = cin; b.width = kWidths[l];
      b.stride = (i == 0 && l > 0) ? 2 : 1;
      b.down = i == 0;
      b.Hin = H; b.Win = W;
      b.Hout = b.stride == 2 ? (H + 1) / 2 : H;      // 3x3/s2/p1 and 1x1/s2: floor((H-1)/2)+1
      b.Wout = b.stride == 2 ? (W + 1) / 2 : W;
      plan_conv(b.c1, cin, b.width, 1, 1);
      plan_bn(b.b1, b.width);
      plan_conv(b.c2, b.width, b.width, 3, b.stride);      // torchvision v1.5: the stride sits on the 3x3
      plan_bn(b.b2, b.width);
      plan_conv(b.c3, b.width, 4 * b.width, 1, 1);
      plan_bn(b.b3, 4 * b.width);
      if (b.down) {
        plan_conv(b.cd, cin, 4 * b.width, 1, b.stride);
        plan_bn(b.bd, 4 * b.width);
      }
      n.blocks.push_back(b);
      cin = 4 * b.width;
      H = b.Hout; W = b.Wout;
    }
  }
  n.feat_dim = cin;   // 2048
  n.head_off = poff;
  plan_bn(n.bn_m0, 3 * n.feat_dim);
  n.fc1_w = poff; poff += (long long)c.size_features * 3 * n.feat_dim;
  n.fc1_b = poff; poff += c.size_features;
  plan_bn(n.bn_m4, c.size_features);
  n.fc2_w = poff; poff += (long long)c.num_classes * c.size_features;
  n.fc2_b = poff; poff += c.num_classes;
  n.n_params = poff;
  // ---- buffers (running_mean, running_var per BatchNorm, module order)
  plan_bn_buf(n.bn0);
  for (auto& b : n.blocks) {
    plan_bn_buf(b.b1);
    plan_bn_buf(b.b2);
    plan_bn_buf(b.b3);
    if (b.down) plan_bn_buf(b.bd);
  }
  plan_bn_buf(n.bn_m0);
  plan_bn_buf(n.bn_m4);
  n.n_buffers = boff;
  n.n_fold = foff;
  n.arena_elems = aoff;

  // ---- workspace
  Bump bp(ws);
  n.arena = bp.take<__nv_bfloat16>(aoff);
  n.jobs_dev = bp.take<RepackJob>(n.jobs.size());
  n.fold_jobs_dev = bp.take<BnFoldJob>(n.fold_jobs.size());
  n.fold_scale = bp.take<float>(foff);
  n.fold_shift = bp.take<float>(foff);
  n.scratch_sum = bp.take<float>(256);
  const long long Bi = n.Bi;
  const int tr = n.training;
  if (tr) {
    n.fold_mean = bp.take<float>(foff);
    n.fold_rstd = bp.take<float>(foff);
    n.ones = bp.take<float>(2048);
    n.big = bp.take<float>(2048);
    bp.take<uint8_t>(0);
    const size_t z0 = (bp.off + 255) & ~size_t(255);
    n.st_sum = bp.take<float>(foff);
    n.st_sq = bp.take<float>(foff);
    n.d_sum = bp.take<float>(foff);
    n.d_sq = bp.take<float>(foff);
    n.scratch_c = bp.take<float>(2 * 2048);
    bp.take<uint8_t>(0);
    const size_t z1 = (bp.off + 255) & ~size_t(255);
    n.zero_begin = ws ? ws + z0 : nullptr;
    n.zero_bytes = z1 - z0;
    bp.off = z1;
  }
  n.S0 = bp.take<__nv_bfloat16>(Bi * n.Hs * n.Ws * 64);
  n.pool_idx = bp.take<uint8_t>(Bi * n.H1 * n.W1 * 64);
  long long mx_x = Bi * n.H1 * n.W1 * 64, mx_c1 = 0, mx_c2 = 0, mx_c3 = 0, mx_s2 = 0, mx_xs = 0;
  for (auto& b : n.blocks) {
    const long long Min = Bi * b.Hin * b.Win, Mout = Bi * b.Hout * b.Wout;
    mx_x = std::max(mx_x, Mout * 4 * b.width);
    mx_c1 = std::max(mx_c1, Min * b.width);
    mx_c2 = std::max(mx_c2, Mout * b.width);
    mx_c3 = std::max(mx_c3, Mout * 4 * b.width);
    if (b.stride == 2) {
      mx_s2 = std::max(mx_s2, Mout * 4 * b.width);
      mx_xs = std::max(mx_xs, Mout * b.cin);
    }
  }
  if (!tr) {
    n.Xa = bp.take<__nv_bfloat16>(mx_x);
    n.Xb = bp.take<__nv_bfloat16>(mx_x);
    n.C1 = bp.take<__nv_bfloat16>(mx_c1);
    n.C2 = bp.take<__nv_bfloat16>(mx_c2);
    n.C3 = bp.take<__nv_bfloat16>(mx_c3);
    n.CD = bp.take<__nv_bfloat16>(mx_c3);
    n.S2 = bp.take<__nv_bfloat16>(mx_s2);
    n.XS = bp.take<__nv_bfloat16>(mx_xs);
  } else {
    // every activation a backward kernel reads is kept (bf16), per bottleneck
    n.X0 = bp.take<__nv_bfloat16>(Bi * n.H1 * n.W1 * 64);
    long long mx_in = Bi * n.H1 * n.W1 * 64;
    for (auto& b : n.blocks) {
      const long long Min = Bi * b.Hin * b.Win, Mout = Bi * b.Hout * b.Wout;
      b.sC1 = bp.take<__nv_bfloat16>(Min * b.width);
      b.sC2 = bp.take<__nv_bfloat16>(Mout * b.width);
      b.sC3 = bp.take<__nv_bfloat16>(Mout * 4 * b.width);
      b.sOut = bp.take<__nv_bfloat16>(Mout * 4 * b.width);
      if (b.down) b.sCD = bp.take<__nv_bfloat16>(Mout * 4 * b.width);
      if (b.stride == 2) {
        b.sS2 = bp.take<__nv_bfloat16>(Mout * 4 * b.width);
        b.sXS = bp.take<__nv_bfloat16>(Mout * b.cin);
      }
      mx_in = std::max(mx_in, Min * b.cin);
    }
    long long mx_wg = 0;
    for (auto& b : n.blocks) mx_wg = std::max(mx_wg, (b.stride == 2 ? 16ll : 9ll) * b.width * b.width);
    n.wg_scratch_elems = mx_wg;
    n.wg_scratch = bp.take<float>(mx_wg);
    n.dy0 = bp.take<__nv_bfloat16>(Bi * n.Hs * n.Ws * 64);
    n.Da = bp.take<__nv_bfloat16>(std::max(mx_x, mx_in));
    n.Db = bp.take<__nv_bfloat16>(std::max(mx_x, mx_in));
    n.DC3 = bp.take<__nv_bfloat16>(mx_c3);
    n.DCD = bp.take<__nv_bfloat16>(mx_c3);
    n.DZ2 = bp.take<__nv_bfloat16>(mx_c2);
    n.DZ1 = bp.take<__nv_bfloat16>(mx_c1);
    n.DS2 = bp.take<__nv_bfloat16>(mx_s2);
    n.DXS = bp.take<__nv_bfloat16>(mx_xs);
    const long long F3 = 3ll * n.feat_dim, SF = c.size_features, NC = c.num_classes;
    n.y0m = bp.take<float>(c.B * F3);
    n.y1 = bp.take<float>(c.B * SF);
    n.y1m = bp.take<float>(c.B * SF);
    n.sv_mean0 = bp.take<float>(F3);
    n.sv_rstd0 = bp.take<float>(F3);
    n.sv_mean1 = bp.take<float>(SF);
    n.sv_rstd1 = bp.take<float>(SF);
    n.dlogits = bp.take<float>(c.B * NC);
    n.dy1 = bp.take<float>(c.B * SF);
    n.dh1 = bp.take<float>(c.B * SF);
    n.dy0v = bp.take<float>(c.B * F3);
    n.dcat = bp.take<float>(c.B * F3);
    n.dfeat = bp.take<float>(Bi * n.feat_dim);
    n.loss_rows = bp.take<float>(c.B);
  }
  n.feat = bp.take<float>(Bi * n.feat_dim);
  n.cat = bp.take<float>((long long)c.B * 3 * n.feat_dim);
  n.h0 = bp.take<float>((long long)c.B * 3 * n.feat_dim);
  n.h1 = bp.take<float>((long long)c.B * c.size_features);
  n.h2 = bp.take<float>((long long)c.B * c.size_features);
  n.logits = bp.take<float>((long long)c.B * c.num_classes);
  bp.take<uint8_t>(0);
  return ((bp.off + 255) & ~size_t(255)) + 256;
}

int check_cfg(const rxb_rn50_config* c) {
  RXB_CHECK_ARG(c != nullptr, "rn50: null config");
  RXB_CHECK_ARG(c->B >= 1 && c->B <= 4096, "rn50: bad batch %d", c->B);
  RXB_CHECK_ARG(c->G >= 3 && c->G % 3 == 0, "rn50: G=%d images per sample must be a positive multiple of 3 (image / negative / positive thirds)", c->G);
  RXB_CHECK_ARG(c->H >= 32 && c->W >= 32 && c->H % 4 == 0 && c->W % 4 == 0, "rn50: H, W must be multiples of 4 (>= 32)");
  RXB_CHECK_ARG(c->num_classes >= 1 && c->size_features >= 8, "rn50: bad head sizes");
  return RXB_OK;
}

// out[.., 0:cout] = conv(A) with an optional BatchNorm+ReLU prologue on A (fold arrays of the BatchNorm in front)
int conv(const rxb_rn50& n, int B, int H, int W, const __nv_bfloat16* A, int cin, const RnConv& cv, int taps, int pad,
         const RnBn* pro, __nv_bfloat16* out, cudaStream_t st) {
  GemmParams p = {};
  p.B = B; p.H = H; p.W = W;
  p.n_total = cv.cout;
  p.taps_x = p.taps_y = taps;
  p.pad_x = p.pad_y = pad;
  p.cin = cin;
  p.epi_mode = EPI_STORE;
  if (pro) { p.scale = n.fold_scale + pro->fold_off; p.shift = n.fold_shift + pro->fold_off; }
  return launch_conv_gemm(p, A, cin, n.arena + cv.fwd_off, out, cv.cout, 0, nullptr, 0, cin <= 32 ? 32 : 64, pro != nullptr, st);
}

int forward_eval(rxb_rn50& n, const void* input, cudaStream_t st) {
  const rxb_rn50_config& c = n.cfg;
  const int Bi = n.Bi;
  RXB_TRY(bn_fold_eval_all(n.params, n.buffers, n.fold_jobs_dev, (int)n.fold_jobs.size(), n.max_bn_c, c.bn_eps, n.fold_scale,
                           n.fold_shift, st));
  // stem: 7x7/2 as a 4x4-tap conv over the 2x2 space-to-depth input, then BN + ReLU + maxpool 3x3/2
  RXB_TRY(conv(n, Bi, n.Hs, n.Ws, static_cast<const __nv_bfloat16*>(input), 32, n.conv0, 4, 2, nullptr, n.S0, st));
  RXB_CUDA(cudaMemsetAsync(n.scratch_sum, 0, 256 * sizeof(float), st));
  RXB_TRY(stem_bn_relu_maxpool(n.S0, Bi, n.Hs, n.Ws, n.fold_scale + n.bn0.fold_off, n.fold_shift + n.bn0.fold_off, n.Xa, 64,
                               n.pool_idx, n.scratch_sum, n.scratch_sum + 64, st));
  __nv_bfloat16 *x = n.Xa, *y = n.Xb;
  for (auto& b : n.blocks) {
    const long long Mout = (long long)Bi * b.Hout * b.Wout;
    RXB_TRY(conv(n, Bi, b.Hin, b.Win, x, b.cin, b.c1, 1, 0, nullptr, n.C1, st));
    if (b.stride == 1) {
      RXB_TRY(conv(n, Bi, b.Hin, b.Win, n.C1, b.width, b.c2, 3, 1, &b.b1, n.C2, st));
    } else {
      RXB_TRY(s2d_bn_relu(n.C1, Bi, b.Hin, b.Win, b.width, n.fold_scale + b.b1.fold_off, n.fold_shift + b.b1.fold_off, n.S2, st));
      RXB_TRY(conv(n, Bi, b.Hout, b.Wout, n.S2, 4 * b.width, b.c2, 2, 1, nullptr, n.C2, st));
    }
    RXB_TRY(conv(n, Bi, b.Hout, b.Wout, n.C2, b.width, b.c3, 1, 0, &b.b2, n.C3, st));
    const __nv_bfloat16* idn = x;
    if (b.down) {
      const __nv_bfloat16* xs = x;
      if (b.stride == 2) {
        RXB_TRY(subsample2(x, Bi, b.Hin, b.Win, b.cin, n.XS, st));
        xs = n.XS;
      }
      RXB_TRY(conv(n, Bi, b.Hout, b.Wout, xs, b.cin, b.cd, 1, 0, nullptr, n.CD, st));
      idn = n.CD;
    }
    RXB_TRY(bn_add_relu(n.C3, n.fold_scale + b.b3.fold_off, n.fold_shift + b.b3.fold_off, idn,
                        b.down ? n.fold_scale + b.bd.fold_off : nullptr, b.down ? n.fold_shift + b.bd.fold_off : nullptr,
                        Mout, 4 * b.width, y, st));
    std::swap(x, y);
  }
  const RnBlock& last = n.blocks.back();
  RXB_TRY(gap_mean(x, Bi, last.Hout * last.Wout, n.feat_dim, n.feat, st));
  // head (models.py:31-39, 44-55), fp32: concat of the thirds' means -> BN1d -> [Dropout] -> Linear -> ReLU -> BN1d
  // -> [Dropout] -> Linear
  const int F3 = 3 * n.feat_dim, SF = c.size_features, NC = c.num_classes;
  RXB_TRY(two_sites_concat(n.feat, c.B, c.G, n.feat_dim, n.cat, st));
  RXB_TRY(affine_rows(n.cat, c.B, F3, n.fold_scale + n.bn_m0.fold_off, n.fold_shift + n.bn_m0.fold_off, 0, n.h0, st));
  RXB_TRY(sgemm_strided(c.B, SF, F3, n.h0, F3, 1, n.params + n.fc1_w, 1, F3, n.params + n.fc1_b, n.h1, SF, 1, st));
  RXB_TRY(affine_rows(n.h1, c.B, SF, n.fold_scale + n.bn_m4.fold_off, n.fold_shift + n.bn_m4.fold_off, 1, n.h2, st));
  RXB_TRY(sgemm_strided(c.B, NC, SF, n.h2, SF, 1, n.params + n.fc2_w, 1, SF, n.params + n.fc2_b, n.logits, NC, 1, st));
  return RXB_OK;
}


// ================================================================================================ training
BnFold fold_of(const rxb_rn50& n, const RnBn& bn) {
  BnFold f;
  f.scale = n.fold_scale + bn.fold_off; f.shift = n.fold_shift + bn.fold_off;
  f.mean = n.fold_mean + bn.fold_off; f.rstd = n.fold_rstd + bn.fold_off;
  return f;
}

// forward conv with per-output-channel batch sums for the BatchNorm behind it (`stat`), and an optional BatchNorm+ReLU
// prologue whose fold is derived inside the kernel from the sums of the BatchNorm in front (`pro`, count = pixels)
int conv_t(const rxb_rn50& n, int B, int H, int W, const __nv_bfloat16* A, int cin, const RnConv& cv, int taps, int pad,
           const RnBn* pro, float pro_count, const RnBn& stat, __nv_bfloat16* out, cudaStream_t st) {
  GemmParams p = {};
  p.B = B; p.H = H; p.W = W;
  p.n_total = cv.cout;
  p.taps_x = p.taps_y = taps;
  p.pad_x = p.pad_y = pad;
  p.cin = cin;
  p.epi_mode = EPI_STORE;
  p.do_stats = 1;
  p.ch_sum = n.st_sum + stat.fold_off;
  p.ch_sumsq = n.st_sq + stat.fold_off;
  if (pro) {
    const BnFold f = fold_of(n, *pro);
    p.scale = f.scale; p.shift = f.shift;
    BnPrepArgs a = {};
    a.sum = n.st_sum + pro->fold_off; a.sumsq = n.st_sq + pro->fold_off;
    a.gamma = n.params + pro->gamma_off; a.beta = n.params + pro->beta_off;
    a.rmean = n.buffers + pro->rm_off; a.rvar = n.buffers + pro->rv_off;
    a.count = pro_count; a.eps = n.cfg.bn_eps; a.momentum = n.cfg.bn_momentum; a.training = 1;
    a.f_scale = f.scale; a.f_shift = f.shift; a.f_mean = f.mean; a.f_rstd = f.rstd;
    p.prep = a;
  }
  return launch_conv_gemm(p, A, cin, n.arena + cv.fwd_off, out, cv.cout, 0, nullptr, 0, cin <= 32 ? 32 : 64, pro != nullptr, st);
}

int prep_t(const rxb_rn50& n, const RnBn& bn, float count, cudaStream_t st) {
  return bn_prep(n.st_sum + bn.fold_off, n.st_sq + bn.fold_off, count, n.params + bn.gamma_off, n.params + bn.beta_off,
                 n.buffers + bn.rm_off, n.buffers + bn.rv_off, n.cfg.bn_eps, n.cfg.bn_momentum, 1, bn.C, fold_of(n, bn), st);
}

// plain forward-style GEMM used as a data gradient: out[p, 0:n_out] = sum dOut(p shifted)[k] * Wd[tap][n_out][k]
int dgrad_plain(const rxb_rn50& n, int B, int H, int W, const __nv_bfloat16* dOut, int k_contract, long long wd_off, int n_out,
                int taps, int pad, __nv_bfloat16* out, cudaStream_t st) {
  GemmParams p = {};
  p.B = B; p.H = H; p.W = W;
  p.n_total = n_out;
  p.taps_x = p.taps_y = taps;
  p.pad_x = p.pad_y = pad;
  p.cin = k_contract;
  p.epi_mode = EPI_STORE;
  return launch_conv_gemm(p, dOut, k_contract, n.arena + wd_off, out, n_out, 0, nullptr, 0, k_contract <= 32 ? 32 : 64, false, st);
}

// data gradient fused with the ReLU / BatchNorm backward of the layer that produced the conv input (bn over X): dz out,
// sum dz -> d_sum[bn]; degenerate channels also get sum dz*x -> d_sq[bn] (see bn_degenerate)
int dgrad_bn(const rxb_rn50& n, int B, int H, int W, const __nv_bfloat16* dOut, int k_contract, long long wd_off, int n_out,
             int taps, int pad, const __nv_bfloat16* X, const RnBn& bn, __nv_bfloat16* out, cudaStream_t st) {
  GemmParams p = {};
  p.B = B; p.H = H; p.W = W;
  p.n_total = n_out;
  p.taps_x = p.taps_y = taps;
  p.pad_x = p.pad_y = pad;
  p.cin = k_contract;
  p.epi_mode = EPI_DGRAD_BN;
  p.out_mode = OUT_DY;
  p.do_stats = 1;
  p.ch_sum = n.d_sum + bn.fold_off; p.ch_sumsq = n.d_sq + bn.fold_off;
  const BnFold f = fold_of(n, bn);
  p.e_scale = f.scale; p.e_shift = f.shift;
  p.e_gamma = n.params + bn.gamma_off; p.e_beta = n.params + bn.beta_off;
  return launch_conv_gemm(p, dOut, k_contract, n.arena + wd_off, out, n_out, 0, X, n_out, k_contract <= 32 ? 32 : 64, false, st);
}

// acc += plain data gradient: the L2 reduce-add epilogue with an always-true ReLU test (scale 1, shift 1e30); `like` is
// any [M, n_out] bf16 tensor (the kernel stages its result over a tile of it)
int dgrad_accumulate(const rxb_rn50& n, int B, int H, int W, const __nv_bfloat16* dOut, int k_contract, long long wd_off,
                     int n_out, const __nv_bfloat16* like, __nv_bfloat16* acc, cudaStream_t st) {
  GemmParams p = {};
  p.B = B; p.H = H; p.W = W;
  p.n_total = n_out;
  p.taps_x = p.taps_y = 1;
  p.cin = k_contract;
  p.epi_mode = EPI_DGRAD_BN;
  p.out_mode = OUT_G_ACCUM;
  p.do_stats = 1;
  p.ch_sum = n.scratch_c; p.ch_sumsq = nullptr;
  p.e_scale = n.ones; p.e_shift = n.big;
  return launch_conv_gemm(p, dOut, k_contract, n.arena + wd_off, acc, n_out, 0, like, n_out, 64, false, st);
}

int wgrad(int B, int H, int W, const __nv_bfloat16* A, int cin, int taps, int pad, const BnFold* pro, const __nv_bfloat16* dOut,
          int cout, float* dW, int w_mode, cudaStream_t st, int n_tile_override = 0) {
  const int n_tile = n_tile_override ? n_tile_override
                                     : cout <= 256 ? cout : (cout % 256 == 0 ? 256 : (cout % 128 == 0 ? 128 : 64));
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
    RXB_TRY(launch_conv_wgrad(p, A, cin, dOut, cout, st));
  }
  return RXB_OK;
}

int finalize_t(const rxb_rn50& n, const RnBn& bn, const RnConv* consumer, float count, cudaStream_t st) {
  // consumer != nullptr: sum dy*x from the consumer conv's W.dW (its weight gradient is complete); else d_sq already
  // holds sum dy*xhat from a direct reduction
  const int taps = consumer ? consumer->k * consumer->k : 0;
  return bn_bwd_finalize(1, consumer ? n.params + consumer->w_off : nullptr, consumer ? n.grads + consumer->w_off : nullptr,
                         consumer ? consumer->cout : 0, taps, n.d_sum + bn.fold_off, n.d_sq + bn.fold_off, fold_of(n, bn), count,
                         bn.C, n.grads + bn.gamma_off, n.grads + bn.beta_off, nullptr, nullptr,
                         consumer ? n.params + bn.gamma_off : nullptr, consumer ? n.params + bn.beta_off : nullptr, st);
}

int apply_t(const rxb_rn50& n, const RnBn& bn, __nv_bfloat16* dy, const __nv_bfloat16* X, long long M, __nv_bfloat16* dst,
            cudaStream_t st) {
  return bn_bwd_apply(dy, X, M, bn.C, fold_of(n, bn), n.d_sum + bn.fold_off, n.d_sq + bn.fold_off, st, dst);
}

int forward_train(rxb_rn50& n, const void* input, const float* mask0, const float* mask1, cudaStream_t st) {
  const rxb_rn50_config& c = n.cfg;
  const int Bi = n.Bi;
  RXB_CUDA(cudaMemsetAsync(n.zero_begin, 0, n.zero_bytes, st));
  RXB_TRY(conv_t(n, Bi, n.Hs, n.Ws, static_cast<const __nv_bfloat16*>(input), 32, n.conv0, 4, 2, nullptr, 0.f, n.bn0, n.S0, st));
  RXB_TRY(prep_t(n, n.bn0, (float)((long long)Bi * n.Hs * n.Ws), st));
  RXB_TRY(stem_bn_relu_maxpool(n.S0, Bi, n.Hs, n.Ws, n.fold_scale + n.bn0.fold_off, n.fold_shift + n.bn0.fold_off, n.X0, 64,
                               n.pool_idx, n.scratch_c, n.scratch_c + 64, st));
  const __nv_bfloat16* x = n.X0;
  for (auto& b : n.blocks) {
    const long long Min = (long long)Bi * b.Hin * b.Win, Mout = (long long)Bi * b.Hout * b.Wout;
    RXB_TRY(conv_t(n, Bi, b.Hin, b.Win, x, b.cin, b.c1, 1, 0, nullptr, 0.f, b.b1, b.sC1, st));
    if (b.stride == 1) {
      RXB_TRY(conv_t(n, Bi, b.Hin, b.Win, b.sC1, b.width, b.c2, 3, 1, &b.b1, (float)Min, b.b2, b.sC2, st));
    } else {
      RXB_TRY(prep_t(n, b.b1, (float)Min, st));
      RXB_TRY(s2d_bn_relu(b.sC1, Bi, b.Hin, b.Win, b.width, n.fold_scale + b.b1.fold_off, n.fold_shift + b.b1.fold_off, b.sS2, st));
      RXB_TRY(conv_t(n, Bi, b.Hout, b.Wout, b.sS2, 4 * b.width, b.c2, 2, 1, nullptr, 0.f, b.b2, b.sC2, st));
    }
    RXB_TRY(conv_t(n, Bi, b.Hout, b.Wout, b.sC2, b.width, b.c3, 1, 0, &b.b2, (float)Mout, b.b3, b.sC3, st));
    const __nv_bfloat16* idn = x;
    if (b.down) {
      const __nv_bfloat16* xs = x;
      if (b.stride == 2) {
        RXB_TRY(subsample2(x, Bi, b.Hin, b.Win, b.cin, b.sXS, st));
        xs = b.sXS;
      }
      RXB_TRY(conv_t(n, Bi, b.Hout, b.Wout, xs, b.cin, b.cd, 1, 0, nullptr, 0.f, b.bd, b.sCD, st));
      RXB_TRY(prep_t(n, b.bd, (float)Mout, st));
      idn = b.sCD;
    }
    RXB_TRY(prep_t(n, b.b3, (float)Mout, st));
    RXB_TRY(bn_add_relu(b.sC3, n.fold_scale + b.b3.fold_off, n.fold_shift + b.b3.fold_off, idn,
                        b.down ? n.fold_scale + b.bd.fold_off : nullptr, b.down ? n.fold_shift + b.bd.fold_off : nullptr,
                        Mout, 4 * b.width, b.sOut, st));
    x = b.sOut;
  }
  const RnBlock& last = n.blocks.back();
  RXB_TRY(gap_mean(x, Bi, last.Hout * last.Wout, n.feat_dim, n.feat, st));
  const int F3 = 3 * n.feat_dim, SF = c.size_features, NC = c.num_classes;
  RXB_TRY(two_sites_concat(n.feat, c.B, c.G, n.feat_dim, n.cat, st));
  RXB_TRY(bn1d_train_fwd(n.cat, c.B, F3, 0, n.params + n.bn_m0.gamma_off, n.params + n.bn_m0.beta_off, n.buffers + n.bn_m0.rm_off,
                         n.buffers + n.bn_m0.rv_off, c.bn_eps, c.bn_momentum, n.h0, n.sv_mean0, n.sv_rstd0, st));
  RXB_TRY(mul_elems(n.h0, mask0, (long long)c.B * F3, n.y0m, st));
  RXB_TRY(sgemm_strided(c.B, SF, F3, n.y0m, F3, 1, n.params + n.fc1_w, 1, F3, n.params + n.fc1_b, n.h1, SF, 1, st));
  RXB_TRY(bn1d_train_fwd(n.h1, c.B, SF, 1, n.params + n.bn_m4.gamma_off, n.params + n.bn_m4.beta_off, n.buffers + n.bn_m4.rm_off,
                         n.buffers + n.bn_m4.rv_off, c.bn_eps, c.bn_momentum, n.y1, n.sv_mean1, n.sv_rstd1, st));
  RXB_TRY(mul_elems(n.y1, mask1, (long long)c.B * SF, n.y1m, st));
  RXB_TRY(sgemm_strided(c.B, NC, SF, n.y1m, SF, 1, n.params + n.fc2_w, 1, SF, n.params + n.fc2_b, n.logits, NC, 1, st));
  return RXB_OK;
}

int backward(rxb_rn50& n, const void* input, const float* mask0, const float* mask1, cudaStream_t st) {
  const rxb_rn50_config& c = n.cfg;
  const int Bi = n.Bi;
  const int F3 = 3 * n.feat_dim, SF = c.size_features, NC = c.num_classes;
  // ---- head (models.py:31-39 backwards)
  RXB_TRY(sgemm_strided(NC, SF, c.B, n.dlogits, 1, NC, n.y1m, SF, 1, nullptr, n.grads + n.fc2_w, SF, 1, st));
  RXB_TRY(column_sum(n.dlogits, c.B, NC, NC, n.grads + n.fc2_b, st));
  RXB_TRY(sgemm_strided(c.B, SF, NC, n.dlogits, NC, 1, n.params + n.fc2_w, SF, 1, nullptr, n.dy1, SF, 1, st));
  RXB_TRY(mul_elems(n.dy1, mask1, (long long)c.B * SF, n.dy1, st));
  RXB_TRY(bn1d_bwd(n.dy1, n.h1, c.B, SF, 1, n.params + n.bn_m4.gamma_off, n.sv_mean1, n.sv_rstd1, n.dh1, n.grads + n.bn_m4.gamma_off,
                   n.grads + n.bn_m4.beta_off, st));
  RXB_TRY(sgemm_strided(SF, F3, c.B, n.dh1, 1, SF, n.y0m, F3, 1, nullptr, n.grads + n.fc1_w, F3, 1, st));
  RXB_TRY(column_sum(n.dh1, c.B, SF, SF, n.grads + n.fc1_b, st));
  RXB_TRY(sgemm_strided(c.B, F3, SF, n.dh1, SF, 1, n.params + n.fc1_w, F3, 1, nullptr, n.dy0v, F3, 1, st));
  RXB_TRY(mul_elems(n.dy0v, mask0, (long long)c.B * F3, n.dy0v, st));
  RXB_TRY(bn1d_bwd(n.dy0v, n.cat, c.B, F3, 0, n.params + n.bn_m0.gamma_off, n.sv_mean0, n.sv_rstd0, n.dcat, n.grads + n.bn_m0.gamma_off,
                   n.grads + n.bn_m0.beta_off, st));
  RXB_TRY(two_sites_concat_bwd(n.dcat, c.B, c.G, n.feat_dim, n.dfeat, st));
  const RnBlock& last = n.blocks.back();
  __nv_bfloat16 *D = n.Da, *Dn = n.Db;
  RXB_TRY(gap_mean_bwd(n.dfeat, Bi, last.Hout * last.Wout, n.feat_dim, D, st));
  // ---- bottlenecks, last to first
  for (int bi = (int)n.blocks.size() - 1; bi >= 0; --bi) {
    RnBlock& b = n.blocks[bi];
    const long long Min = (long long)Bi * b.Hin * b.Win, Mout = (long long)Bi * b.Hout * b.Wout;
    const int w = b.width, C4 = 4 * b.width;
    const __nv_bfloat16* xin = bi > 0 ? n.blocks[bi - 1].sOut : n.X0;
    // du = D*[out>0] in place, BatchNorm3 (and downsample BatchNorm) reductions
    RXB_TRY(relu_bwd_sums(D, b.sOut, b.sC3, fold_of(n, b.b3), b.down ? b.sCD : nullptr, b.down ? fold_of(n, b.bd) : fold_of(n, b.b3),
                          Mout, C4, n.d_sum + b.b3.fold_off, n.d_sq + b.b3.fold_off, b.down ? n.d_sum + b.bd.fold_off : nullptr,
                          b.down ? n.d_sq + b.bd.fold_off : nullptr, st));
    RXB_TRY(finalize_t(n, b.b3, nullptr, (float)Mout, st));
    RXB_TRY(apply_t(n, b.b3, D, b.sC3, Mout, n.DC3, st));                               // dc3
    if (b.down) {
      RXB_TRY(finalize_t(n, b.bd, nullptr, (float)Mout, st));
      RXB_TRY(apply_t(n, b.bd, D, b.sCD, Mout, n.DCD, st));                             // dcd
    }
    // conv3
    const BnFold f2 = fold_of(n, b.b2), f1 = fold_of(n, b.b1);
    RXB_TRY(wgrad(Bi, b.Hout, b.Wout, b.sC2, w, 1, 0, &f2, n.DC3, C4, n.grads + b.c3.w_off, 0, st));
    RXB_TRY(dgrad_bn(n, Bi, b.Hout, b.Wout, n.DC3, C4, b.c3.dgrad_off, w, 1, 0, b.sC2, b.b2, n.DZ2, st));
    RXB_TRY(finalize_t(n, b.b2, &b.c3, (float)Mout, st));
    RXB_TRY(apply_t(n, b.b2, n.DZ2, b.sC2, Mout, nullptr, st));                         // DZ2 := dc2
    // conv2
    if (b.stride == 1) {
      if (w <= 128) {
        // 3x3 weight gradient, Cin <= 128: 32 output channels per launch keep the nine tap accumulators inside the 512
        // TMEM columns, so the activation tile is loaded and transformed ONCE per pixel tile and the taps are offsets
        // into one full-halo dOut box (conv_gemm.cu shift_dout mode) instead of nine shifted re-loads of A
        RXB_TRY(wgrad(Bi, b.Hin, b.Win, b.sC1, w, 3, 1, &f1, n.DZ2, w, n.grads + b.c2.w_off, 0, st, 32));
      } else {
        // wider: per-tap A boxes; result into the tap-major scratch by bulk L2 reduce-adds, then transposed into OIHW
        RXB_CUDA(cudaMemsetAsync(n.wg_scratch, 0, sizeof(float) * 9ll * w * w, st));
        RXB_TRY(wgrad(Bi, b.Hin, b.Win, b.sC1, w, 3, 1, &f1, n.DZ2, w, n.wg_scratch, 3, st));
        RXB_TRY(wgrad_finish(n.wg_scratch, w, w, 3, 0, n.grads + b.c2.w_off, st));
      }
      RXB_TRY(dgrad_bn(n, Bi, b.Hin, b.Win, n.DZ2, w, b.c2.dgrad_off, w, 3, 1, b.sC1, b.b1, n.DZ1, st));
      RXB_TRY(finalize_t(n, b.b1, &b.c2, (float)Min, st));
    } else {
      RXB_CUDA(cudaMemsetAsync(n.wg_scratch, 0, sizeof(float) * 16ll * w * w, st));
      RXB_TRY(wgrad(Bi, b.Hout, b.Wout, b.sS2, 4 * w, 2, 1, nullptr, n.DZ2, w, n.wg_scratch, 3, st));
      RXB_TRY(wgrad_finish(n.wg_scratch, w, w, 3, 1, n.grads + b.c2.w_off, st));
      RXB_TRY(dgrad_plain(n, Bi, b.Hout, b.Wout, n.DZ2, w, b.c2.dgrad_off, 4 * w, 2, 0, n.DS2, st));
      RXB_TRY(s2d_bn_relu_bwd(n.DS2, b.sC1, Bi, b.Hin, b.Win, w, f1, n.DZ1, n.d_sum + b.b1.fold_off, n.d_sq + b.b1.fold_off, st));
      RXB_TRY(finalize_t(n, b.b1, nullptr, (float)Min, st));
    }
    RXB_TRY(apply_t(n, b.b1, n.DZ1, b.sC1, Min, nullptr, st));                          // DZ1 := dc1
    // conv1 and the gradient of the block input
    RXB_TRY(wgrad(Bi, b.Hin, b.Win, xin, b.cin, 1, 0, nullptr, n.DZ1, w, n.grads + b.c1.w_off, 0, st));
    if (!b.down) {
      // identity: dL/dx = du + conv1 data gradient, accumulated into the buffer that holds du
      RXB_TRY(dgrad_accumulate(n, Bi, b.Hin, b.Win, n.DZ1, w, b.c1.dgrad_off, b.cin, xin, D, st));
    } else {
      RXB_TRY(dgrad_plain(n, Bi, b.Hin, b.Win, n.DZ1, w, b.c1.dgrad_off, b.cin, 1, 0, Dn, st));
      const __nv_bfloat16* xs = b.stride == 2 ? b.sXS : xin;
      RXB_TRY(wgrad(Bi, b.Hout, b.Wout, xs, b.cin, 1, 0, nullptr, n.DCD, C4, n.grads + b.cd.w_off, 0, st));
      if (b.stride == 2) {
        RXB_TRY(dgrad_plain(n, Bi, b.Hout, b.Wout, n.DCD, C4, b.cd.dgrad_off, b.cin, 1, 0, n.DXS, st));
        RXB_TRY(upsample2_add(n.DXS, Bi, b.Hin, b.Win, b.cin, Dn, st));
      } else {
        RXB_TRY(dgrad_accumulate(n, Bi, b.Hin, b.Win, n.DCD, C4, b.cd.dgrad_off, b.cin, xin, Dn, st));
      }
      std::swap(D, Dn);
    }
  }
  // ---- stem: max-pool / ReLU / BatchNorm backward, weight gradient of the 7x7 convolution
  RXB_TRY(stem_pool_bwd(D, n.pool_idx, n.S0, Bi, n.Hs, n.Ws, fold_of(n, n.bn0), n.dy0, n.d_sum + n.bn0.fold_off, n.d_sq + n.bn0.fold_off, st));
  const long long Ms = (long long)Bi * n.Hs * n.Ws;
  RXB_TRY(finalize_t(n, n.bn0, nullptr, (float)Ms, st));
  RXB_TRY(apply_t(n, n.bn0, n.dy0, n.S0, Ms, nullptr, st));
  RXB_TRY(wgrad(Bi, n.Hs, n.Ws, static_cast<const __nv_bfloat16*>(input), 32, 4, 2, nullptr, n.dy0, 64, n.grads + n.conv0.w_off, 1, st));
  return RXB_OK;
}

}  // namespace
}  // namespace rxb

extern "C" {

int64_t rxb_rn50_param_count(const rxb_rn50_config* cfg) {
  if (rxb::check_cfg(cfg)) return -1;
  rxb_rn50 n;
  n.cfg = *cfg;
  rxb::plan(n, nullptr);
  return n.n_params;
}

int64_t rxb_rn50_buffer_count(const rxb_rn50_config* cfg) {
  if (rxb::check_cfg(cfg)) return -1;
  rxb_rn50 n;
  n.cfg = *cfg;
  rxb::plan(n, nullptr);
  return n.n_buffers;
}

size_t rxb_rn50_workspace_bytes(const rxb_rn50_config* cfg, int training) {
  if (rxb::check_cfg(cfg)) return 0;
  rxb_rn50 n;
  n.cfg = *cfg;
  n.training = training;
  return rxb::plan(n, nullptr);
}

int rxb_rn50_create(const rxb_rn50_config* cfg, float* params, float* grads, float* momentum, float* buffers, void* workspace,
                    size_t workspace_bytes, int training, rxb_rn50** out) {
  using namespace rxb;
  int rc = check_cfg(cfg);
  if (rc) return rc;
  RXB_CHECK_ARG(params && buffers && workspace && out, "rxb_rn50_create: null pointer");
  RXB_CHECK_ARG(!training || (grads && momentum), "rxb_rn50_create: training needs grads and momentum");
  RXB_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "rxb_rn50_create: workspace must be 256B aligned");
  rc = rxb_check_device();
  if (rc) return rc;
  rxb_rn50* n = new rxb_rn50();
  n->cfg = *cfg;
  n->training = training;
  n->params = params;
  n->grads = grads;
  n->momentum = momentum;
  n->buffers = buffers;
  const size_t need = plan(*n, static_cast<uint8_t*>(workspace));
  if (need > workspace_bytes) {
    delete n;
    return set_error(RXB_ERR_INVALID, "rxb_rn50_create: workspace %zu B < required %zu B", workspace_bytes, need);
  }
  cudaError_t ce = cudaMemcpy(n->jobs_dev, n->jobs.data(), n->jobs.size() * sizeof(RepackJob), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess)
    ce = cudaMemcpy(n->fold_jobs_dev, n->fold_jobs.data(), n->fold_jobs.size() * sizeof(BnFoldJob), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess && training) {
    std::vector<float> ones(2048, 1.f), big(2048, 1e30f);
    ce = cudaMemcpy(n->ones, ones.data(), ones.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(n->big, big.data(), big.size() * sizeof(float), cudaMemcpyHostToDevice);
  }
  if (ce != cudaSuccess) {
    delete n;
    return set_error(RXB_ERR_CUDA, "rxb_rn50_create: upload of the job tables failed: %s", cudaGetErrorString(ce));
  }
  *out = n;
  return RXB_OK;
}

void rxb_rn50_destroy(rxb_rn50* net) { delete net; }

int rxb_rn50_sync_weights(rxb_rn50* net, rxb_stream_t stream) {
  RXB_CHECK_ARG(net, "rxb_rn50_sync_weights: null");
  return rxb::repack_weights(net->params, net->arena, net->jobs_dev, (int)net->jobs.size(), net->max_job_elems,
                             rxb::as_stream(stream));
}

int rxb_rn50_forward(rxb_rn50* net, const void* input_s2d, float* logits_out, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(net && input_s2d && logits_out, "rxb_rn50_forward: null pointer");
  RXB_CHECK_ARG(!net->training, "rxb_rn50_forward: plan was created for training (its buffers are laid out for backward)");
  cudaStream_t st = as_stream(stream);
  RXB_TRY(forward_eval(*net, input_s2d, st));
  RXB_CUDA(cudaMemcpyAsync(logits_out, net->logits, sizeof(float) * net->cfg.B * net->cfg.num_classes, cudaMemcpyDeviceToDevice, st));
  return RXB_OK;
}

int64_t rxb_rn50_head_offset(const rxb_rn50_config* cfg) {
  if (rxb::check_cfg(cfg)) return -1;
  rxb_rn50 n;
  n.cfg = *cfg;
  rxb::plan(n, nullptr);
  return n.head_off;
}

int rxb_rn50_train_step(rxb_rn50* net, const void* input_s2d, const int64_t* target, const float* drop_mask0,
                        const float* drop_mask1, int global_batch, float* loss_out, float* logits_out, float* feat_out,
                        rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(net && input_s2d && target && drop_mask0 && drop_mask1, "rxb_rn50_train_step: null pointer");
  RXB_CHECK_ARG(net->training, "rxb_rn50_train_step: plan was created for inference");
  RXB_CHECK_ARG(global_batch >= 1, "rxb_rn50_train_step: bad global_batch");
  cudaStream_t st = as_stream(stream);
  rxb_rn50& n = *net;
  RXB_CUDA(cudaMemsetAsync(n.grads, 0, sizeof(float) * n.n_params, st));
  RXB_TRY(forward_train(n, input_s2d, drop_mask0, drop_mask1, st));
  RXB_TRY(rxb_softmax_ce(n.logits, n.cfg.num_classes, target, n.cfg.B, n.cfg.num_classes, n.loss_rows, n.dlogits,
                         1.f / (float)global_batch, stream));
  if (loss_out) RXB_TRY(sum_scale(n.loss_rows, n.cfg.B, 1.f / (float)global_batch, loss_out, st));
  if (logits_out)
    RXB_CUDA(cudaMemcpyAsync(logits_out, n.logits, sizeof(float) * n.cfg.B * n.cfg.num_classes, cudaMemcpyDeviceToDevice, st));
  if (feat_out)
    RXB_CUDA(cudaMemcpyAsync(feat_out, n.feat, sizeof(float) * n.Bi * n.feat_dim, cudaMemcpyDeviceToDevice, st));
  RXB_TRY(backward(n, input_s2d, drop_mask0, drop_mask1, st));
  return RXB_OK;
}

int rxb_rn50_sgd(rxb_rn50* net, float lr, float mu, float wd, int nesterov, float grad_scale, int head_only,
                 rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(net && net->training, "rxb_rn50_sgd: needs a training plan");
  const long long b = head_only ? net->head_off : 0;
  RXB_TRY(rxb_sgd_step(net->params + b, net->grads + b, net->momentum + b, net->n_params - b, lr, mu, wd, nesterov, grad_scale,
                       stream));
  return repack_weights(net->params, net->arena, net->jobs_dev, (int)net->jobs.size(), net->max_job_elems, as_stream(stream));
}

}  // extern "C"
