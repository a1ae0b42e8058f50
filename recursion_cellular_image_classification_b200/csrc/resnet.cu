// resnet.cu — the reference's REAL model on the device: TwoSitesNN (reference cell_classifier/models.py:7-57) =
// torchvision ResNet-50 trunk with the 6-channel stem (:16-29, fc -> Identity), per-sample feature means of the
// image / negative-control / positive-control thirds concatenated (:44-53) and the
// BatchNorm1d -> Dropout -> Linear(6144,1024) -> ReLU -> BatchNorm1d -> Dropout -> Linear(1024,1108) head (:31-39).
//
// This file is the evaluation-mode executor (BatchNorm from running statistics, Dropout = identity): what
// reference test.py:23-27 runs for every test batch, and what a checkpoint trained with the reference (models/
// best_model_<id>.pth, `module.`-prefixed, main.py:147) needs to be served on a B200.  Parameters and buffers are flat
// fp32 arrays in the reference model's own named_parameters() / named_buffers() order.
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
  long long w_off = 0, fwd_off = 0;
  int cin = 0, cout = 0, k = 1, stride = 1;
};
struct RnBlock {
  int cin = 0, width = 0, stride = 1, down = 0;
  int Hin = 0, Win = 0, Hout = 0, Wout = 0;
  RnConv c1, c2, c3, cd;
  RnBn b1, b2, b3, bd;
};

}  // namespace rxb

struct rxb_rn50 {
  rxb_rn50_config cfg;
  int training = 0;
  float *params = nullptr, *buffers = nullptr;
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
    if (k == 1) cv.fwd_off = add_job(cv.w_off, RP_1x1_FWD, cout, cin, (long long)cout * cin);
    else if (stride == 1) cv.fwd_off = add_job(cv.w_off, RP_3x3_FWD, cout, cin, 9ll * cout * cin);
    else cv.fwd_off = add_job(cv.w_off, RP_3x3S2_FWD, cout, cin, 16ll * cout * cin);
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
  n.Xa = bp.take<__nv_bfloat16>(mx_x);
  n.Xb = bp.take<__nv_bfloat16>(mx_x);
  n.C1 = bp.take<__nv_bfloat16>(mx_c1);
  n.C2 = bp.take<__nv_bfloat16>(mx_c2);
  n.C3 = bp.take<__nv_bfloat16>(mx_c3);
  n.CD = bp.take<__nv_bfloat16>(mx_c3);
  n.S2 = bp.take<__nv_bfloat16>(mx_s2);
  n.XS = bp.take<__nv_bfloat16>(mx_xs);
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

size_t rxb_rn50_workspace_bytes(const rxb_rn50_config* cfg) {
  if (rxb::check_cfg(cfg)) return 0;
  rxb_rn50 n;
  n.cfg = *cfg;
  return rxb::plan(n, nullptr);
}

int rxb_rn50_create(const rxb_rn50_config* cfg, float* params, float* buffers, void* workspace, size_t workspace_bytes,
                    rxb_rn50** out) {
  using namespace rxb;
  int rc = check_cfg(cfg);
  if (rc) return rc;
  RXB_CHECK_ARG(params && buffers && workspace && out, "rxb_rn50_create: null pointer");
  RXB_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "rxb_rn50_create: workspace must be 256B aligned");
  rc = rxb_check_device();
  if (rc) return rc;
  rxb_rn50* n = new rxb_rn50();
  n->cfg = *cfg;
  n->params = params;
  n->buffers = buffers;
  const size_t need = plan(*n, static_cast<uint8_t*>(workspace));
  if (need > workspace_bytes) {
    delete n;
    return set_error(RXB_ERR_INVALID, "rxb_rn50_create: workspace %zu B < required %zu B", workspace_bytes, need);
  }
  cudaError_t ce = cudaMemcpy(n->jobs_dev, n->jobs.data(), n->jobs.size() * sizeof(RepackJob), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess)
    ce = cudaMemcpy(n->fold_jobs_dev, n->fold_jobs.data(), n->fold_jobs.size() * sizeof(BnFoldJob), cudaMemcpyHostToDevice);
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
  cudaStream_t st = as_stream(stream);
  RXB_TRY(forward_eval(*net, input_s2d, st));
  RXB_CUDA(cudaMemcpyAsync(logits_out, net->logits, sizeof(float) * net->cfg.B * net->cfg.num_classes, cudaMemcpyDeviceToDevice, st));
  return RXB_OK;
}

}  // extern "C"
