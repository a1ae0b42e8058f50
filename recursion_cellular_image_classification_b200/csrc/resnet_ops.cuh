// resnet_ops.cuh — the HBM-bound kernels the ResNet-50 TwoSitesNN executor (resnet.cu) needs beyond elementwise.cuh:
// the residual join, the space-to-depth form of a stride-2 3x3 convolution's input, the 1x1 stride-2 subsampling, the
// global average pool, the reference's site/control feature concat (models.py:46-53) and the MLP head's row-wise
// BatchNorm1d / ReLU (models.py:31-39).
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace rxb {

// out[b,i,j,(py*2+px)*C + c] = relu(X[b,2i+py,2j+px,c]*scale[c] + shift[c])  (0 where the source pixel is outside the
// image).  X bf16 [B,H,W,C]; out bf16 [B,(H+1)/2,(W+1)/2,4C].  A 3x3 stride-2 pad-1 convolution over relu(bn(X)) is
// the stride-1 2x2-tap pad-1 convolution over this tensor with the weights of RP_3x3S2_FWD (dy = 2*sy+py-1).
// scale == nullptr: plain copy (no BatchNorm / ReLU).
int s2d_bn_relu(const __nv_bfloat16* X, int B, int H, int W, int C, const float* scale, const float* shift,
                __nv_bfloat16* out, cudaStream_t st);
// out[b,i,j,c] = X[b,2i,2j,c] — the input of a 1x1 stride-2 convolution (torchvision's downsample branch).
int subsample2(const __nv_bfloat16* X, int B, int H, int W, int C, __nv_bfloat16* out, cudaStream_t st);
// out = relu(c3*s3 + h3 + (sd ? idn*sd + hd : idn))   bf16 [M,C] dense — the bottleneck's residual join
// (torchvision Bottleneck.forward: out = relu(bn3(conv3) + identity)).
int bn_add_relu(const __nv_bfloat16* c3, const float* s3, const float* h3, const __nv_bfloat16* idn, const float* sd,
                const float* hd, long long M, int C, __nv_bfloat16* out, cudaStream_t st);
// feat[b,c] = mean_p X[b,p,c]   (adaptive_avg_pool2d(1) + flatten; X is post-ReLU already)
int gap_mean(const __nv_bfloat16* X, int B, int HW, int C, float* feat, cudaStream_t st);
// models.py:44-53: feat f32 [bs*G, F] -> cat f32 [bs, 3F] = [mean of the first third of the G images | second | third]
int two_sites_concat(const float* feat, int bs, int G, int F, float* cat, cudaStream_t st);
// y[r,f] = (pre_relu ? max(x[r,f],0) : x[r,f]) * scale[f] + shift[f]     (BatchNorm1d in eval mode, after an optional ReLU)
int affine_rows(const float* x, int rows, int F, const float* scale, const float* shift, int pre_relu, float* y,
                cudaStream_t st);

// Evaluation-mode BatchNorm folds for a whole network in ONE launch: scale = gamma*rsqrt(var+eps), shift = beta-mean*scale.
struct BnFoldJob {
  long long gamma_off, beta_off;   // into the flat fp32 parameters
  long long rm_off, rv_off;        // into the flat fp32 buffers
  long long fold_off;              // into the flat fold arrays
  int C, pad;
};
int bn_fold_eval_all(const float* params, const float* buffers, const BnFoldJob* jobs_dev, int n_jobs, int max_c, float eps,
                     float* fold_scale, float* fold_shift, cudaStream_t st);

}  // namespace rxb
