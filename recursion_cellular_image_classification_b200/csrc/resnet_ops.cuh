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

// ---------------------------------------------------------------------------------------------- training pieces
namespace rxb {

// Residual-join backward (torchvision Bottleneck: out = relu(bn3(c3) + idn)), first half:
//   du = D * [out > 0]   written over D (bf16 [M,C] dense);
//   dsum3 += sum du ; dsq3 += sum du*xhat(c3)           (BatchNorm3 backward reductions, xhat from f3.mean / f3.rstd)
//   dsumd / dsqd likewise against cd when the block has a downsample BatchNorm (cd != nullptr).
int relu_bwd_sums(__nv_bfloat16* D, const __nv_bfloat16* out, const __nv_bfloat16* c3, BnFold f3, const __nv_bfloat16* cd,
                  BnFold fd, long long M, int C, float* dsum3, float* dsq3, float* dsumd, float* dsqd, cudaStream_t st);
// Backward of s2d_bn_relu: dz[b,y,x,c] = DS[b,y/2,x/2,(y%2*2+x%2)*C+c] * [X*scale+shift > 0]   (bf16 [B,H,W,C]);
// dsum += sum dz ; dsq += sum dz*xhat(X).
int s2d_bn_relu_bwd(const __nv_bfloat16* DS, const __nv_bfloat16* X, int B, int H, int W, int C, BnFold f,
                    __nv_bfloat16* dz, float* dsum, float* dsq, cudaStream_t st);
// Backward of subsample2: Din[b,2i,2j,c] += DXS[b,i,j,c]
int upsample2_add(const __nv_bfloat16* DXS, int B, int H, int W, int C, __nv_bfloat16* Din, cudaStream_t st);
// Backward of gap_mean: D[b,p,c] = dfeat[b,c] / HW   (bf16)
int gap_mean_bwd(const float* dfeat, int B, int HW, int C, __nv_bfloat16* D, cudaStream_t st);
// Backward of two_sites_concat: dfeat[b*G+g, f] = dcat[b, third(g)*F + f] / (images in that third)
int two_sites_concat_bwd(const float* dcat, int bs, int G, int F, float* dfeat, cudaStream_t st);
// BatchNorm1d in training mode over rows [rows, F] (models.py:32,36), optionally on relu(x) (the ReLU of models.py:35
// sits between Linear and BatchNorm1d): y = (x' - mean) * rstd * gamma + beta with the batch's biased variance,
// running statistics updated with momentum and the unbiased variance; mean / rstd saved for backward.
int bn1d_train_fwd(const float* x, int rows, int F, int pre_relu, const float* gamma, const float* beta, float* rmean,
                   float* rvar, float eps, float momentum, float* y, float* save_mean, float* save_rstd, cudaStream_t st);
// dx = gamma*rstd*(dy - mean(dy) - xhat*mean(dy*xhat)) (times [x > 0] when pre_relu); dgamma = sum dy*xhat; dbeta = sum dy
int bn1d_bwd(const float* dy, const float* x, int rows, int F, int pre_relu, const float* gamma, const float* save_mean,
             const float* save_rstd, float* dx, float* dgamma, float* dbeta, cudaStream_t st);
// Weight gradient from the tap-major scratch of conv_wgrad's w_mode 3 into torch's OIHW layout (overwrites dW):
//   s2d == 0: scratch [k*k][N][K]            -> dW[n][c][t] = scratch[t][n][c]                       (k x k filter)
//   s2d == 1: scratch [4][N][4K] (2x2 taps over the space-to-depth input) -> dW[n][c][dy][dx] (3x3 stride-2 filter),
//             dy = 2*sy+py-1, dx = 2*sx+px-1, channel (py*2+px)*K + c
int wgrad_finish(const float* scratch, int N, int K, int k, int s2d, float* dW, cudaStream_t st);
// y[i] = x[i] * m[i]   (Dropout with an explicit mask holding 0 or 1/(1-p); the same call is its backward)
int mul_elems(const float* x, const float* m, long long n, float* y, cudaStream_t st);

}  // namespace rxb
