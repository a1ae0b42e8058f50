// elementwise.cuh — launchers of the HBM-bound kernels around the convolutions: BatchNorm parameter
// folding, the fused BN+ReLU+pool forwards, and the BatchNorm/ReLU/pool backward pieces.
#pragma once
#include "common.cuh"

namespace rxb {

// Folded BatchNorm of one layer.  All arrays are fp32 [C] in the workspace.
struct BnFold {
  float* scale;   // gamma * rstd
  float* shift;   // beta - mean * scale
  float* mean;
  float* rstd;
};

// scale/shift/mean/rstd from batch sums (training) or running stats (eval); training also updates the
// running stats with momentum (unbiased variance), like torch.nn.BatchNorm2d.
int bn_prep(const float* sum, const float* sumsq, float count, const float* gamma, const float* beta,
            float* running_mean, float* running_var, float eps, float momentum, int training, int C, BnFold f,
            cudaStream_t st);

// S0 raw bf16 [B,Hs,Ws,64] -> out[B,Hs/2,Ws/2, ld_out] channels [0,64) = maxpool3x3/s2/p1(relu(bn(S0)));
// idx u8 [B,Hs/2,Ws/2,64] = argmax position in the 3x3 window; sum/sumsq += stats of the bf16 output.
int stem_bn_relu_maxpool(const __nv_bfloat16* S0, int B, int Hs, int Ws, const float* scale, const float* shift,
                         __nv_bfloat16* out, int ld_out, uint8_t* idx, float* sum, float* sumsq, cudaStream_t st);

// P[B,H/2,W/2,C] = avgpool2x2(relu(bn(X[B,H,W,C] (row stride ldx))))
int transition_pool_fwd(const __nv_bfloat16* X, int ldx, int B, int H, int W, int C, const float* scale,
                        const float* shift, __nv_bfloat16* P, cudaStream_t st);

// feat[b,c] = mean_p relu(bn(X[b,p,c]))   (norm5 -> relu -> adaptive_avg_pool2d(1))
int final_bn_relu_gap(const __nv_bfloat16* X, int ldx, int B, int HW, int C, const float* scale,
                      const float* shift, float* feat, cudaStream_t st);

// Backward through ReLU and the *local* part of BatchNorm for a consumer whose upstream gradient is
//   mode 0: da[p,c] = 0.25 * dP[b, y/2, x/2, c]     (avgpool 2x2 backward, transition)
//   mode 1: da[p,c] = dfeat[b,c] / (H*W)            (global average pool backward)
// dy = da * [x*scale+shift > 0];  G[p,c] = scale[c]*dy (bf16, overwrites);  dsum += sum dy,
// dsq += sum dy*xhat.
int bn_relu_bwd_to_G(int mode, const void* upstream, const __nv_bfloat16* X, int ldx, int B, int H, int W, int C,
                     BnFold f, __nv_bfloat16* G, float* dsum, float* dsq, cudaStream_t st);

// After the reductions of one BatchNorm's backward are complete (dsum = sum dy; dsq = sum dy*xhat):
//   dgamma = sum dy*xhat ; dbeta = sum dy
//   mode 0 (consumer of a concat buffer): corrA[c] += scale*dsum/M ; corrB[c] += scale*dsq/M
//   mode 1 (single consumer):             dsum[c] = dsum/M ; dsq[c] = dsq/M    (consumed by bn_bwd_apply)
// When W/dW are given (the conv that consumed relu(bn(x)), fp32 OIHW [K][C][taps] weights and their finished
// gradient), dsq is not read: with z = scale*x+shift and dz = dy,  sum_p dy*z = sum_{k,tap} W*dW  per input
// channel (both sides equal sum_p dL/dA' * A'), hence sum dy*x = (W.dW - shift*sum dy)/scale.  The dgrad kernel
// therefore only reduces sum dy - except for the channels bn_degenerate(gamma, beta) flags (gamma/beta: the
// BatchNorm's own weight and bias, may be null), whose sum dy*x the dgrad epilogue reduced directly into dsq.
int bn_bwd_finalize(int mode, const float* W, const float* dW, int K, int taps, float* dsum, float* dsq, BnFold f,
                    float count, int C, float* dgamma, float* dbeta, float* corrA, float* corrB, const float* gamma,
                    const float* beta, cudaStream_t st);

// out[c] = sum_p dy*x per input channel from W.dW (see above); the standalone form behind rxb_bn_sum_dyx_from_wdw.
int sum_dyx_from_wdw_launch(const float* W, const float* dW, int K, int C, int taps, const float* scale,
                            const float* shift, const float* sum_dy, float* out, cudaStream_t st);

// dx[p,c] = scale[c] * (dy[p,c] - m1[c] - xhat[p,c]*m2[c])  (bf16 [M,C] dense), in place on dy, or into dst when given.
// raw != nullptr (C = 64 or 128 ...: the row-walking kernel): m1 / m2 hold the RAW reductions sum(dy) and W.dW (direct
// sum(dy*x) for the channels bn_degenerate flags) as the fused 3x3 data+weight-gradient kernel leaves them; the
// kernel derives the means itself (bn_bwd_finalize's mode-1 arithmetic, no separate launch) and block 0 writes
// dgamma / dbeta.
struct BnRawSums {
  const float* gamma;
  const float* beta;
  float* dgamma;
  float* dbeta;
  float inv_count;
};
int bn_bwd_apply(__nv_bfloat16* dy, const __nv_bfloat16* X, long long M, int C, BnFold f, const float* m1,
                 const float* m2, cudaStream_t st, __nv_bfloat16* dst = nullptr, const BnRawSums* raw = nullptr,
                 int reverse = 0);   // reverse: rows from the last to the first (the executor alternates directions)

// dst[p, c] = G[p, c0+c] - corrA[c0+c] - xhat[p, c0+c]*corrB[c0+c]   for c in [0, nch)   (bf16 dense out)
int grad_fixup(const __nv_bfloat16* G, const __nv_bfloat16* X, int ld, long long M, int c0, int nch,
               const float* mean, const float* rstd, const float* corrA, const float* corrB, __nv_bfloat16* dst,
               cudaStream_t st);

// maxpool3x3/s2/p1 backward (gather through idx) + ReLU mask of the stem: dy0 bf16 [B,Hs,Ws,64];
// dsum += sum dy0 ; dsq += sum dy0*xhat0
int stem_pool_bwd(const __nv_bfloat16* dPool, const uint8_t* idx, const __nv_bfloat16* S0, int B, int Hs, int Ws,
                  BnFold f, __nv_bfloat16* dy0, float* dsum, float* dsq, cudaStream_t st);

// C[i,j] = sum_l A(i,l) * Bm(l,j) (+ bias[j]) (+ C if accumulate), fp32, arbitrary strides (tiny head GEMMs).
int sgemm_strided(int M, int N, int K, const float* A, long long a_i, long long a_l, const float* Bm, long long b_l,
                  long long b_j, const float* bias, float* C, long long c_i, long long c_j, cudaStream_t st);
// out[j] = sum_i A[i*ld + j]
int column_sum(const float* A, int rows, int cols, long long ld, float* out, cudaStream_t st);
// out[0] = scale * sum_i v[i]
int sum_scale(const float* v, int n, float scale, float* out, cudaStream_t st);

// fp32 OIHW master weights -> bf16 GEMM operand layouts (one launch over a job table).
enum RepackType { RP_1x1_FWD = 0, RP_1x1_DGRAD = 1, RP_3x3_FWD = 2, RP_3x3_DGRAD = 3, RP_STEM_FWD = 4,
                  // 3x3 stride-2 pad-1 weights as the 2x2-tap stride-1 operand over the 2x2 space-to-depth input
                  // (resnet_ops.cuh s2d_bn_relu): dst[(sy,sx)][n][(py*2+px)*K + c] = W[n][c][2sy+py-1][2sx+px-1] (0 outside)
                  RP_3x3S2_FWD = 5,
                  // its data-gradient operand (a plain 2x2-tap pad-0 conv over dOut producing the space-to-depth
                  // gradient): dst[(1-sy,1-sx)][(py*2+px)*K + c][n] = W[n][c][2sy+py-1][2sx+px-1]
                  RP_3x3S2_DGRAD = 6 };
struct RepackJob {
  long long src_off;   // into the flat fp32 parameter buffer
  long long dst_off;   // into the bf16 operand arena (elements)
  int type, N, K, pad;
};
int repack_weights(const float* params, __nv_bfloat16* arena, const RepackJob* jobs_dev, int n_jobs,
                   long long max_elems, cudaStream_t st);

}  // namespace rxb
