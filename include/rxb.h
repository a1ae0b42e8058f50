/*
 * rxb.h — C ABI of librxb.so: the B200 (sm_100a) hot path of the RxRx1 cellular image classifier.
 *
 * The reference (antoinecollas/recursion-cellular-image-classification) is pure Python and has no
 * FFI layer; its "plugin API" for this path is five Python callables.  The entry
 * points below are what those callables bind to (via ctypes).  Each one cites the
 * reference code it replaces.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes only.  No torch / C++ types cross the boundary.
 *   - return 0 on success, a negative rxb_status on failure; rxb_last_error() gives the message
 *     (thread-local).  Nothing throws across the boundary.
 *   - every pointer is a DEVICE pointer owned by the caller unless the name says host; the library
 *     never frees or retains caller memory beyond the lifetime of a plan it was bound to.
 *   - every call takes a cudaStream_t (as void*) and is asynchronous on it.
 *   - no hidden device allocations: workspaces are sized by *_workspace_bytes() and caller-provided.
 *   - there is no CPU fallback: on a machine without an sm_100 device the calls fail loudly.
 */
#ifndef RXB_H_
#define RXB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rxb_stream_t; /* cudaStream_t */

enum rxb_status {
  RXB_OK = 0,
  RXB_ERR_INVALID = -1,   /* bad argument */
  RXB_ERR_CUDA = -2,      /* CUDA runtime / driver error */
  RXB_ERR_UNSUPPORTED = -3,
  RXB_ERR_NO_DEVICE = -4  /* no sm_100 device: there is no fallback */
};
/* Collectives are not part of this library: the gradient all-reduce, the statistics all-reduce and the test-time
 * all-gather are issued by the host side through torch.distributed (NCCL) between the phases the executor exposes
 * (rxb_dn121_train_step's `phase`), see recursion_cellular_image_classification_b200/parallel.py. */

int rxb_version(void);
const char* rxb_last_error(void);
/* 0 if cuda:current is an sm_100 part, RXB_ERR_NO_DEVICE otherwise. */
int rxb_check_device(void);

/* ------------------------------------------------------------------------------------------------
 * Family 1a — per-experiment channel statistics.
 * Replaces the hot loop of compute_mean_std(), compute_stats_experiments.py:13-20 (count / sum(x) /
 * sum(x^2) per channel of x = u8/255), with exact integer accumulation.
 *   imgs    u8 [n, C, H, W] planar — what dataloader.py:141-146 decodes and compute_stats_experiments.py:15 reads
 *           (one file per channel).  `layout` must be RXB_LAYOUT_NCHW; any other value is RXB_ERR_UNSUPPORTED.
 *   exp_id  i32[n], experiment slot of every image, 0 <= exp_id < n_exp
 *   sum, sumsq, count  u64[n_exp, C], ACCUMULATED into (caller zeroes them once); count is in pixels
 *               (compute_stats_experiments.py:21 multiplies the image count by 512*512).
 * H*W must be a multiple of 16.
 */
enum rxb_layout { RXB_LAYOUT_NCHW = 0 };
int rxb_stats_accumulate(const uint8_t* imgs, const int32_t* exp_id, int64_t n, int H, int W, int C,
                         int layout, int n_exp, unsigned long long* sum, unsigned long long* sumsq,
                         unsigned long long* count, rxb_stream_t stream);
/* mean/std of x/255 in f64 (compute_stats_experiments.py:22-23).  If pre_mean/pre_std (f64[n_exp,C],
 * device) are non-NULL the statistics are those of (x/255 - pre_mean)/pre_std, i.e. the reference's
 * verification pass (compute_stats_experiments.py:16-17, 51-57), derived from the same sums. */
int rxb_stats_finalize(const unsigned long long* sum, const unsigned long long* sumsq,
                       const unsigned long long* count, int n_exp, int C, const double* pre_mean,
                       const double* pre_std, double* mean, double* std, rxb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Family 1b — fused D4 augmentation + crop + per-experiment normalisation loader.
 * Replaces ImagesDS._transform, dataloader.py:128-139 (albumentations VerticalFlip, HorizontalFlip,
 * rotation restricted to multiples of 90 degrees, Random/CenterCrop, Normalize) on already-decoded
 * u8 planes, and the fp32 H2D copy of train.py:44.
 *   src       u8 [n_src, 6, H, W] planar
 *   src_idx   i32[B]   which source image each output image is cut from
 *   exp_id    i32[B]   row of norm_m / norm_d to use
 *   aug_code  u8[B]    bit0 vflip, bit1 hflip, bits2-3 k (number of 90-degree CCW turns, applied
 *                      after the flips like dataloader.py:42-46), bit4 = reference-compatible rotation
 *                      (the cv2.warpAffine gather about (W/2,H/2) with BORDER_REFLECT_101)
 *   crop_yx   i32[B,2] top-left of the (Ho,Wo) crop in the augmented image (RandomCrop/CenterCrop)
 *   norm_m    f32[n_exp,6] = f32(mean)*255 ;  norm_d f32[n_exp,6] = 1/(f32(std)*255)   (A.1)
 *   dst       format RXB_OUT_F32_NCHW  : f32  [B,6,Ho,Wo]   = (f32(x) - m) * d, bit-exact vs numpy
 *             format RXB_OUT_BF16_NHWC8: bf16 [B,Ho,Wo,8]   channels 6,7 = 0
 *             format RXB_OUT_BF16_S2D32: bf16 [B,Ho/2,Wo/2,32] 2x2 space-to-depth of NHWC8
 *                                        (channel = (y&1)*16 + (x&1)*8 + c) — the stem conv's input
 * H and W must be equal (square, D4) and multiples of 16; Ho, Wo even for S2D32.
 */
enum rxb_out_format { RXB_OUT_F32_NCHW = 0, RXB_OUT_BF16_NHWC8 = 1, RXB_OUT_BF16_S2D32 = 2 };
int rxb_load_norm_aug(const uint8_t* src, int64_t n_src, int H, int W, const int32_t* src_idx,
                      const int32_t* exp_id, const uint8_t* aug_code, const int32_t* crop_yx,
                      const float* norm_m, const float* norm_d, int n_exp, void* dst, int B, int Ho,
                      int Wo, int out_format, rxb_stream_t stream);

/* Baseline grayscale JPEG decode on the device.  Replaces ImagesDS._load_from_buffer,
 * dataloader.py:141-146 (cv2.imdecode(buffer, -1) per channel file; files written by png_to_jpeg.py:11-15).
 * Bit-identical to cv2.imdecode / libjpeg's default decoder (Huffman, accurate integer IDCT).
 *   blob     u8  device: the n files' bytes back to back
 *   begin, end  i64[n] device: file i is blob[begin[i] .. end[i]).  For files packed back to back pass an
 *            offsets array o[n+1] as begin = o, end = o + 1; any subset / order of files works the same way
 *   dst      u8 [n,H,W] device — e.g. [B,6,H,W] planar, the input of rxb_stats_accumulate / rxb_load_norm_*
 *   status   i32[n] device: 0 ok; 1 not a JPEG / truncated headers; 2 unsupported (progressive, lossless,
 *            arithmetic, 12-bit, multi-component); 3 missing or malformed table; 4 frame size != (H,W);
 *            5 corrupt entropy-coded data.  A file with a non-zero status leaves (part of) its plane unwritten.
 *   workspace  device, 16-byte aligned, rxb_jpeg_decode_workspace_bytes(n,H,W) bytes (the quantised coefficients,
 *            128 bytes per 8x8 block), or NULL.  With a workspace all 32 lanes of a file's warp decode
 *            (speculative, self-synchronising subsequences); without one a single lane per file decodes — same
 *            result, several times slower.  Files with restart intervals always take the single-lane kernel. */
size_t rxb_jpeg_decode_workspace_bytes(int n, int H, int W);
int rxb_jpeg_decode_gray(const uint8_t* blob, const int64_t* begin, const int64_t* end, int n, int H, int W,
                         uint8_t* dst, int32_t* status, void* workspace, size_t workspace_bytes,
                         rxb_stream_t stream);

/* Arbitrary-angle variant of the loader: the reference's full train transform
 * VerticalFlip -> HorizontalFlip -> ShiftScaleRotate(rotate_limit=180) -> RandomCrop -> Normalize
 * (dataloader.py:42-48, 128-139).  ShiftScaleRotate is cv2.warpAffine(img, M, (W,H), INTER_LINEAR,
 * BORDER_REFLECT_101) on u8; the kernel restates OpenCV's fixed-point arithmetic (1/32-pixel
 * coordinates, 10-bit weights, round-half-up) so the u8 gather is bit-identical to OpenCV's.
 *   flip_code u8[B]     bit0 vflip, bit1 hflip (applied to the source before the warp)
 *   M         f64[B,2,3] FORWARD matrices, exactly what cv2.warpAffine receives (host-side
 *                        cv2.getRotationMatrix2D((W/2,H/2), angle, 1.0)); inverted on the device the
 *                        way cv::warpAffine does
 *   the other arguments are rxb_load_norm_aug's; H and W need not be equal or multiples of 16. */
int rxb_load_norm_affine(const uint8_t* src, int64_t n_src, int H, int W, const int32_t* src_idx,
                         const int32_t* exp_id, const uint8_t* flip_code, const double* M,
                         const int32_t* crop_yx, const float* norm_m, const float* norm_d, int n_exp,
                         void* dst, int B, int Ho, int Wo, int out_format, rxb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Family 4 — test-time averaging, plate-group masking, rescale and greedy assignment.
 * Replaces test.py:27 (softmax), :42-46 (mask + rescale) and :48-56 (greedy loop).
 */
/* probs[n,c] = rescale(mask(mean_v softmax(logits[v,n,:]))).  logits f32 [V,N,C].  plate i32[N];
 * group_col i32[C] = plate_groups[:, experiment_type] (main.py:157-166).  A class is kept for row n
 * iff group_col[c] == plate[n] (test.py:42-45).  rescale: row /= row-sum, zero rows unchanged.
 * plate == NULL skips the mask. */
int rxb_tta_softmax_avg_mask(const float* logits, int V, int N, int C, const int32_t* plate,
                             const int32_t* group_col, float* probs, rxb_stream_t stream);
/* In place on f32 probabilities that were produced elsewhere: preds[mask]=0 ; preds=rescale(preds). */
int rxb_mask_rescale(float* preds, int N, int C, const int32_t* plate, const int32_t* group_col,
                     rxb_stream_t stream);
size_t rxb_greedy_assign_workspace_bytes(int N, int C);
/* The loop of test.py:48-56, bit-exact (numpy's float32 pairwise row sums are reproduced).
 * preds f32[N,C] is read only; result i32[N].  workspace from rxb_greedy_assign_workspace_bytes,
 * N <= 8*SM count. */
int rxb_greedy_assign(const float* preds, int N, int C, int32_t* result, void* workspace,
                      rxb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Family 3 — classifier head loss.  Replaces nn.CrossEntropyLoss (train.py:37) forward + backward.
 * logits f32 [B,C] (row stride ld), target i64[B].  loss_rows f32[B] (per-sample NLL); dlogits f32
 * [B,C] (row stride ld) = (softmax - onehot) * grad_scale (pass 1/global_batch for the mean loss);
 * dlogits may be NULL (eval).
 */
int rxb_softmax_ce(const float* logits, int ld, const int64_t* target, int B, int C, float* loss_rows,
                   float* dlogits, float grad_scale, rxb_stream_t stream);

/* SGD with momentum / nesterov / weight decay, torch.optim.SGD semantics (main.py:89-93):
 * g = grad*grad_scale + wd*p ; m = mu*m + g ; p -= lr * (nesterov ? g + mu*m : m).
 * (momentum buffers start at 0, which equals torch's first-step "m = g".) */
int rxb_sgd_step(float* p, const float* grad, float* mom, int64_t n, float lr, float mu, float wd,
                 int nesterov, float grad_scale, rxb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Family 2 — convolutions as tcgen05 implicit GEMMs (TMA-fed, TMEM accumulators).
 * Low-level entry (used by tests and by the executor below):
 *   out[p, c_off + n] = sum_{tap,k} A(p shifted by tap)[k] * Wt[tap][n][k]        (bf16 in, fp32 accum)
 * with an optional per-channel affine+ReLU applied to A on the way in (pre-activation BatchNorm,
 * torchvision densenet _DenseLayer: norm -> relu -> conv), zero padding applied AFTER it.
 */
typedef struct rxb_conv_desc {
  int B, H, W;        /* output (== input, stride 1) spatial size and batch */
  int Cin;            /* channels read per tap (multiple of 8) */
  int ldA;            /* channel stride of the input tensor  (>= Cin; concat buffers are wider) */
  int Cout;           /* 16..256, multiple of 16 */
  int ldC;            /* channel stride of the output tensor */
  int c_off;          /* first output channel (concat-by-offset) */
  int taps_y, taps_x; /* 1x1, 3x3, 4x4 (space-to-depth stem) */
  int pad_y, pad_x;   /* input row = y + ty - pad_y */
  int prologue;       /* 1: A := relu(A*scale[k] + shift[k]) */
  int stats;          /* 1: accumulate per-out-channel sum and sum of squares (of the bf16-rounded out) */
} rxb_conv_desc;
int rxb_conv_fwd(const rxb_conv_desc* d, const void* A_bf16, const void* W_bf16 /*[taps][Cout][Cin]*/,
                 const float* scale, const float* shift, void* out_bf16, float* ch_sum, float* ch_sumsq,
                 rxb_stream_t stream);
/* Data gradient fused with the ReLU / BatchNorm backward of the layer that produced the conv's input
 * (torch autograd runs conv dgrad, threshold_backward and batch_norm_backward as separate kernels):
 *   acc[p, k] = sum_{tap,n} dOut(p shifted)[n] * Wt[tap][k][n]      (Wt: the dgrad operand layout, [taps][k][n])
 *   dy        = acc * [X[p,k]*bn_scale[k] + bn_shift[k] > 0]
 *   sum_dy[k] += sum_p dy
 *   out_mode 0: out = dy ; 1: out = bn_scale*dy ; 2: out += bn_scale*dy      (bf16 [B,H,W,ldC], channels 0..Cout;
 *   mode 2 is an L2 reduce-add issued by TMA: the running gradient is never loaded into the SM)
 * In the descriptor Cin is the contraction size (channels of dOut per tap), Cout (>= 64) the width of the result.
 * The second BatchNorm-backward reduction, sum_p dy*X, is not reduced here: see rxb_bn_sum_dyx_from_wdw. */
int rxb_conv_dgrad_bn(const rxb_conv_desc* d, const void* dOut_bf16, const void* Wt_bf16, const void* X_bf16,
                      int ldX, const float* bn_scale, const float* bn_shift, int out_mode, void* out_bf16,
                      float* sum_dy, rxb_stream_t stream);
/* The same launch with the BatchNorm's own weight and bias (f32[Cout]) given: channels with |gamma| < 1e-3 or
 * |gamma| < 0.05*|beta| ("degenerate": the W.dW identity below divides by gamma*rstd and cancels catastrophically
 * there) get sum_dy AND sum_dyx[k] += sum_p dy*X reduced directly in the epilogue, in fp32 from the unrounded dy;
 * sum_dyx is left untouched for every other channel.  This is what the executor calls. */
int rxb_conv_dgrad_bn_ex(const rxb_conv_desc* d, const void* dOut_bf16, const void* Wt_bf16, const void* X_bf16,
                         int ldX, const float* bn_scale, const float* bn_shift, const float* bn_gamma,
                         const float* bn_beta, int out_mode, void* out_bf16, float* sum_dy, float* sum_dyx,
                         rxb_stream_t stream);
/* The same launch for a 1x1 convolution (taps 1x1, Cin <= 128) that ALSO accumulates that convolution's weight
 * gradient from the tiles it already holds in shared memory (torch autograd: a separate convolution_backward weight
 * kernel that re-reads both tensors):
 *   dW[k][c] += sum_p dOut[p,k] * A'[p,c],  A' = bf16(relu(X[p,c]*bf16(bn_scale[c]) + bf16(bn_shift[c])))
 * dW: f32 [Cin][Cout] = the forward convolution's OIHW weight gradient (its output channels are this launch's
 * contraction channels).  A' is exactly what the forward prologue of rxb_conv_fwd fed the convolution, and the ReLU
 * mask of dy becomes the test A' > 0.  bn_gamma / bn_beta / sum_dyx may be NULL together.
 * Also accepted: the dense layers' 3x3 data gradient (taps 3x3, pad 1, Cin 32 -> Cout 128, H > 8, W > 4), with
 * dW f32 [32][128][3][3] (OIHW of the forward 3x3 convolution) - its full-halo dOut box serves both contractions. */
int rxb_conv_dgrad_bn_wgrad(const rxb_conv_desc* d, const void* dOut_bf16, const void* Wt_bf16, const void* X_bf16,
                            int ldX, const float* bn_scale, const float* bn_shift, const float* bn_gamma,
                            const float* bn_beta, int out_mode, void* out_bf16, float* sum_dy, float* sum_dyx,
                            float* dW, rxb_stream_t stream);
/* The 3x3 form of rxb_conv_dgrad_bn_wgrad with dOut NOT given as a dense tensor but derived on load from a DenseNet
 * block's concat buffers (what a separate element-wise pass would write first):
 *   dOut[p][c] = G[p][c0+c] + kb[c]*Xc[p][c0+c] + kc[c],  kb = -rstd[c0+c]*corrB[c0+c],  kc = mean*rstd*corrB - corrA
 * (the exact gradient G - corrA - xhat*corrB of the convolution's 32 output channels under the lazy BatchNorm
 * backward), evaluated in packed bf16 in shared memory: t = bf16(fma(x, bf16(kb), bf16(kc))), dOut = bf16(g + t).
 * G, Xc: bf16 [B,H,W,ld]; mean, rstd, corrA, corrB: f32 per concat channel.  d->Cin = 32, d->Cout = 128, 3x3, pad 1. */
int rxb_conv_dgrad3x3_bn_wgrad_fixup(const rxb_conv_desc* d, const void* G_bf16, const void* Xc_bf16, int ld, int c0,
                                     const float* mean, const float* rstd, const float* corrA, const float* corrB,
                                     const void* Wt_bf16, const void* X_bf16, int ldX, const float* bn_scale,
                                     const float* bn_shift, int out_mode, void* out_bf16, float* sum_dy, float* dW,
                                     rxb_stream_t stream);
/* sum_dyx[c] = sum_p dy[p,c]*X[p,c] for the BatchNorm in front of a convolution, from that convolution's weights and
 * finished weight gradient (fp32 OIHW [Cout][Cin][taps]):  with z = bn_scale*x + bn_shift,
 *   sum_p dy*z = sum_{k,tap} W[k][c][tap]*dW[k][c][tap]   (both equal sum_p dL/dA' * A', A' = relu(z)),
 * so sum_dyx = (W.dW - bn_shift*sum_dy) / bn_scale.  Replaces one of torch's batch_norm_backward reductions. */
int rxb_bn_sum_dyx_from_wdw(const float* W, const float* dW, int Cout, int Cin, int taps, const float* bn_scale,
                            const float* bn_shift, const float* sum_dy, float* sum_dyx, rxb_stream_t stream);
/* dW[tap][n][k] += sum_p dOut[p, n] * A'(p shifted by tap)[k]  (fp32 atomics into dW). */
int rxb_conv_wgrad(const rxb_conv_desc* d, const void* A_bf16, const float* scale, const float* shift,
                   const void* dOut_bf16, int ldD, float* dW, rxb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * DenseNet-121 executor (6-channel stem, growth 32, blocks 6/12/24/16, 1108 classes):
 * forward, loss, backward and the SGD update of one data-parallel rank, enqueued natively.
 * Replaces TwoSitesNN.forward (models.py:41-57) with the north star's DenseNet-121 trunk, the
 * train step of train.py:44 and the optimizer of main.py:89-93.
 */
typedef struct rxb_dn121 rxb_dn121; /* opaque */
typedef struct rxb_dn121_config {
  int B;            /* images per rank */
  int H, W;         /* input size (multiple of 32) */
  int num_classes;  /* 1108 */
  float bn_eps;     /* 1e-5 */
  float bn_momentum;/* 0.1 */
} rxb_dn121_config;
/* number of fp32 parameters and of fp32 buffer elements (running mean/var), torchvision order */
int64_t rxb_dn121_param_count(const rxb_dn121_config* cfg);
int64_t rxb_dn121_buffer_count(const rxb_dn121_config* cfg);
size_t rxb_dn121_workspace_bytes(const rxb_dn121_config* cfg, int training);
/* params/grads/momentum: fp32[param_count] flat, torchvision parameter order (features.conv0.weight,
 * features.norm0.weight, ...), conv weights in torch OIHW layout.  buffers: fp32[buffer_count]
 * (running_mean, running_var per BN in module order).  All device pointers, bound for the plan's life. */
int rxb_dn121_create(const rxb_dn121_config* cfg, float* params, float* grads, float* momentum,
                     float* buffers, void* workspace, size_t workspace_bytes, int training,
                     rxb_dn121** out);
void rxb_dn121_destroy(rxb_dn121* net);
/* Re-derive the bf16 GEMM operand copies from the fp32 master parameters (after an external update). */
int rxb_dn121_sync_weights(rxb_dn121* net, rxb_stream_t stream);
/* input: bf16 S2D32 [B,H/2,W/2,32] from rxb_load_norm_aug.  logits_out f32 [B,num_classes] (may be
 * NULL).  training=1 uses batch statistics and updates running stats; 0 uses running stats. */
int rxb_dn121_forward(rxb_dn121* net, const void* input_s2d, float* logits_out, int training,
                      rxb_stream_t stream);
/* forward (training) + mean CE loss over `global_batch` samples + backward into grads (overwritten).
 * loss_out: f32[1] device, sum over this rank's samples of NLL / global_batch.
 * phase: -1 = everything; otherwise 0..rxb_dn121_num_phases()-1 runs one slice so the caller can
 * interleave collectives: phase 0 = forward+loss+head backward, later phases walk the dense blocks
 * backwards; after phase p the gradient range rxb_dn121_phase_grad_range(p) is final. */
int rxb_dn121_num_phases(void);
int rxb_dn121_phase_grad_range(const rxb_dn121* net, int phase, int64_t* begin, int64_t* end);
int rxb_dn121_train_step(rxb_dn121* net, const void* input_s2d, const int64_t* target,
                         int global_batch, float* loss_out, int phase, rxb_stream_t stream);
/* p -= lr * nesterov(grad*grad_scale + wd*p) on the flat buffers, then refresh the bf16 operands. */
int rxb_dn121_sgd(rxb_dn121* net, float lr, float mu, float wd, int nesterov, float grad_scale,
                  rxb_stream_t stream);
/* ------------------------------------------------------------------------------------------------
 * The reference's real model, evaluation mode: TwoSitesNN = torchvision ResNet-50 trunk with the
 * 6-channel stem, fc -> Identity (models.py:16-29), per-sample feature means of the image / negative-control /
 * positive-control thirds concatenated (models.py:44-53), and the BatchNorm1d -> Dropout -> Linear(6144,
 * size_features) -> ReLU -> BatchNorm1d -> Dropout -> Linear(size_features, num_classes) head (models.py:31-39).
 * Replaces TwoSitesNN.forward as test.py:23-27 calls it (model.eval(): running statistics, Dropout = identity), so a
 * checkpoint trained with the reference (main.py:147) can be served natively, and the reference's train step
 * (rxb_rn50_train_step / rxb_rn50_sgd below).  Convolutions are the tcgen05 implicit GEMMs above (stride-2 3x3 as a
 * 2x2-tap convolution over a space-to-depth pass, see csrc/resnet.cu).
 */
typedef struct rxb_rn50 rxb_rn50; /* opaque */
typedef struct rxb_rn50_config {
  int B;             /* samples per call */
  int G;             /* images per sample: 3 (train/val item) or 6 (test item); a multiple of 3 */
  int H, W;          /* image size, multiples of 4 (364 in the reference's training, 512 at test time) */
  int num_classes;   /* 1108 */
  int size_features; /* 1024 (models.py:11) */
  float bn_eps;      /* 1e-5 */
  float bn_momentum; /* 0.1 (training: running-statistics update) */
} rxb_rn50_config;
/* fp32 parameters in the reference model's named_parameters() order (base_nn.conv1.weight, base_nn.bn1.weight, ...,
 * mlp.6.bias; conv weights OIHW) and fp32 buffers (running_mean, running_var per BatchNorm in module order; the
 * int64 num_batches_tracked entries are not part of it). */
int64_t rxb_rn50_param_count(const rxb_rn50_config* cfg);
int64_t rxb_rn50_buffer_count(const rxb_rn50_config* cfg);
int64_t rxb_rn50_head_offset(const rxb_rn50_config* cfg);   /* first mlp.* parameter: the head is the flat buffer's tail */
size_t rxb_rn50_workspace_bytes(const rxb_rn50_config* cfg, int training);
/* grads / momentum (fp32[param_count]) are needed for training plans only. */
int rxb_rn50_create(const rxb_rn50_config* cfg, float* params, float* grads, float* momentum, float* buffers,
                    void* workspace, size_t workspace_bytes, int training, rxb_rn50** out);
void rxb_rn50_destroy(rxb_rn50* net);
/* Re-derive the bf16 GEMM operand copies from the fp32 parameters (after loading a checkpoint). */
int rxb_rn50_sync_weights(rxb_rn50* net, rxb_stream_t stream);
/* input: bf16 S2D32 [B*G, H/2, W/2, 32] from rxb_load_norm_aug (sample-major: the G images of a sample are
 * consecutive, thirds in the reference's order).  logits_out f32 [B, num_classes]. */
int rxb_rn50_forward(rxb_rn50* net, const void* input_s2d, float* logits_out, rxb_stream_t stream);
/* The reference's train step on this rank's B samples (train.py:37,44: zero_grad, forward in training mode,
 * CrossEntropy mean over `global_batch`, backward): BatchNorm2d / BatchNorm1d with batch statistics (running statistics
 * updated), Dropout(p) as multiplication by the caller's masks (f32 [B, 3*2048] and [B, size_features] holding 0 or
 * 1/(1-p): the caller owns the random stream, which is what makes parity testable), gradients of every parameter
 * into grads (overwritten).  loss_out f32[1]: this rank's share of the mean loss; logits_out f32[B,num_classes] or
 * NULL: the training-mode logits; feat_out f32[B*G,2048] or NULL: the trunk's pooled features (base_nn output). */
int rxb_rn50_train_step(rxb_rn50* net, const void* input_s2d, const int64_t* target, const float* drop_mask0,
                        const float* drop_mask1, int global_batch, float* loss_out, float* logits_out,
                        float* feat_out, rxb_stream_t stream);
/* nesterov SGD (main.py:89-93) on the flat buffers, then refresh the bf16 operands.  head_only = 1 updates the mlp.*
 * parameters alone: the reference's first two epochs with a pretrained trunk (train.py:46-58 freezes every child of
 * the model except 'mlp'). */
int rxb_rn50_sgd(rxb_rn50* net, float lr, float mu, float wd, int nesterov, float grad_scale, int head_only,
                 rxb_stream_t stream);

/* number of kernels the last forward/train_step/sgd call enqueued (bench.py's gpu_launches). */
int64_t rxb_launch_count(void);
void rxb_launch_count_reset(void);
/* Per-kernel-family timing for bench.py: while enabled every launch is bracketed by CUDA events on its
 * stream.  rxb_profile_collect synchronises, writes milliseconds and launch counts per category
 * and clears the record.  Categories (ncat >= 17): 0 stats, 1 loader, 2 dense-layer 1x1 forward, 3 dense-layer 1x1
 * data gradient, 4 dense-layer 1x1 weight gradient, 5 elementwise (other than 14-16), 6 head, 7 optimizer+repack,
 * 8 TTA/assignment, 9 dense-layer 3x3 forward, 10 3x3 data gradient, 11 3x3 weight gradient, 12 other forward-style
 * GEMMs (stem, transitions and their data gradient), 13 other weight gradients (stem, transitions),
 * 14 bn_bwd_apply, 15 grad_fixup, 16 bn_bwd_finalize. */
void rxb_profile_enable(int on);
int rxb_profile_collect(float* ms, long long* launches, int ncat);

#ifdef __cplusplus
}
#endif
#endif /* RXB_H_ */
