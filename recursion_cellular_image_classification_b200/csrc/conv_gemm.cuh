// conv_gemm.cuh — parameter blocks and host launchers of the tcgen05 implicit-GEMM convolution kernels.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace rxb {

// Pixel-space tiling: a GEMM M-tile is 128 pixels = tb images x th rows x tw columns (powers of two),
// fetched by ONE 4-D TMA box (channels, x, y, image); shifting the box by a filter tap is the implicit
// im2col, and TMA's out-of-bounds zero fill is the convolution's zero padding.
struct PixelTiling {
  int W, H, B;
  int tw_log2, th_log2, tb_log2;
  int tiles_x, tiles_y, tiles_b;
  int x_step;   // distance between tile origins in x: 1 << tw_log2, or tw - (taps_x - 1) when the M tile carries its x halo
  int rev;      // 0, or the tile count: tiles are then visited from the last to the first (GemmParams::reverse)
};
PixelTiling make_tiling(int B, int H, int W);       // wide tiles (tw as large as possible)
PixelTiling make_tiling_tall(int B, int H, int W);  // tw <= 8: tall tiles for the row-halo tap reuse
bool conv_dgrad3x3_wgrad_fusable(int B, int H, int W);   // GemmParams::wg_dW may be given for a 3x3 (32 -> 128) dgrad

enum EpiMode {
  EPI_STORE = 0,     // out[p, c_off+n] = bf16(acc) ; optional per-channel sum / sum-of-squares of the stored value
  EPI_DGRAD_BN = 1,  // dy = acc * [x*es+et > 0] ; ch_sum += sum(dy) ; out per out_mode (G modes store es*dy)
};
enum OutMode { OUT_DY = 0, OUT_G_WRITE = 1, OUT_G_ACCUM = 2 };

// Fused BatchNorm fold for the A prologue: when sum != nullptr the kernel derives relu(x*scale+shift)'s scale/shift
// itself from the per-channel batch sums (training) or running statistics (eval) - the arithmetic of bn_prep_kernel -
// and CTA (0,0) stores the fold (scale, shift, mean, rstd) for the backward kernels and updates the running statistics.
struct BnPrepArgs {
  const float* sum;
  const float* sumsq;
  const float* gamma;
  const float* beta;
  float* rmean;
  float* rvar;
  float count, eps, momentum;
  int training;
  float *f_scale, *f_shift, *f_mean, *f_rstd;
};

// Fused weight gradient (GemmParams::wg_dW) only: the consumer BatchNorm's backward reductions folded into the kernel
// tail instead of a bn_bwd_finalize launch.  Everything bn_bwd_finalize derives is LINEAR in (sum dy, W.dW), so every
// CTA contributes its own partial sums: t = sum bf16(W)*dW_cta per channel (the CTA's weight-gradient partial is in
// its registers on the way out) and s = its sum(dy).
//   mode 1 (concat BatchNorm, 1x1): dbeta += s, dgamma += q, corrA += scale*s/count, corrB += scale*q/count with
//          q = rstd*(raw - mean*s), raw = (t - bf16(shift)*s)/bf16(scale)  (degenerate channels: raw = direct sum dy*x)
//   mode 2 (bottleneck BatchNorm, 3x3): ch_sumsq += t for the non-degenerate channels (direct sum dy*x for the others,
//          as before); bn_bwd_apply's raw mode turns (ch_sum, ch_sumsq) into the means itself
struct BnTailArgs {
  int mode;             // 0: none
  const float* W;       // fp32 OIHW weights of the fused convolution (mode 1; mode 2 uses the bf16 panel in shared memory)
  const float* mean;    // fold of the consumer BatchNorm, per N channel (mode 1)
  const float* rstd;
  float inv_count;
  float* dgamma;
  float* dbeta;
  float* corrA;
  float* corrB;
};

// EPI 4 only (optional): the exact gradient of the convolution's 32 output channels is not read from a dense tensor
// but derived from the block's concat buffers on the way in (what grad_fixup_kernel wrote before):
//   dZ[p][c] = G[p][c0+c] - corrA[c0+c] - xhat[p][c0+c]*corrB[c0+c] = G + kb*X + kc,  kb = -rstd*corrB, kc = mean*rstd*corrB - corrA
// Every A stage then holds TWO full-halo boxes (the G slice and the X slice, strided 64-byte runs of the concat tensors)
// and the epilogue group turns the G box into dZ in place one tile ahead (packed bf16: t = fma(x, kb, kc), dZ = g + t;
// pixels outside the image keep TMA's zeros = the convolution's padding).
struct FixupArgs {
  const void* G;        // bf16 concat gradient [B,H,W,ld]
  const void* X;        // bf16 concat activations [B,H,W,ld]
  long long ld;
  int c0;               // first channel of the slice
  const float* mean;    // per concat channel (index c0 + c)
  const float* rstd;
  const float* corrA;
  const float* corrB;
};

struct GemmParams {
  int B, H, W;      // pixel space shared by A and the output (stride-1 convolutions)
  int n_total;      // valid N (multiple of 32)
  int taps_x, taps_y, pad_x, pad_y;
  int cin;          // valid A channels per tap
  int epi_mode;
  int out_mode;
  int do_stats;
  float* ch_sum;    // [n_total] (EPI_STORE: of the channel range being written; DGRAD: sum dy)
  float* ch_sumsq;  // [n_total] (EPI_STORE: sum of squares; DGRAD: sum(dy*x), written ONLY for the channels
                    //  bn_degenerate(e_gamma, e_beta) flags - for all others it follows from W.dW, see bn_bwd_finalize)
  // prologue (A := relu(A*scale + shift)), indexed by A channel
  const float* scale;
  const float* shift;
  BnPrepArgs prep;  // alternative to scale/shift: fold computed in the kernel (prep.sum != nullptr)
  // EPI_DGRAD_BN: folded BatchNorm of the consumer, per N channel
  const float* e_scale;
  const float* e_shift;
  const float* e_gamma;   // optional (with e_beta and ch_sumsq): BatchNorm weight / bias per N channel; channels that
  const float* e_beta;    // bn_degenerate() flags get direct sum(dy) / sum(dy*x) reductions in the epilogue
  // (also for the dense layers' 3x3, cin 32 -> n_total 128, when conv_dgrad3x3_wgrad_fusable(): fp32 OIHW [32][128][3][3])
  // EPI_DGRAD_BN of a 1x1 convolution with cin <= 128 (optional): fp32 [cin][n_total] = the convolution's OIHW weight
  // gradient, dW[k][c] += sum_p A[p][k] * relu(X[p][c]*e_scale[c] + e_shift[c]) (fold operands rounded to bf16 like
  // the forward prologue), accumulated by the same kernel from the tiles it already holds (conv_gemm.cu, EPI 3)
  float* wg_dW;
  int reverse;          // 1: walk the pixel tiles from the last to the first.  Consecutive kernels of the executor alternate
                        // direction, so each starts on the part of its input the previous one touched last (still in L2)
  BnTailArgs tail;      // with wg_dW: the BatchNorm-backward reductions that follow, folded into the tail (mode != 0)
  FixupArgs fix;        // with wg_dW of a 3x3 (EPI 4): dOut derived from the concat buffers on load (fix.G != nullptr)
  // ---- filled by launch_conv_gemm
  PixelTiling t;
  int n_tiles, bn, kb_per_tap;
  int halo;         // 1: an A stage holds th+taps_y-1 image rows; row taps are descriptor offsets into it
                    // 2: one box with the full x and y halo, every tap a descriptor offset
                    // 3: x-merged: the M tile is 16 wide INCLUDING its x halo, the taps_x taps of a filter row are one
                    //    N = taps_x*Cout MMA and the epilogue adds the neighbours' partial sums with warp shuffles
  int rows_a;       // pixel rows of one A stage (128, or (th+taps_y-1)*tw with halo)
  int stages;       // A (and streamed B) pipeline depth
  int n_stg;        // output staging buffers (stores: 1 or 2; dgrad: 2..4 activation-tile buffers staged in place)
  int b_resident;   // 1: the whole [taps][k-blocks] weight panel of the N tile is loaded once per CTA
  int n_acc;        // TMEM accumulator stages of kAccStride columns: 2, or 4 for the narrow store epilogue
  int mma_stats;    // 1: per-channel sums of the stored tile are accumulated by tcgen05.mma over the staging buffer
  int col_narrow;   // narrow store epilogue (bn 32 or 64): 1 = statistics by a pass of the group over the staged tile
                    // (like the 128-wide epilogue), 0 = warp-shuffle column sums in the epilogue
  unsigned long long* dbg;  // development: per-role clock64 timeline of CTA (0,0) (RXB_DBG_TIMELINE=1), else nullptr
};

// A: bf16 activation [B,H,W,ldA] (first `cin` channels used per tap); Wt: bf16 [taps][n_total][cin].
// out: bf16 [B,H,W,ldc], written (or read-modified-written) at channels c_off..c_off+n_total.
// X (EPI_DGRAD_BN): the activation the consumer's BatchNorm saw, bf16 [B,H,W,ldx], channels 0..n_total.
// bk = 64 (128B swizzle) or 32 (64B swizzle).
int launch_conv_gemm(GemmParams p, const void* A, long long ldA, const void* Wt, void* out, long long ldc,
                     int c_off, const void* X, long long ldx, int bk, bool prologue, cudaStream_t stream);

struct WgradParams {
  PixelTiling t;
  int taps_x, taps_y, pad_x, pad_y;
  int cin;              // A channels per tap
  int bkc;              // channels per A box: 64 or 32
  int boxes_per_tap;    // ceil(cin / bkc)
  int boxes_per_chunk;  // 128 / bkc
  int n_chunks;         // accumulator groups (128 A-rows each) in total
  int chunks_per_cta;   // <= 512 / n
  int n;                // Cout (32, or a multiple of 64 up to 256)
  int n_off;            // first dOut channel of this launch (N tiling for Cout > 256)
  int pix_tiles_per_cta;
  int prologue;
  int shift_dout;       // 1: multi-tap with cin <= 128: A' tile loaded+transformed ONCE per pixel tile, dOut shifted per tap
  const float* scale;
  const float* shift;
  float* dW;            // fp32, torch OIHW [Cout_total][cin_w][taps_y_w][taps_x_w], atomically accumulated
  int cout_total;
  int w_mode;           // 0: generic OIHW (k -> (tap, channel)) ; 1: space-to-depth stem (7x7 stride 2, 6 ch) ;
                        // 2: 3x3 stride-2 weights behind the 2x2-tap space-to-depth conv (cin = 4 x real channels)
                        // 3: multi-tap, TAP-MAJOR fp32 scratch dW[tap][cout_total][cin] (zeroed by the caller), added by
                        //    bulk L2 reduce-adds of contiguous channel runs instead of 4-byte atomics; a finish kernel
                        //    (resnet_ops wgrad_finish) transposes it into OIHW
  int a_halo;           // 1: multi-tap, bkc*taps_x == 128, no prologue: ONE full-halo A box per pixel tile; chunk = filter
                        //    row ty whose taps_x taps are M atoms one pixel row apart (the stem's 4x4 taps)
  unsigned long long* dbg;  // development timeline of CTA (0,0) (RXB_DBG_TIMELINE), else nullptr
  int bulk_out;         // 1: result staged in shared memory and added to dW by cp.reduce.async.bulk (else fp32 atomics)
  int bulk_bufs;        // staging buffers for the 1x1 bulk path (1 or 2)
};
// A: bf16 [B,H,W,ldA]; dOut: bf16 [B,H,W,ldD].
int launch_conv_wgrad(WgradParams p, const void* A, long long ldA, const void* dOut, long long ldD,
                      cudaStream_t stream);

}  // namespace rxb
