// api_conv.cu — C-ABI wrappers of the implicit-GEMM convolution kernels (rxb_conv_fwd / rxb_conv_wgrad).
#include "conv_gemm.cuh"
#include "elementwise.cuh"

namespace rxb {

int pick_bn(int n) { return n < 128 ? n : 128; }

}  // namespace rxb

extern "C" {

int rxb_conv_fwd(const rxb_conv_desc* d, const void* A_bf16, const void* W_bf16, const float* scale,
                 const float* shift, void* out_bf16, float* ch_sum, float* ch_sumsq, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(d && A_bf16 && W_bf16 && out_bf16, "rxb_conv_fwd: null pointer");
  RXB_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0, "rxb_conv_fwd: bad spatial size");
  RXB_CHECK_ARG(d->Cin > 0 && d->Cin % 8 == 0 && d->ldA >= d->Cin && d->ldA % 8 == 0, "rxb_conv_fwd: bad Cin/ldA");
  RXB_CHECK_ARG(d->Cout >= 32 && d->Cout % 32 == 0, "rxb_conv_fwd: Cout must be a multiple of 32");
  RXB_CHECK_ARG(d->ldC >= d->c_off + d->Cout && d->ldC % 8 == 0 && d->c_off % 8 == 0, "rxb_conv_fwd: bad ldC/c_off");
  RXB_CHECK_ARG(d->taps_x >= 1 && d->taps_y >= 1 && d->taps_x <= 8 && d->taps_y <= 8, "rxb_conv_fwd: bad taps");
  RXB_CHECK_ARG(!d->prologue || (scale && shift), "rxb_conv_fwd: prologue needs scale/shift");
  RXB_CHECK_ARG(!d->stats || (ch_sum && ch_sumsq), "rxb_conv_fwd: stats needs ch_sum/ch_sumsq");
  int rc = rxb_check_device();
  if (rc) return rc;
  const int bk = d->Cin <= 32 ? 32 : 64;
  if (d->prologue && bk != 64) return set_error(RXB_ERR_UNSUPPORTED, "rxb_conv_fwd: prologue needs Cin > 32");
  GemmParams p = {};
  p.B = d->B; p.H = d->H; p.W = d->W;
  p.n_total = d->Cout;
  p.taps_x = d->taps_x; p.taps_y = d->taps_y; p.pad_x = d->pad_x; p.pad_y = d->pad_y;
  p.cin = d->Cin;
  p.epi_mode = EPI_STORE;
  p.out_mode = OUT_DY;
  p.do_stats = d->stats;
  p.ch_sum = ch_sum ? ch_sum + d->c_off : nullptr;
  p.ch_sumsq = ch_sumsq ? ch_sumsq + d->c_off : nullptr;
  p.scale = scale;
  p.shift = shift;
  return launch_conv_gemm(p, A_bf16, d->ldA, W_bf16, out_bf16, d->ldC, d->c_off, nullptr, 0, bk, d->prologue != 0,
                          as_stream(stream));
}

int rxb_conv_dgrad_bn(const rxb_conv_desc* d, const void* dOut_bf16, const void* Wt_bf16, const void* X_bf16, int ldX,
                      const float* bn_scale, const float* bn_shift, int out_mode, void* out_bf16, float* sum_dy,
                      rxb_stream_t stream) {
  return rxb_conv_dgrad_bn_ex(d, dOut_bf16, Wt_bf16, X_bf16, ldX, bn_scale, bn_shift, nullptr, nullptr, out_mode,
                              out_bf16, sum_dy, nullptr, stream);
}

int rxb_conv_dgrad_bn_ex(const rxb_conv_desc* d, const void* dOut_bf16, const void* Wt_bf16, const void* X_bf16,
                         int ldX, const float* bn_scale, const float* bn_shift, const float* bn_gamma,
                         const float* bn_beta, int out_mode, void* out_bf16, float* sum_dy, float* sum_dyx,
                         rxb_stream_t stream) {
  return rxb_conv_dgrad_bn_wgrad(d, dOut_bf16, Wt_bf16, X_bf16, ldX, bn_scale, bn_shift, bn_gamma, bn_beta, out_mode,
                                 out_bf16, sum_dy, sum_dyx, nullptr, stream);
}

int rxb_conv_dgrad_bn_wgrad(const rxb_conv_desc* d, const void* dOut_bf16, const void* Wt_bf16, const void* X_bf16,
                            int ldX, const float* bn_scale, const float* bn_shift, const float* bn_gamma,
                            const float* bn_beta, int out_mode, void* out_bf16, float* sum_dy, float* sum_dyx,
                            float* dW, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(d && dOut_bf16 && Wt_bf16 && X_bf16 && bn_scale && bn_shift && out_bf16 && sum_dy,
                "rxb_conv_dgrad_bn: null pointer");
  RXB_CHECK_ARG((bn_gamma == nullptr) == (bn_beta == nullptr) && (bn_gamma == nullptr || sum_dyx != nullptr),
                "rxb_conv_dgrad_bn_ex: bn_gamma, bn_beta and sum_dyx come together");
  RXB_CHECK_ARG(d->Cin > 0 && d->Cin % 8 == 0 && d->ldA >= d->Cin && d->ldA % 8 == 0, "rxb_conv_dgrad_bn: bad Cin/ldA");
  RXB_CHECK_ARG(d->Cout >= 64 && d->Cout % 32 == 0 && d->ldC >= d->Cout && ldX >= d->Cout, "rxb_conv_dgrad_bn: bad Cout");
  RXB_CHECK_ARG(out_mode >= OUT_DY && out_mode <= OUT_G_ACCUM, "rxb_conv_dgrad_bn: bad out_mode");
  int rc = rxb_check_device();
  if (rc) return rc;
  GemmParams p = {};
  p.B = d->B; p.H = d->H; p.W = d->W;
  p.n_total = d->Cout;
  p.taps_x = d->taps_x; p.taps_y = d->taps_y; p.pad_x = d->pad_x; p.pad_y = d->pad_y;
  p.cin = d->Cin;
  p.epi_mode = EPI_DGRAD_BN;
  p.out_mode = out_mode;
  p.do_stats = 1;
  p.ch_sum = sum_dy;
  p.ch_sumsq = sum_dyx;
  p.e_scale = bn_scale;
  p.e_shift = bn_shift;
  p.e_gamma = bn_gamma;
  p.e_beta = bn_beta;
  p.wg_dW = dW;
  return launch_conv_gemm(p, dOut_bf16, d->ldA, Wt_bf16, out_bf16, d->ldC, 0, X_bf16, ldX, d->Cin <= 32 ? 32 : 64, false,
                          as_stream(stream));
}

int rxb_conv_dgrad3x3_bn_wgrad_fixup(const rxb_conv_desc* d, const void* G_bf16, const void* Xc_bf16, int ld, int c0,
                                     const float* mean, const float* rstd, const float* corrA, const float* corrB,
                                     const void* Wt_bf16, const void* X_bf16, int ldX, const float* bn_scale,
                                     const float* bn_shift, int out_mode, void* out_bf16, float* sum_dy, float* dW,
                                     rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(d && G_bf16 && Xc_bf16 && mean && rstd && corrA && corrB && Wt_bf16 && X_bf16 && bn_scale && bn_shift &&
                out_bf16 && sum_dy && dW, "rxb_conv_dgrad3x3_bn_wgrad_fixup: null pointer");
  RXB_CHECK_ARG(d->Cin == 32 && d->Cout == 128 && d->taps_x == 3 && d->taps_y == 3 && d->pad_x == 1 && d->pad_y == 1,
                "rxb_conv_dgrad3x3_bn_wgrad_fixup: 3x3, pad 1, 32 -> 128 channels");
  RXB_CHECK_ARG(ld % 8 == 0 && c0 % 8 == 0 && c0 + 32 <= ld && d->ldC >= d->Cout && ldX >= d->Cout,
                "rxb_conv_dgrad3x3_bn_wgrad_fixup: bad ld / c0");
  RXB_CHECK_ARG(out_mode >= OUT_DY && out_mode <= OUT_G_ACCUM, "rxb_conv_dgrad3x3_bn_wgrad_fixup: bad out_mode");
  int rc = rxb_check_device();
  if (rc) return rc;
  GemmParams p = {};
  p.B = d->B; p.H = d->H; p.W = d->W;
  p.n_total = d->Cout;
  p.taps_x = p.taps_y = 3; p.pad_x = p.pad_y = 1;
  p.cin = 32;
  p.epi_mode = EPI_DGRAD_BN;
  p.out_mode = out_mode;
  p.do_stats = 1;
  p.ch_sum = sum_dy;
  p.e_scale = bn_scale;
  p.e_shift = bn_shift;
  p.wg_dW = dW;
  p.fix.G = G_bf16; p.fix.X = Xc_bf16; p.fix.ld = ld; p.fix.c0 = c0;
  p.fix.mean = mean; p.fix.rstd = rstd; p.fix.corrA = corrA; p.fix.corrB = corrB;
  return launch_conv_gemm(p, nullptr, 32, Wt_bf16, out_bf16, d->ldC, 0, X_bf16, ldX, 32, false, as_stream(stream));
}

int rxb_bn_sum_dyx_from_wdw(const float* W, const float* dW, int Cout, int Cin, int taps, const float* bn_scale,
                            const float* bn_shift, const float* sum_dy, float* sum_dyx, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(W && dW && bn_scale && bn_shift && sum_dy && sum_dyx, "rxb_bn_sum_dyx_from_wdw: null pointer");
  RXB_CHECK_ARG(Cout > 0 && Cin > 0 && taps > 0, "rxb_bn_sum_dyx_from_wdw: bad shape");
  int rc = rxb_check_device();
  if (rc) return rc;
  return sum_dyx_from_wdw_launch(W, dW, Cout, Cin, taps, bn_scale, bn_shift, sum_dy, sum_dyx, as_stream(stream));
}

int rxb_conv_wgrad(const rxb_conv_desc* d, const void* A_bf16, const float* scale, const float* shift,
                   const void* dOut_bf16, int ldD, float* dW, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(d && A_bf16 && dOut_bf16 && dW, "rxb_conv_wgrad: null pointer");
  RXB_CHECK_ARG(d->Cin > 0 && d->Cin % 8 == 0 && d->ldA >= d->Cin && d->ldA % 8 == 0, "rxb_conv_wgrad: bad Cin/ldA");
  RXB_CHECK_ARG(d->Cout == 32 || d->Cout % 64 == 0, "rxb_conv_wgrad: Cout must be 32 or a multiple of 64");
  RXB_CHECK_ARG(ldD >= d->Cout && ldD % 8 == 0, "rxb_conv_wgrad: bad ldD");
  RXB_CHECK_ARG(!d->prologue || (scale && shift), "rxb_conv_wgrad: prologue needs scale/shift");
  int rc = rxb_check_device();
  if (rc) return rc;
  const int n_tile = d->Cout <= 256 ? d->Cout : (d->Cout % 256 == 0 ? 256 : (d->Cout % 128 == 0 ? 128 : 64));
  for (int n_off = 0; n_off < d->Cout; n_off += n_tile) {
    WgradParams p = {};
    p.t = make_tiling(d->B, d->H, d->W);
    p.taps_x = d->taps_x; p.taps_y = d->taps_y; p.pad_x = d->pad_x; p.pad_y = d->pad_y;
    p.cin = d->Cin;
    p.bkc = d->Cin <= 32 ? 32 : 64;
    p.n = n_tile;
    p.n_off = n_off;
    p.prologue = d->prologue;
    p.scale = scale;
    p.shift = shift;
    p.dW = dW;
    p.cout_total = d->Cout;
    p.w_mode = 0;
    rc = launch_conv_wgrad(p, A_bf16, d->ldA, dOut_bf16, ldD, as_stream(stream));
    if (rc) return rc;
  }
  return RXB_OK;
}

}  // extern "C"
