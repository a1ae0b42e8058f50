// warp_fixed.cuh — OpenCV's fixed-point affine-warp arithmetic (cv::warpAffine, INTER_LINEAR, uint8,
// BORDER_REFLECT_101), restated operation by operation.  Used by loader_affine_kernel (loader.cu); the
// functions are also host-compilable so that tests/warp_host.cpp can check this very code against
// cv2.warpAffine on the CPU (build the host side with -ffp-contract=off: OpenCV's own build does not
// fuse these multiply-adds, and the device side uses the explicit round-to-nearest intrinsics).
//
// Constants of OpenCV's imgwarp.cpp: AB_BITS = 10 (coordinates carry 10 fractional bits), INTER_BITS = 5
// (rounded to 1/32 pixel with round_delta = 16), INTER_REMAP_COEF_BITS = 15 — for bilinear weights the
// 15-bit table entries are exactly 32*(32-fx|fx)*(32-fy|fy), hence the 10-bit weights and (acc+512)>>10.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RXB_HD __host__ __device__ __forceinline__
#else
#define RXB_HD inline
#endif

namespace rxb {

#if defined(__CUDA_ARCH__)
RXB_HD double wf_mul(double a, double b) { return __dmul_rn(a, b); }
RXB_HD double wf_add(double a, double b) { return __dadd_rn(a, b); }
RXB_HD double wf_sub(double a, double b) { return __dsub_rn(a, b); }
RXB_HD double wf_div(double a, double b) { return __ddiv_rn(a, b); }
RXB_HD int wf_round(double v) { return __double2int_rn(v); }   // cvRound: nearest, ties to even
#else
RXB_HD double wf_mul(double a, double b) { return a * b; }
RXB_HD double wf_add(double a, double b) { return a + b; }
RXB_HD double wf_sub(double a, double b) { return a - b; }
RXB_HD double wf_div(double a, double b) { return a / b; }
RXB_HD int wf_round(double v) { return (int)lrint(v); }
#endif

// The inversion cv::warpAffine applies to a forward 2x3 matrix M (row-major) -> mi.
RXB_HD void warp_invert(const double* M, double* mi) {
  double D = wf_sub(wf_mul(M[0], M[4]), wf_mul(M[1], M[3]));
  D = D != 0.0 ? wf_div(1.0, D) : 0.0;
  const double m0 = wf_mul(M[4], D), m4 = wf_mul(M[0], D);
  const double m1 = wf_mul(M[1], -D), m3 = wf_mul(M[3], -D);
  mi[0] = m0;
  mi[1] = m1;
  mi[3] = m3;
  mi[4] = m4;
  mi[2] = wf_sub(wf_mul(-m0, M[2]), wf_mul(m1, M[5]));
  mi[5] = wf_sub(wf_mul(-m3, M[2]), wf_mul(m4, M[5]));
}

// adelta[x] / bdelta[x]: saturate_cast<int>(m * x * AB_SCALE)
RXB_HD int warp_col_delta(double m, int x) { return wf_round(wf_mul(wf_mul(m, (double)x), 1024.0)); }

// X0 / Y0 of a destination row: saturate_cast<int>((m_y * y + m_c) * AB_SCALE) + round_delta
RXB_HD int warp_row_base(double m_y, double m_c, int y) {
  return wf_round(wf_mul(wf_add(wf_mul(m_y, (double)y), m_c), 1024.0)) + 16;
}

// cv::borderInterpolate(p, n, BORDER_REFLECT_101)
RXB_HD int warp_reflect101(int p, int n) {
  if (n == 1) return 0;
  while ((unsigned)p >= (unsigned)n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

struct WarpTaps {
  int xa, xb, ya, yb;          // reflected source columns / rows of the 2x2 neighbourhood
  int w00, w01, w10, w11;      // 10-bit weights, sum 1024
};

RXB_HD WarpTaps warp_taps(int row_x, int row_y, int col_x, int col_y, int W, int H) {
  const int X = (row_x + col_x) >> 5, Y = (row_y + col_y) >> 5;     // 1/32-pixel coordinates
  int sx = X >> 5, sy = Y >> 5;
  sx = sx < -32768 ? -32768 : (sx > 32767 ? 32767 : sx);            // saturate_cast<short>
  sy = sy < -32768 ? -32768 : (sy > 32767 ? 32767 : sy);
  const int fx = X & 31, fy = Y & 31;
  WarpTaps t;
  t.xa = warp_reflect101(sx, W);
  t.xb = warp_reflect101(sx + 1, W);
  t.ya = warp_reflect101(sy, H);
  t.yb = warp_reflect101(sy + 1, H);
  t.w00 = (32 - fx) * (32 - fy);
  t.w01 = fx * (32 - fy);
  t.w10 = (32 - fx) * fy;
  t.w11 = fx * fy;
  return t;
}

RXB_HD int warp_blend(const WarpTaps& t, int p00, int p01, int p10, int p11) {
  return (t.w00 * p00 + t.w01 * p01 + t.w10 * p10 + t.w11 * p11 + 512) >> 10;
}

}  // namespace rxb
