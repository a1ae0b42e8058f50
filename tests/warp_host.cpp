// warp_host.cpp — TEST INFRASTRUCTURE ONLY.  Runs the product's own fixed-point warp arithmetic
// (recursion_cellular_image_classification_b200/csrc/warp_fixed.cuh, the header loader_affine_kernel is built from)
// on the CPU with the kernel's per-pixel control flow, so tests/test_oracle_cpu.py can check that code against
// cv2.warpAffine without a GPU.  Built by the test with: g++ -O2 -ffp-contract=off -shared -fPIC.
#include "warp_fixed.cuh"

extern "C" void warp_host_planar_u8(const uint8_t* src /*[6,H,W]*/, int H, int W, const double* M /*[2,3] forward*/,
                                    int vflip, int hflip, int y0, int x0, int Ho, int Wo, uint8_t* dst /*[6,Ho,Wo]*/) {
  using namespace rxb;
  double mi[6];
  warp_invert(M, mi);
  const long long plane = (long long)H * W;
  for (int oy = 0; oy < Ho; ++oy)
    for (int ox = 0; ox < Wo; ++ox) {
      const int col_x = warp_col_delta(mi[0], ox + x0), col_y = warp_col_delta(mi[3], ox + x0);
      WarpTaps t = warp_taps(warp_row_base(mi[1], mi[2], oy + y0), warp_row_base(mi[4], mi[5], oy + y0), col_x, col_y,
                             W, H);
      if (hflip) { t.xa = W - 1 - t.xa; t.xb = W - 1 - t.xb; }
      if (vflip) { t.ya = H - 1 - t.ya; t.yb = H - 1 - t.yb; }
      const uint8_t* ra = src + (long long)t.ya * W;
      const uint8_t* rb = src + (long long)t.yb * W;
      for (int c = 0; c < 6; ++c)
        dst[((long long)c * Ho + oy) * Wo + ox] =
            (uint8_t)warp_blend(t, ra[c * plane + t.xa], ra[c * plane + t.xb], rb[c * plane + t.xa], rb[c * plane + t.xb]);
    }
}
