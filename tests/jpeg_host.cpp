// jpeg_host.cpp — TEST INFRASTRUCTURE ONLY.  Runs the product's baseline-JPEG decoding code
// (recursion_cellular_image_classification_b200/csrc/jpeg_fixed.cuh, the header jpeg_decode_kernel is built from) on
// the CPU, block by block like the kernel, so tests/test_oracle_cpu.py can check it against cv2.imdecode without a
// GPU.  Built by the test with g++ -O2 -shared -fPIC.
#include "jpeg_fixed.cuh"

extern "C" int jpeg_host_decode_gray(const uint8_t* data, int len, int H, int W, uint8_t* dst /*[H,W]*/) {
  using namespace rxb::jpg;
  static HuffTable dc, ac;
  Frame f;
  int st = parse_headers(data, len, &f, &dc, &ac);
  if (st) return st;
  if (f.H != H || f.W != W) return RXB_JPG_BAD_SIZE;
  // the kernel's control flow: a sliding window over the entropy-coded bytes, one block at a time
  static uint8_t win[kWin];
  int file_pos = f.scan;
  int valid = win_slide(win, 0, 0, data + file_pos, len - file_pos, 0, 1);
  BitReader br;
  br_init(&br, win, win + valid);
  const int bw = (W + 7) / 8, bh = (H + 7) / 8;
  int pred = 0, err = 0;
  for (int blk = 0; blk < bw * bh; ++blk) {
    const int consumed = (int)(br.p - win);
    const int more = len - (file_pos + valid);
    if (valid - consumed < kWinGuard && more > 0) {
      valid = win_slide(win, valid, consumed, data + file_pos + valid, more, 0, 1);
      file_pos += consumed;
      br.p = win;
      br.end = win + valid;
    }
    if (f.restart_interval && blk && blk % f.restart_interval == 0) {
      br_restart(&br);
      pred = 0;
    }
    int coef[64] = {0};
    decode_block(&br, &dc, &ac, f.quant, &pred, coef, &err);
    uint32_t px[16];
    idct_islow(coef, px);
    const int by = blk / bw, bx = blk % bw;
    for (int r = 0; r < 8; ++r)
      for (int c = 0; c < 8; ++c) {
        const int y = by * 8 + r, x = bx * 8 + c;
        if (y < H && x < W) dst[(long long)y * W + x] = (uint8_t)(px[2 * r + (c >> 2)] >> (8 * (c & 3)));
      }
  }
  return err;
}
