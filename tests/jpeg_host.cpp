// jpeg_host.cpp — TEST INFRASTRUCTURE ONLY.  Runs the product's baseline-JPEG decoding code
// (recursion_cellular_image_classification_b200/csrc/jpeg_fixed.cuh, the header jpeg_decode_kernel is built from) on
// the CPU, block by block like the kernel, so tests/test_oracle_cpu.py can check it against cv2.imdecode without a
// GPU.  Built by the test with g++ -O2 -shared -fPIC.
#include <cstddef>

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
  return err ? err : (br_overran(&br) ? (int)RXB_JPG_BAD_CODE : 0);
}

// The speculative parallel path, with the warp emulated lane by lane (shuffles = reads of the previous round's
// arrays): cooperative unstuffing, speculative scan, synchronisation rounds, prefix sum, output pass, DC running sum,
// IDCT.  Returns the status; *rounds receives the largest number of synchronisation rounds a chunk needed.
extern "C" int jpeg_host_decode_gray_parallel(const uint8_t* data, int len, int H, int W, uint8_t* dst, int* rounds) {
  using namespace rxb::jpg;
  static HuffTable dc, ac;
  Frame f;
  int st = parse_headers(data, len, &f, &dc, &ac);
  if (st) return st;
  if (f.H != H || f.W != W) return RXB_JPG_BAD_SIZE;
  if (f.restart_interval) return -1;                          // sequential path
  const int bw = (W + 7) / 8, bh = (H + 7) / 8, nblk = bw * bh;
  int16_t* coef = new int16_t[(size_t)nblk * 64]();
  alignas(16) static uint8_t win[kChunkBytes + kSlack];
  const uint8_t* raw = data + f.scan;
  const int raw_len = len - f.scan;
  int raw_pos = 0, ended = 0;
  const int cap = kChunkBytes + kSlack;
  auto fill = [&](int filled) {
    int valid = cap;
    while (filled < cap) {
      if (ended || raw_pos >= raw_len) {
        for (int i = filled; i < cap; ++i) win[i] = 0;
        valid = filled;
        break;
      }
      const int room = cap - filled;
      const int limit = raw_len < raw_pos + (room < 128 ? room : 128) ? raw_len : raw_pos + (room < 128 ? room : 128);
      int total = 0, found = 0;
      for (int lane = 0; lane < 32 && !found; ++lane) {
        int keep, marker;
        uint8_t by[4];
        classify4(raw, raw_len, raw_pos + 4 * lane, limit, &keep, &marker, by);
        for (int j = 0; j < 4; ++j)
          if ((keep >> j) & 1) win[filled + total++] = by[j];
        if (marker < 4) found = 1;
      }
      filled += total;
      if (found) ended = 1; else raw_pos = limit;
    }
    return valid;
  };
  int valid = fill(0);
  SubState carry = {0, 0};
  int carry_blocks = 0, max_rounds = 0;
  bool overran = false;
  while (carry_blocks < nblk && valid > 0) {
    SubState start[32], ex[32];
    int cnt[32];
    bool active[32];
    for (int l = 0; l < 32; ++l) {
      active[l] = l * kSubBits < valid * 8;
      start[l] = l ? SubState{l * kSubBits, 0} : carry;
      ex[l] = start[l];
      cnt[l] = 0;
      if (active[l]) sub_decode<false>(win, &ex[l], (l + 1) * kSubBits, &dc, &ac, &cnt[l], nullptr, 0, nullptr);
    }
    int r = 0;
    for (;; ++r) {
      SubState prev[32];
      for (int l = 0; l < 32; ++l) prev[l] = l ? ex[l - 1] : carry;
      int any = 0;
      for (int l = 1; l < 32; ++l)
        if (active[l] && (prev[l].pos != start[l].pos || prev[l].k != start[l].k)) {
          any = 1;
          start[l] = prev[l];
          ex[l] = start[l];
          cnt[l] = 0;
          sub_decode<false>(win, &ex[l], (l + 1) * kSubBits, &dc, &ac, &cnt[l], nullptr, 0, nullptr);
        }
      if (!any) break;
    }
    if (r > max_rounds) max_rounds = r;
    int base = carry_blocks;
    for (int l = 0; l < 32; ++l) {
      SubState s2 = start[l];
      int bi = base;
      int done_pos = -1;
      if (active[l]) sub_decode<true>(win, &s2, (l + 1) * kSubBits, &dc, &ac, &bi, coef, nblk, &done_pos);
      if (done_pos > valid * 8) overran = true;               // the last block used bits that are not in the file
      base += cnt[l];
    }
    carry = SubState{ex[31].pos - kChunkBytes * 8, ex[31].k};
    carry_blocks = base;
    for (int i = 0; i < kSlack; ++i) win[i] = win[kChunkBytes + i];
    const int fresh = fill(kSlack);
    valid = valid < cap ? (valid > kChunkBytes ? valid - kChunkBytes : 0) : fresh;
  }
  const bool truncated = carry_blocks < nblk || overran;
  int pred = 0;
  for (int b = 0; b < nblk; ++b) {
    pred += coef[(size_t)b * 64];
    coef[(size_t)b * 64] = (int16_t)pred;
    uint32_t px[16];
    idct_islow_q(coef + (size_t)b * 64, f.quant, px);
    const int by = b / bw, bx = b % bw;
    for (int r = 0; r < 8; ++r)
      for (int c = 0; c < 8; ++c) {
        const int y = by * 8 + r, x = bx * 8 + c;
        if (y < H && x < W) dst[(long long)y * W + x] = (uint8_t)(px[2 * r + (c >> 2)] >> (8 * (c & 3)));
      }
  }
  delete[] coef;
  *rounds = max_rounds;
  return truncated ? (int)RXB_JPG_BAD_CODE : 0;
}
