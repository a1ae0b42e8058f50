// jpeg_fixed.cuh — baseline JPEG (SOF0, 8-bit, one component, Huffman) decoding arithmetic, restated from the
// libjpeg(-turbo) algorithms that cv2.imdecode runs for the reference's loader (dataloader.py:141-146; files written
// by png_to_jpeg.py:11-15 — PIL 'L' mode, quality 95): marker parsing (jdmarker.c), canonical Huffman decoding with a
// look-ahead table (jdhuff.c jpeg_make_d_derived_tbl / decode_mcu), and the accurate integer inverse DCT
// (jidctint.c jpeg_idct_islow, the library's default dct_method) with its range-limit table.  Every step is integer
// arithmetic, so the decoded plane is bit-identical to cv2.imdecode(buf, -1).
//
// Used by jpeg_decode_kernel (jpeg.cu).  The functions are also host-compilable: tests/jpeg_host.cpp runs this very
// code on the CPU against cv2.imdecode.  Not supported (status RXB_JPG_UNSUPPORTED): progressive / lossless /
// arithmetic-coded / 12-bit / multi-component files.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RXB_JFN __device__ __forceinline__
#define RXB_JLD(p) (*(p))       // generic loads: the bit reader's window lives in shared memory
#define RXB_JSYNC() __syncwarp()
#else
#define RXB_JFN inline
#define RXB_JLD(p) (*(p))
#define RXB_JSYNC()
#endif

namespace rxb {
namespace jpg {

enum Status {
  RXB_JPG_OK = 0,
  RXB_JPG_NOT_JPEG = 1,      // no SOI, or markers run off the end before SOS
  RXB_JPG_UNSUPPORTED = 2,   // not baseline / 8-bit / single-component Huffman
  RXB_JPG_BAD_TABLE = 3,     // missing or malformed DQT / DHT
  RXB_JPG_BAD_SIZE = 4,      // frame size differs from the size the caller expects
  RXB_JPG_BAD_CODE = 5       // corrupt entropy-coded data (invalid code, coefficient index past 63)
};

constexpr int kLook = 9;     // look-ahead bits (libjpeg uses 8; longer codes take the canonical-code walk)

// natural (row-major) position -> zigzag position (inverse of jpeg_natural_order).  Coefficients are kept in the
// order they are decoded; the inverse DCT reads them through this map, which folds away in its unrolled loops.
RXB_JFN constexpr int zigzag_of_natural(int j) {
  constexpr uint8_t t[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
                             41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
                             46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};
  return t[j];
}

struct HuffTable {
  int maxcode[18];           // largest code of each length (-1 if none); [17] is a sentinel
  int valoffset[17];         // huffval index of the first code of a length, minus that code
  uint16_t lut[1 << kLook];  // (length << 8) | symbol for codes of up to kLook bits, 0 otherwise
  uint8_t huffval[256];
};

struct Frame {
  int H, W;
  int restart_interval;      // MCUs between RSTn markers, 0 = none
  int scan;                  // offset of the first entropy-coded byte
  uint16_t quant[64];        // in zigzag order, as DQT stores it
};

RXB_JFN int rd16(const uint8_t* p) { return (RXB_JLD(p) << 8) | RXB_JLD(p + 1); }

// jpeg_make_d_derived_tbl: canonical codes from the 16 length counts of a DHT segment.
RXB_JFN int build_huff(const uint8_t* seg /*16 counts then the symbols*/, int avail, HuffTable* t) {
  int total = 0;
  for (int l = 1; l <= 16; ++l) total += RXB_JLD(seg + l - 1);
  if (total > 256 || 16 + total > avail) return RXB_JPG_BAD_TABLE;
  for (int i = 0; i < total; ++i) t->huffval[i] = RXB_JLD(seg + 16 + i);
  for (int i = 0; i < (1 << kLook); ++i) t->lut[i] = 0;
  int code = 0, p = 0;
  for (int l = 1; l <= 16; ++l) {
    const int cnt = RXB_JLD(seg + l - 1);
    t->valoffset[l] = p - code;
    for (int i = 0; i < cnt; ++i, ++code, ++p) {
      if (code >= (1 << l)) return RXB_JPG_BAD_TABLE;         // more codes than the length allows
      if (l <= kLook) {
        const int first = code << (kLook - l);
        const uint16_t e = (uint16_t)((l << 8) | t->huffval[p]);
        for (int j = 0; j < (1 << (kLook - l)); ++j) t->lut[first + j] = e;
      }
    }
    t->maxcode[l] = cnt ? code - 1 : -1;
    code <<= 1;
  }
  t->maxcode[17] = 0xFFFFF;
  t->maxcode[0] = -1;
  t->valoffset[0] = 0;
  return RXB_JPG_OK;
}

// Marker walk (jdmarker.c): SOI, then segments until SOS.  Keeps the LAST quantisation / Huffman tables defined
// under the ids the frame and scan headers name (two passes over the same few hundred bytes).
RXB_JFN int parse_headers(const uint8_t* d, int len, Frame* f, HuffTable* dc, HuffTable* ac) {
  if (len < 4 || RXB_JLD(d) != 0xFF || RXB_JLD(d + 1) != 0xD8) return RXB_JPG_NOT_JPEG;
  int tq = -1, td = -1, ta = -1;
  bool have_q = false, have_dc = false, have_ac = false, have_sof = false;
  f->restart_interval = 0;
  for (int pass = 0; pass < 2; ++pass) {
    int p = 2;
    for (;;) {
      if (p + 4 > len) return RXB_JPG_NOT_JPEG;
      if (RXB_JLD(d + p) != 0xFF) { ++p; continue; }          // next_marker: skip garbage
      const int m = RXB_JLD(d + p + 1);
      if (m == 0xFF) { ++p; continue; }                       // fill bytes
      if (m == 0x00 || m == 0x01 || (m >= 0xD0 && m <= 0xD8)) { p += 2; continue; }   // no parameters
      if (m == 0xD9) return RXB_JPG_NOT_JPEG;                 // EOI before SOS
      const int seglen = rd16(d + p + 2);
      if (seglen < 2 || p + 2 + seglen > len) return RXB_JPG_NOT_JPEG;
      const uint8_t* s = d + p + 4;
      const int n = seglen - 2;
      if (m == 0xC0 || m == 0xC1) {                           // baseline / extended-sequential Huffman frame
        if (n < 9 || RXB_JLD(s) != 8 || RXB_JLD(s + 5) != 1) return RXB_JPG_UNSUPPORTED;
        f->H = rd16(s + 1);
        f->W = rd16(s + 3);
        tq = RXB_JLD(s + 8) & 3;
        have_sof = true;
      } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) ||
                 m == 0xCC) {
        return RXB_JPG_UNSUPPORTED;                           // progressive, lossless, arithmetic
      } else if (m == 0xDD) {
        if (n < 2) return RXB_JPG_NOT_JPEG;
        f->restart_interval = rd16(s);
      } else if (m == 0xDB && pass == 1) {
        int q = 0;
        while (q < n) {
          const int pq = RXB_JLD(s + q) >> 4, id = RXB_JLD(s + q) & 15;
          const int bytes = pq ? 128 : 64;
          if (q + 1 + bytes > n) return RXB_JPG_BAD_TABLE;
          if (id == tq) {
            for (int i = 0; i < 64; ++i) f->quant[i] = (uint16_t)(pq ? rd16(s + q + 1 + 2 * i) : RXB_JLD(s + q + 1 + i));
            have_q = true;
          }
          q += 1 + bytes;
        }
      } else if (m == 0xC4 && pass == 1) {
        int q = 0;
        while (q + 17 <= n) {
          const int cls = RXB_JLD(s + q) >> 4, id = RXB_JLD(s + q) & 15;
          int total = 0;
          for (int l = 0; l < 16; ++l) total += RXB_JLD(s + q + 1 + l);
          if (q + 17 + total > n) return RXB_JPG_BAD_TABLE;
          if (cls == 0 && id == td) {
            if (build_huff(s + q + 1, n - q - 1, dc)) return RXB_JPG_BAD_TABLE;
            have_dc = true;
          } else if (cls == 1 && id == ta) {
            if (build_huff(s + q + 1, n - q - 1, ac)) return RXB_JPG_BAD_TABLE;
            have_ac = true;
          }
          q += 17 + total;
        }
      } else if (m == 0xDA) {                                 // SOS
        if (!have_sof) return RXB_JPG_NOT_JPEG;
        if (n < 6 || RXB_JLD(s) != 1) return RXB_JPG_UNSUPPORTED;
        td = RXB_JLD(s + 2) >> 4;
        ta = RXB_JLD(s + 2) & 15;
        if (RXB_JLD(s + 3) != 0 || RXB_JLD(s + 4) != 63) return RXB_JPG_UNSUPPORTED;   // spectral selection
        f->scan = p + 2 + seglen;
        break;
      }
      p += 2 + seglen;
    }
  }
  return (have_q && have_dc && have_ac) ? RXB_JPG_OK : RXB_JPG_BAD_TABLE;
}

// MSB-first bit reader with 0xFF00 unstuffing; after a marker (or the end of the data) it feeds zero bits, like
// jdhuff.c's jpeg_fill_bit_buffer does once cinfo->unread_marker is set.
struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t acc;
  int nbits;
  int marker;                // 0 = none pending
  int fake;                  // zero bits appended after the end of the data (a consumed one means truncation)
};

RXB_JFN void br_init(BitReader* b, const uint8_t* p, const uint8_t* end) {
  b->p = p; b->end = end; b->acc = 0; b->nbits = 0; b->marker = 0; b->fake = 0;
}

RXB_JFN void br_fill(BitReader* b) {
  while (b->nbits <= 56) {
    int c = 0;
    if (!b->marker) {
      if (b->p >= b->end) {
        b->marker = 0xD9;                                    // ran off the data: behave as if EOI was seen
      } else {
        c = RXB_JLD(b->p++);
        if (c == 0xFF) {
          int c2 = 0xFF;
          while (c2 == 0xFF && b->p < b->end) c2 = RXB_JLD(b->p++);   // FF FF .. are fill bytes
          if (c2 == 0xFF) c2 = 0xD9;
          if (c2 != 0) { b->marker = c2; c = 0; }             // a real marker: zero-fill from here on
        }
      }
    }
    if (b->marker) b->fake += 8;
    b->acc = (b->acc << 8) | (uint64_t)c;
    b->nbits += 8;
  }
}

// True if decoding consumed bits that were not in the file (the scan ended before the last block did).
RXB_JFN bool br_overran(const BitReader* b) { return b->fake > b->nbits; }

// At least 32 valid bits (a code of up to 16 bits plus up to 15 value bits).  Common case: the next four bytes hold
// no 0xFF, so they are appended at once; otherwise the byte-wise path sorts out stuffing and markers.
RXB_JFN void br_need32(BitReader* b) {
  if (b->nbits >= 32) return;
  if (!b->marker && b->p + 4 <= b->end) {
    const uint32_t b0 = RXB_JLD(b->p), b1 = RXB_JLD(b->p + 1), b2 = RXB_JLD(b->p + 2), b3 = RXB_JLD(b->p + 3);
    if (b0 != 0xFF && b1 != 0xFF && b2 != 0xFF && b3 != 0xFF) {
      b->acc = (b->acc << 32) | (uint64_t)((b0 << 24) | (b1 << 16) | (b2 << 8) | b3);
      b->nbits += 32;
      b->p += 4;
      return;
    }
  }
  br_fill(b);
}

RXB_JFN int br_peek(const BitReader* b, int n) { return (int)((b->acc >> (b->nbits - n)) & ((1u << n) - 1)); }

RXB_JFN int huff_decode(BitReader* b, const HuffTable* t, int* err) {
  const uint16_t e = t->lut[br_peek(b, kLook)];
  if (e) {
    b->nbits -= e >> 8;
    return e & 255;
  }
  int l = kLook + 1;
  int code = br_peek(b, l);
  while (l <= 16 && code > t->maxcode[l]) {
    ++l;
    code = br_peek(b, l);
  }
  if (l > 16) {
    *err = RXB_JPG_BAD_CODE;
    b->nbits -= 16;
    return 0;
  }
  b->nbits -= l;
  return t->huffval[(code + t->valoffset[l]) & 255];
}

RXB_JFN int receive_extend(BitReader* b, int s) {
  const int x = br_peek(b, s);
  b->nbits -= s;
  return x < (1 << (s - 1)) ? x - (1 << s) + 1 : x;          // HUFF_EXTEND
}

// Restart boundary (jdhuff.c process_restart + jdmarker.c read_restart_marker): drop the partial byte, step over the
// RSTn marker, reset the DC predictor.
RXB_JFN void br_restart(BitReader* b) {
  b->acc = 0;
  b->nbits = 0;
  b->fake = 0;
  if (!b->marker) {                                           // marker not reached yet: scan forward to it
    while (b->p + 1 < b->end && !(RXB_JLD(b->p) == 0xFF && RXB_JLD(b->p + 1) >= 0xD0 && RXB_JLD(b->p + 1) <= 0xD7))
      ++b->p;
    b->p += 2;
    if (b->p > b->end) b->p = b->end;
  } else if (b->marker >= 0xD0 && b->marker <= 0xD7) {
    b->marker = 0;                                            // br_fill already consumed both marker bytes
  }
}

// decode_mcu for one 8x8 block.  `coef` (64 ints, zeroed by the caller) receives DEQUANTISED coefficients in ZIGZAG
// order: the JCOEF (short) value times the quantiser, which is what jpeg_idct_islow's DEQUANTIZE computes.
RXB_JFN void decode_block(BitReader* b, const HuffTable* dc, const HuffTable* ac, const uint16_t* quant, int* pred,
                          int* coef, int* err) {
  br_need32(b);
  int s = huff_decode(b, dc, err) & 15;                       // a DC category is at most 11 (15 for 12-bit files)
  if (s) s = receive_extend(b, s);
  *pred += s;
  coef[0] = (int)(int16_t)*pred * (int)quant[0];
  for (int k = 1; k < 64; ++k) {
    br_need32(b);
    const int rs = huff_decode(b, ac, err);
    const int r = rs >> 4;
    s = rs & 15;
    if (s) {
      k += r;
      if (k > 63) { *err = RXB_JPG_BAD_CODE; return; }
      coef[k] = (int)(int16_t)receive_extend(b, s) * (int)quant[k];
    } else {
      if (r != 15) return;                                    // EOB
      k += 15;                                                // ZRL
    }
  }
}

// The entropy-coded bytes are read through a sliding window (shared memory on the device): `win` holds file bytes
// [file_pos, file_pos + valid).  Before a block is decoded at least kWinGuard unread bytes must be in the window
// (a block consumes at most 64 symbols x 31 bits, doubled by byte stuffing = 496 bytes) unless the file ends there.
// `lane`/`nlanes`: the copy is shared by the lanes of a warp (host: 0, 1).  All lanes pass the same arguments;
// returns the new number of valid bytes (the unread bytes now start at win[0]).
constexpr int kWin = 2048, kWinGuard = 512;

RXB_JFN int win_slide(uint8_t* win, int valid, int consumed, const uint8_t* rest, int rest_len, int lane, int nlanes) {
  const int keep = valid - consumed;                          // < kWinGuard <= consumed: source and target disjoint
  RXB_JSYNC();                                                // the decoding lane is done with the old window
  for (int i = lane; i < keep; i += nlanes) win[i] = win[consumed + i];
  RXB_JSYNC();                                                // the fresh bytes may land on the bytes just moved
  const int fresh = rest_len < kWin - keep ? rest_len : kWin - keep;
  for (int i = lane; i < fresh; i += nlanes) win[keep + i] = RXB_JLD(rest + i);
  RXB_JSYNC();
  return keep + fresh;
}

// ------------------------------------------------------------------------------------------------------------
// Speculative parallel decoding inside ONE file (the warp-scope form of self-synchronising Huffman decoding).
// A Huffman-coded stream can only be decoded from a known symbol boundary, but a decoder started at a wrong bit
// falls into step with the true symbol sequence after a few symbols.  The unstuffed ("clean") stream is cut into
// chunks of 32 subsequences of kSubBits bits, one per lane:
//   1. every lane decodes its subsequence speculatively from its first bit, assuming a block starts there, and
//      records its exit state (bit position just past the subsequence end, zigzag index k) and block count;
//   2. lane l takes lane l-1's exit state as its true entry state (lane 0: the carry from the previous chunk); lanes
//      whose entry changed re-decode; repeated until no entry changes (2-4 rounds in practice, at most 31);
//   3. a prefix sum over the block counts gives every lane its first block index;
//   4. every lane decodes once more from its true entry state and writes the quantised coefficients (JCOEF, zigzag
//      order, DC still a difference) to the file's coefficient buffer.
// A state is (bit position, k): k = 0 means a DC symbol comes next.  The DC predictor is not part of it — DC values
// are a running sum, taken afterwards.  Files with restart intervals take the sequential path.
constexpr int kSubBits = 512;
constexpr int kChunkBytes = 32 * kSubBits / 8;   // 2048 clean bytes per round
constexpr int kSlack = 64;                       // look-ahead bytes behind the chunk (a symbol may straddle its end)

struct SubState {
  int pos;   // bit position relative to the chunk start
  int k;     // zigzag index of the next coefficient; 0 = a DC symbol is next
};

RXB_JFN uint32_t bswap32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(x, 0, 0x0123);
#else
  return __builtin_bswap32(x);
#endif
}

// 32 bits of the clean stream starting at bit `pos` (win is 4-byte aligned)
RXB_JFN uint32_t cr_peek32(const uint8_t* win, int pos) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(win) + (pos >> 5);
  const uint32_t a = bswap32(w[0]), b = bswap32(w[1]);
  const int sh = pos & 31;
  return sh ? (a << sh) | (b >> (32 - sh)) : a;
}

RXB_JFN int huff_lookup(const HuffTable* t, uint32_t bits, int* len) {
  const uint16_t e = t->lut[bits >> (32 - kLook)];
  if (e) {
    *len = e >> 8;
    return e & 255;
  }
  int l = kLook + 1;
  int code = (int)(bits >> (32 - l));
  while (l <= 16 && code > t->maxcode[l]) {
    ++l;
    code = (int)(bits >> (32 - l));
  }
  if (l > 16) {
    *len = 16;
    return 0;
  }
  *len = l;
  return t->huffval[(code + t->valoffset[l]) & 255];
}

RXB_JFN int extend_bits(uint32_t bits, int len, int s) {       // the s value bits that follow a len-bit code
  const int x = (int)((bits << len) >> (32 - s));
  return x < (1 << (s - 1)) ? x - (1 << s) + 1 : x;
}

// Decode from *st until the bit position reaches end_bit (always stopping on a symbol boundary).  *block is the index
// of the block in progress.  With OUT, coefficients of blocks below nblk are stored to coef[block*64 + k] and
// *done_pos receives the bit position at which the file's last block (nblk - 1) ended, if that happens here.
template <bool OUT>
RXB_JFN void sub_decode(const uint8_t* win, SubState* st, int end_bit, const HuffTable* dc, const HuffTable* ac,
                        int* block, int16_t* coef, int nblk, int* done_pos) {
  int pos = st->pos, k = st->k, bi = *block;
  while (pos < end_bit) {
    const uint32_t bits = cr_peek32(win, pos);
    int len;
    if (k == 0) {
      const int s = huff_lookup(dc, bits, &len) & 15;
      if (OUT && bi < nblk) coef[(long long)bi * 64] = (int16_t)(s ? extend_bits(bits, len, s) : 0);
      pos += len + s;
      k = 1;
    } else {
      const int rs = huff_lookup(ac, bits, &len);
      const int r = rs >> 4, s = rs & 15;
      if (s) {
        k += r;
        if (k > 63) k = 63;                                   // corrupt data: keep the store in bounds
        if (OUT && bi < nblk) coef[(long long)bi * 64 + k] = (int16_t)extend_bits(bits, len, s);
        pos += len + s;
        ++k;
      } else {
        pos += len;
        k = r == 15 ? k + 16 : 64;                            // ZRL / EOB
      }
      if (k >= 64) {
        k = 0;
        ++bi;
        if (OUT && bi == nblk) *done_pos = pos;
      }
    }
  }
  st->pos = pos;
  st->k = k;
  *block = bi;
}

// Unstuffing, one lane's share: up to four raw bytes at [base, base+4) below `limit`.  keep bit j = byte j is
// entropy-coded data (a 0x00 that follows 0xFF is stuffing); *marker = index of the first byte that starts a marker
// (0xFF followed by anything but 0x00, or by the end of the file), 4 if none — the data ends there.
RXB_JFN void classify4(const uint8_t* raw, int raw_len, int base, int limit, int* keep, int* marker, uint8_t* bytes) {
  *keep = 0;
  *marker = 4;
  int prev = (base > 0 && base < limit) ? RXB_JLD(raw + base - 1) : 0;   // lanes past the limit read nothing
  for (int j = 0; j < 4; ++j) {
    const int i = base + j;
    if (i >= limit) break;
    const int b = RXB_JLD(raw + i);
    const int next = i + 1 < raw_len ? RXB_JLD(raw + i + 1) : 0xD9;
    if (b == 0xFF && next != 0x00) {
      *marker = j;
      break;
    }
    if (!(b == 0x00 && prev == 0xFF)) *keep |= 1 << j;
    bytes[j] = (uint8_t)b;
    prev = b;
  }
}

// Inverse DCT of one block of the coefficient buffer: JCOEF values (zigzag order) times the quantiser (DEQUANTIZE).
RXB_JFN void idct_islow_q(const int16_t* zz, const uint16_t* quant, uint32_t* px);

// ---- jidctint.c jpeg_idct_islow (CONST_BITS 13, PASS1_BITS 2) ----
constexpr int kF0298 = 2446, kF0390 = 3196, kF0541 = 4433, kF0765 = 6270, kF0899 = 7373, kF1175 = 9633,
              kF1501 = 12299, kF1847 = 15137, kF1961 = 16069, kF2053 = 16819, kF2562 = 20995, kF3072 = 25172;

RXB_JFN int shl(int a, int n) { return (int)((unsigned)a << n); }
RXB_JFN int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// One 1-D pass over 8 values spaced `stride` apart; results descaled by `shift`.
RXB_JFN void idct_1d(const int* in, int stride, int shift, int* o) {
  int z2 = in[2 * stride], z3 = in[6 * stride];
  int z1 = (z2 + z3) * kF0541;
  int tmp2 = z1 + z3 * (-kF1847);
  int tmp3 = z1 + z2 * kF0765;
  z2 = in[0];
  z3 = in[4 * stride];
  int tmp0 = shl(z2 + z3, 13), tmp1 = shl(z2 - z3, 13);
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7 * stride];
  tmp1 = in[5 * stride];
  tmp2 = in[3 * stride];
  tmp3 = in[1 * stride];
  z1 = tmp0 + tmp3;
  z2 = tmp1 + tmp2;
  z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * kF1175;
  tmp0 *= kF0298;
  tmp1 *= kF2053;
  tmp2 *= kF3072;
  tmp3 *= kF1501;
  z1 *= -kF0899;
  z2 *= -kF2562;
  z3 *= -kF1961;
  z4 *= -kF0390;
  z3 += z5;
  z4 += z5;
  tmp0 += z1 + z3;
  tmp1 += z2 + z4;
  tmp2 += z2 + z3;
  tmp3 += z1 + z4;
  o[0] = descale(tmp10 + tmp3, shift);
  o[7] = descale(tmp10 - tmp3, shift);
  o[1] = descale(tmp11 + tmp2, shift);
  o[6] = descale(tmp11 - tmp2, shift);
  o[2] = descale(tmp12 + tmp1, shift);
  o[5] = descale(tmp12 - tmp1, shift);
  o[3] = descale(tmp13 + tmp0, shift);
  o[4] = descale(tmp13 - tmp0, shift);
}

// range_limit[(x) & RANGE_MASK] of jdmaster.c prepare_range_limit_table, centred for the IDCT (x = sample - 128)
RXB_JFN int range_limit(int x) {
  const int i = x & 1023;
  return i < 128 ? i + 128 : (i < 512 ? 255 : (i < 896 ? 0 : i - 896));
}

// zz: 64 dequantised coefficients in zigzag order -> 8 rows of 8 samples, each row packed little-endian into two
// 32-bit words (row r: px[2r] = samples 0-3, px[2r+1] = samples 4-7).
RXB_JFN void idct_islow(const int* zz, uint32_t* px) {
  int coef[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) coef[j] = zz[zigzag_of_natural(j)];
  int ws[64];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    int o[8];
    idct_1d(coef + c, 8, 13 - 2, o);                          // columns: DESCALE(., CONST_BITS - PASS1_BITS)
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[r * 8 + c] = o[r];
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int o[8];
    idct_1d(ws + r * 8, 1, 13 + 2 + 3, o);                    // rows: DESCALE(., CONST_BITS + PASS1_BITS + 3)
    px[2 * r] = (uint32_t)range_limit(o[0]) | ((uint32_t)range_limit(o[1]) << 8) |
                ((uint32_t)range_limit(o[2]) << 16) | ((uint32_t)range_limit(o[3]) << 24);
    px[2 * r + 1] = (uint32_t)range_limit(o[4]) | ((uint32_t)range_limit(o[5]) << 8) |
                    ((uint32_t)range_limit(o[6]) << 16) | ((uint32_t)range_limit(o[7]) << 24);
  }
}

RXB_JFN void idct_islow_q(const int16_t* zz, const uint16_t* quant, uint32_t* px) {
  int deq[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) deq[k] = (int)zz[k] * (int)quant[k];
  idct_islow(deq, px);
}

}  // namespace jpg
}  // namespace rxb
