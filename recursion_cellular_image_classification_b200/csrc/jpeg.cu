// jpeg.cu — baseline grayscale JPEG decode on the device.
//
// Replaces ImagesDS._load_from_buffer (reference cell_classifier/dataloader.py:141-146: cv2.imdecode of six
// single-channel JPEG buffers per site; files written by png_to_jpeg.py:11-15) — the host libjpeg call that bounds
// the loader once normalisation and augmentation run at HBM speed (75 % of compute_mean_std's time).
// The arithmetic (marker parsing, canonical Huffman decoding, libjpeg's accurate integer IDCT) is in
// jpeg_fixed.cuh and is bit-identical to cv2.imdecode(buf, -1).
//
// Parallelisation.  Huffman-coded data is sequential within a file (variable-length codes, DC prediction), so the
// unit of parallelism is the file: a training batch holds B*G*6 of them (768 per 128 single-site images).  One WARP
// owns one file.  Lane 0 parses the headers, builds the two look-ahead tables in shared memory and decodes 32 blocks
// at a time into a shared coefficient tile (dequantised, zigzag order; row pitch 65 words so the 32 lanes' blocks
// fall in different banks), reading the entropy-coded bytes through a 2 KB shared-memory window that the warp
// refills cooperatively between blocks; then every lane inverse-transforms one block in registers and stores its 8x8 samples —
// 32 neighbouring blocks make 256 contiguous bytes per image row.  No coefficient buffer in HBM, no workspace.
#include "common.cuh"
#include "jpeg_fixed.cuh"

namespace rxb {

constexpr int kJpWarps = 4;
constexpr int kJpPitch = 65;
constexpr int kJpDeferred = 100;   // internal status: "decode this file with the sequential kernel"

struct JpegWarpShared {
  jpg::HuffTable dc, ac;
  jpg::Frame frame;
  int status;
  int coef[32 * kJpPitch];
  uint8_t win[jpg::kWin];       // sliding window over the entropy-coded bytes (the bit reader reads shared memory)
};

__global__ void __launch_bounds__(kJpWarps * 32)
jpeg_decode_kernel(const uint8_t* __restrict__ blob, const int64_t* __restrict__ begin,
                   const int64_t* __restrict__ endp, int n, int H, int W, uint8_t* __restrict__ dst,
                   int32_t* status, int only_deferred) {
  extern __shared__ __align__(16) uint8_t jpeg_smem[];         // kJpWarps x JpegWarpShared (53 KB: opt-in size)
  JpegWarpShared* sh = reinterpret_cast<JpegWarpShared*>(jpeg_smem);
  const int lane = threadIdx.x & 31;
  const int file = blockIdx.x * kJpWarps + (threadIdx.x >> 5);
  if (file >= n) return;                                       // whole warps leave together
  if (only_deferred && status[file] != kJpDeferred) return;    // second launch: files the parallel kernel passed on
  JpegWarpShared& S = sh[threadIdx.x >> 5];

  const int64_t beg = begin[file], end = endp[file];
  const uint8_t* data = blob + beg;
  const int len = (int)min(end - beg, (int64_t)0x7fffffff);
  if (lane == 0) {
    int st = end > beg ? jpg::parse_headers(data, len, &S.frame, &S.dc, &S.ac) : (int)jpg::RXB_JPG_NOT_JPEG;
    if (st == jpg::RXB_JPG_OK && (S.frame.H != H || S.frame.W != W)) st = jpg::RXB_JPG_BAD_SIZE;
    S.status = st;
  }
  __syncwarp();
  if (S.status != jpg::RXB_JPG_OK) {
    if (lane == 0) status[file] = S.status;
    return;
  }
  int file_pos = S.frame.scan;                                 // file offset of win[0]
  int valid = jpg::win_slide(S.win, 0, 0, data + file_pos, len - file_pos, lane, 32);
  jpg::BitReader br;
  jpg::br_init(&br, S.win, S.win + valid);                     // only lane 0's copy is used

  const int bw = (W + 7) >> 3, bh = (H + 7) >> 3, nblk = bw * bh;
  const int restart = S.frame.restart_interval;
  uint8_t* plane = dst + (long long)file * H * W;
  const bool vec_ok = (W & 7) == 0 && (reinterpret_cast<uintptr_t>(plane) & 7) == 0;
  int pred = 0, err = 0;

  for (int base = 0; base < nblk; base += 32) {
    int* mine = S.coef + lane * kJpPitch;
#pragma unroll 8
    for (int j = 0; j < 64; ++j) mine[j] = 0;
    __syncwarp();
    const int cnt = min(32, nblk - base);
    for (int b = 0; b < cnt; ++b) {                            // warp-uniform: the window slides between blocks
      const int consumed = __shfl_sync(0xffffffffu, (int)(br.p - S.win), 0);
      const int more = len - (file_pos + valid);
      if (valid - consumed < jpg::kWinGuard && more > 0) {
        valid = jpg::win_slide(S.win, valid, consumed, data + file_pos + valid, more, lane, 32);
        file_pos += consumed;
        br.p = S.win;
        br.end = S.win + valid;
      }
      if (lane == 0) {
        if (restart && (base + b) && (base + b) % restart == 0) {
          jpg::br_restart(&br);
          pred = 0;
        }
        jpg::decode_block(&br, &S.dc, &S.ac, S.frame.quant, &pred, S.coef + b * kJpPitch, &err);
      }
    }
    __syncwarp();
    const int blk = base + lane;
    if (blk < nblk) {
      uint32_t px[16];
      jpg::idct_islow(mine, px);                               // reads the zigzag-ordered tile through the folded map
      const int by = blk / bw, bx = blk - by * bw;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int y = by * 8 + r;
        if (y >= H) break;
        uint8_t* row = plane + (long long)y * W + bx * 8;
        if (vec_ok) {
          *reinterpret_cast<uint2*>(row) = make_uint2(px[2 * r], px[2 * r + 1]);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (bx * 8 + c < W) row[c] = (uint8_t)(px[2 * r + (c >> 2)] >> (8 * (c & 3)));
        }
      }
    }
    __syncwarp();
  }
  if (lane == 0) status[file] = err ? err : (jpg::br_overran(&br) ? (int)jpg::RXB_JPG_BAD_CODE : (int)jpg::RXB_JPG_OK);
}

// ------------------------------------------------------------------------------------------------------------
// Speculative parallel path (jpeg_fixed.cuh, "Speculative parallel decoding"): still one warp per file, but all 32
// lanes decode — each owns a 512-bit subsequence of the current 2 KB chunk of the unstuffed stream.  Quantised
// coefficients go to a per-file buffer in the caller's workspace (JCOEF [blocks][64], zigzag order); the warp then
// takes the running sum of the DC differences and inverse-transforms 32 blocks at a time.
struct JpegParShared {
  jpg::HuffTable dc, ac;
  jpg::Frame frame;
  int status;
  alignas(16) uint8_t win[jpg::kChunkBytes + jpg::kSlack];
};
constexpr int kJpParWarps = 4;

// Append unstuffed bytes to win[filled, cap): 128 raw bytes per step, four per lane; stuffed zeros are dropped with a
// ballot-free prefix sum over the lanes' keep counts; the first marker ends the data and the rest is zero-filled.
// Returns the number of bytes of win that hold file data (cap unless the data ended inside this window).
__device__ __forceinline__ int fill_clean_warp(uint8_t* win, int filled, int cap, const uint8_t* raw, int raw_len,
                                               int& raw_pos, int& ended, int lane) {
  int valid = cap;
  while (filled < cap) {
    if (ended || raw_pos >= raw_len) {
      for (int i = filled + lane; i < cap; i += 32) win[i] = 0;
      valid = filled;
      break;
    }
    const int room = cap - filled;
    const int limit = min(raw_len, raw_pos + min(128, room));
    int keep, marker;
    uint8_t by[4];
    jpg::classify4(raw, raw_len, raw_pos + 4 * lane, limit, &keep, &marker, by);
    const unsigned mball = __ballot_sync(0xffffffffu, marker < 4);
    if (mball && lane > __ffs(mball) - 1) keep = 0;            // nothing after the marker is data
    const int cnt = __popc(keep);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    int off = filled + incl - cnt;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((keep >> j) & 1) win[off++] = by[j];
    filled += __shfl_sync(0xffffffffu, incl, 31);
    if (mball) ended = 1; else raw_pos = limit;
  }
  __syncwarp();
  return valid;
}

__global__ void __launch_bounds__(kJpParWarps * 32)
jpeg_decode_par_kernel(const uint8_t* __restrict__ blob, const int64_t* __restrict__ begin,
                       const int64_t* __restrict__ endp, int n, int H, int W, uint8_t* __restrict__ dst,
                       int32_t* status, int16_t* coef_ws) {
  __shared__ JpegParShared sh[kJpParWarps];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int file = blockIdx.x * kJpParWarps + (threadIdx.x >> 5);
  if (file >= n) return;
  JpegParShared& S = sh[threadIdx.x >> 5];

  const int64_t beg = begin[file], end = endp[file];
  const uint8_t* data = blob + beg;
  const int len = (int)min(end - beg, (int64_t)0x7fffffff);
  if (lane == 0) {
    int st = end > beg ? jpg::parse_headers(data, len, &S.frame, &S.dc, &S.ac) : (int)jpg::RXB_JPG_NOT_JPEG;
    if (st == jpg::RXB_JPG_OK && (S.frame.H != H || S.frame.W != W)) st = jpg::RXB_JPG_BAD_SIZE;
    if (st == jpg::RXB_JPG_OK && S.frame.restart_interval) st = kJpDeferred;   // restart markers: sequential kernel
    S.status = st;
  }
  __syncwarp();
  if (S.status != jpg::RXB_JPG_OK) {
    if (lane == 0) status[file] = S.status;
    return;
  }

  const int bw = (W + 7) >> 3, bh = (H + 7) >> 3, nblk = bw * bh;
  int16_t* coef = coef_ws + (long long)file * nblk * 64;
  {
    uint4* z = reinterpret_cast<uint4*>(coef);                 // 128 bytes per block
    for (int i = lane; i < nblk * 8; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  const uint8_t* raw = data + S.frame.scan;
  const int raw_len = len - S.frame.scan;
  const int cap = jpg::kChunkBytes + jpg::kSlack;
  int raw_pos = 0, ended = 0;
  int valid = fill_clean_warp(S.win, 0, cap, raw, raw_len, raw_pos, ended, lane);   // ends with __syncwarp()

  jpg::SubState carry = {0, 0};
  int carry_blocks = 0;
  bool overran = false;
  while (carry_blocks < nblk && valid > 0) {
    const int my_end = (lane + 1) * jpg::kSubBits;
    // Subsequences that begin behind the end of the data hold no blocks; left alone, their zero fill (a periodic
    // stream never falls into step) would cost the full 31 rounds in the last chunk of every file.
    const bool active = lane * jpg::kSubBits < valid * 8;
    jpg::SubState start = carry;
    if (lane) { start.pos = lane * jpg::kSubBits; start.k = 0; }
    jpg::SubState ex = start;
    int cnt = 0;
    if (active) jpg::sub_decode<false>(S.win, &ex, my_end, &S.dc, &S.ac, &cnt, nullptr, 0, nullptr);
    for (;;) {                                                 // synchronisation rounds
      int ppos = __shfl_up_sync(full, ex.pos, 1), pk = __shfl_up_sync(full, ex.k, 1);
      if (lane == 0) { ppos = carry.pos; pk = carry.k; }
      const bool changed = active && (ppos != start.pos || pk != start.k);
      if (!__any_sync(full, changed)) break;
      if (changed) {
        start.pos = ppos; start.k = pk;
        ex = start;
        cnt = 0;
        jpg::sub_decode<false>(S.win, &ex, my_end, &S.dc, &S.ac, &cnt, nullptr, 0, nullptr);
      }
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(full, incl, d);
      if (lane >= d) incl += t;
    }
    int done_pos = -1;
    if (active) {
      jpg::SubState out = start;
      int bi = carry_blocks + incl - cnt;
      jpg::sub_decode<true>(S.win, &out, my_end, &S.dc, &S.ac, &bi, coef, nblk, &done_pos);
    }
    overran |= __any_sync(full, done_pos > valid * 8);       // the last block used bits that are not in the file
    carry.pos = __shfl_sync(full, ex.pos, 31) - jpg::kChunkBytes * 8;
    carry.k = __shfl_sync(full, ex.k, 31);
    carry_blocks += __shfl_sync(full, incl, 31);
    __syncwarp();
    for (int i = lane; i < jpg::kSlack; i += 32) S.win[i] = S.win[jpg::kChunkBytes + i];
    __syncwarp();
    const int fresh = fill_clean_warp(S.win, jpg::kSlack, cap, raw, raw_len, raw_pos, ended, lane);
    valid = valid < cap ? max(valid - jpg::kChunkBytes, 0) : fresh;   // the data end slides with the window
  }
  __syncwarp();                                                // every lane's coefficient stores are visible
  const bool truncated = carry_blocks < nblk || overran;       // the data ended before the last block did

  uint8_t* plane = dst + (long long)file * H * W;
  const bool vec_ok = (W & 7) == 0 && (reinterpret_cast<uintptr_t>(plane) & 7) == 0;
  int run = 0;                                                 // DC predictor carried across groups of 32 blocks
  for (int g = 0; g < nblk; g += 32) {
    const int blk = g + lane;
    const int16_t* zz = coef + (long long)blk * 64;
    int v = blk < nblk ? (int)zz[0] : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(full, v, d);
      if (lane >= d) v += t;
    }
    v += run;
    run = __shfl_sync(full, v, 31);
    if (blk < nblk) {
      int deq[64];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 w = reinterpret_cast<const uint4*>(zz)[q];
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 8; ++e)
          deq[q * 8 + e] = (int)(int16_t)(ww[e >> 1] >> (16 * (e & 1))) * (int)S.frame.quant[q * 8 + e];
      }
      deq[0] = (int)(int16_t)v * (int)S.frame.quant[0];
      uint32_t px[16];
      jpg::idct_islow(deq, px);
      const int by = blk / bw, bx = blk - by * bw;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int y = by * 8 + r;
        if (y >= H) break;
        uint8_t* row = plane + (long long)y * W + bx * 8;
        if (vec_ok) {
          *reinterpret_cast<uint2*>(row) = make_uint2(px[2 * r], px[2 * r + 1]);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (bx * 8 + c < W) row[c] = (uint8_t)(px[2 * r + (c >> 2)] >> (8 * (c & 3)));
        }
      }
    }
  }
  if (lane == 0) status[file] = truncated ? (int)jpg::RXB_JPG_BAD_CODE : (int)jpg::RXB_JPG_OK;
}

}  // namespace rxb

extern "C" size_t rxb_jpeg_decode_workspace_bytes(int n, int H, int W) {
  if (n <= 0 || H <= 0 || W <= 0) return 0;
  return (size_t)n * ((H + 7) / 8) * ((W + 7) / 8) * 64 * sizeof(int16_t);
}

extern "C" int rxb_jpeg_decode_gray(const uint8_t* blob, const int64_t* begin, const int64_t* end, int n, int H,
                                    int W, uint8_t* dst, int32_t* status, void* workspace, size_t workspace_bytes,
                                    rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(n >= 0, "rxb_jpeg_decode_gray: negative file count");
  if (n == 0) return RXB_OK;
  RXB_CHECK_ARG(blob && begin && end && dst && status, "rxb_jpeg_decode_gray: null pointer");
  RXB_CHECK_ARG(H > 0 && W > 0 && H <= 65535 && W <= 65535, "rxb_jpeg_decode_gray: bad image size %dx%d", H, W);
  RXB_CHECK_ARG(((reinterpret_cast<uintptr_t>(begin) | reinterpret_cast<uintptr_t>(end)) & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(status) & 3) == 0,
                "rxb_jpeg_decode_gray: begin/end must be 8-byte aligned, status 4-byte aligned");
  if (workspace) {
    RXB_CHECK_ARG(workspace_bytes >= rxb_jpeg_decode_workspace_bytes(n, H, W),
                  "rxb_jpeg_decode_gray: workspace of %zu bytes, need %zu", workspace_bytes,
                  rxb_jpeg_decode_workspace_bytes(n, H, W));
    RXB_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                  "rxb_jpeg_decode_gray: workspace must be 16-byte aligned");
  }
  int rc = rxb_check_device();
  if (rc) return rc;
  const size_t smem = kJpWarps * sizeof(JpegWarpShared);
  RXB_CUDA(cudaFuncSetAttribute(jpeg_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RXB_PROF(as_stream(stream), PROF_LOADER);
  if (workspace) {
    jpeg_decode_par_kernel<<<ceil_div(n, kJpParWarps), kJpParWarps * 32, 0, as_stream(stream)>>>(
        blob, begin, end, n, H, W, dst, status, reinterpret_cast<int16_t*>(workspace));
    RXB_LAUNCH_OK();
  }
  // without a workspace every file, otherwise the files the parallel kernel passed on (restart intervals)
  jpeg_decode_kernel<<<ceil_div(n, kJpWarps), kJpWarps * 32, smem, as_stream(stream)>>>(blob, begin, end, n, H, W,
                                                                                        dst, status, workspace ? 1 : 0);
  RXB_LAUNCH_OK();
  return RXB_OK;
}
