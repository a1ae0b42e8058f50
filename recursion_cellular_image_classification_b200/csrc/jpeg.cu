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
                   int32_t* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t jpeg_smem[];         // kJpWarps x JpegWarpShared (53 KB: opt-in size)
  JpegWarpShared* sh = reinterpret_cast<JpegWarpShared*>(jpeg_smem);
  const int lane = threadIdx.x & 31;
  const int file = blockIdx.x * kJpWarps + (threadIdx.x >> 5);
  if (file >= n) return;                                       // whole warps leave together
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
  if (lane == 0) status[file] = err;
}

}  // namespace rxb

extern "C" int rxb_jpeg_decode_gray(const uint8_t* blob, const int64_t* begin, const int64_t* end, int n, int H,
                                    int W, uint8_t* dst, int32_t* status, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(n >= 0, "rxb_jpeg_decode_gray: negative file count");
  if (n == 0) return RXB_OK;
  RXB_CHECK_ARG(blob && begin && end && dst && status, "rxb_jpeg_decode_gray: null pointer");
  RXB_CHECK_ARG(H > 0 && W > 0 && H <= 65535 && W <= 65535, "rxb_jpeg_decode_gray: bad image size %dx%d", H, W);
  RXB_CHECK_ARG(((reinterpret_cast<uintptr_t>(begin) | reinterpret_cast<uintptr_t>(end)) & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(status) & 3) == 0,
                "rxb_jpeg_decode_gray: begin/end must be 8-byte aligned, status 4-byte aligned");
  int rc = rxb_check_device();
  if (rc) return rc;
  const size_t smem = kJpWarps * sizeof(JpegWarpShared);
  RXB_CUDA(cudaFuncSetAttribute(jpeg_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RXB_PROF(as_stream(stream), PROF_LOADER);
  jpeg_decode_kernel<<<ceil_div(n, kJpWarps), kJpWarps * 32, smem, as_stream(stream)>>>(blob, begin, end, n, H, W,
                                                                                      dst, status);
  RXB_LAUNCH_OK();
  return RXB_OK;
}
