// loader.cu — kernel family 1b: fused D4 flip/rotate + crop + per-experiment normalisation.
//
// Restates ImagesDS._transform (reference cell_classifier/dataloader.py:128-139) for rotations that
// are multiples of 90 degrees: VerticalFlip -> HorizontalFlip -> rotate(k*90) -> crop -> Normalize
// (albumentations 0.3.0 Normalize: (f32(x) - f32(mean)*255) * (1/(f32(std)*255))), and
// writes the tensor the stem convolution consumes (bf16 NHWC8 or its 2x2 space-to-depth) directly, so
// the fp32 NCHW host tensor and its H2D copy (dataloader.py:177, train.py:44) disappear.
//
// HBM-bound: 6 B read + 12 B (16 B with the two pad channels) written per pixel.
// One CTA produces a 64x64 output tile.  The D4 image of that tile is a 64x64 source square, fetched
// for all 6 planes by ONE TMA box load (u8; the box is 80 bytes wide because TMA needs the box start
// 16-byte aligned in the innermost dimension while crops start at any x; the 80-byte pitch also keeps
// the transposing maps at 4-way shared-memory bank conflicts); threads then gather bytes through the
// affine index map and emit one coalesced 16-byte store per pixel.  Reference-compatible rotations (the cv2.warpAffine gather
// with BORDER_REFLECT_101, not affine at the border) take a plain global-gather path.
#include "common.cuh"
#include "ptx.cuh"
#include "warp_fixed.cuh"

namespace rxb {

constexpr int kLdTile = 64;
constexpr int kLdThreads = 256;
constexpr int kLdPlanes = 6;
constexpr int kLdPitch = kLdTile + 16;  // box width: 64 + slack for 16-byte alignment of the box start

// (cy,cx): coordinates in the augmented SxS image -> (sy,sx) in the source image.
__device__ __forceinline__ void d4_source_coord(int code, int S, int cy, int cx, int& sy, int& sx) {
  const int k = (code >> 2) & 3;
  const bool compat = (code >> 4) & 1;
  // "mirror" index used by the rotation: canonical S-1-t ; reference-compatible reflect101(S-t)
  auto mir = [&](int t) {
    if (!compat) return S - 1 - t;
    int u = S - t;
    return u <= S - 1 ? u : 2 * (S - 1) - u;
  };
  int ry, rx;
  switch (k) {
    case 0: ry = cy; rx = cx; break;
    case 1: ry = cx; rx = mir(cy); break;
    case 2: ry = mir(cy); rx = mir(cx); break;
    default: ry = mir(cx); rx = cy; break;
  }
  sy = (code & 1) ? S - 1 - ry : ry;
  sx = (code & 2) ? S - 1 - rx : rx;
}

struct LoaderArgs {
  const uint8_t* src;
  const int32_t* src_idx;
  const int32_t* exp_id;
  const uint8_t* aug_code;
  const int32_t* crop_yx;
  const float* norm_m;
  const float* norm_d;
  void* dst;
  int S, Ho, Wo, n_exp, fmt;
  long long n_src;
  int tiles_x, tiles_y;
};

__device__ __forceinline__ void emit_pixel(const LoaderArgs& a, int b, int oy, int ox, const float (&v)[6]) {
  if (a.fmt == RXB_OUT_F32_NCHW) {
    float* d = reinterpret_cast<float*>(a.dst) + ((long long)b * kLdPlanes * a.Ho + oy) * a.Wo + ox;
    const long long plane = (long long)a.Ho * a.Wo;
#pragma unroll
    for (int c = 0; c < kLdPlanes; ++c) d[c * plane] = v[c];
  } else {
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = 0u;
    long long off;  // in 16-byte units
    if (a.fmt == RXB_OUT_BF16_NHWC8) {
      off = ((long long)b * a.Ho + oy) * a.Wo + ox;
    } else {  // S2D32: [b][oy/2][ox/2][(oy&1)*16 + (ox&1)*8 + c]
      off = ((((long long)b * (a.Ho >> 1) + (oy >> 1)) * (a.Wo >> 1) + (ox >> 1)) << 2) + ((oy & 1) << 1) +
            (ox & 1);
    }
    st_stream_v4(reinterpret_cast<uint4*>(a.dst) + off, o);
  }
}

__global__ void __launch_bounds__(kLdThreads)
loader_kernel(const __grid_constant__ CUtensorMap tmap_src, const LoaderArgs a) {
  __shared__ __align__(128) uint8_t tile[kLdPlanes * kLdTile * kLdPitch];  // 30 KB: [plane][row][80]
  __shared__ __align__(8) uint64_t bar;
  __shared__ float s_m[kLdPlanes], s_d[kLdPlanes];

  const int b = blockIdx.y;
  const int tx0 = (blockIdx.x % a.tiles_x) * kLdTile;
  const int ty0 = (blockIdx.x / a.tiles_x) * kLdTile;
  const int code = a.aug_code[b];
  const int y0 = a.crop_yx[2 * b], x0 = a.crop_yx[2 * b + 1];
  const int img = a.src_idx[b];
  const int S = a.S;
  const bool compat = (code >> 4) & 1;

  if (threadIdx.x < kLdPlanes) {
    int e = a.exp_id[b];
    e = e < 0 ? 0 : (e >= a.n_exp ? a.n_exp - 1 : e);
    s_m[threadIdx.x] = a.norm_m[e * kLdPlanes + threadIdx.x];
    s_d[threadIdx.x] = a.norm_d[e * kLdPlanes + threadIdx.x];
  }

  // affine source map of this tile (canonical D4): s = s00 + dy*(s10-s00) + dx*(s01-s00)
  int sy00, sx00, sy10, sx10, sy01, sx01;
  d4_source_coord(code & 15, S, ty0 + y0, tx0 + x0, sy00, sx00);
  d4_source_coord(code & 15, S, ty0 + y0 + 1, tx0 + x0, sy10, sx10);
  d4_source_coord(code & 15, S, ty0 + y0, tx0 + x0 + 1, sy01, sx01);
  const int dyy = sy10 - sy00, dxy = sx10 - sx00;  // d(source)/d(dy)
  const int dyx = sy01 - sy00, dxx = sx01 - sx00;  // d(source)/d(dx)
  const int sy_min = sy00 + min(0, (kLdTile - 1) * dyy) + min(0, (kLdTile - 1) * dyx);
  const int sx_min = sx00 + min(0, (kLdTile - 1) * dxy) + min(0, (kLdTile - 1) * dxx);
  const int sx_al = (sx_min >> 4) << 4;  // floor to a multiple of 16 (arithmetic shift: works for negatives)

  if (!compat) {
    if (threadIdx.x == 0) {
      ptx::mbar_init(&bar, 1);
      ptx::fence_barrier_init();
      ptx::mbar_arrive_expect_tx(&bar, kLdPlanes * kLdTile * kLdPitch);
      ptx::tma_load_3d(tile, &tmap_src, &bar, sx_al, sy_min, img * kLdPlanes);
    }
  }
  __syncthreads();
  if (!compat) ptx::mbar_wait(&bar, 0, 100);

  const int dx = threadIdx.x & (kLdTile - 1);
  const int ox = tx0 + dx;
  float m[kLdPlanes], d[kLdPlanes];
#pragma unroll
  for (int c = 0; c < kLdPlanes; ++c) {
    m[c] = s_m[c];
    d[c] = s_d[c];
  }
  const uint8_t* gsrc = a.src + (long long)img * kLdPlanes * S * S;

  for (int dy = threadIdx.x >> 6; dy < kLdTile; dy += kLdThreads / kLdTile) {
    const int oy = ty0 + dy;
    if (oy >= a.Ho || ox >= a.Wo) continue;
    float v[kLdPlanes];
    if (!compat) {
      const int ly = sy00 + dy * dyy + dx * dyx - sy_min;
      const int lx = sx00 + dy * dxy + dx * dxx - sx_al;
      const int phys = ly * kLdPitch + lx;
#pragma unroll
      for (int c = 0; c < kLdPlanes; ++c) {
        float x = (float)tile[c * kLdTile * kLdPitch + phys];
        v[c] = __fmul_rn(__fsub_rn(x, m[c]), d[c]);
      }
    } else {
      int sy, sx;
      d4_source_coord(code, S, oy + y0, ox + x0, sy, sx);
#pragma unroll
      for (int c = 0; c < kLdPlanes; ++c) {
        float x = (float)__ldg(gsrc + ((long long)c * S + sy) * S + sx);
        v[c] = __fmul_rn(__fsub_rn(x, m[c]), d[c]);
      }
    }
    emit_pixel(a, b, oy, ox, v);
  }
}


// ------------------------------------------------------------------------------------------------
// Arbitrary-angle variant: VerticalFlip -> HorizontalFlip -> ShiftScaleRotate -> crop ->
// Normalize, i.e. the reference's full train transform (dataloader.py:42-48).  ShiftScaleRotate is
// cv2.warpAffine(img, M, (W,H), INTER_LINEAR, BORDER_REFLECT_101) on u8; its arithmetic is integer and is
// restated in warp_fixed.cuh so the u8 result is bit-identical to OpenCV's: the matrix is inverted in
// double like cv::warpAffine, source coordinates are fixed point rounded to 1/32 pixel, the four taps go
// through BORDER_REFLECT_101 and are blended with 10-bit integer weights.
// The gather reads the source through the read-only path (an image's six planes are 1.5 MB: L2-resident);
// a rotated footprint is not a TMA box.  Output formats and stores are the D4 loader's.
struct AffineArgs {
  LoaderArgs l;       // aug_code carries the flips only (bit0 vflip, bit1 hflip)
  const double* M;    // [B][2][3] forward matrices as passed to cv2.warpAffine
  int H, W;
};

constexpr int kAfTileW = 64, kAfTileH = 16, kAfThreads = 256;   // 8 warps x 8 columns
static_assert(kAfThreads / 32 * 8 == kAfTileW, "one 8-column strip per warp");

__global__ void __launch_bounds__(kAfThreads) loader_affine_kernel(const AffineArgs a) {
  __shared__ double s_mi[6];
  __shared__ float s_m[kLdPlanes], s_d[kLdPlanes];
  const int b = blockIdx.y;
  const int tx0 = (blockIdx.x % a.l.tiles_x) * kAfTileW;
  const int ty0 = (blockIdx.x / a.l.tiles_x) * kAfTileH;

  if (threadIdx.x == 0) warp_invert(a.M + 6 * (long long)b, s_mi);
  if (threadIdx.x >= 32 && threadIdx.x < 32 + kLdPlanes) {
    const int c = threadIdx.x - 32;
    int e = a.l.exp_id[b];
    e = e < 0 ? 0 : (e >= a.l.n_exp ? a.l.n_exp - 1 : e);
    s_m[c] = a.l.norm_m[e * kLdPlanes + c];
    s_d[c] = a.l.norm_d[e * kLdPlanes + c];
  }
  __syncthreads();

  const int code = a.l.aug_code[b];
  const bool vflip = code & 1, hflip = code & 2;
  const int y0 = a.l.crop_yx[2 * b], x0 = a.l.crop_yx[2 * b + 1];
  const int H = a.H, W = a.W;
  const long long plane = (long long)H * W;
  const long long img = min(max((long long)a.l.src_idx[b], 0ll), a.l.n_src - 1);
  const uint8_t* gsrc = a.l.src + img * kLdPlanes * plane;

  // A warp covers an 8x4 pixel patch (lane = 4 rows x 8 columns), so its rotated source footprint stays compact
  // (about 10x10 pixels: a third of the sectors a 32x1 row of pixels touches) and its stores still form full
  // 128/256-byte runs in the NHWC8 / space-to-depth layouts.  Warp w owns columns [8w, 8w+8) of the 64x16 tile.
  const int lane = threadIdx.x & 31;
  const int ox = tx0 + (threadIdx.x >> 5) * 8 + (lane & 7);
  if (ox >= a.l.Wo) return;
  const int col_x = warp_col_delta(s_mi[0], ox + x0);
  const int col_y = warp_col_delta(s_mi[3], ox + x0);
  float m[kLdPlanes], d[kLdPlanes];
#pragma unroll
  for (int c = 0; c < kLdPlanes; ++c) {
    m[c] = s_m[c];
    d[c] = s_d[c];
  }

  for (int dy = lane >> 3; dy < kAfTileH; dy += 4) {
    const int oy = ty0 + dy;
    if (oy >= a.l.Ho) break;
    WarpTaps t = warp_taps(warp_row_base(s_mi[1], s_mi[2], oy + y0), warp_row_base(s_mi[4], s_mi[5], oy + y0),
                           col_x, col_y, W, H);
    if (hflip) { t.xa = W - 1 - t.xa; t.xb = W - 1 - t.xb; }   // the warp reads the flipped image
    if (vflip) { t.ya = H - 1 - t.ya; t.yb = H - 1 - t.yb; }
    const uint8_t* ra = gsrc + (long long)t.ya * W;
    const uint8_t* rb = gsrc + (long long)t.yb * W;
    float v[kLdPlanes];
#pragma unroll
    for (int c = 0; c < kLdPlanes; ++c) {
      const int px = warp_blend(t, __ldg(ra + c * plane + t.xa), __ldg(ra + c * plane + t.xb),
                                __ldg(rb + c * plane + t.xa), __ldg(rb + c * plane + t.xb));
      v[c] = __fmul_rn(__fsub_rn((float)px, m[c]), d[c]);
    }
    emit_pixel(a.l, b, oy, ox, v);
  }
}

}  // namespace rxb

extern "C" int rxb_load_norm_aug(const uint8_t* src, int64_t n_src, int H, int W, const int32_t* src_idx,
                                 const int32_t* exp_id, const uint8_t* aug_code, const int32_t* crop_yx,
                                 const float* norm_m, const float* norm_d, int n_exp, void* dst, int B,
                                 int Ho, int Wo, int out_format, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(src && src_idx && exp_id && aug_code && crop_yx && norm_m && norm_d && dst,
                "rxb_load_norm_aug: null pointer");
  RXB_CHECK_ARG(H == W, "rxb_load_norm_aug: D4 needs square images (H=%d W=%d)", H, W);
  RXB_CHECK_ARG(H % 16 == 0 && H >= 16, "rxb_load_norm_aug: H must be a multiple of 16");
  RXB_CHECK_ARG(Ho > 0 && Wo > 0 && Ho <= H && Wo <= W, "rxb_load_norm_aug: bad crop size");
  RXB_CHECK_ARG(n_src > 0 && n_src * 6 < (1ll << 31) && n_exp > 0 && B >= 0, "rxb_load_norm_aug: bad sizes");
  RXB_CHECK_ARG(out_format >= RXB_OUT_F32_NCHW && out_format <= RXB_OUT_BF16_S2D32,
                "rxb_load_norm_aug: bad out_format");
  if (out_format == RXB_OUT_BF16_S2D32)
    RXB_CHECK_ARG(Ho % 2 == 0 && Wo % 2 == 0, "rxb_load_norm_aug: S2D32 needs even Ho, Wo");
  RXB_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                "rxb_load_norm_aug: src/dst must be 16-byte aligned");
  RXB_CHECK_ARG(B <= 65535, "rxb_load_norm_aug: B > 65535");
  if (B == 0) return RXB_OK;
  int rc = rxb_check_device();
  if (rc) return rc;

  CUtensorMap tm;
  uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)n_src * 6};
  uint64_t strides[2] = {(uint64_t)W, (uint64_t)W * H};
  uint32_t box[3] = {kLdPitch, kLdTile, kLdPlanes};
  rc = make_tmap(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(src), dims, strides, box,
                 CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc) return rc;

  LoaderArgs a;
  a.src = src; a.src_idx = src_idx; a.exp_id = exp_id; a.aug_code = aug_code; a.crop_yx = crop_yx;
  a.norm_m = norm_m; a.norm_d = norm_d; a.dst = dst;
  a.S = H; a.Ho = Ho; a.Wo = Wo; a.n_exp = n_exp; a.fmt = out_format; a.n_src = n_src;
  a.tiles_x = ceil_div(Wo, kLdTile);
  a.tiles_y = ceil_div(Ho, kLdTile);
  dim3 grid(a.tiles_x * a.tiles_y, B);
  RXB_PROF(as_stream(stream), PROF_LOADER);
  loader_kernel<<<grid, kLdThreads, 0, as_stream(stream)>>>(tm, a);
  RXB_LAUNCH_OK();
  return RXB_OK;
}

extern "C" int rxb_load_norm_affine(const uint8_t* src, int64_t n_src, int H, int W, const int32_t* src_idx,
                                    const int32_t* exp_id, const uint8_t* flip_code, const double* M,
                                    const int32_t* crop_yx, const float* norm_m, const float* norm_d, int n_exp,
                                    void* dst, int B, int Ho, int Wo, int out_format, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(B >= 0, "rxb_load_norm_affine: negative batch");
  if (B == 0) return RXB_OK;  // an empty batch has no addresses to check
  RXB_CHECK_ARG(src && src_idx && exp_id && flip_code && M && crop_yx && norm_m && norm_d && dst,
                "rxb_load_norm_affine: null pointer");
  RXB_CHECK_ARG(H > 0 && W > 0 && H <= 32767 && W <= 32767, "rxb_load_norm_affine: bad image size %dx%d", H, W);
  RXB_CHECK_ARG(Ho > 0 && Wo > 0 && Ho <= H && Wo <= W, "rxb_load_norm_affine: bad crop size");
  RXB_CHECK_ARG(n_src > 0 && n_exp > 0 && B >= 0, "rxb_load_norm_affine: bad sizes");
  RXB_CHECK_ARG(out_format >= RXB_OUT_F32_NCHW && out_format <= RXB_OUT_BF16_S2D32,
                "rxb_load_norm_affine: bad out_format");
  if (out_format == RXB_OUT_BF16_S2D32)
    RXB_CHECK_ARG(Ho % 2 == 0 && Wo % 2 == 0, "rxb_load_norm_affine: S2D32 needs even Ho, Wo");
  RXB_CHECK_ARG((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(M) & 7) == 0,
                "rxb_load_norm_affine: dst must be 16-byte aligned, M 8-byte aligned");
  RXB_CHECK_ARG(B <= 65535, "rxb_load_norm_affine: B > 65535");
  int rc = rxb_check_device();
  if (rc) return rc;

  AffineArgs a;
  a.l.src = src; a.l.src_idx = src_idx; a.l.exp_id = exp_id; a.l.aug_code = flip_code; a.l.crop_yx = crop_yx;
  a.l.norm_m = norm_m; a.l.norm_d = norm_d; a.l.dst = dst;
  a.l.S = H; a.l.Ho = Ho; a.l.Wo = Wo; a.l.n_exp = n_exp; a.l.fmt = out_format; a.l.n_src = n_src;
  a.l.tiles_x = ceil_div(Wo, kAfTileW);
  a.l.tiles_y = ceil_div(Ho, kAfTileH);
  a.M = M; a.H = H; a.W = W;
  dim3 grid(a.l.tiles_x * a.l.tiles_y, B);
  RXB_PROF(as_stream(stream), PROF_LOADER);
  loader_affine_kernel<<<grid, kAfThreads, 0, as_stream(stream)>>>(a);
  RXB_LAUNCH_OK();
  return RXB_OK;
}
