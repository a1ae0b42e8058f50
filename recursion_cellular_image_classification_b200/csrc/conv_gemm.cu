// conv_gemm.cu — kernel family 2: convolutions as tcgen05 implicit GEMMs.
//
// Forward / data-gradient kernel (conv_gemm_kernel):
//     D[p, n] = sum_{tap, k} A'(p shifted by tap)[k] * Wt[tap][n][k]          bf16 x bf16 -> fp32 (TMEM)
//   * persistent, warp-specialised CTA (one per SM): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer and
//     TMEM owner, warps 2-5 = epilogue, warps 6-13 = A-operand transform.
//   * the M tile is 128 pixels fetched by one 4-D TMA box; taps are box shifts, padding is TMA zero fill.
//     For multi-row filters the box carries th+taps_y-1 image rows ("row halo") and the taps_y row taps are
//     shared-memory descriptor offsets into the SAME stage, so an activation row is fetched from L2 (and
//     transformed) taps_x times instead of taps_x*taps_y times.
//   * DenseNet's pre-activation BatchNorm+ReLU (torchvision _DenseLayer: norm -> relu -> conv) is applied
//     to the A tile IN SHARED MEMORY between the TMA and the MMA (zero padding stays zero), so the
//     normalised activation never exists in HBM.
//   * two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
//   * epilogues run through a swizzled shared-memory staging tile: results leave with ONE TMA store per
//     64-channel box (coalesced, clipped at tensor edges, written at a channel offset of a wider tensor =
//     DenseNet concat-by-offset); the fused ReLU/BatchNorm-backward epilogue gets the activation tile and
//     the running concat-gradient tile by TMA loads issued by the producer warp.
//
// Weight-gradient kernel (conv_wgrad_kernel):
//     dW[tap][n][k] += sum_p A'(p shifted by tap)[k] * dOut[p][n]
//   both operands are MN-major (the contraction runs over pixels), 128 A-channels per accumulator group,
//   up to 512 TMEM columns of groups per CTA, pixel range split over CTAs, fp32 atomics into torch OIHW.
//   For 3x3 filters the A' tile is loaded and transformed once per pixel tile and dOut is shifted instead.
//
// Reference: torchvision densenet121 as swapped in for TwoSitesNN's trunk (reference
// cell_classifier/models.py:16-29, 45); the convs there are cuDNN calls.
#include "conv_gemm.cuh"
#include <stdlib.h>

namespace rxb {

constexpr int kXformThreads = 256;                     // 8 A-operand transform warps
constexpr int kGemmThreads = 192 + kXformThreads;      // wgrad kernel: TMA, MMA, 4 epilogue warps, 8 transform warps
// forward / dgrad kernel: warp 0 TMA loads, warp 1 MMA issue, warp 2 TMA stores, warps 3-18 sixteen workers:
//   store epilogue with A prologue : workers 0-7 transform the A operand, workers 8-15 are the epilogue
//   store epilogue, no prologue    : workers 8-15 are the epilogue
//   dgrad epilogue                 : all sixteen workers are the epilogue
// (a worker's TMEM lane quarter is warp % 4; four consecutive warps cover the four quarters)
constexpr int kConvThreads = 96 + 16 * 32;
constexpr int kWorker0 = 3;
constexpr int kMaxStages = 8;
constexpr int kAccStride = 128;   // TMEM columns per accumulator stage (bn <= 128)
constexpr int kGramCol = 256;     // TMEM columns [256,384): running Gram matrix of the stored tiles (diag = sum of squares)
constexpr int kSumCol = 384;      // TMEM columns [384,400): running column sums of the stored tiles
constexpr int kMaxPrologueC = 1024;
constexpr int kMaxBN = 128;

struct __align__(16) GemmAux {
  __nv_bfloat16 s_scale[kMaxPrologueC + 64];   // prologue fold as bf16 pairs: the operands of fma.rn.relu.bf16x2
  __nv_bfloat16 s_shift[kMaxPrologueC + 64];
  float e_scale[kMaxBN];
  float e_shift[kMaxBN];
  uint32_t e_thr2[kMaxBN / 2];   // dgrad ReLU mask as a packed-bf16 threshold test: (x ^ sgn) > thr, two columns per word
  uint32_t e_sgn2[kMaxBN / 2];
  uint32_t e_flag16[kMaxBN / 16];   // dgrad: bit j of word c/16 = column c+j takes direct sum(dy) / sum(dy*x) reductions
  uint32_t e_flag_any4[4];          // OR of e_flag16 per 32 columns (all 0 for every freshly initialised network)
  float s_stat[2][kMaxBN];
  uint64_t full[kMaxStages];
  uint64_t xform[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[4];    // per TMEM accumulator stage (p.n_acc = 2, or 4 for the narrow store epilogue)
  uint64_t tmem_empty[4];
  uint64_t epi_in_full[4][2];   // [buffer][epilogue group that will consume the tile]: each barrier has ONE waiting
                                // group that observes every one of its phases (a parity wait tells only adjacent
                                // phases apart, so two groups must not alternate on one barrier)
  uint64_t epi_in_empty[4];
  uint64_t stg_full[4];    // staged output tile complete in shared memory (epilogue -> MMA warp, store warp)
  uint64_t stg_free[4];    // store warp has read the staged tile and the statistics MMAs over it are complete
  uint64_t b_full;         // resident weights landed
  uint64_t stats_done;
  // fused 1x1 weight gradient (EPI 3): thresholds on the RAW activation for the direct reductions of degenerate
  // channels (e_thr2 / e_sgn2 then hold the test "A' > 0" on the transformed tile), and its barriers
  uint32_t d_thr2[kMaxBN / 2];
  uint32_t d_sgn2[kMaxBN / 2];
  uint64_t xa_ready[4];      // [buffer] activation tile transformed to A' = relu(bn(x)) in place (epilogue group -> MMA warp)
  uint64_t wg_done[4][2];    // [buffer][epilogue group] the weight-gradient MMAs have finished reading the A' tile
  uint64_t wg_final;         // every weight-gradient MMA of this CTA has completed
  uint64_t dz_ready[kMaxStages];   // EPI 4 with FixupArgs: the stage's G box has become dZ (epilogue group -> MMA warp)
  __align__(16) __nv_bfloat16 f_kb[32];    // ... its per-channel constants as bf16 (dZ = g + fma(x, kb, kc))
  __align__(16) __nv_bfloat16 f_kc[32];
  uint32_t tmem_base;
  uint32_t pad;
};

static int log2_ceil(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

PixelTiling make_tiling(int B, int H, int W) {
  PixelTiling t;
  t.W = W; t.H = H; t.B = B;
  int twl = log2_ceil(W); if (twl > 7) twl = 7;
  int thl = log2_ceil(H); if (thl > 7 - twl) thl = 7 - twl;
  int tbl = 7 - twl - thl;
  t.tw_log2 = twl; t.th_log2 = thl; t.tb_log2 = tbl;
  t.tiles_x = ceil_div(W, 1 << twl);
  t.tiles_y = ceil_div(H, 1 << thl);
  t.tiles_b = ceil_div(B, 1 << tbl);
  t.x_step = 1 << twl;
  t.rev = 0;
  return t;
}

// The dense layers' 3x3 data gradient can carry its weight gradient (EPI 4) when the tall 8x16 single-image tiling
// applies, i.e. when launch_conv_gemm picks the full-halo tile for it.
bool conv_dgrad3x3_wgrad_fusable(int B, int H, int W) {
  (void)B;
  static const int off = getenv("RXB_DBG_NO_WGFUSE3") ? atoi(getenv("RXB_DBG_NO_WGFUSE3")) : 0;
  return !off && H > 8 && W > 4;
}

PixelTiling make_tiling_tall(int B, int H, int W) {
  PixelTiling t;
  t.W = W; t.H = H; t.B = B;
  int twl = log2_ceil(W); if (twl > 3) twl = 3;
  int thl = log2_ceil(H); if (thl > 7 - twl) thl = 7 - twl;
  int tbl = 7 - twl - thl;
  t.tw_log2 = twl; t.th_log2 = thl; t.tb_log2 = tbl;
  t.tiles_x = ceil_div(W, 1 << twl);
  t.tiles_y = ceil_div(H, 1 << thl);
  t.tiles_b = ceil_div(B, 1 << tbl);
  t.x_step = 1 << twl;
  t.rev = 0;
  return t;
}

__device__ __forceinline__ void tile_origin(const PixelTiling& t, int m_tile, int& x0, int& y0, int& b0) {
  if (t.rev) m_tile = t.rev - 1 - m_tile;
  const int tx = m_tile % t.tiles_x;
  const int rest = m_tile / t.tiles_x;
  const int ty = rest % t.tiles_y;
  const int tb = rest / t.tiles_y;
  x0 = tx * t.x_step;
  y0 = ty << t.th_log2;
  b0 = tb << t.tb_log2;
}

// ---- A-operand transform: in-place relu(x*scale+shift) on [rows][64 ch] bf16 rows stored with the 128-byte
// swizzle.  kXformThreads threads: thread t owns 16-byte chunk j = t&7 (channels 8j..8j+7) of rows (t>>3) + 32 i.
// (bx, by, bb) is the image coordinate of box row 0; box rows run x fastest, then y (box_h rows), then image.
// Rows whose pixel lies outside the image keep the zeros TMA wrote (conv zero padding); boxes entirely
// inside the image take the path without per-row coordinate arithmetic.
// relu(x*s + h) on two packed bf16 lanes in ONE instruction (single rounding of the exact fused result; the
// BatchNorm scale/shift are rounded to bf16 like every other GEMM operand).  Measured alternative (round 2, removed
// again): fp32 scale/shift with cvt.rn.relu.bf16x2.f32, five instructions per pair - training-mode logits 0.0127
// instead of 0.0144 relative at 512x512 B=16 (both under the 2e-2 bar), step time +3.5 %; see DESIGN.md 5.3.
__device__ __forceinline__ uint32_t fma_relu_bf16x2(uint32_t x, uint32_t s, uint32_t h) {
  uint32_t d;
  asm("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(s), "r"(h));
  return d;
}

__device__ __forceinline__ uint32_t fma_bf16x2(uint32_t x, uint32_t s, uint32_t h) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(s), "r"(h));
  return d;
}
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

__device__ __forceinline__ void transform_chunk(uint4* p, const uint32_t (&s)[4], const uint32_t (&h)[4]) {
  uint4 v = *p;
  v.x = fma_relu_bf16x2(v.x, s[0], h[0]);
  v.y = fma_relu_bf16x2(v.y, s[1], h[1]);
  v.z = fma_relu_bf16x2(v.z, s[2], h[2]);
  v.w = fma_relu_bf16x2(v.w, s[3], h[3]);
  *p = v;
}
__device__ __forceinline__ void transform_box_sw128_regs(uint8_t* tile, int rows, const uint32_t (&s)[4],
                                                         const uint32_t (&h)[4], int t, const PixelTiling& til, int box_w,
                                                         int box_h, int bx, int by, int bb) {
  const int j = t & 7;
  const int tb = 1 << til.tb_log2;
  const bool interior = bx >= 0 && bx + box_w <= til.W && by >= 0 && by + box_h <= til.H && bb + tb <= til.B;
  constexpr int kRowsPerIter = kXformThreads / 8;
  if (interior) {
#pragma unroll 4
    for (int row = t >> 3; row < rows; row += kRowsPerIter)
      transform_chunk(reinterpret_cast<uint4*>(tile + row * 128 + ((j ^ (row & 7)) << 4)), s, h);
  } else {
    for (int row = t >> 3; row < rows; row += kRowsPerIter) {
      const int r2 = row / box_w;
      const int xi = row - r2 * box_w;
      const int bi = tb == 1 ? 0 : r2 / box_h;     // halo boxes hold one image: no second division
      const int yi = r2 - bi * box_h;
      const int x = bx + xi, y = by + yi, b = bb + bi;
      if (x < 0 || x >= til.W || y < 0 || y >= til.H || b >= til.B) continue;
      transform_chunk(reinterpret_cast<uint4*>(tile + row * 128 + ((j ^ (row & 7)) << 4)), s, h);
    }
  }
}
__device__ __forceinline__ void transform_box_sw128(uint8_t* tile, int rows, const __nv_bfloat16* sc,
                                                    const __nv_bfloat16* sh, int t, const PixelTiling& til, int box_w,
                                                    int box_h, int bx, int by, int bb) {
  const int j = t & 7;
  // the 8 channels of this thread's chunk: one 16-byte load each for scale and shift (already bf16 pairs)
  const uint4 s4 = *reinterpret_cast<const uint4*>(sc + j * 8), h4 = *reinterpret_cast<const uint4*>(sh + j * 8);
  const uint32_t s[4] = {s4.x, s4.y, s4.z, s4.w}, h[4] = {h4.x, h4.y, h4.z, h4.w};
  const int tb = 1 << til.tb_log2;
  const bool interior = bx >= 0 && bx + box_w <= til.W && by >= 0 && by + box_h <= til.H && bb + tb <= til.B;
  constexpr int kRowsPerIter = kXformThreads / 8;
  if (interior) {
#pragma unroll 4
    for (int row = t >> 3; row < rows; row += kRowsPerIter)
      transform_chunk(reinterpret_cast<uint4*>(tile + row * 128 + ((j ^ (row & 7)) << 4)), s, h);
  } else {
    for (int row = t >> 3; row < rows; row += kRowsPerIter) {
      const int r2 = row / box_w;
      const int xi = row - r2 * box_w;
      const int bi = tb == 1 ? 0 : r2 / box_h;     // halo boxes hold one image: no second division
      const int yi = r2 - bi * box_h;
      const int x = bx + xi, y = by + yi, b = bb + bi;
      if (x < 0 || x >= til.W || y < 0 || y >= til.H || b >= til.B) continue;
      transform_chunk(reinterpret_cast<uint4*>(tile + row * 128 + ((j ^ (row & 7)) << 4)), s, h);
    }
  }
}

// Column sums over the 32 rows held by a warp for 32 columns: lane L ends with the total of column L.
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = upper ? v[i] : v[i + o];
      const float keep = upper ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// Address of the 16-byte chunk holding channels [col, col+8) of pixel row `row` in a staging tile made of
// boxes [128 rows][cw channels] (cw = 64: 128-byte rows, 128B swizzle; cw = 32: 64-byte rows, 64B swizzle).
__device__ __forceinline__ uint4* staging_chunk(uint8_t* base, int cw, int row, int col) {
  if (cw == 64) {
    const int box = col >> 6, j = (col & 63) >> 3;
    return reinterpret_cast<uint4*>(base + box * (128 * 128) + row * 128 + ((j ^ (row & 7)) << 4));
  }
  const int j = (col & 31) >> 3;
  return reinterpret_cast<uint4*>(base + row * 64 + ((j ^ ((row >> 1) & 3)) << 4));
}

// Two packed-bf16 comparisons a > b in one instruction.
__device__ __forceinline__ void gt_bf16x2(uint32_t a, uint32_t b, bool& lo, bool& hi) {
  uint32_t l, h;
  asm("{\n\t.reg .pred p, q;\n\tsetp.gt.bf16x2 p|q, %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\tselp.u32 %1, 1, 0, q;\n\t}"
      : "=r"(l), "=r"(h)
      : "r"(a), "r"(b));
  lo = l != 0;
  hi = h != 0;
}
// fp32 -> bf16 bits rounded toward -infinity (so that for every bf16 x:  x > t  <=>  x > round_down(t))
__device__ __forceinline__ uint32_t bf16_round_down_bits(float t) {
  const uint32_t u = __float_as_uint(t);
  uint32_t b = u >> 16;
  if ((u & 0x80000000u) && (u & 0xffffu) && (u & 0x7f800000u) != 0x7f800000u) b += 1;   // negative: truncation went up
  return b & 0xffffu;
}
// ReLU mask of relu(es*x + eh) as a threshold on x: returns thr bits, sets sgn (0x8000 flips x when es < 0)
__device__ __forceinline__ uint32_t relu_threshold_bits(float es, float eh, uint32_t& sgn) {
  sgn = 0;
  if (es > 0.f) return bf16_round_down_bits(-eh / es);             // x > -eh/es
  if (es < 0.f) { sgn = 0x8000u; return bf16_round_down_bits(eh / es); }   // x < -eh/es  <=>  -x > eh/es
  return eh > 0.f ? 0xff80u : 0x7f80u;                              // constant mask: thr = -inf (always) / +inf (never)
}

// Direct BatchNorm-backward reductions for the degenerate channels of one 16-column chunk (rare; see bn_degenerate):
// sum(dy) and sum(dy*x) of the UNSCALED fp32 dy over the warp's 32 pixel rows, added to the CTA's shared totals.
// Out of line on purpose, and self-contained: it re-reads the accumulators from TMEM and the activation chunk from
// shared memory (not yet overwritten), so the hot epilogue only tests the chunk's flag word and shares no registers.
__device__ __noinline__ void dgrad_direct_sums(uint32_t fl16, uint32_t taddr, const uint4* xc0, const uint4* xc1,
                                               const uint32_t* thr2, const uint32_t* sgn2, bool row_valid, float* s_dy,
                                               float* s_dyx, int lane) {
  uint32_t r16[16];
  ptx::tmem_ld_32x32b_x16(taddr, r16);
  const uint4 xa = *xc0, xb = *xc1;
  const uint32_t xin[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
  ptx::tmem_ld_wait();
#pragma unroll
  for (int jc = 0; jc < 16; ++jc) {
    if (!((fl16 >> jc) & 1u)) continue;
    bool m0, m1;
    gt_bf16x2(xin[jc >> 1] ^ sgn2[jc >> 1], thr2[jc >> 1], m0, m1);
    const bool m = (jc & 1) ? m1 : m0;
    const float xv = (jc & 1) ? bf16_hi(xin[jc >> 1]) : bf16_lo(xin[jc >> 1]);
    const float d = (m && row_valid) ? __uint_as_float(r16[jc]) : 0.f;
    const float s1 = warp_sum(d), s2 = warp_sum(d * xv);
    if (lane == 0) {
      atomicAdd(s_dy + jc, s1);
      atomicAdd(s_dyx + jc, s2);
    }
  }
}

// development timeline: role r (0 producer, 1 mma, 2 epilogue leader), tile it < 16, event ev < 8
#define RXB_TL(r, it, ev)                                                                         \
  do {                                                                                            \
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && (it) < 16) p.dbg[((r) * 16 + (it)) * 8 + (ev)] = clock64(); \
  } while (0)

// EPI: 0 = store epilogue for 128-wide tiles (statistics, if any, on the tensor pipe); 1 = store epilogue for
// narrow tiles (bn < 128: shuffle statistics, x-merged 3x3 tiles); 2 = fused ReLU/BatchNorm-backward dgrad epilogue;
// 3 = EPI 2 of a 1x1 convolution that ALSO accumulates that convolution's weight gradient (see "fused weight
// gradient" below): the kernel already holds both of its operands in shared memory.
// Compile-time so each variant keeps only its own epilogue (register pressure under the 96-register cap).
//
// Fused weight gradient (EPI 3).  dW[k][c] = sum_p dY[p][k] * A'[p][c] with A' = relu(bn(X)) contracts over pixels the
// very tiles this kernel loads for the data gradient: dY (its A operand, two 64-channel stages) and X (the activation
// tile of the epilogue).  The owning epilogue group turns the X tile into A' in place as soon as it lands (the
// forward prologue's fma.rn.relu.bf16x2, so the ReLU mask becomes the test A' > 0 - exactly the forward's mask), the
// MMA warp then issues 2 x 8 MN-major MMAs  D[c][k] += A'^T dY  into TMEM columns [256,384) (unused by the dgrad
// variant) after the tile's data-gradient MMAs, releases the dY stages and tells the group that the A' tile may be
// overwritten by the staged result.  At the end four warps transpose the 128x128 fp32 accumulator through the dead
// pipeline stages and add it to dW with bulk L2 reduce-adds, like conv_wgrad_kernel.  The kernel is HBM-bound with
// the tensor pipe and shared memory mostly idle, so the extra MMAs are free and the separate weight-gradient launch
// (re-reading X and dY from HBM) disappears.
//
// EPI 4 is the same for the dense layers' 3x3 convolution (dZ 32 channels -> 128): the data gradient's A stage IS the
// full-halo dZ box the weight-gradient kernel's "shifted dOut" mode reads (same origin, same 8x16 tall tiling), so a
// filter row's three taps are ONE N = 96 MMA  D_ty[c][(tx,n)] += A'^T dZ(shifted)  - 3 x 8 MMAs per tile into TMEM
// columns [128,416).  That leaves room for ONE data-gradient accumulator stage, so this variant runs one epilogue
// group of sixteen warps (four column groups) that reads the whole accumulator into registers at once, hands the
// stage back, and only then waits for the weight-gradient MMAs before overwriting the A' tile with the staged result:
// the tensor pipe alternates data-gradient and weight-gradient MMAs without waiting for the epilogue.
template <int BK, bool PROLOGUE, int EPI>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmA2, const GemmParams p) {
  static_assert(BK == 64 || BK == 32, "BK");
  static_assert(!PROLOGUE || BK == 64, "the in-smem BatchNorm+ReLU transform is written for 128B rows");
  constexpr int ROW_BYTES = BK * 2;
  constexpr uint32_t kSwz = BK == 64 ? ptx::kSwizzle128B : ptx::kSwizzle64B;
  constexpr uint32_t kSBO = 8 * ROW_BYTES;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = p.stages;
  const int taps = p.taps_x * p.taps_y;
  const int tps = p.halo == 1 ? p.taps_y : 1;                  // row taps served by one stage of streamed weights
  const int a_tx = p.rows_a * ROW_BYTES;                       // bytes TMA delivers per A box
  const bool fix = EPI == 4 && p.fix.G != nullptr;             // stage = G box + X box (FixupArgs)
  const int a_box = (a_tx + 1023) & ~1023;                     // (both boxes of a stage start 1024-byte aligned: the
  const int a_stage = (fix ? 2 : 1) * a_box;                   //  swizzle pattern follows the absolute address)
  const int b_tap = p.bn * ROW_BYTES;                          // one (tap, k-block) weight tile
  const int b_stage = tps * b_tap;
  const int b_total = p.b_resident ? taps * p.kb_per_tap * b_tap : stages * b_stage;
  const int cw = p.bn >= 64 ? 64 : 32;                         // channels per staging / store box
  const int n_boxes = (p.bn + cw - 1) / cw;
  const int stage_tile = 128 * n_boxes * cw * 2;
  constexpr bool dgrad = EPI >= 2;
  constexpr bool wg = EPI >= 3;
  constexpr bool wg3 = EPI == 4;                      // 3x3: one accumulator stage, one epilogue group
  // EPI 3 splits the sixteen workers like the forward kernel: workers 0-7 turn each activation tile into A' the moment
  // it lands (the weight-gradient MMAs never wait for an epilogue group to come round), workers 8-15 are two epilogue
  // groups of four warps.  EPI 2 and EPI 4 use all sixteen workers as epilogue.
  constexpr bool epi16 = dgrad && EPI != 3;
  constexpr int kWgCol = wg3 ? kAccStride : kGramCol;   // TMEM columns of the weight-gradient accumulators
  constexpr int kSumC = wg3 ? 448 : kSumCol;            // ... and of the running column sums
  constexpr bool narrow = EPI == 1;
  uint8_t* smA = smem;
  uint8_t* smB = smA + (size_t)stages * a_stage;
  uint8_t* st_out = smB + b_total;
  // dgrad: the two activation-tile buffers double as the output staging (the epilogue overwrites x in place)
  uint8_t* st_x = st_out;
  uint8_t* ones = st_out + p.n_stg * stage_tile;   // 1 KB of bf16 1.0: B operand of the column-sum MMA
  GemmAux* aux = reinterpret_cast<GemmAux*>(ones + 1024);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = p.t.tiles_x * p.t.tiles_y * p.t.tiles_b;
  const int my_tiles = blockIdx.x < m_tiles ? (m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int n0 = blockIdx.y * p.bn;
  const int groups = p.halo >= 2 ? 1 : p.halo == 1 ? p.taps_x : taps;  // A loads per k-block
  const int tw = 1 << p.t.tw_log2, th = 1 << p.t.th_log2;
  const int box_h = p.halo ? th + p.taps_y - 1 : th;
  const int box_w = p.halo == 2 ? tw + p.taps_x - 1 : tw;   // halo 3: the 16-wide M tile already contains its x halo
  const bool xmerge = narrow && p.halo == 3;
  const int out_w = xmerge ? tw - (p.taps_x - 1) : tw;       // valid output columns of a tile
  const int n_epi_threads = epi16 ? 512 : 256;                 // EPI 2/4: sixteen worker warps ; stores, EPI 3: workers 8-15

  // ---- one-time setup
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmOut);
    if (dgrad) ptx::prefetch_tmap(&tmX);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&aux->full[s], 1);
      ptx::mbar_init(&aux->xform[s], kXformThreads / 32);   // one arrival per transform warp
      ptx::mbar_init(&aux->empty[s], 1);
    }
    for (int a = 0; a < 4; ++a) {
      ptx::mbar_init(&aux->tmem_full[a], 1);
      ptx::mbar_init(&aux->tmem_empty[a], wg3 ? n_epi_threads / 32 : n_epi_threads / 64);  // one arrival per warp of the stage's epilogue group
    }
    for (int a = 0; a < 4; ++a) {
      ptx::mbar_init(&aux->epi_in_full[a][0], 1);
      ptx::mbar_init(&aux->epi_in_full[a][1], 1);
      ptx::mbar_init(&aux->epi_in_empty[a], p.mma_stats ? 2 : 1);   // dgrad: store warp has read it + stats MMAs done
      ptx::mbar_init(&aux->stg_full[a], 1);
      // store warp has read it (+ statistics MMAs done; + shared-buffer mode: the owning group's column pass is done)
      ptx::mbar_init(&aux->stg_free[a], (p.mma_stats || (!dgrad && p.n_stg == 1 && EPI == 0 && p.do_stats)) ? 2 : 1);
    }
    ptx::mbar_init(&aux->b_full, 1);
    ptx::mbar_init(&aux->stats_done, 1);
    if (wg) {
      for (int a = 0; a < 4; ++a) {
        ptx::mbar_init(&aux->xa_ready[a], wg3 ? 1 : kXformThreads / 32);   // EPI 3: one arrival per transform warp
        ptx::mbar_init(&aux->wg_done[a][0], 1);
        ptx::mbar_init(&aux->wg_done[a][1], 1);
      }
      ptx::mbar_init(&aux->wg_final, 1);
      for (int s = 0; s < stages; ++s) ptx::mbar_init(&aux->dz_ready[s], 1);
    }
    ptx::fence_barrier_init();
  }
  if (p.mma_stats) {
    for (int i = threadIdx.x; i < 256; i += kConvThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
    ptx::fence_proxy_async_smem();
  }
  if (warp == 1) ptx::tmem_alloc<512>(&aux->tmem_base);
  // everything above is private to the CTA and overlaps the previous kernel's drain; global memory from here on
  pdl_sync();
  if (PROLOGUE) {
    const int padded = p.kb_per_tap * BK;
    if (p.prep.sum != nullptr) {
      // fused BatchNorm fold (the arithmetic of bn_prep_kernel); CTA (0,0) publishes it for the backward pass
      const bool publish = blockIdx.x == 0 && blockIdx.y == 0;
      for (int c = threadIdx.x; c < padded; c += kConvThreads) {
        float sc = 0.f, sh = 0.f;
        if (c < p.cin) {
          float mean, var;
          if (p.prep.training) {
            mean = p.prep.sum[c] / p.prep.count;
            var = fmaxf(p.prep.sumsq[c] / p.prep.count - mean * mean, 0.f);
          } else {
            mean = p.prep.rmean[c];
            var = p.prep.rvar[c];
          }
          const float rstd = rsqrtf(var + p.prep.eps);
          sc = p.prep.gamma[c] * rstd;
          sh = p.prep.beta[c] - mean * sc;
          if (publish) {
            p.prep.f_scale[c] = sc;
            p.prep.f_shift[c] = sh;
            p.prep.f_mean[c] = mean;
            p.prep.f_rstd[c] = rstd;
            if (p.prep.training && p.prep.rmean != nullptr) {
              const float unbiased = p.prep.count > 1.f ? var * (p.prep.count / (p.prep.count - 1.f)) : var;
              p.prep.rmean[c] = (1.f - p.prep.momentum) * p.prep.rmean[c] + p.prep.momentum * mean;
              p.prep.rvar[c] = (1.f - p.prep.momentum) * p.prep.rvar[c] + p.prep.momentum * unbiased;
            }
          }
        }
        aux->s_scale[c] = __float2bfloat16_rn(sc);
        aux->s_shift[c] = __float2bfloat16_rn(sh);
      }
    } else {
      for (int c = threadIdx.x; c < padded; c += kConvThreads) {
        aux->s_scale[c] = __float2bfloat16_rn(c < p.cin ? p.scale[c] : 0.f);
        aux->s_shift[c] = __float2bfloat16_rn(c < p.cin ? p.shift[c] : 0.f);
      }
    }
  }
  if (fix) {
    for (int c = threadIdx.x; c < 32; c += kConvThreads) {
      const int ch = p.fix.c0 + c;
      const float rb = p.fix.rstd[ch] * p.fix.corrB[ch];
      aux->f_kb[c] = __float2bfloat16_rn(-rb);
      aux->f_kc[c] = __float2bfloat16_rn(p.fix.mean[ch] * rb - p.fix.corrA[ch]);
    }
  }
  for (int c = threadIdx.x; c < kMaxBN; c += kConvThreads) {
    aux->s_stat[0][c] = 0.f;
    aux->s_stat[1][c] = 0.f;
    const bool in = dgrad && n0 + c < p.n_total;
    aux->e_scale[c] = in ? p.e_scale[n0 + c] : 0.f;
    aux->e_shift[c] = in ? p.e_shift[n0 + c] : 0.f;
    if (wg) {   // the forward prologue's fold operands: the same fp32 fold rounded to bf16
      aux->s_scale[c] = __float2bfloat16_rn(in ? p.e_scale[n0 + c] : 0.f);
      aux->s_shift[c] = __float2bfloat16_rn(in ? p.e_shift[n0 + c] : 0.f);
    }
  }
  if (dgrad) {
    for (int c2 = threadIdx.x; c2 < kMaxBN / 2; c2 += kConvThreads) {
      uint32_t thr = 0, sg = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = 2 * c2 + h;
        uint32_t sg1 = 0, t1 = 0x7f80u;   // columns past n_total: never
        if (n0 + c < p.n_total) t1 = relu_threshold_bits(p.e_scale[n0 + c], p.e_shift[n0 + c], sg1);
        thr |= t1 << (16 * h);
        sg |= sg1 << (16 * h);
      }
      // fused weight gradient: the epilogue sees A' = relu(bn(x)), whose mask is A' > 0 (thr = +0, no sign flip)
      aux->e_thr2[c2] = wg ? 0u : thr;
      aux->e_sgn2[c2] = wg ? 0u : sg;
      aux->d_thr2[c2] = thr;
      aux->d_sgn2[c2] = sg;
    }
    if (threadIdx.x < kMaxBN) {   // warps 0-3: one column per lane, the warp's ballot is two flag words
      const int c = n0 + (int)threadIdx.x;
      const bool f = p.e_gamma != nullptr && c < p.n_total && bn_degenerate(p.e_gamma[c], p.e_beta[c]);
      const uint32_t b = __ballot_sync(0xffffffffu, f);
      if (lane == 0) {
        aux->e_flag16[2 * warp] = b & 0xffffu;
        aux->e_flag16[2 * warp + 1] = b >> 16;
        aux->e_flag_any4[warp] = b;
      }
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = aux->tmem_base;

  if (warp == 0) {
    // =============================== TMA producer
    if (lane == 0) {
      if (p.b_resident) {
        // the whole weight panel of this N tile stays in shared memory for the life of the CTA
        ptx::mbar_arrive_expect_tx(&aux->b_full, taps * p.kb_per_tap * b_tap);
        for (int tap = 0; tap < taps; ++tap)
          for (int kb = 0; kb < p.kb_per_tap; ++kb) {
            // x-merged: slots ordered [ty][kb][tx] so the taps of a filter row are contiguous rows of ONE B operand
            const int ty = tap / p.taps_x, tx = tap - ty * p.taps_x;
            const int slot = p.halo == 3 ? (ty * p.kb_per_tap + kb) * p.taps_x + tx : tap * p.kb_per_tap + kb;
            ptx::tma_load_3d(smB + (size_t)slot * b_tap, &tmB, &aux->b_full, kb * BK, n0, tap);
          }
      }
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int m_tile = blockIdx.x; m_tile < m_tiles; m_tile += gridDim.x, ++it) {
        int x0, y0, b0;
        tile_origin(p.t, m_tile, x0, y0, b0);
        RXB_TL(0, it, 0);
        if (dgrad) {
          // input of this tile's epilogue: the activation tile the consumer's BatchNorm saw (double-buffered)
          const int xb = it % p.n_stg;
          ptx::mbar_wait(&aux->epi_in_empty[xb], ((it / p.n_stg) & 1) ^ 1, 6);
          RXB_TL(0, it, 1);
          uint64_t* xbar = &aux->epi_in_full[xb][wg3 ? 0 : (it & 1)];   // the barrier of the group that owns the tile
          ptx::mbar_arrive_expect_tx(xbar, stage_tile);
          for (int bx = 0; bx < n_boxes; ++bx)
            ptx::tma_load_4d(st_x + (size_t)xb * stage_tile + bx * (128 * cw * 2), &tmX, xbar, n0 + bx * cw, x0, y0, b0);
        }
        for (int g = 0; g < groups; ++g) {
          const int gy = p.halo ? 0 : g / p.taps_x, gx = p.halo >= 2 ? 0 : p.halo == 1 ? g : g - gy * p.taps_x;
          for (int kb = 0; kb < p.kb_per_tap; ++kb) {
            ptx::mbar_wait(&aux->empty[stage], phase ^ 1, 1);
            ptx::mbar_arrive_expect_tx(&aux->full[stage], (fix ? 2 : 1) * a_tx + (p.b_resident ? 0 : b_stage));
            ptx::tma_load_4d(smA + (size_t)stage * a_stage, &tmA, &aux->full[stage], kb * BK, x0 + gx - p.pad_x,
                             y0 + gy - p.pad_y, b0);
            if (fix)   // the X slice's box behind the G slice's
              ptx::tma_load_4d(smA + (size_t)stage * a_stage + a_box, &tmA2, &aux->full[stage], kb * BK, x0 + gx - p.pad_x,
                               y0 + gy - p.pad_y, b0);
            if (!p.b_resident) {
              for (int ty = 0; ty < tps; ++ty) {
                const int tap = p.halo == 1 ? ty * p.taps_x + gx : g;
                ptx::tma_load_3d(smB + (size_t)stage * b_stage + ty * b_tap, &tmB, &aux->full[stage], kb * BK, n0, tap);
              }
            }
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
        RXB_TL(0, it, 2);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer: the whole warp walks the pipeline (so every address is warp-uniform);
    // one elected lane issues the tcgen05 instructions
    {
      const uint32_t idesc = ptx::make_idesc_bf16(128, p.halo == 3 ? p.bn * p.taps_x : p.bn, 0, 0);
      const uint32_t row_tap16 = ((uint32_t)tw * ROW_BYTES) >> 4;  // one image row of the box, in descriptor units
      const uint64_t desc0 = ptx::make_smem_desc(0, 16, kSBO, kSwz);
      const uint32_t d_hi = ptx::desc_hi(desc0), d_lo0 = ptx::desc_lo(desc0);
      // full-halo A box: the 8 pixels of a core-matrix group are one tile row, consecutive tile rows lie box_w box
      // rows apart, and a tap (ty,tx) is a start offset of ty*box_w+tx rows (the 128B swizzle follows the absolute
      // shared-memory address, so unaligned starts and strides are exact - probed on B200, profiles/r01_umma_probe.log)
      const uint32_t a_hi = p.halo == 2 ? ptx::desc_hi(ptx::make_smem_desc(0, 16, (uint32_t)box_w * ROW_BYTES, kSwz)) : d_hi;
      const uint32_t box_row16 = (uint32_t)ROW_BYTES >> 4;
      const uint32_t smA16 = ptx::smem_u32(smA) >> 4, smB16 = ptx::smem_u32(smB) >> 4;
      const uint32_t a_stage16 = (uint32_t)a_stage >> 4, b_stage16 = (uint32_t)b_stage >> 4, b_tap16 = (uint32_t)b_tap >> 4;
      // Column statistics of the stored tiles on the tensor pipe: with S = the staged bf16 tile [128 px][128 ch],
      //   Gram += S^T S  (diagonal = per-channel sum of squares)      sums += S^T 1  (per-channel sum)
      // both MN-major operands straight from the staging buffer the TMA store reads.
      const uint32_t idesc_gram = ptx::make_idesc_bf16(128, 128, 1, 1);
      const uint32_t idesc_sum = ptx::make_idesc_bf16(128, 16, 1, 0);
      const uint64_t d_ones = ptx::make_smem_desc(ptx::smem_u32(ones), 128, 256, ptx::kSwizzleNone);
      auto issue_stats = [&](int j) {
        const int sb = j % p.n_stg;
        const int use = j / p.n_stg;
        ptx::mbar_wait(&aux->stg_full[sb], use & 1, 8);
        ptx::tcgen05_fence_after();
        if (lane == 0) RXB_TL(1, j, 3);
        if (ptx::elect_one()) {
          const uint64_t ds0 = ptx::make_smem_desc(ptx::smem_u32(st_out + (size_t)sb * stage_tile), 16384, 1024,
                                                   ptx::kSwizzle128B);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t ds = ds0 + (uint64_t)(ks * (2048 >> 4));
            const uint32_t accumulate = (j > 0 || ks > 0) ? 1u : 0u;
            if (!dgrad) ptx::umma_bf16_ss(tmem_base + kGramCol, ds, ds, idesc_gram, accumulate);
            ptx::umma_bf16_ss(tmem_base + kSumC, ds, d_ones, idesc_sum, accumulate);
          }
          ptx::umma_commit(dgrad ? &aux->epi_in_empty[sb] : &aux->stg_free[sb]);
        }
        __syncwarp();
      };
      if (p.b_resident) {
        ptx::mbar_wait(&aux->b_full, 0, 9);
        ptx::tcgen05_fence_after();
      }
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      // fused weight gradient: D[c][k] (128 lanes x 64 columns per dY stage) += A'^T dY, both operands MN-major
      const uint32_t idesc_wg = ptx::make_idesc_bf16(128, 64, 1, 1);
      const uint32_t idesc_wg128 = ptx::make_idesc_bf16(128, 128, 1, 1);
      for (int m_tile = blockIdx.x; m_tile < m_tiles; m_tile += gridDim.x, ++it) {
        const int wg_stage0 = stage;   // first dY stage of this tile (released after the weight-gradient MMAs)
        // accumulator stage it % n_acc, used (it / n_acc) times before: the MMA warp runs up to n_acc tiles ahead of
        // the epilogue groups (each group holds a stage for the whole of its TMEM reads)
        const int acc = it % p.n_acc;
        const uint32_t acc_phase = (uint32_t)(it / p.n_acc) & 1u;
        ptx::mbar_wait(&aux->tmem_empty[acc], acc_phase ^ 1, 2);
        ptx::tcgen05_fence_after();
        if (lane == 0) RXB_TL(1, it, 0);
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        uint32_t accumulate = 0;
        for (int g = 0; g < groups; ++g) {
          for (int kb = 0; kb < p.kb_per_tap; ++kb) {
            ptx::mbar_wait(fix ? &aux->dz_ready[stage] : PROLOGUE ? &aux->xform[stage] : &aux->full[stage], phase, 3);
            ptx::tcgen05_fence_after();
            if (ptx::elect_one()) {
              const uint32_t a_lo = d_lo0 + smA16 + (uint32_t)stage * a_stage16;
              if (p.halo == 3) {
                // x-merged: per filter row ONE MMA chain with N = taps_x*bn; accumulator columns [tx][n]
                for (int ty = 0; ty < p.taps_y; ++ty) {
                  const uint32_t b_lo = d_lo0 + smB16 + (uint32_t)((ty * p.kb_per_tap + kb) * p.taps_x) * b_tap16;
                  const uint32_t a_lo_t = a_lo + (uint32_t)ty * row_tap16;
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k) {
                    ptx::umma_bf16_ss_parts(d_tmem, a_lo_t + 2 * k, d_hi, b_lo + 2 * k, d_hi, idesc, accumulate);
                    accumulate = 1;
                  }
                }
              } else if (p.halo == 2) {
                for (int tap = 0; tap < taps; ++tap) {
                  const int ty = tap / p.taps_x, tx = tap - ty * p.taps_x;
                  const uint32_t b_lo = d_lo0 + smB16 + (uint32_t)(tap * p.kb_per_tap + kb) * b_tap16;
                  const uint32_t a_lo_t = a_lo + (uint32_t)(ty * box_w + tx) * box_row16;
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k) {
                    ptx::umma_bf16_ss_parts(d_tmem, a_lo_t + 2 * k, a_hi, b_lo + 2 * k, d_hi, idesc, accumulate);
                    accumulate = 1;
                  }
                }
              } else {
                for (int ty = 0; ty < tps; ++ty) {
                  const int tap = p.halo ? ty * p.taps_x + g : g;
                  const uint32_t b_lo = d_lo0 + smB16 + (p.b_resident ? (uint32_t)(tap * p.kb_per_tap + kb) * b_tap16
                                                                      : (uint32_t)stage * b_stage16 + (uint32_t)ty * b_tap16);
                  const uint32_t a_lo_t = a_lo + (uint32_t)ty * row_tap16;
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k) {
                    ptx::umma_bf16_ss_parts(d_tmem, a_lo_t + 2 * k, d_hi, b_lo + 2 * k, d_hi, idesc, accumulate);
                    accumulate = 1;
                  }
                }
              }
              if (!wg) ptx::umma_commit(&aux->empty[stage]);
            }
            __syncwarp();
            accumulate = 1;
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
        if (ptx::elect_one()) ptx::umma_commit(&aux->tmem_full[acc]);
        __syncwarp();
        if (lane == 0) RXB_TL(1, it, 2);
        if (wg) {
          // this tile's weight-gradient MMAs: A = the activation tile turned into A' in place by its epilogue group,
          // B = the tile's dY stages (still held); then the stages go back to the producer
          const int sb = it % p.n_stg;
          ptx::mbar_wait(&aux->xa_ready[sb], (uint32_t)(it / p.n_stg) & 1u, 20);
          ptx::tcgen05_fence_after();
          if (wg3) {
            if (ptx::elect_one()) {
              // A = the A' tile (two 64-channel boxes, MN-major); B = the full-halo dZ box of this tile's A stage: a
              // filter row's taps are N atoms ONE box row apart in descending tx order, an 8-pixel K group is one tile
              // row, consecutive groups lie box_w rows apart (conv_wgrad_kernel's shift_dout == 2 addressing)
              const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(st_out + (size_t)sb * stage_tile), 16384, 1024,
                                                       ptx::kSwizzle128B);
              const uint64_t db0 = ptx::make_smem_desc(ptx::smem_u32(smA + (size_t)wg_stage0 * a_stage), ROW_BYTES,
                                                       (uint32_t)box_w * ROW_BYTES, kSwz);
              const uint32_t idesc_wg3 = ptx::make_idesc_bf16(128, 96, 1, 1);
              const uint32_t kstep_b = (2u * (uint32_t)box_w * ROW_BYTES) >> 4;   // 16 pixels = two tile rows
              for (int ty = 0; ty < 3; ++ty) {
                const uint64_t db_t = db0 + (uint64_t)(((uint32_t)((2 - ty) * box_w) * ROW_BYTES) >> 4);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                  ptx::umma_bf16_ss(tmem_base + kWgCol + ty * 96, da0 + (uint64_t)(ks * (2048 >> 4)),
                                    db_t + (uint64_t)(ks * kstep_b), idesc_wg3, (it > 0 || ks > 0) ? 1u : 0u);
              }
              ptx::umma_commit(&aux->empty[wg_stage0]);
              ptx::umma_commit(&aux->wg_done[sb][0]);
            }
          } else if (ptx::elect_one()) {
            const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(st_out + (size_t)sb * stage_tile), 16384, 1024,
                                                     ptx::kSwizzle128B);
            int s2 = wg_stage0;
            if (p.kb_per_tap == 2 && wg_stage0 + 1 < stages) {
              // the tile's two dY stages are adjacent (no wrap of the stage ring between them): ONE N = 128 MMA chain
              // whose two 64-channel atoms lie a stage (16 KB) apart - half the MMAs
              const uint64_t db0 = ptx::make_smem_desc(ptx::smem_u32(smA + (size_t)s2 * a_stage), (uint32_t)a_stage, 1024,
                                                       ptx::kSwizzle128B);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                ptx::umma_bf16_ss(tmem_base + kWgCol, da0 + (uint64_t)(ks * (2048 >> 4)),
                                  db0 + (uint64_t)(ks * (2048 >> 4)), idesc_wg128, (it > 0 || ks > 0) ? 1u : 0u);
              ptx::umma_commit(&aux->empty[s2]);
              ptx::umma_commit(&aux->empty[s2 + 1]);
            } else
            for (int kb = 0; kb < p.kb_per_tap; ++kb) {
              const uint64_t db0 = ptx::make_smem_desc(ptx::smem_u32(smA + (size_t)s2 * a_stage), 16384, 1024,
                                                       ptx::kSwizzle128B);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                ptx::umma_bf16_ss(tmem_base + kWgCol + kb * 64, da0 + (uint64_t)(ks * (2048 >> 4)),
                                  db0 + (uint64_t)(ks * (2048 >> 4)), idesc_wg, (it > 0 || ks > 0) ? 1u : 0u);
              ptx::umma_commit(&aux->empty[s2]);
              if (++s2 == stages) s2 = 0;
            }
            ptx::umma_commit(&aux->wg_done[sb][it & 1]);
          }
          __syncwarp();
        }
        if (p.mma_stats && it > 0) issue_stats(it - 1);   // the previous tile's epilogue ran under this tile's main loop
      }
      if (wg) {
        if (ptx::elect_one()) ptx::umma_commit(&aux->wg_final);
        __syncwarp();
      }
      if (p.mma_stats) {
        if (it > 0) issue_stats(it - 1);
        if (ptx::elect_one()) ptx::umma_commit(&aux->stats_done);
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // =============================== TMA store warp: staged tile -> global, one TMA store (or L2 reduce-add into the
    // running gradient) per 64-channel box, clipped at the tensor edges.  Keeps store issue and the wait for the
    // TMA engine to drain the staging buffer off the epilogue warps' critical path.
    if (lane == 0) {
      int it = 0;
      for (int m_tile = blockIdx.x; m_tile < m_tiles; m_tile += gridDim.x, ++it) {
        int x0, y0, b0;
        tile_origin(p.t, m_tile, x0, y0, b0);
        const int sb = it % p.n_stg;
        const int use = it / p.n_stg;
        const uint8_t* so = st_out + (size_t)sb * stage_tile;
        ptx::mbar_wait(&aux->stg_full[sb], use & 1, 16);
        for (int bx = 0; bx < n_boxes; ++bx) {
          if (n0 + bx * cw >= p.n_total) break;
          if (dgrad && p.out_mode == OUT_G_ACCUM)
            ptx::tma_reduce_add_4d(&tmOut, so + bx * (128 * cw * 2), n0 + bx * cw, x0, y0, b0);
          else
            ptx::tma_store_4d(&tmOut, so + bx * (128 * cw * 2), n0 + bx * cw, x0, y0, b0);
        }
        ptx::tma_store_commit();
        ptx::tma_store_wait_read();
        // (shared staging buffer, n_stg == 1: the release goes to the barrier of the group that stages the NEXT tile)
        ptx::mbar_arrive(dgrad ? &aux->epi_in_empty[sb] : &aux->stg_free[p.n_stg == 1 ? ((it + 1) & 1) : sb]);
      }
      ptx::tma_store_wait_all();
    }
  } else if (epi16 || warp >= kWorker0 + 8) {
    // =============================== epilogue: two groups of warps take alternate tiles (group g owns TMEM accumulator
    // stage g), so one group's TMEM reads overlap the other's arithmetic and shared-memory traffic.  Within a
    // group a warp reads TMEM lanes (warp & 3) * 32 .. ; dgrad has two warps per lane quarter that split the
    // 32-column chunks round-robin.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int first_epi_warp = epi16 ? kWorker0 : kWorker0 + 8;
    const int warps_per_group = wg3 ? 16 : epi16 ? 8 : 4;   // EPI 4: ONE group of sixteen warps takes every tile
    const int e = warp - first_epi_warp;
    const int g2 = e / warps_per_group;                       // epilogue group = accumulator stage
    const int grp = (e - g2 * warps_per_group) >> 2;          // column group inside the epilogue group
    const int n_grp = wg3 ? 4 : epi16 ? 2 : 1;
    const int group_threads = warps_per_group * 32;
    const int et = threadIdx.x - (first_epi_warp + g2 * warps_per_group) * 32;   // 0..group_threads-1
    const bool leader = et == 0;
    const uint32_t bar_threads = (uint32_t)group_threads;
    const uint32_t bar_id = 1 + g2;
    // fused weight gradient: tile `t` (buffer t % n_stg) of this CTA becomes A' = relu(bn(x)) in place (both 64-channel
    // boxes, every thread of the group) and is handed to the MMA warp
    auto wg_transform_tile = [&](int t) {
      const int tsb = t % p.n_stg;
      uint8_t* tso = st_out + (size_t)tsb * stage_tile;
      int tx0, ty0, tb0;
      tile_origin(p.t, blockIdx.x + t * gridDim.x, tx0, ty0, tb0);
      if (wg3) {   // 512 threads: one 64-channel box per half
        const int bx = et >> 8;
        transform_box_sw128(tso + bx * (128 * 128), 128, aux->s_scale + bx * 64, aux->s_shift + bx * 64, et & 255, p.t, tw,
                            th, tx0, ty0, tb0);
      } else {   // EPI 3 (only CTAs with degenerate channels come here): 128 threads do the 256-thread pattern twice
        for (int bx = 0; bx < n_boxes; ++bx)
          for (int hh = 0; hh < 2; ++hh)
            transform_box_sw128(tso + bx * (128 * 128), 128, aux->s_scale + bx * 64, aux->s_shift + bx * 64, et + 128 * hh,
                                p.t, tw, th, tx0, ty0, tb0);
      }
      ptx::fence_proxy_async_smem();     // every writing thread orders its stores before the MMA's async reads
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_threads) : "memory");
      if (wg3 ? leader : et < kXformThreads / 32) ptx::mbar_arrive(&aux->xa_ready[tsb]);   // the barrier's arrival count
    };
    bool any_flag = false;
    if constexpr (dgrad) {
      const uint4 any4 = *reinterpret_cast<const uint4*>(aux->e_flag_any4);
      any_flag = (any4.x | any4.y | any4.z | any4.w) != 0;
    }
    // EPI 4 runs the transform one tile AHEAD (software pipelining): tile it+1 is transformed between the TMEM reads
    // of tile it and the wait for tile it's weight-gradient MMAs, so those MMAs never wait for this group to come
    // round again.  (CTAs with degenerate BatchNorm channels need the raw x after tmem_full and keep the simple order.)
    // EPI 0 statistics without the tensor pipe: each of the group's 128 threads owns one channel PAIR and one half of
    // the tile's rows, and walks the staged bf16 tile once it is complete (a warp reads one 128-byte row per
    // instruction: conflict-free under the swizzle); the sums live in registers for the life of the CTA.  Replaces the
    // Gram + column-sum MMAs, whose operand reads (100 KB per tile) competed with the main loop for shared-memory
    // bandwidth - the resource these kernels run out of.
    const bool col_stats = EPI == 0 && p.do_stats && !p.mma_stats;
    // EPI 1 (bn = 32 or 64): the same pass instead of warp-shuffle column sums - a shuffle costs a shared-memory
    // wavefront like a load, and the butterfly needed 124 of them per warp and tile against 14-32 row loads
    const bool col_narrow = EPI == 1 && p.do_stats && !p.mma_stats && p.col_narrow;
    float cs0 = 0.f, cs1 = 0.f, cq0 = 0.f, cq1 = 0.f;
    // EPI 4 with FixupArgs: tile t's stage holds the G and X slices' full-halo boxes [box_h][box_w] pixels x 32 channels
    // (64-byte rows, 64B swizzle); the G box becomes dZ = g + fma(x, kb, kc) in place.  Pixels outside the image keep
    // the zeros TMA wrote.  One tile ahead, like the A' transform.
    auto dz_transform_tile = [&](int t) {
      const int st = t % stages;
      ptx::mbar_wait(&aux->full[st], (uint32_t)(t / stages) & 1u, 26);
      uint8_t* gbox = smA + (size_t)st * a_stage;
      const uint8_t* xbox = gbox + a_box;
      int tx0, ty0, tb0;
      tile_origin(p.t, blockIdx.x + t * gridDim.x, tx0, ty0, tb0);
      const int n_chunks = p.rows_a * 4;                     // 16-byte chunks of a box
      for (int idx = et; idx < n_chunks; idx += (int)bar_threads) {
        const int rowb = idx >> 2, j = idx & 3;
        const int yi = rowb / box_w, xi = rowb - yi * box_w;
        const int x = tx0 - p.pad_x + xi, y = ty0 - p.pad_y + yi;
        if (x < 0 || x >= p.t.W || y < 0 || y >= p.t.H) continue;
        const int off = rowb * 64 + ((j ^ ((rowb >> 1) & 3)) << 4);
        const uint4 kb4 = *reinterpret_cast<const uint4*>(aux->f_kb + j * 8), kc4 = *reinterpret_cast<const uint4*>(aux->f_kc + j * 8);
        const uint4 xv = *reinterpret_cast<const uint4*>(xbox + off);
        uint4 gv = *reinterpret_cast<uint4*>(gbox + off);
        gv.x = add_bf16x2(gv.x, fma_bf16x2(xv.x, kb4.x, kc4.x));
        gv.y = add_bf16x2(gv.y, fma_bf16x2(xv.y, kb4.y, kc4.y));
        gv.z = add_bf16x2(gv.z, fma_bf16x2(xv.z, kb4.z, kc4.z));
        gv.w = add_bf16x2(gv.w, fma_bf16x2(xv.w, kb4.w, kc4.w));
        *reinterpret_cast<uint4*>(gbox + off) = gv;
      }
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_threads) : "memory");
      if (leader) ptx::mbar_arrive(&aux->dz_ready[st]);
    };
    if (fix && my_tiles > 0) dz_transform_tile(0);
    const bool wg_ahead = wg3 && !any_flag;
    if (wg_ahead && my_tiles > 0) {
      ptx::mbar_wait(&aux->epi_in_full[0][0], 0, 23);
      wg_transform_tile(0);
    }
    for (int it = g2; it < my_tiles; it += (wg3 ? 1 : 2)) {
      const int acc = it % p.n_acc;                              // n_acc is even (or 1 with one group): a stage always belongs to one group
      const uint32_t acc_phase = (uint32_t)(it / p.n_acc) & 1u;
      const int m_tile = blockIdx.x + it * gridDim.x;
      int x0, y0, b0;
      tile_origin(p.t, m_tile, x0, y0, b0);
      const int sb = it % p.n_stg;
      const int use = it / p.n_stg;
      uint8_t* so = st_out + (size_t)sb * stage_tile;
      if (leader) RXB_TL(2, it, 0);
      // the staging buffer is free once the TMA store issued n_stg tiles ago has read it and (statistics on the
      // tensor pipe) the MMAs over it have completed; dgrad stages in place over the activation tile it owns
      // (stores always have TWO staging buffers, one per epilogue group, so a group waits on every phase of its buffer)
      if (!dgrad && p.n_stg == 1) {
        // ONE staging buffer shared by the two groups (32 KB more for the operand pipeline): tile it waits for the
        // release of tile it-1, which arrives on this group's own barrier - its ((it-1)/2)-th phase
        if (it > 0) ptx::mbar_wait(&aux->stg_free[g2], (uint32_t)((it - 1) >> 1) & 1u, 10);
      } else if (!dgrad && use > 0) ptx::mbar_wait(&aux->stg_free[sb], (use - 1) & 1, 10);
      if (leader) RXB_TL(2, it, 2);
      // this group's own barrier of the buffer, used (it / period) times before; period = lcm(n_stg, 2) tiles
      const int period = (wg3 || !(p.n_stg & 1)) ? p.n_stg : 2 * p.n_stg;
      if (dgrad) ptx::mbar_wait(&aux->epi_in_full[sb][g2], (it / period) & 1, 7);
      // fused weight gradient: the activation tile is transformed as soon as it lands (EPI 4: already done, one tile
      // ahead).  CTAs with degenerate BatchNorm channels (rare) need the raw x for their direct reductions and
      // transform after that pass, below.
      // (EPI 3: workers 0-7 transform the tile; EPI 4: done one tile ahead, below)
      if (leader) RXB_TL(2, it, 3);
      // rows whose pixel lies outside the image are clipped by the TMA store; keep them out of the channel sums
      // (a multi-tap filter gives them non-zero accumulators from their in-image neighbours)
      const int r2 = row >> p.t.tw_log2;
      const int xx = row & (tw - 1);                       // column inside the M tile
      const int xo = xmerge ? xx - p.pad_x : xx;           // output column inside the tile (x-merged: minus the halo)
      const bool row_valid = xo >= 0 && xo < out_w && x0 + xo < p.t.W && y0 + (r2 & (th - 1)) < p.t.H &&
                             b0 + (r2 >> p.t.th_log2) < p.t.B;
      const int srow = xmerge ? (r2 * out_w + min(max(xo, 0), out_w - 1)) : row;   // staging row of this thread
      ptx::mbar_wait(&aux->tmem_full[acc], acc_phase, 4);
      ptx::tcgen05_fence_after();
      if (leader) RXB_TL(2, it, 4);
      if constexpr (dgrad) {
        // degenerate BatchNorm channels (kernel-uniform test, rare): a separate pass over the flagged 16-column chunks
        // BEFORE the activation tile is overwritten, so the main loop below carries no trace of it
        if (any_flag) {
          for (int c = grp * 32; c < p.bn; c += n_grp * 32) {
            for (int hc = 0; hc < 32; hc += 16) {
              const int cc = c + hc;
              const uint32_t fl16 = aux->e_flag16[cc >> 4];
              if (fl16 != 0)
                dgrad_direct_sums(fl16, tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride + cc,
                                  staging_chunk(so, cw, row, cc), staging_chunk(so, cw, row, cc + 8),
                                  aux->d_thr2 + (cc >> 1), aux->d_sgn2 + (cc >> 1), row_valid, &aux->s_stat[0][cc],
                                  &aux->s_stat[1][cc], lane);
            }
          }
          if constexpr (wg) {
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_threads) : "memory");   // all raw-x reads done
            wg_transform_tile(it);
          }
        }
      }
      if constexpr (wg && !wg3) {
        // the staged result overwrites the A' tile: its weight-gradient MMAs must have completed (per-group barrier,
        // like epi_in_full: a group observes every phase of the barrier it waits on)
        ptx::mbar_wait(&aux->wg_done[sb][g2], (it / period) & 1, 22);
        ptx::tcgen05_fence_after();
      }
      for (int c = grp * 32; c < p.bn; c += n_grp * 32) {
        if (n0 + c >= p.n_total) break;
        if constexpr (dgrad) {
          // fused ReLU / BatchNorm backward: dy = acc * [es*x+eh > 0]; staged value = dy (OUT_DY) or es*dy (G modes),
          // written over the activation chunk this thread just read.  The mask is a packed-bf16 threshold test.
          // Sixteen columns at a time (register pressure).
          const bool scaled = p.out_mode != OUT_DY;
          uint32_t pkd[wg3 ? 16 : 1];   // EPI 4: the 32 columns' packed results, written after the accumulator is released
#pragma unroll
          for (int hc = 0; hc < 32; hc += 16) {
            const int cc = c + hc;
            uint32_t r16[16];
            ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride + cc, r16);
            uint4* xc0 = staging_chunk(so, cw, row, cc);
            uint4* xc1 = staging_chunk(so, cw, row, cc + 8);
            const uint4 xa = *xc0, xb = *xc1;
            const uint32_t xin[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            // per-column constants are warp-uniform 16-byte loads - each costs shared-memory wavefronts like a data row,
            // and they were as much traffic as the tile itself (ncu source view).  The fused variants need none for the
            // mask (A' > 0: thr = +0, no sign flip) and take the scale as the bf16 pairs the forward prologue used
            // (one load per eight columns instead of two)
            const uint4* thr4 = reinterpret_cast<const uint4*>(aux->e_thr2 + (cc >> 1));
            const uint4* sgn4 = reinterpret_cast<const uint4*>(aux->e_sgn2 + (cc >> 1));
            const float4* es4 = reinterpret_cast<const float4*>(aux->e_scale + cc);
            const uint4* esb4 = reinterpret_cast<const uint4*>(aux->s_scale + cc);
            ptx::tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int i4 = 0; i4 < 2; ++i4) {
              uint32_t thv[4] = {0u, 0u, 0u, 0u}, sgv[4] = {0u, 0u, 0u, 0u};
              if constexpr (!wg) {
                const uint4 th4 = thr4[i4], sg4 = sgn4[i4];
                thv[0] = th4.x; thv[1] = th4.y; thv[2] = th4.z; thv[3] = th4.w;
                sgv[0] = sg4.x; sgv[1] = sg4.y; sgv[2] = sg4.z; sgv[3] = sg4.w;
              }
              float esv[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
              if (scaled) {
                if constexpr (wg) {
                  const uint4 eb = esb4[i4];
                  esv[0] = bf16_lo(eb.x); esv[1] = bf16_hi(eb.x); esv[2] = bf16_lo(eb.y); esv[3] = bf16_hi(eb.y);
                  esv[4] = bf16_lo(eb.z); esv[5] = bf16_hi(eb.z); esv[6] = bf16_lo(eb.w); esv[7] = bf16_hi(eb.w);
                } else {
                  const float4 e0 = es4[2 * i4], e1 = es4[2 * i4 + 1];
                  esv[0] = e0.x; esv[1] = e0.y; esv[2] = e0.z; esv[3] = e0.w;
                  esv[4] = e1.x; esv[5] = e1.y; esv[6] = e1.z; esv[7] = e1.w;
                }
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int pi = 4 * i4 + j;   // column pair (cc + 2*pi, cc + 2*pi + 1)
                bool m0, m1;
                gt_bf16x2(xin[pi] ^ sgv[j], thv[j], m0, m1);
                const float a0 = __uint_as_float(r16[2 * pi]) * esv[2 * j], a1 = __uint_as_float(r16[2 * pi + 1]) * esv[2 * j + 1];
                pk[pi] = pack_bf16x2(m0 ? a0 : 0.f, m1 ? a1 : 0.f);
              }
            }
            if constexpr (wg3) {
#pragma unroll
              for (int i = 0; i < 8; ++i) pkd[hc / 2 + i] = row_valid ? pk[i] : 0u;
            } else {
              *xc0 = row_valid ? make_uint4(pk[0], pk[1], pk[2], pk[3]) : make_uint4(0u, 0u, 0u, 0u);
              *xc1 = row_valid ? make_uint4(pk[4], pk[5], pk[6], pk[7]) : make_uint4(0u, 0u, 0u, 0u);
            }
          }
          if constexpr (wg3) {
            // the accumulator is in registers: hand the stage back (the next tile's data-gradient MMAs queue behind this
            // tile's weight-gradient MMAs), then wait for those MMAs before overwriting the A' tile they read
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&aux->tmem_empty[acc]);
            if (fix && it + 1 < my_tiles) dz_transform_tile(it + 1);   // the next tile's dZ, ahead of its MMAs
            if (wg_ahead && it + 1 < my_tiles) {   // the next tile's transform, ahead of its weight-gradient MMAs
              ptx::mbar_wait(&aux->epi_in_full[(it + 1) % p.n_stg][0], (uint32_t)((it + 1) / p.n_stg) & 1u, 24);
              wg_transform_tile(it + 1);
            }
            ptx::mbar_wait(&aux->wg_done[sb][0], (it / period) & 1, 22);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *staging_chunk(so, cw, row, c + 8 * i) = make_uint4(pkd[4 * i], pkd[4 * i + 1], pkd[4 * i + 2], pkd[4 * i + 3]);
          }
          continue;
        }
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride + c, r);
        ptx::tmem_ld_wait();
        if (xmerge) {
          // out[x] = Z_0[x-1] + Z_1[x] + Z_2[x+1] (pad 1): partial sums of the neighbouring input columns sit in the
          // neighbouring lanes (16-lane segments = tile rows; the segment ends are halo columns, not stored).
          // Sixteen columns at a time keeps the live registers at 64.
#pragma unroll
          for (int hc = 0; hc < 32; hc += 16) {
            uint32_t h1[16], h2[16];
            ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride + p.bn + c + hc, h1);
            ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride + 2 * p.bn + c + hc, h2);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(r[hc + i]), 1);
              const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(h2[i]), 1);
              r[hc + i] = __float_as_uint(left + __uint_as_float(h1[i]) + right);
            }
          }
        }
        uint32_t packed[16];
        if (!dgrad) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            packed[i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            if (p.do_stats && !row_valid) packed[i] = 0u;
          }
          if (!xmerge || (xo >= 0 && xo < out_w)) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              *staging_chunk(so, cw, srow, c + 8 * i) =
                  make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
          }
          if (narrow && p.do_stats && !p.mma_stats && !col_narrow) {
            // sums, then squares re-derived from the packed values: the two 32-value arrays are never live together
            float v[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v[2 * i] = bf16_lo(packed[i]);
              v[2 * i + 1] = bf16_hi(packed[i]);
            }
            const float cs = warp_column_sums(v, lane);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float lo = bf16_lo(packed[i]), hi = bf16_hi(packed[i]);
              v[2 * i] = lo * lo;
              v[2 * i + 1] = hi * hi;
            }
            const float cq = warp_column_sums(v, lane);
            atomicAdd(&aux->s_stat[0][c + lane], cs);
            atomicAdd(&aux->s_stat[1][c + lane], cq);
          }
        }
      }
      if (leader) RXB_TL(2, it, 5);
      if constexpr (!wg3) {
        ptx::tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&aux->tmem_empty[acc]);
      }
      // hand the staged tile to the store warp (and to the MMA warp for the column statistics)
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_threads) : "memory");
      if (leader) {
        RXB_TL(2, it, 6);
        ptx::mbar_arrive(&aux->stg_full[sb]);
      }
      if constexpr (EPI == 0) {
        if (col_stats) {
          const int pr = et & 63, half = et >> 6;          // channel pair, row half
          const uint8_t* colp = so + (pr >> 5) * (128 * 128) + (pr & 3) * 4;
          const int j = (pr & 31) >> 2;                      // 16-byte chunk of the 128-byte row
#pragma unroll 8
          for (int r = half * 64; r < half * 64 + 64; ++r) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(colp + r * 128 + ((j ^ (r & 7)) << 4));
            const float lo = bf16_lo(v), hi = bf16_hi(v);
            cs0 += lo; cs1 += hi;
            cq0 = fmaf(lo, lo, cq0); cq1 = fmaf(hi, hi, cq1);
          }
          if (p.n_stg == 1) {   // shared buffer: the other group may overwrite it only after this pass
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_threads) : "memory");
            if (leader) ptx::mbar_arrive(&aux->stg_free[(it + 1) & 1]);
          }
        }
      }
      if constexpr (EPI == 1) {
        if (col_narrow) {
          const int n_rows = xmerge ? th * out_w : 128;        // staged rows of a tile (every one rewritten per tile)
          const int wq = et >> 5;                                 // warp of the group
          if (cw == 64) {
            // 64 channels = 32 pairs: lane = pair, a warp reads one 128-byte row per instruction, warps split the rows
            const uint8_t* colp = so + (lane & 3) * 4;
            const int j = lane >> 2;
            for (int r = wq; r < n_rows; r += 4) {
              const uint32_t v = *reinterpret_cast<const uint32_t*>(colp + r * 128 + ((j ^ (r & 7)) << 4));
              const float lo = bf16_lo(v), hi = bf16_hi(v);
              cs0 += lo; cs1 += hi;
              cq0 = fmaf(lo, lo, cq0); cq1 = fmaf(hi, hi, cq1);
            }
          } else {
            // 32 channels = 16 pairs: lanes 0-15 read an even row, lanes 16-31 the odd row after it (64-byte rows: the
            // two halves of one 128-byte bank line), warps split the row pairs
            const int pr = lane & 15, odd = lane >> 4;
            const uint8_t* colp = so + (pr & 3) * 4;
            const int j = pr >> 2;
            for (int q2 = wq; 2 * q2 + odd < n_rows; q2 += 4) {
              const int r = 2 * q2 + odd;
              const uint32_t v = *reinterpret_cast<const uint32_t*>(colp + r * 64 + ((j ^ ((r >> 1) & 3)) << 4));
              const float lo = bf16_lo(v), hi = bf16_hi(v);
              cs0 += lo; cs1 += hi;
              cq0 = fmaf(lo, lo, cq0); cq1 = fmaf(hi, hi, cq1);
            }
          }
        }
      }
    }
    if constexpr (EPI == 1) {
      if (col_narrow) {   // (s_stat was zeroed at kernel start; the common tail adds it to the global sums)
        const int c0 = 2 * (cw == 64 ? lane : (lane & 15));
        atomicAdd(&aux->s_stat[0][c0], cs0);
        atomicAdd(&aux->s_stat[0][c0 + 1], cs1);
        atomicAdd(&aux->s_stat[1][c0], cq0);
        atomicAdd(&aux->s_stat[1][c0 + 1], cq1);
      }
    }
    if constexpr (EPI == 0) {
      if (col_stats) {   // s_stat was zeroed at kernel start; the common tail below adds it to the global sums
        const int c0 = 2 * (et & 63);
        atomicAdd(&aux->s_stat[0][c0], cs0);
        atomicAdd(&aux->s_stat[0][c0 + 1], cs1);
        atomicAdd(&aux->s_stat[1][c0], cq0);
        atomicAdd(&aux->s_stat[1][c0 + 1], cq1);
      }
    }
    float tail_s = 0.f, tail_dyx = 0.f;   // this CTA's sum(dy) / direct sum(dy*x) of channel n0 + row (fused BatchNorm tail)
    bool tail_deg = false;
    if (!narrow && p.do_stats && p.mma_stats) {
      if (dgrad) asm volatile("bar.sync 3, %0;" ::"r"((uint32_t)n_epi_threads) : "memory");   // s_stat of both groups
      // per-channel totals of this CTA from TMEM: lane = channel; Gram diagonal and the sums column
      if (g2 == 0 && grp == 0 && my_tiles > 0) {
        ptx::mbar_wait(&aux->stats_done, 0, 12);
        ptx::tcgen05_fence_after();
        const int ch = n0 + row;
        uint32_t s16[16];
        ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + kSumC, s16);
        ptx::tmem_ld_wait();
        float total = __uint_as_float(s16[0]);
        if (dgrad) {
          if ((aux->e_flag16[row >> 4] >> (row & 15)) & 1u) {
            // degenerate channel: the direct fp32 reductions replace the tensor-pipe sum of the (scaled) staged tile
            total = aux->s_stat[0][row];
            const float dyx = aux->s_stat[1][row];
            tail_deg = true;
            tail_dyx = dyx;
            if (ch < p.n_total && dyx != 0.f && !(wg && p.tail.mode == 1)) atomicAdd(p.ch_sumsq + ch, dyx);
          } else if (p.out_mode != OUT_DY) {  // the staged value was es*dy (fused variants: es as the bf16 the epilogue applied)
            const float es = wg ? bf16_round(aux->e_scale[row]) : aux->e_scale[row];
            total = es != 0.f ? total / es : 0.f;
          }
        } else {
          uint32_t g[32];
          ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + kGramCol + q * 32, g);
          ptx::tmem_ld_wait();
          float sq = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) sq = lane == i ? __uint_as_float(g[i]) : sq;
          if (ch < p.n_total && sq != 0.f) atomicAdd(p.ch_sumsq + ch, sq);
        }
        tail_s = total;
        if (ch < p.n_total && total != 0.f && !(wg && p.tail.mode == 1)) atomicAdd(p.ch_sum + ch, total);
      }
    } else if (p.do_stats) {
      asm volatile("bar.sync 3, %0;" ::"r"((uint32_t)n_epi_threads) : "memory");   // both groups
      for (int c = threadIdx.x - first_epi_warp * 32; c < p.bn && n0 + c < p.n_total; c += n_epi_threads) {
        const float a = aux->s_stat[0][c], bq = aux->s_stat[1][c];
        if (a != 0.f) atomicAdd(p.ch_sum + n0 + c, a);
        if (bq != 0.f) atomicAdd(p.ch_sumsq + n0 + c, bq);
      }
    }
    if constexpr (wg) {
      // fused weight gradient out: TMEM lane = input channel c of this N tile, column = dY channel k.  Every MMA of
      // the CTA has completed, so the A pipeline stages are dead: the fp32 result is transposed into them in torch's
      // OIHW order (row k = 128 consecutive input channels) and leaves as one bulk L2 reduce-add per k.  (The staging
      // area runs from the A stages into the weight area that follows them - both dead.)
      if (wg3 && g2 == 0 && grp == 0 && my_tiles > 0) {
        // 3x3: accumulator group (ty, j) holds tap (ty, 2-j), lane = input channel c, column = output channel n.  OIHW
        // row n = [c][tap] is 128*9 floats; sixteen rows (72 KB) are staged at a time.
        ptx::mbar_wait(&aux->wg_final, 0, 21);
        ptx::tcgen05_fence_after();
        if (p.tail.mode == 2) {
          // W.dW of this CTA's partial, per input channel c = row: the accumulators against the bf16 weight panel still
          // resident in shared memory (slot 8-tp, row c, 32 output channels n in four 64B-swizzled chunks)
          float t = 0.f;
          for (int tp = 0; tp < 9; ++tp) {
            const int tyy = tp / 3, txx = tp - tyy * 3;
            const uint8_t* wrow = smB + (size_t)(8 - tp) * b_tap + row * 64;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t r16[16];
              ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + kWgCol + (tyy * 3 + (2 - txx)) * 32 + half * 16, r16);
              const uint4 w0 = *reinterpret_cast<const uint4*>(wrow + (((2 * half) ^ ((row >> 1) & 3)) << 4));
              const uint4 w1 = *reinterpret_cast<const uint4*>(wrow + (((2 * half + 1) ^ ((row >> 1) & 3)) << 4));
              const uint32_t wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                t = fmaf(bf16_lo(wv[i]), __uint_as_float(r16[2 * i]), t);
                t = fmaf(bf16_hi(wv[i]), __uint_as_float(r16[2 * i + 1]), t);
              }
            }
          }
          if (!tail_deg && t != 0.f) atomicAdd(p.ch_sumsq + row, t);
        }
        float* stg = reinterpret_cast<float*>(smA);
        const int rowlen = 128 * 9;
        for (int half = 0; half < 2; ++half) {
          if (half) {   // the reduce-adds of the first half must be done with the staging area
            ptx::tma_store_wait_read();
            asm volatile("bar.sync 4, 128;" ::: "memory");
          }
          for (int tp = 0; tp < 9; ++tp) {
            const int tyy = tp / 3, txx = tp - tyy * 3;
            uint32_t r16[16];
            ptx::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + kWgCol + (tyy * 3 + (2 - txx)) * 32 + half * 16, r16);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) stg[i * rowlen + row * 9 + tp] = __uint_as_float(r16[i]);
          }
          ptx::fence_proxy_async_smem();
          asm volatile("bar.sync 4, 128;" ::: "memory");
          if (et < 16)
            ptx::bulk_reduce_add_f32(p.wg_dW + (long long)(half * 16 + et) * rowlen, stg + et * rowlen, (uint32_t)rowlen * 4u);
          ptx::tma_store_commit();
        }
        ptx::tma_store_wait_all();
      } else if (!wg3 && g2 == 0 && grp == 0 && my_tiles > 0) {
        ptx::mbar_wait(&aux->wg_final, 0, 21);
        ptx::tcgen05_fence_after();
        float* stg = reinterpret_cast<float*>(smA);
        const int kcols = p.kb_per_tap * BK;
        const bool ch_ok = n0 + row < p.n_total;
        const bool want_t = p.tail.mode == 1 && ch_ok && !tail_deg;
        float t = 0.f;   // W.dW of this CTA's partial for input channel n0 + row (fused BatchNorm tail)
        for (int c = 0; c < kcols; c += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + kWgCol + c, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) stg[(c + i) * 128 + row] = __uint_as_float(r[i]);
          if (want_t) {
            const float* wp = p.tail.W + (long long)c * p.n_total + n0 + row;   // W[k][channel], k = c + i
#pragma unroll
            for (int hb = 0; hb < 32; hb += 16) {
              float wv[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) wv[i] = c + hb + i < p.cin ? __ldg(wp + (long long)(hb + i) * p.n_total) : 0.f;
#pragma unroll
              for (int i = 0; i < 16; ++i) t = fmaf(bf16_round(wv[i]), __uint_as_float(r[hb + i]), t);
            }
          }
        }
        if (p.tail.mode == 1 && ch_ok) {
          // bn_bwd_finalize's arithmetic on this CTA's partial sums (linear, so the CTAs' contributions add up)
          const int ch = n0 + row;
          const float es = bf16_round(aux->e_scale[row]), eh = bf16_round(aux->e_shift[row]);
          const float raw = tail_deg ? tail_dyx : (es != 0.f ? (t - eh * tail_s) / es : 0.f);
          const float qv = p.tail.rstd[ch] * (raw - p.tail.mean[ch] * tail_s);
          const float sc = es * p.tail.inv_count;   // the scale the staged gradient carried (bf16, like the forward's fold)
          if (qv != 0.f) atomicAdd(p.tail.dgamma + ch, qv);
          if (tail_s != 0.f) atomicAdd(p.tail.dbeta + ch, tail_s);
          if (tail_s != 0.f) atomicAdd(p.tail.corrA + ch, sc * tail_s);
          if (qv != 0.f) atomicAdd(p.tail.corrB + ch, sc * qv);
        }
        ptx::fence_proxy_async_smem();
        asm volatile("bar.sync 4, 128;" ::: "memory");
        const int valid = min(128, p.n_total - n0);
        for (int k = et; k < p.cin; k += 128)
          ptx::bulk_reduce_add_f32(p.wg_dW + (long long)k * p.n_total + n0, stg + k * 128, (uint32_t)valid * 4u);
        ptx::tma_store_commit();
        ptx::tma_store_wait_all();
      }
    }
  } else {
    // =============================== workers 0-7: A-operand transform (pre-activation BatchNorm + ReLU)
    if constexpr (EPI == 3) {
      // fused 1x1 weight gradient: every activation tile becomes A' = relu(bn(x)) in place as soon as it lands and is
      // handed to the MMA warp.  (CTAs with degenerate BatchNorm channels leave it to the owning epilogue group, which
      // needs the raw x first.)
      const uint4 any4 = *reinterpret_cast<const uint4*>(aux->e_flag_any4);
      if ((any4.x | any4.y | any4.z | any4.w) == 0) {
        const int t = threadIdx.x - kWorker0 * 32;  // 0..kXformThreads-1
        const int period = (p.n_stg & 1) ? 2 * p.n_stg : p.n_stg;
        // this thread's fold constants (its 16-byte chunk of both 64-channel boxes) stay in registers for the life of
        // the CTA: re-loading them per tile cost a fifth of the transform's shared-memory wavefronts
        uint32_t fs[2][4], fh[2][4];
        for (int bx = 0; bx < 2; ++bx) {
          const uint4 s4 = *reinterpret_cast<const uint4*>(aux->s_scale + bx * 64 + (t & 7) * 8);
          const uint4 h4 = *reinterpret_cast<const uint4*>(aux->s_shift + bx * 64 + (t & 7) * 8);
          fs[bx][0] = s4.x; fs[bx][1] = s4.y; fs[bx][2] = s4.z; fs[bx][3] = s4.w;
          fh[bx][0] = h4.x; fh[bx][1] = h4.y; fh[bx][2] = h4.z; fh[bx][3] = h4.w;
        }
        for (int it = 0; it < my_tiles; ++it) {
          int x0, y0, b0;
          tile_origin(p.t, blockIdx.x + it * gridDim.x, x0, y0, b0);
          const int sb = it % p.n_stg;
          uint8_t* so = st_out + (size_t)sb * stage_tile;
          ptx::mbar_wait(&aux->epi_in_full[sb][it & 1], (uint32_t)(it / period) & 1u, 25);
#pragma unroll
          for (int bx = 0; bx < 2; ++bx)
            if (bx < n_boxes)
              transform_box_sw128_regs(so + bx * (128 * 128), 128, fs[bx], fh[bx], t, p.t, tw, th, x0, y0, b0);
          ptx::fence_proxy_async_smem();     // every writing thread orders its stores before the MMA's async reads
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&aux->xa_ready[sb]);
        }
      }
    }
    if (PROLOGUE) {
      const int t = threadIdx.x - kWorker0 * 32;  // 0..kXformThreads-1
      int stage = 0;
      uint32_t phase = 0;
      // (measured and rejected: keeping the fold constants of up to two k-blocks in registers instead of re-loading them
      // per stage - 3x3 forward 0.293 -> 0.353 ms, the role's registers spill under the kernel's 96-register cap)
      for (int m_tile = blockIdx.x; m_tile < m_tiles; m_tile += gridDim.x) {
        int x0, y0, b0;
        tile_origin(p.t, m_tile, x0, y0, b0);
        for (int g = 0; g < groups; ++g) {
          const int gy = p.halo ? 0 : g / p.taps_x, gx = p.halo >= 2 ? 0 : p.halo == 1 ? g : g - gy * p.taps_x;
          for (int kb = 0; kb < p.kb_per_tap; ++kb) {
            ptx::mbar_wait(&aux->full[stage], phase, 5);
            transform_box_sw128(smA + (size_t)stage * a_stage, p.rows_a, aux->s_scale + kb * BK, aux->s_shift + kb * BK,
                                t, p.t, box_w, box_h, x0 + gx - p.pad_x, y0 + gy - p.pad_y, b0);
            ptx::fence_proxy_async_smem();     // every writing thread orders its stores before the MMA's async reads
            __syncwarp();
            if ((threadIdx.x & 31) == 0) ptx::mbar_arrive(&aux->xform[stage]);
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  }

  // ---- teardown
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================================================
// Weight gradient
struct __align__(16) WgradAux {
  __nv_bfloat16 s_scale[kMaxPrologueC + 64];   // prologue fold as bf16 pairs: the operands of fma.rn.relu.bf16x2
  __nv_bfloat16 s_shift[kMaxPrologueC + 64];
  uint64_t full[kMaxStages];
  uint64_t xform[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t d_full[2];    // classic mode: the dOut tile of a pixel tile (loaded once, shared by all channel chunks)
  uint64_t d_empty[2];
  uint64_t tmem_full;
  uint32_t tmem_base;
  uint32_t pad;
};

constexpr int kWgA_BYTES = 128 * 128 * 2;  // 128 pixels x 128 channels

__global__ void __launch_bounds__(kGemmThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmD,
                  const WgradParams p, const int stages) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int taps = p.taps_x * p.taps_y;
  const int d_tile = 128 * p.n * 2;                     // one dOut tile
  const int nd = p.shift_dout ? taps : 1;               // dOut tiles (accumulator groups) per stage
  // shift_dout == 2: ONE dOut box with the full (tw+taps_x-1) x (th+taps_y-1) halo serves every tap
  const int halo_w = (1 << p.t.tw_log2) + p.taps_x - 1, halo_h = (1 << p.t.th_log2) + p.taps_y - 1;
  const int halo_tx = halo_w * halo_h * p.n * 2;        // bytes TMA delivers for the halo box
  const int b_bytes = p.shift_dout == 2 ? ((halo_tx + 1023) & ~1023) : nd * d_tile;
  uint8_t* smA = smem;
  uint8_t* smB = smem + (size_t)stages * kWgA_BYTES;
  // classic mode: two dOut tile buffers (one load per pixel tile); shifted-dOut modes: one dOut group per stage
  const int n_dbuf = p.shift_dout ? stages : 2;
  WgradAux* aux = reinterpret_cast<WgradAux*>(smB + (size_t)n_dbuf * b_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = p.t.tiles_x * p.t.tiles_y * p.t.tiles_b;
  const int total_boxes = (p.shift_dout ? 1 : taps) * p.boxes_per_tap;
  const int chunk0 = blockIdx.y * p.chunks_per_cta;
  // accumulators held by this CTA: channel chunks (classic) or filter taps (shift_dout)
  const int n_local = p.shift_dout ? taps : min(p.chunks_per_cta, p.n_chunks - chunk0);
  const int stages_per_tile = (p.shift_dout || p.a_halo) ? 1 : n_local;
  const int a_halo_tx = halo_w * halo_h * p.bkc * 2;     // a_halo: bytes of the full-halo A box
  const int tile_begin = blockIdx.x * p.pix_tiles_per_cta;
  const int tile_end = min(m_tiles, tile_begin + p.pix_tiles_per_cta);
  const int a_row_bytes = p.bkc * 2;
  const int a_box_bytes = 128 * a_row_bytes;
  const int d_boxes = p.n >= 64 ? p.n / 64 : 1;
  const int th = 1 << p.t.th_log2;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmD);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&aux->full[s], 1);
      ptx::mbar_init(&aux->xform[s], kXformThreads / 32);   // one arrival per transform warp
      ptx::mbar_init(&aux->empty[s], 1);
    }
    ptx::mbar_init(&aux->tmem_full, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&aux->d_full[a], 1);
      ptx::mbar_init(&aux->d_empty[a], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(&aux->tmem_base);
  pdl_sync();   // CTA-private setup above overlaps the previous kernel; global memory from here on
  if (p.prologue) {
    const int padded = p.boxes_per_tap * p.bkc;
    for (int c = threadIdx.x; c < padded; c += kGemmThreads) {
      aux->s_scale[c] = __float2bfloat16_rn(c < p.cin ? p.scale[c] : 0.f);
      aux->s_shift[c] = __float2bfloat16_rn(c < p.cin ? p.shift[c] : 0.f);
    }
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = aux->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        int x0, y0, b0;
        tile_origin(p.t, tile, x0, y0, b0);
        RXB_TL(0, tile - tile_begin, 0);
        if (!p.shift_dout) {
          const int tc = tile - tile_begin, db = tc & 1;
          ptx::mbar_wait(&aux->d_empty[db], ((tc >> 1) & 1) ^ 1, 17);
          RXB_TL(0, tc, 1);
          ptx::mbar_arrive_expect_tx(&aux->d_full[db], d_tile);
          for (int j = 0; j < d_boxes; ++j)
            ptx::tma_load_4d(smB + (size_t)db * d_tile + (size_t)j * 16384, &tmD, &aux->d_full[db], j * 64, x0, y0, b0);
        }
        for (int cl = 0; cl < stages_per_tile; ++cl) {
          ptx::mbar_wait(&aux->empty[stage], phase ^ 1, 11);
          ptx::mbar_arrive_expect_tx(&aux->full[stage], (p.a_halo ? a_halo_tx : kWgA_BYTES) +
                                                            (p.shift_dout == 2 ? halo_tx : p.shift_dout ? b_bytes : 0));
          uint8_t* a_dst = smA + (size_t)stage * kWgA_BYTES;
          if (p.a_halo) {   // dW[t] = sum_q dOut[q] * A[q + t - pad]: the box starts pad pixels before the tile
            ptx::tma_load_4d(a_dst, &tmA, &aux->full[stage], 0, x0 - p.pad_x, y0 - p.pad_y, b0);
          } else
          for (int i = 0; i < p.boxes_per_chunk; ++i) {
            const int kk = (chunk0 + cl) * p.boxes_per_chunk + i;
            int tp = 0, c0 = p.boxes_per_tap * p.bkc;  // fully out of bounds -> zero box
            if (kk < total_boxes) {
              tp = kk / p.boxes_per_tap;
              c0 = (kk - tp * p.boxes_per_tap) * p.bkc;
            }
            int ax = x0, ay = y0;
            if (!p.shift_dout) {
              const int ty = tp / p.taps_x, tx = tp - ty * p.taps_x;
              ax += tx - p.pad_x;
              ay += ty - p.pad_y;
            }
            ptx::tma_load_4d(a_dst + (size_t)i * a_box_bytes, &tmA, &aux->full[stage], c0, ax, ay, b0);
          }
          uint8_t* d_dst = smB + (size_t)stage * b_bytes;
          if (p.shift_dout == 2) {
            // dW[t] = sum_q A'[q] * dOut[q - (t - pad)]: the box starts (taps-1-pad) pixels before the tile
            ptx::tma_load_4d(d_dst, &tmD, &aux->full[stage], 0, x0 - (p.taps_x - 1 - p.pad_x), y0 - (p.taps_y - 1 - p.pad_y), b0);
          } else if (p.shift_dout) {
            for (int t = 0; t < nd; ++t) {  // dW[t] = sum_q A'[q] * dOut[q - (t - pad)]
              const int ty = t / p.taps_x, tx = t - ty * p.taps_x;
              const int dx = x0 - (tx - p.pad_x), dy = y0 - (ty - p.pad_y);
              for (int j = 0; j < d_boxes; ++j)
                ptx::tma_load_4d(d_dst + (size_t)t * d_tile + (size_t)j * 16384, &tmD, &aux->full[stage], j * 64, dx, dy, b0);
            }
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        RXB_TL(0, tile - tile_begin, 2);
      }
    }
  } else if (warp == 1) {
    // whole warp walks the pipeline; one elected lane issues (keeps the descriptor arithmetic warp-uniform)
    {
      // Shifted-dOut mode with Cout <= 64: one dOut tile is one swizzle atom along N and the taps' tiles lie d_tile
      // bytes apart - exactly the descriptor's leading-dimension stride - so the taps_x tiles of a filter row are
      // ONE MMA of N = taps_x*Cout whose accumulator columns are the per-tap accumulators side by side.
      const int tiles_per_mma = (p.shift_dout && p.n <= 64 && p.taps_x * p.n <= 256) ? p.taps_x : 1;
      // Full-halo mode: tap (ty,tx) reads the halo box from pixel row (taps_y-1-ty)*halo_w + (taps_x-1-tx); an 8-pixel
      // K group is one tile row (tw = 8), consecutive groups lie halo_w rows apart, and the taps_x taps of a filter row
      // are N atoms ONE pixel row apart (leading-dimension stride = one row), in DESCENDING tx order - so accumulator
      // group (ty, j) holds tap (ty, taps_x-1-j).  (MN-major operands with unaligned starts: probed on B200,
      // profiles/r01_umma_probe_unaligned_operands.log.)
      const uint32_t idesc = ptx::make_idesc_bf16(128, p.n * tiles_per_mma, 1, 1);
      const uint32_t a_swz = p.bkc == 64 ? ptx::kSwizzle128B : ptx::kSwizzle64B;
      const uint32_t a_sbo = 8 * a_row_bytes;
      const uint32_t a_kstep16 = (16 * a_row_bytes) >> 4;
      const uint32_t d_row_bytes = p.n >= 64 ? 128 : 64;
      const uint32_t d_swz = p.n >= 64 ? ptx::kSwizzle128B : ptx::kSwizzle64B;
      const uint32_t d_sbo = 8 * d_row_bytes;
      const uint32_t d_kstep16 = (16 * d_row_bytes) >> 4;
      const uint32_t d_lbo = 128 * d_row_bytes;
      const uint64_t da0 = ptx::make_smem_desc(ptx::smem_u32(smA), a_box_bytes, a_sbo, a_swz);
      const uint64_t db0 = p.shift_dout == 2
                               ? ptx::make_smem_desc(ptx::smem_u32(smB), d_row_bytes, (uint32_t)halo_w * d_row_bytes, d_swz)
                               : ptx::make_smem_desc(ptx::smem_u32(smB), d_lbo, d_sbo, d_swz);
      const uint32_t d_row16 = d_row_bytes >> 4;
      const uint32_t d_kstep16_h = (2u * (uint32_t)halo_w * d_row_bytes) >> 4;   // 16 pixels = two tile rows
      const uint32_t a_hi = ptx::desc_hi(da0), b_hi = ptx::desc_hi(db0);
      const uint32_t a_stage16 = kWgA_BYTES >> 4, b_stage16 = (uint32_t)b_bytes >> 4, d_tile16 = (uint32_t)d_tile >> 4;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int tc = tile - tile_begin, db = tc & 1;
        if (lane == 0) RXB_TL(1, tc, 0);
        if (!p.shift_dout) ptx::mbar_wait(&aux->d_full[db], (tc >> 1) & 1, 18);
        if (lane == 0) RXB_TL(1, tc, 1);
        for (int cl = 0; cl < stages_per_tile; ++cl) {
          ptx::mbar_wait(p.prologue ? &aux->xform[stage] : &aux->full[stage], phase, 13);
          ptx::tcgen05_fence_after();
          if (lane == 0 && cl < 4) RXB_TL(1, tc, 2 + cl);
          if (ptx::elect_one()) {
            const uint32_t a_lo = ptx::desc_lo(da0) + (uint32_t)stage * a_stage16;
            const uint32_t b_lo = ptx::desc_lo(db0) + (uint32_t)(p.shift_dout ? stage : db) * b_stage16;
            const uint32_t accumulate = tile > tile_begin ? 1u : 0u;
            if (p.a_halo) {
              // chunk ty = filter row: its taps_x taps are M atoms one pixel row apart (leading stride = one row),
              // an 8-pixel K group is one tile row, consecutive groups lie halo_w rows apart
              const uint64_t dah = ptx::make_smem_desc(0, (uint32_t)a_row_bytes, (uint32_t)halo_w * a_row_bytes, a_swz);
              const uint32_t ah_hi = ptx::desc_hi(dah);
              const uint32_t ah_lo = ptx::desc_lo(dah) + (ptx::smem_u32(smA) >> 4) + (uint32_t)stage * a_stage16;
              const uint32_t ah_kstep16 = (2u * (uint32_t)halo_w * a_row_bytes) >> 4;
              for (int ty = 0; ty < n_local; ++ty) {
                const uint32_t acc = tmem_base + (uint32_t)ty * p.n;
                const uint32_t a_lo_t = ah_lo + (((uint32_t)((chunk0 + ty) * halo_w) * a_row_bytes) >> 4);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                  ptx::umma_bf16_ss_parts(acc, a_lo_t + ks * ah_kstep16, ah_hi, b_lo + ks * d_kstep16, b_hi, idesc,
                                          ks > 0 ? 1u : accumulate);
              }
            } else if (p.shift_dout == 2) {
              for (int ty = 0; ty < p.taps_y; ++ty) {
                const uint32_t acc = tmem_base + (uint32_t)(ty * p.taps_x) * p.n;
                const uint32_t b_lo_t = b_lo + (uint32_t)((p.taps_y - 1 - ty) * halo_w) * d_row16;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                  ptx::umma_bf16_ss_parts(acc, a_lo + ks * a_kstep16, a_hi, b_lo_t + ks * d_kstep16_h, b_hi, idesc,
                                          ks > 0 ? 1u : accumulate);
              }
            } else
            for (int t = 0; t < nd; t += tiles_per_mma) {
              const uint32_t acc = tmem_base + (p.shift_dout ? t : cl) * p.n;
              const uint32_t b_lo_t = b_lo + (uint32_t)t * d_tile16;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                ptx::umma_bf16_ss_parts(acc, a_lo + ks * a_kstep16, a_hi, b_lo_t + ks * d_kstep16, b_hi, idesc,
                                        ks > 0 ? 1u : accumulate);
            }
            ptx::umma_commit(&aux->empty[stage]);
            if (!p.shift_dout && cl == stages_per_tile - 1) ptx::umma_commit(&aux->d_empty[db]);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
      if (ptx::elect_one()) ptx::umma_commit(&aux->tmem_full);
      __syncwarp();
    }
  } else if (warp < 6) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (tile_end > tile_begin) {
      if (threadIdx.x == 64) RXB_TL(2, 0, 0);
      ptx::mbar_wait(&aux->tmem_full, 0, 14);
      ptx::tcgen05_fence_after();
      const int et = threadIdx.x - 64;
      if (et == 0) RXB_TL(2, 0, 1);
      if (p.bulk_out) {
        // Every MMA of this CTA has completed, so the pipeline stages are dead: the fp32 result is transposed into
        // them in the order torch's OIHW gradient has in memory and leaves as contiguous L2 reduce-adds issued by
        // the TMA engine (cp.reduce.async.bulk), one per output channel n, instead of 4-byte atomics.
        float* stg = reinterpret_cast<float*>(smem);
        if (p.shift_dout) {
          const int rowlen = p.cin * taps;   // floats of one n row: [cin][taps]
          for (int tp = 0; tp < taps; ++tp) {
            // accumulator group of tap tp (full-halo mode keeps a filter row's taps in descending tx order)
            const int tyy = tp / p.taps_x, txx = tp - tyy * p.taps_x;
            const int grp_col = (p.shift_dout == 2 ? tyy * p.taps_x + (p.taps_x - 1 - txx) : tp) * p.n;
            for (int c = 0; c < p.n; c += 32) {
              uint32_t r[32];
              ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + grp_col + c, r);
              ptx::tmem_ld_wait();
              if (row < p.cin) {
#pragma unroll
                for (int i = 0; i < 32; ++i) stg[(size_t)(c + i) * rowlen + row * taps + tp] = __uint_as_float(r[i]);
              }
            }
          }
          ptx::fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int n = et; n < p.n; n += 128)
            ptx::bulk_reduce_add_f32(p.dW + (long long)(p.n_off + n) * rowlen, stg + (size_t)n * rowlen, rowlen * 4);
          ptx::tma_store_commit();
        } else if (taps > 1) {
          // classic multi-tap mode into the TAP-MAJOR scratch layout dW[tap][n][cin] (w_mode 3): chunk cl holds the two
          // 64-channel boxes kk = (chunk0+cl)*2 + {0,1}, each a contiguous run of one (tap, n) row of the scratch
          const int nbuf = p.bulk_bufs;
          for (int cl = 0; cl < n_local; ++cl) {
            float* sb = stg + (size_t)(cl % nbuf) * p.n * 128;
            if (cl >= nbuf) {
              if (nbuf == 2) ptx::tma_store_wait_read_pending<1>(); else ptx::tma_store_wait_read_pending<0>();
              asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            for (int c = 0; c < p.n; c += 32) {
              uint32_t r[32];
              ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + cl * p.n + c, r);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) sb[(c + i) * 128 + row] = __uint_as_float(r[i]);
            }
            ptx::fence_proxy_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int ib2 = 0; ib2 < 2; ++ib2) {
              const int kk = (chunk0 + cl) * 2 + ib2;
              if (kk >= total_boxes) break;
              const int tp = kk / p.boxes_per_tap;
              const int c0 = (kk - tp * p.boxes_per_tap) * 64;
              const int valid = min(64, p.cin - c0);
              if (valid <= 0) continue;
              for (int n = et; n < p.n; n += 128)
                ptx::bulk_reduce_add_f32(p.dW + ((long long)tp * p.cout_total + p.n_off + n) * p.cin + c0, sb + n * 128 + ib2 * 64,
                                         valid * 4);
            }
            ptx::tma_store_commit();
          }
        } else {
          // 1x1: chunk cl holds input channels [ (chunk0+cl)*128, +128 )
          const int nbuf = p.bulk_bufs;
          for (int cl = 0; cl < n_local; ++cl) {
            const int ch0 = (chunk0 + cl) * 128;
            const int valid = min(128, p.cin - ch0);
            float* sb = stg + (size_t)(cl % nbuf) * p.n * 128;
            if (cl >= nbuf) {  // the reduce that read this buffer nbuf chunks ago must be done with shared memory
              if (nbuf == 2) ptx::tma_store_wait_read_pending<1>(); else ptx::tma_store_wait_read_pending<0>();
              asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            for (int c = 0; c < p.n; c += 32) {
              uint32_t r[32];
              ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + cl * p.n + c, r);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) sb[(c + i) * 128 + row] = __uint_as_float(r[i]);
            }
            ptx::fence_proxy_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (valid > 0)
              for (int n = et; n < p.n; n += 128)
                ptx::bulk_reduce_add_f32(p.dW + (long long)(p.n_off + n) * p.cin + ch0, sb + n * 128, valid * 4);
            ptx::tma_store_commit();
          }
        }
        ptx::tma_store_wait_all();
        if (et == 0) RXB_TL(2, 0, 2);
      } else {
        const int ib = row / p.bkc, ch_in = row - ib * p.bkc;
        for (int cl = 0; cl < n_local; ++cl) {
          int tp, ch;
          bool ok;
          if (p.shift_dout) {
            tp = cl;
            if (p.shift_dout == 2) {  // accumulator group cl = (ty, j) holds tap (ty, taps_x-1-j)
              const int tyy = cl / p.taps_x, jj = cl - tyy * p.taps_x;
              tp = tyy * p.taps_x + (p.taps_x - 1 - jj);
            }
            ch = row;
            ok = ch < p.cin;
          } else {
            const int kk = (chunk0 + cl) * p.boxes_per_chunk + ib;
            tp = kk / p.boxes_per_tap;
            ch = (kk - tp * p.boxes_per_tap) * p.bkc + ch_in;
            ok = kk < total_boxes && ch < p.cin;
          }
          long long base = 0, nstride = 0;
          if (p.w_mode == 0) {
            base = (long long)ch * taps + tp;
            nstride = (long long)p.cin * taps;
          } else if (p.w_mode == 2) {
            // 3x3 stride-2 weights behind the 2x2-tap space-to-depth form (resnet.cu): tap (sy,sx) of 2x2,
            // ch = (py*2+px)*K + c with K = cin/4  ->  W[n][c][dy][dx], dy = 2*sy+py-1, dx = 2*sx+px-1
            const int K4 = p.cin >> 2;
            const int sy = tp / p.taps_x, sx = tp - sy * p.taps_x;
            const int qq = ch / K4, c = ch - qq * K4;
            const int dy = 2 * sy + (qq >> 1) - 1, dx = 2 * sx + (qq & 1) - 1;
            ok = ok && dy >= 0 && dy < 3 && dx >= 0 && dx < 3;
            base = ((long long)c * 3 + dy) * 3 + dx;
            nstride = (long long)K4 * 9;
          } else {  // space-to-depth stem: tap (sy,sx) of 4x4, ch = (py*2+px)*8 + c  ->  W[n][c][dy][dx], 7x7, 6 ch
            const int sy = tp / p.taps_x, sx = tp - sy * p.taps_x;
            const int c = ch & 7, px = (ch >> 3) & 1, py = (ch >> 4) & 1;
            const int dy = 2 * sy + py - 1, dx = 2 * sx + px - 1;
            ok = ok && c < 6 && dy >= 0 && dy < 7 && dx >= 0 && dx < 7;
            base = ((long long)c * 7 + dy) * 7 + dx;
            nstride = 6 * 49;
          }
          for (int c = 0; c < p.n; c += 32) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + cl * p.n + c, r);
            ptx::tmem_ld_wait();
            if (ok) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float g = __uint_as_float(r[i]);
                if (g != 0.f) atomicAdd(p.dW + (long long)(p.n_off + c + i) * nstride + base, g);
              }
            }
          }
        }
      }
    }
  } else {
    if (p.prologue) {
      const int t = threadIdx.x - 192;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        int x0, y0, b0;
        tile_origin(p.t, tile, x0, y0, b0);
        for (int cl = 0; cl < stages_per_tile; ++cl) {
          ptx::mbar_wait(&aux->full[stage], phase, 15);
          for (int i = 0; i < p.boxes_per_chunk; ++i) {
            const int kk = (chunk0 + cl) * p.boxes_per_chunk + i;
            if (kk >= total_boxes) continue;
            const int tp = kk / p.boxes_per_tap;
            const int c0 = (kk - tp * p.boxes_per_tap) * p.bkc;
            int ax = x0, ay = y0;
            if (!p.shift_dout) {
              const int ty = tp / p.taps_x, tx = tp - ty * p.taps_x;
              ax += tx - p.pad_x;
              ay += ty - p.pad_y;
            }
            transform_box_sw128(smA + (size_t)stage * kWgA_BYTES + (size_t)i * a_box_bytes, 128, aux->s_scale + c0,
                                aux->s_shift + c0, t, p.t, 1 << p.t.tw_log2, th, ax, ay, b0);
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if ((threadIdx.x & 31) == 0) ptx::mbar_arrive(&aux->xform[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ host
static CUtensorMapSwizzle swizzle_for(int box_c) {
  return box_c * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : box_c * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                         : CU_TENSOR_MAP_SWIZZLE_32B;
}

// activation map: dims (channels, W, H, B) with channel stride 1 and pixel stride ld; box (box_c, tw, box_h, tb)
static int make_act_tmap(CUtensorMap* tm, const void* base, const PixelTiling& t, int channels, long long ld,
                         int box_c, int box_h, int box_w = 0) {
  uint64_t dims[4] = {(uint64_t)channels, (uint64_t)t.W, (uint64_t)t.H, (uint64_t)t.B};
  uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * t.W, (uint64_t)ld * 2 * t.W * t.H};
  uint32_t box[4] = {(uint32_t)box_c, box_w ? (uint32_t)box_w : 1u << t.tw_log2, (uint32_t)box_h, 1u << t.tb_log2};
  return make_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box,
                   swizzle_for(box_c));
}

int launch_conv_gemm(GemmParams p, const void* A, long long ldA, const void* Wt, void* out, long long ldc,
                     int c_off, const void* X, long long ldx, int bk, bool prologue, cudaStream_t stream) {
  if (!(bk == 64 || bk == 32)) return set_error(RXB_ERR_INVALID, "conv_gemm: bk must be 32 or 64");
  if (prologue && bk != 64) return set_error(RXB_ERR_INVALID, "conv_gemm: prologue needs bk=64");
  if (p.n_total % 32 || p.n_total < 32) return set_error(RXB_ERR_INVALID, "conv_gemm: n_total=%d", p.n_total);
  if ((ldA * 2) % 16 || (reinterpret_cast<uintptr_t>(A) & 15) || (ldc * 2) % 16 || (c_off * 2) % 16 ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(RXB_ERR_INVALID, "conv_gemm: tensors must be 16-byte aligned with channel strides multiple of 8");
  if ((p.cin * 2) % 16) return set_error(RXB_ERR_INVALID, "conv_gemm: cin must be a multiple of 8");
  const bool dgrad = p.epi_mode == EPI_DGRAD_BN;
  if (dgrad && (!X || (ldx * 2) % 16 || (reinterpret_cast<uintptr_t>(X) & 15)))
    return set_error(RXB_ERR_INVALID, "conv_gemm: dgrad epilogue needs an aligned X");

  if (dgrad && prologue) return set_error(RXB_ERR_INVALID, "conv_gemm: the dgrad epilogue has no A prologue");
  if (dgrad && p.n_total < 64) return set_error(RXB_ERR_INVALID, "conv_gemm: dgrad epilogue needs n_total >= 64");
  const bool wg3 = p.wg_dW != nullptr && p.taps_x == 3 && p.taps_y == 3;
  if (p.tail.mode != 0) {
    if (p.wg_dW == nullptr || !(p.tail.mode == 1 || p.tail.mode == 2) || (p.tail.mode == 2) != wg3 || !p.do_stats)
      return set_error(RXB_ERR_INVALID, "conv_gemm: the fused BatchNorm tail needs the fused weight gradient (mode 1: 1x1, mode 2: 3x3)");
    if (p.tail.mode == 1 && !(p.tail.W && p.tail.mean && p.tail.rstd && p.tail.dgamma && p.tail.dbeta && p.tail.corrA && p.tail.corrB))
      return set_error(RXB_ERR_INVALID, "conv_gemm: fused BatchNorm tail (mode 1): null pointer");
    if (p.tail.mode == 2 && p.ch_sumsq == nullptr)
      return set_error(RXB_ERR_INVALID, "conv_gemm: fused BatchNorm tail (mode 2) needs ch_sumsq");
  }
  if (p.wg_dW != nullptr && !wg3 && !(dgrad && p.taps_x == 1 && p.taps_y == 1 && bk == 64 && p.cin <= 128))
    return set_error(RXB_ERR_INVALID, "conv_gemm: the fused weight gradient is for 1x1 data gradients with cin <= 128");
  if (wg3 && !(dgrad && bk == 32 && p.cin == 32 && p.n_total == 128 && p.pad_x == 1 && p.pad_y == 1 &&
               conv_dgrad3x3_wgrad_fusable(p.B, p.H, p.W)))
    return set_error(RXB_ERR_INVALID, "conv_gemm: the fused 3x3 weight gradient needs 32 -> 128 channels, pad 1 and an "
                                      "image the 8x16 full-halo tiling covers (H > 8, W > 4)");
  // dgrad, and stores of >= 128 channels with statistics, run 128-wide N tiles whose column sums come from the
  // tensor pipe (columns past n_total are zero weights / clipped stores)
  if (p.e_gamma != nullptr && (p.e_beta == nullptr || p.ch_sumsq == nullptr))
    return set_error(RXB_ERR_INVALID, "conv_gemm: e_gamma needs e_beta and ch_sumsq");
  p.bn = (dgrad || p.n_total >= kMaxBN) ? kMaxBN : p.n_total;
  // statistics of 128-wide store tiles: by default a pass of the epilogue group over the staged tile (col_stats in the
  // kernel); RXB_DBG_GRAM_STATS=1 restores the Gram / column-sum MMAs on the tensor pipe (kept for comparison)
  static const int dbg_gram = getenv("RXB_DBG_GRAM_STATS") ? atoi(getenv("RXB_DBG_GRAM_STATS")) : 0;
  p.mma_stats = (dgrad || (dbg_gram && p.do_stats && p.bn == kMaxBN)) ? 1 : 0;
  static const int dbg_dgrad = getenv("RXB_DBG_DGRAD") ? atoi(getenv("RXB_DBG_DGRAD")) : 0;   // timing experiments only
  if (dgrad && (dbg_dgrad & 1)) { p.mma_stats = 0; p.do_stats = 0; }
  if (dgrad && (dbg_dgrad & 2) && p.out_mode == OUT_G_ACCUM) p.out_mode = OUT_G_WRITE;
  // TMEM accumulator stages: the narrow store epilogue (shuffle statistics, no Gram / sum columns in TMEM) has all 512
  // columns for accumulators - four stages let the MMA warp run ahead of the x-merge epilogue, which holds a stage for
  // ~4k cycles per tile (profiles/r02_timelines.log)
  static const int dbg_nacc = getenv("RXB_DBG_NACC") ? atoi(getenv("RXB_DBG_NACC")) : 4;
  p.n_acc = (!dgrad && p.bn < kMaxBN && !p.mma_stats && dbg_nacc == 4) ? 4 : 2;
  static const int dbg_shfl = getenv("RXB_DBG_SHFL_STATS") ? atoi(getenv("RXB_DBG_SHFL_STATS")) : 0;
  p.col_narrow = (!dgrad && !dbg_shfl && (p.bn == 32 || p.bn == 64)) ? 1 : 0;
  if (wg3) p.n_acc = 1;   // TMEM: 128 accumulator + 288 weight-gradient + 16 sum columns
  p.n_tiles = ceil_div(p.n_total, p.bn);
  p.kb_per_tap = ceil_div(p.cin, bk);
  if (prologue && p.kb_per_tap * bk > kMaxPrologueC + 64) return set_error(RXB_ERR_INVALID, "conv_gemm: cin too large");
  // tiling: multi-row filters prefer tall tiles so the row halo is cheap
  p.t = make_tiling(p.B, p.H, p.W);
  p.halo = 0;
  if (p.taps_y > 1) {
    PixelTiling tall = make_tiling_tall(p.B, p.H, p.W);
    if (tall.tb_log2 == 0 && tall.tw_log2 == 3) {
      p.t = tall;
      p.halo = 1;
      // 128-byte rows: the whole (tw+taps_x-1) x (th+taps_y-1) neighbourhood is ONE box and every tap is a descriptor
      // offset into it (the weight panel must then be resident: one A stage serves all taps of a k-block)
      static const int dbg_no_full_halo = getenv("RXB_DBG_NO_FULL_HALO") ? atoi(getenv("RXB_DBG_NO_FULL_HALO")) : 0;
      // (64-byte rows, BK = 32: the same holds for the 64 B swizzle - probed too; dbg value 2 keeps them on row halo)
      if ((bk == 64 || dbg_no_full_halo != 2) && dbg_no_full_halo != 1) p.halo = 2;
    }
  }
  const int taps = p.taps_x * p.taps_y;
  if (p.halo == 2) {
    const long long panel = (long long)taps * p.kb_per_tap * p.bn * bk * 2;
    if (panel > 96 * 1024) p.halo = 1;
  }
  // x-merged tiles for narrow outputs: a tcgen05.mma costs the same ~70 cycles for every N <= 128, so the taps_x taps
  // of a filter row become ONE N = taps_x*Cout MMA over a 16-wide M tile that carries its own x halo (14 valid
  // columns); the epilogue adds the neighbouring columns' partial sums with warp shuffles.  3x fewer MMAs for 3x3/32.
  static const int dbg_no_xmerge = getenv("RXB_DBG_NO_XMERGE") ? atoi(getenv("RXB_DBG_NO_XMERGE")) : 0;
  if (!dbg_no_xmerge && p.halo == 2 && !dgrad && bk == 64 && p.taps_x == 3 && p.pad_x == 1 && p.n_tiles == 1 &&
      p.bn == 32 && p.H >= 8 && p.W >= 56) {   // narrower images waste too many of the 14-wide tiles' columns
    p.halo = 3;
    p.t.tw_log2 = 4; p.t.th_log2 = 3; p.t.tb_log2 = 0;
    p.t.x_step = 16 - (p.taps_x - 1);
    p.t.tiles_x = ceil_div(p.W, p.t.x_step);
    p.t.tiles_y = ceil_div(p.H, 8);
    p.t.tiles_b = p.B;
  }
  const int tw = 1 << p.t.tw_log2, th = 1 << p.t.th_log2;
  const int box_h = p.halo ? th + p.taps_y - 1 : th;
  const int box_w = p.halo == 2 ? tw + p.taps_x - 1 : tw;
  p.rows_a = p.halo ? box_h * box_w : 128;
  if (box_h > 256) return set_error(RXB_ERR_INVALID, "conv_gemm: halo box too tall");

  CUtensorMap tmA, tmB, tmOut, tmX, tmA2;
  const bool fix = p.fix.G != nullptr;
  if (fix && !(wg3 && p.halo == 2 && p.fix.X && p.fix.mean && p.fix.rstd && p.fix.corrA && p.fix.corrB &&
               (p.fix.ld * 2) % 16 == 0 && p.fix.c0 % 8 == 0 && p.fix.c0 + p.cin <= p.fix.ld))
    return set_error(RXB_ERR_INVALID, "conv_gemm: FixupArgs need the fused 3x3 kernel (EPI 4) and aligned concat slices");
  int rc;
  if (fix) {   // the A operand is derived from the two concat slices
    rc = make_act_tmap(&tmA, static_cast<const __nv_bfloat16*>(p.fix.G) + p.fix.c0, p.t, p.cin, p.fix.ld, bk, box_h, box_w);
    if (rc) return rc;
    rc = make_act_tmap(&tmA2, static_cast<const __nv_bfloat16*>(p.fix.X) + p.fix.c0, p.t, p.cin, p.fix.ld, bk, box_h, box_w);
    if (rc) return rc;
  } else {
    rc = make_act_tmap(&tmA, A, p.t, p.cin, ldA, bk, box_h, box_w);
    if (rc) return rc;
    tmA2 = tmA;
  }
  {
    uint64_t dims[3] = {(uint64_t)p.cin, (uint64_t)p.n_total, (uint64_t)taps};
    uint64_t strides[2] = {(uint64_t)p.cin * 2, (uint64_t)p.cin * 2 * p.n_total};
    uint32_t box[3] = {(uint32_t)bk, (uint32_t)p.bn, 1};
    rc = make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(Wt), dims, strides, box,
                   bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  const int cw = p.bn >= 64 ? 64 : 32;
  rc = make_act_tmap(&tmOut, static_cast<__nv_bfloat16*>(out) + c_off, p.t, p.n_total, ldc, cw, th,
                     p.halo == 3 ? p.t.x_step : 0);
  if (rc) return rc;
  if (dgrad) {
    rc = make_act_tmap(&tmX, X, p.t, p.n_total, ldx, cw, th);
    if (rc) return rc;
  } else {
    tmX = tmOut;
  }

  // ---- shared-memory plan: [A stages][B: resident panel or per-stage][staging x n_stg][x tile x 2 (dgrad)][ones][aux]
  const int row_bytes = bk * 2;
  const long long a_stage = (fix ? 2 : 1) * (long long)((p.rows_a * row_bytes + 1023) & ~1023);
  const long long b_tap = (long long)p.bn * row_bytes;
  const long long b_stage = (p.halo == 1 ? p.taps_y : 1) * b_tap;
  const long long b_panel = (long long)taps * p.kb_per_tap * b_tap;
  const long long stage_tile = 128ll * ceil_div(p.bn, cw) * cw * 2;
  // dgrad: 2-3 activation-tile buffers that double as the output staging
  const long long fixed = (long long)sizeof(GemmAux) + 1024 /*alignment*/ + 1024 /*ones*/ + (dgrad ? stage_tile : 0);
  const long long budget = 227 * 1024;
  const int m_tiles = p.t.tiles_x * p.t.tiles_y * p.t.tiles_b;
  p.t.rev = p.reverse ? m_tiles : 0;
  int gx = num_sms() / p.n_tiles;
  if (gx < 1) gx = 1;
  if (gx > m_tiles) gx = m_tiles;
  if (gx <= 0) return RXB_OK;
  p.b_resident = 0;
  p.n_stg = 1;
  // stores: two staging buffers (one per epilogue group); dgrad: two activation/staging buffers, a third if it fits
  long long per_stage = a_stage + b_stage, avail = budget - fixed - stage_tile - (dgrad ? 0 : stage_tile);
  static const int dbg_no_resident = getenv("RXB_DBG_NO_RESIDENT") ? atoi(getenv("RXB_DBG_NO_RESIDENT")) : 0;
  const bool allow_res = !(dbg_no_resident == 1 || (dbg_no_resident == 2 && p.bn < 128));
  if ((allow_res || p.halo >= 2) && b_panel <= 96 * 1024 && (m_tiles > gx || p.halo >= 2) &&
      (avail - b_panel) / a_stage >= (p.halo >= 2 ? 2 : 3)) {
    p.b_resident = 1;
    per_stage = a_stage;
    avail -= b_panel;
  }
  if (dgrad) {
    // the load of an activation tile is on the epilogue's dependency chain (buffer freed -> TMA load -> epilogue),
    // so a third buffer hides one load latency
    // An EVEN number of buffers, so that a buffer always belongs to the same epilogue group (the groups take alternate
    // tiles): with three buffers shared by both groups the kernel faulted intermittently in 4-GPU runs (never at N<=2);
    // two or four buffers ran clean.  RXB_DBG_NX=3 restores the odd count for investigation.
    static const int dbg_nx = getenv("RXB_DBG_NX") ? atoi(getenv("RXB_DBG_NX")) : 3;   // 3: odd count, per-group barriers
    p.n_stg = 2;
    if (dbg_nx == 3) {
      // (FixupArgs: a stage is two boxes; two stages = two tiles of prefetch are enough for the small dZ loads)
      if ((avail - stage_tile) / per_stage >= (fix ? 2 : 3)) { p.n_stg = 3; avail -= stage_tile; }
    } else if (dbg_nx >= 4 && (avail - 2 * stage_tile) / per_stage >= 3) {
      p.n_stg = 4;
      avail -= 2 * stage_tile;
    }
  } else {
    // one staging buffer per epilogue group (the groups alternate tiles; a shared buffer would make their parity
    // waits ambiguous)
    p.n_stg = 2;
    // Forward 1x1 with more than 128 input channels: ONE 32 KB staging buffer shared by the two groups - two more
    // operand stages in flight for a kernel bound by bytes in flight (per-group release barriers keep the waits
    // unambiguous).  Measured per shape: Cin 224 0.339 -> 0.327 ms, 256 (64x64) -3 %, 992 -5 %; with 64 input
    // channels the tile is too short and the groups wait for each other (0.193 -> 0.266 ms), hence the threshold.
    // RXB_DBG_STG1=0 disables, =2 forces it for every Cin.
    static const int dbg_stg1 = getenv("RXB_DBG_STG1") ? atoi(getenv("RXB_DBG_STG1")) : 1;
    if (dbg_stg1 && (p.kb_per_tap >= 3 || dbg_stg1 == 2) && prologue && p.bn == kMaxBN && bk == 64 && p.taps_x == 1 &&
        p.taps_y == 1 && !p.mma_stats) {
      p.n_stg = 1;
      avail += stage_tile;
    }
  }
  long long stages = avail / per_stage;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error(RXB_ERR_INVALID, "conv_gemm: tile too large for shared memory");
  if (p.halo >= 2 && !p.b_resident) return set_error(RXB_ERR_INVALID, "conv_gemm: full-halo tile without resident weights");
  p.stages = (int)stages;
  // (the result is staged in the A stages AND the weight area behind them: both are dead once every MMA has completed)
  if (wg3 && (p.halo != 2 || stages * a_stage + b_panel < 16ll * 128 * 9 * 4))
    return set_error(RXB_ERR_INVALID, "conv_gemm: fused 3x3 weight gradient: needs the full-halo tile and 72 KB of staging");
  if (p.wg_dW != nullptr && !wg3 &&
      stages * a_stage + (p.b_resident ? b_panel : stages * b_stage) < (long long)p.kb_per_tap * bk * 512)
    return set_error(RXB_ERR_INVALID, "conv_gemm: fused weight gradient: pipeline too small to stage the result");
  const size_t smem = (size_t)(stages * per_stage + (p.b_resident ? b_panel : 0) +
                               (dgrad ? p.n_stg - 1 : p.n_stg) * stage_tile + fixed);
  dim3 grid(gx, p.n_tiles);
  static const int dbg_tl = getenv("RXB_DBG_TIMELINE") ? atoi(getenv("RXB_DBG_TIMELINE")) : 0;
  static unsigned long long* tl_dev = nullptr;
  p.dbg = nullptr;
  if (dbg_tl) {
    if (!tl_dev) cudaMalloc(&tl_dev, 3 * 16 * 8 * 8);
    cudaMemsetAsync(tl_dev, 0, 3 * 16 * 8 * 8, stream);
    p.dbg = tl_dev;
  }

  RXB_PROF(stream, dgrad ? (taps > 1 ? PROF_CONV_DGRAD_3X3 : PROF_CONV_DGRAD)
                         : !prologue ? PROF_CONV_OTHER : (taps > 1 ? PROF_CONV_FWD_3X3 : PROF_CONV_FWD));
#define RXB_LAUNCH_GEMM(BK_, PRO_, EPI_)                                                                        \
  do {                                                                                                          \
    RXB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BK_, PRO_, EPI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem));                                                                  \
    RXB_CUDA(launch_k((conv_gemm_kernel<BK_, PRO_, EPI_>), grid, dim3(kConvThreads), smem, stream, tmA, tmB, tmOut, \
                      tmX, tmA2, p));                                                                           \
  } while (0)
  const int epi = dgrad ? (wg3 ? 4 : p.wg_dW != nullptr ? 3 : 2) : (p.bn < kMaxBN ? 1 : 0);
  if (bk == 64 && prologue) { if (epi == 1) RXB_LAUNCH_GEMM(64, true, 1); else RXB_LAUNCH_GEMM(64, true, 0); }
  else if (bk == 64) { if (epi == 3) RXB_LAUNCH_GEMM(64, false, 3); else if (epi == 2) RXB_LAUNCH_GEMM(64, false, 2); else if (epi == 1) RXB_LAUNCH_GEMM(64, false, 1); else RXB_LAUNCH_GEMM(64, false, 0); }
  else { if (epi == 4) RXB_LAUNCH_GEMM(32, false, 4); else if (epi == 2) RXB_LAUNCH_GEMM(32, false, 2); else if (epi == 1) RXB_LAUNCH_GEMM(32, false, 1); else RXB_LAUNCH_GEMM(32, false, 0); }
#undef RXB_LAUNCH_GEMM
  RXB_LAUNCH_OK();
  if (dbg_tl) {   // development: print the timeline of CTA (0,0), cycles relative to its first event
    unsigned long long h[3 * 16 * 8];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, tl_dev, sizeof(h), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (int i = 0; i < 3 * 16 * 8; ++i) if (h[i] && h[i] < t0) t0 = h[i];
    static int printed = 0;
    if (printed++ < dbg_tl) {
      printf("timeline epi=%d n=%d cin=%d taps=%d stages=%d n_stg=%d res=%d halo=%d\n", p.epi_mode, p.n_total, p.cin, taps,
             p.stages, p.n_stg, p.b_resident, p.halo);
      const char* role[3] = {"prod", "mma ", "epi "};
      for (int it = 0; it < 12; ++it)
        for (int r = 0; r < 3; ++r) {
          printf("  it%02d %s:", it, role[r]);
          for (int ev = 0; ev < 8; ++ev) {
            const unsigned long long v = h[(r * 16 + it) * 8 + ev];
            if (v) printf(" %7llu", v - t0); else printf("       -");
          }
          printf("\n");
        }
    }
  }
  return RXB_OK;
}

int launch_conv_wgrad(WgradParams p, const void* A, long long ldA, const void* dOut, long long ldD,
                      cudaStream_t stream) {
  if (!(p.bkc == 64 || p.bkc == 32)) return set_error(RXB_ERR_INVALID, "conv_wgrad: bkc must be 32 or 64");
  if (p.prologue && p.bkc != 64) return set_error(RXB_ERR_INVALID, "conv_wgrad: prologue needs bkc=64");
  if (!(p.n == 32 || (p.n % 64 == 0 && p.n >= 64 && p.n <= 256)))
    return set_error(RXB_ERR_INVALID, "conv_wgrad: n=%d must be 32 or a multiple of 64 up to 256", p.n);
  const int taps = p.taps_x * p.taps_y;
  p.boxes_per_tap = ceil_div(p.cin, p.bkc);
  p.boxes_per_chunk = 128 / p.bkc;
  p.shift_dout = (taps > 1 && p.bkc == 64 && p.cin <= 128 && taps * p.n <= 512 && p.w_mode != 3) ? 1 : 0;
  if (p.w_mode == 3 && (taps == 1 || p.bkc != 64 || p.cin % 4))
    return set_error(RXB_ERR_INVALID, "conv_wgrad: the tap-major scratch layout is for multi-tap filters with cin >= 64");
  if (p.shift_dout && p.n <= 64 && p.taps_x * p.n <= 256 && p.taps_x > 1) {
    // full-halo dOut box: needs 8-pixel tile rows (one K group per row); the tall tiling also keeps the halo small
    static const int dbg_no_wg_halo = getenv("RXB_DBG_NO_WG_HALO") ? atoi(getenv("RXB_DBG_NO_WG_HALO")) : 0;
    PixelTiling tall = make_tiling_tall(p.t.B, p.t.H, p.t.W);
    if (!dbg_no_wg_halo && tall.tw_log2 == 3 && tall.tb_log2 == 0) {
      p.t = tall;
      p.shift_dout = 2;
    }
  }
  int m_tiles = p.t.tiles_x * p.t.tiles_y * p.t.tiles_b;
  int chunk_groups;
  if (p.shift_dout) {
    p.n_chunks = 1;
    p.chunks_per_cta = 1;
    chunk_groups = 1;
  } else {
    p.n_chunks = ceil_div(taps * p.boxes_per_tap, p.boxes_per_chunk);
    const int max_cpc = 512 / p.n < p.n_chunks ? 512 / p.n : p.n_chunks;
    // Every CTA adds its partial dW into global memory, so the reduce traffic is (pixel-split CTAs) x |dW|, while
    // every channel group re-reads dOut.  Pick the chunks-per-CTA that minimises the modelled traffic.
    const double dw_bytes = (double)taps * p.boxes_per_tap * p.bkc * p.n * 4.0;
    const double dout_bytes = (double)m_tiles * 128.0 * p.n * 2.0;
    double best = 0;
    p.chunks_per_cta = max_cpc;
    for (int cpc = 1; cpc <= max_cpc; ++cpc) {
      const int cg = ceil_div(p.n_chunks, cpc);
      int pc = num_sms() / cg;
      if (pc < 1) pc = 1;
      if (pc > m_tiles) pc = m_tiles;
      const double cost = 2.0 * pc * dw_bytes + (double)cg * dout_bytes;
      if (cpc == 1 || cost < best) { best = cost; p.chunks_per_cta = cpc; }
    }
    chunk_groups = ceil_div(p.n_chunks, p.chunks_per_cta);
  }
  // full-halo A box (the stem's 4x4 taps over 32 channels): needs a filter row per 128-row chunk, no A prologue,
  // 8-pixel tile rows and every chunk in one CTA
  p.a_halo = 0;
  {
    static const int dbg_no_a_halo = getenv("RXB_DBG_NO_A_HALO") ? atoi(getenv("RXB_DBG_NO_A_HALO")) : 0;
    PixelTiling tall = make_tiling_tall(p.t.B, p.t.H, p.t.W);
    if (!dbg_no_a_halo && !p.shift_dout && taps > 1 && !p.prologue && p.bkc * p.taps_x == 128 && p.boxes_per_tap == 1 &&
        p.n_chunks == p.taps_y && p.taps_y * p.n <= 512 && tall.tw_log2 == 3 && tall.tb_log2 == 0 &&
        (8 + p.taps_x - 1) * ((1 << tall.th_log2) + p.taps_y - 1) * p.bkc * 2 <= kWgA_BYTES) {
      p.a_halo = 1;
      p.t = tall;
      p.chunks_per_cta = p.n_chunks;
      chunk_groups = 1;
    }
  }
  if (p.prologue && p.boxes_per_tap * p.bkc > kMaxPrologueC + 64)
    return set_error(RXB_ERR_INVALID, "conv_wgrad: cin too large");
  m_tiles = p.t.tiles_x * p.t.tiles_y * p.t.tiles_b;
  int pix_ctas = num_sms() / chunk_groups;
  static const int dbg_wg_pix = getenv("RXB_DBG_WG_PIX") ? atoi(getenv("RXB_DBG_WG_PIX")) : 0;   // experiment: cap
  if (dbg_wg_pix > 0 && pix_ctas > dbg_wg_pix) pix_ctas = dbg_wg_pix;
  if (pix_ctas < 1) pix_ctas = 1;
  if (pix_ctas > m_tiles) pix_ctas = m_tiles;
  p.pix_tiles_per_cta = ceil_div(m_tiles, pix_ctas);
  pix_ctas = ceil_div(m_tiles, p.pix_tiles_per_cta);

  CUtensorMap tmA, tmD;
  const int th = 1 << p.t.th_log2;
  const int tw = 1 << p.t.tw_log2;
  int rc = p.a_halo ? make_act_tmap(&tmA, A, p.t, p.cin, ldA, p.bkc, th + p.taps_y - 1, tw + p.taps_x - 1)
                    : make_act_tmap(&tmA, A, p.t, p.cin, ldA, p.bkc, th);
  if (rc) return rc;
  if (p.shift_dout == 2)
    rc = make_act_tmap(&tmD, static_cast<const __nv_bfloat16*>(dOut) + p.n_off, p.t, p.n, ldD, p.n >= 64 ? 64 : 32,
                       th + p.taps_y - 1, tw + p.taps_x - 1);
  else
    rc = make_act_tmap(&tmD, static_cast<const __nv_bfloat16*>(dOut) + p.n_off, p.t, p.n, ldD, p.n >= 64 ? 64 : 32, th);
  if (rc) return rc;

  const int b_bytes = p.shift_dout == 2 ? (((tw + p.taps_x - 1) * (th + p.taps_y - 1) * p.n * 2 + 1023) & ~1023)
                                        : (p.shift_dout ? taps : 1) * 128 * p.n * 2;
  const size_t budget = 226 * 1024;
  // classic mode: two dOut tile buffers shared by the channel chunks + A stages; shifted modes: dOut rides in the stage
  const size_t d_fixed = p.shift_dout ? 0 : 2 * (size_t)b_bytes;
  const size_t per_stage = kWgA_BYTES + (p.shift_dout ? b_bytes : 0);
  int stages = (int)((budget - sizeof(WgradAux) - 1024 - d_fixed) / per_stage);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error(RXB_ERR_INVALID, "conv_wgrad: tile too large for shared memory");
  const size_t smem = (size_t)stages * per_stage + d_fixed + sizeof(WgradAux) + 1024;
  // result leaves through shared memory as bulk reduce-adds when its global layout is contiguous per output
  // channel: 1x1 filters, or the shifted-dOut multi-tap mode; the staging area is the dead pipeline
  const size_t pipe_bytes = (size_t)stages * per_stage + d_fixed;
  p.bulk_out = 0;
  p.bulk_bufs = 1;
  if (p.w_mode == 3 && !p.a_halo) {
    if ((size_t)p.n * 512 > pipe_bytes) return set_error(RXB_ERR_INVALID, "conv_wgrad: staging does not fit");
    p.bulk_out = 1;
    p.bulk_bufs = (size_t)p.n * 1024 <= pipe_bytes ? 2 : 1;
  }
  if (p.w_mode == 0 && p.cin % 8 == 0) {
    if (p.shift_dout && (size_t)p.n * p.cin * taps * 4 <= pipe_bytes) p.bulk_out = 1;
    if (!p.shift_dout && taps == 1 && p.bkc == 64 && (size_t)p.n * 512 <= pipe_bytes) {
      p.bulk_out = 1;
      p.bulk_bufs = (size_t)p.n * 1024 <= pipe_bytes ? 2 : 1;
    }
  }
  RXB_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RXB_PROF(stream, !p.prologue ? PROF_WGRAD_OTHER : (taps > 1 ? PROF_CONV_WGRAD_3X3 : PROF_CONV_WGRAD));
  static const int dbg_tl = getenv("RXB_DBG_TIMELINE") ? atoi(getenv("RXB_DBG_TIMELINE")) : 0;
  static unsigned long long* tl_dev = nullptr;
  p.dbg = nullptr;
  if (dbg_tl) {
    if (!tl_dev) cudaMalloc(&tl_dev, 3 * 16 * 8 * 8);
    cudaMemsetAsync(tl_dev, 0, 3 * 16 * 8 * 8, stream);
    p.dbg = tl_dev;
  }
  RXB_CUDA(launch_k(conv_wgrad_kernel, dim3(pix_ctas, chunk_groups), dim3(kGemmThreads), smem, stream, tmA, tmD, p, stages));
  RXB_LAUNCH_OK();
  if (dbg_tl) {   // development: timeline of CTA (0,0); rows: producer / MMA per pixel tile, epilogue in row it00
    unsigned long long h[3 * 16 * 8];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, tl_dev, sizeof(h), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (int i = 0; i < 3 * 16 * 8; ++i) if (h[i] && h[i] < t0) t0 = h[i];
    static int printed = 0;
    if (printed++ < dbg_tl) {
      printf("wgrad timeline cin=%d n=%d taps=%d stages=%d shift=%d a_halo=%d chunks/cta=%d tiles/cta=%d grid=(%d,%d)\n", p.cin,
             p.n, taps, stages, p.shift_dout, p.a_halo, p.chunks_per_cta, p.pix_tiles_per_cta, pix_ctas, chunk_groups);
      const char* role[3] = {"prod", "mma ", "epi "};
      for (int it = 0; it < 14; ++it)
        for (int r = 0; r < 3; ++r) {
          if (r == 2 && it > 0) continue;
          printf("  it%02d %s:", it, role[r]);
          for (int ev = 0; ev < 8; ++ev) {
            const unsigned long long v = h[(r * 16 + it) * 8 + ev];
            if (v) printf(" %7llu", v - t0); else printf("       -");
          }
          printf("\n");
        }
    }
  }
  return RXB_OK;
}

}  // namespace rxb
