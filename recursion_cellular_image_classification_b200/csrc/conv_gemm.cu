// conv_gemm.cu — kernel family 2: convolutions as tcgen05 implicit GEMMs.
//
// Forward / data-gradient kernel (conv_gemm_kernel):
//     D[p, n] = sum_{tap, k} A'(p shifted by tap)[k] * Wt[tap][n][k]          bf16 x bf16 -> fp32 (TMEM)
//   * a persistent, warp-specialised CTA per SM: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer and
//     TMEM owner, warps 2-5 = epilogue (TMEM -> registers -> global), warps 6-9 = A-operand transform.
//   * the M tile is 128 pixels fetched by one 4-D TMA box; taps are box shifts, padding is TMA zero fill.
//   * DenseNet's pre-activation BatchNorm+ReLU (torchvision _DenseLayer: norm -> relu -> conv) is applied
//     to the A tile IN SHARED MEMORY between the TMA and the MMA (zero padding stays zero), so the
//     normalised activation never exists in HBM.
//   * two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
//   * epilogues: bf16 store at a channel offset of a wider tensor (DenseNet concat-by-offset) with
//     per-channel sum / sum-of-squares for the NEXT BatchNorm, or the fused ReLU/BatchNorm backward
//     (mask, per-channel reductions, accumulate into the concat gradient).
//
// Weight-gradient kernel (conv_wgrad_kernel):
//     dW[tap][n][k] += sum_p A'(p shifted by tap)[k] * dOut[p][n]
//   both operands are MN-major (the contraction runs over pixels), 128 A-channels per accumulator group,
//   up to 512 TMEM columns of groups per CTA, pixel range split over CTAs, fp32 atomics into torch OIHW.
//
// Reference: torchvision densenet121 as swapped in for TwoSitesNN's trunk (reference
// cell_classifier/models.py:16-29, 45); the convs there are cuDNN calls.
#include "conv_gemm.cuh"

namespace rxb {

constexpr int kGemmThreads = 320;
constexpr int kMaxStages = 8;
constexpr int kAccStride = 256;   // TMEM columns per accumulator stage
constexpr int kMaxPrologueC = 1024;
constexpr int kMaxStatN = 1024;

struct __align__(16) GemmAux {
  float s_scale[kMaxPrologueC + 64];
  float s_shift[kMaxPrologueC + 64];
  float s_stat[2][kMaxStatN];
  uint64_t full[kMaxStages];
  uint64_t xform[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

PixelTiling make_tiling(int B, int H, int W) {
  PixelTiling t;
  t.W = W; t.H = H; t.B = B;
  auto log2_ceil = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
  int twl = log2_ceil(W); if (twl > 7) twl = 7;
  int thl = log2_ceil(H); if (thl > 7 - twl) thl = 7 - twl;
  int tbl = 7 - twl - thl;
  t.tw_log2 = twl; t.th_log2 = thl; t.tb_log2 = tbl;
  t.tiles_x = ceil_div(W, 1 << twl);
  t.tiles_y = ceil_div(H, 1 << thl);
  t.tiles_b = ceil_div(B, 1 << tbl);
  return t;
}

__device__ __forceinline__ void tile_origin(const PixelTiling& t, int m_tile, int& x0, int& y0, int& b0) {
  const int tx = m_tile % t.tiles_x;
  const int rest = m_tile / t.tiles_x;
  const int ty = rest % t.tiles_y;
  const int tb = rest / t.tiles_y;
  x0 = tx << t.tw_log2;
  y0 = ty << t.th_log2;
  b0 = tb << t.tb_log2;
}
__device__ __forceinline__ void row_coord(const PixelTiling& t, int row, int& xi, int& yi, int& bi) {
  xi = row & ((1 << t.tw_log2) - 1);
  yi = (row >> t.tw_log2) & ((1 << t.th_log2) - 1);
  bi = row >> (t.tw_log2 + t.th_log2);
}

// In-place relu(x*scale+shift) on one [128 rows][64 ch] bf16 tile stored with the 128-byte swizzle.
// 128 threads: thread t owns 16-byte chunk j = t&7 (channels 8j..8j+7) of rows (t>>3) + 16 i.
// Rows whose (shifted) pixel lies outside the image keep the zeros TMA wrote (conv zero padding).
__device__ __forceinline__ void transform_tile_sw128(uint8_t* tile, const float* sc, const float* sh, int t,
                                                     const PixelTiling& til, int x0, int y0, int b0, int dx,
                                                     int dy) {
  const int j = t & 7;
  float s[8], h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    s[e] = sc[j * 8 + e];
    h[e] = sh[j * 8 + e];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = (t >> 3) + 16 * i;
    int xi, yi, bi;
    row_coord(til, row, xi, yi, bi);
    const int x = x0 + xi + dx, y = y0 + yi + dy, b = b0 + bi;
    if (x < 0 || x >= til.W || y < 0 || y >= til.H || b >= til.B) continue;
    uint4* p = reinterpret_cast<uint4*>(tile + row * 128 + ((j ^ (row & 7)) << 4));
    uint4 v = *p;
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float lo = fmaxf(fmaf(bf16_lo(w[e]), s[2 * e], h[2 * e]), 0.f);
      float hi = fmaxf(fmaf(bf16_hi(w[e]), s[2 * e + 1], h[2 * e + 1]), 0.f);
      w[e] = pack_bf16x2(lo, hi);
    }
    *p = make_uint4(w[0], w[1], w[2], w[3]);
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

template <int BK, bool PROLOGUE>
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p, const int stages) {
  static_assert(BK == 64 || BK == 32, "BK");
  static_assert(!PROLOGUE || BK == 64, "the in-smem BatchNorm+ReLU transform is written for 128B rows");
  constexpr int A_BYTES = 128 * BK * 2;
  constexpr uint32_t kSwz = BK == 64 ? ptx::kSwizzle128B : ptx::kSwizzle64B;
  constexpr uint32_t kSBO = 8 * BK * 2;  // 8 rows of BK bf16

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.bn * BK * 2;
  uint8_t* smA = smem;
  uint8_t* smB = smem + (size_t)stages * A_BYTES;
  GemmAux* aux = reinterpret_cast<GemmAux*>(smB + (size_t)stages * b_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = p.t.tiles_x * p.t.tiles_y * p.t.tiles_b;
  const int total_tiles = m_tiles * p.n_tiles;
  const int taps = p.taps_x * p.taps_y;
  const int kb_total = taps * p.kb_per_tap;

  // ---- one-time setup
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&aux->full[s], 1);
      ptx::mbar_init(&aux->xform[s], 128);
      ptx::mbar_init(&aux->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&aux->tmem_full[a], 1);
      ptx::mbar_init(&aux->tmem_empty[a], 128);
    }
    ptx::fence_barrier_init();
  }
  if (PROLOGUE) {
    const int padded = p.kb_per_tap * BK;
    for (int c = threadIdx.x; c < padded; c += kGemmThreads) {
      aux->s_scale[c] = c < p.cin ? p.scale[c] : 0.f;
      aux->s_shift[c] = c < p.cin ? p.shift[c] : 0.f;
    }
  }
  if (p.do_stats)
    for (int c = threadIdx.x; c < 2 * kMaxStatN; c += kGemmThreads) (&aux->s_stat[0][0])[c] = 0.f;
  if (warp == 1) ptx::tmem_alloc<512>(&aux->tmem_base);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = aux->tmem_base;

  if (warp == 0) {
    // =============================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
        int x0, y0, b0;
        tile_origin(p.t, m_tile, x0, y0, b0);
        const int n0 = n_tile * p.bn;
        for (int tp = 0; tp < taps; ++tp) {
          const int ty = tp / p.taps_x, tx = tp - ty * p.taps_x;
          for (int kb = 0; kb < p.kb_per_tap; ++kb) {
            ptx::mbar_wait(&aux->empty[stage], phase ^ 1, 1);
            ptx::mbar_arrive_expect_tx(&aux->full[stage], A_BYTES + b_bytes);
            ptx::tma_load_4d(smA + (size_t)stage * A_BYTES, &tmA, &aux->full[stage], kb * BK, x0 + tx - p.pad_x,
                             y0 + ty - p.pad_y, b0);
            ptx::tma_load_3d(smB + (size_t)stage * b_bytes, &tmB, &aux->full[stage], kb * BK, n0, tp);
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, p.bn, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&aux->tmem_empty[acc], acc_phase ^ 1, 2);
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < kb_total; ++kb) {
          ptx::mbar_wait(PROLOGUE ? &aux->xform[stage] : &aux->full[stage], phase, 3);
          ptx::tcgen05_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smA + (size_t)stage * A_BYTES);
          const uint32_t b_addr = ptx::smem_u32(smB + (size_t)stage * b_bytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = ptx::make_smem_desc(a_addr + k * 32, 16, kSBO, kSwz);
            const uint64_t db = ptx::make_smem_desc(b_addr + k * 32, 16, kSBO, kSwz);
            ptx::umma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&aux->empty[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&aux->tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp < 6) {
    // =============================== epilogue: TMEM lanes (warp & 3) * 32 ..
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
      int x0, y0, b0, xi, yi, bi;
      tile_origin(p.t, m_tile, x0, y0, b0);
      row_coord(p.t, row, xi, yi, bi);
      const int x = x0 + xi, y = y0 + yi, b = b0 + bi;
      const bool valid = x < p.t.W && y < p.t.H && b < p.t.B;
      const long long pix = ((long long)b * p.t.H + y) * p.t.W + x;
      const int n0 = n_tile * p.bn;

      ptx::mbar_wait(&aux->tmem_full[acc], acc_phase, 4);
      ptx::tcgen05_fence_after();
      for (int c = 0; c < p.bn; c += 32) {
        if (n0 + c >= p.n_total) break;
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccStride + c, r);
        ptx::tmem_ld_wait();
        float v[32];
        if (p.epi_mode == EPI_STORE) {
          uint32_t packed[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            packed[i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            v[2 * i] = valid ? bf16_lo(packed[i]) : 0.f;
            v[2 * i + 1] = valid ? bf16_hi(packed[i]) : 0.f;
          }
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + pix * p.ldc + p.c_off + n0 + c);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
          }
          if (p.do_stats) {
            float sq[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) sq[i] = v[i] * v[i];
            const float cs = warp_column_sums(v, lane);
            const float cq = warp_column_sums(sq, lane);
            atomicAdd(&aux->s_stat[0][n0 + c + lane], cs);
            atomicAdd(&aux->s_stat[1][n0 + c + lane], cq);
          }
        } else {  // EPI_DGRAD_BN
          float xh[32];
          uint32_t xin[16];
          if (valid) {
            const uint4* xs = reinterpret_cast<const uint4*>(p.X + pix * p.ldx + n0 + c);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 t4 = __ldg(xs + i);
              xin[4 * i] = t4.x; xin[4 * i + 1] = t4.y; xin[4 * i + 2] = t4.z; xin[4 * i + 3] = t4.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) xin[i] = 0u;
          }
          uint32_t gin[16];
          __nv_bfloat16* gp = p.out + pix * p.ldc + p.c_off + n0 + c;
          if (p.out_mode == OUT_G_ACCUM && valid) {
            const uint4* gs = reinterpret_cast<const uint4*>(gp);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 t4 = gs[i];
              gin[4 * i] = t4.x; gin[4 * i + 1] = t4.y; gin[4 * i + 2] = t4.z; gin[4 * i + 3] = t4.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) gin[i] = 0u;
          }
          uint32_t packed[16];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int ch = n0 + c + i;
            const float xv = (i & 1) ? bf16_hi(xin[i >> 1]) : bf16_lo(xin[i >> 1]);
            const float es = __ldg(p.e_scale + ch), et = __ldg(p.e_shift + ch);
            const float pre = fmaf(xv, es, et);
            float dy = (valid && pre > 0.f) ? __uint_as_float(r[i]) : 0.f;
            v[i] = dy;
            xh[i] = valid ? dy * ((xv - __ldg(p.e_mean + ch)) * __ldg(p.e_rstd + ch)) : 0.f;
            float o;
            if (p.out_mode == OUT_DY) {
              o = dy;
            } else {
              const float g0 = (i & 1) ? bf16_hi(gin[i >> 1]) : bf16_lo(gin[i >> 1]);
              o = fmaf(es, dy, g0);
            }
            const uint32_t ob = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(o));
            if (i & 1) packed[i >> 1] |= ob << 16; else packed[i >> 1] = ob;
          }
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(gp);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
          }
          if (p.do_stats) {
            const float cs = warp_column_sums(v, lane);
            const float cq = warp_column_sums(xh, lane);
            atomicAdd(&aux->s_stat[0][n0 + c + lane], cs);
            atomicAdd(&aux->s_stat[1][n0 + c + lane], cq);
          }
        }
      }
      ptx::tcgen05_fence_before();
      ptx::mbar_arrive(&aux->tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.do_stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int et = threadIdx.x - 64;  // 0..127
      const int stat_off = p.epi_mode == EPI_STORE ? p.c_off : 0;
      for (int c = et; c < p.n_total; c += 128) {
        const float a = aux->s_stat[0][c], bq = aux->s_stat[1][c];
        if (a != 0.f) atomicAdd(p.ch_sum + stat_off + c, a);
        if (bq != 0.f) atomicAdd(p.ch_sumsq + stat_off + c, bq);
      }
    }
  } else {
    // =============================== A-operand transform (pre-activation BatchNorm + ReLU)
    if (PROLOGUE) {
      const int t = threadIdx.x - 192;  // 0..127
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles;
        int x0, y0, b0;
        tile_origin(p.t, m_tile, x0, y0, b0);
        for (int tp = 0; tp < taps; ++tp) {
          const int ty = tp / p.taps_x, tx = tp - ty * p.taps_x;
          for (int kb = 0; kb < p.kb_per_tap; ++kb) {
            ptx::mbar_wait(&aux->full[stage], phase, 5);
            transform_tile_sw128(smA + (size_t)stage * A_BYTES, aux->s_scale + kb * BK, aux->s_shift + kb * BK, t,
                                 p.t, x0, y0, b0, tx - p.pad_x, ty - p.pad_y);
            ptx::fence_proxy_async_smem();
            ptx::mbar_arrive(&aux->xform[stage]);
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
  float s_scale[kMaxPrologueC + 64];
  float s_shift[kMaxPrologueC + 64];
  uint64_t full[kMaxStages];
  uint64_t xform[kMaxStages];
  uint64_t empty[kMaxStages];
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
  const int b_bytes = 128 * p.n * 2;
  uint8_t* smA = smem;
  uint8_t* smB = smem + (size_t)stages * kWgA_BYTES;
  WgradAux* aux = reinterpret_cast<WgradAux*>(smB + (size_t)stages * b_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = p.t.tiles_x * p.t.tiles_y * p.t.tiles_b;
  const int taps = p.taps_x * p.taps_y;
  const int total_boxes = taps * p.boxes_per_tap;
  const int chunk0 = blockIdx.y * p.chunks_per_cta;
  const int n_local = min(p.chunks_per_cta, p.n_chunks - chunk0);
  const int tile_begin = blockIdx.x * p.pix_tiles_per_cta;
  const int tile_end = min(m_tiles, tile_begin + p.pix_tiles_per_cta);
  const int a_row_bytes = p.bkc * 2;
  const int a_box_bytes = 128 * a_row_bytes;
  const int d_boxes = p.n >= 64 ? p.n / 64 : 1;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmD);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&aux->full[s], 1);
      ptx::mbar_init(&aux->xform[s], 128);
      ptx::mbar_init(&aux->empty[s], 1);
    }
    ptx::mbar_init(&aux->tmem_full, 1);
    ptx::fence_barrier_init();
  }
  if (p.prologue) {
    const int padded = p.boxes_per_tap * p.bkc;
    for (int c = threadIdx.x; c < padded; c += kGemmThreads) {
      aux->s_scale[c] = c < p.cin ? p.scale[c] : 0.f;
      aux->s_shift[c] = c < p.cin ? p.shift[c] : 0.f;
    }
  }
  if (warp == 1) ptx::tmem_alloc<512>(&aux->tmem_base);
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
        for (int cl = 0; cl < n_local; ++cl) {
          ptx::mbar_wait(&aux->empty[stage], phase ^ 1, 11);
          ptx::mbar_arrive_expect_tx(&aux->full[stage], kWgA_BYTES + b_bytes);
          uint8_t* a_dst = smA + (size_t)stage * kWgA_BYTES;
          for (int i = 0; i < p.boxes_per_chunk; ++i) {
            const int kk = (chunk0 + cl) * p.boxes_per_chunk + i;
            int tp = 0, c0 = p.boxes_per_tap * p.bkc;  // fully out of bounds -> zero box
            if (kk < total_boxes) {
              tp = kk / p.boxes_per_tap;
              c0 = (kk - tp * p.boxes_per_tap) * p.bkc;
            }
            const int ty = tp / p.taps_x, tx = tp - ty * p.taps_x;
            ptx::tma_load_4d(a_dst + (size_t)i * a_box_bytes, &tmA, &aux->full[stage], c0, x0 + tx - p.pad_x,
                             y0 + ty - p.pad_y, b0);
          }
          uint8_t* d_dst = smB + (size_t)stage * b_bytes;
          for (int j = 0; j < d_boxes; ++j)
            ptx::tma_load_4d(d_dst + (size_t)j * 16384, &tmD, &aux->full[stage], j * 64, x0, y0, b0);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, p.n, 1, 1);
      const uint32_t a_swz = p.bkc == 64 ? ptx::kSwizzle128B : ptx::kSwizzle64B;
      const uint32_t a_sbo = 8 * a_row_bytes;
      const uint32_t a_kstep = 16 * a_row_bytes;
      const uint32_t d_row_bytes = p.n >= 64 ? 128 : 64;
      const uint32_t d_swz = p.n >= 64 ? ptx::kSwizzle128B : ptx::kSwizzle64B;
      const uint32_t d_sbo = 8 * d_row_bytes;
      const uint32_t d_kstep = 16 * d_row_bytes;
      const uint32_t d_lbo = 128 * d_row_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        for (int cl = 0; cl < n_local; ++cl) {
          ptx::mbar_wait(p.prologue ? &aux->xform[stage] : &aux->full[stage], phase, 13);
          ptx::tcgen05_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smA + (size_t)stage * kWgA_BYTES);
          const uint32_t d_addr = ptx::smem_u32(smB + (size_t)stage * b_bytes);
          const uint32_t acc = tmem_base + cl * p.n;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t da = ptx::make_smem_desc(a_addr + ks * a_kstep, a_box_bytes, a_sbo, a_swz);
            const uint64_t db = ptx::make_smem_desc(d_addr + ks * d_kstep, d_lbo, d_sbo, d_swz);
            ptx::umma_bf16_ss(acc, da, db, idesc, (tile > tile_begin || ks > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&aux->empty[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
      ptx::umma_commit(&aux->tmem_full);
    }
  } else if (warp < 6) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (tile_end > tile_begin) {
      ptx::mbar_wait(&aux->tmem_full, 0, 14);
      ptx::tcgen05_fence_after();
      const int ib = row / p.bkc, ch_in = row - ib * p.bkc;
      for (int cl = 0; cl < n_local; ++cl) {
        const int kk = (chunk0 + cl) * p.boxes_per_chunk + ib;
        const int tp = kk / p.boxes_per_tap;
        const int ch = (kk - tp * p.boxes_per_tap) * p.bkc + ch_in;
        bool ok = kk < total_boxes && ch < p.cin;
        long long base = 0, nstride = 0;
        if (p.w_mode == 0) {
          base = (long long)ch * taps + tp;
          nstride = (long long)p.cin * taps;
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
  } else {
    if (p.prologue) {
      const int t = threadIdx.x - 192;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        int x0, y0, b0;
        tile_origin(p.t, tile, x0, y0, b0);
        for (int cl = 0; cl < n_local; ++cl) {
          ptx::mbar_wait(&aux->full[stage], phase, 15);
          for (int i = 0; i < p.boxes_per_chunk; ++i) {
            const int kk = (chunk0 + cl) * p.boxes_per_chunk + i;
            if (kk >= total_boxes) continue;
            const int tp = kk / p.boxes_per_tap;
            const int c0 = (kk - tp * p.boxes_per_tap) * p.bkc;
            const int ty = tp / p.taps_x, tx = tp - ty * p.taps_x;
            transform_tile_sw128(smA + (size_t)stage * kWgA_BYTES + (size_t)i * a_box_bytes, aux->s_scale + c0,
                                 aux->s_shift + c0, t, p.t, x0, y0, b0, tx - p.pad_x, ty - p.pad_y);
          }
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive(&aux->xform[stage]);
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
static int make_act_tmap(CUtensorMap* tm, const void* base, const PixelTiling& t, int channels, long long ld,
                         int box_c) {
  uint64_t dims[4] = {(uint64_t)channels, (uint64_t)t.W, (uint64_t)t.H, (uint64_t)t.B};
  uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * t.W, (uint64_t)ld * 2 * t.W * t.H};
  uint32_t box[4] = {(uint32_t)box_c, 1u << t.tw_log2, 1u << t.th_log2, 1u << t.tb_log2};
  CUtensorMapSwizzle swz = box_c * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                           : box_c * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                             : CU_TENSOR_MAP_SWIZZLE_32B;
  return make_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, swz);
}

int launch_conv_gemm(const GemmParams& p, const void* A, long long ldA, const void* Wt, int bk, bool prologue,
                     cudaStream_t stream) {
  if (!(bk == 64 || bk == 32)) return set_error(RXB_ERR_INVALID, "conv_gemm: bk must be 32 or 64");
  if (prologue && bk != 64) return set_error(RXB_ERR_INVALID, "conv_gemm: prologue needs bk=64");
  if (p.bn % 32 || p.bn < 32 || p.bn > 256) return set_error(RXB_ERR_INVALID, "conv_gemm: bn=%d", p.bn);
  if (p.n_total % 32) return set_error(RXB_ERR_INVALID, "conv_gemm: n_total=%d not a multiple of 32", p.n_total);
  if (p.n_total > kMaxStatN && p.do_stats) return set_error(RXB_ERR_INVALID, "conv_gemm: n_total too large for stats");
  if (prologue && p.kb_per_tap * bk > kMaxPrologueC + 64) return set_error(RXB_ERR_INVALID, "conv_gemm: cin too large");
  if ((ldA * 2) % 16 || (reinterpret_cast<uintptr_t>(A) & 15))
    return set_error(RXB_ERR_INVALID, "conv_gemm: A not 16-byte aligned / ldA not a multiple of 8");
  CUtensorMap tmA, tmB;
  int rc = make_act_tmap(&tmA, A, p.t, p.cin, ldA, bk);
  if (rc) return rc;
  {
    const int taps = p.taps_x * p.taps_y;
    uint64_t dims[3] = {(uint64_t)p.cin, (uint64_t)p.n_total, (uint64_t)taps};
    uint64_t strides[2] = {(uint64_t)p.cin * 2, (uint64_t)p.cin * 2 * p.n_total};
    uint32_t box[3] = {(uint32_t)bk, (uint32_t)p.bn, 1};
    if ((p.cin * 2) % 16) return set_error(RXB_ERR_INVALID, "conv_gemm: cin must be a multiple of 8");
    rc = make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(Wt), dims, strides, box,
                   bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  const int a_bytes = 128 * bk * 2, b_bytes = p.bn * bk * 2;
  const size_t budget = 225 * 1024;
  int stages = (int)((budget - sizeof(GemmAux) - 1024) / (size_t)(a_bytes + b_bytes));
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error(RXB_ERR_INVALID, "conv_gemm: tile too large for shared memory");
  const size_t smem = (size_t)stages * (a_bytes + b_bytes) + sizeof(GemmAux) + 1024;
  const int total_tiles = p.t.tiles_x * p.t.tiles_y * p.t.tiles_b * p.n_tiles;
  int grid = num_sms();
  if (grid > total_tiles) grid = total_tiles;
  if (grid <= 0) return RXB_OK;

  RXB_PROF(stream, p.epi_mode == EPI_STORE ? PROF_CONV_FWD : PROF_CONV_DGRAD);
#define RXB_LAUNCH_GEMM(BK_, PRO_)                                                                             \
  do {                                                                                                         \
    RXB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BK_, PRO_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                  (int)smem));                                                                 \
    conv_gemm_kernel<BK_, PRO_><<<grid, kGemmThreads, smem, stream>>>(tmA, tmB, p, stages);                    \
  } while (0)
  if (bk == 64 && prologue) RXB_LAUNCH_GEMM(64, true);
  else if (bk == 64) RXB_LAUNCH_GEMM(64, false);
  else RXB_LAUNCH_GEMM(32, false);
#undef RXB_LAUNCH_GEMM
  RXB_LAUNCH_OK();
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
  p.n_chunks = ceil_div(taps * p.boxes_per_tap, p.boxes_per_chunk);
  p.chunks_per_cta = 512 / p.n;
  if (p.chunks_per_cta > p.n_chunks) p.chunks_per_cta = p.n_chunks;
  const int chunk_groups = ceil_div(p.n_chunks, p.chunks_per_cta);
  if (p.prologue && p.boxes_per_tap * p.bkc > kMaxPrologueC + 64)
    return set_error(RXB_ERR_INVALID, "conv_wgrad: cin too large");
  const int m_tiles = p.t.tiles_x * p.t.tiles_y * p.t.tiles_b;
  int pix_ctas = num_sms() / chunk_groups;
  if (pix_ctas < 1) pix_ctas = 1;
  if (pix_ctas > m_tiles) pix_ctas = m_tiles;
  p.pix_tiles_per_cta = ceil_div(m_tiles, pix_ctas);
  pix_ctas = ceil_div(m_tiles, p.pix_tiles_per_cta);

  CUtensorMap tmA, tmD;
  int rc = make_act_tmap(&tmA, A, p.t, p.cin, ldA, p.bkc);
  if (rc) return rc;
  rc = make_act_tmap(&tmD, static_cast<const __nv_bfloat16*>(dOut) + p.n_off, p.t, p.n, ldD, p.n >= 64 ? 64 : 32);
  if (rc) return rc;

  const int b_bytes = 128 * p.n * 2;
  const size_t budget = 225 * 1024;
  int stages = (int)((budget - sizeof(WgradAux) - 1024) / (size_t)(kWgA_BYTES + b_bytes));
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error(RXB_ERR_INVALID, "conv_wgrad: tile too large for shared memory");
  const size_t smem = (size_t)stages * (kWgA_BYTES + b_bytes) + sizeof(WgradAux) + 1024;
  RXB_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RXB_PROF(stream, PROF_CONV_WGRAD);
  conv_wgrad_kernel<<<dim3(pix_ctas, chunk_groups), kGemmThreads, smem, stream>>>(tmA, tmD, p, stages);
  RXB_LAUNCH_OK();
  return RXB_OK;
}

}  // namespace rxb
