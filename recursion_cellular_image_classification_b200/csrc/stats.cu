// stats.cu — kernel family 1a: per-experiment channel statistics (sum x, sum x^2, pixel count).
//
// Restates the hot loop of compute_mean_std (reference compute_stats_experiments.py:13-20) on decoded
// u8 planes.  The reference accumulates x/255 in f64; here the sums are EXACT integers (u32 per thread
// via IDP4A, u64 across threads), so the f64 finalisation reproduces the reference to ~1e-16 relative.
//
// HBM-bound: 1 byte read per pixel-channel, nothing written.  One CTA streams one (image, channel)
// plane with 16-byte no-allocate loads, 8 in flight per thread; a warp-shuffle + shared-memory tree
// folds the CTA and one u64 atomic per quantity lands in the per-(experiment, channel) slot.
#include "common.cuh"

namespace rxb {

constexpr int kStatsThreads = 256;
constexpr int kStatsUnroll = 8;

__device__ __forceinline__ void acc16(const uint4& v, unsigned& s, unsigned& q) {
  s = __dp4a(v.x, 0x01010101u, s);
  s = __dp4a(v.y, 0x01010101u, s);
  s = __dp4a(v.z, 0x01010101u, s);
  s = __dp4a(v.w, 0x01010101u, s);
  q = __dp4a(v.x, v.x, q);
  q = __dp4a(v.y, v.y, q);
  q = __dp4a(v.z, v.z, q);
  q = __dp4a(v.w, v.w, q);
}

// grid.x = n*C planes.  plane_vecs = H*W/16.
__global__ void __launch_bounds__(kStatsThreads)
stats_planar_kernel(const uint4* __restrict__ imgs, const int32_t* __restrict__ exp_id, int C,
                    int plane_vecs, int n_exp, unsigned long long* __restrict__ sum,
                    unsigned long long* __restrict__ sumsq, unsigned long long* __restrict__ count) {
  const long long plane = blockIdx.x;
  const int img = (int)(plane / C);
  const int ch = (int)(plane - (long long)img * C);
  const uint4* p = imgs + plane * (long long)plane_vecs;

  unsigned long long S = 0, Q = 0;
  // u32 partials are flushed to u64 every kFlush vectors per thread: 16*255^2*kFlush < 2^32.
  constexpr int kFlush = 2048;
  int i = threadIdx.x;
  while (i < plane_vecs) {
    unsigned s = 0, q = 0;
    int budget = kFlush;  // multiple of kStatsUnroll
    // main body: kStatsUnroll independent loads in flight
    while (budget > 0 && i + (kStatsUnroll - 1) * kStatsThreads < plane_vecs) {
      uint4 v[kStatsUnroll];
#pragma unroll
      for (int u = 0; u < kStatsUnroll; ++u) v[u] = ld_stream_v4(p + i + u * kStatsThreads);
#pragma unroll
      for (int u = 0; u < kStatsUnroll; ++u) acc16(v[u], s, q);
      i += kStatsUnroll * kStatsThreads;
      budget -= kStatsUnroll;
    }
    // tail: fewer than kStatsUnroll vectors left for this thread
    while (budget > 0 && i < plane_vecs && i + (kStatsUnroll - 1) * kStatsThreads >= plane_vecs) {
      uint4 v = ld_stream_v4(p + i);
      acc16(v, s, q);
      i += kStatsThreads;
      --budget;
    }
    S += s;
    Q += q;
  }

  S = warp_sum_u64(S);
  Q = warp_sum_u64(Q);
  __shared__ unsigned long long sh[2][kStatsThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    sh[0][wid] = S;
    sh[1][wid] = Q;
  }
  __syncthreads();
  if (wid == 0) {
    S = lane < kStatsThreads / 32 ? sh[0][lane] : 0ull;
    Q = lane < kStatsThreads / 32 ? sh[1][lane] : 0ull;
    S = warp_sum_u64(S);
    Q = warp_sum_u64(Q);
    if (lane == 0) {
      int e = exp_id[img];
      if (e >= 0 && e < n_exp) {
        atomicAdd(&sum[e * C + ch], S);
        atomicAdd(&sumsq[e * C + ch], Q);
        atomicAdd(&count[e * C + ch], (unsigned long long)plane_vecs * 16ull);
      }
    }
  }
}

// mean/std of x/255 (and of the pre-normalised variable in verification mode), f64 like the reference.
__global__ void stats_finalize_kernel(const unsigned long long* __restrict__ sum,
                                      const unsigned long long* __restrict__ sumsq,
                                      const unsigned long long* __restrict__ count, int total,
                                      const double* __restrict__ pre_mean,
                                      const double* __restrict__ pre_std, double* __restrict__ mean,
                                      double* __restrict__ std) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  double n = (double)count[i];
  double sx = (double)sum[i] / 255.0;                // sum of x/255
  double sx2 = (double)sumsq[i] / (255.0 * 255.0);   // sum of (x/255)^2
  if (pre_mean != nullptr && pre_std != nullptr) {
    // z = (x/255 - m)/s :  sum z = (sx - n m)/s ; sum z^2 = (sx2 - 2 m sx + n m^2)/s^2
    double m = pre_mean[i], s = pre_std[i];
    double sz = (sx - n * m) / s;
    double sz2 = (sx2 - 2.0 * m * sx + n * m * m) / (s * s);
    sx = sz;
    sx2 = sz2;
  }
  double mu = n > 0 ? sx / n : 0.0;
  double var = n > 0 ? sx2 / n - mu * mu : 0.0;
  mean[i] = mu;
  std[i] = sqrt(var > 0.0 ? var : 0.0);
}

}  // namespace rxb

extern "C" {

int rxb_stats_accumulate(const uint8_t* imgs, const int32_t* exp_id, int64_t n, int H, int W, int C,
                         int layout, int n_exp, unsigned long long* sum, unsigned long long* sumsq,
                         unsigned long long* count, rxb_stream_t stream) {
  RXB_CHECK_ARG(n >= 0 && H > 0 && W > 0 && C > 0 && n_exp > 0, "rxb_stats_accumulate: bad sizes");
  RXB_CHECK_ARG(sum && sumsq && count, "rxb_stats_accumulate: null pointer");
  if (n == 0) return RXB_OK;  // an empty chunk is a no-op (its data pointer may be null)
  RXB_CHECK_ARG(imgs && exp_id, "rxb_stats_accumulate: null pointer");
  RXB_CHECK_ARG(((long long)H * W) % 16 == 0, "rxb_stats_accumulate: H*W must be a multiple of 16");
  RXB_CHECK_ARG((reinterpret_cast<uintptr_t>(imgs) & 15) == 0, "rxb_stats_accumulate: imgs not 16B aligned");
  if (layout != RXB_LAYOUT_NCHW)
    return rxb::set_error(RXB_ERR_UNSUPPORTED, "rxb_stats_accumulate: only planar NCHW u8 is supported");
  RXB_CHECK_ARG(n * C < (1ll << 31), "rxb_stats_accumulate: too many planes for one call");
  int rc = rxb_check_device();
  if (rc) return rc;
  int plane_vecs = (int)(((long long)H * W) / 16);
  RXB_PROF(rxb::as_stream(stream), rxb::PROF_STATS);
  rxb::stats_planar_kernel<<<(unsigned)(n * C), rxb::kStatsThreads, 0, rxb::as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(imgs), exp_id, C, plane_vecs, n_exp, sum, sumsq, count);
  RXB_LAUNCH_OK();
  return RXB_OK;
}

int rxb_stats_finalize(const unsigned long long* sum, const unsigned long long* sumsq,
                       const unsigned long long* count, int n_exp, int C, const double* pre_mean,
                       const double* pre_std, double* mean, double* std, rxb_stream_t stream) {
  RXB_CHECK_ARG(sum && sumsq && count && mean && std, "rxb_stats_finalize: null pointer");
  RXB_CHECK_ARG(n_exp > 0 && C > 0, "rxb_stats_finalize: bad sizes");
  int rc = rxb_check_device();
  if (rc) return rc;
  int total = n_exp * C;
  RXB_PROF(rxb::as_stream(stream), rxb::PROF_STATS);
  rxb::stats_finalize_kernel<<<rxb::ceil_div(total, 128), 128, 0, rxb::as_stream(stream)>>>(
      sum, sumsq, count, total, pre_mean, pre_std, mean, std);
  RXB_LAUNCH_OK();
  return RXB_OK;
}

}  // extern "C"
