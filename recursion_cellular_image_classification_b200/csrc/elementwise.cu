// elementwise.cu — the HBM-bound kernels around the convolutions of the DenseNet-121 executor.
//
// torchvision's densenet (the trunk the north star swaps into reference models.py:16) runs BatchNorm,
// ReLU, pooling and concatenation as separate ATen kernels.  Here the forward
// BN+ReLU lives inside the conv kernels' A-operand path; what remains are the pooling fusions, the
// parameter folding and the backward pieces, all bf16 NHWC, 16-byte vectorised (8 channels per thread),
// with per-channel reductions folded warp -> shared -> one global atomic per channel per CTA.
#include "elementwise.cuh"

namespace rxb {

constexpr int kEwThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Per-thread partials for channel group cg (8 channels) -> shared -> global.  `sh` holds 2*C floats.
__device__ __forceinline__ void block_channel_reduce(const float (&s)[8], const float (&q)[8], int cg, int C,
                                                     float* gsum, float* gsq, float* sh) {
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (s[e] != 0.f) atomicAdd(&sh[cg * 8 + e], s[e]);
    if (q[e] != 0.f) atomicAdd(&sh[C + cg * 8 + e], q[e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    if (sh[i] != 0.f) atomicAdd(gsum + i, sh[i]);
    if (sh[C + i] != 0.f) atomicAdd(gsq + i, sh[C + i]);
  }
}

static int ew_grid(long long items, int per_block) {
  long long blocks = ceil_div<long long>(items, per_block);
  long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ------------------------------------------------------------------------------------------------ bn_prep
__global__ void bn_prep_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq, float count,
                               const float* __restrict__ gamma, const float* __restrict__ beta,
                               float* __restrict__ rmean, float* __restrict__ rvar, float eps, float momentum,
                               int training, int C, BnFold f) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    mean = sum[c] / count;
    var = fmaxf(sumsq[c] / count - mean * mean, 0.f);
    if (rmean != nullptr) {
      const float unbiased = count > 1.f ? var * (count / (count - 1.f)) : var;
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * unbiased;
    }
  } else {
    mean = rmean[c];
    var = rvar[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  f.scale[c] = sc;
  f.shift[c] = beta[c] - mean * sc;
  f.mean[c] = mean;
  f.rstd[c] = rstd;
}

int bn_prep(const float* sum, const float* sumsq, float count, const float* gamma, const float* beta,
            float* running_mean, float* running_var, float eps, float momentum, int training, int C, BnFold f,
            cudaStream_t st) {
  RXB_PROF(st, PROF_ELEMENTWISE);
  RXB_CUDA(launch_k(bn_prep_kernel, dim3(ceil_div(C, 128)), dim3(128), (size_t)(0), st, sum, sumsq, count, gamma, beta, running_mean, running_var, eps,
                                                  momentum, training, C, f));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

// ------------------------------------------------------------------------------------------------ stem pool fwd
// (three blocks per SM: 102 -> 80 registers with a few spills, 721 -> 580 us; four blocks spill too much, and the same
// cap slowed stem_pool_bwd 578 -> 692 us and changed nothing for bn_relu_bwd_to_G)
__global__ void __launch_bounds__(kEwThreads, 3)
stem_bn_relu_maxpool_kernel(const __nv_bfloat16* __restrict__ S0, int B, int Hs, int Ws,
                            const float* __restrict__ scale, const float* __restrict__ shift,
                            __nv_bfloat16* __restrict__ out, int ld_out, uint8_t* __restrict__ idx,
                            float* __restrict__ sum, float* __restrict__ sumsq) {
  pdl_sync();
  __shared__ float sh[2 * 64];
  const int cg = threadIdx.x & 7;
  const int Ho = Hs >> 1, Wo = Ws >> 1;
  const long long total = (long long)B * Ho * Wo;
  float sc[8], sf[8], as[8], aq[8];
  load8f(scale + cg * 8, sc);
  load8f(shift + cg * 8, sf);
#pragma unroll
  for (int e = 0; e < 8; ++e) as[e] = aq[e] = 0.f;
  for (long long pix = (long long)blockIdx.x * 32 + (threadIdx.x >> 3); pix < total; pix += (long long)gridDim.x * 32) {
    const int ox = (int)(pix % Wo);
    const long long r = pix / Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    // (measured and rejected, round 2: comparing the raw activations - BN+ReLU is monotone - and folding once after the
    // window: fewer instructions per tap, but 718 -> 798 us)
    float best[8];
    int bi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; bi[e] = 0; }
    // all nine window loads are issued before any is used (clamped addresses; out-of-image taps are skipped below)
    uint4 win[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int iy = min(max(2 * oy - 1 + k / 3, 0), Hs - 1), ix = min(max(2 * ox - 1 + k % 3, 0), Ws - 1);
      win[k] = ld_stream_v4(S0 + (((long long)b * Hs + iy) * Ws + ix) * 64 + cg * 8);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int iy = 2 * oy - 1 + k / 3, ix = 2 * ox - 1 + k % 3;
      if (iy < 0 || iy >= Hs || ix < 0 || ix >= Ws) continue;
      float x[8];
      unpack8(win[k], x);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float a = fmaxf(fmaf(x[e], sc[e], sf[e]), 0.f);
        if (a > best[e]) { best[e] = a; bi[e] = k; }
      }
    }
    const uint4 o = pack8(best);
    *reinterpret_cast<uint4*>(out + pix * ld_out + cg * 8) = o;
    float rb[8];
    unpack8(o, rb);
#pragma unroll
    for (int e = 0; e < 8; ++e) { as[e] += rb[e]; aq[e] += rb[e] * rb[e]; }
    uint2 ib;
    ib.x = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
    ib.y = (uint32_t)bi[4] | ((uint32_t)bi[5] << 8) | ((uint32_t)bi[6] << 16) | ((uint32_t)bi[7] << 24);
    *reinterpret_cast<uint2*>(idx + pix * 64 + cg * 8) = ib;
  }
  block_channel_reduce(as, aq, cg, 64, sum, sumsq, sh);
}

int stem_bn_relu_maxpool(const __nv_bfloat16* S0, int B, int Hs, int Ws, const float* scale, const float* shift,
                         __nv_bfloat16* out, int ld_out, uint8_t* idx, float* sum, float* sumsq, cudaStream_t st) {
  RXB_PROF(st, PROF_ELEMENTWISE);
  const long long total = (long long)B * (Hs / 2) * (Ws / 2);
  RXB_CUDA(launch_k(stem_bn_relu_maxpool_kernel, dim3(ew_grid(total, 32)), dim3(kEwThreads), (size_t)(0), st, S0, B, Hs, Ws, scale, shift, out, ld_out,
                                                                        idx, sum, sumsq));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

// ------------------------------------------------------------------------------------------------ transition pool
__global__ void __launch_bounds__(kEwThreads)
transition_pool_fwd_kernel(const __nv_bfloat16* __restrict__ X, int ldx, int B, int H, int W, int C,
                           const float* __restrict__ scale, const float* __restrict__ shift,
                           __nv_bfloat16* __restrict__ P) {
  pdl_sync();
  const int groups = C >> 3;
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)B * Ho * Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % groups);
    const long long pix = i / groups;
    const int ox = (int)(pix % Wo);
    const long long r = pix / Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    float sc[8], sf[8], acc[8];
    load8f(scale + cg * 8, sc);
    load8f(shift + cg * 8, sf);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const long long ip = ((long long)b * H + 2 * oy + dy) * W + 2 * ox + dx;
        float x[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(X + ip * ldx + cg * 8)), x);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += fmaxf(fmaf(x[e], sc[e], sf[e]), 0.f);
      }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= 0.25f;
    *reinterpret_cast<uint4*>(P + pix * C + cg * 8) = pack8(acc);
  }
}

int transition_pool_fwd(const __nv_bfloat16* X, int ldx, int B, int H, int W, int C, const float* scale,
                        const float* shift, __nv_bfloat16* P, cudaStream_t st) {
  RXB_PROF(st, PROF_ELEMENTWISE);
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
  RXB_CUDA(launch_k(transition_pool_fwd_kernel, dim3(ew_grid(total, kEwThreads)), dim3(kEwThreads), (size_t)(0), st, X, ldx, B, H, W, C, scale, shift, P));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

// ------------------------------------------------------------------------------------------------ final GAP
__global__ void __launch_bounds__(kEwThreads)
final_bn_relu_gap_kernel(const __nv_bfloat16* __restrict__ X, int ldx, int B, int HW, int C,
                         const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ feat) {
  pdl_sync();
  const int groups = C >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * groups) return;
  const int cg = i % groups, b = i / groups;
  float sc[8], sf[8], acc[8];
  load8f(scale + cg * 8, sc);
  load8f(shift + cg * 8, sf);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int p = 0; p < HW; ++p) {
    float x[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(X + ((long long)b * HW + p) * ldx + cg * 8)), x);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += fmaxf(fmaf(x[e], sc[e], sf[e]), 0.f);
  }
  const float inv = 1.f / (float)HW;
#pragma unroll
  for (int e = 0; e < 8; ++e) feat[(long long)b * C + cg * 8 + e] = acc[e] * inv;
}

int final_bn_relu_gap(const __nv_bfloat16* X, int ldx, int B, int HW, int C, const float* scale,
                      const float* shift, float* feat, cudaStream_t st) {
  RXB_PROF(st, PROF_ELEMENTWISE);
  const int total = B * (C / 8);
  RXB_CUDA(launch_k(final_bn_relu_gap_kernel, dim3(ceil_div(total, 128)), dim3(128), (size_t)(0), st, X, ldx, B, HW, C, scale, shift, feat));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

// ------------------------------------------------------------------------------------------------ BN/ReLU bwd -> G
// A thread keeps one channel group; per element: mask*da selected once, two accumulators (sum dy, sum dy*x) and one
// multiply for G.  The scale factors (0.25 or 1/HW, rstd, mean) are applied to the per-thread sums at the end:
//   sum dy*xhat = rstd * (sum dy*x - mean * sum dy).
template <int MODE>
__global__ void __launch_bounds__(kEwThreads)
bn_relu_bwd_to_G_kernel(const void* __restrict__ upstream, const __nv_bfloat16* __restrict__ X, int ldx, int B,
                        int H, int W, int C, BnFold f, __nv_bfloat16* __restrict__ G, float* __restrict__ dsum,
                        float* __restrict__ dsq) {
  pdl_sync();
  extern __shared__ float sh[];
  const int groups = C >> 3;
  const int cg = threadIdx.x % groups;
  const int ppi = blockDim.x / groups;  // pixels per block iteration
  const long long total = (long long)B * H * W;
  const float k = MODE == 0 ? 0.25f : 1.f / (float)(H * W);   // avgpool 2x2 / global average pool backward
  float sc[8], sf[8], gk[8], as[8], aq[8];
  load8f(f.scale + cg * 8, sc);
  load8f(f.shift + cg * 8, sf);
#pragma unroll
  for (int e = 0; e < 8; ++e) { as[e] = aq[e] = 0.f; gk[e] = sc[e] * k; }
  const long long pstride = (long long)gridDim.x * ppi;
  long long pix = (long long)blockIdx.x * ppi + threadIdx.x / groups;
  // avgpool backward: four pixels in flight per thread (all loads issued before any is used)
  if (MODE == 0) {
    const __nv_bfloat16* dP = static_cast<const __nv_bfloat16*>(upstream);
    for (; pix + 3 * pstride < total; pix += 4 * pstride) {
      uint4 dv[4], xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long pu = pix + u * pstride;
        const int x_ = (int)(pu % W);
        const long long r = pu / W;
        const int y_ = (int)(r % H);
        const int b = (int)(r / H);
        const long long pp = ((long long)b * (H >> 1) + (y_ >> 1)) * (W >> 1) + (x_ >> 1);
        dv[u] = __ldg(reinterpret_cast<const uint4*>(dP + pp * C + cg * 8));
        xv[u] = ld_stream_v4(X + pu * ldx + cg * 8);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float da[8], x[8], g[8];
        unpack8(dv[u], da);
        unpack8(xv[u], x);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dy = fmaf(x[e], sc[e], sf[e]) > 0.f ? da[e] : 0.f;
          as[e] += dy;
          aq[e] = fmaf(dy, x[e], aq[e]);
          g[e] = gk[e] * dy;
        }
        st_stream_v4(G + (pix + u * pstride) * ldx + cg * 8, pack8(g));
      }
    }
  }
  for (; pix < total; pix += pstride) {
    const int x_ = (int)(pix % W);
    const long long r = pix / W;
    const int y_ = (int)(r % H);
    const int b = (int)(r / H);
    float da[8];
    if (MODE == 0) {
      const __nv_bfloat16* dP = static_cast<const __nv_bfloat16*>(upstream);
      const long long pp = ((long long)b * (H >> 1) + (y_ >> 1)) * (W >> 1) + (x_ >> 1);
      unpack8(__ldg(reinterpret_cast<const uint4*>(dP + pp * C + cg * 8)), da);
    } else {
      load8f(static_cast<const float*>(upstream) + (long long)b * C + cg * 8, da);
    }
    float x[8], g[8];
    unpack8(ld_stream_v4(X + pix * ldx + cg * 8), x);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float dy = fmaf(x[e], sc[e], sf[e]) > 0.f ? da[e] : 0.f;   // unscaled: k is applied at the end / in gk
      as[e] += dy;
      aq[e] = fmaf(dy, x[e], aq[e]);
      g[e] = gk[e] * dy;
    }
    st_stream_v4(G + pix * ldx + cg * 8, pack8(g));
  }
  {
    float mu[8], rs[8];
    load8f(f.mean + cg * 8, mu);
    load8f(f.rstd + cg * 8, rs);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      aq[e] = k * rs[e] * (aq[e] - mu[e] * as[e]);
      as[e] *= k;
    }
  }
  block_channel_reduce(as, aq, cg, C, dsum, dsq, sh);
}

int bn_relu_bwd_to_G(int mode, const void* upstream, const __nv_bfloat16* X, int ldx, int B, int H, int W, int C,
                     BnFold f, __nv_bfloat16* G, float* dsum, float* dsq, cudaStream_t st) {
  RXB_PROF(st, PROF_ELEMENTWISE);
  const int groups = C / 8;
  if (groups > kEwThreads || kEwThreads % groups) return set_error(RXB_ERR_INVALID, "bn_relu_bwd_to_G: C=%d", C);
  const int ppi = kEwThreads / groups;
  const int grid = ew_grid((long long)B * H * W, ppi * 4);
  const size_t smem = 2 * (size_t)C * sizeof(float);
  if (mode == 0)
    RXB_CUDA(launch_k((bn_relu_bwd_to_G_kernel<0>), dim3(grid), dim3(kEwThreads), (size_t)(smem), st, upstream, X, ldx, B, H, W, C, f, G, dsum, dsq));
  else
    RXB_CUDA(launch_k((bn_relu_bwd_to_G_kernel<1>), dim3(grid), dim3(kEwThreads), (size_t)(smem), st, upstream, X, ldx, B, H, W, C, f, G, dsum, dsq));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

// ------------------------------------------------------------------------------------------------ BN bwd finalize
// sum_p dy*x of input channel c from the finished weight gradient of the conv that consumed A' = relu(scale*x+shift):
// sum_p dy*A' = sum_{k,tap} W[k][c][tap]*dW[k][c][tap] (both sides equal sum_p dL/dA' * A'), and on the ReLU's support
// A' = scale*x + shift, so sum dy*x = (W.dW - shift*sum dy)/scale.  The identity is evaluated with the operands the
// kernels really used: W rounded to bf16 (the GEMM operand), scale/shift rounded to bf16 (the prologue's fold operands)
// - otherwise the 2^-9 operand mismatch is amplified by 1/scale.
// One warp per channel (the K*taps products are strided through the OIHW tensors); every lane returns the result.
__device__ __forceinline__ float sum_dyx_from_wdw(const float* __restrict__ W, const float* __restrict__ dW, int K, int C,
                                                  int taps, int c, float es, float eh, float sum_dy, int lane) {
  float t = 0.f;
  const int n = K * taps;
  for (int i = lane; i < n; i += 32) {
    const int k = i / taps, tp = i - k * taps;
    const long long at = ((long long)k * C + c) * taps + tp;
    t = fmaf(bf16_round(__ldg(W + at)), __ldg(dW + at), t);
  }
  t = warp_sum(t);
  es = bf16_round(es);
  eh = bf16_round(eh);
  return es != 0.f ? (t - eh * sum_dy) / es : 0.f;
}

__global__ void __launch_bounds__(256)
sum_dyx_from_wdw_kernel(const float* __restrict__ W, const float* __restrict__ dW, int K, int C, int taps,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ sum_dy, float* __restrict__ out) {
  pdl_sync();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  const float r = sum_dyx_from_wdw(W, dW, K, C, taps, c, scale[c], shift[c], sum_dy[c], lane);
  if (lane == 0) out[c] = r;
}

// one warp per channel
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(int mode, const float* __restrict__ W, const float* __restrict__ dW, int K, int taps,
                       float* __restrict__ dsum, float* __restrict__ dsq, BnFold f, float count, int C,
                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ corrA,
                       float* __restrict__ corrB, const float* __restrict__ gamma, const float* __restrict__ beta) {
  pdl_sync();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  const float s = dsum[c];
  float q;
  if (W) {
    // degenerate channels: the data-gradient epilogue reduced sum dy*x directly into dsq (conv_gemm.cu)
    const bool direct = gamma != nullptr && bn_degenerate(gamma[c], beta[c]);
    const float raw = direct ? dsq[c] : sum_dyx_from_wdw(W, dW, K, C, taps, c, f.scale[c], f.shift[c], s, lane);
    q = f.rstd[c] * (raw - f.mean[c] * s);                       // sum dy*xhat
  } else {
    q = dsq[c];
  }
  if (lane != 0) return;
  dgamma[c] = q;
  dbeta[c] = s;
  const float m1 = s / count, m2 = q / count;
  if (mode == 0) {
    corrA[c] += f.scale[c] * m1;
    corrB[c] += f.scale[c] * m2;
  } else {
    dsum[c] = m1;
    dsq[c] = m2;
  }
}

int bn_bwd_finalize(int mode, const float* W, const float* dW, int K, int taps, float* dsum, float* dsq, BnFold f,
                    float count, int C, float* dgamma, float* dbeta, float* corrA, float* corrB, const float* gamma,
                    const float* beta, cudaStream_t st) {
  RXB_PROF(st, PROF_EW_FINALIZE);
  RXB_CUDA(launch_k(bn_bwd_finalize_kernel, dim3(ceil_div(C, 8)), dim3(256), (size_t)(0), st, mode, W, dW, K, taps, dsum, dsq, f, count, C, dgamma, dbeta,
                                                        corrA, corrB, gamma, beta));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

// ------------------------------------------------------------------------------------------------ BN bwd apply
// dx = scale*(dy - m1 - xhat*m2) = scale*dy + cb*x + cc  with  cb = -scale*rstd*m2,  cc = scale*(rstd*m2*mean - m1).
// A thread keeps ONE channel group (its folded constants live in registers) and walks pixels, four rows in flight.
__global__ void __launch_bounds__(kEwThreads)
bn_bwd_apply_kernel(__nv_bfloat16* dy, const __nv_bfloat16* __restrict__ X, long long M, int C,
                    BnFold f, const float* __restrict__ m1, const float* __restrict__ m2, __nv_bfloat16* dst,
                    const int raw_on, const BnRawSums raw, const int reverse) {
  pdl_sync();
  const int groups = C >> 3;                         // divides the block size (C = 64 or 128)
  const int cg = threadIdx.x % groups;
  const int rows_per_block = kEwThreads / groups;
  float sc[8], cb[8], cc[8];
  {
    float mu[8], rs[8], a1[8], a2[8];
    load8f(f.scale + cg * 8, sc);
    load8f(f.mean + cg * 8, mu);
    load8f(f.rstd + cg * 8, rs);
    load8f(m1 + cg * 8, a1);
    load8f(m2 + cg * 8, a2);
    if (raw_on) {
      // a1 = sum dy, a2 = W.dW (or the direct sum dy*x): the means, as bn_bwd_finalize mode 1 derives them
      float sh[8], ga[8], be[8];
      load8f(f.shift + cg * 8, sh);
      load8f(raw.gamma + cg * 8, ga);
      load8f(raw.beta + cg * 8, be);
      const bool writer = blockIdx.x == 0 && (int)threadIdx.x < groups;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float s = a1[e];
        const float es = bf16_round(sc[e]), eh = bf16_round(sh[e]);
        const float rawv = bn_degenerate(ga[e], be[e]) ? a2[e] : (es != 0.f ? (a2[e] - eh * s) / es : 0.f);
        const float q = rs[e] * (rawv - mu[e] * s);
        if (writer) {
          raw.dgamma[cg * 8 + e] = q;
          raw.dbeta[cg * 8 + e] = s;
        }
        a1[e] = s * raw.inv_count;
        a2[e] = q * raw.inv_count;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      cb[e] = -sc[e] * rs[e] * a2[e];
      cc[e] = sc[e] * (rs[e] * a2[e] * mu[e] - a1[e]);
    }
  }
  const long long stride = (long long)gridDim.x * rows_per_block;
  long long row = (long long)blockIdx.x * rows_per_block + threadIdx.x / groups;
  // (reverse: the same rows, visited from the last to the first)
  const long long rbase = reverse ? M - 1 : 0, rsign = reverse ? -1 : 1;
  for (; row + 3 * stride < M; row += 4 * stride) {
    uint4 dv[4], xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = rbase + rsign * (row + u * stride);
      dv[u] = *reinterpret_cast<const uint4*>(dy + r * C + cg * 8);   // in place: coherent load
      xv[u] = ld_stream_v4(X + r * C + cg * 8);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long r = rbase + rsign * (row + u * stride);
      float d[8], x[8];
      unpack8(dv[u], d);
      unpack8(xv[u], x);
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] = fmaf(sc[e], d[e], fmaf(cb[e], x[e], cc[e]));
      st_stream_v4(dst + r * C + cg * 8, pack8(d));
    }
  }
  for (; row < M; row += stride) {
    const long long r = rbase + rsign * row;
    float d[8], x[8];
    unpack8(*reinterpret_cast<const uint4*>(dy + r * C + cg * 8), d);
    unpack8(ld_stream_v4(X + r * C + cg * 8), x);
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] = fmaf(sc[e], d[e], fmaf(cb[e], x[e], cc[e]));
    st_stream_v4(dst + r * C + cg * 8, pack8(d));
  }
}

// Channel counts that do not divide the 256-thread block (ResNet's 2048-wide maps are fine; anything with C/8 > 256 or
// not a power of two): a plain grid-stride variant, one 8-channel group per thread per iteration.
__global__ void __launch_bounds__(kEwThreads)
bn_bwd_apply_wide_kernel(const __nv_bfloat16* dy, const __nv_bfloat16* __restrict__ X, long long M, int C, BnFold f,
                         const float* __restrict__ m1, const float* __restrict__ m2, __nv_bfloat16* dst) {
  pdl_sync();
  const int groups = C >> 3;
  const long long total = M * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % groups);
    float sc[8], mu[8], rs[8], a1[8], a2[8], d[8], x[8];
    load8f(f.scale + cg * 8, sc);
    load8f(f.mean + cg * 8, mu);
    load8f(f.rstd + cg * 8, rs);
    load8f(m1 + cg * 8, a1);
    load8f(m2 + cg * 8, a2);
    unpack8(*reinterpret_cast<const uint4*>(dy + i * 8), d);
    unpack8(ld_stream_v4(X + i * 8), x);
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] = sc[e] * (d[e] - a1[e] - (x[e] - mu[e]) * rs[e] * a2[e]);
    st_stream_v4(dst + i * 8, pack8(d));
  }
}

int bn_bwd_apply(__nv_bfloat16* dy, const __nv_bfloat16* X, long long M, int C, BnFold f, const float* m1,
                 const float* m2, cudaStream_t st, __nv_bfloat16* dst, const BnRawSums* raw, int reverse) {
  if (C % 8) return set_error(RXB_ERR_INVALID, "bn_bwd_apply: C=%d must be a multiple of 8", C);
  if (raw != nullptr && (C / 8 > kEwThreads || kEwThreads % (C / 8)))
    return set_error(RXB_ERR_INVALID, "bn_bwd_apply: raw sums need a channel count that divides the block (C=%d)", C);
  if (dst == nullptr) dst = dy;
  RXB_PROF(st, PROF_EW_BN_APPLY);
  if (C / 8 > kEwThreads || kEwThreads % (C / 8)) {
    RXB_CUDA(launch_k(bn_bwd_apply_wide_kernel, dim3(ew_grid(M * (C / 8), kEwThreads * 2)), dim3(kEwThreads), (size_t)(0), st,
                      (const __nv_bfloat16*)dy, X, M, C, f, m1, m2, dst));
  } else {
    const int rows_per_block = kEwThreads / (C / 8);
    const BnRawSums none = {nullptr, nullptr, nullptr, nullptr, 0.f};
    RXB_CUDA(launch_k(bn_bwd_apply_kernel, dim3(ew_grid(M, rows_per_block * 4)), dim3(kEwThreads), (size_t)(0), st, dy, X, M, C, f, m1, m2, dst,
                      raw != nullptr ? 1 : 0, raw != nullptr ? *raw : none, reverse));
  }
  RXB_LAUNCH_OK();
  return RXB_OK;
}

// ------------------------------------------------------------------------------------------------ grad fixup
// dX = G - corrA - xhat*corrB = G + kb*x + kc  with  kb = -rstd*corrB,  kc = mean*rstd*corrB - corrA.
// A thread keeps ONE 8-channel group of the slice (its two folded constants live in registers) and walks pixels, four
// rows in flight: the slice is a strided 16*groups-byte run per pixel of the concat buffers, so bytes in flight - not
// instruction count - set the rate.
__global__ void __launch_bounds__(kEwThreads)
grad_fixup_rows_kernel(const __nv_bfloat16* __restrict__ G, const __nv_bfloat16* __restrict__ X, int ld, long long M,
                       int c0, int nch, const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ corrA, const float* __restrict__ corrB, __nv_bfloat16* __restrict__ dst) {
  pdl_sync();
  const int groups = nch >> 3;                       // divides the block size
  const int cg = threadIdx.x % groups;
  const int rows_per_block = kEwThreads / groups;
  const int c = c0 + cg * 8;
  float kb[8], kc[8];
  {
    float mu[8], rs[8], ca[8], cb[8];
    load8f(mean + c, mu);
    load8f(rstd + c, rs);
    load8f(corrA + c, ca);
    load8f(corrB + c, cb);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      kb[e] = -rs[e] * cb[e];
      kc[e] = mu[e] * rs[e] * cb[e] - ca[e];
    }
  }
  const long long stride = (long long)gridDim.x * rows_per_block;
  long long row = (long long)blockIdx.x * rows_per_block + threadIdx.x / groups;
  for (; row + 3 * stride < M; row += 4 * stride) {
    uint4 gv[4], xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      gv[u] = ld_stream_v4(G + (row + u * stride) * ld + c);
      xv[u] = ld_stream_v4(X + (row + u * stride) * ld + c);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float g[8], x[8];
      unpack8(gv[u], g);
      unpack8(xv[u], x);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = g[e] + fmaf(kb[e], x[e], kc[e]);
      *reinterpret_cast<uint4*>(dst + (row + u * stride) * nch + cg * 8) = pack8(g);
    }
  }
  for (; row < M; row += stride) {
    float g[8], x[8];
    unpack8(ld_stream_v4(G + row * ld + c), g);
    unpack8(ld_stream_v4(X + row * ld + c), x);
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = g[e] + fmaf(kb[e], x[e], kc[e]);
    *reinterpret_cast<uint4*>(dst + row * nch + cg * 8) = pack8(g);
  }
}

__global__ void __launch_bounds__(kEwThreads)
grad_fixup_kernel(const __nv_bfloat16* __restrict__ G, const __nv_bfloat16* __restrict__ X, int ld, long long M,
                  int c0, int nch, const float* __restrict__ mean, const float* __restrict__ rstd,
                  const float* __restrict__ corrA, const float* __restrict__ corrB, __nv_bfloat16* __restrict__ dst) {
  pdl_sync();
  const int groups = nch >> 3;
  const long long total = M * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % groups);
    const long long pix = i / groups;
    const int c = c0 + cg * 8;
    float mu[8], rs[8], ca[8], cb[8], g[8], x[8];
    load8f(mean + c, mu);
    load8f(rstd + c, rs);
    load8f(corrA + c, ca);
    load8f(corrB + c, cb);
    unpack8(__ldg(reinterpret_cast<const uint4*>(G + pix * ld + c)), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(X + pix * ld + c)), x);
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = g[e] - ca[e] - (x[e] - mu[e]) * rs[e] * cb[e];
    *reinterpret_cast<uint4*>(dst + pix * nch + cg * 8) = pack8(g);
  }
}

int grad_fixup(const __nv_bfloat16* G, const __nv_bfloat16* X, int ld, long long M, int c0, int nch,
               const float* mean, const float* rstd, const float* corrA, const float* corrB, __nv_bfloat16* dst,
               cudaStream_t st) {
  RXB_PROF(st, PROF_EW_FIXUP);
  const int groups = nch / 8;
  if (nch % 8 == 0 && groups >= 1 && groups <= kEwThreads && kEwThreads % groups == 0) {
    const int rows_per_block = kEwThreads / groups;
    RXB_CUDA(launch_k(grad_fixup_rows_kernel, dim3(ew_grid(M, rows_per_block * 4)), dim3(kEwThreads), (size_t)(0), st, G, X,
                      ld, M, c0, nch, mean, rstd, corrA, corrB, dst));
    RXB_LAUNCH_OK();
    return RXB_OK;
  }
  RXB_CUDA(launch_k(grad_fixup_kernel, dim3(ew_grid(M * (nch / 8), kEwThreads * 2)), dim3(kEwThreads), (size_t)(0), st, G, X, ld, M, c0, nch, mean, rstd,
                                                                                 corrA, corrB, dst));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

// ------------------------------------------------------------------------------------------------ stem pool bwd
// One thread = 8 channels of a 2x2 block of stem pixels {2q,2q+1} x {2r,2r+1}: the four pooling windows that can
// contain them (oy in {q,q+1}, ox in {r,r+1}) are loaded once, all loads are independent and issued up front.
__global__ void __launch_bounds__(kEwThreads)
stem_pool_bwd_kernel(const __nv_bfloat16* __restrict__ dPool, const uint8_t* __restrict__ idx,
                     const __nv_bfloat16* __restrict__ S0, int B, int Hs, int Ws, BnFold f,
                     __nv_bfloat16* __restrict__ dy0, float* __restrict__ dsum, float* __restrict__ dsq) {
  pdl_sync();
  __shared__ float sh[2 * 64];
  const int cg = threadIdx.x & 7;
  const int Ho = Hs >> 1, Wo = Ws >> 1;
  const long long total = (long long)B * Ho * Wo;
  float sc[8], sf[8], mu[8], rs[8], as[8], aq[8];
  load8f(f.scale + cg * 8, sc);
  load8f(f.shift + cg * 8, sf);
  load8f(f.mean + cg * 8, mu);
  load8f(f.rstd + cg * 8, rs);
#pragma unroll
  for (int e = 0; e < 8; ++e) as[e] = aq[e] = 0.f;
  for (long long quad = (long long)blockIdx.x * 32 + (threadIdx.x >> 3); quad < total; quad += (long long)gridDim.x * 32) {
    const int r = (int)(quad % Wo);
    const long long t = quad / Wo;
    const int q = (int)(t % Ho);
    const int b = (int)(t / Ho);
    uint2 wi[2][2];
    uint4 wd[2][2];
    uint4 xs[2][2];
#pragma unroll
    for (int wy = 0; wy < 2; ++wy)
#pragma unroll
      for (int wx = 0; wx < 2; ++wx) {
        const int oy = q + wy, ox = r + wx;
        const bool ok = oy < Ho && ox < Wo;
        const long long op = ((long long)b * Ho + (ok ? oy : q)) * Wo + (ok ? ox : r);
        wi[wy][wx] = __ldg(reinterpret_cast<const uint2*>(idx + op * 64 + cg * 8));
        wd[wy][wx] = __ldg(reinterpret_cast<const uint4*>(dPool + op * 64 + cg * 8));
        if (!ok) wi[wy][wx] = make_uint2(0xffffffffu, 0xffffffffu);   // matches no window position
      }
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px)
        xs[py][px] = ld_stream_v4(S0 + (((long long)b * Hs + 2 * q + py) * Ws + 2 * r + px) * 64 + cg * 8);
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        float g[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = 0.f;
        // window (q+wy, r+wx) contains pixel (2q+py, 2r+px) at ky = py - 2wy + 1, kx = px - 2wx + 1 (when in 0..2)
#pragma unroll
        for (int wy = 0; wy <= py; ++wy)
#pragma unroll
          for (int wx = 0; wx <= px; ++wx) {
            const uint32_t code = (uint32_t)((py - 2 * wy + 1) * 3 + (px - 2 * wx + 1));
            float d[8];
            unpack8(wd[wy][wx], d);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint32_t w = e < 4 ? wi[wy][wx].x : wi[wy][wx].y;
              if (((w >> (8 * (e & 3))) & 0xffu) == code) g[e] += d[e];
            }
          }
        float x[8];
        unpack8(xs[py][px], x);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float dy = fmaf(x[e], sc[e], sf[e]) > 0.f ? g[e] : 0.f;
          g[e] = dy;
          as[e] += dy;
          aq[e] = fmaf(dy, x[e], aq[e]);     // sum dy*x; turned into sum dy*xhat once, below
        }
        st_stream_v4(dy0 + (((long long)b * Hs + 2 * q + py) * Ws + 2 * r + px) * 64 + cg * 8, pack8(g));
      }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) aq[e] = rs[e] * (aq[e] - mu[e] * as[e]);
  block_channel_reduce(as, aq, cg, 64, dsum, dsq, sh);
}

int stem_pool_bwd(const __nv_bfloat16* dPool, const uint8_t* idx, const __nv_bfloat16* S0, int B, int Hs, int Ws,
                  BnFold f, __nv_bfloat16* dy0, float* dsum, float* dsq, cudaStream_t st) {
  RXB_PROF(st, PROF_ELEMENTWISE);
  RXB_CUDA(launch_k(stem_pool_bwd_kernel, dim3(ew_grid((long long)B * (Hs / 2) * (Ws / 2), 32 * 2)), dim3(kEwThreads), (size_t)(0), st, dPool, idx, S0, B, Hs,
                                                                                                  Ws, f, dy0, dsum, dsq));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

// ------------------------------------------------------------------------------------------------ head GEMMs
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(int M, int N, int K, const float* __restrict__ A, long long a_i, long long a_l,
                     const float* __restrict__ Bm, long long b_l, long long b_j, const float* __restrict__ bias,
                     float* __restrict__ C, long long c_i, long long c_j) {
  pdl_sync();
  __shared__ float sA[32][33], sB[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int l0 = 0; l0 < K; l0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int rr = ty + 8 * r;
      const int i = i0 + rr, l = l0 + tx;
      sA[rr][tx] = (i < M && l < K) ? A[i * a_i + l * a_l] : 0.f;
      const int l2 = l0 + rr, j = j0 + tx;
      sB[rr][tx] = (l2 < K && j < N) ? Bm[l2 * b_l + j * b_j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < 32; ++l) {
      const float bv = sB[l][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(sA[ty + 8 * r][l], bv, acc[r]);
    }
    __syncthreads();
  }
  const int j = j0 + tx;
  if (j < N) {
    const float bj = bias ? bias[j] : 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty + 8 * r;
      if (i < M) C[i * c_i + j * c_j] = acc[r] + bj;
    }
  }
}

int sgemm_strided(int M, int N, int K, const float* A, long long a_i, long long a_l, const float* Bm, long long b_l,
                  long long b_j, const float* bias, float* C, long long c_i, long long c_j, cudaStream_t st) {
  RXB_PROF(st, PROF_HEAD);
  dim3 grid(ceil_div(N, 32), ceil_div(M, 32));
  RXB_CUDA(launch_k(sgemm_strided_kernel, dim3(grid), dim3(256), (size_t)(0), st, M, N, K, A, a_i, a_l, Bm, b_l, b_j, bias, C, c_i, c_j));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void column_sum_kernel(const float* __restrict__ A, int rows, int cols, long long ld,
                                  float* __restrict__ out) {
  pdl_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  float s = 0.f;
  for (int i = 0; i < rows; ++i) s += A[i * ld + j];
  out[j] = s;
}
int column_sum(const float* A, int rows, int cols, long long ld, float* out, cudaStream_t st) {
  RXB_PROF(st, PROF_ELEMENTWISE);
  RXB_CUDA(launch_k(column_sum_kernel, dim3(ceil_div(cols, 128)), dim3(128), (size_t)(0), st, A, rows, cols, ld, out));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void sum_scale_kernel(const float* __restrict__ v, int n, float scale, float* __restrict__ out) {
  pdl_sync();
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) out[0] = s * scale;
  }
}
int sum_scale(const float* v, int n, float scale, float* out, cudaStream_t st) {
  RXB_PROF(st, PROF_ELEMENTWISE);
  RXB_CUDA(launch_k(sum_scale_kernel, dim3(1), dim3(256), (size_t)(0), st, v, n, scale, out));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

// ------------------------------------------------------------------------------------------------ weight repack
__global__ void __launch_bounds__(256)
repack_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ arena, const RepackJob* __restrict__ jobs) {
  pdl_sync();
  const RepackJob j = jobs[blockIdx.y];
  const float* src = params + j.src_off;
  __nv_bfloat16* dst = arena + j.dst_off;
  const int N = j.N, K = j.K;
  long long total;
  switch (j.type) {
    case RP_1x1_FWD:
    case RP_1x1_DGRAD: total = (long long)N * K; break;
    case RP_3x3_FWD:
    case RP_3x3_DGRAD: total = 9ll * N * K; break;
    case RP_3x3S2_FWD:
    case RP_3x3S2_DGRAD: total = 16ll * N * K; break;   // [4 taps][N][4K] / [4 taps][4K][N]
    default: total = 16ll * N * 32; break;  // stem: [16 taps][N][32]
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v = 0.f;
    switch (j.type) {
      case RP_1x1_FWD: v = src[i]; break;                      // dst[n][k] = W[n][k]
      case RP_1x1_DGRAD: {                                      // dst[k][n] = W[n][k]
        const int n = (int)(i % N), k = (int)(i / N);
        v = src[(long long)n * K + k];
      } break;
      case RP_3x3_FWD: {                                        // dst[tap][n][k] = W[n][k][tap]
        const int k = (int)(i % K);
        const long long r = i / K;
        const int n = (int)(r % N), tap = (int)(r / N);
        v = src[((long long)n * K + k) * 9 + tap];
      } break;
      case RP_3x3_DGRAD: {                                      // dst[8-tap][k][n] = W[n][k][tap]
        const int n = (int)(i % N);
        const long long r = i / N;
        const int k = (int)(r % K), tapf = (int)(r / K);
        v = src[((long long)n * K + k) * 9 + (8 - tapf)];
      } break;
      case RP_3x3S2_FWD: {                                      // dst[(sy,sx)][n][(py*2+px)*K + c] = W[n][c][dy][dx]
        const int cc = (int)(i % (4 * K));
        const long long r = i / (4 * K);
        const int n = (int)(r % N), tap = (int)(r / N);
        const int sy = tap >> 1, sx = tap & 1;
        const int q = cc / K, c = cc - q * K;
        const int dy = 2 * sy + (q >> 1) - 1, dx = 2 * sx + (q & 1) - 1;
        if (dy >= 0 && dy < 3 && dx >= 0 && dx < 3) v = src[(((long long)n * K + c) * 3 + dy) * 3 + dx];
      } break;
      case RP_3x3S2_DGRAD: {                                    // dst[flipped tap][(py*2+px)*K + c][n]
        const int n = (int)(i % N);
        const long long r = i / N;
        const int cc = (int)(r % (4 * K)), tapf = (int)(r / (4 * K));
        const int sy = 1 - (tapf >> 1), sx = 1 - (tapf & 1);
        const int q = cc / K, c = cc - q * K;
        const int dy = 2 * sy + (q >> 1) - 1, dx = 2 * sx + (q & 1) - 1;
        if (dy >= 0 && dy < 3 && dx >= 0 && dx < 3) v = src[(((long long)n * K + c) * 3 + dy) * 3 + dx];
      } break;
      default: {                                                // stem: dst[(sy,sx)][n][(py,px,c)] = W[n][c][dy][dx]
        const int cc = (int)(i % 32);
        const long long r = i / 32;
        const int n = (int)(r % N), tap = (int)(r / N);
        const int sy = tap >> 2, sx = tap & 3;
        const int c = cc & 7, px = (cc >> 3) & 1, py = (cc >> 4) & 1;
        const int dy = 2 * sy + py - 1, dx = 2 * sx + px - 1;
        if (c < 6 && dy >= 0 && dy < 7 && dx >= 0 && dx < 7) v = src[(((long long)n * 6 + c) * 7 + dy) * 7 + dx];
      } break;
    }
    dst[i] = __float2bfloat16_rn(v);
  }
}

int repack_weights(const float* params, __nv_bfloat16* arena, const RepackJob* jobs_dev, int n_jobs,
                   long long max_elems, cudaStream_t st) {
  RXB_PROF(st, PROF_OPTIM);
  int gx = (int)ceil_div<long long>(max_elems, 256 * 8);
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  RXB_CUDA(launch_k(repack_kernel, dim3(dim3(gx, n_jobs)), dim3(256), (size_t)(0), st, params, arena, jobs_dev));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

int sum_dyx_from_wdw_launch(const float* W, const float* dW, int K, int C, int taps, const float* scale,
                            const float* shift, const float* sum_dy, float* out, cudaStream_t st) {
  RXB_PROF(st, PROF_ELEMENTWISE);
  RXB_CUDA(launch_k(sum_dyx_from_wdw_kernel, dim3(ceil_div(C, 8)), dim3(256), (size_t)(0), st, W, dW, K, C, taps, scale, shift, sum_dy, out));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

}  // namespace rxb
