// resnet_ops.cu — see resnet_ops.cuh.  All kernels are bf16 NHWC, 16-byte vectorised (8 channels per thread) like
// elementwise.cu; none is on the DenseNet path.
#include "resnet_ops.cuh"

namespace rxb {

namespace {
constexpr int kThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
int grid_for(long long items) {
  long long blocks = ceil_div<long long>(items, kThreads);
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
}  // namespace

__global__ void __launch_bounds__(kThreads)
s2d_bn_relu_kernel(const __nv_bfloat16* __restrict__ X, int B, int H, int W, int C, const float* __restrict__ scale,
                   const float* __restrict__ shift, __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  const int groups = C >> 3, Ho = (H + 1) >> 1, Wo = (W + 1) >> 1;
  const long long total = (long long)B * Ho * Wo * 4 * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % groups);
    long long r = i / groups;
    const int q = (int)(r & 3);
    r >>= 2;
    const int ox = (int)(r % Wo);
    r /= Wo;
    const int oy = (int)(r % Ho), b = (int)(r / Ho);
    const int y = 2 * oy + (q >> 1), x = 2 * ox + (q & 1);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (y < H && x < W) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(X + (((long long)b * H + y) * W + x) * C + cg * 8));
      if (scale != nullptr) {
        float f[8], s[8], h[8];
        unpack8(v, f);
        load8f(scale + cg * 8, s);
        load8f(shift + cg * 8, h);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaxf(fmaf(f[e], s[e], h[e]), 0.f);
        o = pack8(f);
      } else {
        o = v;
      }
    }
    *reinterpret_cast<uint4*>(out + ((((long long)b * Ho + oy) * Wo + ox) * 4 + q) * C + cg * 8) = o;
  }
}

int s2d_bn_relu(const __nv_bfloat16* X, int B, int H, int W, int C, const float* scale, const float* shift,
                __nv_bfloat16* out, cudaStream_t st) {
  if (C % 8) return set_error(RXB_ERR_INVALID, "s2d_bn_relu: C=%d must be a multiple of 8", C);
  RXB_PROF(st, PROF_ELEMENTWISE);
  const long long total = (long long)B * ((H + 1) / 2) * ((W + 1) / 2) * 4 * (C / 8);
  RXB_CUDA(launch_k(s2d_bn_relu_kernel, dim3(grid_for(total)), dim3(kThreads), (size_t)0, st, X, B, H, W, C, scale, shift, out));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void __launch_bounds__(kThreads)
subsample2_kernel(const __nv_bfloat16* __restrict__ X, int B, int H, int W, int C, __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  const int groups = C >> 3, Ho = (H + 1) >> 1, Wo = (W + 1) >> 1;
  const long long total = (long long)B * Ho * Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % groups);
    long long r = i / groups;
    const int ox = (int)(r % Wo);
    r /= Wo;
    const int oy = (int)(r % Ho), b = (int)(r / Ho);
    *reinterpret_cast<uint4*>(out + (((long long)b * Ho + oy) * Wo + ox) * C + cg * 8) =
        __ldg(reinterpret_cast<const uint4*>(X + (((long long)b * H + 2 * oy) * W + 2 * ox) * C + cg * 8));
  }
}

int subsample2(const __nv_bfloat16* X, int B, int H, int W, int C, __nv_bfloat16* out, cudaStream_t st) {
  if (C % 8) return set_error(RXB_ERR_INVALID, "subsample2: C=%d must be a multiple of 8", C);
  RXB_PROF(st, PROF_ELEMENTWISE);
  const long long total = (long long)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  RXB_CUDA(launch_k(subsample2_kernel, dim3(grid_for(total)), dim3(kThreads), (size_t)0, st, X, B, H, W, C, out));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void __launch_bounds__(kThreads)
bn_add_relu_kernel(const __nv_bfloat16* __restrict__ c3, const float* __restrict__ s3, const float* __restrict__ h3,
                   const __nv_bfloat16* __restrict__ idn, const float* __restrict__ sd, const float* __restrict__ hd,
                   long long M, int C, __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  const int groups = C >> 3;
  const long long total = M * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % groups);
    float a[8], r[8], s[8], h[8];
    unpack8(ld_stream_v4(c3 + i * 8), a);
    unpack8(ld_stream_v4(idn + i * 8), r);
    load8f(s3 + cg * 8, s);
    load8f(h3 + cg * 8, h);
    if (sd != nullptr) {
      float s2[8], h2[8];
      load8f(sd + cg * 8, s2);
      load8f(hd + cg * 8, h2);
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = fmaf(r[e], s2[e], h2[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = fmaxf(fmaf(a[e], s[e], h[e]) + r[e], 0.f);
    *reinterpret_cast<uint4*>(out + i * 8) = pack8(a);
  }
}

int bn_add_relu(const __nv_bfloat16* c3, const float* s3, const float* h3, const __nv_bfloat16* idn, const float* sd,
                const float* hd, long long M, int C, __nv_bfloat16* out, cudaStream_t st) {
  if (C % 8) return set_error(RXB_ERR_INVALID, "bn_add_relu: C=%d must be a multiple of 8", C);
  RXB_PROF(st, PROF_ELEMENTWISE);
  RXB_CUDA(launch_k(bn_add_relu_kernel, dim3(grid_for(M * (C / 8))), dim3(kThreads), (size_t)0, st, c3, s3, h3, idn, sd, hd, M, C, out));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void __launch_bounds__(kThreads)
gap_mean_kernel(const __nv_bfloat16* __restrict__ X, int B, int HW, int C, float* __restrict__ feat) {
  pdl_sync();
  const int groups = C >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * groups) return;
  const int cg = i % groups, b = i / groups;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int p = 0; p < HW; ++p) {
    float x[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(X + ((long long)b * HW + p) * C + cg * 8)), x);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += x[e];
  }
  const float inv = 1.f / (float)HW;
#pragma unroll
  for (int e = 0; e < 8; ++e) feat[(long long)b * C + cg * 8 + e] = acc[e] * inv;
}

int gap_mean(const __nv_bfloat16* X, int B, int HW, int C, float* feat, cudaStream_t st) {
  if (C % 8) return set_error(RXB_ERR_INVALID, "gap_mean: C=%d must be a multiple of 8", C);
  RXB_PROF(st, PROF_ELEMENTWISE);
  RXB_CUDA(launch_k(gap_mean_kernel, dim3(ceil_div(B * (C / 8), 128)), dim3(128), (size_t)0, st, X, B, HW, C, feat));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void two_sites_concat_kernel(const float* __restrict__ feat, int bs, int G, int F, float* __restrict__ cat) {
  pdl_sync();
  const long long total = (long long)bs * 3 * F;
  const int per = G / 3;                                   // models.py:46: shape = int(G/3)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long r = i / F;
    const int third = (int)(r % 3), b = (int)(r / 3);
    // the last third takes everything from 2*per on (models.py:49: features[:, 2*shape:, :])
    const int g0 = third * per, g1 = third == 2 ? G : g0 + per;
    float s = 0.f;
    for (int g = g0; g < g1; ++g) s += feat[((long long)b * G + g) * F + f];
    cat[i] = s / (float)(g1 - g0);
  }
}

int two_sites_concat(const float* feat, int bs, int G, int F, float* cat, cudaStream_t st) {
  if (G < 3) return set_error(RXB_ERR_INVALID, "two_sites_concat: G=%d images per sample, need at least 3", G);
  RXB_PROF(st, PROF_HEAD);
  RXB_CUDA(launch_k(two_sites_concat_kernel, dim3(grid_for((long long)bs * 3 * F)), dim3(kThreads), (size_t)0, st, feat, bs, G, F, cat));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void affine_rows_kernel(const float* __restrict__ x, long long total, int F, const float* __restrict__ scale,
                                   const float* __restrict__ shift, int pre_relu, float* __restrict__ y) {
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    float v = x[i];
    if (pre_relu) v = fmaxf(v, 0.f);
    y[i] = fmaf(v, scale[f], shift[f]);
  }
}

int affine_rows(const float* x, int rows, int F, const float* scale, const float* shift, int pre_relu, float* y,
                cudaStream_t st) {
  RXB_PROF(st, PROF_HEAD);
  const long long total = (long long)rows * F;
  RXB_CUDA(launch_k(affine_rows_kernel, dim3(grid_for(total)), dim3(kThreads), (size_t)0, st, x, total, F, scale, shift, pre_relu, y));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void bn_fold_eval_all_kernel(const float* __restrict__ params, const float* __restrict__ buffers,
                                        const BnFoldJob* __restrict__ jobs, float eps, float* __restrict__ fold_scale,
                                        float* __restrict__ fold_shift) {
  pdl_sync();
  const BnFoldJob j = jobs[blockIdx.y];
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < j.C; c += gridDim.x * blockDim.x) {
    const float sc = params[j.gamma_off + c] * rsqrtf(buffers[j.rv_off + c] + eps);
    fold_scale[j.fold_off + c] = sc;
    fold_shift[j.fold_off + c] = params[j.beta_off + c] - buffers[j.rm_off + c] * sc;
  }
}

int bn_fold_eval_all(const float* params, const float* buffers, const BnFoldJob* jobs_dev, int n_jobs, int max_c, float eps,
                     float* fold_scale, float* fold_shift, cudaStream_t st) {
  RXB_PROF(st, PROF_ELEMENTWISE);
  int gx = ceil_div(max_c, 256);
  if (gx > 8) gx = 8;
  RXB_CUDA(launch_k(bn_fold_eval_all_kernel, dim3(gx, n_jobs), dim3(256), (size_t)0, st, params, buffers, jobs_dev, eps, fold_scale, fold_shift));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

}  // namespace rxb

// ================================================================================================ training pieces
namespace rxb {

namespace {
// Per-thread partials for channel group cg (8 channels) -> shared -> one global atomic per channel per CTA.
__device__ __forceinline__ void block_channel_reduce2(const float (&s)[8], const float (&q)[8], int cg, int C, float* gsum,
                                                      float* gsq, float* sh) {
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
  __syncthreads();
}
}  // namespace

template <bool DOWN>
__global__ void __launch_bounds__(kThreads)
relu_bwd_sums_kernel(__nv_bfloat16* D, const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ c3, BnFold f3,
                     const __nv_bfloat16* __restrict__ cd, BnFold fd, long long M, int C, float* __restrict__ dsum3,
                     float* __restrict__ dsq3, float* __restrict__ dsumd, float* __restrict__ dsqd) {
  pdl_sync();
  extern __shared__ float sh[];
  const int groups = C >> 3;
  const int cg = threadIdx.x % groups;
  const int ppi = blockDim.x / groups;
  float as[8], a3[8], ad[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) as[e] = a3[e] = ad[e] = 0.f;
  for (long long pix = (long long)blockIdx.x * ppi + threadIdx.x / groups; pix < M; pix += (long long)gridDim.x * ppi) {
    const long long at = pix * C + cg * 8;
    float d[8], o[8], x3[8];
    unpack8(*reinterpret_cast<const uint4*>(D + at), d);
    unpack8(ld_stream_v4(out + at), o);
    unpack8(ld_stream_v4(c3 + at), x3);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      d[e] = o[e] > 0.f ? d[e] : 0.f;
      as[e] += d[e];
      a3[e] = fmaf(d[e], x3[e], a3[e]);
    }
    if (DOWN) {
      float xd[8];
      unpack8(ld_stream_v4(cd + at), xd);
#pragma unroll
      for (int e = 0; e < 8; ++e) ad[e] = fmaf(d[e], xd[e], ad[e]);
    }
    *reinterpret_cast<uint4*>(D + at) = pack8(d);
  }
  {
    float mu[8], rs[8];
    load8f(f3.mean + cg * 8, mu);
    load8f(f3.rstd + cg * 8, rs);
#pragma unroll
    for (int e = 0; e < 8; ++e) a3[e] = rs[e] * (a3[e] - mu[e] * as[e]);     // sum du*xhat
  }
  block_channel_reduce2(as, a3, cg, C, dsum3, dsq3, sh);
  if (DOWN) {
    float mu[8], rs[8];
    load8f(fd.mean + cg * 8, mu);
    load8f(fd.rstd + cg * 8, rs);
#pragma unroll
    for (int e = 0; e < 8; ++e) ad[e] = rs[e] * (ad[e] - mu[e] * as[e]);
    block_channel_reduce2(as, ad, cg, C, dsumd, dsqd, sh);
  }
}

int relu_bwd_sums(__nv_bfloat16* D, const __nv_bfloat16* out, const __nv_bfloat16* c3, BnFold f3, const __nv_bfloat16* cd,
                  BnFold fd, long long M, int C, float* dsum3, float* dsq3, float* dsumd, float* dsqd, cudaStream_t st) {
  const int groups = C / 8;
  if (C % 8 || groups > kThreads || kThreads % groups) return set_error(RXB_ERR_INVALID, "relu_bwd_sums: C=%d", C);
  RXB_PROF(st, PROF_ELEMENTWISE);
  const int ppi = kThreads / groups;
  long long blocks = ceil_div<long long>(M, (long long)ppi * 4);
  const long long cap = (long long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const size_t smem = 2 * (size_t)C * sizeof(float);
  if (cd != nullptr)
    RXB_CUDA(launch_k((relu_bwd_sums_kernel<true>), dim3((unsigned)blocks), dim3(kThreads), smem, st, D, out, c3, f3, cd, fd, M, C, dsum3, dsq3, dsumd, dsqd));
  else
    RXB_CUDA(launch_k((relu_bwd_sums_kernel<false>), dim3((unsigned)blocks), dim3(kThreads), smem, st, D, out, c3, f3, cd, fd, M, C, dsum3, dsq3, dsumd, dsqd));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void __launch_bounds__(kThreads)
s2d_bn_relu_bwd_kernel(const __nv_bfloat16* __restrict__ DS, const __nv_bfloat16* __restrict__ X, int B, int H, int W, int C,
                       BnFold f, __nv_bfloat16* __restrict__ dz, float* __restrict__ dsum, float* __restrict__ dsq) {
  pdl_sync();
  extern __shared__ float sh[];
  const int groups = C >> 3, Ho = (H + 1) >> 1, Wo = (W + 1) >> 1;
  const int cg = threadIdx.x % groups;
  const int ppi = blockDim.x / groups;
  const long long M = (long long)B * H * W;
  float sc[8], sf[8], as[8], aq[8];
  load8f(f.scale + cg * 8, sc);
  load8f(f.shift + cg * 8, sf);
#pragma unroll
  for (int e = 0; e < 8; ++e) as[e] = aq[e] = 0.f;
  for (long long pix = (long long)blockIdx.x * ppi + threadIdx.x / groups; pix < M; pix += (long long)gridDim.x * ppi) {
    const int x_ = (int)(pix % W);
    const long long r = pix / W;
    const int y_ = (int)(r % H), b = (int)(r / H);
    const int q = ((y_ & 1) << 1) | (x_ & 1);
    float g[8], x[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(DS + ((((long long)b * Ho + (y_ >> 1)) * Wo + (x_ >> 1)) * 4 + q) * C + cg * 8)), g);
    unpack8(ld_stream_v4(X + pix * C + cg * 8), x);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      g[e] = fmaf(x[e], sc[e], sf[e]) > 0.f ? g[e] : 0.f;
      as[e] += g[e];
      aq[e] = fmaf(g[e], x[e], aq[e]);
    }
    *reinterpret_cast<uint4*>(dz + pix * C + cg * 8) = pack8(g);
  }
  {
    float mu[8], rs[8];
    load8f(f.mean + cg * 8, mu);
    load8f(f.rstd + cg * 8, rs);
#pragma unroll
    for (int e = 0; e < 8; ++e) aq[e] = rs[e] * (aq[e] - mu[e] * as[e]);
  }
  block_channel_reduce2(as, aq, cg, C, dsum, dsq, sh);
}

int s2d_bn_relu_bwd(const __nv_bfloat16* DS, const __nv_bfloat16* X, int B, int H, int W, int C, BnFold f,
                    __nv_bfloat16* dz, float* dsum, float* dsq, cudaStream_t st) {
  const int groups = C / 8;
  if (C % 8 || groups > kThreads || kThreads % groups) return set_error(RXB_ERR_INVALID, "s2d_bn_relu_bwd: C=%d", C);
  RXB_PROF(st, PROF_ELEMENTWISE);
  const int ppi = kThreads / groups;
  long long blocks = ceil_div<long long>((long long)B * H * W, (long long)ppi * 4);
  const long long cap = (long long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  RXB_CUDA(launch_k(s2d_bn_relu_bwd_kernel, dim3((unsigned)blocks), dim3(kThreads), 2 * (size_t)C * sizeof(float), st, DS, X, B, H, W, C, f, dz, dsum, dsq));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void __launch_bounds__(kThreads)
upsample2_add_kernel(const __nv_bfloat16* __restrict__ DXS, int B, int H, int W, int C, __nv_bfloat16* Din) {
  pdl_sync();
  const int groups = C >> 3, Ho = (H + 1) >> 1, Wo = (W + 1) >> 1;
  const long long total = (long long)B * Ho * Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % groups);
    long long r = i / groups;
    const int ox = (int)(r % Wo);
    r /= Wo;
    const int oy = (int)(r % Ho), b = (int)(r / Ho);
    __nv_bfloat16* dst = Din + (((long long)b * H + 2 * oy) * W + 2 * ox) * C + cg * 8;
    float a[8], d[8];
    unpack8(*reinterpret_cast<const uint4*>(dst), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(DXS + i * 8)), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += d[e];
    *reinterpret_cast<uint4*>(dst) = pack8(a);
  }
}

int upsample2_add(const __nv_bfloat16* DXS, int B, int H, int W, int C, __nv_bfloat16* Din, cudaStream_t st) {
  if (C % 8) return set_error(RXB_ERR_INVALID, "upsample2_add: C=%d", C);
  RXB_PROF(st, PROF_ELEMENTWISE);
  const long long total = (long long)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  RXB_CUDA(launch_k(upsample2_add_kernel, dim3(grid_for(total)), dim3(kThreads), (size_t)0, st, DXS, B, H, W, C, Din));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void __launch_bounds__(kThreads)
gap_mean_bwd_kernel(const float* __restrict__ dfeat, int B, int HW, int C, __nv_bfloat16* __restrict__ D) {
  pdl_sync();
  const int groups = C >> 3;
  const long long total = (long long)B * HW * groups;
  const float inv = 1.f / (float)HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % groups);
    const int b = (int)(i / ((long long)HW * groups));
    float v[8];
    load8f(dfeat + (long long)b * C + cg * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= inv;
    *reinterpret_cast<uint4*>(D + i * 8) = pack8(v);
  }
}

int gap_mean_bwd(const float* dfeat, int B, int HW, int C, __nv_bfloat16* D, cudaStream_t st) {
  if (C % 8) return set_error(RXB_ERR_INVALID, "gap_mean_bwd: C=%d", C);
  RXB_PROF(st, PROF_ELEMENTWISE);
  RXB_CUDA(launch_k(gap_mean_bwd_kernel, dim3(grid_for((long long)B * HW * (C / 8))), dim3(kThreads), (size_t)0, st, dfeat, B, HW, C, D));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void two_sites_concat_bwd_kernel(const float* __restrict__ dcat, int bs, int G, int F, float* __restrict__ dfeat) {
  pdl_sync();
  const long long total = (long long)bs * G * F;
  const int per = G / 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const long long r = i / F;
    const int g = (int)(r % G), b = (int)(r / G);
    int third = g / per;
    if (third > 2) third = 2;
    const int cnt = third == 2 ? G - 2 * per : per;
    dfeat[i] = dcat[((long long)b * 3 + third) * F + f] / (float)cnt;
  }
}

int two_sites_concat_bwd(const float* dcat, int bs, int G, int F, float* dfeat, cudaStream_t st) {
  RXB_PROF(st, PROF_HEAD);
  RXB_CUDA(launch_k(two_sites_concat_bwd_kernel, dim3(grid_for((long long)bs * G * F)), dim3(kThreads), (size_t)0, st, dcat, bs, G, F, dfeat));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void bn1d_train_fwd_kernel(const float* __restrict__ x, int rows, int F, int pre_relu, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar, float eps,
                                      float momentum, float* __restrict__ y, float* __restrict__ save_mean,
                                      float* __restrict__ save_rstd) {
  pdl_sync();
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) {
    float v = x[(long long)r * F + f];
    if (pre_relu) v = fmaxf(v, 0.f);
    s += v;
  }
  const float mean = s / (float)rows;
  float q = 0.f;
  for (int r = 0; r < rows; ++r) {
    float v = x[(long long)r * F + f];
    if (pre_relu) v = fmaxf(v, 0.f);
    q = fmaf(v - mean, v - mean, q);
  }
  const float var = q / (float)rows;
  const float rstd = rsqrtf(var + eps);
  const float g = gamma[f], bta = beta[f];
  for (int r = 0; r < rows; ++r) {
    float v = x[(long long)r * F + f];
    if (pre_relu) v = fmaxf(v, 0.f);
    y[(long long)r * F + f] = fmaf((v - mean) * rstd, g, bta);
  }
  save_mean[f] = mean;
  save_rstd[f] = rstd;
  const float unbiased = rows > 1 ? var * ((float)rows / (float)(rows - 1)) : var;
  rmean[f] = (1.f - momentum) * rmean[f] + momentum * mean;
  rvar[f] = (1.f - momentum) * rvar[f] + momentum * unbiased;
}

int bn1d_train_fwd(const float* x, int rows, int F, int pre_relu, const float* gamma, const float* beta, float* rmean,
                   float* rvar, float eps, float momentum, float* y, float* save_mean, float* save_rstd, cudaStream_t st) {
  RXB_PROF(st, PROF_HEAD);
  RXB_CUDA(launch_k(bn1d_train_fwd_kernel, dim3(ceil_div(F, 128)), dim3(128), (size_t)0, st, x, rows, F, pre_relu, gamma, beta, rmean, rvar, eps,
                    momentum, y, save_mean, save_rstd));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void bn1d_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, int rows, int F, int pre_relu,
                                const float* __restrict__ gamma, const float* __restrict__ save_mean,
                                const float* __restrict__ save_rstd, float* __restrict__ dx, float* __restrict__ dgamma,
                                float* __restrict__ dbeta) {
  pdl_sync();
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float mean = save_mean[f], rstd = save_rstd[f], g = gamma[f];
  float s1 = 0.f, s2 = 0.f;
  for (int r = 0; r < rows; ++r) {
    float v = x[(long long)r * F + f];
    if (pre_relu) v = fmaxf(v, 0.f);
    const float d = dy[(long long)r * F + f];
    s1 += d;
    s2 = fmaf(d, (v - mean) * rstd, s2);
  }
  dgamma[f] = s2;
  dbeta[f] = s1;
  const float m1 = s1 / (float)rows, m2 = s2 / (float)rows;
  for (int r = 0; r < rows; ++r) {
    const float raw = x[(long long)r * F + f];
    const float v = pre_relu ? fmaxf(raw, 0.f) : raw;
    float d = g * rstd * (dy[(long long)r * F + f] - m1 - (v - mean) * rstd * m2);
    if (pre_relu && !(raw > 0.f)) d = 0.f;
    dx[(long long)r * F + f] = d;
  }
}

int bn1d_bwd(const float* dy, const float* x, int rows, int F, int pre_relu, const float* gamma, const float* save_mean,
             const float* save_rstd, float* dx, float* dgamma, float* dbeta, cudaStream_t st) {
  RXB_PROF(st, PROF_HEAD);
  RXB_CUDA(launch_k(bn1d_bwd_kernel, dim3(ceil_div(F, 128)), dim3(128), (size_t)0, st, dy, x, rows, F, pre_relu, gamma, save_mean, save_rstd, dx,
                    dgamma, dbeta));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void mul_elems_kernel(const float* x, const float* __restrict__ m, long long n, float* y) {   // y may alias x
  pdl_sync();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = x[i] * m[i];
}

int mul_elems(const float* x, const float* m, long long n, float* y, cudaStream_t st) {
  RXB_PROF(st, PROF_HEAD);
  RXB_CUDA(launch_k(mul_elems_kernel, dim3(grid_for(n)), dim3(kThreads), (size_t)0, st, x, m, n, y));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

__global__ void wgrad_finish_kernel(const float* __restrict__ scratch, int N, int K, int k, int s2d, float* __restrict__ dW) {
  pdl_sync();
  const int taps = k * k;
  const long long total = (long long)N * K * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const long long r = i / taps;
    const int c = (int)(r % K), n = (int)(r / K);
    float v;
    if (!s2d) {
      v = scratch[((long long)t * N + n) * K + c];
    } else {
      const int dy = t / 3, dx = t - dy * 3;                 // k == 3
      const int sy = dy == 0 ? 0 : 1, py = dy == 1 ? 0 : 1;   // dy = 2*sy + py - 1
      const int sx = dx == 0 ? 0 : 1, px = dx == 1 ? 0 : 1;
      v = scratch[((long long)(sy * 2 + sx) * N + n) * (4 * K) + (py * 2 + px) * K + c];
    }
    dW[i] = v;
  }
}

int wgrad_finish(const float* scratch, int N, int K, int k, int s2d, float* dW, cudaStream_t st) {
  RXB_PROF(st, PROF_WGRAD_OTHER);
  RXB_CUDA(launch_k(wgrad_finish_kernel, dim3(grid_for((long long)N * K * k * k)), dim3(kThreads), (size_t)0, st, scratch, N, K, k, s2d, dW));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

}  // namespace rxb
