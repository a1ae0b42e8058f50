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
