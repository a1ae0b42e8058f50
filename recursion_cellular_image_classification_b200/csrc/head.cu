// head.cu — kernel family 3 (loss part) and the optimizer.
//
// rxb_softmax_ce restates nn.CrossEntropyLoss forward+backward (reference cell_classifier/train.py:37):
// per-row log-sum-exp in fp32, NLL of the target class, and d(logits) = (softmax - onehot) * scale.
// rxb_sgd_step restates torch.optim.SGD with momentum/nesterov/weight-decay (reference main.py:89-93).
// Both are HBM-bound elementwise/row kernels.
#include <math.h>
#include "common.cuh"

namespace rxb {

constexpr int kCeThreads = 256;

__global__ void __launch_bounds__(kCeThreads)
softmax_ce_kernel(const float* __restrict__ logits, int ld, const long long* __restrict__ target, int C,
                  float* __restrict__ loss_rows, float* __restrict__ dlogits, float grad_scale) {
  pdl_sync();
  __shared__ float red[kCeThreads / 32];
  __shared__ float bc;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float* x = logits + (long long)b * ld;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < C; c += kCeThreads) m = fmaxf(m, x[c]);
  m = warp_max(m);
  if (lane == 0) red[wid] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = red[0];
    for (int i = 1; i < kCeThreads / 32; ++i) mm = fmaxf(mm, red[i]);
    bc = mm;
  }
  __syncthreads();
  m = bc;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += kCeThreads) s += expf(x[c] - m);
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float ss = 0.f;
    for (int i = 0; i < kCeThreads / 32; ++i) ss += red[i];
    bc = ss;
  }
  __syncthreads();
  s = bc;
  const long long t = target[b];
  const float lse = m + logf(s);
  if (threadIdx.x == 0) loss_rows[b] = (t >= 0 && t < C) ? lse - x[t] : 0.f;
  if (dlogits != nullptr) {
    float* g = dlogits + (long long)b * ld;
    const float inv = 1.f / s;
    for (int c = threadIdx.x; c < C; c += kCeThreads) {
      float p = expf(x[c] - m) * inv;
      g[c] = (p - (c == t ? 1.f : 0.f)) * grad_scale;
    }
  }
}

__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ mom, long long n, float lr,
           float mu, float wd, int nesterov, float grad_scale) {
  pdl_sync();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float w = p[i];
    float g = fmaf(wd, w, grad[i] * grad_scale);
    float m = fmaf(mu, mom[i], g);
    mom[i] = m;
    float step = nesterov ? fmaf(mu, m, g) : m;
    p[i] = w - lr * step;
  }
}

}  // namespace rxb

extern "C" {

int rxb_softmax_ce(const float* logits, int ld, const int64_t* target, int B, int C, float* loss_rows,
                   float* dlogits, float grad_scale, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(logits && target && loss_rows, "rxb_softmax_ce: null pointer");
  RXB_CHECK_ARG(B >= 0 && C >= 1 && ld >= C, "rxb_softmax_ce: bad sizes");
  if (B == 0) return RXB_OK;
  int rc = rxb_check_device();
  if (rc) return rc;
  RXB_PROF(as_stream(stream), PROF_HEAD);
  RXB_CUDA(launch_k(softmax_ce_kernel, dim3(B), dim3(kCeThreads), (size_t)(0), as_stream(stream), logits, ld, reinterpret_cast<const long long*>(target),
                                                             C, loss_rows, dlogits, grad_scale));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

int rxb_sgd_step(float* p, const float* grad, float* mom, int64_t n, float lr, float mu, float wd, int nesterov,
                 float grad_scale, rxb_stream_t stream) {
  using namespace rxb;
  RXB_CHECK_ARG(p && grad && mom, "rxb_sgd_step: null pointer");
  RXB_CHECK_ARG(n >= 0, "rxb_sgd_step: bad size");
  if (n == 0) return RXB_OK;
  int rc = rxb_check_device();
  if (rc) return rc;
  long long blocks = ceil_div<long long>(n, 256);
  long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  RXB_PROF(as_stream(stream), PROF_OPTIM);
  RXB_CUDA(launch_k(sgd_kernel, dim3((unsigned)blocks), dim3(256), (size_t)(0), as_stream(stream), p, grad, mom, n, lr, mu, wd, nesterov, grad_scale));
  RXB_LAUNCH_OK();
  return RXB_OK;
}

}  // extern "C"
