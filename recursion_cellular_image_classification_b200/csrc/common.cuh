// common.cuh — error plumbing and small device helpers shared by every translation unit of librxb.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <utility>
#include "../../include/rxb.h"

namespace rxb {

// thread-local last error message (rxb_last_error)
char* err_buf();
int set_error(int code, const char* fmt, ...);
extern long long g_launches;  // kernels enqueued since the last reset (gpu_launches in bench.py)

#define RXB_CHECK_ARG(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return rxb::set_error(RXB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define RXB_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return rxb::set_error(RXB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,         \
                            cudaGetErrorString(e__));                                          \
  } while (0)

// after a <<<>>> launch
#define RXB_LAUNCH_OK()                                                                        \
  do {                                                                                         \
    cudaError_t e__ = cudaGetLastError();                                                      \
    if (e__ != cudaSuccess)                                                                    \
      return rxb::set_error(RXB_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__,     \
                            cudaGetErrorString(e__));                                          \
    ++rxb::g_launches;                                                                         \
  } while (0)

// Optional per-launch timing (bench.py's kernel-family breakdown): when enabled, every launcher brackets its
// kernel with CUDA events on the launching stream; rxb_profile_collect sums them per category.
enum ProfCat { PROF_STATS = 0, PROF_LOADER, PROF_CONV_FWD, PROF_CONV_DGRAD, PROF_CONV_WGRAD, PROF_ELEMENTWISE,
               PROF_HEAD, PROF_OPTIM, PROF_TTA,
               // finer split (rxb.h): 2/3/4 are the dense layers' 1x1 kernels, 5 the remaining elementwise kernels
               PROF_CONV_FWD_3X3, PROF_CONV_DGRAD_3X3, PROF_CONV_WGRAD_3X3, PROF_CONV_OTHER, PROF_WGRAD_OTHER,
               PROF_EW_BN_APPLY, PROF_EW_FIXUP, PROF_EW_FINALIZE, PROF_NCAT };
extern bool g_prof_on;
struct ProfScope {
  cudaStream_t st;
  int slot;
  ProfScope(cudaStream_t s, int cat);
  ~ProfScope();
};
#define RXB_PROF(stream, cat) rxb::ProfScope prof_scope__((stream), (cat))

inline cudaStream_t as_stream(rxb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch.  Kernels of the training/inference executor are enqueued with
// cudaLaunchAttributeProgrammaticStreamSerialization: a kernel's CTAs may become resident and run their private
// prologue (barrier init, TMEM allocation, tensor-map prefetch) while the previous kernel of the stream drains;
// pdl_sync() is the point after which they may touch global memory (the predecessor has completed and flushed).
// EVERY kernel launched through launch_k must call pdl_sync() before its first global read or write.
extern bool g_dbg_sync;
extern bool g_pdl;   // RXB_PDL=1 enables the attribute (default off: kernels serialise as usual and pdl_sync() is a no-op)
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  memset(&attr, 0, sizeof(attr));
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
  // RXB_DBG_SYNC=1 (development): wait for THIS kernel on its stream so a device fault is reported at the launch that
  // caused it (other streams, e.g. NCCL's, keep running concurrently)
  if (e == cudaSuccess && g_dbg_sync) e = cudaStreamSynchronize(st);
  return e;
}
int num_sms();

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 16-byte load that does not pollute L1 (read-once data)
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_v4(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ float bf16_round(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
}

// BatchNorm channels whose backward cannot use the W.dW identity (it recovers sum(dy*x) as (W.dW - shift*sum dy)/scale,
// which cancels catastrophically when |gamma| is small against |beta| and is undefined for gamma == 0): for these the
// data-gradient epilogue reduces sum(dy) and sum(dy*x) directly.  Freshly initialised networks (gamma = 1, beta = 0)
// have none; trained / weight-decayed checkpoints do.  One predicate shared by the kernel that reduces and the kernel
// that consumes.
__host__ __device__ inline bool bn_degenerate(float gamma, float beta) {
  const float g = gamma < 0.f ? -gamma : gamma, b = beta < 0.f ? -beta : beta;
  return g < 1e-3f || g < 0.05f * b;
}

}  // namespace rxb
