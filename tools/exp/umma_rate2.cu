// umma_rate2.cu — development probe (not part of the library): issue rate of tcgen05.mma kind::f16 M=128 for the operand
// layouts the conv kernels use (K-major / MN-major, 128 B / 64 B swizzle, halo strides).  Values are irrelevant.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate2 umma_rate2.cu -I../../recursion_cellular_image_classification_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace rxb;

struct Cfg { int N, a_mn, b_mn; uint32_t swz_a, lbo_a, sbo_a, kstep_a, swz_b, lbo_b, sbo_b, kstep_b; int ksteps; const char* name; };

template <int KS>
__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, int iters, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(&tmem_base_s);
  ptx::fence_proxy_async_smem();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, c.N, c.a_mn, c.b_mn);
    const uint32_t a_addr = ptx::smem_u32(smem), b_addr = ptx::smem_u32(smem + 96 * 1024);
    const uint64_t da0 = ptx::make_smem_desc(a_addr, c.lbo_a, c.sbo_a, c.swz_a);
    const uint64_t db0 = ptx::make_smem_desc(b_addr, c.lbo_b, c.sbo_b, c.swz_b);
    const uint32_t a_lo = ptx::desc_lo(da0), a_hi = ptx::desc_hi(da0), b_lo = ptx::desc_lo(db0), b_hi = ptx::desc_hi(db0);
    const uint32_t ka = c.kstep_a >> 4, kb = c.kstep_b >> 4;
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ks = j % KS;
        ptx::umma_bf16_ss_parts(tmem + (j & 1) * 256, a_lo + ks * ka, a_hi, b_lo + ks * kb, b_hi, idesc, 1u);
      }
    }
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0, 1);
    long long t1 = clock64();
    if (blockIdx.x == 0) out_cycles[0] = t1 - t0;
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tcgen05_fence_after(); ptx::tmem_dealloc<512>(tmem); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

int main() {
  const uint32_t S128 = ptx::kSwizzle128B, S64 = ptx::kSwizzle64B, S0 = ptx::kSwizzleNone;
  const Cfg cfgs[] = {
    {128, 0, 0, S128, 16, 1024, 32, S128, 16, 1024, 32, 4, "K-major A/B SW128 N=128 (1x1 forward / dgrad main loop)"},
    {64, 0, 0, S128, 16, 1024, 32, S128, 16, 1024, 32, 4, "K-major A/B SW128 N=64"},
    {256, 0, 0, S128, 16, 1024, 32, S128, 16, 1024, 32, 4, "K-major A/B SW128 N=256"},
    {128, 1, 1, S128, 16384, 1024, 2048, S128, 16384, 1024, 2048, 8, "MN-major A/B SW128 N=128 (Gram statistics; fused 1x1 wgrad)"},
    {64, 1, 1, S128, 16384, 1024, 2048, S128, 16384, 1024, 2048, 8, "MN-major A/B SW128 N=64"},
    {256, 1, 1, S128, 16384, 1024, 2048, S128, 16384, 1024, 2048, 8, "MN-major A/B SW128 N=256"},
    {16, 1, 0, S128, 16384, 1024, 2048, S0, 128, 256, 0, 8, "MN-major A SW128, K-major B no swizzle N=16 (column-sum statistics)"},
    {96, 1, 1, S128, 16384, 1024, 2048, S64, 64, 640, 1280, 8, "MN-major A SW128, MN-major B SW64 halo N=96 (3x3 wgrad)"},
    {128, 0, 0, S64, 16, 640, 32, S64, 16, 512, 32, 2, "K-major A SW64 halo, K-major B SW64 N=128 (3x3 dgrad, BK=32)"},
    {128, 0, 0, S64, 16, 512, 32, S64, 16, 512, 32, 2, "K-major A SW64 aligned, K-major B SW64 N=128"},
    {96, 0, 0, S128, 16, 2304, 32, S128, 16, 1024, 32, 4, "K-major A SW128 halo (18-row stride), K-major B SW128 N=96 (3x3 forward x-merged)"},
    {64, 0, 0, S64, 16, 704, 32, S64, 16, 512, 32, 2, "K-major A SW64 halo, K-major B SW64 N=64 (stem, BK=32)"},
    {128, 1, 0, S128, 16384, 1024, 2048, S128, 16, 1024, 32, 8, "MN-major A SW128, K-major B SW128 N=128"},
    {128, 0, 1, S128, 16, 1024, 32, S128, 16384, 1024, 2048, 4, "K-major A SW128, MN-major B SW128 N=128"},
  };
  CK(cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(rate_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(rate_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long* dc; CK(cudaMalloc(&dc, 8));
  for (const Cfg& c : cfgs) {
    const int iters = 4096;
    for (int r = 0; r < 2; ++r) {
      if (c.ksteps == 2) rate_kernel<2><<<148, 128, 200 * 1024>>>(c, iters, dc);
      else if (c.ksteps == 4) rate_kernel<4><<<148, 128, 200 * 1024>>>(c, iters, dc);
      else rate_kernel<8><<<148, 128, 200 * 1024>>>(c, iters, dc);
      CK(cudaDeviceSynchronize());
    }
    long long cyc; CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
    printf("rate: %-90s : %6.1f cycles/MMA (floor %d)\n", c.name, (double)cyc / iters, c.N / 2);
  }
  return 0;
}
