// umma_probe.cu — development probe (not part of the library): (1) issue rate of tcgen05.mma kind::f16 M=128 for several N,
// (2) correctness of K-major SWIZZLE_128B A-operand descriptors whose start address is NOT 1024-byte aligned (row shifts of
// 128 B) and whose stride-byte-offset is not a multiple of 1024 (used by the full-halo 3x3 convolution tile).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu -I../../recursion_cellular_image_classification_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace rxb;

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int a_rows_step, long long* out_cycles, int Mrows = 128) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(&tmem_base_s);
  ptx::fence_proxy_async_smem();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(Mrows, N, 0, 0);
    const uint32_t a_addr = ptx::smem_u32(smem), b_addr = ptx::smem_u32(smem + 64 * 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t da = ptx::make_smem_desc(a_addr + ((i * a_rows_step) & 63) * 128 + (i & 3) * 32, 16, 1024, ptx::kSwizzle128B);
      const uint64_t db = ptx::make_smem_desc(b_addr + (i & 3) * 32, 16, 1024, ptx::kSwizzle128B);
      ptx::umma_bf16_ss(tmem + (i & 1) * 256, da, db, idesc, i > 1 ? 1u : 0u);
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

// A: rows x 64 bf16 in smem (SW128 by absolute row), B: 32 x 64.  D[i, n] = sum_k A[r0 + (i/8)*G + i%8][k] * B[n][k],
// where G = sbo/128 rows.
__global__ void __launch_bounds__(128, 1) shift_kernel(const __nv_bfloat16* A, int a_rows, const __nv_bfloat16* Bm, int r0,
                                                        int sbo, int use_base_off, float* D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 64 * 1024;
  for (int i = threadIdx.x; i < a_rows * 8; i += 128) {
    const int row = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sA + row * 128 + ((c ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + row * 64 + c * 8);
  }
  for (int i = threadIdx.x; i < 32 * 8; i += 128) {
    const int row = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sB + row * 128 + ((c ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(Bm + row * 64 + c * 8);
  }
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<32>(&tmem_base_s);
  ptx::fence_proxy_async_smem();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, 32, 0, 0);
    const uint32_t a_addr = ptx::smem_u32(sA) + r0 * 128, b_addr = ptx::smem_u32(sB);
    for (int k = 0; k < 4; ++k) {
      uint64_t da = ptx::make_smem_desc(a_addr + k * 32, 16, sbo, ptx::kSwizzle128B);
      if (use_base_off) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
      const uint64_t db = ptx::make_smem_desc(b_addr + k * 32, 16, 1024, ptx::kSwizzle128B);
      ptx::umma_bf16_ss(tmem, da, db, idesc, k > 0 ? 1u : 0u);
    }
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0, 2);
  }
  __syncthreads();
  ptx::tcgen05_fence_after();
  uint32_t r[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ptx::tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16), r);
  ptx::tmem_ld_wait();
  for (int n = 0; n < 32; ++n) D[(warp * 32 + lane) * 32 + n] = __uint_as_float(r[n]);
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tcgen05_fence_after(); ptx::tmem_dealloc<32>(tmem); }
}

// (3) MN-major SWIZZLE_64B B operand (the wgrad dOut tile: rows = pixels (K), 32 channels (N) per 64-byte row) whose
// start is an arbitrary pixel row and whose 8-pixel groups lie `sbo` bytes apart:
//   D[m, n] = sum_{k<64} A[m][k] * Bm[r0 + (k/8)*G + k%8][n],  G = sbo/64 rows.   A: K-major SW128 [128][64].
__global__ void __launch_bounds__(128, 1) mn_shift_kernel(const __nv_bfloat16* A, const __nv_bfloat16* Bm, int b_rows, int r0,
                                                           int sbo, float* D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 32 * 1024;
  for (int i = threadIdx.x; i < 128 * 8; i += 128) {
    const int row = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sA + row * 128 + ((c ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + row * 64 + c * 8);
  }
  for (int i = threadIdx.x; i < b_rows * 4; i += 128) {
    const int row = i >> 2, c = i & 3;
    *reinterpret_cast<uint4*>(sB + row * 64 + ((c ^ ((row >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(Bm + row * 32 + c * 8);
  }
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<32>(&tmem_base_s);
  ptx::fence_proxy_async_smem();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, 32, 0, 1);   // A K-major, B MN-major
    const uint32_t a_addr = ptx::smem_u32(sA), b_addr = ptx::smem_u32(sB) + r0 * 64;
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = ptx::make_smem_desc(a_addr + k * 32, 16, 1024, ptx::kSwizzle128B);
      const uint64_t db = ptx::make_smem_desc(b_addr + k * 2 * sbo, 8192, sbo, ptx::kSwizzle64B);
      ptx::umma_bf16_ss(tmem, da, db, idesc, k > 0 ? 1u : 0u);
    }
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0, 2);
  }
  __syncthreads();
  ptx::tcgen05_fence_after();
  uint32_t r[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ptx::tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16), r);
  ptx::tmem_ld_wait();
  for (int n = 0; n < 32; ++n) D[(warp * 32 + lane) * 32 + n] = __uint_as_float(r[n]);
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tcgen05_fence_after(); ptx::tmem_dealloc<32>(tmem); }
}

// (4) K-major SWIZZLE_64B A operand (dgrad 3x3: 32 channels per 64-byte pixel row) with unaligned start / group stride:
//   D[i, n] = sum_{k<32} A[r0 + (i/8)*G + i%8][k] * B[n][k],  G = sbo/64 rows.
__global__ void __launch_bounds__(128, 1) k64_shift_kernel(const __nv_bfloat16* A, int a_rows, const __nv_bfloat16* Bm, int r0,
                                                            int sbo, float* D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 64 * 1024;
  for (int i = threadIdx.x; i < a_rows * 4; i += 128) {
    const int row = i >> 2, c = i & 3;
    *reinterpret_cast<uint4*>(sA + row * 64 + ((c ^ ((row >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(A + row * 32 + c * 8);
  }
  for (int i = threadIdx.x; i < 32 * 4; i += 128) {
    const int row = i >> 2, c = i & 3;
    *reinterpret_cast<uint4*>(sB + row * 64 + ((c ^ ((row >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(Bm + row * 32 + c * 8);
  }
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<32>(&tmem_base_s);
  ptx::fence_proxy_async_smem();
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, 32, 0, 0);
    const uint32_t a_addr = ptx::smem_u32(sA) + r0 * 64, b_addr = ptx::smem_u32(sB);
    for (int k = 0; k < 2; ++k) {
      const uint64_t da = ptx::make_smem_desc(a_addr + k * 32, 16, sbo, ptx::kSwizzle64B);
      const uint64_t db = ptx::make_smem_desc(b_addr + k * 32, 16, 512, ptx::kSwizzle64B);
      ptx::umma_bf16_ss(tmem, da, db, idesc, k > 0 ? 1u : 0u);
    }
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0, 2);
  }
  __syncthreads();
  ptx::tcgen05_fence_after();
  uint32_t r[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ptx::tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16), r);
  ptx::tmem_ld_wait();
  for (int n = 0; n < 32; ++n) D[(warp * 32 + lane) * 32 + n] = __uint_as_float(r[n]);
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tcgen05_fence_after(); ptx::tmem_dealloc<32>(tmem); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

int main() {
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  long long* dc; CK(cudaMalloc(&dc, 8));
  const int Ns[] = {16, 32, 64, 96, 128, 192, 256};
  for (int step = 0; step <= 1; ++step)
    for (int N : Ns) {
      const int iters = 2048;
      rate_kernel<<<148, 128, 100 * 1024>>>(N, iters, step, dc);
      CK(cudaDeviceSynchronize());
      rate_kernel<<<148, 128, 100 * 1024>>>(N, iters, step, dc);
      CK(cudaDeviceSynchronize());
      long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
      printf("rate: M=128 N=%3d K=16 a_row_step=%d : %.1f cycles/MMA (floor %d)\n", N, step, (double)c / iters, N / 2);
    }
  for (int N : {64, 128, 256}) {
    const int iters = 2048;
    rate_kernel<<<148, 128, 100 * 1024>>>(N, iters, 0, dc, 64);
    CK(cudaDeviceSynchronize());
    rate_kernel<<<148, 128, 100 * 1024>>>(N, iters, 0, dc, 64);
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
    printf("rate: M= 64 N=%3d K=16 : %.1f cycles/MMA\n", N, (double)c / iters);
  }
  // ---- shift correctness
  const int a_rows = 400;
  std::vector<__nv_bfloat16> hA(a_rows * 64), hB(32 * 64);
  std::vector<float> fA(a_rows * 64), fB(32 * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 17 - 8) / 8.f; hA[i] = __float2bfloat16(v); fA[i] = v; }
  for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 13 - 6) / 4.f; hB[i] = __float2bfloat16(v); fB[i] = v; }
  __nv_bfloat16 *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * 32 * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  std::vector<float> hD(128 * 32);
  const int sbos[] = {1024, 1280, 2304};
  for (int sbo : sbos)
    for (int bo = 0; bo <= 1; ++bo) {
      printf("shift: sbo=%d base_off=%d : ", sbo, bo);
      for (int r0 = 0; r0 < 12; ++r0) {
        shift_kernel<<<1, 128, 100 * 1024>>>(dA, a_rows, dB, r0, sbo, bo, dD);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        const int G = sbo / 128;
        for (int i = 0; i < 128; ++i)
          for (int n = 0; n < 32; ++n) {
            const int row = r0 + (i / 8) * G + (i % 8);
            double ref = 0;
            for (int k = 0; k < 64; ++k) ref += (double)fA[row * 64 + k] * fB[n * 64 + k];
            maxerr = fmax(maxerr, fabs(ref - hD[i * 32 + n]));
          }
        printf("%s", maxerr < 1e-3 ? "ok " : "BAD ");
      }
      printf("\n");
    }
  // ---- (3) MN-major SW64 B with unaligned pixel-row start and custom group stride
  {
    CK(cudaFuncSetAttribute(mn_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(k64_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    const int b_rows = 400;
    std::vector<__nv_bfloat16> hA2(128 * 64), hB2(b_rows * 32);
    std::vector<float> fA2(128 * 64), fB2(b_rows * 32);
    for (size_t i = 0; i < hA2.size(); ++i) { float v = (rand() % 17 - 8) / 8.f; hA2[i] = __float2bfloat16(v); fA2[i] = v; }
    for (size_t i = 0; i < hB2.size(); ++i) { float v = (rand() % 13 - 6) / 4.f; hB2[i] = __float2bfloat16(v); fB2[i] = v; }
    __nv_bfloat16 *dA2, *dB2;
    CK(cudaMalloc(&dA2, hA2.size() * 2)); CK(cudaMalloc(&dB2, hB2.size() * 2));
    CK(cudaMemcpy(dA2, hA2.data(), hA2.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB2, hB2.data(), hB2.size() * 2, cudaMemcpyHostToDevice));
    const int sbos2[] = {512, 640, 1152};
    for (int sbo : sbos2) {
      printf("mn_shift (B MN-major SW64): sbo=%d : ", sbo);
      for (int r0 = 0; r0 < 12; ++r0) {
        mn_shift_kernel<<<1, 128, 100 * 1024>>>(dA2, dB2, b_rows, r0, sbo, dD);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        const int G = sbo / 64;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 32; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) ref += (double)fA2[m * 64 + k] * fB2[(r0 + (k / 8) * G + (k % 8)) * 32 + n];
            maxerr = fmax(maxerr, fabs(ref - hD[m * 32 + n]));
          }
        printf("%s", maxerr < 1e-3 ? "ok " : "BAD ");
      }
      printf("\n");
    }
    // ---- (4) K-major SW64 A with unaligned start: reuse hB2 as A rows [b_rows][32], B = first 32 rows of hA2 viewed [32][32]
    std::vector<__nv_bfloat16> hB3(32 * 32);
    std::vector<float> fB3(32 * 32);
    for (size_t i = 0; i < hB3.size(); ++i) { float v = (rand() % 11 - 5) / 4.f; hB3[i] = __float2bfloat16(v); fB3[i] = v; }
    __nv_bfloat16* dB3;
    CK(cudaMalloc(&dB3, hB3.size() * 2));
    CK(cudaMemcpy(dB3, hB3.data(), hB3.size() * 2, cudaMemcpyHostToDevice));
    for (int sbo : sbos2) {
      printf("k64_shift (A K-major SW64): sbo=%d : ", sbo);
      for (int r0 = 0; r0 < 12; ++r0) {
        k64_shift_kernel<<<1, 128, 100 * 1024>>>(dB2, b_rows, dB3, r0, sbo, dD);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        const int G = sbo / 64;
        for (int i = 0; i < 128; ++i)
          for (int n = 0; n < 32; ++n) {
            const int row = r0 + (i / 8) * G + (i % 8);
            double ref = 0;
            for (int k = 0; k < 32; ++k) ref += (double)fB2[row * 32 + k] * fB3[n * 32 + k];
            maxerr = fmax(maxerr, fabs(ref - hD[i * 32 + n]));
          }
        printf("%s", maxerr < 1e-3 ? "ok " : "BAD ");
      }
      printf("\n");
    }
  }
  return 0;
}
