// tensor-pipe ceiling of SS-mode tcgen05.mma (operands resident in shared memory, no loads): TFLOP/s for a given N and CTAs per SM
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../multigrid-neural-architectures_b200/csrc -I../include mma_rate.cu -o mma_rate
#include "umma_common.cuh"
#include <cstdio>
#include <cstdlib>
__global__ void __launch_bounds__(64, 1) mma_rate_kernel(int n_tile, int iters, int tmem_cols, int nstage, int distinct) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int stage_bytes = 16384 + n_tile * 128;
  for (int i = tid; i < nstage * stage_bytes / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i * 2654435761u & 0x00ff00ffu);
  if (tid == 0) { mbar_init(&done_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int s = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t a_lo = desc_lo_k_sw128(smem_u32(smem + (size_t)s * stage_bytes));
      const uint32_t b_lo = desc_lo_k_sw128(smem_u32(smem + (size_t)s * stage_bytes + 16384));
      if (leader) {
#pragma unroll
        for (int q = 0; q < 4; ++q) tc_mma_bf16_lohi(tmem_base, a_lo + q * 2, b_lo + q * 2, DESC_HI_SW128, idesc, (it | q) != 0);
      }
      if (distinct && ++s == nstage) s = 0;
    }
    if (leader) tc_commit(&done_bar);
    mbar_wait(&done_bar, 0);
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory"); }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1) mma_rate_pair_kernel(int n_tile, int iters, int tmem_cols, int nstage, int distinct) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int stage_bytes = 16384 + (n_tile / 2) * 128;
  for (int i = tid; i < nstage * stage_bytes / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i * 2654435761u & 0x00ff00ffu);
  if (tid == 0) { mbar_init(&done_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 1) {
    if (rank == 0) {
      const bool leader = elect_one();
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_tile >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int s = 0;
      for (int it = 0; it < iters; ++it) {
        const uint32_t a_lo = desc_lo_k_sw128(smem_u32(smem + (size_t)s * stage_bytes));
        const uint32_t b_lo = desc_lo_k_sw128(smem_u32(smem + (size_t)s * stage_bytes + 16384));
        if (leader) {
#pragma unroll
          for (int q = 0; q < 4; ++q) tc_mma_bf16_lohi_pair(tmem_base, a_lo + q * 2, b_lo + q * 2, DESC_HI_SW128, idesc, (it | q) != 0);
        }
        if (distinct && ++s == nstage) s = 0;
      }
      if (leader) tc_commit_pair(&done_bar);
    }
    mbar_wait_cluster(&done_bar, 0);
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory"); }
}
int main() {
  cudaFuncSetAttribute(mma_rate_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int n : {32, 64, 96, 128, 192, 256})
    for (int ctas : {1, 2}) {
      const int iters = 20000, distinct = 1;
      int cols = 32; while (cols < n) cols <<= 1;
      if (ctas * cols > 512) continue;
      const int stage = 16384 + (n / 2) * 128, nstage = (ctas == 1 ? 180 * 1024 : 90 * 1024) / stage;
      const size_t smem = (size_t)nstage * stage + 1024;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      mma_rate_pair_kernel<<<148 * ctas, 64, smem>>>(n, 100, cols, nstage, distinct);
      cudaEventRecord(e0);
      mma_rate_pair_kernel<<<148 * ctas, 64, smem>>>(n, iters, cols, nstage, distinct);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      const double flops = 2.0 * 128 * n * 64 * (double)iters * 148 * ctas;
      const double cyc_per_mma = ms * 1e-3 * 1.965e9 / (iters * 4.0 * ctas);
      printf("PAIR N=%3d ctas/SM=%d: %8.3f ms  %7.1f TFLOP/s  %6.1f cycles per pair-MMA @1965MHz (floor %d)  %s\n", n, ctas, ms, flops / ms / 1e9, cyc_per_mma, 128 * n / 512,
             err == cudaSuccess ? "" : cudaGetErrorString(err));
    }

  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 20000;
  for (int n : {32, 64, 96, 128, 192, 256})
    for (int ctas : {1, 2})
      for (int distinct : {0, 1}) {
        int cols = 32; while (cols < n) cols <<= 1;
        if (ctas * cols > 512) continue;
        const int stage = 16384 + n * 128, nstage = distinct ? (ctas == 1 ? 180 * 1024 : 90 * 1024) / stage : 1;
        const size_t smem = (size_t)nstage * stage + 1024;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        mma_rate_kernel<<<148 * ctas, 64, smem>>>(n, 100, cols, nstage, distinct);
        cudaEventRecord(e0);
        mma_rate_kernel<<<148 * ctas, 64, smem>>>(n, iters, cols, nstage, distinct);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 128 * n * 64 * (double)iters * 148 * ctas;
        const double cyc_per_mma = ms * 1e-3 * 1.965e9 / (iters * 4.0 * ctas);
        printf("N=%3d ctas/SM=%d %s: %8.3f ms  %7.1f TFLOP/s  %6.1f cycles/MMA/SM @1965MHz (floor %d)  %s\n", n, ctas, distinct ? "ring  " : "1 stage", ms, flops / ms / 1e9,
               cyc_per_mma, 128 * n / 256, err == cudaSuccess ? "" : cudaGetErrorString(err));
      }
  return 0;
}
