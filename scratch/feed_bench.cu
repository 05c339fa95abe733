// How fast can shared memory be filled from L2-resident data?  cp.async 16B (LDGSTS) vs cp.async.bulk.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int DEPTH>
__global__ void __launch_bounds__(128) ldgsts_kernel(const uint4* __restrict__ src, size_t n_vec, int iters, int stage_vecs) {
  extern __shared__ uint4 sm[];
  const int tid = threadIdx.x;
  size_t base = ((size_t)blockIdx.x * 7919) % (n_vec - (size_t)stage_vecs * (iters + 1));
  for (int it = 0; it < iters + DEPTH; ++it) {
    if (it < iters) {
      const uint4* s = src + base + (size_t)it * stage_vecs;
      uint32_t dst = smem_u32(sm + (size_t)(it % (DEPTH + 1)) * stage_vecs);
      for (int i = tid; i < stage_vecs; i += 128)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 16), "l"(s + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH) : "memory");
  }
}

__global__ void __launch_bounds__(128) bulk_kernel(const uint4* __restrict__ src, size_t n_vec, int iters, int stage_vecs, int depth) {
  extern __shared__ __align__(128) uint4 sm[];
  __shared__ uint64_t bar[8];
  const int tid = threadIdx.x;
  if (tid == 0) for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
  __syncthreads();
  size_t base = ((size_t)blockIdx.x * 7919) % (n_vec - (size_t)stage_vecs * (iters + 1));
  if (tid == 0) {
    uint32_t bytes = stage_vecs * 16;
    for (int it = 0; it < iters + depth; ++it) {
      int s = it % depth;
      if (it >= depth) {  // wait for the copy issued `depth` iterations ago
        uint32_t par = ((it / depth) - 1) & 1, done = 0;
        while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(par) : "memory");
      }
      if (it < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + (size_t)s * stage_vecs)),
                     "l"(src + base + (size_t)it * stage_vecs), "r"(bytes), "r"(smem_u32(&bar[s])) : "memory");
      }
    }
  }
}

int main(int argc, char** argv) {
  size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 64) << 20;   // working set (MB): 64 MB stays in L2
  size_t n_vec = bytes / 16;
  uint4* d; cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
  int sms = 148;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 400;
  for (int stage_kb : {16, 24}) {
    int stage_vecs = stage_kb * 1024 / 16;
    for (int ctas : {1, 2, 4}) {
      // LDGSTS, depth 1 and 3
      for (int depth : {1, 3}) {
        int smem = (depth + 1) * stage_kb * 1024;
        if (smem * ctas > 220 * 1024) continue;
        auto k = depth == 1 ? ldgsts_kernel<1> : ldgsts_kernel<3>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k<<<sms * ctas, 128, smem>>>(d, n_vec, iters, stage_vecs);
        cudaEventRecord(e0);
        k<<<sms * ctas, 128, smem>>>(d, n_vec, iters, stage_vecs);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("ldgsts  stage %2d KB  ctas/SM %d  depth %d : %7.2f TB/s  (%s)\n", stage_kb, ctas, depth,
               (double)sms * ctas * iters * stage_kb * 1024 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
      }
      for (int depth : {2, 4}) {
        int smem = depth * stage_kb * 1024;
        if (smem * ctas > 220 * 1024) continue;
        cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        bulk_kernel<<<sms * ctas, 128, smem>>>(d, n_vec, iters, stage_vecs, depth);
        cudaEventRecord(e0);
        bulk_kernel<<<sms * ctas, 128, smem>>>(d, n_vec, iters, stage_vecs, depth);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("bulk    stage %2d KB  ctas/SM %d  depth %d : %7.2f TB/s  (%s)\n", stage_kb, ctas, depth,
               (double)sms * ctas * iters * stage_kb * 1024 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}
