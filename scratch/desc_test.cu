// Does a K-major SWIZZLE_128B UMMA descriptor work when its start address is offset by j rows of 128 B
// (not 1024-aligned)?  A = [rows][64] bf16 pattern in smem written with absolute-address swizzle,
// B = 64x64 identity (K-major), D[m][n] must equal A[j+m][n] for the first 16 K elements ... we use K=64 (4 MMAs).
#include "../multigrid-neural-architectures_b200/csrc/umma_common.cuh"
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint64_t desc_k(uint32_t saddr, uint32_t sbo, uint32_t base_off) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
}

constexpr int ROWS = 160;
// variant with explicit base_offset = (start>>7)&7
__global__ void __launch_bounds__(128) k2(float* out, int j, int mode, int use_bo) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar; __shared__ uint32_t tb;
  uint8_t* A = sm; uint8_t* B = sm + ROWS * 128;
  const int tid = threadIdx.x;
  for (int i = tid; i < ROWS * 64; i += 128) {
    int r = i / 64, c = i % 64;
    float val = (mode == 0) ? (float)r : (float)c;
    uint32_t addr = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(A + addr) = __float2bfloat16(val);
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    int n = i / 64, c = i % 64;
    uint32_t addr = n * 128 + (((c >> 3) ^ (n & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(B + addr) = __float2bfloat16(n == c ? 1.f : 0.f);
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tb;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t a_addr = smem_u32(A) + j * 128, b_addr = smem_u32(B);
    uint32_t bo = use_bo ? ((a_addr >> 7) & 7) : 0;
    for (int q = 0; q < 4; ++q)
      tc_mma_bf16(tmem, desc_k(a_addr + q * 32, 1024, bo), desc_k(b_addr + q * 32, 1024, 0), idesc, q != 0);
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = tid >> 5;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t v[16];
    ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc_wait_ld();
    for (int x = 0; x < 16; ++x) out[tid * 64 + c0 + x] = __uint_as_float(v[x]);
  }
  tc_fence_before(); __syncthreads();
  if (tid < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory"); }
}

int main() {
  float* d; cudaMalloc(&d, 128 * 64 * 4);
  std::vector<float> h(128 * 64);
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int use_bo = 0; use_bo < 2; ++use_bo)
    for (int j : {0, 1, 2, 3, 7, 8, 9, 17}) {
      int bad_r = 0, bad_c = 0;
      for (int mode = 0; mode < 2; ++mode) {
        k2<<<1, 128, 64 * 1024>>>(d, j, mode, use_bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("j=%d error %s\n", j, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            float want = mode == 0 ? (float)(j + m) : (float)n;
            if (h[m * 64 + n] != want) (mode == 0 ? bad_r : bad_c)++;
          }
      }
      printf("base_offset %s  row shift j=%2d : row-pattern mismatches %5d, column-pattern mismatches %5d\n", use_bo ? "(addr>>7)&7" : "0          ", j, bad_r, bad_c);
    }
  return 0;
}
