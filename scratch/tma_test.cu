// Hardware questions behind the TMA halo loaders (answers recorded in DESIGN.md):
//  (1) does cp.async.bulk.tensor with CU_TENSOR_MAP_SWIZZLE_128B swizzle by the ABSOLUTE shared-memory address, i.e. may the
//      destination be any 128-byte row of a 1024-aligned buffer and still produce the layout a UMMA SW128 descriptor reads?
//  (2) out-of-bounds fill: a box (64 ch, W+1, 1, 1) over an NHWC tensor yields the zero padding slot x == W, the zero row
//      y == H, zero rows for n < 0 / n >= N and zero channels beyond Cp.
//  (3) does cuTensorMapEncodeTiled accept a ZERO stride (dims (C, 2, Ws, Hs, N), stride 0 on the "2"): nearest-neighbour
//      up-sampling in x by the copy engine itself.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_test tma_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// rows[i] = (y, n) of buffer row i; dst row offset `off0` rows into a 1024-aligned buffer
__global__ void __launch_bounds__(128) k_rows(const __grid_constant__ CUtensorMap tm, int nrows, int Wp, int c0, const int* ys, const int* ns,
                                              int off0, __nv_bfloat16* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  const int total_rows = off0 + nrows * Wp;
  for (int i = tid; i < (total_rows + 8) * 32; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x7fc07fc0u;   // NaN fill: untouched bytes show up
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(nrows * Wp * 128)) : "memory");
    for (int i = 0; i < nrows; ++i) {
      const uint32_t dst = smem_u32(sm) + (uint32_t)(off0 + i * Wp) * 128u;
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                   "l"(&tm), "r"(smem_u32(&bar)), "r"(c0), "r"(0), "r"(ys[i]), "r"(ns[i])
                   : "memory");
    }
  }
  // wait
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
  }
  // un-swizzle by absolute row index: slot s (row index from the 1024-aligned base), 16-byte chunk v at ((v ^ (s & 7)) << 4)
  for (int i = tid; i < nrows * Wp * 64; i += 128) {
    const int s = off0 + i / 64, c = i % 64;
    const uint32_t addr = s * 128 + (((c >> 3) ^ (s & 7)) << 4) + (c & 7) * 2;
    out[i] = *reinterpret_cast<__nv_bfloat16*>(sm + addr);
  }
}

__global__ void __launch_bounds__(128) k_up(const __grid_constant__ CUtensorMap tm, int Wp2, int c0, int ys, int n, int off0, __nv_bfloat16* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  for (int i = tid; i < (off0 + Wp2 + 8) * 32; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x7fc07fc0u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(Wp2 * 128)) : "memory");
    const uint32_t dst = smem_u32(sm) + (uint32_t)off0 * 128u;
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
                 "l"(&tm), "r"(smem_u32(&bar)), "r"(c0), "r"(0), "r"(0), "r"(ys), "r"(n)
                 : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
  }
  for (int i = tid; i < Wp2 * 64; i += 128) {
    const int s = off0 + i / 64, c = i % 64;
    const uint32_t addr = s * 128 + (((c >> 3) ^ (s & 7)) << 4) + (c & 7) * 2;
    out[i] = *reinterpret_cast<__nv_bfloat16*>(sm + addr);
  }
}

static float val(int n, int y, int x, int c) { return (float)((n * 7 + y) * 16 + x) + c / 128.0f; }   // exact in bf16? keep small: checked via bf16 rounding below

int main() {
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int N = 2, H = 5, W = 7, Cp = 72, Wp = W + 1;
  std::vector<__nv_bfloat16> h((size_t)N * H * W * Cp);
  for (int n = 0; n < N; ++n) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < Cp; ++c)
    h[(((size_t)n * H + y) * W + x) * Cp + c] = __float2bfloat16((float)(((n * H + y) * W + x) % 61) + (float)(c % 64) * 0.0078125f * 0 + (float)c);
  __nv_bfloat16* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  auto ref = [&](int n, int y, int x, int c) -> float {
    if (n < 0 || n >= N || y < 0 || y >= H || x < 0 || x >= W || c >= Cp) return 0.f;
    return __bfloat162float(h[(((size_t)n * H + y) * W + x) * Cp + c]);
  };
  int fails = 0;
  // ---- (1) + (2)
  {
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)Wp, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode 4d: %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 2;
    for (int c0 : {0, 64}) for (int off0 : {0, 1, 3, 5, 13}) {
      std::vector<int> ys = {4, 5, 0, 1, -1, 2}, ns = {0, 0, 1, 1, 0, 2};
      const int nrows = (int)ys.size();
      int *dys, *dns; cudaMalloc(&dys, nrows * 4); cudaMalloc(&dns, nrows * 4);
      cudaMemcpy(dys, ys.data(), nrows * 4, cudaMemcpyHostToDevice); cudaMemcpy(dns, ns.data(), nrows * 4, cudaMemcpyHostToDevice);
      __nv_bfloat16* dout; cudaMalloc(&dout, (size_t)nrows * Wp * 64 * 2);
      k_rows<<<1, 128, 40 * 1024>>>(tm, nrows, Wp, c0, dys, dns, off0, dout);
      cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("k_rows failed: %s\n", cudaGetErrorString(e)); return 3; }
      std::vector<__nv_bfloat16> o((size_t)nrows * Wp * 64);
      cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int i = 0; i < nrows; ++i) for (int x = 0; x < Wp; ++x) for (int c = 0; c < 64; ++c) {
        const float got = __bfloat162float(o[((size_t)i * Wp + x) * 64 + c]);
        const float want = ref(ns[i], ys[i], x, c0 + c);
        if (!(got == want)) { if (bad < 5) printf("  c0 %d off0 %d row %d x %d c %d: got %f want %f\n", c0, off0, i, x, c, got, want); ++bad; }
      }
      printf("rows c0=%d off0=%d: %s (%d bad)\n", c0, off0, bad ? "FAIL" : "ok", bad);
      fails += bad != 0;
    }
  }
  // ---- (3) zero stride: up-sample x by the copy engine
  {
    const int Hs = 3, Ws = 4, Cq = 64;
    std::vector<__nv_bfloat16> hs((size_t)N * Hs * Ws * Cq);
    for (size_t i = 0; i < hs.size(); ++i) hs[i] = __float2bfloat16((float)((i / Cq) % 97) + (float)(i % Cq) * 0.f);
    __nv_bfloat16* ds; cudaMalloc(&ds, hs.size() * 2); cudaMemcpy(ds, hs.data(), hs.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[5] = {(cuuint64_t)Cq, 2, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)N};
    cuuint64_t strides[4] = {0, (cuuint64_t)Cq * 2, (cuuint64_t)Ws * Cq * 2, (cuuint64_t)Hs * Ws * Cq * 2};
    cuuint32_t box[5] = {64, 2, (cuuint32_t)Ws + 1, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, ds, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode 5d zero-stride: %d\n", (int)r);
    if (r == CUDA_SUCCESS) {
      const int Wp2 = 2 * (Ws + 1);
      __nv_bfloat16* dout; cudaMalloc(&dout, (size_t)Wp2 * 64 * 2);
      for (int off0 : {0, 3}) {
        k_up<<<1, 128, 40 * 1024>>>(tm, Wp2, 0, 1, 1, off0, dout);
        cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("k_up failed: %s\n", cudaGetErrorString(e)); return 4; }
        std::vector<__nv_bfloat16> o((size_t)Wp2 * 64);
        cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int x = 0; x < Wp2; ++x) for (int c = 0; c < 64; ++c) {
          const int xs = x >> 1;
          const float want = xs < Ws ? __bfloat162float(hs[(((size_t)1 * Hs + 1) * Ws + xs) * Cq + c]) : 0.f;
          const float got = __bfloat162float(o[(size_t)x * 64 + c]);
          if (!(got == want)) { if (bad < 5) printf("  up off0 %d x %d c %d: got %f want %f\n", off0, x, c, got, want); ++bad; }
        }
        printf("up off0=%d: %s (%d bad)\n", off0, bad ? "FAIL" : "ok", bad);
        fails += bad != 0;
      }
    }
  }
  printf("tma_test: %s\n", fails ? "FAILURES" : "all ok");
  return 0;
}
