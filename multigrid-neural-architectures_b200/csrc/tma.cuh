// TMA (cp.async.bulk.tensor) support of the halo kernels: tensor-map construction + cache on the host, copy wrappers on the device.
//
// A halo is staged ROW BY ROW of the zero-padded slot space: slot row R = n*(H+1) + y holds the W pixels of image row (n, y)
// followed by one zero slot (x == W); rows y == H and rows outside the batch are all zero.  One 4-D tensor map per source grid,
// dims (Cp, W, H, N), box (64 channels, W+1, 1, 1), 128-byte swizzle, out-of-bounds fill = zero, produces exactly that: the copy
// engine writes the padding (x == W, y == H, n < 0, n >= N, channels >= Cp) itself.  The swizzle is a function of the ABSOLUTE
// shared-memory address, so a row may land at any 128-byte row of a 1024-aligned buffer and still be what a UMMA SW128
// descriptor reads (scratch/tma_test.cu, profiles/r2a_tma_swizzle_oob_zero_stride.txt).
// The coarser grid of ResampleConcat (SpatialUpSamplingNearest(2), models/ilsvrc/rnmg.lua:70-76) is replicated by the copy
// engine too: a 5-D map (Cp, 2, Ws, Hs, N) whose "2" dimension has stride ZERO and box (64, 2, Ws, 1, 1) yields the 2*Ws = W
// up-sampled slots of one row; row y reads coarse row y >> 1.  (The pad slot x == W is not written by that box: the kernels zero
// the pad column of their buffers once.)
#pragma once
#include <cuda.h>
#include "common.cuh"
#include <unordered_map>

struct TmapKey {
  const void* ptr; int N, H, W, Cp, kind, box_w;
  bool operator==(const TmapKey& o) const { return ptr == o.ptr && N == o.N && H == o.H && W == o.W && Cp == o.Cp && kind == o.kind && box_w == o.box_w; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = (size_t)k.ptr;
    h = h * 1000003u ^ (size_t)k.N; h = h * 1000003u ^ (size_t)k.H; h = h * 1000003u ^ (size_t)k.W;
    h = h * 1000003u ^ (size_t)k.Cp; h = h * 1000003u ^ (size_t)(k.kind * 1024 + k.box_w);
    return h;
  }
};
struct TmapCache { std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> map; };

typedef CUresult (*mg_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline mg_encode_tiled_fn mg_encode_tiled() {
  static mg_encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (mg_encode_tiled_fn)p;
  }
  return fn;
}

// kind 0: the grid at the conv's own resolution (box_w slots per row, box_w >= W: the surplus is the zero padding);
// kind 1: the coarser grid (g.H = H/2, g.W = W/2) up-sampled x2 in x by a zero-stride dimension (box = W = 2*g.W slots)
static inline int mg_tensor_map(mg_ctx* ctx, const void* ptr, int N, int H, int W, int Cp, int kind, int box_w, CUtensorMap* out) {
  if (!ctx->tmaps) ctx->tmaps = new (std::nothrow) TmapCache();
  MG_REQUIRE(ctx, ctx->tmaps != nullptr, MG_ERR_INVALID_ARG, "tensor map cache: out of memory");
  TmapCache* tc = (TmapCache*)ctx->tmaps;
  const TmapKey key{ptr, N, H, W, Cp, kind, box_w};
  auto it = tc->map.find(key);
  if (it != tc->map.end()) { *out = it->second; return MG_OK; }
  mg_encode_tiled_fn enc = mg_encode_tiled();
  MG_REQUIRE(ctx, enc != nullptr, MG_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  MG_REQUIRE(ctx, ((uintptr_t)ptr & 15) == 0 && (Cp % 8 == 0 || kind == 5), MG_ERR_INVALID_ARG, "tensor map: base %p / Cp %d not 16-byte aligned", ptr, Cp);
  MG_REQUIRE(ctx, box_w >= 1 && box_w <= 256, MG_ERR_UNSUPPORTED, "tensor map: box of %d slots", box_w);
  CUtensorMap tm;
  CUresult r;
  const cuuint64_t row = (cuuint64_t)Cp * 2;
  if (kind == 5) {
    // packed weight image as rows of 128 bytes (N = total rows, box_w = rows per copy): the pair kernels load half a stage per CTA
    cuuint64_t dims[2] = {64, (cuuint64_t)N};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_w};
    cuuint32_t es[2] = {1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (kind == 4) {
    // whole-image boxes: (H+1)*(W+1) slots per image divide the tile (7 x 7 grids: 64 slots), so ONE box of box_w images is the
    // tile -- pad column, pad row and images beyond the batch are out-of-bounds zeros
    cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {row, row * W, row * W * H};
    cuuint32_t box[4] = {64, (cuuint32_t)(W + 1), (cuuint32_t)(H + 1), (cuuint32_t)box_w};
    cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (kind == 3) {
    // output tile of the stem kernel: 16 rows x 8 pixels x 64 channels, 128-byte swizzle (TMA store; clipped at the image edge)
    cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {row, row * W, row * W * H};
    cuuint32_t box[4] = {64, 8, 16, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (kind == 2) {
    // stem (7x7 stride 2): the input patch of a 16 x 8 output tile, box_w columns x 37 rows of 8 channels (16 bytes), no swizzle
    cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {row, row * W, row * W * H};
    cuuint32_t box[4] = {8, (cuuint32_t)box_w, 37, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    MG_REQUIRE(ctx, Cp == 8, MG_ERR_INVALID_ARG, "tensor map: stem planes need Cp = 8, not %d", Cp);
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (kind == 0) {
    cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {row, row * W, row * W * H};
    cuuint32_t box[4] = {64, (cuuint32_t)box_w, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[5] = {(cuuint64_t)Cp, 2, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[4] = {0, row, row * W, row * W * H};
    cuuint32_t box[5] = {64, 2, (cuuint32_t)W, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    MG_REQUIRE(ctx, box_w == 2 * W, MG_ERR_INVALID_ARG, "tensor map: up-sampled box %d != 2 * %d", box_w, W);
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  MG_REQUIRE(ctx, r == CUDA_SUCCESS, MG_ERR_CUDA, "cuTensorMapEncodeTiled(kind %d, N %d H %d W %d Cp %d box %d) failed: %d", kind, N, H, W, Cp, box_w, (int)r);
  if (tc->map.size() > 65536) tc->map.clear();   // hosts that keep re-allocating activations: do not grow without bound
  tc->map.emplace(key, tm);
  *out = tm;
  return MG_OK;
}

#ifdef __CUDACC__
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
               "l"(tm), "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
               "l"(tm), "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// CTA-pair (cta_group::2) variants: the copy lands in the executing CTA's shared memory, the completion bytes are counted on the
// barrier at the same offset of the pair's EVEN CTA (the one that issues the MMAs): bit 24 of a shared::cluster address is the
// CTA's parity within the pair (cute: Sm100MmaPeerBitMask)
constexpr uint32_t MG_PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
               "l"(tm), "r"((uint32_t)__cvta_generic_to_shared(bar) & MG_PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
               "l"(tm), "r"((uint32_t)__cvta_generic_to_shared(bar) & MG_PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(tm), "r"((uint32_t)__cvta_generic_to_shared(bar) & MG_PEER_BIT_MASK), "r"(c0), "r"(c1)
               : "memory");
}
// slot rows [r0, r0 + nr) like tma_load_rows, without arming the barrier (the pair's even CTA arms it with the bytes of both CTAs)
__device__ __forceinline__ void tma_load_rows_pair(const CUtensorMap* tm, int up, uint32_t dst, uint64_t* bar, int c0, int r0, int nr, int W, int Hp, int lane) {
  const int Wp = W + 1;
  for (int i = lane; i < nr; i += 32) {
    const int R = r0 + i;
    int n = -1, y = 0;
    if (R >= 0) { n = R / Hp; y = R - n * Hp; }
    const uint32_t d = dst + (uint32_t)(i * Wp) * 128u;
    if (up) tma_load_5d_pair(d, tm, bar, c0, 0, 0, y >> 1, n);
    else tma_load_4d_pair(d, tm, bar, c0, 0, y, n);
  }
}

// floor(a / b) for b > 0 and any a
__device__ __forceinline__ int floordiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// whole-image mode (kind 4 map): the tile is `imgs` complete images starting at image n0
__device__ __forceinline__ void tma_load_images(const CUtensorMap* tm, uint32_t dst, uint64_t* bar, int c0, int n0, uint32_t bytes, int lane) {
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
    tma_load_4d(dst, tm, bar, c0, 0, 0, n0);
  }
}

// One warp stages slot rows [r0, r0 + nr) (64 channels from c0 of one source grid) at `dst` (row i at dst + i * Wp * 128) and
// arms `bar` with the byte count; every lane issues its own rows.  up: coarser grid through the zero-stride map (W slots per
// row; the pad slot is left alone), else W+1 slots per row.  Hp = H + 1 slot rows per image.
__device__ __forceinline__ void tma_load_rows(const CUtensorMap* tm, int up, uint32_t dst, uint64_t* bar, int c0, int r0, int nr, int W, int Hp, int lane) {
  const int Wp = W + 1;
  const uint32_t row_bytes = (uint32_t)(up ? W : Wp) * 128u;
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"((uint32_t)nr * row_bytes) : "memory");
  }
  __syncwarp();
  for (int i = lane; i < nr; i += 32) {
    const int R = r0 + i;
    int n = -1, y = 0;
    if (R >= 0) { n = R / Hp; y = R - n * Hp; }
    const uint32_t d = dst + (uint32_t)(i * Wp) * 128u;
    if (up) tma_load_5d(d, tm, bar, c0, 0, 0, y >> 1, n);     // y == H -> coarse row H/2: out of bounds -> zeros
    else tma_load_4d(d, tm, bar, c0, 0, y, n);
  }
}
#endif
