// tcgen05 weight gradient of the multigrid convolution (accGradParameters of
// cudnn.SpatialConvolution, models/ilsvrc/rnmg.lua:26,36) for sm_100a.
//
//   dW[(tap, ci)][co] += gscale * sum_pixels  gather(x)[pixel][(tap, ci)] * g[pixel][co]
//
// GEMM view: M = 128 consecutive gathered channels of the K order of the forward kernel (16
// k-vectors: (tap, segment, c8)), N = Cout (<= 256 per launch column), reduction over pixels.
// Both operands are "MN-major" for the tensor core: the shared-memory images are exactly the ones
// the forward kernel builds (rows = pixels, 128 bytes = 64 channels, 128B swizzle), only the
// descriptors differ.  The cross-scale gather (pooled companion | same | up-sampled coarser grid)
// is recomputed here with the same cp.async loader -- the concatenated tensor is not stored for
// backward either.  The pixel range is split across CTAs; partial sums are added to dW with
// fp32 reductions (red.global.add), so dW accumulates like Torch's accGradParameters.
#include "common.cuh"
#include "umma_common.cuh"
#include "tma.cuh"
#include <algorithm>

namespace {

constexpr int WM = 128;           // gathered channels per CTA = UMMA M = 16 k-vectors
constexpr int PIX = 64;           // pixels per pipeline stage = 4 UMMA K steps
constexpr int A_IMG = PIX * 128;  // one 64-channel operand image of a stage
constexpr int N_PROD = 128;
constexpr int W_THREADS = 160;    // 4 producer/epilogue warps + 1 MMA warp
constexpr int W_MAX_STAGES = 8;
constexpr int SMEM_WGRAD = 232448 - 2048;   // 227 KB per CTA minus the kernel's static shared memory

struct WParams {
  USeg seg[MG_MAX_SEG];
  int seg_C[MG_MAX_SEG], seg_cbegin[MG_MAX_SEG];
  int n_seg;
  int k, stride, pad;
  int H, W, Ho, Wo;
  int64_t M;             // N * Ho * Wo (pixels of g)
  int kv_per_tap, nkv;
  const __nv_bfloat16* g;
  int g_cp;              // channel pitch of g
  int Cout, Ccat;
  int n_tile;            // UMMA N (multiple of 16)
  int n_blk;             // 64-channel images of g per stage = ceil(n_tile / 64)
  float* partial;        // [splits][k*k][Cout][Ccat] fp32 partial sums (context workspace), reduced by wgrad_reduce_kernel
  float gscale;
  int64_t pix_per_cta;   // multiple of PIX
  int stages, lag, tmem_cols;
};

__host__ __device__ constexpr uint32_t idesc_bf16_m128_mn(int n) {
  // D fp32, A/B bf16, both MN-major (bits 15, 16), M = 128
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(WM >> 4) << 24);
}

__global__ void __launch_bounds__(W_THREADS, 2) umma_wgrad_kernel(const __grid_constant__ WParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[W_MAX_STAGES], empty_bar[W_MAX_STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ USeg s_seg[MG_MAX_SEG];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  const int stage_bytes = 2 * A_IMG + p.n_blk * A_IMG;
  const int mt = blockIdx.x;                 // which 16 k-vectors
  const int nt = blockIdx.y;                 // which column tile of Cout
  const int64_t pix0 = (int64_t)blockIdx.z * p.pix_per_cta;
  const int64_t pix1 = min(p.M, pix0 + p.pix_per_cta);
  const int n_iters = (int)((pix1 - pix0 + PIX - 1) / PIX);

  if (tid < p.n_seg) s_seg[tid] = p.seg[tid];
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], N_PROD); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 4) {
    // ---- this thread's fixed k-vector of the gathered operand --------------------------------
    const int kvl = tid & 15;                 // 0..15 within the M tile
    const int pg = tid >> 4;                  // pixel lane 0..7: pixels pg, pg+8, ... of each stage
    const int j = mt * 16 + kvl;
    const bool kv_ok = j < p.nkv;
    int tap = 0, r = 0, sg = 0;
    if (kv_ok) {
      tap = j / p.kv_per_tap; r = j - tap * p.kv_per_tap;
      while (sg + 1 < p.n_seg && r >= s_seg[sg + 1].kv_begin) ++sg;
    }
    const USeg sgm = s_seg[sg];
    const int c8 = r - sgm.kv_begin;
    const int dy = tap / p.k - p.pad, dx = tap % p.k - p.pad;
    const int blk = kvl >> 3, v = kvl & 7;
    // running (n, oy, ox) of the next pixel this thread gathers; advances by 8 per copy
    int64_t m = pix0 + pg;
    int ox = (int)(m % p.Wo); int64_t q = m / p.Wo;
    int oy = (int)(q % p.Ho); int n = (int)(q / p.Ho);
    // ---- g operand: chunk (16 B) gc of pixel gp, 8 pixels apart -------------------------------
    const int chunks = p.n_blk * 8;           // 16-byte chunks per pixel row across the g images
    const int g_per_thread = (PIX * chunks) / N_PROD;
    const int L = p.lag;

    Ring rs(S), rpub(S);
    for (int it = 0; it < n_iters + L; ++it, rs.next()) {
      if (it < n_iters) {
        const int s = rs.idx;
        if (it >= S) mbar_wait(&empty_bar[s], rs.phase ^ 1u);
        uint8_t* st = smem + (size_t)s * stage_bytes;
        const uint32_t a_dst = smem_u32(st) + blk * A_IMG;
        const int64_t mbase = pix0 + (int64_t)it * PIX;
        // the kernel is issue bound on the stem (ncu: 66 % of the issue slots): the source address of the gathered operand
        // is carried along the pixel walk (8 output pixels in x per copy = 8 * stride input pixels) and recomputed only when
        // the walk wraps to the next output row
        const char* abase = reinterpret_cast<const char*>(sgm.ptr) + c8 * 16;
        const uint32_t apitch = (uint32_t)sgm.Cp * 2u;
#pragma unroll
        for (int i = 0; i < PIX / 8; ++i) {
          const int prow = i * 8 + pg;
          const int iy = oy * p.stride + dy, ix = ox * p.stride + dx;
          const bool ok = kv_ok && (mbase + prow) < pix1 && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
          const int64_t pixel = ((int64_t)n * sgm.Hs + (iy >> sgm.shift)) * sgm.Ws + (ix >> sgm.shift);
          const char* src = ok ? abase + pixel * apitch : reinterpret_cast<const char*>(sgm.ptr);
          cp_async16(a_dst + prow * 128 + ((v ^ (prow & 7)) << 4), src, ok ? 16u : 0u);
          ox += 8;
          while (ox >= p.Wo) { ox -= p.Wo; if (++oy == p.Ho) { oy = 0; ++n; } }
        }
        const uint32_t g_dst = smem_u32(st) + 2 * A_IMG;
        if (N_PROD % chunks == 0) {   // chunk index of this thread is the same in every round: no division per copy
          const int ch = tid % chunks, gb = ch >> 3, gv = ch & 7, rows_per_round = N_PROD / chunks;
          const int c0 = nt * p.n_tile + ch * 8;
          const bool ch_ok = ch * 8 < p.n_tile && c0 < p.g_cp;
          int prow = tid / chunks;
          const __nv_bfloat16* src0 = p.g + (size_t)(mbase + prow) * p.g_cp + c0;
          for (int i = 0; i < g_per_thread; ++i, prow += rows_per_round, src0 += (size_t)rows_per_round * p.g_cp) {
            const bool ok = ch_ok && (mbase + prow) < pix1;
            cp_async16(g_dst + gb * A_IMG + prow * 128 + ((gv ^ (prow & 7)) << 4), ok ? src0 : p.g, ok ? 16u : 0u);
          }
        } else {
          for (int i = 0; i < g_per_thread; ++i) {
            const int idx = i * N_PROD + tid;       // consecutive threads -> consecutive chunks of one pixel row
            const int prow = idx / chunks, ch = idx - prow * chunks;
            const int gb = ch >> 3, gv = ch & 7;
            const int c0 = nt * p.n_tile + ch * 8;  // first output channel of the chunk
            const bool ok = (mbase + prow) < pix1 && ch * 8 < p.n_tile && c0 < p.g_cp;
            const __nv_bfloat16* src = ok ? p.g + (size_t)(mbase + prow) * p.g_cp + c0 : p.g;
            cp_async16(g_dst + gb * A_IMG + prow * 128 + ((gv ^ (prow & 7)) << 4), src, ok ? 16u : 0u);
          }
        }
      }
      cp_async_commit();
      if (it >= L) {
        cp_async_wait_dyn(L);
        fence_proxy_async();
        mbar_arrive(&full_bar[rpub.idx]);
        rpub.next();
      }
    }
    // ---- epilogue: TMEM -> fp32 reductions into dW --------------------------------------------
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;           // row of D = gathered channel (k-vector row/8, element row%8)
    const int jr = mt * 16 + (row >> 3), e = row & 7;
    int rtap = 0, rr = 0, rsg = 0;
    bool row_ok = jr < p.nkv;
    if (row_ok) {
      rtap = jr / p.kv_per_tap; rr = jr - rtap * p.kv_per_tap;
      while (rsg + 1 < p.n_seg && rr >= s_seg[rsg + 1].kv_begin) ++rsg;
      const int cl = (rr - s_seg[rsg].kv_begin) * 8 + e;
      row_ok = cl < p.seg_C[rsg];
      rr = p.seg_cbegin[rsg] + cl;              // logical concat channel ci
    }
    const int KK = p.k * p.k;
    // plain stores of this CTA's partial sums; the reduction over the pixel splits is a separate pass in a fixed order
    // (wgrad_reduce_kernel): no floating-point atomics, the gradient is bit-identical from run to run
    float* prow = p.partial + (((size_t)blockIdx.z * KK + rtap) * p.Cout) * p.Ccat + rr;
    for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
      uint32_t acc[16];
      tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, acc);
      tc_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          const int co = nt * p.n_tile + c0 + x;
          if (co < p.Cout) prow[(size_t)co * p.Ccat] = __uint_as_float(acc[x]);
        }
      }
    }
    tc_fence_before();
  } else {
    // ---- MMA issuer: the whole warp runs the loop, the elected lane issues (see elect_one) -----------
    {
      const bool leader = elect_one();
      const uint32_t idesc = idesc_bf16_m128_mn(p.n_tile);
      Ring rs(S);
      for (int it = 0; it < n_iters; ++it, rs.next()) {
        const int s = rs.idx;
        mbar_wait(&full_bar[s], rs.phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t a_lo = desc_lo_mn_sw128(a_addr, A_IMG), b_lo = desc_lo_mn_sw128(a_addr + 2 * A_IMG, A_IMG);
        if (leader) {
#pragma unroll
          for (int q = 0; q < PIX / 16; ++q)   // 16 pixels (K) per UMMA: 16 rows of 128 bytes = descriptor address + 128
            tc_mma_bf16_lohi(tmem_base, a_lo + q * 128, b_lo + q * 128, DESC_HI_SW128, idesc, (it | q) != 0);
          tc_commit(&empty_bar[s]);
        }
      }
      if (leader) tc_commit(&tmem_full_bar);
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}


// ---------------------------------------------------------------- halo weight gradient -------------
// 3x3 / stride 1: the same zero-padded slot space as the halo convolution kernels (umma_conv.cu).  A CTA owns one
// chunk of the gathered input (<= 64 channels of ONE source grid), one column tile of Cout (<= 96) and a range of
// 128-slot tiles.  Per tile the copy engine stages the halo of the chunk (whole slot rows, tma.cuh) and the g rows of the
// 128 slots (whole slot rows too: padding slots arrive as zeros, so they add nothing to the sums); the nine taps are nine
// row-shifted views of that halo.  Two taps form one UMMA M = 128 operand (MN-major: its two 64-channel blocks are the same
// buffer LBO = shift_b - shift_a rows apart), so five accumulators hold dW[tap][64 channels][n_tile] for the whole range.
// Warps: 4 epilogue (TMEM lane quarters), one TMA warp, one MMA warp.
constexpr int WH_THREADS = 192, WH_TMA_WARP = 4, WH_MMA_WARP = 5;

struct WHParams {
  CUtensorMap tmap_x[MG_MAX_SEG];   // the source grids (up-sampling map for the coarser grid)
  CUtensorMap tmap_g;               // the gradient grid
  int n_seg, any_up;
  int seg_up[MG_MAX_SEG];           // the grid is the coarser one (zero-stride map, W slots per row)
  int seg_C[MG_MAX_SEG], seg_Cp[MG_MAX_SEG], seg_cbegin[MG_MAX_SEG];   // logical / padded channels, first concat channel
  int H, W, Wp, Hp, HL;
  int nr_max, halo_bytes;    // halo buffer: slot rows, bytes (multiple of 1024)
  int gnr_max, g_bytes;      // g rows of a tile: slot rows, bytes per 64-channel block (multiple of 1024)
  int64_t T;
  int Cout, Ccat;
  int n_tile, n_blk;
  float* partial;            // [splits][9][Cout][Ccat] fp32 partial sums (context workspace)
  int n_slot_tiles, tiles_per_cta;
  int stages, tmem_cols;
  int kt;                    // slots per pipeline stage (K of one stage): 128 or 256
  int img_box;               // whole-image TMA boxes (7 x 7 grids): buffers hold exactly the halo / the tile's g rows
};

__global__ void __launch_bounds__(WH_THREADS, 1) umma_wgrad_halo_kernel(const __grid_constant__ WHParams p) {
  pdl_launch();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[W_MAX_STAGES], empty_bar[W_MAX_STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  const int stage_bytes = p.halo_bytes + p.n_blk * p.g_bytes;
  const int nt = blockIdx.y;
  const int tile0 = blockIdx.z * p.tiles_per_cta;
  const int n_iters = min(p.tiles_per_cta, p.n_slot_tiles - tile0);
  // chunk blockIdx.x = (source grid sg, first channel c0, nch channels)
  int sg = 0, c0 = (int)blockIdx.x * 64;
  while (sg + 1 < p.n_seg && c0 >= ((p.seg_Cp[sg] + 63) & ~63)) { c0 -= (p.seg_Cp[sg] + 63) & ~63; ++sg; }
  const int nch = min(64, p.seg_Cp[sg] - c0);
  const int up = p.seg_up[sg];

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1 + p.n_blk); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WH_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == WH_TMA_WARP && lane == 0) { tma_prefetch_desc(&p.tmap_x[sg]); tma_prefetch_desc(&p.tmap_g); }
  if (p.img_box) {   // margins of the halo (Wp + 1 slots before / after the tile) are never written by the copy engine: zero them once
    const int m = p.Wp + 1;
    for (int i = tid; i < S * 2 * m * 8; i += WH_THREADS) {
      const int q = i & 7, sl = (i >> 3) % (2 * m), st = (i >> 3) / (2 * m);
      const int slot = sl < m ? sl : p.HL - 2 * m + sl;
      *reinterpret_cast<uint4*>(smem + (size_t)st * stage_bytes + (size_t)slot * 128 + q * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
  } else if (up) {   // the up-sampling box leaves the pad slot of every row alone: zero it once (stage buffers are stage_bytes apart)
    for (int i = tid; i < S * p.nr_max * 8; i += WH_THREADS) {
      const int q = i & 7, row = (i >> 3) % p.nr_max, st = (i >> 3) / p.nr_max;
      *reinterpret_cast<uint4*>(smem + (size_t)st * stage_bytes + (size_t)(row * p.Wp + p.W) * 128 + q * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
  }
  pdl_wait();   // the prologue above overlapped the previous kernel's tail; g and the activations are read below
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 4) {
    // ---- epilogue: five accumulators -> partial sums ------------------------------------------------
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;            // rows 0..63: first tap of the pair, 64..127: second
    const int cl = c0 + (row & 63);              // channel within the chunk's grid
    const bool row_ok = (row & 63) < nch && cl < p.seg_C[sg];
    const int ci = p.seg_cbegin[sg] + cl;
    // plain coalesced stores of this CTA's partial sums (consecutive rows = consecutive ci); the
    // reduction over the pixel splits is a separate pass (wgrad_reduce_kernel): no atomics
    float* part = p.partial + (size_t)blockIdx.z * 9 * p.Cout * p.Ccat;
    for (int pair = 0; pair < 5; ++pair) {
      const int tap = pair * 2 + (row >> 6);
      const bool ok = row_ok && tap < 9;
      float* prow = part + (size_t)tap * p.Cout * p.Ccat + ci;
      for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
        uint32_t acc[16];
        tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(pair * p.n_tile + c0), acc);
        tc_wait_ld();
        if (ok) {
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const int co = nt * p.n_tile + c0 + x;
            if (co < p.Cout) prow[(size_t)co * p.Ccat] = __uint_as_float(acc[x]);
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp == WH_TMA_WARP) {
    // ---- producer: halo rows + g rows of every tile -------------------------------------------------
    Ring rs(S);
    for (int it = 0; it < n_iters; ++it, rs.next()) {
      const int s = rs.idx;
      if (it >= S) mbar_wait(&empty_bar[s], rs.phase ^ 1u);
      const int t0 = (tile0 + it) * p.kt;
      const int hs = t0 - p.Wp - 1;
      const int r0 = floordiv(hs, p.Wp);
      const int nr = (hs + p.HL - 1) / p.Wp - r0 + 1;
      const int gr0 = t0 / p.Wp;
      const int gnr = (t0 + p.kt - 1) / p.Wp - gr0 + 1;
      const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
      if (p.img_box) {
        const int n0 = t0 / (p.Hp * p.Wp);
        tma_load_images(&p.tmap_x[sg], st + (uint32_t)(p.Wp + 1) * 128u, &full_bar[s], c0, n0, (uint32_t)p.kt * 128u, lane);
        for (int b = 0; b < p.n_blk; ++b)
          tma_load_images(&p.tmap_g, st + (uint32_t)(p.halo_bytes + b * p.g_bytes), &full_bar[s], nt * p.n_tile + b * 64, n0, (uint32_t)p.kt * 128u, lane);
      } else {
        tma_load_rows(&p.tmap_x[sg], up, st, &full_bar[s], c0, r0, nr, p.W, p.Hp, lane);
        for (int b = 0; b < p.n_blk; ++b)
          tma_load_rows(&p.tmap_g, 0, st + (uint32_t)(p.halo_bytes + b * p.g_bytes), &full_bar[s], nt * p.n_tile + b * 64, gr0, gnr, p.W, p.Hp, lane);
      }
    }
  } else {
    // ---- MMA issuer: whole warp, elected lane issues (see elect_one) --------------------------------
    const bool leader = elect_one();
    const uint32_t idesc = idesc_bf16_m128_mn(p.n_tile);
    const int nq = p.kt >> 4;
    Ring rs(S);
    for (int it = 0; it < n_iters; ++it, rs.next()) {
      const int s = rs.idx;
      const int t0 = (tile0 + it) * p.kt;
      const int hs = t0 - p.Wp - 1;
      const int off = p.img_box ? 0 : hs - floordiv(hs, p.Wp) * p.Wp;          // first halo slot within the row-aligned buffer
      const int goff = p.img_box ? 0 : t0 - (t0 / p.Wp) * p.Wp;                // first g slot within its row-aligned buffer
      mbar_wait(&full_bar[s], rs.phase);
      tc_fence_after();
      const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes) + (uint32_t)off * 128u;
      const uint32_t b_lo = desc_lo_mn_sw128(smem_u32(smem + (size_t)s * stage_bytes) + (uint32_t)p.halo_bytes + (uint32_t)goff * 128u, (uint32_t)p.g_bytes);
      if (leader) {
#pragma unroll 1
        for (int pair = 0; pair < 5; ++pair) {
          const int ta = pair * 2, tb2 = min(pair * 2 + 1, 8);
          const int sa = (ta / 3) * p.Wp + ta % 3, sb = (tb2 / 3) * p.Wp + tb2 % 3;   // halo slot of tile row 0
          const uint32_t a_lo = desc_lo_mn_sw128(a_base + (uint32_t)sa * 128u, (uint32_t)(sb - sa) * 128u);
          const uint32_t d_tmem = tmem_base + pair * p.n_tile;
#pragma unroll 8
          for (int q = 0; q < nq; ++q)   // 16 slots (K) per UMMA = 2048 bytes = descriptor address + 128
            tc_mma_bf16_lohi(d_tmem, a_lo + q * 128, b_lo + q * 128, DESC_HI_SW128, idesc, (it | q) != 0);
        }
        tc_commit(&empty_bar[s]);
      }
    }
    if (leader) tc_commit(&tmem_full_bar);
  }
  __syncthreads();
  if (warp == WH_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------- CTA-pair halo weight gradient -------------
// The same kernel on tcgen05 CTA pairs (cta_group::2): the two CTAs of a cluster own two consecutive CHUNKS (64 input channels
// each) of the same pixel tiles and column tile.  One MMA (M = 256) issued by the even CTA drives both tensor cores: rows 0-127 are
// the even CTA's tap pair of its chunk, rows 128-255 the odd CTA's; the gradient operand g (N columns) is split, each CTA stages
// only its half of the channels.  Per SM: half the g bytes, half the MMA instructions (the per-instruction issue cost is what
// holds N <= 96 MMAs at ~77 % of the tensor peak: scratch/mma_rate.cu), and a stage small enough for a third ring slot.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WH_THREADS, 1) umma_wgrad_halo_pair_kernel(const __grid_constant__ WHParams p) {
  pdl_launch();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[W_MAX_STAGES], empty_bar[W_MAX_STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int S = p.stages;
  const int stage_bytes = p.halo_bytes + p.g_bytes;      // one 64-channel block of g per CTA: its half of the column tile
  const int nt = blockIdx.y;
  const int tile0 = blockIdx.z * p.tiles_per_cta;
  const int n_iters = min(p.tiles_per_cta, p.n_slot_tiles - tile0);
  int sg = 0, c0 = (int)blockIdx.x * 64;
  bool dummy = false;                                    // odd chunk count: the last pair's odd CTA has no chunk (loads zeros)
  {
    int total = 0;
    for (int s2 = 0; s2 < p.n_seg; ++s2) total += (p.seg_Cp[s2] + 63) & ~63;
    dummy = c0 >= total;
  }
  if (!dummy) while (sg + 1 < p.n_seg && c0 >= ((p.seg_Cp[sg] + 63) & ~63)) { c0 -= (p.seg_Cp[sg] + 63) & ~63; ++sg; }
  const int nch = dummy ? 0 : min(64, p.seg_Cp[sg] - c0);
  const int up = p.seg_up[sg];                           // (a CTA without a chunk reads grid 0 beyond its channels: zeros)
  const int half = p.n_tile >> 1;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WH_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp == WH_TMA_WARP && lane == 0) { tma_prefetch_desc(&p.tmap_x[sg]); tma_prefetch_desc(&p.tmap_g); }
  if (p.img_box) {
    const int m = p.Wp + 1;
    for (int i = tid; i < S * 2 * m * 8; i += WH_THREADS) {
      const int q = i & 7, sl = (i >> 3) % (2 * m), st = (i >> 3) / (2 * m);
      const int slot = sl < m ? sl : p.HL - 2 * m + sl;
      *reinterpret_cast<uint4*>(smem + (size_t)st * stage_bytes + (size_t)slot * 128 + q * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
  } else if (up) {
    for (int i = tid; i < S * p.nr_max * 8; i += WH_THREADS) {
      const int q = i & 7, row = (i >> 3) % p.nr_max, st = (i >> 3) / p.nr_max;
      *reinterpret_cast<uint4*>(smem + (size_t)st * stage_bytes + (size_t)(row * p.Wp + p.W) * 128 + q * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 4) {
    // ---- epilogue: this CTA's chunk, all n_tile columns ------------------------------------------------
    mbar_wait_cluster(&tmem_full_bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const int cl = c0 + (row & 63);
    const bool row_ok = !dummy && (row & 63) < nch && cl < p.seg_C[sg];
    const int ci = dummy ? 0 : p.seg_cbegin[sg] + cl;
    float* part = p.partial + (size_t)blockIdx.z * 9 * p.Cout * p.Ccat;
    for (int pair = 0; pair < 5; ++pair) {
      const int tap = pair * 2 + (row >> 6);
      const bool ok = row_ok && tap < 9;
      float* prow = part + (size_t)tap * p.Cout * p.Ccat + ci;
      for (int cc = 0; cc < p.n_tile; cc += 16) {
        uint32_t acc[16];
        tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(pair * p.n_tile + cc), acc);
        tc_wait_ld();
        if (ok) {
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const int co = nt * p.n_tile + cc + x;
            if (co < p.Cout) prow[(size_t)co * p.Ccat] = __uint_as_float(acc[x]);
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp == WH_TMA_WARP) {
    // ---- producer: this CTA's halo chunk + its half of the g columns; bytes counted on the even CTA's barrier ----
    Ring rs(S);
    for (int it = 0; it < n_iters; ++it, rs.next()) {
      const int s = rs.idx;
      if (it >= S) mbar_wait_cluster(&empty_bar[s], rs.phase ^ 1u);
      const int t0 = (tile0 + it) * p.kt;
      const int hs = t0 - p.Wp - 1;
      const int r0 = floordiv(hs, p.Wp);
      const int nr = (hs + p.HL - 1) / p.Wp - r0 + 1;
      const int gr0 = t0 / p.Wp;
      const int gnr = (t0 + p.kt - 1) / p.Wp - gr0 + 1;
      const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
      const int gc0 = nt * p.n_tile + (int)rank * half;
      // a CTA without a chunk (odd count) loads channels beyond the grid: zeros
      const int xc0 = dummy ? (1 << 20) : c0;
      if (p.img_box) {
        const int n0 = t0 / (p.Hp * p.Wp);
        if (lane == 0) {
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 4u * (uint32_t)p.kt * 128u);
          tma_load_4d_pair(st + (uint32_t)(p.Wp + 1) * 128u, &p.tmap_x[sg], &full_bar[s], xc0, 0, 0, n0);
          tma_load_4d_pair(st + (uint32_t)p.halo_bytes, &p.tmap_g, &full_bar[s], gc0, 0, 0, n0);
        }
      } else {
        // the two CTAs of a pair may stage segments of different kinds (same-size / up-sampled): the even CTA needs its peer's row bytes
        if (rank == 0 && lane == 0) {
          int sg1 = 0, c1 = ((int)blockIdx.x + 1) * 64;
          int total = 0;
          for (int s2 = 0; s2 < p.n_seg; ++s2) total += (p.seg_Cp[s2] + 63) & ~63;
          int up1 = p.seg_up[0];
          if (c1 < total) { while (sg1 + 1 < p.n_seg && c1 >= ((p.seg_Cp[sg1] + 63) & ~63)) { c1 -= (p.seg_Cp[sg1] + 63) & ~63; ++sg1; } up1 = p.seg_up[sg1]; }
          const uint32_t bytes = (uint32_t)nr * (uint32_t)((up ? p.W : p.Wp) + (up1 ? p.W : p.Wp)) * 128u + 2u * (uint32_t)gnr * (uint32_t)p.Wp * 128u;
          mbar_arrive_expect_tx(&full_bar[s], bytes);
        }
        __syncwarp();
        tma_load_rows_pair(&p.tmap_x[sg], up, st, &full_bar[s], xc0, r0, nr, p.W, p.Hp, lane);
        tma_load_rows_pair(&p.tmap_g, 0, st + (uint32_t)p.halo_bytes, &full_bar[s], gc0, gr0, gnr, p.W, p.Hp, lane);
      }
    }
  } else if (rank == 0) {
    // ---- MMA issuer (even CTA): M = 256 across the pair ------------------------------------------
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const int nq = p.kt >> 4;
    Ring rs(S);
    for (int it = 0; it < n_iters; ++it, rs.next()) {
      const int s = rs.idx;
      const int t0 = (tile0 + it) * p.kt;
      const int hs = t0 - p.Wp - 1;
      const int off = p.img_box ? 0 : hs - floordiv(hs, p.Wp) * p.Wp;
      const int goff = p.img_box ? 0 : t0 - (t0 / p.Wp) * p.Wp;
      mbar_wait_cluster(&full_bar[s], rs.phase);
      tc_fence_after();
      const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes) + (uint32_t)off * 128u;
      const uint32_t b_lo = desc_lo_mn_sw128(smem_u32(smem + (size_t)s * stage_bytes) + (uint32_t)p.halo_bytes + (uint32_t)goff * 128u, (uint32_t)p.g_bytes);
      if (leader) {
#pragma unroll 1
        for (int pair = 0; pair < 5; ++pair) {
          const int ta = pair * 2, tb2 = min(pair * 2 + 1, 8);
          const int sa = (ta / 3) * p.Wp + ta % 3, sb = (tb2 / 3) * p.Wp + tb2 % 3;
          const uint32_t a_lo = desc_lo_mn_sw128(a_base + (uint32_t)sa * 128u, (uint32_t)(sb - sa) * 128u);
          const uint32_t d_tmem = tmem_base + pair * p.n_tile;
#pragma unroll 8
          for (int q = 0; q < nq; ++q)
            tc_mma_bf16_lohi_pair(d_tmem, a_lo + q * 128, b_lo + q * 128, DESC_HI_SW128, idesc, (it | q) != 0);
        }
        tc_commit_pair(&empty_bar[s]);
      }
    }
    if (leader) tc_commit_pair(&tmem_full_bar);
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == WH_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// dw[co][ci][tap] += gscale * sum_z partial[z][tap][co][ci]   (z in increasing order: deterministic)
// A CTA owns blockDim.x consecutive (co, ci) pairs x blockDim.y tap lanes: the partial planes are read coalesced along ci, the
// sums are transposed through shared memory and dw is updated as ONE contiguous range of blockDim.x * KK floats (Torch's
// [Cout][Ccat][k][k] layout puts the taps innermost: a thread-per-element store would be a 4-byte write every KK * 4 bytes).
__global__ void __launch_bounds__(288) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int Cout, int Ccat, float* __restrict__ dw,
                                                           float gscale, int KK) {
  pdl_launch();
  pdl_wait();
  extern __shared__ float s_t[];   // [blockDim.x][KK]
  const int64_t pairs = (int64_t)Cout * Ccat;
  const int64_t q0 = (int64_t)blockIdx.x * blockDim.x, q = q0 + threadIdx.x;
  if (q < pairs) {
    const size_t zstride = (size_t)KK * pairs;
    for (int tap = threadIdx.y; tap < KK; tap += blockDim.y) {
      const float* src = partial + (size_t)tap * pairs + q;
      float s = 0.f;
      int z = 0;
      for (; z + 4 <= splits; z += 4) {   // four loads in flight, added in split order
        const float a = src[(size_t)z * zstride], b = src[(size_t)(z + 1) * zstride], c = src[(size_t)(z + 2) * zstride], d = src[(size_t)(z + 3) * zstride];
        s += a; s += b; s += c; s += d;
      }
      for (; z < splits; ++z) s += src[(size_t)z * zstride];
      s_t[threadIdx.x * KK + tap] = s;
    }
  }
  __syncthreads();
  const int64_t n_out = min((int64_t)blockDim.x, pairs - q0) * KK;
  float* dst = dw + q0 * KK;
  const int nthr = blockDim.x * blockDim.y;
  for (int64_t j = threadIdx.y * blockDim.x + threadIdx.x; j < n_out; j += nthr) dst[j] += gscale * s_t[j];
}

static inline cudaError_t launch_wgrad_reduce(mg_ctx* ctx, const float* partial, int z, int Cout, int Ccat, float* dw, float gscale, int KK) {
  const int tl = KK <= 9 ? KK : 8;          // tap lanes
  const int pb = KK == 1 ? 256 : 32;        // pairs per CTA
  const int64_t pairs = (int64_t)Cout * Ccat;
  return mg_launch_pdl(wgrad_reduce_kernel, dim3((unsigned)mg_cdiv(pairs, pb)), dim3(pb, tl), (size_t)pb * KK * sizeof(float), ctx->stream, partial, z, Cout, Ccat, dw, gscale, KK);
}


// ---------------------------------------------------------------- stem weight gradient -------------
// accGradParameters of cudnn.SpatialConvolution(3, C, 7,7, 2,2, 3,3) (models/ilsvrc/rnmg.lua:180), C <= 64.  Same staging as the
// forward stem kernel (umma_conv.cu): a tile is 16 x 8 output pixels, the copy engine loads its 37 x 21-pixel input patch and
// four warps split it into the even / odd column planes.  GEMM per tile: D[co][(tap, ch)] += sum over the 128 pixels of
// g[pixel][co] * x[pixel + tap][ch].  A = the g tile (128 rows of 128 bytes, MN-major SW128, loaded by the copy engine; channels
// beyond Cp arrive as zeros).  B = the patch IN PLACE: in a plane the taps kx, kx+2, kx+4, kx+6 of an output pixel are four
// CONSECUTIVE 16-byte pixels, and the next output pixel is the same window shifted by one pixel -- so an MN-major no-swizzle
// descriptor with SBO = 16 bytes (next tap: core matrices that overlap) and LBO = two plane rows (the next eight pixels = the next
// tile row) reads the Toeplitz operand [16 pixels][4 taps x 8 channels] without any gather.  14 accumulators (7 kernel rows x 2
// parities) of 32 columns live in TMEM over ALL tiles of the (persistent) CTA; the partial sums are reduced in a fixed order.
// The MMAs are M = 64 (Cout <= 64): D row m sits in TMEM lane (m % 16) + 32 * (m / 16).
constexpr int SG_PC = 11, SG_PR = 37, SG_ROW = SG_PC * 16, SG_PLANE = 6592, SG_SLOT = 13312, SG_RAW_COLS = 21, SG_RAW = 12544, SG_G = 16384;
constexpr int SG_STAGE = SG_G + SG_SLOT + SG_RAW, SG_MAX_RING = 5, SG_THREADS = 320;   // 4 epilogue warps, TMA, MMA, 4 de-interleave warps

struct StemWParams {
  CUtensorMap tmap_x;   // kind 2: input patches
  CUtensorMap tmap_g;   // kind 3: 16 x 8 x 64-channel tiles of the gradient
  int tiles_x, tiles_y, n_tiles_total, ring;
  int Cout, Cin;
  float* partial;       // [gridDim.x][Cout][Cin][49]
};

__global__ void __launch_bounds__(SG_THREADS, 1) umma_stem_wgrad_kernel(const __grid_constant__ StemWParams p) {
  pdl_launch();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t r_full[SG_MAX_RING], r_empty[SG_MAX_RING], a_full[SG_MAX_RING], a_empty[SG_MAX_RING], g_full[SG_MAX_RING], tmem_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NA = p.ring;
  uint8_t* g_smem = smem;                                  // ring of g tiles (1024-aligned: SW128)
  uint8_t* a_smem = g_smem + (size_t)NA * SG_G;            // ring of plane pairs
  uint8_t* r_smem = a_smem + (size_t)NA * SG_SLOT;         // ring of raw patches
  const int n_my = (p.n_tiles_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < NA; ++s) {
      mbar_init(&r_full[s], 1); mbar_init(&r_empty[s], 128); mbar_init(&a_full[s], 128); mbar_init(&a_empty[s], 1); mbar_init(&g_full[s], 1);
    }
    mbar_init(&tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&p.tmap_x); tma_prefetch_desc(&p.tmap_g); }
  // column 10 of the odd planes belongs to the non-existent tap kx = 7 (its accumulator columns are never read): finite values
  for (int i = tid; i < NA * SG_PR; i += SG_THREADS) {
    const int sl = i / SG_PR, r = i - sl * SG_PR;
    *reinterpret_cast<uint4*>(a_smem + (size_t)sl * SG_SLOT + SG_PLANE + r * SG_ROW + 10 * 16) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp < 4) {
    // ================= epilogue: 14 accumulators -> this CTA's partial dW, Torch layout [co][ci][ky][kx] =================
    mbar_wait(&tmem_full, 0);
    tc_fence_after();
    // M = 64: row m of D lives in TMEM lane (m % 16) + 32 * (m / 16) -- the lower half of every warp's 32-lane quarter
    const int co = warp * 16 + lane;
    const bool co_ok = lane < 16 && co < p.Cout;
    float* out = p.partial + ((size_t)blockIdx.x * p.Cout + (co_ok ? co : 0)) * p.Cin * 49;
    {
      for (int a = 0; a < 14; ++a) {
        const int ky = a >> 1, par = a & 1;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t acc[16];
          tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * 32 + h * 16), acc);
          tc_wait_ld();
          if (co_ok) {
#pragma unroll
            for (int x = 0; x < 16; ++x) {
              const int j = h * 2 + (x >> 3), ci = x & 7, kx = 2 * j + par;
              if (kx < 7 && ci < p.Cin) out[ci * 49 + ky * 7 + kx] = n_my > 0 ? __uint_as_float(acc[x]) : 0.f;
            }
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ================= producer: patch + g tile per tile =================
    if (lane == 0) {
      Ring rr(NA);
      for (int it = 0; it < n_my; ++it, rr.next()) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int n = tile / tiles_per_img, tr = tile - n * tiles_per_img;
        const int ty = tr / p.tiles_x, tx = tr - ty * p.tiles_x;
        const int s = rr.idx;
        if (it >= NA) mbar_wait(&r_empty[s], rr.phase ^ 1u);
        mbar_arrive_expect_tx(&r_full[s], (uint32_t)(SG_PR * SG_RAW_COLS * 16));
        tma_load_4d(smem_u32(r_smem + (size_t)s * SG_RAW), &p.tmap_x, &r_full[s], 0, 2 * tx * 8 - 3, 2 * ty * 16 - 3, n);
        if (it >= NA) mbar_wait(&a_empty[s], rr.phase ^ 1u);     // the MMAs that read this g tile have completed
        mbar_arrive_expect_tx(&g_full[s], (uint32_t)SG_G);
        tma_load_4d(smem_u32(g_smem + (size_t)s * SG_G), &p.tmap_g, &g_full[s], 0, tx * 8, ty * 16, n);
      }
    }
  } else if (warp == 5) {
    // ================= MMA issuer: per tile 8 K steps (two tile rows each) x 14 accumulators =================
    const bool leader = elect_one();
    // M = 64 (the output channels), N = 32, both operands MN-major: half the shared-memory read of an M = 128 instruction per MMA
    // (the A tile is re-read by every one of the 14 MMAs of a K step: the kernel is bound by that read, not by the MMA rate)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
    // B: MN-major, no swizzle: SBO (MN direction, next tap of the same parity) = 16 bytes, LBO (K direction, next 8 pixels) = 2 plane rows
    const uint32_t b_hi = (16u >> 4) | (1u << 14);
    const uint32_t b_lbo = ((uint32_t)(2 * SG_ROW) >> 4) << 16;
    Ring ra(NA);
    for (int it = 0; it < n_my; ++it, ra.next()) {
      const int s = ra.idx;
      mbar_wait(&a_full[s], ra.phase);
      mbar_wait(&g_full[s], ra.phase);
      tc_fence_after();
      const uint32_t a_lo0 = desc_lo_mn_sw128(smem_u32(g_smem + (size_t)s * SG_G), 0);
      const uint32_t pl = smem_u32(a_smem + (size_t)s * SG_SLOT);
      if (leader) {
#pragma unroll 1
        for (int r = 0; r < 8; ++r) {
          const uint32_t a_lo = a_lo0 + (uint32_t)r * 128u;                     // 16 pixels = 2048 bytes
          const uint32_t row0 = pl + (uint32_t)(4 * r) * SG_ROW;                // plane row of output row 2r, tap row 0
#pragma unroll
          for (int a = 0; a < 14; ++a) {
            const uint32_t baddr = row0 + (uint32_t)((a >> 1) * SG_ROW + (a & 1) * SG_PLANE);
            tc_mma_bf16_lohi2(tmem_base + (uint32_t)(a * 32), a_lo, DESC_HI_SW128, ((baddr >> 4) & 0x3FFFu) | b_lbo, b_hi, idesc, (uint32_t)((it | r) != 0));
          }
        }
        tc_commit(&a_empty[s]);
      }
    }
    if (leader) tc_commit(&tmem_full);
  } else {
    // ================= de-interleave: raw patch -> even / odd column planes =================
    const int dt = tid - 6 * 32;
    Ring rr(NA);
    for (int it = 0; it < n_my; ++it, rr.next()) {
      const int s = rr.idx;
      mbar_wait(&r_full[s], rr.phase);
      if (it >= NA) mbar_wait(&a_empty[s], rr.phase ^ 1u);
      const uint8_t* raw = r_smem + (size_t)s * SG_RAW;
      uint8_t* plp = a_smem + (size_t)s * SG_SLOT;
      for (int i = dt; i < SG_PR * SG_RAW_COLS; i += 128) {
        const int r = i / SG_RAW_COLS, x = i - r * SG_RAW_COLS;
        const uint4 v = *reinterpret_cast<const uint4*>(raw + (size_t)i * 16);
        *reinterpret_cast<uint4*>(plp + (x & 1) * SG_PLANE + r * SG_ROW + (x >> 1) * 16) = v;
      }
      fence_proxy_async();
      mbar_arrive(&a_full[s]);
      mbar_arrive(&r_empty[s]);
    }
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// dw[i] += gscale * sum_z partial[z][i]: blockDim.y lanes each sum every blockDim.y-th split in increasing order, the lanes are
// added in lane order (a fixed tree: deterministic)
__global__ void __launch_bounds__(256) stem_wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int n, float* __restrict__ dw, float gscale) {
  pdl_launch();
  pdl_wait();
  __shared__ float s_red[8][32];
  const int i = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (i < n)
    for (int z = threadIdx.y; z < splits; z += 8) s += partial[(size_t)z * n + i];
  s_red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) t += s_red[l][threadIdx.x];
    dw[i] += gscale * t;
  }
}

}  // namespace

int simt_dbias(mg_ctx* ctx, const mg_grid* g, int Cout, float* dbias, float gscale);

bool umma_wgrad_supported(const mg_ctx* ctx, const mg_conv_desc* d) {
  if (ctx->dtype != MG_BF16) return false;
  if (d->n_seg < 1 || d->n_seg > MG_MAX_SEG) return false;
  for (int s = 0; s < d->n_seg; ++s) {
    const mg_grid& g = d->seg[s];
    if (d->seg_mode[s] == MG_SEG_POOL || g.scale || g.shift || g.Cp % 8) return false;
  }
  return true;
}

static bool wgrad_halo_applies(const mg_conv_desc* d) {
  static int on = -1, min_w = -1;
  if (on < 0) { const char* e = getenv("MGCONV_WGRAD_HALO"); on = e ? atoi(e) : 1; }
  if (min_w < 0) { const char* e = getenv("MGCONV_HALO_MIN_W"); min_w = e ? atoi(e) : 7; }
  if (!on || d->ksize != 3 || d->stride != 1 || d->pad != 1) return false;
  if (d->W < min_w || d->W > 64 || d->H > 1023) return false;
  for (int s = 0; s < d->n_seg; ++s) {
    const mg_grid& g = d->seg[s];
    if (d->seg_mode[s] == MG_SEG_UP) { if (g.H * 2 != d->H || g.W * 2 != d->W) return false; }
    else if (g.H != d->H || g.W != d->W) return false;
  }
  return (int64_t)d->seg[0].N * (d->H + 1) * (d->W + 1) < ((int64_t)1 << 31);
}

static int wgrad_halo(mg_ctx* ctx, const mg_conv_desc* d, const mg_grid* g, float* dw, float gscale) {
  WHParams p;
  memset(&p, 0, sizeof(p));
  int Ccat = 0;
  for (int s = 0; s < d->n_seg; ++s) Ccat += d->seg[s].C;
  p.H = d->H; p.W = d->W; p.Wp = d->W + 1; p.Hp = d->H + 1;
  p.T = (int64_t)g->N * p.Hp * p.Wp;
  p.Cout = d->Cout; p.Ccat = Ccat;
  // five accumulators of n_tile columns must fit the 512 TMEM columns: n_tile <= 96
  const int np = mg_round_up(d->Cout, 16);
  const int n_tiles = (np + 95) / 96;
  p.n_tile = mg_round_up((np + n_tiles - 1) / n_tiles, 16);
  p.n_blk = (p.n_tile + 63) / 64;
  // K tile (slots per pipeline stage): 256 when two such stages fit -- the halo overhead per slot and the number of
  // stage hand-shakes halve, and the copy engine is fed with fewer, larger batches (scratch/tma_bw.cu)
  static int kt_env = -1;
  if (kt_env < 0) { const char* e = getenv("MGCONV_WGRAD_KT"); kt_env = e ? atoi(e) : 0; }
  static int img_env = -1;
  if (img_env < 0) { const char* e = getenv("MGCONV_IMG_BOX"); img_env = e ? atoi(e) : 1; }
  bool any_up = false;
  for (int s = 0; s < d->n_seg; ++s) any_up = any_up || d->seg_mode[s] == MG_SEG_UP;
  for (p.kt = (kt_env == 128 ? 128 : 256); ; p.kt = 128) {
    // whole-image TMA boxes when the images ((H+1)*(W+1) slots: 64 on the 7 x 7 grids) divide the K tile: one copy per operand
    p.img_box = (img_env && !any_up && p.kt % (p.Hp * p.Wp) == 0) ? 1 : 0;
    p.HL = p.kt + 2 * p.Wp + 2;
    p.nr_max = (p.HL - 1 + p.Wp - 1) / p.Wp + 1;
    p.halo_bytes = mg_round_up((p.img_box ? p.HL : p.nr_max * p.Wp) * 128, 1024);
    p.gnr_max = (p.kt - 1 + p.Wp - 1) / p.Wp + 1;
    p.g_bytes = mg_round_up((p.img_box ? p.kt : p.gnr_max * p.Wp) * 128, 1024);
    if (p.kt == 128 || 2 * (p.halo_bytes + p.n_blk * p.g_bytes) + 1024 <= SMEM_WGRAD) break;
  }
  p.n_slot_tiles = (int)mg_cdiv(p.T, p.kt);
  int n_chunks = 0;
  for (int s = 0; s < d->n_seg; ++s) n_chunks += (d->seg[s].Cp + 63) / 64;
  int64_t splits = std::max<int64_t>(1, (int64_t)ctx->num_sms / ((int64_t)n_chunks * n_tiles));
  splits = std::min<int64_t>(splits, std::max(1, p.n_slot_tiles / 2));
  p.tiles_per_cta = (int)mg_cdiv(p.n_slot_tiles, splits);
  const int z = (int)mg_cdiv(p.n_slot_tiles, p.tiles_per_cta);
  const int stage_bytes = p.halo_bytes + p.n_blk * p.g_bytes;
  int S = std::min(W_MAX_STAGES, (SMEM_WGRAD - 1024) / stage_bytes);
  S = std::max(2, std::min(S, std::max(2, p.tiles_per_cta)));
  MG_REQUIRE(ctx, S * stage_bytes + 1024 <= SMEM_WGRAD, MG_ERR_UNSUPPORTED, "halo wgrad: %d bytes of shared memory", S * stage_bytes + 1024);
  p.stages = S;
  int cols = 32;
  while (cols < 5 * p.n_tile) cols <<= 1;
  p.tmem_cols = cols;
  static bool attr_set = false;
  if (!attr_set) {
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_WGRAD));
    attr_set = true;
  }
  void* ws = nullptr;
  const size_t plane = (size_t)9 * p.Cout * p.Ccat;
  int rc = mg_ctx_workspace(ctx, (size_t)z * plane * sizeof(float), &ws);
  if (rc) return rc;
  p.partial = (float*)ws;
  const int imgs = p.kt / (p.Hp * p.Wp);
  rc = p.img_box ? mg_tensor_map(ctx, g->data, g->N, g->H, g->W, g->Cp, 4, imgs, &p.tmap_g) : mg_tensor_map(ctx, g->data, g->N, g->H, g->W, g->Cp, 0, p.Wp, &p.tmap_g);
  if (rc) return rc;
  int c = 0;
  p.n_seg = d->n_seg;
  for (int s = 0; s < d->n_seg; ++s) {
    const mg_grid& sg = d->seg[s];
    p.seg_up[s] = d->seg_mode[s] == MG_SEG_UP ? 1 : 0;
    if (p.seg_up[s]) p.any_up = 1;
    rc = p.img_box ? mg_tensor_map(ctx, sg.data, sg.N, sg.H, sg.W, sg.Cp, 4, imgs, &p.tmap_x[s])
                   : mg_tensor_map(ctx, sg.data, sg.N, sg.H, sg.W, sg.Cp, p.seg_up[s], p.seg_up[s] ? 2 * sg.W : p.Wp, &p.tmap_x[s]);
    if (rc) return rc;
    p.seg_C[s] = sg.C; p.seg_Cp[s] = sg.Cp; p.seg_cbegin[s] = c;
    c += sg.C;
  }
  // CTA pairs (cta_group::2): two chunks per cluster share the pixel tiles, each CTA stages half of the g columns.  Taken when the
  // chunk count is even (an odd count leaves one CTA of the last pair without work); MGCONV_WGRAD_PAIR = 0 never, 2 always
  static int pair_env = -1;
  if (pair_env < 0) { const char* e = getenv("MGCONV_WGRAD_PAIR"); pair_env = e ? atoi(e) : 1; }
  const bool pair = pair_env == 2 || (pair_env == 1 && n_chunks % 2 == 0);
  if (pair) {
    const int n_chunks2 = mg_round_up(n_chunks, 2);
    int64_t sp = std::max<int64_t>(1, (int64_t)ctx->num_sms / ((int64_t)n_chunks2 * n_tiles));
    sp = std::min<int64_t>(sp, std::max(1, p.n_slot_tiles / 2));
    p.tiles_per_cta = (int)mg_cdiv(p.n_slot_tiles, sp);
    const int zp = (int)mg_cdiv(p.n_slot_tiles, p.tiles_per_cta);
    const int stage_p = p.halo_bytes + p.g_bytes;
    static int smem_cap = -1;   // (experiment) MGCONV_WGRAD_SMEM_KB: leave shared memory for co-resident CTAs of other lanes
    if (smem_cap < 0) { const char* e = getenv("MGCONV_WGRAD_SMEM_KB"); smem_cap = e ? atoi(e) * 1024 : SMEM_WGRAD; }
    int Sp = std::min(W_MAX_STAGES, (std::min(SMEM_WGRAD, smem_cap) - 1024) / stage_p);
    Sp = std::max(2, std::min(Sp, std::max(2, p.tiles_per_cta)));
    p.stages = Sp;
    rc = mg_ctx_workspace(ctx, (size_t)zp * plane * sizeof(float), &ws);
    if (rc) return rc;
    p.partial = (float*)ws;
    static bool pair_attr = false;
    if (!pair_attr) {
      MG_CUDA(ctx, cudaFuncSetAttribute(umma_wgrad_halo_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_WGRAD));
      pair_attr = true;
    }
    dim3 gridp((unsigned)n_chunks2, (unsigned)n_tiles, (unsigned)zp);
    MG_CUDA(ctx, mg_launch_pdl(umma_wgrad_halo_pair_kernel, gridp, dim3(WH_THREADS), (size_t)(Sp * stage_p + 1024), ctx->stream, p));
    MG_CHECK_LAUNCH(ctx);
    ctx->tc_launches++;
    MG_CUDA(ctx, launch_wgrad_reduce(ctx, p.partial, zp, p.Cout, p.Ccat, dw, gscale, 9));
    MG_CHECK_LAUNCH(ctx);
    return MG_OK;
  }
  dim3 grid((unsigned)n_chunks, (unsigned)n_tiles, (unsigned)z);
  MG_CUDA(ctx, mg_launch_pdl(umma_wgrad_halo_kernel, grid, dim3(WH_THREADS), (size_t)(S * stage_bytes + 1024), ctx->stream, p));
  MG_CHECK_LAUNCH(ctx);
  ctx->tc_launches++;
  MG_CUDA(ctx, launch_wgrad_reduce(ctx, p.partial, z, p.Cout, p.Ccat, dw, gscale, 9));
  MG_CHECK_LAUNCH(ctx);
  return MG_OK;
}

static bool wgrad_stem_applies(const mg_conv_desc* d, const mg_grid* g) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MGCONV_STEM_WGRAD"); on = e ? atoi(e) : 1; }
  if (!on || d->ksize != 7 || d->stride != 2 || d->pad != 3 || d->n_seg != 1 || d->seg_mode[0] != MG_SEG_SAME) return false;
  const mg_grid& x = d->seg[0];
  return x.Cp == 8 && x.H == d->H && x.W == d->W && d->Cout <= 64 && g->Cp % 8 == 0 && !x.scale;
}

static int wgrad_stem(mg_ctx* ctx, const mg_conv_desc* d, const mg_grid* g, float* dw, float gscale) {
  StemWParams p;
  memset(&p, 0, sizeof(p));
  const mg_grid& x = d->seg[0];
  int rc = mg_tensor_map(ctx, x.data, x.N, x.H, x.W, 8, 2, SG_RAW_COLS, &p.tmap_x);
  if (rc) return rc;
  rc = mg_tensor_map(ctx, g->data, g->N, g->H, g->W, g->Cp, 3, 8, &p.tmap_g);
  if (rc) return rc;
  p.tiles_x = (g->W + 7) / 8; p.tiles_y = (g->H + 15) / 16;
  p.n_tiles_total = g->N * p.tiles_x * p.tiles_y;
  p.Cout = d->Cout; p.Cin = x.C;
  const int grid = std::max(1, std::min(ctx->num_sms, p.n_tiles_total));
  p.ring = std::max(2, std::min(SG_MAX_RING, (p.n_tiles_total + grid - 1) / grid));
  const int n = d->Cout * x.C * 49;
  void* ws = nullptr;
  rc = mg_ctx_workspace(ctx, (size_t)grid * n * sizeof(float), &ws);
  if (rc) return rc;
  p.partial = (float*)ws;
  static bool attr_set = false;
  if (!attr_set) {
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SG_MAX_RING * SG_STAGE + 1024));
    attr_set = true;
  }
  MG_CUDA(ctx, mg_launch_pdl(umma_stem_wgrad_kernel, dim3(grid), dim3(SG_THREADS), (size_t)(p.ring * SG_STAGE + 1024), ctx->stream, p));
  MG_CHECK_LAUNCH(ctx);
  ctx->tc_launches++;
  MG_CUDA(ctx, mg_launch_pdl(stem_wgrad_reduce_kernel, dim3((unsigned)mg_cdiv(n, 32)), dim3(32, 8), 0, ctx->stream, (const float*)p.partial, grid, n, dw, gscale));
  MG_CHECK_LAUNCH(ctx);
  return MG_OK;
}

int umma_conv_backward_weight(mg_ctx* ctx, const mg_conv_desc* d, const mg_grid* g, float* dw, float* dbias, float gscale) {
  if (wgrad_stem_applies(d, g)) {
    int rc = wgrad_stem(ctx, d, g, dw, gscale);
    if (rc) return rc;
    if (dbias) return simt_dbias(ctx, g, d->Cout, dbias, gscale);
    return MG_OK;
  }
  if (wgrad_halo_applies(d)) {
    int rc = wgrad_halo(ctx, d, g, dw, gscale);
    if (rc) return rc;
    if (dbias) return simt_dbias(ctx, g, d->Cout, dbias, gscale);
    return MG_OK;
  }
  WParams p;
  memset(&p, 0, sizeof(p));
  p.n_seg = d->n_seg;
  int c = 0, cp = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    const mg_grid& sg = d->seg[s];
    const int m = d->seg_mode[s];
    if (m == MG_SEG_SAME) MG_REQUIRE(ctx, sg.H == d->H && sg.W == d->W, MG_ERR_SHAPE, "wgrad: SAME seg %d is %dx%d, expected %dx%d", s, sg.H, sg.W, d->H, d->W);
    else MG_REQUIRE(ctx, sg.H * 2 == d->H && sg.W * 2 == d->W, MG_ERR_SHAPE, "wgrad: UP seg %d is %dx%d, x2 != %dx%d", s, sg.H, sg.W, d->H, d->W);
    p.seg[s].ptr = (const __nv_bfloat16*)sg.data; p.seg[s].Hs = sg.H; p.seg[s].Ws = sg.W; p.seg[s].Cp = sg.Cp;
    p.seg[s].shift = m == MG_SEG_UP ? 1 : 0; p.seg[s].kv_begin = cp / 8;
    p.seg_C[s] = sg.C; p.seg_cbegin[s] = c;
    c += sg.C; cp += sg.Cp;
  }
  p.k = d->ksize; p.stride = d->stride; p.pad = d->pad; p.H = d->H; p.W = d->W;
  p.Ho = g->H; p.Wo = g->W;
  p.M = (int64_t)g->N * g->H * g->W;
  p.kv_per_tap = cp / 8; p.nkv = p.k * p.k * p.kv_per_tap;
  p.g = (const __nv_bfloat16*)g->data; p.g_cp = g->Cp;
  p.Cout = d->Cout; p.Ccat = c;
  const int np = mg_round_up(d->Cout, 16);
  const int n_tiles = (np + 255) / 256;
  p.n_tile = mg_round_up((np + n_tiles - 1) / n_tiles, 16);
  p.n_blk = (p.n_tile + 63) / 64;
  p.gscale = gscale;
  const int m_tiles = (p.nkv + 15) / 16;
  // split the pixel range so that four CTAs are resident per SM (the loaders are latency bound)
  static int per_sm = -1;
  if (per_sm < 0) { const char* e = getenv("MGCONV_WGRAD_CTAS"); per_sm = e ? atoi(e) : 4; }
  const int64_t want = (int64_t)per_sm * ctx->num_sms;
  int64_t splits = std::max<int64_t>(1, want / ((int64_t)m_tiles * n_tiles));
  splits = std::min<int64_t>(splits, mg_cdiv(p.M, 4 * PIX));
  splits = std::max<int64_t>(1, std::min<int64_t>(splits, 65535));
  p.pix_per_cta = mg_round_up((int)mg_cdiv(p.M, splits), PIX);
  const int z = (int)mg_cdiv(p.M, p.pix_per_cta);
  const int stage_bytes = 2 * A_IMG + p.n_blk * A_IMG;
  static int budget_kb = -1;
  if (budget_kb < 0) { const char* e = getenv("MGCONV_SMEM_KB"); budget_kb = e ? atoi(e) : 54; }
  int S = std::min(W_MAX_STAGES, (budget_kb * 1024) / stage_bytes);
  const int iters = (int)(p.pix_per_cta / PIX);
  S = std::max(2, std::min(S, std::max(2, iters)));
  p.stages = S; p.lag = std::max(0, std::min(S - 2, 3));
  int cols = 32;
  while (cols < p.n_tile) cols <<= 1;
  p.tmem_cols = cols;
  static bool attr_set = false;
  if (!attr_set) {
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 8 * 1024));
    attr_set = true;
  }
  const int KK = p.k * p.k;
  const size_t plane = (size_t)KK * p.Cout * p.Ccat;
  void* ws = nullptr;
  int rc = mg_ctx_workspace(ctx, (size_t)z * plane * sizeof(float), &ws);
  if (rc) return rc;
  p.partial = (float*)ws;
  // rows of the last M tile beyond nkv and pad channels are never written: the reduction must not read garbage there
  // -- every (tap, co, ci) of the plane IS written by exactly one (M tile, column tile) of every split
  dim3 grid((unsigned)m_tiles, (unsigned)n_tiles, (unsigned)z);
  umma_wgrad_kernel<<<grid, W_THREADS, S * stage_bytes + 1024, ctx->stream>>>(p);
  MG_CHECK_LAUNCH(ctx);
  ctx->tc_launches++;
  MG_CUDA(ctx, launch_wgrad_reduce(ctx, p.partial, z, p.Cout, p.Ccat, dw, gscale, KK));
  MG_CHECK_LAUNCH(ctx);
  if (dbias) return simt_dbias(ctx, g, d->Cout, dbias, gscale);
  return MG_OK;
}
