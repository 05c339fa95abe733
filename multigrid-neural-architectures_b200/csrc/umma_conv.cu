// tcgen05 / TMEM implicit-GEMM multigrid convolution for sm_100a (bf16 operands, fp32 accumulate).
//
//   D[m][n] = sum_k A[m][k] * B[n][k]        m = output pixel (n, oy, ox), n = output channel
//
// K runs over "k-vectors" of 8 channels (16 bytes): for tap in k*k, for segment in the conv's
// gather list, for c8 in Cp_seg/8.  A is never materialised: four producer warps copy each
// k-vector of each of the CTA's 128 pixels straight from the source grids into the 128B-swizzled
// K-major operand image in shared memory with cp.async (zero-fill for the conv's zero padding);
// the same-scale grid, the pooled companion of the finer grid and the coarser grid (read at
// (y>>1, x>>1) -- SpatialUpSamplingNearest) are just different base pointers / shifts, so the
// channel concatenation of ResampleConcat (models/ilsvrc/rnmg.lua:41-89) exists only as the order
// of the K loop.  B (the weights) is pre-packed by pack_weights_kernel into exactly the shared
// memory image of each pipeline stage, so one thread moves a stage with a single cp.async.bulk.
// One elected thread issues tcgen05.mma (M=128, N<=256, K=16) into a TMEM accumulator; the
// producer warps drain it with tcgen05.ld, add the bias and store bf16 NHWC rows.
//
// forward:  A = gather(x),  B = W[co][(tap,ci)]                      -> y[m][co]
// dgrad  :  A = g (one SAME segment), B = W^T[cpad][(tap',co)], taps mirrored -> dcat[m][cpad]
// wgrad  :  see umma_wgrad.cu (MN-major operands).
#include "common.cuh"
#include "conv_view.cuh"
#include "umma_common.cuh"
#include <algorithm>
#include <vector>

namespace {

constexpr int BM = 128;           // pixels per CTA = UMMA M
constexpr int KV_PER_STAGE = 8;   // 8 k-vectors of 8 bf16 = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * 128;
constexpr int N_PRODUCERS = 128;
constexpr int MMA_WARP = 4, B_WARP = 5;
constexpr int N_THREADS = 192;
constexpr int MAX_STAGES = 8;

struct UParams {
  USeg seg[MG_MAX_SEG];
  int n_seg;
  int k, stride, pad;
  int H, W;      // logical (concatenated) input size
  int Ho, Wo;
  int64_t M;     // N * Ho * Wo
  int Nimg;
  int kv_per_tap, nkv, n_stages;
  int n_tile;    // UMMA N of this launch (multiple of 16, <= 256)
  const uint8_t* wpack;  // [n_tiles][n_stages][n_tile][128B]
  const float* bias;     // [c_bias] or null
  int c_bias;
  __nv_bfloat16* y;
  int y_pitch;   // elements per output pixel
  int c_valid;   // channels to write (multiple of 8)
  int stages, lag;
  int tmem_cols;
  // halo kernel (3x3, stride 1): M rows are SLOTS of the zero-padded linear image space
  int Wp, Hp;        // slot pitch of an image row (W + 1) and rows per image (H + 1): the extra column / row is the conv's zero padding
  int64_t T;         // N * Hp * Wp slots
  int HL;            // halo slots staged per 64-channel chunk = 128 + 2 * Wp + 2
  int n_chunks;      // ceil(kv_per_tap / 8)
  int halo_bytes;    // HL * 128 rounded up to 1024
  int n_abuf;        // halo buffers (1 or 2)
  long long* timeline;   // debug (MGCONV_TIMELINE=1): per-CTA clock stamps [grid][8], else null
  double* stats;         // halo forward: per-channel (sum y, sum y^2) of the STORED bf16 output accumulated here ([2][c_stats]), or null
  int c_stats;
  // persistent halo kernel
  int m_tiles, n_ntiles, n_items;   // slot tiles, column tiles, work items = m_tiles * n_ntiles
};

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M=128
__host__ __device__ constexpr uint32_t idesc_bf16_m128(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// ---------------------------------------------------------------- the kernel ---------------
template <bool FAST>
__global__ void __launch_bounds__(N_THREADS, 4) umma_conv_kernel(const __grid_constant__ UParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem is only guaranteed 16-byte aligned: round up to the 1024 bytes the swizzle atoms need
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t s_row[BM];
  __shared__ USeg s_seg[MG_MAX_SEG];
  __shared__ float s_bias[256];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages;
  const int b_stage_bytes = p.n_tile * 128;
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + (size_t)S * A_STAGE_BYTES;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int ntile = blockIdx.y;

  if (tid < p.n_seg) s_seg[tid] = p.seg[tid];
  if (tid < BM) {  // packed (n, oy, ox) of this CTA's rows; 0xFFFFFFFF = row beyond M
    int64_t m = m0 + tid;
    uint32_t v = 0xFFFFFFFFu;
    if (m < p.M) {
      if (p.M < ((int64_t)1 << 32)) {   // 32-bit divisions: the prologue of 25 088 CTAs on the stem
        const uint32_t mu = (uint32_t)m, q = mu / (uint32_t)p.Wo, ox = mu - q * (uint32_t)p.Wo;
        const uint32_t n = q / (uint32_t)p.Ho, oy = q - n * (uint32_t)p.Ho;
        v = (n << 20) | (oy << 10) | ox;
      } else {
        int ox = (int)(m % p.Wo); int64_t q = m / p.Wo;
        int oy = (int)(q % p.Ho); int n = (int)(q / p.Ho);
        v = ((uint32_t)n << 20) | ((uint32_t)oy << 10) | (uint32_t)ox;
      }
    }
    s_row[tid] = v;
  }
  if (tid >= BM - 64 && tid < BM + 64) {   // bias tile of this column tile (zero beyond Cout): no global loads in the epilogue
    for (int c = tid - (BM - 64); c < p.n_tile; c += 128) {
      const int ch = blockIdx.y * p.n_tile + c;
      s_bias[c] = (p.bias && ch < p.c_bias) ? p.bias[ch] : 0.f;
    }
  }
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], N_PRODUCERS + 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 4) {
    // ================= A producers: gather k-vectors with cp.async ==========================
    const int v = tid & 7;       // k-vector slot of the stage (16-byte column of the 128-byte row)
    const int rg = tid >> 3;     // row group 0..15; rows rg, rg+16, ...
    const int L = p.lag;
    if (FAST) {
      // stride 1, "same" padding, k in {1,3}: everything that depends on the pixel is computed once per
      // CTA -- linear pixel index (SAME segments), half-resolution pixel index (UP segments) and a
      // validity bit per tap -- so that one k-vector copy costs a handful of instructions.
      uint32_t pix_up[BM / 16], flags[BM / 16];
      const uint32_t pix_same0 = (uint32_t)m0 + rg;   // pixel index of row it*16+rg = pix_same0 + it*16
      const int Hs2 = p.H >> 1, Ws2 = p.W >> 1;
#pragma unroll
      for (int it = 0; it < BM / 16; ++it) {
        const uint32_t ri = s_row[it * 16 + rg];
        const int ox = ri & 1023, oy = (ri >> 10) & 1023, n = ri >> 20;
        uint32_t f = 0;
        if (ri != 0xFFFFFFFFu) {
          for (int t = 0; t < p.k * p.k; ++t) {
            const int iy = oy + t / p.k - p.pad, ix = ox + t % p.k - p.pad;
            if ((unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W) f |= 1u << t;
          }
          f |= (uint32_t)(oy & 1) << 9 | (uint32_t)(ox & 1) << 10;
          pix_up[it] = (uint32_t)(((size_t)n * Hs2 + (oy >> 1)) * Ws2 + (ox >> 1));
        } else {
          pix_up[it] = 0;
        }
        flags[it] = f;
      }
      const uint32_t dst_thread = (uint32_t)(rg * 128 + ((v ^ (rg & 7)) << 4));  // (it*16+rg)&7 == rg&7
      Ring rs(S), rpub(S);
      // this thread's k-vector j = ks * 8 + v as (tap = (ty, tx), r): carried along instead of three divisions per stage
      int tap = v / p.kv_per_tap, r = v - tap * p.kv_per_tap;
      int ty = tap / p.k, tx = tap - ty * p.k;
      for (int ks = 0; ks < p.n_stages + L; ++ks, rs.next()) {
        if (ks < p.n_stages) {
          const int s = rs.idx;
          if (ks >= S) mbar_wait(&empty_bar[s], rs.phase ^ 1u);
          const int j = ks * KV_PER_STAGE + v;
          const bool kv_ok = j < p.nkv;
          int sg = 0;
          if (kv_ok) while (sg + 1 < p.n_seg && r >= s_seg[sg + 1].kv_begin) ++sg;
          const USeg sgm = s_seg[sg];
          const int c8 = r - sgm.kv_begin;
          const int dy = ty - p.pad, dx = tx - p.pad;
          const uint32_t dst0 = smem_u32(a_smem + (size_t)s * A_STAGE_BYTES) + dst_thread;
          const uint32_t tapbit = kv_ok ? (1u << tap) : 0u;
          const uint32_t pitch = (uint32_t)sgm.Cp * 2u;
          const char* base = reinterpret_cast<const char*>(sgm.ptr) + c8 * 16;
          if (sgm.shift == 0) {
            base += (int64_t)(dy * p.W + dx) * (int64_t)pitch;
#pragma unroll
            for (int it = 0; it < BM / 16; ++it) {
              const bool ok = (flags[it] & tapbit) != 0;
              const char* src = ok ? base + (uint64_t)(pix_same0 + it * 16) * pitch : reinterpret_cast<const char*>(sgm.ptr);
              cp_async16(dst0 + it * 2048, src, ok ? 16u : 0u);
            }
          } else {
            // (o+d)>>1 - (o>>1):  d=-1 -> -1 if o even;  d=+1 -> +1 if o odd
#pragma unroll
            for (int it = 0; it < BM / 16; ++it) {
              const uint32_t f = flags[it];
              const bool ok = (f & tapbit) != 0;
              const int py = (f >> 9) & 1, px = (f >> 10) & 1;
              const int dyo = dy < 0 ? py - 1 : (dy > 0 ? py : 0);
              const int dxo = dx < 0 ? px - 1 : (dx > 0 ? px : 0);
              const char* src = ok ? base + (int64_t)((int64_t)pix_up[it] + dyo * Ws2 + dxo) * (int64_t)pitch
                                   : reinterpret_cast<const char*>(sgm.ptr);
              cp_async16(dst0 + it * 2048, src, ok ? 16u : 0u);
            }
          }
          r += KV_PER_STAGE;
          while (r >= p.kv_per_tap) { r -= p.kv_per_tap; ++tap; if (++tx == p.k) { tx = 0; ++ty; } }
        }
        cp_async_commit();
        if (ks >= L) {
          cp_async_wait_dyn(L);
          fence_proxy_async();
          mbar_arrive(&full_bar[rpub.idx]);
          rpub.next();
        }
      }
    } else {
      // general geometry (any k / stride / pad: the 7x7 stride-2 stem).  The kernel is issue bound (ncu: 59 % of the issue
      // slots on the stem), so everything that depends only on the row is decoded once per CTA: image index and the
      // input coordinates of tap (0,0); a copy then costs two adds, two bound checks and one address computation.
      int riy[BM / 16], rix[BM / 16], rn[BM / 16];
#pragma unroll
      for (int it = 0; it < BM / 16; ++it) {
        const uint32_t ri = s_row[it * 16 + rg];
        const bool valid = ri != 0xFFFFFFFFu;
        rn[it] = valid ? (int)(ri >> 20) : 0;
        riy[it] = valid ? (int)((ri >> 10) & 1023) * p.stride - p.pad : -(1 << 20);   // far outside: every bound check fails
        rix[it] = valid ? (int)(ri & 1023) * p.stride - p.pad : -(1 << 20);
      }
      const uint32_t dst_thread = (uint32_t)(rg * 128 + ((v ^ (rg & 7)) << 4));          // (it*16+rg)&7 == rg&7
      Ring rs(S), rpub(S);
      int tap = v / p.kv_per_tap, r = v - tap * p.kv_per_tap;
      int ty = tap / p.k, tx = tap - ty * p.k;
      for (int ks = 0; ks < p.n_stages + L; ++ks, rs.next()) {
        if (ks < p.n_stages) {
          const int s = rs.idx;
          if (ks >= S) mbar_wait(&empty_bar[s], rs.phase ^ 1u);
          // decode this lane's k-vector: (tap = (ty, tx), segment, channel offset), carried along from stage to stage
          const int j = ks * KV_PER_STAGE + v;
          const bool kv_ok = j < p.nkv;
          int sg = 0;
          if (kv_ok) while (sg + 1 < p.n_seg && r >= s_seg[sg + 1].kv_begin) ++sg;
          const USeg sgm = s_seg[sg];
          const int dy = kv_ok ? ty : -(1 << 20), dx = tx;
          const char* base = reinterpret_cast<const char*>(sgm.ptr) + (r - sgm.kv_begin) * 16;
          const uint32_t pitch = (uint32_t)sgm.Cp * 2u;
          const int plane = sgm.Hs * sgm.Ws;
          const uint32_t dst0 = smem_u32(a_smem + (size_t)s * A_STAGE_BYTES) + dst_thread;
#pragma unroll
          for (int it = 0; it < BM / 16; ++it) {
            const int iy = riy[it] + dy, ix = rix[it] + dx;
            const bool ok = (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
            const int64_t pixel = (int64_t)rn[it] * plane + (iy >> sgm.shift) * sgm.Ws + (ix >> sgm.shift);
            const char* src = ok ? base + pixel * pitch : reinterpret_cast<const char*>(sgm.ptr);
            cp_async16(dst0 + it * 2048, src, ok ? 16u : 0u);
          }
          r += KV_PER_STAGE;
          while (r >= p.kv_per_tap) { r -= p.kv_per_tap; if (++tx == p.k) { tx = 0; ++ty; } }
        }
        cp_async_commit();
        if (ks >= L) {  // stage ks-L has landed for this thread: publish it to the tensor core (async proxy)
          cp_async_wait_dyn(L);
          fence_proxy_async();
          mbar_arrive(&full_bar[rpub.idx]);
          rpub.next();
        }
      }
    }
    // ================= epilogue: TMEM -> registers -> bf16 NHWC rows ======================
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const int64_t m = m0 + row;
    const int n_base = ntile * p.n_tile;
    __nv_bfloat16* yrow = p.y + (size_t)(m < p.M ? m : 0) * p.y_pitch;
    for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
      uint32_t acc[16];
      tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, acc);
      tc_wait_ld();
      if (m < p.M) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int n0 = n_base + c0 + h * 8;
          if (n0 + 8 <= p.c_valid) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = __uint_as_float(acc[h * 8 + 2 * e]) + s_bias[c0 + h * 8 + 2 * e];
              const float b = __uint_as_float(acc[h * 8 + 2 * e + 1]) + s_bias[c0 + h * 8 + 2 * e + 1];
              __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
              pk[e] = *reinterpret_cast<uint32_t*>(&t);
            }
            *reinterpret_cast<uint4*>(yrow + n0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
    }
    tc_fence_before();
  } else if (warp == B_WARP) {
    // ================= B loader: one bulk copy per stage =====================================
    if (lane == 0) {
      const uint8_t* wsrc = p.wpack + (size_t)ntile * p.n_stages * b_stage_bytes;
      Ring rs(S);
      for (int ks = 0; ks < p.n_stages; ++ks, rs.next()) {
        const int s = rs.idx;
        if (ks >= S) mbar_wait(&empty_bar[s], rs.phase ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)b_stage_bytes);
        bulk_g2s(smem_u32(b_smem + (size_t)s * b_stage_bytes), wsrc + (size_t)ks * b_stage_bytes, (uint32_t)b_stage_bytes, &full_bar[s]);
      }
    }
  } else {
    // ================= MMA issuer: whole warp, elected lane issues (see elect_one) =================
    {
      const bool leader = elect_one();
      const uint32_t idesc = idesc_bf16_m128(p.n_tile);
      Ring rs(S);
      for (int ks = 0; ks < p.n_stages; ++ks, rs.next()) {
        const int s = rs.idx;
        mbar_wait(&full_bar[s], rs.phase);
        tc_fence_after();
        const uint32_t a_lo = desc_lo_k_sw128(smem_u32(a_smem + (size_t)s * A_STAGE_BYTES));
        const uint32_t b_lo = desc_lo_k_sw128(smem_u32(b_smem + (size_t)s * b_stage_bytes));
        const int kv_here = min(KV_PER_STAGE, p.nkv - ks * KV_PER_STAGE);
        const int ksteps = (kv_here + 1) >> 1;   // 16 bf16 = 2 k-vectors per UMMA K step
        if (leader) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (q < ksteps) tc_mma_bf16_lohi(tmem_base, a_lo + q * 2, b_lo + q * 2, DESC_HI_SW128, idesc, (ks | q) != 0);
          tc_commit(&empty_bar[s]);   // frees the stage once these MMAs have read it
        }
      }
      if (leader) tc_commit(&tmem_full_bar);    // accumulator complete
    }
  }
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}


// ---------------------------------------------------------------- halo kernel ----------------
// 3x3 / stride 1 convolutions.  The CTA's 128 GEMM rows are 128 consecutive SLOTS of the zero-padded
// linear image space t = (n*(H+1) + y)*(W+1) + x (slot x == W and row y == H are the conv's zero
// padding, shared between neighbouring rows / images), so that tap (ky,kx) of every row is the slot
// (ky-1)*(W+1) + (kx-1) further on.  Per 64-channel chunk the producers stage ONE halo of
// 128 + 2*(W+1) + 2 slots; the nine taps are nine UMMA descriptors into that same buffer, offset by
// whole 128-byte rows (the 128B swizzle is a function of the absolute shared-memory address, so a
// descriptor may start at any row -- verified on hardware, scratch/desc_test.cu).  A is fetched from
// L2 once per chunk instead of once per tap (x4.4 - x7 less gather traffic than umma_conv_kernel).
// butterfly reduce-scatter over the 32 lanes of a warp: every lane contributes v[0..15], afterwards lane l
// holds the warp-wide sum of element l & 15 (16 shuffles instead of 16 x 5)
__device__ __forceinline__ float warp_reduce_scatter16(float (&v)[16], int lane) {
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = hi ? v[i] : v[i + o];
      const float keep = hi ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

constexpr int HALO_MAX_SLOTS = 256 + 2 * 65 + 2;   // W <= 64, up to two 128-slot sub-tiles per CTA
constexpr int H_PROD = 256;                        // 8 loader warps (the first 4 also drain TMEM): the gather is issue bound
constexpr int H_MMA_WARP = H_PROD / 32, H_B_WARP = H_MMA_WARP + 1;
constexpr int H_THREADS = H_PROD + 64;

// CL = thread-block cluster size along the tile index: the CL CTAs of a cluster need the same weight stages, so each
// loads 1/CL of a stage and multicasts it to all of them (L2 -> SM weight traffic / CL)
// MT = 128-slot sub-tiles per CTA (1 or 2).  With MT = 2 the CTA owns 256 consecutive slots and two TMEM accumulators:
// every weight stage feeds both sub-tiles, so the weight stream per output row halves and the halo overhead per row
// drops -- the kernel is bound by L2 -> SM bandwidth (measured ~42 B/clk/SM chip-wide), of which the weights that every
// tile re-streams are 70-90 %.  All eight loader warps drain TMEM (warps 0-3 accumulator 0, warps 4-7 accumulator 1).
template <int CL, int MT>
__global__ void __launch_bounds__(H_THREADS, MT == 1 ? 4 : 3) umma_conv_halo_kernel(const __grid_constant__ UParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], a_full[2], a_empty[2], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t s_pix[BM * MT + 132];    // full-resolution pixel index of a halo slot, 0xFFFFFFFF = padding / outside
  __shared__ uint32_t s_pup[BM * MT + 132];    // half-resolution pixel index (UP segments)
  __shared__ USeg s_seg[MG_MAX_SEG];
  __shared__ float s_bias[256];                // bias of this column tile (zero beyond Cout): no global loads in the epilogue

  pdl_launch();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long* tl = p.timeline ? p.timeline + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
  if (tl && tid == 0) { tl[0] = clock64(); unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); tl[7] = sm; }
  const int S = p.stages;
  const int b_stage_bytes = p.n_tile * 128;
  uint8_t* a_smem = smem;                               // two halo buffers
  uint8_t* b_smem = smem + (size_t)p.n_abuf * p.halo_bytes;    // B ring
  const int64_t t0 = (int64_t)blockIdx.x * (BM * MT);
  const int ntile = blockIdx.y;
  const int KK = 9;

  if (tid < p.n_seg) s_seg[tid] = p.seg[tid];

  {
    const int slots_per_img = p.Hp * p.Wp;
    const int Hs2 = p.H >> 1, Ws2 = p.W >> 1;
    for (int h = tid; h < p.HL; h += H_THREADS) {
      const int64_t t = t0 - p.Wp - 1 + h;
      uint32_t pix = 0xFFFFFFFFu, pup = 0;
      if (t >= 0 && t < p.T) {
        // T < 2^31 (halo_applies): 32-bit unsigned divisions (the 64-bit ones cost ~1k cycles of every CTA's prologue)
        const uint32_t tu = (uint32_t)t;
        const uint32_t n = tu / (uint32_t)slots_per_img, rem = tu - n * (uint32_t)slots_per_img;
        const uint32_t yy = rem / (uint32_t)p.Wp, xs = rem - yy * (uint32_t)p.Wp;
        if ((int)yy < p.H && (int)xs < p.W) {
          pix = (n * p.H + yy) * p.W + xs;
          pup = (n * Hs2 + (yy >> 1)) * Ws2 + (xs >> 1);
        }
      }
      s_pix[h] = pix; s_pup[h] = pup;
    }
  }
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CL); }
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], H_PROD); mbar_init(&a_empty[s], 1); }
    mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == H_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // everything above touched only kernel parameters, shared memory and TMEM: it overlapped the tail of the
  // previous kernel; from here on the predecessor's output (activations, packed weights, bias) is read
  pdl_wait();
  if (tid < p.n_tile) {
    const int ch = blockIdx.y * p.n_tile + tid;
    s_bias[tid] = (p.bias && ch < p.c_bias) ? p.bias[ch] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // every CTA's barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tl && tid == 0) tl[1] = clock64();

  if (warp < H_MMA_WARP) {
    // ================= A producers: one halo per 64-channel chunk ===============================
    const int v = tid & 7;       // k-vector column of the chunk
    const int rg = tid >> 3;     // halo slots rg, rg+16, ...
    const int NB = p.n_abuf, lagA = NB - 1;   // one buffer: publish at once; two: publish the previous chunk
    Ring ra(NB), rpub(NB);
    for (int c = 0; c < p.n_chunks + lagA; ++c, ra.next()) {
      if (c < p.n_chunks) {
        const int buf = ra.idx;
        if (c >= NB) mbar_wait(&a_empty[buf], ra.phase ^ 1u);
        const int r = c * KV_PER_STAGE + v;            // k-vector within a tap
        const bool kv_ok = r < p.kv_per_tap;
        int sg = 0;
        if (kv_ok) while (sg + 1 < p.n_seg && r >= s_seg[sg + 1].kv_begin) ++sg;
        const USeg sgm = s_seg[sg];
        const uint32_t pitch = (uint32_t)sgm.Cp * 2u;
        const char* base = reinterpret_cast<const char*>(sgm.ptr) + (r - sgm.kv_begin) * 16;
        const uint32_t* tab = sgm.shift ? s_pup : s_pix;
        const uint32_t dst0 = smem_u32(a_smem + (size_t)buf * p.halo_bytes) + (uint32_t)(v << 4);
        // table reads are batched ahead of the copies (the asm copies are ordered, the compiler cannot hoist them)
        for (int h0 = rg; h0 < p.HL; h0 += 4 * (H_PROD / 8)) {
          uint32_t pv[4], tv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int h = h0 + u * (H_PROD / 8);
            pv[u] = h < p.HL ? s_pix[h] : 0xFFFFFFFFu;
            tv[u] = h < p.HL ? tab[h] : 0u;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int h = h0 + u * (H_PROD / 8);
            if (h < p.HL) {
              const bool ok = kv_ok && pv[u] != 0xFFFFFFFFu;
              const char* src = ok ? base + (uint64_t)tv[u] * pitch : reinterpret_cast<const char*>(sgm.ptr);
              // 16-byte chunk v of slot h, swizzled by the slot's absolute 128-byte row (buffers are 1024-aligned)
              cp_async16((dst0 ^ ((uint32_t)(h & 7) << 4)) + (uint32_t)h * 128, src, ok ? 16u : 0u);
            }
          }
        }
      }
      cp_async_commit();
      if (c >= lagA) {   // chunk c-lagA has landed for this thread
        cp_async_wait_dyn(lagA);
        fence_proxy_async();
        mbar_arrive(&a_full[rpub.idx]);
        rpub.next();
      }
    }
    // ================= epilogue (warps 0-3): TMEM -> registers -> bf16 NHWC rows ===========
    // With p.stats the BatchNorm statistics of this tile (sum, sum of squares of the STORED bf16 values over the
    // valid rows) are reduced here: warp butterfly -> per-warp slots in the (now idle) halo buffer -> one fp64
    // atomic pair per channel and CTA.  The separate statistics pass over y (one HBM read of y) disappears.
    if (warp < 4 * MT) {
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    if (tl && tid == 0) tl[3] = clock64();
    const int row = warp * 32 + lane;                      // sub-tile warp >> 2, TMEM lane quarter warp & 3
    const uint32_t pix = s_pix[row + p.Wp + 1];          // slot t0 + row
    const bool row_ok = pix != 0xFFFFFFFFu;
    const int n_base = ntile * p.n_tile;
    const bool want_stats = p.stats != nullptr;
    float* s_part = reinterpret_cast<float*>(a_smem);     // [4*MT warps][2][256]: every MMA has completed, the halo buffer is free
    __nv_bfloat16* yrow = p.y + (size_t)(row_ok ? pix : 0) * p.y_pitch;
    const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * p.n_tile);
    for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
      uint32_t acc[16];
      tc_ld16(tacc + (uint32_t)c0, acc);
      tc_wait_ld();
      uint32_t pk[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int n0 = n_base + c0 + h * 8;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = __uint_as_float(acc[h * 8 + 2 * e]) + s_bias[c0 + h * 8 + 2 * e];
          const float b = __uint_as_float(acc[h * 8 + 2 * e + 1]) + s_bias[c0 + h * 8 + 2 * e + 1];
          __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
          pk[h][e] = *reinterpret_cast<uint32_t*>(&t);
        }
        if (row_ok && n0 + 8 <= p.c_valid) *reinterpret_cast<uint4*>(yrow + n0) = make_uint4(pk[h][0], pk[h][1], pk[h][2], pk[h][3]);
      }
      if (want_stats) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const bool use = row_ok && n_base + c0 + h * 8 + 8 <= p.c_valid;
          float sv[16];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float ra = use ? __uint_as_float(pk[h][e] << 16) : 0.f, rb = use ? __uint_as_float(pk[h][e] & 0xFFFF0000u) : 0.f;
            sv[2 * e] = ra; sv[2 * e + 1] = rb;
            sv[8 + 2 * e] = ra * ra; sv[8 + 2 * e + 1] = rb * rb;
          }
          const float tot = warp_reduce_scatter16(sv, lane);   // lane l: (l & 15) < 8 -> sum of column, else sum of squares
          if (lane < 16) s_part[(warp * 2 + (lane >> 3)) * 256 + c0 + h * 8 + (lane & 7)] = tot;
        }
      }
    }
    tc_fence_before();
    if (want_stats) {
      asm volatile("bar.sync 1, %0;" ::"n"(128 * MT) : "memory");       // the epilogue warps
      for (int c = tid; c < p.n_tile; c += 128 * MT) {
        const int ch = n_base + c;
        if (ch < p.c_stats) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int w = 0; w < 4 * MT; ++w) { a += s_part[(2 * w) * 256 + c]; b += s_part[(2 * w + 1) * 256 + c]; }
          atomicAdd(p.stats + ch, (double)a);
          atomicAdd(p.stats + p.c_stats + ch, (double)b);
        }
      }
    }
    if (tl && tid == 0) tl[4] = clock64();
    }
  } else if (warp == H_B_WARP) {
    // ================= B loader: one bulk copy per (chunk, tap) stage ============================
    if (lane == 0) {
      const int n_st = p.n_chunks * KK;
      const uint8_t* wsrc = p.wpack + (size_t)ntile * n_st * b_stage_bytes;
      const uint32_t rank = CL > 1 ? cluster_ctarank() : 0;
      const uint32_t slice = (uint32_t)b_stage_bytes / CL;   // n_tile * 128 / CL: a multiple of 16 bytes
      Ring rb(S);
      for (int ks = 0; ks < n_st; ++ks, rb.next()) {
        const int s = rb.idx;
        if (ks >= S) mbar_wait(&empty_bar[s], rb.phase ^ 1u);   // CL arrivals: every CTA of the cluster freed slot s
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)b_stage_bytes);
        if (CL == 1)
          bulk_g2s(smem_u32(b_smem + (size_t)s * b_stage_bytes), wsrc + (size_t)ks * b_stage_bytes, (uint32_t)b_stage_bytes, &full_bar[s]);
        else
          bulk_g2s_mcast(smem_u32(b_smem + (size_t)s * b_stage_bytes) + rank * slice, wsrc + (size_t)ks * b_stage_bytes + rank * slice, slice,
                         &full_bar[s], (uint16_t)((1u << CL) - 1));
      }
    }
  } else {
    // ================= MMA issuer: 9 shifted descriptors per chunk ===============================
    // the whole warp runs the loop (warp-uniform operands), the elected lane issues the MMAs and the commits
    {
      const bool leader = elect_one();
      const uint32_t idesc = idesc_bf16_m128(p.n_tile);
      int ks = 0;
      long long wait_a = 0, wait_b = 0, tq = 0;
      Ring ra(p.n_abuf), rb(S);
      for (int c = 0; c < p.n_chunks; ++c, ra.next()) {
        const int buf = ra.idx;
        if (tl) tq = clock64();
        mbar_wait(&a_full[buf], ra.phase);
        tc_fence_after();
        if (tl && c == 0 && lane == 0) tl[2] = clock64();
        if (tl && c > 0) wait_a += clock64() - tq;
        const uint32_t a_base = smem_u32(a_smem + (size_t)buf * p.halo_bytes);
        const int kv_here = min(KV_PER_STAGE, p.kv_per_tap - c * KV_PER_STAGE);
        const int ksteps = (kv_here + 1) >> 1;
        for (int tap = 0; tap < KK; ++tap, ++ks, rb.next()) {
          const int s = rb.idx;
          if (tl) tq = clock64();
          mbar_wait(&full_bar[s], rb.phase);
          tc_fence_after();
          if (tl) wait_b += clock64() - tq;
          const uint32_t a_lo = desc_lo_k_sw128(a_base + (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * 128u);
          const uint32_t b_lo = desc_lo_k_sw128(smem_u32(b_smem + (size_t)s * b_stage_bytes));
          if (leader) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)   // both sub-tiles consume the same weight stage
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (q < ksteps)
                  tc_mma_bf16_lohi(tmem_base + (uint32_t)(mt * p.n_tile), a_lo + (uint32_t)(mt * BM * 8 + q * 2), b_lo + (uint32_t)(q * 2), DESC_HI_SW128,
                                   idesc, (ks | q) != 0);
            if (CL == 1) tc_commit(&empty_bar[s]); else tc_commit_mcast(&empty_bar[s], (uint16_t)((1u << CL) - 1));
          }
        }
        if (leader) tc_commit(&a_empty[buf]);   // halo buffer free once this chunk's MMAs have read it
      }
      if (leader) tc_commit(&tmem_full_bar);
      if (tl && lane == 0) { tl[5] = wait_a; tl[6] = wait_b; }
    }
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == H_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}


// ---------------------------------------------------------------- persistent halo kernel ----------
// Same operand scheme as umma_conv_halo_kernel for convolutions whose WHOLE packed weight image fits in shared
// memory next to a ring of halo buffers (R-MG-34 block 1: 96 -> 64 at 56x56 is 147 KB, its dgrad 110 KB).  The
// non-persistent kernel re-streams the weights for every 128-slot tile -- 2.4x the activation traffic at N = 64 --
// through a two-deep ring, which makes its main loop a chain of L2 round trips.  Here one CTA per SM loads the
// weights ONCE, then walks slot tiles blockIdx.x, +gridDim.x, ... with decoupled roles:
//   8 loader warps   stage halos into a ring that runs ACROSS chunks and tiles (the next tile's halo is in
//                    flight while the tensor core works on the current one),
//   1 MMA thread     accumulates tile i into TMEM accumulator i & 1 (every operand already in shared memory),
//   4 epilogue warps drain accumulator (i-1) & 1 meanwhile (TMEM double buffering) and keep the BatchNorm
//                    statistics of all the CTA's tiles in shared memory: one fp64 atomic pair per channel and CTA.
constexpr int P_LOAD = 256, P_EPI_WARP0 = 8, P_MMA_WARP = 12, P_B_WARP = 13, P_THREADS = 448;
constexpr int P_MAX_ABUF = 6;

__global__ void __launch_bounds__(P_THREADS, 1) umma_conv_halo_persistent_kernel(const __grid_constant__ UParams p) {
  pdl_launch();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t b_full, a_full[P_MAX_ABUF], a_empty[P_MAX_ABUF], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t s_pix[2][BM + 132], s_pup[2][BM + 132];
  __shared__ USeg s_seg[MG_MAX_SEG];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NA = p.n_abuf;
  const int b_stage_bytes = p.n_tile * 128;
  const int KK = 9;
  const int n_st = p.n_chunks * KK;
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + (size_t)NA * p.halo_bytes;                        // all n_st weight stages, resident
  float* s_bias = reinterpret_cast<float*>(b_smem + (size_t)n_st * b_stage_bytes);   // [n_tile]
  float* s_part = s_bias + p.n_tile;                                         // [4 epilogue warps][2][n_tile]
  const int n_my = (p.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles blockIdx.x, +gridDim.x, ...

  if (tid < p.n_seg) s_seg[tid] = p.seg[tid];
  for (int c = tid; c < 8 * p.n_tile; c += P_THREADS) s_part[c] = 0.f;
  if (tid == 0) {
    mbar_init(&b_full, 1);
    for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], P_LOAD / 2); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == P_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  for (int c = tid; c < p.n_tile; c += P_THREADS) s_bias[c] = (p.bias && c < p.c_bias) ? p.bias[c] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int slots_per_img = p.Hp * p.Wp;

  if (warp < P_EPI_WARP0) {
    // ================= loaders: two independent groups of four warps ===============================
    // Group g stages chunks g, g+2, ... of the CTA's chunk sequence (which runs across tiles): wait for the ring
    // slot, issue, wait for the data, publish.  While one group sits out the L2 latency of its chunk the other
    // one is issuing, and neither ever waits on a buffer the other group's chunk has to release first.
    const int grp = warp >> 2, gt = tid & 127;
    const int v = gt & 7, rg = gt >> 3;            // k-vector column, slot lane 0..15
    const int Hs2 = p.H >> 1, Ws2 = p.W >> 1;
    const int total_chunks = n_my * p.n_chunks;
    int cur_it = -1;
    Ring ra(NA);                       // this group's chunks are ac = grp, grp + 2, ...: the ring advances two slots per iteration
    if (grp) ra.next();
    int it = grp / p.n_chunks, c = grp - it * p.n_chunks;      // (tile, chunk) of ac, carried along
    for (int ac = grp; ac < total_chunks; ac += 2, ra.next(), ra.next(), c += 2) {
      while (c >= p.n_chunks) { c -= p.n_chunks; ++it; }
      if (it != cur_it) {   // this group's slot tables of tile `it`
        cur_it = it;
        const int64_t t0 = (int64_t)(blockIdx.x + it * gridDim.x) * BM;
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");   // old table no longer read
        for (int h = gt; h < p.HL; h += 128) {
          const int64_t t = t0 - p.Wp - 1 + h;
          uint32_t pix = 0xFFFFFFFFu, pup = 0;
          if (t >= 0 && t < p.T) {
            const uint32_t tu = (uint32_t)t;
            const uint32_t n = tu / (uint32_t)slots_per_img, rem = tu - n * (uint32_t)slots_per_img;
            const uint32_t yy = rem / (uint32_t)p.Wp, xs = rem - yy * (uint32_t)p.Wp;
            if ((int)yy < p.H && (int)xs < p.W) {
              pix = (n * p.H + yy) * p.W + xs;
              pup = (n * Hs2 + (yy >> 1)) * Ws2 + (xs >> 1);
            }
          }
          s_pix[grp][h] = pix; s_pup[grp][h] = pup;
        }
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
      }
      const int buf = ra.idx;
      if (ac >= NA) mbar_wait(&a_empty[buf], ra.phase ^ 1u);
      const int r = c * KV_PER_STAGE + v;
      const bool kv_ok = r < p.kv_per_tap;
      int sg = 0;
      if (kv_ok) while (sg + 1 < p.n_seg && r >= s_seg[sg + 1].kv_begin) ++sg;
      const USeg sgm = s_seg[sg];
      const uint32_t pitch = (uint32_t)sgm.Cp * 2u;
      const char* base = reinterpret_cast<const char*>(sgm.ptr) + (r - sgm.kv_begin) * 16;
      const uint32_t* tab = sgm.shift ? s_pup[grp] : s_pix[grp];
      const uint32_t dst0 = smem_u32(a_smem + (size_t)buf * p.halo_bytes) + (uint32_t)(v << 4);
      for (int h0 = rg; h0 < p.HL; h0 += 4 * 16) {
        uint32_t pv[4], tv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int h = h0 + u * 16;
          pv[u] = h < p.HL ? s_pix[grp][h] : 0xFFFFFFFFu;
          tv[u] = h < p.HL ? tab[h] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int h = h0 + u * 16;
          if (h < p.HL) {
            const bool ok = kv_ok && pv[u] != 0xFFFFFFFFu;
            const char* src = ok ? base + (uint64_t)tv[u] * pitch : reinterpret_cast<const char*>(sgm.ptr);
            cp_async16((dst0 ^ ((uint32_t)(h & 7) << 4)) + (uint32_t)h * 128, src, ok ? 16u : 0u);
          }
        }
      }
      cp_async_commit();
      cp_async_wait_dyn(0);
      fence_proxy_async();
      mbar_arrive(&a_full[buf]);
    }
  } else if (warp < P_MMA_WARP) {
    // ================= epilogue warps: drain accumulator it & 1 ====================================
    const int ew = warp - P_EPI_WARP0, et = tid - P_EPI_WARP0 * 32;   // TMEM lane quarter, thread 0..127
    const bool want_stats = p.stats != nullptr;
    for (int it = 0; it < n_my; ++it) {
      const int mt = blockIdx.x + it * gridDim.x;
      const int acc = it & 1;
      // this thread's output pixel: slot t0 + row
      const int row = ew * 32 + lane;
      const int64_t t = (int64_t)mt * BM + row;
      bool row_ok = false; uint32_t pix = 0;
      if (t < p.T) {
        const uint32_t tu = (uint32_t)t;
        const uint32_t n = tu / (uint32_t)slots_per_img, rem = tu - n * (uint32_t)slots_per_img;
        const uint32_t yy = rem / (uint32_t)p.Wp, xs = rem - yy * (uint32_t)p.Wp;
        if ((int)yy < p.H && (int)xs < p.W) { row_ok = true; pix = (n * p.H + yy) * p.W + xs; }
      }
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      __nv_bfloat16* yrow = p.y + (size_t)pix * p.y_pitch;
      const uint32_t tcol = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * p.n_tile);
      for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
        uint32_t a16[16];
        tc_ld16(tcol + (uint32_t)c0, a16);
        tc_wait_ld();
        uint32_t pk[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int n0 = c0 + h * 8;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = __uint_as_float(a16[h * 8 + 2 * e]) + s_bias[c0 + h * 8 + 2 * e];
            const float b = __uint_as_float(a16[h * 8 + 2 * e + 1]) + s_bias[c0 + h * 8 + 2 * e + 1];
            __nv_bfloat162 t2 = __floats2bfloat162_rn(a, b);
            pk[h][e] = *reinterpret_cast<uint32_t*>(&t2);
          }
          if (row_ok && n0 + 8 <= p.c_valid) *reinterpret_cast<uint4*>(yrow + n0) = make_uint4(pk[h][0], pk[h][1], pk[h][2], pk[h][3]);
        }
        if (want_stats) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const bool use = row_ok && c0 + h * 8 + 8 <= p.c_valid;
            float sv[16];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float ra = use ? __uint_as_float(pk[h][e] << 16) : 0.f, rb = use ? __uint_as_float(pk[h][e] & 0xFFFF0000u) : 0.f;
              sv[2 * e] = ra; sv[2 * e + 1] = rb;
              sv[8 + 2 * e] = ra * ra; sv[8 + 2 * e + 1] = rb * rb;
            }
            const float tot = warp_reduce_scatter16(sv, lane);
            // slot owned by this lane of this warp for the whole kernel: accumulate over the CTA's tiles without synchronisation
            if (lane < 16) s_part[(ew * 2 + (lane >> 3)) * p.n_tile + c0 + h * 8 + (lane & 7)] += tot;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);   // 128 arrivals: the accumulator may be overwritten
    }
    if (want_stats) {
      asm volatile("bar.sync 3, 128;" ::: "memory");
      for (int c = et; c < p.n_tile; c += 128)
        if (c < p.c_stats) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) { a += s_part[(w * 2) * p.n_tile + c]; b += s_part[(w * 2 + 1) * p.n_tile + c]; }
          atomicAdd(p.stats + c, (double)a);
          atomicAdd(p.stats + p.c_stats + c, (double)b);
        }
    }
  } else if (warp == P_B_WARP) {
    // ================= weight loader: every stage once =============================================
    if (lane == 0) {
      mbar_arrive_expect_tx(&b_full, (uint32_t)(n_st * b_stage_bytes));
      for (int q = 0; q < n_st; ++q)
        bulk_g2s(smem_u32(b_smem + (size_t)q * b_stage_bytes), p.wpack + (size_t)q * b_stage_bytes, (uint32_t)b_stage_bytes, &b_full);
    }
  } else {
    // ================= MMA issuer (whole warp, elected lane issues; see elect_one) =================
    {
      const bool leader = elect_one();
      const uint32_t idesc = idesc_bf16_m128(p.n_tile);
      const uint32_t b_base = smem_u32(b_smem);
      int ac = 0;
      Ring ra(NA);
      mbar_wait(&b_full, 0);
      for (int it = 0; it < n_my; ++it) {
        const int acc = it & 1;
        if (it >= 2) { mbar_wait(&tmem_empty[acc], ((it >> 1) - 1) & 1); tc_fence_after(); }
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n_tile);
        for (int c = 0; c < p.n_chunks; ++c, ++ac, ra.next()) {
          const int buf = ra.idx;
          mbar_wait(&a_full[buf], ra.phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(a_smem + (size_t)buf * p.halo_bytes);
          const int kv_here = min(KV_PER_STAGE, p.kv_per_tap - c * KV_PER_STAGE);
          const int ksteps = (kv_here + 1) >> 1;
          if (leader) {
            for (int tap = 0; tap < KK; ++tap) {
              const uint32_t a_lo = desc_lo_k_sw128(a_base + (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * 128u);
              const uint32_t b_lo = desc_lo_k_sw128(b_base + (uint32_t)((c * KK + tap) * b_stage_bytes));
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (q < ksteps) tc_mma_bf16_lohi(d_tmem, a_lo + (uint32_t)(q * 2), b_lo + (uint32_t)(q * 2), DESC_HI_SW128, idesc, (c | tap | q) != 0);
            }
            tc_commit(&a_empty[buf]);
          }
        }
        if (leader) tc_commit(&tmem_full[acc]);
      }
    }
  }
  __syncthreads();
  if (warp == P_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------- weight packing -----------
struct PackParams {
  const float* w;   // [Cout][Ccat][k][k]
  uint8_t* out;
  int transposed;
  int k, Ccat, Cout;
  int n_seg;
  int seg_C[MG_MAX_SEG], seg_cbegin[MG_MAX_SEG], seg_kvbegin[MG_MAX_SEG], seg_cpbegin[MG_MAX_SEG];
  int kv_per_tap, nkv, n_stages, n_tile, n_tiles;
  int n_rows_valid;  // forward: Cout; transposed: CcatP (rows that may be non-zero)
  int halo;          // stage order of umma_conv_halo_kernel: stage = chunk * 9 + tap, k-vector = chunk * 8 + v
};

// one thread per (n row, k-vector): writes 16 bytes of the swizzled stage image
__device__ __forceinline__ void pack_one(const PackParams& p, int64_t i) {
  const int kv_total = p.n_stages * KV_PER_STAGE;
  const int64_t total = (int64_t)p.n_tiles * p.n_tile * kv_total;
  if (i >= total) return;
  const int j = (int)(i % kv_total);
  const int nrow = (int)(i / kv_total);          // global row = tile * n_tile + local
  const int tile = nrow / p.n_tile, nl = nrow % p.n_tile;
  const int stage = j / KV_PER_STAGE, v = j % KV_PER_STAGE;
  __nv_bfloat16 vals[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) vals[e] = __float2bfloat16_rn(0.f);
  const int KK = p.k * p.k;
  int tap, r;
  bool kv_ok;
  if (p.halo) { tap = stage % KK; r = (stage / KK) * KV_PER_STAGE + v; kv_ok = r < p.kv_per_tap; }
  else { tap = j / p.kv_per_tap; r = j % p.kv_per_tap; kv_ok = j < p.nkv; }
  if (kv_ok && nrow < p.n_rows_valid) {
    if (!p.transposed) {
      int sg = 0;
      while (sg + 1 < p.n_seg && r >= p.seg_kvbegin[sg + 1]) ++sg;
      const int c0 = (r - p.seg_kvbegin[sg]) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (c0 + e < p.seg_C[sg]) vals[e] = __float2bfloat16_rn(p.w[((size_t)nrow * p.Ccat + p.seg_cbegin[sg] + c0 + e) * KK + tap]);
    } else {
      // row = padded concat channel, K = (mirrored tap, output channel co)
      int sg = 0;
      while (sg + 1 < p.n_seg && nrow >= p.seg_cpbegin[sg + 1]) ++sg;
      const int cl = nrow - p.seg_cpbegin[sg];
      if (cl < p.seg_C[sg]) {
        const int ci = p.seg_cbegin[sg] + cl;
        const int ky = p.k - 1 - tap / p.k, kx = p.k - 1 - tap % p.k;
        const int co0 = r * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (co0 + e < p.Cout) vals[e] = __float2bfloat16_rn(p.w[((size_t)(co0 + e) * p.Ccat + ci) * KK + ky * p.k + kx]);
      }
    }
  }
  uint8_t* dst = p.out + ((size_t)(tile * p.n_stages + stage) * p.n_tile + nl) * 128 + ((v ^ (nl & 7)) << 4);
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(vals);
}

__global__ void pack_weights_kernel(PackParams p) {
  pdl_launch();
  pdl_wait();
  pack_one(p, (int64_t)blockIdx.x * blockDim.x + threadIdx.x);
}

// every convolution of a plan in ONE launch: block b serves job j with blk_begin[j] <= b < blk_begin[j+1]
// (150 launches of ~4 us each, pure launch latency, become one HBM-bound pass)
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const PackParams* __restrict__ jobs, const int* __restrict__ blk_begin, int n_jobs) {
  pdl_launch();
  pdl_wait();
  int lo = 0, hi = n_jobs - 1;
  const int b = (int)blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (blk_begin[mid] <= b) lo = mid; else hi = mid - 1;
  }
  __shared__ PackParams sp;
  const int* src = reinterpret_cast<const int*>(jobs + lo);
  for (int t = threadIdx.x; t < (int)(sizeof(PackParams) / sizeof(int)); t += blockDim.x) reinterpret_cast<int*>(&sp)[t] = src[t];
  __syncthreads();
  pack_one(sp, (int64_t)(b - blk_begin[lo]) * blockDim.x + threadIdx.x);
}

// ---------------------------------------------------------------- host side -----------------
struct Geometry {
  int kv_per_tap, nkv, n_stages, n_tile, n_tiles, n_rows;
  int halo, n_chunks;
};

// the halo kernel serves 3x3 / stride 1 / pad 1 convolutions on grids of at least `MGCONV_HALO_MIN_W`
// (default 7) and at most 64 columns whose UP segments are exactly half size
static bool halo_applies(const mg_conv_desc* d) {
  static int min_w = -1;
  if (min_w < 0) { const char* e = getenv("MGCONV_HALO_MIN_W"); min_w = e ? atoi(e) : 7; }
  if (d->ksize != 3 || d->stride != 1 || d->pad != 1) return false;
  if (d->W < min_w || d->W > 64 || d->H > 1023) return false;
  for (int s = 0; s < d->n_seg; ++s) {
    const mg_grid& g = d->seg[s];
    if (d->seg_mode[s] == MG_SEG_UP) { if (g.H * 2 != d->H || g.W * 2 != d->W) return false; }
    else if (g.H != d->H || g.W != d->W) return false;
  }
  return (int64_t)d->seg[0].N * (d->H + 1) * (d->W + 1) < ((int64_t)1 << 31);
}

static void n_tiling(int n_rows_pad16, int* n_tile, int* n_tiles) {
  *n_tiles = (n_rows_pad16 + 255) / 256;
  *n_tile = mg_round_up((n_rows_pad16 + *n_tiles - 1) / *n_tiles, 16);
}

// geometry of the forward (transposed = 0) or dgrad (transposed = 1) GEMM of a conv descriptor
static Geometry geometry(const mg_conv_desc* d, int transposed) {
  Geometry g;
  int CcatP = 0;
  for (int s = 0; s < d->n_seg; ++s) CcatP += d->seg[s].Cp;
  const int CoutP = mg_round_up(d->Cout, 8);
  const int taps = d->ksize * d->ksize;
  if (!transposed) { g.kv_per_tap = CcatP / 8; g.n_rows = d->Cout; n_tiling(mg_round_up(d->Cout, 16), &g.n_tile, &g.n_tiles); }
  else { g.kv_per_tap = CoutP / 8; g.n_rows = CcatP; n_tiling(mg_round_up(CcatP, 16), &g.n_tile, &g.n_tiles); }
  g.nkv = taps * g.kv_per_tap;
  g.n_stages = (g.nkv + KV_PER_STAGE - 1) / KV_PER_STAGE;
  g.halo = halo_applies(d) ? 1 : 0;
  g.n_chunks = (g.kv_per_tap + KV_PER_STAGE - 1) / KV_PER_STAGE;
  if (g.halo) g.n_stages = g.n_chunks * taps;   // one weight stage per (chunk, tap)
  return g;
}

static int pick_stages(int n_tile, int* smem_bytes) {
  const int stage = A_STAGE_BYTES + n_tile * 128;
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("MGCONV_STAGES"); forced = e ? atoi(e) : 0; }
  static int budget_kb = -1;   // per-CTA shared memory target: ~54 KB keeps four CTAs resident per SM (measured best on R-MG-34)
  if (budget_kb < 0) { const char* e = getenv("MGCONV_SMEM_KB"); budget_kb = e ? atoi(e) : 54; }
  int S = forced > 0 ? forced : std::min(MAX_STAGES, (budget_kb * 1024) / stage);
  S = std::max(2, std::min(S, MAX_STAGES));
  *smem_bytes = S * stage + 1024;  // + alignment slack
  return S;
}

static int launch(mg_ctx* ctx, UParams& p, int n_tiles) {
  int smem = 0;
  p.stages = pick_stages(p.n_tile, &smem);
  p.stages = std::min(p.stages, std::max(2, p.n_stages));
  p.lag = std::max(0, std::min(p.stages - 2, 3));   // lag <= S-2: loaders issue ahead while the MMAs of the current stage run
  int cols = 32;
  while (cols < p.n_tile) cols <<= 1;
  p.tmem_cols = cols;
  static bool attr_set = false;
  if (!attr_set) {
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 8 * 1024));
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 8 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)mg_cdiv(p.M, BM), (unsigned)n_tiles);
  // fast gather: stride 1, "same" padding (output size = input size), 1x1 or 3x3, every UP segment at exactly half size
  bool fast = p.stride == 1 && (p.k == 1 || p.k == 3) && p.pad == p.k / 2 && p.Ho == p.H && p.Wo == p.W && p.M < (int64_t)1 << 31;
  for (int s = 0; s < p.n_seg; ++s) {
    if (p.seg[s].shift) fast = fast && p.seg[s].Hs == p.H / 2 && p.seg[s].Ws == p.W / 2 && p.H % 2 == 0 && p.W % 2 == 0;
    else fast = fast && p.seg[s].Hs == p.H && p.seg[s].Ws == p.W;
  }
  if (fast) umma_conv_kernel<true><<<grid, N_THREADS, smem, ctx->stream>>>(p);
  else umma_conv_kernel<false><<<grid, N_THREADS, smem, ctx->stream>>>(p);
  MG_CHECK_LAUNCH(ctx);
  ctx->tc_launches++;
  return MG_OK;
}

static int launch_halo(mg_ctx* ctx, UParams& p, const Geometry& g, int algo) {
  p.Wp = p.W + 1; p.Hp = p.H + 1;
  p.T = (int64_t)p.Nimg * p.Hp * p.Wp;
  p.n_chunks = g.n_chunks;
  const int b_stage = p.n_tile * 128;
  static int budget_env = -1, mt_env = -1, cl_env = -1, want_tl = -1;
  if (budget_env < 0) { const char* e = getenv("MGCONV_HALO_SMEM_KB"); budget_env = e ? atoi(e) : 0; }
  if (mt_env < 0) { const char* e = getenv("MGCONV_MT"); mt_env = e ? atoi(e) : 0; }
  // measured: multicast clusters couple the CTAs and lose 5-10 % (weights are not the bottleneck): off by default
  if (cl_env < 0) { const char* e = getenv("MGCONV_CLUSTER"); cl_env = e ? atoi(e) : 1; }
  if (want_tl < 0) { const char* e = getenv("MGCONV_TIMELINE"); want_tl = e ? atoi(e) : 0; }
  int CL = algo == MG_ALGO_TILE128_MCAST2 ? 2 : ((algo == MG_ALGO_AUTO || algo == MG_ALGO_RESIDENT) ? cl_env : 1);
  if ((p.n_tile * 128 / 16) % CL != 0 || CL < 1) CL = 1;          // each slice must be whole 16-byte units
  if (CL != 1 && CL != 2 && CL != 4) CL = 1;
  // two sub-tiles per CTA (half the weight stream per row) whenever that still leaves about a CTA per SM
  int MT = 1;
  if (CL == 1 && !want_tl) {
    const int64_t ctas2 = mg_cdiv(p.T, 2 * BM) * g.n_tiles;
    const int want = (algo == MG_ALGO_TILE128 || algo == MG_ALGO_TILE128_DEEP) ? 1
                     : ((algo == MG_ALGO_TILE256 || algo == MG_ALGO_TILE256_DEEP) ? 2 : (ctx->tune_mt ? ctx->tune_mt : mt_env));
    // heuristic (scratch/conv_bench.py on R-MG-34): sharing the weight stages pays off for narrow column tiles with a long K
    // loop (224 -> 64 at 14x14: 59 -> 44 us); wide tiles lose more to the shallower weight ring and the lower CTA count
    const bool heur2 = p.n_tile <= 64 && g.n_chunks >= 3 && ctas2 * 10 >= (int64_t)ctx->num_sms * 8;
    MT = (want == 1 || want == 2) ? want : (heur2 ? 2 : 1);
  }
  p.HL = BM * MT + 2 * p.Wp + 2;
  p.halo_bytes = mg_round_up(p.HL * 128, 1024);
  int cols = 32;
  while (cols < MT * p.n_tile) cols <<= 1;
  p.tmem_cols = cols;
  int budget_kb;
  if (MT == 1) {
    // measured on R-MG-34 (scratch/conv_bench.py): narrow tiles are bound by per-CTA latency chains and want
    // four resident CTAs (54 KB each, one halo buffer); N >= 192 tiles prefer two CTAs with a deeper ring
    budget_kb = p.n_tile >= 192 ? 108 : 54;
  } else {
    // TMEM allows 512 / cols CTAs per SM; at most three, so that the weight ring is at least four stages deep
    const int ctas = std::max(1, std::min(512 / cols, 3));
    budget_kb = ctas == 1 ? 200 : (ctas == 2 ? 108 : 71);
  }
  // "deep" variants trade resident CTAs for two halo buffers and a longer weight ring (more bytes in flight per CTA)
  if (algo == MG_ALGO_TILE128_DEEP) budget_kb = 108;
  if (algo == MG_ALGO_TILE256_DEEP) budget_kb = 200;
  if (budget_env > 0) budget_kb = budget_env;
  // two halo buffers when several chunks follow each other and the budget allows, else one
  const int min_ring = MT == 1 ? 2 : 3;
  p.n_abuf = (g.n_chunks > 1 && 2 * p.halo_bytes + min_ring * b_stage <= budget_kb * 1024) ? 2 : 1;
  int S = (budget_kb * 1024 - p.n_abuf * p.halo_bytes) / b_stage;
  S = std::max(2, std::min(S, MAX_STAGES));
  p.stages = S; p.lag = 1;
  static bool attr_set = false;
  if (!attr_set) {
    const int mx = 227 * 1024 - 8 * 1024;
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_conv_halo_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_conv_halo_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_conv_halo_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_conv_halo_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    attr_set = true;
  }
  const int smem = p.n_abuf * p.halo_bytes + S * b_stage + 1024;
  MG_REQUIRE(ctx, smem <= 227 * 1024 - 8 * 1024, MG_ERR_UNSUPPORTED, "halo conv: %d bytes of shared memory", smem);
  dim3 grid((unsigned)mg_round_up((int)mg_cdiv(p.T, BM * MT), CL), (unsigned)g.n_tiles);   // padding CTAs see only invalid slots
  p.timeline = nullptr;
  if (want_tl) {   // debug: dump per-CTA phase stamps of this launch to $MGCONV_TIMELINE_FILE after it ran
    static long long* d_tl = nullptr; static size_t cap = 0;
    const size_t need = (size_t)grid.x * grid.y * 8;
    if (cap < need) { if (d_tl) cudaFree(d_tl); cudaMalloc(&d_tl, need * sizeof(long long)); cap = need; }
    cudaMemsetAsync(d_tl, 0, need * sizeof(long long), ctx->stream);
    p.timeline = d_tl;
    umma_conv_halo_kernel<1, 1><<<grid, H_THREADS, smem, ctx->stream>>>(p);
    cudaStreamSynchronize(ctx->stream);
    std::vector<long long> h(need);
    cudaMemcpy(h.data(), d_tl, need * sizeof(long long), cudaMemcpyDeviceToHost);
    const char* fn = getenv("MGCONV_TIMELINE_FILE");
    FILE* f = fopen(fn ? fn : "timeline.csv", "w");
    if (f) { for (size_t i = 0; i < need / 8; ++i) fprintf(f, "%lld,%lld,%lld,%lld,%lld,%lld,%lld,%lld\n", h[i*8], h[i*8+1], h[i*8+2], h[i*8+3], h[i*8+4], h[i*8+7], h[i*8+5], h[i*8+6]); fclose(f); }
    MG_CHECK_LAUNCH(ctx);
    ctx->tc_launches++;
    return MG_OK;
  }
  if (CL == 1) {
    if (MT == 2) MG_CUDA(ctx, mg_launch_pdl(umma_conv_halo_kernel<1, 2>, grid, dim3(H_THREADS), (size_t)smem, ctx->stream, p));
    else MG_CUDA(ctx, mg_launch_pdl(umma_conv_halo_kernel<1, 1>, grid, dim3(H_THREADS), (size_t)smem, ctx->stream, p));
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(H_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (CL == 2) MG_CUDA(ctx, cudaLaunchKernelEx(&cfg, umma_conv_halo_kernel<2, 1>, p));
    else MG_CUDA(ctx, cudaLaunchKernelEx(&cfg, umma_conv_halo_kernel<4, 1>, p));
  }
  MG_CHECK_LAUNCH(ctx);
  ctx->tc_launches++;
  return MG_OK;
}

// shared-memory plan of the persistent kernel, or false when the weights do not fit / the launch is too small to pay off
struct PersistPlan { int n_abuf, smem, tmem_cols, grid; };
static bool persist_plan(const mg_ctx* ctx, const mg_conv_desc* d, const Geometry& g, int Nimg, int algo, PersistPlan* pl) {
  static int on = -1, min_tiles_per_sm = -1;
  if (on < 0) { const char* e = getenv("MGCONV_PERSIST"); on = e ? atoi(e) : 1; }
  if (min_tiles_per_sm < 0) { const char* e = getenv("MGCONV_PERSIST_MIN_TILES"); min_tiles_per_sm = e ? atoi(e) : 4; }
  if (algo == MG_ALGO_TILE128 || algo == MG_ALGO_TILE256 || algo == MG_ALGO_TILE128_DEEP || algo == MG_ALGO_TILE256_DEEP ||
      algo == MG_ALGO_TILE128_MCAST2) return false;
  const bool forced = algo == MG_ALGO_RESIDENT || ctx->tune_persist == 1;
  if (!forced && (ctx->tune_persist == 2 || !on)) return false;
  if (!g.halo || g.n_tiles != 1) return false;
  const int Wp = d->W + 1, Hp = d->H + 1;
  const int HL = BM + 2 * Wp + 2;
  const int halo_bytes = mg_round_up(HL * 128, 1024);
  const int64_t m_tiles = mg_cdiv((int64_t)Nimg * Hp * Wp, BM);
  if (!forced && m_tiles < (int64_t)min_tiles_per_sm * ctx->num_sms) return false;
  // heuristic (scratch/conv_bench.py on R-MG-34): pays off when a tile is a single chunk (the loaders then run a whole
  // tile ahead of the tensor core); two-chunk tiles are left to the autotuner
  if (!forced && g.n_chunks > 1) return false;
  const int b_bytes = g.n_chunks * 9 * g.n_tile * 128;
  const int tail = 9 * g.n_tile * 4;                       // bias tile + statistics slots
  const int budget = 232448 - 9 * 1024 - 1024;             // 227 KB per CTA minus static shared memory and alignment slack
  const int room = budget - b_bytes - tail;
  int n_abuf = room / halo_bytes;
  if (n_abuf < 2) return false;
  n_abuf = std::min(n_abuf, std::min(P_MAX_ABUF, 2 * g.n_chunks + 1));
  int cols = 32;
  while (cols < 2 * g.n_tile) cols <<= 1;                  // two accumulators
  if (cols > 512) return false;
  pl->n_abuf = n_abuf; pl->smem = n_abuf * halo_bytes + b_bytes + tail + 1024; pl->tmem_cols = cols;
  pl->grid = (int)std::min<int64_t>(m_tiles, ctx->num_sms);
  return true;
}

static int launch_halo_persistent(mg_ctx* ctx, UParams& p, const Geometry& g, const PersistPlan& pl) {
  p.Wp = p.W + 1; p.Hp = p.H + 1;
  p.T = (int64_t)p.Nimg * p.Hp * p.Wp;
  p.HL = BM + 2 * p.Wp + 2;
  p.n_chunks = g.n_chunks;
  p.halo_bytes = mg_round_up(p.HL * 128, 1024);
  p.m_tiles = (int)mg_cdiv(p.T, BM); p.n_ntiles = 1; p.n_items = p.m_tiles;
  p.tmem_cols = pl.tmem_cols;
  p.n_abuf = pl.n_abuf;
  p.stages = 0; p.lag = 1;
  static bool attr_set = false;
  if (!attr_set) {
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_conv_halo_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 9 * 1024));
    attr_set = true;
  }
  p.timeline = nullptr;
  MG_CUDA(ctx, mg_launch_pdl(umma_conv_halo_persistent_kernel, dim3(pl.grid), dim3(P_THREADS), (size_t)pl.smem, ctx->stream, p));
  MG_CHECK_LAUNCH(ctx);
  ctx->tc_launches++;
  return MG_OK;
}

static bool fused_stats_on() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MGCONV_FUSED_STATS"); on = e ? atoi(e) : 1; }
  return on != 0;
}

}  // namespace

bool umma_wgrad_supported(const mg_ctx* ctx, const mg_conv_desc* d);

bool umma_conv_supported(const mg_ctx* ctx, const mg_conv_desc* d, int kind) {
  if (ctx->dtype != MG_BF16) return false;
  if (kind == 2) return umma_wgrad_supported(ctx, d);
  if (kind == 1 && d->stride != 1) return false;     // dgrad of the strided stem is never needed
  if (d->n_seg < 1 || d->n_seg > MG_MAX_SEG) return false;
  for (int s = 0; s < d->n_seg; ++s) {
    const mg_grid& g = d->seg[s];
    if (d->seg_mode[s] == MG_SEG_POOL) return false;  // gathers are pure copies: POOL operands come as pooled companions
    if (g.scale || g.shift) return false;             // pending affines are materialised by the apply pass
    if (g.Cp % 8) return false;
    if (g.H > 1023 || g.W > 1023 || g.N > 4095) return false;
  }
  return d->H <= 1023 && d->W <= 1023;
}

size_t umma_packed_bytes(const mg_conv_desc* d, int transposed) {
  if (!d || d->n_seg < 1 || d->n_seg > MG_MAX_SEG) return 0;
  if (transposed && d->stride != 1) return 0;
  for (int s = 0; s < d->n_seg; ++s)
    if (d->seg_mode[s] == MG_SEG_POOL || d->seg[s].scale || d->seg[s].Cp % 8) return 0;
  Geometry g = geometry(d, transposed);
  return (size_t)g.n_tiles * g.n_stages * g.n_tile * 128;
}

static PackParams make_pack_params(const mg_conv_desc* d, const float* w, void* wpack, int transposed, int64_t* total) {
  Geometry g = geometry(d, transposed);
  PackParams p;
  memset(&p, 0, sizeof(p));
  p.w = w; p.out = (uint8_t*)wpack; p.transposed = transposed;
  p.k = d->ksize; p.Cout = d->Cout; p.n_seg = d->n_seg;
  int c = 0, cp = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    p.seg_C[s] = d->seg[s].C; p.seg_cbegin[s] = c; p.seg_cpbegin[s] = cp; p.seg_kvbegin[s] = cp / 8;
    c += d->seg[s].C; cp += d->seg[s].Cp;
  }
  p.Ccat = c;
  p.kv_per_tap = g.kv_per_tap; p.nkv = g.nkv; p.n_stages = g.n_stages; p.n_tile = g.n_tile; p.n_tiles = g.n_tiles;
  p.n_rows_valid = g.n_rows; p.halo = g.halo;
  *total = (int64_t)g.n_tiles * g.n_tile * g.n_stages * KV_PER_STAGE;
  return p;
}

int umma_pack_weights(mg_ctx* ctx, const mg_conv_desc* d, const float* w, void* wpack, int transposed) {
  int64_t total = 0;
  PackParams p = make_pack_params(d, w, wpack, transposed, &total);
  mg_launch_pdl(pack_weights_kernel, dim3((unsigned)mg_cdiv(total, 256)), dim3(256), 0, ctx->stream, p);
  MG_CHECK_LAUNCH(ctx);
  return MG_OK;
}

// The job table lives in device memory owned by the context and is re-uploaded only when its contents change
// (plans are static: one upload, then every step is a single launch -- and nothing but that launch is seen by a
// CUDA-graph capture of the step).
int umma_pack_weights_batched(mg_ctx* ctx, int n, const mg_conv_desc* const* descs, const float* const* w, void* const* wpack,
                              const int32_t* transposed) {
  static_assert(sizeof(PackParams) % sizeof(int) == 0, "PackParams is copied as ints");
  std::vector<PackParams> jobs((size_t)n);
  std::vector<int> blk((size_t)n + 1);
  int64_t nb = 0;
  for (int j = 0; j < n; ++j) {
    MG_REQUIRE(ctx, descs[j] && w[j] && wpack[j], MG_ERR_INVALID_ARG, "pack_weights_batched: job %d has a null pointer", j);
    int64_t total = 0;
    jobs[j] = make_pack_params(descs[j], w[j], wpack[j], transposed[j], &total);
    blk[j] = (int)nb;
    nb += mg_cdiv(total, 256);
    MG_REQUIRE(ctx, nb < ((int64_t)1 << 31), MG_ERR_UNSUPPORTED, "pack_weights_batched: too many blocks");
  }
  blk[n] = (int)nb;
  const size_t jb = (size_t)n * sizeof(PackParams), bb = ((size_t)n + 1) * sizeof(int), bytes = jb + bb;
  std::vector<uint8_t> image(bytes);
  memcpy(image.data(), jobs.data(), jb);
  memcpy(image.data() + jb, blk.data(), bb);
  if (ctx->pack_bytes != bytes || !ctx->pack_host || memcmp(ctx->pack_host, image.data(), bytes) != 0) {
    if (ctx->pack_cap < bytes) {
      MG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (ctx->pack_dev) cudaFree(ctx->pack_dev);
      free(ctx->pack_host);
      ctx->pack_dev = nullptr; ctx->pack_host = nullptr; ctx->pack_cap = 0; ctx->pack_bytes = 0;
      MG_CUDA(ctx, cudaMalloc(&ctx->pack_dev, bytes));
      ctx->pack_host = malloc(bytes);
      MG_REQUIRE(ctx, ctx->pack_host != nullptr, MG_ERR_INVALID_ARG, "pack_weights_batched: out of host memory");
      ctx->pack_cap = bytes;
    }
    memcpy(ctx->pack_host, image.data(), bytes);
    ctx->pack_bytes = bytes;
    // pageable source: staged by the driver before the call returns, ordered on the stream after earlier readers
    MG_CUDA(ctx, cudaMemcpyAsync(ctx->pack_dev, ctx->pack_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  const PackParams* djobs = (const PackParams*)ctx->pack_dev;
  const int* dblk = (const int*)((const uint8_t*)ctx->pack_dev + jb);
  if (nb > 0) {
    MG_CUDA(ctx, mg_launch_pdl(pack_weights_batched_kernel, dim3((unsigned)nb), dim3(256), 0, ctx->stream, djobs, dblk, n));
    MG_CHECK_LAUNCH(ctx);
  }
  return MG_OK;
}

int umma_conv_forward(mg_ctx* ctx, const mg_conv_desc* d, const void* wpack, const float* bias, mg_grid* y, double* bn_sums) {
  Geometry g = geometry(d, 0);
  UParams p;
  memset(&p, 0, sizeof(p));
  p.n_seg = d->n_seg;
  int cp = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    const mg_grid& sg = d->seg[s];
    const int m = d->seg_mode[s];
    if (m == MG_SEG_SAME) MG_REQUIRE(ctx, sg.H == d->H && sg.W == d->W, MG_ERR_SHAPE, "conv: SAME seg %d is %dx%d, expected %dx%d", s, sg.H, sg.W, d->H, d->W);
    else MG_REQUIRE(ctx, sg.H * 2 == d->H && sg.W * 2 == d->W, MG_ERR_SHAPE, "conv: UP seg %d is %dx%d, x2 != %dx%d", s, sg.H, sg.W, d->H, d->W);
    MG_REQUIRE(ctx, sg.N == d->seg[0].N, MG_ERR_SHAPE, "conv: seg %d batch", s);
    p.seg[s].ptr = (const __nv_bfloat16*)sg.data; p.seg[s].Hs = sg.H; p.seg[s].Ws = sg.W; p.seg[s].Cp = sg.Cp;
    p.seg[s].shift = m == MG_SEG_UP ? 1 : 0; p.seg[s].kv_begin = cp / 8;
    cp += sg.Cp;
  }
  p.k = d->ksize; p.stride = d->stride; p.pad = d->pad; p.H = d->H; p.W = d->W;
  p.Ho = y->H; p.Wo = y->W; p.Nimg = y->N;
  p.M = (int64_t)y->N * y->H * y->W;
  p.kv_per_tap = g.kv_per_tap; p.nkv = g.nkv; p.n_stages = g.n_stages; p.n_tile = g.n_tile;
  p.wpack = (const uint8_t*)wpack; p.bias = bias; p.c_bias = d->Cout;
  p.y = (__nv_bfloat16*)y->data; p.y_pitch = y->Cp; p.c_valid = y->Cp;
  // the halo kernel reduces the BatchNorm statistics in its epilogue; the other kernels are followed by the statistics pass
  const bool fused_stats = bn_sums && g.halo && fused_stats_on();
  if (fused_stats) { p.stats = bn_sums; p.c_stats = d->Cout; }
  PersistPlan pl;
  int rc = g.halo ? (persist_plan(ctx, d, g, p.Nimg, d->algo_fwd, &pl) ? launch_halo_persistent(ctx, p, g, pl) : launch_halo(ctx, p, g, d->algo_fwd))
                  : launch(ctx, p, g.n_tiles);
  if (rc) return rc;
  if (bn_sums && !fused_stats) return mg_bn_stats(ctx, y, bn_sums);
  return MG_OK;
}

int umma_conv_backward_data(mg_ctx* ctx, const mg_conv_desc* d, const void* wpack_t, const mg_grid* gr, mg_grid* dcat) {
  Geometry g = geometry(d, 1);
  UParams p;
  memset(&p, 0, sizeof(p));
  p.n_seg = 1;
  p.seg[0].ptr = (const __nv_bfloat16*)gr->data; p.seg[0].Hs = gr->H; p.seg[0].Ws = gr->W; p.seg[0].Cp = gr->Cp;
  p.seg[0].shift = 0; p.seg[0].kv_begin = 0;
  MG_REQUIRE(ctx, gr->Cp == mg_round_up(d->Cout, 8), MG_ERR_SHAPE, "dgrad: g.Cp %d", gr->Cp);
  p.k = d->ksize; p.stride = 1; p.pad = d->ksize - 1 - d->pad; p.H = gr->H; p.W = gr->W;
  p.Ho = dcat->H; p.Wo = dcat->W; p.Nimg = dcat->N;
  p.M = (int64_t)dcat->N * dcat->H * dcat->W;
  p.kv_per_tap = g.kv_per_tap; p.nkv = g.nkv; p.n_stages = g.n_stages; p.n_tile = g.n_tile;
  p.wpack = (const uint8_t*)wpack_t; p.bias = nullptr;
  p.y = (__nv_bfloat16*)dcat->data; p.y_pitch = dcat->Cp; p.c_valid = dcat->Cp;
  MG_REQUIRE(ctx, dcat->Cp == g.n_rows, MG_ERR_SHAPE, "dgrad: dcat.Cp %d != %d", dcat->Cp, g.n_rows);
  PersistPlan pl;
  return g.halo ? (persist_plan(ctx, d, g, p.Nimg, d->algo_bwd_data, &pl) ? launch_halo_persistent(ctx, p, g, pl) : launch_halo(ctx, p, g, d->algo_bwd_data))
                : launch(ctx, p, g.n_tiles);
}

