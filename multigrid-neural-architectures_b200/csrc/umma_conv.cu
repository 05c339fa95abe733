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
      int ox = (int)(m % p.Wo); int64_t q = m / p.Wo;
      int oy = (int)(q % p.Ho); int n = (int)(q / p.Ho);
      v = ((uint32_t)n << 20) | ((uint32_t)oy << 10) | (uint32_t)ox;
    }
    s_row[tid] = v;
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
      for (int ks = 0; ks < p.n_stages + L; ++ks) {
        if (ks < p.n_stages) {
          const int s = ks % S;
          if (ks >= S) mbar_wait(&empty_bar[s], ((ks / S) - 1) & 1);
          const int j = ks * KV_PER_STAGE + v;
          const bool kv_ok = j < p.nkv;
          int tap = 0, r = 0, sg = 0;
          if (kv_ok) {
            tap = j / p.kv_per_tap; r = j - tap * p.kv_per_tap;
            while (sg + 1 < p.n_seg && r >= s_seg[sg + 1].kv_begin) ++sg;
          }
          const USeg sgm = s_seg[sg];
          const int c8 = r - sgm.kv_begin;
          const int dy = tap / p.k - p.pad, dx = tap % p.k - p.pad;
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
        }
        cp_async_commit();
        if (ks >= L) {
          cp_async_wait_dyn(L);
          fence_proxy_async();
          mbar_arrive(&full_bar[(ks - L) % S]);
        }
      }
    } else {
      for (int ks = 0; ks < p.n_stages + L; ++ks) {
        if (ks < p.n_stages) {
          const int s = ks % S;
          if (ks >= S) mbar_wait(&empty_bar[s], ((ks / S) - 1) & 1);
          // decode this lane's k-vector: (tap, segment, channel offset)
          const int j = ks * KV_PER_STAGE + v;
          const bool kv_ok = j < p.nkv;
          int tap = 0, r = 0, sg = 0;
          if (kv_ok) {
            tap = j / p.kv_per_tap; r = j - tap * p.kv_per_tap;
            while (sg + 1 < p.n_seg && r >= s_seg[sg + 1].kv_begin) ++sg;
          }
          const USeg sgm = s_seg[sg];
          const int c8 = r - sgm.kv_begin;
          const int dy = tap / p.k - p.pad, dx = tap % p.k - p.pad;
          const uint32_t dst0 = smem_u32(a_smem + (size_t)s * A_STAGE_BYTES);
#pragma unroll
          for (int it = 0; it < BM / 16; ++it) {
            const int row = it * 16 + rg;
            const uint32_t ri = s_row[row];
            const int ox = ri & 1023, oy = (ri >> 10) & 1023, n = ri >> 20;
            const int iy = oy * p.stride + dy, ix = ox * p.stride + dx;
            const bool ok = kv_ok && ri != 0xFFFFFFFFu && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
            const __nv_bfloat16* src = sgm.ptr;
            if (ok) src += ((size_t)((size_t)n * sgm.Hs + (iy >> sgm.shift)) * sgm.Ws + (ix >> sgm.shift)) * sgm.Cp + c8 * 8;
            cp_async16(dst0 + row * 128 + ((v ^ (row & 7)) << 4), src, ok ? 16u : 0u);
          }
        }
        cp_async_commit();
        if (ks >= L) {  // stage ks-L has landed for this thread: publish it to the tensor core (async proxy)
          cp_async_wait_dyn(L);
          fence_proxy_async();
          mbar_arrive(&full_bar[(ks - L) % S]);
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
              float a = __uint_as_float(acc[h * 8 + 2 * e]), b = __uint_as_float(acc[h * 8 + 2 * e + 1]);
              if (p.bias) {
                if (n0 + 2 * e < p.c_bias) a += __ldg(p.bias + n0 + 2 * e);
                if (n0 + 2 * e + 1 < p.c_bias) b += __ldg(p.bias + n0 + 2 * e + 1);
              }
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
      for (int ks = 0; ks < p.n_stages; ++ks) {
        const int s = ks % S;
        if (ks >= S) mbar_wait(&empty_bar[s], ((ks / S) - 1) & 1);
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)b_stage_bytes);
        bulk_g2s(smem_u32(b_smem + (size_t)s * b_stage_bytes), wsrc + (size_t)ks * b_stage_bytes, (uint32_t)b_stage_bytes, &full_bar[s]);
      }
    }
  } else {
    // ================= MMA issuer ================================================================
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_m128(p.n_tile);
      for (int ks = 0; ks < p.n_stages; ++ks) {
        const int s = ks % S;
        mbar_wait(&full_bar[s], (ks / S) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(a_smem + (size_t)s * A_STAGE_BYTES);
        const uint32_t b_addr = smem_u32(b_smem + (size_t)s * b_stage_bytes);
        const int kv_here = min(KV_PER_STAGE, p.nkv - ks * KV_PER_STAGE);
        const int ksteps = (kv_here + 1) >> 1;   // 16 bf16 = 2 k-vectors per UMMA K step
        for (int q = 0; q < ksteps; ++q)
          tc_mma_bf16(tmem_base, smem_desc_k_sw128(a_addr + q * 32), smem_desc_k_sw128(b_addr + q * 32), idesc, (ks | q) != 0);
        tc_commit(&empty_bar[s]);   // frees the stage once these MMAs have read it
      }
      tc_commit(&tmem_full_bar);    // accumulator complete
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
constexpr int HALO_MAX_SLOTS = 128 + 2 * 64 + 2;   // W <= 63
constexpr int H_PROD = 256;                        // 8 loader warps (the first 4 also drain TMEM): the gather is issue bound
constexpr int H_MMA_WARP = H_PROD / 32, H_B_WARP = H_MMA_WARP + 1;
constexpr int H_THREADS = H_PROD + 64;

// CL = thread-block cluster size along the tile index: the CL CTAs of a cluster need the same weight stages, so each
// loads 1/CL of a stage and multicasts it to all of them (L2 -> SM weight traffic / CL)
template <int CL>
__global__ void __launch_bounds__(H_THREADS, 2) umma_conv_halo_kernel(const __grid_constant__ UParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], a_full[2], a_empty[2], tmem_full_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t s_pix[HALO_MAX_SLOTS];   // full-resolution pixel index of a halo slot, 0xFFFFFFFF = padding / outside
  __shared__ uint32_t s_pup[HALO_MAX_SLOTS];   // half-resolution pixel index (UP segments)
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
  const int64_t t0 = (int64_t)blockIdx.x * BM;
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
        const int n = (int)(t / slots_per_img); const int rem = (int)(t - (int64_t)n * slots_per_img);
        const int yy = rem / p.Wp, xs = rem - yy * p.Wp;
        if (yy < p.H && xs < p.W) {
          pix = (uint32_t)(((size_t)n * p.H + yy) * p.W + xs);
          pup = (uint32_t)(((size_t)n * Hs2 + (yy >> 1)) * Ws2 + (xs >> 1));
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
    for (int c = 0; c < p.n_chunks + lagA; ++c) {
      if (c < p.n_chunks) {
        const int buf = c % NB;
        if (c >= NB) mbar_wait(&a_empty[buf], ((c / NB) - 1) & 1);
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
        mbar_arrive(&a_full[(c - lagA) % NB]);
      }
    }
    // ================= epilogue (warps 0-3): TMEM -> registers -> bf16 NHWC rows ===========
    if (warp < 4) {
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    if (tl && tid == 0) tl[3] = clock64();
    const int row = warp * 32 + lane;
    const uint32_t pix = s_pix[row + p.Wp + 1];          // slot t0 + row
    const bool row_ok = pix != 0xFFFFFFFFu;
    const int n_base = ntile * p.n_tile;
    __nv_bfloat16* yrow = p.y + (size_t)(row_ok ? pix : 0) * p.y_pitch;
    for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
      uint32_t acc[16];
      tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, acc);
      tc_wait_ld();
      if (row_ok) {
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
    if (tl && tid == 0) tl[4] = clock64();
    }
  } else if (warp == H_B_WARP) {
    // ================= B loader: one bulk copy per (chunk, tap) stage ============================
    if (lane == 0) {
      const int n_st = p.n_chunks * KK;
      const uint8_t* wsrc = p.wpack + (size_t)ntile * n_st * b_stage_bytes;
      const uint32_t rank = CL > 1 ? cluster_ctarank() : 0;
      const uint32_t slice = (uint32_t)b_stage_bytes / CL;   // n_tile * 128 / CL: a multiple of 16 bytes
      for (int ks = 0; ks < n_st; ++ks) {
        const int s = ks % S;
        if (ks >= S) mbar_wait(&empty_bar[s], ((ks / S) - 1) & 1);   // CL arrivals: every CTA of the cluster freed slot s
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
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_m128(p.n_tile);
      int ks = 0;
      long long wait_a = 0, wait_b = 0, tq = 0;
      for (int c = 0; c < p.n_chunks; ++c) {
        const int buf = c % p.n_abuf;
        if (tl) tq = clock64();
        mbar_wait(&a_full[buf], (c / p.n_abuf) & 1);
        tc_fence_after();
        if (tl && c == 0) tl[2] = clock64();
        if (tl && c > 0) wait_a += clock64() - tq;
        const uint32_t a_base = smem_u32(a_smem + (size_t)buf * p.halo_bytes);
        const int kv_here = min(KV_PER_STAGE, p.kv_per_tap - c * KV_PER_STAGE);
        const int ksteps = (kv_here + 1) >> 1;
        for (int tap = 0; tap < KK; ++tap, ++ks) {
          const int s = ks % S;
          if (tl) tq = clock64();
          mbar_wait(&full_bar[s], (ks / S) & 1);
          tc_fence_after();
          if (tl) wait_b += clock64() - tq;
          const uint32_t a_addr = a_base + (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * 128u;
          const uint32_t b_addr = smem_u32(b_smem + (size_t)s * b_stage_bytes);
          for (int q = 0; q < ksteps; ++q)
            tc_mma_bf16(tmem_base, smem_desc_k_sw128(a_addr + q * 32), smem_desc_k_sw128(b_addr + q * 32), idesc, (ks | q) != 0);
          if (CL == 1) tc_commit(&empty_bar[s]); else tc_commit_mcast(&empty_bar[s], (uint16_t)((1u << CL) - 1));
        }
        tc_commit(&a_empty[buf]);   // halo buffer free once this chunk's MMAs have read it
      }
      tc_commit(&tmem_full_bar);
      if (tl) { tl[5] = wait_a; tl[6] = wait_b; }
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
// Same operand scheme as umma_conv_halo_kernel, but one CTA walks a strided list of (slot tile, column tile)
// work items and its roles run decoupled, so that nothing of a tile's latency chain is exposed:
//   8 loader warps   stage halos into a ring that runs ACROSS chunks and tiles (the next tile's halo is in
//                    flight while the tensor core works on the current one),
//   1 B-loader thread streams the (chunk, tap) weight stages through its own ring,
//   1 MMA thread     accumulates tile i into TMEM accumulator i & 1,
//   4 epilogue warps drain accumulator (i-1) & 1 meanwhile (TMEM double buffering).
constexpr int P_LOAD = 256, P_EPI_WARP0 = 8, P_MMA_WARP = 12, P_B_WARP = 13, P_THREADS = 448;
constexpr int P_MAX_ABUF = 4;

__global__ void __launch_bounds__(P_THREADS, 1) umma_conv_halo_persistent_kernel(const __grid_constant__ UParams p) {
  pdl_launch();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], a_full[P_MAX_ABUF], a_empty[P_MAX_ABUF], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t s_pix[2][HALO_MAX_SLOTS], s_pup[2][HALO_MAX_SLOTS];
  __shared__ float s_bias[2][256];
  __shared__ USeg s_seg[MG_MAX_SEG];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.stages, NA = p.n_abuf;
  const int b_stage_bytes = p.n_tile * 128;
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + (size_t)NA * p.halo_bytes;
  const int KK = 9;
  const int n_my = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // items blockIdx.x, +gridDim.x, ...

  if (tid < p.n_seg) s_seg[tid] = p.seg[tid];
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], P_LOAD); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == P_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int slots_per_img = p.Hp * p.Wp;

  if (warp < P_EPI_WARP0) {
    // ================= loaders ====================================================================
    const int v = tid & 7, rg = tid >> 3;          // k-vector column, slot lane 0..31
    const int Hs2 = p.H >> 1, Ws2 = p.W >> 1;
    const int lagA = NA - 1;
    int ac = 0;                                      // chunks issued so far (ring position)
    const int total_chunks = n_my * p.n_chunks;
    for (int it = 0; it < n_my; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int mt = item % p.m_tiles;
      const int tb = it & 1;
      const int64_t t0 = (int64_t)mt * BM;
      for (int h = tid; h < p.HL; h += P_LOAD) {
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
        s_pix[tb][h] = pix; s_pup[tb][h] = pup;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(P_LOAD) : "memory");
      for (int c = 0; c < p.n_chunks; ++c, ++ac) {
        const int buf = ac % NA;
        if (ac >= NA) mbar_wait(&a_empty[buf], ((ac / NA) - 1) & 1);
        const int r = c * KV_PER_STAGE + v;
        const bool kv_ok = r < p.kv_per_tap;
        int sg = 0;
        if (kv_ok) while (sg + 1 < p.n_seg && r >= s_seg[sg + 1].kv_begin) ++sg;
        const USeg sgm = s_seg[sg];
        const uint32_t pitch = (uint32_t)sgm.Cp * 2u;
        const char* base = reinterpret_cast<const char*>(sgm.ptr) + (r - sgm.kv_begin) * 16;
        const uint32_t* tab = sgm.shift ? s_pup[tb] : s_pix[tb];
        const uint32_t dst0 = smem_u32(a_smem + (size_t)buf * p.halo_bytes) + (uint32_t)(v << 4);
        for (int h0 = rg; h0 < p.HL; h0 += 4 * (P_LOAD / 8)) {
          uint32_t pv[4], tv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int h = h0 + u * (P_LOAD / 8);
            pv[u] = h < p.HL ? s_pix[tb][h] : 0xFFFFFFFFu;
            tv[u] = h < p.HL ? tab[h] : 0u;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int h = h0 + u * (P_LOAD / 8);
            if (h < p.HL) {
              const bool ok = kv_ok && pv[u] != 0xFFFFFFFFu;
              const char* src = ok ? base + (uint64_t)tv[u] * pitch : reinterpret_cast<const char*>(sgm.ptr);
              cp_async16((dst0 ^ ((uint32_t)(h & 7) << 4)) + (uint32_t)h * 128, src, ok ? 16u : 0u);
            }
          }
        }
        cp_async_commit();
        if (ac >= lagA) {
          cp_async_wait_dyn(lagA);
          fence_proxy_async();
          mbar_arrive(&a_full[(ac - lagA) % NA]);
        }
      }
    }
    for (int q = max(0, total_chunks - lagA); q < total_chunks; ++q) {   // publish the chunks still in flight
      cp_async_wait_dyn(total_chunks - 1 - q);
      fence_proxy_async();
      mbar_arrive(&a_full[q % NA]);
    }
  } else if (warp < P_MMA_WARP) {
    // ================= epilogue warps: drain accumulator it & 1 ====================================
    const int ew = warp - P_EPI_WARP0, et = tid - P_EPI_WARP0 * 32;   // TMEM lane quarter, thread 0..127
    for (int it = 0; it < n_my; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int mt = item % p.m_tiles, ntile = item / p.m_tiles;
      const int acc = it & 1;
      for (int c = et; c < p.n_tile; c += 128) {
        const int ch = ntile * p.n_tile + c;
        s_bias[acc][c] = (p.bias && ch < p.c_bias) ? p.bias[ch] : 0.f;
      }
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
      asm volatile("bar.sync 2, 128;" ::: "memory");   // bias tile visible to the four epilogue warps
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      const int n_base = ntile * p.n_tile;
      __nv_bfloat16* yrow = p.y + (size_t)pix * p.y_pitch;
      const uint32_t tcol = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * p.n_tile);
      for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
        uint32_t a16[16];
        tc_ld16(tcol + (uint32_t)c0, a16);
        tc_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int n0 = n_base + c0 + h * 8;
            if (n0 + 8 <= p.c_valid) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float a = __uint_as_float(a16[h * 8 + 2 * e]) + s_bias[acc][c0 + h * 8 + 2 * e];
                const float b = __uint_as_float(a16[h * 8 + 2 * e + 1]) + s_bias[acc][c0 + h * 8 + 2 * e + 1];
                __nv_bfloat162 t2 = __floats2bfloat162_rn(a, b);
                pk[e] = *reinterpret_cast<uint32_t*>(&t2);
              }
              *reinterpret_cast<uint4*>(yrow + n0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);   // 128 arrivals: the accumulator may be overwritten
    }
  } else if (warp == P_B_WARP) {
    // ================= B loader ==================================================================
    if (lane == 0) {
      const int n_st = p.n_chunks * KK;
      int ks = 0;
      for (int it = 0; it < n_my; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int ntile = item / p.m_tiles;
        const uint8_t* wsrc = p.wpack + (size_t)ntile * n_st * b_stage_bytes;
        for (int q = 0; q < n_st; ++q, ++ks) {
          const int s = ks % S;
          if (ks >= S) mbar_wait(&empty_bar[s], ((ks / S) - 1) & 1);
          mbar_arrive_expect_tx(&full_bar[s], (uint32_t)b_stage_bytes);
          bulk_g2s(smem_u32(b_smem + (size_t)s * b_stage_bytes), wsrc + (size_t)q * b_stage_bytes, (uint32_t)b_stage_bytes, &full_bar[s]);
        }
      }
    }
  } else {
    // ================= MMA issuer ================================================================
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_m128(p.n_tile);
      int ks = 0, ac = 0;
      for (int it = 0; it < n_my; ++it) {
        const int acc = it & 1;
        if (it >= 2) { mbar_wait(&tmem_empty[acc], ((it >> 1) - 1) & 1); tc_fence_after(); }
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n_tile);
        for (int c = 0; c < p.n_chunks; ++c, ++ac) {
          const int buf = ac % NA;
          mbar_wait(&a_full[buf], (ac / NA) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(a_smem + (size_t)buf * p.halo_bytes);
          const int kv_here = min(KV_PER_STAGE, p.kv_per_tap - c * KV_PER_STAGE);
          const int ksteps = (kv_here + 1) >> 1;
          for (int tap = 0; tap < KK; ++tap, ++ks) {
            const int s = ks % S;
            mbar_wait(&full_bar[s], (ks / S) & 1);
            tc_fence_after();
            const uint32_t a_addr = a_base + (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * 128u;
            const uint32_t b_addr = smem_u32(b_smem + (size_t)s * b_stage_bytes);
            for (int q = 0; q < ksteps; ++q)
              tc_mma_bf16(d_tmem, smem_desc_k_sw128(a_addr + q * 32), smem_desc_k_sw128(b_addr + q * 32), idesc, (c | tap | q) != 0);
            tc_commit(&empty_bar[s]);
          }
          tc_commit(&a_empty[buf]);
        }
        tc_commit(&tmem_full[acc]);
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
__global__ void pack_weights_kernel(PackParams p) {
  pdl_launch();
  pdl_wait();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
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

// ---------------------------------------------------------------- host side -----------------
struct Geometry {
  int kv_per_tap, nkv, n_stages, n_tile, n_tiles, n_rows;
  int halo, n_chunks;
};

// the halo kernel serves 3x3 / stride 1 / pad 1 convolutions on grids of at least `MGCONV_HALO_MIN_W`
// (default 7) and at most 63 columns whose UP segments are exactly half size
static bool halo_applies(const mg_conv_desc* d) {
  static int min_w = -1;
  if (min_w < 0) { const char* e = getenv("MGCONV_HALO_MIN_W"); min_w = e ? atoi(e) : 7; }
  if (d->ksize != 3 || d->stride != 1 || d->pad != 1) return false;
  if (d->W < min_w || d->W > 63 || d->H > 1023) return false;
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

static int launch_halo(mg_ctx* ctx, UParams& p, const Geometry& g) {
  p.Wp = p.W + 1; p.Hp = p.H + 1;
  p.T = (int64_t)p.Nimg * p.Hp * p.Wp;
  p.HL = BM + 2 * p.Wp + 2;
  p.n_chunks = g.n_chunks;
  p.halo_bytes = mg_round_up(p.HL * 128, 1024);
  const int b_stage = p.n_tile * 128;
  static int budget_env = -1;
  if (budget_env < 0) { const char* e = getenv("MGCONV_HALO_SMEM_KB"); budget_env = e ? atoi(e) : 0; }
  // measured on R-MG-34 (scratch/conv_bench.py): narrow tiles are bound by per-CTA latency chains and want
  // four resident CTAs (54 KB each, one halo buffer); N >= 192 tiles are bound by the weight stream and
  // prefer two CTAs with a deeper ring
  const int budget_kb = budget_env > 0 ? budget_env : (p.n_tile >= 192 ? 108 : 54);
  // two halo buffers when several chunks follow each other and the budget allows, else one
  p.n_abuf = (g.n_chunks > 1 && 2 * p.halo_bytes + 2 * b_stage <= budget_kb * 1024) ? 2 : 1;
  int S = (budget_kb * 1024 - p.n_abuf * p.halo_bytes) / b_stage;
  S = std::max(2, std::min(S, MAX_STAGES));
  p.stages = S; p.lag = 1;
  int cols = 32;
  while (cols < p.n_tile) cols <<= 1;
  p.tmem_cols = cols;
  static bool attr_set = false;
  if (!attr_set) {
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_conv_halo_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 8 * 1024));
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_conv_halo_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 8 * 1024));
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_conv_halo_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 8 * 1024));
    attr_set = true;
  }
  static int cl_env = -1;
  // measured: multicast clusters couple the CTAs and lose 5-10 % (weights are not the bottleneck): off by default
  if (cl_env < 0) { const char* e = getenv("MGCONV_CLUSTER"); cl_env = e ? atoi(e) : 1; }
  int CL = cl_env;
  if ((p.n_tile * 128 / 16) % CL != 0 || CL < 1) CL = 1;          // each slice must be whole 16-byte units
  if (CL != 1 && CL != 2 && CL != 4) CL = 1;
  const int smem = p.n_abuf * p.halo_bytes + S * b_stage + 1024;
  dim3 grid((unsigned)mg_round_up((int)mg_cdiv(p.T, BM), CL), (unsigned)g.n_tiles);   // padding CTAs see only invalid slots
  static int want_tl = -1;
  if (want_tl < 0) { const char* e = getenv("MGCONV_TIMELINE"); want_tl = e ? atoi(e) : 0; }
  p.timeline = nullptr;
  if (want_tl) {   // debug: dump per-CTA phase stamps of this launch to $MGCONV_TIMELINE_FILE after it ran
    static long long* d_tl = nullptr; static size_t cap = 0;
    const size_t need = (size_t)grid.x * grid.y * 8;
    if (cap < need) { if (d_tl) cudaFree(d_tl); cudaMalloc(&d_tl, need * sizeof(long long)); cap = need; }
    cudaMemsetAsync(d_tl, 0, need * sizeof(long long), ctx->stream);
    p.timeline = d_tl;
    umma_conv_halo_kernel<1><<<grid, H_THREADS, smem, ctx->stream>>>(p);
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
    MG_CUDA(ctx, mg_launch_pdl(umma_conv_halo_kernel<1>, grid, dim3(H_THREADS), (size_t)smem, ctx->stream, p));
  } else {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = dim3(H_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (CL == 2) MG_CUDA(ctx, cudaLaunchKernelEx(&cfg, umma_conv_halo_kernel<2>, p));
    else MG_CUDA(ctx, cudaLaunchKernelEx(&cfg, umma_conv_halo_kernel<4>, p));
  }
  MG_CHECK_LAUNCH(ctx);
  ctx->tc_launches++;
  return MG_OK;
}

static int launch_halo_persistent(mg_ctx* ctx, UParams& p, const Geometry& g) {
  p.Wp = p.W + 1; p.Hp = p.H + 1;
  p.T = (int64_t)p.Nimg * p.Hp * p.Wp;
  p.HL = BM + 2 * p.Wp + 2;
  p.n_chunks = g.n_chunks;
  p.halo_bytes = mg_round_up(p.HL * 128, 1024);
  p.m_tiles = (int)mg_cdiv(p.T, BM); p.n_ntiles = g.n_tiles; p.n_items = p.m_tiles * p.n_ntiles;
  const int b_stage = p.n_tile * 128;
  int cols = 32;
  while (cols < 2 * p.n_tile) cols <<= 1;            // two accumulators
  p.tmem_cols = cols;
  // one persistent CTA per SM (two when both accumulators and the rings of two CTAs fit)
  static int cps_env = -1;
  if (cps_env < 0) { const char* e = getenv("MGCONV_PERSIST_CTAS"); cps_env = e ? atoi(e) : 0; }
  int cps = cps_env > 0 ? cps_env : (cols <= 256 ? 2 : 1);
  const int budget = (cps == 1 ? 200 : 104) * 1024;
  p.n_abuf = std::max(2, std::min(P_MAX_ABUF, (budget / 2) / p.halo_bytes));
  int S = (budget - p.n_abuf * p.halo_bytes) / b_stage;
  while (S < 3 && p.n_abuf > 2) { --p.n_abuf; S = (budget - p.n_abuf * p.halo_bytes) / b_stage; }
  S = std::max(2, std::min(S, MAX_STAGES));
  p.stages = S; p.lag = 1;
  static bool attr_set = false;
  if (!attr_set) {
    MG_CUDA(ctx, cudaFuncSetAttribute(umma_conv_halo_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 12 * 1024));
    attr_set = true;
  }
  const int smem = p.n_abuf * p.halo_bytes + S * b_stage + 1024;
  const int grid = std::min(p.n_items, ctx->num_sms * cps);
  p.timeline = nullptr;
  MG_CUDA(ctx, mg_launch_pdl(umma_conv_halo_persistent_kernel, dim3(grid), dim3(P_THREADS), (size_t)smem, ctx->stream, p));
  MG_CHECK_LAUNCH(ctx);
  ctx->tc_launches++;
  return MG_OK;
}

static bool persist_on() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MGCONV_PERSIST"); on = e ? atoi(e) : 0; }
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

int umma_pack_weights(mg_ctx* ctx, const mg_conv_desc* d, const float* w, void* wpack, int transposed) {
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
  const int64_t total = (int64_t)g.n_tiles * g.n_tile * g.n_stages * KV_PER_STAGE;
  mg_launch_pdl(pack_weights_kernel, dim3((unsigned)mg_cdiv(total, 256)), dim3(256), 0, ctx->stream, p);
  MG_CHECK_LAUNCH(ctx);
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
  int rc = g.halo ? (persist_on() ? launch_halo_persistent(ctx, p, g) : launch_halo(ctx, p, g)) : launch(ctx, p, g.n_tiles);
  if (rc) return rc;
  if (bn_sums) return mg_bn_stats(ctx, y, bn_sums);
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
  return g.halo ? (persist_on() ? launch_halo_persistent(ctx, p, g) : launch_halo(ctx, p, g)) : launch(ctx, p, g.n_tiles);
}

