// cudnn.SpatialFullConvolution(nIP, nOP, 2,2, 2,2) of U-MG (models/mnist-cluttered/unmg.lua:35-52) on the tensor cores.
//
//   y[n, 2y+dy, 2x+dx, co] = b[co] + sum_ci x[n, y, x, ci] * w[ci][co][dy][dx]
//
// is four 1x1 convolutions that share their input: with q = 2*dy + dx it is ONE 1x1 convolution x -> Y' with 4*CoutP output
// channels (channel q*CoutP + co, weight W1[q*CoutP+co][ci] = w[ci][co][dy][dx]) followed by a depth-to-space shuffle, and
// its backward is space-to-depth of g followed by the 1x1 dgrad / wgrad of the same descriptor.  Everything heavy runs in
// the tcgen05 kernels of umma_conv.cu / umma_wgrad.cu; this file adds the (tiny) weight re-layouts and the two shuffles.
// The CUDA-core kernels in elementwise.cu remain the fp32 path.  Measured on mnist-cluttered/unmg, B = 128: the six
// up-convolutions were 106 ms of a 113 ms training step on the CUDA cores.
#include "common.cuh"

bool umma_conv_supported(const mg_ctx*, const mg_conv_desc*, int kind);
size_t umma_packed_bytes(const mg_conv_desc*, int transposed);
int umma_pack_weights(mg_ctx*, const mg_conv_desc*, const float*, void*, int transposed);
int umma_conv_forward(mg_ctx*, const mg_conv_desc*, const void*, const float*, mg_grid*, mg_sum*);
int umma_conv_backward_data(mg_ctx*, const mg_conv_desc*, const void*, const mg_grid*, mg_grid*);
int umma_conv_backward_weight(mg_ctx*, const mg_conv_desc*, const mg_grid*, float*, float*, float);

namespace {

typedef __nv_bfloat16 bf16;

// W1[(q*CoutP + co)][ci] = w[ci][co][q] (zero rows for co >= Cout), b1[q*CoutP + co] = bias[co]
__global__ void upconv_w1_kernel(const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ W1, float* __restrict__ b1,
                                 int Cin, int Cout, int CoutP) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * CoutP * Cin) {
    const int row = i / Cin, ci = i - row * Cin;
    const int q = row / CoutP, co = row - q * CoutP;
    W1[i] = co < Cout ? w[((size_t)ci * Cout + co) * 4 + q] : 0.f;
  }
  if (i < 4 * CoutP) {
    const int co = i % CoutP;
    b1[i] = (bias && co < Cout) ? bias[co] : 0.f;
  }
}

// dw[ci][co][q] += dW1[(q*CoutP + co)][ci];  dbias[co] += sum_q db1[q*CoutP + co]   (gscale already applied by the wgrad kernels)
__global__ void upconv_dw_kernel(const float* __restrict__ dW1, const float* __restrict__ db1, float* __restrict__ dw, float* __restrict__ dbias,
                                 int Cin, int Cout, int CoutP) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Cin * Cout * 4) {
    const int q = i & 3, co = (i >> 2) % Cout, ci = (i >> 2) / Cout;
    dw[i] += dW1[((size_t)q * CoutP + co) * Cin + ci];
  }
  if (dbias && i < Cout) dbias[i] += db1[i] + db1[CoutP + i] + db1[2 * CoutP + i] + db1[3 * CoutP + i];
}

// TO_SPACE: y[n, 2y+dy, 2x+dx, c] = Yp[n, y, x, q*CoutP + c];  else Yp[...] = y[...] (space-to-depth);  16-byte vectors
template <bool TO_SPACE>
__global__ void __launch_bounds__(256) upconv_shuffle_kernel(bf16* __restrict__ yp, bf16* __restrict__ y, int N, int H, int W, int CoutP) {
  const int V = CoutP >> 3;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * H * W * 4 * V) return;
  const int v = (int)(i % V); int64_t r = i / V;
  const int q = (int)(r & 3); r >>= 2;
  const int x = (int)(r % W); r /= W;
  const int yy = (int)(r % H); const int n = (int)(r / H);
  uint4* a = reinterpret_cast<uint4*>(yp + ((((int64_t)n * H + yy) * W + x) * 4 + q) * CoutP) + v;
  uint4* b = reinterpret_cast<uint4*>(y + (((int64_t)n * 2 * H + 2 * yy + (q >> 1)) * 2 * W + 2 * x + (q & 1)) * CoutP) + v;
  if (TO_SPACE) *b = *a; else *a = *b;
}

struct Scratch { float *W1, *b1, *dW1, *db1; uint8_t *wp, *wpt; bf16* Yp; };

static inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

// carve the lane's up-convolution scratch
static int scratch(mg_ctx* ctx, const mg_conv_desc* d1, int Cin, int CoutP, int64_t pixels, Scratch* s) {
  const size_t nW = (size_t)4 * CoutP * Cin * sizeof(float), nb = (size_t)4 * CoutP * sizeof(float);
  const size_t npk = umma_packed_bytes(d1, 0), npkt = umma_packed_bytes(d1, 1);
  const size_t nY = (size_t)pixels * 4 * CoutP * sizeof(bf16);
  const size_t total = 2 * up256(nW) + 2 * up256(nb) + up256(npk) + up256(npkt) + up256(nY);
  void** ws = &ctx->up_ws[ctx->cur_lane];
  size_t* cap = &ctx->up_ws_bytes[ctx->cur_lane];
  if (*cap < total) {
    if (*ws) { cudaStreamSynchronize(ctx->stream); cudaFree(*ws); *ws = nullptr; *cap = 0; }
    if (cudaMalloc(ws, total) != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "upconv scratch: cudaMalloc(%zu) failed", total); return MG_ERR_CUDA; }
    *cap = total;
  }
  uint8_t* p = (uint8_t*)*ws;
  s->W1 = (float*)p; p += up256(nW);
  s->dW1 = (float*)p; p += up256(nW);
  s->b1 = (float*)p; p += up256(nb);
  s->db1 = (float*)p; p += up256(nb);
  s->wp = p; p += up256(npk);
  s->wpt = p; p += up256(npkt);
  s->Yp = (bf16*)p;
  return MG_OK;
}

static bool applies(const mg_ctx* ctx, const mg_grid* x, const mg_grid* y, mg_conv_desc* d1, mg_grid* Yp) {
  if (ctx->dtype != MG_BF16 || ctx->impl == MG_IMPL_SIMT) return false;
  if (x->scale || x->shift || x->Cp % 8 || y->Cp % 8 || y->scale) return false;
  memset(d1, 0, sizeof(*d1));
  d1->n_seg = 1; d1->seg[0] = *x; d1->seg_mode[0] = MG_SEG_SAME;
  d1->ksize = 1; d1->stride = 1; d1->pad = 0; d1->Cout = 4 * y->Cp; d1->H = x->H; d1->W = x->W;
  memset(Yp, 0, sizeof(*Yp));
  Yp->N = x->N; Yp->H = x->H; Yp->W = x->W; Yp->C = 4 * y->Cp; Yp->Cp = 4 * y->Cp;
  return umma_conv_supported(ctx, d1, 0) && umma_conv_supported(ctx, d1, 1) && umma_conv_supported(ctx, d1, 2);
}

}  // namespace

// return MG_OK, an error, or -1 when the tensor-core composition does not apply (caller falls back to the CUDA-core kernels)
int upconv_tc_forward(mg_ctx* ctx, const mg_grid* x, const float* w, const float* bias, mg_grid* y) {
  mg_conv_desc d1; mg_grid Yp;
  if (!applies(ctx, x, y, &d1, &Yp)) return -1;
  const int Cin = x->C, Cout = y->C, CoutP = y->Cp;
  const int64_t pixels = (int64_t)x->N * x->H * x->W;
  Scratch s;
  int rc = scratch(ctx, &d1, Cin, CoutP, pixels, &s);
  if (rc) return rc;
  // the conv descriptor addresses its input by concatenated channel: Torch-layout weight [4*CoutP][Cin][1][1] = W1
  upconv_w1_kernel<<<(unsigned)mg_cdiv((int64_t)4 * CoutP * Cin, 256), 256, 0, ctx->stream>>>(w, bias, s.W1, s.b1, Cin, Cout, CoutP);
  MG_CHECK_LAUNCH(ctx);
  rc = umma_pack_weights(ctx, &d1, s.W1, s.wp, 0);
  if (rc) return rc;
  Yp.data = s.Yp;
  rc = umma_conv_forward(ctx, &d1, s.wp, s.b1, &Yp, nullptr);
  if (rc) return rc;
  upconv_shuffle_kernel<true><<<(unsigned)mg_cdiv(pixels * 4 * (CoutP / 8), 256), 256, 0, ctx->stream>>>(s.Yp, (bf16*)y->data, x->N, x->H, x->W, CoutP);
  MG_CHECK_LAUNCH(ctx);
  return MG_OK;
}

int upconv_tc_backward(mg_ctx* ctx, const mg_grid* x, const float* w, const mg_grid* g, mg_grid* dx, float* dw, float* dbias, float gscale) {
  mg_conv_desc d1; mg_grid Gp;
  if (!applies(ctx, x, g, &d1, &Gp)) return -1;
  if (dx && (dx->Cp != x->Cp || dx->scale)) return -1;
  const int Cin = x->C, Cout = g->C, CoutP = g->Cp;
  const int64_t pixels = (int64_t)x->N * x->H * x->W;
  Scratch s;
  int rc = scratch(ctx, &d1, Cin, CoutP, pixels, &s);
  if (rc) return rc;
  Gp.data = s.Yp;
  upconv_shuffle_kernel<false><<<(unsigned)mg_cdiv(pixels * 4 * (CoutP / 8), 256), 256, 0, ctx->stream>>>(s.Yp, (bf16*)g->data, x->N, x->H, x->W, CoutP);
  MG_CHECK_LAUNCH(ctx);
  if (dx) {
    upconv_w1_kernel<<<(unsigned)mg_cdiv((int64_t)4 * CoutP * Cin, 256), 256, 0, ctx->stream>>>(w, nullptr, s.W1, s.b1, Cin, Cout, CoutP);
    MG_CHECK_LAUNCH(ctx);
    rc = umma_pack_weights(ctx, &d1, s.W1, s.wpt, 1);
    if (rc) return rc;
    rc = umma_conv_backward_data(ctx, &d1, s.wpt, &Gp, dx);
    if (rc) return rc;
  }
  if (dw) {
    MG_CUDA(ctx, cudaMemsetAsync(s.dW1, 0, (size_t)4 * CoutP * Cin * sizeof(float), ctx->stream));
    MG_CUDA(ctx, cudaMemsetAsync(s.db1, 0, (size_t)4 * CoutP * sizeof(float), ctx->stream));
    rc = umma_conv_backward_weight(ctx, &d1, &Gp, s.dW1, dbias ? s.db1 : nullptr, gscale);
    if (rc) return rc;
    upconv_dw_kernel<<<(unsigned)mg_cdiv((int64_t)4 * Cout * Cin, 256), 256, 0, ctx->stream>>>(s.dW1, s.db1, dw, dbias, Cin, Cout, CoutP);
    MG_CHECK_LAUNCH(ctx);
  }
  return MG_OK;
}
