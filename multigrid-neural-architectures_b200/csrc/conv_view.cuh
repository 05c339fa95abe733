// Device-side description of one multigrid convolution's gathered input.
#pragma once
#include "common.cuh"

template <typename T>
struct ConvV {
  int n_seg;
  GridV<T> seg[MG_MAX_SEG];
  int mode[MG_MAX_SEG];
  int c_begin[MG_MAX_SEG + 1];   // logical concat channel offsets  [finer | same | coarser]
  int cp_begin[MG_MAX_SEG + 1];  // offsets with every segment padded to its Cp (dcat layout)
  int k, stride, pad;
  int Ccat, CcatP, Cout;
  int N, H, W;    // concatenated-input spatial size
  int Ho, Wo;     // output spatial size

  __device__ __forceinline__ int seg_of(int ci) const {
    int s = 0;
    while (s + 1 < n_seg && ci >= c_begin[s + 1]) ++s;
    return s;
  }
  // value of concat channel ci at input position (iy, ix) (in bounds)
  __device__ __forceinline__ float fetch(int n, int iy, int ix, int ci) const {
    int s = seg_of(ci);
    int c = ci - c_begin[s];
    const GridV<T>& g = seg[s];
    int m = mode[s];
    if (m == MG_SEG_SAME) return g.at(n, iy, ix, c);
    if (m == MG_SEG_POOL) return g.pooled(n, iy, ix, c, nullptr);
    return g.at(n, iy >> 1, ix >> 1, c);
  }
  // padded concat channel -> logical concat channel (or -1 for a pad channel)
  __device__ __forceinline__ int logical_of_padded(int cpad) const {
    if (cpad >= CcatP) return -1;
    int s = 0;
    while (s + 1 < n_seg && cpad >= cp_begin[s + 1]) ++s;
    int c = cpad - cp_begin[s];
    return c < seg[s].C ? c_begin[s] + c : -1;
  }
};

// host: validate the descriptor against the resampling rules of ResampleConcat and build the view
template <typename T>
static int make_conv_view(mg_ctx* ctx, const mg_conv_desc& d, ConvV<T>* out) {
  ConvV<T>& v = *out;
  MG_REQUIRE(ctx, d.n_seg >= 1 && d.n_seg <= MG_MAX_SEG, MG_ERR_INVALID_ARG, "conv: n_seg %d", d.n_seg);
  MG_REQUIRE(ctx, d.ksize >= 1 && d.stride >= 1 && d.pad >= 0 && d.Cout >= 1, MG_ERR_INVALID_ARG,
             "conv: bad ksize/stride/pad/Cout");
  v.n_seg = d.n_seg; v.k = d.ksize; v.stride = d.stride; v.pad = d.pad; v.Cout = d.Cout;
  v.H = d.H; v.W = d.W; v.N = d.seg[0].N;
  v.Ho = (d.H + 2 * d.pad - d.ksize) / d.stride + 1;
  v.Wo = (d.W + 2 * d.pad - d.ksize) / d.stride + 1;
  int c = 0, cp = 0;
  for (int s = 0; s < d.n_seg; ++s) {
    const mg_grid& g = d.seg[s];
    MG_REQUIRE(ctx, g.data != nullptr, MG_ERR_INVALID_ARG, "conv: seg %d null data", s);
    MG_REQUIRE(ctx, g.N == v.N, MG_ERR_SHAPE, "conv: seg %d batch %d != %d", s, g.N, v.N);
    MG_REQUIRE(ctx, g.Cp % 8 == 0 && g.Cp >= g.C, MG_ERR_SHAPE, "conv: seg %d Cp %d (C %d)", s, g.Cp, g.C);
    MG_REQUIRE(ctx, (g.scale == nullptr) == (g.shift == nullptr), MG_ERR_INVALID_ARG, "conv: seg %d scale/shift", s);
    int m = d.seg_mode[s];
    if (m == MG_SEG_SAME) {
      MG_REQUIRE(ctx, g.H == d.H && g.W == d.W, MG_ERR_SHAPE, "conv: SAME seg %d is %dx%d, expected %dx%d", s, g.H, g.W, d.H, d.W);
    } else if (m == MG_SEG_POOL) {  // JoinTable would raise the same size error in the reference
      MG_REQUIRE(ctx, (g.H + 1) / 2 == d.H && (g.W + 1) / 2 == d.W, MG_ERR_SHAPE,
                 "conv: POOL seg %d is %dx%d, ceil/2 != %dx%d", s, g.H, g.W, d.H, d.W);
    } else if (m == MG_SEG_UP) {
      MG_REQUIRE(ctx, g.H * 2 == d.H && g.W * 2 == d.W, MG_ERR_SHAPE,
                 "conv: UP seg %d is %dx%d, x2 != %dx%d", s, g.H, g.W, d.H, d.W);
    } else {
      MG_FAIL(ctx, MG_ERR_INVALID_ARG, "conv: seg %d mode %d", s, m);
    }
    v.seg[s] = make_view<T>(g);
    v.mode[s] = m;
    v.c_begin[s] = c; v.cp_begin[s] = cp;
    c += g.C; cp += g.Cp;
  }
  for (int s = d.n_seg; s <= MG_MAX_SEG; ++s) { v.c_begin[s] = c; v.cp_begin[s] = cp; }
  v.Ccat = c; v.CcatP = cp;
  return MG_OK;
}
