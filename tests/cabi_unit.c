/* Pure-C driver of the plan-level C ABI (include/mgconv.h): one residual multigrid unit
 * (models/ilsvrc/rnmg.lua:91-159) forward + backward with NO Python in between -- the call sequence a LuaJIT host
 * makes through ffi.C.  Reads the problem (shapes, inputs, parameters, output gradients) from a binary file written
 * by tests/test_cabi.py, writes outputs / gradients / running statistics to another; the pytest compares them with the oracle.
 *
 *   gcc -std=c99 tests/cabi_unit.c -Iinclude -I/usr/local/cuda/include -o tests/cabi_unit \
 *       multigrid-neural-architectures_b200/mgconv/libmgconv.so -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,...
 *   tests/cabi_unit <bf16|fp32> in.bin out.bin
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <cuda_runtime_api.h>
#include "mgconv.h"

#define CHECK_CUDA(e) do { cudaError_t _e = (e); if (_e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__); return 2; } } while (0)
#define CHECK_MG(ctx, e) do { int _r = (e); if (_r != 0) { fprintf(stderr, "mgconv status %d at %s:%d: %s\n", _r, __FILE__, __LINE__, mg_last_error(ctx)); return 3; } } while (0)

static float* upload(FILE* f, size_t n) {
  float* h = (float*)malloc(n * sizeof(float));
  float* d = NULL;
  if (!h || fread(h, sizeof(float), n, f) != n) { fprintf(stderr, "short read (%zu floats)\n", n); exit(4); }
  if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess) exit(5);
  cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice);
  free(h);
  return d;
}
static float* zeros(size_t n) {
  float* d = NULL;
  if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess) exit(5);
  cudaMemset(d, 0, n * sizeof(float));
  return d;
}
static void download(FILE* f, const float* d, size_t n) {
  float* h = (float*)malloc(n * sizeof(float));
  cudaMemcpy(h, d, n * sizeof(float), cudaMemcpyDeviceToHost);
  fwrite(h, sizeof(float), n, f);
  free(h);
}

int main(int argc, char** argv) {
  if (argc != 4) { fprintf(stderr, "usage: %s <bf16|fp32> in.bin out.bin\n", argv[0]); return 1; }
  const int dtype = strcmp(argv[1], "bf16") == 0 ? MG_BF16 : MG_F32;
  FILE* in = fopen(argv[2], "rb");
  if (!in) { perror(argv[2]); return 1; }
  int32_t hdr[3 + 5 * MG_STAGE_MAX_GRIDS];
  if (fread(hdr, sizeof(int32_t), 3 + 5 * MG_STAGE_MAX_GRIDS, in) != 3 + 5 * MG_STAGE_MAX_GRIDS) return 4;
  mg_stage_desc d;
  memset(&d, 0, sizeof(d));
  d.n_scales = hdr[0];
  const int batch = hdr[1];
  d.residual = hdr[2];
  for (int i = 0; i < MG_STAGE_MAX_GRIDS; ++i) {
    d.C_in[i] = hdr[3 + i]; d.C_out[i] = hdr[3 + MG_STAGE_MAX_GRIDS + i];
    d.H[i] = hdr[3 + 2 * MG_STAGE_MAX_GRIDS + i]; d.W[i] = hdr[3 + 3 * MG_STAGE_MAX_GRIDS + i];
    d.ksize[i] = hdr[3 + 4 * MG_STAGE_MAX_GRIDS + i];
  }
  d.eps = 1e-5f; d.momentum = 0.1f;
  const int n = d.n_scales, L = d.residual ? 2 : 1;

  cudaStream_t stream;
  CHECK_CUDA(cudaStreamCreate(&stream));
  mg_ctx* ctx = NULL;
  if (mg_ctx_create(0, (void*)stream, dtype, &ctx) != 0) { fprintf(stderr, "mg_ctx_create failed (no sm_100 GPU?)\n"); return 3; }

  const float* x[MG_STAGE_MAX_GRIDS]; const float* dy[MG_STAGE_MAX_GRIDS];
  float* y[MG_STAGE_MAX_GRIDS]; float* dx[MG_STAGE_MAX_GRIDS];
  size_t nx[MG_STAGE_MAX_GRIDS], ny[MG_STAGE_MAX_GRIDS];
  for (int i = 0; i < n; ++i) {
    nx[i] = (size_t)batch * d.C_in[i] * d.H[i] * d.W[i];
    ny[i] = (size_t)batch * d.C_out[i] * d.H[i] * d.W[i];
    x[i] = upload(in, nx[i]);
  }
  for (int i = 0; i < n; ++i) { dy[i] = upload(in, ny[i]); y[i] = zeros(ny[i]); dx[i] = zeros(nx[i]); }
  mg_stage_params prm;
  memset(&prm, 0, sizeof(prm));
  size_t nw[2 * MG_STAGE_MAX_GRIDS], nc[2 * MG_STAGE_MAX_GRIDS];
  for (int l = 0; l < L; ++l)
    for (int i = 0; i < n; ++i) {
      const int k = l * n + i;
      int ccat = (l == 0 ? d.C_in[i] : d.C_out[i]);
      if (i > 0) ccat += (l == 0 ? d.C_in[i - 1] : d.C_out[i - 1]);
      if (i + 1 < n) ccat += (l == 0 ? d.C_in[i + 1] : d.C_out[i + 1]);
      nc[k] = (size_t)d.C_out[i];
      nw[k] = nc[k] * ccat * d.ksize[i] * d.ksize[i];
      prm.conv_w[k] = upload(in, nw[k]); prm.conv_b[k] = upload(in, nc[k]);
      prm.bn_g[k] = upload(in, nc[k]);   prm.bn_b[k] = upload(in, nc[k]);
      prm.bn_rm[k] = upload(in, nc[k]);  prm.bn_rv[k] = upload(in, nc[k]);
      prm.conv_gw[k] = zeros(nw[k]); prm.conv_gb[k] = zeros(nc[k]); prm.bn_gg[k] = zeros(nc[k]); prm.bn_gb[k] = zeros(nc[k]);
    }
  fclose(in);

  mg_stage_plan* plan = NULL;
  CHECK_MG(ctx, mg_plan_create(ctx, &d, batch, &plan));
  const size_t wsb = mg_plan_workspace_bytes(plan);
  void* ws = NULL;
  CHECK_CUDA(cudaMalloc(&ws, wsb));
  CHECK_MG(ctx, mg_stage_forward(plan, ws, x, &prm, y, 1));
  CHECK_MG(ctx, mg_stage_backward(plan, ws, dy, &prm, dx, 1.0f));
  CHECK_MG(ctx, mg_ctx_sync(ctx));
  CHECK_CUDA(cudaGetLastError());

  FILE* out = fopen(argv[3], "wb");
  if (!out) { perror(argv[3]); return 1; }
  for (int i = 0; i < n; ++i) download(out, y[i], ny[i]);
  for (int i = 0; i < n; ++i) download(out, dx[i], nx[i]);
  for (int k = 0; k < L * n; ++k) {
    download(out, prm.conv_gw[k], nw[k]); download(out, prm.conv_gb[k], nc[k]);
    download(out, prm.bn_gg[k], nc[k]);   download(out, prm.bn_gb[k], nc[k]);
    download(out, prm.bn_rm[k], nc[k]);   download(out, prm.bn_rv[k], nc[k]);
  }
  fclose(out);
  int64_t launches = 0, tc = 0;
  mg_ctx_launch_count(ctx, &launches);
  mg_ctx_tc_launch_count(ctx, &tc);
  printf("cabi_unit ok: workspace %zu bytes, %lld kernel launches, %lld tcgen05\n", wsb, (long long)launches, (long long)tc);
  mg_plan_destroy(plan);
  mg_ctx_destroy(ctx);
  return 0;
}
