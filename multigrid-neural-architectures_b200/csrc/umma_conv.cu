// placeholder until the tcgen05 path lands
#include "common.cuh"
bool umma_conv_supported(const mg_ctx*, const mg_conv_desc*, int) { return false; }
size_t umma_packed_bytes(const mg_conv_desc*, int) { return 0; }
int umma_pack_weights(mg_ctx* ctx, const mg_conv_desc*, const float*, void*, int) { MG_FAIL(ctx, MG_ERR_UNSUPPORTED, "tcgen05 path not built"); }
int umma_conv_forward(mg_ctx* ctx, const mg_conv_desc*, const void*, const float*, mg_grid*, double*) { MG_FAIL(ctx, MG_ERR_UNSUPPORTED, "tcgen05 path not built"); }
int umma_conv_backward_data(mg_ctx* ctx, const mg_conv_desc*, const void*, const mg_grid*, mg_grid*) { MG_FAIL(ctx, MG_ERR_UNSUPPORTED, "tcgen05 path not built"); }
int umma_conv_backward_weight(mg_ctx* ctx, const mg_conv_desc*, const mg_grid*, float*, float*, float) { MG_FAIL(ctx, MG_ERR_UNSUPPORTED, "tcgen05 path not built"); }
