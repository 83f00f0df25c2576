// K4 — zstd frame decoder (placeholder until the decoder kernels land).
#include "common.cuh"
struct sq_dec_scratch { int unused; };
void sq_dec_destroy(sq_ctx *ctx) { delete ctx->dec; ctx->dec = nullptr; }
extern "C" int32_t sq_decode_device(sq_ctx *ctx, const void *, const sq_frame *, uint32_t, void *, sq_frame_result *, void *) {
    return sq_set_error(ctx, SQ_ERR_OTHER, "sq_decode_device: decoder not built yet");
}
