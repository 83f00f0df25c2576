// K4 — zstd frame decoder for sm_100a.
//
// Replaces zstd::bulk::decompress(bytes, orig_size) in read_chunks (reference
// src/archive/reader.rs:276-305): one chunk-record payload -> its decoded bytes, accepting what stock
// ZSTD_decompress accepts whole (several frames, skippable frames, frames without content size) and
// failing per payload on corruption, capacity overflow or trailing garbage.
//
// Mapping.  Blocks of one frame chain through repeat offsets, treeless literals and Repeat_Mode tables,
// so a frame decodes in order; parallelism comes from frames.  One persistent WARP owns one payload at a
// time (atomic work counter).  All 32 lanes run the same control flow (zstd_dec_core.h): headers, table
// construction and the serial FSE sequence chain are evaluated redundantly by every lane from shared-memory
// tables, which costs nothing in SIMT and needs no broadcast; lanes split the 4 Huffman literal streams,
// the decode-table fills, and the bytes of every literal run and match copy (coalesced within a copy).
// Per-warp tables (Huffman 4 KB, LL/ML/OF cells 5 KB) live in shared memory; decoded Huffman literals go
// to a 128 KiB per-warp slot in HBM (L2 resident).  Algorithmic traffic = payload bytes read + decoded
// bytes written.
#include "common.cuh"
#include "zstd_dec_core.h"

struct sq_dec_scratch {
    uint8_t *lits;      // per resident warp: Z_BLOCK_MAX + 64
    void *build;        // per resident warp: table-construction scratch (zd::Scratch, 2.3 KB; in HBM so that shared memory holds only
                        // the decode tables and six CTAs fit an SM)
    uint32_t *counter;
    uint32_t warps;
};

namespace {
constexpr uint32_t DEC_WARPS_PER_CTA = 4;
struct WarpState { zd::Tables T; };

__global__ void __launch_bounds__(DEC_WARPS_PER_CTA * 32, 6) zstd_decode_kernel(const uint8_t *__restrict__ comp, const sq_frame *__restrict__ frames,
                                                                             uint32_t n, uint8_t *__restrict__ out, sq_frame_result *__restrict__ res,
                                                                             uint8_t *__restrict__ lits_all, zd::Scratch *__restrict__ build_all,
                                                                             uint32_t *__restrict__ counter) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    WarpState *ws = reinterpret_cast<WarpState *>(smem) + w;
    uint8_t *lits = lits_all + (size_t)(blockIdx.x * DEC_WARPS_PER_CTA + w) * (Z_BLOCK_MAX + 64);
    zd::Scratch *S = build_all + (blockIdx.x * DEC_WARPS_PER_CTA + w);
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(counter, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n) break;
        const sq_frame f = frames[i];
        const int64_t r = zd::decode_payload(comp + f.src_off, f.src_len, out + f.dst_off, f.capacity, &ws->T, S, lits);
        __syncwarp();
        if (lane == 0) {
            sq_frame_result fr;
            fr.out_len = r < 0 ? 0u : (uint32_t)r;
            fr.status = r < 0 ? SQ_ERR_READER : SQ_OK;
            res[i] = fr;
        }
    }
}
}  // namespace

void sq_dec_destroy(sq_ctx *ctx) {
    sq_dec_scratch *d = ctx->dec;
    if (!d) return;
    cudaFree(d->lits); cudaFree(d->build); cudaFree(d->counter);
    delete d;
    ctx->dec = nullptr;
}

extern "C" int32_t sq_decode_device(sq_ctx *ctx, const void *d_comp, const sq_frame *d_frames, uint32_t n, void *d_out,
                                    sq_frame_result *d_results, void *stream) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    if (n == 0) return SQ_OK;
    if (!d_comp || !d_frames || !d_out || !d_results) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_decode_device: null pointer");
    const size_t smem = DEC_WARPS_PER_CTA * sizeof(WarpState);
    if (!ctx->dec) {
        sq_dec_scratch *d = new sq_dec_scratch();
        memset(d, 0, sizeof *d);
        ctx->dec = d;
        int ctas_per_sm = 0;
        SQ_CUDA(ctx, cudaFuncSetAttribute(zstd_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SQ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, zstd_decode_kernel, DEC_WARPS_PER_CTA * 32, smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        d->warps = (uint32_t)ctx->sm_count * (uint32_t)ctas_per_sm * DEC_WARPS_PER_CTA;
        SQ_CUDA(ctx, cudaMalloc(&d->lits, (size_t)d->warps * (Z_BLOCK_MAX + 64)));
        SQ_CUDA(ctx, cudaMalloc(&d->build, (size_t)d->warps * sizeof(zd::Scratch)));
        SQ_CUDA(ctx, cudaMalloc(&d->counter, sizeof(uint32_t)));
    }
    sq_dec_scratch *d = ctx->dec;
    cudaStream_t st = sq_stream(ctx, stream);
    SQ_CUDA(ctx, cudaMemsetAsync(d->counter, 0, sizeof(uint32_t), st));
    uint32_t ctas = d->warps / DEC_WARPS_PER_CTA;
    const uint32_t need = (n + DEC_WARPS_PER_CTA - 1) / DEC_WARPS_PER_CTA;
    if (ctas > need) ctas = need;
    zstd_decode_kernel<<<ctas, DEC_WARPS_PER_CTA * 32, smem, st>>>((const uint8_t *)d_comp, d_frames, n, (uint8_t *)d_out, d_results, d->lits,
                                                                   (zd::Scratch *)d->build, d->counter);
    SQ_LAUNCHED(ctx, 1);
    SQ_CUDA(ctx, cudaGetLastError());
    return SQ_OK;
}
