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
//
// Block-parallel path (calls of at most BP_MAX_FRAMES payloads, i.e. when frames alone cannot fill the GPU; zstd_dec_core.h has the
// reasoning).  A payload that is one frame of several blocks is decoded in two passes, whoever wrote it: bp_snapshot_kernel walks
// the table descriptions of a frame's blocks (serial, cheap) so that every block knows the tables it decodes with;
// bp_first_pass_kernel gives every BLOCK a warp for all work that does not read the output (Huffman literals, FSE chain, repeat
// offsets kept symbolic, positions, and placing the literals); bp_matches_kernel then gives every FRAME a warp that resolves the
// offsets and copies the matches block by block, in order (matches reach back across blocks).  bp_scan_kernel decides from the
// headers which payloads try this; the others run through the one-pass kernel on a second stream beside the passes, and so does,
// afterwards, any frame that broke an assumption (a block that does not regenerate 128 KiB, more sequences than a slot holds,
// anything malformed).
#include "common.cuh"
#include "zstd_dec_core.h"

struct sq_dec_scratch {
    uint8_t *lits;      // per resident warp: Z_BLOCK_MAX + 64
    void *build;        // per resident warp: table-construction scratch (zd::Scratch, 2.3 KB; in HBM so that shared memory holds only
                        // the decode tables and six CTAs fit an SM)
    uint32_t *counter;      // [0] one-pass work counter, [1] bp_first_pass, [2] bp_matches, [3] one-pass for frames the two passes gave back, [4] bp_snapshot
    uint32_t warps;
    // block-parallel path, sized for bp_cap frames
    uint32_t bp_cap;
    uint8_t *bp_elig;         // [bp_cap] 0 = one-pass kernel, 1 = two passes, 2 = two passes gave up: one-pass kernel afterwards
    zd::FrameInfo *bp_info;   // [bp_cap]
    zd::BlockTask *bp_tasks;  // [bp_cap * BP_MAX_BLOCKS]
    zd::BlockState *bp_states;
    zd::Tables *bp_snaps;     // [bp_cap * BP_MAX_BLOCKS] the tables every compressed block decodes with
    uint8_t *bp_lits;         // per resident warp of bp_first_pass_kernel: Z_BLOCK_MAX + 64 (it runs beside the one-pass kernel, which has its own)
    zd::StoredSeq *bp_seqs;   // [bp_cap * BP_MAX_BLOCKS * BP_SEQ_CAP]
    void *bp_build;           // per resident warp of bp_first_pass_kernel: its own table-construction scratch
    cudaStream_t aux;         // the one-pass kernel for the frames that are not eligible runs here, beside the two passes
    cudaEvent_t ev_fork, ev_join;
};

namespace {
constexpr uint32_t BP_MAX_FRAMES = 2048, BP_MAX_BLOCKS = 16, BP_SEQ_CAP = Z_BLOCK_MAX / 6 + 8;  // a block with more than one sequence per 6 bytes falls back
constexpr size_t BP_LIT_STRIDE = Z_BLOCK_MAX + 64;
constexpr uint32_t DEC_WARPS_PER_CTA = 4;
struct WarpState { zd::Tables T; };

__global__ void __launch_bounds__(DEC_WARPS_PER_CTA * 32, 6) zstd_decode_kernel(const uint8_t *__restrict__ comp, const sq_frame *__restrict__ frames,
                                                                             uint32_t n, uint8_t *__restrict__ out, sq_frame_result *__restrict__ res,
                                                                             uint8_t *__restrict__ lits_all, zd::Scratch *__restrict__ build_all,
                                                                             uint32_t *__restrict__ counter, const uint8_t *__restrict__ sel, uint32_t want) {
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
        if (sel && sel[i] != want) continue;  // the block-parallel path has it (or had it and gave it back: want == 2)
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

// ---- block-parallel path ------------------------------------------------------------------------------------------------
__global__ void bp_scan_kernel(const uint8_t *__restrict__ comp, const sq_frame *__restrict__ frames, uint32_t n, uint8_t *__restrict__ elig,
                               zd::FrameInfo *__restrict__ info, zd::BlockTask *__restrict__ tasks) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const sq_frame f = frames[i];
    zd::FrameInfo fi;
    fi.nblocks = 0; fi.has_fcs = 0; fi.fcs = 0; fi.chained = 0; fi.reserved = 0;
    const uint32_t nb = zd::scan_frame(comp + f.src_off, f.src_len, BP_MAX_BLOCKS, tasks + (size_t)i * BP_MAX_BLOCKS, &fi);
    elig[i] = nb ? 1 : 0;
    info[i] = fi;
}

__global__ void __launch_bounds__(DEC_WARPS_PER_CTA * 32, 6) bp_snapshot_kernel(const uint8_t *__restrict__ comp, const sq_frame *__restrict__ frames, uint32_t n,
                                                                                uint8_t *__restrict__ elig, const zd::FrameInfo *__restrict__ info,
                                                                                const zd::BlockTask *__restrict__ tasks, zd::Tables *__restrict__ snaps,
                                                                                zd::Scratch *__restrict__ build_all, uint32_t *__restrict__ counter) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    WarpState *ws = reinterpret_cast<WarpState *>(smem) + w;
    zd::Scratch *S = build_all + (blockIdx.x * DEC_WARPS_PER_CTA + w);
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(counter, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n) break;
        if (elig[i] != 1 || !info[i].chained) continue;  // frames whose blocks bring all their own tables need no snapshots
        const size_t first = (size_t)i * BP_MAX_BLOCKS;
        const bool ok = zd::snapshot_frame_tables(comp + frames[i].src_off, tasks + first, info[i].nblocks, &ws->T, S, snaps + first);
        __syncwarp();
        if (!ok && lane == 0) elig[i] = 2;  // malformed table descriptions: the one-pass kernel reports it
    }
}

__global__ void __launch_bounds__(DEC_WARPS_PER_CTA * 32, 6) bp_first_pass_kernel(const uint8_t *__restrict__ comp, const sq_frame *__restrict__ frames, uint32_t n,
                                                                                  const uint8_t *__restrict__ elig, const zd::FrameInfo *__restrict__ info,
                                                                                  const zd::BlockTask *__restrict__ tasks, const zd::Tables *__restrict__ snaps,
                                                                                  zd::BlockState *__restrict__ states, uint8_t *__restrict__ lits_all, zd::StoredSeq *__restrict__ seqs, uint8_t *__restrict__ out,
                                                                                  zd::Scratch *__restrict__ build_all, uint32_t *__restrict__ counter) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    WarpState *ws = reinterpret_cast<WarpState *>(smem) + w;
    zd::Scratch *S = build_all + (blockIdx.x * DEC_WARPS_PER_CTA + w);
    uint8_t *lits = lits_all + (size_t)(blockIdx.x * DEC_WARPS_PER_CTA + w) * BP_LIT_STRIDE;
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n * BP_MAX_BLOCKS) break;
        const uint32_t i = item / BP_MAX_BLOCKS, b = item % BP_MAX_BLOCKS;
        if (elig[i] != 1 || b >= info[i].nblocks) continue;
        const zd::BlockTask t = tasks[item];
        if (t.type != 2) continue;
        const sq_frame f = frames[i];
        const bool chained = info[i].chained != 0;
        if (chained) {  // the tables this block decodes with
            const uint32_t *from = reinterpret_cast<const uint32_t *>(snaps + item);
            uint32_t *to = reinterpret_cast<uint32_t *>(&ws->T);
            for (uint32_t k = lane; k < sizeof(zd::Tables) / 4; k += 32) to[k] = from[k];
            __syncwarp();
        }
        zd::decode_block_first_pass(comp + f.src_off + t.src_off, t.size, chained, &ws->T, S, lits, out + f.dst_off, t.out_start, f.capacity, b == 0,
                                    seqs + (size_t)item * BP_SEQ_CAP, BP_SEQ_CAP, states + item);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128) bp_matches_kernel(const uint8_t *__restrict__ comp, const sq_frame *__restrict__ frames, uint32_t n, uint8_t *__restrict__ elig,
                                                          const zd::FrameInfo *__restrict__ info, const zd::BlockTask *__restrict__ tasks,
                                                          const zd::BlockState *__restrict__ states, const zd::StoredSeq *__restrict__ seqs, uint8_t *__restrict__ out,
                                                          sq_frame_result *__restrict__ res, uint32_t *__restrict__ counter) {
    const uint32_t lane = threadIdx.x & 31;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(counter, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n) break;
        if (elig[i] != 1) continue;
        const sq_frame f = frames[i];
        const size_t first = (size_t)i * BP_MAX_BLOCKS;
        const int64_t r = zd::execute_frame_matches(comp + f.src_off, tasks + first, info + i, states + first, seqs + first * BP_SEQ_CAP, BP_SEQ_CAP, out + f.dst_off,
                                                    f.capacity);
        __syncwarp();
        if (lane == 0) {
            if (r < 0) elig[i] = 2;  // not self-contained after all, or malformed: the one-pass kernel decodes it (and reports the error)
            else {
                sq_frame_result fr;
                fr.out_len = (uint32_t)r;
                fr.status = SQ_OK;
                res[i] = fr;
            }
        }
    }
}
}  // namespace

void sq_dec_destroy(sq_ctx *ctx) {
    sq_dec_scratch *d = ctx->dec;
    if (!d) return;
    cudaFree(d->lits); cudaFree(d->build); cudaFree(d->counter);
    cudaFree(d->bp_elig); cudaFree(d->bp_info); cudaFree(d->bp_tasks); cudaFree(d->bp_states); cudaFree(d->bp_snaps); cudaFree(d->bp_lits); cudaFree(d->bp_seqs); cudaFree(d->bp_build);
    if (d->aux) cudaStreamDestroy(d->aux);
    if (d->ev_fork) cudaEventDestroy(d->ev_fork);
    if (d->ev_join) cudaEventDestroy(d->ev_join);
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
        SQ_CUDA(ctx, cudaMalloc(&d->counter, 8 * sizeof(uint32_t)));
        SQ_CUDA(ctx, cudaFuncSetAttribute(bp_first_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SQ_CUDA(ctx, cudaFuncSetAttribute(bp_snapshot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SQ_CUDA(ctx, cudaMalloc(&d->bp_build, (size_t)d->warps * sizeof(zd::Scratch)));
        SQ_CUDA(ctx, cudaMalloc(&d->bp_lits, (size_t)d->warps * BP_LIT_STRIDE));
        SQ_CUDA(ctx, cudaStreamCreateWithFlags(&d->aux, cudaStreamNonBlocking));
        SQ_CUDA(ctx, cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
        SQ_CUDA(ctx, cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming));
    }
    sq_dec_scratch *d = ctx->dec;
    cudaStream_t st = sq_stream(ctx, stream);
    SQ_CUDA(ctx, cudaMemsetAsync(d->counter, 0, 8 * sizeof(uint32_t), st));
    uint32_t ctas = d->warps / DEC_WARPS_PER_CTA;
    static const bool bp_off = getenv("SQ_NO_BLOCK_PARALLEL") != nullptr;
    if (n <= BP_MAX_FRAMES && !bp_off) {  // too few frames to fill the GPU with one warp each: give the blocks of eligible frames their own warps
        if (d->bp_cap < n) {
            SQ_CUDA(ctx, cudaStreamSynchronize(st));
            cudaFree(d->bp_elig); cudaFree(d->bp_info); cudaFree(d->bp_tasks); cudaFree(d->bp_states); cudaFree(d->bp_snaps); cudaFree(d->bp_seqs);
            d->bp_elig = nullptr; d->bp_info = nullptr; d->bp_tasks = nullptr; d->bp_states = nullptr; d->bp_snaps = nullptr; d->bp_seqs = nullptr; d->bp_cap = 0;
            const size_t items = (size_t)n * BP_MAX_BLOCKS;
            SQ_CUDA(ctx, cudaMalloc(&d->bp_elig, n));
            SQ_CUDA(ctx, cudaMalloc(&d->bp_info, (size_t)n * sizeof(zd::FrameInfo)));
            SQ_CUDA(ctx, cudaMalloc(&d->bp_tasks, items * sizeof(zd::BlockTask)));
            SQ_CUDA(ctx, cudaMalloc(&d->bp_states, items * sizeof(zd::BlockState)));
            SQ_CUDA(ctx, cudaMalloc(&d->bp_snaps, items * sizeof(zd::Tables)));
            SQ_CUDA(ctx, cudaMalloc(&d->bp_seqs, items * BP_SEQ_CAP * sizeof(zd::StoredSeq)));
            d->bp_cap = n;
        }
        bp_scan_kernel<<<(n + 127) / 128, 128, 0, st>>>((const uint8_t *)d_comp, d_frames, n, d->bp_elig, d->bp_info, d->bp_tasks);
        // fork: the one-pass kernel takes the frames the scan left to it, beside the two passes
        SQ_CUDA(ctx, cudaEventRecord(d->ev_fork, st));
        SQ_CUDA(ctx, cudaStreamWaitEvent(d->aux, d->ev_fork, 0));
        const uint32_t need1 = (n + DEC_WARPS_PER_CTA - 1) / DEC_WARPS_PER_CTA;
        zstd_decode_kernel<<<ctas < need1 ? ctas : need1, DEC_WARPS_PER_CTA * 32, smem, d->aux>>>((const uint8_t *)d_comp, d_frames, n, (uint8_t *)d_out, d_results, d->lits,
                                                                                                    (zd::Scratch *)d->build, d->counter, d->bp_elig, 0u);
        SQ_CUDA(ctx, cudaEventRecord(d->ev_join, d->aux));
        bp_snapshot_kernel<<<ctas < need1 ? ctas : need1, DEC_WARPS_PER_CTA * 32, smem, st>>>((const uint8_t *)d_comp, d_frames, n, d->bp_elig, d->bp_info, d->bp_tasks,
                                                                                                d->bp_snaps, (zd::Scratch *)d->bp_build, d->counter + 4);
        const uint32_t want = (n * BP_MAX_BLOCKS + DEC_WARPS_PER_CTA - 1) / DEC_WARPS_PER_CTA;
        bp_first_pass_kernel<<<want < ctas ? want : ctas, DEC_WARPS_PER_CTA * 32, smem, st>>>((const uint8_t *)d_comp, d_frames, n, d->bp_elig, d->bp_info, d->bp_tasks,
                                                                                                d->bp_snaps, d->bp_states, d->bp_lits, d->bp_seqs, (uint8_t *)d_out,
                                                                                                (zd::Scratch *)d->bp_build, d->counter + 1);
        bp_matches_kernel<<<(n + 3) / 4, 128, 0, st>>>((const uint8_t *)d_comp, d_frames, n, d->bp_elig, d->bp_info, d->bp_tasks, d->bp_states, d->bp_seqs,
                                                       (uint8_t *)d_out, d_results, d->counter + 2);
        SQ_CUDA(ctx, cudaStreamWaitEvent(st, d->ev_join, 0));
        // frames the two passes gave back (none for K3-written archives): one-pass, after everything else (it shares the one-pass scratch)
        zstd_decode_kernel<<<ctas < need1 ? ctas : need1, DEC_WARPS_PER_CTA * 32, smem, st>>>((const uint8_t *)d_comp, d_frames, n, (uint8_t *)d_out, d_results, d->lits,
                                                                                                (zd::Scratch *)d->build, d->counter + 3, d->bp_elig, 2u);
        SQ_LAUNCHED(ctx, 6);
        SQ_CUDA(ctx, cudaGetLastError());
        return SQ_OK;
    }
    const uint32_t need = (n + DEC_WARPS_PER_CTA - 1) / DEC_WARPS_PER_CTA;
    if (ctas > need) ctas = need;
    zstd_decode_kernel<<<ctas, DEC_WARPS_PER_CTA * 32, smem, st>>>((const uint8_t *)d_comp, d_frames, n, (uint8_t *)d_out, d_results, d->lits,
                                                                   (zd::Scratch *)d->build, d->counter, nullptr, 0u);
    SQ_LAUNCHED(ctx, 1);
    SQ_CUDA(ctx, cudaGetLastError());
    return SQ_OK;
}
