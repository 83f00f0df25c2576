// K3 — zstd frame encoder for sm_100a.
//
// Replaces zstd::bulk::compress(chunk, 12) (reference src/util/chunk.rs:89-90): one standard
// zstd frame per selected chunk -- single segment, frame content size present, no checksum,
// no dictionary id (the shape libzstd's one-shot API emits).
//
// Stages (all on the caller's stream):
//   plan      per chunk: block count / frame header size                       (enc_plan_kernel)
//   compress  per 128 KiB block: LZ parse + entropy coding into block scratch  (see zstd_enc_lz.cuh)
//   size      per chunk: frame length from the block bodies                    (enc_size_kernel)
//   place     exclusive scan of frame lengths over the batch -> frame offsets  (enc_scan_kernel)
//   emit      per block: write header + body (or raw fallback) into d_out      (enc_emit_kernel)
#include <stdlib.h>
#include "common.cuh"
#include "zstd_enc_lz.cuh"
#include "zstd_enc_lz2.cuh"

#define SQ_BLOCK_MAX (128u * 1024u)  // zstd Block_Maximum_Size
#define SQ_MAX_BLOCKS 16u            // 2 MiB / 128 KiB

struct sq_block_info {   // one per (chunk, block)
    uint32_t body_len;   // bytes of compressed body in scratch (type 2) / 1 (type 1) / block_len (type 0)
    uint32_t type;       // 0 raw, 1 RLE, 2 compressed
};

struct sq_enc_scratch {
    sq_block_info *blocks;   // [cap_chunks * SQ_MAX_BLOCKS]
    uint32_t *frame_len;     // [cap_chunks]
    uint8_t *bodies;         // compressed block bodies, stride lz::BODY_STRIDE per (chunk, block)
    zc::Seq *seqs;           // [cap_chunks * lz::MAX_SEQ_PER_CHUNK] parsed sequences
    lz::BlockMeta *meta;     // [cap_chunks * SQ_MAX_BLOCKS]
    uint32_t *rec;           // [cap_chunks * lz::REC_PER_CHUNK] per-position parse records
    uint8_t *lits;           // per entropy warp: gathered literals
    uint32_t *sbits;         // per entropy warp: FSE state-transition records + packed symbol codes, 4 x SEQ_PER_BLOCK
    uint32_t ent_warps;
    uint32_t *list;          // [lz_sub * lz2::LIST_STRIDE] the chunks' sorted row lists (index_kernel -> search_kernel)
    uint32_t *words;         // [lz_sub * lz::REC_PER_CHUNK] index_kernel pass 1 -> pass 2: row + tags of every position
    uint32_t *span_start;    // [lz_sub + 1] search spans of a sub-batch, exclusive prefix
    uint32_t lz_sub, search_ctas;
    cudaEvent_t tev[5];      // SQ_FLAG_STAGE_TIMING: before search / after search / after chase / after entropy / after emit
    int tev_valid;
    uint32_t cap_chunks;
    uint32_t *status;        // [0] != 0 => capacity overflow; [1],[2] work counters
};

namespace {

__host__ __device__ inline uint32_t frame_header_len(uint32_t len) { return 4 + 1 + (len <= 255 ? 1 : len <= 65791 ? 2 : 4); }
__host__ __device__ inline uint32_t n_blocks(uint32_t len) { return len == 0 ? 1 : (len + SQ_BLOCK_MAX - 1) / SQ_BLOCK_MAX; }

// v0 "compress": every block raw.  The LZ/entropy stages overwrite blocks[] where they win.
__global__ void enc_plan_kernel(const sq_span *__restrict__ spans, const uint8_t *__restrict__ select, uint32_t n,
                                sq_block_info *__restrict__ blocks) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * SQ_MAX_BLOCKS) return;
    uint32_t c = t / SQ_MAX_BLOCKS, b = t % SQ_MAX_BLOCKS;
    uint32_t len = spans[c].len, nb = n_blocks(len);
    sq_block_info bi = {0, 0};
    if ((!select || select[c]) && b < nb) {
        uint32_t blen = min(SQ_BLOCK_MAX, len - b * SQ_BLOCK_MAX);
        bi.body_len = blen;
        bi.type = 0;
    }
    blocks[t] = bi;
}

__global__ void enc_size_kernel(const sq_span *__restrict__ spans, const uint8_t *__restrict__ select, uint32_t n,
                                const sq_block_info *__restrict__ blocks, uint32_t *__restrict__ frame_len) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    uint32_t fl = 0;
    if (!select || select[c]) {
        uint32_t len = spans[c].len, nb = n_blocks(len);
        fl = frame_header_len(len);
        for (uint32_t b = 0; b < nb; b++) fl += 3 + blocks[c * SQ_MAX_BLOCKS + b].body_len;
    }
    frame_len[c] = fl;
}

// single-CTA exclusive scan (n <= a few 100k): 1024 threads, serial over tiles
__global__ void __launch_bounds__(1024) enc_scan_kernel(const uint32_t *__restrict__ frame_len, uint32_t n, uint64_t out_capacity,
                                                         uint64_t *__restrict__ frame_off, uint32_t *__restrict__ frame_len_out,
                                                         uint64_t *__restrict__ total, uint32_t *__restrict__ status) {
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint64_t v = i < n ? frame_len[i] : 0, x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) warp_sums[w] = x;
        __syncthreads();
        if (w == 0) {
            uint64_t s = warp_sums[lane], z = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, z, d); if (lane >= d) z += y; }
            warp_sums[lane] = z - s;  // exclusive
        }
        __syncthreads();
        uint64_t excl = carry + warp_sums[w] + x - v;
        if (i < n) { frame_off[i] = excl; frame_len_out[i] = (uint32_t)v; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *total = carry;
        if (carry > out_capacity) atomicExch(status, 1u);
    }
}

// Cooperative byte copy by one CTA to an arbitrarily aligned destination.
__device__ void cta_copy(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, uint32_t n) {
    // head: until dst is 16-byte aligned
    uint32_t head = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
    if (head > n) head = n;
    if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
    dst += head; src += head; n -= head;
    const uint32_t nvec = n >> 4;
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3) * 8;
    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
    uint4 *d128 = reinterpret_cast<uint4 *>(dst);
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const uint4 *s128 = reinterpret_cast<const uint4 *>(src);
        for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) d128[i] = __ldg(&s128[i]);
    } else {
        for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) {
            const uint32_t *p = s32 + 4 * i;
            uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3);
            uint4 o;
            if (sh == 0) { o = make_uint4(w0, w1, w2, w3); }
            else {
                uint32_t w4 = __ldg(p + 4);
                o.x = __funnelshift_r(w0, w1, sh); o.y = __funnelshift_r(w1, w2, sh);
                o.z = __funnelshift_r(w2, w3, sh); o.w = __funnelshift_r(w3, w4, sh);
            }
            d128[i] = o;
        }
    }
    const uint32_t done = nvec << 4, tail = n - done;
    if (threadIdx.x < tail) dst[done + threadIdx.x] = src[done + threadIdx.x];
}

// grid = (SQ_MAX_BLOCKS, n): CTA (b, c) writes block b of chunk c; b == 0 also writes the frame header.
__global__ void __launch_bounds__(256) enc_emit_kernel(const uint8_t *__restrict__ data, const sq_span *__restrict__ spans, uint32_t n,
                                                        const sq_block_info *__restrict__ blocks, const uint8_t *__restrict__ bodies,
                                                        const uint64_t *__restrict__ frame_off, const uint32_t *__restrict__ frame_len,
                                                        uint8_t *__restrict__ out, const uint32_t *__restrict__ status) {
    const uint32_t c = blockIdx.y, b = blockIdx.x;
    if (*status) return;            // capacity overflow: write nothing
    if (frame_len[c] == 0) return;  // not selected
    const uint32_t len = spans[c].len, nb = n_blocks(len);
    if (b >= nb) return;
    uint8_t *f = out + frame_off[c];
    uint32_t hl = frame_header_len(len);
    if (b == 0 && threadIdx.x == 0) {
        f[0] = 0x28; f[1] = 0xB5; f[2] = 0x2F; f[3] = 0xFD;  // magic 0xFD2FB528 LE
        if (len <= 255) { f[4] = 0x20; f[5] = (uint8_t)len; }
        else if (len <= 65791) { f[4] = 0x60; uint32_t v = len - 256; f[5] = (uint8_t)v; f[6] = (uint8_t)(v >> 8); }
        else { f[4] = 0xA0; f[5] = (uint8_t)len; f[6] = (uint8_t)(len >> 8); f[7] = (uint8_t)(len >> 16); f[8] = (uint8_t)(len >> 24); }
    }
    uint32_t pos = hl;
    for (uint32_t k = 0; k < b; k++) pos += 3 + blocks[c * SQ_MAX_BLOCKS + k].body_len;
    const sq_block_info bi = blocks[c * SQ_MAX_BLOCKS + b];
    const uint32_t blen = len == 0 ? 0 : min(SQ_BLOCK_MAX, len - b * SQ_BLOCK_MAX);
    if (threadIdx.x == 0) {
        // Block_Header: bit0 last, bits1-2 type, bits 3.. size (regenerated size for raw/RLE, body size for compressed)
        uint32_t h = (b + 1 == nb ? 1u : 0u) | bi.type << 1 | (bi.type == 2 ? bi.body_len : blen) << 3;
        f[pos] = (uint8_t)h; f[pos + 1] = (uint8_t)(h >> 8); f[pos + 2] = (uint8_t)(h >> 16);
    }
    const uint8_t *src = bi.type == 0 ? data + spans[c].off + (uint64_t)b * SQ_BLOCK_MAX
                                      : bodies + ((uint64_t)c * SQ_MAX_BLOCKS + b) * lz::BODY_STRIDE;
    cta_copy(f + pos + 3, src, bi.body_len);
}

}  // namespace


// ---- search: sub-batches of at most LZ_SUB chunks go through span_plan -> index -> search; the list scratch (8 MB per chunk of a
// sub-batch) is what bounds the sub-batch ----
namespace {
// search: three CTAs of eight warps per SM (80 registers, 2048 x 16-bit continuation table per warp); four CTAs at 64
// registers spill and were 9 % slower
constexpr int LZ_SEARCH_THREADS = 256, LZ_SEARCH_MINB = 3, LZ_SEARCH_TLOG = 10, LZ_INDEX_THREADS = 512, LZ_INDEX_PARTS = 4;
constexpr uint32_t LZ_SUB_MAX = 256;
constexpr int SEARCH_SMEM = (int)(lz2::SearchSmem<LZ_SEARCH_TLOG>::PER_WARP * (LZ_SEARCH_THREADS / 32));
}  // namespace

static int32_t enc_scratch(sq_ctx *ctx, uint32_t n, int set) {
    if (!ctx->enc_sets[set]) {
        ctx->enc_sets[set] = new sq_enc_scratch();
        memset(ctx->enc_sets[set], 0, sizeof(sq_enc_scratch));
    }
    sq_enc_scratch *e = ctx->enc_sets[set];
    if (!e->lits) {  // per-resident-worker state, sized once from the SM count
        {   // entropy stage: one warp per block, 4 warps per CTA; shared memory (42.8 KB per CTA) allows 5 CTAs per SM
            const char *ov = getenv("SQ_ENT_WARPS_PER_SM");
            e->ent_warps = (uint32_t)ctx->sm_count * (ov && atoi(ov) > 0 ? (uint32_t)atoi(ov) : 20u);
            SQ_CUDA(ctx, cudaFuncSetAttribute(lz::entropy_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        }
        {   // search: all warps are interchangeable, the grid is what is resident
            SQ_CUDA(ctx, cudaFuncSetAttribute(lz2::index_kernel<LZ_INDEX_THREADS, LZ_INDEX_PARTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(lz2::ROWS * sizeof(uint32_t))));
            int per_sm = 0;
            SQ_CUDA(ctx, cudaFuncSetAttribute(lz2::search_kernel<LZ_SEARCH_THREADS, LZ_SEARCH_MINB, LZ_SEARCH_TLOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SEARCH_SMEM));
            SQ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lz2::search_kernel<LZ_SEARCH_THREADS, LZ_SEARCH_MINB, LZ_SEARCH_TLOG>, LZ_SEARCH_THREADS, SEARCH_SMEM));
            if (per_sm < 1) return sq_set_error(ctx, SQ_ERR_CUDA, "search kernel does not fit this device");
            e->search_ctas = (uint32_t)(per_sm * ctx->sm_count);
            if (getenv("SQ_TIMING")) fprintf(stderr, "[sq] search kernel: %d threads, %d CTAs per SM, %d B of shared memory per CTA\n", LZ_SEARCH_THREADS, per_sm, SEARCH_SMEM);
        }
        SQ_CUDA(ctx, cudaMalloc(&e->lits, (size_t)e->ent_warps * (Z_BLOCK_MAX + 64)));
        SQ_CUDA(ctx, cudaMalloc(&e->sbits, (size_t)e->ent_warps * lz::SBITS_STRIDE * sizeof(uint32_t)));
        SQ_CUDA(ctx, cudaMalloc(&e->status, 8 * sizeof(uint32_t)));
        SQ_CUDA(ctx, cudaMemset(e->status, 0, 8 * sizeof(uint32_t)));
    }
    if (e->cap_chunks < n) {
        SQ_CUDA(ctx, cudaDeviceSynchronize());
        cudaFree(e->blocks); cudaFree(e->frame_len); cudaFree(e->bodies); cudaFree(e->seqs); cudaFree(e->meta); cudaFree(e->rec);
        cudaFree(e->list); cudaFree(e->words); cudaFree(e->span_start); e->list = nullptr; e->words = nullptr; e->span_start = nullptr;
        e->rec = nullptr; e->blocks = nullptr; e->frame_len = nullptr; e->bodies = nullptr; e->seqs = nullptr; e->meta = nullptr; e->cap_chunks = 0;
        uint32_t cap = n;
        SQ_CUDA(ctx, cudaMalloc(&e->blocks, (size_t)cap * SQ_MAX_BLOCKS * sizeof(sq_block_info)));
        SQ_CUDA(ctx, cudaMalloc(&e->frame_len, (size_t)cap * sizeof(uint32_t)));
        SQ_CUDA(ctx, cudaMalloc(&e->bodies, (size_t)cap * SQ_MAX_BLOCKS * lz::BODY_STRIDE));
        SQ_CUDA(ctx, cudaMalloc(&e->seqs, (size_t)cap * lz::MAX_SEQ_PER_CHUNK * sizeof(zc::Seq)));
        SQ_CUDA(ctx, cudaMalloc(&e->meta, (size_t)cap * SQ_MAX_BLOCKS * sizeof(lz::BlockMeta)));
        SQ_CUDA(ctx, cudaMalloc(&e->rec, (size_t)cap * lz::REC_PER_CHUNK * sizeof(uint32_t)));
        e->lz_sub = cap < LZ_SUB_MAX ? cap : LZ_SUB_MAX;
        SQ_CUDA(ctx, cudaMalloc(&e->list, (size_t)e->lz_sub * lz2::LIST_STRIDE * sizeof(uint32_t)));
        SQ_CUDA(ctx, cudaMalloc(&e->words, (size_t)e->lz_sub * lz::REC_PER_CHUNK * sizeof(uint32_t)));
        SQ_CUDA(ctx, cudaMalloc(&e->span_start, (size_t)(e->lz_sub + 1) * sizeof(uint32_t)));
        e->cap_chunks = cap;
    }
    return SQ_OK;
}

void sq_enc_destroy(sq_ctx *ctx) {
    for (int set = 0; set < 2; set++) {
        sq_enc_scratch *e = ctx->enc_sets[set];
        if (!e) continue;
        cudaFree(e->blocks); cudaFree(e->frame_len); cudaFree(e->bodies); cudaFree(e->status); cudaFree(e->seqs); cudaFree(e->meta); cudaFree(e->rec);
        for (int i = 0; i < 5; i++) if (e->tev[i]) cudaEventDestroy(e->tev[i]);
        cudaFree(e->list); cudaFree(e->words); cudaFree(e->span_start); cudaFree(e->lits); cudaFree(e->sbits);
        delete e;
        ctx->enc_sets[set] = nullptr;
        if (ctx->enc_set_done[set]) cudaEventDestroy(ctx->enc_set_done[set]);
        ctx->enc_set_done[set] = nullptr;
    }
}

extern "C" size_t sq_encode_bound(size_t len) {
    size_t nb = len == 0 ? 1 : (len + SQ_BLOCK_MAX - 1) / SQ_BLOCK_MAX;
    return 9 + 3 * nb + len;  // header (<= 9) + raw blocks: the encoder never emits more than this
}

extern "C" int32_t sq_encode_device(sq_ctx *ctx, const void *d_data, const sq_span *d_spans, const uint8_t *d_select, uint32_t n,
                                    void *d_out, uint64_t out_capacity, uint64_t *d_frame_off, uint32_t *d_frame_len, uint64_t *d_total,
                                    void *stream) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    // Calls issued on two different streams get two independent scratch sets and may run concurrently (the tail of one batch
    // then overlaps the head of the next).  A further stream takes over the least recently used set; sq_encode_device_set orders
    // it behind that set's previous work with an event.  The caller orders the K1/K2 calls of consecutive batches itself.
    cudaStream_t st = sq_stream(ctx, stream);
    int set = -1;
    for (int i = 0; i < 2 && set < 0; i++) if (ctx->enc_set_bound[i] && ctx->enc_set_stream[i] == st) set = i;
    for (int i = 0; i < 2 && set < 0; i++) if (!ctx->enc_set_bound[i]) set = i;
    if (set < 0) set = ctx->enc_set_lru;
    ctx->enc_set_lru = set ^ 1;
    return sq_encode_device_set(ctx, set, d_data, d_spans, d_select, n, d_out, out_capacity, d_frame_off, d_frame_len, d_total, st);
}

// `set` selects one of two independent scratch sets, so two batches on two streams can be in the encoder at the same time
int32_t sq_encode_device_set(sq_ctx *ctx, int set, const void *d_data, const sq_span *d_spans, const uint8_t *d_select, uint32_t n, void *d_out,
                             uint64_t out_capacity, uint64_t *d_frame_off, uint32_t *d_frame_len, uint64_t *d_total, cudaStream_t st) {
    if (n == 0) return SQ_OK;
    if (!d_data || !d_spans || !d_out || !d_frame_off || !d_frame_len || !d_total)
        return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_encode_device: null pointer");
    int32_t rc = enc_scratch(ctx, n, set);
    if (rc) return rc;
    sq_enc_scratch *e = ctx->enc_sets[set];
    if (!ctx->enc_set_done[set]) SQ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->enc_set_done[set], cudaEventDisableTiming));
    if (ctx->enc_set_bound[set] && ctx->enc_set_stream[set] != st)  // the set changes hands: wait for its previous user
        SQ_CUDA(ctx, cudaStreamWaitEvent(st, ctx->enc_set_done[set], 0));
    ctx->enc_set_bound[set] = 1;
    ctx->enc_set_stream[set] = st;
    const uint32_t nb = n * SQ_MAX_BLOCKS;
    enc_plan_kernel<<<(nb + 255) / 256, 256, 0, st>>>(d_spans, d_select, n, e->blocks);
    SQ_CUDA(ctx, cudaMemsetAsync(e->status + 1, 0, 3 * sizeof(uint32_t), st));
    {
        static_assert(sizeof(lz::BlockOut) == sizeof(sq_block_info), "block info layout");
        const bool timing = (ctx->flags & SQ_FLAG_STAGE_TIMING) != 0;
        if (timing && !e->tev[0]) for (int i = 0; i < 5; i++) SQ_CUDA(ctx, cudaEventCreate(&e->tev[i]));
        if (timing) SQ_CUDA(ctx, cudaEventRecord(e->tev[0], st));
        // search: per sub-batch, the chunks' sorted row lists (one CTA per chunk), then every warp of the grid on any 256 positions
        static const uint32_t dbg = getenv("SQ_LZ2_DBG") ? (uint32_t)atoi(getenv("SQ_LZ2_DBG")) : 0u;
        for (uint32_t first = 0; first < n; first += e->lz_sub) {
            const uint32_t count = n - first < e->lz_sub ? n - first : e->lz_sub;
            lz2::span_plan_kernel<<<1, 1024, 0, st>>>(d_spans, d_select, first, count, e->span_start, e->status + 1);
            {   // one cluster of LZ_INDEX_PARTS CTAs per chunk in flight, two CTAs per SM
                const uint32_t resident = 2u * (uint32_t)ctx->sm_count / LZ_INDEX_PARTS;
                cudaLaunchConfig_t cfg;
                memset(&cfg, 0, sizeof cfg);
                cfg.gridDim = dim3((count < resident ? count : resident) * LZ_INDEX_PARTS, 1, 1);
                cfg.blockDim = dim3(LZ_INDEX_THREADS, 1, 1);
                cfg.dynamicSmemBytes = lz2::ROWS * sizeof(uint32_t);
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = LZ_INDEX_PARTS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                SQ_CUDA(ctx, cudaLaunchKernelEx(&cfg, lz2::index_kernel<LZ_INDEX_THREADS, LZ_INDEX_PARTS>, (const uint8_t *)d_data, d_spans, d_select, first, count,
                                                e->list, e->words, e->rec, (uint32_t)((ctx->flags & SQ_FLAG_DETERMINISTIC) != 0)));
            }
            lz2::search_kernel<LZ_SEARCH_THREADS, LZ_SEARCH_MINB, LZ_SEARCH_TLOG><<<e->search_ctas, LZ_SEARCH_THREADS, SEARCH_SMEM, st>>>((const uint8_t *)d_data, d_spans, first, count,
                                                                                                                                    e->span_start, e->list, e->rec, e->status + 1, dbg);
            SQ_LAUNCHED(ctx, 3);
        }
        if (timing) SQ_CUDA(ctx, cudaEventRecord(e->tev[1], st));
        static const bool chase_thread = getenv("SQ_LZ2_DBG") && (atoi(getenv("SQ_LZ2_DBG")) & 4);  // debugging reference: the scalar parse, one thread per block
        if (chase_thread) lz2::chase_thread_kernel<<<(n * SQ_MAX_BLOCKS + 63) / 64, 64, 0, st>>>((const uint8_t *)d_data, d_spans, d_select, n, e->rec, e->seqs, e->meta);
        else {
            const uint32_t want = (n * SQ_MAX_BLOCKS + 3) / 4, resident = (uint32_t)ctx->sm_count * 10u;  // 10 CTAs of 4 warps per SM at 48 registers
            lz2::chase_kernel<<<want < resident ? want : resident, 128, 0, st>>>((const uint8_t *)d_data, d_spans, d_select, n, e->rec, e->seqs, e->meta, e->status + 3);
        }
        if (timing) SQ_CUDA(ctx, cudaEventRecord(e->tev[2], st));
        const uint32_t ent_ctas = e->ent_warps / 4;
        lz::entropy_kernel<<<ent_ctas, 128, 0, st>>>((const uint8_t *)d_data, d_spans, d_select, n, e->seqs, e->meta, e->lits, e->bodies,
                                                     reinterpret_cast<lz::BlockOut *>(e->blocks), e->sbits, e->status + 2);
    }
    if (ctx->flags & SQ_FLAG_STAGE_TIMING) SQ_CUDA(ctx, cudaEventRecord(e->tev[3], st));
    enc_size_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_spans, d_select, n, e->blocks, e->frame_len);
    enc_scan_kernel<<<1, 1024, 0, st>>>(e->frame_len, n, out_capacity, d_frame_off, d_frame_len, d_total, e->status);
    enc_emit_kernel<<<dim3(SQ_MAX_BLOCKS, n), 256, 0, st>>>((const uint8_t *)d_data, d_spans, n, e->blocks, e->bodies, d_frame_off, d_frame_len,
                                                            (uint8_t *)d_out, e->status);
    if (ctx->flags & SQ_FLAG_STAGE_TIMING) { SQ_CUDA(ctx, cudaEventRecord(e->tev[4], st)); e->tev_valid = 1; }
    SQ_LAUNCHED(ctx, 6);
    SQ_CUDA(ctx, cudaGetLastError());
    SQ_CUDA(ctx, cudaEventRecord(ctx->enc_set_done[set], st));
    return SQ_OK;
}

// status of the most recent encode on this context (synchronizes): SQ_ERR_CAPACITY if d_out was too small
extern "C" int32_t sq_encode_status(sq_ctx *ctx) {
    if (!ctx) return SQ_OK;
    SQ_CUDA(ctx, cudaDeviceSynchronize());
    for (int set = 0; set < 2; set++) {
        if (!ctx->enc_sets[set]) continue;
        uint32_t s = 0;
        SQ_CUDA(ctx, cudaMemcpy(&s, ctx->enc_sets[set]->status, sizeof s, cudaMemcpyDeviceToHost));
        if (s) {
            cudaMemset(ctx->enc_sets[set]->status, 0, sizeof s);
            return sq_set_error(ctx, SQ_ERR_CAPACITY, "encode output buffer too small");
        }
    }
    return SQ_OK;
}

extern "C" int32_t sq_encode_stage_ms(sq_ctx *ctx, void *stream, float out[4]) {
    if (!ctx || !out) return SQ_ERR_INVALID_ARG;
    cudaStream_t st = sq_stream(ctx, stream);
    for (int set = 0; set < 2; set++) {
        sq_enc_scratch *e = ctx->enc_sets[set];
        if (!e || !e->tev_valid || !ctx->enc_set_bound[set] || ctx->enc_set_stream[set] != st) continue;
        SQ_CUDA(ctx, cudaEventSynchronize(e->tev[4]));
        for (int i = 0; i < 4; i++) SQ_CUDA(ctx, cudaEventElapsedTime(&out[i], e->tev[i], e->tev[i + 1]));
        return SQ_OK;
    }
    return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_encode_stage_ms: no timed encode on this stream (SQ_FLAG_STAGE_TIMING set?)");
}
