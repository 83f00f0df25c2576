// Device + host drivers for the synthetic corpus generators in corpus.h (bench/test support).
#include "common.cuh"
#include "corpus.h"

namespace {
__global__ void corpus_fill_kernel(uint8_t *out, uint64_t len, uint64_t seed, uint64_t id, uint32_t klass) {
    uint64_t page = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t off = page * SQC_PAGE;
    if (off >= len) return;
    uint32_t limit = len - off < SQC_PAGE ? (uint32_t)(len - off) : SQC_PAGE;
    sqc_fill_page(out + off, limit, seed, id, klass, page);
}
__global__ void corpus_fill_slots_kernel(uint8_t *out, uint64_t slot_bytes, const uint64_t *ids, const uint32_t *klass,
                                         uint32_t pages_per_slot, uint64_t total_pages, uint64_t seed) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_pages) return;
    uint32_t slot = (uint32_t)(t / pages_per_slot);
    uint64_t page = t % pages_per_slot, off = page * SQC_PAGE;
    uint32_t limit = slot_bytes - off < SQC_PAGE ? (uint32_t)(slot_bytes - off) : SQC_PAGE;
    sqc_fill_page(out + (uint64_t)slot * slot_bytes + off, limit, seed, ids[slot], klass[slot], page);
}
}  // namespace

extern "C" int32_t sq_corpus_fill_device(sq_ctx *ctx, void *d_out, uint64_t len, uint64_t seed, uint64_t payload_id, uint32_t klass,
                                         void *stream) {
    if (!ctx || (!d_out && len)) return SQ_ERR_INVALID_ARG;
    if (!len) return SQ_OK;
    if (reinterpret_cast<uintptr_t>(d_out) & 7) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "corpus buffer must be 8-byte aligned");
    uint64_t pages = (len + SQC_PAGE - 1) / SQC_PAGE;
    corpus_fill_kernel<<<(unsigned)((pages + 127) / 128), 128, 0, sq_stream(ctx, stream)>>>((uint8_t *)d_out, len, seed, payload_id, klass);
    SQ_LAUNCHED(ctx, 1);
    SQ_CUDA(ctx, cudaGetLastError());
    return SQ_OK;
}

extern "C" int32_t sq_corpus_fill_slots_device(sq_ctx *ctx, void *d_out, uint64_t slot_bytes, const uint64_t *d_ids,
                                               const uint32_t *d_klass, uint32_t n, uint64_t seed, void *stream) {
    if (!ctx || !d_out || !d_ids || !d_klass) return SQ_ERR_INVALID_ARG;
    if (!n || !slot_bytes) return SQ_OK;
    if ((reinterpret_cast<uintptr_t>(d_out) & 7) || (slot_bytes & 7)) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "corpus slots must be 8-byte aligned");
    uint32_t pps = (uint32_t)((slot_bytes + SQC_PAGE - 1) / SQC_PAGE);
    uint64_t total = (uint64_t)pps * n;
    corpus_fill_slots_kernel<<<(unsigned)((total + 127) / 128), 128, 0, sq_stream(ctx, stream)>>>((uint8_t *)d_out, slot_bytes, d_ids, d_klass,
                                                                                                 pps, total, seed);
    SQ_LAUNCHED(ctx, 1);
    SQ_CUDA(ctx, cudaGetLastError());
    return SQ_OK;
}

extern "C" int32_t sq_corpus_fill_host(void *h_out, uint64_t len, uint64_t seed, uint64_t payload_id, uint32_t klass) {
    if (!h_out && len) return SQ_ERR_INVALID_ARG;
    uint8_t *out = (uint8_t *)h_out;
    alignas(8) uint8_t tmp[SQC_PAGE];
    for (uint64_t off = 0, page = 0; off < len; off += SQC_PAGE, page++) {
        uint32_t limit = len - off < SQC_PAGE ? (uint32_t)(len - off) : SQC_PAGE;
        sqc_fill_page(tmp, SQC_PAGE, seed, payload_id, klass, page);  // host buffer may be unaligned: stage
        memcpy(out + off, tmp, limit);
    }
    return SQ_OK;
}
