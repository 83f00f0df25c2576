// Context, error plumbing, pinned staging and the fused batch entry points of the C ABI.
#include <time.h>
#include <stdlib.h>
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"

int32_t sq_xxh3_init(sq_ctx *ctx);

static thread_local char g_create_err[512] = "";

int32_t sq_set_error(sq_ctx *ctx, int32_t code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx ? ctx->err : g_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}

int32_t sq_ensure(sq_ctx *ctx, void **p, size_t *cap, size_t need) {
    if (*cap >= need) return SQ_OK;
    if (*p) { SQ_CUDA(ctx, cudaDeviceSynchronize()); SQ_CUDA(ctx, cudaFree(*p)); *p = nullptr; *cap = 0; }
    size_t want = need + need / 8 + 4096;
    SQ_CUDA(ctx, cudaMalloc(p, want));
    *cap = want;
    return SQ_OK;
}

extern "C" int32_t sq_abi_version(void) { return 1; }
extern "C" int32_t sq_kernel_launches(sq_ctx *ctx, uint64_t *out) {
    if (!ctx || !out) return SQ_ERR_INVALID_ARG;
    *out = ctx->launches;
    return SQ_OK;
}

extern "C" const char *sq_strerror(int32_t s) {
    switch (s) {  // display strings of AppError (reference src/util/errors.rs:7-65)
    case SQ_OK: return "ok";
    case SQ_ERR_IO: return "I/O error";
    case SQ_ERR_READ_DIR: return "Failed to read directory";
    case SQ_ERR_READ_ENTRY: return "Failed to read entry";
    case SQ_ERR_WRITER: return "Error writing to squish";
    case SQ_ERR_READER: return "Error reading from squish";
    case SQ_ERR_FLUSH: return "Failed to flush archive writer";
    case SQ_ERR_COMPRESSION: return "Compression error";
    case SQ_ERR_ARCHIVE: return "Archive format error";
    case SQ_ERR_ENCODER: return "Zstd encoder error";
    case SQ_ERR_LOCK_POISONED: return "Mutex poisoned";
    case SQ_ERR_SENDER: return "Error sending to writer thread";
    case SQ_ERR_CREATE_DIR: return "Error creating directory";
    case SQ_ERR_CREATE_FILE: return "Error creating file";
    case SQ_ERR_FILE_NOT_EXIST: return "Specified file does not exist";
    case SQ_ERR_ILLEGAL_UTF8: return "Illegal UTF8 detected";
    case SQ_ERR_MISSING_CHUNK: return "Missing Chunk for File";
    case SQ_ERR_INVALID_CHUNK_SIZE: return "Invalid chunk size";
    case SQ_ERR_CAP_THREADS: return "Unable to Cap Maximum Threads";
    case SQ_ERR_INVALID_TIMESTAMP: return "Invalid timestamp in squish";
    case SQ_ERR_OTHER: return "Unknown error";
    case SQ_ERR_NO_DEVICE: return "No CUDA device (this library has no CPU fallback)";
    case SQ_ERR_CUDA: return "CUDA runtime error";
    case SQ_ERR_INVALID_ARG: return "Invalid argument";
    case SQ_ERR_CAPACITY: return "Output capacity exceeded";
    default: return "Unknown status";
    }
}

extern "C" const char *sq_last_error(const sq_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" int32_t sq_create(const sq_config *cfg, sq_ctx **out) {
    if (!out) return SQ_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    const bool tm = getenv("SQ_TIMING") != nullptr;  // start-up breakdown on stderr
    auto now = [] { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; };
    const double t0 = now();
    auto mark = [&](const char *what) { if (tm) fprintf(stderr, "[sq_create %7.3f] %s\n", now() - t0, what); };
    cudaError_t e = cudaGetDeviceCount(&ndev);
    mark("driver initialised (cudaGetDeviceCount)");
    if (e != cudaSuccess || ndev == 0)
        return sq_set_error(nullptr, SQ_ERR_NO_DEVICE, "no CUDA device available (%s); libsquish_b200 has no CPU fallback",
                            e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    int dev = cfg ? cfg->device : 0;
    if (dev < 0 || dev >= ndev) return sq_set_error(nullptr, SQ_ERR_INVALID_ARG, "device %d out of range (%d devices)", dev, ndev);
    sq_ctx *ctx = (sq_ctx *)calloc(1, sizeof(sq_ctx));
    if (!ctx) return SQ_ERR_OTHER;
    ctx->device = dev;
    ctx->chunk_size = cfg && cfg->chunk_size ? cfg->chunk_size : SQ_CHUNK_SIZE;
    ctx->max_batch = cfg && cfg->max_batch_chunks ? cfg->max_batch_chunks : 4096;
    ctx->flags = cfg ? cfg->flags : 0;
    ctx->dedup_capacity = cfg && cfg->dedup_capacity ? cfg->dedup_capacity : (1ull << 20);
    int32_t rc = SQ_OK;
    auto fail = [&](int32_t code) { snprintf(g_create_err, sizeof g_create_err, "%s", ctx->err); sq_destroy(ctx); return code; };
    if (ctx->dedup_capacity >= (1ull << 30)) {  // slot indices and key references are 32-bit (slots = 2 x capacity rounded up to a power of two)
        sq_set_error(ctx, SQ_ERR_INVALID_ARG, "dedup_capacity %llu is not below 2^30", (unsigned long long)ctx->dedup_capacity); return fail(SQ_ERR_INVALID_ARG); }
    if (ctx->chunk_size > SQ_CHUNK_SIZE) { sq_set_error(ctx, SQ_ERR_INVALID_CHUNK_SIZE, "chunk_size %u > %u", ctx->chunk_size, SQ_CHUNK_SIZE); return fail(SQ_ERR_INVALID_CHUNK_SIZE); }
    if (cudaSetDevice(dev) != cudaSuccess) { sq_set_error(ctx, SQ_ERR_CUDA, "cudaSetDevice(%d) failed", dev); return fail(SQ_ERR_CUDA); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { sq_set_error(ctx, SQ_ERR_CUDA, "cudaGetDeviceProperties failed"); return fail(SQ_ERR_CUDA); }
    if (prop.major < 10) { sq_set_error(ctx, SQ_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor); return fail(SQ_ERR_NO_DEVICE); }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaFree(nullptr) != cudaSuccess) { sq_set_error(ctx, SQ_ERR_CUDA, "CUDA context creation failed"); return fail(SQ_ERR_CUDA); }
    mark("CUDA context created");
    auto init = [&]() -> int32_t {
        SQ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        SQ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        SQ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            SQ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->slot_stream[i], cudaStreamNonBlocking));
            SQ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->dedup_done[i], cudaEventDisableTiming));
            SQ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->slots[i].h2d_done, cudaEventDisableTiming));
            SQ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->slots[i].compute_done, cudaEventDisableTiming));
            SQ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->uslots[i].h2d_done, cudaEventDisableTiming));
            SQ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->uslots[i].compute_done, cudaEventDisableTiming));
            SQ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->uslots[i].d2h_done, cudaEventDisableTiming));
            SQ_CUDA(ctx, cudaHostAlloc((void **)&ctx->slots[i].h_total, 64, cudaHostAllocDefault));
        }
        SQ_CUDA(ctx, cudaMalloc(&ctx->d_work_counter, 64 * sizeof(uint32_t)));
        SQ_CUDA(ctx, cudaMemset(ctx->d_work_counter, 0, 64 * sizeof(uint32_t)));
        int32_t r = sq_xxh3_init(ctx);
        if (r) return r;
        return sq_dedup_create(ctx);
    };
    rc = init();
    if (rc) return fail(rc);
    mark("streams, events, digest constants ready (the dedup index and all scratch are allocated on first use)");
    *out = ctx;
    return SQ_OK;
}

extern "C" void sq_destroy(sq_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    sq_dedup_destroy(ctx);
    sq_enc_destroy(ctx);
    sq_dec_destroy(ctx);
    if (ctx->d_work_counter) cudaFree(ctx->d_work_counter);
    if (ctx->d_stage_in) cudaFree(ctx->d_stage_in);
    if (ctx->d_stage_out) cudaFree(ctx->d_stage_out);
    if (ctx->d_stage_meta) cudaFree(ctx->d_stage_meta);
    for (int i = 0; i < 2; i++) {
        if (ctx->slots[i].d_in) cudaFree(ctx->slots[i].d_in);
        if (ctx->slots[i].d_out) cudaFree(ctx->slots[i].d_out);
        if (ctx->slots[i].d_meta) cudaFree(ctx->slots[i].d_meta);
        if (ctx->uslots[i].d_in) cudaFree(ctx->uslots[i].d_in);
        if (ctx->uslots[i].d_out) cudaFree(ctx->uslots[i].d_out);
        if (ctx->uslots[i].d_meta) cudaFree(ctx->uslots[i].d_meta);
        if (ctx->uslots[i].h2d_done) cudaEventDestroy(ctx->uslots[i].h2d_done);
        if (ctx->uslots[i].compute_done) cudaEventDestroy(ctx->uslots[i].compute_done);
        if (ctx->uslots[i].d2h_done) cudaEventDestroy(ctx->uslots[i].d2h_done);
        if (ctx->slots[i].h2d_done) cudaEventDestroy(ctx->slots[i].h2d_done);
        if (ctx->slots[i].compute_done) cudaEventDestroy(ctx->slots[i].compute_done);
        if (ctx->slots[i].h_total) cudaFreeHost(ctx->slots[i].h_total);
        if (ctx->slot_stream[i]) cudaStreamDestroy(ctx->slot_stream[i]);
        if (ctx->dedup_done[i]) cudaEventDestroy(ctx->dedup_done[i]);
    }
    for (int i = 0; i < 2; i++) {
        if (ctx->digest_done[i]) cudaEventDestroy(ctx->digest_done[i]);
        if (ctx->verdict_done[i]) cudaEventDestroy(ctx->verdict_done[i]);
    }
    if (ctx->dedup_stream) cudaStreamDestroy(ctx->dedup_stream);
    if (ctx->d_peer_digests) cudaFree(ctx->d_peer_digests);
    if (ctx->d_peer_verdict) cudaFree(ctx->d_peer_verdict);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    free(ctx);
}

// Gives the device memory the context has cached back to the driver: encoder scratch sets, pipeline staging, decoder scratch.
// Everything is allocated again on demand.  Synchronizes the device first.
extern "C" int32_t sq_release_scratch(sq_ctx *ctx) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    SQ_CUDA(ctx, cudaSetDevice(ctx->device));
    SQ_CUDA(ctx, cudaDeviceSynchronize());
    for (int i = 0; i < 2; i++) if (ctx->slots[i].busy || ctx->uslots[i].busy) return sq_set_error(ctx, SQ_ERR_OTHER, "sq_release_scratch: a pipeline slot is in flight");
    sq_enc_destroy(ctx);
    sq_dec_destroy(ctx);
    for (int i = 0; i < 2; i++) {
        ctx->enc_set_bound[i] = 0; ctx->enc_set_stream[i] = nullptr;
        cudaFree(ctx->slots[i].d_in); cudaFree(ctx->slots[i].d_out); cudaFree(ctx->slots[i].d_meta);
        ctx->slots[i].d_in = ctx->slots[i].d_out = ctx->slots[i].d_meta = nullptr;
        ctx->slots[i].in_cap = ctx->slots[i].out_cap = ctx->slots[i].meta_cap = 0;
        cudaFree(ctx->uslots[i].d_in); cudaFree(ctx->uslots[i].d_out); cudaFree(ctx->uslots[i].d_meta);
        ctx->uslots[i].d_in = ctx->uslots[i].d_out = ctx->uslots[i].d_meta = nullptr;
        ctx->uslots[i].in_cap = ctx->uslots[i].out_cap = ctx->uslots[i].meta_cap = 0;
    }
    cudaFree(ctx->d_stage_in); cudaFree(ctx->d_stage_out); cudaFree(ctx->d_stage_meta);
    ctx->d_stage_in = ctx->d_stage_out = ctx->d_stage_meta = nullptr;
    ctx->stage_in_cap = ctx->stage_out_cap = ctx->stage_meta_cap = 0;
    ctx->enc_set_lru = 0;
    return SQ_OK;
}

extern "C" int32_t sq_synchronize(sq_ctx *ctx, void *stream) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    SQ_CUDA(ctx, cudaStreamSynchronize(sq_stream(ctx, stream)));
    return SQ_OK;
}

extern "C" int32_t sq_host_alloc(sq_ctx *ctx, size_t bytes, void **out) {
    if (!ctx || !out) return SQ_ERR_INVALID_ARG;
    SQ_CUDA(ctx, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return SQ_OK;
}
extern "C" int32_t sq_host_free(sq_ctx *ctx, void *p) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    if (p) SQ_CUDA(ctx, cudaFreeHost(p));
    return SQ_OK;
}
extern "C" void sq_free(void *p) { free(p); }

// ---- host convenience for K1 ---------------------------------------------------------
extern "C" int32_t sq_digest_host(sq_ctx *ctx, const void *h_data, size_t data_len, const sq_span *h_spans, uint32_t n,
                                  uint8_t *h_digests) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    if (n == 0) return SQ_OK;
    if (!h_spans || !h_digests || (!h_data && data_len)) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_digest_host: null pointer");
    for (uint32_t i = 0; i < n; i++)
        if (h_spans[i].off + h_spans[i].len > data_len || h_spans[i].len > ctx->chunk_size)
            return sq_set_error(ctx, SQ_ERR_INVALID_CHUNK_SIZE, "span %u [%llu,+%u) outside the batch or larger than a chunk", i,
                                (unsigned long long)h_spans[i].off, h_spans[i].len);
    int32_t rc;
    if ((rc = sq_ensure(ctx, &ctx->d_stage_in, &ctx->stage_in_cap, data_len + 64))) return rc;
    size_t meta = (size_t)n * (sizeof(sq_span) + 16);
    if ((rc = sq_ensure(ctx, &ctx->d_stage_meta, &ctx->stage_meta_cap, meta))) return rc;
    sq_span *d_spans = (sq_span *)ctx->d_stage_meta;
    uint8_t *d_dig = (uint8_t *)ctx->d_stage_meta + (size_t)n * sizeof(sq_span);
    SQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage_in, h_data, data_len, cudaMemcpyHostToDevice, ctx->stream));
    SQ_CUDA(ctx, cudaMemcpyAsync(d_spans, h_spans, (size_t)n * sizeof(sq_span), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = sq_digest_device(ctx, ctx->d_stage_in, d_spans, n, d_dig, ctx->stream))) return rc;
    SQ_CUDA(ctx, cudaMemcpyAsync(h_digests, d_dig, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    SQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SQ_OK;
}
