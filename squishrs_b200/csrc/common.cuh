// Shared internals of libsquish_b200: context, error plumbing, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/squish_b200.h"

#if defined(__CUDACC__)
#define SQ_HD __host__ __device__ __forceinline__
#else
#define SQ_HD inline
#endif

struct sq_dedup_table;  // dedup.cu
struct sq_enc_scratch;  // zstd_enc.cu
struct sq_dec_scratch;  // zstd_dec.cu

struct sq_ctx {
    int device;
    int sm_count;
    uint32_t chunk_size;
    uint32_t max_batch;
    uint32_t flags;           // sq_config.flags
    uint64_t dedup_capacity;
    cudaStream_t stream;      // the context's own stream
    cudaStream_t copy_stream; // second stream for double-buffered uploads
    char err[512];
    uint64_t launches;        // kernels launched through this context (bench: gpu_launches)
    // K1
    uint32_t *d_work_counter; // persistent-warp work queue heads (a few u32)
    // K2
    sq_dedup_table *dedup;
    // K3 / K4 scratch (lazily sized)
    sq_enc_scratch *enc_sets[2];  // two independent encoder scratch sets: the pipeline slots may encode concurrently
    sq_dec_scratch *dec;
    // staging for the *_host entry points (lazily grown)
    void *d_stage_in; size_t stage_in_cap;
    void *d_stage_out; size_t stage_out_cap;
    void *d_stage_meta; size_t stage_meta_cap;
    // double-buffered host pipeline (sq_pack_submit / sq_pack_wait): two slots, each with its own device staging
    struct pack_slot {
        void *d_in, *d_out, *d_meta; size_t in_cap, out_cap, meta_cap;
        cudaEvent_t h2d_done, compute_done;
        uint64_t *h_total;            // pinned
        uint32_t n; uint64_t out_capacity; void *h_out; int busy;
        void *h_results; const void *d_results;  // results are downloaded in sq_pack_wait when the caller's array is not pinned
    } slots[2];
    cudaStream_t d2h_stream;
    // double-buffered unpack pipeline (sq_unpack_submit / sq_unpack_wait): uploads on copy_stream, decode on stream, downloads on d2h_stream
    struct unpack_slot {
        void *d_in, *d_out, *d_meta; size_t in_cap, out_cap, meta_cap;
        cudaEvent_t h2d_done, compute_done, d2h_done;
        int busy;
        void *h_results, *h_out; const void *d_results; size_t out_len; uint32_t n; int deferred;  // pageable host buffers: downloads happen in sq_unpack_wait
    } uslots[2];
    int next_uslot;
    cudaStream_t slot_stream[2];   // one compute stream per pipeline slot: the tail of one batch overlaps the head of the next
    cudaEvent_t dedup_done[2];     // K1/K2 of consecutive batches stay ordered across the two slot streams
    int dedup_done_valid[2];
    int next_slot;
    cudaStream_t enc_set_stream[2];  // which caller stream each encoder scratch set is bound to (sq_encode_device)
    int enc_set_bound[2];
    cudaEvent_t enc_set_done[2];     // recorded after each encode on the set; a new stream taking the set over waits on it
    int enc_set_lru;
    // one dedup index for several devices (sq_share_dedup): K2 of every batch runs on the owner's device, in submission order
    sq_ctx *dedup_owner;           // nullptr or this: the context's own index
    cudaStream_t dedup_stream;     // owner side: the one stream all shared inserts go through
    void *d_peer_digests, *d_peer_verdict; size_t peer_digests_cap, peer_verdict_cap;  // owner side: one batch of digests / verdicts
    cudaEvent_t digest_done[2], verdict_done[2];  // per pipeline slot; verdict_done lives on the owner's device
};

int32_t sq_set_error(sq_ctx *ctx, int32_t code, const char *fmt, ...);
int32_t sq_ensure(sq_ctx *ctx, void **p, size_t *cap, size_t need);

#define SQ_CUDA(ctx, call)                                                                       \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return sq_set_error((ctx), SQ_ERR_CUDA, "%s failed: %s (%s:%d)", #call,              \
                                cudaGetErrorString(e__), __FILE__, __LINE__);                    \
    } while (0)

#define SQ_LAUNCHED(ctx, k) ((ctx)->launches += (k))
static inline cudaStream_t sq_stream(sq_ctx *ctx, void *s) { return s ? (cudaStream_t)s : ctx->stream; }

// entry points implemented per translation unit
int32_t sq_dedup_create(sq_ctx *ctx);
int32_t sq_dedup_ensure(sq_ctx *ctx, uint64_t capacity);  // (re)creates the index for `capacity` inserts if the current one is smaller
void sq_dedup_destroy(sq_ctx *ctx);
void sq_enc_destroy(sq_ctx *ctx);
int32_t sq_encode_device_set(sq_ctx *ctx, int set, const void *d_data, const sq_span *d_spans, const uint8_t *d_select, uint32_t n, void *d_out,
                             uint64_t out_capacity, uint64_t *d_frame_off, uint32_t *d_frame_len, uint64_t *d_total, cudaStream_t st);
void sq_dec_destroy(sq_ctx *ctx);
