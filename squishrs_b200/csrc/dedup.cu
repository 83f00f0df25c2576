// K2 — dedup index: an open-addressed hash set of 128-bit digests in HBM.
//
// Replaces the ChunkStore set `Arc<DashMap<ChunkHash,()>>` and its entry() protocol
// (reference src/util/chunk.rs:19-25,83-99) plus len() (chunk.rs:116-118).
//
// Layout.  Two parallel arrays of `slots` (power of two >= 2 x capacity):
//   key_ref[s]  u32   0 = empty, else 1 + index into `keys` (claimed with atomicCAS)
//   min_gidx[s] u64   the lowest global chunk index that has presented this digest (atomicMin)
// `keys` is an append-only array of the digests that claimed a slot; once written an entry
// never changes, so probing threads compare against immutable memory and no 128-bit CAS is
// needed.  A key is appended (atomicAdd on `count`) BEFORE the slot CAS publishes its index;
// a loser of the CAS leaves its appended key orphaned, so key storage is sized for every
// insert ever made (capacity counts inserts, not distinct digests, at 16 B each).
//
// Winner rule: is_new[i] = 1 iff chunk i carries the lowest global index
// seen for its digest.  Pass 1 inserts and records each chunk's slot; pass 2 (a second
// kernel, i.e. after a grid-wide barrier) reads min_gidx back.  Batches must be submitted
// in ascending gidx order for cross-batch determinism, as the packer does.
#include "common.cuh"

struct sq_dedup_table {
    uint32_t *key_ref;
    unsigned long long *min_gidx;
    uint4 *keys;
    uint32_t *slot_of;  // per batch item: slot index found in pass 1
    unsigned long long *counters;  // [0] = appended keys, [1] = distinct digests, [2] = overflow flag
    uint64_t slots, key_capacity;
    uint32_t slot_of_cap;
};

namespace {

__device__ __forceinline__ uint64_t slot_hash(uint4 d) {
    // the digest is already a high-quality hash: fold the two halves
    uint64_t lo = (uint64_t)d.x | (uint64_t)d.y << 32, hi = (uint64_t)d.z | (uint64_t)d.w << 32;
    return lo ^ (hi * 0x9E3779B97F4A7C15ULL);
}

__global__ void dedup_insert_kernel(const uint4 *__restrict__ digests, const unsigned long long *__restrict__ gidx,
                                    unsigned long long gidx_base, uint32_t n, uint32_t *key_ref, unsigned long long *min_gidx,
                                    uint4 *keys, uint32_t *slot_of, unsigned long long *counters, uint64_t mask,
                                    uint64_t key_capacity) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 d = digests[i];
    const unsigned long long g = gidx ? gidx[i] : gidx_base + i;
    if (g == ~0ULL) { slot_of[i] = 0xFFFFFFFFu; return; }  // padding record of a routed exchange
    uint64_t s = slot_hash(d) & mask;
    uint32_t my_ref = 0;
    for (uint64_t probes = 0; probes <= mask; probes++, s = (s + 1) & mask) {
        uint32_t cur = *(volatile uint32_t *)&key_ref[s];
        if (cur == 0) {
            if (my_ref == 0) {  // append our digest once, then try to publish it
                unsigned long long k = atomicAdd(&counters[0], 1ULL);
                if (k >= key_capacity) { atomicExch(&counters[2], 1ULL); slot_of[i] = 0xFFFFFFFFu; return; }
                keys[k] = d;
                __threadfence();
                my_ref = (uint32_t)k + 1;
            }
            cur = atomicCAS(&key_ref[s], 0u, my_ref);
            if (cur == 0) {  // Entry::Vacant
                atomicAdd(&counters[1], 1ULL);
                atomicMin(&min_gidx[s], g);
                slot_of[i] = (uint32_t)s;
                return;
            }
        }
        __threadfence();
        const uint4 k = __ldcg(&keys[cur - 1]);
        if (k.x == d.x && k.y == d.y && k.z == d.z && k.w == d.w) {  // Entry::Occupied: digest equality is identity
            atomicMin(&min_gidx[s], g);
            slot_of[i] = (uint32_t)s;
            return;
        }
    }
    atomicExch(&counters[2], 1ULL);
    slot_of[i] = 0xFFFFFFFFu;
}

__global__ void dedup_verdict_kernel(const unsigned long long *__restrict__ gidx, unsigned long long gidx_base, uint32_t n,
                                     const unsigned long long *__restrict__ min_gidx, const uint32_t *__restrict__ slot_of,
                                     uint8_t *__restrict__ is_new) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long g = gidx ? gidx[i] : gidx_base + i;
    const uint32_t s = slot_of[i];
    is_new[i] = (s != 0xFFFFFFFFu && min_gidx[s] == g) ? 1 : 0;
}


// ---- multi-GPU: the index is sharded by digest prefix; records travel as 32-byte {digest, gidx, pad} ----
struct RoutedRec { uint4 digest; unsigned long long gidx, pad; };
static_assert(sizeof(RoutedRec) == 32, "routed record layout");

__global__ void route_kernel(const uint4 *__restrict__ digests, unsigned long long gidx_base, uint32_t n, uint32_t world, uint32_t cap,
                             RoutedRec *__restrict__ send, uint32_t *__restrict__ send_pos, uint32_t *__restrict__ peer_count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 d = digests[i];
    const unsigned long long prefix = (unsigned long long)d.x | (unsigned long long)d.y << 32;  // first 8 digest bytes, LE
    const uint32_t owner = (uint32_t)(prefix % world);
    const uint32_t k = atomicAdd(&peer_count[owner], 1u);  // k < cap because cap >= n
    RoutedRec r;
    r.digest = d; r.gidx = gidx_base + i; r.pad = 0;
    send[(size_t)owner * cap + k] = r;
    send_pos[i] = owner * cap + k;
}
__global__ void route_pad_kernel(RoutedRec *__restrict__ send, uint32_t total) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    RoutedRec r;
    r.digest = make_uint4(0, 0, 0, 0); r.gidx = ~0ULL; r.pad = 0;
    send[i] = r;
}
__global__ void routed_split_kernel(const RoutedRec *__restrict__ recv, uint32_t count, uint4 *__restrict__ digests,
                                    unsigned long long *__restrict__ gidx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    digests[i] = recv[i].digest;
    gidx[i] = recv[i].gidx;
}
__global__ void unroute_kernel(const uint8_t *__restrict__ verdict_back, const uint32_t *__restrict__ send_pos, uint32_t n,
                               uint8_t *__restrict__ is_new) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) is_new[i] = verdict_back[send_pos[i]];
}

}  // namespace

// The index is created on first use, sized for `capacity` inserts (a job that never dedups -- unpack, list -- never pays for it,
// and the archive packer sizes it to the job's chunk count instead of the context's maximum).
int32_t sq_dedup_ensure(sq_ctx *ctx, uint64_t capacity) {
    if (capacity < 1024) capacity = 1024;
    if (capacity >= (1ull << 30)) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "dedup index for %llu inserts: not below 2^30", (unsigned long long)capacity);
    if (ctx->dedup && ctx->dedup->key_capacity >= capacity) return SQ_OK;
    if (ctx->dedup) { SQ_CUDA(ctx, cudaDeviceSynchronize()); sq_dedup_destroy(ctx); }
    sq_dedup_table *t = new sq_dedup_table();
    memset(t, 0, sizeof *t);
    uint64_t slots = 1024;
    while (slots < 2 * capacity) slots <<= 1;
    t->slots = slots;
    t->key_capacity = capacity;
    t->slot_of_cap = ctx->max_batch;
    ctx->dedup = t;
    SQ_CUDA(ctx, cudaMalloc(&t->key_ref, slots * sizeof(uint32_t)));
    SQ_CUDA(ctx, cudaMalloc(&t->min_gidx, slots * sizeof(unsigned long long)));
    SQ_CUDA(ctx, cudaMalloc(&t->keys, t->key_capacity * sizeof(uint4)));
    SQ_CUDA(ctx, cudaMalloc(&t->slot_of, (size_t)t->slot_of_cap * sizeof(uint32_t)));
    SQ_CUDA(ctx, cudaMalloc(&t->counters, 4 * sizeof(unsigned long long)));
    return sq_dedup_reset(ctx);
}

int32_t sq_dedup_create(sq_ctx *ctx) { (void)ctx; return SQ_OK; }  // lazily: see sq_dedup_ensure

void sq_dedup_destroy(sq_ctx *ctx) {
    sq_dedup_table *t = ctx->dedup;
    if (!t) return;
    cudaFree(t->key_ref); cudaFree(t->min_gidx); cudaFree(t->keys); cudaFree(t->slot_of); cudaFree(t->counters);
    delete t;
    ctx->dedup = nullptr;
}

extern "C" int32_t sq_dedup_reset(sq_ctx *ctx) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    if (!ctx->dedup) return SQ_OK;  // nothing has been inserted yet: the index is created (empty) on first use
    sq_dedup_table *t = ctx->dedup;
    SQ_CUDA(ctx, cudaMemsetAsync(t->key_ref, 0, t->slots * sizeof(uint32_t), ctx->stream));
    SQ_CUDA(ctx, cudaMemsetAsync(t->min_gidx, 0xFF, t->slots * sizeof(unsigned long long), ctx->stream));
    SQ_CUDA(ctx, cudaMemsetAsync(t->counters, 0, 4 * sizeof(unsigned long long), ctx->stream));
    SQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SQ_OK;
}

extern "C" int32_t sq_dedup_insert_device(sq_ctx *ctx, const void *d_digests, const uint64_t *d_gidx, uint64_t gidx_base,
                                          uint32_t n, uint8_t *d_is_new, void *stream) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    if (n == 0) return SQ_OK;
    if (!d_digests || !d_is_new) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_dedup_insert_device: null pointer");
    if (!ctx->dedup) { int32_t r = sq_dedup_ensure(ctx, ctx->dedup_capacity); if (r) return r; }
    sq_dedup_table *t = ctx->dedup;
    if (n > t->slot_of_cap) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "batch of %u chunks exceeds max_batch_chunks %u", n, t->slot_of_cap);
    cudaStream_t st = sq_stream(ctx, stream);
    const uint32_t tpb = 128, blocks = (n + tpb - 1) / tpb;
    dedup_insert_kernel<<<blocks, tpb, 0, st>>>((const uint4 *)d_digests, (const unsigned long long *)d_gidx, gidx_base, n, t->key_ref,
                                                 t->min_gidx, t->keys, t->slot_of, t->counters, t->slots - 1, t->key_capacity);
    dedup_verdict_kernel<<<blocks, tpb, 0, st>>>((const unsigned long long *)d_gidx, gidx_base, n, t->min_gidx, t->slot_of, d_is_new);
    SQ_LAUNCHED(ctx, 2);
    SQ_CUDA(ctx, cudaGetLastError());
    return SQ_OK;
}

extern "C" int32_t sq_dedup_len(sq_ctx *ctx, uint64_t *out) {
    if (!ctx || !out) return SQ_ERR_INVALID_ARG;
    if (!ctx->dedup) { *out = 0; return SQ_OK; }
    unsigned long long c[4];
    SQ_CUDA(ctx, cudaDeviceSynchronize());
    SQ_CUDA(ctx, cudaMemcpy(c, ctx->dedup->counters, sizeof c, cudaMemcpyDeviceToHost));
    if (c[2]) return sq_set_error(ctx, SQ_ERR_CAPACITY, "dedup index overflow: more than %llu inserts", (unsigned long long)ctx->dedup->key_capacity);
    *out = c[1];
    return SQ_OK;
}

// ---- sharded index entry points ---------------------------------------------------------
extern "C" int32_t sq_route_digests_device(sq_ctx *ctx, const void *d_digests, uint64_t gidx_base, uint32_t n, uint32_t world,
                                           uint32_t cap_per_peer, void *d_send, uint32_t *d_send_pos, void *stream) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    if (!ctx->dedup) { int32_t r = sq_dedup_ensure(ctx, ctx->dedup_capacity); if (r) return r; }
    if (!d_digests || !d_send || !d_send_pos || world == 0 || cap_per_peer < n)
        return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_route_digests_device: bad arguments (cap_per_peer must be >= n)");
    if (world > 60) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "at most 60 ranks");
    cudaStream_t st = sq_stream(ctx, stream);
    uint32_t *peer_count = ctx->d_work_counter + 4;  // world <= 60 counters
    SQ_CUDA(ctx, cudaMemsetAsync(peer_count, 0, world * sizeof(uint32_t), st));
    const uint32_t total = world * cap_per_peer;
    route_pad_kernel<<<(total + 255) / 256, 256, 0, st>>>((RoutedRec *)d_send, total);
    if (n) route_kernel<<<(n + 127) / 128, 128, 0, st>>>((const uint4 *)d_digests, gidx_base, n, world, cap_per_peer, (RoutedRec *)d_send, d_send_pos, peer_count);
    SQ_LAUNCHED(ctx, 2);
    SQ_CUDA(ctx, cudaGetLastError());
    return SQ_OK;
}

extern "C" int32_t sq_dedup_insert_routed_device(sq_ctx *ctx, const void *d_recv, uint32_t count, uint8_t *d_verdict, void *stream) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    if (count == 0) return SQ_OK;
    if (!ctx->dedup) { int32_t r = sq_dedup_ensure(ctx, ctx->dedup_capacity); if (r) return r; }
    if (!d_recv || !d_verdict) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_dedup_insert_routed_device: null pointer");
    cudaStream_t st = sq_stream(ctx, stream);
    int32_t rc = sq_ensure(ctx, &ctx->d_stage_meta, &ctx->stage_meta_cap, (size_t)count * 24 + 256);
    if (rc) return rc;
    uint4 *dig = (uint4 *)ctx->d_stage_meta;
    unsigned long long *gidx = (unsigned long long *)((uint8_t *)ctx->d_stage_meta + (size_t)count * 16);
    routed_split_kernel<<<(count + 255) / 256, 256, 0, st>>>((const RoutedRec *)d_recv, count, dig, gidx);
    SQ_LAUNCHED(ctx, 1);
    return sq_dedup_insert_device(ctx, dig, (const uint64_t *)gidx, 0, count, d_verdict, st);
}

extern "C" int32_t sq_unroute_verdicts_device(sq_ctx *ctx, const uint8_t *d_verdict_back, const uint32_t *d_send_pos, uint32_t n,
                                              uint8_t *d_is_new, void *stream) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    if (n == 0) return SQ_OK;
    if (!d_verdict_back || !d_send_pos || !d_is_new) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_unroute_verdicts_device: null pointer");
    unroute_kernel<<<(n + 255) / 256, 256, 0, sq_stream(ctx, stream)>>>(d_verdict_back, d_send_pos, n, d_is_new);
    SQ_LAUNCHED(ctx, 1);
    SQ_CUDA(ctx, cudaGetLastError());
    return SQ_OK;
}
