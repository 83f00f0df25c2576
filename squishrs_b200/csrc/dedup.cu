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

}  // namespace

int32_t sq_dedup_create(sq_ctx *ctx) {
    sq_dedup_table *t = new sq_dedup_table();
    memset(t, 0, sizeof *t);
    uint64_t slots = 1024;
    while (slots < 2 * ctx->dedup_capacity) slots <<= 1;
    t->slots = slots;
    t->key_capacity = ctx->dedup_capacity;
    t->slot_of_cap = ctx->max_batch;
    ctx->dedup = t;
    SQ_CUDA(ctx, cudaMalloc(&t->key_ref, slots * sizeof(uint32_t)));
    SQ_CUDA(ctx, cudaMalloc(&t->min_gidx, slots * sizeof(unsigned long long)));
    SQ_CUDA(ctx, cudaMalloc(&t->keys, t->key_capacity * sizeof(uint4)));
    SQ_CUDA(ctx, cudaMalloc(&t->slot_of, (size_t)t->slot_of_cap * sizeof(uint32_t)));
    SQ_CUDA(ctx, cudaMalloc(&t->counters, 4 * sizeof(unsigned long long)));
    return sq_dedup_reset(ctx);
}

void sq_dedup_destroy(sq_ctx *ctx) {
    sq_dedup_table *t = ctx->dedup;
    if (!t) return;
    cudaFree(t->key_ref); cudaFree(t->min_gidx); cudaFree(t->keys); cudaFree(t->slot_of); cudaFree(t->counters);
    delete t;
    ctx->dedup = nullptr;
}

extern "C" int32_t sq_dedup_reset(sq_ctx *ctx) {
    if (!ctx || !ctx->dedup) return SQ_ERR_INVALID_ARG;
    sq_dedup_table *t = ctx->dedup;
    SQ_CUDA(ctx, cudaMemsetAsync(t->key_ref, 0, t->slots * sizeof(uint32_t), ctx->stream));
    SQ_CUDA(ctx, cudaMemsetAsync(t->min_gidx, 0xFF, t->slots * sizeof(unsigned long long), ctx->stream));
    SQ_CUDA(ctx, cudaMemsetAsync(t->counters, 0, 4 * sizeof(unsigned long long), ctx->stream));
    SQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SQ_OK;
}

extern "C" int32_t sq_dedup_insert_device(sq_ctx *ctx, const void *d_digests, const uint64_t *d_gidx, uint64_t gidx_base,
                                          uint32_t n, uint8_t *d_is_new, void *stream) {
    if (!ctx || !ctx->dedup) return SQ_ERR_INVALID_ARG;
    if (n == 0) return SQ_OK;
    if (!d_digests || !d_is_new) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_dedup_insert_device: null pointer");
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
    if (!ctx || !ctx->dedup || !out) return SQ_ERR_INVALID_ARG;
    unsigned long long c[4];
    SQ_CUDA(ctx, cudaDeviceSynchronize());
    SQ_CUDA(ctx, cudaMemcpy(c, ctx->dedup->counters, sizeof c, cudaMemcpyDeviceToHost));
    if (c[2]) return sq_set_error(ctx, SQ_ERR_CAPACITY, "dedup index overflow: more than %llu inserts", (unsigned long long)ctx->dedup->key_capacity);
    *out = c[1];
    return SQ_OK;
}
