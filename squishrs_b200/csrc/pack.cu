// Fused batch entry points: ChunkStore::insert for a whole batch (reference src/util/chunk.rs:80-100)
// = K1 digest -> K2 dedup -> K3 encode of the winners, and read_chunks for a batch
// (src/archive/reader.rs:259-314) = K4 decode.  Host variants add the H2D / D2H copies.
#include "common.cuh"

extern "C" int32_t sq_encode_status(sq_ctx *ctx);

namespace {
__global__ void pack_results_kernel(const uint4 *__restrict__ digests, const uint8_t *__restrict__ is_new,
                                    const uint64_t *__restrict__ frame_off, const uint32_t *__restrict__ frame_len, uint32_t n,
                                    sq_chunk_result *__restrict__ res) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sq_chunk_result r;
    *reinterpret_cast<uint4 *>(r.digest) = digests[i];
    r.is_new = is_new[i];
    r.frame_off = r.is_new ? frame_off[i] : 0;
    r.frame_len = r.is_new ? frame_len[i] : 0;
    r.reserved[0] = r.reserved[1] = r.reserved[2] = 0;
    res[i] = r;
}
}  // namespace

// per-batch metadata scratch layout inside ctx->d_stage_meta (device)
struct pack_meta {
    uint8_t *digests; uint8_t *is_new; uint64_t *frame_off; uint32_t *frame_len; uint64_t *total;
    sq_span *spans; sq_chunk_result *results; sq_frame *frames; sq_frame_result *fres;
};
static int32_t meta_layout_in(sq_ctx *ctx, void **buf, size_t *cap, uint32_t n, pack_meta *m) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    size_t o_dig = take((size_t)n * 16), o_new = take(n), o_fo = take((size_t)n * 8), o_fl = take((size_t)n * 4), o_tot = take(8),
           o_sp = take((size_t)n * sizeof(sq_span)), o_res = take((size_t)n * sizeof(sq_chunk_result)),
           o_fr = take((size_t)n * sizeof(sq_frame)), o_frs = take((size_t)n * sizeof(sq_frame_result));
    int32_t rc = sq_ensure(ctx, buf, cap, off);
    if (rc) return rc;
    uint8_t *b = (uint8_t *)*buf;
    m->digests = b + o_dig; m->is_new = b + o_new; m->frame_off = (uint64_t *)(b + o_fo); m->frame_len = (uint32_t *)(b + o_fl);
    m->total = (uint64_t *)(b + o_tot); m->spans = (sq_span *)(b + o_sp); m->results = (sq_chunk_result *)(b + o_res);
    m->frames = (sq_frame *)(b + o_fr); m->fres = (sq_frame_result *)(b + o_frs);
    return SQ_OK;
}

static int32_t meta_layout(sq_ctx *ctx, uint32_t n, pack_meta *m) { return meta_layout_in(ctx, &ctx->d_stage_meta, &ctx->stage_meta_cap, n, m); }

// digest -> dedup -> encode -> results, all asynchronous on `st`, with caller-provided metadata scratch
static int32_t pack_device_impl(sq_ctx *ctx, const void *d_data, const sq_span *d_spans, uint32_t n, uint64_t gidx_base, const pack_meta &m,
                                sq_chunk_result *d_results, void *d_out, uint64_t out_capacity, cudaStream_t st, int set = 0,
                                cudaEvent_t after_dedup = nullptr) {
    int32_t rc;
    if ((rc = sq_digest_device(ctx, d_data, d_spans, n, m.digests, st))) return rc;
    sq_ctx *own = ctx->dedup_owner;
    if (own && ctx->digest_done[set]) {
        // Shared index (several devices, one process): the digests travel to the owner's device, the insert runs there on the one
        // stream every shared insert goes through -- submission order is chunk order, so "lowest chunk index wins" is decided
        // exactly as on one device -- and the verdict bytes travel back.  16 + 1 bytes per chunk over NVLink.
        SQ_CUDA(ctx, cudaEventRecord(ctx->digest_done[set], st));
        SQ_CUDA(ctx, cudaSetDevice(own->device));
        cudaStream_t ds = own->dedup_stream;
        rc = sq_ensure(own, &own->d_peer_digests, &own->peer_digests_cap, (size_t)n * 16);
        if (!rc) rc = sq_ensure(own, &own->d_peer_verdict, &own->peer_verdict_cap, n);
        cudaError_t ce = cudaSuccess;
        if (!rc) {
            if ((ce = cudaStreamWaitEvent(ds, ctx->digest_done[set], 0)) == cudaSuccess &&
                (ce = cudaMemcpyPeerAsync(own->d_peer_digests, own->device, m.digests, ctx->device, (size_t)n * 16, ds)) == cudaSuccess) {
                rc = sq_dedup_insert_device(own, own->d_peer_digests, nullptr, gidx_base, n, (uint8_t *)own->d_peer_verdict, ds);
                if (!rc && (ce = cudaMemcpyPeerAsync(m.is_new, ctx->device, own->d_peer_verdict, own->device, n, ds)) == cudaSuccess)
                    ce = cudaEventRecord(ctx->verdict_done[set], ds);
            }
        }
        cudaSetDevice(ctx->device);
        if (rc) { if (own != ctx) snprintf(ctx->err, sizeof ctx->err, "%s", own->err); return rc; }
        if (ce != cudaSuccess) return sq_set_error(ctx, SQ_ERR_CUDA, "shared dedup insert failed: %s", cudaGetErrorString(ce));
        SQ_CUDA(ctx, cudaStreamWaitEvent(st, ctx->verdict_done[set], 0));
    } else if ((rc = sq_dedup_insert_device(ctx, m.digests, nullptr, gidx_base, n, m.is_new, st))) return rc;
    if (after_dedup) SQ_CUDA(ctx, cudaEventRecord(after_dedup, st));
    if ((rc = sq_encode_device_set(ctx, set, d_data, d_spans, m.is_new, n, d_out, out_capacity, m.frame_off, m.frame_len, m.total, st))) return rc;
    pack_results_kernel<<<(n + 127) / 128, 128, 0, st>>>((const uint4 *)m.digests, m.is_new, m.frame_off, m.frame_len, n, d_results);
    SQ_LAUNCHED(ctx, 1);
    SQ_CUDA(ctx, cudaGetLastError());
    return SQ_OK;
}

extern "C" int32_t sq_pack_device(sq_ctx *ctx, const void *d_data, const sq_span *d_spans, uint32_t n, uint64_t gidx_base,
                                  sq_chunk_result *d_results, void *d_out, uint64_t out_capacity, uint64_t *out_used, void *stream) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    if (n == 0) { if (out_used) *out_used = 0; return SQ_OK; }
    if (!d_data || !d_spans || !d_results || !d_out) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_pack_device: null pointer");
    pack_meta m;
    int32_t rc = meta_layout(ctx, n, &m);
    if (rc) return rc;
    cudaStream_t st = sq_stream(ctx, stream);
    if ((rc = pack_device_impl(ctx, d_data, d_spans, n, gidx_base, m, d_results, d_out, out_capacity, st))) return rc;
    if (out_used) {
        SQ_CUDA(ctx, cudaMemcpyAsync(out_used, m.total, 8, cudaMemcpyDeviceToHost, st));
        SQ_CUDA(ctx, cudaStreamSynchronize(st));
        if ((rc = sq_encode_status(ctx))) return rc;
    }
    return SQ_OK;
}

// Several devices, one index: after this call every pack of `ctx` (sq_pack_submit / sq_pack_device) decides "new or duplicate" in
// `owner`'s index.  The owner shares with itself so that all contexts go through the same ordered stream.  One host thread must
// submit the batches of all sharing contexts, in chunk order.
extern "C" int32_t sq_share_dedup(sq_ctx *ctx, sq_ctx *owner) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    if (!owner) { ctx->dedup_owner = nullptr; return SQ_OK; }  // back to the context's own index
    SQ_CUDA(owner, cudaSetDevice(owner->device));
    if (!owner->dedup_stream) SQ_CUDA(owner, cudaStreamCreateWithFlags(&owner->dedup_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) if (!ctx->verdict_done[i]) SQ_CUDA(owner, cudaEventCreateWithFlags(&ctx->verdict_done[i], cudaEventDisableTiming));
    if (owner != ctx) { cudaDeviceEnablePeerAccess(ctx->device, 0); cudaGetLastError(); }
    SQ_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < 2; i++) if (!ctx->digest_done[i]) SQ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->digest_done[i], cudaEventDisableTiming));
    if (owner != ctx) { cudaDeviceEnablePeerAccess(owner->device, 0); cudaGetLastError(); }
    ctx->dedup_owner = owner;
    return SQ_OK;
}

// ---- double-buffered host pipeline ------------------------------------------------------------------------------
// submit: H2D of the batch on the copy stream -> kernels on the compute stream -> results + total D2H; returns at once.
// wait:   blocks until the batch's kernels are done, then copies exactly the used frame bytes D2H on a third stream.
// With two tickets in flight the upload of batch k+1 and the frame download of batch k-1 overlap the kernels of batch k
// (north_star: "pinned, double-buffered cudaMemcpyAsync uploads").
struct sq_ticket { int slot; };
static sq_ticket g_tickets[2] = {{0}, {1}};

// cudaMemcpyAsync device -> PAGEABLE host memory returns only when the copy has completed: issued inside a submit call it would
// block the caller until the whole batch is done.  Downloads into memory that is not pinned are therefore left to the wait call.
static bool host_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

extern "C" int32_t sq_pack_submit(sq_ctx *ctx, const void *h_data, size_t data_len, const sq_span *h_spans, uint32_t n,
                                  uint64_t gidx_base, sq_chunk_result *h_results, void *h_out, uint64_t out_capacity, sq_ticket **ticket) {
    if (!ctx || !ticket) return SQ_ERR_INVALID_ARG;
    if (n == 0 || !h_spans || !h_results || !h_out || (!h_data && data_len)) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_pack_submit: bad arguments");
    uint64_t bound = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (h_spans[i].off + h_spans[i].len > data_len || h_spans[i].len > ctx->chunk_size || h_spans[i].len == 0)
            return sq_set_error(ctx, SQ_ERR_INVALID_CHUNK_SIZE, "span %u [%llu,+%u) invalid", i, (unsigned long long)h_spans[i].off, h_spans[i].len);
        bound += sq_encode_bound(h_spans[i].len);
    }
    const int si = ctx->next_slot;
    sq_ctx::pack_slot &sl = ctx->slots[si];
    if (sl.busy) return sq_set_error(ctx, SQ_ERR_OTHER, "sq_pack_submit: both pipeline slots are in flight; call sq_pack_wait first");
    int32_t rc;
    if ((rc = sq_ensure(ctx, &sl.d_in, &sl.in_cap, data_len + 64))) return rc;
    if ((rc = sq_ensure(ctx, &sl.d_out, &sl.out_cap, bound + 64))) return rc;
    pack_meta m;
    if ((rc = meta_layout_in(ctx, &sl.d_meta, &sl.meta_cap, n, &m))) return rc;
    SQ_CUDA(ctx, cudaMemcpyAsync(sl.d_in, h_data, data_len, cudaMemcpyHostToDevice, ctx->copy_stream));
    SQ_CUDA(ctx, cudaMemcpyAsync(m.spans, h_spans, (size_t)n * sizeof(sq_span), cudaMemcpyHostToDevice, ctx->copy_stream));
    SQ_CUDA(ctx, cudaEventRecord(sl.h2d_done, ctx->copy_stream));
    // Each slot has its own compute stream and encoder scratch set: the long encode of this batch may overlap the tail of the
    // previous one (a chunk occupies one search CTA for ~0.1 s, so tails would otherwise idle most SMs).  Digest and dedup of
    // consecutive batches stay ordered through dedup_done (they share the digest work counter and the index's batch scratch).
    cudaStream_t cs = ctx->slot_stream[si];
    SQ_CUDA(ctx, cudaStreamWaitEvent(cs, sl.h2d_done, 0));
    if (ctx->dedup_done_valid[si ^ 1]) SQ_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->dedup_done[si ^ 1], 0));
    if ((rc = pack_device_impl(ctx, sl.d_in, m.spans, n, gidx_base, m, m.results, sl.d_out, bound, cs, si, ctx->dedup_done[si]))) return rc;
    ctx->dedup_done_valid[si] = 1;
    sl.h_results = nullptr;
    if (host_pinned(h_results)) SQ_CUDA(ctx, cudaMemcpyAsync(h_results, m.results, (size_t)n * sizeof(sq_chunk_result), cudaMemcpyDeviceToHost, cs));
    else { sl.h_results = h_results; sl.d_results = m.results; }
    SQ_CUDA(ctx, cudaMemcpyAsync(sl.h_total, m.total, 8, cudaMemcpyDeviceToHost, cs));
    SQ_CUDA(ctx, cudaEventRecord(sl.compute_done, cs));
    sl.busy = 1; sl.n = n; sl.out_capacity = out_capacity; sl.h_out = h_out;
    ctx->next_slot = si ^ 1;
    *ticket = &g_tickets[si];
    return SQ_OK;
}

extern "C" int32_t sq_pack_wait(sq_ctx *ctx, sq_ticket *ticket, uint64_t *out_used) {
    if (!ctx || !ticket) return SQ_ERR_INVALID_ARG;
    sq_ctx::pack_slot &sl = ctx->slots[ticket->slot];
    if (!sl.busy) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_pack_wait: ticket is not in flight");
    sl.busy = 0;
    SQ_CUDA(ctx, cudaEventSynchronize(sl.compute_done));
    if (sl.h_results) SQ_CUDA(ctx, cudaMemcpyAsync(sl.h_results, sl.d_results, (size_t)sl.n * sizeof(sq_chunk_result), cudaMemcpyDeviceToHost, ctx->d2h_stream));
    const uint64_t used = *sl.h_total;
    if (used > sl.out_capacity)
        return sq_set_error(ctx, SQ_ERR_CAPACITY, "pack output needs %llu bytes, caller gave %llu", (unsigned long long)used, (unsigned long long)sl.out_capacity);
    if (used) SQ_CUDA(ctx, cudaMemcpyAsync(sl.h_out, sl.d_out, used, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    if (used || sl.h_results) SQ_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));
    if (out_used) *out_used = used;
    return SQ_OK;
}

extern "C" int32_t sq_pack_host(sq_ctx *ctx, const void *h_data, size_t data_len, const sq_span *h_spans, uint32_t n,
                                uint64_t gidx_base, sq_chunk_result *h_results, void *h_out, uint64_t out_capacity, uint64_t *out_used) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    if (n == 0) { if (out_used) *out_used = 0; return SQ_OK; }
    sq_ticket *t = nullptr;
    int32_t rc = sq_pack_submit(ctx, h_data, data_len, h_spans, n, gidx_base, h_results, h_out, out_capacity, &t);
    if (rc) return rc;
    return sq_pack_wait(ctx, t, out_used);
}

// ---- double-buffered unpack pipeline ------------------------------------------------------------------------------
// submit: payload upload on the copy stream -> K4 on the context stream -> results + decoded bytes download on the
// d2h stream; returns at once.  With two tickets in flight the upload of batch k+1 and the download of batch k-1
// overlap the decode of batch k, so unpack runs at the slower of the decoder and the host link.
static sq_ticket g_utickets[2] = {{0}, {1}};

extern "C" int32_t sq_unpack_submit(sq_ctx *ctx, const void *h_comp, size_t comp_len, const sq_frame *h_frames, uint32_t n, void *h_out,
                                    size_t out_len, sq_frame_result *h_results, sq_ticket **ticket) {
    if (!ctx || !ticket) return SQ_ERR_INVALID_ARG;
    if (n == 0 || !h_comp || !h_frames || !h_results || (!h_out && out_len)) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_unpack_submit: bad arguments");
    for (uint32_t i = 0; i < n; i++) {
        if (h_frames[i].src_off + h_frames[i].src_len > comp_len)
            return sq_set_error(ctx, SQ_ERR_READER, "frame %u payload outside the compressed buffer", i);
        if (h_frames[i].dst_off + h_frames[i].capacity > out_len)
            return sq_set_error(ctx, SQ_ERR_CAPACITY, "frame %u capacity outside the output buffer", i);
    }
    const int si = ctx->next_uslot;
    sq_ctx::unpack_slot &sl = ctx->uslots[si];
    if (sl.busy) return sq_set_error(ctx, SQ_ERR_OTHER, "sq_unpack_submit: both pipeline slots are in flight; call sq_unpack_wait first");
    int32_t rc;
    if ((rc = sq_ensure(ctx, &sl.d_in, &sl.in_cap, comp_len + 64))) return rc;
    if ((rc = sq_ensure(ctx, &sl.d_out, &sl.out_cap, out_len + 64))) return rc;
    pack_meta m;
    if ((rc = meta_layout_in(ctx, &sl.d_meta, &sl.meta_cap, n, &m))) return rc;
    SQ_CUDA(ctx, cudaMemcpyAsync(sl.d_in, h_comp, comp_len, cudaMemcpyHostToDevice, ctx->copy_stream));
    SQ_CUDA(ctx, cudaMemcpyAsync(m.frames, h_frames, (size_t)n * sizeof(sq_frame), cudaMemcpyHostToDevice, ctx->copy_stream));
    SQ_CUDA(ctx, cudaEventRecord(sl.h2d_done, ctx->copy_stream));
    SQ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, sl.h2d_done, 0));
    if ((rc = sq_decode_device(ctx, sl.d_in, m.frames, n, sl.d_out, m.fres, ctx->stream))) return rc;
    SQ_CUDA(ctx, cudaEventRecord(sl.compute_done, ctx->stream));
    sl.deferred = !(host_pinned(h_results) && (!out_len || host_pinned(h_out)));
    sl.h_results = h_results; sl.h_out = h_out; sl.d_results = m.fres; sl.out_len = out_len; sl.n = n;
    if (!sl.deferred) {
        SQ_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h_stream, sl.compute_done, 0));
        SQ_CUDA(ctx, cudaMemcpyAsync(h_results, m.fres, (size_t)n * sizeof(sq_frame_result), cudaMemcpyDeviceToHost, ctx->d2h_stream));
        if (out_len) SQ_CUDA(ctx, cudaMemcpyAsync(h_out, sl.d_out, out_len, cudaMemcpyDeviceToHost, ctx->d2h_stream));
        SQ_CUDA(ctx, cudaEventRecord(sl.d2h_done, ctx->d2h_stream));
    }
    sl.busy = 1;
    ctx->next_uslot = si ^ 1;
    *ticket = &g_utickets[si];
    return SQ_OK;
}

extern "C" int32_t sq_unpack_wait(sq_ctx *ctx, sq_ticket *ticket) {
    if (!ctx || !ticket) return SQ_ERR_INVALID_ARG;
    sq_ctx::unpack_slot &sl = ctx->uslots[ticket->slot];
    if (!sl.busy) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_unpack_wait: ticket is not in flight");
    sl.busy = 0;
    if (sl.deferred) {  // pageable destination: the downloads block, so they happen here and not in the submit call
        SQ_CUDA(ctx, cudaEventSynchronize(sl.compute_done));
        SQ_CUDA(ctx, cudaMemcpyAsync(sl.h_results, sl.d_results, (size_t)sl.n * sizeof(sq_frame_result), cudaMemcpyDeviceToHost, ctx->d2h_stream));
        if (sl.out_len) SQ_CUDA(ctx, cudaMemcpyAsync(sl.h_out, sl.d_out, sl.out_len, cudaMemcpyDeviceToHost, ctx->d2h_stream));
        SQ_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));
        return SQ_OK;
    }
    SQ_CUDA(ctx, cudaEventSynchronize(sl.d2h_done));
    return SQ_OK;
}

extern "C" int32_t sq_unpack_host(sq_ctx *ctx, const void *h_comp, size_t comp_len, const sq_frame *h_frames, uint32_t n, void *h_out,
                                  size_t out_len, sq_frame_result *h_results) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    if (n == 0) return SQ_OK;
    if (!h_comp || !h_frames || !h_results || (!h_out && out_len)) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_unpack_host: null pointer");
    for (uint32_t i = 0; i < n; i++) {
        if (h_frames[i].src_off + h_frames[i].src_len > comp_len)
            return sq_set_error(ctx, SQ_ERR_READER, "frame %u payload outside the compressed buffer", i);
        if (h_frames[i].dst_off + h_frames[i].capacity > out_len)
            return sq_set_error(ctx, SQ_ERR_CAPACITY, "frame %u capacity outside the output buffer", i);
    }
    int32_t rc;
    if ((rc = sq_ensure(ctx, &ctx->d_stage_in, &ctx->stage_in_cap, comp_len + 64))) return rc;
    if ((rc = sq_ensure(ctx, &ctx->d_stage_out, &ctx->stage_out_cap, out_len + 64))) return rc;
    pack_meta m;
    if ((rc = meta_layout(ctx, n, &m))) return rc;
    cudaStream_t st = ctx->stream;
    SQ_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage_in, h_comp, comp_len, cudaMemcpyHostToDevice, st));
    SQ_CUDA(ctx, cudaMemcpyAsync(m.frames, h_frames, (size_t)n * sizeof(sq_frame), cudaMemcpyHostToDevice, st));
    if ((rc = sq_decode_device(ctx, ctx->d_stage_in, m.frames, n, ctx->d_stage_out, m.fres, st))) return rc;
    SQ_CUDA(ctx, cudaMemcpyAsync(h_results, m.fres, (size_t)n * sizeof(sq_frame_result), cudaMemcpyDeviceToHost, st));
    SQ_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->d_stage_out, out_len, cudaMemcpyDeviceToHost, st));
    SQ_CUDA(ctx, cudaStreamSynchronize(st));
    return SQ_OK;
}
