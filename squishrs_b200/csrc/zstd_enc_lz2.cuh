// K3 stage "compress", round-2 design: match search by a thread-block CLUSTER per chunk (lz2::search_kernel) and a parse that
// knows about repeat offsets (lz2::chase_kernel).  The entropy stage is lz::entropy_kernel, unchanged.
//
// Why a cluster.  The round-1 search kept one 2 MB hash table per resident CTA (444 of them, 0.94 GB): every row read, ring
// insert and candidate fetch was a random DRAM sector (142 B of DRAM traffic per input byte).  Here G CTAs (a cluster, one CTA
// per SM) work on ONE chunk, so only <= 18 chunks are in flight and their tables (2 MB each), ring heads and input (2 MiB
// each) stay resident in the 126 MB L2.  The CTAs of a cluster take interleaved sub-tiles of SUB positions; one hardware
// cluster barrier per step orders "all inserts of step s" before "all searches of step s".
//
// Table.  2^14 rows x 32 ring entries, keyed by a hash of 5 bytes.  entry = position | ptag << 21 | xtag << 26 where ptag = 5
// more bits of the row hash and xtag = a 6-bit hash of bytes 5..7.  One xor with the searcher's own tag word classifies an
// entry without touching the candidate's bytes:
//     long   all 11 tag bits agree  -> the candidate very likely shares >= 8 bytes           (x = e ^ T < p)
//     short  only ptag agrees       -> it shares the 5-byte prefix but not 8 bytes           ((x & 0x03FFFFFF) < p, x != that)
// Every long candidate is examined (after the continuation filter below); of the short ones only the nearest, because among
// matches of 5..7 bytes only the offset matters.  This is what lets a 32-entry row cost fewer byte comparisons than the
// round-1 16-entry row did (CPU model tests/harness/lz_model2.cc: 2.5-4 comparisons per position, ratio within 1.1 % of
// libzstd level 12 on source code, binaries and small files).
//
// Per group of 32 positions (one warp):
//   rows      four lanes share a 128-byte row (two 16-byte loads each), 8 positions per pass, 4 passes; all 8 loads of a lane are
//             in flight together
//   filter    long candidates -> (position, offset) pairs in the warp's queue; nearest short per position kept aside
//   continue  a long pair (p, o) whose left neighbour (p-1, o) is also a pair continues a match that is (or will be) found one
//             position earlier: dropped, its result arrives by inheritance.  Done in pair space through a direct-mapped table
//             of 16-bit keys; columns 0 and 16 keep everything so inheritance never runs dry.
//   verify    one pair per lane per trip: 40 candidate bytes (five aligned 8-byte loads, all requested up front) against the
//             position's bytes in shared memory, up to CAP = 32 matching bytes; best per position by atomicMax on
//             (2 len - log2 offset, nearer offset first)
//   inherit   one max-scan over the warp hands a match at position j to j + d as (offset, len - d)
//   emit      <= 3 bytes of backward extension, clamp to the block end, one 4-byte record per position (streaming store)
// The lazy decision and everything that depends on the parser's state moved to the chase kernel.
//
// chase_kernel: one warp per 128 KiB block walks the records.  At the cursor it resolves the lazy choice (depth 2, zstd's
// gains) from the 64 records it holds in registers, extends capped matches warp-wide, and -- new -- looks for repeat-offset
// matches the way a sequential encoder does: every literal position in front of the chosen match is tested against the three
// repeat offsets (>= 3 bytes is enough, they cost almost nothing to code), and at the match start a repeat offset wins when
// zstd's rule of thumb says so.  This is where the round-1 parse lost 3-7 % on binaries and records.
#pragma once
#include <cooperative_groups.h>
#include "zstd_enc_lz.cuh"
#include "zstd_enc_parse.h"

namespace lz2 {
namespace cg = cooperative_groups;

constexpr uint32_t ROW_LOG = 14, ROWS = 1u << ROW_LOG, ROW_K = 32;
constexpr uint32_t POS_BITS = 21, POS_MASK = (1u << POS_BITS) - 1, PTAG_BITS = 5, PFX_MASK = (1u << (POS_BITS + PTAG_BITS)) - 1;
constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr uint32_t LEN_BASE = zparse::LEN_BASE, CAP = zparse::CAP, MAX_SHIFT = zparse::MAX_SHIFT, REP_MIN = zparse::REP_MIN;
constexpr uint32_t SMALL_CHUNK = 128u * 1024u;  // up to here matches of 4 bytes are searched (hash of 4 bytes), above 5: libzstd's level-12 parameters make the same switch
constexpr int32_t ACCEPT_THR = zparse::ACCEPT_THR;
constexpr uint32_t LOOKAHEAD = CAP + 16;
constexpr uint32_t QUEUE_WORDS = 32 * ROW_K + 32;  // per warp: every entry of every row could be a long candidate, plus one short per position
constexpr uint32_t SEQ_PER_BLOCK = lz::SEQ_PER_BLOCK, MAX_SEQ_PER_CHUNK = lz::MAX_SEQ_PER_CHUNK, BLOCKS_PER_CHUNK = lz::BLOCKS_PER_CHUNK;
constexpr uint32_t REC_PER_CHUNK = lz::REC_PER_CHUNK;

// row index (14 bits) and ptag (5 bits) from a hash of the first five bytes (hshift = 24) or four (hshift = 32); xtag from bytes 5..7
__device__ __forceinline__ uint32_t hash_row_ptag(uint64_t v, uint32_t hshift) { return (uint32_t)(((v << hshift) * 889523592379ULL) >> (64 - (ROW_LOG + PTAG_BITS))); }
__device__ __forceinline__ uint32_t tag_word(uint32_t hv, uint64_t v) {
    const uint32_t xt = ((uint32_t)(v >> 40) * 0x9E3779B1u) >> 26;
    return (hv & ((1u << PTAG_BITS) - 1)) << POS_BITS | xt << (POS_BITS + PTAG_BITS);
}

// Per-position search record (4 bytes, HBM), 0 = no match at this position:
//   bits 0-20 offset   bits 21-25 verified length - LEN_BASE   bit 26 "may be longer" (the chase extends it)   bits 27-28 backward extension (<= 3)
__device__ __forceinline__ uint32_t pack_rec(uint32_t off, uint32_t len, uint32_t capped, uint32_t back) {
    return off | (len - LEN_BASE) << 21 | capped << 26 | back << 27;
}

template <int G> __device__ __forceinline__ void group_sync() {
    if (G > 1) cg::this_cluster().sync(); else __syncthreads();
}
// split-phase cluster barrier: arrive (release) early, wait (acquire) late
template <int G> __device__ __forceinline__ void cl_arrive() { if (G > 1) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
template <int G> __device__ __forceinline__ void cl_wait() { if (G > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// 8 bytes at position p of the chunk, straight from global memory (any alignment)
__device__ __forceinline__ uint64_t gld8(const uint8_t *__restrict__ in, uint32_t p, uint32_t n) {
    if (p + 12 <= n) return lz::ld8(in, p);
    uint64_t v = 0;
    for (uint32_t k = 0; k < 8 && p + k < n; k++) v |= (uint64_t)in[p + k] << (8 * k);
    return v;
}

// ---- search -----------------------------------------------------------------------------------------------------------
// Step pipeline of one CTA (rank r of a cluster of G); step s covers positions [(s G + r) SUB, + SUB) of the chunk:
//     reserve   ring slots (atomicAdd on the row heads) for my positions of step s+1: the round trips hide behind the search
//     stage     bytes of my sub-tile of step s+1 -> the free half of the stage buffer
//     search    step s (every CTA's entries of steps <= s are in the table)
//     publish   my entries of step s+1 into the reserved slots
//     barrier   cluster-wide (release / acquire): all entries of step s+1 are visible
// A measured alternative -- publish two steps ahead, arrive at the top of the iteration and wait at the bottom, so that a CTA
// only ever waits for one that has not even started the same step -- was 14 % faster but cost 1.2 % of ratio on real files
// (entries of later positions are never candidates, but they take ring slots early); the tolerance decides.
template <int THREADS, int SUB> struct SearchSmem {
    static constexpr uint32_t WARPS = THREADS / 32, SIN = SUB + LOOKAHEAD + 32;
    static constexpr uint32_t OFF_IN = 0, OFF_QUEUE = OFF_IN + 2 * SIN, OFF_CONT = OFF_QUEUE + WARPS * QUEUE_WORDS * 4,
                              OFF_BEST = OFF_CONT + WARPS * 1024 * 4, BYTES = OFF_BEST + SUB * 4;
    static_assert(SIN % 16 == 0, "stage buffer shape");
};

template <int G, int THREADS, int SUB, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) search_kernel(const uint8_t *__restrict__ data, const sq_span *__restrict__ spans,
                                                               const uint8_t *__restrict__ select, uint32_t n_chunks,
                                                               uint32_t *__restrict__ tab_all, uint32_t *__restrict__ head_all,
                                                               uint32_t *__restrict__ rec_all, uint32_t *__restrict__ counter, uint32_t dbg) {
    constexpr uint32_t WARPS = THREADS / 32, PER_THREAD = SUB / THREADS, GROUPS = SUB / 32;
    using L = SearchSmem<THREADS, SUB>;
    constexpr uint32_t SIN = L::SIN, SIN_WORDS = SIN / 4;
    static_assert(SUB % THREADS == 0 && SUB <= 1024, "tile shape");
    extern __shared__ __align__(16) uint8_t s_dyn[];
    uint8_t (*s_in2)[SIN] = reinterpret_cast<uint8_t (*)[SIN]>(s_dyn + L::OFF_IN);
    uint32_t *s_queue = reinterpret_cast<uint32_t *>(s_dyn + L::OFF_QUEUE);
    uint32_t *s_cont = reinterpret_cast<uint32_t *>(s_dyn + L::OFF_CONT);
    uint32_t *s_best = reinterpret_cast<uint32_t *>(s_dyn + L::OFF_BEST);
    __shared__ uint32_t s_chunk, s_gctr[2];  // group counters alternate by step parity

    const uint32_t tid = threadIdx.x, wq = tid >> 5, lane = tid & 31;
    uint32_t rank = 0;
    if (G > 1) rank = cg::this_cluster().block_rank();
    const uint32_t cid = blockIdx.x / G;
    uint32_t *tab = tab_all + (size_t)cid * ROWS * ROW_K;
    uint32_t *head = head_all + (size_t)cid * ROWS;
    uint32_t *queue = s_queue + wq * QUEUE_WORDS;
    uint32_t *T = s_cont + wq * 1024;  // continuation keys: (pair | epoch << 26), never cleared (the epoch changes with every group)
    const uint32_t sub = lane >> 2, part = lane & 3u;
    uint32_t epoch = 0;

    for (;;) {
        // ---- the cluster takes the next selected chunk ----
        if (rank == 0 && tid == 0) {
            uint32_t c;
            do { c = atomicAdd(counter, 1u); } while (c < n_chunks && select && !select[c]);
            if (G > 1) { for (uint32_t r = 0; r < (uint32_t)G; r++) *cg::this_cluster().map_shared_rank(&s_chunk, r) = c; }
            else s_chunk = c;
        }
        if (tid == 0) { s_gctr[0] = WARPS; s_gctr[1] = WARPS; }
        group_sync<G>();  // also: every CTA is done with the previous chunk's table
        const uint32_t chunk = s_chunk;
        if (chunk >= n_chunks) break;
        const uint8_t *in = data + spans[chunk].off;
        const uint32_t n = spans[chunk].len;
        uint32_t *rec = rec_all + (size_t)chunk * REC_PER_CHUNK;
        const bool aligned = (reinterpret_cast<uintptr_t>(in) & 7) == 0;
        const uint32_t MIN_MATCH = n <= SMALL_CHUNK ? 4u : 5u, hshift = n <= SMALL_CHUNK ? 32u : 24u;
        // every CTA clears its slice of the table (entries only: a ring head may start anywhere)
        {
            constexpr uint32_t SLICE4 = ROWS * ROW_K / 4 / G;
            uint4 *t4 = reinterpret_cast<uint4 *>(tab) + (size_t)rank * SLICE4;
            for (uint32_t i = tid; i < SLICE4; i += THREADS) __stcg(t4 + i, make_uint4(EMPTY, EMPTY, EMPTY, EMPTY));
        }
        const uint32_t n_steps = (n + G * SUB - 1) / (G * SUB);
        for (uint32_t i = tid; i < SIN_WORDS; i += THREADS)
            reinterpret_cast<uint32_t *>(s_in2[0])[i] = lz::stage_word(in, rank * SUB + i * 4, n, aligned);
        group_sync<G>();  // the clear is visible everywhere
        // ring-slot reservation and entry store for my positions of one step
        uint32_t r_ent[PER_THREAD], r_idx[PER_THREAD], r_slot[PER_THREAD];
        auto reserve = [&](uint32_t step) {
#pragma unroll
            for (uint32_t k = 0; k < PER_THREAD; k++) {
                const uint32_t p = (step * G + rank) * SUB + tid + k * THREADS;
                r_ent[k] = EMPTY; r_idx[k] = 0; r_slot[k] = 0;
                if (p + 8 <= n) {
                    const uint64_t v = gld8(in, p, n);
                    // inside a run of one byte (the five bytes here are the five bytes one position earlier) nothing is inserted: the run's
                    // first position stands for it, and a hot row is not flooded with interchangeable entries
                    if (p > 0 && (v & 0xFFFFFFFFFFull) == (uint64_t)in[p - 1] * 0x0101010101ull) continue;
                    const uint32_t hv = hash_row_ptag(v, hshift);
                    r_ent[k] = p | tag_word(hv, v);
                    r_idx[k] = (hv >> PTAG_BITS) * ROW_K;
                    r_slot[k] = atomicAdd(&head[hv >> PTAG_BITS], 1u);
                }
            }
        };
        auto publish = [&]() {
#pragma unroll
            for (uint32_t k = 0; k < PER_THREAD; k++)
                if (r_ent[k] != EMPTY) __stcg(&tab[r_idx[k] + (r_slot[k] & (ROW_K - 1))], r_ent[k]);
        };
        reserve(0); publish();
        group_sync<G>();  // step 0 is in the table

        for (uint32_t s = 0; s < n_steps; s++) {
            reserve(s + 1);  // ring slots for my positions of the next step: the atomics' round trips hide behind this step's search
            const uint32_t t0 = (s * G + rank) * SUB, t1 = min(n, t0 + SUB);
            const uint8_t *s_in = s_in2[s & 1];
            {   // stage my sub-tile of the next step
                const uint32_t g1 = ((s + 1) * G + rank) * SUB;
                uint32_t *dst = reinterpret_cast<uint32_t *>(s_in2[(s + 1) & 1]);
                if (g1 < n) for (uint32_t i = tid; i < SIN_WORDS; i += THREADS) dst[i] = lz::stage_word(in, g1 + i * 4, n, aligned);
            }
            // ---- search: warps take groups of 32 positions ----
            uint32_t g = wq;
#pragma unroll 1
            while (g < GROUPS && t0 + g * 32 < n) {
                const uint32_t gl = g * 32, li = gl + lane, p = t0 + li, pg = t0 + gl;
                const bool searchable = p + 8 <= n;
                const bool gfast = pg + 32 + CAP + 16 <= n;
                const uint64_t v_own = lz::smem_u64(s_in, li);
                const uint32_t hv_own = hash_row_ptag(v_own, hshift);
                const uint32_t T_own = tag_word(hv_own, v_own);
                epoch = (epoch + 1u) & 63u;
                // rows: pass k serves position 8 k + sub; this lane reads entries [4 part, +4) and [16 + 4 part, +4) of that row
                uint4 ea[4], eb[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t hk = __shfl_sync(0xffffffffu, hv_own, 8 * k + sub);
                    const uint4 *rp = reinterpret_cast<const uint4 *>(tab + (hk >> PTAG_BITS) * ROW_K) + part;
                    ea[k] = __ldcg(rp); eb[k] = __ldcg(rp + 4);
                }
                s_best[li] = 0u;
                // filter: long candidates (all tag bits agree, before the position) and the nearest short one (ptag only)
                uint32_t lmask = 0, smax[4];
                uint32_t xs[32];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t src = 8 * k + sub;
                    const uint32_t Tk = __shfl_sync(0xffffffffu, T_own, src);
                    const uint32_t pk = pg + src;
                    const uint32_t plim = pk + 8 <= n ? pk : 0u;
                    const uint32_t e[8] = {ea[k].x, ea[k].y, ea[k].z, ea[k].w, eb[k].x, eb[k].y, eb[k].z, eb[k].w};
                    uint32_t sm = 0;
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        const uint32_t x = e[m] ^ Tk, z = x & PFX_MASK;
                        xs[8 * k + m] = x;
                        if (x < plim) lmask |= 1u << (8 * k + m);
                        sm = max(sm, (z < plim && x != z) ? z : 0u);  // candidate position 0 is not worth a special case
                    }
                    sm = max(sm, __shfl_xor_sync(0xffffffffu, sm, 1));
                    sm = max(sm, __shfl_xor_sync(0xffffffffu, sm, 2));
                    smax[k] = sm;
                }
                // long pairs -> queue as (candidate position | group column << 21)
                uint32_t nlong;
                uint32_t wpos = ent::warp_excl_scan(__popc(lmask), lane, &nlong);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t colsh = (8 * k + sub) << POS_BITS;
#pragma unroll
                    for (int m = 0; m < 8; m++)
                        if (lmask >> (8 * k + m) & 1) queue[wpos++] = xs[8 * k + m] | colsh;
                }
                __syncwarp();
                // ---- continuation filter in pair space: (col, c) is dropped when (col - 1, c - 1) is a pair too ----
                const uint32_t ep = epoch << 26;
                for (uint32_t i = lane; i < nlong; i += 32) {
                    const uint32_t pr = queue[i], col = pr >> POS_BITS, o = pg + col - (pr & POS_MASK);
                    T[(o * 37u + col) & 1023u] = pr | ep;
                }
                __syncwarp();
                uint32_t total = 0;
                for (uint32_t base = 0; base < nlong; base += 32) {
                    const uint32_t i = base + lane;
                    const bool valid = i < nlong;
                    const uint32_t pr = valid ? queue[i] : 0u, col = pr >> POS_BITS, o = pg + col - (pr & POS_MASK);
                    const bool hit = T[(o * 37u + col - 1u) & 1023u] == ((pr - (1u << POS_BITS) - 1u) | ep);  // col 0 keeps everything: its key would wrap
                    const bool keep = valid && !(hit && col != 0 && !((dbg & 1u) && col == 16));
                    const uint32_t b = __ballot_sync(0xffffffffu, keep);
                    __syncwarp();
                    if (keep) queue[total + __popc(b & ((1u << lane) - 1u))] = pr;
                    total += __popc(b);
                }
                // nearest short candidate of each position (lane = column)
                {
                    uint32_t sm = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t vk = __shfl_sync(0xffffffffu, smax[k], (lane & 7u) * 4u);
                        if ((int)(lane >> 3) == k) sm = vk;
                    }
                    const bool has = sm != 0;
                    const uint32_t b = __ballot_sync(0xffffffffu, has);
                    if (has) queue[total + __popc(b & ((1u << lane) - 1u))] = sm | lane << POS_BITS;
                    total += __popc(b);
                }
                __syncwarp();
                // ---- verify ----
                if (gfast) {
#pragma unroll 1
                    for (uint32_t i = lane; i < total; i += 32) {
                        const uint32_t pr = queue[i], l0 = gl + (pr >> POS_BITS), c0 = pr & POS_MASK, o = t0 + l0 - c0;
                        const uint32_t *wp = reinterpret_cast<const uint32_t *>(s_in) + (l0 >> 2);
                        const uintptr_t ga = reinterpret_cast<uintptr_t>(in + c0);
                        const uint2 *wc = reinterpret_cast<const uint2 *>(ga & ~(uintptr_t)7);
                        const uint32_t sp = (l0 & 3u) * 8, sc = (uint32_t)(ga & 3u) * 8;
                        const bool up = (ga & 4u) != 0;
                        const uint2 A0 = __ldg(wc), A1 = __ldg(wc + 1), A2 = __ldg(wc + 2), A3 = __ldg(wc + 3), A4 = __ldg(wc + 4);
                        uint32_t pw[9];
#pragma unroll
                        for (int j = 0; j < 9; j++) pw[j] = wp[j];
                        uint32_t cw[9];
                        cw[0] = up ? A0.y : A0.x; cw[1] = up ? A1.x : A0.y; cw[2] = up ? A1.y : A1.x; cw[3] = up ? A2.x : A1.y; cw[4] = up ? A2.y : A2.x;
                        cw[5] = up ? A3.x : A2.y; cw[6] = up ? A3.y : A3.x; cw[7] = up ? A4.x : A3.y; cw[8] = up ? A4.y : A4.x;
                        uint64_t d[4];
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            d[j] = (uint64_t)(__funnelshift_r(pw[2 * j + 1], pw[2 * j + 2], sp) ^ __funnelshift_r(cw[2 * j + 1], cw[2 * j + 2], sc)) << 32 |
                                   (__funnelshift_r(pw[2 * j], pw[2 * j + 1], sp) ^ __funnelshift_r(cw[2 * j], cw[2 * j + 1], sc));
                        uint64_t x = d[0]; uint32_t mb = 0;
                        if (!x) { x = d[1]; mb = 8; if (!x) { x = d[2]; mb = 16; if (!x) { x = d[3]; mb = 24; } } }
                        const uint32_t m = x ? mb + ((uint32_t)(__ffsll((long long)x) - 1) >> 3) : CAP;
                        if (m >= MIN_MATCH)
                            atomicMax(&s_best[l0], (uint32_t)((int32_t)(2 * m) - (int32_t)zc::highbit(o + 3) + 12) << POS_BITS | (POS_MASK - o));
                    }
                } else {  // the last bytes of the chunk: careful scalar comparison
#pragma unroll 1
                    for (uint32_t i = lane; i < total; i += 32) {
                        const uint32_t pr = queue[i], l0 = gl + (pr >> POS_BITS), pp = t0 + l0, o = pp - (pr & POS_MASK);
                        const uint32_t m = lz::match_length(in, pp, pp - o, min(n - pp, CAP), n);
                        if (m >= MIN_MATCH)
                            atomicMax(&s_best[l0], (uint32_t)((int32_t)(2 * m) - (int32_t)zc::highbit(o + 3) + 12) << POS_BITS | (POS_MASK - o));
                    }
                }
                __syncwarp();
                // ---- inherit + emit ----
                uint32_t blen = 0, boff = 0;
                if (searchable) {
                    const uint32_t best = s_best[li];
                    if (best) {
                        boff = POS_MASK - (best & POS_MASK);
                        blen = ((best >> POS_BITS) - 12u + zc::highbit(boff + 3)) >> 1;
                    }
                }
                uint32_t capped = blen >= CAP ? 1u : 0u;
                {
                    const uint32_t own = blen ? (lane + blen) << 22 | capped << 21 | boff : 0u;
                    uint32_t v = own;
                    int32_t e = blen ? (int32_t)(2 * (lane + blen)) - (int32_t)zc::highbit(boff + 3) : -1000;
#pragma unroll
                    for (uint32_t d = 1; d < 32; d <<= 1) {
                        const uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
                        const int32_t eu = __shfl_up_sync(0xffffffffu, e, d);
                        if (lane >= d && eu > e) { v = u; e = eu; }
                    }
                    const uint32_t end = v >> 22;
                    if (v != own && end >= lane + MIN_MATCH && searchable) { blen = end - lane; boff = v & POS_MASK; capped = v >> 21 & 1u; }
                }
                if (p < t1) {
                    uint32_t r = 0;
                    if (blen) {
                        const uint32_t be = min(n, (p / Z_BLOCK_MAX + 1) * Z_BLOCK_MAX);
                        if (p + blen > be) { blen = be - p; capped = 0; }
                        if (blen >= MIN_MATCH) {
                            const uint32_t c = p - boff;
                            uint32_t bback = 0;
                            if (c >= 4 && p + 8 <= n) {  // one unaligned load each side: bytes [x-4, x)
                                const uint32_t diff = (uint32_t)lz::ld8(in, p - 4) ^ (uint32_t)lz::ld8(in, c - 4);
                                bback = diff == 0 ? 3u : (uint32_t)__clz((int)diff) >> 3;
                                if (bback > 3) bback = 3;
                            } else {
                                while (bback < 3 && p > bback && c > bback && in[p - bback - 1] == in[c - bback - 1]) bback++;
                            }
                            r = pack_rec(boff, blen, capped, bback);
                        }
                    }
                    __stcs(&rec[p], r);
                }
                if (lane == 0) g = atomicAdd(&s_gctr[s & 1], 1u);
                g = __shfl_sync(0xffffffffu, g, 0);
            }
            // ---- publish the next step's entries into the slots reserved above; every CTA's entries of step s+1 are visible after the barrier ----
#pragma unroll
            for (uint32_t k = 0; k < PER_THREAD; k++) asm volatile("" : "+r"(r_slot[k]) :: "memory");
            publish();
            cl_arrive<G>();
            cl_wait<G>();
            if (tid == 0) s_gctr[(s + 1) & 1] = WARPS;  // nobody touches the other parity's counter during this iteration
            __syncthreads();
        }
    }
}

// ---- chase: lazy decision + repeat offsets + sequence emission, one warp per block ------------------------------------
__device__ __forceinline__ uint32_t load4(const uint8_t *__restrict__ in, uint32_t pos, uint32_t n) {
    if (pos + 8 <= n) {  // n = ~0u: the caller knows every read of this block stays 8 bytes inside the chunk
        const uintptr_t a = reinterpret_cast<uintptr_t>(in + pos);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        return __funnelshift_r(__ldg(w), __ldg(w + 1), (uint32_t)(a & 3u) * 8);
    }
    uint32_t v = 0;
    for (uint32_t k = 0; k < 4 && pos + k < n; k++) v |= (uint32_t)in[pos + k] << (8 * k);
    return v;
}

__global__ void __launch_bounds__(128) chase_kernel(const uint8_t *__restrict__ data, const sq_span *__restrict__ spans,
                                                     const uint8_t *__restrict__ select, uint32_t n_chunks,
                                                     const uint32_t *__restrict__ rec_all, zc::Seq *__restrict__ seqs_all,
                                                     lz::BlockMeta *__restrict__ meta_all) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (item >= n_chunks * BLOCKS_PER_CHUNK) return;
    const uint32_t chunk = item / BLOCKS_PER_CHUNK, b = item % BLOCKS_PER_CHUNK;
    if (select && !select[chunk]) return;
    const uint32_t n = spans[chunk].len;
    const uint32_t bs = b * Z_BLOCK_MAX;
    if (bs >= n) return;
    const uint32_t be = min(n, bs + Z_BLOCK_MAX);
    const uint8_t *in = data + spans[chunk].off;
    const uint32_t *rec = rec_all + (size_t)chunk * REC_PER_CHUNK;
    zc::Seq *seqs = seqs_all + (size_t)chunk * MAX_SEQ_PER_CHUNK + (size_t)b * SEQ_PER_BLOCK;
    const uint32_t nl = be + 8 <= n ? ~0u : n;  // bound handed to load4: no per-load check except in the chunk's last block
    uint32_t p = bs, anchor = bs, nseq = 0;
    uint32_t r0 = 0, r1 = 0, r2 = 0;  // repeat offsets are unknown at a block start (a raw block must not desynchronise the decoder); the frame's first block knows 1,4,8
    if (b == 0) { r0 = 1; r1 = 4; r2 = 8; }

    auto emit = [&](uint32_t start, uint32_t len, uint32_t off) {
        const uint32_t ll = start - anchor;
        uint32_t ob = off + 3;  // repeat-offset code substitution (RFC 8878 3.1.1.5)
        if (ll) { if (off == r0) ob = 1; else if (off == r1) ob = 2; else if (off == r2) ob = 3; }
        else { if (off == r1) ob = 1; else if (off == r2) ob = 2; else if (r0 > 1 && off == r0 - 1) ob = 3; }
        if (ob > 3) { r2 = r1; r1 = r0; r0 = off; }
        else {
            const uint32_t ix = ob - 1 + (ll ? 0 : 1);
            if (ix == 1) { const uint32_t t = r1; r1 = r0; r0 = t; }
            else if (ix == 2) { const uint32_t t = r2; r2 = r1; r1 = r0; r0 = t; }
            else if (ix == 3) { const uint32_t t = r0 - 1; r2 = r1; r1 = r0; r0 = t; }
        }
        if (lane == 0) { zc::Seq sq; sq.ll = ll; sq.ml = len; sq.off_base = ob; seqs[nseq] = sq; }
        nseq++;
        p = start + len;
        anchor = p;
    };
    // One pass over the literal positions [lo, hi) in front of a match -- and, with `with_start`, over the match start hi itself --
    // for repeat-offset matches (>= 3 bytes): lane j tests position lo + j against the three offsets that are cheap THERE (they
    // differ when the position follows a match directly).  The first position that matches anything wins; its longest candidate is
    // extended by the whole warp only if more than 3 bytes agree.  In the gap a hit is taken at once; at the match start it has to
    // beat the match by zstd's rule of thumb (3 rl > 3 len - log2(offset) + 1).  Returns true if a sequence was emitted.
    auto scan = [&](uint32_t lo, uint32_t hi, bool with_start, uint32_t mlen, uint32_t moff) -> bool {
        if (!(r0 | r1 | r2)) return false;
        const uint32_t end = with_start ? hi + 1 : hi;
        for (uint32_t gb = lo; gb < end; gb += 32) {
            const uint32_t g = gb + lane;
            const bool ll0 = g == anchor;
            const uint32_t c0 = ll0 ? r1 : r0, c1 = ll0 ? r2 : r1, c2 = ll0 ? (r0 > 1 ? r0 - 1 : 0u) : r2;
            uint32_t fl = 0;  // per candidate 2 bits: 0 = no match, 1 = exactly 3 bytes, 2 = 4 or more
            if (g < end && g + 4 <= be) {
                const uint32_t v = load4(in, g, nl);
                if (c0 && c0 <= g) { const uint32_t x = load4(in, g - c0, nl) ^ v; if (!(x & 0xFFFFFFu)) fl |= x ? 1u : 2u; }
                if (c1 && c1 <= g) { const uint32_t x = load4(in, g - c1, nl) ^ v; if (!(x & 0xFFFFFFu)) fl |= (x ? 1u : 2u) << 2; }
                if (c2 && c2 <= g) { const uint32_t x = load4(in, g - c2, nl) ^ v; if (!(x & 0xFFFFFFu)) fl |= (x ? 1u : 2u) << 4; }
            }
            const uint32_t any = __ballot_sync(0xffffffffu, fl != 0);
            if (!any) continue;
            const uint32_t j = (uint32_t)__ffs((int)any) - 1, gs = gb + j;
            const uint32_t F = __shfl_sync(0xffffffffu, fl, j);
            const uint32_t C0 = __shfl_sync(0xffffffffu, c0, j), C1 = __shfl_sync(0xffffffffu, c1, j), C2 = __shfl_sync(0xffffffffu, c2, j);
            uint32_t rl = 0, ro = 0;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const uint32_t f = F >> (2 * i) & 3u, o = i == 0 ? C0 : i == 1 ? C1 : C2;
                if (!f) continue;
                uint32_t l = 3;
                if (f == 2) l = 4 + (gs + 4 < be ? lz::warp_extend(in, n, gs + 4, gs + 4 - o, be - gs - 4, lane) : 0u);
                if (l > rl) { rl = l; ro = o; }
            }
            if (gs < hi) { emit(gs, rl, ro); return true; }
            // the match start itself
            if (ro != moff && (int32_t)(3 * rl) > (int32_t)(3 * mlen) - (int32_t)zc::highbit(moff + 3) + 1) { emit(gs, rl, ro); return true; }
            return false;
        }
        return false;
    };
    // lazy score of a record (what the parser gains by taking it), -1 = unusable
    auto score_of = [&](uint32_t r) -> int32_t {
        if (!r) return -1;
        const uint32_t sl = (r >> 26 & 1u) ? CAP : ((r >> 21) & 31u) + LEN_BASE;
        const int32_t sc = (int32_t)(4 * sl) - (int32_t)zc::highbit((r & POS_MASK) + 3);
        return sc >= ACCEPT_THR ? sc : -1;
    };

    uint32_t base = ~0u - 63u, w0 = 0, w1 = 0, umask = 0;
    while (p < be && nseq < SEQ_PER_BLOCK) {
        const uint32_t nb = p & ~31u;
        if (nb != base) {
            if (nb == base + 32) w0 = w1;
            else w0 = (nb + lane < be) ? __ldg(rec + nb + lane) : 0u;
            w1 = (nb + 32 + lane < be) ? __ldg(rec + nb + 32 + lane) : 0u;  // requested now, first touched when the cursor gets near it
            base = nb;
            umask = __ballot_sync(0xffffffffu, score_of(w0) >= 0);  // usable records of this window, once per window
        }
        const uint32_t usable = umask & (0xffffffffu << (p - base));
        if (!usable) {  // literals up to the end of this window (unless a repeat offset matches)
            const uint32_t ge = min(be, base + 32);
            if (!scan(p, ge, false, 0, 0)) p = ge;
            continue;
        }
        const uint32_t q = base + (uint32_t)__ffs((int)usable) - 1;
        auto rec_at = [&](uint32_t x) -> uint32_t {  // warp-uniform x in [base, base + 64)
            const uint32_t d = x - base;
            if (x >= be) return 0u;
            return d < 32 ? __shfl_sync(0xffffffffu, w0, d) : __shfl_sync(0xffffffffu, w1, d - 32);
        };
        uint32_t start = q;
        uint32_t r = rec_at(q);
        int32_t cur = score_of(r);
        while (!(r >> 26 & 1u) && start - q + 2 <= MAX_SHIFT) {
            const uint32_t ra = rec_at(start + 1);
            const int32_t s1 = score_of(ra);
            if (s1 > cur + 4) { cur = s1; start += 1; r = ra; continue; }
            const uint32_t rb = rec_at(start + 2);
            const int32_t s2 = score_of(rb);
            if (s2 > cur + 7) { cur = s2; start += 2; r = rb; continue; }
            break;
        }
        const uint32_t off = r & POS_MASK;
        uint32_t len = ((r >> 21) & 31u) + LEN_BASE, back = (r >> 27) & 3u;
        if ((r >> 26 & 1u) && start + len < be) len += lz::warp_extend(in, n, start + len, start + len - off, be - start - len, lane);
        if (back > start - anchor) back = start - anchor;
        start -= back; len += back;
        if (scan(min(p, start), start, true, len, off)) continue;
        emit(start, len, off);
    }
    if (lane == 0) {
        lz::BlockMeta m;
        m.seq_start = b * SEQ_PER_BLOCK; m.nseq = nseq; m.last_lits = be - anchor; m.reserved = 0;
        meta_all[(size_t)chunk * BLOCKS_PER_CHUNK + b] = m;
    }
}

// ---- the same parse with ONE THREAD per block (zstd_enc_parse.h, the function the CPU model runs): a debugging reference for
// the warp kernel above (SQ_LZ2_DBG & 4), far too slow to ship -- a lone thread's dependent chain runs at a few hundred cycles per step
__global__ void __launch_bounds__(64) chase_thread_kernel(const uint8_t *__restrict__ data, const sq_span *__restrict__ spans,
                                                    const uint8_t *__restrict__ select, uint32_t n_chunks,
                                                    const uint32_t *__restrict__ rec_all, zc::Seq *__restrict__ seqs_all,
                                                    lz::BlockMeta *__restrict__ meta_all) {
    const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n_chunks * BLOCKS_PER_CHUNK) return;
    const uint32_t chunk = item / BLOCKS_PER_CHUNK, b = item % BLOCKS_PER_CHUNK;
    if (select && !select[chunk]) return;
    const uint32_t n = spans[chunk].len;
    const uint32_t bs = b * Z_BLOCK_MAX;
    if (bs >= n) return;
    const uint32_t be = min(n, bs + Z_BLOCK_MAX);
    uint32_t last_lits;
    const uint32_t nseq = zparse::chase_block(data + spans[chunk].off, n, rec_all + (size_t)chunk * REC_PER_CHUNK, bs, be, b == 0,
                                              seqs_all + (size_t)chunk * MAX_SEQ_PER_CHUNK + (size_t)b * SEQ_PER_BLOCK, SEQ_PER_BLOCK, &last_lits);
    lz::BlockMeta m;
    m.seq_start = b * SEQ_PER_BLOCK; m.nseq = nseq; m.last_lits = last_lits; m.reserved = 0;
    meta_all[(size_t)chunk * BLOCKS_PER_CHUNK + b] = m;
}

}  // namespace lz2
