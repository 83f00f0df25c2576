// K3 stage "compress": LZ match search (lz_search_kernel), parse (lz_chase_kernel) and entropy coding
// (entropy_kernel).
//
// lz_search_kernel -- one persistent CTA per chunk in flight, 256 threads, the chunk walked in tiles of
// 1024 positions.  Per tile:
//   stage   the tile's bytes (+ lookahead) into shared memory
//   insert  every position into the chunk's bucketed hash table in HBM: 2^15 rows x 16 entries keyed by a
//           hash of 6 bytes, entry = (position+1) | 10-bit tag << 22, ring slot taken with atomicAdd on the
//           row head.  The table is never cleared between chunks: a stale entry is just a candidate
//           position, and every candidate is verified against the bytes of the CURRENT chunk.
//   search  every position reads its 64-byte row (L1 bypassed), filters by tag, requests the first 8 bytes
//           of all sixteen candidates at once, extends the survivors 8 bytes per round in lock-step (up to
//           24 bytes; longer matches are extended by the chase) and keeps the best by 2*len - log2(offset); plus how far it extends backwards (<= 3).
//           Results go to a one-tile window (+ 32-position halo of the previous tile) in shared memory.
//   decide  every position resolves the lazy (depth 2) choice "if the parser stands here, which match
//           start does it take" from the ring alone -- no dependence on parser state, so it is parallel --
//           and writes a 4-byte record per position to HBM.
// lz_chase_kernel -- one warp per 128 KiB block follows the records from the block start (windows of 32
// records per coalesced load, hops by shuffle, literal runs skipped by ballot), extends capped matches
// warp-wide, applies backward extension, substitutes repeat-offset codes and appends the block's sequences.
// Matches never cross a block boundary; repeat-offset knowledge is dropped at every block start so a block
// that later falls back to raw cannot desynchronise the decoder's history.
//
// entropy_kernel -- one warp per block: gathers literals, then writes the block body warp-parallel
// (zstd_enc_entropy.cuh: histograms, segment-parallel Huffman / FSE bit packing; tables by one lane).
#pragma once
#include "common.cuh"
#include "zstd_enc_block.h"
#include "zstd_enc_entropy.cuh"

namespace lz {

constexpr uint32_t ROW_LOG = 15, ROWS = 1u << ROW_LOG, ROW_K = 16, TAG_BITS = 10;
constexpr uint32_t TILE = 1024, HALO = 32, RING = TILE + HALO, THREADS = 256, PER_THREAD = TILE / THREADS;
constexpr uint32_t MIN_MATCH = 6, SEARCH_CAP = 24, TARGET_LEN = 24, DEFER = 20, MAX_LAZY_ITERS = 8;
constexpr uint32_t LOOKAHEAD = SEARCH_CAP + 16;  // bytes staged past the tile so the p-side of every comparison is in smem
#ifndef SQ_LZ_ACCEPT
#define SQ_LZ_ACCEPT 8
#endif
constexpr int32_t ACCEPT_THR = SQ_LZ_ACCEPT;
constexpr uint32_t BLOCKS_PER_CHUNK = 16;
constexpr uint32_t BODY_STRIDE = 2 * Z_BLOCK_MAX;  // per-block body slot: literals + <= 8 bytes per sequence always fit
constexpr uint32_t SEQ_PER_BLOCK = Z_BLOCK_MAX / MIN_MATCH + 8, MAX_SEQ_PER_CHUNK = BLOCKS_PER_CHUNK * SEQ_PER_BLOCK;
constexpr uint32_t REC_PER_CHUNK = 2048u * 1024u, MAX_SHIFT = 7;
// per entropy warp: 3 x SEQ_PER_BLOCK words of FSE state-transition records, SEQ_PER_BLOCK words of packed symbol codes, then three
// byte arrays of symbol codes (stride SEQ_CODE_STRIDE)
constexpr uint32_t SEQ_CODE_STRIDE = (SEQ_PER_BLOCK + 15u) & ~15u, SBITS_STRIDE = (4 * SEQ_PER_BLOCK + 3 * SEQ_CODE_STRIDE / 4 + 3u) & ~3u;

#ifdef SQ_LZ_TIMERS
__device__ unsigned long long g_lz_timers[12];
#endif

struct BlockMeta {  // one per (chunk, block), written by lz_kernel, read by entropy_kernel
    uint32_t seq_start, nseq, last_lits, reserved;
};

__device__ __forceinline__ uint64_t ld64_unaligned(const uint8_t *base, uint32_t pos) {
    // reads the aligned 16 bytes around base+pos: callers guarantee pos + 16 <= chunk length
    const uintptr_t a = reinterpret_cast<uintptr_t>(base + pos);
    const uint64_t *w = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t)7);
    const uint32_t sh = (uint32_t)(a & 7u) * 8;
    const uint64_t lo = __ldg(w), hi = __ldg(w + 1);  // branch-free: both loads always issue (memory-level parallelism)
    return (lo >> sh) | ((hi << 1) << (63 - sh));
}

// unaligned 8 bytes as three aligned 32-bit loads + two funnel shifts (caller guarantees pos + 12 <= chunk length)
__device__ __forceinline__ uint64_t ld8(const uint8_t *base, uint32_t pos) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(base + pos);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3u) * 8;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    return (uint64_t)__funnelshift_r(w1, w2, sh) << 32 | __funnelshift_r(w0, w1, sh);
}

// length of the common prefix of in[p..] and in[c..] (c < p), at most maxlen
__device__ __forceinline__ uint32_t match_length(const uint8_t *__restrict__ in, uint32_t p, uint32_t c, uint32_t maxlen, uint32_t n_safe) {
    uint32_t l = 0;
    // word loop only while both unaligned 8-byte reads stay inside the 8-byte padded chunk
    while (l + 8 <= maxlen && p + l + 16 <= n_safe) {
        uint64_t x = ld64_unaligned(in, p + l) ^ ld64_unaligned(in, c + l);
        if (x) return l + (uint32_t)(__ffsll((long long)x) - 1) / 8;
        l += 8;
    }
    while (l < maxlen && in[p + l] == in[c + l]) l++;
    return l;
}

// unaligned 8 bytes at byte index li of a 4-byte aligned shared array
__device__ __forceinline__ uint64_t smem_u64(const uint8_t *s, uint32_t li) {
    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(s) + (li >> 2);
    const uint32_t sh = (li & 3) * 8;
    const uint32_t w0 = s32[0], w1 = s32[1], w2 = s32[2];
    return (uint64_t)__funnelshift_r(w1, w2, sh) << 32 | __funnelshift_r(w0, w1, sh);
}
// hash of the first MIN_MATCH (6) bytes
__device__ __forceinline__ uint32_t hash5(uint64_t v) { return (uint32_t)(((v << 16) * 227718039650203ULL) >> (64 - (ROW_LOG + TAG_BITS))); }

__device__ __forceinline__ int32_t sel_score(uint32_t len, uint32_t off) { return (int32_t)(2 * len) - (int32_t)zc::highbit(off + 3); }
__device__ __forceinline__ int32_t lazy_score(uint32_t len, uint32_t off) { return (int32_t)(4 * len) - (int32_t)zc::highbit(off + 3); }

// Per-position parse record (4 bytes, HBM): what the parser does if its cursor stands on this position.
//   0                      literal
//   bits  0-20 offset      bits 21-25 verified length-1 (<= SEARCH_CAP)      bit 26 "may be longer": the chase extends it
//   bits 27-28 backward extension available (<= 3)    bits 29-31 shift from the position to the match start (<= 7)
__device__ __forceinline__ uint32_t pack_rec(uint32_t off, uint32_t len, uint32_t extend, uint32_t back, uint32_t shift) {
    return off | (len - 1) << 21 | extend << 26 | back << 27 | shift << 29;
}

// ---- kernel A: search + decide, fully parallel, one persistent CTA per chunk in flight ----------------
// Tile pipeline (two barriers per tile):
//   reserve   ring slots for the NEXT tile's positions are taken now (atomicAdd on the row heads), so the round trip
//             of the atomics hides behind this tile's search; the entries themselves are stored after the search,
//             so the search sees exactly the table "everything up to and including this tile"
//   prefetch  the bytes of the tile after next travel through registers into the free half of the double-buffered stage
//   search    warps take groups of 32 positions from a shared counter (balances the warps of a CTA)
//   publish   store the next tile's entries                                                   | barrier
//   decide    lazy choice per position; warp 0 also slides the result window's halo             | barrier
constexpr uint32_t SIN = TILE + LOOKAHEAD + 24, SIN_WORDS = SIN / 4;  // staged bytes per tile (multiple of 16)
static_assert(SIN % 16 == 0 && SIN_WORDS <= 2 * THREADS && SIN >= TILE + SEARCH_CAP + 8, "stage buffer shape");

__device__ __forceinline__ uint32_t stage_word(const uint8_t *__restrict__ in, uint32_t g, uint32_t n, bool aligned) {
    if (g + 4 <= n && aligned) return __ldg(reinterpret_cast<const uint32_t *>(in + g));
    uint32_t w = 0;
    for (uint32_t k = 0; k < 4; k++) if (g + k < n) w |= (uint32_t)in[g + k] << (8 * k);
    return w;
}

#ifndef SQ_LZ_SSTRIDE
#define SQ_LZ_SSTRIDE 2  // default search stride (the kernel takes the mode as an argument): 1 = every position reads its row; 2 = every
                         // second one (1.3x faster; +0.3 % bytes on the mixed log/JSON/binary corpus, +3 % on real text and code);
                         // 0 = chosen per tile from the previous tile's share of positions inside long matches (experimental)
#endif
#ifndef SQ_LZ_MINB
#define SQ_LZ_MINB 3  // resident CTAs per SM the register budget is held to
#endif
__global__ void __launch_bounds__(THREADS, SQ_LZ_MINB) lz_search_kernel(const uint8_t *__restrict__ data, const sq_span *__restrict__ spans,
                                                             const uint8_t *__restrict__ select, uint32_t n_chunks,
                                                             uint32_t *__restrict__ tab_all, uint32_t *__restrict__ head_all,
                                                             uint32_t *__restrict__ rec_all, uint32_t *__restrict__ counter, uint32_t stride_mode, uint32_t adapt_thr) {
    __shared__ __align__(16) uint8_t s_in2[2][SIN];
    __shared__ uint8_t s_len[RING];
    __shared__ uint32_t s_off[RING];
    __shared__ uint8_t s_back[RING];
    __shared__ int16_t s_sc[RING];  // lazy score of the position's usable match, -1 = none
    __shared__ __align__(16) uint32_t s_queue[(THREADS / 32) * 32 * ROW_K];  // per warp: compacted (position | candidate << 10) pairs
    __shared__ uint32_t s_best[TILE];                           // per position: (score+9) << 26 | (len-6) << 21 | off  (0 = none)
    __shared__ uint32_t s_chunk, s_gctr, s_stride, s_long;
#ifdef SQ_LZ_TIMERS
    long long tm[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tc = clock64();
#define LZ_TICK(i) do { long long now_ = clock64(); tm[i] += now_ - tc; tc = now_; } while (0)
#else
#define LZ_TICK(i)
#endif
    uint32_t *tab = tab_all + (size_t)blockIdx.x * ROWS * ROW_K;
    uint32_t *head = head_all + (size_t)blockIdx.x * ROWS;
    const uint32_t tid = threadIdx.x, wq = tid >> 5, lane = tid & 31;
    uint32_t *queue = s_queue + wq * (32 * ROW_K);
    // Search stride 2: only every second position reads its row; the others live on inherited matches (and on the backward
    // extension of their right neighbour's finds).  A group of 32 lanes then spans 64 positions.
    // The stride is chosen per tile: where most positions of the previous tile sat inside long matches (repetitive data: logs,
    // records, padding) the in-between positions have little to add and stride 2 costs almost nothing; elsewhere (text, code)
    // every position is searched.  adapt_thr = number of such positions (of TILE) from which the next tile runs at stride 2.
    constexpr uint32_t TAG_MASK = (1u << TAG_BITS) - 1, STAT_LEN = 16;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            uint32_t c;
            do { c = atomicAdd(counter, 1u); } while (c < n_chunks && select && !select[c]);
            s_chunk = c;
            s_gctr = THREADS / 32;
            s_stride = stride_mode ? stride_mode : 1u;
            s_long = 0;
        }
        __syncthreads();
        const uint32_t chunk = s_chunk;
        if (chunk >= n_chunks) break;
        LZ_TICK(5);
        const uint8_t *in = data + spans[chunk].off;
        const uint32_t n = spans[chunk].len;
        uint32_t *rec = rec_all + (size_t)chunk * REC_PER_CHUNK;
        const bool aligned = (reinterpret_cast<uintptr_t>(in) & 7) == 0;
        for (uint32_t i = tid; i < RING; i += THREADS) { s_len[i] = 0; s_sc[i] = -1; }
        // prologue: stage tiles 0 and 1, insert tile 0
        for (uint32_t i = tid; i < 2 * SIN_WORDS; i += THREADS) {
            const uint32_t b = i >= SIN_WORDS ? 1u : 0u, w = i - b * SIN_WORDS;
            reinterpret_cast<uint32_t *>(s_in2[b])[w] = stage_word(in, b * TILE + w * 4, n, aligned);
        }
        __syncthreads();
#pragma unroll
        for (uint32_t k = 0; k < PER_THREAD; k++) {
            const uint32_t li = tid + k * THREADS;
            if (li + 8 <= n) {
                const uint32_t hv = hash5(smem_u64(s_in2[0], li)), row = hv >> TAG_BITS;
                const uint32_t slot = atomicAdd(&head[row], 1u) & (ROW_K - 1);
                __stcg(&tab[row * ROW_K + slot], (li + 1) | (hv & TAG_MASK) << 22);
            }
        }
        __syncthreads();

        const uint32_t n_tiles = (n + TILE - 1) / TILE;
        for (uint32_t t = 0; t < n_tiles; t++) {
            const uint32_t t0 = t * TILE, t1 = min(n, t0 + TILE);
            const uint32_t be = min(n, (t0 / Z_BLOCK_MAX + 1) * Z_BLOCK_MAX);  // end of the block this tile lies in
            const bool last_tile_of_block = (t1 == be);
            const uint32_t SSTRIDE = s_stride, GSPAN = 32 * SSTRIDE, GROUPS = TILE / GSPAN;
            const uint8_t *s_in = s_in2[t & 1];
            const uint8_t *s_nx = s_in2[(t + 1) & 1];
            // ---- reserve: ring slots for the next tile's positions (results are consumed after the search) ----
            uint32_t slot_raw[PER_THREAD];
#pragma unroll
            for (uint32_t k = 0; k < PER_THREAD; k++) {
                const uint32_t li = tid + k * THREADS, p = t0 + TILE + li;
                slot_raw[k] = 0;
                if (p + 8 <= n) slot_raw[k] = atomicAdd(&head[hash5(smem_u64(s_nx, li)) >> TAG_BITS], 1u);
            }
            // ---- prefetch: bytes of tile t+2 (stored into this tile's stage buffer once the search is done) ----
            uint32_t sw0 = 0, sw1 = 0;
            {
                const uint32_t g2 = t0 + 2 * TILE;
                if (g2 < n) {
                    sw0 = stage_word(in, g2 + tid * 4, n, aligned);
                    if (tid < SIN_WORDS - THREADS) sw1 = stage_word(in, g2 + (THREADS + tid) * 4, n, aligned);
                }
            }
            LZ_TICK(0);
            // ---- search ----
            // Each warp takes 32 consecutive positions at a time.  Lanes first filter their own row by tag (no data access),
            // the warp compacts the surviving (position, candidate) pairs into a dense queue, and the lanes then verify
            // pairs -- not row slots -- so no issue slot is spent on empty slots.  The best candidate per position is kept
            // with a 32-bit atomicMax on (score, length, offset) in shared memory.
            // rows are software-pipelined: the row of the warp's NEXT group is requested before this group's pairs are verified
            // Rows are read cooperatively: four lanes share one 64-byte row (one 16-byte load each), so a load instruction
            // touches 8 rows instead of 32 and the L1 sees a quarter of the tag requests.  Lane l therefore filters, for
            // k = 0..3, entries [4 (l & 3), +4) of the row of group position 8 k + (l >> 2); everything downstream works on
            // (position, offset) pairs and does not care which lane found them.
            uint32_t g = wq, gn = 0;
            if (lane == 0) gn = atomicAdd(&s_gctr, 1u);
            gn = __shfl_sync(0xffffffffu, gn, 0);
            const uint32_t sub = lane >> 2, part = lane & 3u;
            const uint32_t droppable = ((sub & 15u) ? 0x0F0Fu : 0u) | 0xF0F0u;  // positions 0 and 16 of a group keep every pair
            uint32_t hv_cur = hash5(smem_u64(s_in, g * GSPAN + SSTRIDE * lane));
            uint4 ne[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t hk = __shfl_sync(0xffffffffu, hv_cur, 8 * k + sub);
                ne[k] = __ldcg(reinterpret_cast<const uint4 *>(tab + (hk >> TAG_BITS) * ROW_K) + part);
            }
#pragma unroll 1
            while (g < GROUPS) {
                const uint32_t gl = g * GSPAN, li = gl + SSTRIDE * lane, p = t0 + li;
                uint32_t mask = 0, off[ROW_K];  // off[4 k + m] = position - candidate, 0 = not a candidate
                const bool searchable = p + 8 <= n;
                const bool gfast = t0 + gl + GSPAN + SEARCH_CAP + 16 <= n;  // every comparison of the group stays inside the chunk and the staged window
                {
                    static_assert(ROW_K == 16, "the search reads one 16-entry row");
                    const uint32_t e[16] = {ne[0].x, ne[0].y, ne[0].z, ne[0].w, ne[1].x, ne[1].y, ne[1].z, ne[1].w,
                                            ne[2].x, ne[2].y, ne[2].z, ne[2].w, ne[3].x, ne[3].y, ne[3].z, ne[3].w};
                    uint32_t tg[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) tg[k] = (__shfl_sync(0xffffffffu, hv_cur, 8 * k + sub) & TAG_MASK) << 22;
                    // entry = (candidate + 1) | tag << 22: after xor with the tag the value is candidate + 1 iff the tags agree
                    // (anything else is 0 or >= 2^22), so one unsigned compare checks tag, emptiness and candidate < p at once
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t pk = t0 + gl + SSTRIDE * (8 * k + sub);
                        const uint32_t plim = pk + 8 <= n ? pk : 0u;
#pragma unroll
                        for (int m = 0; m < 4; m++) {
                            const uint32_t om1 = pk - (e[4 * k + m] ^ tg[k]);
                            const bool v = om1 < plim;
                            off[4 * k + m] = v ? om1 + 1u : 0u;
                            mask |= (v ? 1u : 0u) << (4 * k + m);
                        }
                    }
                }
                // prefetch the next group's rows only now that this group's entries are consumed: the loads land straight in
                // the registers they are read from one group later (table rows exist for any hash value)
                if (gn < GROUPS) {
                    hv_cur = hash5(smem_u64(s_in, gn * GSPAN + SSTRIDE * lane));
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t hk = __shfl_sync(0xffffffffu, hv_cur, 8 * k + sub);
                        ne[k] = __ldcg(reinterpret_cast<const uint4 *>(tab + (hk >> TAG_BITS) * ROW_K) + part);
                    }
                }
                s_best[li] = 0u;
                LZ_TICK(6);
                // ---- continuation filter ----
                // A pair (p, c) whose left neighbour pair (p-1, c-1) is also a candidate pair continues a match that is verified at
                // its first position; it is dropped here and its result arrives by inheritance below.  Membership is tested through a
                // direct-mapped table of (position, offset) keys that borrows the queue's memory; a key collision only loses a drop.
                // Positions 0 and 16 keep every pair, so a long match is re-verified every 16 positions and inheritance never runs dry.
                // Both passes are branch-free (predicated stores, then sixteen independent loads) so the loads overlap; the column
                // is rotated by bits 1..4 of the offset so that the four lanes filtering one row spread over the banks.
                {
                    uint16_t *T = reinterpret_cast<uint16_t *>(queue);
                    uint4 *Tz = reinterpret_cast<uint4 *>(queue);
#pragma unroll
                    for (int z = 0; z < 4; z++) Tz[lane + 32 * z] = make_uint4(0u, 0u, 0u, 0u);
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 16; q++)
                        if (off[q]) T[(off[q] & 31u) * 32u + ((8 * (q >> 2) + sub + (off[q] & 30u)) & 31u)] = (uint16_t)((off[q] >> 5) + 1u);
                    __syncwarp();
                    uint32_t hit = 0;
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        const uint32_t o = off[q], col = 8 * (q >> 2) + sub;
                        hit |= (T[(o & 31u) * 32u + (((col ? col - 1u : 0u) + (o & 30u)) & 31u)] == (uint16_t)((o >> 5) + 1u) ? 1u : 0u) << q;
                    }
                    mask &= ~(hit & droppable);
                    __syncwarp();
                }
                LZ_TICK(7);
                uint32_t total;
                uint32_t wpos = ent::warp_excl_scan(__popc(mask), lane, &total);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t lk = gl + SSTRIDE * (8 * k + sub), pbase = lk | (t0 + lk) << 10;
#pragma unroll
                    for (int m = 0; m < 4; m++)
                        if (mask >> (4 * k + m) & 1) queue[wpos++] = pbase - (off[4 * k + m] << 10);
                }
                __syncwarp();
                LZ_TICK(8);
                if (gfast) {
#pragma unroll 1
                    for (uint32_t i = lane; i < total; i += 64) {
                        // two pairs per lane per trip: both candidates' bytes (four aligned 8-byte loads each, 24 bytes at any
                        // alignment) are requested before either is examined
                        const bool h1 = i + 32 < total;
                        const uint32_t pr0 = queue[i], pr1 = queue[h1 ? i + 32 : i];
                        const uint32_t l0 = pr0 & 1023u, c0 = pr0 >> 10, l1 = pr1 & 1023u, c1 = pr1 >> 10;
                        const uint32_t *wp0 = reinterpret_cast<const uint32_t *>(s_in) + (l0 >> 2), *wp1 = reinterpret_cast<const uint32_t *>(s_in) + (l1 >> 2);
                        const uintptr_t ga0 = reinterpret_cast<uintptr_t>(in + c0), ga1 = reinterpret_cast<uintptr_t>(in + c1);
                        const uint2 *wc0 = reinterpret_cast<const uint2 *>(ga0 & ~(uintptr_t)7), *wc1 = reinterpret_cast<const uint2 *>(ga1 & ~(uintptr_t)7);
                        const uint32_t sp0 = (l0 & 3u) * 8, sp1 = (l1 & 3u) * 8, sc0 = (uint32_t)(ga0 & 3u) * 8, sc1 = (uint32_t)(ga1 & 3u) * 8;
                        const bool u0 = (ga0 & 4u) != 0, u1 = (ga1 & 4u) != 0;  // the candidate starts in the upper word of its first 8 bytes
                        const uint2 A0 = __ldg(wc0), A1 = __ldg(wc0 + 1), B0 = __ldg(wc1), B1 = __ldg(wc1 + 1);
                        const uint32_t p0w = wp0[0], p1w = wp0[1], p2w = wp0[2], q0w = wp1[0], q1w = wp1[1], q2w = wp1[2];
                        const uint32_t a0w = u0 ? A0.y : A0.x, a1w = u0 ? A1.x : A0.y, a2w = u0 ? A1.y : A1.x;
                        const uint32_t b0w = u1 ? B0.y : B0.x, b1w = u1 ? B1.x : B0.y, b2w = u1 ? B1.y : B1.x;
                        const uint32_t x0lo = __funnelshift_r(p0w, p1w, sp0) ^ __funnelshift_r(a0w, a1w, sc0);
                        const uint32_t x0hi = __funnelshift_r(p1w, p2w, sp0) ^ __funnelshift_r(a1w, a2w, sc0);
                        const uint32_t x1lo = __funnelshift_r(q0w, q1w, sp1) ^ __funnelshift_r(b0w, b1w, sc1);
                        const uint32_t x1hi = __funnelshift_r(q1w, q2w, sp1) ^ __funnelshift_r(b1w, b2w, sc1);
                        // bytes 0..3 must agree (minimum match is 6); the first difference inside bytes 4..7 ends the match there
                        uint32_t m0 = x0lo ? 0u : x0hi ? 4u + (uint32_t)(__ffs((int)x0hi) - 1) / 8 : 8u;
                        uint32_t m1 = x1lo ? 0u : x1hi ? 4u + (uint32_t)(__ffs((int)x1hi) - 1) / 8 : 8u;
                        const bool g0 = m0 == 8u, g1 = m1 == 8u;
                        // bytes 8..23 are requested right away as well (same or next sector as the first 16 bytes): a pair that
                        // matches all 8 first bytes then needs no second round trip
                        const uint2 A2 = __ldg(wc0 + 2), A3 = __ldg(wc0 + 3), B2 = __ldg(wc1 + 2), B3 = __ldg(wc1 + 3);
                        if (g0 | g1) {
                            if (g0) {
                                const uint32_t a3w = u0 ? A2.x : A1.y, a4w = u0 ? A2.y : A2.x, a5w = u0 ? A3.x : A2.y, a6w = u0 ? A3.y : A3.x;
                                const uint32_t p3w = wp0[3], p4w = wp0[4], p5w = wp0[5], p6w = wp0[6];
                                const uint64_t z0 = (uint64_t)(__funnelshift_r(p3w, p4w, sp0) ^ __funnelshift_r(a3w, a4w, sc0)) << 32 |
                                                    (__funnelshift_r(p2w, p3w, sp0) ^ __funnelshift_r(a2w, a3w, sc0));
                                const uint64_t z1 = (uint64_t)(__funnelshift_r(p5w, p6w, sp0) ^ __funnelshift_r(a5w, a6w, sc0)) << 32 |
                                                    (__funnelshift_r(p4w, p5w, sp0) ^ __funnelshift_r(a4w, a5w, sc0));
                                m0 = z0 ? 8u + (uint32_t)(__ffsll((long long)z0) - 1) / 8 : z1 ? 16u + (uint32_t)(__ffsll((long long)z1) - 1) / 8 : SEARCH_CAP;
                            }
                            if (g1) {
                                const uint32_t b3w = u1 ? B2.x : B1.y, b4w = u1 ? B2.y : B2.x, b5w = u1 ? B3.x : B2.y, b6w = u1 ? B3.y : B3.x;
                                const uint32_t q3w = wp1[3], q4w = wp1[4], q5w = wp1[5], q6w = wp1[6];
                                const uint64_t z0 = (uint64_t)(__funnelshift_r(q3w, q4w, sp1) ^ __funnelshift_r(b3w, b4w, sc1)) << 32 |
                                                    (__funnelshift_r(q2w, q3w, sp1) ^ __funnelshift_r(b2w, b3w, sc1));
                                const uint64_t z1 = (uint64_t)(__funnelshift_r(q5w, q6w, sp1) ^ __funnelshift_r(b5w, b6w, sc1)) << 32 |
                                                    (__funnelshift_r(q4w, q5w, sp1) ^ __funnelshift_r(b4w, b5w, sc1));
                                m1 = z0 ? 8u + (uint32_t)(__ffsll((long long)z0) - 1) / 8 : z1 ? 16u + (uint32_t)(__ffsll((long long)z1) - 1) / 8 : SEARCH_CAP;
                            }
                        }
                        if (m0 >= MIN_MATCH) {
                            const uint32_t off0 = t0 + l0 - c0;
                            atomicMax(&s_best[l0], (uint32_t)(sel_score(m0, off0) + 9) << 26 | (m0 - MIN_MATCH) << 21 | off0);
                        }
                        if (h1 && m1 >= MIN_MATCH) {
                            const uint32_t off1 = t0 + l1 - c1;
                            atomicMax(&s_best[l1], (uint32_t)(sel_score(m1, off1) + 9) << 26 | (m1 - MIN_MATCH) << 21 | off1);
                        }
                    }
                } else {  // the last bytes of the chunk: careful scalar comparison, one pair per lane per trip
#pragma unroll 1
                    for (uint32_t i = lane; i < total; i += 32) {
                        const uint32_t pr = queue[i], l0 = pr & 1023u, c0 = pr >> 10, pp = t0 + l0;
                        const uint32_t m0 = match_length(in, pp, c0, min(n - pp, SEARCH_CAP), n);
                        if (m0 >= MIN_MATCH) atomicMax(&s_best[l0], (uint32_t)(sel_score(m0, pp - c0) + 9) << 26 | (m0 - MIN_MATCH) << 21 | (pp - c0));
                    }
                }
                __syncwarp();
                LZ_TICK(9);
                uint32_t blen = 0, boff = 0;
                if (searchable) {
                    const uint32_t best = s_best[li];
                    if (best) { blen = ((best >> 21) & 31u) + MIN_MATCH; boff = best & 0x1FFFFFu; }
                }
                // ---- inheritance: a match (off, len) at group position j is a match (off, len - d) at position j + d.  Candidates
                // are ranked by 2 * end - log2(off), which does not depend on the position, so one max-scan over the warp serves
                // every position -- including the ones between the lanes when the search stride is 2.
                uint32_t known = blen;  // verified bytes; a capped match may be longer (the chase extends it from `known`)
                uint32_t win = 0;  // the scan's winner, also serving the positions between the lanes (search stride > 1)
                {
                    const uint32_t gp = SSTRIDE * lane;
                    const uint32_t own = blen ? (gp + blen) << 22 | (blen >= SEARCH_CAP ? 1u : 0u) << 21 | boff : 0u;
                    uint32_t v = own;
                    int32_t e = blen ? (int32_t)(2 * (gp + blen)) - (int32_t)zc::highbit(boff + 3) : -1000;
#pragma unroll
                    for (uint32_t d = 1; d < 32; d <<= 1) {
                        const uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
                        const int32_t eu = __shfl_up_sync(0xffffffffu, e, d);
                        if (lane >= d && eu > e) { v = u; e = eu; }
                    }
                    const uint32_t end = v >> 22;
                    if (v != own && end >= gp + MIN_MATCH && searchable) {
                        known = end - gp; boff = v & 0x1FFFFFu;
                        blen = (v >> 21 & 1u) ? SEARCH_CAP : known;
                    }
                    win = v;
                }
                // what the parser may take at a position: backward extension (<= 3), clamp to the block end, acceptance rule, lazy score
                auto emit = [&](uint32_t l_e, uint32_t blen_e, uint32_t known_e, uint32_t boff_e) {
                    const uint32_t p_e = t0 + l_e;
                    uint32_t bback = 0;
                    if (blen_e) {
                        const uint32_t c = p_e - boff_e;
                        if (c >= 4 && p_e + 8 <= n) {  // one unaligned load each side: bytes [x-4, x)
                            const uint32_t dp = (uint32_t)ld8(in, p_e - 4), dc = (uint32_t)ld8(in, c - 4);
                            const uint32_t diff = dp ^ dc;  // byte 3 is the byte just before the position
                            bback = diff == 0 ? 3u : (uint32_t)__clz((int)diff) >> 3;
                            if (bback > 3) bback = 3;
                        } else {
                            while (bback < 3 && p_e > bback && c > bback && in[p_e - bback - 1] == in[c - bback - 1]) bback++;
                        }
                    }
                    if (p_e < t1) {
                        if (p_e + blen_e > be) blen_e = be - p_e;
                        if (p_e + known_e > be) known_e = be - p_e;
                        int32_t lsc = -1;
                        if (known_e >= MIN_MATCH) { lsc = lazy_score(blen_e, boff_e); if (lsc < ACCEPT_THR) lsc = -1; }
                        s_sc[l_e + HALO] = (int16_t)lsc;
                        s_len[l_e + HALO] = (uint8_t)(lsc >= 0 ? blen_e : 0u);
                        s_off[l_e + HALO] = boff_e;
                        s_back[l_e + HALO] = (uint8_t)(bback | known_e << 2);
                    }
                    if (stride_mode == 0) {  // adaptive stride: positions that sit inside a long match
                        const uint32_t nl = __popc(__ballot_sync(0xffffffffu, known_e >= STAT_LEN));
                        if (lane == 0 && nl) atomicAdd(&s_long, nl);
                    }
                };
                emit(li, blen, known, boff);
                if (SSTRIDE == 2) {
                    const uint32_t d = 1, end = win >> 22, gp = SSTRIDE * lane + d;
                    uint32_t blen1 = 0, known1 = 0, boff1 = 0;
                    if (end >= gp + MIN_MATCH && p + d + 8 <= n) {
                        known1 = end - gp; boff1 = win & 0x1FFFFFu;
                        blen1 = (win >> 21 & 1u) ? SEARCH_CAP : known1;
                    }
                    emit(li + d, blen1, known1, boff1);
                }
                g = gn;
                if (lane == 0) gn = atomicAdd(&s_gctr, 1u);
                gn = __shfl_sync(0xffffffffu, gn, 0);
            }
            LZ_TICK(2);
            // ---- publish the next tile's entries into the slots reserved above ----
            // (the empty asm keeps the compiler from consuming the atomics' results -- and waiting for them -- before the search)
            static_assert(PER_THREAD == 4, "slot barrier lists four registers");
            asm volatile("" : "+r"(slot_raw[0]), "+r"(slot_raw[1]), "+r"(slot_raw[2]), "+r"(slot_raw[3]), "+r"(sw0), "+r"(sw1) :: "memory");
#pragma unroll
            for (uint32_t k = 0; k < PER_THREAD; k++) {
                const uint32_t li = tid + k * THREADS, p = t0 + TILE + li;
                if (p + 8 <= n) {
                    const uint32_t hv = hash5(smem_u64(s_nx, li));
                    __stcg(&tab[(hv >> TAG_BITS) * ROW_K + (slot_raw[k] & (ROW_K - 1))], (p + 1) | (hv & TAG_MASK) << 22);
                }
            }
            __syncthreads();
            LZ_TICK(1);
            // ---- decide: positions [d0, d1) now have their lookahead window available ----
            const uint32_t d0 = t0 >= DEFER ? t0 - DEFER : 0;
            const uint32_t d1 = last_tile_of_block ? t1 : t1 - DEFER;
            bool first = true;
            for (uint32_t p = d0 + tid; p < d1 || first; p += THREADS) {
                // positions before t0 that belong to the previous block were already decided there
                if (p < d1 && !(p < t0 && (p / Z_BLOCK_MAX) != (t0 / Z_BLOCK_MAX))) {
                    int32_t sc = s_sc[p - t0 + HALO];
                    uint32_t r = 0;
                    if (sc >= 0) {
                        uint32_t start = p;
                        while (s_len[start - t0 + HALO] < TARGET_LEN && start - p + 2 <= MAX_SHIFT) {
                            const int32_t s1 = start + 1 < be ? (int32_t)s_sc[start + 1 - t0 + HALO] : -1;
                            if (s1 > sc + 4) { sc = s1; start += 1; continue; }
                            const int32_t s2 = start + 2 < be ? (int32_t)s_sc[start + 2 - t0 + HALO] : -1;
                            if (s2 > sc + 7) { sc = s2; start += 2; continue; }
                            break;
                        }
                        const uint32_t kb = s_back[start - t0 + HALO];
                        r = pack_rec(s_off[start - t0 + HALO], kb >> 2, s_len[start - t0 + HALO] >= TARGET_LEN, kb & 3u, start - p);
                    }
                    __stcs(&rec[p], r);  // streamed: written once, read once by the chase kernel
                }
                if (first) {
                    first = false;
                    // Only warp 0's first trip reads window slots below HALO (positions before t0), so once that trip is over
                    // warp 0 slides the window: the tile's last HALO results become the next tile's halo.
                    if (tid < HALO) {
                        __syncwarp();
                        s_len[tid] = s_len[TILE + tid]; s_off[tid] = s_off[TILE + tid]; s_back[tid] = s_back[TILE + tid]; s_sc[tid] = s_sc[TILE + tid];
                    }
                }
            }
            // the stage buffer of this tile is free now: it receives tile t+2
            {
                uint32_t *dst = reinterpret_cast<uint32_t *>(s_in2[t & 1]);
                dst[tid] = sw0;
                if (tid < SIN_WORDS - THREADS) dst[THREADS + tid] = sw1;
            }
            if (tid == 0) {
                s_gctr = THREADS / 32;
                if (stride_mode == 0) { s_stride = s_long >= adapt_thr ? 2u : 1u; s_long = 0; }
            }
            __syncthreads();
            LZ_TICK(3);
        }
    }
#ifdef SQ_LZ_TIMERS
    if (tid == 0) for (int i = 0; i < 10; i++) atomicAdd(&g_lz_timers[i], (unsigned long long)tm[i]);
#endif
}

// ---- kernel B: chase.  One warp per 128 KiB block walks the records from the block start: lanes hold a
// window of 32 records (one coalesced load), hops inside the window are shuffles, literal runs are skipped
// with a ballot, capped matches are extended 256 bytes per step by the whole warp. ------------------------
__device__ __forceinline__ uint32_t warp_extend(const uint8_t *__restrict__ in, uint32_t n, uint32_t a, uint32_t b, uint32_t maxlen, uint32_t lane) {
    // common prefix of in[a..] and in[b..] (b < a), at most maxlen; all lanes return the same value
    uint32_t done = 0;
    while (done < maxlen) {
        const uint32_t i = done + lane * 8;
        uint32_t m = 8;  // bytes of this lane's 8-byte slot that match (slots past maxlen count as matching)
        if (i < maxlen) {
            const uint32_t lim = min(8u, maxlen - i);
            if (a + i + 16 <= n) {  // both unaligned 8-byte reads stay inside the chunk
                const uint64_t x = ld64_unaligned(in, a + i) ^ ld64_unaligned(in, b + i);
                m = x ? (uint32_t)(__ffsll((long long)x) - 1) / 8 : 8u;
                if (m >= lim) m = 8;
            } else {
                m = 0;
                while (m < lim && in[a + i + m] == in[b + i + m]) m++;
                if (m == lim) m = 8;
            }
        }
        const uint32_t bad = __ballot_sync(0xffffffffu, m != 8);
        if (bad) {
            const uint32_t first = __ffs((int)bad) - 1;
            const uint32_t mm = __shfl_sync(0xffffffffu, m, first);
            return min(maxlen, done + first * 8 + mm);
        }
        done += 256;
    }
    return maxlen;
}

__global__ void __launch_bounds__(128) lz_chase_kernel(const uint8_t *__restrict__ data, const sq_span *__restrict__ spans,
                                                        const uint8_t *__restrict__ select, uint32_t n_chunks,
                                                        const uint32_t *__restrict__ rec_all, zc::Seq *__restrict__ seqs_all,
                                                        BlockMeta *__restrict__ meta_all) {
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
    uint32_t p = bs, anchor = bs, nseq = 0;
    uint32_t r0 = 0, r1 = 0, r2 = 0;  // repeat offsets are unknown at a block start (see file header); the frame's first block knows 1,4,8
    if (b == 0) { r0 = 1; r1 = 4; r2 = 8; }
    // the window after the current one is requested speculatively: unless a long match jumps over it, its records are
    // already in registers when the cursor gets there
    uint32_t pf_base = ~0u, pf = 0;
    while (p < be) {
        const uint32_t base = p & ~31u;
        uint32_t mine;
        if (base == pf_base) mine = pf;
        else mine = (base + lane < be) ? __ldg(rec + base + lane) : 0u;
        pf_base = base + 32;
        pf = (pf_base + lane < be) ? __ldg(rec + pf_base + lane) : 0u;
        uint32_t live = __ballot_sync(0xffffffffu, mine != 0);
        live &= 0xffffffffu << (p - base);  // records at or after the cursor
        while (live) {
            const uint32_t idx = __ffs((int)live) - 1;  // next non-literal record at or after the cursor
            const uint32_t r = __shfl_sync(0xffffffffu, mine, idx);
            p = base + idx;
            uint32_t off = r & 0x1FFFFFu, len = ((r >> 21) & 31u) + 1, back = (r >> 27) & 3u, start = p + (r >> 29);
            if ((r >> 26 & 1u) && start + len < be) len += warp_extend(in, n, start + len, start + len - off, be - start - len, lane);
            if (back > start - anchor) back = start - anchor;
            start -= back; len += back;
            const uint32_t ll = start - anchor;
            // repeat-offset code substitution (RFC 8878 3.1.1.5)
            uint32_t ob = off + 3;
            if (ll) { if (off == r0) ob = 1; else if (off == r1) ob = 2; else if (off == r2) ob = 3; }
            else { if (off == r1) ob = 1; else if (off == r2) ob = 2; else if (r0 > 1 && off == r0 - 1) ob = 3; }
            if (ob > 3) { r2 = r1; r1 = r0; r0 = off; }
            else {
                const uint32_t ix = ob - 1 + (ll ? 0 : 1);
                if (ix == 1) { const uint32_t tmp = r1; r1 = r0; r0 = tmp; }
                else if (ix == 2) { const uint32_t tmp = r2; r2 = r1; r1 = r0; r0 = tmp; }
                else if (ix == 3) { const uint32_t tmp = r0 - 1; r2 = r1; r1 = r0; r0 = tmp; }
            }
            if (lane == 0 && nseq < SEQ_PER_BLOCK) { zc::Seq sq; sq.ll = ll; sq.ml = len; sq.off_base = ob; seqs[nseq] = sq; }
            nseq++;
            p = start + len;
            anchor = p;
            if (p >= base + 32) { live = 0; break; }
            live &= 0xffffffffu << (p - base);
        }
        if (p < base + 32) p = base + 32;  // only literals left in this window
    }
    if (lane == 0) {
        BlockMeta m;
        m.seq_start = b * SEQ_PER_BLOCK; m.nseq = min(nseq, SEQ_PER_BLOCK); m.last_lits = be - anchor; m.reserved = 0;
        meta_all[(size_t)chunk * BLOCKS_PER_CHUNK + b] = m;
    }
}

// ---- entropy stage: one warp per block ------------------------------------------------------------
struct BlockOut { uint32_t body_len, type; };  // mirrors sq_block_info

__global__ void __launch_bounds__(128) entropy_kernel(const uint8_t *__restrict__ data, const sq_span *__restrict__ spans,
                                                       const uint8_t *__restrict__ select, uint32_t n_chunks,
                                                       const zc::Seq *__restrict__ seqs_all, const BlockMeta *__restrict__ meta_all,
                                                       uint8_t *__restrict__ lit_all, uint8_t *__restrict__ bodies, BlockOut *__restrict__ blocks,
                                                       uint32_t *__restrict__ sbits_all, uint32_t *__restrict__ counter) {
    __shared__ ent::WarpWork s_work[4];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    ent::WarpWork *W = &s_work[threadIdx.x >> 5];
    uint8_t *lits = lit_all + (size_t)warp_global * (Z_BLOCK_MAX + 64);
    uint32_t *sbits = sbits_all + (size_t)warp_global * SBITS_STRIDE;
    uint8_t *cbytes = reinterpret_cast<uint8_t *>(sbits + 4 * SEQ_PER_BLOCK);  // 16-byte aligned: 4 * SEQ_PER_BLOCK and SBITS_STRIDE are multiples of 4 words
    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_chunks * BLOCKS_PER_CHUNK) break;
        const uint32_t chunk = item / BLOCKS_PER_CHUNK, b = item % BLOCKS_PER_CHUNK;
        if (select && !select[chunk]) continue;
        const uint32_t n = spans[chunk].len;
        const uint32_t nb = n == 0 ? 1 : (n + Z_BLOCK_MAX - 1) / Z_BLOCK_MAX;
        if (b >= nb) continue;
        const uint8_t *in = data + spans[chunk].off;
        const uint32_t bs = b * Z_BLOCK_MAX, be = min(n, bs + Z_BLOCK_MAX), blen = be - bs;
        if (blen == 0) { if (lane == 0) { BlockOut o0; o0.type = 0; o0.body_len = 0; blocks[(size_t)chunk * BLOCKS_PER_CHUNK + b] = o0; } continue; }
        const BlockMeta m = meta_all[(size_t)chunk * BLOCKS_PER_CHUNK + b];
        const zc::Seq *seqs = seqs_all + (size_t)chunk * MAX_SEQ_PER_CHUNK + m.seq_start;
        if (m.nseq <= 2) {  // a block this regular may be one repeated byte: then it is an RLE block (1-byte body)
            const uint8_t v0 = in[bs];
            bool same = true;
            for (uint32_t i = bs + lane; i < be && same; i += 32) same = in[i] == v0;
            if (__all_sync(0xffffffffu, same)) {
                if (lane == 0) {
                    uint8_t *dst = bodies + ((size_t)chunk * BLOCKS_PER_CHUNK + b) * (size_t)BODY_STRIDE;
                    dst[0] = v0;
                    BlockOut o1; o1.type = 1; o1.body_len = 1;
                    blocks[(size_t)chunk * BLOCKS_PER_CHUNK + b] = o1;
                }
                continue;
            }
        }
        // gather literals: 32 sequences at a time, a warp scan gives every lane its source and destination offsets
        uint32_t src_pos = bs, lit_pos = 0;
        for (uint32_t base = 0; base < m.nseq; base += 32) {
            const uint32_t i = base + lane;
            uint32_t ll = 0, ml = 0;
            if (i < m.nseq) { ll = seqs[i].ll; ml = seqs[i].ml; }
            uint32_t adv = ll + ml, incl_adv = adv, incl_ll = ll;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t a = __shfl_up_sync(0xffffffffu, incl_adv, d), l2 = __shfl_up_sync(0xffffffffu, incl_ll, d);
                if ((int)lane >= d) { incl_adv += a; incl_ll += l2; }
            }
            const uint32_t my_src = src_pos + incl_adv - adv, my_lit = lit_pos + incl_ll - ll;
            for (uint32_t k = 0; k < ll; k++) lits[my_lit + k] = in[my_src + k];
            src_pos += __shfl_sync(0xffffffffu, incl_adv, 31);
            lit_pos += __shfl_sync(0xffffffffu, incl_ll, 31);
        }
        for (uint32_t k = lane; k < m.last_lits; k += 32) lits[lit_pos + k] = in[src_pos + k];
        const uint32_t nlits = lit_pos + m.last_lits;
        __syncwarp();
        __threadfence_block();
        BlockOut out;
        out.type = 0; out.body_len = blen;  // raw unless the compressed body is smaller
        // the body slot holds literals + <= 8 bytes per sequence; a parse that cannot fit cannot beat raw either
        if (src_pos + m.last_lits == be && zc::block_body_bound(nlits, m.nseq) <= BODY_STRIDE) {
            uint8_t *dst = bodies + ((size_t)chunk * BLOCKS_PER_CHUNK + b) * (size_t)BODY_STRIDE;
            uint32_t sz = ent::warp_write_literals(dst, lits, nlits, W, lane);
            sz += ent::warp_write_sequences(dst + sz, seqs, m.nseq, W, sbits, cbytes, SEQ_CODE_STRIDE, lane);
            if (sz < blen) { out.type = 2; out.body_len = sz; }
        }
        if (lane == 0) blocks[(size_t)chunk * BLOCKS_PER_CHUNK + b] = out;
        __syncwarp();
    }
}

}  // namespace lz
