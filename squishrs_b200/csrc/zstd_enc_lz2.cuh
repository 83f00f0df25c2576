// K3 stage "compress": match search (lz2::index_kernel + lz2::search_kernel) and a parse that knows about repeat offsets
// (lz2::chase_kernel).  The entropy stage is lz::entropy_kernel.
//
// A GPU has the whole chunk before it starts, so the match finder does not have to be the incremental "insert, then search"
// structure of a streaming encoder.  The search is split into two kernels with no synchronisation inside either:
//
//   index_kernel   one CTA per chunk.  Every position is hashed (5 bytes; 4 in chunks <= 128 KiB) into one of 2^14 rows; a
//                  counting sort in shared memory (histogram, prefix sum, scatter in position order) writes the chunk's
//                  LIST: the entries of row 0 in position order, then row 1, ...  entry = position | ptag << 21 | xtag << 26
//                  (ptag = 5 more bits of the row hash, xtag = a 6-bit hash of bytes 5..7).  Every position also gets the
//                  end E of its window in the list: its row's fill level after the 512-position tile the position lies in
//                  (stored in the record array, which the search overwrites with its result).
//   search_kernel  any warp can take any 256 positions of any chunk: the candidates of position p are the 32 list entries in
//                  front of E(p), i.e. the 32 most recent positions of the same row up to the end of p's tile -- what a
//                  32-entry ring per row would hold when a sequential encoder reaches p (plus at most a tile of lookahead),
//                  without the ring, its atomics, its clearing, or any ordering between positions.
//
// History.  Round 1 kept one 2 MB table per resident CTA in HBM (142 B of DRAM traffic per input byte).  The first round-2
// version put a thread-block cluster on each chunk with an L2-resident ring table and a cluster barrier every 2048 positions:
// 14 B of DRAM per byte, but 43 % of all warp stalls sat on that barrier (each warp had ONE group of 32 positions per step, so
// every step ran at the pace of the slowest of 64 warps), and positions inserted ahead of the searcher took ring slots.  The
// list removes both: no barrier, and a window holds at most one tile (512 positions) of later entries instead of 2048+.
//
// One xor with the searcher's own tag word classifies an entry without touching the candidate's bytes:
//     long   all 11 tag bits agree  -> the candidate very likely shares >= 8 bytes           (x = e ^ T < p)
//     short  only ptag agrees       -> it shares the 5-byte prefix but not 8 bytes           ((x & 0x03FFFFFF) < p, x != that)
// Every long candidate is examined (after the continuation filter below); of the short ones only the nearest, because among
// matches of 5..7 bytes only the offset matters (CPU model tests/harness/lz_model2.cc).
//
// Per group of 32 positions (one warp):
//   rows      four lanes share a 128-byte window of the list (two 16-byte loads each, the window start rounded up to 16 bytes:
//             the 29..32 entries in front of E), 8 positions per pass, 4 passes; all 8 loads of a lane are in flight together
//   filter    long candidates -> (position, offset) pairs in the warp's queue; nearest short per position kept aside
//   continue  a long pair (p, o) whose left neighbour (p-1, o) is also a pair continues a match that is (or will be) found one
//             position earlier: dropped, its result arrives by inheritance.  Done in pair space through a direct-mapped table
//             of 16-bit keys (hash of the offset, slot shifted by the column); column 0 keeps everything so inheritance never
//             runs dry.
//   verify    one pair per lane per trip: 40 candidate bytes (five aligned 8-byte loads, all requested up front) against the
//             position's bytes in shared memory, up to CAP = 32 matching bytes; best per position by atomicMax on
//             (2 len - log2 offset, nearer offset first)
//   inherit   one max-scan over the warp hands a match at position j to j + d as (offset, len - d)
//   emit      <= 3 bytes of backward extension, clamp to the block end, one 4-byte record per position (streaming store)
//
// chase_kernel: one warp per 128 KiB block walks the records.  At the cursor it resolves the lazy choice (depth 2, zstd's
// gains) from the 64 records it holds in registers, extends capped matches warp-wide, and looks for repeat-offset
// matches the way a sequential encoder does: every literal position in front of the chosen match is tested against the three
// repeat offsets (>= 3 bytes is enough, they cost almost nothing to code), and at the match start a repeat offset wins when
// zstd's rule of thumb says so.
#pragma once
#include <cooperative_groups.h>
#include "zstd_enc_lz.cuh"
#include "zstd_enc_parse.h"

namespace lz2 {

constexpr uint32_t ROW_LOG = 14, ROWS = 1u << ROW_LOG, ROW_K = 32;
constexpr uint32_t POS_BITS = 21, POS_MASK = (1u << POS_BITS) - 1, PTAG_BITS = 5, PFX_MASK = (1u << (POS_BITS + PTAG_BITS)) - 1;
constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr uint32_t LEN_BASE = zparse::LEN_BASE, CAP = zparse::CAP, MAX_SHIFT = zparse::MAX_SHIFT, REP_MIN = zparse::REP_MIN;
constexpr uint32_t SMALL_CHUNK = 128u * 1024u;  // up to here matches of 4 bytes are searched (hash of 4 bytes), above 5: libzstd's level-12 parameters make the same switch
constexpr int32_t ACCEPT_THR = zparse::ACCEPT_THR;
constexpr uint32_t LOOKAHEAD = CAP + 16;
constexpr uint32_t QUEUE_WORDS = 32 * ROW_K + 32;  // per warp: every entry of every window could be a long candidate, plus one short per position
constexpr uint32_t SEQ_PER_BLOCK = lz::SEQ_PER_BLOCK, MAX_SEQ_PER_CHUNK = lz::MAX_SEQ_PER_CHUNK, BLOCKS_PER_CHUNK = lz::BLOCKS_PER_CHUNK;
constexpr uint32_t REC_PER_CHUNK = lz::REC_PER_CHUNK;
// the list of one chunk: LIST_PAD empty entries (a window may start up to 32 entries in front of the first row), then at most
// one entry per position; slots are 16-byte aligned
constexpr uint32_t LIST_PAD = 64, LIST_STRIDE = REC_PER_CHUNK + LIST_PAD + 64;
constexpr uint32_t SPAN = 256;  // positions a search warp takes at a time

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

// 8 bytes at position p of the chunk, straight from global memory (any alignment)
__device__ __forceinline__ uint64_t gld8(const uint8_t *__restrict__ in, uint32_t p, uint32_t n) {
    if (p + 12 <= n) return lz::ld8(in, p);
    uint64_t v = 0;
    for (uint32_t k = 0; k < 8 && p + k < n; k++) v |= (uint64_t)in[p + k] << (8 * k);
    return v;
}

// ---- plan: 256-position spans of the selected chunks of one sub-batch, as an exclusive prefix (one CTA) ------------------
__global__ void __launch_bounds__(1024) span_plan_kernel(const sq_span *__restrict__ spans, const uint8_t *__restrict__ select, uint32_t first,
                                                          uint32_t count, uint32_t *__restrict__ span_start, uint32_t *__restrict__ counter) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < count; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        uint32_t v = 0;
        if (i < count && (!select || select[first + i])) v = (spans[first + i].len + SPAN - 1) / SPAN;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if ((int)lane >= d) x += y; }
        if (lane == 31) warp_sums[w] = x;
        __syncthreads();
        if (w == 0) {
            const uint32_t s = warp_sums[lane];
            uint32_t z = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, z, d); if ((int)lane >= d) z += y; }
            warp_sums[lane] = z - s;
        }
        __syncthreads();
        const uint32_t excl = carry + warp_sums[w] + x - v;
        if (i < count) span_start[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { span_start[count] = carry; *counter = 0; }  // the search kernel's ticket counter
}

// ---- index: the chunk's list (rows in position order) and every position's window end ---------------------------------
// Inside a run of one byte (the five bytes here are the five bytes one position earlier) nothing is inserted: the run's first
// position stands for it, and a hot row is not flooded with interchangeable entries.
//
// One CTA per chunk (two CTAs of 512 threads per SM), two passes.  Pass 1 hashes every position once -- eight consecutive positions per thread from three aligned
// 8-byte loads -- counts the rows in shared memory and leaves a 26-bit word per position in a scratch array (row + tags +
// "insert" flag).  After a prefix sum over the rows, pass 2 reads those words back (coalesced) and scatters the entries, one tile
// of THREADS positions at a time.  Inside a tile the order of two entries of one row is the order of their shared-memory
// atomics, so a position does not take "its own index" as the end of its window but the row's fill level E at the END of its
// tile (or a little later: one barrier per tile, see pass 2): the window [E - 32, E) then holds every earlier position of the tile regardless of the race (and the tile's later ones,
// which the search discards by position).  A tile is 512 positions.
// A cluster of PARTS CTAs works on one chunk: pass 1 deals the positions to the CTAs (their histograms are summed through
// distributed shared memory), in pass 2 every CTA walks all the words and scatters the rows of its own part.  The scatter is
// 16384 append streams of 4-byte stores per chunk; with one CTA per chunk 296 chunks were in flight and their open 32-byte
// sectors (155 MB) fell out of the L2 half written, every 4-byte store becoming a DRAM sector write (measured: 1.4 TB/s of DRAM
// traffic, 18 ms per 410 chunks whatever the CTA shape).  A quarter of the chunks in flight keeps every open sector in L2 until
// it is full.
constexpr uint32_t IDX_INS = 1u << 25;
__device__ __forceinline__ uint32_t index_word(uint64_t v, uint32_t prev, uint32_t hshift, bool first) {
    const uint32_t hv = hash_row_ptag(v, hshift), xt = ((uint32_t)(v >> 40) * 0x9E3779B1u) >> 26;
    const bool run = !first && (uint32_t)v == prev * 0x01010101u && ((uint32_t)(v >> 32) & 0xFFu) == prev;
    return hv | xt << (ROW_LOG + PTAG_BITS) | (run ? 0u : IDX_INS);
}

template <int THREADS, int PARTS>
__global__ void __launch_bounds__(THREADS, 2048 / THREADS > 2 ? 2 : 2048 / THREADS) index_kernel(const uint8_t *__restrict__ data, const sq_span *__restrict__ spans,
                                                            const uint8_t *__restrict__ select, uint32_t first, uint32_t count,
                                                            uint32_t *__restrict__ list_all, uint32_t *__restrict__ words_all, uint32_t *__restrict__ rec_all,
                                                            uint32_t ordered) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) uint32_t s_cnt[];  // ROWS counters: this CTA's histogram, then (own part only) running list indices
    __shared__ uint32_t s_warp[THREADS / 32], s_part_total, s_turn;
    constexpr uint32_t PART_ROWS = ROWS / PARTS, PER = PART_ROWS / THREADS, PART_SHIFT = ROW_LOG - (PARTS == 4 ? 2 : PARTS == 2 ? 1 : 0);
    static_assert(PARTS == 1 || PARTS == 2 || PARTS == 4, "cluster size");
    static_assert(PART_ROWS % THREADS == 0, "row counters per thread");
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t part = PARTS > 1 ? cluster.block_rank() : 0u;
    for (uint32_t ci = blockIdx.x / PARTS; ci < count; ci += gridDim.x / PARTS) {  // the same for every CTA of a cluster
        const uint32_t chunk = first + ci;
        if (select && !select[chunk]) continue;
        const uint8_t *in = data + spans[chunk].off;
        const uint32_t n = spans[chunk].len;
        uint32_t *rec = rec_all + (size_t)chunk * REC_PER_CHUNK;
        uint32_t *list = list_all + (size_t)ci * LIST_STRIDE + LIST_PAD;
        uint32_t *words = words_all + (size_t)ci * REC_PER_CHUNK;  // pass 1 -> pass 2; not the record array: a CTA that is ahead already writes window ends there
        const uint32_t hshift = n <= SMALL_CHUNK ? 32u : 24u;
        const uint32_t np = n >= 8 ? n - 7 : 0;  // positions that have 8 bytes
        for (uint32_t i = tid; i < ROWS; i += THREADS) s_cnt[i] = 0;
        __syncthreads();
        // pass 1 (positions dealt to the CTAs of the cluster): one word per position, row histogram
        uint32_t done = 0;  // positions covered by the vector path
        if ((reinterpret_cast<uintptr_t>(in) & 7) == 0 && n >= 16) {
            const uint32_t n8 = (n - 16) / 8 + 1;  // groups of eight positions whose 16 bytes lie inside the chunk
            done = n8 * 8;
            for (uint32_t i = part * THREADS + tid; i < n8; i += PARTS * THREADS) {
                const uint32_t p0 = i * 8;
                const uint2 B = __ldg(reinterpret_cast<const uint2 *>(in + p0)), Cw = __ldg(reinterpret_cast<const uint2 *>(in + p0 + 8));
                uint32_t prev = p0 ? __ldg(reinterpret_cast<const uint32_t *>(in + p0 - 4)) >> 24 : 0u;
                const uint32_t wd[4] = {B.x, B.y, Cw.x, Cw.y};
                uint32_t out[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t sh = (k & 3) * 8, a = k >> 2;
                    const uint32_t lo = __funnelshift_r(wd[a], wd[a + 1], sh), hi = __funnelshift_r(wd[a + 1], wd[a + 2], sh);
                    const uint32_t wv = index_word((uint64_t)hi << 32 | lo, prev, hshift, p0 + k == 0);
                    if (wv & IDX_INS) atomicAdd(&s_cnt[(wv >> PTAG_BITS) & (ROWS - 1)], 1u);
                    out[k] = wv;
                    prev = lo & 0xFFu;
                }
                uint4 *dst = reinterpret_cast<uint4 *>(words + p0);
                __stcg(dst, make_uint4(out[0], out[1], out[2], out[3]));
                __stcg(dst + 1, make_uint4(out[4], out[5], out[6], out[7]));
            }
        }
        for (uint32_t p = done + part * THREADS + tid; p < np; p += PARTS * THREADS) {  // the chunk's last positions, or all of an unaligned chunk
            const uint32_t wv = index_word(gld8(in, p, n), p ? in[p - 1] : 0u, hshift, p == 0);
            if (wv & IDX_INS) atomicAdd(&s_cnt[(wv >> PTAG_BITS) & (ROWS - 1)], 1u);
            __stcg(words + p, wv);
        }
        // every CTA sums the cluster's histograms over its own part of the rows
        uint32_t c[PER], sum = 0;
        if (PARTS > 1) cluster.sync(); else __syncthreads();
#pragma unroll
        for (uint32_t k = 0; k < PER; k++) {
            const uint32_t row = part * PART_ROWS + tid * PER + k;
            uint32_t t = 0;
            if (PARTS > 1) { for (uint32_t r = 0; r < (uint32_t)PARTS; r++) t += *cluster.map_shared_rank(&s_cnt[row], r); }
            else t = s_cnt[row];
            c[k] = t; sum += t;
        }
        if (PARTS > 1) cluster.sync();  // nobody reads the partial counts any more (and all words of pass 1 are visible)
        // exclusive prefix over the rows of this part, then shifted by the totals of the parts in front
        {
            uint32_t x = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if ((int)lane >= d) x += y; }
            if (lane == 31) s_warp[w] = x;
            __syncthreads();
            if (w == 0) {
                const uint32_t sv = lane < THREADS / 32 ? s_warp[lane] : 0u;
                uint32_t z = sv;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, z, d); if ((int)lane >= d) z += y; }
                if (lane < THREADS / 32) s_warp[lane] = z - sv;
                if (lane == 31) s_part_total = z;
            }
            if (PARTS > 1) cluster.sync(); else __syncthreads();
            uint32_t run = s_warp[w] + x - sum;
            if (PARTS > 1) for (uint32_t r = 0; r < part; r++) run += *cluster.map_shared_rank(&s_part_total, r);
#pragma unroll
            for (uint32_t k = 0; k < PER; k++) { s_cnt[part * PART_ROWS + tid * PER + k] = run; run += c[k]; }
        }
        if (tid == 0) s_turn = 0;
        __syncthreads();
        // pass 2 (every CTA walks all positions and keeps the rows of its part): scatter, tile by tile; the words of the next
        // DEPTH tiles are requested before this round's tiles are worked on
        constexpr uint32_t DEPTH = 4;
        uint32_t wn[DEPTH];
#pragma unroll
        for (uint32_t k = 0; k < DEPTH; k++) { const uint32_t p = k * THREADS + tid; wn[k] = p < np ? __ldcg(words + p) : 0u; }
        for (uint32_t t0 = 0; t0 < np; t0 += DEPTH * THREADS) {
            uint32_t wc[DEPTH];
#pragma unroll
            for (uint32_t k = 0; k < DEPTH; k++) {
                wc[k] = wn[k];
                const uint32_t p = t0 + (DEPTH + k) * THREADS + tid;
                if (p < np) wn[k] = __ldcg(words + p);
            }
#pragma unroll
            for (uint32_t k = 0; k < DEPTH; k++) {
                const uint32_t p = t0 + k * THREADS + tid, wv = wc[k];
                if (t0 + k * THREADS >= np) break;  // uniform over the CTA
                const uint32_t row = (wv >> PTAG_BITS) & (ROWS - 1);
                const bool mine = p < np && (row >> PART_SHIFT) == part;
                const bool ins = mine && (wv & IDX_INS);
                const uint32_t entry = p | (wv & ((1u << PTAG_BITS) - 1)) << POS_BITS | (wv >> (ROW_LOG + PTAG_BITS) & 63u) << (POS_BITS + PTAG_BITS);
                if (!ordered) {
                    if (ins) list[atomicAdd(&s_cnt[row], 1u)] = entry;
                } else {
                    // SQ_FLAG_DETERMINISTIC: a row's entries land in exact position order.  Lanes of one row rank themselves inside
                    // the warp, one of them adds the group's count, and the warps of a tile take turns at the counters.
                    const uint32_t peers = __match_any_sync(0xffffffffu, ins ? row : 0x80000000u | lane);
                    const uint32_t leader = (uint32_t)__ffs((int)peers) - 1, rank = __popc(peers & ((1u << lane) - 1u));
                    const uint32_t my_turn = ((t0 / THREADS) + k) * (THREADS / 32) + w;
                    if (lane == 0) while (*reinterpret_cast<volatile uint32_t *>(&s_turn) != my_turn) __nanosleep(20);
                    __syncwarp();
                    uint32_t base = 0;
                    if (ins && lane == leader) base = atomicAdd(&s_cnt[row], (uint32_t)__popc(peers));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&s_turn) = my_turn + 1;
                    if (ins) list[base + rank] = entry;
                }
                __syncthreads();
                if (mine) rec[p] = s_cnt[row];  // read while faster warps already scatter the next tile: E may come out a few entries later, which only moves the window
                if (ordered) __syncthreads();   // ... unless the bytes have to be reproducible
            }
        }
        // What a window can see beside real entries is EMPTY, not whatever the memory held before: the pad in front of the first row
        // (windows of the first positions start there) and the entries right behind the end of the list (a window is rounded up
        // to 16 bytes).  Left-over bits could pass the 5-bit prefix test and displace a position's true nearest short candidate.
        if (part == 0 && tid < LIST_PAD) list[(int32_t)tid - (int32_t)LIST_PAD] = EMPTY;
        if (part == PARTS - 1) {
            __syncthreads();
            if (tid < 8) list[s_cnt[ROWS - 1] + tid] = EMPTY;
        }
        // no CTA of the cluster leaves (or reuses s_part_total) while another may still read its shared memory
        if (PARTS > 1) cluster.sync(); else __syncthreads();
    }
}

// ---- search ------------------------------------------------------------------------------------------------------------
template <int TLOG> struct SearchSmem {
    static constexpr uint32_t WIN = SPAN + LOOKAHEAD + 32;  // staged bytes of a span
    static constexpr uint32_t PER_WARP = QUEUE_WORDS * 4 + (4u << TLOG) + WIN + 32 * 4;
    static_assert(WIN % 16 == 0 && PER_WARP % 16 == 0, "per-warp shared memory shape");
};

template <int THREADS, int MINB, int TLOG>
__global__ void __launch_bounds__(THREADS, MINB) search_kernel(const uint8_t *__restrict__ data, const sq_span *__restrict__ spans, uint32_t first,
                                                               uint32_t count, const uint32_t *__restrict__ span_start,
                                                               const uint32_t *__restrict__ list_all, uint32_t *__restrict__ rec_all, uint32_t *__restrict__ counter, uint32_t dbg) {
    constexpr uint32_t WARPS = THREADS / 32, WIN = SearchSmem<TLOG>::WIN, WIN_WORDS = WIN / 4, GROUPS = SPAN / 32, TMASK = (2u << TLOG) - 1;
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const uint32_t tid = threadIdx.x, wq = tid >> 5, lane = tid & 31;
    uint8_t *s_warp = s_dyn + wq * SearchSmem<TLOG>::PER_WARP;
    uint32_t *queue = reinterpret_cast<uint32_t *>(s_warp);
    // continuation keys, 16 bits each (twice the slots of a 32-bit table in the same memory: fewer keys overwritten before they are
    // looked up, so fewer pairs survive the filter for nothing): 10 bits of the offset's hash | epoch << 10; never cleared, the
    // epoch changes with every group
    uint16_t *T = reinterpret_cast<uint16_t *>(queue + QUEUE_WORDS);
    uint8_t *s_in = reinterpret_cast<uint8_t *>(queue + QUEUE_WORDS + (1u << TLOG));
    uint32_t *s_best = reinterpret_cast<uint32_t *>(s_in + WIN);
    const uint32_t sub = lane >> 2, part = lane & 3u;
    uint32_t epoch = 0;
    const uint32_t total_spans = span_start[count];

    // Spans are handed out through one counter, in order: all warps of the grid then work within a few hundred KB of each other,
    // so the chunk they read -- input, list, records -- stays in L2.  (A fixed warp -> span assignment let fast warps run tens of
    // chunks ahead of slow ones: 473 MB of DRAM reads per 2 MiB chunk, L2 hit rate 43 %.)  The next ticket is requested at the
    // top of a span and first looked at when the span is done.
    uint32_t sp = 0;
    if (lane == 0) sp = atomicAdd(counter, 1u);
    sp = __shfl_sync(0xffffffffu, sp, 0);
    while (sp < total_spans) {
        uint32_t sp_next = 0;
        if (lane == 0) sp_next = atomicAdd(counter, 1u);
        // which chunk: the last ci with span_start[ci] <= sp
        uint32_t lo = 0, hi = count;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (span_start[mid] <= sp) lo = mid; else hi = mid; }
        const uint32_t chunk = first + lo, t0 = (sp - span_start[lo]) * SPAN;
        const uint8_t *in = data + spans[chunk].off;
        const uint32_t n = spans[chunk].len;
        uint32_t *rec = rec_all + (size_t)chunk * REC_PER_CHUNK;
        const uint32_t *list = list_all + (size_t)lo * LIST_STRIDE + LIST_PAD;
        const bool aligned = (reinterpret_cast<uintptr_t>(in) & 7) == 0;
        const uint32_t MIN_MATCH = n <= SMALL_CHUNK ? 4u : 5u, hshift = n <= SMALL_CHUNK ? 32u : 24u;
        const uint32_t t1 = min(n, t0 + SPAN);
        __syncwarp();
        // the continuation table starts empty for every span (16 bytes per lane and store): what it answers then depends on this
        // span alone, not on what the warp -- or a kernel before it -- left in shared memory
        for (uint32_t i = lane; i < (2u << TLOG) / 8; i += 32) reinterpret_cast<uint4 *>(T)[i] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        epoch = 0;
        for (uint32_t i = lane; i < WIN_WORDS; i += 32) reinterpret_cast<uint32_t *>(s_in)[i] = lz::stage_word(in, t0 + i * 4, n, aligned);
        // every position's window end, for the whole span (the index kernel left it in the record array)
        uint32_t jv[GROUPS];
#pragma unroll
        for (uint32_t g = 0; g < GROUPS; g++) { const uint32_t p = t0 + g * 32 + lane; jv[g] = p + 8 <= n ? __ldcs(rec + p) : 0u; }
        __syncwarp();
        uint32_t carry = 0;  // offset of a capped match that reaches the end of the previous group of this span
#pragma unroll 1
        for (uint32_t g = 0; g < GROUPS && t0 + g * 32 < n; g++) {
            const uint32_t gl = g * 32, li = gl + lane, p = t0 + li, pg = t0 + gl;
            const bool searchable = p + 8 <= n;
            const bool gfast = pg + 32 + CAP + 16 <= n;
            uint32_t j_own = jv[0];
#pragma unroll
            for (uint32_t k = 1; k < GROUPS; k++) if (g == k) j_own = jv[k];
            // A group that lies inside a match already known to be long is not searched: when the record in front of the group is
            // "32 bytes verified, may be longer" and the next 64 bytes really continue that match, every position of the group
            // inherits (offset, 32, may be longer) -- a parse that took the match never looks at these records, one that arrives
            // from elsewhere finds the same match.  (CPU model: 3-15 % of all groups, no measurable change in size.)
            uint32_t blen = 0, boff = 0, capped = 0;
            bool skipg = false;
            if (carry && gfast) {  // warp-uniform
                bool same = true;
                if (lane < 16) {
                    const uintptr_t ga = reinterpret_cast<uintptr_t>(in + pg - carry) + 4 * lane;
                    const uint32_t *wc = reinterpret_cast<const uint32_t *>(ga & ~(uintptr_t)3);
                    same = __funnelshift_r(__ldg(wc), __ldg(wc + 1), (uint32_t)(ga & 3u) * 8) == reinterpret_cast<const uint32_t *>(s_in)[(gl >> 2) + lane];
                }
                skipg = __all_sync(0xffffffffu, same);
            }
            if (skipg) { blen = CAP; boff = carry; capped = 1; }
            else {
            // windows: pass k serves position 8 k + sub; this lane reads entries [4 part, +4) and [16 + 4 part, +4) of that window
            uint4 ea[4], eb[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int32_t jk = (int32_t)__shfl_sync(0xffffffffu, j_own, 8 * k + sub);
                const uint4 *rp = reinterpret_cast<const uint4 *>(list + ((jk + 3 - (int32_t)ROW_K) & ~3)) + part;
                ea[k] = __ldcg(rp); eb[k] = __ldcg(rp + 4);
            }
            const uint64_t v_own = lz::smem_u64(s_in, li);
            const uint32_t hv_own = hash_row_ptag(v_own, hshift);
            const uint32_t T_own = tag_word(hv_own, v_own);
            epoch = (epoch + 1u) & 63u;
            s_best[lane] = 0u;
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
            const uint32_t ep = epoch << 10;
            for (uint32_t i = lane; i < nlong; i += 32) {
                const uint32_t pr = queue[i], col = pr >> POS_BITS, ho = (pg + col - (pr & POS_MASK)) * 0x9E3779B1u;
                T[((ho >> 11) + col) & TMASK] = (uint16_t)((ho >> 1 & 0x3FFu) | ep);
            }
            __syncwarp();
            uint32_t total = 0;
            for (uint32_t base = 0; base < nlong; base += 32) {
                const uint32_t i = base + lane;
                const bool valid = i < nlong;
                const uint32_t pr = valid ? queue[i] : 0u, col = pr >> POS_BITS, ho = (pg + col - (pr & POS_MASK)) * 0x9E3779B1u;
                const bool hit = T[((ho >> 11) + col - 1u) & TMASK] == (uint16_t)((ho >> 1 & 0x3FFu) | ep);  // the pair one column to the left has the same offset
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
                    const uint32_t sp8 = (l0 & 3u) * 8, sc = (uint32_t)(ga & 3u) * 8;
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
                        d[j] = (uint64_t)(__funnelshift_r(pw[2 * j + 1], pw[2 * j + 2], sp8) ^ __funnelshift_r(cw[2 * j + 1], cw[2 * j + 2], sc)) << 32 |
                               (__funnelshift_r(pw[2 * j], pw[2 * j + 1], sp8) ^ __funnelshift_r(cw[2 * j], cw[2 * j + 1], sc));
                    uint64_t x = d[0]; uint32_t mb = 0;
                    if (!x) { x = d[1]; mb = 8; if (!x) { x = d[2]; mb = 16; if (!x) { x = d[3]; mb = 24; } } }
                    const uint32_t m = x ? mb + ((uint32_t)(__ffsll((long long)x) - 1) >> 3) : CAP;
                    if (m >= MIN_MATCH)
                        atomicMax(&s_best[pr >> POS_BITS], (uint32_t)((int32_t)(2 * m) - (int32_t)zc::highbit(o + 3) + 12) << POS_BITS | (POS_MASK - o));
                }
            } else {  // the last bytes of the chunk: careful scalar comparison
#pragma unroll 1
                for (uint32_t i = lane; i < total; i += 32) {
                    const uint32_t pr = queue[i], l0 = gl + (pr >> POS_BITS), pp = t0 + l0, o = pp - (pr & POS_MASK);
                    const uint32_t m = lz::match_length(in, pp, pp - o, min(n - pp, CAP), n);
                    if (m >= MIN_MATCH)
                        atomicMax(&s_best[pr >> POS_BITS], (uint32_t)((int32_t)(2 * m) - (int32_t)zc::highbit(o + 3) + 12) << POS_BITS | (POS_MASK - o));
                }
            }
            __syncwarp();
            // ---- inherit + emit ----
            if (searchable) {
                const uint32_t best = s_best[lane];
                if (best) {
                    boff = POS_MASK - (best & POS_MASK);
                    blen = ((best >> POS_BITS) - 12u + zc::highbit(boff + 3)) >> 1;
                }
            }
            capped = blen >= CAP ? 1u : 0u;
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
            }  // searched group
            carry = __shfl_sync(0xffffffffu, capped ? boff : 0u, 31);
            if (p < t1) {
                uint32_t r = 0;
                if (blen) {
                    const uint32_t be = min(n, (p / Z_BLOCK_MAX + 1) * Z_BLOCK_MAX);
                    if (p + blen > be) { blen = be - p; capped = 0; }
                    if (blen >= MIN_MATCH) {
                        const uint32_t c = p - boff;
                        uint32_t bback = 0;
                        if (skipg) bback = 0;  // inside a long match: nobody starts here with literals in front
                        else if (c >= 4 && p + 8 <= n) {  // one unaligned load each side: bytes [x-4, x)
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
            __syncwarp();
        }
        sp = __shfl_sync(0xffffffffu, sp_next, 0);
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
                                                     lz::BlockMeta *__restrict__ meta_all, uint32_t *__restrict__ counter) {
    const uint32_t lane = threadIdx.x & 31;
    // blocks are handed out through a counter: a warp that drew a short block (or one of an unselected chunk) takes the next one
    // instead of idling until the slowest warp of its CTA is done (a fixed block -> warp assignment reached 39 % occupancy of 62 %)
    for (;;) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(counter, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_chunks * BLOCKS_PER_CHUNK) return;
    const uint32_t chunk = item / BLOCKS_PER_CHUNK, b = item % BLOCKS_PER_CHUNK;
    if (select && !select[chunk]) continue;
    const uint32_t n = spans[chunk].len;
    const uint32_t bs = b * Z_BLOCK_MAX;
    if (bs >= n) continue;
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
