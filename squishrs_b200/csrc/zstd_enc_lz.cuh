// K3 stage "compress": LZ match search + parse (lz_kernel) and entropy coding (entropy_kernel).
//
// lz_kernel — one persistent CTA per chunk in flight, 256 threads, the chunk walked in tiles of
// 1024 positions.  Per tile:
//   stage   the tile's bytes (+16 lookahead) into shared memory
//   insert  every position into the chunk's bucketed hash table in HBM: 2^15 rows x 32 entries,
//           entry = (position+1) | 10-bit tag << 22, ring slot taken with atomicAdd on the row head.
//           The table is never cleared between chunks: a stale entry is just a candidate position,
//           and every candidate is verified against the bytes of the CURRENT chunk.
//   search  every position reads its 128-byte row (4 x 128-bit loads, L1 bypassed), filters by tag,
//           verifies candidates against the input and keeps the best by 2*len - log2(offset);
//           also how far the match extends backwards (<= 15 bytes).  Results go to a 2-tile ring in smem.
//   decide  every position resolves the lazy (depth 2) choice "if the parser stands here, which match
//           start does it take" from the ring alone -- no dependence on parser state, so it is parallel.
//   chase   one lane follows those decisions from the parser cursor, applies backward extension,
//           substitutes repeat-offset codes, and appends sequences for the block to HBM.
// Matches never cross a 128 KiB block boundary; repeat-offset knowledge is dropped at every block
// start so a block that later falls back to raw cannot desynchronise the decoder's history.
//
// entropy_kernel — one warp per block: gathers literals, then one lane runs the serial block writer
// (zstd_enc_block.h: Huffman literals, FSE sequence tables, bitstreams) into the block's body slot.
#pragma once
#include "common.cuh"
#include "zstd_enc_block.h"

namespace lz {

constexpr uint32_t ROW_LOG = 15, ROWS = 1u << ROW_LOG, ROW_K = 32, TAG_BITS = 10;
constexpr uint32_t TILE = 1024, RING = 2 * TILE, THREADS = 256, PER_THREAD = TILE / THREADS;
constexpr uint32_t MIN_MATCH = 5, SEARCH_CAP = 1024, TARGET_LEN = 64, DEFER = 20, MAX_LAZY_ITERS = 8;
constexpr int32_t ACCEPT_THR = 8;
constexpr uint32_t BLOCKS_PER_CHUNK = 16;
constexpr uint32_t BODY_STRIDE = 2 * Z_BLOCK_MAX;  // per-block body slot: literals + <= 8 bytes per sequence always fit
constexpr uint32_t MAX_SEQ_PER_CHUNK = (2048u * 1024u) / MIN_MATCH + 64;

struct BlockMeta {  // one per (chunk, block), written by lz_kernel, read by entropy_kernel
    uint32_t seq_start, nseq, last_lits, reserved;
};

__device__ __forceinline__ uint64_t ld64_unaligned(const uint8_t *base, uint32_t pos) {
    // reads the aligned 16 bytes around base+pos: callers guarantee pos + 16 <= chunk length
    const uintptr_t a = reinterpret_cast<uintptr_t>(base + pos);
    const uint64_t *w = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t)7);
    const uint32_t sh = (uint32_t)(a & 7u) * 8;
    uint64_t lo = __ldg(w);
    if (sh == 0) return lo;
    uint64_t hi = __ldg(w + 1);
    return (lo >> sh) | (hi << (64 - sh));
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
__device__ __forceinline__ uint32_t hash5(uint64_t v) { return (uint32_t)(((v << 24) * 889523592379ULL) >> (64 - (ROW_LOG + TAG_BITS))); }

__device__ __forceinline__ int32_t sel_score(uint32_t len, uint32_t off) { return (int32_t)(2 * len) - (int32_t)zc::highbit(off + 3); }
__device__ __forceinline__ int32_t lazy_score(uint32_t len, uint32_t off) { return (int32_t)(4 * len) - (int32_t)zc::highbit(off + 3); }

// Candidate the parser may take at position q (ring lookup + block clamp + acceptance rule)
struct Cand { uint32_t len, off; int32_t score; };
__device__ __forceinline__ Cand cand_at(const uint16_t *s_len, const uint32_t *s_off, uint32_t q, uint32_t be) {
    Cand c = {0, 0, 0};
    if (q >= be) return c;
    uint32_t len = s_len[q & (RING - 1)];
    if (!len) return c;
    if (q + len > be) len = be - q;
    if (len < MIN_MATCH) return c;
    const uint32_t off = s_off[q & (RING - 1)];
    const int32_t sc = lazy_score(len, off);
    if (sc < ACCEPT_THR) return c;
    c.len = len; c.off = off; c.score = sc;
    return c;
}

__global__ void __launch_bounds__(THREADS) lz_kernel(const uint8_t *__restrict__ data, const sq_span *__restrict__ spans,
                                                      const uint8_t *__restrict__ select, uint32_t n_chunks, uint32_t *__restrict__ tab_all,
                                                      uint32_t *__restrict__ head_all, zc::Seq *__restrict__ seqs_all,
                                                      BlockMeta *__restrict__ meta_all, uint32_t *__restrict__ counter) {
    __shared__ __align__(16) uint8_t s_in[TILE + 32];
    __shared__ uint16_t s_len[RING];
    __shared__ uint32_t s_off[RING];
    __shared__ uint8_t s_back[RING];
    __shared__ uint8_t s_dec[RING];  // 255 = literal, else delta from p to the chosen match start
    __shared__ uint32_t s_chunk;
    // parser state (owned by thread 0, kept in smem across tiles)
    __shared__ uint32_t s_cursor, s_anchor, s_nseq, s_blk_seq_start, s_rep[3];

    uint32_t *tab = tab_all + (size_t)blockIdx.x * ROWS * ROW_K;
    uint32_t *head = head_all + (size_t)blockIdx.x * ROWS;
    const uint32_t tid = threadIdx.x;

    for (;;) {
        if (tid == 0) {
            uint32_t c;
            do { c = atomicAdd(counter, 1u); } while (c < n_chunks && select && !select[c]);
            s_chunk = c;
        }
        __syncthreads();
        const uint32_t chunk = s_chunk;
        if (chunk >= n_chunks) break;
        const uint8_t *in = data + spans[chunk].off;
        const uint32_t n = spans[chunk].len;
        const uint32_t n_safe = n;  // unaligned 8-byte loads at pos need pos + 16 <= n
        zc::Seq *seqs = seqs_all + (size_t)chunk * MAX_SEQ_PER_CHUNK;
        BlockMeta *meta = meta_all + (size_t)chunk * BLOCKS_PER_CHUNK;
        const bool aligned = (reinterpret_cast<uintptr_t>(in) & 7) == 0;
        if (tid == 0) { s_cursor = 0; s_anchor = 0; s_nseq = 0; s_blk_seq_start = 0; s_rep[0] = 1; s_rep[1] = 4; s_rep[2] = 8; }
        for (uint32_t i = tid; i < RING; i += THREADS) { s_len[i] = 0; s_dec[i] = 255; }
        __syncthreads();

        const uint32_t n_tiles = (n + TILE - 1) / TILE;
        for (uint32_t t = 0; t < n_tiles; t++) {
            const uint32_t t0 = t * TILE, t1 = min(n, t0 + TILE);
            const uint32_t be = min(n, (t0 / Z_BLOCK_MAX + 1) * Z_BLOCK_MAX);  // end of the block this tile lies in
            const bool last_tile_of_block = (t1 == be);
            // ---- stage ----
            for (uint32_t i = tid * 4; i < TILE + 32; i += THREADS * 4) {
                uint32_t w = 0;
                const uint32_t g = t0 + i;
                if (g + 4 <= n && aligned) w = __ldg(reinterpret_cast<const uint32_t *>(in + g));
                else for (uint32_t k = 0; k < 4; k++) if (g + k < n) w |= (uint32_t)in[g + k] << (8 * k);
                *reinterpret_cast<uint32_t *>(&s_in[i]) = w;
            }
            __syncthreads();
            // ---- insert ----
#pragma unroll
            for (uint32_t k = 0; k < PER_THREAD; k++) {
                const uint32_t li = tid + k * THREADS, p = t0 + li;
                const uint32_t hv = hash5(smem_u64(s_in, li));
                if (p + 8 <= n) {
                    const uint32_t row = hv >> TAG_BITS;
                    const uint32_t slot = atomicAdd(&head[row], 1u) & (ROW_K - 1);
                    __stcg(&tab[row * ROW_K + slot], (p + 1) | (hv & ((1u << TAG_BITS) - 1)) << 22);
                }
            }
            __syncthreads();
            // ---- search ----
#pragma unroll 1
            for (uint32_t k = 0; k < PER_THREAD; k++) {
                const uint32_t li = tid + k * THREADS, p = t0 + li;
                uint32_t blen = 0, boff = 0, bback = 0;
                if (p + 8 <= n) {
                    const uint64_t v = smem_u64(s_in, li);
                    const uint32_t hv = hash5(v);
                    const uint32_t row = hv >> TAG_BITS, tag = hv & ((1u << TAG_BITS) - 1);
                    const uint32_t maxlen = min(n - p, SEARCH_CAP);
                    int32_t bscore = -1000;
                    const uint4 *r4 = reinterpret_cast<const uint4 *>(tab + row * ROW_K);
#pragma unroll 1
                    for (uint32_t j = 0; j < ROW_K / 4 && blen < SEARCH_CAP; j++) {
                        const uint4 e4 = __ldcg(r4 + j);
                        const uint32_t e[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            if ((e[q] >> 22) != tag) continue;
                            const uint32_t c = (e[q] & 0x3FFFFFu) - 1u;
                            if (c >= p) continue;
                            // first 8 bytes from registers vs the candidate
                            uint32_t l;
                            if (c + 16 <= n_safe) {
                                const uint64_t x = v ^ ld64_unaligned(in, c);
                                if ((uint32_t)x) continue;  // first 4 bytes differ: tag collision
                                if (x) l = (uint32_t)(__ffsll((long long)x) - 1) / 8;
                                else l = 8 + match_length(in, p + 8, c + 8, maxlen - 8, n_safe);
                            } else {
                                l = match_length(in, p, c, maxlen, n_safe);
                            }
                            if (l < MIN_MATCH) continue;
                            const uint32_t off = p - c;
                            const int32_t sc = sel_score(l, off);
                            if (sc > bscore) { bscore = sc; blen = l; boff = off; }
                        }
                    }
                    if (blen) {
                        const uint32_t c = p - boff;
                        while (bback < 15 && p > bback && c > bback && in[p - bback - 1] == in[c - bback - 1]) bback++;
                    }
                }
                if (p < t1) {
                    s_len[p & (RING - 1)] = (uint16_t)blen;
                    s_off[p & (RING - 1)] = boff;
                    s_back[p & (RING - 1)] = (uint8_t)bback;
                }
            }
            __syncthreads();
            // ---- decide: positions [d0, d1) now have their lookahead window available ----
            const uint32_t d0 = t0 >= DEFER ? t0 - DEFER : 0;
            const uint32_t d1 = last_tile_of_block ? t1 : t1 - DEFER;
            for (uint32_t p = d0 + tid; p < d1; p += THREADS) {
                // positions before t0 that belong to the previous block were already decided there
                if (p < t0 && (p / Z_BLOCK_MAX) != (t0 / Z_BLOCK_MAX)) continue;
                Cand cur = cand_at(s_len, s_off, p, be);
                uint8_t dec = 255;
                if (cur.len) {
                    uint32_t start = p;
                    for (uint32_t it = 0; it < MAX_LAZY_ITERS && cur.len < TARGET_LEN; it++) {
                        const Cand c1 = cand_at(s_len, s_off, start + 1, be);
                        if (c1.len && c1.score > cur.score + 4) { cur = c1; start += 1; continue; }
                        const Cand c2 = cand_at(s_len, s_off, start + 2, be);
                        if (c2.len && c2.score > cur.score + 7) { cur = c2; start += 2; continue; }
                        break;
                    }
                    dec = (uint8_t)(start - p);
                }
                s_dec[p & (RING - 1)] = dec;
            }
            __syncthreads();
            // ---- chase ----
            if (tid == 0) {
                uint32_t p = s_cursor, anchor = s_anchor, nseq = s_nseq;
                uint32_t r0 = s_rep[0], r1 = s_rep[1], r2 = s_rep[2];
                while (p < d1) {
                    const uint32_t d = s_dec[p & (RING - 1)];
                    if (d == 255) { p++; continue; }
                    uint32_t start = p + d;
                    uint32_t len = s_len[start & (RING - 1)];
                    const uint32_t off = s_off[start & (RING - 1)];
                    if (start + len > be) len = be - start;
                    if (len >= SEARCH_CAP && start + len < be) {  // capped by the search: extend (rare; long runs)
                        len += match_length(in, start + len, start + len - off, be - start - len, n_safe);
                    }
                    uint32_t k = s_back[start & (RING - 1)];
                    if (k > start - anchor) k = start - anchor;
                    start -= k; len += k;
                    const uint32_t ll = start - anchor;
                    // repeat-offset code substitution (RFC 8878 3.1.1.5)
                    uint32_t ob = off + 3;
                    if (ll) { if (off == r0) ob = 1; else if (off == r1) ob = 2; else if (off == r2) ob = 3; }
                    else { if (off == r1) ob = 1; else if (off == r2) ob = 2; else if (r0 > 1 && off == r0 - 1) ob = 3; }
                    if (ob > 3) { r2 = r1; r1 = r0; r0 = off; }
                    else {
                        const uint32_t idx = ob - 1 + (ll ? 0 : 1);
                        if (idx == 1) { const uint32_t tmp = r1; r1 = r0; r0 = tmp; }
                        else if (idx == 2) { const uint32_t tmp = r2; r2 = r1; r1 = r0; r0 = tmp; }
                        else if (idx == 3) { const uint32_t tmp = r0 - 1; r2 = r1; r1 = r0; r0 = tmp; }
                    }
                    if (nseq < MAX_SEQ_PER_CHUNK) {
                        zc::Seq sq; sq.ll = ll; sq.ml = len; sq.off_base = ob;
                        seqs[nseq] = sq;
                    }
                    nseq++;
                    p = start + len;
                    anchor = p;
                }
                if (last_tile_of_block) {  // close the block
                    const uint32_t b = t0 / Z_BLOCK_MAX;
                    BlockMeta m;
                    m.seq_start = s_blk_seq_start; m.nseq = min(nseq, MAX_SEQ_PER_CHUNK) - s_blk_seq_start; m.last_lits = be - anchor; m.reserved = 0;
                    meta[b] = m;
                    s_blk_seq_start = min(nseq, MAX_SEQ_PER_CHUNK);
                    anchor = be; p = be;
                    r0 = r1 = r2 = 0;  // forget repeat offsets: the next block must not depend on this one being emitted compressed
                }
                s_cursor = p; s_anchor = anchor; s_nseq = nseq; s_rep[0] = r0; s_rep[1] = r1; s_rep[2] = r2;
            }
            __syncthreads();
        }
    }
}

// ---- entropy stage: one warp per block ------------------------------------------------------------
struct BlockOut { uint32_t body_len, type; };  // mirrors sq_block_info

__global__ void __launch_bounds__(128) entropy_kernel(const uint8_t *__restrict__ data, const sq_span *__restrict__ spans,
                                                       const uint8_t *__restrict__ select, uint32_t n_chunks,
                                                       const zc::Seq *__restrict__ seqs_all, const BlockMeta *__restrict__ meta_all,
                                                       uint8_t *__restrict__ lit_all, uint8_t *__restrict__ bodies, BlockOut *__restrict__ blocks,
                                                       zc::EncWork *__restrict__ work_all, uint32_t *__restrict__ counter) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    zc::EncWork *wk = work_all + warp_global;
    uint8_t *lits = lit_all + (size_t)warp_global * (Z_BLOCK_MAX + 64);
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
        // gather literals: lanes stride over sequences; source positions from a running prefix kept by lane order
        // (serial prefix over sequences by chunks of 32 with warp scan)
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
        BlockOut out;
        out.type = 0; out.body_len = blen;  // raw unless the compressed body is smaller
        if (lane == 0 && src_pos + m.last_lits == be && blen > 0) {
            uint8_t *dst = bodies + ((size_t)chunk * BLOCKS_PER_CHUNK + b) * (size_t)BODY_STRIDE;
            // worst case body is literals + sequences; both fit the slot only if the parse actually saves bytes,
            // so bound the attempt: sequences cost <= 8 bytes each
            if (zc::block_body_bound(nlits, m.nseq) <= BODY_STRIDE) {
                const uint32_t sz = zc::write_block_body(dst, lits, nlits, seqs, m.nseq, wk);
                if (sz < blen) { out.type = 2; out.body_len = sz; }
            }
        }
        if (lane == 0) blocks[(size_t)chunk * BLOCKS_PER_CHUNK + b] = out;
        __syncwarp();
    }
}

}  // namespace lz
