// K3 stage "compress", shared pieces: byte-access helpers, per-block metadata, warp-wide match extension, and the entropy
// stage (entropy_kernel: one warp per 128 KiB block gathers literals and writes the block body warp-parallel, see
// zstd_enc_entropy.cuh).  The match search (index_kernel + search_kernel: the chunk's positions sorted by hash row, then a search
// without any synchronisation) and the parse (chase_kernel) live in zstd_enc_lz2.cuh.
#pragma once
#include "common.cuh"
#include "zstd_enc_block.h"
#include "zstd_enc_entropy.cuh"

namespace lz {

constexpr uint32_t BLOCKS_PER_CHUNK = 16;
constexpr uint32_t BODY_STRIDE = 2 * Z_BLOCK_MAX;  // per-block body slot: literals + <= 8 bytes per sequence always fit
// sequence slots per block: a block that parses into more sequences than this keeps the rest as literals (the parse stops emitting)
constexpr uint32_t SEQ_PER_BLOCK = Z_BLOCK_MAX / 6 + 8, MAX_SEQ_PER_CHUNK = BLOCKS_PER_CHUNK * SEQ_PER_BLOCK;
constexpr uint32_t REC_PER_CHUNK = 2048u * 1024u;  // per-position search records of one chunk
// per entropy warp: 3 x SEQ_PER_BLOCK words of FSE state-transition records, SEQ_PER_BLOCK words of packed symbol codes, then three
// byte arrays of symbol codes (stride SEQ_CODE_STRIDE)
constexpr uint32_t SEQ_CODE_STRIDE = (SEQ_PER_BLOCK + 15u) & ~15u, SBITS_STRIDE = (4 * SEQ_PER_BLOCK + 3 * SEQ_CODE_STRIDE / 4 + 3u) & ~3u;

struct BlockMeta {  // one per (chunk, block), written by the chase kernel, read by entropy_kernel
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
// one 4-byte word of a chunk for the shared-memory stage buffer (bytes past the chunk read as 0)
__device__ __forceinline__ uint32_t stage_word(const uint8_t *__restrict__ in, uint32_t g, uint32_t n, bool aligned) {
    if (g + 4 <= n && aligned) return __ldg(reinterpret_cast<const uint32_t *>(in + g));
    uint32_t w = 0;
    for (uint32_t k = 0; k < 4; k++) if (g + k < n) w |= (uint32_t)in[g + k] << (8 * k);
    return w;
}

// ---- warp-wide match extension (used by the chase, zstd_enc_lz2.cuh) -------------------------------------
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
