// K3 parse of one 128 KiB block from the search kernel's per-position records: lazy decision (depth 2, zstd's gains), repeat
// offsets, sequence emission.  Scalar and __host__ __device__: ONE GPU THREAD walks one block (lz2::chase_kernel), and the CPU
// model (tests/harness/lz_model2.cc) runs the very same function, so the parse the tests study is the parse the GPU ships.
//
// Why a thread and not a warp: the walk is a chain of data-dependent scalar decisions; a warp spent 32 lanes on it (the first
// round-2 version: 95 M warp instructions per chunk, a third of the search kernel's time).  With a thread per block the same
// instructions serve 32 blocks at once.
//
// Search record (one u32 per position, 0 = no match): bits 0-20 offset, 21-25 verified length - 4, 26 "may be longer",
// 27-28 backward extension available (<= 3).
// The shortest match the search reports is 5 bytes (4 in chunks of at most 128 KiB, like libzstd's parameters for small inputs).
#pragma once
#include "zstd_enc_block.h"
#if defined(__CUDACC__)
#define ZMEM __host__ __device__ __forceinline__
#else
#define ZMEM inline
#endif

namespace zparse {

constexpr uint32_t LEN_BASE = 4, CAP = 32, MAX_SHIFT = 7, REP_MIN = 3, POS_MASK = (1u << 21) - 1;  // record length field = length - LEN_BASE
constexpr int32_t ACCEPT_THR = 6;

ZHD uint32_t ld32(const uint8_t *in, uint32_t pos, uint32_t n) {  // 4 bytes at any alignment; bytes past n read as 0
#if defined(__CUDA_ARCH__)
    if (pos + 8 <= n) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(in + pos);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        return __funnelshift_r(__ldg(w), __ldg(w + 1), (uint32_t)(a & 3u) * 8);
    }
#else
    if (pos + 4 <= n) { uint32_t v; memcpy(&v, in + pos, 4); return v; }
#endif
    uint32_t v = 0;
    for (uint32_t k = 0; k < 4 && pos + k < n; k++) v |= (uint32_t)in[pos + k] << (8 * k);
    return v;
}

// common prefix of in[a..] and in[b..] (b < a), at most maxlen (a + maxlen <= n)
ZHD uint32_t common_len(const uint8_t *in, uint32_t n, uint32_t a, uint32_t b, uint32_t maxlen) {
    uint32_t l = 0;
    while (l + 4 <= maxlen) {
        const uint32_t x = ld32(in, a + l, n) ^ ld32(in, b + l, n);
        if (x) {
#if defined(__CUDA_ARCH__)
            return l + ((uint32_t)(__ffs((int)x) - 1) >> 3);
#else
            return l + ((uint32_t)__builtin_ctz(x) >> 3);
#endif
        }
        l += 4;
    }
    while (l < maxlen && in[a + l] == in[b + l]) l++;
    return l;
}

ZHD int32_t rec_score(uint32_t r) {  // what the parser gains by taking this record's match; -1 = unusable
    if (!r) return -1;
    const uint32_t sl = (r >> 26 & 1u) ? CAP : ((r >> 21) & 31u) + LEN_BASE;
    const int32_t sc = (int32_t)(4 * sl) - (int32_t)zc::highbit((r & POS_MASK) + 3);
    return sc >= ACCEPT_THR ? sc : -1;
}

struct Parser {
    const uint8_t *in; uint32_t n, bs, be;
    zc::Seq *seqs; uint32_t cap, nseq;
    uint32_t p, anchor, r0, r1, r2;

    ZMEM void emit(uint32_t start, uint32_t len, uint32_t off) {
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
        zc::Seq sq; sq.ll = ll; sq.ml = len; sq.off_base = ob;
        seqs[nseq++] = sq;
        p = start + len;
        anchor = p;
    }
    // longest repeat-offset match at g (>= 3 bytes), 0 if none; which offsets are cheap depends on whether g follows a match directly
    ZMEM uint32_t rep_at(uint32_t g, uint32_t *off_out) const {
        *off_out = 0;
        if (g + 4 > be || !(r0 | r1 | r2)) return 0;
        const bool ll0 = g == anchor;
        const uint32_t c[3] = {ll0 ? r1 : r0, ll0 ? r2 : r1, ll0 ? (r0 > 1 ? r0 - 1 : 0u) : r2};
        const uint32_t v = ld32(in, g, n);
        uint32_t bl = 0;
        for (int i = 0; i < 3; i++) {
            const uint32_t o = c[i];
            if (!o || o > g) continue;
            if ((ld32(in, g - o, n) ^ v) & 0xFFFFFFu) continue;
            const uint32_t l = 3 + common_len(in, n, g + 3, g + 3 - o, be - g - 3);
            if (l > bl) { bl = l; *off_out = o; }
        }
        return bl;
    }
};

// Walks positions [bs, be) of one block.  Returns the number of sequences; *last_lits = literals after the last sequence.
ZHDN uint32_t chase_block(const uint8_t *in, uint32_t n, const uint32_t *rec, uint32_t bs, uint32_t be, bool first_block,
                          zc::Seq *seqs, uint32_t cap, uint32_t *last_lits) {
    Parser P;
    P.in = in; P.n = n; P.bs = bs; P.be = be; P.seqs = seqs; P.cap = cap; P.nseq = 0; P.p = bs; P.anchor = bs;
    // repeat offsets are unknown at a block start (a block that falls back to raw must not desynchronise the decoder's
    // history); the frame's first block knows 1, 4, 8
    P.r0 = first_block ? 1 : 0; P.r1 = first_block ? 4 : 0; P.r2 = first_block ? 8 : 0;
    while (P.p < be && P.nseq < cap) {
        const uint32_t q = P.p;
#if defined(__CUDA_ARCH__)
        uint32_t r = __ldg(rec + q);
#else
        uint32_t r = rec[q];
#endif
        int32_t cur = rec_score(r);
        uint32_t ro;
        if (cur < 0) {  // a literal, unless a repeat offset matches here
            const uint32_t rl = P.rep_at(q, &ro);
            if (rl >= REP_MIN) P.emit(q, rl, ro); else P.p = q + 1;
            continue;
        }
        // lazy choice: a better match one or two positions later wins (zstd's gains: +4 / +7 in units of a quarter byte)
        uint32_t start = q;
        while (!(r >> 26 & 1u) && start - q + 2 <= MAX_SHIFT) {
#if defined(__CUDA_ARCH__)
            const uint32_t ra = start + 1 < be ? __ldg(rec + start + 1) : 0u, rb = start + 2 < be ? __ldg(rec + start + 2) : 0u;
#else
            const uint32_t ra = start + 1 < be ? rec[start + 1] : 0u, rb = start + 2 < be ? rec[start + 2] : 0u;
#endif
            const int32_t s1 = rec_score(ra);
            if (s1 > cur + 4) { cur = s1; start += 1; r = ra; continue; }
            const int32_t s2 = rec_score(rb);
            if (s2 > cur + 7) { cur = s2; start += 2; r = rb; continue; }
            break;
        }
        const uint32_t off = r & POS_MASK;
        uint32_t len = ((r >> 21) & 31u) + LEN_BASE, back = (r >> 27) & 3u;
        if ((r >> 26 & 1u) && start + len < be) len += common_len(in, n, start + len, start + len - off, be - start - len);
        if (start + len > be) len = be - start;  // the range may end inside a block (sub-block parsing): a match never crosses it
        if (len < REP_MIN) { P.p = q + 1; continue; }
        if (back > start - P.anchor) back = start - P.anchor;
        start -= back; len += back;
        // literal positions in front of the match: the first one where a repeat offset matches takes over
        bool took = false;
        for (uint32_t g = q; g < start; g++) {
            const uint32_t rl = P.rep_at(g, &ro);
            if (rl >= REP_MIN) { P.emit(g, rl, ro); took = true; break; }
        }
        if (took) continue;
        {   // at the match start a repeat offset wins when zstd's rule of thumb says so: 3 rl > 3 len - log2(offset) + 1
            const uint32_t rl = P.rep_at(start, &ro);
            if (rl >= 3 && ro != off && (int32_t)(3 * rl) > (int32_t)(3 * len) - (int32_t)zc::highbit(off + 3) + 1) { P.emit(start, rl, ro); continue; }
        }
        P.emit(start, len, off);
    }
    *last_lits = be - P.anchor;
    return P.nseq;
}

}  // namespace zparse
