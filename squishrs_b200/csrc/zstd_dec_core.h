// zstd frame decoder core (K4), written for warp-uniform execution: on the GPU all 32 lanes of a warp
// run the same control flow over one frame (table parsing and the serial FSE sequence chain are
// computed redundantly and identically by every lane -- free in SIMT), and the lanes split only where
// the format allows: the 4 Huffman literal streams and the byte copies of literal runs and matches.
// On the host the same code compiles with a "warp" of one lane, which is how tests/ check it against
// frames written by stock libzstd.  Format: RFC 8878.
#pragma once
#include "zstd_core.h"

namespace zd {

#if defined(__CUDA_ARCH__)
#define ZD_LANE() (threadIdx.x & 31u)
#define ZD_WARP 32u
#define ZD_SYNC() __syncwarp()
#else
#define ZD_LANE() 0u
#define ZD_WARP 1u
#define ZD_SYNC() ((void)0)
#endif
#if defined(__CUDACC__)
#define ZD_DEV __host__ __device__ inline
#else
#define ZD_DEV static inline
#endif

struct SeqCell {  // one FSE state of a sequence table (4 bytes: the per-warp tables must stay small, they bound residency);
    uint16_t next_base;   // base value / extra-bit count of the code come from the constant code tables
    uint8_t nb_bits;      // state bits to read for the transition
    uint8_t sym;          // the LL / ML / OF code
};

struct Tables {          // per-warp state that persists across the blocks of a frame (treeless / repeat modes)
    uint16_t huf[1u << Z_HUF_MAXBITS];  // symbol | nb_bits << 8
    SeqCell ll[1u << Z_LL_MAXLOG], ml[1u << Z_ML_MAXLOG], of[1u << Z_OF_MAXLOG];
    uint32_t huf_log, ll_log, ml_log, of_log;
    uint32_t have_huf, have_ll, have_ml, have_of;
};

enum { ERR_CORRUPT = -1, ERR_CAPACITY = -2 };

// ---- loads ------------------------------------------------------------------------------------------
ZD_DEV uint64_t load64(const uint8_t *p) {  // unaligned little-endian 8 bytes; caller guarantees p-7 .. p+15 readable
#if defined(__CUDA_ARCH__)
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint64_t *w = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t)7);
    const uint32_t sh = (uint32_t)(a & 7u) * 8;
    const uint64_t lo = w[0], hi = w[1];
    return (lo >> sh) | ((hi << 1) << (63 - sh));
#else
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
#endif
}

// Backward bit reader over [start, start+size): bits are consumed from the top (just below the sentinel).
// A 64-bit window is cached in registers and refilled (one unaligned 8-byte load) only when it runs low.
// Positions are 32-bit: a stream is at most a block (< 2^20 bits).
struct BitReader {
    const uint8_t *start;
    int32_t pos;    // number of unread bits; goes negative on corrupt input (reads then return zeros)
    int32_t wlo;    // bit index of win's bit 0 (multiple of 8; negative near the stream start)
    uint64_t win;   // bits [wlo, wlo+64) of the stream, zero where that lies before the stream
};
ZD_DEV void br_refill(BitReader *b) {
    if (b->pos <= 0) { b->win = 0; b->wlo = b->pos - 64; return; }  // exhausted / overrun: zeros
    // place the window so that pos - wlo is in (56, 64]
    const int32_t lo = ((b->pos + 7) >> 3) * 8 - 64;
    b->wlo = lo;
    if (lo >= 0) { b->win = load64(b->start + (lo >> 3)); return; }
    // within 8 bytes of the stream start: never read before it; the missing low bits are zero
    const uint32_t sh = (uint32_t)(-lo);  // multiple of 8, 8..56
    b->win = load64(b->start) << sh;
}
ZD_DEV int br_init(BitReader *b, const uint8_t *start, uint32_t size) {
    if (size == 0) return ERR_CORRUPT;
    const uint8_t last = start[size - 1];
    if (last == 0) return ERR_CORRUPT;
    b->start = start;
    b->pos = (int32_t)((size - 1) * 8 + zc::highbit(last));
    br_refill(b);
    return 0;
}
ZD_DEV uint32_t br_peek(BitReader *b, uint32_t n) {  // n <= 31; bits [pos-n, pos), zero-filled below the stream start
    const int32_t lo = b->pos - (int32_t)n;
    if (lo < b->wlo) br_refill(b);
    const uint32_t sh = (uint32_t)(lo - b->wlo) & 63u;
    return (uint32_t)(b->win >> sh) & ((1u << n) - 1u);
}
ZD_DEV uint32_t br_read(BitReader *b, uint32_t n) {
    const uint32_t v = br_peek(b, n);
    b->pos -= (int32_t)n;
    return v;
}

// Unchecked read for a caller that has just refilled: n <= 31 and the window still covers [pos - n, pos).  After br_refill
// the window holds more than 56 unread bits (or the stream is exhausted and it holds zeros), so a caller may take up to 56
// bits before the next refill.  Once pos has gone negative (corrupt input) the values are meaningless but the shift stays in
// range; callers check pos < 0 afterwards.
ZD_DEV uint32_t br_take(BitReader *b, uint32_t n) {
    b->pos -= (int32_t)n;
    const uint32_t sh = (uint32_t)(b->pos - b->wlo) & 63u;
    return (uint32_t)(b->win >> sh) & ((1u << n) - 1u);
}

// Forward bit reader (FSE table descriptions)
struct FwdReader { const uint8_t *p; uint32_t size; uint32_t bit; };
ZD_DEV uint32_t fr_peek(const FwdReader *r, uint32_t n) {
    const uint32_t byte = r->bit >> 3;
    uint64_t v = 0;
    for (uint32_t i = 0; i < 5 && byte + i < r->size; i++) v |= (uint64_t)r->p[byte + i] << (8 * i);
    return (uint32_t)(v >> (r->bit & 7)) & ((1u << n) - 1);
}

// FSE table description -> normalised counts.  Returns bytes consumed or < 0.
ZD_DEV int read_ncount(const uint8_t *src, uint32_t size, int16_t *norm, uint32_t max_sym_cap, uint32_t max_log, uint32_t *max_sym, uint32_t *tl) {
    if (size < 1) return ERR_CORRUPT;
    FwdReader r = {src, size, 0};
    const uint32_t al = fr_peek(&r, 4) + 5;
    r.bit += 4;
    if (al > max_log) return ERR_CORRUPT;
    int32_t remaining = 1 << al;
    uint32_t s = 0;
    while (remaining > 0 && s <= max_sym_cap) {
        if ((r.bit >> 3) >= size) return ERR_CORRUPT;
        const uint32_t bits = zc::highbit((uint32_t)remaining + 1) + 1;
        uint32_t v = fr_peek(&r, bits);
        const uint32_t lower = (1u << (bits - 1)) - 1;
        const uint32_t thresh = (1u << bits) - 1 - ((uint32_t)remaining + 1);
        if ((v & lower) < thresh) { r.bit += bits - 1; v &= lower; }
        else { r.bit += bits; if (v > lower) v -= thresh; }
        const int32_t prob = (int32_t)v - 1;
        remaining -= prob < 0 ? 1 : prob;
        norm[s++] = (int16_t)prob;
        if (prob == 0) {
            for (;;) {
                if ((r.bit >> 3) >= size) return ERR_CORRUPT;
                const uint32_t rep = fr_peek(&r, 2);
                r.bit += 2;
                for (uint32_t i = 0; i < rep && s <= max_sym_cap; i++) norm[s++] = 0;
                if (rep != 3) break;
            }
        }
    }
    if (remaining != 0 || s == 0) return ERR_CORRUPT;
    *max_sym = s - 1;
    *tl = al;
    const uint32_t used = (r.bit + 7) >> 3;
    return used > size ? ERR_CORRUPT : (int)used;
}

// ---- Huffman literals ---------------------------------------------------------------------------------
// Reads the tree description at src; fills T->huf.  Returns bytes consumed or < 0.  (Executed uniformly.)
ZD_DEV int read_huf_tree(const uint8_t *src, uint32_t size, Tables *T, uint8_t *weights /* 256 */, zc::FseDCell *wcells /* 64 */) {
    if (size < 1) return ERR_CORRUPT;
    const uint32_t hb = src[0];
    uint32_t nw = 0, used;
    if (hb >= 128) {  // direct 4-bit weights
        nw = hb - 127;
        used = 1 + (nw + 1) / 2;
        if (used > size) return ERR_CORRUPT;
        for (uint32_t i = 0; i < nw; i++) weights[i] = (i & 1) ? (src[1 + i / 2] & 15) : (src[1 + i / 2] >> 4);
    } else {  // FSE-compressed weights: table description (max log 6) then two interleaved states
        used = 1 + hb;
        if (hb == 0 || used > size) return ERR_CORRUPT;
        int16_t norm[16];
        uint32_t max_sym = 0, tl = 0;
        const int hs = read_ncount(src + 1, hb, norm, 12, 6, &max_sym, &tl);
        if (hs < 0 || (uint32_t)hs >= hb) return ERR_CORRUPT;
        uint16_t next[16];
        zc::fse_build_dtable(wcells, norm, max_sym, tl, next);
        BitReader b;
        if (br_init(&b, src + 1 + hs, hb - hs) < 0) return ERR_CORRUPT;
        uint32_t s1 = br_read(&b, tl), s2 = br_read(&b, tl);
        for (;;) {
            if (nw >= 254) return ERR_CORRUPT;
            weights[nw++] = wcells[s1].sym;
            { const uint32_t nb = wcells[s1].nb_bits; if (b.pos < (int32_t)nb) { weights[nw++] = wcells[s2].sym; break; } s1 = wcells[s1].base + br_read(&b, nb); }
            if (nw >= 254) return ERR_CORRUPT;
            weights[nw++] = wcells[s2].sym;
            { const uint32_t nb = wcells[s2].nb_bits; if (b.pos < (int32_t)nb) { weights[nw++] = wcells[s1].sym; break; } s2 = wcells[s2].base + br_read(&b, nb); }
        }
    }
    if (nw == 0 || nw > 255) return ERR_CORRUPT;
    // implied last weight completes the sum of 2^(w-1) to a power of two
    uint32_t sum = 0;
    for (uint32_t i = 0; i < nw; i++) { if (weights[i] > Z_HUF_MAXBITS) return ERR_CORRUPT; if (weights[i]) sum += 1u << (weights[i] - 1); }
    if (sum == 0) return ERR_CORRUPT;
    const uint32_t max_bits = zc::highbit(sum) + 1;
    if (max_bits > Z_HUF_MAXBITS) return ERR_CORRUPT;
    const uint32_t rest = (1u << max_bits) - sum;
    if (rest == 0 || (rest & (rest - 1))) return ERR_CORRUPT;
    weights[nw] = (uint8_t)(zc::highbit(rest) + 1);
    const uint32_t nsym = nw + 1;
    // decode table: weights ascending, symbols ascending, 2^(w-1) consecutive cells each
    uint32_t idx = 0;
    for (uint32_t w = 1; w <= max_bits; w++) {
        const uint32_t span = 1u << (w - 1), nb = max_bits + 1 - w;
        for (uint32_t s = 0; s < nsym; s++) {
            if (weights[s] != w) continue;
            const uint16_t cell = (uint16_t)(s | nb << 8);
            for (uint32_t k = ZD_LANE(); k < span; k += ZD_WARP) T->huf[idx + k] = cell;  // lanes split the fill
            idx += span;
        }
    }
    ZD_SYNC();
    T->huf_log = max_bits;
    T->have_huf = 1;
    return (int)used;
}

// one Huffman stream -> n symbols at out (executed by ONE lane)
ZD_DEV int huf_decode_stream(const uint8_t *src, uint32_t size, uint8_t *out, uint32_t n, const Tables *T) {
    BitReader b;
    if (br_init(&b, src, size) < 0) return ERR_CORRUPT;
    const uint32_t log = T->huf_log;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t cell = T->huf[br_peek(&b, log)];
        out[i] = (uint8_t)cell;
        b.pos -= cell >> 8;
    }
    return b.pos == 0 ? 0 : ERR_CORRUPT;
}

struct Literals { const uint8_t *ptr; uint32_t size; uint32_t rle; uint8_t rle_byte; };

// Literals section.  Returns bytes consumed or < 0.  Decoded Huffman literals land in litbuf (>= 128 KiB + 32).
ZD_DEV int decode_literals(const uint8_t *src, uint32_t size, Tables *T, uint8_t *litbuf, Literals *L, uint8_t *weights, zc::FseDCell *wcells) {
    if (size < 1) return ERR_CORRUPT;
    const uint32_t type = src[0] & 3, sf = (src[0] >> 2) & 3;
    if (type < 2) {  // Raw / RLE
        uint32_t hl, n;
        if (sf == 0 || sf == 2) { hl = 1; n = src[0] >> 3; }
        else if (sf == 1) { if (size < 2) return ERR_CORRUPT; hl = 2; n = (src[0] | src[1] << 8) >> 4; }
        else { if (size < 3) return ERR_CORRUPT; hl = 3; n = (src[0] | src[1] << 8 | (uint32_t)src[2] << 16) >> 4; }
        if (n > Z_BLOCK_MAX) return ERR_CORRUPT;
        if (type == 0) {
            if (hl + n > size) return ERR_CORRUPT;
            L->ptr = src + hl; L->size = n; L->rle = 0;
            return (int)(hl + n);
        }
        if (hl + 1 > size) return ERR_CORRUPT;
        L->ptr = src + hl; L->size = n; L->rle = 1; L->rle_byte = src[hl];
        return (int)(hl + 1);
    }
    // Compressed (2) / Treeless (3)
    uint32_t hl, regen, comp, streams;
    if (sf <= 1) { if (size < 3) return ERR_CORRUPT; hl = 3; const uint32_t v = src[0] | src[1] << 8 | (uint32_t)src[2] << 16; regen = (v >> 4) & 0x3FF; comp = v >> 14; streams = sf == 0 ? 1 : 4; }
    else if (sf == 2) { if (size < 4) return ERR_CORRUPT; hl = 4; const uint32_t v = src[0] | src[1] << 8 | (uint32_t)src[2] << 16 | (uint32_t)src[3] << 24; regen = (v >> 4) & 0x3FFF; comp = v >> 18; streams = 4; }
    else { if (size < 5) return ERR_CORRUPT; hl = 5; const uint64_t v = src[0] | src[1] << 8 | (uint32_t)src[2] << 16 | (uint64_t)src[3] << 24 | (uint64_t)src[4] << 32; regen = (uint32_t)(v >> 4) & 0x3FFFF; comp = (uint32_t)(v >> 22); streams = 4; }
    if (regen > Z_BLOCK_MAX || hl + comp > size || regen == 0) return ERR_CORRUPT;
    const uint8_t *p = src + hl;
    uint32_t left = comp;
    if (type == 2) {
        const int t = read_huf_tree(p, left, T, weights, wcells);
        if (t < 0) return t;
        p += t; left -= (uint32_t)t;
    } else if (!T->have_huf) return ERR_CORRUPT;
    int err = 0;
    if (streams == 1) {
        if (ZD_LANE() == 0) err = huf_decode_stream(p, left, litbuf, regen, T);
    } else {
        if (left < 6) return ERR_CORRUPT;
        const uint32_t s1 = p[0] | p[1] << 8, s2 = p[2] | p[3] << 8, s3 = p[4] | p[5] << 8;
        if (6 + s1 + s2 + s3 > left) return ERR_CORRUPT;  // stream 4 may not be empty
        const uint32_t s4 = left - 6 - s1 - s2 - s3;
        const uint32_t seg = (regen + 3) / 4;
        if (3 * seg > regen) return ERR_CORRUPT;
        const uint32_t off[4] = {6, 6 + s1, 6 + s1 + s2, 6 + s1 + s2 + s3}, sz[4] = {s1, s2, s3, s4};
#if defined(__CUDA_ARCH__)
        const uint32_t k = ZD_LANE();
        if (k < 4) err = huf_decode_stream(p + off[k], sz[k], litbuf + k * seg, k < 3 ? seg : regen - 3 * seg, T);
#else
        for (uint32_t k = 0; k < 4 && !err; k++) err = huf_decode_stream(p + off[k], sz[k], litbuf + k * seg, k < 3 ? seg : regen - 3 * seg, T);
#endif
    }
#if defined(__CUDA_ARCH__)
    err = __any_sync(0xffffffffu, err != 0) ? ERR_CORRUPT : 0;
#endif
    ZD_SYNC();
    if (err) return ERR_CORRUPT;
    L->ptr = litbuf; L->size = regen; L->rle = 0;
    return (int)(hl + comp);
}

// ---- sequence tables -------------------------------------------------------------------------------------
enum { KIND_LL = 0, KIND_OF = 1, KIND_ML = 2 };
ZD_DEV void fill_cell(SeqCell *c, uint32_t kind, uint32_t sym, uint32_t nb, uint32_t base) {
    (void)kind;
    c->next_base = (uint16_t)base;
    c->nb_bits = (uint8_t)nb;
    c->sym = (uint8_t)sym;
}
// builds cells from normalised counts (uniform: every lane executes it; lanes write identical values)
ZD_DEV int build_seq_table(SeqCell *cells, uint32_t kind, const int16_t *norm, uint32_t max_sym, uint32_t tl, zc::FseDCell *tmp /* 512 */) {
    const uint32_t sym_cap = kind == KIND_LL ? 35 : kind == KIND_ML ? 52 : 31;
    if (max_sym > sym_cap) return ERR_CORRUPT;
    uint16_t next[64];
    zc::fse_build_dtable(tmp, norm, max_sym, tl, next);
    const uint32_t size = 1u << tl;
    for (uint32_t u = 0; u < size; u++) fill_cell(&cells[u], kind, tmp[u].sym, tmp[u].nb_bits, tmp[u].base);
    return 0;
}
// one of LL / OF / ML per its 2-bit mode.  Returns bytes consumed or < 0.
ZD_DEV int read_seq_table(const uint8_t *src, uint32_t size, uint32_t mode, uint32_t kind, SeqCell *cells, uint32_t *log, uint32_t *have,
                          zc::FseDCell *tmp) {
    const uint32_t max_log = kind == KIND_LL ? Z_LL_MAXLOG : kind == KIND_ML ? Z_ML_MAXLOG : Z_OF_MAXLOG;
    if (mode == 0) {  // Predefined
        int16_t dn[64];
        uint32_t ms, tl;
        if (kind == KIND_LL) { ms = 35; tl = 6; for (uint32_t i = 0; i <= ms; i++) dn[i] = zc::ZTAB(LL_defnorm)[i]; }
        else if (kind == KIND_ML) { ms = 52; tl = 6; for (uint32_t i = 0; i <= ms; i++) dn[i] = zc::ZTAB(ML_defnorm)[i]; }
        else { ms = 28; tl = 5; for (uint32_t i = 0; i <= ms; i++) dn[i] = zc::ZTAB(OF_defnorm)[i]; }
        if (build_seq_table(cells, kind, dn, ms, tl, tmp) < 0) return ERR_CORRUPT;
        *log = tl; *have = 1;
        return 0;
    }
    if (mode == 1) {  // RLE: a single symbol, zero state bits
        if (size < 1) return ERR_CORRUPT;
        const uint32_t sym = src[0];
        const uint32_t sym_cap = kind == KIND_LL ? 35 : kind == KIND_ML ? 52 : 31;
        if (sym > sym_cap) return ERR_CORRUPT;
        fill_cell(&cells[0], kind, sym, 0, 0);
        *log = 0; *have = 1;
        return 1;
    }
    if (mode == 2) {  // FSE_Compressed
        int16_t norm[64];
        uint32_t max_sym = 0, tl = 0;
        const int used = read_ncount(src, size, norm, kind == KIND_LL ? 35 : kind == KIND_ML ? 52 : 31, max_log, &max_sym, &tl);
        if (used < 0) return used;
        if (build_seq_table(cells, kind, norm, max_sym, tl, tmp) < 0) return ERR_CORRUPT;
        *log = tl; *have = 1;
        return used;
    }
    return *have ? 0 : ERR_CORRUPT;  // Repeat
}

// ---- byte movers -------------------------------------------------------------------------------------------
// one lane moves n bytes (non-overlapping).  On the device 16 bytes travel per round as five aligned 32-bit loads (issued
// together, so a round costs one memory round trip) and funnel shifts; the byte-wise version needed two rounds of eight 1-byte
// loads for the same 16 bytes, and the decoder's copies are latency-bound (ncu: 60 % of the match pass's samples waited on them).
// Whole aligned words are read around the source: up to 3 bytes before it and after its end, never past the next multiple of 4
// (the frame's output starts 16-byte aligned and output buffers end on a multiple of 4 or carry slack).
ZD_DEV void lane_copy(uint8_t *dst, const uint8_t *src, uint32_t n) {
#if defined(__CUDA_ARCH__)
    const uintptr_t sa = reinterpret_cast<uintptr_t>(src);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(sa & ~(uintptr_t)3);
    const uint32_t lead = (uint32_t)(sa & 3u), sh = lead * 8;
    for (uint32_t t = 0; t < n; t += 16, w += 4) {
        const uint32_t m = n - t < 16u ? n - t : 16u, need = (m + lead + 3u) >> 2;  // bytes this round, words that hold them (1..5)
        const uint32_t a0 = w[0], a1 = need > 1 ? w[1] : 0u, a2 = need > 2 ? w[2] : 0u, a3 = need > 3 ? w[3] : 0u, a4 = need > 4 ? w[4] : 0u;
        const uint32_t o[4] = {__funnelshift_r(a0, a1, sh), __funnelshift_r(a1, a2, sh), __funnelshift_r(a2, a3, sh), __funnelshift_r(a3, a4, sh)};
        uint8_t *d = dst + t;
        if (m == 16 && (reinterpret_cast<uintptr_t>(d) & 3u) == 0) {
            uint32_t *dw = reinterpret_cast<uint32_t *>(d);
            dw[0] = o[0]; dw[1] = o[1]; dw[2] = o[2]; dw[3] = o[3];
        } else {
#pragma unroll
            for (uint32_t k = 0; k < 16; k++) if (k < m) d[k] = (uint8_t)(o[k >> 2] >> (8 * (k & 3)));
        }
    }
#else
    for (uint32_t t = 0; t < n; t++) dst[t] = src[t];
#endif
}
// (the warp-cooperative movers below split the bytes over the lanes)
ZD_DEV void copy_bytes(uint8_t *dst, const uint8_t *src, uint32_t n) {
    for (uint32_t i = ZD_LANE(); i < n; i += ZD_WARP) dst[i] = src[i];
}
ZD_DEV void fill_bytes(uint8_t *dst, uint8_t v, uint32_t n) {
    for (uint32_t i = ZD_LANE(); i < n; i += ZD_WARP) dst[i] = v;
}
// match copy with byte-serial semantics: out[pos+i] = out[pos-off+i]; when off < len the source is periodic
ZD_DEV void copy_match(uint8_t *out, uint32_t pos, uint32_t off, uint32_t len) {
    const uint8_t *from = out + pos - off;
    if (off >= len || ZD_WARP == 1) {
        if (ZD_WARP == 1) { for (uint32_t i = 0; i < len; i++) out[pos + i] = from[i]; }
        else for (uint32_t i = ZD_LANE(); i < len; i += ZD_WARP) out[pos + i] = from[i];
    } else {
        for (uint32_t i = ZD_LANE(); i < len; i += ZD_WARP) out[pos + i] = from[i % off];
    }
}

// ---- blocks / frames -----------------------------------------------------------------------------------------
struct Scratch {  // per warp, not persistent across blocks
    uint8_t weights[256];
    zc::FseDCell cells[512];
};

// ---- pieces of a compressed block, shared by the one-pass decoder and the two-pass (block-parallel) one ----------------
// Sequences section header: number of sequences, then the three tables.  On return *pp / *pleft are the FSE bitstream.
ZD_DEV int read_sequences_header(const uint8_t **pp, uint32_t *pleft, uint32_t *pnseq, Tables *T, Scratch *S) {
    const uint8_t *p = *pp;
    uint32_t left = *pleft;
    if (left < 1) return ERR_CORRUPT;
    uint32_t nseq = p[0];
    if (nseq == 0) { p += 1; left -= 1; }
    else if (nseq < 128) { p += 1; left -= 1; }
    else if (nseq < 255) { if (left < 2) return ERR_CORRUPT; nseq = ((nseq - 128) << 8) + p[1]; p += 2; left -= 2; }
    else { if (left < 3) return ERR_CORRUPT; nseq = p[1] + (p[2] << 8) + 0x7F00; p += 3; left -= 3; }
    if (nseq) {
        if (left < 1) return ERR_CORRUPT;
        const uint32_t modes = p[0];
        if (modes & 3) return ERR_CORRUPT;
        p += 1; left -= 1;
        int u = read_seq_table(p, left, modes >> 6, KIND_LL, T->ll, &T->ll_log, &T->have_ll, S->cells);
        if (u < 0) return u;
        p += u; left -= (uint32_t)u;
        u = read_seq_table(p, left, (modes >> 4) & 3, KIND_OF, T->of, &T->of_log, &T->have_of, S->cells);
        if (u < 0) return u;
        p += u; left -= (uint32_t)u;
        u = read_seq_table(p, left, (modes >> 2) & 3, KIND_ML, T->ml, &T->ml_log, &T->have_ml, S->cells);
        if (u < 0) return u;
        p += u; left -= (uint32_t)u;
        ZD_SYNC();
    }
    *pp = p; *pleft = left; *pnseq = nseq;
    return 0;
}

// One step of the serial FSE chain: the sequence's literal length, match length and offset VALUE (1..3 = repeat-offset codes),
// then the state transitions unless it is the block's last sequence.  Identical on every lane.
struct FseStates { uint32_t ll, of, ml; };
ZD_DEV void fse_step(BitReader *b, const Tables *T, FseStates *st, bool last, uint32_t *ll_out, uint32_t *ml_out, uint32_t *ofv_out) {
    const SeqCell cl = T->ll[st->ll], co = T->of[st->of], cm = T->ml[st->ml];
    // One refill per sequence covers its extra bits and state transitions in the common case (the window then holds more than 56
    // bits); the two further refills only happen for very long offsets / lengths.  All branches are warp-uniform.
    const uint32_t ofb = co.sym /* <= 31, checked when the table was built */, mlb = zc::ZTAB(ML_bits)[cm.sym], llb = zc::ZTAB(LL_bits)[cl.sym];
    br_refill(b);
    *ofv_out = (1u << ofb) + br_take(b, ofb);
    if (ofb + mlb + llb > 56) br_refill(b);
    *ml_out = zc::ZTAB(ML_base)[cm.sym] + br_take(b, mlb);
    *ll_out = zc::ZTAB(LL_base)[cl.sym] + br_take(b, llb);
    if (!last) {  // state updates: LL, ML, OF (at most 9 + 9 + 8 bits)
        if (ofb + mlb + llb > 56 - 26) br_refill(b);
        st->ll = cl.next_base + br_take(b, cl.nb_bits);
        st->ml = cm.next_base + br_take(b, cm.nb_bits);
        st->of = co.next_base + br_take(b, co.nb_bits);
    }
}

// Offset value -> offset, with the repeat-offset history update (RFC 8878 3.1.1.5).  Returns 0 for a corrupt code.
ZD_DEV uint32_t resolve_offset(uint32_t ofv, uint32_t ll, uint32_t *r0, uint32_t *r1, uint32_t *r2) {
    if (ofv > 3) { const uint32_t off = ofv - 3; *r2 = *r1; *r1 = *r0; *r0 = off; return off; }
    const uint32_t idx = ofv - 1 + (ll == 0 ? 1 : 0);
    if (idx == 0) return *r0;
    const uint32_t off = idx == 1 ? *r1 : idx == 2 ? *r2 : *r0 - 1;
    if (off == 0) return 0;
    if (idx >= 2) *r2 = *r1;
    *r1 = *r0; *r0 = off;
    return off;
}

// Matches that read output of their own batch go in waves: everything below the destination of the first unfinished match is
// final (literals are all placed, earlier matches are done), so every short match whose source ends there or earlier is copied by
// its own lane, all of them at once; a long or self-overlapping (periodic) match at the front takes the whole warp.  In repetitive
// data offsets are a few hundred bytes, i.e. a dozen sequences back: three or four waves per batch instead of up to 32 warp-wide
// copies one after the other.
#if defined(__CUDA_ARCH__)
ZD_DEV void run_match_waves(uint8_t *out, uint32_t my_match, uint32_t my_ml, uint32_t my_off, uint32_t src_end, uint32_t pending) {
    __syncwarp();  // what was written before (literals, independent matches) is visible
    while (pending) {
        const int first = __ffs((int)pending) - 1;
        const uint32_t first_dst = __shfl_sync(0xffffffffu, my_match, first);
        const bool ready = (pending >> ZD_LANE() & 1u) && src_end <= first_dst && my_ml <= 64u && my_off >= my_ml;
        const uint32_t rmask = __ballot_sync(0xffffffffu, ready);
        if (rmask >> first & 1u) {
            if (ready) lane_copy(out + my_match, out + my_match - my_off, my_ml);
            pending &= ~rmask;
        } else {
            const uint32_t moff = __shfl_sync(0xffffffffu, my_off, first), mlen = __shfl_sync(0xffffffffu, my_ml, first);
            copy_match(out, first_dst, moff, mlen);
            pending &= pending - 1;
        }
        __syncwarp();
    }
}
#endif

// Executes one batch of sequences (lane j holds sequence j of the batch): a warp scan turns lengths into positions, all literal
// runs are copied at once (their sources are never the output), matches that only read output older than the batch are copied
// lane-parallel, the rest in order with the whole warp on each.  Returns the new output position or < 0.
ZD_DEV int64_t execute_batch(uint8_t *out, uint32_t pos, uint32_t cap, uint32_t frame_start, const Literals *L, uint32_t *lit_pos_io,
                             uint32_t my_ll, uint32_t my_ml, uint32_t my_off, uint32_t nbatch, uint32_t *synced) {
    const uint32_t lit_pos = *lit_pos_io;
    uint32_t lit_excl, out_excl, lit_tot, out_tot;
#if defined(__CUDA_ARCH__)
    {
        uint32_t xl = my_ll, xo = my_ll + my_ml;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t yl = __shfl_up_sync(0xffffffffu, xl, d), yo = __shfl_up_sync(0xffffffffu, xo, d);
            if ((int)ZD_LANE() >= d) { xl += yl; xo += yo; }
        }
        lit_excl = xl - my_ll; out_excl = xo - (my_ll + my_ml);
        lit_tot = __shfl_sync(0xffffffffu, xl, 31); out_tot = __shfl_sync(0xffffffffu, xo, 31);
    }
#else
    lit_excl = 0; out_excl = 0; lit_tot = my_ll; out_tot = my_ll + my_ml;
#endif
    if (lit_pos + lit_tot > L->size) return ERR_CORRUPT;
    if ((uint64_t)pos + out_tot > cap) return ERR_CAPACITY;
    const uint32_t my_lit_out = pos + out_excl, my_match = my_lit_out + my_ll;
    const bool mine = ZD_LANE() < nbatch;
    int bad = mine && my_off > my_match - frame_start;
#if defined(__CUDA_ARCH__)
    bad = __any_sync(0xffffffffu, bad);
#endif
    if (bad) return ERR_CORRUPT;
    // literal runs (sources are the literal buffer / the input, never the output): each lane its own run
    if (mine) {
        const uint8_t *ls = L->ptr + lit_pos + lit_excl;
        if (L->rle) for (uint32_t t = 0; t < my_ll; t++) out[my_lit_out + t] = L->rle_byte;
        else lane_copy(out + my_lit_out, ls, my_ll);
    }
    // matches.  A match is independent if everything it reads was synced before this batch started.
    const uint32_t batch_synced = *synced;
    const uint32_t src_end = my_match - my_off + (my_off < my_ml ? my_off : my_ml);
    const bool indep = mine && src_end <= batch_synced;
    if (indep) {
        const uint8_t *from = out + my_match - my_off;  // non-overlapping with anything written in this batch
        if (my_off >= my_ml) lane_copy(out + my_match, from, my_ml);
        else for (uint32_t t = 0; t < my_ml; t++) out[my_match + t] = from[t % my_off];
    }
#if defined(__CUDA_ARCH__)
    run_match_waves(out, my_match, my_ml, my_off, src_end, __ballot_sync(0xffffffffu, mine && !indep));
#else
    if (mine && !indep) copy_match(out, my_match, my_off, my_ml);
#endif
    *lit_pos_io = lit_pos + lit_tot;
    *synced = pos + out_tot;
    return (int64_t)pos + out_tot;
}

// One compressed block: literals + sequences executed into out[pos..].  Returns new pos or < 0.
ZD_DEV int64_t decode_compressed_block(const uint8_t *src, uint32_t size, uint8_t *out, uint32_t pos, uint32_t cap, uint32_t frame_start,
                                       Tables *T, Scratch *S, uint8_t *litbuf, uint32_t rep[3], uint32_t *synced) {
    Literals L;
    const int lused = decode_literals(src, size, T, litbuf, &L, S->weights, S->cells);
    if (lused < 0) return lused;
    const uint8_t *p = src + lused;
    uint32_t left = size - (uint32_t)lused, nseq = 0;
    const int hu = read_sequences_header(&p, &left, &nseq, T, S);
    if (hu < 0) return hu;
    uint32_t lit_pos = 0;
    if (nseq) {
        BitReader b;
        if (br_init(&b, p, left) < 0) return ERR_CORRUPT;
        FseStates st;
        st.ll = br_read(&b, T->ll_log); st.of = br_read(&b, T->of_log); st.ml = br_read(&b, T->ml_log);
        uint32_t r0 = rep[0], r1 = rep[1], r2 = rep[2];
        // Sequences are handled in batches of one per lane: the serial FSE chain is walked identically by every lane and lane j
        // keeps sequence j; execute_batch then places and copies the batch.
        for (uint32_t i0 = 0; i0 < nseq; i0 += ZD_WARP) {
            const uint32_t nbatch = nseq - i0 < ZD_WARP ? nseq - i0 : ZD_WARP;
            uint32_t my_ll = 0, my_ml = 0, my_off = 1;
            int err = 0;
            for (uint32_t k = 0; k < nbatch; k++) {
                uint32_t ll, ml, ofv;
                fse_step(&b, T, &st, i0 + k + 1 >= nseq, &ll, &ml, &ofv);
                uint32_t off = resolve_offset(ofv, ll, &r0, &r1, &r2);
                if (off == 0) { err = 1; off = 1; }
                if (k == ZD_LANE()) { my_ll = ll; my_ml = ml; my_off = off; }
            }
            if (err || b.pos < 0) return ERR_CORRUPT;
            const int64_t np = execute_batch(out, pos, cap, frame_start, &L, &lit_pos, my_ll, my_ml, my_off, nbatch, synced);
            if (np < 0) return np;
            pos = (uint32_t)np;
        }
        if (b.pos != 0) return ERR_CORRUPT;
        rep[0] = r0; rep[1] = r1; rep[2] = r2;
    } else if (left != 0) return ERR_CORRUPT;
    const uint32_t rest = L.size - lit_pos;
    if ((uint64_t)pos + rest > cap) return ERR_CAPACITY;
    if (L.rle) fill_bytes(out + pos, L.rle_byte, rest); else copy_bytes(out + pos, L.ptr + lit_pos, rest);
    ZD_SYNC();
    return (int64_t)pos + rest;
}

// ---- two passes over a frame (block-parallel decoding) ---------------------------------------------------------------------
// The blocks of a frame chain in three ways: Repeat_Mode / Treeless reuse the previous block's tables, repeat-offset codes read
// the previous block's history, and matches read the previous blocks' output.  Only the last one needs the blocks in order:
//   tables   a serial pre-pass per frame (snapshot_frame_tables) reads nothing but the table descriptions of every block and
//            leaves each block a snapshot of the tables it decodes with;
//   history  pass 1 tracks the three repeat offsets symbolically: a slot is either a number or "what slot s held when the block
//            started, minus d"; the stored offsets and the block's final history are resolved in pass 2, in order;
//   output   pass 1 assumes every block but the last regenerates 128 KiB (true for K3 and for libzstd below level 16, checked in
//            pass 2), which tells it where its block starts.
// Pass 1 (one warp per BLOCK, all blocks of all frames at once) does everything that does not read the output: Huffman literals,
// the FSE chain, repeat offsets, positions, and it PLACES THE LITERALS in the output.  What is left for pass 2 (one warp per
// FRAME, in order) is one {destination, length, offset} triple per match.  Anything that breaks an assumption -- and every
// malformed frame -- goes through the one-pass decoder afterwards, which also produces the error.
struct StoredSeq { uint32_t mpos, ml, off; };  // match destination (frame-relative), length, offset (number or symbol)
struct BlockState { int32_t status; uint32_t nseq, regen; uint32_t hist[3]; };  // hist: the history the block leaves (numbers or symbols)
enum { NEED_ONE_PASS = -100 };
constexpr uint32_t HIST_SYM = 0x80000000u;  // symbol: HIST_SYM | slot << 16 | delta  ==  (incoming slot) - delta

// Literals section of a compressed block: its type, header length and body length (for Compressed / Treeless the body is the
// optional tree description followed by the streams).  false if the header is truncated.
ZD_DEV bool literals_section_span(const uint8_t *src, uint32_t size, uint32_t *type_out, uint32_t *hl_out, uint32_t *body_out) {
    if (size < 1) return false;
    const uint32_t type = src[0] & 3, sf = (src[0] >> 2) & 3;
    uint32_t hl, body;
    if (type < 2) {
        uint32_t n;
        if (sf == 0 || sf == 2) { hl = 1; n = src[0] >> 3; }
        else if (sf == 1) { if (size < 2) return false; hl = 2; n = (src[0] | src[1] << 8) >> 4; }
        else { if (size < 3) return false; hl = 3; n = (src[0] | src[1] << 8 | (uint32_t)src[2] << 16) >> 4; }
        body = type == 0 ? n : 1;
    } else {
        if (sf <= 1) { if (size < 3) return false; hl = 3; body = (src[0] | src[1] << 8 | (uint32_t)src[2] << 16) >> 14; }
        else if (sf == 2) { if (size < 4) return false; hl = 4; body = (src[0] | src[1] << 8 | (uint32_t)src[2] << 16 | (uint32_t)src[3] << 24) >> 18; }
        else { if (size < 5) return false; hl = 5; body = (uint32_t)((src[0] | src[1] << 8 | (uint32_t)src[2] << 16 | (uint64_t)src[3] << 24 | (uint64_t)src[4] << 32) >> 22); }
    }
    if (hl + body > size) return false;
    *type_out = type; *hl_out = hl; *body_out = body;
    return true;
}

// pass 1 of one compressed block whose output starts at out[out_start] (frame-relative; out is the frame's output, cap its size);
// *T holds the snapshot of the tables the block decodes with, or anything if the frame's blocks bring all their own (chained == false)
ZD_DEV void decode_block_first_pass(const uint8_t *src, uint32_t size, bool chained, Tables *T, Scratch *S, uint8_t *litbuf, uint8_t *out, uint32_t out_start, uint32_t cap,
                                    bool first_block, StoredSeq *seqs, uint32_t seq_cap, BlockState *bs_out) {
    BlockState bs;
    bs.status = 0; bs.nseq = 0; bs.regen = 0;
    if (!chained) T->have_huf = T->have_ll = T->have_ml = T->have_of = 0;
    // repeat offsets: the frame's first block starts from 1, 4, 8; any other from three symbols
    uint32_t r0 = first_block ? 1u : HIST_SYM, r1 = first_block ? 4u : HIST_SYM | 1u << 16, r2 = first_block ? 8u : HIST_SYM | 2u << 16;
    Literals L;
    const int lused = decode_literals(src, size, T, litbuf, &L, S->weights, S->cells);
    int rc = lused < 0 ? lused : 0;
    uint32_t pos = out_start, lit_pos = 0, nseq = 0;
    if (!rc) {
        const uint8_t *p = src + lused;
        uint32_t left = size - (uint32_t)lused;
        rc = read_sequences_header(&p, &left, &nseq, T, S);
        if (!rc && nseq > seq_cap) rc = ERR_CAPACITY;  // more sequences than the slot holds: the one-pass decoder takes the frame
        if (!rc && nseq) {
            BitReader b;
            if (br_init(&b, p, left) < 0) rc = ERR_CORRUPT;
            FseStates st;
            if (!rc) { st.ll = br_read(&b, T->ll_log); st.of = br_read(&b, T->of_log); st.ml = br_read(&b, T->ml_log); }
            for (uint32_t i0 = 0; i0 < nseq && !rc; i0 += ZD_WARP) {
                const uint32_t nbatch = nseq - i0 < ZD_WARP ? nseq - i0 : ZD_WARP;
                uint32_t my_ll = 0, my_ml = 0, my_off = 1;
                int err = 0;
                for (uint32_t k = 0; k < nbatch; k++) {
                    uint32_t ll, ml, ofv, off;
                    fse_step(&b, T, &st, i0 + k + 1 >= nseq, &ll, &ml, &ofv);
                    if (ofv > 3) { off = ofv - 3; r2 = r1; r1 = r0; r0 = off; }
                    else {
                        const uint32_t idx = ofv - 1 + (ll == 0 ? 1 : 0);
                        if (idx == 0) off = r0;
                        else {
                            if (idx == 3) {  // r0 - 1: a number, or one more on a symbol's delta
                                if (r0 & HIST_SYM) { off = r0 + 1; if ((off & 0xFFFFu) == 0) err = 1; }
                                else { off = r0 - 1; if (off == 0) err = 1; }
                            } else off = idx == 1 ? r1 : r2;
                            if (idx >= 2) r2 = r1;
                            r1 = r0; r0 = off;
                        }
                    }
                    if (k == ZD_LANE()) { my_ll = ll; my_ml = ml; my_off = off; }
                }
                if (err || b.pos < 0) { rc = ERR_CORRUPT; break; }
                // positions
                uint32_t lit_excl, out_excl, lit_tot, out_tot;
#if defined(__CUDA_ARCH__)
                {
                    uint32_t xl = my_ll, xo = my_ll + my_ml;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t yl = __shfl_up_sync(0xffffffffu, xl, d), yo = __shfl_up_sync(0xffffffffu, xo, d);
                        if ((int)ZD_LANE() >= d) { xl += yl; xo += yo; }
                    }
                    lit_excl = xl - my_ll; out_excl = xo - (my_ll + my_ml);
                    lit_tot = __shfl_sync(0xffffffffu, xl, 31); out_tot = __shfl_sync(0xffffffffu, xo, 31);
                }
#else
                lit_excl = 0; out_excl = 0; lit_tot = my_ll; out_tot = my_ll + my_ml;
#endif
                if (lit_pos + lit_tot > L.size) { rc = ERR_CORRUPT; break; }
                if ((uint64_t)pos + out_tot > cap) { rc = ERR_CAPACITY; break; }
                const uint32_t my_lit_out = pos + out_excl, my_match = my_lit_out + my_ll;
                const bool mine = ZD_LANE() < nbatch;
                int bad = mine && !(my_off & HIST_SYM) && my_off > my_match;  // symbols are checked when pass 2 resolves them
#if defined(__CUDA_ARCH__)
                bad = __any_sync(0xffffffffu, bad);
#endif
                if (bad) { rc = ERR_CORRUPT; break; }
                if (mine) {
                    if (L.rle) for (uint32_t t = 0; t < my_ll; t++) out[my_lit_out + t] = L.rle_byte;
                    else lane_copy(out + my_lit_out, L.ptr + lit_pos + lit_excl, my_ll);
                    StoredSeq sq;
                    sq.mpos = my_match; sq.ml = my_ml; sq.off = my_off;
                    seqs[i0 + ZD_LANE()] = sq;
                }
                lit_pos += lit_tot;
                pos += out_tot;
            }
            if (!rc && b.pos != 0) rc = ERR_CORRUPT;
        } else if (!rc && left != 0) rc = ERR_CORRUPT;
        if (!rc) {
            const uint32_t rest = L.size - lit_pos;
            if ((uint64_t)pos + rest > cap) rc = ERR_CAPACITY;
            else {
                if (L.rle) fill_bytes(out + pos, L.rle_byte, rest); else copy_bytes(out + pos, L.ptr + lit_pos, rest);
                pos += rest;
            }
        }
    }
    bs.status = rc; bs.nseq = nseq; bs.regen = pos - out_start;
    bs.hist[0] = r0; bs.hist[1] = r1; bs.hist[2] = r2;
    ZD_SYNC();
    if (ZD_LANE() == 0) *bs_out = bs;
}

// A whole chunk-record payload: >= 1 zstd frames and skippable frames, nothing else (what stock
// ZSTD_decompress accepts).  Returns decoded size or < 0.
ZD_DEV int64_t decode_payload(const uint8_t *src, uint32_t size, uint8_t *out, uint32_t cap, Tables *T, Scratch *S, uint8_t *litbuf) {
    uint32_t ip = 0, pos = 0;
    uint32_t synced = 0;  // output below this position is visible to every lane
    while (ip < size) {
        if (size - ip < 4) return ERR_CORRUPT;
        const uint32_t magic = src[ip] | src[ip + 1] << 8 | (uint32_t)src[ip + 2] << 16 | (uint32_t)src[ip + 3] << 24;
        if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {  // skippable frame
            if (size - ip < 8) return ERR_CORRUPT;
            const uint32_t n = src[ip + 4] | src[ip + 5] << 8 | (uint32_t)src[ip + 6] << 16 | (uint32_t)src[ip + 7] << 24;
            if (n > size - ip - 8) return ERR_CORRUPT;
            ip += 8 + n;
            continue;
        }
        if (magic != Z_MAGIC) return ERR_CORRUPT;
        if (size - ip < 5) return ERR_CORRUPT;
        const uint32_t fhd = src[ip + 4];
        const uint32_t fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, has_checksum = (fhd >> 2) & 1, did = fhd & 3;
        if (fhd & 0x08) return ERR_CORRUPT;  // reserved bit
        uint32_t h = 5;
        if (!single) {
            if (size - ip < h + 1) return ERR_CORRUPT;
            const uint32_t wd = src[ip + h], wlog = 10 + (wd >> 3);
            if (wlog > 31) return ERR_CORRUPT;
            h += 1;
        }
        const uint32_t did_bytes = did == 3 ? 4 : did;
        if (size - ip < h + did_bytes) return ERR_CORRUPT;
        for (uint32_t i = 0; i < did_bytes; i++) if (src[ip + h + i]) return ERR_CORRUPT;  // a dictionary is required: not supported, as in the reference
        h += did_bytes;
        const uint32_t fcs_bytes = fcs_flag == 0 ? single : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
        if (size - ip < h + fcs_bytes) return ERR_CORRUPT;
        uint64_t fcs = 0;
        for (uint32_t i = 0; i < fcs_bytes; i++) fcs |= (uint64_t)src[ip + h + i] << (8 * i);
        if (fcs_flag == 1) fcs += 256;
        h += fcs_bytes;
        if (fcs_bytes && fcs > (uint64_t)cap - pos) return ERR_CAPACITY;
        ip += h;
        const uint32_t frame_start = pos;
        uint32_t rep[3] = {1, 4, 8};
        T->have_huf = T->have_ll = T->have_ml = T->have_of = 0;
        for (;;) {
            if (size - ip < 3) return ERR_CORRUPT;
            const uint32_t bh = src[ip] | src[ip + 1] << 8 | (uint32_t)src[ip + 2] << 16;
            const uint32_t last = bh & 1, type = (bh >> 1) & 3, bsz = bh >> 3;
            ip += 3;
            if (type != 3 && bsz > Z_BLOCK_MAX) return ERR_CORRUPT;
            if (type == 0) {  // raw
                if (bsz > size - ip) return ERR_CORRUPT;
                if (bsz > cap - pos) return ERR_CAPACITY;
                copy_bytes(out + pos, src + ip, bsz);
                ZD_SYNC();
                pos += bsz; ip += bsz; synced = pos;
            } else if (type == 1) {  // RLE
                if (size - ip < 1) return ERR_CORRUPT;
                if (bsz > cap - pos) return ERR_CAPACITY;
                fill_bytes(out + pos, src[ip], bsz);
                ZD_SYNC();
                pos += bsz; ip += 1; synced = pos;
            } else if (type == 2) {
                if (bsz > size - ip || bsz > Z_BLOCK_MAX) return ERR_CORRUPT;
                const int64_t np = decode_compressed_block(src + ip, bsz, out, pos, cap, frame_start, T, S, litbuf, rep, &synced);
                if (np >= 0) synced = (uint32_t)np;  // the block ends with a warp sync
                if (np < 0) return np;
                if ((uint64_t)np - pos > Z_BLOCK_MAX) return ERR_CORRUPT;
                pos = (uint32_t)np; ip += bsz;
            } else return ERR_CORRUPT;
            if (last) break;
        }
        if (fcs_bytes && (uint64_t)(pos - frame_start) != fcs) return ERR_CORRUPT;
        if (has_checksum) { if (size - ip < 4) return ERR_CORRUPT; ip += 4; }  // content checksum is not verified (the reference never writes one)
    }
    return pos;
}

// Two-pass eligibility of a payload: exactly one zstd frame of 2 .. max_blocks well-formed blocks, at least one of them
// compressed.  Anything else -- several frames, skippable frames, malformed headers, a single block (nothing to gain) -- is left
// to decode_payload, which also produces the error.  Fills tasks[] (with the output start of every block under the 128 KiB
// assumption) and returns the number of blocks, or 0 if not eligible.
struct BlockTask { uint32_t src_off, size, type, out_start; };
struct FrameInfo { uint32_t nblocks, has_fcs; uint64_t fcs; uint32_t chained, reserved; };  // chained: some block inherits tables (snapshots needed)
ZD_DEV uint32_t scan_frame(const uint8_t *src, uint32_t size, uint32_t max_blocks, BlockTask *tasks, FrameInfo *fi) {
    if (size < 6) return 0;
    const uint32_t magic = src[0] | src[1] << 8 | (uint32_t)src[2] << 16 | (uint32_t)src[3] << 24;
    if (magic != Z_MAGIC) return 0;
    const uint32_t fhd = src[4];
    const uint32_t fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, has_checksum = (fhd >> 2) & 1, did = fhd & 3;
    if ((fhd & 0x08) || did) return 0;
    uint32_t h = 5;
    if (!single) { if (size < h + 1 || 10 + (src[h] >> 3) > 31) return 0; h += 1; }
    const uint32_t fcs_bytes = fcs_flag == 0 ? single : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
    if (size < h + fcs_bytes) return 0;
    uint64_t fcs = 0;
    for (uint32_t i = 0; i < fcs_bytes; i++) fcs |= (uint64_t)src[h + i] << (8 * i);
    if (fcs_flag == 1) fcs += 256;
    uint32_t ip = h + fcs_bytes, nb = 0, start = 0;
    bool any = false, chained = false;
    for (;;) {
        if (size - ip < 3 || nb >= max_blocks) return 0;
        const uint32_t bh = src[ip] | src[ip + 1] << 8 | (uint32_t)src[ip + 2] << 16;
        const uint32_t last = bh & 1, type = (bh >> 1) & 3, bsz = bh >> 3;
        ip += 3;
        if (type == 3 || bsz > Z_BLOCK_MAX) return 0;
        const uint32_t body = type == 1 ? 1u : bsz;
        if (body > size - ip) return 0;
        tasks[nb].src_off = ip; tasks[nb].size = bsz; tasks[nb].type = type; tasks[nb].out_start = start;
        start += type == 2 ? Z_BLOCK_MAX : bsz;
        any = any || type == 2;
        if (type == 2) {  // Treeless literals or a Repeat_Mode table: the block decodes with tables of an earlier one
            uint32_t lt, hl, body;
            if (!literals_section_span(src + ip, bsz, &lt, &hl, &body) || hl + body >= bsz) chained = true;  // (or malformed: the snapshot pass finds out)
            else {
                const uint8_t *q = src + ip + hl + body;
                const uint32_t left = bsz - hl - body, ns = q[0], nh = ns < 128 ? 1 : ns < 255 ? 2 : 3;
                if (lt == 3) chained = true;
                else if (ns) {
                    if (left < nh + 1) chained = true;
                    else { const uint32_t m = q[nh]; if ((m >> 6) == 3 || ((m >> 4) & 3) == 3 || ((m >> 2) & 3) == 3) chained = true; }
                }
            }
        }
        nb++;
        ip += body;
        if (last) break;
    }
    if (has_checksum) { if (size - ip < 4) return 0; ip += 4; }
    if (ip != size || nb < 2 || !any) return 0;
    fi->nblocks = nb; fi->has_fcs = fcs_bytes != 0; fi->fcs = fcs; fi->chained = chained ? 1u : 0u; fi->reserved = 0;
    return nb;
}

// Serial over the blocks of one frame, but reading only their table descriptions: after it, snaps[b] holds the tables compressed
// block b decodes with (its own, or what Repeat_Mode / Treeless make it inherit).  false = malformed (one-pass decoder).
ZD_DEV bool snapshot_frame_tables(const uint8_t *src, const BlockTask *tasks, uint32_t nblocks, Tables *T, Scratch *S, Tables *snaps) {
    T->have_huf = T->have_ll = T->have_ml = T->have_of = 0;
    for (uint32_t b = 0; b < nblocks; b++) {
        if (tasks[b].type != 2) continue;
        const uint8_t *bp = src + tasks[b].src_off;
        const uint32_t size = tasks[b].size;
        uint32_t type, hl, body;
        if (!literals_section_span(bp, size, &type, &hl, &body)) return false;
        if (type == 2) { if (read_huf_tree(bp + hl, body, T, S->weights, S->cells) < 0) return false; }
        else if (type == 3 && !T->have_huf) return false;
        const uint8_t *p = bp + hl + body;
        uint32_t left = size - hl - body, nseq = 0;
        if (read_sequences_header(&p, &left, &nseq, T, S) < 0) return false;
        ZD_SYNC();
        {   // all lanes copy the table state
            const uint32_t *from = reinterpret_cast<const uint32_t *>(T);
            uint32_t *to = reinterpret_cast<uint32_t *>(snaps + b);
            for (uint32_t i = ZD_LANE(); i < sizeof(Tables) / 4; i += ZD_WARP) to[i] = from[i];
        }
        ZD_SYNC();
    }
    return true;
}

// number or symbol -> number, given the history at the block's start; 0 = not a valid offset
ZD_DEV uint32_t resolve_hist(uint32_t h, const uint32_t R[3]) {
    if (!(h & HIST_SYM)) return h;
    const uint32_t base = R[(h >> 16) & 3u], d = h & 0xFFFFu;
    return base > d ? base - d : 0u;
}

// pass 2 of a whole frame: raw / RLE blocks and the matches of the compressed ones, in order.  Returns the decoded size, or
// NEED_ONE_PASS when pass 1 failed on a block, a block does not start where it was assumed to, or an offset is not valid.
ZD_DEV int64_t execute_frame_matches(const uint8_t *src, const BlockTask *tasks, const FrameInfo *fi, const BlockState *states, const StoredSeq *seqs,
                                     size_t seq_stride, uint8_t *out, uint32_t cap) {
    uint32_t pos = 0;
    uint32_t R[3] = {1, 4, 8};
    for (uint32_t b = 0; b < fi->nblocks; b++) {
        const BlockTask t = tasks[b];
        if (pos != t.out_start) return NEED_ONE_PASS;
        if (t.type == 0) {
            if (t.size > cap - pos) return NEED_ONE_PASS;
            copy_bytes(out + pos, src + t.src_off, t.size);
            ZD_SYNC();
            pos += t.size;
        } else if (t.type == 1) {
            if (t.size > cap - pos) return NEED_ONE_PASS;
            fill_bytes(out + pos, src[t.src_off], t.size);
            ZD_SYNC();
            pos += t.size;
        } else {
            const BlockState bs = states[b];
            if (bs.status < 0 || bs.regen > Z_BLOCK_MAX) return NEED_ONE_PASS;
            const StoredSeq *sq = seqs + (size_t)b * seq_stride;
            StoredSeq next;  // the next batch's triples are requested before this batch's copies (the pass is latency-bound)
            next.mpos = 1; next.ml = 0; next.off = 1;
            if (ZD_LANE() < bs.nseq) next = sq[ZD_LANE()];
            for (uint32_t i0 = 0; i0 < bs.nseq; i0 += ZD_WARP) {
                const uint32_t nbatch = bs.nseq - i0 < ZD_WARP ? bs.nseq - i0 : ZD_WARP;
                StoredSeq mine = next;
                next.mpos = 1; next.ml = 0; next.off = 1;
                if (i0 + ZD_WARP + ZD_LANE() < bs.nseq) next = sq[i0 + ZD_WARP + ZD_LANE()];
                mine.off = resolve_hist(mine.off, R);
                int bad = mine.off == 0 || mine.off > mine.mpos;
#if defined(__CUDA_ARCH__)
                bad = __any_sync(0xffffffffu, bad);
#endif
                if (bad) return NEED_ONE_PASS;
#if defined(__CUDA_ARCH__)
                const uint32_t src_end = mine.mpos - mine.off + (mine.off < mine.ml ? mine.off : mine.ml);
                run_match_waves(out, mine.mpos, mine.ml, mine.off, src_end, nbatch >= 32 ? 0xffffffffu : (1u << nbatch) - 1u);
#else
                copy_match(out, mine.mpos, mine.off, mine.ml);
#endif
            }
            const uint32_t n0 = resolve_hist(bs.hist[0], R), n1 = resolve_hist(bs.hist[1], R), n2 = resolve_hist(bs.hist[2], R);
            if (bs.nseq && (n0 == 0 || n1 == 0 || n2 == 0)) return NEED_ONE_PASS;
            if (bs.nseq) { R[0] = n0; R[1] = n1; R[2] = n2; }
            pos += bs.regen;
        }
    }
    if (fi->has_fcs && (uint64_t)pos != fi->fcs) return NEED_ONE_PASS;
    return pos;
}

}  // namespace zd
