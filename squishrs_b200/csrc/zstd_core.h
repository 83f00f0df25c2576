// zstd format core shared by the encoder (K3) and decoder (K4): code tables, bit I/O, FSE and
// Huffman table construction.  Everything is __host__ __device__ so the serial pieces can be
// exercised on the CPU by tests/ (against stock libzstd) and run by single lanes on the GPU.
// Written from the format specification (RFC 8878), not from libzstd sources.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define ZHD __host__ __device__ __forceinline__
#define ZHDN __host__ __device__ inline
#define ZCONST __constant__ const
#else
#define ZHD static inline
#define ZHDN static inline
#define ZCONST static const
#endif

#define Z_BLOCK_MAX (128u * 1024u)
#define Z_MAGIC 0xFD2FB528u
#define Z_LL_MAXLOG 9
#define Z_ML_MAXLOG 9
#define Z_OF_MAXLOG 8
#define Z_HUF_MAXBITS 11
#define Z_MINMATCH 3

namespace zc {

// ---- code tables (RFC 8878 3.1.1.3.2.1) ---------------------------------------------------
#if defined(__CUDA_ARCH__)
#define ZTAB(name) name##_d
#else
#define ZTAB(name) name##_h
#endif
#define ZDEF_TABLE(type, name, n, ...)            \
    static const type name##_h[n] = __VA_ARGS__;  \
    ZCONST type name##_d[n] = __VA_ARGS__;
#if !defined(__CUDACC__)
#undef ZDEF_TABLE
#define ZDEF_TABLE(type, name, n, ...) static const type name##_h[n] = __VA_ARGS__;
#endif

ZDEF_TABLE(uint32_t, LL_base, 36, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512,
                                    1024, 2048, 4096, 8192, 16384, 32768, 65536})
ZDEF_TABLE(uint8_t, LL_bits, 36, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16})
ZDEF_TABLE(uint32_t, ML_base, 53, {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32,
                                    33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539})
ZDEF_TABLE(uint8_t, ML_bits, 53, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                   1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16})
ZDEF_TABLE(int16_t, LL_defnorm, 36, {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1})
ZDEF_TABLE(int16_t, ML_defnorm, 53, {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                      1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1})
ZDEF_TABLE(int16_t, OF_defnorm, 29, {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1})

ZHD uint32_t highbit(uint32_t v) {  // floor(log2(v)), v > 0
#if defined(__CUDA_ARCH__)
    return 31u - (uint32_t)__clz((int)v);
#else
    return 31u - (uint32_t)__builtin_clz(v);
#endif
}
ZHD uint32_t ll_code(uint32_t ll) {
    if (ll < 16) return ll;
    if (ll >= 64) return highbit(ll) + 19;
    uint32_t c = 16;
    while (c < 24 && ZTAB(LL_base)[c + 1] <= ll) c++;
    return c;
}
ZHD uint32_t ml_code(uint32_t ml) {  // ml = real match length >= 3
    uint32_t b = ml - 3;
    if (b < 32) return b;
    if (b >= 128) return highbit(b) + 36;
    uint32_t c = 32;
    while (c < 42 && ZTAB(ML_base)[c + 1] <= ml) c++;
    return c;
}

// ---- forward (LSB-first) bit writer used for FSE table descriptions ---------------------------
struct FwdWriter {
    uint8_t *p; uint32_t pos; uint64_t acc; uint32_t n;
};
ZHD void fw_init(FwdWriter *w, uint8_t *p) { w->p = p; w->pos = 0; w->acc = 0; w->n = 0; }
ZHD void fw_put(FwdWriter *w, uint32_t v, uint32_t bits) {
    w->acc |= (uint64_t)v << w->n;
    w->n += bits;
    while (w->n >= 8) { w->p[w->pos++] = (uint8_t)w->acc; w->acc >>= 8; w->n -= 8; }
}
ZHD uint32_t fw_finish(FwdWriter *w) {
    if (w->n) { w->p[w->pos++] = (uint8_t)w->acc; w->acc = 0; w->n = 0; }
    return w->pos;
}

// ---- backward bit writer: fields appended LSB-first; the LAST field appended is read FIRST --------
struct BackWriter {
    uint8_t *p; uint32_t pos; uint64_t acc; uint32_t n;
};
ZHD void bw_init(BackWriter *w, uint8_t *p) { w->p = p; w->pos = 0; w->acc = 0; w->n = 0; }
ZHD void bw_put(BackWriter *w, uint32_t v, uint32_t bits) {  // bits <= 32; v must fit in `bits`
    w->acc |= (uint64_t)v << w->n;
    w->n += bits;
    while (w->n >= 8) { w->p[w->pos++] = (uint8_t)w->acc; w->acc >>= 8; w->n -= 8; }
}
ZHD uint32_t bw_finish(BackWriter *w) {  // sentinel 1 bit then zero padding
    bw_put(w, 1, 1);
    if (w->n) { w->p[w->pos++] = (uint8_t)w->acc; w->acc = 0; w->n = 0; }
    return w->pos;
}

// ---- FSE: normalisation, table description writer, encode table -------------------------------
// Fixed-point (8 fractional bits) log2 for cost estimates.
ZHD uint32_t log2_fix8(uint32_t v) {  // v >= 1
    uint32_t hb = highbit(v);
    uint32_t frac = hb >= 8 ? (v >> (hb - 8)) & 0xFF : (v << (8 - hb)) & 0xFF;  // linear mantissa
    return (hb << 8) + frac;
}

// choose a table log for `total` symbols over an alphabet whose highest used symbol is max_sym
ZHD uint32_t fse_table_log(uint32_t max_log, uint32_t total, uint32_t max_sym) {
    uint32_t by_src = total > 1 ? highbit(total - 1) : 1;
    by_src = by_src > 2 ? by_src - 2 : 1;
    uint32_t min_a = highbit(total) + 1, min_b = highbit(max_sym + 1) + 2;
    uint32_t min_bits = min_a < min_b ? min_a : min_b;
    uint32_t tl = max_log;
    if (by_src < tl) tl = by_src;
    if (min_bits > tl) tl = min_bits;
    if (tl < 5) tl = 5;
    if (tl > max_log) tl = max_log;
    return tl;
}

// Normalise counts[0..max_sym] (sum = total > 0, at least two non-zero entries) to sum 2^tl, every used
// symbol >= 1.  Returns 0 on success.
ZHDN int fse_normalize(int16_t *norm, uint32_t tl, const uint32_t *counts, uint32_t total, uint32_t max_sym) {
    const uint32_t size = 1u << tl;
    uint32_t used = 0;
    for (uint32_t s = 0; s <= max_sym; s++) used += counts[s] != 0;
    if (used > size) return -1;
    int32_t remaining = (int32_t)size;
    uint32_t largest = 0, largest_cnt = 0;
    const uint64_t scale = ((uint64_t)size << 32) / total;
    for (uint32_t s = 0; s <= max_sym; s++) {
        if (!counts[s]) { norm[s] = 0; continue; }
        uint64_t x = (uint64_t)counts[s] * scale;           // 32.32 fixed point share of the table
        uint32_t p = (uint32_t)(x >> 32);
        uint32_t frac = (uint32_t)x;
        // round to nearest, but bias small probabilities up (they lose most from truncation)
        if (p == 0) p = 1;
        else if (frac > (p < 8 ? 0x60000000u : 0x80000000u)) p++;
        norm[s] = (int16_t)p;
        remaining -= (int32_t)p;
        if (counts[s] > largest_cnt) { largest_cnt = counts[s]; largest = s; }
    }
    if (remaining > 0) norm[largest] = (int16_t)(norm[largest] + remaining);
    while (remaining < 0) {  // take back from whichever symbol loses least (largest norm/count surplus)
        uint32_t best = 0xFFFFFFFFu;
        uint64_t best_cost = ~0ull;
        for (uint32_t s = 0; s <= max_sym; s++) {
            if (norm[s] < 2) continue;
            // cost of norm -> norm-1  ~ counts * log2(norm/(norm-1)) ~ counts / norm
            uint64_t cost = ((uint64_t)counts[s] << 16) / (uint32_t)norm[s];
            if (cost < best_cost) { best_cost = cost; best = s; }
        }
        if (best == 0xFFFFFFFFu) return -1;
        norm[best]--;
        remaining++;
    }
    return 0;
}

// FSE table description (RFC 8878 4.1.1).  Returns bytes written.
ZHDN uint32_t fse_write_ncount(uint8_t *dst, const int16_t *norm, uint32_t max_sym, uint32_t tl) {
    FwdWriter w;
    fw_init(&w, dst);
    const int32_t size = 1 << tl;
    fw_put(&w, tl - 5, 4);
    int32_t remaining = size + 1, threshold = size;
    uint32_t nb_bits = tl + 1;
    uint32_t s = 0;
    bool prev_zero = false;
    while (remaining > 1 && s <= max_sym) {
        if (prev_zero) {  // run of zero-probability symbols following a zero
            uint32_t start = s;
            while (s <= max_sym && norm[s] == 0) s++;
            uint32_t run = s - start;
            while (run >= 3) { fw_put(&w, 3, 2); run -= 3; }
            fw_put(&w, run, 2);
            if (s > max_sym) break;
        }
        int32_t count = norm[s++];
        const int32_t max = 2 * threshold - 1 - remaining;
        remaining -= count < 0 ? -count : count;
        count++;  // +1: the value -1 ("less than one") is coded as 0
        if (count >= threshold) count += max;
        fw_put(&w, (uint32_t)count, nb_bits - (count < max ? 1 : 0));
        prev_zero = (count == 1);
        while (remaining < threshold) { nb_bits--; threshold >>= 1; }
    }
    return fw_finish(&w);
}

// Encoder-side FSE table.  state values live in [size, 2*size).
struct FseCTable {
    uint16_t next_state[512];      // indexed by (state >> nb_bits) + delta_find_state
    int32_t delta_find_state[64];  // per symbol
    uint32_t delta_nb_bits[64];    // per symbol
    uint32_t tl;
};

// Spread symbols over the table exactly as the decoder will (RFC 8878 4.1.1 "FSE decoding table")
// and derive the encoding transitions.  `scratch` needs 512 bytes.
ZHDN void fse_build_ctable(FseCTable *ct, const int16_t *norm, uint32_t max_sym, uint32_t tl, uint8_t *scratch) {
    const uint32_t size = 1u << tl, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    uint8_t *cell_sym = scratch;
    uint32_t high = size - 1;
    uint16_t cumul[65];
    cumul[0] = 0;
    for (uint32_t s = 0; s <= max_sym; s++) {
        if (norm[s] == -1) { cumul[s + 1] = (uint16_t)(cumul[s] + 1); cell_sym[high--] = (uint8_t)s; }
        else cumul[s + 1] = (uint16_t)(cumul[s] + norm[s]);
    }
    uint32_t pos = 0;
    for (uint32_t s = 0; s <= max_sym; s++)
        for (int32_t i = 0; i < norm[s]; i++) {
            cell_sym[pos] = (uint8_t)s;
            do { pos = (pos + step) & mask; } while (pos > high);
        }
    // cells of one symbol, visited in ascending table index, get consecutive sub-states
    for (uint32_t u = 0; u < size; u++) { uint32_t s = cell_sym[u]; ct->next_state[cumul[s]++] = (uint16_t)(size + u); }
    uint32_t total = 0;
    for (uint32_t s = 0; s <= max_sym; s++) {
        int32_t n = norm[s];
        if (n == 0) { ct->delta_nb_bits[s] = ((tl + 1) << 16) - size; ct->delta_find_state[s] = 0; continue; }
        if (n == -1 || n == 1) {
            ct->delta_nb_bits[s] = (tl << 16) - size;
            ct->delta_find_state[s] = (int32_t)total - 1;
            total += 1;
        } else {
            uint32_t max_bits_out = tl - highbit((uint32_t)n - 1);
            uint32_t min_state_plus = (uint32_t)n << max_bits_out;
            ct->delta_nb_bits[s] = (max_bits_out << 16) - min_state_plus;
            ct->delta_find_state[s] = (int32_t)total - n;
            total += (uint32_t)n;
        }
    }
    ct->tl = tl;
}
ZHD uint32_t fse_init_state(const FseCTable *ct, uint32_t sym) {
    uint32_t nb = (ct->delta_nb_bits[sym] + (1u << 15)) >> 16;
    uint32_t value = (nb << 16) - ct->delta_nb_bits[sym];
    return ct->next_state[(int32_t)(value >> nb) + ct->delta_find_state[sym]];
}
// emits the low bits of `state`, returns the next state
ZHD uint32_t fse_encode(const FseCTable *ct, BackWriter *w, uint32_t state, uint32_t sym) {
    uint32_t nb = (state + ct->delta_nb_bits[sym]) >> 16;
    bw_put(w, state & ((1u << nb) - 1), nb);
    return ct->next_state[(int32_t)(state >> nb) + ct->delta_find_state[sym]];
}
ZHD void fse_flush_state(const FseCTable *ct, BackWriter *w, uint32_t state) { bw_put(w, state & ((1u << ct->tl) - 1), ct->tl); }

// cost in 1/256 bit of coding `counts` with the distribution `norm` at table log tl
ZHDN uint64_t fse_cost(const int16_t *norm, uint32_t tl, const uint32_t *counts, uint32_t max_sym) {
    uint64_t c = 0;
    for (uint32_t s = 0; s <= max_sym; s++) {
        if (!counts[s]) continue;
        int32_t n = norm[s] == -1 ? 1 : norm[s];
        if (n <= 0) return ~0ull >> 2;  // symbol not representable
        c += (uint64_t)counts[s] * ((tl << 8) - log2_fix8((uint32_t)n));
    }
    return c;
}

// ---- FSE decode table (RFC 8878 4.1.1) ----------------------------------------------------------
struct FseDCell { uint8_t sym, nb_bits; uint16_t base; };
ZHDN void fse_build_dtable(FseDCell *t, const int16_t *norm, uint32_t max_sym, uint32_t tl, uint16_t *next /* >= max_sym+1 */) {
    const uint32_t size = 1u << tl, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    uint32_t high = size - 1;
    for (uint32_t s = 0; s <= max_sym; s++) {
        if (norm[s] == -1) { t[high--].sym = (uint8_t)s; next[s] = 1; }
        else next[s] = (uint16_t)norm[s];
    }
    uint32_t pos = 0;
    for (uint32_t s = 0; s <= max_sym; s++)
        for (int32_t i = 0; i < norm[s]; i++) {
            t[pos].sym = (uint8_t)s;
            do { pos = (pos + step) & mask; } while (pos > high);
        }
    for (uint32_t u = 0; u < size; u++) {
        uint32_t s = t[u].sym, x = next[s]++;
        uint32_t nb = tl - highbit(x);
        t[u].nb_bits = (uint8_t)nb;
        t[u].base = (uint16_t)((x << nb) - size);
    }
}

// ---- Huffman --------------------------------------------------------------------------------------
// Code lengths (<= max_bits) for counts[0..255]; lens[s] = 0 for unused symbols.  Needs >= 2 used symbols.
// Package-free construction: sort, two-queue merge, then repair depths over the limit keeping Kraft equality.
// huf_build_lengths_sorted: work[0..n) already holds the used symbols in ascending (count, symbol) order and lens[] is zeroed
// (the GPU entropy stage sorts with all lanes of the warp; the order is the one the stable insertion sort below produces).
ZHDN uint32_t huf_build_lengths_sorted(uint8_t *lens, const uint32_t *counts, uint32_t max_bits, uint16_t *work /* 1280 u16 */, uint32_t n);
ZHDN uint32_t huf_build_lengths(uint8_t *lens, const uint32_t *counts, uint32_t max_bits, uint16_t *work /* 1280 u16 */) {
    uint16_t *order = work;          // 256: symbols sorted by ascending count
    uint32_t n = 0;
    for (uint32_t s = 0; s < 256; s++) { lens[s] = 0; if (counts[s]) order[n++] = (uint16_t)s; }
    if (n < 2) return 0;
    for (uint32_t i = 1; i < n; i++) {  // insertion sort by count (n <= 256)
        uint16_t s = order[i];
        uint32_t c = counts[s], j = i;
        while (j > 0 && counts[order[j - 1]] > c) { order[j] = order[j - 1]; j--; }
        order[j] = s;
    }
    return huf_build_lengths_sorted(lens, counts, max_bits, work, n);
}
ZHDN uint32_t huf_build_lengths_sorted(uint8_t *lens, const uint32_t *counts, uint32_t max_bits, uint16_t *work /* 1280 u16 */, uint32_t n) {
    uint16_t *order = work;          // 256: symbols sorted by ascending count
    uint16_t *parent = work + 256;   // 512: tree parents (leaves 0..n-1, internal n..2n-2)
    if (n < 2) return 0;
    // two-queue Huffman: leaves in `order`, internal node weights in wq[] (monotone)
    uint32_t *wq = (uint32_t *)(work + 768);  // up to 255 internal weights (needs 510 u16 -> within 1280)
    uint32_t li = 0, qi = 0, qn = 0;
    for (uint32_t k = 0; k + 1 < n; k++) {
        uint32_t w = 0;
        for (int t = 0; t < 2; t++) {
            bool leaf = li < n && (qi >= qn || counts[order[li]] <= wq[qi]);
            if (leaf) { w += counts[order[li]]; parent[li] = (uint16_t)(n + k); li++; }
            else { w += wq[qi]; parent[n + qi] = (uint16_t)(n + k); qi++; }
        }
        wq[qn++] = w;
    }
    // depths: root = node 2n-2 has depth 0; walk internal nodes from the root down (parents have larger index)
    parent[2 * n - 2] = 0;
    for (int32_t i = (int32_t)(2 * n - 3); i >= (int32_t)n; i--) parent[i] = (uint16_t)(parent[parent[i]] + 1);  // now holds depth
    uint32_t maxlen = 0;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t d = parent[parent[i]] + 1u;
        if (d > max_bits) d = max_bits;
        lens[order[i]] = (uint8_t)d;
        if (d > maxlen) maxlen = d;
    }
    // Kraft repair in units of 2^-max_bits
    uint32_t kraft = 0;
    for (uint32_t i = 0; i < n; i++) kraft += 1u << (max_bits - lens[order[i]]);
    const uint32_t full = 1u << max_bits;
    while (kraft > full) {  // over-subscribed: lengthen the rarest symbol that is still shorter than max_bits
        for (uint32_t i = 0; i < n; i++) {
            uint8_t &l = lens[order[i]];
            if (l < max_bits) { kraft -= 1u << (max_bits - l - 1); l++; break; }
        }
    }
    while (kraft < full) {  // slack: shorten the most frequent symbol whose gain fits
        bool moved = false;
        for (int32_t i = (int32_t)n - 1; i >= 0; i--) {
            uint8_t &l = lens[order[i]];
            if (l > 1 && kraft + (1u << (max_bits - l)) <= full) { kraft += 1u << (max_bits - l); l--; moved = true; break; }
        }
        if (!moved) break;
    }
    maxlen = 0;
    for (uint32_t i = 0; i < n; i++) if (lens[order[i]] > maxlen) maxlen = lens[order[i]];
    return kraft == full ? maxlen : 0;
}

// canonical codes as the zstd decoder derives them: weights ascending, symbols ascending
ZHDN void huf_assign_codes(uint16_t *codes, const uint8_t *lens, uint32_t maxlen) {
    uint32_t idx = 0;
    for (uint32_t w = 1; w <= maxlen; w++) {  // weight w <=> length maxlen + 1 - w
        uint32_t len = maxlen + 1 - w;
        for (uint32_t s = 0; s < 256; s++)
            if (lens[s] == len) { codes[s] = (uint16_t)(idx >> (w - 1)); idx += 1u << (w - 1); }
    }
}

// Huffman tree description (RFC 8878 4.2.1).  Returns bytes written, 0 if it cannot be described.
// scratch: FseCTable-sized + 512 bytes, passed in by the caller.
ZHDN uint32_t huf_write_tree(uint8_t *dst, const uint8_t *lens, uint32_t maxlen, FseCTable *ct, uint8_t *scratch) {
    int32_t last = 255;
    while (last >= 0 && lens[last] == 0) last--;
    if (last < 1) return 0;
    const uint32_t nw = (uint32_t)last;  // weights listed for symbols 0..last-1; the last one is implied
    uint8_t w8[256];
    uint32_t wcount[13];
    for (uint32_t i = 0; i < 13; i++) wcount[i] = 0;
    for (uint32_t s = 0; s < nw; s++) { w8[s] = lens[s] ? (uint8_t)(maxlen + 1 - lens[s]) : 0; wcount[w8[s]]++; }
    // try FSE-compressed weights first (mandatory when nw > 128)
    uint32_t fse_size = 0;
    uint32_t max_w = 12;
    while (max_w > 0 && wcount[max_w] == 0) max_w--;
    uint32_t distinct = 0;
    for (uint32_t i = 0; i <= max_w; i++) distinct += wcount[i] != 0;
    if (distinct >= 2 && nw > 1) {
        uint32_t tl = fse_table_log(6, nw, max_w);
        int16_t norm[13];
        if (fse_normalize(norm, tl, wcount, nw, max_w) == 0) {
            uint8_t *p = dst + 1;
            uint32_t hs = fse_write_ncount(p, norm, max_w, tl);
            fse_build_ctable(ct, norm, max_w, tl, scratch);
            BackWriter bw;
            bw_init(&bw, p + hs);
            // two interleaved states: even index -> state 1, odd index -> state 2, walking backwards
            int32_t i = (int32_t)nw - 1;
            uint32_t s1, s2;
            if (nw & 1) { s1 = fse_init_state(ct, w8[i--]); s2 = fse_init_state(ct, w8[i--]); }
            else { s2 = fse_init_state(ct, w8[i--]); s1 = fse_init_state(ct, w8[i--]); }
            for (; i >= 0; i--) {
                if (i & 1) s2 = fse_encode(ct, &bw, s2, w8[i]);
                else s1 = fse_encode(ct, &bw, s1, w8[i]);
            }
            fse_flush_state(ct, &bw, s2);
            fse_flush_state(ct, &bw, s1);
            uint32_t bs = bw_finish(&bw);
            fse_size = hs + bs;
            if (fse_size < 128 && (nw > 128 || fse_size < (nw + 1) / 2)) { dst[0] = (uint8_t)fse_size; return 1 + fse_size; }
        }
    }
    if (nw > 128) return 0;
    dst[0] = (uint8_t)(127 + nw);
    for (uint32_t s = 0; s < nw; s += 2) dst[1 + s / 2] = (uint8_t)(w8[s] << 4 | (s + 1 < nw ? w8[s + 1] : 0));
    return 1 + (nw + 1) / 2;
}

}  // namespace zc
