// Serial (single lane) zstd compressed-block writer: literals section + sequences section from a
// parsed block (literal bytes + sequence list).  __host__ __device__: run by one GPU lane per block in
// K3's entropy stage, and by the CPU test harness against stock libzstd.
// Format: RFC 8878 3.1.1.3.
#pragma once
#include "zstd_core.h"

namespace zc {

struct Seq {
    uint32_t ll;        // literal length before the match
    uint32_t ml;        // match length (>= 3)
    uint32_t off_base;  // 1..3 = repeat-offset codes, >= 4 = real offset + 3
};

struct EncWork {  // per-block scratch (~9 KB)
    uint32_t counts[256];
    uint8_t lens[256];
    uint16_t codes[256];
    uint16_t hwork[1280];
    FseCTable ct[3];  // LL, OF, ML
    uint8_t spread[512];
    uint32_t hist[3][64];
    int16_t norm[3][64];
};

// ---- literals -------------------------------------------------------------------------------------
ZHDN uint32_t write_raw_literals(uint8_t *dst, const uint8_t *lits, uint32_t n, bool rle) {
    uint32_t h;
    const uint32_t type = rle ? 1 : 0;
    if (n <= 31) { dst[0] = (uint8_t)(type | n << 3); h = 1; }
    else if (n <= 4095) { uint32_t v = type | 1u << 2 | n << 4; dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); h = 2; }
    else { uint32_t v = type | 3u << 2 | n << 4; dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); h = 3; }
    if (rle) { dst[h] = lits[0]; return h + 1; }
    for (uint32_t i = 0; i < n; i++) dst[h + i] = lits[i];
    return h + n;
}

ZHDN uint32_t huf_encode_stream(uint8_t *dst, const uint8_t *sym, uint32_t n, const uint16_t *codes, const uint8_t *lens) {
    BackWriter w;
    bw_init(&w, dst);
    for (int32_t i = (int32_t)n - 1; i >= 0; i--) bw_put(&w, codes[sym[i]], lens[sym[i]]);
    return bw_finish(&w);
}

// returns bytes written; never more than the raw form would take (falls back to raw / RLE literals)
ZHDN uint32_t write_literals(uint8_t *dst, const uint8_t *lits, uint32_t n, EncWork *wk) {
    if (n == 0) return write_raw_literals(dst, lits, 0, false);
    for (uint32_t s = 0; s < 256; s++) wk->counts[s] = 0;
    for (uint32_t i = 0; i < n; i++) wk->counts[lits[i]]++;
    uint32_t used = 0, maxc = 0;
    for (uint32_t s = 0; s < 256; s++) { used += wk->counts[s] != 0; if (wk->counts[s] > maxc) maxc = wk->counts[s]; }
    if (used == 1) return write_raw_literals(dst, lits, n, true);
    if (n < 64) return write_raw_literals(dst, lits, n, false);
    const uint32_t maxlen = huf_build_lengths(wk->lens, wk->counts, Z_HUF_MAXBITS, wk->hwork);
    if (!maxlen) return write_raw_literals(dst, lits, n, false);
    uint64_t bits = 0;
    for (uint32_t s = 0; s < 256; s++) bits += (uint64_t)wk->counts[s] * wk->lens[s];
    const uint32_t est = (uint32_t)((bits + 7) / 8);
    if (est + 16 + (used + 1) / 2 >= n - (n >> 6)) return write_raw_literals(dst, lits, n, false);
    huf_assign_codes(wk->codes, wk->lens, maxlen);
    const uint32_t lh = 3 + (n >= 1024) + (n >= 16384);
    const bool single = n < 256;
    uint8_t *p = dst + lh;
    const uint32_t tree = huf_write_tree(p, wk->lens, maxlen, &wk->ct[0], wk->spread);
    if (!tree) return write_raw_literals(dst, lits, n, false);
    p += tree;
    if (single) {
        p += huf_encode_stream(p, lits, n, wk->codes, wk->lens);
    } else {
        const uint32_t seg = (n + 3) / 4;
        uint8_t *jump = p;
        p += 6;
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t start = k * seg, cnt = k < 3 ? seg : n - 3 * seg;
            const uint32_t sz = huf_encode_stream(p, lits + start, cnt, wk->codes, wk->lens);
            if (k < 3) {
                if (sz > 0xFFFF) return write_raw_literals(dst, lits, n, false);
                jump[2 * k] = (uint8_t)sz; jump[2 * k + 1] = (uint8_t)(sz >> 8);
            }
            p += sz;
        }
    }
    const uint32_t csize = (uint32_t)(p - dst) - lh;
    if (csize >= n) return write_raw_literals(dst, lits, n, false);
    // Literals_Section_Header: type 2 (Compressed) | size_format | regenerated size | compressed size
    if (lh == 3) {
        uint32_t v = 2u | (single ? 0u : 1u) << 2 | n << 4 | csize << 14;
        dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16);
    } else if (lh == 4) {
        uint32_t v = 2u | 2u << 2 | n << 4 | csize << 18;
        dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); dst[3] = (uint8_t)(v >> 24);
    } else {
        uint64_t v = 2u | 3u << 2 | (uint64_t)n << 4 | (uint64_t)csize << 22;
        for (int i = 0; i < 5; i++) dst[i] = (uint8_t)(v >> (8 * i));
    }
    return lh + csize;
}

// ---- sequences ------------------------------------------------------------------------------------
// picks Predefined / RLE / FSE_Compressed for one symbol stream; writes the table description (if any)
// at *p and builds the encode table.  Returns the mode.
ZHDN uint32_t choose_table(uint8_t **p, FseCTable *ct, int16_t *norm, const uint32_t *hist, uint32_t nseq, uint32_t max_sym,
                           const int16_t *defnorm, uint32_t def_max_sym, uint32_t def_log, uint32_t max_log, uint8_t *spread) {
    uint32_t used = 0, only = 0;
    for (uint32_t s = 0; s <= max_sym; s++) if (hist[s]) { used++; only = s; }
    if (used == 1 && nseq > 2) {  // RLE: one byte, zero state bits
        *(*p)++ = (uint8_t)only;
        for (uint32_t s = 0; s <= max_sym; s++) norm[s] = 0;
        norm[only] = (int16_t)(1 << 5);
        ct->tl = 0;
        return 1;
    }
    uint64_t cost_def = max_sym <= def_max_sym ? fse_cost(defnorm, def_log, hist, max_sym) : ~0ull >> 2;
    uint64_t cost_fse = ~0ull >> 2;
    uint32_t tl = 0;
    if (used >= 2 && nseq >= 16) {
        tl = fse_table_log(max_log, nseq, max_sym);
        if (fse_normalize(norm, tl, hist, nseq, max_sym) == 0)
            cost_fse = fse_cost(norm, tl, hist, max_sym) + ((uint64_t)(max_sym + 4) * 5 / 8 + 1) * 8 * 256;  // + header estimate
    }
    if (cost_fse < cost_def) {
        *p += fse_write_ncount(*p, norm, max_sym, tl);
        fse_build_ctable(ct, norm, max_sym, tl, spread);
        return 2;
    }
    int16_t dn[64];
    for (uint32_t s = 0; s <= def_max_sym; s++) dn[s] = defnorm[s];
    fse_build_ctable(ct, dn, def_max_sym, def_log, spread);
    return 0;
}

ZHDN uint32_t write_sequences(uint8_t *dst, const Seq *seqs, uint32_t nseq, EncWork *wk) {
    uint8_t *p = dst;
    if (nseq < 128) *p++ = (uint8_t)nseq;
    else if (nseq < 0x7F00) { *p++ = (uint8_t)((nseq >> 8) + 128); *p++ = (uint8_t)nseq; }
    else { *p++ = 255; *p++ = (uint8_t)(nseq - 0x7F00); *p++ = (uint8_t)((nseq - 0x7F00) >> 8); }
    if (nseq == 0) return (uint32_t)(p - dst);
    for (int t = 0; t < 3; t++) for (int s = 0; s < 64; s++) wk->hist[t][s] = 0;
    uint32_t max_ll = 0, max_of = 0, max_ml = 0;
    for (uint32_t i = 0; i < nseq; i++) {
        uint32_t a = ll_code(seqs[i].ll), b = highbit(seqs[i].off_base), c = ml_code(seqs[i].ml);
        wk->hist[0][a]++; wk->hist[1][b]++; wk->hist[2][c]++;
        if (a > max_ll) max_ll = a;
        if (b > max_of) max_of = b;
        if (c > max_ml) max_ml = c;
    }
    uint8_t *modes = p++;
    const uint32_t m_ll = choose_table(&p, &wk->ct[0], wk->norm[0], wk->hist[0], nseq, max_ll, ZTAB(LL_defnorm), 35, 6, Z_LL_MAXLOG, wk->spread);
    const uint32_t m_of = choose_table(&p, &wk->ct[1], wk->norm[1], wk->hist[1], nseq, max_of, ZTAB(OF_defnorm), 28, 5, Z_OF_MAXLOG, wk->spread);
    const uint32_t m_ml = choose_table(&p, &wk->ct[2], wk->norm[2], wk->hist[2], nseq, max_ml, ZTAB(ML_defnorm), 52, 6, Z_ML_MAXLOG, wk->spread);
    *modes = (uint8_t)(m_ll << 6 | m_of << 4 | m_ml << 2);
    const FseCTable *cl = &wk->ct[0], *co = &wk->ct[1], *cm = &wk->ct[2];
    BackWriter w;
    bw_init(&w, p);
    // last sequence first: its codes only initialise the three states
    uint32_t i = nseq - 1;
    uint32_t lc = ll_code(seqs[i].ll), oc = highbit(seqs[i].off_base), mc = ml_code(seqs[i].ml);
    uint32_t s_ml = m_ml == 1 ? 0 : fse_init_state(cm, mc);
    uint32_t s_of = m_of == 1 ? 0 : fse_init_state(co, oc);
    uint32_t s_ll = m_ll == 1 ? 0 : fse_init_state(cl, lc);
    bw_put(&w, seqs[i].ll - ZTAB(LL_base)[lc], ZTAB(LL_bits)[lc]);
    bw_put(&w, seqs[i].ml - ZTAB(ML_base)[mc], ZTAB(ML_bits)[mc]);
    bw_put(&w, seqs[i].off_base - (1u << oc), oc);
    while (i-- > 0) {
        lc = ll_code(seqs[i].ll); oc = highbit(seqs[i].off_base); mc = ml_code(seqs[i].ml);
        if (m_of != 1) s_of = fse_encode(co, &w, s_of, oc);
        if (m_ml != 1) s_ml = fse_encode(cm, &w, s_ml, mc);
        if (m_ll != 1) s_ll = fse_encode(cl, &w, s_ll, lc);
        bw_put(&w, seqs[i].ll - ZTAB(LL_base)[lc], ZTAB(LL_bits)[lc]);
        bw_put(&w, seqs[i].ml - ZTAB(ML_base)[mc], ZTAB(ML_bits)[mc]);
        bw_put(&w, seqs[i].off_base - (1u << oc), oc);
    }
    if (m_ml != 1) fse_flush_state(cm, &w, s_ml);
    if (m_of != 1) fse_flush_state(co, &w, s_of);
    if (m_ll != 1) fse_flush_state(cl, &w, s_ll);
    p += bw_finish(&w);
    return (uint32_t)(p - dst);
}

// Worst-case body size the writers above can produce before the caller compares against the raw size.
ZHD uint32_t block_body_bound(uint32_t nlits, uint32_t nseq) { return nlits + 8 + 4 + 3 * 260 + nseq * 8 + 16; }

// Whole compressed-block body.  dst must hold block_body_bound(nlits, nseq) bytes.
ZHDN uint32_t write_block_body(uint8_t *dst, const uint8_t *lits, uint32_t nlits, const Seq *seqs, uint32_t nseq, EncWork *wk) {
    uint32_t n = write_literals(dst, lits, nlits, wk);
    n += write_sequences(dst + n, seqs, nseq, wk);
    return n;
}

}  // namespace zc
