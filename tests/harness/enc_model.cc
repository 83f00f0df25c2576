// CPU model of the K3 encoder (test/dev tool, NOT product path): the same per-position bucketed
// match search + lazy parse + block writer (zstd_enc_block.h) the CUDA kernels implement, run
// serially so ratio and stock-decodability can be checked without a GPU.
//   g++ -O2 -shared -fPIC -I squishrs_b200/csrc tests/harness/enc_model.cc -o tests/harness/libencmodel.so
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "zstd_enc_block.h"

using namespace zc;

struct Params {
    int hash_log;      // rows
    int row_entries;   // entries per row
    int min_match;     // 4 or 5 (hash bytes)
    int lazy_depth;    // 0,1,2
    int rep_mode;      // 0 none (post-hoc only), 1 neighbour-offset approximation, 2 exact
    int tile;          // insertion granularity (GPU inserts a whole tile before searching it)
    int target_len;    // stop lazy when match >= this
    int alt_window;    // how many previous positions' offsets are tried as rep candidates (rep_mode 1)
    int sel_mul;       // 0: longest wins; else maximise len*sel_mul - log2(offset)
    int accept_thr;    // a match is used only if len*4 - log2(off_base) >= accept_thr
    int skip_stride;   // 0 = search every position; S = search p%S==0 first, others inherit if the anchor's match still has >= skip_min left
    int skip_min;
    int precheck;      // 1: a candidate must match the 8 bytes ending at the current best length to be examined; 2: + inherit from anchor position (stride 4)
};

static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t hashN(const uint8_t *p, int mm, int hl) {
    uint64_t v = rd64(p);
    if (mm == 4) return (uint32_t)((uint32_t)v * 2654435761u) >> (32 - hl);
    if (mm == 5) return (uint32_t)(((v << 24) * 889523592379ULL) >> (64 - hl));
    return (uint32_t)(((v << 16) * 227718039650203ULL) >> (64 - hl));
}
static inline uint32_t match_len(const uint8_t *s, uint32_t a, uint32_t b, uint32_t n) {  // a > b
    uint32_t l = 0;
    while (a + l + 8 <= n) {
        uint64_t x = rd64(s + a + l) ^ rd64(s + b + l);
        if (x) return l + (__builtin_ctzll(x) >> 3);
        l += 8;
    }
    while (a + l < n && s[a + l] == s[b + l]) l++;
    return l;
}

struct Match { uint32_t len, off; };

struct Model {
    const uint8_t *s; uint32_t n; Params P;
    std::vector<Match> best, alt;
    std::vector<uint8_t> back, capflag;
    std::vector<uint32_t> srcpos, known;
    uint64_t nskip = 0, nsearch = 0, nverify = 0, ncheap = 0;
    void search() {
        best.assign(n, {0, 0}); alt.assign(n, {0, 0}); back.assign(n, 0); capflag.assign(n, 0); srcpos.assign(n, 0); known.assign(n, 0);
        static const bool cont = getenv("ENC_CONT") != nullptr;
        static const uint32_t cont_fresh = getenv("ENC_CONT_FRESH") ? (uint32_t)atoi(getenv("ENC_CONT_FRESH")) : 16;
        std::vector<uint32_t> prevc, curc;
        int run_e = -1000; uint32_t run_end = 0, run_off = 0, run_src = 0; bool run_cap = false;
        const uint32_t rows = 1u << P.hash_log, K = (uint32_t)P.row_entries;
        std::vector<uint32_t> tab((size_t)rows * K, 0), head(rows, 0);
        const uint32_t T = (uint32_t)P.tile;
        for (uint32_t t0 = 0; t0 < n; t0 += T) {
            uint32_t t1 = std::min(n, t0 + T);
            // insert the whole tile first (ring per row), as the GPU does
            static const uint32_t ins_stride = getenv("ENC_INS_STRIDE") ? (uint32_t)atoi(getenv("ENC_INS_STRIDE")) : 1;
            static const uint32_t search_stride = getenv("ENC_SEARCH_STRIDE") ? (uint32_t)atoi(getenv("ENC_SEARCH_STRIDE")) : 1;
            static const bool slot_by_pos = getenv("ENC_SLOT_BY_POS") != nullptr;
            for (uint32_t p = t0; p < t1 && p + 8 <= n; p++) {
                if (p % ins_stride) continue;
                uint32_t h = hashN(s + p, P.min_match, P.hash_log);
                tab[(size_t)h * K + (slot_by_pos ? (p / ins_stride) % K : head[h]++ % K)] = p + 1;
            }
            for (int round = 0; round < 2; round++)
            for (uint32_t p = t0; p < t1 && p + 8 <= n; p++) {
                if (P.precheck == 2 && !P.skip_stride) { bool anchor_pos = (p & 3) == 0; if ((round == 0) != anchor_pos) continue; }
                else if (P.skip_stride) {
                    bool anchor_pos = (p % P.skip_stride) == 0;
                    if ((round == 0) != anchor_pos) continue;
                    if (!anchor_pos) {
                        uint32_t q = p - (p % P.skip_stride);
                        static const bool skipcap = getenv("ENC_SKIPCAP") != nullptr;
                        if (skipcap) {  // anchor's match is capped: verify only the inherited candidate (one pair) instead of the row
                            if (q >= t0 && best[q].len >= (uint32_t)P.target_len) {
                                uint32_t l = match_len(s, p, p - best[q].off, n); if (l > (uint32_t)P.target_len) l = P.target_len;
                                best[p] = {l, best[q].off}; nskip++; nverify++; continue;
                            }
                        } else
                        if (q >= t0 && best[q].len >= (p - q) + (uint32_t)P.skip_min) { best[p] = {best[q].len - (p - q), best[q].off}; nskip++; continue; }
                    }
                } else if (round && P.precheck != 2) continue;
                nsearch++;
                uint32_t h = hashN(s + p, P.min_match, P.hash_log);
                Match b = {0, 0};
                int exttop = getenv("ENC_EXTTOP") ? atoi(getenv("ENC_EXTTOP")) : 0;
                uint32_t near_off[8]; int nnear = 0;
                if (exttop) {  // the T nearest candidates whose first 8 bytes match
                    for (uint32_t k = 0; k < K; k++) {
                        uint32_t e = tab[(size_t)h * K + k];
                        if (!e || e - 1 >= p) continue;
                        uint32_t c = e - 1;
                        if (rd64(s + c) != rd64(s + p)) continue;
                        uint32_t off = p - c; int i = nnear < exttop ? nnear++ : exttop;
                        if (i == exttop) { if (off >= near_off[exttop - 1]) continue; i = exttop - 1; }
                        near_off[i] = off;
                        while (i > 0 && near_off[i] < near_off[i - 1]) { std::swap(near_off[i], near_off[i - 1]); i--; }
                    }
                }
                if (P.precheck == 2 && (p & 3) && (p & ~3u) >= t0) {
                    uint32_t q = p & ~3u;
                    if (best[q].len >= (p - q) + 8) b = {best[q].len - (p - q), best[q].off};
                }
                if (cont) { prevc.swap(curc); curc.clear(); if (p == t0) prevc.clear(); }
                for (uint32_t k = 0; k < K; k++) {
                    if (p % search_stride) break;
                    uint32_t e = tab[(size_t)h * K + k];
                    if (!e || e - 1 >= p) continue;
                    uint32_t c = e - 1;
                    if (rd32(s + c) != rd32(s + p) || (cont && (rd64(s + c) ^ rd64(s + p)) << 16)) continue;
                    if (cont) {
                        curc.push_back(c);
                        static const int cont_mode = getenv("ENC_CONT_MODE") ? atoi(getenv("ENC_CONT_MODE")) : 0;
                        bool is_cont = c > 0 && std::find(prevc.begin(), prevc.end(), c - 1) != prevc.end();
                        if (cont_mode == 1) is_cont = is_cont && p > 0 && best[p - 1].len && best[p - 1].off == p - c;  // only the represented chain is dropped
                        if (((p - t0) % cont_fresh) != 0 && is_cont) { ncheap++; continue; }
                    }
                    if (P.precheck && b.len >= 16 && p + b.len <= n && rd64(s + c + b.len - 8) != rd64(s + p + b.len - 8)) { ncheap++; continue; }
                    uint32_t l = match_len(s, p, c, n);
                    if (exttop && l > 8) { bool keep = false; for (int i = 0; i < nnear; i++) keep |= near_off[i] == p - c; if (!keep) l = 8; }
                    if (l > (uint32_t)P.target_len && getenv("ENC_CAPLEN")) l = P.target_len;
                    nverify++;
                    if (p > 0 && c > 0 && s[p - 1] == s[c - 1]) ncheap++;  // continuation of a match that started earlier
                    uint32_t off = p - c;
                    if (P.sel_mul == 0) { if (l > b.len || (l == b.len && off < b.off)) b = {l, off}; }
                    else if (b.len == 0 || (int)l * P.sel_mul - (int)highbit(off + 3) > (int)b.len * P.sel_mul - (int)highbit(b.off + 3)) b = {l, off};
                }
                if (b.len < (uint32_t)P.min_match) b = {0, 0};
                if (cont) {  // mirrors lz_search_kernel: max-scan of (2*end - log2(off)) over the 32-position group, expired winner -> own
                    srcpos[p] = p;
                    if (((p - t0) % 32) == 0) run_e = -1000;
                    const bool own_cap = b.len >= (uint32_t)P.target_len;
                    const int eo = b.len ? 2 * (int)(p + b.len) - (int)highbit(b.off + 3) : -1000;
                    const Match own = b;
                    known[p] = b.len; capflag[p] = own_cap;
                    if (run_e > eo && run_end >= p + (uint32_t)P.min_match) {
                        known[p] = run_end - p; capflag[p] = run_cap; srcpos[p] = run_src;
                        b = {run_cap ? (uint32_t)P.target_len : run_end - p, run_off};
                    }
                    if (own.len && eo >= run_e) { run_e = eo; run_end = p + own.len; run_off = own.off; run_cap = own_cap; run_src = p; }
                }
                best[p] = b;
            }
        }
        if (P.rep_mode == 4) {  // dominant ("stride") offsets per tile: the top-2 most frequent best offsets, tested everywhere
            std::vector<uint32_t> offs;
            for (uint32_t t0 = 0; t0 < n; t0 += T) {
                uint32_t t1 = std::min(n, t0 + T);
                offs.clear();
                uint32_t w0 = t0 >= (uint32_t)P.alt_window * T ? t0 - P.alt_window * T : 0;
                for (uint32_t p = w0; p < t1; p++) if (best[p].len) offs.push_back(best[p].off);
                std::sort(offs.begin(), offs.end());
                uint32_t top[2] = {0, 0}, cnt[2] = {0, 0};
                for (size_t i = 0; i < offs.size();) {
                    size_t j = i; while (j < offs.size() && offs[j] == offs[i]) j++;
                    uint32_t c = (uint32_t)(j - i);
                    if (c > cnt[0]) { top[1] = top[0]; cnt[1] = cnt[0]; top[0] = offs[i]; cnt[0] = c; }
                    else if (c > cnt[1]) { top[1] = offs[i]; cnt[1] = c; }
                    i = j;
                }
                for (uint32_t p = t0; p < t1; p++) {
                    Match a = {0, 0};
                    for (int k = 0; k < 2; k++) {
                        uint32_t o = top[k];
                        if (!o || cnt[k] < 4 || o > p || o == best[p].off) continue;
                        uint32_t l = match_len(s, p, p - o, n);
                        if (l >= 3 && l > a.len) a = {l, o};
                    }
                    alt[p] = a;
                }
            }
        }
        // backward extension of each best match (bounded), and neighbour-offset alternates
        for (uint32_t p = 0; p < n; p++) {
            if (best[p].len && cont && !getenv("ENC_CONT_TRUEBACK")) back[p] = (uint8_t)std::min<uint32_t>(3, p - srcpos[p]);
            else if (best[p].len) {
                uint32_t o = best[p].off, k = 0;
                while (k < (uint32_t)(getenv("ENC_BACK") ? atoi(getenv("ENC_BACK")) : 15) && p > k && p - k > o && s[p - k - 1] == s[p - k - 1 - o]) k++;
                back[p] = (uint8_t)k;
            }
            if (P.rep_mode == 3 || P.rep_mode == 5) {  // last two distinct valid offsets seen before p (scan-friendly on the GPU)
                Match a = {0, 0};
                uint32_t seen[2] = {0, 0}; int ns = 0;
                for (int d = 1; d <= P.alt_window && (uint32_t)d <= p && ns < 2; d++) {
                    uint32_t o = best[p - d].off;
                    if (!o || (ns && o == seen[0])) continue;
                    // mode 5: only positions that look like match STARTS (not the tail of the previous position's match)
                    if (P.rep_mode == 5 && p - d > 0 && best[p - d - 1].len > best[p - d].len && best[p - d - 1].off == o) continue;
                    if (P.rep_mode == 5 && p - d > 0 && best[p - d - 1].len >= best[p - d].len + 1) continue;
                    seen[ns++] = o;
                }
                for (int k = 0; k < ns; k++) {
                    uint32_t o = seen[k];
                    if (o == best[p].off || o > p) continue;
                    uint32_t l = match_len(s, p, p - o, n);
                    if (l >= 3 && l > a.len) a = {l, o};
                }
                alt[p] = a;
            }
            if (P.rep_mode == 1) {
                Match a = {0, 0};
                for (int d = 1; d <= P.alt_window && (uint32_t)d <= p; d++) {
                    uint32_t o = best[p - d].off;
                    if (!o || o == best[p].off || o == a.off || o > p) continue;
                    uint32_t l = match_len(s, p, p - o, n);
                    if (l >= 3 && l > a.len) a = {l, o};
                }
                alt[p] = a;
            }
        }
    }
};

struct Cand { uint32_t len, off_base, off; int32_t score; };

static inline int32_t score_of(uint32_t len, uint32_t off_base) { return (int32_t)len * 4 - (int32_t)highbit(off_base); }

// repcode bookkeeping (RFC 8878 3.1.1.5)
static inline uint32_t rep_code_for(uint32_t off, const uint32_t rep[3], bool ll0) {
    if (!ll0) { if (off == rep[0]) return 1; if (off == rep[1]) return 2; if (off == rep[2]) return 3; }
    else { if (off == rep[1]) return 1; if (off == rep[2]) return 2; if (rep[0] > 1 && off == rep[0] - 1) return 3; }
    return 0;
}
static inline void rep_update(uint32_t rep[3], uint32_t off_base, bool ll0) {
    if (off_base > 3) { rep[2] = rep[1]; rep[1] = rep[0]; rep[0] = off_base - 3; return; }
    uint32_t idx = off_base - 1 + (ll0 ? 1 : 0);
    if (idx == 0) return;
    uint32_t v = idx == 3 ? rep[0] - 1 : rep[idx];
    if (idx >= 2) rep[2] = rep[1];
    rep[1] = rep[0];
    rep[0] = v;
}

extern "C" long enc_model_frame(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t cap, const Params *Pp, uint32_t *stats) {
    Params P = *Pp;
    Model M{src, n, P};
    M.search();
    uint8_t *o = dst;
    // frame header
    *o++ = 0x28; *o++ = 0xB5; *o++ = 0x2F; *o++ = 0xFD;
    if (n <= 255) { *o++ = 0x20; *o++ = (uint8_t)n; }
    else if (n <= 65791) { *o++ = 0x60; uint32_t v = n - 256; *o++ = (uint8_t)v; *o++ = (uint8_t)(v >> 8); }
    else { *o++ = 0xA0; for (int i = 0; i < 4; i++) *o++ = (uint8_t)(n >> (8 * i)); }
    uint32_t rep[3] = {1, 4, 8};
    std::vector<Seq> seqs;
    std::vector<uint8_t> lits, body;
    EncWork *wk = new EncWork;
    uint32_t total_seq = 0, total_lit = 0, rep_seq = 0, of_bits = 0, ml_sum = 0;
    uint32_t nblocks = n ? (n + Z_BLOCK_MAX - 1) / Z_BLOCK_MAX : 1;
    const uint32_t win_w = getenv("ENC_WINDOW") ? (uint32_t)atoi(getenv("ENC_WINDOW")) : 0;  // sequential-parser study: steps of W positions
    uint32_t win_lo = ~0u - 64, win_steps = 0;
    for (uint32_t b = 0; b < nblocks; b++) {
        const uint32_t bs = b * Z_BLOCK_MAX, be = std::min(n, bs + Z_BLOCK_MAX);
        seqs.clear(); lits.clear();
        uint32_t p = bs, anchor = bs;
        uint32_t rep_save[3] = {rep[0], rep[1], rep[2]};
        if (getenv("ENC_REP_INVALIDATE") && b > 0) { rep[0] = rep[1] = rep[2] = 0; }
        auto pick = [&](uint32_t q, uint32_t anchor_) -> Cand {
            Cand c = {0, 0, 0, -1000000};
            if (q >= be) return c;
            if (win_w && !(q >= win_lo && q < win_lo + win_w)) { win_lo = q; win_steps++; }
            const bool ll0 = (q == anchor_);
            auto consider = [&](uint32_t len, uint32_t off) {
                if (!len || off > q) return;
                if (q + len > be) len = be - q;
                uint32_t rc = P.rep_mode == 6 ? 0 : rep_code_for(off, rep, ll0);  // mode 6: decisions are rep-blind (parallelisable)
                uint32_t ob = rc ? rc : off + 3;
                uint32_t minl = rc ? 3u : (uint32_t)P.min_match;
                if (len < minl) return;
                int32_t sc = score_of(len, ob);
                if (sc < P.accept_thr) return;
                if (sc > c.score) c = {len, ob, off, sc};
            };
            consider(M.best[q].len, M.best[q].off);
            if (P.rep_mode == 1 || P.rep_mode >= 3) consider(M.alt[q].len, M.alt[q].off);
            if (P.rep_mode == 2) {
                for (int r = 0; r < 3; r++) {
                    uint32_t off = rep[r];
                    if (off && off <= q) consider(match_len(src, q, q - off, n), off);
                }
                if (ll0 && rep[0] > 1 && rep[0] - 1 <= q) consider(match_len(src, q, q - (rep[0] - 1), n), rep[0] - 1);
            }
            return c;
        };
        while (p < be) {
            Cand cur = pick(p, anchor);
            if (cur.len == 0) { p++; continue; }
            uint32_t start = p;
            if (P.lazy_depth > 0 && cur.len < (uint32_t)P.target_len) {
                for (;;) {
                    bool improved = false;
                    for (int d = 1; d <= P.lazy_depth; d++) {
                        Cand c2 = pick(start + d, anchor);
                        if (c2.len && c2.score > cur.score + (d == 1 ? 4 : 7)) { cur = c2; start += d; improved = true; break; }
                    }
                    if (!improved || cur.len >= (uint32_t)P.target_len) break;
                    if (getenv("ENC_MAXSHIFT") && start - p + 2 > (uint32_t)atoi(getenv("ENC_MAXSHIFT"))) break;
                }
            }
            if (getenv("ENC_CONT") && cur.off_base > 3 && M.best[start].off == cur.off && M.capflag[start]) { uint32_t full = match_len(src, start, start - cur.off, n); if (start + full > be) full = be - start; cur.len = full; }
            else if (getenv("ENC_CAPLEN") && cur.len + 3 >= (uint32_t)P.target_len) { uint32_t full = match_len(src, start, start - cur.off, n); if (start + full > be) full = be - start; if (full > cur.len) cur.len = full; }
            // backward extension (only for non-rep matches found by the search; bounded by literal run)
            if (cur.off_base > 3 && M.best[start].off == cur.off && M.best[start].len) {
                uint32_t k = std::min<uint32_t>(M.back[start], start - anchor);
                start -= k; cur.len += k;
            }
            const bool ll0 = start == anchor;
            // the rep code depends on ll0 at the final start
            uint32_t rc = rep_code_for(cur.off, rep, ll0);
            uint32_t ob = rc ? rc : cur.off + 3;
            Seq sq = {start - anchor, cur.len, ob};
            lits.insert(lits.end(), src + anchor, src + start);
            seqs.push_back(sq);
            if (getenv("ENC_DUMP") && total_seq + seqs.size() < 60) fprintf(stderr, "%u ll %u ml %u ob %u off %u | best(len %u off %u) alt(len %u off %u)\n", start, sq.ll, sq.ml, ob, cur.off, M.best[start].len, M.best[start].off, M.alt[start].len, M.alt[start].off);
            if (ob <= 3) rep_seq++;
            of_bits += highbit(ob); ml_sum += cur.len;
            rep_update(rep, ob, ll0);
            p = start + cur.len;
            anchor = p;
        }
        lits.insert(lits.end(), src + anchor, src + be);
        body.resize(block_body_bound((uint32_t)lits.size(), (uint32_t)seqs.size()) + 64);
        uint32_t bl = write_block_body(body.data(), lits.data(), (uint32_t)lits.size(), seqs.data(), (uint32_t)seqs.size(), wk);
        const uint32_t blen = be - bs;
        const bool last = b + 1 == nblocks;
        if ((size_t)(o - dst) + 3 + std::max(bl, blen) > cap) { delete wk; return -1; }
        if (bl >= blen || blen == 0) {  // raw block: repeat-offset history is untouched
            uint32_t h = (last ? 1u : 0u) | 0u << 1 | blen << 3;
            *o++ = (uint8_t)h; *o++ = (uint8_t)(h >> 8); *o++ = (uint8_t)(h >> 16);
            memcpy(o, src + bs, blen); o += blen;
            rep[0] = rep_save[0]; rep[1] = rep_save[1]; rep[2] = rep_save[2];
        } else {
            uint32_t h = (last ? 1u : 0u) | 2u << 1 | bl << 3;
            *o++ = (uint8_t)h; *o++ = (uint8_t)(h >> 8); *o++ = (uint8_t)(h >> 16);
            memcpy(o, body.data(), bl); o += bl;
            total_seq += (uint32_t)seqs.size(); total_lit += (uint32_t)lits.size();
        }
    }
    delete wk;
    if (stats) { stats[0] = total_seq; stats[1] = total_lit; stats[2] = rep_seq; stats[3] = of_bits; stats[4] = ml_sum; stats[5] = win_w ? win_steps : (uint32_t)M.nsearch; stats[6] = getenv("ENC_COUNTCONT") ? (uint32_t)(M.ncheap >> 4) : (uint32_t)M.nskip; stats[7] = (uint32_t)(M.nverify >> 4); }
    return o - dst;
}

// Entropy-coder isolation: encode an externally supplied parse (e.g. libzstd's own level-12 sequences from
// ZSTD_generateSequences) with write_block_body.  seq array: {offset, litLength, matchLength, rep} u32 x4, with
// block delimiters (offset == 0 && matchLength == 0) carrying the trailing literals of each block.
extern "C" long enc_model_from_sequences(const uint8_t *src, uint32_t n, const uint32_t *zs, uint32_t nzs, uint8_t *dst, uint32_t cap) {
    uint8_t *o = dst;
    *o++ = 0x28; *o++ = 0xB5; *o++ = 0x2F; *o++ = 0xFD;
    if (n <= 255) { *o++ = 0x20; *o++ = (uint8_t)n; }
    else if (n <= 65791) { *o++ = 0x60; uint32_t v = n - 256; *o++ = (uint8_t)v; *o++ = (uint8_t)(v >> 8); }
    else { *o++ = 0xA0; for (int i = 0; i < 4; i++) *o++ = (uint8_t)(n >> (8 * i)); }
    uint32_t rep[3] = {1, 4, 8};
    std::vector<Seq> seqs; std::vector<uint8_t> lits, body;
    EncWork *wk = new EncWork;
    uint32_t pos = 0, bs = 0;
    for (uint32_t i = 0; i < nzs; i++) {
        uint32_t off = zs[4 * i], ll = zs[4 * i + 1], ml = zs[4 * i + 2];
        lits.insert(lits.end(), src + pos, src + pos + ll);
        pos += ll;
        if (off == 0 && ml == 0) {  // block delimiter
            body.resize(block_body_bound((uint32_t)lits.size(), (uint32_t)seqs.size()) + 64);
            uint32_t bl = write_block_body(body.data(), lits.data(), (uint32_t)lits.size(), seqs.data(), (uint32_t)seqs.size(), wk);
            uint32_t blen = pos - bs;
            bool last = pos == n;
            if (bl >= blen) { delete wk; return -2; }
            uint32_t h = (last ? 1u : 0u) | 2u << 1 | bl << 3;
            *o++ = (uint8_t)h; *o++ = (uint8_t)(h >> 8); *o++ = (uint8_t)(h >> 16);
            memcpy(o, body.data(), bl); o += bl;
            seqs.clear(); lits.clear(); bs = pos;
            continue;
        }
        bool ll0 = ll == 0;
        uint32_t rc = rep_code_for(off, rep, ll0);
        uint32_t ob = rc ? rc : off + 3;
        seqs.push_back({ll, ml, ob});
        rep_update(rep, ob, ll0);
        pos += ml;
    }
    delete wk;
    return pos == n ? o - dst : -3;
}
