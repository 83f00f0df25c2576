// CPU model of the round-2 K3 parse (dev/test tool, NOT product path): bucketed row search -> rep-blind lazy
// decide -> chase with repeat-offset scanning, written stage by stage the way the CUDA kernels run it, so the
// ratio of a parameter set can be measured against libzstd level 12 without a GPU.
//   g++ -O2 -shared -fPIC -I squishrs_b200/csrc tests/harness/lz_model2.cc -o tests/harness/liblz_model2.so
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "zstd_enc_block.h"
#include "zstd_enc_parse.h"

using namespace zc;

struct P2 {
    int rows_log, K, mm, cap, stride, tile, slot_by_pos;
    int rep_scan, rep_min, rep_at_start, accept_thr, target_len, max_shift, back_max;  // (round-2 exploration knobs; the parse is zstd_enc_parse.h now)
    int hash2_log, hash2_K, hash2_bytes;  // optional second table keyed on a longer hash (0 = off)
    int sel_mul;
    int lazy_rep;     // decide pass: neighbour-offset rep approximation (0 off)
    int chain;        // 1: candidates come from a full per-position hash chain (upper bound study), depth = K
    int short_keep;   // > 0: candidates whose first 8 bytes do not all match ("short") are verified only for the short_keep nearest; 0 = all
    int tag_bits;     // > 0 (with short_keep): long/short classes come from ptag (5 bits of the row hash) and xtag (tag_bits bits of a hash of bytes 5..7) instead of the bytes
    int long_cap;     // > 0: at most this many long candidates per position enter the pair queue (the nearest ones)
    int skip_capped;  // > 0: positions p with p % skip_capped != 0 are not searched when their anchor (p rounded down) found a match still >= CAP long at p
    int ins_stride;   // > 1: only positions p % ins_stride == 0 enter the table (every position is still searched; continuation = same pair ins_stride positions earlier)
    int skip_runs;    // 1: a position whose 5-byte prefix equals the previous position's (inside a run of one byte) is not inserted
    int cont;         // continuation filter: 0 off, 1 exact (pair (p-1,c-1) was a candidate pair), 2 previous byte equal; refresh every 16 positions
    int group_skip;   // 1: a group of 32 positions that lies inside a match already known to be long (the record in front of the group is capped and the
                      // match goes on for 64 more bytes) is not searched: every position inherits (offset, CAP, may-be-longer)
    int sub_len;      // > 0: every block is parsed in independent pieces of this many bytes (repeat offsets unknown at each start), sequences merged afterwards
    int xtag_more;    // > 0: the 6-bit tag covers this many more bytes (from byte 8 on), so "long" means 8 + xtag_more common bytes
    int part_skip;    // > 0: positions at the start of a group covered by the match that crosses into it are not searched, in units of part_skip positions
    int cont_period;  // > 0: the filter keeps everything at positions p % cont_period == 0 (the first column of a search group) and is independent of the tile
};

static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t hashN(uint64_t v, int bytes, int hl) {
    if (bytes == 4) return (uint32_t)(((v << 32) * 889523592379ULL) >> (64 - hl));
    if (bytes == 5) return (uint32_t)(((v << 24) * 889523592379ULL) >> (64 - hl));
    if (bytes == 6) return (uint32_t)(((v << 16) * 227718039650203ULL) >> (64 - hl));
    if (bytes == 7) return (uint32_t)(((v << 8) * 58295818150454627ULL) >> (64 - hl));
    return (uint32_t)((v * 0xCF1BBCDCB7A56463ULL) >> (64 - hl));
}
static inline uint32_t match_len(const uint8_t *s, uint32_t a, uint32_t b, uint32_t n, uint32_t lim = ~0u) {  // a > b; stops counting at lim
    uint32_t l = 0;
    while (a + l + 8 <= n && l < lim) {
        uint64_t x = rd64(s + a + l) ^ rd64(s + b + l);
        if (x) return l + (__builtin_ctzll(x) >> 3);
        l += 8;
    }
    while (a + l < n && l < lim && s[a + l] == s[b + l]) l++;
    return l;
}

extern "C" long lz_model2_frame(const uint8_t *s, uint32_t n, uint8_t *dst, uint32_t cap_dst, const P2 *Pp, uint64_t *stats) {
    const P2 P = *Pp;
    const uint32_t MM = (uint32_t)P.mm, CAP = (uint32_t)P.cap;
    std::vector<uint32_t> blen(n + 64, 0), boff(n + 64, 0);
    std::vector<uint8_t> noback(n + 64, 0);
    uint64_t nverify = 0, nrows = 0, nlong = 0, nskipped = 0;
    // ---- stage S: search ----
    {
        const uint32_t rows = 1u << P.rows_log, K = (uint32_t)P.K;
        std::vector<uint32_t> tab((size_t)rows * K, 0), head(rows, 0);
        const uint32_t rows2 = P.hash2_log ? 1u << P.hash2_log : 0, K2 = (uint32_t)P.hash2_K;
        std::vector<uint32_t> tab2((size_t)rows2 * std::max(K2, 1u), 0), head2(std::max(rows2, 1u), 0);
        std::vector<uint32_t> chain_prev, chain_head;
        if (P.chain) { chain_prev.assign(n, 0); chain_head.assign(1u << 20, 0); }
        const uint32_t T = (uint32_t)P.tile;
        for (uint32_t t0 = 0; t0 < n; t0 += T) {
            const uint32_t t1 = std::min(n, t0 + T);
            for (uint32_t p = t0; p < t1 && p + 8 <= n; p++) {
                if (P.ins_stride > 1 && p % (uint32_t)P.ins_stride) continue;
                if (P.skip_runs && p > 0 && ((rd64(s + p) ^ rd64(s + p - 1)) & 0xFFFFFFFFFFull) == 0) continue;
                const uint32_t h = hashN(rd64(s + p), P.mm, P.rows_log);
                tab[(size_t)h * K + (P.slot_by_pos ? p % K : head[h]++ % K)] = p + 1;
                if (rows2) { const uint32_t h2 = hashN(rd64(s + p), P.hash2_bytes, P.hash2_log); tab2[(size_t)h2 * K2 + (P.slot_by_pos ? p % K2 : head2[h2]++ % K2)] = p + 1; }
                if (P.chain) { const uint32_t hc = hashN(rd64(s + p), P.mm, 20); chain_prev[p] = chain_head[hc]; chain_head[hc] = p + 1; }
            }
            std::vector<uint32_t> prevc, curc, hist[4];
            uint32_t part_until = 0;
            for (int phase = 0; phase < (P.skip_capped ? 2 : 1); phase++)
            for (uint32_t p = t0; p < t1 && p + 8 <= n; p++) {
                if (p % (uint32_t)P.stride) continue;
                if (P.part_skip && (p & 31u) == 0 && p >= 32 && (p & 255u) != 0 && p + 32 + CAP + 16 <= n) {
                    // positions at the start of a group that lie inside the match reaching furthest across the group boundary inherit it
                    uint32_t best_end = 0, bo2 = 0;
                    for (uint32_t d = 0; d < 32; d++) { const uint32_t q = p - 1 - d; if (blen[q] && q + std::min(blen[q], CAP) > best_end) { best_end = q + std::min(blen[q], CAP); bo2 = boff[q]; } }
                    part_until = 0;
                    if (best_end > p + MM) {
                        uint32_t cov = best_end - p - MM + 1;            // positions p .. p+cov-1 still have >= MM bytes of that match
                        cov = (cov / (uint32_t)P.part_skip) * (uint32_t)P.part_skip;  // whole passes only
                        if (cov > 32) cov = 32;
                        for (uint32_t k = 0; k < cov; k++) { blen[p + k] = best_end - (p + k); boff[p + k] = bo2; noback[p + k] = 1; }
                        part_until = p + cov; nskipped += cov;
                    }
                }
                if (P.part_skip && p < part_until) continue;
                if (P.group_skip && (p & 31u) == 0 && p >= 32 && (p & 255u) != 0 && p + 32 + CAP + 16 <= n) {
                    uint32_t q = p - 1, go = 0;
                    for (uint32_t d = 0; d < 32 && !go; d++) if (blen[p - 1 - d] >= CAP && blen[p - 1 - d] > d) { go = boff[p - 1 - d]; q = p - 1 - d; }
                    if (go && match_len(s, p, p - go, n, 64) >= 64) { for (uint32_t k = 0; k < 32; k++) { blen[p + k] = CAP; boff[p + k] = go; noback[p + k] = 1; } nskipped += 32; p += 31; continue; }
                }
                if (P.skip_capped) {
                    const uint32_t a = p - p % (uint32_t)P.skip_capped;
                    if ((phase == 0) != (a == p)) continue;
                    if (a != p && blen[a] >= CAP && match_len(s, p, p - boff[a], n, CAP) >= CAP) { blen[p] = CAP; boff[p] = boff[a]; nskipped++; continue; }
                }
                uint32_t bl = 0, bo = 0; int32_t bs = -1000;
                if (P.ins_stride > 1) { hist[p & 3] = curc; prevc = hist[(p + 4 - P.ins_stride) & 3]; if (p < t0 + (uint32_t)P.ins_stride) prevc.clear(); curc.clear(); }
                else { prevc.swap(curc); curc.clear(); if (p == t0 && !P.cont_period) prevc.clear(); }
                auto consider = [&](uint32_t c) {
                    if (c >= p) return;
                    if (rd32(s + c) != rd32(s + p)) return;
                    if (P.cont) {
                        if (std::find(curc.begin(), curc.end(), c) != curc.end()) return;  // same candidate from the second table
                        curc.push_back(c);
                        if ((P.cont_period ? p % (uint32_t)P.cont_period != 0 : ((p & 15u) != 0 && p > t0)) && c > 0) {
                            const uint32_t st = P.ins_stride > 1 ? (uint32_t)P.ins_stride : 1u;
                            if (P.cont == 1 && c >= st && std::find(prevc.begin(), prevc.end(), c - st) != prevc.end()) return;
                            if (P.cont == 2 && s[p - 1] == s[c - 1]) return;
                        }
                    }
                    uint32_t l = match_len(s, p, c, n, CAP);
                    nverify++;
                    if (l < MM) return;
                    if (l > CAP) l = CAP;
                    const uint32_t off = p - c;
                    const int32_t sc = P.sel_mul * (int32_t)l - (int32_t)highbit(off + 3);
                    if (sc > bs || (sc == bs && off < bo)) { bs = sc; bl = l; bo = off; }
                };
                nrows++;
                if (P.chain) {
                    uint32_t e = chain_prev[p]; int depth = P.K;
                    while (e && depth--) { consider(e - 1); e = chain_prev[e - 1]; }
                } else {
                    const uint32_t h = hashN(rd64(s + p), P.mm, P.rows_log);
                    if (P.short_keep) {
                        uint32_t shorts[64]; int ns = 0, nl = 0;
                        uint32_t longs[64]; int nlq = 0;
                        auto consider_long = [&](uint32_t c) { nlong++; if (P.long_cap) { longs[nlq++] = c; return; } nl++; consider(c); };
                        for (uint32_t k = 0; k < K; k++) {
                            const uint32_t e = tab[(size_t)h * K + k];
                            if (!e || e - 1 >= p) continue;
                            if (P.tag_bits) {
                                const uint64_t vc = rd64(s + e - 1), vp = rd64(s + p);
                                const uint32_t pc = hashN(vc, P.mm, 19) & 31u, pp = hashN(vp, P.mm, 19) & 31u;
                                uint32_t xc = (uint32_t)(((vc >> 40) * 0x9E3779B1u) & 0xFFFFFFFFu) >> (32 - P.tag_bits), xp = (uint32_t)(((vp >> 40) * 0x9E3779B1u) & 0xFFFFFFFFu) >> (32 - P.tag_bits);
                                if (P.xtag_more > 0 && p + 16 <= n) {  // the tag also covers xtag_more bytes from byte 8 on
                                    const uint64_t mk = P.xtag_more >= 8 ? ~0ull : (1ull << (8 * P.xtag_more)) - 1;
                                    const uint64_t wc = (vc >> 40) | (rd64(s + e - 1 + 8) & mk) << 24, wp = (vp >> 40) | (rd64(s + p + 8) & mk) << 24;
                                    xc = (uint32_t)((wc * 0x9E3779B97F4A7C15ull) >> (64 - P.tag_bits)); xp = (uint32_t)((wp * 0x9E3779B97F4A7C15ull) >> (64 - P.tag_bits));
                                }
                                if (pc != pp) continue;
                                if (xc == xp) consider_long(e - 1); else shorts[ns++] = e - 1;
                                continue;
                            }
                            if (rd64(s + e - 1) == rd64(s + p)) consider(e - 1);
                            else if (rd32(s + e - 1) == rd32(s + p) && s[e - 1 + 4] == s[p + 4]) shorts[ns++] = e - 1;
                        }
                        if (P.long_cap) { std::sort(longs, longs + nlq); for (int i = 0; i < P.long_cap && i < nlq; i++) consider(longs[nlq - 1 - i]); }  // the nearest long_cap long candidates
                        std::sort(shorts, shorts + ns);
                        for (int i = 0; i < P.short_keep && i < ns && P.short_keep < 100; i++) consider(shorts[ns - 1 - i]);  // short_keep >= 100: short candidates are dropped
                    } else
                    for (uint32_t k = 0; k < K; k++) { const uint32_t e = tab[(size_t)h * K + k]; if (e) consider(e - 1); }
                    if (rows2) {
                        const uint32_t h2 = hashN(rd64(s + p), P.hash2_bytes, P.hash2_log);
                        for (uint32_t k = 0; k < K2; k++) { const uint32_t e = tab2[(size_t)h2 * K2 + k]; if (e) consider(e - 1); }
                    }
                }
                blen[p] = bl; boff[p] = bo;
            }
        }
        // inheritance: a match (off, len) at p is a match (off, len-1) at p+1 -- serves un-searched positions and capped matches
        for (uint32_t p = 1; p < n; p++) {
            if (blen[p - 1] > MM || (blen[p - 1] >= CAP)) {
                const uint32_t il = blen[p - 1] >= CAP ? CAP : blen[p - 1] - 1, io = boff[p - 1];
                // a capped match stays capped while it really continues; the model checks the true length
                uint32_t tl = il;
                if (blen[p - 1] >= CAP) { tl = match_len(s, p, p - io, n, CAP); if (tl > CAP) tl = CAP; }
                if (tl >= MM) {
                    const int32_t si = P.sel_mul * (int32_t)tl - (int32_t)highbit(io + 3);
                    const int32_t so = blen[p] ? P.sel_mul * (int32_t)blen[p] - (int32_t)highbit(boff[p] + 3) : -1000;
                    if (si > so) { blen[p] = tl; boff[p] = io; }
                }
            }
        }
    }
    // ---- records exactly as the search kernel writes them, then the shipped parse (zstd_enc_parse.h) block by block ----
    std::vector<uint32_t> rec(n + 8, 0);
    for (uint32_t p = 0; p < n; p++) {
        uint32_t l = blen[p], capped = l >= CAP ? 1u : 0u;
        if (!l) continue;
        if (l > CAP) l = CAP;
        const uint32_t be = std::min(n, (p / Z_BLOCK_MAX + 1) * Z_BLOCK_MAX);
        if (p + l > be) { l = be - p; capped = 0; }
        if (l < MM) continue;
        uint32_t k = 0;
        const uint32_t c = p - boff[p];
        while (!noback[p] && k < 3 && p > k && c > k && s[p - k - 1] == s[c - k - 1]) k++;
        rec[p] = boff[p] | (l - zparse::LEN_BASE) << 21 | capped << 26 | k << 27;
    }
    uint8_t *o = dst;
    *o++ = 0x28; *o++ = 0xB5; *o++ = 0x2F; *o++ = 0xFD;
    if (n <= 255) { *o++ = 0x20; *o++ = (uint8_t)n; }
    else if (n <= 65791) { *o++ = 0x60; uint32_t v = n - 256; *o++ = (uint8_t)v; *o++ = (uint8_t)(v >> 8); }
    else { *o++ = 0xA0; for (int i = 0; i < 4; i++) *o++ = (uint8_t)(n >> (8 * i)); }
    std::vector<Seq> seqs(Z_BLOCK_MAX / 3 + 8); std::vector<uint8_t> lits, body;
    EncWork *wk = new EncWork;
    uint64_t total_seq = 0, rep_seq = 0, total_lit = 0;
    const uint32_t nblocks = n ? (n + Z_BLOCK_MAX - 1) / Z_BLOCK_MAX : 1;
    for (uint32_t b = 0; b < nblocks; b++) {
        const uint32_t bs = b * Z_BLOCK_MAX, be = std::min(n, bs + Z_BLOCK_MAX);
        uint32_t last_lits = be - bs, ns = 0;
        if (be > bs && P.sub_len > 0) {
            uint32_t carry = 0;
            std::vector<Seq> tmp(P.sub_len / 6 + 8);
            for (uint32_t sb = bs; sb < be; sb += (uint32_t)P.sub_len) {
                const uint32_t se = std::min(be, sb + (uint32_t)P.sub_len);
                uint32_t ll_last = 0;
                const uint32_t k = zparse::chase_block(s, n, rec.data(), sb, se, sb == 0, tmp.data(), (uint32_t)P.sub_len / 6, &ll_last);
                for (uint32_t i = 0; i < k; i++) { Seq q = tmp[i]; if (i == 0) q.ll += carry; seqs[ns++] = q; }
                carry = k ? ll_last : carry + ll_last;
            }
            last_lits = carry;
        } else
        if (be > bs) ns = zparse::chase_block(s, n, rec.data(), bs, be, b == 0, seqs.data(), Z_BLOCK_MAX / 6 + 8, &last_lits);
        lits.clear();
        uint32_t pos = bs;
        for (uint32_t i = 0; i < ns; i++) { lits.insert(lits.end(), s + pos, s + pos + seqs[i].ll); pos += seqs[i].ll + seqs[i].ml; if (seqs[i].off_base <= 3) rep_seq++; }
        lits.insert(lits.end(), s + pos, s + be);
        body.resize(block_body_bound((uint32_t)lits.size(), ns) + 64);
        const uint32_t bl = write_block_body(body.data(), lits.data(), (uint32_t)lits.size(), seqs.data(), ns, wk);
        const uint32_t blen_b = be - bs;
        const bool last = b + 1 == nblocks;
        if ((size_t)(o - dst) + 3 + std::max(bl, blen_b) > cap_dst) { delete wk; return -1; }
        if (bl >= blen_b || blen_b == 0) {
            const uint32_t h = (last ? 1u : 0u) | 0u << 1 | blen_b << 3;
            *o++ = (uint8_t)h; *o++ = (uint8_t)(h >> 8); *o++ = (uint8_t)(h >> 16);
            memcpy(o, s + bs, blen_b); o += blen_b;
        } else {
            const uint32_t h = (last ? 1u : 0u) | 2u << 1 | bl << 3;
            *o++ = (uint8_t)h; *o++ = (uint8_t)(h >> 8); *o++ = (uint8_t)(h >> 16);
            memcpy(o, body.data(), bl); o += bl;
            total_seq += ns; total_lit += lits.size();
        }
    }
    delete wk;
    if (stats) { stats[0] = total_seq; stats[1] = total_lit; stats[2] = rep_seq; stats[3] = nverify; stats[4] = nrows; stats[5] = nlong; stats[6] = nskipped; }
    return o - dst;
}

// The parameters the GPU ships (zstd_enc_lz2.cuh): 2^14 rows, the 32 most recent earlier positions of the row as candidates
// (tile = 512: the window of a position ends at its row's fill level after the 512-position tile it lies in), 5 + 6 tag bits, nearest
// short candidate only, exact continuation filter refreshed at the first column of every group of 32, 5-byte matches (4 in
// chunks <= 128 KiB), no inserts inside runs of one byte.
extern "C" long lz_model2_shipped(const uint8_t *s, uint32_t n, uint8_t *dst, uint32_t cap_dst) {
    P2 P;
    memset(&P, 0, sizeof P);
    P.rows_log = 14; P.K = 32; P.mm = n <= 128u * 1024u ? 4 : 5; P.cap = 32; P.stride = 1; P.tile = 512; P.sel_mul = 2; P.cont_period = 32; P.group_skip = 1;
    P.short_keep = 1; P.tag_bits = 6; P.cont = 1; P.skip_runs = 1;
    return lz_model2_frame(s, n, dst, cap_dst, &P, nullptr);
}
