// K3 entropy stage, warp-parallel: one warp writes one compressed block body (literals section + sequences
// section, RFC 8878 3.1.1.3) from the block's gathered literals and parsed sequences.
//
// What is parallel and how:
//   histograms        lanes stride over literals / sequences, shared-memory atomics
//   table building    one lane, on shared-memory state (zstd_core.h: Huffman lengths/codes/tree, FSE normalise/
//                     NCount/CTable) -- tiny next to the bit packing
//   Huffman streams   "segment-parallel bit packing": each of the 4 streams is cut into 8 symbol segments (one per
//                     lane).  Pass 1 sums code lengths per segment, a warp prefix sum turns them into bit offsets
//                     (and the stream byte sizes / jump table); pass 2 lets every lane emit its segment at its
//                     offset.  Words strictly inside a lane's bit range are plain stores; the first and last word
//                     of a range are pre-zeroed and merged with atomicOr, so neighbours sharing a word never race.
//   FSE sequences     the three state chains (LL, OF, ML) are serial in sequence order: lanes 0-2 walk them once
//                     and record each transition's (bits, count); then the same segment-parallel packing emits the
//                     interleaved state bits + extra bits of 32 sequence segments.
#pragma once
#include "zstd_enc_block.h"

namespace ent {

struct WarpWork {
    zc::EncWork wk;
    uint32_t scalars[16];
};

// ---- lane-private writer of one contiguous bit range of a stream --------------------------------------------
struct SegWriter {
    uint32_t *words;       // aligned word array containing the stream
    uint64_t acc;
    uint32_t nacc;         // valid bits in acc
    uint32_t widx;         // next word to flush
    uint32_t first_w, last_w;
};
__device__ __forceinline__ void seg_init(SegWriter *w, uint8_t *stream, uint64_t bit_start, uint64_t bit_end) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(stream);
    w->words = reinterpret_cast<uint32_t *>(a & ~(uintptr_t)3);
    const uint64_t g0 = (uint64_t)(a & 3) * 8 + bit_start, g1 = (uint64_t)(a & 3) * 8 + bit_end;
    w->first_w = (uint32_t)(g0 >> 5);
    w->last_w = (uint32_t)((g1 - 1) >> 5);
    w->widx = w->first_w;
    w->nacc = (uint32_t)(g0 & 31);
    w->acc = 0;
}
// zero the bytes of the range's first and last word that belong to the stream's byte extent [stream, stream+size)
__device__ __forceinline__ void seg_zero_edges(const SegWriter *w, uint8_t *stream, uint32_t size) {
    uint8_t *wb = reinterpret_cast<uint8_t *>(w->words);
    uint8_t *lo = stream, *hi = stream + size;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint8_t *p = wb + (size_t)w->first_w * 4 + k;
        if (p >= lo && p < hi) *p = 0;
        uint8_t *q = wb + (size_t)w->last_w * 4 + k;
        if (q >= lo && q < hi) *q = 0;
    }
}
__device__ __forceinline__ void seg_flush_word(SegWriter *w) {
    const uint32_t v = (uint32_t)w->acc;
    if (w->widx == w->first_w || w->widx == w->last_w) { if (v) atomicOr(&w->words[w->widx], v); }
    else w->words[w->widx] = v;
    w->widx++;
    w->acc >>= 32;
    w->nacc -= 32;
}
__device__ __forceinline__ void seg_put(SegWriter *w, uint32_t v, uint32_t bits) {  // bits <= 31
    w->acc |= (uint64_t)v << w->nacc;
    w->nacc += bits;
    if (w->nacc >= 32) seg_flush_word(w);
}
__device__ __forceinline__ void seg_finish(SegWriter *w) {
    if (w->nacc) { const uint32_t v = (uint32_t)w->acc; if (v) atomicOr(&w->words[w->widx], v); }
}

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane, uint32_t *total) {
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if ((int)lane >= d) x += y; }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

// ---- literals section ---------------------------------------------------------------------------------------------
// returns bytes written at dst (all lanes return the same value)
__device__ uint32_t warp_write_literals(uint8_t *dst, const uint8_t *lits, uint32_t n, WarpWork *W, uint32_t lane) {
    zc::EncWork *wk = &W->wk;
    // histogram
    for (uint32_t s = lane; s < 256; s += 32) wk->counts[s] = 0;
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) atomicAdd(&wk->counts[lits[i]], 1u);
    __syncwarp();
    // mode decision + tables by one lane; the symbol sort that the Huffman construction starts from is done by the whole warp
    // (rank = number of used symbols that come before in (count, symbol) order: 8 symbols per lane against all 256)
    enum { M_RAW = 0, M_RLE = 1, M_HUF1 = 2, M_HUF4 = 3 };
    const uint32_t lh = 3 + (n >= 1024) + (n >= 16384);
    uint32_t used = 0;
    {
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) mine += wk->counts[lane + 32 * j] != 0;
        used = mine;
#pragma unroll
        for (int d = 16; d; d >>= 1) used += __shfl_xor_sync(0xffffffffu, used, d);
    }
    if (used >= 2 && n >= 64) {
        uint32_t cs[8], rk[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { cs[j] = wk->counts[lane + 32 * j]; rk[j] = 0; wk->lens[lane + 32 * j] = 0; }
        for (uint32_t t = 0; t < 256; t++) {
            const uint32_t ct = wk->counts[t];
            if (ct == 0) continue;  // warp-uniform
#pragma unroll
            for (int j = 0; j < 8; j++) rk[j] += (ct < cs[j] || (ct == cs[j] && t < lane + 32 * j)) ? 1u : 0u;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) if (cs[j]) wk->hwork[rk[j]] = (uint16_t)(lane + 32 * j);
    }
    __syncwarp();
    if (lane == 0) {
        uint32_t mode = M_RAW, tree = 0, maxlen = 0;
        if (n > 0) {
            if (used == 1) mode = M_RLE;
            else if (n >= 64) {
                maxlen = zc::huf_build_lengths_sorted(wk->lens, wk->counts, Z_HUF_MAXBITS, wk->hwork, used);
                if (maxlen) {
                    uint64_t bits = 0;
                    for (uint32_t s = 0; s < 256; s++) bits += (uint64_t)wk->counts[s] * wk->lens[s];
                    const uint32_t est = (uint32_t)((bits + 7) / 8);
                    if (est + 16 + (used + 1) / 2 < n - (n >> 6)) {
                        zc::huf_assign_codes(wk->codes, wk->lens, maxlen);
                        tree = zc::huf_write_tree(dst + lh, wk->lens, maxlen, &wk->ct[0], wk->spread);
                        if (tree) mode = n < 256 ? M_HUF1 : M_HUF4;
                    }
                }
            }
        }
        W->scalars[0] = mode; W->scalars[1] = tree;
    }
    __syncwarp();
    const uint32_t mode = W->scalars[0], tree = W->scalars[1];
    uint32_t written = 0;
    if (mode == M_HUF4) {
        // pass 1: code-length sums of the 32 symbol segments (lane = stream*8 + segment, segments in BIT order:
        // segment 0 holds the stream's LAST symbols, which are written first)
        const uint32_t seg4 = (n + 3) / 4;
        const uint32_t k = lane >> 3, j = lane & 7;
        const uint32_t s_begin = k * seg4, s_cnt = k < 3 ? seg4 : n - 3 * seg4;
        const uint32_t q = (s_cnt + 7) / 8;
        const uint32_t hi = s_cnt > j * q ? s_cnt - j * q : 0, lo = s_cnt > (j + 1) * q ? s_cnt - (j + 1) * q : 0;  // symbols [lo, hi) of stream k
        // the literal buffer is read a 32-bit word at a time (each lane walks its own segment; byte loads would cost a
        // memory round trip per symbol)
        const uint32_t *lits32 = reinterpret_cast<const uint32_t *>(lits);  // the per-warp literal slot is 64-byte aligned
        uint32_t bits = 0;
        {
            uint32_t cw = 0, cwi = ~0u;
            for (uint32_t i = lo; i < hi; i++) {
                const uint32_t a = s_begin + i;
                if ((a >> 2) != cwi) { cwi = a >> 2; cw = lits32[cwi]; }
                bits += wk->lens[(cw >> (8 * (a & 3u))) & 0xFFu];
            }
        }
        if (j == 7) bits += 1;  // the segment holding symbol 0 ends the stream: sentinel bit
        // prefix within each group of 8 lanes
        uint32_t x = bits;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d, 8); if ((int)j >= d) x += y; }
        const uint32_t start_bit = x - bits;
        const uint32_t stream_bits = __shfl_sync(0xffffffffu, x, 7, 8);
        const uint32_t stream_bytes = (stream_bits + 7) / 8;
        const uint32_t sz0 = __shfl_sync(0xffffffffu, stream_bytes, 0), sz1 = __shfl_sync(0xffffffffu, stream_bytes, 8),
                       sz2 = __shfl_sync(0xffffffffu, stream_bytes, 16), sz3 = __shfl_sync(0xffffffffu, stream_bytes, 24);
        const uint32_t csize = tree + 6 + sz0 + sz1 + sz2 + sz3;
        if (csize < n && sz0 <= 0xFFFF && sz1 <= 0xFFFF && sz2 <= 0xFFFF) {
            uint8_t *p = dst + lh + tree;
            if (lane == 0) {
                p[0] = (uint8_t)sz0; p[1] = (uint8_t)(sz0 >> 8); p[2] = (uint8_t)sz1; p[3] = (uint8_t)(sz1 >> 8); p[4] = (uint8_t)sz2; p[5] = (uint8_t)(sz2 >> 8);
                if (lh == 3) { const uint32_t v = 2u | 1u << 2 | n << 4 | csize << 14; dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); }
                else if (lh == 4) { const uint32_t v = 2u | 2u << 2 | n << 4 | csize << 18; dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); dst[3] = (uint8_t)(v >> 24); }
                else { const uint64_t v = 2u | 3u << 2 | (uint64_t)n << 4 | (uint64_t)csize << 22; for (int i = 0; i < 5; i++) dst[i] = (uint8_t)(v >> (8 * i)); }
            }
            const uint32_t off_k = 6 + (k > 0 ? sz0 : 0) + (k > 1 ? sz1 : 0) + (k > 2 ? sz2 : 0);
            uint8_t *stream = p + off_k;
            SegWriter w;
            const bool active = bits > 0;
            if (active) { seg_init(&w, stream, start_bit, start_bit + bits); seg_zero_edges(&w, stream, stream_bytes); }
            __syncwarp();
            __threadfence_block();
            if (active) {
                uint32_t cw = 0, cwi = ~0u;
                for (uint32_t i = hi; i > lo; i--) {
                    const uint32_t a = s_begin + i - 1;
                    if ((a >> 2) != cwi) { cwi = a >> 2; cw = lits32[cwi]; }
                    const uint32_t s = (cw >> (8 * (a & 3u))) & 0xFFu;
                    seg_put(&w, wk->codes[s], wk->lens[s]);
                }
                if (j == 7) seg_put(&w, 1, 1);
                seg_finish(&w);
            }
            __syncwarp();
            written = lh + csize;
        }
    } else if (mode == M_HUF1) {
        if (lane == 0) {
            uint8_t *p = dst + lh + tree;
            const uint32_t sz = zc::huf_encode_stream(p, lits, n, wk->codes, wk->lens);
            const uint32_t csize = tree + sz;
            uint32_t wr = 0;
            if (csize < n) { const uint32_t v = 2u | 0u << 2 | n << 4 | csize << 14; dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); wr = 3 + csize; }
            W->scalars[2] = wr;
        }
        __syncwarp();
        written = W->scalars[2];
    }
    if (written) return written;
    // Raw / RLE literals
    const bool rle = mode == M_RLE;
    uint32_t h;
    if (n <= 31) h = 1; else if (n <= 4095) h = 2; else h = 3;
    if (lane == 0) {
        const uint32_t type = rle ? 1 : 0;
        if (h == 1) dst[0] = (uint8_t)(type | n << 3);
        else if (h == 2) { const uint32_t v = type | 1u << 2 | n << 4; dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); }
        else { const uint32_t v = type | 3u << 2 | n << 4; dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); }
        if (rle) dst[h] = lits[0];
    }
    if (rle) { __syncwarp(); return h + 1; }
    for (uint32_t i = lane; i < n; i += 32) dst[h + i] = lits[i];
    __syncwarp();
    return h + n;
}

// ---- sequences section ------------------------------------------------------------------------------------------------
// sbits: per-warp scratch of 4 * nseq u32: state-transition bits | count << 16 for LL, OF, ML, then the packed symbol codes
// (LL code | OF code << 8 | ML code << 16), computed once by all lanes so that the serial passes never recompute them
// cbytes: per-warp scratch of 3 byte arrays (stride cstride, 16-byte aligned): the LL / OF / ML code of every sequence, the form the
// serial FSE chains read (16 symbols per load)
__device__ uint32_t warp_write_sequences(uint8_t *dst, const zc::Seq *seqs, uint32_t nseq, WarpWork *W, uint32_t *sbits, uint8_t *cbytes,
                                         uint32_t cstride, uint32_t lane) {
    zc::EncWork *wk = &W->wk;
    if (nseq == 0) { if (lane == 0) dst[0] = 0; __syncwarp(); return 1; }
    // code histograms
    for (uint32_t s = lane; s < 192; s += 32) wk->hist[s / 64][s % 64] = 0;
    __syncwarp();
    uint32_t *codes = sbits + 3 * (size_t)nseq;
    for (uint32_t i = lane; i < nseq; i += 32) {
        const zc::Seq q = seqs[i];
        const uint32_t lc = zc::ll_code(q.ll), oc = zc::highbit(q.off_base), mc = zc::ml_code(q.ml);
        atomicAdd(&wk->hist[0][lc], 1u);
        atomicAdd(&wk->hist[1][oc], 1u);
        atomicAdd(&wk->hist[2][mc], 1u);
        codes[i] = lc | oc << 8 | mc << 16;
        cbytes[i] = (uint8_t)lc; cbytes[cstride + i] = (uint8_t)oc; cbytes[2 * cstride + i] = (uint8_t)mc;
    }
    __syncwarp();
    __threadfence_block();
    if (lane == 0) {
        uint8_t *p = dst;
        if (nseq < 128) *p++ = (uint8_t)nseq;
        else if (nseq < 0x7F00) { *p++ = (uint8_t)((nseq >> 8) + 128); *p++ = (uint8_t)nseq; }
        else { *p++ = 255; *p++ = (uint8_t)(nseq - 0x7F00); *p++ = (uint8_t)((nseq - 0x7F00) >> 8); }
        uint32_t mx[3] = {0, 0, 0};
        for (int t = 0; t < 3; t++) for (uint32_t s = 0; s < 64; s++) if (wk->hist[t][s]) mx[t] = s;
        uint8_t *modes = p++;
        const uint32_t m_ll = zc::choose_table(&p, &wk->ct[0], wk->norm[0], wk->hist[0], nseq, mx[0], zc::ZTAB(LL_defnorm), 35, 6, Z_LL_MAXLOG, wk->spread);
        const uint32_t m_of = zc::choose_table(&p, &wk->ct[1], wk->norm[1], wk->hist[1], nseq, mx[1], zc::ZTAB(OF_defnorm), 28, 5, Z_OF_MAXLOG, wk->spread);
        const uint32_t m_ml = zc::choose_table(&p, &wk->ct[2], wk->norm[2], wk->hist[2], nseq, mx[2], zc::ZTAB(ML_defnorm), 52, 6, Z_ML_MAXLOG, wk->spread);
        *modes = (uint8_t)(m_ll << 6 | m_of << 4 | m_ml << 2);
        W->scalars[0] = (uint32_t)(p - dst);
        W->scalars[1] = m_ll; W->scalars[2] = m_of; W->scalars[3] = m_ml;
    }
    __syncwarp();
    const uint32_t hdr = W->scalars[0];
    const uint32_t mode_of_lane[3] = {W->scalars[1], W->scalars[2], W->scalars[3]};
    // pass A: the three FSE state chains, last sequence first (lane 0 = LL, 1 = OF, 2 = ML)
    if (lane < 3) {
        const zc::FseCTable *ct = &wk->ct[lane];
        const bool rle = mode_of_lane[lane] == 1;
        uint32_t *out = sbits + (size_t)lane * nseq;
        const uint8_t *cb = cbytes + (size_t)lane * cstride;  // this lane's stream of symbol codes
        uint32_t state = 0;
        if (rle) {
            for (uint32_t i = 0; i + 1 < nseq; i++) out[i] = 0;
        } else {
            state = zc::fse_init_state(ct, cb[nseq - 1]);
            // The chain itself is serial (one shared-memory lookup per symbol depends on the previous state); the symbol codes and
            // their per-symbol table entries do not: codes arrive 16 per load, one load ahead, and the table entries of four symbols
            // are fetched before the four dependent steps.
            auto step1 = [&](uint32_t i, uint32_t c) {
                const uint32_t nb = (state + ct->delta_nb_bits[c]) >> 16;
                out[i] = (state & ((1u << nb) - 1)) | nb << 16;
                state = ct->next_state[(int32_t)(state >> nb) + ct->delta_find_state[c]];
            };
            int32_t i = (int32_t)nseq - 2;  // next symbol to encode
            while (i >= 0 && (i & 15) != 15) { step1((uint32_t)i, cb[i]); i--; }
            uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
            if (i >= 15) nxt = *reinterpret_cast<const uint4 *>(cb + i - 15);
            while (i >= 15) {
                const uint4 cur = nxt;
                if (i >= 31) nxt = *reinterpret_cast<const uint4 *>(cb + i - 31);
                const uint32_t wv[4] = {cur.w, cur.z, cur.y, cur.x};  // descending symbol order
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t c[4], dn[4]; int32_t df[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) c[k] = wv[q] >> (8 * (3 - k)) & 0xFFu;
#pragma unroll
                    for (int k = 0; k < 4; k++) { dn[k] = ct->delta_nb_bits[c[k]]; df[k] = ct->delta_find_state[c[k]]; }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t nb = (state + dn[k]) >> 16;
                        out[i - 4 * q - k] = (state & ((1u << nb) - 1)) | nb << 16;
                        state = ct->next_state[(int32_t)(state >> nb) + df[k]];
                    }
                }
                i -= 16;
            }
            while (i >= 0) { step1((uint32_t)i, cb[i]); i--; }
        }
        W->scalars[4 + lane] = rle ? 0 : (state & ((1u << ct->tl) - 1)) | ct->tl << 16;  // final state flush
    }
    __syncwarp();
    // segments of sequences in BIT order: lane 0 holds the last sequences (written first)
    const uint32_t q = (nseq + 31) / 32;
    const uint32_t hi = nseq > lane * q ? nseq - lane * q : 0, lo = nseq > (lane + 1) * q ? nseq - (lane + 1) * q : 0;
    const uint32_t *sb_ll = sbits, *sb_of = sbits + nseq, *sb_ml = sbits + 2 * (size_t)nseq;
    const bool owns_end = hi > 0 && lo == 0;
    // pass B1: bits per segment
    uint32_t bits = 0;
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t cd = codes[i];
        bits += zc::ZTAB(LL_bits)[cd & 0xFFu] + zc::ZTAB(ML_bits)[cd >> 16] + (cd >> 8 & 0xFFu);
        if (i + 1 < nseq) bits += (sb_ll[i] >> 16) + (sb_of[i] >> 16) + (sb_ml[i] >> 16);
    }
    if (owns_end) bits += (W->scalars[4] >> 16) + (W->scalars[5] >> 16) + (W->scalars[6] >> 16) + 1;
    uint32_t total_bits;
    const uint32_t start_bit = warp_excl_scan(bits, lane, &total_bits);
    const uint32_t stream_bytes = (total_bits + 7) / 8;
    uint8_t *stream = dst + hdr;
    SegWriter w;
    const bool active = bits > 0;
    if (active) { seg_init(&w, stream, start_bit, start_bit + bits); seg_zero_edges(&w, stream, stream_bytes); }
    __syncwarp();
    __threadfence_block();
    if (active) {
        for (uint32_t i = hi; i > lo; i--) {
            const uint32_t k = i - 1;
            const zc::Seq s = seqs[k];
            const uint32_t cd = codes[k], lc = cd & 0xFFu, oc = cd >> 8 & 0xFFu, mc = cd >> 16;
            if (k + 1 < nseq) {
                const uint32_t a = sb_of[k], b = sb_ml[k], c = sb_ll[k];
                seg_put(&w, a & 0xFFFF, a >> 16);
                seg_put(&w, b & 0xFFFF, b >> 16);
                seg_put(&w, c & 0xFFFF, c >> 16);
            }
            seg_put(&w, s.ll - zc::ZTAB(LL_base)[lc], zc::ZTAB(LL_bits)[lc]);
            seg_put(&w, s.ml - zc::ZTAB(ML_base)[mc], zc::ZTAB(ML_bits)[mc]);
            seg_put(&w, s.off_base - (1u << oc), oc);
        }
        if (owns_end) {
            seg_put(&w, W->scalars[6] & 0xFFFF, W->scalars[6] >> 16);  // ML, OF, LL states, then the sentinel
            seg_put(&w, W->scalars[5] & 0xFFFF, W->scalars[5] >> 16);
            seg_put(&w, W->scalars[4] & 0xFFFF, W->scalars[4] >> 16);
            seg_put(&w, 1, 1);
        }
        seg_finish(&w);
    }
    __syncwarp();
    return hdr + stream_bytes;
}

}  // namespace ent
