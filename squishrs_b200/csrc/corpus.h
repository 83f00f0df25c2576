// Synthetic corpus payload generators (bench/test support).
// Pure integer arithmetic, identical bytes on host and device.  A payload is a stream of
// independent 4 KiB pages; page p of payload `id` depends only on (seed, id, klass, p), so
// the device fills one page per thread and the host fills pages in a loop.
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define SQC_HD __host__ __device__ inline
#else
#define SQC_HD static inline
#endif

#define SQC_PAGE 4096u
enum { SQC_TEXT = 0, SQC_LOG = 1, SQC_JSON = 2, SQC_BINARY = 3, SQC_RANDOM = 4, SQC_ZERO = 5, SQC_FSMIX = 6 };

struct sqc_rng { uint64_t s; };
SQC_HD uint64_t sqc_next(sqc_rng *r) {
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
SQC_HD uint64_t sqc_mix(uint64_t x) {
    sqc_rng r = {x};
    return sqc_next(&r);
}

// byte sink over one page: packs into u64 words, never writes at or past `limit`
struct sqc_out { uint8_t *base; uint32_t pos, limit; uint64_t acc; };
SQC_HD void sqc_flush(sqc_out *o) {
    // called when pos is a multiple of 8 (or at the end of the page)
    uint32_t start = (o->pos - 1) & ~7u;
    if (start + 8 <= o->limit) *(uint64_t *)(o->base + start) = o->acc;
    else for (uint32_t i = start; i < o->limit; i++) o->base[i] = (uint8_t)(o->acc >> (8 * (i - start)));
    o->acc = 0;
}
SQC_HD void sqc_put(sqc_out *o, uint8_t c) {
    if (o->pos >= SQC_PAGE) return;
    o->acc |= (uint64_t)c << (8 * (o->pos & 7));
    o->pos++;
    if ((o->pos & 7) == 0) sqc_flush(o);
}
SQC_HD int sqc_full(const sqc_out *o) { return o->pos >= SQC_PAGE; }
SQC_HD void sqc_puts(sqc_out *o, const char *s) { while (*s) sqc_put(o, (uint8_t)*s++); }
SQC_HD void sqc_putdec(sqc_out *o, uint64_t v, int min_digits) {
    char tmp[20];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n < min_digits) tmp[n++] = '0';
    while (n) sqc_put(o, (uint8_t)tmp[--n]);
}
SQC_HD void sqc_puthex(sqc_out *o, uint64_t v, int digits) {
    for (int i = digits - 1; i >= 0; i--) { uint32_t d = (uint32_t)(v >> (4 * i)) & 15; sqc_put(o, (uint8_t)(d < 10 ? '0' + d : 'a' + d - 10)); }
}
// log-uniform rank in [0, 2^bits): a Zipf(1)-like popularity curve in integers only
SQC_HD uint32_t sqc_zipf(sqc_rng *r, int bits) {
    uint64_t x = sqc_next(r);
    uint32_t b = (uint32_t)(x % (uint32_t)(bits + 1));
    if (b == 0) return 0;
    return (1u << (b - 1)) + ((uint32_t)(x >> 32) & ((1u << (b - 1)) - 1));
}
// pseudo-word for a vocabulary rank: short for popular ranks, lowercase letters
SQC_HD void sqc_word(sqc_out *o, uint32_t rank) {
    uint64_t h = sqc_mix(0x5157ULL + rank);
    int bl = 0;
    for (uint32_t t = rank; t; t >>= 1) bl++;
    int len = 1 + bl / 2 + (int)(h % 3);
    h >>= 8;
    const char *cons = "tnshrdlcmwfgypbvk", *vow = "eaoiu";
    for (int i = 0; i < len; i++) {
        if (i & 1) sqc_put(o, (uint8_t)vow[h % 5]); else sqc_put(o, (uint8_t)cons[h % 17]);
        h = h / 17 + (h << 7);
    }
}

SQC_HD void sqc_page_text(sqc_out *o, sqc_rng *r) {
    uint32_t col = 0;
    while (!sqc_full(o)) {
        uint32_t before = o->pos;
        sqc_word(o, sqc_zipf(r, 12));
        col += o->pos - before + 1;
        if (col >= 72 + (uint32_t)(sqc_next(r) & 15)) { sqc_put(o, '\n'); col = 0; } else sqc_put(o, ' ');
    }
}
SQC_HD void sqc_page_log(sqc_out *o, sqc_rng *r, uint64_t page_no) {
    const char *lvl[4] = {"INFO ", "DEBUG", "WARN ", "ERROR"};
    const char *mod[8] = {"net.http", "db.pool", "auth.session", "cache.lru", "sched.worker", "fs.sync", "rpc.client", "gc.heap"};
    const char *msg[8] = {"request completed", "connection acquired", "token refreshed", "evicted entries", "job dispatched",
                          "flushed dirty pages", "retrying call", "pause finished"};
    uint64_t ts = 1760000000000ULL + page_no * 40000ULL;
    while (!sqc_full(o)) {
        uint64_t x = sqc_next(r);
        ts += x & 1023;
        uint64_t sec = ts / 1000;
        sqc_puts(o, "2025-10-");
        sqc_putdec(o, 1 + (sec / 86400) % 28, 2);
        sqc_put(o, 'T');
        sqc_putdec(o, (sec / 3600) % 24, 2); sqc_put(o, ':');
        sqc_putdec(o, (sec / 60) % 60, 2); sqc_put(o, ':');
        sqc_putdec(o, sec % 60, 2); sqc_put(o, '.');
        sqc_putdec(o, ts % 1000, 3);
        sqc_puts(o, "Z ");
        uint32_t t = (uint32_t)(x >> 10) & 7;
        sqc_puts(o, lvl[((x >> 13) & 15) < 11 ? 0 : ((x >> 13) & 3)]);
        sqc_put(o, ' ');
        sqc_puts(o, mod[t]);
        sqc_puts(o, ": ");
        sqc_puts(o, msg[t]);
        sqc_puts(o, " status=");
        sqc_putdec(o, ((x >> 20) & 31) < 28 ? 200 : 500 + ((x >> 25) & 3), 3);
        sqc_puts(o, " dur_ms=");
        sqc_putdec(o, (x >> 28) & 4095, 1);
        sqc_puts(o, " user=u");
        sqc_putdec(o, sqc_zipf(r, 14), 1);
        sqc_puts(o, " req=");
        sqc_puthex(o, x * 0x9E3779B97F4A7C15ULL, 12);
        sqc_put(o, '\n');
    }
}
SQC_HD void sqc_page_json(sqc_out *o, sqc_rng *r, uint64_t page_no) {
    const char *st[4] = {"ok", "ok", "pending", "failed"};
    const char *kind[4] = {"click", "view", "purchase", "scroll"};
    uint64_t n = page_no * 24;
    while (!sqc_full(o)) {
        uint64_t x = sqc_next(r), y = sqc_next(r);
        sqc_puts(o, "{\"id\":\"");
        sqc_puthex(o, x, 16); sqc_puthex(o, y, 16);
        sqc_puts(o, "\",\"seq\":");
        sqc_putdec(o, n++, 1);
        sqc_puts(o, ",\"ts\":");
        sqc_putdec(o, 1760000000ULL + n * 3 + (x & 3), 1);
        sqc_puts(o, ",\"user\":\"user_");
        sqc_putdec(o, sqc_zipf(r, 16), 1);
        sqc_puts(o, "\",\"event\":\"");
        sqc_puts(o, kind[(y >> 8) & 3]);
        sqc_puts(o, "\",\"status\":\"");
        sqc_puts(o, st[(y >> 12) & 3]);
        sqc_puts(o, "\",\"value\":");
        sqc_putdec(o, (y >> 16) & 1023, 1); sqc_put(o, '.'); sqc_putdec(o, (y >> 28) & 63, 2);
        sqc_puts(o, ",\"tags\":[\"t");
        sqc_putdec(o, sqc_zipf(r, 6), 1);
        sqc_puts(o, "\",\"t");
        sqc_putdec(o, sqc_zipf(r, 6), 1);
        sqc_puts(o, "\"]}\n");
    }
}
SQC_HD void sqc_page_binary(sqc_out *o, sqc_rng *r, uint64_t page_no) {
    // 20-byte records: u32 counter | u16 type | u16 0 | u32 small | 4 random bytes | u32 0
    uint32_t ctr = (uint32_t)(page_no * 205);
    while (!sqc_full(o)) {
        uint64_t x = sqc_next(r);
        uint32_t c = ctr++;
        for (int i = 0; i < 4; i++) sqc_put(o, (uint8_t)(c >> (8 * i)));
        sqc_put(o, (uint8_t)(x & 7)); sqc_put(o, 0); sqc_put(o, 0); sqc_put(o, 0);
        uint32_t small = (uint32_t)(x >> 8) & 0x3FF;
        for (int i = 0; i < 4; i++) sqc_put(o, (uint8_t)(small >> (8 * i)));
        for (int i = 0; i < 4; i++) sqc_put(o, (uint8_t)(x >> (32 + 8 * i)));
        for (int i = 0; i < 4; i++) sqc_put(o, 0);
    }
}

// Fill page `page_no` of payload (seed, id, klass); writes only bytes < limit (limit <= 4096).
SQC_HD void sqc_fill_page(uint8_t *page_base, uint32_t limit, uint64_t seed, uint64_t id, uint32_t klass, uint64_t page_no) {
    sqc_rng r = {seed ^ (id * 0x9E3779B97F4A7C15ULL) ^ (page_no * 0xD1B54A32D192ED03ULL) ^ ((uint64_t)klass << 56)};
    (void)sqc_next(&r);
    sqc_out o = {page_base, 0, limit, 0};
    if (klass == SQC_FSMIX) {
        uint32_t pick = (uint32_t)(sqc_mix(id * 31 + page_no / 16) % 10);  // runs of 64 KiB of one kind
        klass = pick < 4 ? SQC_TEXT : pick < 6 ? SQC_JSON : pick < 9 ? SQC_BINARY : SQC_ZERO;
    }
    switch (klass) {
    case SQC_TEXT: sqc_page_text(&o, &r); break;
    case SQC_LOG: sqc_page_log(&o, &r, page_no); break;
    case SQC_JSON: sqc_page_json(&o, &r, page_no); break;
    case SQC_BINARY: sqc_page_binary(&o, &r, page_no); break;
    case SQC_RANDOM:
        for (uint32_t i = 0; i < SQC_PAGE / 8; i++) { o.acc = sqc_next(&r); o.pos += 8; sqc_flush(&o); }
        break;
    default:
        for (uint32_t i = 0; i < SQC_PAGE / 8; i++) { o.acc = 0; o.pos += 8; sqc_flush(&o); }
        break;
    }
}
