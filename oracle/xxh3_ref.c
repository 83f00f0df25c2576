/* ORACLE — test infrastructure only (see xxh3_ref.h). */
#include "xxh3_ref.h"
#include <string.h>

static const uint8_t kSecret[192] = {
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c,
    0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f,
    0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21,
    0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c,
    0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3,
    0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8,
    0xa8, 0xfa, 0x76, 0x3f, 0xe3, 0x9c, 0x34, 0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d,
    0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31, 0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64,
    0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5, 0x81, 0xa0, 0xff, 0xfa, 0x13, 0x63, 0xeb,
    0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e,
    0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce,
    0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e,
};

#define P32_1 0x9E3779B1U
#define P32_2 0x85EBCA77U
#define P32_3 0xC2B2AE3DU
#define P64_1 0x9E3779B185EBCA87ULL
#define P64_2 0xC2B2AE3D27D4EB4FULL
#define P64_3 0x165667B19E3779F9ULL
#define P64_4 0x85EBCA77C2B2AE63ULL
#define P64_5 0x27D4EB2F165667C5ULL
#define MX1 0x165667919E3779F9ULL
#define MX2 0x9FB21C651E98DF25ULL

typedef struct { uint64_t lo, hi; } u128_t;

static uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static uint64_t rd64(const uint8_t *p) { return (uint64_t)rd32(p) | (uint64_t)rd32(p + 4) << 32; }
static uint32_t bswap32(uint32_t x) { return x >> 24 | (x >> 8 & 0xFF00) | (x << 8 & 0xFF0000) | x << 24; }
static uint64_t bswap64(uint64_t x) { return (uint64_t)bswap32((uint32_t)x) << 32 | bswap32((uint32_t)(x >> 32)); }
static uint32_t rotl32(uint32_t x, int r) { return x << r | x >> (32 - r); }
static u128_t mul128(uint64_t a, uint64_t b) {
    unsigned __int128 m = (unsigned __int128)a * b;
    u128_t r = { (uint64_t)m, (uint64_t)(m >> 64) };
    return r;
}
static uint64_t fold64(uint64_t a, uint64_t b) { u128_t m = mul128(a, b); return m.lo ^ m.hi; }
static uint64_t xxh64_avalanche(uint64_t h) { h ^= h >> 33; h *= P64_2; h ^= h >> 29; h *= P64_3; h ^= h >> 32; return h; }
static uint64_t xxh3_avalanche(uint64_t h) { h ^= h >> 37; h *= MX1; h ^= h >> 32; return h; }

static u128_t len_1to3(const uint8_t *in, size_t len) {
    uint32_t c1 = in[0], c2 = in[len >> 1], c3 = in[len - 1];
    uint32_t cl = c1 << 16 | c2 << 24 | c3 | (uint32_t)len << 8;
    uint32_t ch = rotl32(bswap32(cl), 13);
    uint64_t fl = (uint64_t)(rd32(kSecret) ^ rd32(kSecret + 4));
    uint64_t fh = (uint64_t)(rd32(kSecret + 8) ^ rd32(kSecret + 12));
    u128_t r = { xxh64_avalanche((uint64_t)cl ^ fl), xxh64_avalanche((uint64_t)ch ^ fh) };
    return r;
}
static u128_t len_4to8(const uint8_t *in, size_t len) {
    uint32_t ilo = rd32(in), ihi = rd32(in + len - 4);
    uint64_t i64 = ilo + ((uint64_t)ihi << 32);
    uint64_t flip = rd64(kSecret + 16) ^ rd64(kSecret + 24);
    u128_t m = mul128(i64 ^ flip, P64_1 + ((uint64_t)len << 2));
    m.hi += m.lo << 1;
    m.lo ^= m.hi >> 3;
    m.lo ^= m.lo >> 35; m.lo *= MX2; m.lo ^= m.lo >> 28;
    m.hi = xxh3_avalanche(m.hi);
    return m;
}
static u128_t len_9to16(const uint8_t *in, size_t len) {
    uint64_t fl = rd64(kSecret + 32) ^ rd64(kSecret + 40);
    uint64_t fh = rd64(kSecret + 48) ^ rd64(kSecret + 56);
    uint64_t ilo = rd64(in), ihi = rd64(in + len - 8);
    u128_t m = mul128(ilo ^ ihi ^ fl, P64_1);
    m.lo += (uint64_t)(len - 1) << 54;
    ihi ^= fh;
    m.hi += ihi + (uint64_t)(uint32_t)ihi * (uint64_t)(P32_2 - 1);
    m.lo ^= bswap64(m.hi);
    u128_t h = mul128(m.lo, P64_2);
    h.hi += m.hi * P64_2;
    h.lo = xxh3_avalanche(h.lo);
    h.hi = xxh3_avalanche(h.hi);
    return h;
}
static uint64_t mix16(const uint8_t *in, const uint8_t *sec, uint64_t seed) {
    return fold64(rd64(in) ^ (rd64(sec) + seed), rd64(in + 8) ^ (rd64(sec + 8) - seed));
}
static u128_t mix32(u128_t acc, const uint8_t *a, const uint8_t *b, const uint8_t *sec, uint64_t seed) {
    acc.lo += mix16(a, sec, seed);
    acc.lo ^= rd64(b) + rd64(b + 8);
    acc.hi += mix16(b, sec + 16, seed);
    acc.hi ^= rd64(a) + rd64(a + 8);
    return acc;
}
static u128_t mid_final(u128_t acc, size_t len) {
    u128_t h;
    h.lo = acc.lo + acc.hi;
    h.hi = acc.lo * P64_1 + acc.hi * P64_4 + (uint64_t)len * P64_2;
    h.lo = xxh3_avalanche(h.lo);
    h.hi = 0 - xxh3_avalanche(h.hi);
    return h;
}
static u128_t len_17to128(const uint8_t *in, size_t len) {
    u128_t acc = { (uint64_t)len * P64_1, 0 };
    int i = (int)((len - 1) / 32);
    for (; i >= 0; i--) acc = mix32(acc, in + 16 * i, in + len - 16 * (i + 1), kSecret + 32 * i, 0);
    return mid_final(acc, len);
}
static u128_t len_129to240(const uint8_t *in, size_t len) {
    u128_t acc = { (uint64_t)len * P64_1, 0 };
    unsigned i;
    for (i = 32; i < 160; i += 32) acc = mix32(acc, in + i - 32, in + i - 16, kSecret + i - 32, 0);
    acc.lo = xxh3_avalanche(acc.lo);
    acc.hi = xxh3_avalanche(acc.hi);
    for (i = 160; i <= len; i += 32) acc = mix32(acc, in + i - 32, in + i - 16, kSecret + 3 + i - 160, 0);
    acc = mix32(acc, in + len - 16, in + len - 32, kSecret + 136 - 17 - 16, 0);
    return mid_final(acc, len);
}

static void accumulate_stripe(uint64_t acc[8], const uint8_t *in, const uint8_t *sec) {
    for (int i = 0; i < 8; i++) {
        uint64_t v = rd64(in + 8 * i);
        uint64_t k = v ^ rd64(sec + 8 * i);
        acc[i ^ 1] += v;
        acc[i] += (k & 0xFFFFFFFFULL) * (k >> 32);
    }
}
static void scramble(uint64_t acc[8], const uint8_t *sec) {
    for (int i = 0; i < 8; i++) {
        uint64_t a = acc[i];
        a ^= a >> 47;
        a ^= rd64(sec + 8 * i);
        acc[i] = a * P32_1;
    }
}
static uint64_t merge_accs(const uint64_t acc[8], const uint8_t *sec, uint64_t start) {
    uint64_t r = start;
    for (int j = 0; j < 4; j++) r += fold64(acc[2 * j] ^ rd64(sec + 16 * j), acc[2 * j + 1] ^ rd64(sec + 16 * j + 8));
    return xxh3_avalanche(r);
}
static u128_t hash_long(const uint8_t *in, size_t len) {
    uint64_t acc[8] = { P32_3, P64_1, P64_2, P64_3, P64_4, P32_2, P64_5, P32_1 };
    const size_t stripes_per_block = (192 - 64) / 8; /* 16 */
    const size_t block_len = 64 * stripes_per_block; /* 1024 */
    size_t nb_blocks = (len - 1) / block_len;
    for (size_t n = 0; n < nb_blocks; n++) {
        for (size_t s = 0; s < stripes_per_block; s++) accumulate_stripe(acc, in + n * block_len + 64 * s, kSecret + 8 * s);
        scramble(acc, kSecret + 192 - 64);
    }
    size_t nb_stripes = ((len - 1) - block_len * nb_blocks) / 64;
    for (size_t s = 0; s < nb_stripes; s++) accumulate_stripe(acc, in + nb_blocks * block_len + 64 * s, kSecret + 8 * s);
    accumulate_stripe(acc, in + len - 64, kSecret + 192 - 64 - 7);
    u128_t h;
    h.lo = merge_accs(acc, kSecret + 11, (uint64_t)len * P64_1);
    h.hi = merge_accs(acc, kSecret + 192 - 64 - 11, ~((uint64_t)len * P64_2));
    return h;
}

void sqo_xxh3_128(const uint8_t *data, size_t len, uint64_t *low64, uint64_t *high64) {
    u128_t h;
    if (len == 0) {
        h.lo = xxh64_avalanche(rd64(kSecret + 64) ^ rd64(kSecret + 72));
        h.hi = xxh64_avalanche(rd64(kSecret + 80) ^ rd64(kSecret + 88));
    } else if (len <= 3) h = len_1to3(data, len);
    else if (len <= 8) h = len_4to8(data, len);
    else if (len <= 16) h = len_9to16(data, len);
    else if (len <= 128) h = len_17to128(data, len);
    else if (len <= 240) h = len_129to240(data, len);
    else h = hash_long(data, len);
    *low64 = h.lo;
    *high64 = h.hi;
}

void sqo_hash_chunk(const uint8_t *data, size_t len, uint8_t out16[16]) {
    uint64_t lo, hi;
    sqo_xxh3_128(data, len, &lo, &hi);
    for (int i = 0; i < 8; i++) { out16[i] = (uint8_t)(lo >> (8 * i)); out16[8 + i] = (uint8_t)(hi >> (8 * i)); }
}
