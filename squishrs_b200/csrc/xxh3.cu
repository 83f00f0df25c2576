// K1 — XXH3-128 (seed 0, default secret) digest kernel for sm_100a.
//
// Replaces hash_chunk (reference src/util/chunk.rs:46-49): one 16-byte digest per
// chunk, stored low64 LE || high64 LE.  HBM-bound: algorithmic traffic = the chunk
// bytes, read exactly once.
//
// Mapping.  XXH3's long path consumes 1 KiB "blocks" of 16 stripes x 8 u64 lanes.
// Inside a block every stripe contributes  acc[i^1] += v ; acc[i] += lo32(v^k)*hi32(v^k)
// -- independent of acc, wrapping adds -- so a block is a 128-way parallel sum, and only
// the per-block scramble is a serial chain.  One WARP owns one chunk: lane l loads the
// 16 B at byte 16*l of each 512 B half block (one coalesced LDG.128 per lane, a full
// 512 B per warp request), i.e. stripe l/4 (+8 for the second half), u64 lanes
// 2*(l%4), 2*(l%4)+1, with its four secret words fixed in registers for the whole chunk.
// Three xor-butterfly shuffle steps (4,8,16) sum the 8 lanes that share an acc pair, then
// every lane scrambles its copy of the pair.  Warps are persistent and pull chunk indices
// from an atomic counter so ragged tails balance.  Four blocks (8 x LDG.128 per lane,
// 4 KiB per warp) are kept in flight by a register double buffer.
#include "common.cuh"

namespace {

__constant__ uint64_t c_secret64[24];   // LE64(kSecret + 8k)
__constant__ uint64_t c_secret_last[8]; // LE64(kSecret + 121 + 8i): last-stripe secret (192-64-7)
__constant__ uint8_t c_secret[192];
__constant__ uint64_t c_acc_init[8];

const uint8_t h_secret[192] = {
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

constexpr uint32_t P32_1 = 0x9E3779B1U, P32_2 = 0x85EBCA77U, P32_3 = 0xC2B2AE3DU;
constexpr uint64_t P64_1 = 0x9E3779B185EBCA87ULL, P64_2 = 0xC2B2AE3D27D4EB4FULL, P64_3 = 0x165667B19E3779F9ULL,
                   P64_4 = 0x85EBCA77C2B2AE63ULL, P64_5 = 0x27D4EB2F165667C5ULL, MX1 = 0x165667919E3779F9ULL,
                   MX2 = 0x9FB21C651E98DF25ULL;

__device__ __forceinline__ uint64_t rd64u(const uint8_t *p) {  // unaligned LE64 from global
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
    return v;
}
__device__ __forceinline__ uint32_t rd32u(const uint8_t *p) {
    return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
}
__device__ __forceinline__ uint64_t sec64(int off) { return rd64u(c_secret + off); }
__device__ __forceinline__ uint64_t fold64(uint64_t a, uint64_t b) { return (a * b) ^ __umul64hi(a, b); }
__device__ __forceinline__ uint64_t avalanche3(uint64_t h) { h ^= h >> 37; h *= MX1; h ^= h >> 32; return h; }
__device__ __forceinline__ uint64_t avalanche64(uint64_t h) { h ^= h >> 33; h *= P64_2; h ^= h >> 29; h *= P64_3; h ^= h >> 32; return h; }
__device__ __forceinline__ uint64_t mul32x32(uint64_t k) { return (uint64_t)(uint32_t)k * (uint64_t)(uint32_t)(k >> 32); }

// ---- lengths 1..240: one lane, straight restatement of the short paths -------------
__device__ uint64_t mix16(const uint8_t *in, int so) {
    return fold64(rd64u(in) ^ sec64(so), rd64u(in + 8) ^ sec64(so + 8));
}
__device__ void mix32(uint64_t &lo, uint64_t &hi, const uint8_t *a, const uint8_t *b, int so) {
    lo += mix16(a, so);
    lo ^= rd64u(b) + rd64u(b + 8);
    hi += mix16(b, so + 16);
    hi ^= rd64u(a) + rd64u(a + 8);
}
__device__ void xxh3_short(const uint8_t *in, uint32_t len, uint64_t &lo, uint64_t &hi) {
    if (len == 0) {
        lo = avalanche64(sec64(64) ^ sec64(72));
        hi = avalanche64(sec64(80) ^ sec64(88));
    } else if (len <= 3) {
        uint32_t c1 = in[0], c2 = in[len >> 1], c3 = in[len - 1];
        uint32_t cl = c1 << 16 | c2 << 24 | c3 | len << 8;
        uint32_t ch = __funnelshift_l(__byte_perm(cl, 0, 0x0123), __byte_perm(cl, 0, 0x0123), 13);
        lo = avalanche64((uint64_t)cl ^ (uint64_t)(rd32u(c_secret) ^ rd32u(c_secret + 4)));
        hi = avalanche64((uint64_t)ch ^ (uint64_t)(rd32u(c_secret + 8) ^ rd32u(c_secret + 12)));
    } else if (len <= 8) {
        uint64_t i64 = (uint64_t)rd32u(in) + ((uint64_t)rd32u(in + len - 4) << 32);
        uint64_t keyed = i64 ^ (sec64(16) ^ sec64(24));
        uint64_t m = P64_1 + ((uint64_t)len << 2);
        uint64_t mlo = keyed * m, mhi = __umul64hi(keyed, m);
        mhi += mlo << 1;
        mlo ^= mhi >> 3;
        mlo ^= mlo >> 35; mlo *= MX2; mlo ^= mlo >> 28;
        lo = mlo;
        hi = avalanche3(mhi);
    } else if (len <= 16) {
        uint64_t fl = sec64(32) ^ sec64(40), fh = sec64(48) ^ sec64(56);
        uint64_t ilo = rd64u(in), ihi = rd64u(in + len - 8);
        uint64_t x = ilo ^ ihi ^ fl;
        uint64_t mlo = x * P64_1, mhi = __umul64hi(x, P64_1);
        mlo += (uint64_t)(len - 1) << 54;
        ihi ^= fh;
        mhi += ihi + (uint64_t)(uint32_t)ihi * (uint64_t)(P32_2 - 1);
        mlo ^= ((uint64_t)__byte_perm((uint32_t)mhi, 0, 0x0123) << 32) | __byte_perm((uint32_t)(mhi >> 32), 0, 0x0123);
        uint64_t hlo = mlo * P64_2, hhi = __umul64hi(mlo, P64_2) + mhi * P64_2;
        lo = avalanche3(hlo);
        hi = avalanche3(hhi);
    } else {
        uint64_t alo = (uint64_t)len * P64_1, ahi = 0;
        if (len <= 128) {
            for (int i = (int)((len - 1) / 32); i >= 0; i--) mix32(alo, ahi, in + 16 * i, in + len - 16 * (i + 1), 32 * i);
        } else {
            for (uint32_t i = 32; i < 160; i += 32) mix32(alo, ahi, in + i - 32, in + i - 16, (int)i - 32);
            alo = avalanche3(alo);
            ahi = avalanche3(ahi);
            for (uint32_t i = 160; i <= len; i += 32) mix32(alo, ahi, in + i - 32, in + i - 16, 3 + (int)i - 160);
            mix32(alo, ahi, in + len - 16, in + len - 32, 136 - 17 - 16);
        }
        lo = avalanche3(alo + ahi);
        hi = 0 - avalanche3(alo * P64_1 + ahi * P64_4 + (uint64_t)len * P64_2);
    }
}

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {  // read-once data: keep it out of L1
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
template <bool ALIGNED>
__device__ __forceinline__ void load16(const uint8_t *p, uint64_t &v0, uint64_t &v1) {
    if (ALIGNED) {
        uint4 q = ldg_stream(reinterpret_cast<const uint4 *>(p));
        v0 = (uint64_t)q.x | (uint64_t)q.y << 32;
        v1 = (uint64_t)q.z | (uint64_t)q.w << 32;
    } else {
        v0 = rd64u(p);
        v1 = rd64u(p + 8);
    }
}
__device__ __forceinline__ uint64_t shfl_xor64(uint64_t v, int m) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m), hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
    return (uint64_t)hi << 32 | lo;
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return (uint64_t)hi << 32 | lo;
}
__device__ __forceinline__ void stripe_acc(uint64_t &a0, uint64_t &a1, uint64_t v0, uint64_t v1, uint64_t k0, uint64_t k1) {
    a0 += v1 + mul32x32(v0 ^ k0);
    a1 += v0 + mul32x32(v1 ^ k1);
}

// One warp, one chunk, len > 240.
template <bool ALIGNED>
__device__ void xxh3_long_warp(const uint8_t *in, uint32_t len, int lane, uint64_t &lo, uint64_t &hi) {
    const int s = lane >> 2, j = lane & 3;
    const uint64_t k00 = c_secret64[s + 2 * j], k01 = c_secret64[s + 2 * j + 1];
    const uint64_t k10 = c_secret64[s + 8 + 2 * j], k11 = c_secret64[s + 8 + 2 * j + 1];
    const uint64_t ks0 = c_secret64[16 + 2 * j], ks1 = c_secret64[16 + 2 * j + 1];  // scramble secret (offset 128)
    uint64_t acc0 = c_acc_init[2 * j], acc1 = c_acc_init[2 * j + 1];  // every lane of an acc-pair group holds the same copy

    const uint32_t nb_blocks = (len - 1) >> 10;
    const uint8_t *p = in + 16 * lane;
    uint32_t b = 0;
    // main loop: two blocks per iteration, software-pipelined in registers: the loads of the NEXT two blocks are
    // issued before the current two are reduced, so 8 x 16 B per lane (4 KiB per warp) are in flight
    uint64_t v0 = 0, v1 = 0, w0 = 0, w1 = 0, x0 = 0, x1 = 0, y0 = 0, y1 = 0;
    if (nb_blocks >= 2) {
        load16<ALIGNED>(p, v0, v1);
        load16<ALIGNED>(p + 512, w0, w1);
        load16<ALIGNED>(p + 1024, x0, x1);
        load16<ALIGNED>(p + 1536, y0, y1);
    }
    for (; b + 2 <= nb_blocks; b += 2, p += 2048) {
        uint64_t nv0 = 0, nv1 = 0, nw0 = 0, nw1 = 0, nx0 = 0, nx1 = 0, ny0 = 0, ny1 = 0;
        if (b + 4 <= nb_blocks) {
            load16<ALIGNED>(p + 2048, nv0, nv1);
            load16<ALIGNED>(p + 2560, nw0, nw1);
            load16<ALIGNED>(p + 3072, nx0, nx1);
            load16<ALIGNED>(p + 3584, ny0, ny1);
        }
        uint64_t a0 = 0, a1 = 0;
        stripe_acc(a0, a1, v0, v1, k00, k01);
        stripe_acc(a0, a1, w0, w1, k10, k11);
#pragma unroll
        for (int m = 4; m <= 16; m <<= 1) { a0 += shfl_xor64(a0, m); a1 += shfl_xor64(a1, m); }
        acc0 += a0; acc1 += a1;
        acc0 = ((acc0 ^ (acc0 >> 47)) ^ ks0) * P32_1;
        acc1 = ((acc1 ^ (acc1 >> 47)) ^ ks1) * P32_1;
        a0 = 0; a1 = 0;
        stripe_acc(a0, a1, x0, x1, k00, k01);
        stripe_acc(a0, a1, y0, y1, k10, k11);
#pragma unroll
        for (int m = 4; m <= 16; m <<= 1) { a0 += shfl_xor64(a0, m); a1 += shfl_xor64(a1, m); }
        acc0 += a0; acc1 += a1;
        acc0 = ((acc0 ^ (acc0 >> 47)) ^ ks0) * P32_1;
        acc1 = ((acc1 ^ (acc1 >> 47)) ^ ks1) * P32_1;
        v0 = nv0; v1 = nv1; w0 = nw0; w1 = nw1; x0 = nx0; x1 = nx1; y0 = ny0; y1 = ny1;
    }
    for (; b < nb_blocks; b++, p += 1024) {
        uint64_t v0, v1, w0, w1;
        load16<ALIGNED>(p, v0, v1);
        load16<ALIGNED>(p + 512, w0, w1);
        uint64_t a0 = 0, a1 = 0;
        stripe_acc(a0, a1, v0, v1, k00, k01);
        stripe_acc(a0, a1, w0, w1, k10, k11);
#pragma unroll
        for (int m = 4; m <= 16; m <<= 1) { a0 += shfl_xor64(a0, m); a1 += shfl_xor64(a1, m); }
        acc0 += a0; acc1 += a1;
        acc0 = ((acc0 ^ (acc0 >> 47)) ^ ks0) * P32_1;
        acc1 = ((acc1 ^ (acc1 >> 47)) ^ ks1) * P32_1;
    }
    // tail: nb_stripes full stripes of the last (partial) block, then the last 64 input bytes
    {
        const uint32_t nb_stripes = ((len - 1) - (nb_blocks << 10)) >> 6;  // 0..15
        uint64_t a0 = 0, a1 = 0;
        if ((uint32_t)s < nb_stripes) {
            uint64_t v0, v1;
            load16<ALIGNED>(p, v0, v1);
            stripe_acc(a0, a1, v0, v1, k00, k01);
        }
        if ((uint32_t)(s + 8) < nb_stripes) {
            uint64_t w0, w1;
            load16<ALIGNED>(p + 512, w0, w1);
            stripe_acc(a0, a1, w0, w1, k10, k11);
        }
        if (lane < 4) {  // last stripe: in + len - 64, secret offset 121 (generally unaligned)
            const uint8_t *q = in + len - 64 + 16 * j;
            uint64_t v0 = rd64u(q), v1 = rd64u(q + 8);
            stripe_acc(a0, a1, v0, v1, c_secret_last[2 * j], c_secret_last[2 * j + 1]);
        }
#pragma unroll
        for (int m = 4; m <= 16; m <<= 1) { a0 += shfl_xor64(a0, m); a1 += shfl_xor64(a1, m); }
        acc0 += a0; acc1 += a1;
    }
    // finalize: gather the 8 accumulators (lane j holds acc[2j], acc[2j+1])
    uint64_t acc[8];
#pragma unroll
    for (int q = 0; q < 4; q++) { acc[2 * q] = shfl64(acc0, q); acc[2 * q + 1] = shfl64(acc1, q); }
    if (lane == 0) {
        uint64_t rl = (uint64_t)len * P64_1, rh = ~((uint64_t)len * P64_2);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            rl += fold64(acc[2 * q] ^ sec64(11 + 16 * q), acc[2 * q + 1] ^ sec64(11 + 16 * q + 8));
            rh += fold64(acc[2 * q] ^ sec64(192 - 64 - 11 + 16 * q), acc[2 * q + 1] ^ sec64(192 - 64 - 11 + 16 * q + 8));
        }
        lo = avalanche3(rl);
        hi = avalanche3(rh);
    }
}

__global__ void __launch_bounds__(256, 3) xxh3_128_kernel(const uint8_t *__restrict__ data, const sq_span *__restrict__ spans,
                                                         uint32_t n, uint8_t *__restrict__ out, uint32_t *__restrict__ counter) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = atomicAdd(counter, 1u);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= n) break;
        const sq_span sp = spans[idx];
        const uint8_t *in = data + sp.off;
        uint64_t lo = 0, hi = 0;
        if (sp.len <= 240) {
            if (lane == 0) xxh3_short(in, sp.len, lo, hi);
        } else if ((reinterpret_cast<uintptr_t>(in) & 15) == 0) {
            xxh3_long_warp<true>(in, sp.len, lane, lo, hi);
        } else {
            xxh3_long_warp<false>(in, sp.len, lane, lo, hi);
        }
        if (lane == 0) {
            uint2 *o = reinterpret_cast<uint2 *>(out + (size_t)idx * 16);  // low64 LE || high64 LE
            o[0] = make_uint2((uint32_t)lo, (uint32_t)(lo >> 32));
            o[1] = make_uint2((uint32_t)hi, (uint32_t)(hi >> 32));
        }
    }
}

}  // namespace

int32_t sq_xxh3_init(sq_ctx *ctx) {
    uint64_t s64[24], slast[8];
    for (int k = 0; k < 24; k++) memcpy(&s64[k], h_secret + 8 * k, 8);
    for (int i = 0; i < 8; i++) memcpy(&slast[i], h_secret + 121 + 8 * i, 8);
    SQ_CUDA(ctx, cudaMemcpyToSymbol(c_secret64, s64, sizeof s64));
    SQ_CUDA(ctx, cudaMemcpyToSymbol(c_secret_last, slast, sizeof slast));
    SQ_CUDA(ctx, cudaMemcpyToSymbol(c_secret, h_secret, sizeof h_secret));
    const uint64_t init[8] = {P32_3, P64_1, P64_2, P64_3, P64_4, P32_2, P64_5, P32_1};
    SQ_CUDA(ctx, cudaMemcpyToSymbol(c_acc_init, init, sizeof init));
    return SQ_OK;
}

extern "C" int32_t sq_digest_device(sq_ctx *ctx, const void *d_data, const sq_span *d_spans, uint32_t n, void *d_digests,
                                    void *stream) {
    if (!ctx) return SQ_ERR_INVALID_ARG;
    if (n == 0) return SQ_OK;
    if (!d_data || !d_spans || !d_digests) return sq_set_error(ctx, SQ_ERR_INVALID_ARG, "sq_digest_device: null pointer");
    cudaStream_t st = sq_stream(ctx, stream);
    SQ_CUDA(ctx, cudaMemsetAsync(ctx->d_work_counter, 0, sizeof(uint32_t), st));
    // persistent warps: 8 warps per CTA, up to 8 CTAs per SM, never more warps than chunks
    uint32_t ctas = (n + 7) / 8, max_ctas = (uint32_t)ctx->sm_count * 3;
    if (ctas > max_ctas) ctas = max_ctas;
    xxh3_128_kernel<<<ctas, 256, 0, st>>>((const uint8_t *)d_data, d_spans, n, (uint8_t *)d_digests, ctx->d_work_counter);
    SQ_LAUNCHED(ctx, 1);
    SQ_CUDA(ctx, cudaGetLastError());
    return SQ_OK;
}
