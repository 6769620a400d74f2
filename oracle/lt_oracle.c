/* oracle/lt_oracle.c — CPU restatement of longtail's chunk -> hash -> compress indexing path.
 *
 * TEST INFRASTRUCTURE ONLY (see lt_oracle.h).  Plain C, single-threaded, written from the
 * behavioural description in SURVEY.md Appendix A; each function cites the reference
 * file:line it follows.  Parity status: PINNED (tests/test_oracle.py).
 */
#include "lt_oracle.h"

#include <errno.h>
#include <stdlib.h>
#include <string.h>

void lto_free(void* p) { free(p); }

static inline uint32_t rotl32(uint32_t x, unsigned r) { r &= 31u; return r ? (x << r) | (x >> (32u - r)) : x; }
static inline uint32_t rotr32(uint32_t x, unsigned r) { r &= 31u; return r ? (x >> r) | (x << (32u - r)) : x; }
static inline uint32_t load32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint64_t load64(const uint8_t* p) { return (uint64_t)load32(p) | ((uint64_t)load32(p + 4) << 32); }

/* ================================================================ HPCDC chunker */

#define HPCDC_WINDOW 48u /* longtail_hpcdcchunker.c:12 */

static const uint32_t hpcdc_table[256] = {
#include "hpcdc_table.inc"
};

uint32_t lto_hpcdc_discriminator(uint32_t avg)
{
    /* longtail_hpcdcchunker.c:126-129 — evaluated in double, truncated */
    double a = (double)avg;
    return (uint32_t)(a / (-1.42888852e-7 * a + 1.33237515));
}

uint32_t lto_hpcdc_window_hash(const uint8_t* end)
{
    /* seed loop longtail_hpcdcchunker.c:273-279; the rolling update :295-297 keeps exactly this
     * value for the window ending at every later position */
    const uint8_t* w = end - HPCDC_WINDOW;
    uint32_t h = 0;
    for (uint32_t i = 0; i < HPCDC_WINDOW; ++i)
        h ^= rotl32(hpcdc_table[w[i]], (HPCDC_WINDOW - i - 1u) & 31u);
    return h;
}

int lto_hpcdc_chunk(const uint8_t* data, uint64_t size, uint32_t min, uint32_t avg, uint32_t max,
                    uint32_t* out_lens, uint64_t cap, uint64_t* out_count)
{
    /* longtail_hpcdcchunker.c:146-150 */
    if (min < HPCDC_WINDOW || min > max || min > avg || avg > max) return EINVAL;
    const uint32_t d = lto_hpcdc_discriminator(avg);
    if (d == 0) return EINVAL;
    uint64_t s = 0;
    uint64_t n = 0;
    while (s < size) /* :250-255 all done when nothing is left */
    {
        uint64_t left = size - s;
        uint32_t len;
        if (left <= min) /* :257-264 */
            len = (uint32_t)left;
        else
        {
            uint32_t lim = left > max ? max : (uint32_t)left; /* :285 */
            const uint8_t* b = data + s;
            /* :273-306 rolling form, kept literal here (the CUDA path uses the stateless form) */
            uint32_t h = 0;
            for (uint32_t i = 0; i < HPCDC_WINDOW; ++i)
                h ^= rotl32(hpcdc_table[b[min - HPCDC_WINDOW + i]], (HPCDC_WINDOW - i - 1u) & 31u);
            uint32_t pos = min;
            while (pos < lim)
            {
                uint8_t in = b[pos];
                uint8_t out = b[pos - HPCDC_WINDOW];
                ++pos;
                h = rotl32(h, 1) ^ rotl32(hpcdc_table[out], HPCDC_WINDOW & 31u) ^ hpcdc_table[in];
                if (h % d == d - 1) break;
            }
            len = pos;
        }
        if (n >= cap) return ENOSPC;
        out_lens[n++] = len;
        s += len;
    }
    *out_count = n;
    return 0;
}

/* ================================================================ BLAKE3 (lib/blake3/ext) */

static const uint32_t B3_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                  0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u}; /* blake3_impl.h:76-78 */
static const uint8_t B3_PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8}; /* blake3_impl.h:80-95 */
enum { B3_CHUNK_START = 1, B3_CHUNK_END = 2, B3_PARENT = 4, B3_ROOT = 8 }; /* blake3_impl.h:13-21 */

#define B3_G(a, b, c, d, x, y)                 \
    do {                                       \
        a = a + b + (x); d = rotr32(d ^ a, 16); \
        c = c + d;       b = rotr32(b ^ c, 12); \
        a = a + b + (y); d = rotr32(d ^ a, 8);  \
        c = c + d;       b = rotr32(b ^ c, 7);  \
    } while (0)

/* blake3_portable.c:46-122: out[0..8) = truncated compression output (the chaining value) */
static void b3_compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter,
                        uint32_t block_len, uint32_t flags, uint32_t out[8])
{
    uint32_t v[16], m[16], t[16];
    memcpy(m, block, sizeof(m));
    for (int i = 0; i < 8; ++i) v[i] = cv[i];
    for (int i = 0; i < 4; ++i) v[8 + i] = B3_IV[i];
    v[12] = (uint32_t)counter;
    v[13] = (uint32_t)(counter >> 32);
    v[14] = block_len;
    v[15] = flags;
    for (int r = 0; r < 7; ++r)
    {
        B3_G(v[0], v[4], v[8], v[12], m[0], m[1]);
        B3_G(v[1], v[5], v[9], v[13], m[2], m[3]);
        B3_G(v[2], v[6], v[10], v[14], m[4], m[5]);
        B3_G(v[3], v[7], v[11], v[15], m[6], m[7]);
        B3_G(v[0], v[5], v[10], v[15], m[8], m[9]);
        B3_G(v[1], v[6], v[11], v[12], m[10], m[11]);
        B3_G(v[2], v[7], v[8], v[13], m[12], m[13]);
        B3_G(v[3], v[4], v[9], v[14], m[14], m[15]);
        for (int i = 0; i < 16; ++i) t[i] = m[B3_PERM[i]];
        memcpy(m, t, sizeof(m));
    }
    for (int i = 0; i < 8; ++i) out[i] = v[i] ^ v[i + 8];
}

/* one 1 KiB (or shorter, final) leaf: blake3.c:89-116 chunk state */
static void b3_leaf(const uint8_t* p, uint64_t len, uint64_t leaf_index, uint32_t extra_flags, uint32_t out[8])
{
    uint32_t cv[8];
    memcpy(cv, B3_IV, sizeof(cv));
    uint64_t blocks = len ? (len + 63) / 64 : 1;
    for (uint64_t b = 0; b < blocks; ++b)
    {
        uint8_t buf[64];
        uint64_t n = len - b * 64 > 64 ? 64 : len - b * 64;
        memset(buf, 0, sizeof(buf));
        if (n) memcpy(buf, p + b * 64, n);
        uint32_t w[16];
        for (int i = 0; i < 16; ++i) w[i] = load32(buf + 4 * i);
        uint32_t flags = 0;
        if (b == 0) flags |= B3_CHUNK_START;
        if (b == blocks - 1) flags |= B3_CHUNK_END | extra_flags;
        b3_compress(cv, w, leaf_index, (uint32_t)n, flags, cv);
    }
    memcpy(out, cv, sizeof(cv));
}

/* subtree over leaves [first, first+count): left side takes the largest power of two strictly
 * below count (blake3.c:161-166) */
static void b3_subtree(const uint8_t* data, uint64_t len, uint64_t first_leaf, uint64_t leaf_count, int is_root, uint32_t out[8])
{
    if (leaf_count == 1)
    {
        uint64_t off = first_leaf * 1024;
        uint64_t n = len - off > 1024 ? 1024 : len - off;
        b3_leaf(data + off, n, first_leaf, is_root ? B3_ROOT : 0, out);
        return;
    }
    uint64_t left = 1;
    while (left * 2 < leaf_count) left *= 2;
    uint32_t block[16];
    b3_subtree(data, len, first_leaf, left, 0, block);
    b3_subtree(data, len, first_leaf + left, leaf_count - left, 0, block + 8);
    b3_compress(B3_IV, block, 0, 64, B3_PARENT | (is_root ? B3_ROOT : 0), out);
}

uint64_t lto_blake3_64(const void* data, uint64_t len)
{
    uint64_t leaves = len ? (len + 1023) / 1024 : 1;
    uint32_t out[8];
    b3_subtree((const uint8_t*)data, len, 0, leaves, 1, out);
    return (uint64_t)out[0] | ((uint64_t)out[1] << 32); /* longtail_blake3.c:97-100 */
}

/* ================================================================ BLAKE2s-64 (lib/blake2/ext/blake2s.c) */

static const uint8_t B2_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

static void b2s_compress(uint32_t h[8], const uint8_t block[64], uint64_t t, int last)
{
    uint32_t m[16], v[16];
    for (int i = 0; i < 16; ++i) m[i] = load32(block + 4 * i);
    for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[8 + i] = B3_IV[i]; } /* same IV as BLAKE3, blake2s.c:42 */
    v[12] ^= (uint32_t)t;
    v[13] ^= (uint32_t)(t >> 32);
    if (last) v[14] ^= 0xFFFFFFFFu;
    for (int r = 0; r < 10; ++r)
    {
        const uint8_t* s = B2_SIGMA[r];
        B3_G(v[0], v[4], v[8], v[12], m[s[0]], m[s[1]]);
        B3_G(v[1], v[5], v[9], v[13], m[s[2]], m[s[3]]);
        B3_G(v[2], v[6], v[10], v[14], m[s[4]], m[s[5]]);
        B3_G(v[3], v[7], v[11], v[15], m[s[6]], m[s[7]]);
        B3_G(v[0], v[5], v[10], v[15], m[s[8]], m[s[9]]);
        B3_G(v[1], v[6], v[11], v[12], m[s[10]], m[s[11]]);
        B3_G(v[2], v[7], v[8], v[13], m[s[12]], m[s[13]]);
        B3_G(v[3], v[4], v[9], v[14], m[s[14]], m[s[15]]);
    }
    for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
}

uint64_t lto_blake2s_64(const void* data, uint64_t len)
{
    const uint8_t* p = (const uint8_t*)data;
    uint32_t h[8];
    memcpy(h, B3_IV, sizeof(h));
    h[0] ^= 0x01010000u ^ 8u; /* digest_length 8, fanout 1, depth 1 (blake2s.c:85-103) */
    uint64_t t = 0;
    while (len - t > 64)
    {
        b2s_compress(h, p + t, t + 64, 0);
        t += 64;
    }
    uint8_t last[64];
    memset(last, 0, sizeof(last));
    if (len - t) memcpy(last, p + t, len - t);
    b2s_compress(h, last, len, 1);
    return (uint64_t)h[0] | ((uint64_t)h[1] << 32);
}


/* ================================================================ Meow hash 0.5/calico, low 64 bits
 * (lib/meowhash/ext/meow_hash_x64_aesni.h:181-198 macros, :234 seed, :480-560 absorb, :583-700 MeowEnd;
 *  lib/meowhash/longtail_meowhash.c:43-50).  The x86 AESDEC round is restated with one inverse T-table. */

typedef struct { uint32_t w[4]; } m128; /* little-endian words of one xmm register */

static uint32_t MEOW_TD0[256];
static int meow_tables_ready;

static uint8_t gf_mul(uint8_t a, uint8_t b)
{
    uint8_t r = 0;
    while (b)
    {
        if (b & 1) r ^= a;
        a = (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1b : 0));
        b >>= 1;
    }
    return r;
}

static void meow_init_tables(void)
{
    if (meow_tables_ready) return;
    uint8_t sbox[256], inv[256];
    /* AES S-box: multiplicative inverse in GF(2^8) followed by the affine map (FIPS-197 5.1.1) */
    for (int x = 0; x < 256; ++x)
    {
        uint8_t y = 0;
        if (x) for (int c = 1; c < 256; ++c) if (gf_mul((uint8_t)x, (uint8_t)c) == 1) { y = (uint8_t)c; break; }
        uint8_t z = y;
        for (int k = 1; k <= 4; ++k) z ^= (uint8_t)((y << k) | (y >> (8 - k)));
        sbox[x] = z ^ 0x63;
    }
    for (int x = 0; x < 256; ++x) inv[sbox[x]] = (uint8_t)x;
    /* column of InvMixColumns hit by a row-0 byte: (0e, 09, 0d, 0b) (FIPS-197 5.3.3) */
    for (int x = 0; x < 256; ++x)
    {
        uint8_t y = inv[x];
        MEOW_TD0[x] = (uint32_t)gf_mul(y, 0x0e) | ((uint32_t)gf_mul(y, 0x09) << 8) | ((uint32_t)gf_mul(y, 0x0d) << 16) | ((uint32_t)gf_mul(y, 0x0b) << 24);
    }
    meow_tables_ready = 1;
}

/* _mm_aesdec_si128(a, k) = InvMixColumns(InvSubBytes(InvShiftRows(a))) ^ k; byte 4c+r of the register is state[r][c] */
static m128 meow_aesdec(m128 a, m128 k)
{
    m128 o;
    for (int c = 0; c < 4; ++c)
    {
        uint32_t b0 = a.w[c] & 0xff, b1 = (a.w[(c + 3) & 3] >> 8) & 0xff, b2 = (a.w[(c + 2) & 3] >> 16) & 0xff, b3 = a.w[(c + 1) & 3] >> 24;
        o.w[c] = MEOW_TD0[b0] ^ rotl32(MEOW_TD0[b1], 8) ^ rotl32(MEOW_TD0[b2], 16) ^ rotl32(MEOW_TD0[b3], 24) ^ k.w[c];
    }
    return o;
}
static m128 meow_paddq(m128 a, m128 b)
{
    uint64_t a0 = a.w[0] | ((uint64_t)a.w[1] << 32), a1 = a.w[2] | ((uint64_t)a.w[3] << 32);
    uint64_t b0 = b.w[0] | ((uint64_t)b.w[1] << 32), b1 = b.w[2] | ((uint64_t)b.w[3] << 32);
    a0 += b0; a1 += b1;
    m128 o = {{(uint32_t)a0, (uint32_t)(a0 >> 32), (uint32_t)a1, (uint32_t)(a1 >> 32)}};
    return o;
}
static m128 meow_pxor(m128 a, m128 b) { m128 o = {{a.w[0] ^ b.w[0], a.w[1] ^ b.w[1], a.w[2] ^ b.w[2], a.w[3] ^ b.w[3]}}; return o; }
static m128 meow_load(const uint8_t* p) { m128 o = {{load32(p), load32(p + 4), load32(p + 8), load32(p + 12)}}; return o; }

/* MEOW_MIX_REG (:181-189) on state lanes (r1..r5) with the four 16-byte inputs */
static void meow_mix_reg(m128* x, int r1, int r2, int r3, int r4, int r5, m128 i1, m128 i2, m128 i3, m128 i4)
{
    x[r1] = meow_aesdec(x[r1], x[r2]);
    x[r3] = meow_paddq(x[r3], i1);
    x[r2] = meow_pxor(x[r2], i2);
    x[r2] = meow_aesdec(x[r2], x[r4]);
    x[r5] = meow_paddq(x[r5], i3);
    x[r4] = meow_pxor(x[r4], i4);
}
/* MEOW_MIX (:191-192): inputs are the loads at +15, +0, +1, +16 of a 32-byte lane */
static void meow_mix(m128* x, int s, const uint8_t* p)
{
    /* lane roles rotate by one register per 32 bytes (:494-501): (0,4,6,1,2), (1,5,7,2,3), ... */
    meow_mix_reg(x, s & 7, (s + 4) & 7, (s + 6) & 7, (s + 1) & 7, (s + 2) & 7, meow_load(p + 15), meow_load(p), meow_load(p + 1), meow_load(p + 16));
}
/* MEOW_SHUFFLE (:194-200) */
static void meow_shuffle(m128* x, int r1, int r2, int r3, int r4, int r5, int r6)
{
    x[r1] = meow_aesdec(x[r1], x[r4]);
    x[r2] = meow_paddq(x[r2], x[r5]);
    x[r4] = meow_pxor(x[r4], x[r6]);
    x[r4] = meow_aesdec(x[r4], x[r2]);
    x[r5] = meow_paddq(x[r5], x[r6]);
    x[r2] = meow_pxor(x[r2], x[r3]);
}

static const uint8_t MEOW_SEED[128] = { /* :234-252, "an encoding of Pi" */
    0x32, 0x43, 0xF6, 0xA8, 0x88, 0x5A, 0x30, 0x8D, 0x31, 0x31, 0x98, 0xA2, 0xE0, 0x37, 0x07, 0x34, 0x4A, 0x40, 0x93, 0x82, 0x22, 0x99, 0xF3, 0x1D,
    0x00, 0x82, 0xEF, 0xA9, 0x8E, 0xC4, 0xE6, 0xC8, 0x94, 0x52, 0x82, 0x1E, 0x63, 0x8D, 0x01, 0x37, 0x7B, 0xE5, 0x46, 0x6C, 0xF3, 0x4E, 0x90, 0xC6,
    0xCC, 0x0A, 0xC2, 0x9B, 0x7C, 0x97, 0xC5, 0x0D, 0xD3, 0xF8, 0x4D, 0x5B, 0x5B, 0x54, 0x70, 0x91, 0x79, 0x21, 0x6D, 0x5D, 0x98, 0x97, 0x9F, 0xB1,
    0xBD, 0x13, 0x10, 0xBA, 0x69, 0x8D, 0xFB, 0x5A, 0xC2, 0xFF, 0xD7, 0x2D, 0xBD, 0x01, 0xAD, 0xFB, 0x7B, 0x8E, 0x1A, 0xFE, 0xD6, 0xA2, 0x67, 0xE9,
    0x6B, 0xA7, 0xC9, 0x04, 0x5F, 0x12, 0xC7, 0xF9, 0x92, 0x4A, 0x19, 0x94, 0x7B, 0x39, 0x16, 0xCF, 0x70, 0x80, 0x1F, 0x2E, 0x28, 0x58, 0xEF, 0xC1,
    0x66, 0x36, 0x92, 0x0D, 0x87, 0x15, 0x74, 0xE6};

uint64_t lto_meow_64(const void* data, uint64_t len)
{
    meow_init_tables();
    const uint8_t* p = (const uint8_t*)data;
    m128 x[8];
    for (int i = 0; i < 8; ++i) x[i] = meow_load(MEOW_SEED + 16 * i);
    /* full 256-byte blocks (MeowAbsorbBlocks :480-540) */
    uint64_t blocks = len >> 8;
    for (uint64_t b = 0; b < blocks; ++b, p += 256)
        for (int s = 0; s < 8; ++s) meow_mix(x, s, p + 32 * s);
    /* MeowEnd (:583-700): the residual (< 256 bytes) sits in a zero-padded buffer */
    uint8_t buf[256 + 32];
    memset(buf, 0, sizeof(buf));
    uint32_t rest = (uint32_t)(len & 255u);
    if (rest) memcpy(buf, p, rest);
    uint8_t tail[32]; /* xmm11 (bytes 0..15) then xmm9 (bytes 16..31) */
    memset(tail, 0, sizeof(tail));
    const uint8_t* last = buf + (len & 0xf0);
    uint32_t len8 = (uint32_t)(len & 0xf);
    memcpy(tail + 16, last, len8); /* masked load of the ragged 16 bytes */
    if (len & 0x10)
    {
        memcpy(tail, tail + 16, 16);      /* xmm11 = xmm9 */
        memcpy(tail + 16, last - 16, 16); /* xmm9 = the full 16 bytes before */
    }
    /* xmm8 = palignr(xmm9, xmm11, 15), xmm10 = palignr(xmm9, xmm11, 1): byte windows +15 and +1 of the pair xmm11:xmm9 */
    m128 xmm8 = meow_load(tail + 15), xmm9 = meow_load(tail + 16), xmm10 = meow_load(tail + 1), xmm11 = meow_load(tail);
    /* length lanes: xmm15 = (len, 0), xmm12 = palignr(0, xmm15, 15), xmm14 = palignr(0, xmm15, 1), xmm13 = 0 */
    uint8_t l15[32];
    memset(l15, 0, sizeof(l15));
    for (int i = 0; i < 8; ++i) l15[i] = (uint8_t)(len >> (8 * i));
    m128 xmm15 = meow_load(l15), xmm12 = meow_load(l15 + 15), xmm14 = meow_load(l15 + 1), xmm13 = {{0, 0, 0, 0}};
    meow_mix_reg(x, 0, 4, 6, 1, 2, xmm8, xmm9, xmm10, xmm11);
    meow_mix_reg(x, 1, 5, 7, 2, 3, xmm12, xmm13, xmm14, xmm15);
    uint32_t lanes = (uint32_t)((len >> 5) & 7);
    for (uint32_t s = 0; s < lanes; ++s) meow_mix(x, 2 + (int)s, buf + 32 * s);
    for (int s = 0; s < 12; ++s) meow_shuffle(x, s & 7, (s + 1) & 7, (s + 2) & 7, (s + 4) & 7, (s + 5) & 7, (s + 6) & 7);
    x[0] = meow_paddq(x[0], x[2]);
    x[1] = meow_paddq(x[1], x[3]);
    x[4] = meow_paddq(x[4], x[6]);
    x[5] = meow_paddq(x[5], x[7]);
    x[0] = meow_pxor(x[0], x[1]);
    x[4] = meow_pxor(x[4], x[5]);
    x[0] = meow_paddq(x[0], x[4]);
    return x[0].w[0] | ((uint64_t)x[0].w[1] << 32);
}

int lto_hash_buffer(uint32_t hash_type, const void* data, uint64_t len, uint64_t* out_hash)
{
    if (hash_type == LTO_HASH_BLAKE3) { *out_hash = lto_blake3_64(data, len); return 0; }
    if (hash_type == LTO_HASH_BLAKE2) { *out_hash = lto_blake2s_64(data, len); return 0; }
    if (hash_type == LTO_HASH_MEOW) { *out_hash = lto_meow_64(data, len); return 0; }
    return EINVAL;
}

int lto_hash_segments(uint32_t hash_type, const uint8_t* base, uint64_t count,
                      const uint64_t* offsets, const uint32_t* lens, uint64_t* out_hashes)
{
    for (uint64_t i = 0; i < count; ++i)
    {
        int err = lto_hash_buffer(hash_type, base + offsets[i], lens[i], &out_hashes[i]);
        if (err) return err;
    }
    return 0;
}

/* ================================================================ LZ4 block (lib/lz4/ext/lz4.c) */

uint64_t lto_lz4_bound(uint64_t size) { return size + size / 255 + 16; } /* lz4.h:215 */

static inline uint32_t lz4_hash(const uint8_t* p, int by_u16)
{
    if (by_u16) return (load32(p) * 2654435761u) >> 19;                 /* lz4.c:777-783, 13 bits */
    return (uint32_t)(((load64(p) << 24) * 889523592379ull) >> 52);     /* lz4.c:785-795, 12 bits */
}

static uint8_t* lz4_put_length(uint8_t* op, uint64_t len)
{
    for (; len >= 255; len -= 255) *op++ = 255;
    *op++ = (uint8_t)len;
    return op;
}

int lto_lz4_compress(const uint8_t* src, uint64_t size, uint8_t* dst, uint64_t cap, uint64_t* out_size)
{
    if (cap < lto_lz4_bound(size) || size > 0x7E000000u) return ENOMEM; /* only the notLimited path is restated (lz4.c:1388) */
    const uint64_t n = size;
    uint8_t* op = dst;
    uint64_t anchor = 0;
    if (n >= 13) /* LZ4_minLength, lz4.c:1001 */
    {
        const int by_u16 = n < 65547; /* LZ4_64Klimit, lz4.c:710,1389 */
        uint32_t* table = (uint32_t*)calloc(by_u16 ? 8192 : 4096, sizeof(uint32_t));
        if (!table) return ENOMEM;
        const uint64_t mflimit_plus_one = n - 11;
        const uint64_t matchlimit = n - 5;
        uint64_t ip = 0;
        table[lz4_hash(src, by_u16)] = 0; /* lz4.c:1004-1010 */
        ++ip;
        uint32_t forward_h = lz4_hash(src + ip, by_u16);
        for (;;)
        {
            uint64_t match;
            uint8_t* token;
            { /* lz4.c:1043-1100 */
                uint64_t forward_ip = ip;
                uint32_t step = 1;
                uint32_t search_nb = 64;
                for (;;)
                {
                    uint32_t h = forward_h;
                    uint64_t current = forward_ip;
                    uint64_t match_index = table[h];
                    ip = forward_ip;
                    forward_ip += step;
                    step = search_nb++ >> 6;
                    if (forward_ip > mflimit_plus_one) goto last_literals;
                    forward_h = lz4_hash(src + forward_ip, by_u16);
                    table[h] = (uint32_t)current;
                    if (!by_u16 && match_index + 65535 < current) continue;
                    if (load32(src + match_index) == load32(src + ip)) { match = match_index; break; }
                }
            }
            while (ip > anchor && match > 0 && src[ip - 1] == src[match - 1]) { --ip; --match; } /* lz4.c:1104-1109 */
            { /* lz4.c:1112-1137 */
                uint64_t lit = ip - anchor;
                token = op++;
                if (lit >= 15) { *token = 0xF0; op = lz4_put_length(op, lit - 15); }
                else *token = (uint8_t)(lit << 4);
                memmove(op, src + anchor, lit); /* memmove: the in-place test (tests/test_oracle.py) runs with dst in front of src in ONE buffer */
                op += lit;
            }
        next_match:
            { /* lz4.c:1157-1226 */
                uint64_t off = ip - match;
                *op++ = (uint8_t)off;
                *op++ = (uint8_t)(off >> 8);
                uint64_t a = ip + 4, b = match + 4;
                while (a < matchlimit && src[a] == src[b]) { ++a; ++b; }
                uint64_t code = a - (ip + 4);
                ip = a;
                if (code >= 15) { *token += 15; op = lz4_put_length(op, code - 15); }
                else *token += (uint8_t)code;
            }
            anchor = ip;
            if (ip >= mflimit_plus_one) break; /* lz4.c:1233 */
            table[lz4_hash(src + ip - 2, by_u16)] = (uint32_t)(ip - 2); /* lz4.c:1236-1243 */
            { /* lz4.c:1256-1294 */
                uint32_t h = lz4_hash(src + ip, by_u16);
                uint64_t match_index = table[h];
                table[h] = (uint32_t)ip;
                if ((by_u16 || match_index + 65535 >= ip) && load32(src + match_index) == load32(src + ip))
                {
                    token = op++;
                    *token = 0;
                    match = match_index;
                    goto next_match;
                }
            }
            forward_h = lz4_hash(src + ++ip, by_u16); /* lz4.c:1298 */
        }
    last_literals:
        free(table);
    }
    { /* lz4.c:1302-1329 */
        uint64_t last = n - anchor;
        if (last >= 15) { *op++ = 0xF0; op = lz4_put_length(op, last - 15); }
        else *op++ = (uint8_t)(last << 4);
        memmove(op, src + anchor, last);
        op += last;
    }
    *out_size = (uint64_t)(op - dst);
    return 0;
}

int lto_lz4_decompress(const uint8_t* src, uint64_t size, uint8_t* dst, uint64_t cap, uint64_t* out_size)
{
    /* LZ4 block format; used for round-trip property tests only */
    uint64_t ip = 0, op = 0;
    while (ip < size)
    {
        uint32_t token = src[ip++];
        uint64_t lit = token >> 4;
        if (lit == 15) { uint8_t b; do { if (ip >= size) return EBADF; b = src[ip++]; lit += b; } while (b == 255); }
        if (ip + lit > size || op + lit > cap) return EBADF;
        memcpy(dst + op, src + ip, lit);
        ip += lit;
        op += lit;
        if (ip >= size) break;
        if (ip + 2 > size) return EBADF;
        uint64_t off = (uint64_t)src[ip] | ((uint64_t)src[ip + 1] << 8);
        ip += 2;
        uint64_t len = token & 15;
        if (len == 15) { uint8_t b; do { if (ip >= size) return EBADF; b = src[ip++]; len += b; } while (b == 255); }
        len += 4;
        if (off == 0 || off > op || op + len > cap) return EBADF;
        for (uint64_t i = 0; i < len; ++i) dst[op + i] = dst[op - off + i];
        op += len;
    }
    *out_size = op;
    return 0;
}

/* ================================================================ open-addressing u64 -> u32 map */

struct u64map
{
    uint64_t* keys;
    uint32_t* vals; /* value + 1; 0 = empty */
    uint64_t mask;
};

static int u64map_init(struct u64map* m, uint64_t n)
{
    uint64_t cap = 16;
    while (cap < n * 2) cap <<= 1;
    m->keys = (uint64_t*)malloc(cap * sizeof(uint64_t));
    m->vals = (uint32_t*)calloc(cap, sizeof(uint32_t));
    m->mask = cap - 1;
    return (m->keys && m->vals) ? 0 : ENOMEM;
}
static void u64map_free(struct u64map* m) { free(m->keys); free(m->vals); }
/* returns pointer to the stored value+1 slot (existing or new, *is_new says which) */
static uint32_t* u64map_slot(struct u64map* m, uint64_t key, int* is_new)
{
    uint64_t i = (key * 0x9E3779B97F4A7C15ull) >> 20 & m->mask;
    for (;; i = (i + 1) & m->mask)
    {
        if (!m->vals[i]) { m->keys[i] = key; *is_new = 1; return &m->vals[i]; }
        if (m->keys[i] == key) { *is_new = 0; return &m->vals[i]; }
    }
}

/* ================================================================ CreateVersionIndex */

struct chunk_list
{
    uint64_t* hashes;
    uint32_t* sizes;
    uint32_t* tags;
    uint64_t count, cap;
};

static int chunk_list_push(struct chunk_list* l, uint64_t h, uint32_t size, uint32_t tag)
{
    if (l->count == l->cap)
    {
        l->cap = l->cap ? l->cap * 2 : 1024;
        l->hashes = (uint64_t*)realloc(l->hashes, l->cap * sizeof(uint64_t));
        l->sizes = (uint32_t*)realloc(l->sizes, l->cap * sizeof(uint32_t));
        l->tags = (uint32_t*)realloc(l->tags, l->cap * sizeof(uint32_t));
        if (!l->hashes || !l->sizes || !l->tags) return ENOMEM;
    }
    l->hashes[l->count] = h;
    l->sizes[l->count] = size;
    l->tags[l->count] = tag;
    ++l->count;
    return 0;
}

struct version_parts
{
    uint32_t asset_count;
    uint64_t* path_hashes;
    uint64_t* content_hashes;
    uint32_t* asset_chunk_counts;
    uint32_t* asset_chunk_starts;
    struct chunk_list all;       /* every chunk of every asset in order */
    uint32_t* asset_chunk_index; /* [all.count] -> unique index */
    struct chunk_list unique;
};

static void version_parts_free(struct version_parts* v)
{
    free(v->path_hashes); free(v->content_hashes); free(v->asset_chunk_counts); free(v->asset_chunk_starts);
    free(v->all.hashes); free(v->all.sizes); free(v->all.tags);
    free(v->unique.hashes); free(v->unique.sizes); free(v->unique.tags);
    free(v->asset_chunk_index);
}

static int build_version_parts(uint32_t count, const char** paths, const uint8_t** datas, const uint64_t* sizes,
                               const uint32_t* tags, uint32_t hash_type, uint32_t target, struct version_parts* v)
{
    memset(v, 0, sizeof(*v));
    v->asset_count = count;
    v->path_hashes = (uint64_t*)calloc(count ? count : 1, sizeof(uint64_t));
    v->content_hashes = (uint64_t*)calloc(count ? count : 1, sizeof(uint64_t));
    v->asset_chunk_counts = (uint32_t*)calloc(count ? count : 1, sizeof(uint32_t));
    v->asset_chunk_starts = (uint32_t*)calloc(count ? count : 1, sizeof(uint32_t));
    /* src/longtail.c:1985-1987, 2111-2113 with GetMinChunkSize() == 48 */
    const uint32_t mn = target / 8 < 48 ? 48 : target / 8;
    const uint32_t av = target / 2 < 48 ? 48 : target / 2;
    const uint32_t mx = target * 2 < 48 ? 48 : target * 2;
    const uint64_t part_size = (uint64_t)target * 1024; /* :2396 */
    uint64_t lens_cap = 1024;
    uint32_t* lens = (uint32_t*)malloc(lens_cap * sizeof(uint32_t));
    int err = 0;
    for (uint32_t a = 0; a < count && !err; ++a)
    {
        err = lto_hash_buffer(hash_type, paths[a], strlen(paths[a]), &v->path_hashes[a]); /* :1281-1297, :2008 */
        v->asset_chunk_starts[a] = (uint32_t)v->all.count;
        uint64_t parts = 1 + sizes[a] / part_size; /* :2402 — an exact multiple yields a trailing empty part */
        for (uint64_t p = 0; p < parts && !err; ++p)
        {
            uint64_t start = p * part_size;
            uint64_t n = sizes[a] - start > part_size ? part_size : sizes[a] - start;
            if (n == 0) continue; /* :2015-2019 */
            uint64_t need = n / mn + 2; /* min == max degenerates to fixed chunks of min bytes */
            if (need > lens_cap) { lens_cap = need; lens = (uint32_t*)realloc(lens, lens_cap * sizeof(uint32_t)); }
            uint64_t nchunks = 0;
            err = lto_hpcdc_chunk(datas[a] + start, n, mn, av, mx, lens, lens_cap, &nchunks); /* :2051-2296 */
            uint64_t off = start;
            for (uint64_t c = 0; c < nchunks && !err; ++c)
            {
                uint64_t h;
                err = lto_hash_buffer(hash_type, datas[a] + off, lens[c], &h);
                if (!err) err = chunk_list_push(&v->all, h, lens[c], tags ? tags[a] : 0);
                off += lens[c];
            }
        }
        v->asset_chunk_counts[a] = (uint32_t)(v->all.count - v->asset_chunk_starts[a]);
        /* :2518-2537 content hash = hash of the asset's chunk-hash array (may be empty) */
        if (!err)
            err = lto_hash_buffer(hash_type, v->all.hashes ? &v->all.hashes[v->asset_chunk_starts[a]] : (const void*)"",
                                  8ull * v->asset_chunk_counts[a], &v->content_hashes[a]);
    }
    free(lens);
    if (err) return err;
    /* :2952-2970 first-occurrence compaction */
    struct u64map map;
    if (u64map_init(&map, v->all.count)) return ENOMEM;
    v->asset_chunk_index = (uint32_t*)malloc((v->all.count ? v->all.count : 1) * sizeof(uint32_t));
    for (uint64_t c = 0; c < v->all.count && !err; ++c)
    {
        int is_new;
        uint32_t* slot = u64map_slot(&map, v->all.hashes[c], &is_new);
        if (is_new)
        {
            *slot = (uint32_t)v->unique.count + 1;
            err = chunk_list_push(&v->unique, v->all.hashes[c], v->all.sizes[c], v->all.tags[c]);
        }
        v->asset_chunk_index[c] = *slot - 1;
    }
    u64map_free(&map);
    return err;
}

static uint8_t* put_bytes(uint8_t* p, const void* src, size_t n) { if (n) memcpy(p, src, n); return p + n; }
static uint8_t* put_u32(uint8_t* p, uint32_t v) { return put_bytes(p, &v, 4); }

/* serialised layout src/longtail.c:2566-2584 / :2630-2704 */
static int serialise_version(const struct version_parts* v, uint32_t count, const char** paths, const uint64_t* sizes,
                             const uint16_t* perms, uint32_t hash_type, uint32_t target, void** out_buf, uint64_t* out_size)
{
    size_t name_bytes = 0;
    for (uint32_t a = 0; a < count; ++a) name_bytes += strlen(paths[a]) + 1;
    size_t total = 24 + (size_t)count * (8 + 8 + 8 + 4 + 4 + 4 + 2) + 4 * v->all.count + 16 * v->unique.count + name_bytes;
    uint8_t* buf = (uint8_t*)malloc(total);
    if (!buf) return ENOMEM;
    uint8_t* p = buf;
    p = put_u32(p, 2); /* Longtail_CurrentVersionIndexVersion, src/longtail.c:16-22 */
    p = put_u32(p, hash_type);
    p = put_u32(p, target);
    p = put_u32(p, count);
    p = put_u32(p, (uint32_t)v->unique.count);
    p = put_u32(p, (uint32_t)v->all.count);
    p = put_bytes(p, v->path_hashes, 8 * (size_t)count);
    p = put_bytes(p, v->content_hashes, 8 * (size_t)count);
    p = put_bytes(p, sizes, 8 * (size_t)count);
    p = put_bytes(p, v->asset_chunk_counts, 4 * (size_t)count);
    p = put_bytes(p, v->asset_chunk_starts, 4 * (size_t)count);
    p = put_bytes(p, v->asset_chunk_index, 4 * v->all.count);
    p = put_bytes(p, v->unique.hashes, 8 * v->unique.count);
    p = put_bytes(p, v->unique.sizes, 4 * v->unique.count);
    p = put_bytes(p, v->unique.tags, 4 * v->unique.count);
    uint32_t off = 0;
    for (uint32_t a = 0; a < count; ++a) { p = put_u32(p, off); off += (uint32_t)strlen(paths[a]) + 1; }
    for (uint32_t a = 0; a < count; ++a) { uint16_t pm = perms ? perms[a] : 0644; p = put_bytes(p, &pm, 2); }
    for (uint32_t a = 0; a < count; ++a) p = put_bytes(p, paths[a], strlen(paths[a]) + 1);
    *out_buf = buf;
    *out_size = (uint64_t)(p - buf);
    return (size_t)(p - buf) == total ? 0 : EFAULT;
}

int lto_create_version_index(uint32_t count, const char** paths, const uint8_t** datas, const uint64_t* sizes,
                             const uint16_t* perms, const uint32_t* tags, uint32_t hash_type,
                             uint32_t target_chunk_size, void** out_buf, uint64_t* out_size)
{
    struct version_parts v;
    int err = build_version_parts(count, paths, datas, sizes, tags, hash_type, target_chunk_size, &v);
    if (!err) err = serialise_version(&v, count, paths, sizes, perms, hash_type, target_chunk_size, out_buf, out_size);
    version_parts_free(&v);
    return err;
}

/* ================================================================ fresh-store upsync */

struct outbuf
{
    uint8_t* data;
    uint64_t size, cap;
};

static int outbuf_reserve(struct outbuf* o, uint64_t extra)
{
    if (o->size + extra > o->cap)
    {
        uint64_t cap = o->cap ? o->cap : 4096;
        while (cap < o->size + extra) cap *= 2;
        uint8_t* p = (uint8_t*)realloc(o->data, cap);
        if (!p) return ENOMEM;
        o->data = p;
        o->cap = cap;
    }
    return 0;
}

int lto_upsync(uint32_t count, const char** paths, const uint8_t** datas, const uint64_t* sizes,
               const uint16_t* perms, const uint32_t* tags, uint32_t hash_type,
               uint32_t target_chunk_size, uint32_t max_block_size, uint32_t max_chunks_per_block,
               void** out_buf, uint64_t* out_size)
{
    struct version_parts v;
    int err = build_version_parts(count, paths, datas, sizes, tags, hash_type, target_chunk_size, &v);
    if (err) { version_parts_free(&v); return err; }

    /* first occurrence of every unique chunk: (asset, offset) — CreateAssetPartLookup src/longtail.c:4429-4500 */
    uint64_t ucount = v.unique.count;
    uint32_t* first_asset = (uint32_t*)malloc((ucount ? ucount : 1) * sizeof(uint32_t));
    uint64_t* first_offset = (uint64_t*)malloc((ucount ? ucount : 1) * sizeof(uint64_t));
    uint8_t* seen = (uint8_t*)calloc(ucount ? ucount : 1, 1);
    for (uint32_t a = 0; a < count; ++a)
    {
        uint64_t off = 0;
        for (uint32_t c = 0; c < v.asset_chunk_counts[a]; ++c)
        {
            uint32_t u = v.asset_chunk_index[v.asset_chunk_starts[a] + c];
            if (!seen[u]) { seen[u] = 1; first_asset[u] = a; first_offset[u] = off; }
            off += v.unique.sizes[u];
        }
    }
    free(seen);

    struct outbuf out = {0, 0, 0};
    err = outbuf_reserve(&out, 4);
    out.size = 4;
    uint32_t block_count = 0;
    /* Longtail_CreateStoreIndex src/longtail.c:6796-6860: greedy packing in unique-chunk order
     * (DiffHashes against an empty store keeps version-index order, :6718-6740) */
    uint64_t i = 0;
    while (i < ucount && !err)
    {
        uint64_t first = i;
        uint32_t tag = v.unique.tags[i];
        uint32_t n = 1;
        uint32_t cur = v.unique.sizes[i];
        while (i + 1 < ucount)
        {
            if (v.unique.tags[i + 1] != tag) break;
            if (n == max_chunks_per_block) break;
            if (cur + v.unique.sizes[i + 1] > max_block_size + max_block_size / 10) break;
            cur += v.unique.sizes[i + 1];
            ++n;
            ++i;
        }
        ++i;
        /* Longtail_CreateBlockIndex :3712-3770 */
        uint64_t block_hash;
        err = lto_hash_buffer(hash_type, &v.unique.hashes[first], 8ull * n, &block_hash);
        if (err) break;
        /* WriteContentBlockJob :4640-4741 gathers the payload; CompressBlock compressblockstore:67-141 */
        uint8_t* payload = (uint8_t*)malloc(cur ? cur : 1);
        uint64_t w = 0;
        for (uint32_t c = 0; c < n; ++c)
        {
            uint64_t u = first + c;
            memcpy(payload + w, datas[first_asset[u]] + first_offset[u], v.unique.sizes[u]);
            w += v.unique.sizes[u];
        }
        uint64_t index_bytes = 8 + 4 + 4 + 4 + 12ull * n; /* Longtail_GetBlockIndexDataSize :3585-3597 */
        uint64_t data_bytes;
        uint8_t* data;
        if (tag == 0)
        {
            data = payload;
            data_bytes = cur;
        }
        else if (tag == LTO_COMPRESSION_LZ4)
        {
            uint64_t bound = lto_lz4_bound(cur);
            data = (uint8_t*)malloc(8 + bound);
            uint64_t csize = 0;
            err = lto_lz4_compress(payload, cur, data + 8, bound, &csize);
            uint32_t hdr[2] = {cur, (uint32_t)csize}; /* compressblockstore:135-137 */
            memcpy(data, hdr, 8);
            data_bytes = 8 + csize;
            free(payload);
        }
        else
        {
            free(payload);
            err = ENOTSUP;
            break;
        }
        if (!err) err = outbuf_reserve(&out, 16 + index_bytes + data_bytes);
        if (!err)
        {
            uint8_t* p = out.data + out.size;
            uint64_t total = index_bytes + data_bytes;
            p = put_bytes(p, &block_hash, 8);
            p = put_bytes(p, &total, 8);
            /* Longtail_WriteStoredBlockToBuffer :4111-4150 */
            p = put_bytes(p, &block_hash, 8);
            p = put_u32(p, hash_type);
            p = put_u32(p, n);
            p = put_u32(p, tag);
            p = put_bytes(p, &v.unique.hashes[first], 8ull * n);
            p = put_bytes(p, &v.unique.sizes[first], 4ull * n);
            p = put_bytes(p, data, data_bytes);
            out.size = (uint64_t)(p - out.data);
            ++block_count;
        }
        free(data);
    }
    free(first_asset);
    free(first_offset);
    if (!err)
    {
        memcpy(out.data, &block_count, 4);
        void* vbuf = 0;
        uint64_t vsize = 0;
        err = serialise_version(&v, count, paths, sizes, perms, hash_type, target_chunk_size, &vbuf, &vsize);
        if (!err) err = outbuf_reserve(&out, 8 + vsize);
        if (!err)
        {
            memcpy(out.data + out.size, &vsize, 8);
            memcpy(out.data + out.size + 8, vbuf, vsize);
            out.size += 8 + vsize;
        }
        free(vbuf);
    }
    version_parts_free(&v);
    if (err) { free(out.data); return err; }
    *out_buf = out.data;
    *out_size = out.size;
    return 0;
}
