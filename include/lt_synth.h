/* lt_synth.h — deterministic synthetic asset bytes, identical on host (C) and device (CUDA).
 *
 * Counter-based: the 16 bytes at (stream key, block index) are a pure function of those two
 * numbers, so any byte range of any asset can be produced anywhere with no state.  Used by
 * bench.py / tests to build the BASELINE.json workloads in HBM and, for the
 * CPU baseline, in host RAM — never by the indexing path itself.
 *
 * Asset model:
 *   - an asset is a sequence of 1 MiB segments;
 *   - a segment is either FRESH (keyed by the asset) or SHARED (drawn from a pool of
 *     `pool_segments` segments keyed by the pool), chosen per segment with probability
 *     `shared_permille`/1000  -> byte redundancy;
 *   - each segment has an entropy class: 0 = uniform random bytes, 1 = 4-bit entropy
 *     (byte & 0x0f), 2 = text-like (order-0 skewed letter distribution); `class_mode` selects
 *     "all random" (0) or "one third each, chosen per segment" (1).
 *
 * PAK-like model (`class_mode` 2): the asset is a concatenation of MEMBERS of
 * 64 KiB .. 64 MiB (log-uniform, multiples of 16 bytes), each followed by zero padding up to the next 2 KiB
 * boundary; a member has one entropy class (one third each).  So that any byte range can be produced without
 * walking the file from its start, members are laid out per 64 MiB SLOT: the member list of slot s is a pure
 * function of (seed, asset, s) and the slot's last member is cut at the slot end.
 */
#ifndef LT_SYNTH_H
#define LT_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define LT_SYNTH_FN __host__ __device__ static inline
#else
#define LT_SYNTH_FN static inline
#endif

#define LT_SYNTH_SEGMENT_BYTES (1u << 20)

struct lt_synth_spec
{
    uint64_t seed;
    uint32_t shared_permille; /* 0..1000: chance that a 1 MiB segment comes from the shared pool */
    uint32_t pool_segments;   /* size of the shared pool in segments (>=1 when shared_permille>0) */
    uint32_t class_mode;      /* 0: all uniform random; 1: random / 4-bit / text-like by segment; 2: PAK-like; 16 + c: all of class c */
    uint32_t reserved;
};

LT_SYNTH_FN uint64_t lt_synth_mix(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* 64 skewed "letters": index = 6 random bits folded through a table that repeats frequent
 * symbols, giving roughly 4.2 bits/byte like order-0 English text */
LT_SYNTH_FN uint8_t lt_synth_text_byte(uint8_t r)
{
    const char* alphabet = "eeeeeeee tttttt aaaaa ooooo iiii nnnn ssss hhh rrr dd ll cu\nmwfgyp";
    return (uint8_t)alphabet[r & 63u];
}

/* key and class of the segment holding byte `offset` of asset `asset_id` */
LT_SYNTH_FN void lt_synth_segment(const struct lt_synth_spec* s, uint64_t asset_id, uint64_t offset,
                                  uint64_t* out_key, uint64_t* out_base, uint32_t* out_class)
{
    uint64_t seg = offset / LT_SYNTH_SEGMENT_BYTES;
    uint64_t h = lt_synth_mix(s->seed * 0x9E3779B97F4A7C15ull + asset_id * 0xD1B54A32D192ED03ull + seg);
    uint64_t key;
    uint64_t base;
    if (s->shared_permille && (h % 1000u) < s->shared_permille)
    {
        uint64_t p = (h >> 20) % (s->pool_segments ? s->pool_segments : 1u);
        key = lt_synth_mix(s->seed ^ 0x5851F42D4C957F2Dull);
        base = p * (uint64_t)LT_SYNTH_SEGMENT_BYTES;
        h = lt_synth_mix(key + p);
    }
    else
    {
        key = lt_synth_mix(s->seed + (asset_id + 1) * 0xA0761D6478BD642Full);
        base = seg * (uint64_t)LT_SYNTH_SEGMENT_BYTES;
    }
    *out_key = key;
    *out_base = base;
    *out_class = s->class_mode >= 16u ? (s->class_mode - 16u) % 3u : s->class_mode ? (uint32_t)((h >> 40) % 3u) : 0u; /* 16 + c: every segment of class c */
}

/* the 16 bytes of 16-byte block `block` (= byte offset / 16 inside the keyed stream) */
LT_SYNTH_FN void lt_synth_block16(uint64_t key, uint64_t block, uint32_t cls, uint8_t out[16])
{
    uint64_t x = key + (block + 1) * 0x9E3779B97F4A7C15ull;
    uint64_t a = lt_synth_mix(x);
    uint64_t b = lt_synth_mix(x ^ 0xD6E8FEB86659FD93ull);
    for (int i = 0; i < 8; ++i)
    {
        out[i] = (uint8_t)(a >> (8 * i));
        out[8 + i] = (uint8_t)(b >> (8 * i));
    }
    if (cls == 1)
    {
        for (int i = 0; i < 16; ++i) out[i] &= 0x0f;
    }
    else if (cls == 2)
    {
        for (int i = 0; i < 16; ++i) out[i] = lt_synth_text_byte(out[i]);
    }
}

#define LT_SYNTH_PAK_SLOT_BYTES (64ull << 20)
#define LT_SYNTH_PAK_ALIGN 2048u

/* size of member m of slot `slot`: 65536 * 2^(10 u), u uniform in [0, 1) with 16 fractional bits, rounded to 16 bytes */
LT_SYNTH_FN uint64_t lt_synth_pak_member_size(uint64_t slot_key, uint32_t m)
{
    uint64_t h = lt_synth_mix(slot_key + (uint64_t)(m + 1) * 0xD6E8FEB86659FD93ull);
    uint32_t e = (uint32_t)(h >> 48) & 0xffffu;  /* exponent in 1/65536 units of the 10 octaves */
    uint32_t oct = (e * 10u) >> 16;              /* whole octaves 0..9 */
    uint32_t frac = (e * 10u) & 0xffffu;         /* position inside the octave: size grows linearly from 2^oct to 2^(oct+1) */
    uint64_t size = ((uint64_t)65536 << oct) + ((((uint64_t)65536 << oct) * frac) >> 16);
    return (size + 15u) & ~(uint64_t)15u;
}

/* where byte `offset` of a PAK-like asset comes from: *out_pad = 1 for padding (zero bytes), else the keyed stream position */
LT_SYNTH_FN void lt_synth_pak_locate(const struct lt_synth_spec* s, uint64_t asset_id, uint64_t offset,
                                     uint64_t* out_key, uint64_t* out_rel, uint32_t* out_class, uint32_t* out_pad)
{
    uint64_t slot = offset / LT_SYNTH_PAK_SLOT_BYTES;
    uint64_t in_slot = offset % LT_SYNTH_PAK_SLOT_BYTES;
    uint64_t slot_key = lt_synth_mix(s->seed * 0x9E3779B97F4A7C15ull + asset_id * 0xD1B54A32D192ED03ull + slot * 0xA0761D6478BD642Full + 0x51ull);
    uint64_t pos = 0;
    for (uint32_t m = 0;; ++m)
    {
        uint64_t size = lt_synth_pak_member_size(slot_key, m);
        uint64_t end = pos + size;
        if (in_slot < end || end >= LT_SYNTH_PAK_SLOT_BYTES)
        {
            uint64_t h = lt_synth_mix(slot_key ^ ((uint64_t)(m + 1) * 0x94D049BB133111EBull));
            *out_key = h;
            *out_rel = in_slot - pos;
            *out_class = (uint32_t)((h >> 40) % 3u);
            *out_pad = 0;
            return;
        }
        uint64_t padded = (end + (LT_SYNTH_PAK_ALIGN - 1u)) & ~(uint64_t)(LT_SYNTH_PAK_ALIGN - 1u);
        if (in_slot < padded)
        {
            *out_key = 0;
            *out_rel = 0;
            *out_class = 0;
            *out_pad = 1;
            return;
        }
        pos = padded;
    }
}

/* the 16 bytes at `offset` (a multiple of 16) of asset `asset_id` under any class_mode */
LT_SYNTH_FN void lt_synth_asset_block16(const struct lt_synth_spec* s, uint64_t asset_id, uint64_t offset, uint8_t out[16])
{
    if (s->class_mode == 2)
    {
        uint64_t key, rel;
        uint32_t cls, pad;
        lt_synth_pak_locate(s, asset_id, offset, &key, &rel, &cls, &pad);
        if (pad)
        {
            for (int i = 0; i < 16; ++i) out[i] = 0;
            return;
        }
        lt_synth_block16(key, rel / 16, cls, out);
        return;
    }
    uint64_t key, base;
    uint32_t cls;
    lt_synth_segment(s, asset_id, offset, &key, &base, &cls);
    lt_synth_block16(key, (base + offset % LT_SYNTH_SEGMENT_BYTES) / 16, cls, out);
}

/* Fill dst[0..len) with asset bytes [offset, offset+len).  offset must be a multiple of 16;
 * len may be ragged. */
LT_SYNTH_FN void lt_synth_fill(const struct lt_synth_spec* s, uint64_t asset_id, uint64_t offset,
                               uint8_t* dst, uint64_t len)
{
    if (s->class_mode == 2)
    {
        for (uint64_t i = 0; i < len; i += 16)
        {
            uint8_t tmp[16];
            lt_synth_asset_block16(s, asset_id, offset + i, tmp);
            uint64_t m = len - i < 16 ? len - i : 16;
            for (uint64_t j = 0; j < m; ++j) dst[i + j] = tmp[j];
        }
        return;
    }
    uint64_t done = 0;
    while (done < len)
    {
        uint64_t key, base;
        uint32_t cls;
        uint64_t off = offset + done;
        lt_synth_segment(s, asset_id, off, &key, &base, &cls);
        uint64_t seg_end = (off / LT_SYNTH_SEGMENT_BYTES + 1) * (uint64_t)LT_SYNTH_SEGMENT_BYTES;
        uint64_t n = seg_end - off;
        if (n > len - done) n = len - done;
        uint64_t in_seg = off % LT_SYNTH_SEGMENT_BYTES;
        for (uint64_t i = 0; i < n; i += 16)
        {
            uint8_t tmp[16];
            lt_synth_block16(key, (base + in_seg + i) / 16, cls, tmp);
            uint64_t m = n - i < 16 ? n - i : 16;
            for (uint64_t j = 0; j < m; ++j) dst[done + i + j] = tmp[j];
        }
        done += n;
    }
}

#endif /* LT_SYNTH_H */
