// zstd.cu — bit-exact ZStd "level 3" frame encoder ('ztd1' / 'ztd2' of lib/zstd/longtail_zstd.c:43-62,107-140) on sm_100a.
//
// One stored block (<= ~9.2 MiB) is one ZStd frame, compressed by ZSTD_compressCCtx(level 3) with a fresh context
// (SURVEY.md A.5): double-fast matcher (lib/zstd/ext/compress/zstd_double_fast.c:105-311) over a frame-wide pair of hash
// tables, 128 KiB blocks, Huffman literals (huf_compress.c), FSE sequences (zstd_compress_sequences.c, fse_compress.c).
//
// Dependencies inside a frame are strictly serial and cross the entropy stage: block k+1's matcher starts from the hash
// tables block k left behind, from repcodes that are only "confirmed" when block k was actually emitted compressed
// (zstd_compress.c:4372-4375), and its literals may reuse block k's Huffman table.  Frames are independent.  Mapping:
//   * one WARP per frame, taken from a global queue by a persistent grid (frames differ 100x in cost: incompressible data
//     accelerates the matcher's stride, compressible data does not);
//   * per-warp state (two hash tables = 768 KiB, sequence / literal / code staging, entropy tables) lives in a
//     workspace slab in HBM — it is far too large for shared memory and is touched at random, so it is served by L2;
//   * table reset, block copy-out and the all-bytes-equal test are lane-parallel; the parse and the bit-serial entropy
//     coders run on lane 0 in this first version (the parse is latency bound: ~2 dependent L2 round trips per probe).
//
// Every decision point that fixes output bytes follows the reference exactly; the file:line of each is cited where it is
// restated.  tests/test_gpu_zstd.py compares frames byte for byte with the CPU oracle / the unmodified reference.
#include "lt_device.cuh"
#include "lt_kernels.h"

#include <stdlib.h>

#include <stddef.h>

namespace ltb {

namespace {

constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t ZS_BLOCK_MAX = 128u << 10;      // ZSTD_BLOCKSIZE_MAX
constexpr uint32_t ZS_MAX_SEQ = ZS_BLOCK_MAX / 4 + 2; // every sequence holds a match of >= 4 bytes at level 3
constexpr uint32_t ZS_ERR = 0xffffffffu;

struct FseTable
{
    uint32_t table_log, max_symbol;
    uint16_t next_state[512]; // sorted by symbol; <= 9 bits for sequences, 6 for Huffman weights
    uint32_t delta_nb_bits[64];
    int32_t delta_find_state[64];
};
struct HufTable
{
    uint32_t table_log, max_symbol; // CTable header, huf_compress.c:219-241
    uint8_t nb_bits[256];
    uint16_t code[256];
    int32_t repeat; // HUF_repeat: 0 none, 1 check, 2 valid
};
struct HufNode
{
    uint32_t count;
    uint16_t parent;
    uint8_t byte, nb_bits;
};

struct alignas(16) ZstdWorker
{
    // Table entries carry the bytes found at their position next to the index ({index, 8 bytes} / {index, 4 bytes}), so the match test
    // of a candidate (MEM_read64(matchl0) == MEM_read64(ip), zstd_double_fast.c:186; MEM_read32(matchs0) == MEM_read32(ip), :207) needs
    // no second dependent load from the source: table -> compare instead of table -> source -> compare.  The tables are this encoder's
    // private state; what they answer is unchanged.
    uint4 hash_long[1u << 17];  // {index, bytes 0..3, bytes 4..7, 0}
    uint2 hash_small[1u << 16]; // {index, bytes 0..3}
    uint32_t seq_lit[ZS_MAX_SEQ], seq_len[ZS_MAX_SEQ], seq_off[ZS_MAX_SEQ];
    uint8_t lits[ZS_BLOCK_MAX + 64];
    uint8_t ll_code[ZS_MAX_SEQ + 6], of_code[ZS_MAX_SEQ + 6], ml_code[ZS_MAX_SEQ + 6];
    uint16_t enc_of[ZS_MAX_SEQ + 6], enc_ml[ZS_MAX_SEQ + 6], enc_ll[ZS_MAX_SEQ + 6]; // per sequence: FSE state bits | count << 12
    uint8_t scratch[3 * ZS_BLOCK_MAX + 4096]; // a block that expands is discarded afterwards, so it needs room
    HufTable huf_prev, huf_next, huf_fresh;
    FseTable ct_ll, ct_of, ct_ml, ct_w;
    HufNode nodes[516];
    uint32_t count[256];
    uint16_t cumul[260];
    uint16_t rank_base[192], rank_curr[192];
    uint8_t spread[512];
    uint8_t weights[260];
    int16_t norm[64];
    int32_t sort_stack[64 * 3];
    uint32_t rep[3];
    uint32_t dict_limit;
    uint32_t window_log, chain_log, hash_log, min_match;
    unsigned long long t_phase[4]; // diagnostic: cycles spent in the matcher, the literal stage, the sequence stage, copy-out
};

// ------------------------------------------------------------------ unaligned little-endian loads from a 4-byte aligned base
__device__ __forceinline__ uint32_t rd32(const uint8_t* __restrict__ s, uint32_t pos)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(s + (pos & ~3u));
    return __funnelshift_r(w[0], w[1], (pos & 3u) * 8u);
}
__device__ __forceinline__ uint64_t rd64(const uint8_t* __restrict__ s, uint32_t pos)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(s + (pos & ~3u));
    const uint32_t sh = (pos & 3u) * 8u;
    const uint32_t a = w[0], b = w[1], c = w[2];
    return (uint64_t)__funnelshift_r(a, b, sh) | ((uint64_t)__funnelshift_r(b, c, sh) << 32);
}
__device__ __forceinline__ uint32_t hibit(uint32_t v) { return 31u - (uint32_t)__clz(v); }

// warp-wide byte copy, any alignment: aligned 4-byte stores assembled from the source with a funnel shift.  The source must
// be readable up to 4 bytes past its end (every source on this path is).  CG = read through L2 (data written by atomics).
template <bool CG>
__device__ void copy_w(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n, uint32_t lane)
{
    uint32_t done = 0;
    if (n >= 64)
    {
        const uint32_t head = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u;
        if (lane < head) dst[lane] = CG ? __ldcg(src + lane) : src[lane];
        const uint32_t words = (n - head) >> 2;
        const uint8_t* s2 = src + head;
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(s2) & 3u) * 8u;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(s2 - (reinterpret_cast<uintptr_t>(s2) & 3u));
        uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + head);
        for (uint32_t w = lane; w < words; w += 32)
        {
            const uint32_t a = CG ? __ldcg(sw + w) : sw[w], b = CG ? __ldcg(sw + w + 1) : sw[w + 1];
            d4[w] = __funnelshift_r(a, b, sh);
        }
        done = head + words * 4u;
    }
    for (uint32_t i = done + lane; i < n; i += 32) dst[i] = CG ? __ldcg(src + i) : src[i];
}

// ------------------------------------------------------------------ forward bit writer (common/bitstream.h:150-241)
struct BitW
{
    uint8_t* out;
    uint32_t pos;
    uint64_t acc;
    uint32_t nbits;
};
__device__ __forceinline__ void bw_init(BitW& w, uint8_t* out) { w.out = out; w.pos = 0; w.acc = 0; w.nbits = 0; }
__device__ __forceinline__ void bw_add(BitW& w, uint64_t value, uint32_t n)
{
    if (!n) return;
    value &= (1ull << n) - 1ull; // n <= 31 everywhere on this path
    w.acc |= value << w.nbits;
    w.nbits += n;
    while (w.nbits >= 8)
    {
        w.out[w.pos++] = (uint8_t)w.acc;
        w.acc >>= 8;
        w.nbits -= 8;
    }
}
__device__ __forceinline__ uint32_t bw_close(BitW& w)
{
    bw_add(w, 1, 1); // end mark
    if (w.nbits) { w.out[w.pos++] = (uint8_t)w.acc; w.nbits = 0; }
    return w.pos;
}

// ================================================================== FSE (lib/zstd/ext/compress/fse_compress.c)

__device__ uint32_t fse_min_table_log(uint32_t n, uint32_t max_symbol) // :343-351
{
    const uint32_t by_src = hibit(n) + 1, by_sym = hibit(max_symbol) + 2;
    return by_src < by_sym ? by_src : by_sym;
}
__device__ uint32_t fse_optimal_table_log(uint32_t max_log, uint32_t n, uint32_t max_symbol, uint32_t minus) // :357-369
{
    const uint32_t by_src = hibit(n - 1) - minus;
    uint32_t log = max_log;
    const uint32_t min_bits = fse_min_table_log(n, max_symbol);
    if (by_src < log) log = by_src;
    if (min_bits > log) log = min_bits;
    if (log < 5) log = 5;
    if (log > 12) log = 12;
    return log;
}

__device__ int fse_normalize_m2(int16_t* norm, uint32_t log, const uint32_t* count, uint64_t total, uint32_t max_symbol, int16_t low_prob) // :379-463
{
    const int16_t UNSET = -2;
    uint32_t distributed = 0, to_give;
    const uint32_t low_threshold = (uint32_t)(total >> log);
    uint32_t low_one = (uint32_t)((total * 3) >> (log + 1));
    for (uint32_t s = 0; s <= max_symbol; ++s)
    {
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= low_threshold) { norm[s] = low_prob; distributed++; total -= count[s]; continue; }
        if (count[s] <= low_one) { norm[s] = 1; distributed++; total -= count[s]; continue; }
        norm[s] = UNSET;
    }
    to_give = (1u << log) - distributed;
    if (to_give == 0) return 0;
    if ((total / to_give) > low_one)
    {
        low_one = (uint32_t)((total * 3) / (to_give * 2));
        for (uint32_t s = 0; s <= max_symbol; ++s)
            if (norm[s] == UNSET && count[s] <= low_one) { norm[s] = 1; distributed++; total -= count[s]; }
        to_give = (1u << log) - distributed;
    }
    if (distributed == max_symbol + 1)
    {
        uint32_t best = 0, best_count = 0;
        for (uint32_t s = 0; s <= max_symbol; ++s)
            if (count[s] > best_count) { best = s; best_count = count[s]; }
        norm[best] = (int16_t)(norm[best] + (int16_t)to_give);
        return 0;
    }
    if (total == 0)
    {
        for (uint32_t s = 0; to_give > 0; s = (s + 1) % (max_symbol + 1))
            if (norm[s] > 0) { to_give--; norm[s]++; }
        return 0;
    }
    {
        const uint64_t v_log = 62 - log;
        const uint64_t mid = (1ull << (v_log - 1)) - 1;
        const uint64_t r_step = (((uint64_t)1 << v_log) * to_give + mid) / (uint32_t)total;
        uint64_t run = mid;
        for (uint32_t s = 0; s <= max_symbol; ++s)
        {
            if (norm[s] != UNSET) continue;
            const uint64_t end = run + count[s] * r_step;
            const uint32_t w = (uint32_t)(end >> v_log) - (uint32_t)(run >> v_log);
            if (w < 1) return -1;
            norm[s] = (int16_t)w;
            run = end;
        }
    }
    return 0;
}

// :465-526; returns the table log, 0 for the single-symbol case, -1 on error
__device__ int fse_normalize(int16_t* norm, uint32_t log, const uint32_t* count, uint32_t total, uint32_t max_symbol, bool use_low_prob)
{
    const uint32_t round_up_threshold[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
    if (log < 5 || log > 12) return -1;
    if (log < fse_min_table_log(total, max_symbol)) return -1;
    const int16_t low_prob = use_low_prob ? -1 : 1;
    const uint64_t scale = 62 - log;
    const uint64_t step = ((uint64_t)1 << 62) / total;
    const uint64_t v_step = 1ull << (scale - 20);
    int left = 1 << log;
    uint32_t largest = 0;
    int16_t largest_p = 0;
    const uint32_t low_threshold = total >> log;
    for (uint32_t s = 0; s <= max_symbol; ++s)
    {
        if (count[s] == total) return 0;
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= low_threshold) { norm[s] = low_prob; left--; continue; }
        int16_t p = (int16_t)((count[s] * step) >> scale);
        if (p < 8)
        {
            const uint64_t rest_to_beat = v_step * round_up_threshold[p];
            p = (int16_t)(p + ((count[s] * step) - ((uint64_t)p << scale) > rest_to_beat));
        }
        if (p > largest_p) { largest_p = p; largest = s; }
        norm[s] = p;
        left -= p;
    }
    if (-left >= (norm[largest] >> 1))
    {
        if (fse_normalize_m2(norm, log, count, total, max_symbol, low_prob)) return -1;
    }
    else
        norm[largest] = (int16_t)(norm[largest] + (int16_t)left);
    return (int)log;
}

__device__ uint32_t fse_write_ncount(uint8_t* out, const int16_t* norm, uint32_t max_symbol, uint32_t log) // :238-330
{
    const int table_size = 1 << log;
    int remaining = table_size + 1, threshold = table_size, nb_bits = (int)log + 1;
    uint32_t bits = log - 5;
    int bit_count = 4;
    uint32_t symbol = 0;
    const uint32_t alphabet = max_symbol + 1;
    bool previous_is_0 = false;
    uint32_t pos = 0;
    while (symbol < alphabet && remaining > 1)
    {
        if (previous_is_0)
        {
            uint32_t start = symbol;
            while (symbol < alphabet && !norm[symbol]) symbol++;
            if (symbol == alphabet) break;
            while (symbol >= start + 24)
            {
                start += 24;
                bits += 0xFFFFu << bit_count;
                out[pos++] = (uint8_t)bits;
                out[pos++] = (uint8_t)(bits >> 8);
                bits >>= 16;
            }
            while (symbol >= start + 3) { start += 3; bits += 3u << bit_count; bit_count += 2; }
            bits += (symbol - start) << bit_count;
            bit_count += 2;
            if (bit_count > 16) { out[pos++] = (uint8_t)bits; out[pos++] = (uint8_t)(bits >> 8); bits >>= 16; bit_count -= 16; }
        }
        {
            int count = norm[symbol++];
            const int max = (2 * threshold - 1) - remaining;
            remaining -= count < 0 ? -count : count;
            count++;
            if (count >= threshold) count += max;
            bits += (uint32_t)count << bit_count;
            bit_count += nb_bits;
            bit_count -= (count < max);
            previous_is_0 = (count == 1);
            if (remaining < 1) return ZS_ERR;
            while (remaining < threshold) { nb_bits--; threshold >>= 1; }
        }
        if (bit_count > 16) { out[pos++] = (uint8_t)bits; out[pos++] = (uint8_t)(bits >> 8); bits >>= 16; bit_count -= 16; }
    }
    if (remaining != 1) return ZS_ERR;
    out[pos] = (uint8_t)bits;
    out[pos + 1] = (uint8_t)(bits >> 8);
    pos += (uint32_t)((bit_count + 7) / 8);
    return pos;
}

__device__ void fse_build_ctable(ZstdWorker* W, FseTable* ct, const int16_t* norm, uint32_t max_symbol, uint32_t log) // :68-209
{
    const uint32_t size = 1u << log, mask = size - 1;
    const uint32_t step = (size >> 1) + (size >> 3) + 3;
    uint16_t* cumul = W->cumul;
    uint8_t* spread = W->spread;
    uint32_t high = size - 1;
    ct->table_log = log;
    ct->max_symbol = max_symbol;
    cumul[0] = 0;
    for (uint32_t u = 1; u <= max_symbol + 1; ++u)
    {
        if (norm[u - 1] == -1) { cumul[u] = (uint16_t)(cumul[u - 1] + 1); spread[high--] = (uint8_t)(u - 1); }
        else cumul[u] = (uint16_t)(cumul[u - 1] + (uint16_t)norm[u - 1]);
    }
    cumul[max_symbol + 1] = (uint16_t)(size + 1);
    {
        uint32_t position = 0;
        for (uint32_t s = 0; s <= max_symbol; ++s)
            for (int i = 0; i < norm[s]; ++i)
            {
                spread[position] = (uint8_t)s;
                position = (position + step) & mask;
                while (position > high) position = (position + step) & mask;
            }
    }
    for (uint32_t u = 0; u < size; ++u) ct->next_state[cumul[spread[u]]++] = (uint16_t)(size + u);
    uint32_t total = 0;
    for (uint32_t s = 0; s <= max_symbol; ++s)
    {
        if (norm[s] == 0) { ct->delta_nb_bits[s] = ((log + 1) << 16) - (1u << log); ct->delta_find_state[s] = 0; }
        else if (norm[s] == -1 || norm[s] == 1)
        {
            ct->delta_nb_bits[s] = (log << 16) - (1u << log);
            ct->delta_find_state[s] = (int32_t)(total - 1);
            total++;
        }
        else
        {
            const uint32_t max_bits_out = log - hibit((uint32_t)norm[s] - 1);
            const uint32_t min_state_plus = (uint32_t)norm[s] << max_bits_out;
            ct->delta_nb_bits[s] = (max_bits_out << 16) - min_state_plus;
            ct->delta_find_state[s] = (int32_t)(total - (uint32_t)norm[s]);
            total += (uint32_t)norm[s];
        }
    }
}
__device__ void fse_build_ctable_rle(FseTable* ct, uint32_t symbol) // :532-552
{
    ct->table_log = 0;
    ct->max_symbol = symbol;
    ct->next_state[0] = ct->next_state[1] = 0;
    ct->delta_nb_bits[symbol] = 0;
    ct->delta_find_state[symbol] = 0;
}

// common/fse.h:437-476
struct FseState
{
    const FseTable* ct;
    uint32_t value;
};
__device__ __forceinline__ void fse_init_state(FseState& st, const FseTable* ct, uint32_t symbol)
{
    const uint32_t nb = (ct->delta_nb_bits[symbol] + (1u << 15)) >> 16;
    const uint32_t v = (nb << 16) - ct->delta_nb_bits[symbol];
    st.ct = ct;
    st.value = ct->next_state[(int32_t)(v >> nb) + ct->delta_find_state[symbol]];
}
__device__ __forceinline__ void fse_encode(BitW& w, FseState& st, uint32_t symbol)
{
    const uint32_t nb = (st.value + st.ct->delta_nb_bits[symbol]) >> 16;
    bw_add(w, st.value, nb);
    st.value = st.ct->next_state[(int32_t)(st.value >> nb) + st.ct->delta_find_state[symbol]];
}
__device__ __forceinline__ void fse_flush_state(BitW& w, const FseState& st) { bw_add(w, st.value, st.ct->table_log); }

// ================================================================== Huffman literals (lib/zstd/ext/compress/huf_compress.c)

__device__ uint32_t hist(uint32_t* count, uint32_t* max_symbol, const uint8_t* src, uint32_t n) // hist.c:29-56
{
    uint32_t m = *max_symbol, largest = 0;
    for (uint32_t s = 0; s <= m; ++s) count[s] = 0;
    if (!n) { *max_symbol = 0; return 0; }
    for (uint32_t i = 0; i < n; ++i) count[src[i]]++;
    while (!count[m]) m--;
    *max_symbol = m;
    for (uint32_t s = 0; s <= m; ++s) if (count[s] > largest) largest = count[s];
    return largest;
}

// the weights of the table description, FSE-compressed (:127-182); 0 / 1 = not worth it
__device__ uint32_t huf_compress_weights(ZstdWorker* W, uint8_t* dst, const uint8_t* weights, uint32_t n)
{
    uint32_t count[13], max_symbol = 12;
    int16_t* norm = W->norm;
    if (n <= 1) return 0;
    {
        const uint32_t most = hist(count, &max_symbol, weights, n);
        if (most == n) return 1;
        if (most == 1) return 0;
    }
    const uint32_t log = fse_optimal_table_log(6, n, max_symbol, 2);
    if (fse_normalize(norm, log, count, n, max_symbol, false) < 0) return ZS_ERR;
    uint32_t pos = fse_write_ncount(dst, norm, max_symbol, log);
    if (pos == ZS_ERR) return ZS_ERR;
    fse_build_ctable(W, &W->ct_w, norm, max_symbol, log);
    if (n <= 2) return 0;
    // FSE_compress_usingCTable_generic (fse_compress.c:560-610): two interleaved states, input read backwards
    BitW w;
    FseState s1, s2;
    const uint8_t* ip = weights + n;
    bw_init(w, dst + pos);
    if (n & 1)
    {
        fse_init_state(s1, &W->ct_w, *--ip);
        fse_init_state(s2, &W->ct_w, *--ip);
        fse_encode(w, s1, *--ip);
    }
    else
    {
        fse_init_state(s2, &W->ct_w, *--ip);
        fse_init_state(s1, &W->ct_w, *--ip);
    }
    while (ip > weights)
    {
        fse_encode(w, s2, *--ip);
        fse_encode(w, s1, *--ip);
    }
    fse_flush_state(w, s2);
    fse_flush_state(w, s1);
    return pos + bw_close(w);
}

__device__ uint32_t huf_write_table(ZstdWorker* W, uint8_t* dst, const HufTable* t) // :248-290
{
    uint8_t bits_to_weight[13];
    uint8_t* weights = W->weights;
    const uint32_t max_symbol = t->max_symbol;
    bits_to_weight[0] = 0;
    for (uint32_t n = 1; n < t->table_log + 1; ++n) bits_to_weight[n] = (uint8_t)(t->table_log + 1 - n);
    for (uint32_t n = 0; n < max_symbol; ++n) weights[n] = bits_to_weight[t->nb_bits[n]];
    {
        const uint32_t h = huf_compress_weights(W, dst + 1, weights, max_symbol);
        if (h == ZS_ERR) return ZS_ERR;
        if (h > 1 && h < max_symbol / 2) { dst[0] = (uint8_t)h; return h + 1; }
    }
    if (max_symbol > 128) return ZS_ERR;
    dst[0] = (uint8_t)(128 + (max_symbol - 1));
    weights[max_symbol] = 0;
    for (uint32_t n = 0; n < max_symbol; n += 2) dst[n / 2 + 1] = (uint8_t)((weights[n] << 4) + weights[n + 1]);
    return (max_symbol + 1) / 2 + 1;
}

// :530-665: bucket sort by count, descending; counts >= 166 share log2 buckets sorted by an unstable quicksort whose exact
// element moves decide the order of equal counts, hence are reproduced move for move (recursion unrolled onto a stack:
// the sub-ranges are disjoint, so the order in which they are processed does not matter)
__device__ __forceinline__ uint32_t huf_bucket(uint32_t count) { return count < 166 ? count : hibit(count) + 158; }
__device__ void huf_insertion_sort(HufNode* a, int low, int high)
{
    const int size = high - low + 1;
    a += low;
    for (int i = 1; i < size; ++i)
    {
        const HufNode key = a[i];
        int j = i - 1;
        while (j >= 0 && a[j].count < key.count) { a[j + 1] = a[j]; j--; }
        a[j + 1] = key;
    }
}
__device__ int huf_partition(HufNode* a, int low, int high)
{
    const uint32_t pivot = a[high].count;
    int i = low - 1;
    for (int j = low; j < high; ++j)
        if (a[j].count > pivot)
        {
            i++;
            const HufNode t = a[i]; a[i] = a[j]; a[j] = t;
        }
    const HufNode t = a[i + 1]; a[i + 1] = a[high]; a[high] = t;
    return i + 1;
}
__device__ void huf_quick_sort(ZstdWorker* W, HufNode* a, int low0, int high0)
{
    int32_t* stack = W->sort_stack;
    int sp = 0;
    stack[sp++] = low0; stack[sp++] = high0;
    while (sp)
    {
        int high = stack[--sp], low = stack[--sp];
        if (high - low < 8) { huf_insertion_sort(a, low, high); continue; } // the threshold is only tested on entry (:590-594)
        while (low < high)
        {
            const int idx = huf_partition(a, low, high);
            if (idx - low < high - idx) { stack[sp++] = low; stack[sp++] = idx - 1; low = idx + 1; }
            else { stack[sp++] = idx + 1; stack[sp++] = high; high = idx - 1; }
        }
    }
}
__device__ void huf_sort(ZstdWorker* W, HufNode* node, const uint32_t* count, uint32_t max_symbol)
{
    uint16_t* base = W->rank_base;
    uint16_t* curr = W->rank_curr;
    for (int i = 0; i < 192; ++i) { base[i] = 0; curr[i] = 0; }
    for (uint32_t n = 0; n <= max_symbol; ++n) base[huf_bucket(count[n])]++;
    for (uint32_t n = 191; n > 0; --n) { base[n - 1] = (uint16_t)(base[n - 1] + base[n]); curr[n - 1] = base[n - 1]; }
    for (uint32_t n = 0; n <= max_symbol; ++n)
    {
        const uint32_t r = huf_bucket(count[n]) + 1;
        const uint32_t pos = curr[r]++;
        node[pos].count = count[n];
        node[pos].byte = (uint8_t)n;
    }
    for (uint32_t n = 166; n < 191; ++n)
    {
        const int size = (int)curr[n] - (int)base[n];
        if (size > 1) huf_quick_sort(W, node + base[n], 0, size - 1);
    }
}

__device__ uint32_t huf_set_max_height(HufNode* node, uint32_t last_non_null, uint32_t target) // :376-505
{
    const uint32_t largest = node[last_non_null].nb_bits;
    if (largest <= target) return largest;
    int total_cost = 0;
    const uint32_t base_cost = 1u << (largest - target);
    int n = (int)last_non_null;
    while (node[n].nb_bits > target)
    {
        total_cost += (int)(base_cost - (1u << (largest - node[n].nb_bits)));
        node[n].nb_bits = (uint8_t)target;
        n--;
    }
    while (node[n].nb_bits == target) --n;
    total_cost >>= (largest - target);
    const uint32_t NONE = 0xF0F0F0F0u;
    uint32_t rank_last[14];
    for (int i = 0; i < 14; ++i) rank_last[i] = NONE;
    {
        uint32_t current = target;
        for (int pos = n; pos >= 0; pos--)
        {
            if (node[pos].nb_bits >= current) continue;
            current = node[pos].nb_bits;
            rank_last[target - current] = (uint32_t)pos;
        }
    }
    while (total_cost > 0)
    {
        uint32_t dec = hibit((uint32_t)total_cost) + 1;
        for (; dec > 1; dec--)
        {
            const uint32_t high_pos = rank_last[dec], low_pos = rank_last[dec - 1];
            if (high_pos == NONE) continue;
            if (low_pos == NONE) break;
            if (node[high_pos].count <= 2 * node[low_pos].count) break;
        }
        while (dec <= 12 && rank_last[dec] == NONE) dec++;
        total_cost -= 1 << (dec - 1);
        node[rank_last[dec]].nb_bits++;
        if (rank_last[dec - 1] == NONE) rank_last[dec - 1] = rank_last[dec];
        if (rank_last[dec] == 0) rank_last[dec] = NONE;
        else
        {
            rank_last[dec]--;
            if (node[rank_last[dec]].nb_bits != target - dec) rank_last[dec] = NONE;
        }
    }
    while (total_cost < 0)
    {
        if (rank_last[1] == NONE)
        {
            while (node[n].nb_bits == target) n--;
            node[n + 1].nb_bits--;
            rank_last[1] = (uint32_t)(n + 1);
            total_cost++;
            continue;
        }
        node[rank_last[1] + 1].nb_bits--;
        rank_last[1]++;
        total_cost++;
    }
    return target;
}

// :681-800: sorted leaves -> tree -> depth limit -> canonical codes
__device__ uint32_t huf_build_table(ZstdWorker* W, HufTable* t, const uint32_t* count, uint32_t max_symbol, uint32_t max_bits)
{
    HufNode* const node = W->nodes + 1; // node[-1] is the sentinel
    const int START = 256;
    for (int i = 0; i < 514; ++i) { W->nodes[i].count = 0; W->nodes[i].parent = 0; W->nodes[i].byte = 0; W->nodes[i].nb_bits = 0; }
    huf_sort(W, node, count, max_symbol);
    int non_null = (int)max_symbol;
    while (node[non_null].count == 0) non_null--;
    {
        int low_s = non_null, node_nb = START;
        const int node_root = node_nb + low_s - 1;
        int low_n = node_nb;
        node[node_nb].count = node[low_s].count + node[low_s - 1].count;
        node[low_s].parent = node[low_s - 1].parent = (uint16_t)node_nb;
        node_nb++;
        low_s -= 2;
        for (int n = node_nb; n <= node_root; ++n) node[n].count = 1u << 30;
        node[-1].count = 1u << 31;
        while (node_nb <= node_root)
        {
            const int n1 = (node[low_s].count < node[low_n].count) ? low_s-- : low_n++;
            const int n2 = (node[low_s].count < node[low_n].count) ? low_s-- : low_n++;
            node[node_nb].count = node[n1].count + node[n2].count;
            node[n1].parent = node[n2].parent = (uint16_t)node_nb;
            node_nb++;
        }
        node[node_root].nb_bits = 0;
        for (int n = node_root - 1; n >= START; --n) node[n].nb_bits = (uint8_t)(node[node[n].parent].nb_bits + 1);
        for (int n = 0; n <= non_null; ++n) node[n].nb_bits = (uint8_t)(node[node[n].parent].nb_bits + 1);
    }
    max_bits = huf_set_max_height(node, (uint32_t)non_null, max_bits);
    uint16_t per_rank[13], val_per_rank[13];
    for (int i = 0; i < 13; ++i) { per_rank[i] = 0; val_per_rank[i] = 0; }
    for (int n = 0; n <= non_null; ++n) per_rank[node[n].nb_bits]++;
    uint16_t min = 0;
    for (int n = (int)max_bits; n > 0; --n) { val_per_rank[n] = min; min = (uint16_t)(min + per_rank[n]); min >>= 1; }
    for (int i = 0; i < 256; ++i) { t->nb_bits[i] = 0; t->code[i] = 0; }
    for (uint32_t n = 0; n <= max_symbol; ++n) t->nb_bits[node[n].byte] = node[n].nb_bits;
    for (uint32_t n = 0; n <= max_symbol; ++n) t->code[n] = t->nb_bits[n] ? val_per_rank[t->nb_bits[n]]++ : (uint16_t)0;
    t->table_log = max_bits;
    t->max_symbol = max_symbol;
    return max_bits;
}

__device__ uint32_t huf_estimate(const HufTable* t, const uint32_t* count, uint32_t max_symbol) // :802-811
{
    uint64_t bits = 0;
    for (uint32_t s = 0; s <= max_symbol; ++s) bits += (uint64_t)t->nb_bits[s] * count[s];
    return (uint32_t)(bits >> 3);
}
__device__ bool huf_validate(const HufTable* t, const uint32_t* count, uint32_t max_symbol) // :813-828
{
    bool bad = false;
    if (t->max_symbol < max_symbol) return false;
    for (uint32_t s = 0; s <= max_symbol; ++s) bad |= (count[s] != 0) & (t->nb_bits[s] == 0);
    return !bad;
}

// ---- warp-level pieces of the literal stage.  WarpShared = this warp's slice of shared memory.
struct WarpShared
{
    uint32_t hist[256]; // byte histogram
    uint32_t tab[256];  // Huffman table of the block being encoded: code | nb_bits << 16
};

// HIST_count (hist.c:29-56) with all lanes: shared-memory atomics, then the counts go to `count` (global) for the serial stages
__device__ uint32_t hist_w(WarpShared* sh, uint32_t* count, uint32_t* max_symbol, const uint8_t* src, uint32_t n, uint32_t lane)
{
    for (uint32_t i = lane; i < 256; i += 32) sh->hist[i] = 0;
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) atomicAdd(&sh->hist[src[i]], 1u);
    __syncwarp();
    uint32_t largest = 0, top = 0;
    for (uint32_t i = lane; i < 256; i += 32)
    {
        const uint32_t c = sh->hist[i];
        count[i] = c;
        if (c > largest) largest = c;
        if (c) top = i;
    }
    for (int d = 16; d; d >>= 1)
    {
        largest = max(largest, __shfl_xor_sync(FULL, largest, d));
        top = max(top, __shfl_xor_sync(FULL, top, d));
    }
    __syncwarp();
    *max_symbol = n ? top : 0;
    return largest;
}

__device__ void load_tab_w(WarpShared* sh, const HufTable* t, uint32_t lane)
{
    for (uint32_t i = lane; i < 256; i += 32) sh->tab[i] = (uint32_t)t->code[i] | ((uint32_t)t->nb_bits[i] << 16);
    __syncwarp();
}

// One Huffman stream (huf_compress.c:984-1110: symbols last to first, codes LSB first, closed by a 1 bit) with all lanes: the
// input is cut into 32 contiguous pieces, a first pass sizes every piece, a suffix scan gives each piece its bit offset (the
// LAST piece comes first in the stream), and every lane ORs its bits into the zeroed output words.
__device__ uint32_t huf_encode_1x_w(const WarpShared* sh, uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane)
{
    const uint32_t piece = (n + 31) / 32;
    const uint32_t lo = min(n, lane * piece), hi = min(n, lo + piece);
    uint32_t bits = 0;
    for (uint32_t i = lo; i < hi; ++i) bits += sh->tab[src[i]] >> 16;
    // suffix sum over lanes: offset = bits of all higher lanes
    uint32_t incl = bits;
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t v = __shfl_down_sync(FULL, incl, d);
        if (lane + d < 32) incl += v;
    }
    const uint32_t total = __shfl_sync(FULL, incl, 0);
    const uint32_t size = (total + 8) >> 3; // + the end mark, rounded up to bytes
    for (uint32_t i = lane; i < size; i += 32) dst[i] = 0;
    __syncwarp();
    uint32_t* const words = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(dst) & ~(uintptr_t)3);
    uint32_t cur = (incl - bits) + (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u) * 8u;
    uint32_t widx = cur >> 5, fill = cur & 31u;
    uint64_t acc = 0;
    for (uint32_t i = hi; i-- > lo;)
    {
        const uint32_t e = sh->tab[src[i]];
        acc |= (uint64_t)(e & 0xffffu) << fill;
        fill += e >> 16;
        if (fill >= 32) { atomicOr(&words[widx++], (uint32_t)acc); acc >>= 32; fill -= 32; }
    }
    if (lane == 0) { acc |= 1ull << fill; fill += 1; } // the end mark follows the first symbol, which lane 0 holds
    if (fill) atomicOr(&words[widx], (uint32_t)acc);
    if (fill > 32) atomicOr(&words[widx + 1], (uint32_t)(acc >> 32));
    __syncwarp();
    return size;
}
__device__ uint32_t huf_encode_4x_w(const WarpShared* sh, uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane) // :1168-1213
{
    const uint32_t seg = (n + 3) / 4;
    uint32_t pos = 6;
    if (n < 12) return 0;
    for (int i = 0; i < 4; ++i)
    {
        const uint32_t len = i < 3 ? seg : n - 3 * seg;
        const uint32_t c = huf_encode_1x_w(sh, dst + pos, src + (uint32_t)i * seg, len, lane);
        if (c == 0 || c > 65535) return 0;
        if (i < 3 && lane == 0) { dst[2 * i] = (uint8_t)c; dst[2 * i + 1] = (uint8_t)(c >> 8); }
        pos += c;
    }
    __syncwarp();
    return pos;
}
__device__ uint32_t huf_encode_with_w(WarpShared* sh, uint8_t* dst, uint32_t head, const uint8_t* src, uint32_t n, bool four, const HufTable* t,
                                      uint32_t lane) // :1223-1239
{
    load_tab_w(sh, t, lane);
    const uint32_t c = four ? huf_encode_4x_w(sh, dst + head, src, n, lane) : huf_encode_1x_w(sh, dst + head, src, n, lane);
    if (c == 0) return 0;
    if (head + c >= n - 1) return 0;
    return head + c;
}

// HUF_compress_internal (:1334-1431), warp-level: histogram and bit packing use all lanes, the table construction and its
// description (a few thousand dependent steps over <= 256 symbols) run on lane 0.  `t` enters as the previous block's table
// and leaves as the table the next block may reuse; *repeat enters as that table's status (all lanes hold the same value).
// Returns 0 = not compressible, 1 = single symbol, ZS_ERR.
__device__ uint32_t huf_compress_w(ZstdWorker* W, WarpShared* sh, uint8_t* dst, const uint8_t* src, uint32_t n, bool four, HufTable* t, int* repeat,
                                   bool prefer_repeat, bool suspect, uint32_t lane)
{
    uint32_t* count = W->count;
    uint32_t max_symbol = 255;
    if (!n) return 0;
    if (prefer_repeat && *repeat == 2) return huf_encode_with_w(sh, dst, 0, src, n, four, t, lane);
    if (suspect && n >= 4096 * 10)
    {
        uint32_t m = 255;
        uint32_t total = hist_w(sh, count, &m, src, 4096, lane);
        m = 255;
        total += hist_w(sh, count, &m, src + n - 4096, 4096, lane);
        if (total <= ((2 * 4096) >> 7) + 4) return 0;
    }
    {
        const uint32_t largest = hist_w(sh, count, &max_symbol, src, n, lane);
        if (largest == n) { if (lane == 0) dst[0] = src[0]; return 1; }
        if (largest <= (n >> 7) + 4) return 0;
    }
    // lane 0: table decisions.  verdict: 0 = return `h` as the result, 1 = encode with the old table, 2 = encode with the new one
    uint32_t verdict = 0, h = 0;
    int rep_now = *repeat;
    if (lane == 0)
    {
        if (rep_now == 1 && !huf_validate(t, count, max_symbol)) rep_now = 0;
        if (prefer_repeat && rep_now != 0) verdict = 1;
        else
        {
            HufTable* fresh = &W->huf_fresh;
            huf_build_table(W, fresh, count, max_symbol, fse_optimal_table_log(11, n, max_symbol, 1)); // HUF_optimalTableLog w/o depth search
            h = huf_write_table(W, dst, fresh);
            if (h != ZS_ERR)
            {
                bool use_old = false;
                if (rep_now != 0)
                {
                    const uint32_t old_size = huf_estimate(t, count, max_symbol), new_size = huf_estimate(fresh, count, max_symbol);
                    use_old = old_size <= h + new_size || h + 12 >= n;
                }
                if (use_old) verdict = 1;
                else if (h + 12 >= n) { verdict = 0; h = 0; }
                else
                {
                    verdict = 2;
                    rep_now = 0;
                    fresh->repeat = t->repeat;
                    *t = *fresh;
                }
            }
        }
    }
    verdict = __shfl_sync(FULL, verdict, 0);
    h = __shfl_sync(FULL, h, 0);
    *repeat = __shfl_sync(FULL, rep_now, 0);
    __syncwarp();
    if (verdict == 0) return h; // 0 or ZS_ERR
    return huf_encode_with_w(sh, dst, verdict == 2 ? h : 0, src, n, four, t, lane);
}

__device__ uint32_t lit_header_plain(uint8_t* dst, uint32_t type, uint32_t n) // zstd_compress_literals.c:39-63, :78-104
{
    const uint32_t fl = 1 + (n > 31) + (n > 4095);
    if (fl == 1) dst[0] = (uint8_t)(type + (n << 3));
    else if (fl == 2) { const uint32_t v = type + (1u << 2) + (n << 4); dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); }
    else { const uint32_t v = type + (3u << 2) + (n << 4); dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); dst[3] = (uint8_t)(v >> 24); }
    return fl;
}
__device__ uint32_t lit_raw_w(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane)
{
    const uint32_t fl = 1 + (n > 31) + (n > 4095);
    if (lane == 0) lit_header_plain(dst, 0, n);
    copy_w<false>(dst + fl, src, n, lane);
    __syncwarp();
    return n + fl;
}
__device__ uint32_t lit_rle_w(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane)
{
    const uint32_t fl = 1 + (n > 31) + (n > 4095);
    if (lane == 0) { lit_header_plain(dst, 1, n); dst[fl] = src[0]; }
    __syncwarp();
    return fl + 1;
}
__device__ void copy_huf_table_w(HufTable* dst, const HufTable* src, uint32_t lane)
{
    const uint32_t* a = reinterpret_cast<const uint32_t*>(src);
    uint32_t* b = reinterpret_cast<uint32_t*>(dst);
    for (uint32_t i = lane; i < sizeof(HufTable) / 4; i += 32) b[i] = a[i];
    __syncwarp();
}

// ZSTD_compressLiterals at strategy dfast (zstd_compress_literals.c:129-235), warp-level
__device__ uint32_t compress_literals_w(ZstdWorker* W, WarpShared* sh, uint8_t* dst, const uint8_t* src, uint32_t n, const HufTable* prev, HufTable* next,
                                        bool suspect, uint32_t lane)
{
    const uint32_t lh = 3 + (n >= 1024) + (n >= 16384);
    bool single = n < 256;
    uint32_t type = 2; // set_compressed
    copy_huf_table_w(next, prev, lane);
    int repeat = prev->repeat;
    if (n < (repeat == 2 ? 6u : 64u)) return lit_raw_w(dst, src, n, lane);
    if (repeat == 2 && lh == 3) single = true;
    const uint32_t c = huf_compress_w(W, sh, dst + lh, src, n, !single, next, &repeat, n <= 1024, suspect, lane);
    if (repeat != 0) type = 3; // set_repeat
    {
        const uint32_t min_gain = (n >> 6) + 2;
        if (c == 0 || c == ZS_ERR || c >= n - min_gain) { copy_huf_table_w(next, prev, lane); return lit_raw_w(dst, src, n, lane); }
    }
    if (c == 1)
    {
        bool same = true;
        if (n < 8) for (uint32_t i = 1; i < n; ++i) if (src[i] != src[0]) same = false;
        if (n >= 8 || same) { copy_huf_table_w(next, prev, lane); return lit_rle_w(dst, src, n, lane); }
    }
    if (lane == 0)
    {
        if (type == 2) next->repeat = 1; // HUF_repeat_check
        if (lh == 3) { const uint32_t v = type + ((uint32_t)(!single) << 2) + (n << 4) + (c << 14); dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); }
        else if (lh == 4) { const uint32_t v = type + (2u << 2) + (n << 4) + (c << 18); dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); dst[3] = (uint8_t)(v >> 24); }
        else { const uint32_t v = type + (3u << 2) + (n << 4) + (c << 22); dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); dst[3] = (uint8_t)(v >> 24); dst[4] = (uint8_t)(c >> 10); }
    }
    __syncwarp();
    return lh + c;
}

// ================================================================== sequences

// extra bits per code (RFC 8878 3.1.1.3.2.1.1; zstd_internal.h:118-146) and the predefined distributions (:124-165)
__constant__ uint8_t c_ll_bits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
__constant__ uint8_t c_ml_bits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                      1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
__constant__ int16_t c_ll_default[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
__constant__ int16_t c_ml_default[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                         1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
__constant__ int16_t c_of_default[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};

// the code of a value is the last code whose baseline does not exceed it (zstd_compress_internal.h:517-546)
__device__ __forceinline__ uint32_t ll_code_of(uint32_t v)
{
    if (v < 16) return v;
    if (v < 24) return 16 + ((v - 16) >> 1);
    if (v < 32) return 20 + ((v - 24) >> 2);
    if (v < 48) return 22 + ((v - 32) >> 3);
    if (v < 64) return 24;
    return hibit(v) + 19;
}
__device__ __forceinline__ uint32_t ml_code_of(uint32_t v) // v = match length - 3
{
    if (v < 32) return v;
    if (v < 40) return 32 + ((v - 32) >> 1);
    if (v < 48) return 36 + ((v - 40) >> 2);
    if (v < 64) return 38 + ((v - 48) >> 3);
    if (v < 96) return 40 + ((v - 64) >> 4);
    if (v < 128) return 42;
    return hibit(v) + 36;
}

// ZSTD_selectEncodingType for strategy < lazy (zstd_compress_sequences.c:157-206); 0 basic, 1 rle, 2 compressed.  Without a
// dictionary the repeat mode never becomes "valid", so set_repeat is unreachable.
__device__ int select_encoding(uint32_t most, uint32_t nb_seq, uint32_t default_log, bool default_allowed)
{
    if (most == nb_seq) return (default_allowed && nb_seq <= 2) ? 0 : 1;
    if (default_allowed)
    {
        const uint32_t dynamic_min = ((1u << default_log) * (10 - 2)) >> 3;
        if (nb_seq < dynamic_min || most < (nb_seq >> (default_log - 1))) return 0;
    }
    return 2;
}

// ZSTD_buildCTable (zstd_compress_sequences.c:243-290): bytes of table description written
__device__ uint32_t build_seq_table(ZstdWorker* W, uint8_t* dst, FseTable* ct, uint32_t max_log, int type, uint32_t* count, uint32_t max,
                                    const uint8_t* codes, uint32_t nb_seq, const int16_t* default_norm, uint32_t default_log, uint32_t default_max)
{
    if (type == 1) { fse_build_ctable_rle(ct, max); dst[0] = codes[0]; return 1; }
    if (type == 0)
    {
        for (uint32_t i = 0; i <= default_max; ++i) W->norm[i] = default_norm[i];
        fse_build_ctable(W, ct, W->norm, default_max, default_log);
        return 0;
    }
    uint32_t n1 = nb_seq;
    const uint32_t log = fse_optimal_table_log(max_log, nb_seq, max, 2);
    if (count[codes[nb_seq - 1]] > 1) { count[codes[nb_seq - 1]]--; n1--; }
    if (fse_normalize(W->norm, log, count, n1, max, n1 >= 2048) < 0) return ZS_ERR;
    const uint32_t h = fse_write_ncount(dst, W->norm, max, log);
    if (h == ZS_ERR) return ZS_ERR;
    fse_build_ctable(W, ct, W->norm, max, log);
    return h;
}

// the sequence section of one block (zstd_compress.c:2930-2990), warp-level; `op` follows the literal section.
//   * code histograms with all lanes, table construction (normalise, NCount, state table: a few hundred dependent steps) on lane 0;
//   * the three FSE state chains (offset, match length, literal length) are independent of each other and of the extra bits, so
//     lanes 0-2 walk one chain each and record, per sequence, the bits that transition emits;
//   * the bitstream (ZSTD_encodeSequences, zstd_compress_sequences.c:293-385: last sequence first) is then packed by all lanes:
//     pieces of the sequence range are sized, suffix-scanned and ORed into the zeroed output, like the Huffman streams.
__device__ uint32_t sequences_compress_w(ZstdWorker* W, WarpShared* sh, uint8_t* dst, uint8_t* op, uint32_t nb_seq, uint32_t lane)
{
    if (lane == 0)
    {
        if (nb_seq < 128) op[0] = (uint8_t)nb_seq;
        else if (nb_seq < 0x7F00) { op[0] = (uint8_t)((nb_seq >> 8) + 0x80); op[1] = (uint8_t)nb_seq; }
        else { op[0] = 0xFF; const uint32_t v = nb_seq - 0x7F00; op[1] = (uint8_t)v; op[2] = (uint8_t)(v >> 8); }
    }
    op += nb_seq < 128 ? 1 : nb_seq < 0x7F00 ? 2 : 3;
    if (nb_seq == 0) return (uint32_t)(op - dst);

    const uint8_t* ll_code = W->ll_code;
    const uint8_t* of_code = W->of_code;
    const uint8_t* ml_code = W->ml_code;
    uint8_t* seq_head = op++;
    uint32_t* count = W->count;
    uint32_t last_count_size = 0;
    uint32_t types[3];
    for (int k = 0; k < 3; ++k) // 0 literal lengths, 1 offsets, 2 match lengths: the order of the table descriptions
    {
        const uint8_t* codes = k == 0 ? ll_code : k == 1 ? of_code : ml_code;
        uint32_t max = k == 0 ? 35 : k == 1 ? 31 : 52;
        const uint32_t most = hist_w(sh, count, &max, codes, nb_seq, lane);
        uint32_t type = 0, h = 0;
        if (lane == 0)
        {
            if (k == 0) { type = select_encoding(most, nb_seq, 6, true); h = build_seq_table(W, op, &W->ct_ll, 9, type, count, max, codes, nb_seq, c_ll_default, 6, 35); }
            else if (k == 1) { type = select_encoding(most, nb_seq, 5, max <= 28); h = build_seq_table(W, op, &W->ct_of, 8, type, count, max, codes, nb_seq, c_of_default, 5, 28); }
            else { type = select_encoding(most, nb_seq, 6, true); h = build_seq_table(W, op, &W->ct_ml, 9, type, count, max, codes, nb_seq, c_ml_default, 6, 52); }
        }
        type = __shfl_sync(FULL, type, 0);
        h = __shfl_sync(FULL, h, 0);
        __syncwarp();
        if (h == ZS_ERR) return ZS_ERR;
        if (type == 2) last_count_size = h;
        types[k] = type;
        op += h;
    }
    if (lane == 0) *seq_head = (uint8_t)((types[0] << 6) + (types[1] << 4) + (types[2] << 2));

    // ---- state chains: lane 0 offsets, lane 1 match lengths, lane 2 literal lengths (fse.h:452-470)
    uint32_t final_state = 0, final_log = 0;
    if (lane < 3)
    {
        const FseTable* ct = lane == 0 ? &W->ct_of : lane == 1 ? &W->ct_ml : &W->ct_ll;
        const uint8_t* codes = lane == 0 ? of_code : lane == 1 ? ml_code : ll_code;
        uint16_t* enc = lane == 0 ? W->enc_of : lane == 1 ? W->enc_ml : W->enc_ll;
        uint32_t n = nb_seq - 1;
        uint32_t value;
        {
            const uint32_t sym = codes[n];
            const uint32_t nbi = (ct->delta_nb_bits[sym] + (1u << 15)) >> 16;
            const uint32_t v = (nbi << 16) - ct->delta_nb_bits[sym];
            value = ct->next_state[(int32_t)(v >> nbi) + ct->delta_find_state[sym]];
        }
        while (n-- > 0)
        {
            const uint32_t sym = codes[n];
            const uint32_t nbo = (value + ct->delta_nb_bits[sym]) >> 16;
            enc[n] = (uint16_t)((value & ((1u << nbo) - 1u)) | (nbo << 12));
            value = ct->next_state[(int32_t)(value >> nbo) + ct->delta_find_state[sym]];
        }
        final_state = value;
        final_log = ct->table_log;
    }
    __syncwarp();
    const uint32_t fs_of = __shfl_sync(FULL, final_state, 0), fl_of = __shfl_sync(FULL, final_log, 0);
    const uint32_t fs_ml = __shfl_sync(FULL, final_state, 1), fl_ml = __shfl_sync(FULL, final_log, 1);
    const uint32_t fs_ll = __shfl_sync(FULL, final_state, 2), fl_ll = __shfl_sync(FULL, final_log, 2);

    // ---- pack
    const uint32_t piece = (nb_seq + 31) / 32;
    const uint32_t lo = min(nb_seq, lane * piece), hi = min(nb_seq, lo + piece);
    uint32_t bits = 0;
    for (uint32_t n = lo; n < hi; ++n)
    {
        bits += c_ll_bits[ll_code[n]] + c_ml_bits[ml_code[n]] + of_code[n];
        if (n + 1 < nb_seq) bits += (W->enc_of[n] >> 12) + (W->enc_ml[n] >> 12) + (W->enc_ll[n] >> 12);
    }
    if (lane == 0) bits += fl_ml + fl_of + fl_ll;
    uint32_t incl = bits;
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t v = __shfl_down_sync(FULL, incl, d);
        if (lane + d < 32) incl += v;
    }
    const uint32_t total = __shfl_sync(FULL, incl, 0);
    const uint32_t stream = (total + 8) >> 3;
    for (uint32_t i = lane; i < stream; i += 32) op[i] = 0;
    __syncwarp();
    {
        uint32_t* const words = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(op) & ~(uintptr_t)3);
        const uint32_t cur = (incl - bits) + (uint32_t)(reinterpret_cast<uintptr_t>(op) & 3u) * 8u;
        uint32_t widx = cur >> 5, fill = cur & 31u;
        uint64_t acc = 0;
#define ZS_PUT(value, count_)                                                                                   \
    {                                                                                                           \
        const uint32_t nb_ = (count_);                                                                          \
        acc |= (uint64_t)((value) & ((1u << nb_) - 1u)) << fill;                                                \
        fill += nb_;                                                                                            \
        if (fill >= 32) { atomicOr(&words[widx++], (uint32_t)acc); acc >>= 32; fill -= 32; }                    \
    }
        for (uint32_t n = hi; n-- > lo;)
        {
            if (n + 1 < nb_seq)
            {
                const uint32_t eo = W->enc_of[n], em = W->enc_ml[n], el = W->enc_ll[n];
                ZS_PUT(eo, eo >> 12);
                ZS_PUT(em, em >> 12);
                ZS_PUT(el, el >> 12);
            }
            ZS_PUT(W->seq_lit[n], c_ll_bits[ll_code[n]]);
            ZS_PUT(W->seq_len[n] - 3, c_ml_bits[ml_code[n]]);
            ZS_PUT(W->seq_off[n], of_code[n]);
        }
        if (lane == 0)
        {
            ZS_PUT(fs_ml, fl_ml);
            ZS_PUT(fs_of, fl_of);
            ZS_PUT(fs_ll, fl_ll);
            ZS_PUT(1u, 1u); // end mark
        }
#undef ZS_PUT
        if (fill) atomicOr(&words[widx], (uint32_t)acc);
    }
    __syncwarp();
    op += stream;
    if (last_count_size && last_count_size + stream < 4) return 0; // zstd <= 1.3.4 decoder quirk, zstd_compress.c:2982-2988
    return (uint32_t)(op - dst);
}

// literals + sequences of one block (zstd_compress.c:2876-2990), warp-level; 0 = emit the block raw, ZS_ERR on error
__device__ uint32_t entropy_compress_w(ZstdWorker* W, WarpShared* sh, uint8_t* dst, uint32_t lit_size, uint32_t nb_seq, uint32_t lane)
{
    const long long t0 = clock64();
    const uint32_t lit_bytes = compress_literals_w(W, sh, dst, W->lits, lit_size, &W->huf_prev, &W->huf_next, (nb_seq == 0) || (lit_size / nb_seq >= 20), lane);
    const long long t1 = clock64();
    // sequence codes with all lanes (ZSTD_seqToCodes, zstd_compress.c:2681-2705)
    for (uint32_t i = lane; i < nb_seq; i += 32)
    {
        W->ll_code[i] = (uint8_t)ll_code_of(W->seq_lit[i]);
        W->of_code[i] = (uint8_t)hibit(W->seq_off[i]);
        W->ml_code[i] = (uint8_t)ml_code_of(W->seq_len[i] - 3);
    }
    __syncwarp();
    const uint32_t c = sequences_compress_w(W, sh, dst, dst + lit_bytes, nb_seq, lane);
    if (lane == 0) { W->t_phase[1] += (unsigned long long)(t1 - t0); W->t_phase[2] += (unsigned long long)(clock64() - t1); }
    return c;
}

// ================================================================== double-fast matcher

__device__ __forceinline__ uint32_t hash_long(uint64_t v, uint32_t bits) { return (uint32_t)((v * 0xCF1BBCDCB7A56463ull) >> (64 - bits)); }
__device__ __forceinline__ uint32_t hash_short_w(uint64_t w, uint32_t bits, uint32_t mls) // zstd_compress_internal.h:803-841, from the 8 bytes at the position
{
    if (mls == 5) return (uint32_t)(((w << 24) * 889523592379ull) >> (64 - bits));
    return ((uint32_t)w * 2654435761u) >> (32 - bits);
}
__device__ __forceinline__ uint4 entry_long(uint32_t index, uint64_t w) { return make_uint4(index, (uint32_t)w, (uint32_t)(w >> 32), 0u); }
__device__ __forceinline__ uint2 entry_short(uint32_t index, uint64_t w) { return make_uint2(index, (uint32_t)w); }
// ZSTD_count (zstd_compress_internal.h:744-767) with all lanes: common bytes of s[a..) and s[b..), a > b, a bounded by end
__device__ uint32_t count_equal_w(const uint8_t* __restrict__ s, uint32_t a, uint32_t b, uint32_t end, uint32_t lane)
{
    uint32_t total = 0;
    for (;;)
    {
        const uint32_t pa = a + 4 * lane;
        uint32_t eq = 0;
        bool stop = true;
        if (pa < end)
        {
            const uint32_t x = rd32(s, pa) ^ rd32(s, b + 4 * lane);
            eq = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
            const uint32_t room = end - pa;
            if (eq > room) eq = room;
            stop = eq < 4u;
        }
        const uint32_t stops = __ballot_sync(FULL, stop);
        if (stops)
        {
            const uint32_t f = (uint32_t)__ffs(stops) - 1u;
            return total + 4u * f + __shfl_sync(FULL, eq, f);
        }
        total += 128;
        a += 128;
        b += 128;
    }
}
// backward catch-up (zstd_double_fast.c:192, :256, :265) with all lanes: how many bytes before ip / m are equal
__device__ uint32_t catch_up_w(const uint8_t* __restrict__ s, uint32_t ip, uint32_t m, uint32_t anchor, uint32_t lowest, uint32_t lane)
{
    uint32_t back = 0;
    for (;;)
    {
        bool eq = false;
        if (ip > anchor + lane && m > lowest + lane) eq = s[ip - 1 - lane] == s[m - 1 - lane];
        const uint32_t mask = __ballot_sync(FULL, eq);
        const uint32_t run = mask == FULL ? 32u : (uint32_t)__ffs(~mask) - 1u;
        back += run;
        if (run < 32) return back;
        ip -= 32;
        m -= 32;
    }
}

// ZSTD_compressBlock_doubleFast_noDict_generic (zstd_double_fast.c:105-311) by one warp.  Positions are frame offsets; the
// match index of frame byte p is p + 2 (ZSTD_WINDOW_START_INDEX, zstd_compress_internal.h:211), 0 in a table = empty.
//
// The parse is sequential, but between two matches it only walks a deterministic stride schedule (step grows by one every
// 256 bytes, :253-258) and every probe does the same thing: read both tables, write both tables, test repcode / long / short
// candidates.  The warp therefore SPECULATES the next 32 probes at once: lane k takes the k-th position of the schedule; a
// table read that an earlier lane of the same batch would have overwritten is resolved inside the warp (__match_any_sync:
// the candidate is the closest lower lane with the same hash, else the table); the first lane whose probe matches wins
// (__ballot_sync) and only the table writes up to that lane are committed — exactly the state the sequential loop would have
// reached.  Match extension, catch-up and literal copies are lane-parallel; the few table insertions after a match and the
// immediate-repcode loop are warp-uniform.
__device__ uint32_t dfast_block_w(ZstdWorker* W, const uint8_t* __restrict__ s, uint32_t block_start, uint32_t block_size, uint32_t rep[3],
                                  uint32_t* lit_total, uint32_t lane)
{
    const uint32_t hl_bits = W->hash_log, hs_bits = W->chain_log, mls = W->min_match;
    uint4* const hash_l = W->hash_long;
    uint2* const hash_s = W->hash_small;
    uint8_t* const lits = W->lits;
    const uint32_t max_dist = 1u << W->window_log;
    const uint32_t dict_limit = W->dict_limit;
    const uint32_t end_index = block_start + 2 + block_size;
    const uint32_t lowest_index = (end_index - dict_limit > max_dist) ? end_index - max_dist : dict_limit;
    const uint32_t lowest = lowest_index - 2; // as a position
    const uint32_t iend = block_start + block_size;
    const int32_t ilimit = (int32_t)iend - 8; // may be negative for a 7-byte frame: positions are < 2^31, compare signed
    uint32_t offset_1 = rep[0], offset_2 = rep[1], saved_1 = 0, saved_2 = 0;
    uint32_t nb = 0, nlit = 0;
    uint32_t anchor = block_start;
    uint32_t ip = block_start;

    ip += (ip == lowest);
    {
        const uint32_t current = ip + 2;
        const uint32_t window_low = (current - dict_limit > max_dist) ? current - max_dist : dict_limit;
        const uint32_t max_rep = current - window_low;
        if (offset_2 > max_rep) { saved_2 = offset_2; offset_2 = 0; }
        if (offset_1 > max_rep) { saved_1 = offset_1; offset_1 = 0; }
    }
    for (;;) // one iteration per stored match (the reference's outer loop)
    {
        // schedule state at the start of the next unprobed iteration
        uint32_t q_ip = ip, q_ip1 = ip + 1, q_step = 1, q_next = ip + 256; // kSearchStrength = 8 (:35)
        bool found = false, finished = false;
        uint32_t P = 0, P1 = 0, f_step = 0, f_type = 0, f_cand = 0, hl1 = 0, idxl1 = 0;
        uint64_t w_f = 0, w1 = 0, tagl1 = 0;
        for (;;) // batches of 32 speculative probes
        {
            // lane k's iteration: position cp, next position cp1, stride in force cs
            uint32_t cp, cp1, cs = q_step, cn = q_next;
            if (q_ip1 + 32u * q_step < q_next) // no stride change inside this batch (the common case)
            {
                cp = lane ? q_ip1 + (lane - 1) * q_step : q_ip;
                cp1 = q_ip1 + lane * q_step;
            }
            else
            {
                cp = q_ip; cp1 = q_ip1;
                for (uint32_t j = 0; j < lane; ++j)
                {
                    if (cp1 >= cn) { cs++; cn += 256; }
                    cp = cp1;
                    cp1 += cs;
                }
            }
            const bool readable = (int32_t)cp <= ilimit;  // 8 bytes can be hashed here
            const bool valid = (int32_t)cp1 <= ilimit;    // the sequential loop executes this iteration
            const uint64_t w = readable ? rd64(s, cp) : 0ull;
            const uint32_t hl = readable ? hash_long(w, hl_bits) : 0xffffffffu - lane;
            const uint32_t hs = readable ? hash_short_w(w, hs_bits, mls) : 0xffffffffu - lane;
            // both table entries are requested BEFORE the intra-batch conflict resolution: they do not depend on it, and the two MATCH.ANY
            // plus the dependent shuffles take about as long as the loads (a lane that ends up using a lower lane's position drops them)
            uint4 el = make_uint4(0u, 0u, 0u, 0u);
            uint2 es = make_uint2(0u, 0u);
            if (readable)
            {
                el = hash_l[hl];
                es = hash_s[hs];
            }
            const uint32_t valids = __ballot_sync(FULL, valid);
            const uint32_t below = (1u << lane) - 1u;
            const uint32_t same_l = __match_any_sync(FULL, hl);
            const uint32_t same_s = __match_any_sync(FULL, hs);
            const uint32_t low_l = same_l & valids & below, low_s = same_s & valids & below;
            const int src_l = low_l ? 31 - __clz(low_l) : (int)lane, src_s = low_s ? 31 - __clz(low_s) : (int)lane;
            const uint32_t from_l = __shfl_sync(FULL, cp, src_l);
            const uint32_t from_s = __shfl_sync(FULL, cp, src_s);
            const uint64_t from_wl = __shfl_sync(FULL, w, src_l);
            const uint32_t from_ws = __shfl_sync(FULL, (uint32_t)w, src_s);
            const uint32_t cand_l = low_l ? from_l + 2 : el.x;
            const uint32_t cand_s = low_s ? from_s + 2 : es.x;
            const uint64_t tag_l = low_l ? from_wl : ((uint64_t)el.y | ((uint64_t)el.z << 32)); // the 8 bytes at the long candidate
            const uint32_t tag_s = low_s ? from_ws : es.y;                                       // the 4 bytes at the short candidate
            uint32_t type = 0; // 1 repcode at ip+1, 2 long match at ip, 3 short match at ip (then the long match at ip1 is preferred)
            if (valid)
            {
                if (offset_1 > 0 && rd32(s, cp + 1 - offset_1) == (uint32_t)(w >> 8)) type = 1;
                else if (cand_l > lowest_index && tag_l == w) type = 2;
                else if (cand_s > lowest_index && tag_s == (uint32_t)w) type = 3;
            }
            const uint32_t hits = __ballot_sync(FULL, type != 0);
            const uint32_t first = hits ? (uint32_t)__ffs(hits) - 1u : 32u;
            const uint32_t commit = valids & (first >= 31u ? FULL : ((2u << first) - 1u));
            __syncwarp(); // every table read of the batch precedes every table write
            if ((commit >> lane) & 1u)
            {
                if ((31 - __clz(same_l & commit)) == (int)lane) hash_l[hl] = entry_long(cp + 2, w); // the last writer of a slot wins
                if ((31 - __clz(same_s & commit)) == (int)lane) hash_s[hs] = entry_short(cp + 2, w);
            }
            __syncwarp();
            if (hits)
            {
                P = __shfl_sync(FULL, cp, first);
                P1 = __shfl_sync(FULL, cp1, first);
                f_step = __shfl_sync(FULL, cs, first);
                f_type = __shfl_sync(FULL, type, first);
                f_cand = __shfl_sync(FULL, f_type == 2 ? cand_l : cand_s, first);
                w_f = __shfl_sync(FULL, w, first);
                if (first < 31)
                {
                    // the next lane probed ip1: its hash and its (batch-resolved) long candidate are hl1 / idxl1 of :175, :199
                    hl1 = __shfl_sync(FULL, hl, first + 1);
                    idxl1 = __shfl_sync(FULL, cand_l, first + 1);
                    tagl1 = __shfl_sync(FULL, tag_l, first + 1);
                    w1 = __shfl_sync(FULL, w, first + 1);
                }
                else
                {
                    w1 = rd64(s, P1);
                    hl1 = hash_long(w1, hl_bits);
                    const uint4 e1 = hash_l[hl1];
                    idxl1 = e1.x;
                    tagl1 = (uint64_t)e1.y | ((uint64_t)e1.z << 32);
                }
                found = true;
                break;
            }
            if (valids != FULL) { finished = true; break; }
            // all 32 iterations ran without a match: advance the schedule by the state lane 31 ended in
            {
                uint32_t n_cs = cs, n_cn = cn;
                if (cp1 >= n_cn) { n_cs++; n_cn += 256; }
                q_ip = __shfl_sync(FULL, cp1, 31);
                q_ip1 = __shfl_sync(FULL, cp1 + n_cs, 31);
                q_step = __shfl_sync(FULL, n_cs, 31);
                q_next = __shfl_sync(FULL, n_cn, 31);
            }
        }
        if (finished || !found) break;

        uint32_t m_len, offset = 0;
        const uint32_t curr = P + 2;
        ip = P;
        if (f_type == 1)
        {
            m_len = count_equal_w(s, ip + 1 + 4, ip + 1 + 4 - offset_1, iend, lane) + 4;
            ip++;
        }
        else
        {
            uint32_t m;
            if (f_type == 2)
            {
                m = f_cand - 2;
                m_len = count_equal_w(s, ip + 8, m + 8, iend, lane) + 8;
            }
            else if (idxl1 > lowest_index && tagl1 == w1) // _search_next_long (:238-259)
            {
                ip = P1;
                m = idxl1 - 2;
                m_len = count_equal_w(s, ip + 8, m + 8, iend, lane) + 8;
            }
            else
            {
                m = f_cand - 2;
                m_len = count_equal_w(s, ip + 4, m + 4, iend, lane) + 4;
            }
            offset = ip - m;
            const uint32_t back = catch_up_w(s, ip, m, anchor, lowest, lane);
            ip -= back;
            m_len += back;
            offset_2 = offset_1;
            offset_1 = offset;
            if (f_step < 4 && lane == 0) hash_l[hl1] = entry_long(P1 + 2, w1); // :261-273
        }
        {
            const uint32_t ll = ip - anchor;
            copy_w<false>(lits + nlit, s + anchor, ll, lane);
            if (lane == 0) { W->seq_lit[nb] = ll; W->seq_len[nb] = m_len; W->seq_off[nb] = f_type == 1 ? 1u : offset + 3u; }
            nlit += ll;
            nb++;
        }
        ip += m_len;
        anchor = ip;
        __syncwarp();
        if ((int32_t)ip <= ilimit)
        {
            // complementary insertions (:286-291), in the reference's order; all lanes compute, lane 0 stores
            const uint32_t insert = curr + 2; // an index; its position is curr
            const uint64_t w_a = rd64(s, insert - 2), w_b = rd64(s, ip - 2), w_d = rd64(s, ip - 1);
            const uint32_t h_a = hash_long(w_a, hl_bits), h_b = hash_long(w_b, hl_bits);
            const uint32_t h_c = hash_short_w(w_a, hs_bits, mls), h_d = hash_short_w(w_d, hs_bits, mls);
            if (lane == 0)
            {
                hash_l[h_a] = entry_long(insert, w_a);
                hash_l[h_b] = entry_long(ip, w_b);      // index of position ip - 2
                hash_s[h_c] = entry_short(insert, w_a);
                hash_s[h_d] = entry_short(ip + 1, w_d); // index of position ip - 1
            }
            // immediate repcodes (:294-308)
            while ((int32_t)ip <= ilimit && offset_2 > 0 && rd32(s, ip) == rd32(s, ip - offset_2))
            {
                const uint32_t r_len = count_equal_w(s, ip + 4, ip + 4 - offset_2, iend, lane) + 4;
                const uint32_t t = offset_2; offset_2 = offset_1; offset_1 = t;
                const uint64_t wi = rd64(s, ip);
                if (lane == 0)
                {
                    hash_s[hash_short_w(wi, hs_bits, mls)] = entry_short(ip + 2, wi);
                    hash_l[hash_long(wi, hl_bits)] = entry_long(ip + 2, wi);
                    W->seq_lit[nb] = 0; W->seq_len[nb] = r_len; W->seq_off[nb] = 1;
                }
                nb++;
                ip += r_len;
                anchor = ip;
            }
            __syncwarp();
        }
    }
    saved_2 = (saved_1 != 0 && offset_1 != 0) ? saved_1 : saved_2;
    rep[0] = offset_1 ? offset_1 : saved_1;
    rep[1] = offset_2 ? offset_2 : saved_2;
    {
        const uint32_t ll = iend - anchor;
        copy_w<false>(lits + nlit, s + anchor, ll, lane);
        nlit += ll;
    }
    __syncwarp();
    *lit_total = nlit;
    return nb;
}

// level 3 rows of clevels.h:25-132 and ZSTD_adjustCParams_internal (zstd_compress.c:1464-1602) for a known size, no dictionary
__device__ void zs_set_params(ZstdWorker* W, uint32_t n)
{
    const uint32_t row = (n <= (256u << 10)) + (n <= (128u << 10)) + (n <= (16u << 10));
    uint32_t wl = row == 0 ? 21 : row == 1 ? 18 : row == 2 ? 17 : 14;
    uint32_t cl = row == 0 ? 16 : row == 1 ? 16 : row == 2 ? 15 : 14;
    uint32_t hl = row == 0 ? 17 : row == 1 ? 16 : row == 2 ? 16 : 15;
    const uint32_t mm = (row == 0 || row == 2) ? 5 : 4;
    const uint32_t src_log = n < 64 ? 6 : hibit(n - 1) + 1;
    if (wl > src_log) wl = src_log;
    if (hl > wl + 1) hl = wl + 1;
    if (cl > wl) cl = wl;
    if (wl < 10) wl = 10;
    W->window_log = wl; W->chain_log = cl; W->hash_log = hl; W->min_match = mm;
}

// one frame by one warp; returns the frame size (warp-uniform)
__device__ uint32_t zstd_compress_frame(ZstdWorker* W, WarpShared* sh, const uint8_t* __restrict__ src, uint32_t size, uint8_t* __restrict__ dst, uint32_t lane)
{
    if (lane == 0)
    {
        zs_set_params(W, size);
        W->dict_limit = 2;
        W->rep[0] = 1; W->rep[1] = 4; W->rep[2] = 8;
        W->huf_prev.table_log = 0; W->huf_prev.max_symbol = 0; W->huf_prev.repeat = 0;
    }
    __syncwarp();
    const uint32_t window_log = W->window_log;
    {
        // fresh context: both tables start zeroed (ZSTD_reset_matchState, zstd_compress.c:1970-2050)
        uint4* hl = W->hash_long;
        uint4* hs = reinterpret_cast<uint4*>(W->hash_small);
        const uint32_t nl = 1u << W->hash_log, ns = (1u << W->chain_log) / 2;
        for (uint32_t i = lane; i < nl; i += 32) hl[i] = make_uint4(0, 0, 0, 0);
        for (uint32_t i = lane; i < ns; i += 32) hs[i] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    uint32_t op = 0;
    if (lane == 0)
    {
        // ZSTD_writeFrameHeader (zstd_compress.c:4575-4623): content size on, no checksum, no dictionary id
        const uint32_t single = ((uint64_t)1 << window_log) >= size;
        const uint32_t fcs = (size >= 256) + (size >= 65536 + 256);
        dst[0] = 0x28; dst[1] = 0xB5; dst[2] = 0x2F; dst[3] = 0xFD;
        uint32_t p = 4;
        dst[p++] = (uint8_t)((single << 5) + (fcs << 6));
        if (!single) dst[p++] = (uint8_t)((window_log - 10) << 3);
        if (fcs == 0) { if (single) dst[p++] = (uint8_t)size; }
        else if (fcs == 1) { const uint32_t v = size - 256; dst[p++] = (uint8_t)v; dst[p++] = (uint8_t)(v >> 8); }
        else { dst[p++] = (uint8_t)size; dst[p++] = (uint8_t)(size >> 8); dst[p++] = (uint8_t)(size >> 16); dst[p++] = (uint8_t)(size >> 24); }
        if (size == 0) { dst[p++] = 1; dst[p++] = 0; dst[p++] = 0; } // ZSTD_writeEpilogue: one empty raw last block (:5244-5252)
        op = p;
    }
    op = __shfl_sync(FULL, op, 0);
    const uint32_t max_dist = 1u << window_log;
    const uint32_t window_size = size ? (max_dist < size ? max_dist : size) : 1u;
    const uint32_t block_max = window_size < ZS_BLOCK_MAX ? window_size : ZS_BLOCK_MAX;
    bool first_block = true;
    uint32_t pos = 0;
    while (pos < size)
    {
        const uint32_t bs = (size - pos) < block_max ? (size - pos) : block_max;
        const uint32_t last = (pos + bs == size);
        uint32_t c = 0;
        // ZSTD_window_enforceMaxDist, called with the block START (zstd_compress.c:4525)
        if (lane == 0 && pos + 2 > max_dist)
        {
            const uint32_t low = pos + 2 - max_dist;
            if (W->dict_limit < low) W->dict_limit = low;
        }
        __syncwarp();
        if (bs >= 7) // MIN_CBLOCK_SIZE + block header + 2 (zstd_compress.c:3212)
        {
            uint32_t next_rep[3] = {W->rep[0], W->rep[1], W->rep[2]};
            uint32_t lit_size = 0;
            const long long tm = clock64();
            const uint32_t nb = dfast_block_w(W, src, pos, bs, next_rep, &lit_size, lane);
            if (lane == 0) W->t_phase[0] += (unsigned long long)(clock64() - tm);
            c = entropy_compress_w(W, sh, W->scratch, lit_size, nb, lane);
            if (c != ZS_ERR)
            {
                if (c && c >= bs - ((bs >> 6) + 2)) c = 0; // ZSTD_minGain gate, zstd_compress.c:3021-3024
                if (!first_block && c < 25)                 // RLE block, zstd_compress.c:4359-4370
                {
                    const uint8_t v = src[pos];
                    bool differs = false;
                    for (uint32_t i = 1 + lane; i < bs && !differs; i += 32) differs = src[pos + i] != v;
                    if (!__any_sync(FULL, differs)) { c = 1; if (lane == 0) W->scratch[0] = v; }
                }
                if (c > 1) // ZSTD_blockState_confirmRepcodesAndEntropyTables
                {
                    copy_huf_table_w(&W->huf_prev, &W->huf_next, lane);
                    if (lane == 0) { W->rep[0] = next_rep[0]; W->rep[1] = next_rep[1]; W->rep[2] = next_rep[2]; }
                }
            }
        }
        if (lane == 0)
        {
            const uint32_t h = c == 0 || c == ZS_ERR ? last + (0u << 1) + (bs << 3) : c == 1 ? last + (1u << 1) + (bs << 3) : last + (2u << 1) + (c << 3);
            dst[op] = (uint8_t)h; dst[op + 1] = (uint8_t)(h >> 8); dst[op + 2] = (uint8_t)(h >> 16);
        }
        if (c == ZS_ERR) return ZS_ERR;
        __syncwarp();
        const long long tc = clock64();
        op += 3;
        if (c == 0)
        {
            copy_w<false>(dst + op, src + pos, bs, lane);
            op += bs;
        }
        else
        {
            copy_w<true>(dst + op, W->scratch, c, lane); // the bit packers wrote through L2 atomics
            op += c;
        }
        __syncwarp();
        if (lane == 0) W->t_phase[3] += (unsigned long long)(clock64() - tc);
        pos += bs;
        first_block = false;
    }
    return op;
}

} // namespace

// one warp per frame from a global queue; output = the compress block store's {u32 raw size, u32 compressed size} header
// (lib/compressblockstore/longtail_compressblockstore.c:127-131) followed by the frame
__global__ void __launch_bounds__(128, 5)
k_zstd_frames(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ raw_off, const uint32_t* __restrict__ raw_len, uint8_t* __restrict__ out,
              const uint64_t* __restrict__ out_off, uint32_t* __restrict__ out_len, uint32_t frame_count, ZstdWorker* workers, uint32_t* queue)
{
    __shared__ WarpShared s_warp[4];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    ZstdWorker* W = workers + warp;
    WarpShared* sh = &s_warp[threadIdx.x >> 5];
    if (lane == 0) { W->t_phase[0] = 0; W->t_phase[1] = 0; W->t_phase[2] = 0; W->t_phase[3] = 0; }
    __syncwarp();
    for (;;)
    {
        uint32_t f = 0;
        if (lane == 0) f = atomicAdd(queue, 1u);
        f = __shfl_sync(FULL, f, 0);
        if (f >= frame_count) break;
        const uint32_t n = raw_len[f];
        uint8_t* dst = out + out_off[f];
        const uint32_t c = zstd_compress_frame(W, sh, raw + raw_off[f], n, dst + 8, lane);
        if (lane == 0)
        {
            if (c == ZS_ERR) out_len[f] = 0xffffffffu;
            else
            {
                dst[0] = (uint8_t)n; dst[1] = (uint8_t)(n >> 8); dst[2] = (uint8_t)(n >> 16); dst[3] = (uint8_t)(n >> 24);
                dst[4] = (uint8_t)c; dst[5] = (uint8_t)(c >> 8); dst[6] = (uint8_t)(c >> 16); dst[7] = (uint8_t)(c >> 24);
                out_len[f] = c + 8;
            }
        }
        __syncwarp();
    }
}

size_t zstd_worker_bytes() { return sizeof(ZstdWorker); }
size_t zstd_worker_phase_offset() { return offsetof(ZstdWorker, t_phase); }

uint32_t zstd_worker_count(uint32_t frame_count, int sm_count)
{
    const uint32_t resident = (uint32_t)sm_count * 20u; // 5 CTAs of 4 warps per SM (register bound)
    return frame_count < resident ? (frame_count + 3u) & ~3u : resident;
}

cudaError_t launch_zstd_frames(const uint8_t* d_raw, const uint64_t* d_raw_off, const uint32_t* d_raw_len, uint8_t* d_out, const uint64_t* d_out_off,
                               uint32_t* d_out_len, uint32_t frame_count, void* d_workers, uint32_t worker_count, uint32_t* d_queue, cudaStream_t st)
{
    if (!frame_count) return cudaSuccess;
    static int carveout = -2;
    if (carveout == -2)
    {
        // percent of the SM's L1/shared storage given to shared memory: 5 CTAs x 8 KiB need 20 %; the rest stays L1 for the source window and
        // the hot table lines (measured on the 16 GiB probe: default 2 184 ms, 20 %: 2 121 ms, 30 %: 2 141 ms, 45 %: 2 205 ms)
        const char* e = getenv("LT_B200_ZSTD_CARVEOUT");
        carveout = e ? atoi(e) : 20;
        if (carveout >= 0) cudaFuncSetAttribute(k_zstd_frames, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
    }
    cudaMemsetAsync(d_queue, 0, sizeof(uint32_t), st);
    k_zstd_frames<<<worker_count / 4, 128, 0, st>>>(d_raw, d_raw_off, d_raw_len, d_out, d_out_off, d_out_len, frame_count,
                                                     static_cast<ZstdWorker*>(d_workers), d_queue);
    return cudaGetLastError();
}

} // namespace ltb
