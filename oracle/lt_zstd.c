/* oracle/lt_zstd.c — CPU restatement of ZStd "level 3" as longtail reaches it.
 *
 * TEST INFRASTRUCTURE ONLY (see lt_oracle.h).  Parity status: PINNED — tests/test_oracle.py compares the
 * output byte for byte with the unmodified reference (ZStdCompressionAPI_Compress through oracle/ref_shim.c,
 * vendored zstd 1.5.6) on seeded inputs of every parameter row, and with committed fixtures (tests/golden/).
 *
 * Path restated (all under /root/reference/lib/zstd): longtail_zstd.c:107-140 -> ZSTD_compressCCtx ->
 *   parameters   ext/compress/clevels.h:25-132 (level-3 rows), zstd_compress.c:1464-1602 (adjust), :7074-7097 (row pick)
 *   frame        zstd_compress.c:4575-4623 (header), :4493-4572 (block loop), :5231-5260 (epilogue)
 *   block        zstd_compress.c:4317-4384, :3202-3360 (seq store), :2996-3033 (compressibility gate)
 *   matcher      zstd_double_fast.c:105-311 (noDict), zstd_compress_internal.h:803-841 (hashes), :1305-1316 (window)
 *   literals     zstd_compress_literals.c:129-235, huf_compress.c (sort :620-665, tree :681-718, depth limit :376-505,
 *                codes :730-753, table description :248-290, weights FSE :127-182, streams :1064-1213, driver :1334-1431)
 *   sequences    zstd_compress.c:2681-2705 (codes), :2750-2870 (statistics), :2876-2990; zstd_compress_sequences.c:157-240
 *                (encoding type), :243-290 (tables), :293-385 (bitstream)
 *   FSE          fse_compress.c:68-209 (table), :238-330 (NCount), :357-369 (table log), :379-526 (normalisation), :560-610
 *
 * Only 'ztd1' (level 0 == default) and 'ztd2' (level 3) are covered: both are level 3 = double-fast strategy.
 * Output capacity is assumed to be at least ZSTD_COMPRESSBOUND (what compressblockstore passes), so the reference's
 * "destination too small" fallbacks never alter a decision (they only ever fire for output larger than the input,
 * which the compressibility gate rejects anyway).
 */
#include "lt_oracle.h"

#include <errno.h>
#include <stdlib.h>
#include <string.h>

#define ZS_BLOCK_MAX (128u << 10)
#define ZS_ERR ((size_t)-1)

static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t hibit(uint32_t v) { return 31u - (uint32_t)__builtin_clz(v); }

uint64_t lto_zstd_bound(uint64_t n) /* zstd.h:232 */
{
    return n + (n >> 8) + (n < (128u << 10) ? ((128u << 10) - n) >> 11 : 0);
}

/* ------------------------------------------------------------------ forward bit writer (bitstream.h:150-241)
 * bits are appended LSB first; a stream is closed with a single 1 bit and padded to a byte. */
typedef struct
{
    uint8_t* out;
    size_t pos;
    uint64_t acc;
    uint32_t nbits;
} bitw;

static void bw_init(bitw* w, uint8_t* out) { w->out = out; w->pos = 0; w->acc = 0; w->nbits = 0; }
static void bw_add(bitw* w, uint64_t value, uint32_t n)
{
    if (!n) return;
    value &= (n >= 64) ? ~0ull : ((1ull << n) - 1);
    w->acc |= value << w->nbits;
    w->nbits += n;
    while (w->nbits >= 8)
    {
        w->out[w->pos++] = (uint8_t)w->acc;
        w->acc >>= 8;
        w->nbits -= 8;
    }
}
static size_t bw_close(bitw* w)
{
    bw_add(w, 1, 1);
    if (w->nbits) { w->out[w->pos++] = (uint8_t)w->acc; w->nbits = 0; }
    return w->pos;
}

/* ================================================================== FSE (fse_compress.c) */

#define FSE_MIN_LOG 5
#define FSE_MAX_LOG 12
#define FSE_MAX_SYMS 256

typedef struct
{
    uint32_t table_log, max_symbol;
    uint16_t next_state[1u << 9]; /* sorted by symbol; sequences use <= 9 bits, Huffman weights 6 */
    uint32_t delta_nb_bits[64];
    int32_t delta_find_state[64];
} fse_ctable;

/* fse_compress.c:343-369 */
static uint32_t fse_min_table_log(size_t n, uint32_t max_symbol)
{
    uint32_t by_src = hibit((uint32_t)n) + 1, by_sym = hibit(max_symbol) + 2;
    return by_src < by_sym ? by_src : by_sym;
}
static uint32_t fse_optimal_table_log(uint32_t max_log, size_t n, uint32_t max_symbol, uint32_t minus)
{
    uint32_t by_src = hibit((uint32_t)(n - 1)) - minus;
    uint32_t log = max_log;
    uint32_t min_bits = fse_min_table_log(n, max_symbol);
    if (by_src < log) log = by_src;
    if (min_bits > log) log = min_bits;
    if (log < FSE_MIN_LOG) log = FSE_MIN_LOG;
    if (log > FSE_MAX_LOG) log = FSE_MAX_LOG;
    return log;
}

/* fse_compress.c:379-463, the fallback normalisation */
static int fse_normalize_m2(int16_t* norm, uint32_t log, const uint32_t* count, size_t total, uint32_t max_symbol, int16_t low_prob)
{
    const int16_t UNSET = -2;
    uint32_t distributed = 0, to_give;
    uint32_t low_threshold = (uint32_t)(total >> log);
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
        uint64_t v_log = 62 - log;
        uint64_t mid = (1ull << (v_log - 1)) - 1;
        uint64_t r_step = (((uint64_t)1 << v_log) * to_give + mid) / (uint32_t)total;
        uint64_t run = mid;
        for (uint32_t s = 0; s <= max_symbol; ++s)
        {
            if (norm[s] != UNSET) continue;
            uint64_t end = run + count[s] * r_step;
            uint32_t w = (uint32_t)(end >> v_log) - (uint32_t)(run >> v_log);
            if (w < 1) return -1;
            norm[s] = (int16_t)w;
            run = end;
        }
    }
    return 0;
}

/* fse_compress.c:465-526; returns the table log, 0 for the single-symbol case, -1 on error */
static int fse_normalize(int16_t* norm, uint32_t log, const uint32_t* count, size_t total, uint32_t max_symbol, int use_low_prob)
{
    static const uint32_t round_up_threshold[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
    if (log < FSE_MIN_LOG || log > FSE_MAX_LOG) return -1;
    if (log < fse_min_table_log(total, max_symbol)) return -1;
    int16_t low_prob = use_low_prob ? -1 : 1;
    uint64_t scale = 62 - log;
    uint64_t step = ((uint64_t)1 << 62) / (uint32_t)total;
    uint64_t v_step = 1ull << (scale - 20);
    int left = 1 << log;
    uint32_t largest = 0;
    int16_t largest_p = 0;
    uint32_t low_threshold = (uint32_t)(total >> log);
    for (uint32_t s = 0; s <= max_symbol; ++s)
    {
        if (count[s] == total) return 0;
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= low_threshold) { norm[s] = low_prob; left--; continue; }
        int16_t p = (int16_t)((count[s] * step) >> scale);
        if (p < 8)
        {
            uint64_t rest_to_beat = v_step * round_up_threshold[p];
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

/* fse_compress.c:238-330: the normalised counts as the decoder reads them */
static size_t fse_write_ncount(uint8_t* out, const int16_t* norm, uint32_t max_symbol, uint32_t log)
{
    const int table_size = 1 << log;
    int remaining = table_size + 1, threshold = table_size, nb_bits = (int)log + 1;
    uint32_t bits = log - FSE_MIN_LOG;
    int bit_count = 4;
    uint32_t symbol = 0, alphabet = max_symbol + 1;
    int previous_is_0 = 0;
    size_t pos = 0;
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
            int max = (2 * threshold - 1) - remaining;
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
    pos += (size_t)((bit_count + 7) / 8);
    return pos;
}

/* fse_compress.c:68-209 */
static void fse_build_ctable(fse_ctable* ct, const int16_t* norm, uint32_t max_symbol, uint32_t log)
{
    const uint32_t size = 1u << log, mask = size - 1;
    const uint32_t step = (size >> 1) + (size >> 3) + 3;
    uint16_t cumul[FSE_MAX_SYMS + 2];
    uint8_t spread[1u << 9];
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
    {
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
                uint32_t max_bits_out = log - hibit((uint32_t)norm[s] - 1);
                uint32_t min_state_plus = (uint32_t)norm[s] << max_bits_out;
                ct->delta_nb_bits[s] = (max_bits_out << 16) - min_state_plus;
                ct->delta_find_state[s] = (int32_t)(total - (uint32_t)norm[s]);
                total += (uint32_t)norm[s];
            }
        }
    }
}
/* fse_compress.c:532-552: one symbol, zero bits per symbol */
static void fse_build_ctable_rle(fse_ctable* ct, uint32_t symbol)
{
    ct->table_log = 0;
    ct->max_symbol = symbol;
    ct->next_state[0] = ct->next_state[1] = 0;
    ct->delta_nb_bits[symbol] = 0;
    ct->delta_find_state[symbol] = 0;
}

/* fse.h:437-476 */
typedef struct { const fse_ctable* ct; uint32_t value; } fse_state;
static void fse_init_state(fse_state* st, const fse_ctable* ct, uint32_t symbol)
{
    uint32_t nb = (ct->delta_nb_bits[symbol] + (1u << 15)) >> 16;
    uint32_t v = (nb << 16) - ct->delta_nb_bits[symbol];
    st->ct = ct;
    st->value = ct->next_state[(int32_t)(v >> nb) + ct->delta_find_state[symbol]];
}
static void fse_encode(bitw* w, fse_state* st, uint32_t symbol)
{
    uint32_t nb = (st->value + st->ct->delta_nb_bits[symbol]) >> 16;
    bw_add(w, st->value, nb);
    st->value = st->ct->next_state[(int32_t)(st->value >> nb) + st->ct->delta_find_state[symbol]];
}
static void fse_flush_state(bitw* w, const fse_state* st) { bw_add(w, st->value, st->ct->table_log); }

/* ================================================================== Huffman literals (huf_compress.c) */

typedef struct
{
    uint32_t table_log, max_symbol; /* the CTable header (huf_compress.c:219-241) */
    uint8_t nb_bits[256];
    uint16_t code[256];
    int repeat; /* HUF_repeat: 0 none, 1 check, 2 valid */
} huf_table;

typedef struct { uint32_t count; uint16_t parent; uint8_t byte, nb_bits; } huf_node;

static uint32_t hist(uint32_t* count, uint32_t* max_symbol, const uint8_t* src, size_t n) /* hist.c:29-56 */
{
    uint32_t m = *max_symbol, largest = 0;
    memset(count, 0, (m + 1) * sizeof(uint32_t));
    if (!n) { *max_symbol = 0; return 0; }
    for (size_t i = 0; i < n; ++i) count[src[i]]++;
    while (!count[m]) m--;
    *max_symbol = m;
    for (uint32_t s = 0; s <= m; ++s) if (count[s] > largest) largest = count[s];
    return largest;
}

/* the Huffman table description's weights, FSE-compressed (huf_compress.c:127-182); 0 / 1 = not worth it */
static size_t huf_compress_weights(uint8_t* dst, const uint8_t* weights, size_t n)
{
    uint32_t count[13], max_symbol = 12;
    int16_t norm[13];
    fse_ctable ct;
    if (n <= 1) return 0;
    {
        uint32_t most = hist(count, &max_symbol, weights, n);
        if (most == n) return 1;
        if (most == 1) return 0;
    }
    uint32_t log = fse_optimal_table_log(6, n, max_symbol, 2);
    int r = fse_normalize(norm, log, count, n, max_symbol, 0);
    if (r < 0) return ZS_ERR;
    size_t pos = fse_write_ncount(dst, norm, max_symbol, log);
    if (pos == ZS_ERR) return ZS_ERR;
    fse_build_ctable(&ct, norm, max_symbol, log);
    /* FSE_compress_usingCTable_generic (fse_compress.c:560-610): two interleaved states, input read backwards */
    if (n <= 2) return 0;
    {
        bitw w;
        fse_state s1, s2;
        const uint8_t* ip = weights + n;
        bw_init(&w, dst + pos);
        if (n & 1)
        {
            fse_init_state(&s1, &ct, *--ip);
            fse_init_state(&s2, &ct, *--ip);
            fse_encode(&w, &s1, *--ip);
        }
        else
        {
            fse_init_state(&s2, &ct, *--ip);
            fse_init_state(&s1, &ct, *--ip);
        }
        while (ip > weights)
        {
            fse_encode(&w, &s2, *--ip);
            fse_encode(&w, &s1, *--ip);
        }
        fse_flush_state(&w, &s2);
        fse_flush_state(&w, &s1);
        pos += bw_close(&w);
    }
    return pos;
}

/* huf_compress.c:248-290 */
static size_t huf_write_table(uint8_t* dst, const huf_table* t)
{
    uint8_t bits_to_weight[13], weights[256];
    uint32_t max_symbol = t->max_symbol;
    bits_to_weight[0] = 0;
    for (uint32_t n = 1; n < t->table_log + 1; ++n) bits_to_weight[n] = (uint8_t)(t->table_log + 1 - n);
    for (uint32_t n = 0; n < max_symbol; ++n) weights[n] = bits_to_weight[t->nb_bits[n]];
    {
        size_t h = huf_compress_weights(dst + 1, weights, max_symbol);
        if (h == ZS_ERR) return ZS_ERR;
        if (h > 1 && h < max_symbol / 2) { dst[0] = (uint8_t)h; return h + 1; }
    }
    if (max_symbol > 128) return ZS_ERR;
    dst[0] = (uint8_t)(128 + (max_symbol - 1));
    weights[max_symbol] = 0;
    for (uint32_t n = 0; n < max_symbol; n += 2) dst[n / 2 + 1] = (uint8_t)((weights[n] << 4) + weights[n + 1]);
    return (max_symbol + 1) / 2 + 1;
}

/* huf_compress.c:530-665: bucket sort by count, descending; large counts share log2 buckets sorted by an unstable quicksort
 * whose exact element moves decide the order of equal counts, hence are reproduced step by step */
static uint32_t huf_bucket(uint32_t count) { return count < 166 ? count : hibit(count) + 158; }
static void huf_swap(huf_node* a, huf_node* b) { huf_node t = *a; *a = *b; *b = t; }
static void huf_insertion_sort(huf_node* a, int low, int high)
{
    int size = high - low + 1;
    a += low;
    for (int i = 1; i < size; ++i)
    {
        huf_node key = a[i];
        int j = i - 1;
        while (j >= 0 && a[j].count < key.count) { a[j + 1] = a[j]; j--; }
        a[j + 1] = key;
    }
}
static int huf_partition(huf_node* a, int low, int high)
{
    uint32_t pivot = a[high].count;
    int i = low - 1;
    for (int j = low; j < high; ++j)
        if (a[j].count > pivot) { i++; huf_swap(&a[i], &a[j]); }
    huf_swap(&a[i + 1], &a[high]);
    return i + 1;
}
static void huf_quick_sort(huf_node* a, int low, int high)
{
    if (high - low < 8) { huf_insertion_sort(a, low, high); return; }
    while (low < high)
    {
        int idx = huf_partition(a, low, high);
        if (idx - low < high - idx) { huf_quick_sort(a, low, idx - 1); low = idx + 1; }
        else { huf_quick_sort(a, idx + 1, high); high = idx - 1; }
    }
}
static void huf_sort(huf_node* node, const uint32_t* count, uint32_t max_symbol)
{
    struct { uint16_t base, curr; } rank[192];
    memset(rank, 0, sizeof(rank));
    for (uint32_t n = 0; n <= max_symbol; ++n) rank[huf_bucket(count[n])].base++;
    for (uint32_t n = 191; n > 0; --n) { rank[n - 1].base = (uint16_t)(rank[n - 1].base + rank[n].base); rank[n - 1].curr = rank[n - 1].base; }
    for (uint32_t n = 0; n <= max_symbol; ++n)
    {
        uint32_t r = huf_bucket(count[n]) + 1;
        uint32_t pos = rank[r].curr++;
        node[pos].count = count[n];
        node[pos].byte = (uint8_t)n;
    }
    for (uint32_t n = 166; n < 191; ++n)
    {
        int size = rank[n].curr - rank[n].base;
        if (size > 1) huf_quick_sort(node + rank[n].base, 0, size - 1);
    }
}

/* huf_compress.c:376-505: enforce the maximum code length on the sorted leaves */
static uint32_t huf_set_max_height(huf_node* node, uint32_t last_non_null, uint32_t target)
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
    {
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
                uint32_t high_pos = rank_last[dec], low_pos = rank_last[dec - 1];
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
    }
    return target;
}

/* huf_compress.c:681-800: sorted leaves -> tree -> depth limit -> canonical codes; returns the longest code length */
static uint32_t huf_build_table(huf_table* t, const uint32_t* count, uint32_t max_symbol, uint32_t max_bits)
{
    huf_node storage[2 * 256 + 1];
    huf_node* const node = storage + 1; /* node[-1] is the sentinel */
    const int START = 256;
    memset(storage, 0, sizeof(storage));
    huf_sort(node, count, max_symbol);
    int non_null = (int)max_symbol;
    while (node[non_null].count == 0) non_null--;
    {
        int low_s = non_null, node_nb = START, node_root = node_nb + low_s - 1, low_n = node_nb;
        node[node_nb].count = node[low_s].count + node[low_s - 1].count;
        node[low_s].parent = node[low_s - 1].parent = (uint16_t)node_nb;
        node_nb++;
        low_s -= 2;
        for (int n = node_nb; n <= node_root; ++n) node[n].count = 1u << 30;
        node[-1].count = 1u << 31;
        while (node_nb <= node_root)
        {
            int n1 = (node[low_s].count < node[low_n].count) ? low_s-- : low_n++;
            int n2 = (node[low_s].count < node[low_n].count) ? low_s-- : low_n++;
            node[node_nb].count = node[n1].count + node[n2].count;
            node[n1].parent = node[n2].parent = (uint16_t)node_nb;
            node_nb++;
        }
        node[node_root].nb_bits = 0;
        for (int n = node_root - 1; n >= START; --n) node[n].nb_bits = (uint8_t)(node[node[n].parent].nb_bits + 1);
        for (int n = 0; n <= non_null; ++n) node[n].nb_bits = (uint8_t)(node[node[n].parent].nb_bits + 1);
    }
    max_bits = huf_set_max_height(node, (uint32_t)non_null, max_bits);
    {
        uint16_t per_rank[13] = {0}, val_per_rank[13] = {0};
        for (int n = 0; n <= non_null; ++n) per_rank[node[n].nb_bits]++;
        uint16_t min = 0;
        for (int n = (int)max_bits; n > 0; --n) { val_per_rank[n] = min; min = (uint16_t)(min + per_rank[n]); min >>= 1; }
        memset(t->nb_bits, 0, sizeof(t->nb_bits));
        memset(t->code, 0, sizeof(t->code));
        for (uint32_t n = 0; n <= max_symbol; ++n) t->nb_bits[node[n].byte] = node[n].nb_bits;
        for (uint32_t n = 0; n <= max_symbol; ++n) t->code[n] = t->nb_bits[n] ? val_per_rank[t->nb_bits[n]]++ : 0;
        t->table_log = max_bits;
        t->max_symbol = max_symbol;
    }
    return max_bits;
}

static size_t huf_estimate(const huf_table* t, const uint32_t* count, uint32_t max_symbol) /* huf_compress.c:802-811 */
{
    size_t bits = 0;
    for (uint32_t s = 0; s <= max_symbol; ++s) bits += (size_t)t->nb_bits[s] * count[s];
    return bits >> 3;
}
static int huf_validate(const huf_table* t, const uint32_t* count, uint32_t max_symbol) /* huf_compress.c:813-828 */
{
    int bad = 0;
    if (t->max_symbol < max_symbol) return 0;
    for (uint32_t s = 0; s <= max_symbol; ++s) bad |= (count[s] != 0) & (t->nb_bits[s] == 0);
    return !bad;
}

/* one Huffman stream: symbols last to first, codes LSB first, closed by a 1 bit (huf_compress.c:984-1110) */
static size_t huf_encode_1x(uint8_t* dst, const uint8_t* src, size_t n, const huf_table* t)
{
    bitw w;
    bw_init(&w, dst);
    for (size_t i = n; i-- > 0;) bw_add(&w, t->code[src[i]], t->nb_bits[src[i]]);
    return bw_close(&w);
}
/* four streams with a 6-byte jump table (huf_compress.c:1168-1213) */
static size_t huf_encode_4x(uint8_t* dst, const uint8_t* src, size_t n, const huf_table* t)
{
    size_t seg = (n + 3) / 4, pos = 6;
    if (n < 12) return 0;
    for (int i = 0; i < 4; ++i)
    {
        const uint8_t* p = src + (size_t)i * seg;
        size_t len = i < 3 ? seg : n - 3 * seg;
        size_t c = huf_encode_1x(dst + pos, p, len, t);
        if (c == 0 || c > 65535) return 0;
        if (i < 3) { dst[2 * i] = (uint8_t)c; dst[2 * i + 1] = (uint8_t)(c >> 8); }
        pos += c;
    }
    return pos;
}
/* huf_compress.c:1223-1239; `head` bytes of table description already sit in dst */
static size_t huf_encode_with(uint8_t* dst, size_t head, const uint8_t* src, size_t n, int four, const huf_table* t)
{
    size_t c = four ? huf_encode_4x(dst + head, src, n, t) : huf_encode_1x(dst + head, src, n, t);
    if (c == 0) return 0;
    if (head + c >= n - 1) return 0;
    return head + c;
}

/* HUF_compress_internal (huf_compress.c:1334-1431).  `t` enters as the previous block's table and leaves as the table the
 * next block may reuse; *repeat enters as that table's status.  Returns 0 = not compressible, 1 = single symbol, ZS_ERR. */
static size_t huf_compress(uint8_t* dst, const uint8_t* src, size_t n, int four, huf_table* t, int* repeat, int prefer_repeat, int suspect)
{
    uint32_t count[256], max_symbol = 255;
    huf_table fresh;
    if (!n) return 0;
    if (prefer_repeat && *repeat == 2) return huf_encode_with(dst, 0, src, n, four, t);
    if (suspect && n >= 4096 * 10)
    {
        uint32_t m = 255, total;
        total = hist(count, &m, src, 4096);
        m = 255;
        total += hist(count, &m, src + n - 4096, 4096);
        if (total <= ((2 * 4096) >> 7) + 4) return 0;
    }
    {
        uint32_t largest = hist(count, &max_symbol, src, n);
        if (largest == n) { dst[0] = src[0]; return 1; }
        if (largest <= (n >> 7) + 4) return 0;
    }
    if (*repeat == 1 && !huf_validate(t, count, max_symbol)) *repeat = 0;
    if (prefer_repeat && *repeat != 0) return huf_encode_with(dst, 0, src, n, four, t);
    {
        uint32_t log = fse_optimal_table_log(11, n, max_symbol, 1); /* HUF_optimalTableLog without the depth search */
        huf_build_table(&fresh, count, max_symbol, log);
    }
    {
        size_t h = huf_write_table(dst, &fresh);
        if (h == ZS_ERR) return ZS_ERR;
        if (*repeat != 0)
        {
            size_t old_size = huf_estimate(t, count, max_symbol), new_size = huf_estimate(&fresh, count, max_symbol);
            if (old_size <= h + new_size || h + 12 >= n) return huf_encode_with(dst, 0, src, n, four, t);
        }
        if (h + 12 >= n) return 0;
        *repeat = 0;
        fresh.repeat = t->repeat;
        *t = fresh;
        return huf_encode_with(dst, h, src, n, four, t);
    }
}

static size_t lit_raw(uint8_t* dst, const uint8_t* src, size_t n) /* zstd_compress_literals.c:39-63 */
{
    uint32_t fl = 1 + (n > 31) + (n > 4095);
    if (fl == 1) dst[0] = (uint8_t)(0 + (n << 3));
    else if (fl == 2) { uint16_t v = (uint16_t)(0 + (1 << 2) + (n << 4)); memcpy(dst, &v, 2); }
    else { uint32_t v = (uint32_t)(0 + (3 << 2) + (n << 4)); memcpy(dst, &v, 4); }
    memcpy(dst + fl, src, n);
    return n + fl;
}
static size_t lit_rle(uint8_t* dst, const uint8_t* src, size_t n) /* zstd_compress_literals.c:78-104 */
{
    uint32_t fl = 1 + (n > 31) + (n > 4095);
    if (fl == 1) dst[0] = (uint8_t)(1 + (n << 3));
    else if (fl == 2) { uint16_t v = (uint16_t)(1 + (1 << 2) + (n << 4)); memcpy(dst, &v, 2); }
    else { uint32_t v = (uint32_t)(1 + (3 << 2) + (n << 4)); memcpy(dst, &v, 4); }
    dst[fl] = src[0];
    return fl + 1;
}

/* ZSTD_compressLiterals at strategy dfast (zstd_compress_literals.c:129-235) */
static size_t compress_literals(uint8_t* dst, const uint8_t* src, size_t n, const huf_table* prev, huf_table* next, int suspect)
{
    const size_t lh = 3 + (n >= 1024) + (n >= 16384);
    int single = n < 256;
    int type = 2; /* set_compressed */
    *next = *prev;
    if (n < (prev->repeat == 2 ? 6u : 64u)) return lit_raw(dst, src, n);
    int repeat = prev->repeat;
    if (repeat == 2 && lh == 3) single = 1;
    size_t c = huf_compress(dst + lh, src, n, !single, next, &repeat, n <= 1024, suspect);
    if (repeat != 0) type = 3; /* set_repeat */
    {
        size_t min_gain = (n >> 6) + 2;
        if (c == 0 || c == ZS_ERR || c >= n - min_gain) { *next = *prev; return lit_raw(dst, src, n); }
    }
    if (c == 1)
    {
        int same = 1;
        if (n < 8) for (size_t i = 1; i < n; ++i) if (src[i] != src[0]) same = 0;
        if (n >= 8 || same) { *next = *prev; return lit_rle(dst, src, n); }
    }
    if (type == 2) next->repeat = 1; /* HUF_repeat_check */
    if (lh == 3) { uint32_t v = (uint32_t)type + ((uint32_t)(!single) << 2) + ((uint32_t)n << 4) + ((uint32_t)c << 14); dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); }
    else if (lh == 4) { uint32_t v = (uint32_t)type + (2u << 2) + ((uint32_t)n << 4) + ((uint32_t)c << 18); memcpy(dst, &v, 4); }
    else { uint32_t v = (uint32_t)type + (3u << 2) + ((uint32_t)n << 4) + ((uint32_t)c << 22); memcpy(dst, &v, 4); dst[4] = (uint8_t)(c >> 10); }
    return lh + c;
}

/* ================================================================== sequences */

typedef struct { uint32_t lit_len, match_len, off_base; } zs_seq;

/* extra bits per code as RFC 8878 3.1.1.3.2.1.1 tabulates them (zstd_internal.h:118-146); the code of a value is the last
 * code whose baseline does not exceed it (zstd_compress_internal.h:517-546) */
static const uint8_t LL_BITS[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const uint8_t ML_BITS[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                    1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static const int16_t LL_DEFAULT[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static const int16_t ML_DEFAULT[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
static const int16_t OF_DEFAULT[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};

static uint32_t code_of(uint32_t v, const uint8_t* bits, uint32_t codes)
{
    uint32_t base = 0;
    for (uint32_t c = 0; c + 1 < codes; ++c)
    {
        uint32_t next = base + (1u << bits[c]);
        if (v < next) return c;
        base = next;
    }
    return codes - 1;
}

/* ZSTD_selectEncodingType for strategy < lazy (zstd_compress_sequences.c:157-206); 0 basic, 1 rle, 2 compressed.
 * Without a dictionary the repeat mode never becomes "valid", so set_repeat is unreachable here. */
static int select_encoding(uint32_t most, size_t nb_seq, uint32_t default_log, int default_allowed)
{
    if (most == nb_seq) return (default_allowed && nb_seq <= 2) ? 0 : 1;
    if (default_allowed)
    {
        size_t dynamic_min = (((size_t)1 << default_log) * (10 - 2)) >> 3;
        if (nb_seq < dynamic_min || most < (nb_seq >> (default_log - 1))) return 0;
    }
    return 2;
}

/* ZSTD_buildCTable (zstd_compress_sequences.c:243-290): returns bytes of table description written */
static size_t build_seq_table(uint8_t* dst, fse_ctable* ct, uint32_t max_log, int type, uint32_t* count, uint32_t max, const uint8_t* codes,
                              size_t nb_seq, const int16_t* default_norm, uint32_t default_log, uint32_t default_max)
{
    if (type == 1) { fse_build_ctable_rle(ct, max); dst[0] = codes[0]; return 1; }
    if (type == 0) { fse_build_ctable(ct, default_norm, default_max, default_log); return 0; }
    {
        int16_t norm[64];
        size_t n1 = nb_seq;
        uint32_t log = fse_optimal_table_log(max_log, nb_seq, max, 2);
        if (count[codes[nb_seq - 1]] > 1) { count[codes[nb_seq - 1]]--; n1--; }
        if (fse_normalize(norm, log, count, n1, max, n1 >= 2048) < 0) return ZS_ERR;
        size_t h = fse_write_ncount(dst, norm, max, log);
        if (h == ZS_ERR) return ZS_ERR;
        fse_build_ctable(ct, norm, max, log);
        return h;
    }
}

/* literals + sequences of one block (zstd_compress.c:2876-2990); 0 = emit the block raw, ZS_ERR on error */
static size_t entropy_compress(uint8_t* dst, const uint8_t* lits, size_t lit_size, const zs_seq* seqs, size_t nb_seq, const huf_table* prev_huf,
                               huf_table* next_huf, uint8_t* code_buf)
{
    uint8_t* op = dst;
    {
        int suspect = (nb_seq == 0) || (lit_size / nb_seq >= 20);
        op += compress_literals(op, lits, lit_size, prev_huf, next_huf, suspect);
    }
    if (nb_seq < 128) *op++ = (uint8_t)nb_seq;
    else if (nb_seq < 0x7F00) { op[0] = (uint8_t)((nb_seq >> 8) + 0x80); op[1] = (uint8_t)nb_seq; op += 2; }
    else { op[0] = 0xFF; uint16_t v = (uint16_t)(nb_seq - 0x7F00); memcpy(op + 1, &v, 2); op += 3; }
    if (nb_seq == 0) return (size_t)(op - dst);

    uint8_t* ll_code = code_buf;
    uint8_t* of_code = code_buf + nb_seq;
    uint8_t* ml_code = code_buf + 2 * nb_seq;
    for (size_t i = 0; i < nb_seq; ++i)
    {
        ll_code[i] = (uint8_t)code_of(seqs[i].lit_len, LL_BITS, 36);
        of_code[i] = (uint8_t)hibit(seqs[i].off_base);
        ml_code[i] = (uint8_t)code_of(seqs[i].match_len - 3, ML_BITS, 53);
    }
    uint8_t* seq_head = op++;
    fse_ctable ct_ll, ct_of, ct_ml;
    uint32_t count[64];
    size_t last_count_size = 0;
    int ll_type, of_type, ml_type;
    {
        uint32_t max = 35, most = hist(count, &max, ll_code, nb_seq);
        ll_type = select_encoding(most, nb_seq, 6, 1);
        size_t h = build_seq_table(op, &ct_ll, 9, ll_type, count, max, ll_code, nb_seq, LL_DEFAULT, 6, 35);
        if (h == ZS_ERR) return ZS_ERR;
        if (ll_type == 2) last_count_size = h;
        op += h;
    }
    {
        uint32_t max = 31, most = hist(count, &max, of_code, nb_seq);
        of_type = select_encoding(most, nb_seq, 5, max <= 28);
        size_t h = build_seq_table(op, &ct_of, 8, of_type, count, max, of_code, nb_seq, OF_DEFAULT, 5, 28);
        if (h == ZS_ERR) return ZS_ERR;
        if (of_type == 2) last_count_size = h;
        op += h;
    }
    {
        uint32_t max = 52, most = hist(count, &max, ml_code, nb_seq);
        ml_type = select_encoding(most, nb_seq, 6, 1);
        size_t h = build_seq_table(op, &ct_ml, 9, ml_type, count, max, ml_code, nb_seq, ML_DEFAULT, 6, 52);
        if (h == ZS_ERR) return ZS_ERR;
        if (ml_type == 2) last_count_size = h;
        op += h;
    }
    *seq_head = (uint8_t)((ll_type << 6) + (of_type << 4) + (ml_type << 2));
    {
        /* ZSTD_encodeSequences (zstd_compress_sequences.c:293-385): last sequence first */
        bitw w;
        fse_state st_ml, st_of, st_ll;
        bw_init(&w, op);
        size_t n = nb_seq - 1;
        fse_init_state(&st_ml, &ct_ml, ml_code[n]);
        fse_init_state(&st_of, &ct_of, of_code[n]);
        fse_init_state(&st_ll, &ct_ll, ll_code[n]);
        bw_add(&w, seqs[n].lit_len, LL_BITS[ll_code[n]]);
        bw_add(&w, seqs[n].match_len - 3, ML_BITS[ml_code[n]]);
        bw_add(&w, seqs[n].off_base, of_code[n]);
        while (n-- > 0)
        {
            fse_encode(&w, &st_of, of_code[n]);
            fse_encode(&w, &st_ml, ml_code[n]);
            fse_encode(&w, &st_ll, ll_code[n]);
            bw_add(&w, seqs[n].lit_len, LL_BITS[ll_code[n]]);
            bw_add(&w, seqs[n].match_len - 3, ML_BITS[ml_code[n]]);
            bw_add(&w, seqs[n].off_base, of_code[n]);
        }
        fse_flush_state(&w, &st_ml);
        fse_flush_state(&w, &st_of);
        fse_flush_state(&w, &st_ll);
        size_t stream = bw_close(&w);
        op += stream;
        if (last_count_size && last_count_size + stream < 4) return 0; /* zstd <= 1.3.4 decoder quirk, :2982-2988 */
    }
    return (size_t)(op - dst);
}

/* ================================================================== double-fast matcher */

typedef struct
{
    uint32_t window_log, chain_log, hash_log, min_match;
} zs_params;

/* level 3 rows of clevels.h:25-132 and ZSTD_adjustCParams_internal (zstd_compress.c:1464-1602) for a known size, no dictionary */
static zs_params zs_get_params(uint64_t n)
{
    static const zs_params rows[4] = {{21, 16, 17, 5}, {18, 16, 16, 4}, {17, 15, 16, 5}, {14, 14, 15, 4}};
    zs_params p = rows[(n <= (256u << 10)) + (n <= (128u << 10)) + (n <= (16u << 10))];
    if (n <= (1ull << 30))
    {
        uint32_t t = (uint32_t)n;
        uint32_t src_log = t < 64 ? 6 : hibit(t - 1) + 1;
        if (p.window_log > src_log) p.window_log = src_log;
    }
    if (p.hash_log > p.window_log + 1) p.hash_log = p.window_log + 1;
    if (p.chain_log > p.window_log) p.chain_log = p.window_log;
    if (p.window_log < 10) p.window_log = 10;
    return p;
}

static inline uint32_t hash_long(const uint8_t* p, uint32_t bits) { return (uint32_t)((rd64(p) * 0xCF1BBCDCB7A56463ull) >> (64 - bits)); }
static inline uint32_t hash_short(const uint8_t* p, uint32_t bits, uint32_t mls)
{
    if (mls == 5) return (uint32_t)(((rd64(p) << 24) * 889523592379ull) >> (64 - bits));
    return (rd32(p) * 2654435761u) >> (32 - bits);
}
static size_t count_equal(const uint8_t* a, const uint8_t* b, const uint8_t* a_end) /* ZSTD_count */
{
    const uint8_t* s = a;
    while (a < a_end && *a == *b) { a++; b++; }
    return (size_t)(a - s);
}

typedef struct
{
    zs_params p;
    uint32_t* hash_long_tab;
    uint32_t* hash_small_tab;
    uint32_t dict_limit; /* == low limit: no external dictionary ever exists on this path */
    uint32_t rep[3];     /* confirmed repcodes of the previous compressed block */
    huf_table huf;       /* confirmed literal table */
} zs_frame;

/* ZSTD_compressBlock_doubleFast_noDict_generic (zstd_double_fast.c:105-311).  `base` points at frame byte 0 whose match
 * index is 2 (ZSTD_WINDOW_START_INDEX).  Writes sequences, returns their count; *last_lits = literals after the last match;
 * rep[0..1] are updated as the reference's `rep` argument. */
static size_t dfast_block(zs_frame* f, const uint8_t* base0, size_t block_start, size_t block_size, zs_seq* seqs, uint32_t rep[3], size_t* last_lits)
{
    const uint8_t* const base = base0 - 2; /* index space */
    const uint32_t hl_bits = f->p.hash_log, hs_bits = f->p.chain_log, mls = f->p.min_match;
    uint32_t* const hash_l = f->hash_long_tab;
    uint32_t* const hash_s = f->hash_small_tab;
    const uint8_t* const istart = base0 + block_start;
    const uint8_t* anchor = istart;
    const uint32_t max_dist = 1u << f->p.window_log;
    const uint32_t end_index = (uint32_t)(block_start + 2 + block_size);
    const uint32_t prefix_lowest_index = (end_index - f->dict_limit > max_dist) ? end_index - max_dist : f->dict_limit;
    const uint8_t* const prefix_lowest = base + prefix_lowest_index;
    const uint8_t* const iend = istart + block_size;
    const uint8_t* const ilimit = iend - 8;
    uint32_t offset_1 = rep[0], offset_2 = rep[1], saved_1 = 0, saved_2 = 0;
    size_t nb = 0;
    const uint8_t* ip = istart;

    ip += (ip == prefix_lowest);
    {
        uint32_t current = (uint32_t)(ip - base);
        uint32_t window_low = (current - f->dict_limit > max_dist) ? current - max_dist : f->dict_limit;
        uint32_t max_rep = current - window_low;
        if (offset_2 > max_rep) { saved_2 = offset_2; offset_2 = 0; }
        if (offset_1 > max_rep) { saved_1 = offset_1; offset_1 = 0; }
    }
    for (;;)
    {
        size_t step = 1;
        const uint8_t* next_step = ip + 256;
        const uint8_t* ip1 = ip + 1;
        size_t m_len = 0;
        uint32_t offset = 0, curr = 0, hl1 = 0;
        int found = 0; /* 1 = repcode stored, 2 = match to store */
        if (ip1 > ilimit) break;
        uint32_t hl0 = hash_long(ip, hl_bits);
        uint32_t idxl0 = hash_l[hl0];
        do
        {
            const uint32_t hs0 = hash_short(ip, hs_bits, mls);
            const uint32_t idxs0 = hash_s[hs0];
            const uint8_t* match_l0 = base + idxl0;
            const uint8_t* match_s0 = base + idxs0;
            curr = (uint32_t)(ip - base);
            hash_l[hl0] = hash_s[hs0] = curr;
            if (offset_1 > 0 && rd32(ip + 1 - offset_1) == rd32(ip + 1))
            {
                m_len = count_equal(ip + 1 + 4, ip + 1 + 4 - offset_1, iend) + 4;
                ip++;
                seqs[nb].lit_len = (uint32_t)(ip - anchor); seqs[nb].match_len = (uint32_t)m_len; seqs[nb].off_base = 1; nb++;
                found = 1;
                break;
            }
            hl1 = hash_long(ip1, hl_bits);
            if (idxl0 > prefix_lowest_index && rd64(match_l0) == rd64(ip))
            {
                m_len = count_equal(ip + 8, match_l0 + 8, iend) + 8;
                offset = (uint32_t)(ip - match_l0);
                while (ip > anchor && match_l0 > prefix_lowest && ip[-1] == match_l0[-1]) { ip--; match_l0--; m_len++; }
                found = 2;
                break;
            }
            const uint32_t idxl1 = hash_l[hl1];
            if (idxs0 > prefix_lowest_index && rd32(match_s0) == rd32(ip))
            {
                const uint8_t* match_l1 = base + idxl1;
                if (idxl1 > prefix_lowest_index && rd64(match_l1) == rd64(ip1))
                {
                    ip = ip1;
                    m_len = count_equal(ip + 8, match_l1 + 8, iend) + 8;
                    offset = (uint32_t)(ip - match_l1);
                    while (ip > anchor && match_l1 > prefix_lowest && ip[-1] == match_l1[-1]) { ip--; match_l1--; m_len++; }
                }
                else
                {
                    m_len = count_equal(ip + 4, match_s0 + 4, iend) + 4;
                    offset = (uint32_t)(ip - match_s0);
                    while (ip > anchor && match_s0 > prefix_lowest && ip[-1] == match_s0[-1]) { ip--; match_s0--; m_len++; }
                }
                found = 2;
                break;
            }
            if (ip1 >= next_step) { step++; next_step += 256; }
            ip = ip1;
            ip1 += step;
            hl0 = hl1;
            idxl0 = idxl1;
        } while (ip1 <= ilimit);
        if (!found) break;
        if (found == 2)
        {
            offset_2 = offset_1;
            offset_1 = offset;
            if (step < 4) hash_l[hl1] = (uint32_t)(ip1 - base);
            seqs[nb].lit_len = (uint32_t)(ip - anchor); seqs[nb].match_len = (uint32_t)m_len; seqs[nb].off_base = offset + 3; nb++;
        }
        ip += m_len;
        anchor = ip;
        if (ip <= ilimit)
        {
            const uint32_t insert = curr + 2;
            hash_l[hash_long(base + insert, hl_bits)] = insert;
            hash_l[hash_long(ip - 2, hl_bits)] = (uint32_t)(ip - 2 - base);
            hash_s[hash_short(base + insert, hs_bits, mls)] = insert;
            hash_s[hash_short(ip - 1, hs_bits, mls)] = (uint32_t)(ip - 1 - base);
            while (ip <= ilimit && offset_2 > 0 && rd32(ip) == rd32(ip - offset_2))
            {
                size_t r_len = count_equal(ip + 4, ip + 4 - offset_2, iend) + 4;
                uint32_t t = offset_2; offset_2 = offset_1; offset_1 = t;
                hash_s[hash_short(ip, hs_bits, mls)] = (uint32_t)(ip - base);
                hash_l[hash_long(ip, hl_bits)] = (uint32_t)(ip - base);
                seqs[nb].lit_len = 0; seqs[nb].match_len = (uint32_t)r_len; seqs[nb].off_base = 1; nb++;
                ip += r_len;
                anchor = ip;
            }
        }
    }
    saved_2 = (saved_1 != 0 && offset_1 != 0) ? saved_1 : saved_2;
    rep[0] = offset_1 ? offset_1 : saved_1;
    rep[1] = offset_2 ? offset_2 : saved_2;
    *last_lits = (size_t)(iend - anchor);
    return nb;
}

/* ================================================================== frame */

int lto_zstd_compress(const uint8_t* src, uint64_t size, uint8_t* dst, uint64_t cap, uint64_t* out_size)
{
    if (size >= (1ull << 31)) return EINVAL; /* index overflow correction (zstd_compress.c:4437-4484) is not restated */
    if (cap < lto_zstd_bound(size)) return EINVAL;
    zs_frame f;
    memset(&f, 0, sizeof(f));
    f.p = zs_get_params(size);
    f.hash_long_tab = (uint32_t*)calloc((size_t)1 << f.p.hash_log, 4);
    f.hash_small_tab = (uint32_t*)calloc((size_t)1 << f.p.chain_log, 4);
    const size_t window_size = size ? ((1ull << f.p.window_log) < size ? (size_t)(1ull << f.p.window_log) : (size_t)size) : 1;
    const size_t block_max = window_size < ZS_BLOCK_MAX ? window_size : ZS_BLOCK_MAX;
    zs_seq* seqs = (zs_seq*)malloc(sizeof(zs_seq) * (block_max / 4 + 2));
    uint8_t* lits = (uint8_t*)malloc(block_max + 16);
    uint8_t* codes = (uint8_t*)malloc(3 * (block_max / 4 + 2));
    uint8_t* scratch = (uint8_t*)malloc(3 * block_max + 4096); /* a block that expands is discarded, so give it room */
    if (!f.hash_long_tab || !f.hash_small_tab || !seqs || !lits || !codes || !scratch) return ENOMEM;
    f.dict_limit = 2;
    f.rep[0] = 1; f.rep[1] = 4; f.rep[2] = 8;

    uint8_t* op = dst;
    {
        /* ZSTD_writeFrameHeader (zstd_compress.c:4575-4623): content size on, no checksum, no dictionary id */
        uint32_t single = (1ull << f.p.window_log) >= size;
        uint32_t fcs = (size >= 256) + (size >= 65536 + 256) + (size >= 0xFFFFFFFFull);
        uint32_t magic = 0xFD2FB528u;
        memcpy(op, &magic, 4);
        op += 4;
        *op++ = (uint8_t)((single << 5) + (fcs << 6));
        if (!single) *op++ = (uint8_t)((f.p.window_log - 10) << 3);
        if (fcs == 0) { if (single) *op++ = (uint8_t)size; }
        else if (fcs == 1) { uint16_t v = (uint16_t)(size - 256); memcpy(op, &v, 2); op += 2; }
        else if (fcs == 2) { uint32_t v = (uint32_t)size; memcpy(op, &v, 4); op += 4; }
        else { memcpy(op, &size, 8); op += 8; }
    }
    if (size == 0)
    {
        op[0] = 1; op[1] = 0; op[2] = 0; /* ZSTD_writeEpilogue: one empty raw last block (zstd_compress.c:5244-5252) */
        op += 3;
    }
    const uint32_t max_dist = 1u << f.p.window_log;
    int first_block = 1;
    size_t pos = 0;
    while (pos < size)
    {
        const size_t bs = (size - pos) < block_max ? (size_t)(size - pos) : block_max;
        const uint32_t last = (pos + bs == size);
        /* ZSTD_window_enforceMaxDist, called with the block START (zstd_compress.c:4525) */
        if ((uint32_t)(pos + 2) > max_dist)
        {
            uint32_t low = (uint32_t)(pos + 2) - max_dist;
            if (f.dict_limit < low) f.dict_limit = low;
        }
        size_t c = 0;
        huf_table next_huf = f.huf;
        uint32_t next_rep[3] = {f.rep[0], f.rep[1], f.rep[2]};
        if (bs >= 7) /* MIN_CBLOCK_SIZE + block header + 2 (zstd_compress.c:3212) */
        {
            size_t last_lits = 0;
            size_t nb = dfast_block(&f, src, pos, bs, seqs, next_rep, &last_lits);
            size_t lit_size = 0;
            {
                const uint8_t* p = src + pos;
                for (size_t i = 0; i < nb; ++i)
                {
                    memcpy(lits + lit_size, p, seqs[i].lit_len);
                    lit_size += seqs[i].lit_len;
                    p += seqs[i].lit_len + seqs[i].match_len;
                }
                memcpy(lits + lit_size, p, last_lits);
                lit_size += last_lits;
            }
            c = entropy_compress(scratch, lits, lit_size, seqs, nb, &f.huf, &next_huf, codes);
            if (c == ZS_ERR) { free(f.hash_long_tab); free(f.hash_small_tab); free(seqs); free(lits); free(codes); free(scratch); return EINVAL; }
            if (c && c >= bs - ((bs >> 6) + 2)) c = 0; /* ZSTD_minGain, zstd_compress.c:3021-3024 */
            if (!first_block && c < 25)
            {
                int rle = 1; /* ZSTD_isRLE, zstd_compress.c:4359-4370 */
                for (size_t i = 1; i < bs; ++i) if (src[pos + i] != src[pos]) { rle = 0; break; }
                if (rle) { c = 1; scratch[0] = src[pos]; }
            }
            if (c > 1) /* ZSTD_blockState_confirmRepcodesAndEntropyTables */
            {
                f.huf = next_huf;
                f.rep[0] = next_rep[0]; f.rep[1] = next_rep[1]; f.rep[2] = next_rep[2];
            }
        }
        if (c == 0)
        {
            uint32_t h = last + (0u << 1) + (uint32_t)(bs << 3);
            op[0] = (uint8_t)h; op[1] = (uint8_t)(h >> 8); op[2] = (uint8_t)(h >> 16);
            memcpy(op + 3, src + pos, bs);
            op += 3 + bs;
        }
        else
        {
            uint32_t h = c == 1 ? last + (1u << 1) + (uint32_t)(bs << 3) : last + (2u << 1) + (uint32_t)(c << 3);
            op[0] = (uint8_t)h; op[1] = (uint8_t)(h >> 8); op[2] = (uint8_t)(h >> 16);
            memcpy(op + 3, scratch, c);
            op += 3 + c;
        }
        pos += bs;
        first_block = 0;
    }
    *out_size = (uint64_t)(op - dst);
    free(f.hash_long_tab); free(f.hash_small_tab); free(seqs); free(lits); free(codes); free(scratch);
    return 0;
}
