// zstd_dec.cu — ZStd frame decoder on sm_100a: the decompress direction of the compress block store
// (CompressBlockStore_GetStoredBlock / DecompressBlock, lib/compressblockstore/longtail_compressblockstore.c:271-455, and
// ZStdCompressionAPI_Decompress, lib/zstd/longtail_zstd.c:143-176 -> ZSTD_decompressDCtx).
//
// Written from the format (RFC 8878) and checked against the reference decoder's behaviour where the text leaves room
// (lib/zstd/ext/common/entropy_common.c:42-187 NCount, :242-300 Huffman weights; common/fse_decompress.c:173-234 the two-state tail rule;
// decompress/huf_decompress.c:385-520 table layout; decompress/zstd_decompress_block.c:485-604 sequence tables, :1229-1346 sequence
// decoding order and repcode rules).  Any frame a conforming encoder produces is accepted (all block / literal / sequence modes,
// repeat tables, treeless literals); dictionaries and frame checksums are not part of longtail's use and are rejected / skipped.
//
// One warp per frame from a queue.  Per block: the entropy tables (<= 512 + 4096 entries) are built by lane 0, the four Huffman
// streams of the literals are decoded by four lanes, the sequence bitstream (three interleaved FSE states: a serial chain) by lane 0
// into a sequence list, and the sequences are then EXECUTED by the whole warp — literal and match copies are lane-parallel, an
// overlapping match (offset < length) repeats its period, like the LZ4 decoder.  Malformed input ends in out_len = 0xffffffff.
//
// The frame decoder below is also compiled for the HOST by tests/zstd_dec_host.cpp (LT_ZSTD_DEC_HOST: one lane, the warp
// primitives reduced to identities) so that its serial logic is checked against the reference's encoder output in the CPU suite.
#ifndef LT_ZSTD_DEC_HOST
#include "lt_device.cuh"
#include "lt_kernels.h"
#define ZD_LANES 32u
#endif

namespace ltb {

namespace {

constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t ZD_BLOCK_MAX = 128u << 10;
constexpr uint32_t ZD_MAX_SEQ = ZD_BLOCK_MAX / 3 + 8; // a sequence copies at least 3 match bytes
constexpr uint32_t ZD_BAD = 0xffffffffu;

struct FseDEntry
{
    uint16_t next_base; // new state = next_base + read(nb_bits)
    uint8_t nb_bits;
    uint8_t symbol;
};
struct FseDTable
{
    uint32_t log;
    uint32_t valid;
    FseDEntry e[512];
};

struct ZstdDecWorker
{
    uint16_t huf[4096]; // symbol | nb_bits << 8, indexed by the next huf_log bits of the stream
    uint32_t huf_log, huf_valid;
    FseDTable ll, of, ml, wt; // wt: the Huffman weights' own FSE table (log <= 6)
    uint8_t lits[ZD_BLOCK_MAX + 64];
    uint32_t seq_ll[ZD_MAX_SEQ], seq_ml[ZD_MAX_SEQ], seq_off[ZD_MAX_SEQ];
    int16_t norm[64];
    uint16_t symbol_next[64];
    uint8_t spread[512];
    uint8_t weights[256];
    uint32_t rank_count[16], rank_start[16];
    uint32_t rep[3];
};

__device__ __forceinline__ uint32_t hibit(uint32_t v) { return 31u - (uint32_t)__clz(v); }

// 4 bytes at an arbitrary address (the caller guarantees 7 readable bytes from the aligned word below it)
__device__ __forceinline__ uint32_t ld32u(const uint8_t* p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    return __funnelshift_r(w[0], w[1], (uint32_t)(a & 3u) * 8u);
}

// ------------------------------------------------------------------ backward bit reader (streams are read from their last byte)
struct BitR
{
    const uint8_t* p;
    int32_t pos; // bits not yet consumed; goes negative when the stream is over-read
};
__device__ bool br_init(BitR& r, const uint8_t* p, uint32_t size)
{
    if (!size) return false;
    const uint32_t last = p[size - 1];
    if (!last) return false; // the end mark is missing
    r.p = p;
    r.pos = (int32_t)(size - 1) * 8 + (int32_t)hibit(last);
    return true;
}
// the next n (<= 24) bits, most significant first; bits before the start of the stream read as zero
__device__ __forceinline__ uint32_t br_peek(const BitR& r, uint32_t n)
{
    const int32_t lo = r.pos - (int32_t)n;
    if (lo >= 0) return (ld32u(r.p + (lo >> 3)) >> (lo & 7)) & ((1u << n) - 1u);
    if (r.pos <= 0) return 0;
    return ((ld32u(r.p) & ((1u << r.pos) - 1u)) << (uint32_t)(-lo)) & ((1u << n) - 1u);
}
__device__ __forceinline__ uint32_t br_read(BitR& r, uint32_t n)
{
    if (!n) return 0;
    uint32_t v;
    if (n > 24) { v = br_peek(r, n - 16) << 16; r.pos -= (int32_t)(n - 16); v |= br_peek(r, 16); r.pos -= 16; }
    else { v = br_peek(r, n); r.pos -= (int32_t)n; }
    return v;
}

// ------------------------------------------------------------------ FSE tables
// normalised counts as the encoder wrote them (RFC 8878 4.1.1; entropy_common.c:42-187); returns bytes consumed or ZD_BAD
__device__ uint32_t read_ncount(int16_t* norm, uint32_t* max_symbol, uint32_t* table_log, const uint8_t* src, uint32_t size, uint32_t max_log)
{
    uint32_t bitpos = 0;
    const uint32_t total_bits = size * 8;
    auto peek = [&](uint32_t n) -> uint32_t {
        uint32_t v = 0;
        for (uint32_t i = 0; i < 3; ++i) // n <= 14 bits span at most 3 bytes
        {
            const uint32_t b = (bitpos >> 3) + i;
            if (b < size) v |= (uint32_t)src[b] << (8 * i);
        }
        return (v >> (bitpos & 7)) & ((1u << n) - 1u);
    };
    const uint32_t log = peek(4) + 5;
    bitpos += 4;
    if (log > max_log) return ZD_BAD;
    int remaining = (1 << log) + 1, threshold = 1 << log;
    uint32_t nb_bits = log + 1, symbol = 0;
    const uint32_t limit = *max_symbol + 1;
    bool previous0 = false;
    for (uint32_t i = 0; i < limit; ++i) norm[i] = 0;
    while (remaining > 1 && symbol < limit)
    {
        if (previous0)
        {
            for (;;)
            {
                const uint32_t r = peek(2);
                bitpos += 2;
                symbol += r;
                if (r != 3) break;
                if (bitpos > total_bits) return ZD_BAD;
            }
            if (symbol >= limit) return ZD_BAD;
        }
        const int max = (2 * threshold - 1) - remaining;
        int count;
        const uint32_t low = peek(nb_bits - 1);
        if ((int)low < max) { count = (int)low; bitpos += nb_bits - 1; }
        else
        {
            count = (int)peek(nb_bits);
            if (count >= threshold) count -= max;
            bitpos += nb_bits;
        }
        count--;
        remaining -= count < 0 ? -count : count;
        norm[symbol++] = (int16_t)count;
        previous0 = count == 0;
        if (remaining < 1) return ZD_BAD;
        while (remaining < threshold) { nb_bits--; threshold >>= 1; }
        if (bitpos > total_bits) return ZD_BAD;
    }
    if (remaining != 1 || symbol > limit) return ZD_BAD;
    *max_symbol = symbol - 1;
    *table_log = log;
    return (bitpos + 7) >> 3;
}

// decoding table from normalised counts (zstd_decompress_block.c:485-604 / fse_decompress.c:58-130)
__device__ void build_fse_dtable(ZstdDecWorker* W, FseDTable* t, const int16_t* norm, uint32_t max_symbol, uint32_t log)
{
    const uint32_t size = 1u << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    uint32_t high = size - 1;
    uint8_t* spread = W->spread;
    uint16_t* next = W->symbol_next;
    for (uint32_t s = 0; s <= max_symbol; ++s)
    {
        if (norm[s] == -1) { spread[high--] = (uint8_t)s; next[s] = 1; }
        else next[s] = (uint16_t)norm[s];
    }
    uint32_t position = 0;
    for (uint32_t s = 0; s <= max_symbol; ++s)
        for (int i = 0; i < norm[s]; ++i)
        {
            spread[position] = (uint8_t)s;
            position = (position + step) & mask;
            while (position > high) position = (position + step) & mask;
        }
    for (uint32_t u = 0; u < size; ++u)
    {
        const uint32_t s = spread[u];
        const uint32_t ns = next[s]++;
        const uint32_t nb = log - hibit(ns);
        t->e[u].symbol = (uint8_t)s;
        t->e[u].nb_bits = (uint8_t)nb;
        t->e[u].next_base = (uint16_t)((ns << nb) - size);
    }
    t->log = log;
    t->valid = 1;
}
__device__ void build_fse_rle(FseDTable* t, uint32_t symbol)
{
    t->e[0].symbol = (uint8_t)symbol;
    t->e[0].nb_bits = 0;
    t->e[0].next_base = 0;
    t->log = 0;
    t->valid = 1;
}

// extra bits and baselines of the length / offset codes (RFC 8878 3.1.1.3.2.1.1)
__constant__ uint8_t d_ll_bits[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
__constant__ uint8_t d_ml_bits[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                      1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
__constant__ uint32_t d_ll_base[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024,
                                       2048, 4096, 8192, 16384, 32768, 65536};
__constant__ uint32_t d_ml_base[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
                                       35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
__constant__ int16_t d_ll_default[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
__constant__ int16_t d_ml_default[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                         1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
__constant__ int16_t d_of_default[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};

// one sequence-symbol table per its mode byte field (0 predefined, 1 RLE, 2 FSE, 3 repeat); returns bytes consumed or ZD_BAD
__device__ uint32_t read_seq_table(ZstdDecWorker* W, FseDTable* t, uint32_t mode, const uint8_t* src, uint32_t size, uint32_t max_symbol, uint32_t max_log,
                                   const int16_t* default_norm, uint32_t default_max, uint32_t default_log)
{
    if (mode == 0)
    {
        for (uint32_t i = 0; i <= default_max; ++i) W->norm[i] = default_norm[i];
        build_fse_dtable(W, t, W->norm, default_max, default_log);
        return 0;
    }
    if (mode == 1)
    {
        if (size < 1 || src[0] > max_symbol) return ZD_BAD;
        build_fse_rle(t, src[0]);
        return 1;
    }
    if (mode == 2)
    {
        uint32_t ms = max_symbol, log = 0;
        const uint32_t used = read_ncount(W->norm, &ms, &log, src, size, max_log);
        if (used == ZD_BAD || used > size) return ZD_BAD;
        build_fse_dtable(W, t, W->norm, ms, log);
        return used;
    }
    return t->valid ? 0 : ZD_BAD; // repeat
}

// ------------------------------------------------------------------ Huffman literals
// tree description -> decoding table (entropy_common.c:242-300, huf_decompress.c:385-520); returns bytes consumed or ZD_BAD
__device__ uint32_t read_huf_table(ZstdDecWorker* W, const uint8_t* src, uint32_t size)
{
    if (!size) return ZD_BAD;
    uint8_t* weights = W->weights;
    uint32_t count, used;
    const uint32_t hb = src[0];
    if (hb >= 128)
    {
        count = hb - 127;
        used = 1 + (count + 1) / 2;
        if (used > size) return ZD_BAD;
        for (uint32_t n = 0; n < count; n += 2)
        {
            weights[n] = src[1 + n / 2] >> 4;
            weights[n + 1] = src[1 + n / 2] & 15;
        }
    }
    else
    {
        used = 1 + hb;
        if (used > size || hb < 2) return ZD_BAD;
        uint32_t ms = 12, log = 0;
        const uint32_t head = read_ncount(W->norm, &ms, &log, src + 1, hb, 6);
        if (head == ZD_BAD || head >= hb) return ZD_BAD;
        build_fse_dtable(W, &W->wt, W->norm, ms, log);
        BitR r;
        if (!br_init(r, src + 1 + head, hb - head)) return ZD_BAD;
        // two interleaved states; the stream ends when a state update over-reads (fse_decompress.c:215-230)
        uint32_t s1 = br_read(r, log), s2 = br_read(r, log);
        if (r.pos < 0) return ZD_BAD;
        count = 0;
        for (;;)
        {
            if (count > 253) return ZD_BAD;
            FseDEntry e = W->wt.e[s1];
            weights[count++] = e.symbol;
            s1 = e.next_base + br_read(r, e.nb_bits);
            if (r.pos < 0) { weights[count++] = W->wt.e[s2].symbol; break; }
            if (count > 253) return ZD_BAD;
            e = W->wt.e[s2];
            weights[count++] = e.symbol;
            s2 = e.next_base + br_read(r, e.nb_bits);
            if (r.pos < 0) { weights[count++] = W->wt.e[s1].symbol; break; }
        }
    }
    // the last weight is implied: the weights must complete a power of two
    uint32_t total = 0;
    for (uint32_t i = 0; i < 16; ++i) W->rank_count[i] = 0;
    for (uint32_t n = 0; n < count; ++n)
    {
        if (weights[n] > 12) return ZD_BAD;
        W->rank_count[weights[n]]++;
        total += (1u << weights[n]) >> 1;
    }
    if (!total) return ZD_BAD;
    const uint32_t log = hibit(total) + 1;
    if (log > 12) return ZD_BAD;
    const uint32_t rest = (1u << log) - total;
    if (!rest || (rest & (rest - 1))) return ZD_BAD;
    const uint32_t last_weight = hibit(rest) + 1;
    weights[count] = (uint8_t)last_weight;
    W->rank_count[last_weight]++;
    count++;
    if (W->rank_count[1] < 2 || (W->rank_count[1] & 1)) return ZD_BAD;
    // table: weights ascending (longest codes first), symbols ascending inside a weight, 2^(w-1) entries per symbol
    uint32_t start = 0;
    for (uint32_t w = 1; w <= log; ++w)
    {
        W->rank_start[w] = start;
        start += W->rank_count[w] << (w - 1);
    }
    if (start != (1u << log)) return ZD_BAD;
    for (uint32_t s = 0; s < count; ++s)
    {
        const uint32_t w = weights[s];
        if (!w) continue;
        const uint32_t len = 1u << (w - 1);
        const uint16_t entry = (uint16_t)(s | ((log + 1 - w) << 8));
        const uint32_t at = W->rank_start[w];
        for (uint32_t i = 0; i < len; ++i) W->huf[at + i] = entry;
        W->rank_start[w] = at + len;
    }
    W->huf_log = log;
    W->huf_valid = 1;
    return used;
}

// one Huffman stream of `count` symbols into dst; false when the stream is malformed
__device__ bool huf_decode_stream(const ZstdDecWorker* W, const uint8_t* src, uint32_t size, uint8_t* dst, uint32_t count)
{
    BitR r;
    if (!br_init(r, src, size)) return false;
    const uint32_t log = W->huf_log;
    for (uint32_t i = 0; i < count; ++i)
    {
        const uint32_t e = W->huf[br_peek(r, log)];
        dst[i] = (uint8_t)e;
        r.pos -= (int32_t)(e >> 8);
    }
    return r.pos == 0; // every bit consumed, none over-read
}

// warp-wide byte copy helpers
__device__ void copy_fwd_w(uint8_t* dst, const uint8_t* src, uint32_t n, uint32_t lane)
{
    for (uint32_t i = lane; i < n; i += ZD_LANES) dst[i] = src[i];
}

// decode one frame; returns decoded size or ZD_BAD (warp-uniform)
__device__ uint32_t zstd_decode_frame(ZstdDecWorker* W, const uint8_t* __restrict__ src, uint32_t n, uint8_t* __restrict__ dst, uint32_t cap, uint32_t lane)
{
    if (n < 6 || src[0] != 0x28 || src[1] != 0xB5 || src[2] != 0x2F || src[3] != 0xFD) return ZD_BAD;
    const uint32_t fhd = src[4];
    const uint32_t fcs_flag = fhd >> 6, single = (fhd >> 5) & 1u, checksum = (fhd >> 2) & 1u, dict_flag = fhd & 3u;
    if (fhd & 0x08) return ZD_BAD; // reserved bit
    if (dict_flag) return ZD_BAD;  // longtail never uses dictionaries
    uint32_t ip = 5;
    if (!single) ip += 1; // window descriptor: the whole output buffer is the window here
    const uint32_t fcs_bytes = fcs_flag == 0 ? single : fcs_flag == 1 ? 2u : fcs_flag == 2 ? 4u : 8u;
    if (ip + fcs_bytes > n) return ZD_BAD;
    uint64_t content = ~0ull;
    if (fcs_bytes)
    {
        content = 0;
        for (uint32_t i = 0; i < fcs_bytes; ++i) content |= (uint64_t)src[ip + i] << (8 * i);
        if (fcs_bytes == 2) content += 256;
        if (content > cap) return ZD_BAD;
    }
    ip += fcs_bytes;
    if (lane == 0)
    {
        W->rep[0] = 1; W->rep[1] = 4; W->rep[2] = 8;
        W->huf_valid = 0; W->ll.valid = 0; W->of.valid = 0; W->ml.valid = 0;
    }
    __syncwarp();
    uint32_t op = 0;
    for (;;)
    {
        if (ip + 3 > n) return ZD_BAD;
        const uint32_t bh = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16);
        ip += 3;
        const uint32_t last = bh & 1u, type = (bh >> 1) & 3u, bsize = bh >> 3;
        if (type == 3) return ZD_BAD;
        if (type == 0) // raw
        {
            if (ip + bsize > n || bsize > cap - op) return ZD_BAD;
            copy_fwd_w(dst + op, src + ip, bsize, lane);
            ip += bsize;
            op += bsize;
        }
        else if (type == 1) // RLE
        {
            if (ip + 1 > n || bsize > cap - op) return ZD_BAD;
            const uint8_t v = src[ip];
            for (uint32_t i = lane; i < bsize; i += ZD_LANES) dst[op + i] = v;
            ip += 1;
            op += bsize;
        }
        else
        {
            if (bsize < 2 || ip + bsize > n || bsize > ZD_BLOCK_MAX) return ZD_BAD;
            const uint8_t* b = src + ip;
            // ---- literals section
            const uint32_t lt = b[0] & 3u, sf = (b[0] >> 2) & 3u;
            uint32_t lh, regen, csize = 0, streams = 1;
            if (lt < 2)
            {
                if (sf == 0 || sf == 2) { lh = 1; regen = b[0] >> 3; }
                else if (sf == 1) { lh = 2; if (bsize < 2) return ZD_BAD; regen = (b[0] >> 4) | ((uint32_t)b[1] << 4); }
                else { lh = 3; if (bsize < 3) return ZD_BAD; regen = (b[0] >> 4) | ((uint32_t)b[1] << 4) | ((uint32_t)b[2] << 12); }
            }
            else
            {
                if (bsize < 5) return ZD_BAD;
                const uint32_t v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
                if (sf < 2) { lh = 3; regen = (v >> 4) & 0x3ffu; csize = (v >> 14) & 0x3ffu; streams = sf == 0 ? 1 : 4; }
                else if (sf == 2) { lh = 4; regen = (v >> 4) & 0x3fffu; csize = v >> 18; streams = 4; }
                else { lh = 5; regen = (v >> 4) & 0x3ffffu; csize = (v >> 22) | ((uint32_t)b[4] << 10); streams = 4; }
            }
            if (regen > ZD_BLOCK_MAX) return ZD_BAD;
            uint32_t lit_end; // offset in the block where the sequences section starts
            const uint8_t* lit_src = W->lits;
            if (lt == 0)
            {
                if (lh + regen > bsize) return ZD_BAD;
                lit_src = b + lh; // raw literals are used in place
                lit_end = lh + regen;
            }
            else if (lt == 1)
            {
                if (lh + 1 > bsize) return ZD_BAD;
                const uint8_t v = b[lh];
                for (uint32_t i = lane; i < regen; i += ZD_LANES) W->lits[i] = v;
                lit_end = lh + 1;
            }
            else
            {
                if (lh + csize > bsize) return ZD_BAD;
                uint32_t tree = 0;
                if (lt == 2)
                {
                    if (lane == 0) tree = read_huf_table(W, b + lh, csize);
                    tree = __shfl_sync(FULL, tree, 0);
                    if (tree == ZD_BAD) return ZD_BAD;
                }
                else if (!W->huf_valid) return ZD_BAD;
                __syncwarp();
                const uint8_t* s0 = b + lh + tree;
                const uint32_t ssize = csize - tree;
                bool ok = true;
                if (streams == 1)
                {
                    if (lane == 0) ok = huf_decode_stream(W, s0, ssize, W->lits, regen);
                }
                else
                {
                    if (ssize < 10) return ZD_BAD;
                    const uint32_t c1 = s0[0] | ((uint32_t)s0[1] << 8), c2 = s0[2] | ((uint32_t)s0[3] << 8), c3 = s0[4] | ((uint32_t)s0[5] << 8);
                    if (6 + c1 + c2 + c3 >= ssize) return ZD_BAD;
                    const uint32_t c4 = ssize - 6 - c1 - c2 - c3;
                    const uint32_t seg = (regen + 3) / 4;
                    if (3 * seg > regen) return ZD_BAD;
                    for (uint32_t q = lane; q < 4; q += ZD_LANES) // four lanes, one stream each
                    {
                        const uint32_t off = 6 + (q > 0 ? c1 : 0) + (q > 1 ? c2 : 0) + (q > 2 ? c3 : 0);
                        const uint32_t sz = q == 0 ? c1 : q == 1 ? c2 : q == 2 ? c3 : c4;
                        const uint32_t cnt = q < 3 ? seg : regen - 3 * seg;
                        ok = huf_decode_stream(W, s0 + off, sz, W->lits + q * seg, cnt) && ok;
                    }
                }
                if (__any_sync(FULL, !ok)) return ZD_BAD;
                lit_end = lh + csize;
            }
            __syncwarp();
            // ---- sequences section
            if (lit_end >= bsize) return ZD_BAD;
            const uint8_t* sq = b + lit_end;
            const uint32_t sq_size = bsize - lit_end;
            uint32_t nb_seq = sq[0], sp = 1;
            if (nb_seq >= 128)
            {
                if (nb_seq == 255) { if (sq_size < 3) return ZD_BAD; nb_seq = sq[1] + ((uint32_t)sq[2] << 8) + 0x7F00; sp = 3; }
                else { if (sq_size < 2) return ZD_BAD; nb_seq = ((nb_seq - 128) << 8) + sq[1]; sp = 2; }
            }
            if (nb_seq > ZD_MAX_SEQ) return ZD_BAD;
            uint32_t produced = 0; // bytes this block regenerates
            if (nb_seq)
            {
                uint32_t status = 0;
                if (lane == 0)
                {
                    status = ZD_BAD;
                    do
                    {
                        if (sp + 1 > sq_size) break;
                        const uint32_t modes = sq[sp++];
                        if (modes & 3u) break;
                        uint32_t used = read_seq_table(W, &W->ll, modes >> 6, sq + sp, sq_size - sp, 35, 9, d_ll_default, 35, 6);
                        if (used == ZD_BAD) break;
                        sp += used;
                        used = read_seq_table(W, &W->of, (modes >> 4) & 3u, sq + sp, sq_size - sp, 31, 8, d_of_default, 28, 5);
                        if (used == ZD_BAD) break;
                        sp += used;
                        used = read_seq_table(W, &W->ml, (modes >> 2) & 3u, sq + sp, sq_size - sp, 52, 9, d_ml_default, 52, 6);
                        if (used == ZD_BAD) break;
                        sp += used;
                        BitR r;
                        if (sp >= sq_size || !br_init(r, sq + sp, sq_size - sp)) break;
                        uint32_t s_ll = br_read(r, W->ll.log), s_of = br_read(r, W->of.log), s_ml = br_read(r, W->ml.log);
                        uint32_t rep0 = W->rep[0], rep1 = W->rep[1], rep2 = W->rep[2];
                        uint64_t lit_used = 0, out_bytes = 0;
                        bool bad = r.pos < 0;
                        for (uint32_t i = 0; i < nb_seq && !bad; ++i)
                        {
                            const FseDEntry e_ll = W->ll.e[s_ll], e_of = W->of.e[s_of], e_ml = W->ml.e[s_ml];
                            const uint32_t of_code = e_of.symbol;
                            if (of_code > 31) { bad = true; break; }
                            const uint32_t off_base = (1u << of_code) + br_read(r, of_code);
                            const uint32_t ml = d_ml_base[e_ml.symbol] + br_read(r, d_ml_bits[e_ml.symbol]);
                            const uint32_t ll = d_ll_base[e_ll.symbol] + br_read(r, d_ll_bits[e_ll.symbol]);
                            uint32_t offset;
                            if (off_base > 3) { offset = off_base - 3; rep2 = rep1; rep1 = rep0; rep0 = offset; }
                            else
                            {
                                const uint32_t idx = off_base - 1 + (ll == 0);
                                if (idx == 0) offset = rep0;
                                else
                                {
                                    offset = idx == 1 ? rep1 : idx == 2 ? rep2 : rep0 - 1;
                                    if (!offset) { bad = true; break; }
                                    if (idx != 1) rep2 = rep1;
                                    rep1 = rep0;
                                    rep0 = offset;
                                }
                            }
                            W->seq_ll[i] = ll; W->seq_ml[i] = ml; W->seq_off[i] = offset;
                            lit_used += ll;
                            out_bytes += (uint64_t)ll + ml;
                            if (i + 1 < nb_seq)
                            {
                                s_ll = e_ll.next_base + br_read(r, e_ll.nb_bits);
                                s_ml = e_ml.next_base + br_read(r, e_ml.nb_bits);
                                s_of = e_of.next_base + br_read(r, e_of.nb_bits);
                            }
                            if (r.pos < 0) bad = true;
                        }
                        if (bad || r.pos != 0 || lit_used > regen) break;
                        out_bytes += regen - lit_used;
                        if (out_bytes > ZD_BLOCK_MAX) break;
                        W->rep[0] = rep0; W->rep[1] = rep1; W->rep[2] = rep2;
                        status = (uint32_t)out_bytes;
                    } while (false);
                }
                status = __shfl_sync(FULL, status, 0);
                if (status == ZD_BAD) return ZD_BAD;
                produced = status;
            }
            else
            {
                if (sp != sq_size) return ZD_BAD;
                produced = regen;
            }
            if (produced > cap - op) return ZD_BAD;
            __syncwarp();
            // ---- execute: literals, then the match, sequence by sequence; all lanes copy
            uint32_t lp = 0;
            for (uint32_t i = 0; i < nb_seq; ++i)
            {
                const uint32_t ll = W->seq_ll[i], ml = W->seq_ml[i], off = W->seq_off[i];
                copy_fwd_w(dst + op, lit_src + lp, ll, lane);
                lp += ll;
                op += ll;
                if (off > op) return ZD_BAD; // reaches before the start of the frame
                __syncwarp(); // the literals just written may be the source of the match
                const uint32_t from = op - off;
                if (off >= ml)
                    for (uint32_t k = lane; k < ml; k += ZD_LANES) dst[op + k] = dst[from + k];
                else
                    for (uint32_t k = lane; k < ml; k += ZD_LANES) dst[op + k] = dst[from + k % off];
                op += ml;
                __syncwarp();
            }
            copy_fwd_w(dst + op, lit_src + lp, regen - lp, lane);
            op += regen - lp;
            ip += bsize;
        }
        __syncwarp();
        if (last) break;
    }
    if (checksum) ip += 4; // XXH64 of the content: not verified (longtail never asks for one)
    if (ip != n) return ZD_BAD;
    if (content != ~0ull && content != op) return ZD_BAD;
    return op;
}

} // namespace

#ifndef LT_ZSTD_DEC_HOST
__global__ void __launch_bounds__(128)
k_zstd_decode(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
              uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
              uint32_t* __restrict__ out_len, uint32_t frame_count, ZstdDecWorker* workers, uint32_t* queue)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    ZstdDecWorker* W = workers + warp;
    for (;;)
    {
        uint32_t f = 0;
        if (lane == 0) f = atomicAdd(queue, 1u);
        f = __shfl_sync(FULL, f, 0);
        if (f >= frame_count) break;
        const uint32_t got = zstd_decode_frame(W, in_base + in_off[f], in_len[f], out_base + out_off[f], out_cap[f], lane);
        if (lane == 0) out_len[f] = got;
        __syncwarp();
    }
}

size_t zstd_dec_worker_bytes() { return sizeof(ZstdDecWorker); }
uint32_t zstd_dec_worker_count(uint32_t frame_count, int sm_count)
{
    const uint32_t resident = (uint32_t)sm_count * 16u;
    return frame_count < resident ? (frame_count + 3u) & ~3u : resident;
}

cudaError_t launch_zstd_decode(const uint8_t* d_in, const uint64_t* d_in_off, const uint32_t* d_in_len, uint8_t* d_out, const uint64_t* d_out_off,
                               const uint32_t* d_out_cap, uint32_t* d_out_len, uint32_t frame_count, void* d_workers, uint32_t worker_count,
                               uint32_t* d_queue, cudaStream_t st)
{
    if (!frame_count) return cudaSuccess;
    cudaMemsetAsync(d_queue, 0, sizeof(uint32_t), st);
    k_zstd_decode<<<worker_count / 4, 128, 0, st>>>(d_in, d_in_off, d_in_len, d_out, d_out_off, d_out_cap, d_out_len, frame_count,
                                                     static_cast<ZstdDecWorker*>(d_workers), d_queue);
    return cudaGetLastError();
}
#endif // !LT_ZSTD_DEC_HOST

} // namespace ltb
