// lz4.cu — bit-exact LZ4 block encoder (LZ4_compress_fast, acceleration 1, notLimited / noDict) on sm_100a.
//
// Re-design of lib/lz4/ext/lz4.c:930-1338 (LZ4_compress_generic_validated as reached from
// LZ4CompressionAPI_Compress, lib/lz4/longtail_lz4.c:52-77) for one WARP per stored block:
//
//  * the greedy parse is inherently sequential (the hash table's content depends on the parse), so a warp owns one block
//    and its 16 KiB table in shared memory; parallelism across blocks comes from the grid;
//  * inside the block the warp speculates the next 32 probe positions of the deterministic skip schedule
//    (step = searchMatchNb++ >> 6, lz4.c:1023-1053): every lane hashes its position, intra-batch table conflicts are
//    resolved with __match_any_sync (a lane's candidate is the closest lower lane with the same hash, else the table), the
//    first matching lane wins (__ballot_sync) and only the table writes up to it are committed — exactly the state the
//    sequential loop would have reached;
//  * match extension (LZ4_count, lz4.c:1181), backward catch-up (:1104-1109), literal / length emission (:1112-1226) are
//    lane-parallel.
//
// Output bytes are identical to the reference's for every input (tests/test_gpu_lz4.py).
#include "lt_device.cuh"
#include "lt_kernels.h"

#include <stdlib.h>

namespace ltb {

namespace {

constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t LZ4_TABLE_BYTES = 32768; // byU32: 4096 x {position, the 4 bytes at that position}; byU16: 8192 x u16 in the first half
constexpr uint32_t LZ4_64K_LIMIT = 65536 + 11; // LZ4_64Klimit, lz4.c:710
constexpr uint32_t LZ4_MAX_DISTANCE = 65535;

__device__ __forceinline__ uint32_t rd32(const uint8_t* __restrict__ s, uint32_t pos)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(s + (pos & ~3u));
    return __funnelshift_r(w[0], w[1], (pos & 3u) * 8u); // w[1] is only consumed when pos is unaligned (then it is in bounds)
}
__device__ __forceinline__ uint64_t rd64(const uint8_t* __restrict__ s, uint32_t pos)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(s + (pos & ~3u));
    const uint32_t sh = (pos & 3u) * 8u;
    const uint32_t a = w[0], b = w[1], c = w[2];
    return (uint64_t)__funnelshift_r(a, b, sh) | ((uint64_t)__funnelshift_r(b, c, sh) << 32);
}

template <bool U16>
__device__ __forceinline__ uint32_t lz4_hash(const uint8_t* __restrict__ s, uint32_t pos)
{
    if (U16) return (rd32(s, pos) * 2654435761u) >> 19;                   // LZ4_hash4, 13 bits (lz4.c:777-783)
    return (uint32_t)(((rd64(s, pos) << 24) * 889523592379ull) >> 52);    // LZ4_hash5, 12 bits (lz4.c:785-795)
}
template <bool U16>
__device__ __forceinline__ uint32_t lz4_hash_word(uint64_t w)
{
    if (U16) return ((uint32_t)w * 2654435761u) >> 19;
    return (uint32_t)(((w << 24) * 889523592379ull) >> 52);
}
// A byU32 entry carries the 4 bytes found at its position next to the position, so the match test of a candidate
// (LZ4_read32(match) == LZ4_read32(ip), lz4.c:1092, :1262) needs no second dependent load from the source: table -> compare instead
// of table -> source -> compare.  The table is this encoder's private state; what it answers is unchanged.  byU16 blocks
// (< 64 KiB) keep plain 16-bit positions and read the source.
struct TabEntry
{
    uint32_t pos, tag;
};
template <bool U16>
__device__ __forceinline__ TabEntry tab_get(const uint32_t* t32, uint32_t h)
{
    TabEntry e;
    if (U16) { e.pos = (uint32_t)reinterpret_cast<const uint16_t*>(t32)[h]; e.tag = 0; }
    else { const uint2 v = reinterpret_cast<const uint2*>(t32)[h]; e.pos = v.x; e.tag = v.y; }
    return e;
}
template <bool U16>
__device__ __forceinline__ void tab_put(uint32_t* t32, uint32_t h, uint32_t v, uint32_t tag)
{
    if (U16) reinterpret_cast<uint16_t*>(t32)[h] = (uint16_t)v; else reinterpret_cast<uint2*>(t32)[h] = make_uint2(v, tag);
}
// does the candidate start with the 4 bytes `word`?
template <bool U16>
__device__ __forceinline__ bool cand_matches(const uint8_t* __restrict__ s, uint32_t cand, uint32_t tag, uint32_t word)
{
    return U16 ? rd32(s, cand) == word : tag == word;
}

// position of the k-th probe of a search that started at S (k = 0, 1, ...): advances are 1 for k = 0 and (63+k)>>6 after
__device__ __forceinline__ uint32_t probe_pos(uint32_t S, uint32_t k)
{
    if (k == 0) return S;
    const uint32_t u = k - 1, q = u >> 6, r = u & 63u;
    return S + 1u + 32u * q * (q + 1u) + r * (q + 1u);
}
__device__ __forceinline__ uint32_t probe_advance(uint32_t k) { return k == 0 ? 1u : (63u + k) >> 6; }

// lane-parallel: dst[0..n) = 255 repeated, then the tail byte — the LZ4 length continuation (lz4.c:1126-1130, 1213-1224)
__device__ __forceinline__ uint32_t put_length(uint8_t* __restrict__ dst, uint32_t op, uint32_t len, uint32_t lane)
{
    const uint32_t full = len / 255u;
    for (uint32_t i = lane; i < full; i += 32) dst[op + i] = 255;
    if (lane == 0) dst[op + full] = (uint8_t)(len - full * 255u);
    return op + full + 1u;
}

__device__ __forceinline__ void copy_bytes(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t n, uint32_t lane)
{
    // literal runs: short in compressible data, the whole block in incompressible data -> aligned word stores assembled
    // from the (arbitrarily aligned) source with a funnel shift; head and tail by bytes
    uint32_t i = 0;
    if (n >= 128)
    {
        const uint32_t head = (4u - ((uintptr_t)dst & 3u)) & 3u;
        if (lane < head) dst[lane] = src[lane];
        const uint32_t words = (n - head - 4u) >> 2; // the funnel reads one word ahead: stay 4 bytes clear of the end
        const uint8_t* s2 = src + head;
        const uint32_t sh = (uint32_t)((uintptr_t)s2 & 3u) * 8u;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(s2 - ((uintptr_t)s2 & 3u));
        uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + head);
        for (uint32_t w = lane; w < words; w += 32) d4[w] = __funnelshift_r(sw[w], sw[w + 1], sh);
        i = head + words * 4u;
    }
    for (uint32_t j = i + lane; j < n; j += 32) dst[j] = src[j];
}

// Literal runs of LZ4_DEFER_MIN bytes or more are not copied by the (latency-bound) parsing warp: their position in the
// output is fixed by the parse, so the warp only records {dst, src, len} and k_lz4_copy moves the bytes afterwards with the
// whole GPU.  Incompressible blocks are one such run.
constexpr uint32_t LZ4_DEFER_MIN = 2048;

struct CopyJobs
{
    uint3* jobs; // {dst offset, src offset, length}
    uint32_t count;
};

__device__ __forceinline__ void emit_literals(uint8_t* __restrict__ dst, uint32_t op, const uint8_t* __restrict__ src, uint32_t from, uint32_t lit,
                                              uint32_t lane, CopyJobs& cj)
{
    if (lit >= LZ4_DEFER_MIN)
    {
        if (lane == 0) cj.jobs[cj.count] = make_uint3(op, from, lit);
        ++cj.count;
    }
    else if (lit)
        copy_bytes(dst + op, src + from, lit, lane);
}

template <bool U16>
__device__ uint32_t lz4_encode_block(const uint8_t* __restrict__ src, uint32_t n, uint8_t* __restrict__ dst, uint32_t* table, uint32_t lane,
                                     CopyJobs& cj)
{
    uint32_t op = 0;
    uint32_t anchor = 0;
    if (n >= 13) // LZ4_minLength (lz4.c:1001)
    {
        {
            // zeroed table = every slot points at position 0 (a legal candidate, lz4.c:1004-1010), so the tags start as its 4 bytes
            const uint32_t first4 = U16 ? 0u : rd32(src, 0);
            for (uint32_t i = lane; i < LZ4_TABLE_BYTES / 4; i += 32) table[i] = (i & 1u) ? first4 : 0u;
        }
        __syncwarp();
        const uint32_t mflimit_plus_one = n - 11;
        const uint32_t matchlimit = n - 5;
        if (lane == 0) tab_put<U16>(table, lz4_hash<U16>(src, 0), 0, rd32(src, 0)); // lz4.c:1004-1010
        __syncwarp();
        uint32_t S = 1;   // start of the current search
        uint32_t k0 = 0;  // probes of this search already committed
        for (;;)
        {
            // ---------------- search: 32 probes per step (lz4.c:1043-1100)
            uint32_t ip, match;
            bool found = false, finished = false;
            // the probe positions of a search are known in advance, so the 8 source bytes of the NEXT batch are requested
            // before the current batch is resolved (wasted when a match ends the search, which is harmless)
            uint32_t p_next = probe_pos(S, k0 + lane);
            bool valid_next = p_next + probe_advance(k0 + lane) <= mflimit_plus_one && p_next < n;
            uint64_t w_next = valid_next ? rd64(src, p_next) : 0ull;
            for (;;)
            {
                const uint32_t p = p_next;
                const bool valid = valid_next; // else: goto _last_literals before probing
                const uint64_t w = w_next;
                {
                    const uint32_t kn = k0 + 32 + lane;
                    p_next = probe_pos(S, kn);
                    valid_next = p_next + probe_advance(kn) <= mflimit_plus_one && p_next < n;
                    w_next = valid_next ? rd64(src, p_next) : 0ull;
                }
                const uint32_t h = valid ? lz4_hash_word<U16>(w) : 0xffffffffu - lane;
                // the table entry is requested BEFORE the intra-batch conflict resolution: it does not depend on it, and MATCH.ANY plus
                // the dependent shuffles take about as long as the load itself (a lane that ends up using a lower lane's position drops it)
                TabEntry e = {0u, 0u};
                if (valid) e = tab_get<U16>(table, h);
                const uint32_t same = __match_any_sync(FULL, h);
                const uint32_t lower = same & ((1u << lane) - 1u);
                // candidate = what the sequential loop would find in the table: the closest lower lane of this batch with the
                // same hash, else the committed table
                const int from = lower ? 31 - __clz(lower) : (int)lane;
                const uint32_t from_lane = __shfl_sync(FULL, p, from);
                const uint32_t from_tag = __shfl_sync(FULL, (uint32_t)w, from);
                const uint32_t cand = lower ? from_lane : e.pos;
                const uint32_t tag = lower ? from_tag : e.tag;
                bool hit = false;
                if (valid && (U16 || cand + LZ4_MAX_DISTANCE >= p)) hit = cand_matches<U16>(src, cand, tag, (uint32_t)w);
                const uint32_t hits = __ballot_sync(FULL, hit);
                const uint32_t valids = __ballot_sync(FULL, valid);
                // lanes whose table write is committed: valid lanes up to and including the first hit
                const uint32_t first_hit = hits ? (uint32_t)__ffs(hits) - 1u : 32u;
                const uint32_t commit = valids & (first_hit >= 31u ? FULL : ((2u << first_hit) - 1u));
                const uint32_t group = same & commit;
                if ((commit >> lane) & 1u)
                    if ((31 - __clz(group)) == (int)lane) tab_put<U16>(table, h, p, (uint32_t)w); // the last writer of a hash value wins
                __syncwarp();
                if (hits)
                {
                    ip = __shfl_sync(FULL, p, first_hit);
                    match = __shfl_sync(FULL, cand, first_hit);
                    found = true;
                    break;
                }
                if (valids != FULL)
                {
                    finished = true;
                    break;
                }
                k0 += 32;
            }
            if (finished) break;

            // ---------------- catch up (lz4.c:1104-1109)
            for (;;)
            {
                bool eq = false;
                if (ip > anchor + lane && match > lane) eq = src[ip - 1 - lane] == src[match - 1 - lane];
                const uint32_t m = __ballot_sync(FULL, eq);
                const uint32_t run = m == FULL ? 32u : (uint32_t)__ffs(~m) - 1u;
                ip -= run;
                match -= run;
                if (run < 32) break;
            }

            bool literals_done = false;
            for (;;) // _next_match (lz4.c:1140-1296)
            {
                // match length beyond MINMATCH: LZ4_count(ip+4, match+4, matchlimit)
                uint32_t a = ip + 4, b = match + 4, code = 0;
                for (;;)
                {
                    const uint32_t pa = a + 4 * lane;
                    uint32_t eqb = 0; // equal bytes this lane contributes (0..4)
                    bool stop = true;
                    if (pa < matchlimit)
                    {
                        const uint32_t x = rd32(src, pa) ^ rd32(src, b + 4 * lane);
                        eqb = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
                        const uint32_t room = matchlimit - pa;
                        if (eqb > room) eqb = room;
                        stop = eqb < 4u;
                    }
                    const uint32_t stops = __ballot_sync(FULL, stop);
                    if (stops)
                    {
                        const uint32_t f = (uint32_t)__ffs(stops) - 1u;
                        code += 4u * f + __shfl_sync(FULL, eqb, f);
                        break;
                    }
                    code += 128;
                    a += 128;
                    b += 128;
                }
                // token, literals (lz4.c:1112-1137) — skipped for the zero-literal sequences found by the post-match test
                const uint32_t lit = literals_done ? 0u : ip - anchor;
                const uint32_t token_pos = op++;
                if (lit >= 15) op = put_length(dst, op, lit - 15, lane);
                emit_literals(dst, op, src, anchor, lit, lane, cj);
                op += lit;
                const uint32_t off = ip - match;
                if (lane == 0)
                {
                    dst[op] = (uint8_t)off; // LZ4_writeLE16 (lz4.c:1157-1163)
                    dst[op + 1] = (uint8_t)(off >> 8);
                    dst[token_pos] = (uint8_t)((lit >= 15 ? 15u : lit) << 4 | (code >= 15 ? 15u : code));
                }
                op += 2;
                if (code >= 15) op = put_length(dst, op, code - 15, lane);
                ip += code + 4;
                anchor = ip;
                if (ip >= mflimit_plus_one) { finished = true; break; } // lz4.c:1233
                // fill table with ip-2, then test ip itself (lz4.c:1236-1294)
                const uint64_t w2 = rd64(src, ip - 2), w0 = rd64(src, ip);
                const uint32_t h2 = lz4_hash_word<U16>(w2);
                const uint32_t h = lz4_hash_word<U16>(w0);
                uint32_t cand = 0, tag = 0;
                if (lane == 0)
                {
                    tab_put<U16>(table, h2, ip - 2, (uint32_t)w2);
                    const TabEntry e = tab_get<U16>(table, h);
                    cand = e.pos;
                    tag = e.tag;
                    tab_put<U16>(table, h, ip, (uint32_t)w0);
                }
                cand = __shfl_sync(FULL, cand, 0);
                tag = __shfl_sync(FULL, tag, 0);
                __syncwarp();
                if ((U16 || cand + LZ4_MAX_DISTANCE >= ip) && cand_matches<U16>(src, cand, tag, (uint32_t)w0))
                {
                    match = cand;
                    literals_done = true; // token = 0 literals, straight to the next match
                    continue;
                }
                break;
            }
            if (finished) break;
            S = ip + 1; // lz4.c:1298: forwardH = hash(++ip), a fresh search (step 1, searchMatchNb reset)
            k0 = 0;
            (void)found;
        }
    }
    // last literals (lz4.c:1302-1329)
    const uint32_t last = n - anchor;
    const uint32_t token_pos = op++;
    if (lane == 0) dst[token_pos] = (uint8_t)((last >= 15 ? 15u : last) << 4);
    if (last >= 15) op = put_length(dst, op, last - 15, lane);
    emit_literals(dst, op, src, anchor, last, lane, cj);
    op += last;
    return op;
}


// ======================================================================================================================
// v2 encoder — byU32 blocks of at most 16 MiB (every block of the default 8 MiB / 1024-chunk packing): the hash table lives in
// SHARED memory as ONE 32-bit word per slot, {position mod 2^17 | valid | 14-bit tag of the 4 bytes at that position}, 16 KiB per
// warp, 13 warps per SM.  What the table answers is unchanged (lz4.c:1085-1092: the latest position with this hash, usable when it
// is at most 65 535 bytes back and starts with the same 4 bytes):
//   * a position is only ever compared with positions at most 65 535 bytes ahead of it, so 17 bits identify it as long as no
//     entry older than 2^17 bytes survives: every <= 48 KiB of progress the warp sweeps the table and invalidates the entries
//     that are out of reach for good (128 words per lane, ~3 % of the probe work on dense data);
//   * the tag answers LZ4_read32(match) == LZ4_read32(ip) without touching the source; a tag match (true match, or one in 2^14
//     by chance) is verified against the source in the same round trip that fetches the bytes of the match extension.
// A probe batch (32 positions of the skip schedule) is: read the slots, write the own entries, read back.  If every lane reads
// its own entry back, the 32 hashes were distinct; if in addition no tag matched, the table already is what the sequential
// loop would have left and the batch is done — no MATCH.ANY, no global memory round trip.  Everything else (intra-batch
// conflicts, a hit, the end of the block) goes through the general path, which resolves the batch exactly like v1 and
// repairs the slots the eager writes touched.
namespace v2 {

constexpr uint32_t POS_MASK = 0x1ffffu;
constexpr uint32_t VALID = 0x20000u;
constexpr uint32_t HI_MASK = 0xfffe0000u;  // valid bit + tag
constexpr uint32_t TAG_MASK = 0xfffc0000u;
constexpr uint32_t TABLE_WORDS = 4096;
constexpr uint32_t SWEEP_TRIGGER = 49152;  // sweep when a probe would lie this far beyond the last sweep position
constexpr uint32_t MAX_N = 16u << 20;      // probe batches span < 32 KiB up to here (step <= 725)
constexpr uint32_t DEFER_MIN = 256;        // literal runs from this length on are copied by k_lz4_copy
constexpr uint32_t JOB_PIECE = 65536;      // ... in pieces of at most this many bytes (one warp each)

__device__ __forceinline__ uint32_t entry_of(uint32_t p, uint32_t w32) { return ((w32 * 2246822519u) & TAG_MASK) | (p & POS_MASK) | VALID; }
__device__ __forceinline__ uint32_t hash5(uint64_t w) { return (uint32_t)(((w << 24) * 889523592379ull) >> 52); }
__device__ __forceinline__ uint64_t ld_word(const uint8_t* __restrict__ s, uint32_t p, uint32_t n) { return p < n ? rd64(s, p) : 0ull; }

// invalidate every entry more than 65 535 bytes behind q.  Exact while all valid entries are less than 2^17 bytes behind q.
__device__ __forceinline__ void sweep(uint32_t* table, uint32_t q, uint32_t lane)
{
    uint4* t4 = reinterpret_cast<uint4*>(table);
#pragma unroll 4
    for (uint32_t i = lane; i < TABLE_WORDS / 4; i += 32)
    {
        uint4 v = t4[i];
        if (((q - v.x) & POS_MASK) > LZ4_MAX_DISTANCE) v.x = 0;
        if (((q - v.y) & POS_MASK) > LZ4_MAX_DISTANCE) v.y = 0;
        if (((q - v.z) & POS_MASK) > LZ4_MAX_DISTANCE) v.z = 0;
        if (((q - v.w) & POS_MASK) > LZ4_MAX_DISTANCE) v.w = 0;
        t4[i] = v;
    }
}
// Invariant: every valid entry lies at or after sweep_base - 65535.  `top` = the highest position inserted so far.
__device__ __noinline__ uint32_t sweep_to(uint32_t* table, uint32_t sweep_base, uint32_t q, uint32_t top, uint32_t lane)
{
    __syncwarp();
    if (q - sweep_base > 65536u) // only after a match longer than 16 KiB
    {
        if (q > top + LZ4_MAX_DISTANCE)
        {
            uint4* t4 = reinterpret_cast<uint4*>(table);
            for (uint32_t i = lane; i < TABLE_WORDS / 4; i += 32) t4[i] = make_uint4(0u, 0u, 0u, 0u);
            __syncwarp();
            return q;
        }
        while (q - sweep_base > 65536u)
        {
            sweep(table, sweep_base + 65536u, lane);
            sweep_base += 65536u;
        }
    }
    sweep(table, q, lane);
    __syncwarp();
    return q;
}

// lane l: fx = the words at ip0 + 4l and m0 + 4l xor-ed (lane 0 = the 4 bytes of the match itself, lanes 1.. = LZ4_count's first
// 124 bytes, lz4.c:1181); beq = the bytes l + 1 before both positions are equal (catch-up, lz4.c:1104-1109).  One round trip.
__device__ __forceinline__ void load_match_words(const uint8_t* __restrict__ src, uint32_t ip0, uint32_t m0, uint32_t anchor, uint32_t matchlimit,
                                                 uint32_t lane, uint32_t& fx, bool& beq)
{
    const uint32_t pa = ip0 + 4u * lane;
    fx = 0xffffffffu;
    if (lane == 0 || pa < matchlimit) fx = rd32(src, pa) ^ rd32(src, m0 + 4u * lane);
    beq = false;
    if (ip0 > anchor + lane && m0 > lane) beq = src[ip0 - 1u - lane] == src[m0 - 1u - lane];
}

// dst[0..n) = src[0..n) by one warp: 16-byte stores assembled from two aligned 16-byte loads of the (arbitrarily aligned) source
template <uint32_t WS>
__device__ __forceinline__ uint4 shift_words(const uint4 a, const uint4 b, uint32_t sh)
{
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    return make_uint4(__funnelshift_r(w[WS], w[WS + 1], sh), __funnelshift_r(w[WS + 1], w[WS + 2], sh), __funnelshift_r(w[WS + 2], w[WS + 3], sh),
                      __funnelshift_r(w[WS + 3], w[WS + 4], sh));
}
template <uint32_t WS>
__device__ __forceinline__ void warp_copy_vecs(uint4* __restrict__ d4, const uint4* __restrict__ sv, uint32_t vecs, uint32_t sh, uint32_t lane)
{
#pragma unroll 4
    for (uint32_t v = lane; v < vecs; v += 32) d4[v] = shift_words<WS>(sv[v], sv[v + 1], sh);
}
__device__ __forceinline__ void warp_copy(uint8_t* __restrict__ d, const uint8_t* __restrict__ s, uint32_t n, uint32_t lane)
{
    const uint32_t head = min(n, (uint32_t)((16u - ((uintptr_t)d & 15u)) & 15u));
    if (lane < head) d[lane] = s[lane];
    const uint8_t* s2 = s + head;
    uint4* d4 = reinterpret_cast<uint4*>(d + head);
    // a vector reads the 32 aligned source bytes around it: stay that far clear of the end
    const uint32_t vecs = n - head >= 48u ? (n - head - 32u) >> 4 : 0u;
    const uint32_t mis = (uint32_t)((uintptr_t)s2 & 15u), sh = (mis & 3u) * 8u;
    const uint4* sv = reinterpret_cast<const uint4*>(s2 - mis);
    switch (mis >> 2)
    {
    case 0: warp_copy_vecs<0>(d4, sv, vecs, sh, lane); break;
    case 1: warp_copy_vecs<1>(d4, sv, vecs, sh, lane); break;
    case 2: warp_copy_vecs<2>(d4, sv, vecs, sh, lane); break;
    default: warp_copy_vecs<3>(d4, sv, vecs, sh, lane); break;
    }
    for (uint32_t i = head + vecs * 16u + lane; i < n; i += 32) d[i] = s[i];
}

__device__ __forceinline__ void emit_runs(uint8_t* __restrict__ dst, uint32_t op, const uint8_t* __restrict__ src, uint32_t from, uint32_t lit,
                                              uint32_t lane, CopyJobs& cj)
{
    if (lit >= DEFER_MIN && !cj.jobs)
    {
        // in-place layout (the source sits at the end of the destination slot): nothing may be deferred, the output overtakes
        // source bytes more than 64 KiB behind the parse.  The copy runs forward, 512 bytes per round, far below the gap.
        warp_copy(dst + op, src + from, lit, lane);
    }
    else if (lit >= DEFER_MIN)
    {
        const uint32_t pieces = (lit + JOB_PIECE - 1u) / JOB_PIECE;
        for (uint32_t i = lane; i < pieces; i += 32)
        {
            const uint32_t o = i * JOB_PIECE;
            cj.jobs[cj.count + i] = make_uint3(op + o, from + o, min(JOB_PIECE, lit - o));
        }
        cj.count += pieces;
    }
    else if (lit)
    {
        // fewer than 256 bytes: two bytes per lane and round, both loads in flight before the stores
        const uint8_t* s = src + from;
        uint8_t* d = dst + op;
        for (uint32_t i = lane; i < lit; i += 64)
        {
            const bool two = i + 32u < lit;
            const uint8_t a = s[i];
            uint8_t b = 0;
            if (two) b = s[i + 32u];
            d[i] = a;
            if (two) d[i + 32u] = b;
        }
    }
}

// the three aligned words around position p, requested now and shifted into place only when they are used (two batches later):
// the funnel shift would otherwise wait for the load right where it is issued
struct Raw
{
    uint32_t a, b, c;
};
__device__ __forceinline__ Raw load_raw(const uint8_t* __restrict__ s, uint32_t p, uint32_t n)
{
    Raw r = {0u, 0u, 0u};
    if (p < n)
    {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(s + (p & ~3u));
        r.a = w[0];
        r.b = w[1];
        r.c = w[2];
    }
    return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ uint32_t encode_block(const uint8_t* __restrict__ src, const uint32_t n, uint8_t* __restrict__ dst, uint32_t* table, const uint32_t lane,
                                 CopyJobs& cj)
{
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t mflimit_plus_one = n - 11, matchlimit = n - 5;
    uint32_t op = 0, anchor = 0;
    {
        // zeroed table = every slot points at position 0 (lz4.c:1004-1010); inserting position 0 itself changes nothing
        const uint32_t init = entry_of(0u, rd32(src, 0));
        uint4* t4 = reinterpret_cast<uint4*>(table);
        for (uint32_t i = lane; i < TABLE_WORDS / 4; i += 32) t4[i] = make_uint4(init, init, init, init);
    }
    __syncwarp();
    uint32_t sweep_base = 0;
    uint32_t S = 1; // start of the current search
    bool finished = false;
    uint32_t pA = S + lane;
    Raw rawA = load_raw(src, pA, n);
    for (;;)
    {
        // ---------------- search (lz4.c:1043-1100): 32 probes per batch — lane l takes probe k0 + l of the search that started at S — with
        // the source words two batches ahead in flight.  The first 65 probes of a search advance by one byte (`lean`): a batch is 32
        // consecutive positions from `base`, no closed forms needed; most searches of compressible data end there.
        uint32_t k0 = 0;
        uint32_t base = S; // lean: position of lane 0's probe
        bool lean = true;
        uint32_t pB = S + 32u + lane;
        Raw rawB = load_raw(src, pB, n);
        uint32_t ip0 = 0, m0 = 0, fx = 0;
        bool beq = false;
        for (;;)
        {
            if (lean && k0 + 31u > 64u) // the batch leaves the step-1 stretch: the words in flight are not its positions
            {
                lean = false;
                pA = probe_pos(S, k0 + lane);
                rawA = load_raw(src, pA, n);
                pB = probe_pos(S, k0 + 32u + lane);
                rawB = load_raw(src, pB, n);
            }
            const uint32_t p = pA;
            const uint32_t sh = (p & 3u) * 8u;
            const uint32_t lo = __funnelshift_r(rawA.a, rawA.b, sh);
            const uint32_t hi = __funnelshift_r(rawA.b, rawA.c, sh);
            pA = pB;
            rawA = rawB;
            uint32_t pf, p_last, adv_last;
            if (lean)
            {
                pB = pA + 32u; // (re-derived when the search outlives the stretch)
                rawB = load_raw(src, pB, n);
                if (lane == 0 && base + 4096u < n) prefetch_l2(src + base + 4096u); // the stream ahead: HBM -> L2
                pf = base;
                p_last = base + 31u;
                adv_last = 1u;
            }
            else
            {
                pB = probe_pos(S, k0 + 64u + lane);
                rawB = load_raw(src, pB, n);
                const uint32_t far = pB + 2048u + 8u * (pB - pA);
                if (far < n) prefetch_l2(src + far);
                pf = probe_pos(S, k0);
                p_last = probe_pos(S, k0 + 31u);
                adv_last = probe_advance(k0 + 31u);
            }
            if (p_last >= sweep_base + SWEEP_TRIGGER) sweep_base = sweep_to(table, sweep_base, pf, pf, lane);
            const bool all_valid = p_last + adv_last <= mflimit_plus_one; // else: goto _last_literals inside this batch
            const uint32_t h = hash5((uint64_t)lo | ((uint64_t)hi << 32));
            const uint32_t mine = entry_of(p, lo);
            uint32_t e = 0, r = 0;
            bool valid = true;
            if (all_valid)
            {
                e = table[h];
                __syncwarp();
                table[h] = mine;
                __syncwarp();
                r = table[h];
                const bool pot = ((e ^ mine) & HI_MASK) == 0u && ((p - e) & POS_MASK) <= LZ4_MAX_DISTANCE;
                if (!__any_sync(FULL, pot || r != mine))
                {
                    k0 += 32;
                    base += 32;
                    continue;
                }
            }
            else
            {
                valid = p < n && p + probe_advance(k0 + lane) <= mflimit_plus_one;
                if (valid) e = table[h];
                __syncwarp();
            }
            // ---- short matches resolved inside the batch.  In a step-1 batch (the first 64 probes of a search: lane l probes pf + l) whose
            // 32 hashes are distinct, every lane's table entry `e` is exact whatever happens to the other lanes — no lane of the batch
            // touches its slot.  So all tag matches of the batch can be verified in ONE round trip (each lane compares the 8 bytes at its
            // candidate with its own), and a match of 4..7 bytes that ends inside the batch is followed by exactly what the batch already
            // did: the insertion of ip-2 and the probe of ip (lz4.c:1236-1294) are the eager writes / entries of lanes f+L-2 and f+L, and
            // the next search (step 1 again) continues with lane f+L+1.  The warp therefore walks the verified matches in order, emits
            // their sequences, takes back the writes of the positions the matches skip, and goes on with the next batch as probes
            // 32 - c0 .. of the search that started at lane c0 — no per-match round trip, no general path, no restart.  Anything else (a
            // longer match, a match that extends backwards, one that ends beyond the batch) is handed to the general path at that lane.
            if (all_valid && lean && pf + 48u <= matchlimit && __ballot_sync(FULL, r != mine) == 0u) // step 1: lane l probes pf + l
            {
                const uint32_t age = (p - e) & POS_MASK;
                const uint32_t cand0 = p - age;
                const bool pot0 = ((e ^ mine) & HI_MASK) == 0u && age <= LZ4_MAX_DISTANCE;
                uint32_t mlen = 0;
                bool backs = false;
                if (pot0)
                {
                    const uint64_t cw = rd64(src, cand0);
                    const uint32_t x_lo = lo ^ (uint32_t)cw, x_hi = hi ^ (uint32_t)(cw >> 32);
                    if (x_lo == 0u) mlen = 4u + (x_hi ? (uint32_t)(__ffs(x_hi) - 1) >> 3 : 4u); // 8 = all eight bytes equal: longer than this path takes
                    backs = cand0 > 0u && src[p - 1u] == src[cand0 - 1u];
                }
                const uint32_t V = __ballot_sync(FULL, mlen != 0u);
                if (V == 0u) // tag matches by chance only: the batch stands as it is
                {
                    k0 += 32;
                    base += 32;
                    continue;
                }
                const uint32_t LONGM = __ballot_sync(FULL, mlen == 8u), BACKM = __ballot_sync(FULL, backs);
                uint32_t cancel = 0, cur = 0, handoff = 32, prev_end = 0xffffffffu;
                for (;;)
                {
                    const uint32_t m = cur < 32u ? V & ~((1u << cur) - 1u) : 0u;
                    if (!m) break;
                    const uint32_t f = (uint32_t)__ffs(m) - 1u;
                    const uint32_t L = __shfl_sync(FULL, mlen, f);
                    const bool chained = f == prev_end; // the post-match test of the previous match hit: no literals, no catch-up (lz4.c:1262-1294)
                    if (((LONGM >> f) & 1u) || (((BACKM >> f) & 1u) && !chained) || f + L > 31u)
                    {
                        handoff = f;
                        break;
                    }
                    const uint32_t pfm = pf + f, mf = __shfl_sync(FULL, cand0, f);
                    const uint32_t lit = pfm - anchor;
                    const uint32_t token_pos = op++;
                    if (lit >= 15) op = put_length(dst, op, lit - 15, lane);
                    emit_runs(dst, op, src, anchor, lit, lane, cj);
                    op += lit;
                    if (lane == 0)
                    {
                        const uint32_t off = pfm - mf;
                        dst[op] = (uint8_t)off;
                        dst[op + 1] = (uint8_t)(off >> 8);
                        dst[token_pos] = (uint8_t)((lit >= 15 ? 15u : lit) << 4 | (L - 4u));
                    }
                    op += 2;
                    anchor = pfm + L;
                    // positions f+1 .. f+L-1 are inside the match: only ip-2 = f+L-2 is inserted
                    cancel |= (((1u << L) - 1u) << f) & ~(1u << f) & ~(1u << (f + L - 2u));
                    cur = prev_end = f + L;
                }
                if (handoff == 32u)
                {
                    // every match of the batch is out.  The next search started at lane c0 = prev_end + 1 and lanes c0 .. 31 were its first
                    // 32 - c0 probes (none of them hit): it simply goes on with the next batch, whose positions — consecutive as before —
                    // are the ones already in flight
                    if ((cancel >> lane) & 1u) table[h] = e;
                    __syncwarp();
                    const uint32_t c0 = prev_end + 1u;
                    S = pf + c0;
                    k0 = 32u - c0;
                    base = pf + 32u;
                    continue;
                }
                // hand the match at lane `handoff` to the general sequence code: the table as the sequential loop leaves it there
                {
                    const uint32_t undo = cancel | (handoff < 31u ? ~((2u << handoff) - 1u) : 0u);
                    if ((undo >> lane) & 1u) table[h] = e;
                    __syncwarp();
                    ip0 = pf + handoff;
                    m0 = __shfl_sync(FULL, cand0, handoff);
                    load_match_words(src, ip0, m0, anchor, matchlimit, lane, fx, beq);
                    break;
                }
            }
            // ---- general path.  A lane's candidate is the closest lower lane of this batch with the same hash, else the table entry
            // it read before the batch wrote anything.  `same` = the lanes with this lane's hash: found from the eager writes (a lane
            // that did not read its own entry back shares its slot), MATCH.ANY only for the batch at the end of the block
            uint32_t same = 1u << lane;
            if (all_valid)
            {
                uint32_t confl = __ballot_sync(FULL, r != mine);
                while (confl)
                {
                    const uint32_t hc = __shfl_sync(FULL, h, __ffs(confl) - 1);
                    const uint32_t g = __ballot_sync(FULL, h == hc);
                    if (h == hc) same = g;
                    confl &= ~g;
                }
            }
            else
                same = __match_any_sync(FULL, valid ? h : 0x10000u + lane);
            const uint32_t lower = same & lt_mask;
            const int from = lower ? 31 - __clz(lower) : (int)lane;
            uint32_t lo_from = lo, p_from = p;
            if (__any_sync(FULL, lower != 0u))
            {
                lo_from = __shfl_sync(FULL, lo, from);
                p_from = __shfl_sync(FULL, p, from);
            }
            uint32_t cand;
            bool pot;
            if (lower)
            {
                cand = p_from;
                pot = valid && lo_from == lo && p - cand <= LZ4_MAX_DISTANCE;
            }
            else
            {
                const uint32_t age = (p - e) & POS_MASK;
                cand = p - age;
                pot = valid && ((e ^ mine) & HI_MASK) == 0u && age <= LZ4_MAX_DISTANCE;
            }
            uint32_t pots = __ballot_sync(FULL, pot);
            const uint32_t valids = __ballot_sync(FULL, valid);
            uint32_t first_hit = 32;
            while (pots) // the first candidate that really starts with the same 4 bytes (lane 0 of fx compares them)
            {
                const uint32_t f = (uint32_t)__ffs(pots) - 1u;
                ip0 = __shfl_sync(FULL, p, f);
                m0 = __shfl_sync(FULL, cand, f);
                load_match_words(src, ip0, m0, anchor, matchlimit, lane, fx, beq);
                if ((__ballot_sync(FULL, fx != 0u) & 1u) == 0u)
                {
                    first_hit = f;
                    break;
                }
                pots &= pots - 1u;
            }
            // table: the writes of the valid lanes up to and including the first hit are committed (the last writer of a hash value
            // wins); a slot only lanes beyond the hit wrote eagerly gets its old content back
            const uint32_t commit = valids & (first_hit >= 31u ? FULL : ((2u << first_hit) - 1u));
            const uint32_t group = same & commit;
            if (valid)
            {
                if (group)
                {
                    if ((uint32_t)(31 - __clz(group)) == lane) table[h] = mine;
                }
                else if (all_valid && (uint32_t)__ffs(same) - 1u == lane)
                    table[h] = e;
            }
            __syncwarp();
            if (first_hit < 32) break;
            if (valids != FULL)
            {
                finished = true;
                break;
            }
            k0 += 32;
            base += 32;
        }
        if (finished) break;

        // ---------------- catch up (lz4.c:1104-1109)
        uint32_t top = ip0;
        uint32_t back;
        {
            const uint32_t m = __ballot_sync(FULL, beq);
            back = m == FULL ? 32u : (uint32_t)__ffs(~m) - 1u;
        }
        if (back == 32u)
            for (;;)
            {
                bool eq = false;
                const uint32_t d = back + lane;
                if (ip0 > anchor + d && m0 > d) eq = src[ip0 - 1u - d] == src[m0 - 1u - d];
                const uint32_t m = __ballot_sync(FULL, eq);
                const uint32_t run = m == FULL ? 32u : (uint32_t)__ffs(~m) - 1u;
                back += run;
                if (run < 32) break;
            }
        uint32_t ip = ip0 - back, match = m0 - back; // where the sequence's match starts; ip0 / m0 / fx describe its forward part
        uint32_t extra = back;                        // matched bytes in front of ip0
        for (;;) // _next_match (lz4.c:1140-1296)
        {
            // match length beyond MINMATCH: LZ4_count(ip0+4, m0+4, matchlimit) — lanes 1..31 hold the first 124 bytes in fx
            uint32_t fwd;
            {
                const uint32_t pa = ip0 + 4u * lane;
                uint32_t eqb = 4;
                bool stop = false;
                if (lane)
                {
                    eqb = 0;
                    if (pa < matchlimit)
                    {
                        eqb = fx ? (uint32_t)(__ffs(fx) - 1) >> 3 : 4u;
                        const uint32_t room = matchlimit - pa;
                        if (eqb > room) eqb = room;
                    }
                    stop = eqb < 4u;
                }
                const uint32_t stops = __ballot_sync(FULL, stop);
                if (stops)
                {
                    const uint32_t f = (uint32_t)__ffs(stops) - 1u;
                    fwd = 4u * (f - 1u) + __shfl_sync(FULL, eqb, f);
                }
                else
                {
                    fwd = 124;
                    uint32_t a = ip0 + 128, b = m0 + 128;
                    for (;;)
                    {
                        const uint32_t pa2 = a + 4u * lane;
                        uint32_t eq2 = 0;
                        bool stop2 = true;
                        if (pa2 < matchlimit)
                        {
                            const uint32_t x = rd32(src, pa2) ^ rd32(src, b + 4u * lane);
                            eq2 = x ? (uint32_t)(__ffs(x) - 1) >> 3 : 4u;
                            const uint32_t room = matchlimit - pa2;
                            if (eq2 > room) eq2 = room;
                            stop2 = eq2 < 4u;
                        }
                        const uint32_t stops2 = __ballot_sync(FULL, stop2);
                        if (stops2)
                        {
                            const uint32_t f = (uint32_t)__ffs(stops2) - 1u;
                            fwd += 4u * f + __shfl_sync(FULL, eq2, f);
                            break;
                        }
                        fwd += 128;
                        a += 128;
                        b += 128;
                    }
                }
            }
            // token, literals, offset, match length (lz4.c:1112-1226); a chained match (post-match test below) has no literals
            const uint32_t code = extra + fwd;
            const uint32_t lit = ip - anchor;
            const uint32_t token_pos = op++;
            if (lit >= 15) op = put_length(dst, op, lit - 15, lane);
            emit_runs(dst, op, src, anchor, lit, lane, cj);
            op += lit;
            const uint32_t off = ip - match;
            if (lane == 0)
            {
                dst[op] = (uint8_t)off; // LZ4_writeLE16 (lz4.c:1157-1163)
                dst[op + 1] = (uint8_t)(off >> 8);
                dst[token_pos] = (uint8_t)((lit >= 15 ? 15u : lit) << 4 | (code >= 15 ? 15u : code));
            }
            op += 2;
            if (code >= 15) op = put_length(dst, op, code - 15, lane);
            ip = ip0 + 4u + fwd;
            anchor = ip;
            if (ip >= mflimit_plus_one) // lz4.c:1233
            {
                finished = true;
                break;
            }
            // fill the table with ip-2, then test ip itself (lz4.c:1236-1294); every lane does the same (no broadcast needed).  The
            // words of the next search's first batch are requested in the same round trip.
            if (ip >= sweep_base + SWEEP_TRIGGER) sweep_base = sweep_to(table, sweep_base, ip - 2u, top, lane);
            const uint64_t w2 = rd64(src, ip - 2u), w0 = rd64(src, ip);
            S = ip + 1u;
            pA = S + lane;
            rawA = load_raw(src, pA, n);
            const uint32_t mine0 = entry_of(ip, (uint32_t)w0);
            const uint32_t h0 = hash5(w0);
            table[hash5(w2)] = entry_of(ip - 2u, (uint32_t)w2);
            __syncwarp();
            const uint32_t e = table[h0];
            __syncwarp();
            table[h0] = mine0;
            __syncwarp();
            const uint32_t age = (ip - e) & POS_MASK;
            if (((e ^ mine0) & HI_MASK) == 0u && age <= LZ4_MAX_DISTANCE)
            {
                const uint32_t cand = ip - age;
                load_match_words(src, ip, cand, ip, matchlimit, lane, fx, beq);
                if (__shfl_sync(FULL, fx, 0) == 0u)
                {
                    ip0 = ip;
                    m0 = cand;
                    match = cand;
                    extra = 0;
                    top = ip;
                    continue; // token = 0 literals, straight to the next match
                }
            }
            break;
        }
        if (finished) break;
        // lz4.c:1298: forwardH = hash(++ip), a fresh search (step 1, searchMatchNb reset) from S = ip + 1
    }
    // last literals (lz4.c:1302-1329)
    const uint32_t last = n - anchor;
    const uint32_t token_pos = op++;
    if (lane == 0) dst[token_pos] = (uint8_t)((last >= 15 ? 15u : last) << 4);
    if (last >= 15) op = put_length(dst, op, last - 15, lane);
    emit_runs(dst, op, src, anchor, last, lane, cj);
    op += last;
    return op;
}

} // namespace v2

} // namespace

// one warp (= one CTA) per block.  dst layout per block: [u32 raw_size][u32 compressed_size][codec bytes]  — the payload
// header compressblockstore writes (lib/compressblockstore/longtail_compressblockstore.c:103-137).
// v1 kernel: blocks the v2 kernel does not take (byU16 blocks below 64 KiB, blocks above 16 MiB), or every block when v2 is off.
__device__ __forceinline__ bool lz4_v2_takes(uint32_t n) { return n >= LZ4_64K_LIMIT && n <= v2::MAX_N; }

__global__ void __launch_bounds__(32)
k_lz4_blocks(const uint8_t* __restrict__ raw_base, const uint64_t* __restrict__ raw_off, const uint32_t* __restrict__ raw_len,
             uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, uint32_t* __restrict__ out_len,
             uint3* __restrict__ copy_jobs, const uint32_t* __restrict__ copy_job_start, uint32_t* __restrict__ copy_job_count,
             uint32_t block_count, uint32_t* __restrict__ g_tables, uint32_t v2_on)
{
    extern __shared__ __align__(16) uint32_t s_shared_table[];
    const uint32_t b = blockIdx.x;
    if (b >= block_count) return;
    const uint32_t n = raw_len[b];
    if (v2_on && lz4_v2_takes(n)) return;
    // the hash table lives in HBM/L2 when the caller provides room (32 instead of 7 resident warps per SM), else in shared memory
    uint32_t* s_table = g_tables ? g_tables + (size_t)b * (LZ4_TABLE_BYTES / 4) : s_shared_table;
    const uint32_t lane = threadIdx.x;
    const uint8_t* src = raw_base + raw_off[b];
    uint8_t* dst = out_base + out_off[b];
    CopyJobs cj = {copy_jobs + copy_job_start[b], 0};
    uint32_t c = n < LZ4_64K_LIMIT ? lz4_encode_block<true>(src, n, dst + 8, s_table, lane, cj)
                                   : lz4_encode_block<false>(src, n, dst + 8, s_table, lane, cj);
    if (lane == 0)
    {
        reinterpret_cast<uint32_t*>(dst)[0] = n;
        reinterpret_cast<uint32_t*>(dst)[1] = c;
        out_len[b] = c + 8;
        copy_job_count[b] = cj.count;
    }
}

// v2 kernel: persistent warps (one per CTA, 13 per SM: 16 KiB of table + 1 KiB of system shared memory each) pull blocks from a queue
__global__ void __launch_bounds__(32)
k_lz4_blocks_v2(const uint8_t* __restrict__ raw_base, const uint64_t* __restrict__ raw_off, const uint32_t* __restrict__ raw_len,
                uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, uint32_t* __restrict__ out_len,
                uint3* __restrict__ copy_jobs, const uint32_t* __restrict__ copy_job_start, uint32_t* __restrict__ copy_job_count,
                uint32_t block_count, uint32_t* __restrict__ queue)
{
    extern __shared__ __align__(16) uint32_t s_shared_table[];
    const uint32_t lane = threadIdx.x;
    for (;;)
    {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(queue, 1u);
        b = __shfl_sync(FULL, b, 0);
        if (b >= block_count) return;
        const uint32_t n = raw_len[b];
        if (!lz4_v2_takes(n)) continue;
        const uint8_t* src = raw_base + raw_off[b];
        uint8_t* dst = out_base + out_off[b];
        const uint32_t job_start = copy_job_start[b];
        CopyJobs cj = {job_start == LZ4_JOBS_INLINE ? nullptr : copy_jobs + job_start, 0};
        const uint32_t c = v2::encode_block(src, n, dst + 8, s_shared_table, lane, cj);
        if (lane == 0)
        {
            reinterpret_cast<uint32_t*>(dst)[0] = n;
            reinterpret_cast<uint32_t*>(dst)[1] = c;
            out_len[b] = c + 8;
            copy_job_count[b] = cj.count;
        }
        __syncwarp();
    }
}

// the deferred literal runs of the encoders: grid = (block, slice).  A block with many jobs (v2: pieces of <= 64 KiB) hands them to
// the warps of its CTAs round robin; a block with a few long jobs (v1: the whole incompressible block) is cut into slices instead.
constexpr uint32_t LZ4_COPY_SLICES = 8;
constexpr uint32_t LZ4_COPY_WARPS = 8;
__global__ void __launch_bounds__(LZ4_COPY_WARPS * 32)
k_lz4_copy(const uint8_t* __restrict__ raw_base, const uint64_t* __restrict__ raw_off, uint8_t* __restrict__ out_base,
           const uint64_t* __restrict__ out_off, const uint3* __restrict__ copy_jobs, const uint32_t* __restrict__ copy_job_start,
           const uint32_t* __restrict__ copy_job_count)
{
    const uint32_t b = blockIdx.x;
    const uint32_t njobs = copy_job_count[b];
    const uint3* jobs = copy_jobs + copy_job_start[b];
    const uint8_t* src = raw_base + raw_off[b];
    uint8_t* dst = out_base + out_off[b] + 8;
    if (njobs >= 2 * LZ4_COPY_SLICES * LZ4_COPY_WARPS)
    {
        const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
        for (uint32_t j = blockIdx.y * LZ4_COPY_WARPS + warp; j < njobs; j += LZ4_COPY_SLICES * LZ4_COPY_WARPS)
        {
            const uint3 job = jobs[j];
            v2::warp_copy(dst + job.x, src + job.y, job.z, lane);
        }
        return;
    }
    for (uint32_t j = 0; j < njobs; ++j)
    {
        const uint3 job = jobs[j];
        // slice boundaries on 16-byte multiples of the job
        const uint32_t per = ((job.z + LZ4_COPY_SLICES - 1) / LZ4_COPY_SLICES + 15u) & ~15u;
        const uint32_t lo = min(job.z, per * blockIdx.y), hi = min(job.z, lo + per);
        if (lo >= hi) continue;
        uint8_t* d = dst + job.x + lo;
        const uint8_t* s = src + job.y + lo;
        const uint32_t n = hi - lo;
        const uint32_t head = min(n, (uint32_t)((16u - ((uintptr_t)d & 15u)) & 15u));
        for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) d[i] = s[i];
        const uint8_t* s2 = s + head;
        uint4* d4 = reinterpret_cast<uint4*>(d + head);
        const uint32_t vecs = n - head >= 20u ? (n - head - 4u) >> 4 : 0u;
        const uint32_t sh = (uint32_t)((uintptr_t)s2 & 3u) * 8u;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(s2 - ((uintptr_t)s2 & 3u));
#pragma unroll 4
        for (uint32_t v = threadIdx.x; v < vecs; v += blockDim.x)
        {
            const uint32_t* w = sw + 4 * v;
            const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
            d4[v] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
        }
        for (uint32_t i = head + vecs * 16u + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
    }
}

// payload gather (WriteContentBlockJob, src/longtail.c:4640-4721): block payload = its chunks' bytes, back to back.
// One CTA per chunk copy job.
__global__ void __launch_bounds__(256)
k_gather_chunks(const uint8_t* __restrict__ arena, const uint64_t* __restrict__ src_off, const uint64_t* __restrict__ dst_off,
                const uint32_t* __restrict__ len, uint8_t* __restrict__ out, uint32_t count)
{
    const uint32_t c = blockIdx.x;
    if (c >= count) return;
    const uint8_t* s = arena + src_off[c];
    uint8_t* d = out + dst_off[c];
    const uint32_t n = len[c];
    // align the destination to 16 bytes, then move 16-byte vectors assembled from the (arbitrarily aligned) source
    const uint32_t head = min(n, (uint32_t)((16u - ((uintptr_t)d & 15u)) & 15u));
    for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) d[i] = s[i];
    const uint8_t* s2 = s + head;
    uint8_t* d2 = d + head;
    const uint32_t vecs = n - head >= 20u ? (n - head - 4u) >> 4 : 0u; // the funnel reads one word ahead: stay clear of the end
    const uint32_t sh = (uint32_t)((uintptr_t)s2 & 3u) * 8u;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s2 - ((uintptr_t)s2 & 3u));
    for (uint32_t v = threadIdx.x; v < vecs; v += blockDim.x)
    {
        const uint32_t* w = sw + 4 * v;
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
        uint4 o;
        if (sh)
        {
            const uint32_t w4 = w[4];
            o = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
        }
        else
            o = make_uint4(w0, w1, w2, w3);
        reinterpret_cast<uint4*>(d2)[v] = o;
    }
    for (uint32_t i = head + vecs * 16u + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

// The block index in front of every payload (Longtail_WriteStoredBlockToBuffer, src/longtail.c:4111-4150; layout of
// Longtail_BlockIndex's data, :3585-3637): u64 block hash, u32 hash id, u32 chunk count, u32 tag, u64 chunk hashes[], u32 chunk sizes[].
// One CTA per block; the image starts at a 4-byte aligned offset, so 64-bit fields go out as two words.
__global__ void __launch_bounds__(128)
k_block_headers(uint8_t* __restrict__ out_base, const uint64_t* __restrict__ img_off, const uint32_t* __restrict__ blk_first,
                const uint32_t* __restrict__ blk_count, const uint32_t* __restrict__ blk_tag, const uint64_t* __restrict__ blk_hash,
                const uint64_t* __restrict__ chunk_hashes, const uint32_t* __restrict__ chunk_sizes, uint32_t hash_type, uint32_t nb)
{
    const uint32_t b = blockIdx.x;
    if (b >= nb) return;
    uint32_t* dst = reinterpret_cast<uint32_t*>(out_base + img_off[b]);
    const uint32_t first = blk_first[b], count = blk_count[b];
    if (threadIdx.x == 0)
    {
        const uint64_t h = blk_hash[b];
        dst[0] = (uint32_t)h;
        dst[1] = (uint32_t)(h >> 32);
        dst[2] = hash_type;
        dst[3] = count;
        dst[4] = blk_tag[b];
    }
    for (uint32_t i = threadIdx.x; i < count; i += blockDim.x)
    {
        const uint64_t h = chunk_hashes[first + i];
        dst[5 + 2 * i] = (uint32_t)h;
        dst[6 + 2 * i] = (uint32_t)(h >> 32);
        dst[5 + 2 * count + i] = chunk_sizes[first + i];
    }
}

void launch_block_headers(uint8_t* d_out, const uint64_t* d_img_off, const uint32_t* d_blk_first, const uint32_t* d_blk_count,
                          const uint32_t* d_blk_tag, const uint64_t* d_blk_hash, const uint64_t* d_chunk_hashes, const uint32_t* d_chunk_sizes,
                          uint32_t hash_type, uint32_t nb, cudaStream_t st)
{
    if (!nb) return;
    k_block_headers<<<nb, 128, 0, st>>>(d_out, d_img_off, d_blk_first, d_blk_count, d_blk_tag, d_blk_hash, d_chunk_hashes, d_chunk_sizes, hash_type, nb);
}

// LZ4 block decoder (LZ4_decompress_safe semantics, lib/lz4/longtail_lz4.c:79-101): one warp per block, the token stream is
// walked by all lanes in lock step (uniform control flow), literal and match bytes are moved lane-parallel.  An overlapping
// match (offset < length) repeats its first `offset` bytes, so byte i of the match is byte (i mod offset) of that period.
// out_len[b] = decoded size, or 0xffffffff for a malformed / overflowing stream (the caller maps it to EBADF).
__global__ void __launch_bounds__(32)
k_lz4_decode(const uint8_t* __restrict__ in_base, const uint64_t* __restrict__ in_off, const uint32_t* __restrict__ in_len,
             uint8_t* __restrict__ out_base, const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ out_cap,
             uint32_t* __restrict__ out_len, uint32_t block_count)
{
    const uint32_t b = blockIdx.x;
    if (b >= block_count) return;
    const uint32_t lane = threadIdx.x;
    const uint8_t* src = in_base + in_off[b];
    const uint32_t n = in_len[b];
    uint8_t* dst = out_base + out_off[b];
    const uint32_t cap = out_cap[b];
    uint32_t ip = 0, op = 0;
    bool bad = n == 0;
    while (!bad && ip < n)
    {
        const uint32_t token = src[ip++];
        uint32_t lit = token >> 4;
        if (lit == 15)
        {
            uint32_t v;
            do
            {
                if (ip >= n) { bad = true; break; }
                v = src[ip++];
                lit += v;
            } while (v == 255);
            if (bad) break;
        }
        if (lit > n - ip || lit > cap - op) { bad = true; break; }
        for (uint32_t i = lane; i < lit; i += 32) dst[op + i] = src[ip + i];
        ip += lit;
        op += lit;
        if (ip >= n) break; // the last sequence carries literals only
        if (n - ip < 2) { bad = true; break; }
        const uint32_t off = (uint32_t)src[ip] | ((uint32_t)src[ip + 1] << 8);
        ip += 2;
        uint32_t len = token & 15u;
        if (len == 15)
        {
            uint32_t v;
            do
            {
                if (ip >= n) { bad = true; break; }
                v = src[ip++];
                len += v;
            } while (v == 255);
            if (bad) break;
        }
        len += 4;
        if (off == 0 || off > op || len > cap - op) { bad = true; break; }
        __syncwarp(); // the literals just written may be the source of this match
        const uint32_t from = op - off;
        if (off >= len)
            for (uint32_t i = lane; i < len; i += 32) dst[op + i] = dst[from + i];
        else
            for (uint32_t i = lane; i < len; i += 32) dst[op + i] = dst[from + i % off];
        op += len;
        __syncwarp();
    }
    if (lane == 0) out_len[b] = bad ? 0xffffffffu : op;
}

cudaError_t launch_lz4_decode(const uint8_t* d_in, const uint64_t* d_in_off, const uint32_t* d_in_len, uint8_t* d_out,
                              const uint64_t* d_out_off, const uint32_t* d_out_cap, uint32_t* d_out_len, uint32_t block_count, cudaStream_t st)
{
    if (!block_count) return cudaSuccess;
    k_lz4_decode<<<block_count, 32, 0, st>>>(d_in, d_in_off, d_in_len, d_out, d_out_off, d_out_cap, d_out_len, block_count);
    return cudaGetLastError();
}

uint32_t lz4_copy_job_capacity(uint32_t raw_len) { return raw_len / v2::DEFER_MIN + 2; }

static bool lz4_v2_enabled()
{
    const char* e = getenv("LT_B200_LZ4_V1");
    return !(e && e[0] == '1');
}
bool lz4_block_in_place(uint32_t raw_len) { return lz4_v2_enabled() && raw_len >= LZ4_64K_LIMIT && raw_len <= v2::MAX_N; }
uint64_t lz4_in_place_offset(uint32_t raw_len) { return ((uint64_t)raw_len / 255 + 16 + 65536 + 256 + 15) & ~15ull; }

size_t lz4_v2_scratch_bytes() { return 256; }

cudaError_t launch_lz4_blocks(const uint8_t* d_raw, const uint64_t* d_raw_off, const uint32_t* d_raw_len, uint8_t* d_out,
                              const uint64_t* d_out_off, uint32_t* d_out_len, uint3* d_copy_jobs, const uint32_t* d_copy_job_start,
                              uint32_t* d_copy_job_count, uint32_t block_count, uint32_t* d_tables, void* d_v2_scratch, int sm_count,
                              cudaStream_t st)
{
    if (!block_count) return cudaSuccess;
    static int v2_on = -1, smem_per_sm = 13;
    if (v2_on < 0)
    {
        const char* e = getenv("LT_B200_LZ4_V1"); // A/B knob: 1 = the round-1 encoder (tables in HBM/L2) for every block
        v2_on = e && atoi(e) ? 0 : 1;
        const char* m = getenv("LT_B200_LZ4_SMEM_CTAS"); // A/B knob: shared-memory CTAs per SM (default: all 13 that fit)
        if (m) smem_per_sm = atoi(m) < 1 ? 1 : (atoi(m) > 13 ? 13 : atoi(m));
        // 13 CTAs of 16 KiB + 1 KiB per SM need the largest shared-memory carve-out
        cudaFuncSetAttribute(k_lz4_blocks_v2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    // v1 takes the blocks v2 leaves (byU16, > 16 MiB): d_tables (32 KiB per block, in HBM/L2) lifts its residency from 7 warps per SM to 32
    k_lz4_blocks<<<block_count, 32, d_tables ? 0 : LZ4_TABLE_BYTES, st>>>(d_raw, d_raw_off, d_raw_len, d_out, d_out_off, d_out_len, d_copy_jobs,
                                                                         d_copy_job_start, d_copy_job_count, block_count, d_tables, (uint32_t)v2_on);
    if (v2_on)
    {
        uint32_t* queue = static_cast<uint32_t*>(d_v2_scratch);
        cudaMemsetAsync(queue, 0, 4, st);
        const uint32_t ctas = min(block_count, (uint32_t)(smem_per_sm * sm_count));
        k_lz4_blocks_v2<<<ctas, 32, v2::TABLE_WORDS * 4, st>>>(d_raw, d_raw_off, d_raw_len, d_out, d_out_off, d_out_len, d_copy_jobs, d_copy_job_start,
                                                             d_copy_job_count, block_count, queue);
    }
    k_lz4_copy<<<dim3(block_count, LZ4_COPY_SLICES), LZ4_COPY_WARPS * 32, 0, st>>>(d_raw, d_raw_off, d_out, d_out_off, d_copy_jobs, d_copy_job_start,
                                                                                 d_copy_job_count);
    return cudaGetLastError();
}

void launch_gather_chunks(const uint8_t* d_arena, const uint64_t* d_src_off, const uint64_t* d_dst_off, const uint32_t* d_len,
                          uint8_t* d_out, uint32_t count, cudaStream_t st)
{
    if (!count) return;
    k_gather_chunks<<<count, 256, 0, st>>>(d_arena, d_src_off, d_dst_off, d_len, d_out, count);
}

// host -> device upload by the SMs: the sources are pinned host memory mapped into the device's address space, read over PCIe with
// 16-byte loads.  Unlike cudaMemcpyAsync this does not occupy the host -> device copy engine, whose queue is first in first out across
// streams: the small table uploads of the kernels that run meanwhile would otherwise wait behind gigabytes of asset bytes.
// A small persistent grid (PCIe needs ~100 KiB in flight, not SMs): 64 KiB tiles of all segments dealt round robin to the CTAs, so the
// kernels working on the previous batch keep the machine.  Sources and destinations 16-byte aligned.
constexpr uint32_t UPLOAD_CTAS = 64, UPLOAD_THREADS = 128, UPLOAD_TILE_VECS = 4096;
__global__ void __launch_bounds__(UPLOAD_THREADS)
k_upload_segments(const UploadSeg* __restrict__ segs, uint32_t count)
{
    uint64_t tile_base = 0;
    for (uint32_t i = 0; i < count; ++i)
    {
        const UploadSeg sg = segs[i];
        const uint64_t vecs = sg.len >> 4;
        const uint64_t tiles = (vecs + UPLOAD_TILE_VECS - 1) / UPLOAD_TILE_VECS;
        const uint4* s = reinterpret_cast<const uint4*>(sg.src);
        uint4* d = reinterpret_cast<uint4*>(sg.dst);
        // my first tile of this segment: the smallest t with (tile_base + t) % gridDim.x == blockIdx.x
        uint64_t t = (blockIdx.x + gridDim.x - (uint32_t)(tile_base % gridDim.x)) % gridDim.x;
        for (; t < tiles; t += gridDim.x)
        {
            const uint64_t lo = t * UPLOAD_TILE_VECS, hi = min(vecs, lo + UPLOAD_TILE_VECS);
#pragma unroll 8
            for (uint64_t v = lo + threadIdx.x; v < hi; v += UPLOAD_THREADS) __stcs(d + v, __ldcs(s + v));
        }
        if ((uint32_t)(tile_base % gridDim.x) == blockIdx.x)
            for (uint64_t b = (vecs << 4) + threadIdx.x; b < sg.len; b += UPLOAD_THREADS) sg.dst[b] = sg.src[b];
        tile_base += tiles;
    }
}

void launch_upload_segments(const UploadSeg* d_segs, uint32_t count, cudaStream_t st)
{
    if (!count) return;
    k_upload_segments<<<UPLOAD_CTAS, UPLOAD_THREADS, 0, st>>>(d_segs, count);
}

} // namespace ltb
