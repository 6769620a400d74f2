// blake3.cu — batched BLAKE3 over variable-length byte segments on sm_100a.
//
// Re-design of the HashAPI.HashBuffer call per chunk (lib/blake3/longtail_blake3.c:81-102 ->
// blake3_hasher_update/finalize, ext/blake3.c:462,571) as two kernels over a whole batch of segments:
//
//   k_blake3_leaves  one thread per 1 KiB leaf ("chunk" in BLAKE3 terms): 16 chained 64-byte block compressions
//                    (ext/blake3_portable.c:46-122).  A warp stages the 32 leaves' next 64 bytes with 16-byte
//                    cp.async into shared memory (coalesced across lanes) and every lane reads its own, arbitrarily
//                    aligned, bytes back.
//   k_blake3_merge   one warp per segment: bottom-up pairing of leaf chaining values; an odd node is promoted
//                    unchanged, which reproduces the left-full tree of ext/blake3.c:161-166 exactly.
//
// Output = digest bytes 0..7 as a little-endian u64 (longtail_blake3.c:97-100).
#include "lt_device.cuh"
#include "lt_kernels.h"

namespace ltb {

namespace {

constexpr uint32_t IV0 = 0x6A09E667u, IV1 = 0xBB67AE85u, IV2 = 0x3C6EF372u, IV3 = 0xA54FF53Au;
constexpr uint32_t IV4 = 0x510E527Fu, IV5 = 0x9B05688Cu, IV6 = 0x1F83D9ABu, IV7 = 0x5BE0CD19u; // ext/blake3_impl.h:76-78
constexpr uint32_t F_CHUNK_START = 1, F_CHUNK_END = 2, F_PARENT = 4, F_ROOT = 8;                // ext/blake3_impl.h:13-21

// The G function is 14 integer ops; IADD3 / LOP3 / SHF / PRMT all issue on the ALU pipe, which saturates first
// (profiles/r01b_leaves_ncu.txt: ALU 94.7 %, FMA 11 %).  Every addition is therefore written as x * one + y with `one` an
// opaque kernel argument (== 1), which ptxas must keep as IMAD on the FMA pipe: 8 ALU + 6 FMA ops per G instead of 10 + 2.
#define B3_ADD(x, y) ((x) * one + (y))
#define B3_G(a, b, c, d, x, y)                                          \
    a = B3_ADD(b, a); a = B3_ADD((x), a); d = __byte_perm(d ^ a, 0, 0x1032); \
    c = B3_ADD(d, c);                     b = rotr32(b ^ c, 12);            \
    a = B3_ADD(b, a); a = B3_ADD((y), a); d = __byte_perm(d ^ a, 0, 0x0321); \
    c = B3_ADD(d, c);                     b = rotr32(b ^ c, 7);

#define B3_ROUND(m0, m1, m2, m3, m4, m5, m6, m7, m8, m9, m10, m11, m12, m13, m14, m15) \
    B3_G(v0, v4, v8, v12, m0, m1)   B3_G(v1, v5, v9, v13, m2, m3)                        \
    B3_G(v2, v6, v10, v14, m4, m5)  B3_G(v3, v7, v11, v15, m6, m7)                       \
    B3_G(v0, v5, v10, v15, m8, m9)  B3_G(v1, v6, v11, v12, m10, m11)                     \
    B3_G(v2, v7, v8, v13, m12, m13) B3_G(v3, v4, v9, v14, m14, m15)

// cv <- first 8 words of compress(cv, m, counter, block_len, flags)   (ext/blake3_portable.c:46-122)
__device__ __forceinline__ void b3_compress(uint32_t (&cv)[8], const uint32_t (&m)[16], uint32_t counter_lo,
                                            uint32_t block_len, uint32_t flags, uint32_t one)
{
    uint32_t v0 = cv[0], v1 = cv[1], v2 = cv[2], v3 = cv[3], v4 = cv[4], v5 = cv[5], v6 = cv[6], v7 = cv[7];
    uint32_t v8 = IV0, v9 = IV1, v10 = IV2, v11 = IV3, v12 = counter_lo, v13 = 0, v14 = block_len, v15 = flags;
    B3_ROUND(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13], m[14], m[15]);
    B3_ROUND(m[2], m[6], m[3], m[10], m[7], m[0], m[4], m[13], m[1], m[11], m[12], m[5], m[9], m[14], m[15], m[8]);
    B3_ROUND(m[3], m[4], m[10], m[12], m[13], m[2], m[7], m[14], m[6], m[5], m[9], m[0], m[11], m[15], m[8], m[1]);
    B3_ROUND(m[10], m[7], m[12], m[9], m[14], m[3], m[13], m[15], m[4], m[0], m[11], m[2], m[5], m[8], m[1], m[6]);
    B3_ROUND(m[12], m[13], m[9], m[11], m[15], m[10], m[14], m[8], m[7], m[2], m[5], m[3], m[0], m[1], m[6], m[4]);
    B3_ROUND(m[9], m[14], m[11], m[5], m[8], m[12], m[15], m[1], m[13], m[3], m[0], m[10], m[2], m[6], m[4], m[7]);
    B3_ROUND(m[11], m[15], m[5], m[0], m[1], m[9], m[8], m[6], m[14], m[10], m[2], m[12], m[3], m[4], m[7], m[13]);
    cv[0] = v0 ^ v8;  cv[1] = v1 ^ v9;  cv[2] = v2 ^ v10; cv[3] = v3 ^ v11;
    cv[4] = v4 ^ v12; cv[5] = v5 ^ v13; cv[6] = v6 ^ v14; cv[7] = v7 ^ v15;
}

constexpr int LEAF_THREADS = 256;
constexpr int LEAF_WARPS = LEAF_THREADS / 32;
constexpr int LEAF_LANE_BYTES = 80;                       // 64 bytes at any alignment fit in five 16-byte pieces
constexpr int LEAF_STAGE_BYTES = 32 * LEAF_LANE_BYTES;    // per warp, per stage

} // namespace

__global__ void k_leaf_counts(const uint32_t* __restrict__ len, uint32_t count, uint32_t* __restrict__ leaf_count)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count)
    {
        uint32_t l = len[i];
        leaf_count[i] = l ? (l + 1023u) >> 10 : 1u; // an empty input is one empty leaf (ext/blake3.c:571-616)
    }
}

__global__ void __launch_bounds__(LEAF_THREADS, 3)
k_blake3_leaves(const uint8_t* __restrict__ base, uint64_t base_size, const uint64_t* __restrict__ seg_off,
                const uint32_t* __restrict__ seg_len, const uint32_t* __restrict__ leaf_prefix, uint32_t seg_count,
                uint32_t total_leaves, uint32_t* __restrict__ cvs, uint64_t* __restrict__ hash_out, uint32_t one,
                uint4* __restrict__ merge_items, uint32_t* __restrict__ merge_counts)
{
    __shared__ __align__(16) uint8_t s_stage[LEAF_WARPS][2][LEAF_STAGE_BYTES];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t leaf = (blockIdx.x * LEAF_WARPS + warp) * 32u + lane;
    const bool live = leaf < total_leaves;

    // which segment owns this leaf: largest s with leaf_prefix[s] <= leaf
    uint32_t seg = 0;
    if (live)
    {
        uint32_t lo = 0, hi = seg_count; // invariant: leaf_prefix[lo] <= leaf < leaf_prefix[hi]
        while (hi - lo > 1)
        {
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&leaf_prefix[mid]) <= leaf) lo = mid; else hi = mid;
        }
        seg = lo;
    }
    uint64_t addr = 0;      // byte offset of my leaf inside base
    uint32_t len = 0;       // bytes in my leaf (0..1024)
    uint32_t leaf_in_seg = 0;
    uint32_t seg_leaves = 0;
    bool single = false;
    if (live)
    {
        const uint32_t first = __ldg(&leaf_prefix[seg]);
        seg_leaves = __ldg(&leaf_prefix[seg + 1]) - first;
        const uint32_t slen = __ldg(&seg_len[seg]);
        leaf_in_seg = leaf - first;
        addr = __ldg(&seg_off[seg]) + (uint64_t)leaf_in_seg * 1024u;
        len = min(1024u, slen - leaf_in_seg * 1024u);
        single = slen <= 1024u;
    }
    const uint32_t nblocks = live ? max(1u, (len + 63u) >> 6) : 0u;
    const uint32_t max_blocks = __reduce_max_sync(0xffffffffu, nblocks);

    const uint32_t stage0 = smem_u32(&s_stage[warp][0][0]);
    const uint32_t sh = (uint32_t)addr & 15u; // my bytes start this far into the first 16-byte piece
    const uint64_t first_piece = addr & ~(uint64_t)15;
    const uint64_t my_end = addr + len;

    // The warp copies, for each lane q, the five 16-byte pieces covering q's next 64 bytes; piece ids are spread over lanes
    // so that consecutive lanes fetch consecutive 16 bytes (coalesced 80-byte runs).  Which piece a lane fetches does not
    // depend on the block index, so source pointer / bytes-left / destination are set up once and advanced by 64 per block.
    const uint8_t* p_src[5];
    int32_t p_left[5];
    uint32_t p_dst[5];
#pragma unroll
    for (uint32_t r = 0; r < 5; ++r)
    {
        const uint32_t id = r * 32u + lane;
        const uint32_t q = id / 5u, k = id - q * 5u;
        const uint64_t q_first = __shfl_sync(0xffffffffu, first_piece, q);
        const uint64_t q_end = __shfl_sync(0xffffffffu, my_end, q);
        const uint64_t src = q_first + (uint64_t)k * 16u;
        p_src[r] = base + src;
        p_left[r] = (int32_t)(int64_t)(q_end - src); // <= 1024 + 15; <= 0 when the piece holds nothing of the leaf
        p_dst[r] = stage0 + q * LEAF_LANE_BYTES + k * 16u;
    }
    auto issue = [&](uint32_t st) {
#pragma unroll
        for (uint32_t r = 0; r < 5; ++r)
        {
            if (p_left[r] > 0) cp_async16(p_dst[r] + st * LEAF_STAGE_BYTES, p_src[r], (uint32_t)min(p_left[r], 16));
            p_src[r] += 64;
            p_left[r] -= 64;
        }
        cp_async_commit();
    };

    uint32_t cv[8] = {IV0, IV1, IV2, IV3, IV4, IV5, IV6, IV7};
    if (max_blocks > 0) issue(0);
    for (uint32_t j = 0; j < max_blocks; ++j)
    {
        const uint32_t st = j & 1u;
        if (j + 1 < max_blocks)
        {
            issue(st ^ 1u);
            cp_async_wait<1>();
        }
        else
            cp_async_wait<0>();
        __syncwarp();
        if (j < nblocks)
        {
            const uint32_t rd = stage0 + st * LEAF_STAGE_BYTES + lane * LEAF_LANE_BYTES + (sh & ~3u);
            const uint32_t bs = (sh & 3u) * 8u;
            uint32_t u[17];
#pragma unroll
            for (int i = 0; i < 17; ++i) u[i] = lds32(rd + 4 * i);
            uint32_t m[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) m[i] = __funnelshift_r(u[i], u[i + 1], bs);
            const uint32_t nb = min(64u, len - j * 64u); // bytes of input in this block (0 only for the empty leaf)
            if (nb < 64u)
            {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                {
                    int rem = (int)nb - 4 * i;
                    m[i] = rem >= 4 ? m[i] : (rem <= 0 ? 0u : (m[i] & ((1u << (8 * rem)) - 1u)));
                }
            }
            uint32_t flags = (j == 0 ? F_CHUNK_START : 0u);
            if (j == nblocks - 1) flags |= F_CHUNK_END | (single ? F_ROOT : 0u);
            b3_compress(cv, m, leaf_in_seg, nb, flags, one);
        }
        __syncwarp(); // everyone is done reading stage st before it is refilled two iterations later
    }
    if (live)
    {
        if (single)
            hash_out[seg] = (uint64_t)cv[0] | ((uint64_t)cv[1] << 32);
        else
        {
            uint4* dst = reinterpret_cast<uint4*>(cvs + (size_t)leaf * 8u);
            dst[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
            dst[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
        }
    }
    // level-0 work list of the tree merge: every even leaf that has a right sibling
    const bool pair = live && !single && (leaf_in_seg & 1u) == 0 && leaf_in_seg + 1 < seg_leaves;
    const uint32_t pairs = __ballot_sync(0xffffffffu, pair);
    if (pairs)
    {
        uint32_t at = 0;
        if (lane == 0) at = atomicAdd(&merge_counts[0], (uint32_t)__popc(pairs));
        at = __shfl_sync(0xffffffffu, at, 0);
        if (pair) merge_items[at + __popc(pairs & ((1u << lane) - 1u))] = make_uint4(leaf, leaf_in_seg, seg_leaves, seg);
    }
}

// Tree merge, one launch per level, one LANE per parent node (ext/blake3.c:161-166 restated bottom-up: nodes stay in the
// slot of their first leaf; at level l the node at leaf k = 0 mod 2^(l+1) absorbs its sibling at k + 2^l when that exists,
// an unpaired node simply stays — which is exactly "the left subtree takes the largest power of two").  The work list of a
// level holds {slot, k, leaves, segment} of every left node that has a sibling; a lane that still has a sibling one level up
// re-queues itself, so the lists stay dense and every warp is full at every level.
constexpr int MERGE_THREADS = 128;

__global__ void __launch_bounds__(MERGE_THREADS)
k_blake3_merge_level(const uint4* __restrict__ items_in, uint4* __restrict__ items_out, uint32_t* __restrict__ counts, uint32_t level,
                     uint32_t* __restrict__ cvs, uint64_t* __restrict__ hash_out, uint32_t one)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t t = blockIdx.x * MERGE_THREADS + threadIdx.x;
    const uint32_t count = counts[level];
    bool again = false;
    uint4 it = make_uint4(0, 0, 0, 0);
    if (t < count)
    {
        it = items_in[t];
        const uint32_t span = 1u << level;
        uint4* left = reinterpret_cast<uint4*>(cvs + (size_t)it.x * 8u);
        const uint4* right = reinterpret_cast<const uint4*>(cvs + (size_t)(it.x + span) * 8u);
        const uint4 a0 = left[0], a1 = left[1], b0 = right[0], b1 = right[1];
        uint32_t m[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t cv[8] = {IV0, IV1, IV2, IV3, IV4, IV5, IV6, IV7};
        const bool root = it.y == 0 && 2u * span >= it.z;
        b3_compress(cv, m, 0, 64, F_PARENT | (root ? F_ROOT : 0u), one); // parent node: key = IV, counter 0 (ext/blake3.c:89-116)
        if (root)
            hash_out[it.w] = (uint64_t)cv[0] | ((uint64_t)cv[1] << 32);
        else
        {
            left[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
            left[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
            again = (it.y & (4u * span - 1u)) == 0 && it.y + 2u * span < it.z;
        }
    }
    const uint32_t m2 = __ballot_sync(0xffffffffu, again);
    if (m2)
    {
        uint32_t at = 0;
        if (lane == 0) at = atomicAdd(&counts[level + 1], (uint32_t)__popc(m2));
        at = __shfl_sync(0xffffffffu, at, 0);
        if (again) items_out[at + __popc(m2 & ((1u << lane) - 1u))] = it;
    }
}

void launch_leaf_counts(const uint32_t* d_len, uint32_t count, uint32_t* d_leaf_count, cudaStream_t st)
{
    if (!count) return;
    k_leaf_counts<<<(count + 255) / 256, 256, 0, st>>>(d_len, count, d_leaf_count);
}

void launch_blake3_leaves(const uint8_t* d_base, uint64_t base_size, const uint64_t* d_off, const uint32_t* d_len,
                          const uint32_t* d_leaf_prefix, uint32_t count, uint32_t total_leaves, uint32_t* d_cvs,
                          uint64_t* d_hash_out, uint4* d_merge_items, uint32_t* d_merge_counts, cudaStream_t st)
{
    if (!count || !total_leaves) return;
    cudaMemsetAsync(d_merge_counts, 0, sizeof(uint32_t) * BLAKE3_MAX_LEVELS, st);
    const uint32_t leaves_per_block = LEAF_WARPS * 32;
    k_blake3_leaves<<<(total_leaves + leaves_per_block - 1) / leaves_per_block, LEAF_THREADS, 0, st>>>(
        d_base, base_size, d_off, d_len, d_leaf_prefix, count, total_leaves, d_cvs, d_hash_out, 1u, d_merge_items, d_merge_counts);
}

// levels = ceil(log2(largest leaf count of a segment)); items_a and items_b each hold total_leaves/2 + 2 entries
uint32_t launch_blake3_merge(uint32_t total_leaves, uint32_t segment_count, uint32_t max_segment_leaves, uint32_t* d_cvs, uint64_t* d_hash_out,
                             uint4* d_items_a, uint4* d_items_b, uint32_t* d_merge_counts, cudaStream_t st)
{
    uint32_t launches = 0;
    for (uint32_t level = 0; (1u << level) < max_segment_leaves && level + 1 < BLAKE3_MAX_LEVELS; ++level)
    {
        // a segment of n leaves has ceil((n - 2^l) / 2^(l+1)) < n / 2^(l+1) + 1/2 parents at level l, and never more than at level 0
        uint64_t bound64 = ((uint64_t)total_leaves >> (level + 1)) + segment_count / 2 + 2;
        if (bound64 > (uint64_t)total_leaves / 2 + 1) bound64 = (uint64_t)total_leaves / 2 + 1;
        const uint32_t bound = (uint32_t)bound64;
        const uint4* in = (level & 1u) ? d_items_b : d_items_a;
        uint4* out = (level & 1u) ? d_items_a : d_items_b;
        k_blake3_merge_level<<<(bound + MERGE_THREADS - 1) / MERGE_THREADS, MERGE_THREADS, 0, st>>>(in, out, d_merge_counts, level, d_cvs, d_hash_out, 1u);
        ++launches;
    }
    return launches;
}

} // namespace ltb
