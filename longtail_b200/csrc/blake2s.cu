// blake2s.cu — batched BLAKE2s with an 8-byte digest over variable-length byte segments on sm_100a.
//
// Re-design of Blake2Hash_HashBuffer (lib/blake2/longtail_blake2.c:95-112 -> blake2s(out, 8, data, len), ext/blake2s.c):
// BLAKE2s chains 64-byte blocks serially over the whole input, so the only parallelism is across segments.  Every LANE
// owns one segment at a time and pulls the next one from a global counter when it finishes (segments differ 16x in
// length, a static assignment would idle most lanes); the 32 lanes' next 64 bytes are staged with coalesced 16-byte
// cp.async exactly like blake3.cu.  Digest bytes 0..7 are returned as a little-endian u64; the digest length 8 is part of
// the parameter block (it is NOT a truncated BLAKE2s-256).
#include "lt_device.cuh"
#include "lt_kernels.h"

namespace ltb {

namespace {

constexpr uint32_t IV0 = 0x6A09E667u, IV1 = 0xBB67AE85u, IV2 = 0x3C6EF372u, IV3 = 0xA54FF53Au;
constexpr uint32_t IV4 = 0x510E527Fu, IV5 = 0x9B05688Cu, IV6 = 0x1F83D9ABu, IV7 = 0x5BE0CD19u; // ext/blake2s.c:42
constexpr uint32_t PARAM0 = 0x01010000u ^ 8u; // digest_length 8, key_length 0, fanout 1, depth 1 (ext/blake2s.c:85-103)

#define B2_ADD(x, y) ((x) * one + (y))
#define B2_G(a, b, c, d, x, y)                                           \
    a = B2_ADD(b, a); a = B2_ADD((x), a); d = __byte_perm(d ^ a, 0, 0x1032); \
    c = B2_ADD(d, c);                     b = rotr32(b ^ c, 12);            \
    a = B2_ADD(b, a); a = B2_ADD((y), a); d = __byte_perm(d ^ a, 0, 0x0321); \
    c = B2_ADD(d, c);                     b = rotr32(b ^ c, 7);
#define B2_ROUND(m0, m1, m2, m3, m4, m5, m6, m7, m8, m9, m10, m11, m12, m13, m14, m15) \
    B2_G(v0, v4, v8, v12, m0, m1)   B2_G(v1, v5, v9, v13, m2, m3)                        \
    B2_G(v2, v6, v10, v14, m4, m5)  B2_G(v3, v7, v11, v15, m6, m7)                       \
    B2_G(v0, v5, v10, v15, m8, m9)  B2_G(v1, v6, v11, v12, m10, m11)                     \
    B2_G(v2, v7, v8, v13, m12, m13) B2_G(v3, v4, v9, v14, m14, m15)

// h <- F(h, m, t, final)   (ext/blake2s.c:147-199)
__device__ __forceinline__ void b2s_compress(uint32_t (&h)[8], const uint32_t (&m)[16], uint32_t t0, bool final_block, uint32_t one)
{
    uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint32_t v8 = IV0, v9 = IV1, v10 = IV2, v11 = IV3, v12 = IV4 ^ t0, v13 = IV5, v14 = final_block ? ~IV6 : IV6, v15 = IV7;
    B2_ROUND(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13], m[14], m[15]);
    B2_ROUND(m[14], m[10], m[4], m[8], m[9], m[15], m[13], m[6], m[1], m[12], m[0], m[2], m[11], m[7], m[5], m[3]);
    B2_ROUND(m[11], m[8], m[12], m[0], m[5], m[2], m[15], m[13], m[10], m[14], m[3], m[6], m[7], m[1], m[9], m[4]);
    B2_ROUND(m[7], m[9], m[3], m[1], m[13], m[12], m[11], m[14], m[2], m[6], m[5], m[10], m[4], m[0], m[15], m[8]);
    B2_ROUND(m[9], m[0], m[5], m[7], m[2], m[4], m[10], m[15], m[14], m[1], m[11], m[12], m[6], m[8], m[3], m[13]);
    B2_ROUND(m[2], m[12], m[6], m[10], m[0], m[11], m[8], m[3], m[4], m[13], m[7], m[5], m[15], m[14], m[1], m[9]);
    B2_ROUND(m[12], m[5], m[1], m[15], m[14], m[13], m[4], m[10], m[0], m[7], m[6], m[3], m[9], m[2], m[8], m[11]);
    B2_ROUND(m[13], m[11], m[7], m[14], m[12], m[1], m[3], m[9], m[5], m[0], m[15], m[4], m[8], m[6], m[2], m[10]);
    B2_ROUND(m[6], m[15], m[14], m[9], m[11], m[3], m[0], m[8], m[12], m[2], m[13], m[7], m[1], m[4], m[10], m[5]);
    B2_ROUND(m[10], m[2], m[8], m[4], m[7], m[6], m[1], m[5], m[15], m[11], m[9], m[14], m[3], m[12], m[13], m[0]);
    h[0] ^= v0 ^ v8;  h[1] ^= v1 ^ v9;  h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

constexpr int B2_THREADS = 256;
constexpr int B2_WARPS = B2_THREADS / 32;
constexpr int B2_LANE_BYTES = 80;
constexpr int B2_STAGE_BYTES = 32 * B2_LANE_BYTES;

} // namespace

__global__ void __launch_bounds__(B2_THREADS, 3)
k_blake2s_segments(const uint8_t* __restrict__ base, const uint64_t* __restrict__ seg_off, const uint32_t* __restrict__ seg_len,
                   uint32_t seg_count, uint32_t* __restrict__ next_segment, uint64_t* __restrict__ hash_out, uint32_t one)
{
    __shared__ __align__(16) uint8_t s_stage[B2_WARPS][B2_STAGE_BYTES];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t stage = smem_u32(&s_stage[warp][0]);

    uint32_t seg = 0xffffffffu; // my current segment, none yet
    uint64_t pos = 0;           // byte offset in base of my next block
    uint32_t left = 0;          // bytes of my segment not yet consumed
    uint32_t total = 0;         // its length
    uint32_t h[8];
    bool exhausted = false;     // the queue is empty for this lane

    for (;;)
    {
        // ---- lanes without work take the next segments from the global queue (one atomic per warp)
        const bool need = !exhausted && seg == 0xffffffffu;
        const uint32_t needers = __ballot_sync(0xffffffffu, need);
        if (needers)
        {
            uint32_t first = 0;
            if (lane == 0) first = atomicAdd(next_segment, (uint32_t)__popc(needers));
            first = __shfl_sync(0xffffffffu, first, 0);
            if (need)
            {
                const uint32_t mine = first + __popc(needers & ((1u << lane) - 1u));
                if (mine < seg_count)
                {
                    seg = mine;
                    pos = __ldg(&seg_off[mine]);
                    total = left = __ldg(&seg_len[mine]);
                    h[0] = IV0 ^ PARAM0; h[1] = IV1; h[2] = IV2; h[3] = IV3; h[4] = IV4; h[5] = IV5; h[6] = IV6; h[7] = IV7;
                }
                else
                    exhausted = true;
            }
        }
        const bool active = seg != 0xffffffffu;
        if (!__ballot_sync(0xffffffffu, active)) break;

        // ---- stage every lane's next 64 bytes: 160 sixteen-byte pieces, consecutive lanes fetch consecutive pieces
        const uint64_t first_piece = pos & ~(uint64_t)15;
        const uint64_t my_end = pos + (active ? min(left, 64u) : 0u);
#pragma unroll
        for (uint32_t r = 0; r < 5; ++r)
        {
            const uint32_t id = r * 32u + lane;
            const uint32_t q = id / 5u, k = id - q * 5u;
            const uint64_t q_first = __shfl_sync(0xffffffffu, first_piece, q);
            const uint64_t q_end = __shfl_sync(0xffffffffu, my_end, q);
            const uint64_t src = q_first + (uint64_t)k * 16u;
            if (src < q_end) cp_async16(stage + q * B2_LANE_BYTES + k * 16u, base + src, (uint32_t)min((uint64_t)16, q_end - src));
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();

        if (active)
        {
            const uint32_t sh = (uint32_t)pos & 15u;
            const uint32_t rd = stage + lane * B2_LANE_BYTES + (sh & ~3u);
            const uint32_t bs = (sh & 3u) * 8u;
            uint32_t u[17];
#pragma unroll
            for (int i = 0; i < 17; ++i) u[i] = lds32(rd + 4 * i);
            uint32_t m[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) m[i] = __funnelshift_r(u[i], u[i + 1], bs);
            const uint32_t nb = min(left, 64u);
            if (nb < 64u)
            {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                {
                    int rem = (int)nb - 4 * i;
                    m[i] = rem >= 4 ? m[i] : (rem <= 0 ? 0u : (m[i] & ((1u << (8 * rem)) - 1u)));
                }
            }
            const bool last = left <= 64u; // the final block is the last (possibly full, possibly empty) one, ext/blake2s.c:223-245
            left -= nb;
            pos += nb;
            b2s_compress(h, m, total - left, last, one); // t counts the bytes consumed so far, segments are < 4 GiB so t1 = 0
            if (last)
            {
                hash_out[seg] = (uint64_t)h[0] | ((uint64_t)h[1] << 32);
                seg = 0xffffffffu;
            }
        }
        __syncwarp(); // the stage is reused by the next iteration
    }
}

void launch_blake2s_segments(const uint8_t* d_base, const uint64_t* d_off, const uint32_t* d_len, uint32_t count, uint32_t* d_counter,
                             uint64_t* d_hash_out, int sm_count, cudaStream_t st)
{
    if (!count) return;
    cudaMemsetAsync(d_counter, 0, sizeof(uint32_t), st);
    uint32_t blocks = (count + B2_THREADS - 1) / B2_THREADS;
    const uint32_t max_blocks = (uint32_t)sm_count * 3u;
    if (blocks > max_blocks) blocks = max_blocks;
    k_blake2s_segments<<<blocks, B2_THREADS, 0, st>>>(d_base, d_off, d_len, count, d_counter, d_hash_out, 1u);
}

} // namespace ltb
