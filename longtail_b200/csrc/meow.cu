// meow.cu — batched Meow hash 0.5/calico (low 64 bits, default seed) over variable-length byte segments on sm_100a.
//
// Re-design of MeowHash_HashBuffer (lib/meowhash/longtail_meowhash.c:43-50 -> MeowBegin / MeowAbsorb / MeowEnd,
// lib/meowhash/ext/meow_hash_x64_aesni.h:459-700).  Meow chains 256-byte blocks serially over the whole input (8 MEOW_MIX per
// block, each 2 AESDEC + 2 PADDQ + 2 PXOR over loads at +15/+0/+1/+16, :181-192), so the only parallelism is across
// segments: every LANE owns one segment at a time and pulls the next one from a global counter when it finishes (like
// blake2s.cu).  The x86 AESDEC round is one inverse T-table (InvSubBytes + InvMixColumns of a row-0 byte) kept in shared
// memory replicated per bank (entry v of lane l at v*128 + l*4), so the 16 data-dependent lookups of a round never
// conflict; the other three tables are byte rotations of it (PRMT).  The lanes' next 256 bytes are staged with coalesced
// 16-byte cp.async into 272-byte rows and read back at the lane's own (arbitrary) alignment with funnel shifts.
//
// The finalisation (MeowEnd: residual mix, length mix, up to 7 lane mixes, 12 MEOW_SHUFFLE, fold; :583-700) costs about
// as much as 2.5 blocks, and a warp pays for it whenever ANY lane runs it; finished lanes therefore wait until a quarter
// of the warp is ready (or nobody has block work left) so the two code paths are not both executed every iteration.
#include "lt_device.cuh"
#include "lt_kernels.h"

namespace ltb {

namespace {

constexpr int MW_THREADS = 256;
constexpr int MW_WARPS = MW_THREADS / 32;
constexpr int MW_ROW = 272;                   // 256 bytes + 16 of alignment slack per lane
constexpr int MW_STAGE_BYTES = 32 * MW_ROW;
constexpr int MW_TABLE_BYTES = 256 * 128;     // 256 entries x 32 banks
constexpr int MW_SMEM_BYTES = MW_TABLE_BYTES + MW_WARPS * MW_STAGE_BYTES;

// MeowDefaultSeed (:234-252) as little-endian words, 8 registers of 4 words
__constant__ uint32_t c_meow_seed[32] = {
    0xA8F64332u, 0x8D305A88u, 0xA2983131u, 0x340737E0u, 0x8293404Au, 0x1DF39922u, 0xA9EF8200u, 0xC8E6C48Eu,
    0x1E825294u, 0x37018D63u, 0x6C46E57Bu, 0xC6904EF3u, 0x9BC20ACCu, 0x0DC5977Cu, 0x5B4DF8D3u, 0x9170545Bu,
    0x5D6D2179u, 0xB19F9798u, 0xBA1013BDu, 0x5AFB8D69u, 0x2DD7FFC2u, 0xFBAD01BDu, 0xFE1A8E7Bu, 0xE967A2D6u,
    0x04C9A76Bu, 0xF9C7125Fu, 0x94194A92u, 0xCF16397Bu, 0x2E1F8070u, 0xC1EF5828u, 0x0D923666u, 0xE6741587u};

struct X128
{
    uint32_t w[4];
};

// _mm_aesdec_si128(a, k) = InvMixColumns(InvSubBytes(InvShiftRows(a))) ^ k; byte 4c+r of the register is state[r][c].
// t = shared address of this lane's copy of the table
__device__ __forceinline__ X128 aesdec(const X128& a, const X128& k, uint32_t t)
{
    X128 o;
#pragma unroll
    for (int c = 0; c < 4; ++c)
    {
        const uint32_t b0 = a.w[c] & 0xffu;
        const uint32_t b1 = (a.w[(c + 3) & 3] >> 8) & 0xffu;
        const uint32_t b2 = (a.w[(c + 2) & 3] >> 16) & 0xffu;
        const uint32_t b3 = a.w[(c + 1) & 3] >> 24;
        const uint32_t t0 = lds32(t + b0 * 128u);
        const uint32_t t1 = __byte_perm(lds32(t + b1 * 128u), 0, 0x2103); // rotl 8
        const uint32_t t2 = __byte_perm(lds32(t + b2 * 128u), 0, 0x1032); // rotl 16
        const uint32_t t3 = __byte_perm(lds32(t + b3 * 128u), 0, 0x0321); // rotl 24
        o.w[c] = t0 ^ t1 ^ t2 ^ t3 ^ k.w[c];
    }
    return o;
}
__device__ __forceinline__ void paddq(X128& a, const X128& b)
{
    uint64_t a0 = a.w[0] | ((uint64_t)a.w[1] << 32), a1 = a.w[2] | ((uint64_t)a.w[3] << 32);
    a0 += b.w[0] | ((uint64_t)b.w[1] << 32);
    a1 += b.w[2] | ((uint64_t)b.w[3] << 32);
    a.w[0] = (uint32_t)a0; a.w[1] = (uint32_t)(a0 >> 32); a.w[2] = (uint32_t)a1; a.w[3] = (uint32_t)(a1 >> 32);
}
__device__ __forceinline__ void pxor(X128& a, const X128& b)
{
    a.w[0] ^= b.w[0]; a.w[1] ^= b.w[1]; a.w[2] ^= b.w[2]; a.w[3] ^= b.w[3];
}

// MEOW_MIX_REG (:181-189)
template <int R1, int R2, int R3, int R4, int R5>
__device__ __forceinline__ void mix_reg(X128 (&x)[8], const X128& i1, const X128& i2, const X128& i3, const X128& i4, uint32_t t)
{
    x[R1] = aesdec(x[R1], x[R2], t);
    paddq(x[R3], i1);
    pxor(x[R2], i2);
    x[R2] = aesdec(x[R2], x[R4], t);
    paddq(x[R5], i3);
    pxor(x[R4], i4);
}
// MEOW_SHUFFLE (:194-200)
template <int R1, int R2, int R3, int R4, int R5, int R6>
__device__ __forceinline__ void shuffle(X128 (&x)[8], uint32_t t)
{
    x[R1] = aesdec(x[R1], x[R4], t);
    paddq(x[R2], x[R5]);
    pxor(x[R4], x[R6]);
    x[R4] = aesdec(x[R4], x[R2], t);
    paddq(x[R5], x[R6]);
    pxor(x[R2], x[R3]);
}

// the four 16-byte loads of MEOW_MIX (:191-192) at byte offsets +15, +0, +1, +16 of the 32-byte lane that starts `sh`
// bytes (0..15) into the 16-byte aligned shared row position `row`
struct MixInputs
{
    X128 i1, i2, i3, i4;
};
__device__ __forceinline__ MixInputs load_mix_inputs(uint32_t row, uint32_t sh)
{
    const uint32_t a = row + (sh & ~3u);
    const uint32_t b = (sh & 3u) * 8u;
    uint32_t w[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) w[i] = lds32(a + 4 * i);
    uint32_t u[5]; // the same byte stream 3 bytes further on: u[j] = bytes 4(j+3)+3 ...
#pragma unroll
    for (int j = 0; j < 5; ++j) u[j] = (j + 4 < 9) ? __funnelshift_r(w[j + 3], w[j + 4], 24) : (w[j + 3] >> 24);
    MixInputs m;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        m.i2.w[i] = __funnelshift_r(w[i], w[i + 1], b);            // +0
        m.i4.w[i] = __funnelshift_r(w[i + 4], w[i + 5], b);        // +16
        m.i3.w[i] = __funnelshift_rc(w[i], w[i + 1], b + 8u);      // +1  (b + 8 == 32 selects w[i+1])
        m.i1.w[i] = __funnelshift_r(u[i], u[i + 1], b);            // +15
    }
    return m;
}

template <int S>
__device__ __forceinline__ void mix_lane(X128 (&x)[8], uint32_t row, uint32_t sh, uint32_t t)
{
    const MixInputs m = load_mix_inputs(row, sh);
    mix_reg<S & 7, (S + 4) & 7, (S + 6) & 7, (S + 1) & 7, (S + 2) & 7>(x, m.i1, m.i2, m.i3, m.i4, t);
}

} // namespace

__global__ void __launch_bounds__(MW_THREADS, 2)
k_meow_segments(const uint8_t* __restrict__ base, const uint64_t* __restrict__ seg_off, const uint32_t* __restrict__ seg_len,
                uint32_t seg_count, uint32_t* __restrict__ next_segment, uint64_t* __restrict__ hash_out, const uint32_t* __restrict__ td0)
{
    extern __shared__ __align__(128) uint8_t s_mem[];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    {
        uint32_t* tab = reinterpret_cast<uint32_t*>(s_mem);
        for (uint32_t i = threadIdx.x; i < 256u * 32u; i += MW_THREADS) tab[i] = __ldg(&td0[i >> 5]);
    }
    __syncthreads();
    const uint32_t t = smem_u32(s_mem) + lane * 4u;
    const uint32_t stage = smem_u32(s_mem + MW_TABLE_BYTES + warp * MW_STAGE_BYTES);
    const uint32_t my_row = stage + lane * MW_ROW;

    uint32_t seg = 0xffffffffu; // my current segment, none yet
    uint64_t pos = 0;           // byte offset in base of my next unread byte
    uint32_t left = 0;          // bytes of my segment not yet absorbed
    uint32_t total = 0;
    X128 x[8];
    bool exhausted = false;

    for (;;)
    {
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
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) x[i].w[j] = c_meow_seed[4 * i + j];
                }
                else
                    exhausted = true;
            }
        }
        const bool active = seg != 0xffffffffu;
        const uint32_t actives = __ballot_sync(0xffffffffu, active);
        if (!actives) break;
        const bool has_block = active && left >= 256u;
        const uint32_t blockers = __ballot_sync(0xffffffffu, has_block);
        const uint32_t finishers = actives & ~blockers;
        const bool run_tail = finishers && (__popc(finishers) >= 8 || !blockers);
        const bool do_tail = run_tail && active && !has_block;

        // ---- stage: a full block for the lanes that have one, the zero-padded residual for the lanes that finish now
        const uint64_t first_piece = pos & ~(uint64_t)15;
        const uint64_t my_end = pos + min(left, 256u);
        const uint32_t stagers = __ballot_sync(0xffffffffu, has_block || do_tail); // waiting and idle lanes stage nothing
#pragma unroll
        for (uint32_t r = 0; r < 17; ++r)
        {
            const uint32_t id = r * 32u + lane;
            const uint32_t q = id / 17u, k = id - q * 17u;
            const uint64_t q_first = __shfl_sync(0xffffffffu, first_piece, q);
            const uint64_t q_end = __shfl_sync(0xffffffffu, my_end, q);
            const uint64_t src = q_first + (uint64_t)k * 16u;
            if ((stagers >> q) & 1u) // bytes at and beyond the end of the segment are zero-filled (src-size 0 reads nothing)
            {
                const uint32_t nbytes = src < q_end ? (uint32_t)min((uint64_t)16, q_end - src) : 0u;
                cp_async16(stage + q * MW_ROW + k * 16u, nbytes ? base + src : base, nbytes);
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();

        const uint32_t sh = (uint32_t)pos & 15u;
        if (has_block)
        {
            // MeowAbsorbBlocks (:480-540)
            mix_lane<0>(x, my_row + 0x00, sh, t);
            mix_lane<1>(x, my_row + 0x20, sh, t);
            mix_lane<2>(x, my_row + 0x40, sh, t);
            mix_lane<3>(x, my_row + 0x60, sh, t);
            mix_lane<4>(x, my_row + 0x80, sh, t);
            mix_lane<5>(x, my_row + 0xa0, sh, t);
            mix_lane<6>(x, my_row + 0xc0, sh, t);
            mix_lane<7>(x, my_row + 0xe0, sh, t);
            pos += 256u;
            left -= 256u;
        }
        else if (do_tail)
        {
            // MeowEnd (:583-700); `left` < 256 bytes sit zero-padded in my row
            const uint32_t tail_off = left & 0xe0u; // the < 32-byte residual follows the full 32-byte lanes
            const MixInputs m = load_mix_inputs(my_row + tail_off, sh);
            // m.i2 = first 16 bytes of the residual, m.i4 = the next 16 (zero beyond the end): with Len & 0x10 these are
            // xmm9 / xmm11 of the reference, without it xmm9 holds the ragged bytes and xmm11 is zero — the same two values
            const X128 &xmm9 = m.i2, &xmm11 = m.i4;
            X128 xmm8, xmm10;
            xmm8.w[0] = __funnelshift_r(xmm11.w[3], xmm9.w[0], 24); // palignr(xmm9, xmm11, 15)
            xmm8.w[1] = __funnelshift_r(xmm9.w[0], xmm9.w[1], 24);
            xmm8.w[2] = __funnelshift_r(xmm9.w[1], xmm9.w[2], 24);
            xmm8.w[3] = __funnelshift_r(xmm9.w[2], xmm9.w[3], 24);
            xmm10.w[0] = __funnelshift_r(xmm11.w[0], xmm11.w[1], 8); // palignr(xmm9, xmm11, 1)
            xmm10.w[1] = __funnelshift_r(xmm11.w[1], xmm11.w[2], 8);
            xmm10.w[2] = __funnelshift_r(xmm11.w[2], xmm11.w[3], 8);
            xmm10.w[3] = __funnelshift_r(xmm11.w[3], xmm9.w[0], 8);
            mix_reg<0, 4, 6, 1, 2>(x, xmm8, xmm9, xmm10, xmm11, t);
            // length lanes: xmm15 = (Len, 0); xmm12 = palignr(0, xmm15, 15) = 0 for Len < 2^56; xmm14 = palignr(0, xmm15, 1)
            X128 xmm12 = {{0, 0, 0, 0}}, xmm13 = {{0, 0, 0, 0}}, xmm14 = {{total >> 8, 0, 0, 0}}, xmm15 = {{total, 0, 0, 0}};
            mix_reg<1, 5, 7, 2, 3>(x, xmm12, xmm13, xmm14, xmm15, t);
            const uint32_t lanes = left >> 5;
            if (lanes > 0) mix_lane<2>(x, my_row + 0x00, sh, t);
            if (lanes > 1) mix_lane<3>(x, my_row + 0x20, sh, t);
            if (lanes > 2) mix_lane<4>(x, my_row + 0x40, sh, t);
            if (lanes > 3) mix_lane<5>(x, my_row + 0x60, sh, t);
            if (lanes > 4) mix_lane<6>(x, my_row + 0x80, sh, t);
            if (lanes > 5) mix_lane<7>(x, my_row + 0xa0, sh, t);
            if (lanes > 6) mix_lane<0>(x, my_row + 0xc0, sh, t);
            shuffle<0, 1, 2, 4, 5, 6>(x, t);
            shuffle<1, 2, 3, 5, 6, 7>(x, t);
            shuffle<2, 3, 4, 6, 7, 0>(x, t);
            shuffle<3, 4, 5, 7, 0, 1>(x, t);
            shuffle<4, 5, 6, 0, 1, 2>(x, t);
            shuffle<5, 6, 7, 1, 2, 3>(x, t);
            shuffle<6, 7, 0, 2, 3, 4>(x, t);
            shuffle<7, 0, 1, 3, 4, 5>(x, t);
            shuffle<0, 1, 2, 4, 5, 6>(x, t);
            shuffle<1, 2, 3, 5, 6, 7>(x, t);
            shuffle<2, 3, 4, 6, 7, 0>(x, t);
            shuffle<3, 4, 5, 7, 0, 1>(x, t);
            paddq(x[0], x[2]);
            paddq(x[1], x[3]);
            paddq(x[4], x[6]);
            paddq(x[5], x[7]);
            pxor(x[0], x[1]);
            pxor(x[4], x[5]);
            paddq(x[0], x[4]);
            hash_out[seg] = (uint64_t)x[0].w[0] | ((uint64_t)x[0].w[1] << 32);
            seg = 0xffffffffu;
        }
        __syncwarp(); // the stage is reused by the next iteration
    }
}

// Td0[v] = InvMixColumns column of a row-0 byte InvSubBytes(v): (0e, 09, 0d, 0b) * InvSbox[v] (FIPS-197 5.3.2, 5.3.3)
void meow_build_table(uint32_t out[256])
{
    auto mul = [](uint32_t a, uint32_t b) {
        uint32_t r = 0;
        for (; b; b >>= 1)
        {
            if (b & 1u) r ^= a;
            a = ((a << 1) ^ ((a & 0x80u) ? 0x1bu : 0u)) & 0xffu;
        }
        return r;
    };
    uint8_t inv_sbox[256];
    for (uint32_t v = 0; v < 256; ++v)
    {
        uint32_t y = 0;
        if (v)
            for (uint32_t c = 1; c < 256; ++c)
                if (mul(v, c) == 1u) { y = c; break; }
        uint32_t z = y;
        for (int k = 1; k <= 4; ++k) z ^= ((y << k) | (y >> (8 - k))) & 0xffu;
        inv_sbox[(z ^ 0x63u) & 0xffu] = (uint8_t)v;
    }
    for (uint32_t v = 0; v < 256; ++v)
    {
        const uint32_t y = inv_sbox[v];
        out[v] = mul(y, 0x0e) | (mul(y, 0x09) << 8) | (mul(y, 0x0d) << 16) | (mul(y, 0x0b) << 24);
    }
}

cudaError_t launch_meow_segments(const uint8_t* d_base, const uint64_t* d_off, const uint32_t* d_len, uint32_t count, uint32_t* d_counter,
                                 uint64_t* d_hash_out, const uint32_t* d_td0, int sm_count, cudaStream_t st)
{
    if (!count) return cudaSuccess;
    static bool configured = false;
    if (!configured)
    {
        cudaError_t e = cudaFuncSetAttribute(k_meow_segments, cudaFuncAttributeMaxDynamicSharedMemorySize, MW_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    cudaMemsetAsync(d_counter, 0, sizeof(uint32_t), st);
    uint32_t blocks = (count + MW_THREADS - 1) / MW_THREADS;
    const uint32_t max_blocks = (uint32_t)sm_count * 2u;
    if (blocks > max_blocks) blocks = max_blocks;
    k_meow_segments<<<blocks, MW_THREADS, MW_SMEM_BYTES, st>>>(d_base, d_off, d_len, count, d_counter, d_hash_out, d_td0);
    return cudaGetLastError();
}

} // namespace ltb
