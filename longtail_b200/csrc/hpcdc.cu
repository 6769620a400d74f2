// hpcdc.cu — content-defined chunk boundaries on sm_100a.
//
// Re-design of lib/hpcdcchunker/longtail_hpcdcchunker.c:225-310 (Longtail_HPCDCNextChunk) as driven by
// src/longtail.c:2231-2296 (DynamicChunking) for a whole batch of parts at once:
//
//   k_tile_desc   tile -> (part, tile-in-part) table for the batch
//   k_hpcdc_scan  every position's 48-byte Buzhash (the window hash is a pure function of the trailing 48
//                 bytes) is evaluated once; positions with hash % d == d-1 ("candidates") are
//                 written per 64 KiB tile, sorted, to a small slot list
//   k_hpcdc_walk  one CTA per part: compacts the part's candidates and walks the min / max selection rule
//                 (:257-264 left <= min, :285 lim = min(left, max), first hit in [min+1, lim]) sequentially
//                 over the sparse list, 32 candidates per step
//
// All arithmetic is u32 and bit-exact; nothing here is a heuristic.
#include "lt_device.cuh"
#include "lt_kernels.h"

namespace ltb {

__global__ void k_tile_desc(const PartDesc* __restrict__ parts, uint32_t part_count, uint2* __restrict__ tile_desc)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= part_count) return;
    PartDesc pd = parts[p];
    uint32_t n = (pd.size + SCAN_TILE - 1) / SCAN_TILE;
    for (uint32_t t = 0; t < n; ++t) tile_desc[pd.tile_start + t] = make_uint2(p, t);
}

// exact re-evaluation of one 16-byte group after the fast divisibility filter fired (rare)
__device__ __noinline__ void scan_group_exact(uint32_t in_addr, uint32_t out_addr, uint32_t h, uint32_t d,
                                              const uint32_t* __restrict__ g_table, uint32_t* bitmap_row, uint32_t first_bit)
{
    for (uint32_t k = 0; k < 16; ++k)
    {
        uint32_t in, out;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(in) : "r"(in_addr + k));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(out) : "r"(out_addr + k));
        h = rotl32(h, 1) ^ rotl32(__ldg(&g_table[out]), 16) ^ __ldg(&g_table[in]); // longtail_hpcdcchunker.c:295-297
        if (h % d == d - 1)                                                       // :298
        {
            uint32_t bit = first_bit + k;
            bitmap_row[bit >> 5] |= 1u << (bit & 31);
        }
    }
}

__global__ void __launch_bounds__(SCAN_THREADS, 1)
k_hpcdc_scan(const uint8_t* __restrict__ arena, const PartDesc* __restrict__ parts, const uint2* __restrict__ tile_desc,
             uint32_t num_tiles, ChunkParams cp, const uint32_t* __restrict__ g_table,
             uint32_t* __restrict__ tile_count, uint32_t* __restrict__ tile_slots)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint32_t* s_table = reinterpret_cast<uint32_t*>(smem);
    uint8_t* s_rows = smem + SCAN_TABLE_BYTES;
    uint32_t* s_bitmap = reinterpret_cast<uint32_t*>(smem + SCAN_TABLE_BYTES + 2 * SCAN_ROWS_BYTES);
    uint32_t* s_warp = s_bitmap + SCAN_THREADS * (SCAN_SEG / 32);

    const uint32_t tid = threadIdx.x;
    const uint32_t lane4 = (tid & 31u) * 4u;

    // bank-replicated substitution table: entry v occupies 256 B = 32 lanes x T[v] then 32 lanes x rotl(T[v],16),
    // so lane l always reads bank l and a lookup is conflict free for any byte values
    for (uint32_t i = tid; i < 256 * 64; i += SCAN_THREADS)
    {
        uint32_t t = __ldg(&g_table[i >> 6]);
        s_table[i] = (i & 32u) ? rotl32(t, 16) : t;
    }
    for (uint32_t i = tid; i < SCAN_THREADS * (SCAN_SEG / 32); i += SCAN_THREADS) s_bitmap[i] = 0;

    const uint32_t rows_addr = smem_u32(s_rows);

    auto issue_tile = [&](uint32_t tile, uint32_t buf) {
        const uint2 td = tile_desc[tile];
        const PartDesc pd = parts[td.x];
        const uint32_t tile_off = td.y * (uint32_t)SCAN_TILE;
        const uint8_t* src = arena + pd.data_off + tile_off;
        const uint32_t dst = rows_addr + buf * SCAN_ROWS_BYTES;
#pragma unroll 4
        for (uint32_t i = tid; i < SCAN_TILE / 16; i += SCAN_THREADS)
        {
            uint32_t off = tile_off + i * 16u;
            uint32_t nb = pd.size > off ? min(16u, pd.size - off) : 0u;
            cp_async16(dst + ((i >> 4) + 1u) * SCAN_ROW + (i & 15u) * 16u, nb ? src + i * 16u : arena, nb);
        }
        if (tid < 3)
        {
            // 48-byte halo in front of the tile; a part's first tile sees zeros (those positions can never be cuts:
            // a cut needs at least min >= 48 bytes of the same part in front of it)
            cp_async16(dst + (SCAN_SEG - SCAN_WINDOW) + tid * 16u, td.y ? src - SCAN_WINDOW + tid * 16u : arena, td.y ? 16u : 0u);
        }
        cp_async_commit();
    };

    uint32_t tile = blockIdx.x;
    uint32_t buf = 0;
    if (tile < num_tiles) issue_tile(tile, 0);

    for (; tile < num_tiles; tile += gridDim.x, buf ^= 1u)
    {
        const uint32_t next = tile + gridDim.x;
        if (next < num_tiles)
        {
            issue_tile(next, buf ^ 1u);
            cp_async_wait<1>();
        }
        else
            cp_async_wait<0>();
        __syncthreads();

        const uint32_t my = rows_addr + buf * SCAN_ROWS_BYTES + (tid + 1u) * SCAN_ROW;
        const uint32_t prev = my - SCAN_ROW + (SCAN_SEG - SCAN_WINDOW);
        const uint32_t tab = smem_u32(s_table);

        // seed: hash of the 48 bytes in front of my segment (longtail_hpcdcchunker.c:273-279)
        uint32_t h = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j)
        {
            uint4 w = lds128(prev + 16 * j);
            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 16; ++k)
            {
                uint32_t idx = __byte_perm(ws[k >> 2], lane4, 0x5504 | ((k & 3) << 4));
                h ^= rotl32(lds32(tab + idx), (47 - (16 * j + k)) & 31);
            }
        }

        uint32_t* my_bits = s_bitmap + tid * (SCAN_SEG / 32);
#pragma unroll 2
        for (int j = 0; j < SCAN_SEG / 16; ++j)
        {
            const uint32_t in_addr = my + 16 * j;
            const uint32_t out_addr = j < 3 ? prev + 16 * j : my + 16 * (j - 3);
            const uint4 cin = lds128(in_addr);
            const uint4 cout = lds128(out_addr);
            const uint32_t wi[4] = {cin.x, cin.y, cin.z, cin.w};
            const uint32_t wo[4] = {cout.x, cout.y, cout.z, cout.w};
            const uint32_t h0 = h;
            uint32_t best = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < 16; ++k)
            {
                const uint32_t sel = 0x5504 | ((k & 3) << 4);
                uint32_t tin = lds32(tab + __byte_perm(wi[k >> 2], lane4, sel));
                uint32_t tout = lds32(tab + 128u + __byte_perm(wo[k >> 2], lane4, sel));
                h = rotl32(h, 1) ^ tout ^ tin;              // :295-297 with rotl(T[out],48&31) pre-rotated in the table
                best = min(best, h * cp.d_odd_inv + cp.d_odd_inv); // (h+1)/odd(d) exact-division test, superset of :298
            }
            if (best <= cp.d_odd_thr) scan_group_exact(in_addr, out_addr, h0, cp.d, g_table, my_bits, 16 * j);
        }
        __syncthreads();

        // ordered compaction of this tile's candidate bits into its slot list
        const uint2 td = tile_desc[tile];
        const uint32_t part_size = parts[td.x].size;
        const uint32_t seg_first = td.y * (uint32_t)SCAN_TILE + tid * SCAN_SEG; // part-relative position of my first byte
        uint32_t words[SCAN_SEG / 32];
        uint32_t cnt = 0;
#pragma unroll
        for (int w = 0; w < SCAN_SEG / 32; ++w)
        {
            uint32_t v = my_bits[w];
            my_bits[w] = 0;
            // a cut after byte q is position q+1; keep it only inside the part
            uint32_t first = seg_first + 32 * w + 1;
            if (first > part_size) v = 0;
            else if (first + 31 > part_size) v &= (1u << (part_size - first + 1)) - 1u;
            words[w] = v;
            cnt += __popc(v);
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += t;
        }
        if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
        __syncthreads();
        uint32_t base = incl - cnt;
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; ++w)
        {
            uint32_t v = s_warp[w];
            if (w < (int)(tid >> 5)) base += v;
            total += v;
        }
        if (cnt)
        {
            uint32_t* out = tile_slots + (size_t)tile * cp.slots;
#pragma unroll
            for (int w = 0; w < SCAN_SEG / 32; ++w)
            {
                uint32_t v = words[w];
                while (v)
                {
                    uint32_t b = __ffs(v) - 1;
                    v &= v - 1;
                    if (base < cp.slots) out[base] = seg_first + 32 * w + b + 1;
                    ++base;
                }
            }
        }
        if (tid == 0) tile_count[tile] = total;
        __syncthreads();
    }
}

// exact window hash for positions [lo, hi] of one part, 32 positions per step (dense-tile fallback)
__device__ uint32_t walk_exact_first(const uint8_t* __restrict__ part, uint32_t lo, uint32_t hi, uint32_t d,
                                     const uint32_t* __restrict__ g_table, uint32_t lane)
{
    for (uint32_t base = lo; base <= hi; base += 32)
    {
        uint32_t p = base + lane;
        bool hit = false;
        if (p <= hi)
        {
            uint32_t h = 0;
            const uint8_t* w = part + p - SCAN_WINDOW;
            for (uint32_t i = 0; i < SCAN_WINDOW; ++i) h ^= rotl32(__ldg(&g_table[w[i]]), (SCAN_WINDOW - 1 - i) & 31);
            hit = (h % d == d - 1);
        }
        uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (m) return base + __ffs(m) - 1;
        if (base > 0xffffffffu - 32) break;
    }
    return 0xffffffffu;
}

constexpr int WALK_THREADS = 128;

__global__ void __launch_bounds__(WALK_THREADS)
k_hpcdc_walk(const uint8_t* __restrict__ arena, const PartDesc* __restrict__ parts, ChunkParams cp,
             const uint32_t* __restrict__ g_table, const uint32_t* __restrict__ tile_count,
             const uint32_t* __restrict__ tile_slots, uint32_t* __restrict__ cand,
             uint64_t* __restrict__ stage_off, uint32_t* __restrict__ stage_len, uint32_t* __restrict__ part_chunk_count)
{
    __shared__ uint32_t s_warp[WALK_THREADS / 32];
    __shared__ uint32_t s_total;
    const PartDesc pd = parts[blockIdx.x];
    const uint32_t tid = threadIdx.x;
    const uint32_t n = pd.size;
    const uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t* my_cand = cand + (size_t)pd.tile_start * cp.slots;

    // phase 1: ordered compaction of the per-tile slot lists into one dense, sorted list for the part.
    // A tile whose list overflowed contributes a single marker entry instead.
    const uint32_t per = (ntiles + WALK_THREADS - 1) / WALK_THREADS;
    const uint32_t t0 = min(ntiles, tid * per), t1 = min(ntiles, t0 + per);
    uint32_t mine = 0;
    for (uint32_t t = t0; t < t1; ++t)
    {
        uint32_t c = tile_count[pd.tile_start + t];
        mine += c <= cp.slots ? c : 1u;
    }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    uint32_t base = incl - mine;
    for (uint32_t w = 0; w < (tid >> 5); ++w) base += s_warp[w];
    if (tid == WALK_THREADS - 1) s_total = base + mine;
    for (uint32_t t = t0; t < t1; ++t)
    {
        uint32_t c = tile_count[pd.tile_start + t];
        if (c <= cp.slots)
        {
            const uint32_t* src = tile_slots + (size_t)(pd.tile_start + t) * cp.slots;
            for (uint32_t i = 0; i < c; ++i) my_cand[base + i] = src[i];
            base += c;
        }
        else
            my_cand[base++] = CAND_OVERFLOW | t;
    }
    __syncthreads();
    if (tid >= 32) return;

    // phase 2: the sequential min/max selection, one warp, 32 sorted candidates per load
    const uint32_t ncand = s_total;
    const uint32_t lane = tid;
    const uint8_t* part = arena + pd.data_off;
    uint64_t* out_off = stage_off + pd.chunk_start;
    uint32_t* out_len = stage_len + pd.chunk_start;
    const uint32_t NONE = 0x7fffffffu;
    uint32_t ci = 0;
    uint32_t e = ci + lane < ncand ? my_cand[ci + lane] : NONE;
    uint32_t s = 0, cnt = 0;
    while (s < n)
    {
        const uint32_t left = n - s;
        uint32_t cut;
        if (left <= cp.min) // longtail_hpcdcchunker.c:257-264
            cut = n;
        else
        {
            const uint32_t lim = min(left, cp.max); // :285
            uint32_t x = s + cp.min + 1;             // first position the rolling loop can stop at (:289-306)
            const uint32_t y = s + lim;
            cut = y;
            while (x <= y)
            {
                // largest cut position an entry can stand for: itself, or the end of an overflowed tile
                uint32_t key = (e & CAND_OVERFLOW) ? ((e & ~CAND_OVERFLOW) + 1u) * (uint32_t)SCAN_TILE : e;
                uint32_t m = __ballot_sync(0xffffffffu, key >= x);
                if (m == 0)
                {
                    ci += 32;
                    e = ci + lane < ncand ? my_cand[ci + lane] : NONE;
                    continue;
                }
                uint32_t ef = __shfl_sync(0xffffffffu, e, __ffs(m) - 1);
                if (!(ef & CAND_OVERFLOW))
                {
                    if (ef <= y) cut = ef; // includes NONE > y
                    break;
                }
                const uint32_t tile_first = (ef & ~CAND_OVERFLOW) * (uint32_t)SCAN_TILE + 1u;
                if (tile_first > y) break;
                const uint32_t lo = max(x, tile_first);
                const uint32_t hi = min(y, tile_first + (uint32_t)SCAN_TILE - 1u);
                uint32_t r = walk_exact_first(part, lo, hi, cp.d, g_table, lane);
                if (r != 0xffffffffu)
                {
                    cut = r;
                    break;
                }
                x = hi + 1;
            }
        }
        if (lane == 0)
        {
            out_off[cnt] = pd.data_off + s;
            out_len[cnt] = cut - s;
        }
        ++cnt;
        s = cut;
    }
    if (lane == 0) part_chunk_count[blockIdx.x] = cnt;
}

// gather the per-part staged chunk lists into dense arrays in part order
__global__ void k_compact_chunks(const PartDesc* __restrict__ parts, const uint32_t* __restrict__ part_chunk_count,
                                 const uint32_t* __restrict__ part_chunk_base, const uint64_t* __restrict__ stage_off,
                                 const uint32_t* __restrict__ stage_len, uint64_t* __restrict__ chunk_off,
                                 uint32_t* __restrict__ chunk_len, uint32_t* __restrict__ chunk_tag)
{
    const PartDesc pd = parts[blockIdx.x];
    const uint32_t n = part_chunk_count[blockIdx.x];
    const uint32_t base = part_chunk_base[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    {
        chunk_off[base + i] = stage_off[pd.chunk_start + i];
        chunk_len[base + i] = stage_len[pd.chunk_start + i];
        chunk_tag[base + i] = pd.tag;
    }
}

// ---------------------------------------------------------------- host launchers

void launch_tile_desc(const PartDesc* d_parts, uint32_t part_count, uint2* d_tile_desc, cudaStream_t st)
{
    if (!part_count) return;
    k_tile_desc<<<(part_count + 127) / 128, 128, 0, st>>>(d_parts, part_count, d_tile_desc);
}

cudaError_t launch_hpcdc_scan(const uint8_t* d_arena, const PartDesc* d_parts, const uint2* d_tile_desc, uint32_t num_tiles,
                              const ChunkParams& cp, const uint32_t* d_table, uint32_t* d_tile_count, uint32_t* d_tile_slots,
                              int sm_count, cudaStream_t st)
{
    if (!num_tiles) return cudaSuccess;
    static bool configured = false;
    if (!configured)
    {
        cudaError_t e = cudaFuncSetAttribute(k_hpcdc_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, SCAN_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    uint32_t grid = num_tiles < (uint32_t)sm_count ? num_tiles : (uint32_t)sm_count;
    k_hpcdc_scan<<<grid, SCAN_THREADS, SCAN_SMEM_BYTES, st>>>(d_arena, d_parts, d_tile_desc, num_tiles, cp, d_table, d_tile_count, d_tile_slots);
    return cudaGetLastError();
}

void launch_hpcdc_walk(const uint8_t* d_arena, const PartDesc* d_parts, uint32_t part_count, const ChunkParams& cp,
                       const uint32_t* d_table, const uint32_t* d_tile_count, const uint32_t* d_tile_slots, uint32_t* d_cand,
                       uint64_t* d_stage_off, uint32_t* d_stage_len, uint32_t* d_part_chunk_count, cudaStream_t st)
{
    if (!part_count) return;
    k_hpcdc_walk<<<part_count, WALK_THREADS, 0, st>>>(d_arena, d_parts, cp, d_table, d_tile_count, d_tile_slots, d_cand,
                                                      d_stage_off, d_stage_len, d_part_chunk_count);
}

void launch_compact_chunks(const PartDesc* d_parts, uint32_t part_count, const uint32_t* d_part_chunk_count,
                           const uint32_t* d_part_chunk_base, const uint64_t* d_stage_off, const uint32_t* d_stage_len,
                           uint64_t* d_chunk_off, uint32_t* d_chunk_len, uint32_t* d_chunk_tag, cudaStream_t st)
{
    if (!part_count) return;
    k_compact_chunks<<<part_count, 256, 0, st>>>(d_parts, d_part_chunk_count, d_part_chunk_base, d_stage_off, d_stage_len,
                                                 d_chunk_off, d_chunk_len, d_chunk_tag);
}

} // namespace ltb
