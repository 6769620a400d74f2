// hpcdc.cu — content-defined chunk boundaries on sm_100a.
//
// Re-design of lib/hpcdcchunker/longtail_hpcdcchunker.c:225-310 (Longtail_HPCDCNextChunk) as driven by
// src/longtail.c:2231-2296 (DynamicChunking) for a whole batch of parts at once:
//
//   k_tile_part   tile -> part table for the batch
//   k_hpcdc_scan  every position's 48-byte Buzhash (the window hash is a pure function of the trailing 48
//                 bytes) is evaluated once; positions with hash % d == d-1 ("candidates") are
//                 written per 8 KiB tile (one warp per tile), sorted, to a small slot list
//   k_hpcdc_walk  one CTA per part: compacts the part's candidates and walks the min / max selection rule
//                 (:257-264 left <= min, :285 lim = min(left, max), first hit in [min+1, lim]) sequentially
//                 over the sparse list, 32 candidates per step
//
// All arithmetic is u32 and bit-exact; nothing here is a heuristic.
#include "lt_device.cuh"
#include "lt_kernels.h"

namespace ltb {

// tile -> part lookup table for the batch (binary search over the parts' first-tile indices)
__global__ void k_tile_part(const PartDesc* __restrict__ parts, uint32_t part_count, uint32_t num_tiles, uint32_t* __restrict__ tile_part)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_tiles) return;
    uint32_t lo = 0, hi = part_count; // parts[lo].tile_start <= t < parts[hi].tile_start (hi == part_count: +inf)
    while (hi - lo > 1)
    {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&parts[mid].tile_start) <= t) lo = mid; else hi = mid;
    }
    // empty parts share their tile_start with the next part: the search lands on the last of them, step back is not needed
    // because `<=` always prefers the right-most part with tile_start <= t, which is the one that owns tiles
    tile_part[t] = lo;
}

// Re-evaluation of one 8-byte group after the group-level divisibility filter fired (about one group in 400 at the
// default parameters): same table lookups as the main loop, the filter again per byte and the true `%` only where the filter
// passes.  The main loop only RECORDS such groups (start hash + position, two per lane in registers); they are re-walked after
// the tile's main loop, when the 48 ring registers are dead and every lane with a recorded group works at the same time.
__device__ __forceinline__ void scan_group_exact(uint32_t in_addr, uint32_t out_addr, uint32_t h, uint32_t d, uint32_t d_odd_inv,
                                              uint32_t d_odd_thr, uint32_t tab_lane, uint32_t bitmap_row, uint32_t first_bit)
{
    // both groups are 8-byte aligned in their rows: two 8-byte loads, then the same PRMT-formed table addresses as the main loop
    uint32_t wi[2], wo[2];
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(wi[0]), "=r"(wi[1]) : "r"(in_addr));
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(wo[0]), "=r"(wo[1]) : "r"(out_addr));
    uint32_t found = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
        const uint32_t sel = 0x7604 | ((k & 3) << 4);
        h = rotl32(h, 1) ^ lds32_off128(__byte_perm(wo[k >> 2], tab_lane, sel)) ^ lds32(__byte_perm(wi[k >> 2], tab_lane, sel)); // :295-297
        if (h * d_odd_inv + d_odd_inv <= d_odd_thr && h % d == d - 1) found |= 1u << k;                                           // :298
    }
    if (found)
    {
        const uint32_t a = bitmap_row + (first_bit >> 5) * 4u; // the group lies inside one bitmap word (first_bit is a multiple of 8)
        const uint32_t v = lds32(a) | (found << (first_bit & 31u));
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
    }
}

// third and later recorded group of one lane in one tile (dense candidates: tiny discriminators, degenerate data): resolved on the spot
__device__ __noinline__ void scan_group_exact_now(uint32_t in_addr, uint32_t out_addr, uint32_t h, uint32_t d, uint32_t d_odd_inv,
                                                  uint32_t d_odd_thr, uint32_t tab_lane, uint32_t bitmap_row, uint32_t first_bit)
{
    scan_group_exact(in_addr, out_addr, h, d, d_odd_inv, d_odd_thr, tab_lane, bitmap_row, first_bit);
}

// Recorded groups live in the 16 padding bytes behind the lane's own row: three start hashes (12 B) and three group indices
// (1 B each).  A fourth recorded group in one lane-tile (probability ~1e-6 at the default parameters) marks the tile as
// overflowed, which k_hpcdc_walk resolves exactly.
constexpr uint32_t SCAN_PEND_MAX = 3;

// Sixteen positions of the rolling hash.  ring[] holds the table VALUE of each of the last 48 bytes (one register each, slots
// are compile-time constants): one PRMT extracts a byte and forms its table address, one shared-memory load per byte fetches T
// when the byte enters the window; when it leaves, 48 positions later, its rotl(T,16) is one more PRMT on the kept register
// (longtail_hpcdcchunker.c:295-297).  Measured equal to a second lookup at address + 128 (26.3 vs 25.9 ms on 64 GiB): the loop is
// bound by instruction issue (seven per byte either way), not by the shared-memory pipe.
template <int SLOT, bool SPARSE>
__device__ __forceinline__ void scan_step16(uint32_t (&ring)[SCAN_WINDOW], uint32_t& h, uint32_t& npend, uint32_t pad_addr, uint32_t in_addr,
                                            uint32_t out_addr, uint32_t first_bit, const ChunkParams& cp, uint32_t tab_lane, uint32_t bits_addr)
{
    const uint4 cin = lds128(in_addr);
    const uint32_t wi[4] = {cin.x, cin.y, cin.z, cin.w};
#pragma unroll
    for (int g = 0; g < 2; ++g)
    {
        const uint32_t h0 = h;
        uint32_t best = 0xffffffffu;
#pragma unroll
        for (int k = 8 * g; k < 8 * g + 8; ++k)
        {
            const uint32_t t = lds32(__byte_perm(wi[k >> 2], tab_lane, 0x7604 | ((k & 3) << 4)));
            h = rotl32(h, 1) ^ __byte_perm(ring[SLOT + k], 0u, 0x1032) ^ t;
            ring[SLOT + k] = t;
            best = min(best, h * cp.d_odd_inv + cp.d_odd_inv); // (h+1)/odd(d) exact-division test, superset of :298
        }
        if (best <= cp.d_odd_thr)
        {
            if (SPARSE)
            {
                if (npend < SCAN_PEND_MAX)
                {
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(pad_addr + 4u * npend), "r"(h0) : "memory");
                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(pad_addr + 12u + npend), "r"((first_bit >> 3) + g) : "memory");
                }
                ++npend;
            }
            else
                scan_group_exact_now(in_addr + 8 * g, out_addr + 8 * g, h0, cp.d, cp.d_odd_inv, cp.d_odd_thr, tab_lane, bits_addr, first_bit + 8 * g);
        }
    }
}

// One lane's 256-byte segment of a staged tile: seed, rolling hash with the group filter, re-walk of the recorded groups.
// Deliberately NOT inlined: the 48 ring registers plus the lookups in flight need the whole register file, and a call boundary
// makes the compiler park the tile loop's state once per tile (outside) instead of spilling ring slots inside the hot loop.
// Returns the number of groups this lane recorded (SPARSE) — more than SCAN_PEND_MAX means the tile must be flagged as overflowed.
template <bool SPARSE>
__device__ __noinline__ uint32_t scan_segment(uint32_t my, uint32_t prev, uint32_t tab_lane, uint32_t bits_addr, const ChunkParams cp)
{
    // seed: hash of the 48 bytes in front of my segment (longtail_hpcdcchunker.c:273-279); their table addresses fill the ring
    uint32_t h = 0;
    uint32_t ring[SCAN_WINDOW];
    uint32_t npend = 0;
    const uint32_t pad = my + SCAN_SEG;
#pragma unroll
    for (int j = 0; j < 3; ++j)
    {
        uint4 w = lds128(prev + 16 * j);
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
            const uint32_t t = lds32(__byte_perm(ws[k >> 2], tab_lane, 0x7604 | ((k & 3) << 4)));
            h ^= rotl32(t, (47 - (16 * j + k)) & 31);
            ring[16 * j + k] = t;
        }
    }

    // 256 = 5 * 48 + 16 positions; the byte leaving the window at position q sits in ring slot q % 48
    scan_step16<0, SPARSE>(ring, h, npend, pad, my, prev, 0, cp, tab_lane, bits_addr);
    scan_step16<16, SPARSE>(ring, h, npend, pad, my + 16, prev + 16, 16, cp, tab_lane, bits_addr);
    scan_step16<32, SPARSE>(ring, h, npend, pad, my + 32, prev + 32, 32, cp, tab_lane, bits_addr);
#pragma unroll 1
    for (uint32_t q = 48; q < SCAN_SEG - 16; q += 48)
    {
        scan_step16<0, SPARSE>(ring, h, npend, pad, my + q, my + q - 48, q, cp, tab_lane, bits_addr);
        scan_step16<16, SPARSE>(ring, h, npend, pad, my + q + 16, my + q - 32, q + 16, cp, tab_lane, bits_addr);
        scan_step16<32, SPARSE>(ring, h, npend, pad, my + q + 32, my + q - 16, q + 32, cp, tab_lane, bits_addr);
    }
    scan_step16<0, SPARSE>(ring, h, npend, pad, my + (SCAN_SEG - 16), my + (SCAN_SEG - 64), SCAN_SEG - 16, cp, tab_lane, bits_addr);
    // the recorded groups, every lane its own (about one lane-group per tile at the default parameters)
    if (SPARSE)
    {
        const uint32_t np = min(npend, SCAN_PEND_MAX);
#pragma unroll 1
        for (uint32_t e = 0; e < np; ++e)
        {
            uint32_t fb;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(fb) : "r"(pad + 12u + e));
            fb *= 8u;
            scan_group_exact(my + fb, fb < (uint32_t)SCAN_WINDOW ? prev + fb : my + fb - SCAN_WINDOW, lds32(pad + 4u * e), cp.d, cp.d_odd_inv,
                             cp.d_odd_thr, tab_lane, bits_addr, fb);
        }
    }
    return npend;
}

// Shared-memory layout (ScanLayout, computed on the host from the probed base of the dynamic window): the 64 KiB table is
// placed at a shared address that is a multiple of 64 KiB, so that "table base + byte * 256 + lane * 4" is produced by the
// single PRMT that extracts the byte — no address add per lookup.  `warps_before` per-warp buffers sit in front of the
// table, the rest behind it.
// SPARSE: groups that pass the filter are recorded and re-walked after the main loop (the common case, odd(d) >= 1024);
// otherwise (tiny discriminators: most groups pass) they are re-walked on the spot.
template <bool SPARSE>
__global__ void __launch_bounds__(SCAN_THREADS, 1)
k_hpcdc_scan(const uint8_t* __restrict__ arena, const PartDesc* __restrict__ parts, const uint32_t* __restrict__ tile_part,
             uint32_t num_tiles, ChunkParams cp, const uint32_t* __restrict__ g_table,
             uint32_t* __restrict__ tile_count, uint32_t* __restrict__ tile_slots, ScanLayout lay)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint32_t* s_table = reinterpret_cast<uint32_t*>(smem + lay.table_off);

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;

    // bank-replicated substitution table: entry v occupies 256 B = 32 lanes x T[v] then 32 lanes x rotl(T[v],16),
    // so lane l always reads bank l and a lookup is conflict free for any byte values
    for (uint32_t i = tid; i < 256 * 64; i += SCAN_THREADS)
    {
        uint32_t t = __ldg(&g_table[i >> 6]);
        s_table[i] = (i & 32u) ? rotl32(t, 16) : t;
    }
    uint8_t* s_mine = smem + (warp < lay.warps_before ? warp * SCAN_WARP_BYTES
                                                      : lay.table_off + SCAN_TABLE_BYTES + (warp - lay.warps_before) * SCAN_WARP_BYTES);
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(s_mine + SCAN_ROWS_BYTES) + lane * (SCAN_SEG / 32);
#pragma unroll
    for (int w = 0; w < SCAN_SEG / 32; ++w) s_bits[w] = 0;
    __syncthreads(); // the only block-wide barrier: the table is ready

    const uint32_t tab = smem_u32(s_table);   // multiple of 65536 by construction
    const uint32_t tab_lane = tab | (lane * 4u); // bytes 2..3: table base, byte 1: free for the data byte, byte 0: lane * 4
    const uint32_t rows = smem_u32(s_mine);
    const uint32_t my = rows + (lane + 1u) * SCAN_ROW;
    const uint32_t prev = my - SCAN_ROW + (SCAN_SEG - SCAN_WINDOW);
    const uint32_t bits_addr = smem_u32(s_bits);
    const uint32_t stride = gridDim.x * SCAN_WARPS;

    uint32_t tile = blockIdx.x * SCAN_WARPS + warp;
    uint32_t part_idx = tile < num_tiles ? __ldg(&tile_part[tile]) : 0u;
    for (; tile < num_tiles; tile += stride)
    {
        const PartDesc pd = parts[part_idx];
        const uint32_t tip = tile - pd.tile_start;
        const uint32_t tile_off = tip * (uint32_t)SCAN_TILE;
        const uint8_t* src = arena + pd.data_off + tile_off;
        // stage the tile: 512 sixteen-byte pieces, consecutive lanes fetch consecutive pieces (coalesced); bytes past the
        // end of the part are zero-filled
        if (tile_off + (uint32_t)SCAN_TILE <= pd.size)
        {
            // interior tile (all but the last tile of a part): two base registers, sixteen copies with immediate offsets
            const uint32_t dst0 = rows + ((lane >> 4) + 1u) * SCAN_ROW + (lane & 15u) * 16u;
            const uint8_t* src0 = src + lane * 16u;
#pragma unroll
            for (uint32_t it = 0; it < SCAN_TILE / 512; ++it) cp_async16_full(dst0 + it * 2u * SCAN_ROW, src0 + it * 512u);
        }
        else
        {
#pragma unroll 4
            for (uint32_t i = lane; i < SCAN_TILE / 16; i += 32)
            {
                uint32_t off = tile_off + i * 16u;
                uint32_t nb = pd.size > off ? min(16u, pd.size - off) : 0u;
                cp_async16(rows + ((i >> 4) + 1u) * SCAN_ROW + (i & 15u) * 16u, nb ? src + i * 16u : arena, nb);
            }
        }
        if (lane < 3)
        {
            // 48-byte halo in front of the tile; a part's first tile sees zeros (those positions can never be cuts:
            // a cut needs at least min >= 48 bytes of the same part in front of it)
            cp_async16(rows + (SCAN_SEG - SCAN_WINDOW) + lane * 16u, tip ? src - SCAN_WINDOW + lane * 16u : arena, tip ? 16u : 0u);
        }
        cp_async_commit();
        // while the copy is in flight: pull the next tile of this warp towards L2 and fetch its descriptor
        const uint32_t next = tile + stride;
        uint32_t next_part = 0;
        if (next < num_tiles)
        {
            next_part = __ldg(&tile_part[next]);
            const uint64_t noff = (uint64_t)tile_off + (uint64_t)stride * SCAN_TILE + lane * 256u;
            if (noff + 256u <= pd.size) // only when the next tile lies in the same part (the common case)
            {
                const uint8_t* nsrc = arena + pd.data_off + noff;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nsrc));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nsrc + 128u));
            }
        }
        cp_async_wait<0>();
        __syncwarp();

        bool overflow = false;
        {
            const uint32_t npend = scan_segment<SPARSE>(my, prev, tab_lane, bits_addr, cp);
            overflow = __any_sync(0xffffffffu, npend > SCAN_PEND_MAX);
        }
        __syncwarp(); // every lane is done with the rows: the next iteration may overwrite them

        // ordered compaction of this tile's candidate bits into its slot list (lane order == position order)
        const uint32_t seg_first = tile_off + lane * SCAN_SEG; // part-relative position of my first byte
        uint32_t words[SCAN_SEG / 32];
        uint32_t cnt = 0;
        {
            // most tiles hold no candidate at all (0.33 expected per tile at the default parameters): one vote decides
            static_assert(SCAN_SEG / 32 == 8, "two 16-byte loads cover a lane's bitmap row");
            const uint4 b0 = lds128(bits_addr), b1 = lds128(bits_addr + 16u);
            words[0] = b0.x; words[1] = b0.y; words[2] = b0.z; words[3] = b0.w;
            words[4] = b1.x; words[5] = b1.y; words[6] = b1.z; words[7] = b1.w;
            const uint32_t orv = (b0.x | b0.y) | (b0.z | b0.w) | (b1.x | b1.y) | (b1.z | b1.w);
            if (orv)
            {
#pragma unroll
                for (int w = 0; w < SCAN_SEG / 32; ++w)
                {
                    uint32_t v = words[w];
                    // always clear: a candidate bit beyond the end of the part (zero-filled tail of a ragged last tile) is masked out
                    // below and must not survive into the next tile this warp scans
                    s_bits[w] = 0;
                    // a cut after byte q is position q+1; keep it only inside the part
                    uint32_t first = seg_first + 32 * w + 1;
                    if (first > pd.size) v = 0;
                    else if (first + 31 > pd.size) v &= (1u << (pd.size - first + 1)) - 1u;
                    words[w] = v;
                    cnt += __popc(v);
                }
            }
        }
        const uint32_t any = __ballot_sync(0xffffffffu, cnt != 0);
        uint32_t total = 0;
        if (any)
        {
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            total = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t base = incl - cnt;
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
        }
        if (lane == 0) tile_count[tile] = overflow ? cp.slots + 1u : total; // > slots: k_hpcdc_walk hashes the queried interval itself
        part_idx = next_part;
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

void launch_tile_part(const PartDesc* d_parts, uint32_t part_count, uint32_t num_tiles, uint32_t* d_tile_part, cudaStream_t st)
{
    if (!part_count || !num_tiles) return;
    k_tile_part<<<(num_tiles + 255) / 256, 256, 0, st>>>(d_parts, part_count, num_tiles, d_tile_part);
}

__global__ void k_probe_dynamic_smem_base(uint32_t* out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    *out = smem_u32(smem);
}

// finds where the 64 KiB-aligned table fits inside the dynamic shared window of this device
cudaError_t make_scan_layout(ScanLayout* lay, cudaStream_t st)
{
    uint32_t* d_base = nullptr;
    uint32_t base = 0;
    cudaError_t e = cudaMalloc(&d_base, sizeof(uint32_t));
    if (e != cudaSuccess) return e;
    k_probe_dynamic_smem_base<<<1, 32, 1024, st>>>(d_base);
    e = cudaMemcpyAsync(&base, d_base, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_base);
    if (e != cudaSuccess) return e;
    int dev = 0, max_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    uint32_t best_total = 0xffffffffu;
    for (uint32_t k = 0; k <= (uint32_t)SCAN_WARPS; ++k)
    {
        uint32_t t0 = (base + k * SCAN_WARP_BYTES + 65535u) & ~65535u; // shared address of the table
        uint32_t total = t0 - base + SCAN_TABLE_BYTES + (SCAN_WARPS - k) * SCAN_WARP_BYTES;
        if (total < best_total)
        {
            best_total = total;
            lay->warps_before = k;
            lay->table_off = t0 - base;
            lay->total_bytes = total;
        }
    }
    if (best_total > (uint32_t)max_optin) return cudaErrorInvalidConfiguration;
    e = cudaFuncSetAttribute(k_hpcdc_scan<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay->total_bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_hpcdc_scan<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay->total_bytes);
}

cudaError_t launch_hpcdc_scan(const uint8_t* d_arena, const PartDesc* d_parts, const uint32_t* d_tile_part, uint32_t num_tiles,
                              const ChunkParams& cp, const uint32_t* d_table, uint32_t* d_tile_count, uint32_t* d_tile_slots,
                              const ScanLayout& lay, int sm_count, cudaStream_t st)
{
    if (!num_tiles) return cudaSuccess;
    uint32_t grid = (num_tiles + SCAN_WARPS - 1) / SCAN_WARPS;
    if (grid > (uint32_t)sm_count) grid = (uint32_t)sm_count;
    uint32_t d_odd = cp.d;
    while (d_odd && !(d_odd & 1u)) d_odd >>= 1;
    if (d_odd >= 1024u) // expected recorded groups per lane-tile = 256 / odd(d) <= 0.25
        k_hpcdc_scan<true><<<grid, SCAN_THREADS, lay.total_bytes, st>>>(d_arena, d_parts, d_tile_part, num_tiles, cp, d_table, d_tile_count, d_tile_slots, lay);
    else
        k_hpcdc_scan<false><<<grid, SCAN_THREADS, lay.total_bytes, st>>>(d_arena, d_parts, d_tile_part, num_tiles, cp, d_table, d_tile_count, d_tile_slots, lay);
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
