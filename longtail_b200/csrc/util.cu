// util.cu — device-wide exclusive scan, first-occurrence dedup table, synthetic asset fill.
#include "lt_device.cuh"
#include "lt_kernels.h"
#include "../../include/lt_synth.h"

namespace ltb {

// ---------------------------------------------------------------- exclusive scan (u32)

namespace {
constexpr int PS_THREADS = 256;
constexpr int PS_ITEMS = 8;
constexpr int PS_BLOCK = PS_THREADS * PS_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t* s_warp, uint32_t* out_total)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0, total = 0;
    for (uint32_t w = 0; w < blockDim.x / 32; ++w)
    {
        uint32_t t = s_warp[w];
        if (w < warp) base += t;
        total += t;
    }
    __syncthreads();
    *out_total = total;
    return base + incl - v;
}
} // namespace

__global__ void __launch_bounds__(PS_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, uint32_t count, uint32_t* __restrict__ block_sums)
{
    __shared__ uint32_t s_warp[PS_THREADS / 32];
    const uint32_t first = blockIdx.x * PS_BLOCK + threadIdx.x * PS_ITEMS;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < PS_ITEMS; ++i)
        if (first + i < count) v += in[first + i];
    uint32_t total;
    block_exclusive(v, s_warp, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of block_sums[0..nblocks), total appended at [nblocks]
__global__ void __launch_bounds__(1024) k_scan_tops(uint32_t* __restrict__ block_sums, uint32_t nblocks)
{
    __shared__ uint32_t s_warp[32];
    uint32_t carry = 0;
    for (uint32_t b = 0; b < nblocks; b += 1024)
    {
        uint32_t i = b + threadIdx.x;
        uint32_t v = i < nblocks ? block_sums[i] : 0;
        uint32_t total;
        uint32_t ex = block_exclusive(v, s_warp, &total);
        if (i < nblocks) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) block_sums[nblocks] = carry;
}

__global__ void __launch_bounds__(PS_THREADS) k_scan_final(const uint32_t* __restrict__ in, uint32_t count, const uint32_t* __restrict__ block_sums,
                                                           uint32_t nblocks, uint32_t* __restrict__ out)
{
    __shared__ uint32_t s_warp[PS_THREADS / 32];
    const uint32_t first = blockIdx.x * PS_BLOCK + threadIdx.x * PS_ITEMS;
    uint32_t item[PS_ITEMS];
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < PS_ITEMS; ++i)
    {
        item[i] = first + i < count ? in[first + i] : 0;
        v += item[i];
    }
    uint32_t total;
    uint32_t run = block_sums[blockIdx.x] + block_exclusive(v, s_warp, &total);
#pragma unroll
    for (int i = 0; i < PS_ITEMS; ++i)
    {
        if (first + i < count) out[first + i] = run;
        run += item[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[count] = block_sums[nblocks];
}

__global__ void k_set_u32(uint32_t* p, uint32_t v) { *p = v; }

size_t scan_tmp_words(uint32_t count) { return (size_t)(count + PS_BLOCK - 1) / PS_BLOCK + 2; }

void launch_exclusive_scan(const uint32_t* d_in, uint32_t count, uint32_t* d_out, uint32_t* d_tmp, cudaStream_t st)
{
    if (count == 0)
    {
        k_set_u32<<<1, 1, 0, st>>>(d_out, 0);
        return;
    }
    const uint32_t nblocks = (count + PS_BLOCK - 1) / PS_BLOCK;
    k_scan_reduce<<<nblocks, PS_THREADS, 0, st>>>(d_in, count, d_tmp);
    k_scan_tops<<<1, 1024, 0, st>>>(d_tmp, nblocks);
    k_scan_final<<<nblocks, PS_THREADS, 0, st>>>(d_in, count, d_tmp, nblocks, d_out);
}

__global__ void k_fill_u32(uint32_t* __restrict__ d, uint32_t value, size_t count)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) d[i] = value;
}

void launch_fill_u32(uint32_t* d, uint32_t value, size_t count, cudaStream_t st)
{
    if (!count) return;
    size_t blocks = (count + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_fill_u32<<<(unsigned)blocks, 256, 0, st>>>(d, value, count);
}

// ---------------------------------------------------------------- exclusive scan (u32 values, u64 sums) and greedy block packing
// Longtail_CreateStoreIndex (src/longtail.c:6796-6860) packs the chunks, in order, into blocks: a block closes on a tag change, at
// max_chunks chunks, or when the next chunk would push it beyond max_block_size + max_block_size / 10.  Sequential as written, but the
// block that STARTS at chunk i ends at a place that depends on i alone — next(i) = min(end of i's tag run, i + max_chunks, first j with
// sum(i..j) > limit) — so next() is computed for every chunk at once (a binary search in the prefix sums), and the chunks reachable from
// chunk 0 through next() are marked by pointer doubling: log2(count) rounds instead of count steps.

namespace {
__device__ __forceinline__ uint64_t block_exclusive64(uint64_t v, uint64_t* s_warp, uint64_t* out_total)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t base = 0, total = 0;
    for (uint32_t w = 0; w < blockDim.x / 32; ++w)
    {
        uint64_t t = s_warp[w];
        if (w < warp) base += t;
        total += t;
    }
    __syncthreads();
    *out_total = total;
    return base + incl - v;
}
} // namespace

__global__ void __launch_bounds__(PS_THREADS) k_scan64_reduce(const uint32_t* __restrict__ in, uint32_t count, uint64_t* __restrict__ block_sums)
{
    __shared__ uint64_t s_warp[PS_THREADS / 32];
    const uint32_t first = blockIdx.x * PS_BLOCK + threadIdx.x * PS_ITEMS;
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < PS_ITEMS; ++i)
        if (first + i < count) v += in[first + i];
    uint64_t total;
    block_exclusive64(v, s_warp, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k_scan64_tops(uint64_t* __restrict__ block_sums, uint32_t nblocks)
{
    __shared__ uint64_t s_warp[32];
    uint64_t carry = 0;
    for (uint32_t b = 0; b < nblocks; b += 1024)
    {
        uint32_t i = b + threadIdx.x;
        uint64_t v = i < nblocks ? block_sums[i] : 0;
        uint64_t total;
        uint64_t ex = block_exclusive64(v, s_warp, &total);
        if (i < nblocks) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) block_sums[nblocks] = carry;
}
__global__ void __launch_bounds__(PS_THREADS) k_scan64_final(const uint32_t* __restrict__ in, uint32_t count, const uint64_t* __restrict__ block_sums,
                                                             uint32_t nblocks, uint64_t* __restrict__ out)
{
    __shared__ uint64_t s_warp[PS_THREADS / 32];
    const uint32_t first = blockIdx.x * PS_BLOCK + threadIdx.x * PS_ITEMS;
    uint32_t item[PS_ITEMS];
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < PS_ITEMS; ++i)
    {
        item[i] = first + i < count ? in[first + i] : 0;
        v += item[i];
    }
    uint64_t total;
    uint64_t run = block_sums[blockIdx.x] + block_exclusive64(v, s_warp, &total);
#pragma unroll
    for (int i = 0; i < PS_ITEMS; ++i)
    {
        if (first + i < count) out[first + i] = run;
        run += item[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[count] = block_sums[nblocks];
}
size_t scan64_tmp_words(uint32_t count) { return (size_t)(count + PS_BLOCK - 1) / PS_BLOCK + 2; }
void launch_exclusive_scan64(const uint32_t* d_in, uint32_t count, uint64_t* d_out, uint64_t* d_tmp, cudaStream_t st)
{
    if (count == 0)
    {
        cudaMemsetAsync(d_out, 0, sizeof(uint64_t), st);
        return;
    }
    const uint32_t nblocks = (count + PS_BLOCK - 1) / PS_BLOCK;
    k_scan64_reduce<<<nblocks, PS_THREADS, 0, st>>>(d_in, count, d_tmp);
    k_scan64_tops<<<1, 1024, 0, st>>>(d_tmp, nblocks);
    k_scan64_final<<<nblocks, PS_THREADS, 0, st>>>(d_in, count, d_tmp, nblocks, d_out);
}

// flag[i] = 1 where a tag run starts
__global__ void k_pack_run_flags(const uint32_t* __restrict__ tag, uint32_t count, uint32_t* __restrict__ flag)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) flag[i] = (i == 0 || tag[i] != tag[i - 1]) ? 1u : 0u;
}
// run_start[r] = first chunk of tag run r (rscan = exclusive scan of flag)
__global__ void k_pack_run_starts(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rscan, uint32_t count, uint32_t* __restrict__ run_start)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count && flag[i]) run_start[rscan[i]] = i;
}
// next[i] = where the block that starts at chunk i ends (= where the following block starts); next[count] = count
__global__ void k_pack_next(const uint64_t* __restrict__ prefix, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rscan,
                            const uint32_t* __restrict__ run_start, uint32_t count, uint64_t limit, uint32_t max_chunks, uint32_t* __restrict__ next)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > count) return;
    if (i == count)
    {
        next[i] = count;
        return;
    }
    const uint32_t runs = rscan[count];
    const uint32_t rid = rscan[i] + flag[i] - 1u;
    uint32_t e = rid + 1u < runs ? run_start[rid + 1u] : count;        // tag change
    if (max_chunks < e - i) e = i + max_chunks;                        // chunk count
    // size: chunk j joins while prefix[j + 1] - prefix[i] <= limit; the first chunk always does
    const uint64_t t = prefix[i] + limit;
    uint32_t lo = i + 1u, hi = e; // smallest j in [i + 1, e) with prefix[j + 1] > t, else e
    while (lo < hi)
    {
        const uint32_t mid = lo + (hi - lo) / 2u;
        if (prefix[mid + 1u] > t) hi = mid; else lo = mid + 1u;
    }
    next[i] = lo;
}
// one round of pointer doubling: every marked chunk marks the chunk 2^r blocks ahead; jump_out = jump o jump
__global__ void k_pack_jump(const uint32_t* __restrict__ jump_in, uint32_t* __restrict__ jump_out, uint32_t* __restrict__ mark, uint32_t count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > count) return;
    const uint32_t j = jump_in[i];
    jump_out[i] = jump_in[j];
    if (i < count && j < count && mark[i]) mark[j] = 1u;
}
// blk_first[b] = first chunk of block b, blk_end[b] = payload bytes of blocks 0 .. b (mscan = exclusive scan of mark)
__global__ void k_pack_emit(const uint32_t* __restrict__ mark, const uint32_t* __restrict__ mscan, const uint32_t* __restrict__ next,
                            const uint64_t* __restrict__ prefix, uint32_t count, uint32_t* __restrict__ blk_first, uint64_t* __restrict__ blk_end)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count || !mark[i]) return;
    const uint32_t b = mscan[i];
    blk_first[b] = i;
    blk_end[b] = prefix[next[i]];
}

void launch_pack_blocks(const uint32_t* d_len, const uint32_t* d_tag, uint32_t count, uint64_t limit, uint32_t max_chunks, const PackBuffers& b, cudaStream_t st)
{
    if (!count) return;
    const uint32_t g = (count + 1 + 255) / 256;
    launch_exclusive_scan64(d_len, count, b.prefix, b.tmp64, st);
    k_pack_run_flags<<<g, 256, 0, st>>>(d_tag, count, b.flag);
    launch_exclusive_scan(b.flag, count, b.rscan, b.tmp32, st);
    k_pack_run_starts<<<g, 256, 0, st>>>(b.flag, b.rscan, count, b.run_start);
    k_pack_next<<<g, 256, 0, st>>>(b.prefix, b.flag, b.rscan, b.run_start, count, limit, max_chunks, b.next);
    cudaMemsetAsync(b.mark, 0, sizeof(uint32_t) * (size_t)count, st);
    const uint32_t one = 1;
    cudaMemcpyAsync(b.mark, &one, sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    // jump_a starts as next; after r rounds the marked set is everything within 2^r blocks of chunk 0
    cudaMemcpyAsync(b.jump_a, b.next, sizeof(uint32_t) * ((size_t)count + 1), cudaMemcpyDeviceToDevice, st);
    uint32_t* in = b.jump_a;
    uint32_t* out = b.jump_b;
    for (uint64_t reach = 1; reach < (uint64_t)count + 1; reach <<= 1)
    {
        k_pack_jump<<<g, 256, 0, st>>>(in, out, b.mark, count);
        uint32_t* t = in; in = out; out = t;
    }
    launch_exclusive_scan(b.mark, count, b.mscan, b.tmp32, st);
    k_pack_emit<<<g, 256, 0, st>>>(b.mark, b.mscan, b.next, b.prefix, count, b.blk_first, b.blk_end);
}

// ---------------------------------------------------------------- first-occurrence dedup
// The reference keeps, for every distinct chunk hash, the first occurrence in asset/part/chunk order
// (src/longtail.c:2952-2970, LookupTable_PutUnique).  Order-independent restatement: the representative of a
// hash is the occurrence with the smallest ordinal -> atomicMin into an open-addressing table.

namespace {
constexpr uint64_t DEDUP_EMPTY = 0xffffffffffffffffull;
__device__ __forceinline__ uint32_t dedup_slot(uint64_t key, uint32_t mask)
{
    return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 32) & mask;
}
} // namespace

__global__ void k_dedup_insert(const uint64_t* __restrict__ hash, uint32_t count, uint64_t* keys, uint32_t* vals, uint32_t capacity)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t key = hash[i];
    if (key == DEDUP_EMPTY)
    {
        atomicMin(&vals[capacity], i); // the one key that cannot live in the table has a side slot
        return;
    }
    const uint32_t mask = capacity - 1;
    for (uint32_t s = dedup_slot(key, mask);; s = (s + 1) & mask)
    {
        unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(&keys[s]), (unsigned long long)DEDUP_EMPTY, (unsigned long long)key);
        if (prev == DEDUP_EMPTY || prev == key)
        {
            atomicMin(&vals[s], i);
            return;
        }
    }
}

__global__ void k_dedup_lookup(const uint64_t* __restrict__ hash, uint32_t count, const uint64_t* __restrict__ keys,
                               const uint32_t* __restrict__ vals, uint32_t capacity, uint32_t* __restrict__ first,
                               uint32_t* __restrict__ is_first)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t key = hash[i];
    uint32_t f;
    if (key == DEDUP_EMPTY)
        f = vals[capacity];
    else
    {
        const uint32_t mask = capacity - 1;
        uint32_t s = dedup_slot(key, mask);
        while (keys[s] != key) s = (s + 1) & mask;
        f = vals[s];
    }
    first[i] = f;
    is_first[i] = f == i ? 1u : 0u;
}

// ---- the same, split by hash across `world` ranks: every rank holds the whole (allgathered) hash list but inserts and resolves only
// the keys it owns; an all-reduce (sum) of `first` — zero where a rank does not own the key — then gives every rank every answer, and
// the random-access part of the dedup is 1/world of the single-GPU work on each rank.
__device__ __forceinline__ uint32_t dedup_owner(uint64_t key, uint32_t world) { return (uint32_t)(((key * 0xD6E8FEB86659FD93ull) >> 40) % world); }

__global__ void k_dedup_insert_part(const uint64_t* __restrict__ hash, uint32_t count, uint64_t* keys, uint32_t* vals, uint32_t capacity,
                                    uint32_t world, uint32_t rank)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t key = hash[i];
    if (dedup_owner(key, world) != rank) return;
    if (key == DEDUP_EMPTY)
    {
        atomicMin(&vals[capacity], i);
        return;
    }
    const uint32_t mask = capacity - 1;
    for (uint32_t s = dedup_slot(key, mask);; s = (s + 1) & mask)
    {
        unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(&keys[s]), (unsigned long long)DEDUP_EMPTY, (unsigned long long)key);
        if (prev == DEDUP_EMPTY || prev == key)
        {
            atomicMin(&vals[s], i);
            return;
        }
    }
}

__global__ void k_dedup_lookup_part(const uint64_t* __restrict__ hash, uint32_t count, const uint64_t* __restrict__ keys,
                                    const uint32_t* __restrict__ vals, uint32_t capacity, uint32_t* __restrict__ first, uint32_t world, uint32_t rank)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t key = hash[i];
    uint32_t f = 0;
    if (dedup_owner(key, world) == rank)
    {
        if (key == DEDUP_EMPTY)
            f = vals[capacity];
        else
        {
            const uint32_t mask = capacity - 1;
            uint32_t s = dedup_slot(key, mask);
            while (keys[s] != key) s = (s + 1) & mask;
            f = vals[s];
        }
    }
    first[i] = f;
}

__global__ void k_mark_first(const uint32_t* __restrict__ first, uint32_t count, uint32_t* __restrict__ is_first)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) is_first[i] = first[i] == i ? 1u : 0u;
}

void launch_dedup_insert_part(const uint64_t* d_hash, uint32_t count, const DedupBuffers& b, uint32_t world, uint32_t rank, cudaStream_t st)
{
    if (!count) return;
    k_dedup_insert_part<<<(count + 255) / 256, 256, 0, st>>>(d_hash, count, b.keys, b.vals, b.capacity, world, rank);
}
void launch_dedup_lookup_part(const uint64_t* d_hash, uint32_t count, const DedupBuffers& b, uint32_t world, uint32_t rank, cudaStream_t st)
{
    if (!count) return;
    k_dedup_lookup_part<<<(count + 255) / 256, 256, 0, st>>>(d_hash, count, b.keys, b.vals, b.capacity, b.first, world, rank);
}
void launch_mark_first(const DedupBuffers& b, uint32_t count, cudaStream_t st)
{
    if (!count) return;
    k_mark_first<<<(count + 255) / 256, 256, 0, st>>>(b.first, count, b.is_first);
}

__global__ void k_dedup_emit(const uint64_t* __restrict__ hash, const uint32_t* __restrict__ len, const uint32_t* __restrict__ tag,
                             uint32_t count, const uint32_t* __restrict__ first, const uint32_t* __restrict__ is_first,
                             const uint32_t* __restrict__ uidx, uint32_t* __restrict__ asset_chunk_index,
                             uint64_t* __restrict__ unique_hash, uint32_t* __restrict__ unique_len, uint32_t* __restrict__ unique_tag,
                             const uint64_t* __restrict__ chunk_off, uint64_t* __restrict__ unique_off, uint32_t* __restrict__ unique_first)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    asset_chunk_index[i] = uidx[first[i]];
    if (is_first[i])
    {
        const uint32_t u = uidx[i];
        if (unique_first) unique_first[u] = i; // ordinal of the first occurrence: which rank holds the bytes, and where
        unique_hash[u] = hash[i];
        unique_len[u] = len[i];
        unique_tag[u] = tag[i]; // tag of the first occurrence (src/longtail.c:2962)
        if (chunk_off) unique_off[u] = chunk_off[i]; // where the first occurrence's bytes live (CreateAssetPartLookup, :4429-4500)
    }
}

// membership of every hash in the set previously inserted with k_dedup_insert (DiffHashes, src/longtail.c:6620-6743: which of the
// version's chunks does the store already hold): 1 = present
__global__ void k_set_contains(const uint64_t* __restrict__ hash, uint32_t count, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                               uint32_t capacity, uint8_t* __restrict__ present)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t key = hash[i];
    bool found;
    if (key == DEDUP_EMPTY)
        found = vals[capacity] != 0xffffffffu;
    else
    {
        const uint32_t mask = capacity - 1;
        uint32_t s = dedup_slot(key, mask);
        while (keys[s] != key && keys[s] != DEDUP_EMPTY) s = (s + 1) & mask;
        found = keys[s] == key;
    }
    present[i] = found ? 1 : 0;
}

void launch_set_contains(const uint64_t* d_hash, uint32_t count, const DedupBuffers& b, uint8_t* d_present, cudaStream_t st)
{
    if (!count) return;
    k_set_contains<<<(count + 255) / 256, 256, 0, st>>>(d_hash, count, b.keys, b.vals, b.capacity, d_present);
}

void launch_dedup_insert(const uint64_t* d_hash, uint32_t count, const DedupBuffers& b, cudaStream_t st)
{
    if (!count) return;
    k_dedup_insert<<<(count + 255) / 256, 256, 0, st>>>(d_hash, count, b.keys, b.vals, b.capacity);
}

void launch_dedup_lookup(const uint64_t* d_hash, uint32_t count, const DedupBuffers& b, cudaStream_t st)
{
    if (!count) return;
    k_dedup_lookup<<<(count + 255) / 256, 256, 0, st>>>(d_hash, count, b.keys, b.vals, b.capacity, b.first, b.is_first);
}

void launch_dedup_emit(const uint64_t* d_hash, const uint32_t* d_len, const uint32_t* d_tag, uint32_t count, const DedupBuffers& b,
                       uint32_t* d_asset_chunk_index, uint64_t* d_unique_hash, uint32_t* d_unique_len, uint32_t* d_unique_tag,
                       const uint64_t* d_chunk_off, uint64_t* d_unique_off, cudaStream_t st, uint32_t* d_unique_first)
{
    if (!count) return;
    k_dedup_emit<<<(count + 255) / 256, 256, 0, st>>>(d_hash, d_len, d_tag, count, b.first, b.is_first, b.uidx, d_asset_chunk_index,
                                                      d_unique_hash, d_unique_len, d_unique_tag, d_chunk_off, d_unique_off, d_unique_first);
}

// ---------------------------------------------------------------- synthetic assets (bench / test inputs only)

__global__ void k_synth_fill(uint8_t* __restrict__ dst, uint64_t len, lt_synth_spec spec, uint64_t asset_id, uint64_t offset)
{
    const uint64_t blocks = (len + 15) / 16;
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; b < blocks; b += stride)
    {
        const uint64_t off = offset + b * 16;
        uint8_t tmp[16];
        lt_synth_asset_block16(&spec, asset_id, off, tmp);
        if (b * 16 + 16 <= len && (((uintptr_t)dst) & 15) == 0)
            *reinterpret_cast<uint4*>(dst + b * 16) = *reinterpret_cast<uint4*>(tmp);
        else
            for (uint64_t i = 0; i < 16 && b * 16 + i < len; ++i) dst[b * 16 + i] = tmp[i];
    }
}

void launch_synth_fill(uint8_t* d_dst, uint64_t len, const lt_synth_spec_dev& s, uint64_t asset_id, uint64_t offset, cudaStream_t st)
{
    if (!len) return;
    lt_synth_spec spec;
    spec.seed = s.seed;
    spec.shared_permille = s.shared_permille;
    spec.pool_segments = s.pool_segments;
    spec.class_mode = s.class_mode;
    spec.reserved = 0;
    uint64_t blocks = ((len + 15) / 16 + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    k_synth_fill<<<(unsigned)blocks, 256, 0, st>>>(d_dst, len, spec, asset_id, offset);
}

} // namespace ltb
