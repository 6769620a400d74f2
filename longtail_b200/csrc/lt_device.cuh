// lt_device.cuh — shared declarations for the sm_100a kernels of the chunk -> hash -> compress path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ltb {

// ---------------------------------------------------------------- geometry of the Buzhash scan
// One WARP scans one tile, independently of every other warp (no block barriers in the steady state).  Every lane owns
// SCAN_SEG contiguous bytes; rows are padded by 16 B in shared memory so that the 32 lanes' 16-byte reads fall into
// distinct bank groups.  16 warps and the 64 KiB bank-replicated table share one SM.
constexpr int SCAN_WARPS = 16;
constexpr int SCAN_THREADS = SCAN_WARPS * 32;
constexpr int SCAN_SEG = 256;
constexpr int SCAN_TILE = 32 * SCAN_SEG;           // 8192 bytes of asset data per (warp) tile
constexpr int SCAN_ROW = SCAN_SEG + 16;
constexpr int SCAN_WINDOW = 48;                    // lib/hpcdcchunker/longtail_hpcdcchunker.c:12
constexpr int SCAN_TABLE_BYTES = 256 * 256;        // 256 entries x (32 lanes x T, 32 lanes x rotl16(T))
constexpr int SCAN_ROWS_BYTES = 33 * SCAN_ROW;     // halo row + 32 lane rows
constexpr int SCAN_BITMAP_BYTES = 32 * (SCAN_SEG / 32) * 4;
constexpr int SCAN_WARP_BYTES = SCAN_ROWS_BYTES + SCAN_BITMAP_BYTES;
constexpr int SCAN_SMEM_BYTES = SCAN_TABLE_BYTES + SCAN_WARPS * SCAN_WARP_BYTES;

struct ScanLayout
{
    uint32_t table_off;    // offset of the table inside the dynamic shared window (its shared ADDRESS is a multiple of 64 KiB)
    uint32_t warps_before; // how many per-warp buffers lie in front of the table
    uint32_t total_bytes;  // dynamic shared memory to request
};

constexpr uint32_t CAND_OVERFLOW = 0x80000000u;    // dense-list marker: tile whose slot list overflowed

struct PartDesc
{
    uint64_t data_off;    // byte offset of the part in the arena (multiple of 16)
    uint32_t size;        // bytes in this part (<= target_chunk_size * 1024)
    uint32_t tile_start;  // first global tile index of this part
    uint32_t chunk_start; // first slot of this part in the per-part chunk staging arrays
    uint32_t asset;       // index of the asset this part belongs to
    uint32_t tag;         // the asset's compression tag (copied to every chunk of the part)
    uint32_t pad;
};

struct ChunkParams
{
    uint32_t min, avg, max; // src/longtail.c:1985-1987
    uint32_t d;             // discriminator, longtail_hpcdcchunker.c:126-129 (computed on the host in double)
    uint32_t d_odd_inv;     // inverse mod 2^32 of the odd part of d
    uint32_t d_odd_thr;     // floor((2^32-1) / odd part of d)
    uint32_t slots;         // candidate slots per tile
};

__device__ __forceinline__ uint32_t rotl32(uint32_t x, uint32_t r) { return __funnelshift_l(x, x, r); }
__device__ __forceinline__ uint32_t rotr32(uint32_t x, uint32_t r) { return __funnelshift_r(x, x, r); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32_off128(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+128];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// 16-byte async copy global -> shared; bytes beyond src_bytes are zero-filled (src_bytes may be 0)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16_full(uint32_t dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

} // namespace ltb
