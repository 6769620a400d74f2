// lt_kernels.h — host-callable launchers of the sm_100a kernels (internal to liblongtail_b200.so)
#pragma once

#include "lt_device.cuh"

namespace ltb {

// ---- hpcdc.cu
void launch_tile_part(const PartDesc* d_parts, uint32_t part_count, uint32_t num_tiles, uint32_t* d_tile_part, cudaStream_t st);
cudaError_t launch_hpcdc_scan(const uint8_t* d_arena, const PartDesc* d_parts, const uint32_t* d_tile_part, uint32_t num_tiles,
                              const ChunkParams& cp, const uint32_t* d_table, uint32_t* d_tile_count, uint32_t* d_tile_slots,
                              const ScanLayout& lay, int sm_count, cudaStream_t st);
cudaError_t make_scan_layout(ScanLayout* lay, cudaStream_t st);
void launch_hpcdc_walk(const uint8_t* d_arena, const PartDesc* d_parts, uint32_t part_count, const ChunkParams& cp,
                       const uint32_t* d_table, const uint32_t* d_tile_count, const uint32_t* d_tile_slots, uint32_t* d_cand,
                       uint64_t* d_stage_off, uint32_t* d_stage_len, uint32_t* d_part_chunk_count, cudaStream_t st);
void launch_compact_chunks(const PartDesc* d_parts, uint32_t part_count, const uint32_t* d_part_chunk_count,
                           const uint32_t* d_part_chunk_base, const uint64_t* d_stage_off, const uint32_t* d_stage_len,
                           uint64_t* d_chunk_off, uint32_t* d_chunk_len, uint32_t* d_chunk_tag, cudaStream_t st);

// ---- blake3.cu : hash `count` byte segments [off, off+len) of `base` (device memory of `base_size` bytes)
// leaf_prefix has count+1 entries: exclusive prefix of max(1, ceil(len/1024)); cvs holds 8 u32 per leaf.
void launch_leaf_counts(const uint32_t* d_len, uint32_t count, uint32_t* d_leaf_count, cudaStream_t st);
constexpr uint32_t BLAKE3_MAX_LEVELS = 24; // segments are < 4 GiB = 2^22 leaves
void launch_blake3_leaves(const uint8_t* d_base, uint64_t base_size, const uint64_t* d_off, const uint32_t* d_len,
                          const uint32_t* d_leaf_prefix, uint32_t count, uint32_t total_leaves, uint32_t* d_cvs,
                          uint64_t* d_hash_out, uint4* d_merge_items, uint32_t* d_merge_counts, cudaStream_t st);
uint32_t launch_blake3_merge(uint32_t total_leaves, uint32_t segment_count, uint32_t max_segment_leaves, uint32_t* d_cvs, uint64_t* d_hash_out,
                             uint4* d_items_a, uint4* d_items_b, uint32_t* d_merge_counts, cudaStream_t st);

// ---- blake2s.cu : BLAKE2s-64 over segments; d_counter is one u32 of scratch (the work queue head)
void launch_blake2s_segments(const uint8_t* d_base, const uint64_t* d_off, const uint32_t* d_len, uint32_t count, uint32_t* d_counter,
                             uint64_t* d_hash_out, int sm_count, cudaStream_t st);

// ---- meow.cu : Meow 0.5 (low 64 bits) over segments; d_td0 = the 256-entry inverse T-table of meow_build_table in device memory
void meow_build_table(uint32_t out[256]);
cudaError_t launch_meow_segments(const uint8_t* d_base, const uint64_t* d_off, const uint32_t* d_len, uint32_t count, uint32_t* d_counter,
                                 uint64_t* d_hash_out, const uint32_t* d_td0, int sm_count, cudaStream_t st);

// ---- lz4.cu
uint32_t lz4_copy_job_capacity(uint32_t raw_len);
// In-place layout of the write path: a block the shared-memory-table encoder takes is gathered to the END of its own output slot and
// compressed towards the front (the output never reaches source bytes the parse can still refer to when the source starts
// lz4_in_place_offset() bytes behind the output; literal runs are then copied by the parsing warp, copy_job_start = LZ4_JOBS_INLINE).
bool lz4_block_in_place(uint32_t raw_len);
uint64_t lz4_in_place_offset(uint32_t raw_len);
constexpr uint32_t LZ4_JOBS_INLINE = 0xffffffffu;
cudaError_t launch_lz4_decode(const uint8_t* d_in, const uint64_t* d_in_off, const uint32_t* d_in_len, uint8_t* d_out,
                              const uint64_t* d_out_off, const uint32_t* d_out_cap, uint32_t* d_out_len, uint32_t block_count, cudaStream_t st);
// d_v2_scratch: lz4_v2_scratch_bytes() bytes (the work-queue head of the shared-memory-table encoder)
size_t lz4_v2_scratch_bytes();
cudaError_t launch_lz4_blocks(const uint8_t* d_raw, const uint64_t* d_raw_off, const uint32_t* d_raw_len, uint8_t* d_out,
                              const uint64_t* d_out_off, uint32_t* d_out_len, uint3* d_copy_jobs, const uint32_t* d_copy_job_start,
                              uint32_t* d_copy_job_count, uint32_t block_count, uint32_t* d_tables, void* d_v2_scratch, int sm_count,
                              cudaStream_t st);
constexpr size_t LZ4_TABLE_BYTES_PER_BLOCK = 32768; // d_tables: one hash table per block (nullptr = shared memory, 7 warps per SM)
void launch_block_headers(uint8_t* d_out, const uint64_t* d_img_off, const uint32_t* d_blk_first, const uint32_t* d_blk_count,
                          const uint32_t* d_blk_tag, const uint64_t* d_blk_hash, const uint64_t* d_chunk_hashes, const uint32_t* d_chunk_sizes,
                          uint32_t hash_type, uint32_t nb, cudaStream_t st);
// bytes [src, src + len) of device-mapped pinned host memory -> [dst, dst + len) of device memory, both 16-byte aligned; read by the SMs
struct UploadSeg { const uint8_t* src; uint8_t* dst; uint64_t len; };
void launch_upload_segments(const UploadSeg* d_segs, uint32_t count, cudaStream_t st);
void launch_gather_chunks(const uint8_t* d_arena, const uint64_t* d_src_off, const uint64_t* d_dst_off, const uint32_t* d_len,
                          uint8_t* d_out, uint32_t count, cudaStream_t st);

// ---- zstd.cu : ZStd level 3 frames ('ztd1' / 'ztd2'), one warp per frame from a queue; output = {u32 raw, u32 comp} + frame.
// d_workers = zstd_worker_count() slabs of zstd_worker_bytes(); d_out_len[i] = 8 + frame size, 0xffffffff on an encoder error
size_t zstd_worker_bytes();
size_t zstd_worker_phase_offset(); // 4 x u64 cycle counters inside a worker slab (matcher, literals, sequences, copy-out)
uint32_t zstd_worker_count(uint32_t frame_count, int sm_count);
cudaError_t launch_zstd_frames(const uint8_t* d_raw, const uint64_t* d_raw_off, const uint32_t* d_raw_len, uint8_t* d_out, const uint64_t* d_out_off,
                               uint32_t* d_out_len, uint32_t frame_count, void* d_workers, uint32_t worker_count, uint32_t* d_queue, cudaStream_t st);

// ---- zstd_dec.cu : ZStd frame decoder (any level's frames), one warp per frame from a queue; d_out_len[i] = 0xffffffff on a malformed frame
size_t zstd_dec_worker_bytes();
uint32_t zstd_dec_worker_count(uint32_t frame_count, int sm_count);
cudaError_t launch_zstd_decode(const uint8_t* d_in, const uint64_t* d_in_off, const uint32_t* d_in_len, uint8_t* d_out, const uint64_t* d_out_off,
                               const uint32_t* d_out_cap, uint32_t* d_out_len, uint32_t frame_count, void* d_workers, uint32_t worker_count,
                               uint32_t* d_queue, cudaStream_t st);

// ---- util.cu
// exclusive prefix sum of count u32 values into out[0..count]; out[count] = total.  tmp: >= scan_tmp_words(count) u32.
size_t scan_tmp_words(uint32_t count);
void launch_exclusive_scan(const uint32_t* d_in, uint32_t count, uint32_t* d_out, uint32_t* d_tmp, cudaStream_t st);

// exclusive prefix sum of count u32 values as u64 sums into out[0..count]; out[count] = total.  tmp: >= scan64_tmp_words(count) u64.
size_t scan64_tmp_words(uint32_t count);
void launch_exclusive_scan64(const uint32_t* d_in, uint32_t count, uint64_t* d_out, uint64_t* d_tmp, cudaStream_t st);

// Longtail_CreateStoreIndex's greedy packing (src/longtail.c:6796-6860) of `count` chunks (sizes d_len, tags d_tag, store order) on the
// device: next(i) for every chunk by binary search in the prefix sums, the block starts marked by pointer doubling from chunk 0.
// Results: mscan[count] = block count B; blk_first[b] (b < B) = first chunk of block b; blk_end[b] = payload bytes of blocks 0..b;
// prefix[i] = bytes of chunks 0..i-1.  All arrays hold count + 1 entries (tmp32: scan_tmp_words(count + 1), tmp64: scan64_tmp_words(count)).
struct PackBuffers
{
    uint64_t* prefix;
    uint64_t* tmp64;
    uint64_t* blk_end;
    uint32_t *flag, *rscan, *run_start, *next, *jump_a, *jump_b, *mark, *mscan, *tmp32, *blk_first;
};
void launch_pack_blocks(const uint32_t* d_len, const uint32_t* d_tag, uint32_t count, uint64_t limit, uint32_t max_chunks, const PackBuffers& b, cudaStream_t st);

// first-occurrence dedup of chunk hashes (src/longtail.c:2952-2970)
struct DedupBuffers
{
    uint64_t* keys;     // [capacity] open-addressing table, capacity is a power of two >= 2 * count
    uint32_t* vals;     // [capacity] smallest ordinal seen for the key
    uint32_t capacity;
    uint32_t* first;    // [count]   ordinal of the first occurrence of chunk i's hash
    uint32_t* is_first; // [count]   1 when chunk i is that first occurrence
    uint32_t* uidx;     // [count+1] exclusive scan of is_first
};
void launch_dedup_insert(const uint64_t* d_hash, uint32_t count, const DedupBuffers& b, cudaStream_t st);
void launch_dedup_lookup(const uint64_t* d_hash, uint32_t count, const DedupBuffers& b, cudaStream_t st);
// present[i] = 1 when d_hash[i] was inserted into b (launch_dedup_insert) before
void launch_set_contains(const uint64_t* d_hash, uint32_t count, const DedupBuffers& b, uint8_t* d_present, cudaStream_t st);
void launch_dedup_emit(const uint64_t* d_hash, const uint32_t* d_len, const uint32_t* d_tag, uint32_t count, const DedupBuffers& b,
                       uint32_t* d_asset_chunk_index, uint64_t* d_unique_hash, uint32_t* d_unique_len, uint32_t* d_unique_tag,
                       const uint64_t* d_chunk_off, uint64_t* d_unique_off, cudaStream_t st, uint32_t* d_unique_first = nullptr);
// the dedup split by hash across ranks (every rank holds the whole hash list): insert / resolve only the keys rank `rank` owns, first[] = 0
// elsewhere; after an all-reduce (sum) of first[] over the ranks, launch_mark_first derives is_first[]
void launch_dedup_insert_part(const uint64_t* d_hash, uint32_t count, const DedupBuffers& b, uint32_t world, uint32_t rank, cudaStream_t st);
void launch_dedup_lookup_part(const uint64_t* d_hash, uint32_t count, const DedupBuffers& b, uint32_t world, uint32_t rank, cudaStream_t st);
void launch_mark_first(const DedupBuffers& b, uint32_t count, cudaStream_t st);
void launch_fill_u32(uint32_t* d, uint32_t value, size_t count, cudaStream_t st);

// ---- synth.cu : deterministic synthetic asset bytes (include/lt_synth.h)
struct lt_synth_spec_dev
{
    uint64_t seed;
    uint32_t shared_permille, pool_segments, class_mode, reserved;
};
void launch_synth_fill(uint8_t* d_dst, uint64_t len, const lt_synth_spec_dev& spec, uint64_t asset_id, uint64_t offset, cudaStream_t st);

} // namespace ltb
