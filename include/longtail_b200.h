/* longtail_b200.h — C ABI of liblongtail_b200.so: the B200 (sm_100a) implementation of longtail's
 * chunk -> hash -> compress indexing hot path.
 *
 * Plain C: pointers, sizes and errno-style int returns (0 = success), exactly like the reference's own
 * API (src/longtail.h).  No CUDA or torch types appear in any signature; device memory is passed as
 * `void*` / `const uint8_t*` device addresses (cudaMalloc'd by the caller, by torch, or by
 * lt_b200_device_alloc).
 *
 * Layers, bottom up:
 *   1. device primitives over data already resident in HBM      (lt_b200_chunk_ranges, lt_b200_hash_segments, ...)
 *   2. batched verbs replacing the reference's job fan-out       (lt_b200_index_assets ~ ChunkAssets + CreateVersionIndex)
 *   3. drop-in Longtail_*API objects and verbs                   (include/longtail_b200_api.h)
 *
 * Reference interfaces replaced (all paths relative to the reference tree):
 *   lt_b200_chunk_ranges        <- Longtail_ChunkerAPI.NextChunk loop, lib/hpcdcchunker/longtail_hpcdcchunker.c:225-310,
 *                                  driven per part by DynamicChunking, src/longtail.c:1989-2311
 *   lt_b200_hash_segments       <- Longtail_HashAPI.HashBuffer, src/longtail.h:209-217 (lib/blake3/longtail_blake3.c:81-102)
 *   lt_b200_index_*             <- ChunkAssets + Longtail_CreateVersionIndex, src/longtail.c:2343-2550, :2808-3017
 */
#ifndef LONGTAIL_B200_H
#define LONGTAIL_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define LT_B200_EXPORT __attribute__((visibility("default")))
#else
#define LT_B200_EXPORT
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define LT_B200_HASH_BLAKE3 0x626c6b33u /* 'blk3', lib/blake3/longtail_blake3.c:6 */
#define LT_B200_HASH_BLAKE2 0x626c6b32u /* 'blk2', lib/blake2/longtail_blake2.c:9 */
#define LT_B200_HASH_MEOW 0x6d656f77u   /* 'meow', lib/meowhash/longtail_meowhash.c:7 */
#define LT_B200_COMPRESSION_LZ4 0x6c7a3432u /* 'lz42', lib/lz4/longtail_lz4.c:10 */
#define LT_B200_COMPRESSION_ZSTD_MIN 0x7a746431u     /* 'ztd1' -> ZStd level 0 == default == 3, lib/zstd/longtail_zstd.c:19,47 */
#define LT_B200_COMPRESSION_ZSTD_DEFAULT 0x7a746432u /* 'ztd2' -> ZStd level 3, lib/zstd/longtail_zstd.c:20,49 */

typedef struct lt_b200_context lt_b200_context;

/* ---- context: one per GPU (one process per GPU in multi-GPU runs) */
LT_B200_EXPORT int lt_b200_context_create(int device_ordinal, lt_b200_context** out_context);
LT_B200_EXPORT void lt_b200_context_destroy(lt_b200_context* context);
/* text of the last failure on this context ("" when none); valid until the next call */
LT_B200_EXPORT const char* lt_b200_last_error(const lt_b200_context* context);
/* number of kernel launches issued by this context since creation (bench.py reports it as gpu_launches) */
LT_B200_EXPORT uint64_t lt_b200_launch_count(const lt_b200_context* context);
/* frees the context's grow-only device workspace (it comes back on demand); the resident chunk table of the last index call is dropped */
LT_B200_EXPORT int lt_b200_trim(lt_b200_context* context);
/* blocks until all work queued by this context is done */
LT_B200_EXPORT int lt_b200_synchronize(lt_b200_context* context);
/* the CUDA stream (cudaStream_t as void*) this context launches on — for CUDA-event timing by the caller */
LT_B200_EXPORT void* lt_b200_stream(lt_b200_context* context);

/* Optional per-kernel timing with CUDA events on the context stream (what bench.py's roofline block reads).
 * `bytes` is the algorithmic byte count the launches covered (asset bytes for the scan and the leaf hash). */
enum
{
    LT_B200_KERNEL_HPCDC_SCAN = 0,
    LT_B200_KERNEL_HPCDC_WALK = 1,
    LT_B200_KERNEL_BLAKE3_LEAVES = 2,
    LT_B200_KERNEL_BLAKE3_MERGE = 3,
    LT_B200_KERNEL_GATHER = 4,
    LT_B200_KERNEL_LZ4 = 5,
    LT_B200_KERNEL_BLAKE2S = 6,
    LT_B200_KERNEL_MEOW = 7,
    LT_B200_KERNEL_ZSTD = 8,
    LT_B200_KERNEL_LZ4_DECODE = 9,
    LT_B200_KERNEL_ZSTD_DECODE = 10,
    LT_B200_KERNEL_COUNT = 16
};
LT_B200_EXPORT int lt_b200_profile_enable(lt_b200_context* context, int on);
LT_B200_EXPORT int lt_b200_profile_reset(lt_b200_context* context);
LT_B200_EXPORT int lt_b200_profile_read(lt_b200_context* context, uint32_t kernel, double* out_ms, uint64_t* out_launches, uint64_t* out_bytes);

/* ---- device memory helpers (so that C callers need no CUDA headers) */
LT_B200_EXPORT int lt_b200_device_alloc(lt_b200_context* context, uint64_t bytes, void** out_device_ptr);
LT_B200_EXPORT int lt_b200_device_free(lt_b200_context* context, void* device_ptr);
LT_B200_EXPORT int lt_b200_host_alloc_pinned(lt_b200_context* context, uint64_t bytes, void** out_host_ptr);
LT_B200_EXPORT int lt_b200_host_free_pinned(lt_b200_context* context, void* host_ptr);
LT_B200_EXPORT int lt_b200_copy_to_device(lt_b200_context* context, void* device_dst, const void* host_src, uint64_t bytes);
LT_B200_EXPORT int lt_b200_copy_to_host(lt_b200_context* context, void* host_dst, const void* device_src, uint64_t bytes);
/* copy_to_device queued on the context's stream without waiting: host_src must stay valid until the next synchronising call on the context */
LT_B200_EXPORT int lt_b200_copy_to_device_async(lt_b200_context* context, void* device_dst, const void* host_src, uint64_t bytes);

/* Deterministic synthetic asset bytes written straight into HBM (include/lt_synth.h).  Bench/test input only. */
struct lt_b200_synth_spec
{
    uint64_t seed;
    uint32_t shared_permille;
    uint32_t pool_segments;
    uint32_t class_mode;
    uint32_t reserved;
};
LT_B200_EXPORT int lt_b200_synth_fill(lt_b200_context* context, void* device_dst, uint64_t bytes, const struct lt_b200_synth_spec* spec,
                                      uint64_t asset_id, uint64_t asset_offset);

/* ---- layer 1: device primitives */

/* One range = one chunker instance = one "part" of the reference (src/longtail.c:2396-2457): bytes
 * [arena_offset, arena_offset + size) of the device arena.  arena_offset must be a multiple of 16. */
struct lt_b200_range
{
    uint64_t arena_offset;
    uint32_t size; /* < 2^31 */
    uint32_t tag;  /* carried to every chunk of the range (asset compression tag) */
};

/* Result of chunk + hash over a list of ranges.  Arrays live in pinned host memory owned by the context and
 * stay valid until the next lt_b200_chunk_ranges call on it; the device copies stay resident for
 * lt_b200_finalize_index / lt_b200_write_blocks. */
struct lt_b200_chunk_table
{
    uint32_t range_count;
    uint32_t chunk_count;               /* total over all ranges */
    const uint32_t* range_chunk_counts; /* [range_count] */
    const uint64_t* chunk_hashes;       /* [chunk_count] in range order, then stream order */
    const uint32_t* chunk_sizes;        /* [chunk_count] */
    const uint32_t* chunk_tags;         /* [chunk_count] */
    const uint64_t* chunk_offsets;      /* [chunk_count] arena offset of each chunk */
};

/* Content-defined chunking (HPCDC) of every range followed by the hash of every chunk.
 * min/avg/max as computed by the caller from target_chunk_size (src/longtail.c:1985-1987); returns EINVAL for
 * parameter combinations the reference asserts on (longtail_hpcdcchunker.c:146-150).
 * want_host = 0 skips the device->host copy of the table (out_table then only carries the counts). */
LT_B200_EXPORT int lt_b200_chunk_ranges(lt_b200_context* context, const uint8_t* device_arena, uint64_t arena_size,
                                        const struct lt_b200_range* ranges, uint32_t range_count,
                                        uint32_t min_chunk_size, uint32_t avg_chunk_size, uint32_t max_chunk_size,
                                        uint32_t hash_type, int want_host, struct lt_b200_chunk_table* out_table);

/* HashAPI.HashBuffer over `count` segments [offsets[i], offsets[i]+sizes[i]) of one device buffer.  offsets/sizes/out are
 * HOST arrays. */
LT_B200_EXPORT int lt_b200_hash_segments(lt_b200_context* context, uint32_t hash_type, const uint8_t* device_base, uint64_t base_size,
                                         const uint64_t* offsets, const uint32_t* sizes, uint32_t count, uint64_t* out_hashes);

/* ---- layer 2: batched verbs */

/* Input of the index build: what Longtail_FileInfos + optional_asset_tags carry (src/longtail.h:1684-1692) plus, per asset,
 * how many chunks its ranges produced. */
struct lt_b200_assets
{
    uint32_t asset_count;
    uint32_t path_data_size;            /* bytes in path_data, NUL terminators included */
    const uint64_t* sizes;              /* [asset_count] */
    const uint32_t* path_start_offsets; /* [asset_count] */
    const uint16_t* permissions;        /* [asset_count] */
    const char* path_data;              /* relative paths, NUL terminated, directories end with '/' */
};

/* Builds the serialised VersionIndex (src/longtail.c:2566-2584 layout, Longtail_WriteVersionIndexToBuffer :3415-3439) from a
 * chunk table: per-asset content hashes (:2518-2537), path hashes (:1281-1297), first-occurrence dedup (:2952-2970) and the
 * zero-parse layout (:2709-2806), all on the device.
 *
 * chunk arrays are HOST pointers covering all assets in asset order (chunk_count entries; asset a owns
 * asset_chunk_counts[a] consecutive entries) — after a multi-GPU allgather they are the merged table.  Pass NULL for
 * chunk_hashes/sizes/tags to use the table still resident on the device from the last lt_b200_chunk_ranges call.
 * *out_buffer points at pinned host memory owned by the context, valid until the next index call on it (the drop-in
 * verb copies it into Longtail_Alloc memory); do not free it. */
LT_B200_EXPORT int lt_b200_build_version_index(lt_b200_context* context, const struct lt_b200_assets* assets,
                                               const uint32_t* asset_chunk_counts, uint32_t chunk_count,
                                               const uint64_t* chunk_hashes, const uint32_t* chunk_sizes, const uint32_t* chunk_tags,
                                               uint32_t hash_type, uint32_t target_chunk_size,
                                               const void** out_buffer, uint64_t* out_size);

/* Device addresses of the chunk table left resident by the last lt_b200_chunk_ranges call (u64 hashes, u32 sizes, u32 tags),
 * e.g. to hand them to an NCCL allgather without a host round trip. */
LT_B200_EXPORT int lt_b200_resident_table(lt_b200_context* context, void** out_device_hashes, void** out_device_sizes,
                                          void** out_device_tags, uint32_t* out_chunk_count);

/* lt_b200_build_version_index with the (merged) chunk table given as DEVICE arrays. */
LT_B200_EXPORT int lt_b200_build_version_index_device(lt_b200_context* context, const struct lt_b200_assets* assets,
                                                      const uint32_t* asset_chunk_counts, uint32_t chunk_count,
                                                      const void* device_chunk_hashes, const void* device_chunk_sizes,
                                                      const void* device_chunk_tags, uint32_t hash_type, uint32_t target_chunk_size,
                                                      const void** out_buffer, uint64_t* out_size);

/* Whole CreateVersionIndex over assets resident in one device arena: splits every asset into parts of
 * target_chunk_size*1024 bytes (src/longtail.c:2396-2437), chunks + hashes them and builds the index.
 * asset_arena_offsets[a] (multiples of 16) locate asset a's bytes in the arena; asset_tags may be NULL. */
LT_B200_EXPORT int lt_b200_index_device_assets(lt_b200_context* context, const uint8_t* device_arena, uint64_t arena_size,
                                               const struct lt_b200_assets* assets, const uint64_t* asset_arena_offsets,
                                               const uint32_t* asset_tags, uint32_t hash_type, uint32_t target_chunk_size,
                                               const void** out_buffer, uint64_t* out_size);

/* Same, with asset bytes in HOST memory (asset_data[a] may be pageable or pinned): host->device copies are part of the
 * call and are pipelined against the kernels in batches of whole parts. */
LT_B200_EXPORT int lt_b200_index_host_assets(lt_b200_context* context, const struct lt_b200_assets* assets,
                                             const uint8_t* const* asset_data, const uint32_t* asset_tags,
                                             uint32_t hash_type, uint32_t target_chunk_size,
                                             const void** out_buffer, uint64_t* out_size);

/* ---- the WriteContent half: block packing, block hashes, payload gather, compression, StoredBlock serialisation
 *
 * Replaces, for chunks whose bytes are resident in a device arena: Longtail_CreateStoreIndex's packing
 * (src/longtail.c:6745-6880), Longtail_CreateBlockIndex (:3712-3770), WriteContentBlockJob's payload gather (:4559-4758) and
 * compressblockstore's CompressBlock (lib/compressblockstore/longtail_compressblockstore.c:67-141) with the LZ4 backend
 * (lib/lz4/longtail_lz4.c:52-77) or the ZStd level 3 backend (lib/zstd/longtail_zstd.c:107-140).  `chunk_*` are HOST arrays listing the chunks to store, in store order (for a fresh store:
 * the unique chunks of the VersionIndex in order — DiffHashes keeps that order, src/longtail.c:6718-6740).
 * Every finished block is handed to `sink` in store order as the exact byte image Longtail_WriteStoredBlockToBuffer
 * (src/longtail.c:4111-4150) would produce; the memory is only valid during the call.  Tags: 0 = stored raw, 'lz42' = LZ4,
 * 'ztd1' / 'ztd2' = ZStd level 3; anything else returns ENOTSUP. */
struct lt_b200_stored_block_view
{
    uint64_t block_hash;
    const void* data;          /* serialised stored block: block index + payload */
    uint64_t size;
    uint32_t chunk_count;
    uint32_t tag;
    uint32_t raw_payload_size; /* uncompressed payload bytes (what compressblockstore's stats count) */
    uint32_t first_chunk;      /* index of the block's first chunk in the caller's arrays */
};
typedef int (*lt_b200_block_sink)(void* user, const struct lt_b200_stored_block_view* block);
LT_B200_EXPORT int lt_b200_write_blocks_device(lt_b200_context* context, const uint8_t* device_arena, uint64_t arena_size,
                                               uint32_t chunk_count, const uint64_t* chunk_hashes, const uint32_t* chunk_sizes,
                                               const uint32_t* chunk_tags, const uint64_t* chunk_arena_offsets, uint32_t hash_type,
                                               uint32_t max_block_size, uint32_t max_chunks_per_block, lt_b200_block_sink sink, void* user);

/* lt_b200_write_blocks_device with flags.  LT_B200_WRITE_DEVICE_SINK: the serialised images stay in HBM — `data` of every view is a
 * DEVICE address (for a consumer that reads device memory: a GPUDirect NIC / storage path, a peer GPU); without it every image is
 * copied once into pinned host staging, batch k's copies overlapping batch k + 1's kernels when device memory allows two batches. */
#define LT_B200_WRITE_DEVICE_SINK 1u
LT_B200_EXPORT int lt_b200_write_blocks_device_ex(lt_b200_context* context, const uint8_t* device_arena, uint64_t arena_size,
                                                  uint32_t chunk_count, const uint64_t* chunk_hashes, const uint32_t* chunk_sizes,
                                                  const uint32_t* chunk_tags, const uint64_t* chunk_arena_offsets, uint32_t hash_type,
                                                  uint32_t max_block_size, uint32_t max_chunks_per_block, uint32_t flags,
                                                  lt_b200_block_sink sink, void* user);

/* an lt_b200_block_sink that only counts (benchmarks, dry runs): user = uint64_t[4] {blocks, stored bytes, raw payload bytes, xor of the
 * block hashes}; it never dereferences `data`, so it also serves LT_B200_WRITE_DEVICE_SINK */
LT_B200_EXPORT int lt_b200_counting_sink(void* user, const struct lt_b200_stored_block_view* block);

/* lt_b200_write_blocks_device for blocks whose composition is GIVEN (a Longtail_StoreIndex: block b holds the next block_chunk_counts[b]
 * chunks of the chunk arrays; its tag is the tag of its first chunk) instead of packed greedily — what Longtail_WriteContent does with the
 * store index it is handed (src/longtail.c:4760-4912). */
LT_B200_EXPORT int lt_b200_write_given_blocks_device(lt_b200_context* context, const uint8_t* device_arena, uint64_t arena_size,
                                                     uint32_t chunk_count, const uint64_t* chunk_hashes, const uint32_t* chunk_sizes,
                                                     const uint32_t* chunk_tags, const uint64_t* chunk_arena_offsets, uint32_t hash_type,
                                                     uint32_t block_count, const uint32_t* block_chunk_counts, lt_b200_block_sink sink, void* user);

/* DiffHashes of Longtail_CreateMissingContent (src/longtail.c:6620-6743, :6882-6998) on the device: out_missing[i] = 1 when
 * chunk_hashes[i] (the version's unique chunks, HOST array) is absent from existing_hashes (the chunk hashes of the store index, HOST
 * array).  The chunks to write are the flagged ones in their given order; pass them to lt_b200_write_blocks_device. */
LT_B200_EXPORT int lt_b200_missing_chunks(lt_b200_context* context, uint32_t chunk_count, const uint64_t* chunk_hashes,
                                          uint32_t existing_count, const uint64_t* existing_hashes, uint8_t* out_missing);

/* Host helper (no GPU needed): Longtail_CreateStoreIndex's greedy block packing (src/longtail.c:6796-6860) over chunks in store
 * order — a block closes on a tag change, at max_chunks_per_block chunks, or when the next chunk would exceed
 * max_block_size + max_block_size/10.  out_block_first / out_block_count need room for chunk_count entries.  Used by planners
 * that shard WriteContent by block across GPUs (longtail_b200/distributed.py). */
LT_B200_EXPORT int lt_b200_pack_blocks(uint32_t chunk_count, const uint32_t* chunk_sizes, const uint32_t* chunk_tags, uint32_t max_block_size,
                                       uint32_t max_chunks_per_block, uint32_t* out_block_first, uint32_t* out_block_count, uint32_t* out_blocks);

/* CompressionAPI.Compress / Decompress (src/longtail.h:266-272) for 'lz42' over `count` independent HOST buffers in one
 * launch: LZ4_compress_fast(acceleration 1) / LZ4_decompress_safe semantics (lib/lz4/longtail_lz4.c:52-101), output bytes
 * identical to the reference codec.  dst_capacity[i] must be at least lt_b200_lz4_bound(src_size[i]) for compression
 * (the reference asserts the same through GetMaxCompressedSize).  Errors: ENOMEM capacity too small, EBADF malformed input. */
LT_B200_EXPORT uint64_t lt_b200_lz4_bound(uint64_t size); /* LZ4_COMPRESSBOUND, lib/lz4/ext/lz4.h:215 */
LT_B200_EXPORT int lt_b200_lz4_compress_host(lt_b200_context* context, uint32_t count, const void* const* src, const uint32_t* src_size,
                                             void* const* dst, const uint64_t* dst_capacity, uint64_t* out_size);
LT_B200_EXPORT int lt_b200_lz4_decompress_host(lt_b200_context* context, uint32_t count, const void* const* src, const uint32_t* src_size,
                                               void* const* dst, const uint64_t* dst_capacity, uint64_t* out_size);

/* CompressionAPI.Compress (src/longtail.h:266-272) for 'ztd1' / 'ztd2' over `count` independent HOST buffers in one launch:
 * ZSTD_compressCCtx(level 3) semantics (lib/zstd/longtail_zstd.c:107-140), one frame per buffer, output bytes identical to the
 * reference codec (vendored zstd 1.5.6).  dst_capacity[i] must be at least lt_b200_zstd_bound(src_size[i]) (EINVAL otherwise,
 * the reference's answer to a ZStd error).  The other ZStd quality ids ('ztd3' / 'ztd4' / 'ztd5': levels 22 / 8 / 22) have no
 * device encoder: ENOTSUP. */
/* diagnostic: GPU cycles the ZStd workers have spent in {matcher, literal stage, sequence stage, copy-out}, summed over workers */
LT_B200_EXPORT int lt_b200_zstd_phase_cycles(lt_b200_context* context, uint64_t out_cycles[4]);
LT_B200_EXPORT uint64_t lt_b200_zstd_bound(uint64_t size); /* ZSTD_COMPRESSBOUND, lib/zstd/ext/zstd.h:232 */
LT_B200_EXPORT int lt_b200_zstd_compress_host(lt_b200_context* context, uint32_t compression_type, uint32_t count, const void* const* src,
                                              const uint32_t* src_size, void* const* dst, const uint64_t* dst_capacity, uint64_t* out_size);

/* CompressionAPI.Decompress (src/longtail.h:266-272) for every ZStd type id ('ztd1' .. 'ztd5'; lib/zstd/longtail_zstd.c:143-176 ->
 * ZSTD_decompressDCtx) over `count` independent HOST buffers in one launch, one warp per frame.  A frame written by any level
 * decodes (raw / RLE / compressed blocks, raw / RLE / Huffman / treeless literals, predefined / RLE / FSE / repeat sequence
 * tables); dictionaries are not supported (longtail uses none).  out_size[i] = the decoded size; EBADF on a malformed frame or
 * a frame whose content does not fit dst_capacity[i]. */
LT_B200_EXPORT int lt_b200_zstd_decompress_host(lt_b200_context* context, uint32_t count, const void* const* src, const uint32_t* src_size,
                                                void* const* dst, const uint64_t* dst_capacity, uint64_t* out_size);

/* ---- the step after PutStoredBlock: an on-disk block sink in the layout of the reference's fsblockstore
 * (lib/fsblockstore/longtail_fsblockstore.c): <store>/chunks/<4 hex>/0x<16 hex>.lrb per block (:66-129) and <store>/store.lsi, the
 * serialised Longtail_StoreIndex (src/longtail.c:8913-8977).  An unmodified longtail opens the directory with
 * Longtail_CreateFSBlockStoreAPI.  Host code, no GPU needed.
 *   open     creates the directory; `writer_threads` file writers work behind the sink (0 = write inside the sink call)
 *   sink     an lt_b200_block_sink (pass the store as `user` to lt_b200_write_blocks_device): existing block files are not rewritten
 *            (SafeWriteStoredBlock :243-320), new ones are written under a temporary name and renamed
 *   flush    waits for the writers; store.lsi := blocks added since the last flush, in sink order, merged in front of the index on disk
 *            (UpdateStoreIndex :330-352, Longtail_MergeStoreIndex src/longtail.c:9155-9290) under flock(store.lsi.sync)
 *   existing_chunks  the chunk hashes store.lsi lists — the `existing_hashes` of lt_b200_missing_chunks (out_hashes may be NULL to size)
 *   close    flush + release */
#define LT_B200_STORE_INDEX_VERSION 0x01000000u /* LONGTAIL_STORE_INDEX_VERSION_1_0_0, src/longtail.c:16-23 */
typedef struct lt_b200_fs_store lt_b200_fs_store;
LT_B200_EXPORT int lt_b200_fs_store_open(const char* store_path, uint32_t writer_threads, lt_b200_fs_store** out_store);
LT_B200_EXPORT int lt_b200_fs_store_sink(void* store, const struct lt_b200_stored_block_view* block);
LT_B200_EXPORT int lt_b200_fs_store_flush(lt_b200_fs_store* store);
LT_B200_EXPORT int lt_b200_fs_store_existing_chunks(lt_b200_fs_store* store, uint64_t* out_hashes, uint32_t capacity, uint32_t* out_count);
LT_B200_EXPORT int lt_b200_fs_store_stats(lt_b200_fs_store* store, uint64_t* out_blocks_written, uint64_t* out_bytes_written,
                                          uint64_t* out_blocks_skipped);
LT_B200_EXPORT int lt_b200_fs_store_close(lt_b200_fs_store* store);

/* Arena offsets of the unique chunks (first occurrences, VersionIndex order) found by the last lt_b200_index_device_assets
 * call on this context — the `chunk_arena_offsets` of a fresh-store lt_b200_write_blocks_device. */
LT_B200_EXPORT int lt_b200_unique_chunk_offsets(lt_b200_context* context, uint64_t* out_offsets, uint32_t count);

/* Same, with the asset bytes pulled through a callback: the library hands out batches of read jobs whose destinations are
 * pinned staging buffers it owns; the callee fills them (in parallel if it likes — the drop-in verb fans them out over the
 * caller's Longtail_JobAPI) and returns 0 or an errno, which aborts the call.  While batch k+1 is being read on the host,
 * batch k is already on its way to the device. */
struct lt_b200_read_job
{
    uint32_t asset_index;
    uint32_t size;
    uint64_t offset; /* byte offset inside the asset */
    void* dst;
};
typedef int (*lt_b200_read_batch_func)(void* user, const struct lt_b200_read_job* jobs, uint32_t job_count);
LT_B200_EXPORT int lt_b200_index_stream_assets(lt_b200_context* context, const struct lt_b200_assets* assets, const uint32_t* asset_tags,
                                               uint32_t hash_type, uint32_t target_chunk_size, lt_b200_read_batch_func read_batch,
                                               void* user, const void** out_buffer, uint64_t* out_size);

/* ---- the step before the hot path: Longtail_GetFilesRecursively2 (src/longtail.c:1656-1893) over the
 * file system and a threaded reader behind lt_b200_index_stream_assets.  Host code.
 *   scan_directory   lists `root_path` recursively with `threads` workers; entries, their order (strcmp over the relative names before
 *                    directories get their trailing '/'), sizes and permissions (st_mode & 0x1FF) are the reference's.  Entries that are
 *                    neither regular files nor directories: ENOTSUP.
 *   file_list_assets the result as the lt_b200_assets every index verb takes (valid until file_list_free)
 *   index_file_list  CreateVersionIndex over the scanned tree: `reader_threads` threads pread the parts into the pinned staging of the
 *                    streaming verb while the previous batch is on the device; asset_tags may be NULL */
typedef struct lt_b200_file_list lt_b200_file_list;
LT_B200_EXPORT int lt_b200_scan_directory(const char* root_path, uint32_t threads, lt_b200_file_list** out_list);
LT_B200_EXPORT const struct lt_b200_assets* lt_b200_file_list_assets(const lt_b200_file_list* list);
LT_B200_EXPORT void lt_b200_file_list_free(lt_b200_file_list* list);
LT_B200_EXPORT int lt_b200_index_file_list(lt_b200_context* context, const lt_b200_file_list* list, const uint32_t* asset_tags,
                                           uint32_t hash_type, uint32_t target_chunk_size, uint32_t reader_threads,
                                           const void** out_buffer, uint64_t* out_size);

/* cmd/main.c:UpSync (:972-1153) for a scanned tree that fits one GPU: the tree is read once into a device arena (reader threads ->
 * pinned staging -> HBM), then CreateVersionIndex, CreateMissingContent against the chunks `store` already lists and WriteContent all run
 * on the resident bytes; stored blocks leave through lt_b200_fs_store_sink and the store is flushed.  *out_version_index is the serialised
 * VersionIndex (pinned memory owned by the context, valid until its next index call) — what UpSync writes to the .lvi file.
 * ENOMEM when the tree does not fit the device. */
LT_B200_EXPORT int lt_b200_upsync_file_list(lt_b200_context* context, const lt_b200_file_list* list, const uint32_t* asset_tags,
                                            uint32_t hash_type, uint32_t target_chunk_size, uint32_t max_block_size,
                                            uint32_t max_chunks_per_block, uint32_t reader_threads, lt_b200_fs_store* store,
                                            const void** out_version_index, uint64_t* out_size, uint32_t* out_blocks_written);

/* cmd/main.c:UpSync (:972-1153) for assets in HOST memory (pinned for full PCIe speed) that fit one GPU: one host -> device copy per asset
 * into a device arena, then CreateVersionIndex, CreateMissingContent against `existing_hashes` (NULL / 0: a fresh store) and WriteContent
 * on the resident bytes; the stored blocks leave through `sink` (flags as lt_b200_write_blocks_device_ex).  *out_version_index: pinned
 * memory owned by the context, valid until its next index call.  ENOMEM when the assets do not fit the device. */
LT_B200_EXPORT int lt_b200_upsync_host_assets(lt_b200_context* context, const struct lt_b200_assets* assets, const uint8_t* const* asset_data,
                                              const uint32_t* asset_tags, uint32_t hash_type, uint32_t target_chunk_size, uint32_t max_block_size,
                                              uint32_t max_chunks_per_block, uint32_t existing_count, const uint64_t* existing_hashes,
                                              uint32_t flags, lt_b200_block_sink sink, void* user, const void** out_version_index,
                                              uint64_t* out_size, uint32_t* out_chunks_written);

/* The same upsync in ONE streaming pass, for host-resident assets of any total size (the version never has to fit the device): batches of
 * whole parts (`batch_bytes` each, 0 = 16 GiB, never less than one part of target_chunk_size * 1024 bytes) travel host -> device into
 * two arenas; while batch k + 1 is on the bus batch k is chunked + hashed, its chunks are matched against the hashes seen so far
 * (first occurrence wins, src/longtail.c:2952-2970; the set starts with `existing_hashes`, Longtail_CreateMissingContent :7257-7340),
 * and every stored block the greedy packing (:6796-6860) can no longer change is compressed and handed to `sink` — so host -> device
 * and device -> host copies overlap.  VersionIndex, stored blocks and their order are identical to lt_b200_upsync_host_assets'. */
LT_B200_EXPORT int lt_b200_upsync_stream_host_assets(lt_b200_context* context, const struct lt_b200_assets* assets,
                                                     const uint8_t* const* asset_data, const uint32_t* asset_tags, uint32_t hash_type,
                                                     uint32_t target_chunk_size, uint32_t max_block_size, uint32_t max_chunks_per_block,
                                                     uint32_t existing_count, const uint64_t* existing_hashes, uint32_t flags,
                                                     uint64_t batch_bytes, lt_b200_block_sink sink, void* user,
                                                     const void** out_version_index, uint64_t* out_size, uint32_t* out_chunks_written);

/* ---- multi-GPU: one process per GPU of one node, NCCL over NVLink / NVSwitch.  The reference has no distributed
 * runtime (its only parallelism is the bikeshed thread pool behind Longtail_JobAPI); what is sharded here is its own job unit, one
 * (asset, part) pair with parts of target_chunk_size * 1024 bytes (src/longtail.c:2396-2457): parts never share chunker state, so ranks
 * chunk + hash disjoint slices and exchange only the chunk tables.
 *
 *   comm_unique_id    rank 0 makes the NCCL id and hands it to the other processes by any means (a file, an environment variable, MPI, a
 *                     torch.distributed store: the library does not care)
 *   comm_create       joins `world` processes (collective); one communicator per context
 *   plan_shards       host helper: out_first_job[r] .. out_first_job[r + 1] are rank r's jobs — contiguous slices of the job list in asset /
 *                     part order (empty parts left out), balanced by bytes; shard_jobs describes them so that the caller can put each
 *                     job's bytes into its rank's device arena
 *   index_sharded     Longtail_CreateVersionIndex over all ranks (collective): chunk + hash of this rank's jobs (job_arena_offsets[j] = where
 *                     the bytes of its j-th job sit in device_arena), variable-size all-gather of the chunk tables, first-occurrence dedup
 *                     split by hash with one all-reduce, the same VersionIndex laid out on every rank; the host copy only where want_host
 *   write_blocks_sharded  Longtail_CreateMissingContent (fresh store) + Longtail_WriteContent over all ranks (collective, after index_sharded):
 *                     every rank derives the same block packing, takes a contiguous run of blocks balanced by bytes, receives the chunks whose
 *                     first occurrence lives on another rank point to point, and hands its finished blocks to its own sink in store order.
 *                     The union over the ranks, in rank order, is the block list of the reference's single-process upsync. */
#define LT_B200_COMM_ID_BYTES 128
typedef struct lt_b200_comm lt_b200_comm;
struct lt_b200_shard_job
{
    uint32_t asset_index;
    uint32_t size;
    uint64_t offset; /* byte offset inside the asset */
};
LT_B200_EXPORT int lt_b200_comm_unique_id(uint8_t out_id[LT_B200_COMM_ID_BYTES]);
LT_B200_EXPORT int lt_b200_comm_create(lt_b200_context* context, const uint8_t id[LT_B200_COMM_ID_BYTES], uint32_t rank, uint32_t world,
                                       lt_b200_comm** out_comm);
LT_B200_EXPORT void lt_b200_comm_destroy(lt_b200_comm* comm);
LT_B200_EXPORT int lt_b200_plan_shards(const struct lt_b200_assets* assets, uint32_t target_chunk_size, uint32_t world,
                                       uint32_t* out_first_job /* [world + 1] */, uint32_t* out_job_count);
LT_B200_EXPORT int lt_b200_shard_jobs(const struct lt_b200_assets* assets, uint32_t target_chunk_size, uint32_t first_job, uint32_t job_count,
                                      struct lt_b200_shard_job* out_jobs);
LT_B200_EXPORT int lt_b200_index_sharded(lt_b200_context* context, lt_b200_comm* comm, const uint8_t* device_arena, uint64_t arena_size,
                                         const struct lt_b200_assets* assets, const uint32_t* asset_tags, const uint64_t* job_arena_offsets,
                                         uint32_t hash_type, uint32_t target_chunk_size, int want_host, const void** out_buffer, uint64_t* out_size);
LT_B200_EXPORT int lt_b200_write_blocks_sharded(lt_b200_context* context, lt_b200_comm* comm, uint32_t max_block_size, uint32_t max_chunks_per_block,
                                                uint32_t flags, lt_b200_block_sink sink, void* user, uint32_t* out_my_blocks,
                                                uint32_t* out_total_blocks);

#ifdef __cplusplus
}
#endif
#endif /* LONGTAIL_B200_H */
