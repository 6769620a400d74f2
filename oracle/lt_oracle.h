/* oracle/lt_oracle.h — CPU restatement of longtail's chunk -> hash -> compress indexing path.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg as the checker.  The product (longtail_b200/) never links or loads it.
 *
 * Parity status: PINNED.  Checked in tests/test_oracle.py against the reference's golden
 * vectors (test/test.cpp:3422-3445 chunker vector, :460/:472 hash KATs, :2185-2192 LZ4 size pin)
 * and, where oracle/_ref is built, byte-for-byte against the unmodified reference
 * (oracle/ref_shim.c) on seeded inputs.
 */
#ifndef LT_ORACLE_H
#define LT_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTO_HASH_BLAKE3 0x626c6b33u /* 'blk3' lib/blake3/longtail_blake3.c:6 */
#define LTO_HASH_BLAKE2 0x626c6b32u /* 'blk2' lib/blake2/longtail_blake2.c:9 */
#define LTO_HASH_MEOW 0x6d656f77u   /* 'meow' lib/meowhash/longtail_meowhash.c:7 */
#define LTO_COMPRESSION_LZ4 0x6c7a3432u /* 'lz42' lib/lz4/longtail_lz4.c:10 */

void lto_free(void* p);

/* lib/hpcdcchunker/longtail_hpcdcchunker.c:126-129 */
uint32_t lto_hpcdc_discriminator(uint32_t avg);

/* Window hash of the 48 bytes ending at p (exclusive): the stateless form of
 * longtail_hpcdcchunker.c:273-306. */
uint32_t lto_hpcdc_window_hash(const uint8_t* end);

/* Chunk one part (one chunker instance fed to exhaustion):
 * longtail_hpcdcchunker.c:225-310 as driven by src/longtail.c:2231-2296. */
int lto_hpcdc_chunk(const uint8_t* data, uint64_t size, uint32_t min, uint32_t avg, uint32_t max,
                    uint32_t* out_lens, uint64_t cap, uint64_t* out_count);

/* lib/blake3/longtail_blake3.c:81-102: unkeyed BLAKE3, digest bytes 0..7 as LE u64 */
uint64_t lto_blake3_64(const void* data, uint64_t len);
/* lib/blake2/longtail_blake2.c:95-112: blake2s with outlen 8 */
uint64_t lto_blake2s_64(const void* data, uint64_t len);
/* lib/meowhash/longtail_meowhash.c:43-50: Meow 0.5/calico with the default seed, low 64 bits of the 128-bit hash */
uint64_t lto_meow_64(const void* data, uint64_t len);
/* HashAPI.HashBuffer by type id; returns 0 or EINVAL */
int lto_hash_buffer(uint32_t hash_type, const void* data, uint64_t len, uint64_t* out_hash);
int lto_hash_segments(uint32_t hash_type, const uint8_t* base, uint64_t count,
                      const uint64_t* offsets, const uint32_t* lens, uint64_t* out_hashes);

/* lib/lz4/ext/lz4.h:215 and lib/lz4/ext/lz4.c:930-1338 (LZ4_compress_fast, acceleration 1,
 * notLimited / noDict); lib/lz4/longtail_lz4.c:52-77 */
uint64_t lto_lz4_bound(uint64_t size);
int lto_lz4_compress(const uint8_t* src, uint64_t size, uint8_t* dst, uint64_t cap, uint64_t* out_size);
int lto_lz4_decompress(const uint8_t* src, uint64_t size, uint8_t* dst, uint64_t cap, uint64_t* out_size);

/* lib/zstd/longtail_zstd.c:107-140 with 'ztd1' / 'ztd2' (both ZStd level 3): one frame per call, content size in the header, no
 * checksum; restated in lt_zstd.c from lib/zstd/ext (zstd 1.5.6).  cap must be >= lto_zstd_bound(size) (zstd.h:232). */
#define LTO_COMPRESSION_ZSTD_DEFAULT 0x7a746432u /* 'ztd2' lib/zstd/longtail_zstd.c:20 */
#define LTO_COMPRESSION_ZSTD_MIN 0x7a746431u     /* 'ztd1' -> level 0 == default == 3, lib/zstd/longtail_zstd.c:19,47 */
uint64_t lto_zstd_bound(uint64_t size);
int lto_zstd_compress(const uint8_t* src, uint64_t size, uint8_t* dst, uint64_t cap, uint64_t* out_size);

/* src/longtail.c:2808-3017 (Longtail_CreateVersionIndex) + :3415-3439 (serialise).
 * Assets are in-memory; paths are relative, directories end with '/'.  *out_buf is malloc'd. */
int lto_create_version_index(uint32_t count, const char** paths, const uint8_t** datas, const uint64_t* sizes,
                             const uint16_t* perms, const uint32_t* tags, uint32_t hash_type,
                             uint32_t target_chunk_size, void** out_buf, uint64_t* out_size);

/* Fresh-store upsync: CreateVersionIndex -> CreateMissingContent(empty store) -> WriteContent
 * through compressblockstore (src/longtail.c:6882-6998, :6745-6880, :4559-4758;
 * lib/compressblockstore/longtail_compressblockstore.c:67-141).  Output format identical to
 * ref_upsync in oracle/ref_shim.c.  Tags: 0 = store raw, LTO_COMPRESSION_LZ4 = LZ4. */
int lto_upsync(uint32_t count, const char** paths, const uint8_t** datas, const uint64_t* sizes,
               const uint16_t* perms, const uint32_t* tags, uint32_t hash_type,
               uint32_t target_chunk_size, uint32_t max_block_size, uint32_t max_chunks_per_block,
               void** out_buf, uint64_t* out_size);

#ifdef __cplusplus
}
#endif
#endif
