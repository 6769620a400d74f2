/* longtail_b200_api.h — drop-in Longtail_*API objects and batched verbs backed by the B200 kernels.
 *
 * Everything here speaks longtail's own plugin ABI (include/longtail_abi.h == reference src/longtail.h), so the objects
 * can be handed to the unmodified reference core (Longtail_CreateVersionIndex, Longtail_WriteContent, the registries,
 * cmd/main.c:UpSync) exactly where the reference's own backends go:
 *
 *   Longtail_CreateB200ChunkerAPI()        replaces Longtail_CreateHPCDCChunkerAPI()   lib/hpcdcchunker/longtail_hpcdcchunker.c:563
 *   Longtail_CreateB200Blake3HashAPI()     replaces Longtail_CreateBlake3HashAPI()     lib/blake3/longtail_blake3.c:104
 *   Longtail_B200_CreateVersionIndex()     replaces Longtail_CreateVersionIndex()      src/longtail.c:2808   (same parameter list)
 *
 * Error behaviour follows the reference: errno-style ints, 0 = success, ESPIPE from NextChunk at end of stream
 * (lib/hpcdcchunker/longtail_hpcdcchunker.c:420-423), objects freed through m_API.Dispose (SAFE_DISPOSE_API).
 * Objects returned here are allocated with the host's Longtail_Alloc when liblongtail is present in the process
 * (resolved with dlsym at first use) and with malloc otherwise, so that Longtail_Free / SAFE_DISPOSE_API on the
 * caller's side does the right thing in both settings.
 */
#ifndef LONGTAIL_B200_API_H
#define LONGTAIL_B200_API_H

#include "longtail_abi.h"
#include "longtail_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Device used by the objects created below (default 0).  Call before the first Create*; returns EBUSY afterwards. */
LT_B200_EXPORT int Longtail_B200_SetDevice(int device_ordinal);

/* ChunkerAPI (src/longtail.h:586-594).  GetMinChunkSize reports 48.  The first NextChunk on a handle drains the feeder
 * (one part of at most target_chunk_size*1024 bytes, src/longtail.c:2396), runs the scan + selection + BLAKE3 kernels once
 * over it and serves the remaining NextChunk calls — and the HashBuffer calls of the B200 HashAPI on those very ranges —
 * from the result.  NextChunkFromBuffer is not on the CreateVersionIndex path (src/longtail.c:2453 forces the feeder
 * branch) and returns ENOTSUP. */
LT_B200_EXPORT struct Longtail_ChunkerAPI* Longtail_CreateB200ChunkerAPI(void);

/* HashAPI (src/longtail.h:209-217), identifier 'blk3'.  HashBuffer on a range handed out by a B200 chunker returns the
 * hash computed in the chunker's pass; any other buffer is hashed on the GPU on the spot (correct, but one round trip per
 * call — the batched verb below is the fast path). */
LT_B200_EXPORT struct Longtail_HashAPI* Longtail_CreateB200Blake3HashAPI(void);
/* identifier 'blk2': BLAKE2s with an 8-byte digest, replaces Longtail_CreateBlake2HashAPI() (lib/blake2/longtail_blake2.c:114) */
LT_B200_EXPORT struct Longtail_HashAPI* Longtail_CreateB200Blake2HashAPI(void);
/* identifier 'meow': Meow hash 0.5/calico, default seed, low 64 bits; replaces Longtail_CreateMeowHashAPI() (lib/meowhash/longtail_meowhash.c:73) */
LT_B200_EXPORT struct Longtail_HashAPI* Longtail_CreateB200MeowHashAPI(void);

/* Longtail_CreateVersionIndex with the reference's parameter list (src/longtail.h:1134-1147).  storage_api supplies the
 * bytes (ConcatPath / OpenReadFile / Read / CloseFile, exactly the calls DynamicChunking makes); job_api, when given, runs
 * the storage reads of one batch in parallel while the GPU works on the previous batch.  hash_api / chunker_api only select
 * the algorithms (GetIdentifier must be 'blk3', 'blk2' or 'meow'; the chunker must report a minimum of 48): the work is done by
 * lt_b200_index_host_assets' kernels, not by calling back into them.  enable_file_map is ignored, as in the reference.
 * *out_version_index is allocated with Longtail_Alloc (see above) and is released by the caller with Longtail_Free. */
LT_B200_EXPORT int Longtail_B200_CreateVersionIndex(
    struct Longtail_StorageAPI* storage_api,
    struct Longtail_HashAPI* hash_api,
    struct Longtail_ChunkerAPI* chunker_api,
    struct Longtail_JobAPI* job_api,
    struct Longtail_ProgressAPI* progress_api,
    struct Longtail_CancelAPI* optional_cancel_api,
    Longtail_CancelAPI_HCancelToken optional_cancel_token,
    const char* root_path,
    const struct Longtail_FileInfos* file_infos,
    const uint32_t* optional_asset_tags,
    uint32_t target_chunk_size,
    int enable_file_map,
    struct Longtail_VersionIndex** out_version_index);

/* CompressionAPI (src/longtail.h:266-272) for the 'lz42' type id: LZ4_compress_fast(acceleration 1) / LZ4_decompress_safe on the
 * GPU, byte-identical to lib/lz4/longtail_lz4.c.  Compress returns ENOMEM when max_compressed_size is below
 * GetMaxCompressedSize (the reference returns ENOMEM when LZ4 reports 0), Decompress returns EBADF on a malformed stream. */
LT_B200_EXPORT struct Longtail_CompressionAPI* Longtail_CreateB200LZ4CompressionAPI(void);
/* Longtail_CompressionRegistry_CreateForTypeFunc (lib/compressionregistry/longtail_compression_registry.h:9): hand it to
 * Longtail_CreateDefaultCompressionRegistry in place of Longtail_CompressionRegistry_CreateForLZ4 */
LT_B200_EXPORT struct Longtail_CompressionAPI* Longtail_CompressionRegistry_CreateForB200LZ4(uint32_t compression_type, uint32_t* out_settings);

/* CompressionAPI for the ZStd type ids (lib/zstd/longtail_zstd.c:31-42, :107-140): Compress with settings 'ztd1' / 'ztd2' is
 * ZSTD_compressCCtx(level 3) on the GPU, one frame per call, byte-identical to the reference (vendored zstd 1.5.6);
 * GetMaxCompressedSize is ZSTD_COMPRESSBOUND; Decompress decodes the frames of EVERY ZStd type id on the GPU (k_zstd_decode: frames of all
 * levels, lib/zstd/longtail_zstd.c:143-176), EINVAL on a malformed frame like the reference.  Compress with the other quality ids ('ztd3'
 * level 22, 'ztd4' level 8, 'ztd5') returns ENOTSUP: there is no device encoder for them and no CPU fallback. */
LT_B200_EXPORT struct Longtail_CompressionAPI* Longtail_CreateB200ZStdCompressionAPI(void);
/* replaces Longtail_CompressionRegistry_CreateForZstd (lib/zstd/longtail_zstd.c:31) in Longtail_CreateDefaultCompressionRegistry */
LT_B200_EXPORT struct Longtail_CompressionAPI* Longtail_CompressionRegistry_CreateForB200ZStd(uint32_t compression_type, uint32_t* out_settings);

/* BlockStoreAPI decorator, replaces Longtail_CreateCompressBlockStoreAPI (lib/compressblockstore/longtail_compressblockstore.c:622):
 * PutStoredBlock compresses on the GPU — concurrent calls from the JobAPI workers are gathered into one launch by a
 * background thread, the caller's block stays alive until OnComplete as in the reference (:151-176) — and forwards the
 * compressed block to the backing store; GetStoredBlock fetches from the backing store and decodes on the GPU;
 * tag 0 passes through untouched; PruneBlocks returns ENOTSUP (:483-493); stats count what the reference counts (:199-201,
 * :362-363).  PutStoredBlock handles 'lz42', 'ztd1' and 'ztd2' (ZStd level 3); GetStoredBlock decodes 'lz42' and every 'ztd?' id (frames of all
 * levels).  Compression types without a device kernel make PutStoredBlock / GetStoredBlock fail with ENOTSUP.
 * `compression_registry` is accepted for signature compatibility and not used. */
LT_B200_EXPORT struct Longtail_BlockStoreAPI* Longtail_CreateB200CompressBlockStoreAPI(
    struct Longtail_BlockStoreAPI* backing_block_store,
    struct Longtail_CompressionRegistryAPI* compression_registry);

/* Longtail_WriteContent (src/longtail.c:4760-4912) TOGETHER WITH the compress block store it normally writes into: same parameter list,
 * but `backing_block_store_api` is the store that would sit BELOW Longtail_CreateCompressBlockStoreAPI — it receives the finished
 * (compressed) stored blocks, byte-identical to what compressblockstore would have forwarded.  The version's assets are read once through
 * source_storage_api (fanned out over job_api) into a device arena; payload gather (WriteContentBlockJob :4559-4758), block hashes,
 * compression and serialisation run on the device for the blocks exactly as store_index lists them.  Errors as the reference: EINVAL for
 * missing arguments or a store-index chunk the version does not hold (:4826-4832), the first storage / PutStoredBlock error otherwise;
 * ENOTSUP for block tags without a device codec. */
LT_B200_EXPORT int Longtail_B200_WriteContent(
    struct Longtail_StorageAPI* source_storage_api, struct Longtail_BlockStoreAPI* backing_block_store_api, struct Longtail_JobAPI* job_api,
    struct Longtail_ProgressAPI* progress_api, struct Longtail_CancelAPI* optional_cancel_api, Longtail_CancelAPI_HCancelToken optional_cancel_token,
    struct Longtail_StoreIndex* store_index, struct Longtail_VersionIndex* version_index, const char* assets_folder);

/* Longtail_SetMonitor (src/longtail.h:860, src/longtail.c:762-776) for the B200 verbs.  Longtail_B200_WriteContent fires, per block of the
 * store index and in block order, BlockCompose when the block's batch is launched, BlockSave (with the size of the serialised block) when
 * it is handed to the backing store and BlockSaved with the store's answer — the events of WriteContentBlockJob (src/longtail.c:4586-4753) —
 * and AssetOpen / AssetClose around the one read of every asset that holds a needed chunk.  AssetRead, which the reference fires per chunk
 * copy, is not fired: the gather runs on the device.  NULL switches the events off; the table is copied. */
LT_B200_EXPORT void Longtail_B200_SetMonitor(const struct Longtail_Monitor* monitor);

#ifdef __cplusplus
}
#endif
#endif /* LONGTAIL_B200_API_H */
