/* longtail_abi.h — the subset of longtail's public callback-struct ABI (reference src/longtail.h) that the
 * chunk -> hash -> compress hot path crosses, declared so that liblongtail_b200.so can be built and used without
 * the reference tree.
 *
 * These are INTERFACE declarations: struct member order and function-pointer signatures are the binary contract
 * between longtail's core (which calls through the structs) and any backend; they are restated here field for
 * field and checked against the real header by tests/dropin/abi_table.c (compiled with -DLONGTAIL_B200_USE_LONGTAIL_H
 * wherever /root/reference is available).
 *
 * A consumer that already includes the real longtail.h defines LONGTAIL_B200_USE_LONGTAIL_H before including
 * longtail_b200_api.h and none of the declarations below are emitted.
 */
#ifndef LONGTAIL_ABI_H
#define LONGTAIL_ABI_H

#ifdef LONGTAIL_B200_USE_LONGTAIL_H
#include "longtail.h"
#else

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t TLongtail_Hash;

/* src/longtail.h:43-46 — every API struct starts with this */
struct Longtail_API;
typedef void (*Longtail_DisposeFunc)(struct Longtail_API* api);
struct Longtail_API
{
    Longtail_DisposeFunc Dispose;
};

/* ---- CancelAPI / ProgressAPI / JobAPI: consumed, never implemented here (src/longtail.h:62-112, :494-511, :513-560) */
struct Longtail_CancelAPI;
typedef struct Longtail_CancelAPI_CancelToken* Longtail_CancelAPI_HCancelToken;
struct Longtail_CancelAPI
{
    struct Longtail_API m_API;
    int (*CreateToken)(struct Longtail_CancelAPI* cancel_api, Longtail_CancelAPI_HCancelToken* out_token);
    int (*Cancel)(struct Longtail_CancelAPI* cancel_api, Longtail_CancelAPI_HCancelToken token);
    int (*IsCancelled)(struct Longtail_CancelAPI* cancel_api, Longtail_CancelAPI_HCancelToken token);
    int (*DisposeToken)(struct Longtail_CancelAPI* cancel_api, Longtail_CancelAPI_HCancelToken token);
};

struct Longtail_ProgressAPI
{
    struct Longtail_API m_API;
    void (*OnProgress)(struct Longtail_ProgressAPI* progress_api, uint32_t total_count, uint32_t done_count);
};

struct Longtail_JobAPI;
typedef void* Longtail_JobAPI_Jobs;
typedef void* Longtail_JobAPI_Group;
typedef int (*Longtail_JobAPI_JobFunc)(void* context, uint32_t job_id, int detected_error);
struct Longtail_JobAPI
{
    struct Longtail_API m_API;
    uint32_t (*GetWorkerCount)(struct Longtail_JobAPI* job_api);
    int (*ReserveJobs)(struct Longtail_JobAPI* job_api, uint32_t job_count, Longtail_JobAPI_Group* out_job_group);
    int (*CreateJobs)(struct Longtail_JobAPI* job_api, Longtail_JobAPI_Group job_group, struct Longtail_ProgressAPI* progress_api,
                      struct Longtail_CancelAPI* optional_cancel_api, Longtail_CancelAPI_HCancelToken optional_cancel_token,
                      uint32_t job_count, Longtail_JobAPI_JobFunc job_funcs[], void* job_contexts[], uint8_t job_channel,
                      Longtail_JobAPI_Jobs* out_jobs);
    int (*AddDependecies)(struct Longtail_JobAPI* job_api, uint32_t job_count, Longtail_JobAPI_Jobs jobs, uint32_t dependency_job_count,
                          Longtail_JobAPI_Jobs dependency_jobs);
    int (*ReadyJobs)(struct Longtail_JobAPI* job_api, uint32_t job_count, Longtail_JobAPI_Jobs jobs);
    int (*WaitForAllJobs)(struct Longtail_JobAPI* job_api, Longtail_JobAPI_Group job_group, struct Longtail_ProgressAPI* progress_api,
                          struct Longtail_CancelAPI* optional_cancel_api, Longtail_CancelAPI_HCancelToken optional_cancel_token);
    int (*ResumeJob)(struct Longtail_JobAPI* job_api, uint32_t job_id);
    int (*GetMaxBatchCount)(struct Longtail_JobAPI* job_api, uint32_t* out_max_job_batch_count, uint32_t* out_max_dependency_batch_count);
};

/* ---- HashAPI, src/longtail.h:197-217 */
struct Longtail_HashAPI;
typedef struct Longtail_HashAPI_Context* Longtail_HashAPI_HContext;
struct Longtail_HashAPI
{
    struct Longtail_API m_API;
    uint32_t (*GetIdentifier)(struct Longtail_HashAPI* hash_api);
    int (*BeginContext)(struct Longtail_HashAPI* hash_api, Longtail_HashAPI_HContext* out_context);
    void (*Hash)(struct Longtail_HashAPI* hash_api, Longtail_HashAPI_HContext context, uint32_t length, const void* data);
    uint64_t (*EndContext)(struct Longtail_HashAPI* hash_api, Longtail_HashAPI_HContext context);
    int (*HashBuffer)(struct Longtail_HashAPI* hash_api, uint32_t length, const void* data, uint64_t* out_hash);
};

/* ---- CompressionAPI + registry, src/longtail.h:258-294 */
struct Longtail_CompressionAPI
{
    struct Longtail_API m_API;
    size_t (*GetMaxCompressedSize)(struct Longtail_CompressionAPI* compression_api, uint32_t settings_id, size_t size);
    int (*Compress)(struct Longtail_CompressionAPI* compression_api, uint32_t settings_id, const char* uncompressed, char* compressed,
                    size_t uncompressed_size, size_t max_compressed_size, size_t* out_compressed_size);
    int (*Decompress)(struct Longtail_CompressionAPI* compression_api, const char* compressed, char* uncompressed, size_t compressed_size,
                      size_t max_uncompressed_size, size_t* out_uncompressed_size);
};

struct Longtail_CompressionRegistryAPI
{
    struct Longtail_API m_API;
    int (*GetCompressionAPI)(struct Longtail_CompressionRegistryAPI* compression_registry, uint32_t compression_type,
                             struct Longtail_CompressionAPI** out_compression_api, uint32_t* out_settings_id);
};

/* ---- StorageAPI, src/longtail.h:306-393.  The hot path calls OpenReadFile, GetSize, Read, CloseFile and ConcatPath only;
 * the other members are declared (as untyped slots) to keep the offsets right. */
typedef struct Longtail_StorageAPI_OpenFile* Longtail_StorageAPI_HOpenFile;
struct Longtail_StorageAPI
{
    struct Longtail_API m_API;
    int (*OpenReadFile)(struct Longtail_StorageAPI* storage_api, const char* path, Longtail_StorageAPI_HOpenFile* out_open_file);
    int (*GetSize)(struct Longtail_StorageAPI* storage_api, Longtail_StorageAPI_HOpenFile f, uint64_t* out_size);
    int (*Read)(struct Longtail_StorageAPI* storage_api, Longtail_StorageAPI_HOpenFile f, uint64_t offset, uint64_t length, void* output);
    void* OpenWriteFile;
    void* Write;
    void* SetSize;
    void* SetPermissions;
    void* GetPermissions;
    void (*CloseFile)(struct Longtail_StorageAPI* storage_api, Longtail_StorageAPI_HOpenFile f);
    void* CreateDir;
    void* RenameFile;
    char* (*ConcatPath)(struct Longtail_StorageAPI* storage_api, const char* root_path, const char* sub_path);
    void* IsDir;
    void* IsFile;
    void* RemoveDir;
    void* RemoveFile;
    void* StartFind;
    void* FindNext;
    void* CloseFind;
    void* GetEntryProperties;
    void* LockFile;
    void* UnlockFile;
    void* GetParentPath;
    void* MapFile;
    void* UnMapFile;
    void* OpenAppendFile;
};

/* ---- ChunkerAPI, src/longtail.h:562-594 */
struct Longtail_ChunkerAPI;
typedef struct Longtail_ChunkerAPI_Chunker* Longtail_ChunkerAPI_HChunker;
typedef int (*Longtail_Chunker_Feeder)(void* context, Longtail_ChunkerAPI_HChunker chunker, uint32_t requested_size, char* buffer, uint32_t* out_size);
struct Longtail_Chunker_ChunkRange
{
    const uint8_t* buf;
    uint64_t offset;
    uint32_t len;
};
struct Longtail_ChunkerAPI
{
    struct Longtail_API m_API;
    int (*GetMinChunkSize)(struct Longtail_ChunkerAPI* chunker_api, uint32_t* out_min_chunk_size);
    int (*CreateChunker)(struct Longtail_ChunkerAPI* chunker_api, uint32_t min_chunk_size, uint32_t avg_chunk_size, uint32_t max_chunk_size,
                         Longtail_ChunkerAPI_HChunker* out_chunker);
    int (*NextChunk)(struct Longtail_ChunkerAPI* chunker_api, Longtail_ChunkerAPI_HChunker chunker, Longtail_Chunker_Feeder feeder,
                     void* feeder_context, struct Longtail_Chunker_ChunkRange* out_chunk_range);
    int (*DisposeChunker)(struct Longtail_ChunkerAPI* chunker_api, Longtail_ChunkerAPI_HChunker chunker);
    int (*NextChunkFromBuffer)(struct Longtail_ChunkerAPI* chunker_api, Longtail_ChunkerAPI_HChunker chunker, const void* buffer,
                               uint64_t buffer_size, const void** out_next_chunk_start);
};

/* ---- data formats on the path */
struct Longtail_FileInfos /* src/longtail.h:1684-1692 */
{
    uint32_t m_Count;
    uint32_t m_PathDataSize;
    uint64_t* m_Sizes;
    uint32_t* m_PathStartOffsets;
    uint16_t* m_Permissions;
    char* m_PathData;
};

struct Longtail_VersionIndex /* src/longtail.h:1856-1881: pointers into one serialised buffer */
{
    uint32_t* m_Version;
    uint32_t* m_HashIdentifier;
    uint32_t* m_TargetChunkSize;
    uint32_t* m_AssetCount;
    uint32_t* m_ChunkCount;
    uint32_t* m_AssetChunkIndexCount;
    TLongtail_Hash* m_PathHashes;
    TLongtail_Hash* m_ContentHashes;
    uint64_t* m_AssetSizes;
    uint32_t* m_AssetChunkCounts;
    uint32_t* m_AssetChunkIndexStarts;
    uint32_t* m_AssetChunkIndexes;
    TLongtail_Hash* m_ChunkHashes;
    uint32_t* m_ChunkSizes;
    uint32_t* m_ChunkTags;
    uint32_t* m_NameOffsets;
    uint32_t m_NameDataSize;
    uint16_t* m_Permissions;
    char* m_NameData;
};

struct Longtail_BlockIndex /* src/longtail.h:1652-1660 */
{
    TLongtail_Hash* m_BlockHash;
    uint32_t* m_HashIdentifier;
    uint32_t* m_ChunkCount;
    uint32_t* m_Tag;
    TLongtail_Hash* m_ChunkHashes;
    uint32_t* m_ChunkSizes;
};

struct Longtail_StoredBlock;
typedef int (*Longtail_StoredBlock_DisposeFunc)(struct Longtail_StoredBlock* stored_block);
struct Longtail_StoredBlock /* src/longtail.h:1669-1675 */
{
    Longtail_StoredBlock_DisposeFunc Dispose;
    struct Longtail_BlockIndex* m_BlockIndex;
    void* m_BlockData;
    uint32_t m_BlockChunksDataSize;
};

struct Longtail_StoreIndex /* src/longtail.h:1699-1711 */
{
    uint32_t* m_Version;
    uint32_t* m_HashIdentifier;
    uint32_t* m_BlockCount;
    uint32_t* m_ChunkCount;
    TLongtail_Hash* m_BlockHashes;
    TLongtail_Hash* m_ChunkHashes;
    uint32_t* m_BlockChunksOffsets;
    uint32_t* m_BlockChunkCounts;
    uint32_t* m_BlockTags;
    uint32_t* m_ChunkSizes;
};

/* ---- BlockStoreAPI and its completion callbacks, src/longtail.h:596-799 */
struct Longtail_AsyncPutStoredBlockAPI
{
    struct Longtail_API m_API;
    void (*OnComplete)(struct Longtail_AsyncPutStoredBlockAPI* async_complete_api, int err);
};
struct Longtail_AsyncGetStoredBlockAPI
{
    struct Longtail_API m_API;
    void (*OnComplete)(struct Longtail_AsyncGetStoredBlockAPI* async_complete_api, struct Longtail_StoredBlock* stored_block, int err);
};
struct Longtail_AsyncGetExistingContentAPI
{
    struct Longtail_API m_API;
    void (*OnComplete)(struct Longtail_AsyncGetExistingContentAPI* async_complete_api, struct Longtail_StoreIndex* store_index, int err);
};
struct Longtail_AsyncPruneBlocksAPI
{
    struct Longtail_API m_API;
    void (*OnComplete)(struct Longtail_AsyncPruneBlocksAPI* async_complete_api, uint32_t pruned_block_count, int err);
};
struct Longtail_AsyncPreflightStartedAPI
{
    struct Longtail_API m_API;
    void (*OnComplete)(struct Longtail_AsyncPreflightStartedAPI* async_complete_api, uint32_t block_count, TLongtail_Hash* block_hashes, int err);
};
struct Longtail_AsyncFlushAPI
{
    struct Longtail_API m_API;
    void (*OnComplete)(struct Longtail_AsyncFlushAPI* async_complete_api, int err);
};

enum /* src/longtail.h:743-775 */
{
    Longtail_BlockStoreAPI_StatU64_GetStoredBlock_Count,
    Longtail_BlockStoreAPI_StatU64_GetStoredBlock_RetryCount,
    Longtail_BlockStoreAPI_StatU64_GetStoredBlock_FailCount,
    Longtail_BlockStoreAPI_StatU64_GetStoredBlock_Chunk_Count,
    Longtail_BlockStoreAPI_StatU64_GetStoredBlock_Byte_Count,
    Longtail_BlockStoreAPI_StatU64_PutStoredBlock_Count,
    Longtail_BlockStoreAPI_StatU64_PutStoredBlock_RetryCount,
    Longtail_BlockStoreAPI_StatU64_PutStoredBlock_FailCount,
    Longtail_BlockStoreAPI_StatU64_PutStoredBlock_Chunk_Count,
    Longtail_BlockStoreAPI_StatU64_PutStoredBlock_Byte_Count,
    Longtail_BlockStoreAPI_StatU64_GetExistingContent_Count,
    Longtail_BlockStoreAPI_StatU64_GetExistingContent_RetryCount,
    Longtail_BlockStoreAPI_StatU64_GetExistingContent_FailCount,
    Longtail_BlockStoreAPI_StatU64_PruneBlocks_Count,
    Longtail_BlockStoreAPI_StatU64_PruneBlocks_RetryCount,
    Longtail_BlockStoreAPI_StatU64_PruneBlocks_FailCount,
    Longtail_BlockStoreAPI_StatU64_PreflightGet_Count,
    Longtail_BlockStoreAPI_StatU64_PreflightGet_RetryCount,
    Longtail_BlockStoreAPI_StatU64_PreflightGet_FailCount,
    Longtail_BlockStoreAPI_StatU64_Flush_Count,
    Longtail_BlockStoreAPI_StatU64_Flush_FailCount,
    Longtail_BlockStoreAPI_StatU64_GetStats_Count,
    Longtail_BlockStoreAPI_StatU64_Count
};
struct Longtail_BlockStore_Stats
{
    uint64_t m_StatU64[Longtail_BlockStoreAPI_StatU64_Count];
};

struct Longtail_BlockStoreAPI
{
    struct Longtail_API m_API;
    int (*PutStoredBlock)(struct Longtail_BlockStoreAPI* block_store_api, struct Longtail_StoredBlock* stored_block,
                          struct Longtail_AsyncPutStoredBlockAPI* async_complete_api);
    int (*PreflightGet)(struct Longtail_BlockStoreAPI* block_store_api, uint32_t block_count, const TLongtail_Hash* block_hashes,
                        struct Longtail_AsyncPreflightStartedAPI* optional_async_complete_api);
    int (*GetStoredBlock)(struct Longtail_BlockStoreAPI* block_store_api, uint64_t block_hash, struct Longtail_AsyncGetStoredBlockAPI* async_complete_api);
    int (*GetExistingContent)(struct Longtail_BlockStoreAPI* block_store_api, uint32_t chunk_count, const TLongtail_Hash* chunk_hashes,
                              uint32_t min_block_usage_percent, struct Longtail_AsyncGetExistingContentAPI* async_complete_api);
    int (*PruneBlocks)(struct Longtail_BlockStoreAPI* block_store_api, uint32_t block_keep_count, const TLongtail_Hash* block_keep_hashes,
                       struct Longtail_AsyncPruneBlocksAPI* async_complete_api);
    int (*GetStats)(struct Longtail_BlockStoreAPI* block_store_api, struct Longtail_BlockStore_Stats* out_stats);
    int (*Flush)(struct Longtail_BlockStoreAPI* block_store_api, struct Longtail_AsyncFlushAPI* async_complete_api);
};

/* src/longtail.h:826-858: progress / tracing callbacks.  The reference keeps the table it is given in a private static
 * (Longtail_SetMonitor, src/longtail.c:762-776) that a library outside longtail.c cannot read, so the B200 verbs take the same
 * struct through Longtail_B200_SetMonitor (include/longtail_b200_api.h). */
typedef void (*Longtail_MonitorGetStoredBlockPrepare)(const struct Longtail_StoreIndex* store_index, uint32_t block_index);
typedef void (*Longtail_MonitorGetStoredBlockLoad)(const struct Longtail_StoreIndex* store_index, uint32_t block_index);
typedef void (*Longtail_MonitorGetStoredBlockLoaded)(const struct Longtail_StoreIndex* store_index, uint32_t block_index, int err);
typedef void (*Longtail_MonitorGetStoredBlockComplete)(const struct Longtail_StoreIndex* store_index, uint32_t block_index, int err);
typedef void (*Longtail_MonitorAssetRemove)(const struct Longtail_VersionIndex* source_version_index, uint32_t asset_index, int err);
typedef void (*Longtail_MonitorAssetOpen)(const struct Longtail_VersionIndex* target_version_index, uint32_t asset_index, int err);
typedef void (*Longtail_MonitorAssetWrite)(const struct Longtail_StoreIndex* target_store_index, const struct Longtail_VersionIndex* version_index,
                                           uint32_t asset_index, uint64_t write_offset, uint32_t size, uint32_t chunk_index, uint32_t chunk_index_in_block,
                                           uint32_t chunk_count_in_block, uint32_t block_index, uint32_t block_data_offset, int err);
typedef void (*Longtail_MonitorChunkRead)(const struct Longtail_StoreIndex* store_index, const struct Longtail_VersionIndex* target_version_index,
                                          uint32_t block_index, uint32_t chunk_index, uint32_t chunk_index_in_block, int err);
typedef void (*Longtail_MonitorBlockCompose)(const struct Longtail_StoreIndex* store_index, uint32_t block_index);
typedef void (*Longtail_MonitorBlockSave)(const struct Longtail_StoreIndex* store_index, uint32_t block_index, uint64_t block_size);
typedef void (*Longtail_MonitorBlockSaved)(const struct Longtail_StoreIndex* store_index, uint32_t block_index, int err);
typedef void (*Longtail_MonitorAssetRead)(const struct Longtail_StoreIndex* store_index, const struct Longtail_VersionIndex* version_index,
                                          uint32_t asset_index, uint64_t read_offset, uint32_t size, TLongtail_Hash chunk_hash, uint32_t block_index,
                                          uint32_t block_data_offset, int err);
typedef void (*Longtail_MonitorAssetClose)(const struct Longtail_VersionIndex* version_index, uint32_t asset_index);
struct Longtail_Monitor
{
    uint64_t StructSize;
    Longtail_MonitorGetStoredBlockPrepare BlockPrepare;
    Longtail_MonitorGetStoredBlockLoad BlockLoad;
    Longtail_MonitorGetStoredBlockLoaded BlockLoaded;
    Longtail_MonitorGetStoredBlockComplete BlockLoadComplete;
    Longtail_MonitorAssetRemove AssetRemove;
    Longtail_MonitorAssetOpen AssetOpen;
    Longtail_MonitorAssetWrite AssetWrite;
    Longtail_MonitorChunkRead ChunkRead;
    Longtail_MonitorBlockCompose BlockCompose;
    Longtail_MonitorBlockSave BlockSave;
    Longtail_MonitorBlockSaved BlockSaved;
    Longtail_MonitorAssetClose AssetClose;
    Longtail_MonitorAssetRead AssetRead;
};

#ifdef __cplusplus
}
#endif
#endif /* LONGTAIL_B200_USE_LONGTAIL_H */
#endif /* LONGTAIL_ABI_H */
