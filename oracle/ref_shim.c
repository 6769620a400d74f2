/* oracle/ref_shim.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin ctypes-friendly C surface over the UNMODIFIED reference (compiled from
 * /root/reference into oracle/_ref/liblongtail_ref.a by oracle/Makefile).  It lets
 * tests/ and bench.py's cpu_baseline / --impl reference leg run the reference's own
 * chunker, hashes, codecs and the Longtail_CreateVersionIndex / Longtail_CreateMissingContent /
 * Longtail_WriteContent verbs on in-memory assets, exactly as cmd/main.c:UpSync wires them
 * (cmd/main.c:972-1153): bikeshed JobAPI, HPCDC chunker, BLAKE3/BLAKE2/Meow hash,
 * full compression registry, compressblockstore on top of a capturing sink or
 * fsblockstore->memstorage.
 *
 * Everything here is this repository's own code; the reference is only #included and linked.
 */
#include "src/longtail.h"
#include "lib/bikeshed/longtail_bikeshed.h"
#include "lib/blake2/longtail_blake2.h"
#include "lib/blake3/longtail_blake3.h"
#include "lib/meowhash/longtail_meowhash.h"
#include "lib/hpcdcchunker/longtail_hpcdcchunker.h"
#include "lib/lz4/longtail_lz4.h"
#include "lib/zstd/longtail_zstd.h"
#include "lib/brotli/longtail_brotli.h"
#include "lib/compressblockstore/longtail_compressblockstore.h"
#include "lib/compressionregistry/longtail_full_compression_registry.h"
#include "lib/filestorage/longtail_filestorage.h"
#include "lib/fsblockstore/longtail_fsblockstore.h"
#include "lib/memstorage/longtail_memstorage.h"
#include "lib/longtail_platform.h"

#include <errno.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define REF_EXPORT __attribute__((visibility("default")))

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

REF_EXPORT void ref_free(void* p) { free(p); }
REF_EXPORT uint32_t ref_cpu_count(void) { return Longtail_GetCPUCount(); }

/* ---------------------------------------------------------------- hashes */

static struct Longtail_HashAPI* make_hash(uint32_t hash_type)
{
    if (hash_type == Longtail_GetBlake3HashType()) return Longtail_CreateBlake3HashAPI();
    if (hash_type == Longtail_GetBlake2HashType()) return Longtail_CreateBlake2HashAPI();
    if (hash_type == Longtail_GetMeowHashType()) return Longtail_CreateMeowHashAPI();
    return 0;
}

REF_EXPORT uint32_t ref_hash_type_blake3(void) { return Longtail_GetBlake3HashType(); }
REF_EXPORT uint32_t ref_hash_type_blake2(void) { return Longtail_GetBlake2HashType(); }
REF_EXPORT uint32_t ref_hash_type_meow(void) { return Longtail_GetMeowHashType(); }

REF_EXPORT int ref_hash_buffer(uint32_t hash_type, const void* data, uint32_t len, uint64_t* out_hash)
{
    struct Longtail_HashAPI* h = make_hash(hash_type);
    if (!h) return EINVAL;
    int err = h->HashBuffer(h, len, data, out_hash);
    SAFE_DISPOSE_API(h);
    return err;
}

/* hash `count` segments (offset,len) of one buffer; used to check per-chunk hashes in bulk */
REF_EXPORT int ref_hash_segments(uint32_t hash_type, const uint8_t* base, uint64_t count,
                                 const uint64_t* offsets, const uint32_t* lens, uint64_t* out_hashes)
{
    struct Longtail_HashAPI* h = make_hash(hash_type);
    if (!h) return EINVAL;
    int err = 0;
    for (uint64_t i = 0; i < count && !err; ++i)
        err = h->HashBuffer(h, lens[i], base + offsets[i], &out_hashes[i]);
    SAFE_DISPOSE_API(h);
    return err;
}

/* ---------------------------------------------------------------- chunker */

struct mem_feeder
{
    const uint8_t* data;
    uint64_t size;
    uint64_t off;
};

static int mem_feed(void* context, Longtail_ChunkerAPI_HChunker chunker, uint32_t requested, char* buffer, uint32_t* out_size)
{
    (void)chunker;
    struct mem_feeder* f = (struct mem_feeder*)context;
    uint64_t n = f->size - f->off;
    if (n > requested) n = requested;
    memcpy(buffer, f->data + f->off, n);
    f->off += n;
    *out_size = (uint32_t)n;
    return 0;
}

/* Longtail_ChunkerAPI.NextChunk over a memory buffer until ESPIPE (the path DynamicChunking
 * takes, src/longtail.c:2231-2296).  Writes chunk lengths; returns 0 or errno. */
REF_EXPORT int ref_hpcdc_chunk(const uint8_t* data, uint64_t size, uint32_t min, uint32_t avg, uint32_t max,
                               uint32_t* out_lens, uint64_t cap, uint64_t* out_count)
{
    struct Longtail_ChunkerAPI* api = Longtail_CreateHPCDCChunkerAPI();
    if (!api) return ENOMEM;
    Longtail_ChunkerAPI_HChunker c;
    int err = api->CreateChunker(api, min, avg, max, &c);
    if (err) { SAFE_DISPOSE_API(api); return err; }
    struct mem_feeder f = {data, size, 0};
    uint64_t n = 0;
    uint64_t expect_off = 0;
    for (;;)
    {
        struct Longtail_Chunker_ChunkRange r;
        err = api->NextChunk(api, c, mem_feed, &f, &r);
        if (err == ESPIPE) { err = 0; break; }
        if (err) break;
        if (r.offset != expect_off) { err = EFAULT; break; }
        if (n >= cap) { err = ENOSPC; break; }
        out_lens[n++] = r.len;
        expect_off += r.len;
    }
    *out_count = n;
    api->DisposeChunker(api, c);
    SAFE_DISPOSE_API(api);
    return err;
}

/* ---------------------------------------------------------------- codecs */

static struct Longtail_CompressionAPI* make_codec(uint32_t type, uint32_t* settings)
{
    struct Longtail_CompressionAPI* c = Longtail_CompressionRegistry_CreateForLZ4(type, settings);
    if (c) return c;
    c = Longtail_CompressionRegistry_CreateForZstd(type, settings);
    if (c) return c;
    return Longtail_CompressionRegistry_CreateForBrotli(type, settings);
}

REF_EXPORT uint64_t ref_compress_bound(uint32_t compression_type, uint64_t size)
{
    uint32_t settings = 0;
    struct Longtail_CompressionAPI* c = make_codec(compression_type, &settings);
    if (!c) return 0;
    uint64_t r = c->GetMaxCompressedSize(c, settings, size);
    SAFE_DISPOSE_API(c);
    return r;
}

REF_EXPORT int ref_compress(uint32_t compression_type, const void* src, uint64_t size, void* dst, uint64_t cap, uint64_t* out_size)
{
    uint32_t settings = 0;
    struct Longtail_CompressionAPI* c = make_codec(compression_type, &settings);
    if (!c) return EINVAL;
    size_t n = 0;
    int err = c->Compress(c, settings, (const char*)src, (char*)dst, size, cap, &n);
    *out_size = n;
    SAFE_DISPOSE_API(c);
    return err;
}

REF_EXPORT int ref_decompress(uint32_t compression_type, const void* src, uint64_t size, void* dst, uint64_t cap, uint64_t* out_size)
{
    uint32_t settings = 0;
    struct Longtail_CompressionAPI* c = make_codec(compression_type, &settings);
    if (!c) return EINVAL;
    size_t n = 0;
    int err = c->Decompress(c, (const char*)src, (char*)dst, size, cap, &n);
    *out_size = n;
    SAFE_DISPOSE_API(c);
    return err;
}

/* ---------------------------------------------------------------- asset storage
 * A read-only Longtail_StorageAPI over caller-owned host buffers (the five functions the
 * hot path uses: ConcatPath, OpenReadFile, GetSize, Read, CloseFile — SURVEY.md §8b). */

struct asset_set
{
    uint32_t count;
    const char** paths;
    const uint8_t** datas;
    const uint64_t* sizes;
    uint32_t* slots; /* open addressing: path hash -> index+1 */
    uint32_t slot_mask;
};

struct asset_storage
{
    struct Longtail_StorageAPI api;
    struct asset_set set;
};

static uint64_t fnv1a(const char* s)
{
    uint64_t h = 0xcbf29ce484222325ull;
    while (*s) { h ^= (uint8_t)*s++; h *= 0x100000001b3ull; }
    return h;
}

static void asset_storage_dispose(struct Longtail_API* api)
{
    struct asset_storage* s = (struct asset_storage*)api;
    free(s->set.slots);
    free(s);
}

static char* asset_concat(struct Longtail_StorageAPI* api, const char* root, const char* sub)
{
    (void)api;
    size_t a = strlen(root), b = strlen(sub);
    char* p = (char*)Longtail_Alloc("ref_shim", a + b + 2);
    memcpy(p, root, a);
    p[a] = '/';
    memcpy(p + a + 1, sub, b + 1);
    return p;
}

static int asset_open(struct Longtail_StorageAPI* api, const char* path, Longtail_StorageAPI_HOpenFile* out)
{
    struct asset_storage* s = (struct asset_storage*)api;
    const char* rel = strchr(path, '/'); /* root is a single component without '/' */
    if (!rel) return ENOENT;
    ++rel;
    uint64_t h = fnv1a(rel);
    for (uint32_t i = (uint32_t)h & s->set.slot_mask;; i = (i + 1) & s->set.slot_mask)
    {
        uint32_t v = s->set.slots[i];
        if (!v) return ENOENT;
        if (strcmp(s->set.paths[v - 1], rel) == 0)
        {
            *out = (Longtail_StorageAPI_HOpenFile)(uintptr_t)v;
            return 0;
        }
    }
}

static int asset_size(struct Longtail_StorageAPI* api, Longtail_StorageAPI_HOpenFile f, uint64_t* out)
{
    struct asset_storage* s = (struct asset_storage*)api;
    *out = s->set.sizes[(uintptr_t)f - 1];
    return 0;
}

static int asset_read(struct Longtail_StorageAPI* api, Longtail_StorageAPI_HOpenFile f, uint64_t offset, uint64_t length, void* output)
{
    struct asset_storage* s = (struct asset_storage*)api;
    uint32_t i = (uint32_t)((uintptr_t)f - 1);
    if (offset + length > s->set.sizes[i]) return EIO;
    memcpy(output, s->set.datas[i] + offset, length);
    return 0;
}

static void asset_close(struct Longtail_StorageAPI* api, Longtail_StorageAPI_HOpenFile f) { (void)api; (void)f; }

static struct Longtail_StorageAPI* make_asset_storage(uint32_t count, const char** paths, const uint8_t** datas, const uint64_t* sizes)
{
    struct asset_storage* s = (struct asset_storage*)calloc(1, sizeof(*s));
    s->api.m_API.Dispose = asset_storage_dispose;
    s->api.OpenReadFile = asset_open;
    s->api.GetSize = asset_size;
    s->api.Read = asset_read;
    s->api.CloseFile = asset_close;
    s->api.ConcatPath = asset_concat;
    s->set.count = count;
    s->set.paths = paths;
    s->set.datas = datas;
    s->set.sizes = sizes;
    uint32_t n = 16;
    while (n < count * 2u) n <<= 1;
    s->set.slot_mask = n - 1;
    s->set.slots = (uint32_t*)calloc(n, sizeof(uint32_t));
    for (uint32_t a = 0; a < count; ++a)
    {
        uint32_t i = (uint32_t)fnv1a(paths[a]) & s->set.slot_mask;
        while (s->set.slots[i]) i = (i + 1) & s->set.slot_mask;
        s->set.slots[i] = a + 1;
    }
    return &s->api;
}

/* Longtail_FileInfos laid out the way Longtail_GetFilesRecursively2 produces it
 * (src/longtail.c:1435-1655): one allocation, paths NUL-terminated back to back. */
static struct Longtail_FileInfos* make_file_infos(uint32_t count, const char** paths, const uint64_t* sizes, const uint16_t* perms)
{
    size_t path_bytes = 0;
    for (uint32_t i = 0; i < count; ++i) path_bytes += strlen(paths[i]) + 1;
    size_t total = sizeof(struct Longtail_FileInfos) + count * (sizeof(uint64_t) + sizeof(uint32_t) + sizeof(uint16_t)) + path_bytes + 16;
    char* mem = (char*)calloc(1, total);
    struct Longtail_FileInfos* fi = (struct Longtail_FileInfos*)mem;
    char* p = mem + sizeof(*fi);
    fi->m_Count = count;
    fi->m_PathDataSize = (uint32_t)path_bytes;
    fi->m_Sizes = (uint64_t*)p; p += count * sizeof(uint64_t);
    fi->m_PathStartOffsets = (uint32_t*)p; p += count * sizeof(uint32_t);
    fi->m_Permissions = (uint16_t*)p; p += count * sizeof(uint16_t);
    fi->m_PathData = p;
    uint32_t off = 0;
    for (uint32_t i = 0; i < count; ++i)
    {
        size_t n = strlen(paths[i]) + 1;
        memcpy(fi->m_PathData + off, paths[i], n);
        fi->m_PathStartOffsets[i] = off;
        fi->m_Sizes[i] = sizes[i];
        fi->m_Permissions[i] = perms ? perms[i] : 0644;
        off += (uint32_t)n;
    }
    return fi;
}

/* ---------------------------------------------------------------- CreateVersionIndex */

/* Runs the reference Longtail_CreateVersionIndex (src/longtail.c:2808) with
 * `workers` bikeshed worker threads (0 = calling thread only, test.cpp:2061) and returns the
 * serialised index (Longtail_WriteVersionIndexToBuffer, src/longtail.c:3415).  *out_seconds is the
 * wall time of the verb alone. */
REF_EXPORT int ref_create_version_index(uint32_t count, const char** paths, const uint8_t** datas, const uint64_t* sizes,
                                        const uint16_t* perms, const uint32_t* tags, uint32_t hash_type,
                                        uint32_t target_chunk_size, uint32_t workers,
                                        void** out_buf, uint64_t* out_size, double* out_seconds)
{
    struct Longtail_HashAPI* hash = make_hash(hash_type);
    if (!hash) return EINVAL;
    struct Longtail_JobAPI* job = Longtail_CreateBikeshedJobAPI(workers, 0);
    struct Longtail_ChunkerAPI* chunker = Longtail_CreateHPCDCChunkerAPI();
    struct Longtail_StorageAPI* storage = make_asset_storage(count, paths, datas, sizes);
    struct Longtail_FileInfos* fi = make_file_infos(count, paths, sizes, perms);
    struct Longtail_VersionIndex* vi = 0;
    double t0 = now_s();
    int err = Longtail_CreateVersionIndex(storage, hash, chunker, job, 0, 0, 0, "root", fi, tags, target_chunk_size, 0, &vi);
    double t1 = now_s();
    if (out_seconds) *out_seconds = t1 - t0;
    if (!err && out_buf)
    {
        void* buf = 0;
        size_t size = 0;
        err = Longtail_WriteVersionIndexToBuffer(vi, &buf, &size);
        if (!err)
        {
            *out_buf = malloc(size ? size : 1);
            memcpy(*out_buf, buf, size);
            *out_size = size;
            Longtail_Free(buf);
        }
    }
    Longtail_Free(vi);
    free(fi);
    SAFE_DISPOSE_API(storage);
    SAFE_DISPOSE_API(chunker);
    SAFE_DISPOSE_API(job);
    SAFE_DISPOSE_API(hash);
    return err;
}

/* ---------------------------------------------------------------- WriteContent with a capturing sink */

/* 64-bit digest of a byte string (four interleaved multiply-rotate lanes, ~10 GB/s): lets a 128 GiB upsync be compared block by block
 * — the reference's serialised StoredBlocks here, the B200 path's in ref_digest_sink — without keeping 64 GiB of blocks around */
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
REF_EXPORT uint64_t ref_digest64(const void* data, uint64_t n)
{
    const uint8_t* p = (const uint8_t*)data;
    uint64_t h[4] = {0x9E3779B97F4A7C15ull ^ n, 0xBF58476D1CE4E5B9ull, 0x94D049BB133111EBull, 0xD6E8FEB86659FD93ull};
    uint64_t i = 0;
    for (; i + 32 <= n; i += 32)
    {
        uint64_t w[4];
        memcpy(w, p + i, 32);
        h[0] = rotl64(h[0] ^ w[0], 29) * 0xA0761D6478BD642Full;
        h[1] = rotl64(h[1] ^ w[1], 31) * 0xE7037ED1A0B428DBull;
        h[2] = rotl64(h[2] ^ w[2], 33) * 0x8EBC6AF09C88C6E3ull;
        h[3] = rotl64(h[3] ^ w[3], 27) * 0x589965CC75374CC3ull;
    }
    uint64_t tail[4] = {0, 0, 0, 0};
    memcpy(tail, p + i, n - i);
    for (int k = 0; k < 4; ++k) h[k] = rotl64(h[k] ^ tail[k], 23) * 0x1D8E4E27C47D124Full;
    uint64_t x = h[0] ^ rotl64(h[1], 17) ^ rotl64(h[2], 34) ^ rotl64(h[3], 51);
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* an lt_b200_block_sink (include/longtail_b200.h) that records {block hash, size, digest} of every block it is handed.
 * user = a ref_digest_list made by ref_digest_list_create */
struct ref_digest_list
{
    uint64_t* rec; /* 3 x u64 per block */
    uint64_t count, cap;
};
struct ref_block_view /* = struct lt_b200_stored_block_view */
{
    uint64_t block_hash;
    const void* data;
    uint64_t size;
    uint32_t chunk_count, tag, raw_payload_size, first_chunk;
};
REF_EXPORT struct ref_digest_list* ref_digest_list_create(void) { return (struct ref_digest_list*)calloc(1, sizeof(struct ref_digest_list)); }
REF_EXPORT void ref_digest_list_free(struct ref_digest_list* l) { if (l) { free(l->rec); free(l); } }
REF_EXPORT const uint64_t* ref_digest_list_data(const struct ref_digest_list* l, uint64_t* out_count) { *out_count = l->count; return l->rec; }
REF_EXPORT int ref_digest_sink(void* user, const struct ref_block_view* v)
{
    struct ref_digest_list* l = (struct ref_digest_list*)user;
    if (l->count == l->cap)
    {
        l->cap = l->cap ? l->cap * 2 : 1024;
        l->rec = (uint64_t*)realloc(l->rec, l->cap * 24);
        if (!l->rec) return ENOMEM;
    }
    uint64_t* r = l->rec + 3 * l->count++;
    r[0] = v->block_hash;
    r[1] = v->size;
    r[2] = ref_digest64(v->data, v->size);
    return 0;
}

struct captured_block
{
    uint64_t hash;
    void* data;
    size_t size;
    uint64_t digest;
};

struct capture_store
{
    struct Longtail_BlockStoreAPI api;
    pthread_mutex_t lock;
    struct captured_block* blocks;
    uint32_t count, cap;
    int keep_bytes;
    uint64_t total_bytes;
};

static void capture_dispose(struct Longtail_API* api)
{
    struct capture_store* s = (struct capture_store*)api;
    for (uint32_t i = 0; i < s->count; ++i) free(s->blocks[i].data);
    free(s->blocks);
    pthread_mutex_destroy(&s->lock);
    free(s);
}

static int capture_put(struct Longtail_BlockStoreAPI* api, struct Longtail_StoredBlock* block, struct Longtail_AsyncPutStoredBlockAPI* async)
{
    struct capture_store* s = (struct capture_store*)api;
    void* buf = 0;
    size_t size = 0;
    int err = Longtail_WriteStoredBlockToBuffer(block, &buf, &size);
    if (!err)
    {
        const uint64_t digest = s->keep_bytes == 2 ? ref_digest64(buf, size) : 0; /* outside the lock: the workers digest in parallel */
        pthread_mutex_lock(&s->lock);
        if (s->count == s->cap)
        {
            s->cap = s->cap ? s->cap * 2 : 64;
            s->blocks = (struct captured_block*)realloc(s->blocks, s->cap * sizeof(*s->blocks));
        }
        struct captured_block* b = &s->blocks[s->count++];
        b->hash = *block->m_BlockIndex->m_BlockHash;
        b->size = size;
        b->data = 0;
        b->digest = digest;
        if (s->keep_bytes == 1)
        {
            b->data = malloc(size);
            memcpy(b->data, buf, size);
        }
        s->total_bytes += size;
        pthread_mutex_unlock(&s->lock);
        Longtail_Free(buf);
    }
    async->OnComplete(async, err);
    return 0;
}

static int capture_flush(struct Longtail_BlockStoreAPI* api, struct Longtail_AsyncFlushAPI* async)
{
    (void)api;
    if (async) async->OnComplete(async, 0);
    return 0;
}

static struct capture_store* make_capture_store(int keep_bytes)
{
    struct capture_store* s = (struct capture_store*)calloc(1, sizeof(*s));
    s->api.m_API.Dispose = capture_dispose;
    s->api.PutStoredBlock = capture_put;
    s->api.Flush = capture_flush;
    s->keep_bytes = keep_bytes;
    pthread_mutex_init(&s->lock, 0);
    return s;
}

/* Reference upsync of a fresh store (cmd/main.c:1052-1153 with an empty existing store index):
 * CreateVersionIndex -> CreateMissingContent(empty) -> WriteContent into
 * compressblockstore(full registry) -> capturing sink.
 *
 * Output buffer (malloc'd): u32 block_count, then per block IN STORE-INDEX ORDER
 * { u64 block_hash, u64 size, u8 serialised_stored_block[size] } (Longtail_WriteStoredBlockToBuffer),
 * followed by the serialised VersionIndex { u64 size, bytes }.  With keep_bytes == 0 only sizes are
 * recorded (timing runs); keep_bytes == 2 records { u64 block_hash, u64 size, u64 ref_digest64 } per block instead of the bytes.  seconds[0..2] = CreateVersionIndex, CreateMissingContent, WriteContent. */
static int upsync_impl(uint32_t count, const char** paths, const uint8_t** datas, const uint64_t* sizes,
                          const uint16_t* perms, const uint32_t* tags, uint32_t hash_type,
                          uint32_t target_chunk_size, uint32_t max_block_size, uint32_t max_chunks_per_block,
                          uint32_t workers, int keep_bytes,
                          void** out_buf, uint64_t* out_size, double* seconds, uint64_t* out_stored_bytes,
                       uint32_t existing_count, const uint64_t* existing_hashes)
{
    struct Longtail_HashAPI* hash = make_hash(hash_type);
    if (!hash) return EINVAL;
    struct Longtail_JobAPI* job = Longtail_CreateBikeshedJobAPI(workers, 0);
    struct Longtail_ChunkerAPI* chunker = Longtail_CreateHPCDCChunkerAPI();
    struct Longtail_StorageAPI* storage = make_asset_storage(count, paths, datas, sizes);
    struct Longtail_FileInfos* fi = make_file_infos(count, paths, sizes, perms);
    struct Longtail_CompressionRegistryAPI* registry = Longtail_CreateFullCompressionRegistry();
    struct capture_store* sink = make_capture_store(keep_bytes);
    struct Longtail_BlockStoreAPI* store = Longtail_CreateCompressBlockStoreAPI(&sink->api, registry);
    struct Longtail_VersionIndex* vi = 0;
    struct Longtail_StoreIndex* empty = 0;
    struct Longtail_StoreIndex* missing = 0;

    double t0 = now_s();
    int err = Longtail_CreateVersionIndex(storage, hash, chunker, job, 0, 0, 0, "root", fi, tags, target_chunk_size, 0, &vi);
    double t1 = now_s();
    if (!err && existing_count == 0) err = Longtail_CreateStoreIndexFromBlocks(0, 0, &empty);
    if (!err && existing_count)
    {
        /* a store that already holds `existing_count` chunks: one synthetic block lists them (sizes are irrelevant to DiffHashes) */
        uint32_t* idx = (uint32_t*)malloc(sizeof(uint32_t) * existing_count);
        uint32_t* szs = (uint32_t*)malloc(sizeof(uint32_t) * existing_count);
        for (uint32_t i = 0; i < existing_count; ++i) { idx[i] = i; szs[i] = 1; }
        struct Longtail_BlockIndex* bi = 0;
        err = Longtail_CreateBlockIndex(hash, 0, existing_count, idx, existing_hashes, szs, &bi);
        if (!err)
        {
            const struct Longtail_BlockIndex* list[1] = {bi};
            err = Longtail_CreateStoreIndexFromBlocks(1, list, &empty);
        }
        Longtail_Free(bi);
        free(idx);
        free(szs);
    }
    double t2 = now_s();
    if (!err) err = Longtail_CreateMissingContent(hash, empty, vi, max_block_size, max_chunks_per_block, &missing);
    double t3 = now_s();
    if (!err) err = Longtail_WriteContent(storage, store, job, 0, 0, 0, missing, vi, "root");
    double t4 = now_s();
    if (seconds) { seconds[0] = t1 - t0; seconds[1] = t3 - t2; seconds[2] = t4 - t3; }
    if (out_stored_bytes) *out_stored_bytes = sink->total_bytes;

    if (!err && out_buf)
    {
        void* vbuf = 0;
        size_t vsize = 0;
        err = Longtail_WriteVersionIndexToBuffer(vi, &vbuf, &vsize);
        if (!err)
        {
            uint32_t block_count = *missing->m_BlockCount;
            size_t total = 4 + 8 + vsize;
            for (uint32_t i = 0; i < sink->count; ++i) total += 16 + (keep_bytes == 1 ? sink->blocks[i].size : 8);
            uint8_t* out = (uint8_t*)malloc(total);
            uint8_t* p = out;
            memcpy(p, &block_count, 4); p += 4;
            for (uint32_t b = 0; b < block_count && !err; ++b)
            {
                uint64_t h = missing->m_BlockHashes[b];
                uint32_t j = b < sink->count && sink->blocks[b].hash == h ? b : 0; /* a single worker stores in order */
                while (j < sink->count && sink->blocks[j].hash != h) ++j;
                if (j == sink->count) { err = ENOENT; break; }
                uint64_t sz = sink->blocks[j].size;
                memcpy(p, &h, 8); p += 8;
                memcpy(p, &sz, 8); p += 8;
                if (keep_bytes == 1) { memcpy(p, sink->blocks[j].data, sz); p += sz; }
                else if (keep_bytes == 2) { memcpy(p, &sink->blocks[j].digest, 8); p += 8; }
            }
            uint64_t vs = vsize;
            memcpy(p, &vs, 8); p += 8;
            memcpy(p, vbuf, vsize); p += vsize;
            *out_buf = out;
            *out_size = (uint64_t)(p - out);
            Longtail_Free(vbuf);
        }
    }
    Longtail_Free(missing);
    Longtail_Free(empty);
    Longtail_Free(vi);
    free(fi);
    SAFE_DISPOSE_API(store);
    SAFE_DISPOSE_API(&sink->api);
    SAFE_DISPOSE_API(registry);
    SAFE_DISPOSE_API(storage);
    SAFE_DISPOSE_API(chunker);
    SAFE_DISPOSE_API(job);
    SAFE_DISPOSE_API(hash);
    return err;
}

REF_EXPORT int ref_upsync(uint32_t count, const char** paths, const uint8_t** datas, const uint64_t* sizes,
                          const uint16_t* perms, const uint32_t* tags, uint32_t hash_type,
                          uint32_t target_chunk_size, uint32_t max_block_size, uint32_t max_chunks_per_block,
                          uint32_t workers, int keep_bytes,
                          void** out_buf, uint64_t* out_size, double* seconds, uint64_t* out_stored_bytes)
{
    return upsync_impl(count, paths, datas, sizes, perms, tags, hash_type, target_chunk_size, max_block_size, max_chunks_per_block, workers,
                       keep_bytes, out_buf, out_size, seconds, out_stored_bytes, 0, 0);
}

/* upsync into a store that already holds the chunks `existing_hashes`: Longtail_CreateMissingContent (src/longtail.c:6882-6998) keeps
 * only the version's chunks the store lacks, in version order (DiffHashes, :6620-6743), and packs those */
REF_EXPORT int ref_upsync_existing(uint32_t count, const char** paths, const uint8_t** datas, const uint64_t* sizes,
                                   const uint16_t* perms, const uint32_t* tags, uint32_t hash_type,
                                   uint32_t target_chunk_size, uint32_t max_block_size, uint32_t max_chunks_per_block,
                                   uint32_t workers, int keep_bytes,
                                   void** out_buf, uint64_t* out_size, double* seconds, uint64_t* out_stored_bytes,
                                   uint32_t existing_count, const uint64_t* existing_hashes)
{
    return upsync_impl(count, paths, datas, sizes, perms, tags, hash_type, target_chunk_size, max_block_size, max_chunks_per_block, workers,
                       keep_bytes, out_buf, out_size, seconds, out_stored_bytes, existing_count, existing_hashes);
}


/* ---------------------------------------------------------------- upsync into an fsblockstore directory
 * Reference behaviour the B200 fs sink (longtail_b200/csrc/fs_store.cpp) is checked against: cmd/main.c:UpSync (:1052-1153) with
 * compressblockstore -> fsblockstore -> filestorage.  The existing content is asked from the store itself, so a second call on the
 * same directory is an incremental upsync. */
struct sync_existing
{
    struct Longtail_AsyncGetExistingContentAPI api;
    struct Longtail_StoreIndex* index;
    int err;
    volatile int done;
};
static void sync_existing_done(struct Longtail_AsyncGetExistingContentAPI* api, struct Longtail_StoreIndex* index, int err)
{
    struct sync_existing* s = (struct sync_existing*)api;
    s->index = index;
    s->err = err;
    __sync_synchronize();
    s->done = 1;
}
struct sync_flush
{
    struct Longtail_AsyncFlushAPI api;
    int err;
    volatile int done;
};
static void sync_flush_done(struct Longtail_AsyncFlushAPI* api, int err)
{
    struct sync_flush* s = (struct sync_flush*)api;
    s->err = err;
    __sync_synchronize();
    s->done = 1;
}
struct sync_get
{
    struct Longtail_AsyncGetStoredBlockAPI api;
    struct Longtail_StoredBlock* block;
    int err;
    volatile int done;
};
static void sync_get_done(struct Longtail_AsyncGetStoredBlockAPI* api, struct Longtail_StoredBlock* block, int err)
{
    struct sync_get* s = (struct sync_get*)api;
    s->block = block;
    s->err = err;
    __sync_synchronize();
    s->done = 1;
}

REF_EXPORT int ref_upsync_to_dir(uint32_t count, const char** paths, const uint8_t** datas, const uint64_t* sizes,
                                 const uint16_t* perms, const uint32_t* tags, uint32_t hash_type,
                                 uint32_t target_chunk_size, uint32_t max_block_size, uint32_t max_chunks_per_block,
                                 uint32_t workers, const char* dir, uint32_t* out_blocks_written)
{
    struct Longtail_HashAPI* hash = make_hash(hash_type);
    if (!hash) return EINVAL;
    struct Longtail_JobAPI* job = Longtail_CreateBikeshedJobAPI(workers, 0);
    struct Longtail_ChunkerAPI* chunker = Longtail_CreateHPCDCChunkerAPI();
    struct Longtail_StorageAPI* storage = make_asset_storage(count, paths, datas, sizes);
    struct Longtail_FileInfos* fi = make_file_infos(count, paths, sizes, perms);
    struct Longtail_CompressionRegistryAPI* registry = Longtail_CreateFullCompressionRegistry();
    struct Longtail_StorageAPI* fs = Longtail_CreateFSStorageAPI();
    struct Longtail_BlockStoreAPI* fs_store = Longtail_CreateFSBlockStoreAPI(job, fs, dir, 0, 0);
    struct Longtail_BlockStoreAPI* store = Longtail_CreateCompressBlockStoreAPI(fs_store, registry);
    struct Longtail_VersionIndex* vi = 0;
    struct Longtail_StoreIndex* missing = 0;
    struct sync_existing ex;
    memset(&ex, 0, sizeof(ex));
    ex.api.OnComplete = sync_existing_done;
    int err = Longtail_CreateVersionIndex(storage, hash, chunker, job, 0, 0, 0, "root", fi, tags, target_chunk_size, 0, &vi);
    if (!err) err = store->GetExistingContent(store, *vi->m_ChunkCount, vi->m_ChunkHashes, 0, &ex.api);
    if (!err)
    {
        while (!ex.done) sched_yield();
        err = ex.err;
    }
    if (!err) err = Longtail_CreateMissingContent(hash, ex.index, vi, max_block_size, max_chunks_per_block, &missing);
    if (!err && out_blocks_written) *out_blocks_written = *missing->m_BlockCount;
    if (!err) err = Longtail_WriteContent(storage, store, job, 0, 0, 0, missing, vi, "root");
    if (!err)
    {
        struct sync_flush fl;
        memset(&fl, 0, sizeof(fl));
        fl.api.OnComplete = sync_flush_done;
        err = store->Flush(store, &fl.api);
        if (!err)
        {
            while (!fl.done) sched_yield();
            err = fl.err;
        }
    }
    Longtail_Free(missing);
    Longtail_Free(ex.index);
    Longtail_Free(vi);
    free(fi);
    SAFE_DISPOSE_API(store);
    SAFE_DISPOSE_API(fs_store);
    SAFE_DISPOSE_API(fs);
    SAFE_DISPOSE_API(registry);
    SAFE_DISPOSE_API(storage);
    SAFE_DISPOSE_API(chunker);
    SAFE_DISPOSE_API(job);
    SAFE_DISPOSE_API(hash);
    return err;
}

/* The UNMODIFIED reference opens `dir` as a block store (fsblockstore behind compressblockstore), takes the block list from store.lsi and
 * reads every block back, decompressed.  out[0] = blocks, out[1] = chunks, out[2] = uncompressed payload bytes,
 * out[3] = order-independent digest (sum over blocks of fnv1a(block hash, chunk hashes, chunk sizes, payload)). */
REF_EXPORT int ref_read_store_dir(const char* dir, uint64_t out[4])
{
    struct Longtail_JobAPI* job = Longtail_CreateBikeshedJobAPI(0, 0);
    struct Longtail_CompressionRegistryAPI* registry = Longtail_CreateFullCompressionRegistry();
    struct Longtail_StorageAPI* fs = Longtail_CreateFSStorageAPI();
    struct Longtail_BlockStoreAPI* fs_store = Longtail_CreateFSBlockStoreAPI(job, fs, dir, 0, 0);
    struct Longtail_BlockStoreAPI* store = Longtail_CreateCompressBlockStoreAPI(fs_store, registry);
    struct Longtail_StoreIndex* index = 0;
    char* index_path = fs->ConcatPath(fs, dir, "store.lsi");
    int err = Longtail_ReadStoreIndex(fs, index_path, &index);
    Longtail_Free(index_path);
    memset(out, 0, sizeof(uint64_t) * 4);
    for (uint32_t b = 0; !err && b < *index->m_BlockCount; ++b)
    {
        struct sync_get g;
        memset(&g, 0, sizeof(g));
        g.api.OnComplete = sync_get_done;
        err = store->GetStoredBlock(store, index->m_BlockHashes[b], &g.api);
        if (err) break;
        while (!g.done) sched_yield();
        err = g.err;
        if (err) break;
        const struct Longtail_BlockIndex* bi = g.block->m_BlockIndex;
        const uint32_t n = *bi->m_ChunkCount;
        if (*bi->m_BlockHash != index->m_BlockHashes[b] || n != index->m_BlockChunkCounts[b]) err = EBADF;
        uint64_t h = 1469598103934665603ull;
        const uint8_t* parts[4] = {(const uint8_t*)bi->m_BlockHash, (const uint8_t*)bi->m_ChunkHashes, (const uint8_t*)bi->m_ChunkSizes, (const uint8_t*)g.block->m_BlockData};
        const size_t lens[4] = {8, 8 * (size_t)n, 4 * (size_t)n, g.block->m_BlockChunksDataSize};
        for (int k = 0; k < 4; ++k)
            for (size_t i = 0; i < lens[k]; ++i) { h ^= parts[k][i]; h *= 1099511628211ull; }
        out[0] += 1;
        out[1] += n;
        out[2] += g.block->m_BlockChunksDataSize;
        out[3] += h;
        if (g.block->Dispose) g.block->Dispose(g.block);
    }
    Longtail_Free(index);
    SAFE_DISPOSE_API(store);
    SAFE_DISPOSE_API(fs_store);
    SAFE_DISPOSE_API(fs);
    SAFE_DISPOSE_API(registry);
    SAFE_DISPOSE_API(job);
    return err;
}

/* ---------------------------------------------------------------- directory scan + index through the reference's file storage
 * (SURVEY.md section 8f row 4): cmd/main.c:UpSync's Longtail_GetFilesRecursively2 + Longtail_CreateVersionIndex over a real directory.
 * ref_scan_directory serialises the FileInfos: u32 count, then per entry { u64 size, u16 permissions, u32 path length, path bytes }. */
REF_EXPORT int ref_scan_directory(const char* root, uint32_t workers, void** out_buf, uint64_t* out_size)
{
    struct Longtail_StorageAPI* fs = Longtail_CreateFSStorageAPI();
    struct Longtail_JobAPI* job = workers ? Longtail_CreateBikeshedJobAPI(workers, 0) : 0;
    struct Longtail_FileInfos* fi = 0;
    int err = Longtail_GetFilesRecursively2(fs, job, 0, 0, 0, root, &fi);
    if (!err)
    {
        size_t total = 4;
        for (uint32_t i = 0; i < fi->m_Count; ++i) total += 8 + 2 + 4 + strlen(&fi->m_PathData[fi->m_PathStartOffsets[i]]);
        uint8_t* out = (uint8_t*)malloc(total);
        uint8_t* p = out;
        memcpy(p, &fi->m_Count, 4); p += 4;
        for (uint32_t i = 0; i < fi->m_Count; ++i)
        {
            const char* path = &fi->m_PathData[fi->m_PathStartOffsets[i]];
            uint32_t n = (uint32_t)strlen(path);
            memcpy(p, &fi->m_Sizes[i], 8); p += 8;
            memcpy(p, &fi->m_Permissions[i], 2); p += 2;
            memcpy(p, &n, 4); p += 4;
            memcpy(p, path, n); p += n;
        }
        *out_buf = out;
        *out_size = (uint64_t)(p - out);
    }
    Longtail_Free(fi);
    if (job) SAFE_DISPOSE_API(job);
    SAFE_DISPOSE_API(fs);
    return err;
}

REF_EXPORT int ref_index_directory(const char* root, uint32_t hash_type, uint32_t target_chunk_size, uint32_t workers, uint32_t tag,
                                   void** out_buf, uint64_t* out_size)
{
    struct Longtail_HashAPI* hash = make_hash(hash_type);
    if (!hash) return EINVAL;
    struct Longtail_StorageAPI* fs = Longtail_CreateFSStorageAPI();
    struct Longtail_JobAPI* job = Longtail_CreateBikeshedJobAPI(workers, 0);
    struct Longtail_ChunkerAPI* chunker = Longtail_CreateHPCDCChunkerAPI();
    struct Longtail_FileInfos* fi = 0;
    struct Longtail_VersionIndex* vi = 0;
    int err = Longtail_GetFilesRecursively2(fs, job, 0, 0, 0, root, &fi);
    uint32_t* tags = 0;
    if (!err)
    {
        tags = (uint32_t*)malloc(sizeof(uint32_t) * (fi->m_Count ? fi->m_Count : 1));
        for (uint32_t i = 0; i < fi->m_Count; ++i) tags[i] = tag;
        err = Longtail_CreateVersionIndex(fs, hash, chunker, job, 0, 0, 0, root, fi, tags, target_chunk_size, 0, &vi);
    }
    if (!err)
    {
        void* buf = 0;
        size_t size = 0;
        err = Longtail_WriteVersionIndexToBuffer(vi, &buf, &size);
        if (!err)
        {
            *out_buf = malloc(size ? size : 1);
            memcpy(*out_buf, buf, size);
            *out_size = size;
            Longtail_Free(buf);
        }
    }
    free(tags);
    Longtail_Free(vi);
    Longtail_Free(fi);
    SAFE_DISPOSE_API(chunker);
    SAFE_DISPOSE_API(job);
    SAFE_DISPOSE_API(fs);
    SAFE_DISPOSE_API(hash);
    return err;
}

/* cmd/main.c:UpSync over real directories: source tree `source_root` (file storage) -> store directory `store_dir` (compressblockstore ->
 * fsblockstore); the serialised VersionIndex is returned.  A second call with the same store is an incremental upsync. */
REF_EXPORT int ref_upsync_dir_to_dir(const char* source_root, const char* store_dir, uint32_t hash_type, uint32_t target_chunk_size,
                                     uint32_t max_block_size, uint32_t max_chunks_per_block, uint32_t workers, uint32_t tag,
                                     void** out_buf, uint64_t* out_size, uint32_t* out_blocks_written)
{
    struct Longtail_HashAPI* hash = make_hash(hash_type);
    if (!hash) return EINVAL;
    struct Longtail_StorageAPI* fs = Longtail_CreateFSStorageAPI();
    struct Longtail_JobAPI* job = Longtail_CreateBikeshedJobAPI(workers, 0);
    struct Longtail_ChunkerAPI* chunker = Longtail_CreateHPCDCChunkerAPI();
    struct Longtail_CompressionRegistryAPI* registry = Longtail_CreateFullCompressionRegistry();
    struct Longtail_BlockStoreAPI* fs_store = Longtail_CreateFSBlockStoreAPI(job, fs, store_dir, 0, 0);
    struct Longtail_BlockStoreAPI* store = Longtail_CreateCompressBlockStoreAPI(fs_store, registry);
    struct Longtail_FileInfos* fi = 0;
    struct Longtail_VersionIndex* vi = 0;
    struct Longtail_StoreIndex* missing = 0;
    uint32_t* tags = 0;
    struct sync_existing ex;
    memset(&ex, 0, sizeof(ex));
    ex.api.OnComplete = sync_existing_done;
    int err = Longtail_GetFilesRecursively2(fs, job, 0, 0, 0, source_root, &fi);
    if (!err)
    {
        tags = (uint32_t*)malloc(sizeof(uint32_t) * (fi->m_Count ? fi->m_Count : 1));
        for (uint32_t i = 0; i < fi->m_Count; ++i) tags[i] = tag;
        err = Longtail_CreateVersionIndex(fs, hash, chunker, job, 0, 0, 0, source_root, fi, tags, target_chunk_size, 0, &vi);
    }
    if (!err) err = store->GetExistingContent(store, *vi->m_ChunkCount, vi->m_ChunkHashes, 0, &ex.api);
    if (!err)
    {
        while (!ex.done) sched_yield();
        err = ex.err;
    }
    if (!err) err = Longtail_CreateMissingContent(hash, ex.index, vi, max_block_size, max_chunks_per_block, &missing);
    if (!err && out_blocks_written) *out_blocks_written = *missing->m_BlockCount;
    if (!err) err = Longtail_WriteContent(fs, store, job, 0, 0, 0, missing, vi, source_root);
    if (!err)
    {
        struct sync_flush fl;
        memset(&fl, 0, sizeof(fl));
        fl.api.OnComplete = sync_flush_done;
        err = store->Flush(store, &fl.api);
        if (!err)
        {
            while (!fl.done) sched_yield();
            err = fl.err;
        }
    }
    if (!err)
    {
        void* buf = 0;
        size_t size = 0;
        err = Longtail_WriteVersionIndexToBuffer(vi, &buf, &size);
        if (!err)
        {
            *out_buf = malloc(size ? size : 1);
            memcpy(*out_buf, buf, size);
            *out_size = size;
            Longtail_Free(buf);
        }
    }
    free(tags);
    Longtail_Free(missing);
    Longtail_Free(ex.index);
    Longtail_Free(vi);
    Longtail_Free(fi);
    SAFE_DISPOSE_API(store);
    SAFE_DISPOSE_API(fs_store);
    SAFE_DISPOSE_API(registry);
    SAFE_DISPOSE_API(chunker);
    SAFE_DISPOSE_API(job);
    SAFE_DISPOSE_API(fs);
    SAFE_DISPOSE_API(hash);
    return err;
}
