// longtail_api.cpp — drop-in Longtail_*API objects over the C ABI (include/longtail_b200_api.h).
//
// Host code only.  It mirrors the reference's object conventions: callback structs whose first member is Longtail_API
// (src/longtail.h:43-46), errno returns, ESPIPE at end of stream, Dispose through the struct.  All data-path work is done
// by the kernels behind lt_b200_*; nothing here hashes or scans a byte on the CPU.
#include "../../include/longtail_b200_api.h"

#include <dlfcn.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

// ---------------------------------------------------------------- allocation: the host's Longtail_Alloc when present
typedef void* (*AllocFn)(const char*, size_t);
typedef void (*FreeFn)(void*);
AllocFn g_alloc = nullptr;
FreeFn g_free = nullptr;
std::once_flag g_alloc_once;

void resolve_alloc()
{
    g_alloc = reinterpret_cast<AllocFn>(dlsym(RTLD_DEFAULT, "Longtail_Alloc"));
    g_free = reinterpret_cast<FreeFn>(dlsym(RTLD_DEFAULT, "Longtail_Free"));
    if (!g_alloc || !g_free) { g_alloc = nullptr; g_free = nullptr; }
}
void* lt_alloc(const char* what, size_t n)
{
    std::call_once(g_alloc_once, resolve_alloc);
    return g_alloc ? g_alloc(what, n) : malloc(n ? n : 1);
}
void lt_free(void* p)
{
    std::call_once(g_alloc_once, resolve_alloc);
    if (g_free) g_free(p); else free(p);
}

// ---------------------------------------------------------------- one shared device context, serialised
std::mutex g_gpu;            // guards every use of g_ctx (one stream, one workspace)
lt_b200_context* g_ctx = nullptr;
int g_device = 0;
void* g_arena = nullptr;     // device scratch for the per-call objects
uint64_t g_arena_cap = 0;

int ensure_ctx()
{
    if (g_ctx) return 0;
    return lt_b200_context_create(g_device, &g_ctx);
}
// The two batched verbs (Longtail_B200_CreateVersionIndex / Longtail_B200_WriteContent) call back into foreign code — the caller's
// StorageAPI and JobAPI, the backing store's PutStoredBlock — while they own a context.  Those callbacks may use the B200 hash /
// compression / chunker objects from any thread, so the verbs run on a context and a lock of their own: a callback that needs g_gpu
// never waits for the verb that is waiting for it.
std::mutex g_verb;
lt_b200_context* g_verb_ctx = nullptr;
int ensure_verb_ctx()
{
    if (g_verb_ctx) return 0;
    return lt_b200_context_create(g_device, &g_verb_ctx);
}
int ensure_arena(uint64_t bytes)
{
    if (g_arena_cap >= bytes) return 0;
    if (g_arena) lt_b200_device_free(g_ctx, g_arena);
    g_arena = nullptr;
    g_arena_cap = 0;
    uint64_t want = bytes + bytes / 4 + 4096;
    int err = lt_b200_device_alloc(g_ctx, want, &g_arena);
    if (err) return err;
    g_arena_cap = want;
    return 0;
}

const uint32_t HASH_BLK3 = LT_B200_HASH_BLAKE3;

// ---------------------------------------------------------------- chunker
struct ChunkRec
{
    uint64_t offset;
    uint32_t len;
    uint64_t hash;
};

struct B200Chunker
{
    uint32_t mn, av, mx;
    uint32_t hash_type = LT_B200_HASH_BLAKE3; // which HashAPI the cached chunk hashes belong to
    uint8_t* buf = nullptr; // pinned host memory holding the whole part
    uint64_t cap = 0;
    uint64_t size = 0;
    bool scanned = false;
    std::vector<ChunkRec> chunks;
    size_t next = 0;
};

struct B200ChunkerAPI
{
    struct Longtail_ChunkerAPI api;
};

std::mutex g_registry_lock;
std::vector<B200Chunker*> g_live; // chunkers whose ranges may be presented to HashBuffer

int chunker_get_min(struct Longtail_ChunkerAPI*, uint32_t* out)
{
    if (!out) return EINVAL;
    *out = 48; // lib/hpcdcchunker/longtail_hpcdcchunker.c:332-346
    return 0;
}

int chunker_create(struct Longtail_ChunkerAPI*, uint32_t mn, uint32_t av, uint32_t mx, Longtail_ChunkerAPI_HChunker* out)
{
    if (!out) return EINVAL;
    if (mn < 48 || mn > mx || mn > av || av > mx) return EINVAL; // longtail_hpcdcchunker.c:146-150
    B200Chunker* c = new (std::nothrow) B200Chunker();
    if (!c) return ENOMEM;
    c->mn = mn;
    c->av = av;
    c->mx = mx;
    *out = reinterpret_cast<Longtail_ChunkerAPI_HChunker>(c);
    return 0;
}

void chunker_release_buffer(B200Chunker* c)
{
    if (c->buf)
    {
        std::lock_guard<std::mutex> g(g_gpu);
        if (g_ctx) lt_b200_host_free_pinned(g_ctx, c->buf);
        c->buf = nullptr;
    }
}

int chunker_dispose(struct Longtail_ChunkerAPI*, Longtail_ChunkerAPI_HChunker h)
{
    B200Chunker* c = reinterpret_cast<B200Chunker*>(h);
    if (!c) return EINVAL;
    {
        std::lock_guard<std::mutex> g(g_registry_lock);
        g_live.erase(std::remove(g_live.begin(), g_live.end(), c), g_live.end());
    }
    chunker_release_buffer(c);
    delete c;
    return 0;
}

// drain the feeder, then one GPU pass over the part: boundaries + chunk hashes
int chunker_scan(B200Chunker* c, Longtail_Chunker_Feeder feeder, void* feeder_context)
{
    // the reference asks its feeder for at most 4*max bytes at a time (longtail_hpcdcchunker.c:160-164,206); a part is at most
    // target*1024 = 512*max bytes, so start there and grow if a caller feeds more
    uint64_t want = (uint64_t)c->mx * 512 + 4096;
    for (;;)
    {
        if (c->cap < want)
        {
            std::lock_guard<std::mutex> g(g_gpu);
            int err = ensure_ctx();
            if (err) return err;
            void* nb = nullptr;
            err = lt_b200_host_alloc_pinned(g_ctx, want, &nb);
            if (err) return err;
            if (c->buf)
            {
                memcpy(nb, c->buf, c->size);
                lt_b200_host_free_pinned(g_ctx, c->buf);
            }
            c->buf = static_cast<uint8_t*>(nb);
            c->cap = want;
        }
        uint64_t room = c->cap - c->size;
        uint32_t ask = room > 0x40000000u ? 0x40000000u : (uint32_t)room;
        uint32_t got = 0;
        int err = feeder(feeder_context, reinterpret_cast<Longtail_ChunkerAPI_HChunker>(c), ask, reinterpret_cast<char*>(c->buf + c->size), &got);
        if (err) return err;
        c->size += got;
        if (got == 0) break;
        if (c->size == c->cap) want = c->cap * 2;
    }
    if (c->size >= 0x80000000ull) return E2BIG;
    if (c->size)
    {
        std::lock_guard<std::mutex> g(g_gpu);
        int err = ensure_ctx();
        if (!err) err = ensure_arena(c->size);
        if (!err) err = lt_b200_copy_to_device(g_ctx, g_arena, c->buf, c->size);
        if (err) return err;
        lt_b200_range r = {0, (uint32_t)c->size, 0};
        lt_b200_chunk_table t;
        err = lt_b200_chunk_ranges(g_ctx, static_cast<const uint8_t*>(g_arena), g_arena_cap, &r, 1, c->mn, c->av, c->mx, HASH_BLK3, 1, &t);
        if (err) return err;
        c->chunks.resize(t.chunk_count);
        for (uint32_t i = 0; i < t.chunk_count; ++i) c->chunks[i] = {t.chunk_offsets[i], t.chunk_sizes[i], t.chunk_hashes[i]};
    }
    c->scanned = true;
    std::lock_guard<std::mutex> g(g_registry_lock);
    g_live.push_back(c);
    return 0;
}

int chunker_next(struct Longtail_ChunkerAPI*, Longtail_ChunkerAPI_HChunker h, Longtail_Chunker_Feeder feeder, void* feeder_context,
                 struct Longtail_Chunker_ChunkRange* out)
{
    B200Chunker* c = reinterpret_cast<B200Chunker*>(h);
    if (!c || !feeder || !out) return EINVAL;
    if (!c->scanned)
    {
        int err = chunker_scan(c, feeder, feeder_context);
        if (err) return err;
    }
    if (c->next >= c->chunks.size())
    {
        out->buf = nullptr; // longtail_hpcdcchunker.c:250-255,420-423: {0, total, 0} and ESPIPE
        out->offset = c->size;
        out->len = 0;
        return ESPIPE;
    }
    const ChunkRec& r = c->chunks[c->next++];
    out->buf = c->buf + r.offset;
    out->offset = r.offset;
    out->len = r.len;
    return 0;
}

int chunker_next_from_buffer(struct Longtail_ChunkerAPI*, Longtail_ChunkerAPI_HChunker, const void*, uint64_t, const void**)
{
    return ENOTSUP; // mmap branch: unreachable from Longtail_CreateVersionIndex (src/longtail.c:2453); see header
}

void chunker_api_dispose(struct Longtail_API* api) { lt_free(api); }

// ---------------------------------------------------------------- hash
struct B200HashAPI
{
    struct Longtail_HashAPI api;
    uint32_t type;
};
uint32_t hash_type_of(struct Longtail_HashAPI* api) { return reinterpret_cast<B200HashAPI*>(api)->type; }

struct HashStream
{
    std::vector<uint8_t> bytes;
};

uint32_t hash_identifier(struct Longtail_HashAPI* api) { return hash_type_of(api); }

// the chunker hashed its chunks with BLAKE3 in its own pass; another HashAPI asking for them gets the whole part re-hashed
// with its algorithm once (one upload, one launch), not chunk by chunk
int rehash_chunker(B200Chunker* c, uint32_t type)
{
    std::lock_guard<std::mutex> g(g_gpu);
    int err = ensure_ctx();
    if (!err) err = ensure_arena(c->size + 64);
    if (!err && c->size) err = lt_b200_copy_to_device(g_ctx, g_arena, c->buf, c->size);
    if (err) return err;
    const uint32_t n = (uint32_t)c->chunks.size();
    std::vector<uint64_t> off(n), out(n);
    std::vector<uint32_t> len(n);
    for (uint32_t i = 0; i < n; ++i) { off[i] = c->chunks[i].offset; len[i] = c->chunks[i].len; }
    err = lt_b200_hash_segments(g_ctx, type, static_cast<const uint8_t*>(g_arena), g_arena_cap, off.data(), len.data(), n, out.data());
    if (err) return err;
    for (uint32_t i = 0; i < n; ++i) c->chunks[i].hash = out[i];
    c->hash_type = type;
    return 0;
}

int hash_on_device(uint32_t type, uint32_t length, const void* data, uint64_t* out)
{
    std::lock_guard<std::mutex> g(g_gpu);
    int err = ensure_ctx();
    if (!err) err = ensure_arena((uint64_t)length + 64);
    if (!err && length) err = lt_b200_copy_to_device(g_ctx, g_arena, data, length);
    if (err) return err;
    uint64_t off = 0;
    return lt_b200_hash_segments(g_ctx, type, static_cast<const uint8_t*>(g_arena), g_arena_cap, &off, &length, 1, out);
}

int hash_buffer(struct Longtail_HashAPI* api, uint32_t length, const void* data, uint64_t* out)
{
    if (!out || (!data && length)) return EINVAL;
    const uint32_t type = hash_type_of(api);
    const uint8_t* p = static_cast<const uint8_t*>(data);
    // a range handed out by one of our chunkers: its hash was computed in the chunker's GPU pass.  The registry lock only covers the
    // search for the owning chunker; the chunker itself is driven by this thread alone, so the lookup and a re-hash with
    // another algorithm (an upload plus a launch) run without it and do not serialise the other hashing threads
    B200Chunker* owner = nullptr;
    {
        std::lock_guard<std::mutex> g(g_registry_lock);
        for (B200Chunker* c : g_live)
            if (c->buf && p >= c->buf && p < c->buf + c->size)
            {
                owner = c;
                break;
            }
    }
    if (owner)
    {
        const uint64_t off = (uint64_t)(p - owner->buf);
        auto it = std::lower_bound(owner->chunks.begin(), owner->chunks.end(), off, [](const ChunkRec& r, uint64_t o) { return r.offset < o; });
        if (it != owner->chunks.end() && it->offset == off && it->len == length && (owner->hash_type == type || rehash_chunker(owner, type) == 0))
        {
            *out = it->hash;
            return 0;
        }
    }
    return hash_on_device(type, length, data, out);
}

int hash_begin(struct Longtail_HashAPI*, Longtail_HashAPI_HContext* out)
{
    if (!out) return EINVAL;
    HashStream* s = new (std::nothrow) HashStream();
    if (!s) return ENOMEM;
    *out = reinterpret_cast<Longtail_HashAPI_HContext>(s);
    return 0;
}

void hash_update(struct Longtail_HashAPI*, Longtail_HashAPI_HContext h, uint32_t length, const void* data)
{
    HashStream* s = reinterpret_cast<HashStream*>(h);
    const uint8_t* p = static_cast<const uint8_t*>(data);
    s->bytes.insert(s->bytes.end(), p, p + length);
}

uint64_t hash_end(struct Longtail_HashAPI* api, Longtail_HashAPI_HContext h)
{
    HashStream* s = reinterpret_cast<HashStream*>(h);
    uint64_t out = 0;
    // EndContext frees the context and has no error channel (lib/blake3/longtail_blake3.c:60-79): a device failure is at least reported
    const int err = hash_on_device(hash_type_of(api), (uint32_t)s->bytes.size(), s->bytes.data(), &out);
    if (err) fprintf(stderr, "longtail_b200: HashAPI.EndContext: device hash failed with %d, returning 0\n", err);
    delete s;
    return out;
}

void hash_api_dispose(struct Longtail_API* api) { lt_free(api); }

// ---------------------------------------------------------------- CreateVersionIndex
struct ReadCtx
{
    struct Longtail_StorageAPI* storage;
    const char* root;
    const struct Longtail_FileInfos* infos;
    struct Longtail_JobAPI* jobs;
    struct Longtail_ProgressAPI* progress;
    struct Longtail_CancelAPI* cancel;
    Longtail_CancelAPI_HCancelToken token;
    uint32_t total_jobs;
    uint32_t done_jobs;
};

struct OneRead
{
    ReadCtx* ctx;
    lt_b200_read_job job;
};

// the storage calls of DynamicChunking (src/longtail.c:2024-2026, 2076, 1950)
int read_one(void* context, uint32_t, int detected_error)
{
    if (detected_error) return 0; // src/longtail.c:2000-2004
    OneRead* r = static_cast<OneRead*>(context);
    struct Longtail_StorageAPI* st = r->ctx->storage;
    const char* rel = &r->ctx->infos->m_PathData[r->ctx->infos->m_PathStartOffsets[r->job.asset_index]];
    char* path = st->ConcatPath(st, r->ctx->root, rel);
    if (!path) return ENOMEM;
    Longtail_StorageAPI_HOpenFile f;
    int err = st->OpenReadFile(st, path, &f);
    if (!err)
    {
        err = st->Read(st, f, r->job.offset, r->job.size, r->job.dst);
        st->CloseFile(st, f);
    }
    lt_free(path);
    return err;
}

int read_batch(void* user, const lt_b200_read_job* jobs, uint32_t count)
{
    ReadCtx* c = static_cast<ReadCtx*>(user);
    if (c->cancel && c->cancel->IsCancelled(c->cancel, c->token) == ECANCELED) return ECANCELED;
    std::vector<OneRead> reads(count);
    for (uint32_t i = 0; i < count; ++i) reads[i] = {c, jobs[i]};
    int err = 0;
    if (c->jobs && count > 1)
    {
        std::vector<Longtail_JobAPI_JobFunc> funcs(count, read_one);
        std::vector<void*> ctxs(count);
        for (uint32_t i = 0; i < count; ++i) ctxs[i] = &reads[i];
        uint32_t max_batch = 0;
        err = c->jobs->GetMaxBatchCount(c->jobs, &max_batch, 0);
        if (err) return err;
        if (max_batch == 0) max_batch = count;
        Longtail_JobAPI_Group group = 0;
        err = c->jobs->ReserveJobs(c->jobs, count, &group);
        if (err) return err;
        for (uint32_t s = 0; s < count && !err;)
        {
            uint32_t n = std::min(max_batch, count - s);
            Longtail_JobAPI_Jobs handle;
            err = c->jobs->CreateJobs(c->jobs, group, 0, c->cancel, c->token, n, &funcs[s], &ctxs[s], 0, &handle);
            if (!err) err = c->jobs->ReadyJobs(c->jobs, n, handle);
            s += n;
        }
        int werr = c->jobs->WaitForAllJobs(c->jobs, group, 0, c->cancel, c->token);
        if (!err) err = werr;
    }
    else
    {
        for (uint32_t i = 0; i < count && !err; ++i) err = read_one(&reads[i], 0, 0);
    }
    c->done_jobs += count;
    if (!err && c->progress) c->progress->OnProgress(c->progress, c->total_jobs, c->done_jobs);
    return err;
}

// Longtail_VersionIndex = pointer struct followed by the serialised data (src/longtail.c:2630-2704 InitVersionIndexFromData)
struct Longtail_VersionIndex* wrap_version_index(const void* data, uint64_t size)
{
    uint8_t* mem = static_cast<uint8_t*>(lt_alloc("Longtail_B200_CreateVersionIndex", sizeof(struct Longtail_VersionIndex) + size));
    if (!mem) return nullptr;
    struct Longtail_VersionIndex* v = reinterpret_cast<struct Longtail_VersionIndex*>(mem);
    uint8_t* p = mem + sizeof(struct Longtail_VersionIndex);
    memcpy(p, data, size);
    uint8_t* start = p;
    v->m_Version = reinterpret_cast<uint32_t*>(p); p += 4;
    v->m_HashIdentifier = reinterpret_cast<uint32_t*>(p); p += 4;
    v->m_TargetChunkSize = reinterpret_cast<uint32_t*>(p); p += 4;
    v->m_AssetCount = reinterpret_cast<uint32_t*>(p); p += 4;
    v->m_ChunkCount = reinterpret_cast<uint32_t*>(p); p += 4;
    v->m_AssetChunkIndexCount = reinterpret_cast<uint32_t*>(p); p += 4;
    const size_t A = *v->m_AssetCount, C = *v->m_ChunkCount, I = *v->m_AssetChunkIndexCount;
    v->m_PathHashes = reinterpret_cast<TLongtail_Hash*>(p); p += 8 * A;
    v->m_ContentHashes = reinterpret_cast<TLongtail_Hash*>(p); p += 8 * A;
    v->m_AssetSizes = reinterpret_cast<uint64_t*>(p); p += 8 * A;
    v->m_AssetChunkCounts = reinterpret_cast<uint32_t*>(p); p += 4 * A;
    v->m_AssetChunkIndexStarts = reinterpret_cast<uint32_t*>(p); p += 4 * A;
    v->m_AssetChunkIndexes = reinterpret_cast<uint32_t*>(p); p += 4 * I;
    v->m_ChunkHashes = reinterpret_cast<TLongtail_Hash*>(p); p += 8 * C;
    v->m_ChunkSizes = reinterpret_cast<uint32_t*>(p); p += 4 * C;
    v->m_ChunkTags = reinterpret_cast<uint32_t*>(p); p += 4 * C;
    v->m_NameOffsets = reinterpret_cast<uint32_t*>(p); p += 4 * A;
    v->m_Permissions = reinterpret_cast<uint16_t*>(p); p += 2 * A;
    v->m_NameDataSize = (uint32_t)(size - (uint64_t)(p - start));
    v->m_NameData = reinterpret_cast<char*>(p);
    return v;
}

// ---------------------------------------------------------------- LZ4 CompressionAPI
const uint32_t TYPE_LZ4 = LT_B200_COMPRESSION_LZ4;

struct B200CompressionAPI
{
    struct Longtail_CompressionAPI api;
};

size_t lz4_max_size(struct Longtail_CompressionAPI*, uint32_t, size_t size) { return (size_t)lt_b200_lz4_bound(size); }

int lz4_compress(struct Longtail_CompressionAPI*, uint32_t, const char* uncompressed, char* compressed, size_t uncompressed_size,
                 size_t max_compressed_size, size_t* out_compressed_size)
{
    if (!compressed || !out_compressed_size || (!uncompressed && uncompressed_size) || uncompressed_size > 0x7E000000u) return EINVAL;
    std::lock_guard<std::mutex> g(g_gpu);
    int err = ensure_ctx();
    if (err) return err;
    const void* src = uncompressed;
    void* dst = compressed;
    uint32_t n = (uint32_t)uncompressed_size;
    uint64_t cap = max_compressed_size, out = 0;
    err = lt_b200_lz4_compress_host(g_ctx, 1, &src, &n, &dst, &cap, &out);
    if (err) return err;
    *out_compressed_size = (size_t)out;
    return 0;
}

int lz4_decompress(struct Longtail_CompressionAPI*, const char* compressed, char* uncompressed, size_t compressed_size,
                   size_t max_uncompressed_size, size_t* out_uncompressed_size)
{
    if (!compressed || !uncompressed || !out_uncompressed_size || compressed_size > 0x7fffffffu) return EINVAL;
    std::lock_guard<std::mutex> g(g_gpu);
    int err = ensure_ctx();
    if (err) return err;
    const void* src = compressed;
    void* dst = uncompressed;
    uint32_t n = (uint32_t)compressed_size;
    uint64_t cap = max_uncompressed_size > 0xfffffff0u ? 0xfffffff0u : max_uncompressed_size, out = 0;
    err = lt_b200_lz4_decompress_host(g_ctx, 1, &src, &n, &dst, &cap, &out);
    if (err) return err;
    *out_uncompressed_size = (size_t)out;
    return 0;
}

void compression_api_dispose(struct Longtail_API* api) { lt_free(api); }

// ---------------------------------------------------------------- ZStd CompressionAPI ('ztd1' / 'ztd2' = level 3; lib/zstd/longtail_zstd.c)
bool is_zstd_level3(uint32_t type) { return type == LT_B200_COMPRESSION_ZSTD_DEFAULT || type == LT_B200_COMPRESSION_ZSTD_MIN; }
bool is_zstd_type(uint32_t type) { return (type & 0xffffff00u) == 0x7a746400u; } // lib/zstd/longtail_zstd.c:33

size_t zstd_max_size(struct Longtail_CompressionAPI*, uint32_t, size_t size) { return (size_t)lt_b200_zstd_bound(size); } // :83-86

int zstd_compress(struct Longtail_CompressionAPI*, uint32_t settings_id, const char* uncompressed, char* compressed, size_t uncompressed_size,
                  size_t max_compressed_size, size_t* out_compressed_size)
{
    if (!compressed || !out_compressed_size || (!uncompressed && uncompressed_size) || uncompressed_size > 0x7E000000u) return EINVAL;
    if (!is_zstd_level3(settings_id)) return ENOTSUP; // levels 8 / 22 ('ztd3', 'ztd4', 'ztd5') have no device encoder
    std::lock_guard<std::mutex> g(g_gpu);
    int err = ensure_ctx();
    if (err) return err;
    const void* src = uncompressed;
    void* dst = compressed;
    uint32_t n = (uint32_t)uncompressed_size;
    uint64_t cap = max_compressed_size, out = 0;
    err = lt_b200_zstd_compress_host(g_ctx, settings_id, 1, &src, &n, &dst, &cap, &out);
    if (err) return err;
    *out_compressed_size = (size_t)out;
    return 0;
}

// ZStdCompressionAPI_Decompress (lib/zstd/longtail_zstd.c:143-176): frames of every level decode on the device (zstd_dec.cu)
int zstd_decompress(struct Longtail_CompressionAPI*, const char* compressed, char* uncompressed, size_t compressed_size,
                    size_t max_uncompressed_size, size_t* out_uncompressed_size)
{
    if (!compressed || !uncompressed || !out_uncompressed_size || compressed_size > 0x7fffffffu) return EINVAL;
    std::lock_guard<std::mutex> g(g_gpu);
    int err = ensure_ctx();
    if (err) return err;
    const void* src = compressed;
    void* dst = uncompressed;
    uint32_t n = (uint32_t)compressed_size;
    uint64_t cap = max_uncompressed_size > 0xfffffff0u ? 0xfffffff0u : max_uncompressed_size, out = 0;
    err = lt_b200_zstd_decompress_host(g_ctx, 1, &src, &n, &dst, &cap, &out);
    if (err) return err == EBADF ? EINVAL : err; // the reference maps every ZSTD error to EINVAL (:168-172)
    *out_uncompressed_size = (size_t)out;
    return 0;
}

// ---------------------------------------------------------------- compress block store
size_t block_index_data_size(uint32_t chunk_count) { return 8 + 4 + 4 + 4 + 12 * (size_t)chunk_count; } // src/longtail.c:3585-3597

// StoredBlock + BlockIndex + index data + payload in one allocation (the shape CompressBlock / DecompressBlock build,
// lib/compressblockstore/longtail_compressblockstore.c:103-137, :300-312)
int owned_block_dispose(struct Longtail_StoredBlock* b)
{
    lt_free(b);
    return 0;
}

struct Longtail_StoredBlock* make_owned_block(const struct Longtail_BlockIndex* index, size_t payload_capacity)
{
    const uint32_t n = *index->m_ChunkCount;
    const size_t ids = block_index_data_size(n);
    uint8_t* mem = static_cast<uint8_t*>(lt_alloc("B200CompressBlockStore", sizeof(struct Longtail_StoredBlock) + sizeof(struct Longtail_BlockIndex) + ids + payload_capacity));
    if (!mem) return nullptr;
    struct Longtail_StoredBlock* b = reinterpret_cast<struct Longtail_StoredBlock*>(mem);
    struct Longtail_BlockIndex* bi = reinterpret_cast<struct Longtail_BlockIndex*>(mem + sizeof(struct Longtail_StoredBlock));
    uint8_t* p = reinterpret_cast<uint8_t*>(bi) + sizeof(struct Longtail_BlockIndex);
    memcpy(p, index->m_BlockHash, ids); // the index data is contiguous from m_BlockHash on (Longtail_InitBlockIndex, src/longtail.c:3599-3627)
    bi->m_BlockHash = reinterpret_cast<TLongtail_Hash*>(p);
    bi->m_HashIdentifier = reinterpret_cast<uint32_t*>(p + 8);
    bi->m_ChunkCount = reinterpret_cast<uint32_t*>(p + 12);
    bi->m_Tag = reinterpret_cast<uint32_t*>(p + 16);
    bi->m_ChunkHashes = reinterpret_cast<TLongtail_Hash*>(p + 20);
    bi->m_ChunkSizes = reinterpret_cast<uint32_t*>(p + 20 + 8 * (size_t)n);
    b->Dispose = owned_block_dispose;
    b->m_BlockIndex = bi;
    b->m_BlockData = p + ids;
    b->m_BlockChunksDataSize = 0;
    return b;
}

struct B200CompressStore;

struct PutRequest
{
    struct Longtail_AsyncPutStoredBlockAPI api; // handed to the backing store; first member so the pointer converts back
    B200CompressStore* store;
    struct Longtail_StoredBlock* compressed;
    struct Longtail_AsyncPutStoredBlockAPI* caller;
};

struct GetRequest
{
    struct Longtail_AsyncGetStoredBlockAPI api;
    B200CompressStore* store;
    struct Longtail_AsyncGetStoredBlockAPI* caller;
};

struct PendingPut
{
    struct Longtail_StoredBlock* block;
    struct Longtail_AsyncPutStoredBlockAPI* caller;
};

struct B200CompressStore
{
    struct Longtail_BlockStoreAPI api;
    struct Longtail_BlockStoreAPI* backing;
    std::atomic<uint64_t> stats[Longtail_BlockStoreAPI_StatU64_Count];
    std::mutex lock;
    std::condition_variable wake;
    std::vector<PendingPut> queue;
    std::vector<struct Longtail_AsyncFlushAPI*> flushers;
    int pending = 0; // requests not yet completed (guarded by lock)
    bool stop = false;
    std::thread worker;
};

void store_complete_request(B200CompressStore* s)
{
    std::vector<struct Longtail_AsyncFlushAPI*> fire;
    {
        std::lock_guard<std::mutex> g(s->lock);
        if (--s->pending == 0) fire.swap(s->flushers);
    }
    for (auto* f : fire) f->OnComplete(f, 0); // longtail_compressblockstore.c:30-50
}

void put_backing_complete(struct Longtail_AsyncPutStoredBlockAPI* api, int err)
{
    PutRequest* r = reinterpret_cast<PutRequest*>(api);
    B200CompressStore* s = r->store;
    if (err) s->stats[Longtail_BlockStoreAPI_StatU64_PutStoredBlock_FailCount]++;
    if (r->compressed) r->compressed->Dispose(r->compressed);
    r->caller->OnComplete(r->caller, err);
    lt_free(r);
    store_complete_request(s);
}

// forwards one (already compressed or pass-through) block to the backing store; on a synchronous error the caller's
// completion is fired with it, because PutStoredBlock itself has long returned 0 for queued blocks
void forward_put(B200CompressStore* s, struct Longtail_StoredBlock* to_store, struct Longtail_StoredBlock* owned,
                 struct Longtail_AsyncPutStoredBlockAPI* caller)
{
    PutRequest* r = static_cast<PutRequest*>(lt_alloc("B200CompressBlockStore", sizeof(PutRequest)));
    if (!r)
    {
        if (owned) owned->Dispose(owned);
        s->stats[Longtail_BlockStoreAPI_StatU64_PutStoredBlock_FailCount]++;
        caller->OnComplete(caller, ENOMEM);
        store_complete_request(s);
        return;
    }
    r->api.m_API.Dispose = 0;
    r->api.OnComplete = put_backing_complete;
    r->store = s;
    r->compressed = owned;
    r->caller = caller;
    int err = s->backing->PutStoredBlock(s->backing, to_store, &r->api);
    if (err) put_backing_complete(&r->api, err);
}

void store_worker(B200CompressStore* s)
{
    for (;;)
    {
        std::vector<PendingPut> batch;
        {
            std::unique_lock<std::mutex> g(s->lock);
            s->wake.wait(g, [&] { return s->stop || !s->queue.empty(); });
            if (s->queue.empty() && s->stop) return;
            // give concurrently running WriteContentBlockJobs a moment to queue their blocks too: one launch serves them all
            g.unlock();
            std::this_thread::sleep_for(std::chrono::microseconds(300));
            g.lock();
            batch.swap(s->queue);
        }
        const uint32_t n = (uint32_t)batch.size();
        std::vector<struct Longtail_StoredBlock*> out(n, nullptr);
        std::vector<const void*> src(n);
        std::vector<void*> dst(n);
        std::vector<uint32_t> src_size(n);
        std::vector<uint64_t> cap(n), got(n, 0);
        int err = 0;
        for (uint32_t i = 0; i < n && !err; ++i)
        {
            const uint32_t raw = batch[i].block->m_BlockChunksDataSize;
            cap[i] = *batch[i].block->m_BlockIndex->m_Tag == TYPE_LZ4 ? lt_b200_lz4_bound(raw) : lt_b200_zstd_bound(raw);
            out[i] = make_owned_block(batch[i].block->m_BlockIndex, 8 + (size_t)cap[i]);
            if (!out[i]) { err = ENOMEM; break; }
            src[i] = batch[i].block->m_BlockData;
            src_size[i] = raw;
            dst[i] = static_cast<uint8_t*>(out[i]->m_BlockData) + 8;
        }
        if (!err)
        {
            std::lock_guard<std::mutex> g(g_gpu);
            err = ensure_ctx();
            // one launch per codec present in the batch
            for (int codec = 0; codec < 2 && !err; ++codec)
            {
                std::vector<uint32_t> idx;
                for (uint32_t i = 0; i < n; ++i)
                    if ((*batch[i].block->m_BlockIndex->m_Tag == TYPE_LZ4) == (codec == 0)) idx.push_back(i);
                if (idx.empty()) continue;
                const uint32_t m = (uint32_t)idx.size();
                std::vector<const void*> s2(m);
                std::vector<void*> d2(m);
                std::vector<uint32_t> z2(m);
                std::vector<uint64_t> c2(m), g2(m, 0);
                for (uint32_t k = 0; k < m; ++k) { s2[k] = src[idx[k]]; d2[k] = dst[idx[k]]; z2[k] = src_size[idx[k]]; c2[k] = cap[idx[k]]; }
                err = codec == 0 ? lt_b200_lz4_compress_host(g_ctx, m, s2.data(), z2.data(), d2.data(), c2.data(), g2.data())
                                 : lt_b200_zstd_compress_host(g_ctx, LT_B200_COMPRESSION_ZSTD_DEFAULT, m, s2.data(), z2.data(), d2.data(), c2.data(), g2.data());
                for (uint32_t k = 0; k < m; ++k) got[idx[k]] = g2[k];
            }
        }
        for (uint32_t i = 0; i < n; ++i)
        {
            if (err)
            {
                if (out[i]) out[i]->Dispose(out[i]);
                s->stats[Longtail_BlockStoreAPI_StatU64_PutStoredBlock_FailCount]++;
                batch[i].caller->OnComplete(batch[i].caller, err);
                store_complete_request(s);
                continue;
            }
            uint32_t* header = static_cast<uint32_t*>(out[i]->m_BlockData); // longtail_compressblockstore.c:135-137
            header[0] = src_size[i];
            header[1] = (uint32_t)got[i];
            out[i]->m_BlockChunksDataSize = 8 + (uint32_t)got[i];
            forward_put(s, out[i], out[i], batch[i].caller);
        }
    }
}

int store_put(struct Longtail_BlockStoreAPI* api, struct Longtail_StoredBlock* block, struct Longtail_AsyncPutStoredBlockAPI* async)
{
    B200CompressStore* s = reinterpret_cast<B200CompressStore*>(api);
    if (!block || !async) return EINVAL;
    const uint32_t n = *block->m_BlockIndex->m_ChunkCount;
    s->stats[Longtail_BlockStoreAPI_StatU64_PutStoredBlock_Count]++;
    s->stats[Longtail_BlockStoreAPI_StatU64_PutStoredBlock_Chunk_Count] += n;
    s->stats[Longtail_BlockStoreAPI_StatU64_PutStoredBlock_Byte_Count] += block_index_data_size(n) + block->m_BlockChunksDataSize;
    const uint32_t tag = *block->m_BlockIndex->m_Tag;
    if (tag != 0 && tag != TYPE_LZ4 && !is_zstd_level3(tag))
    {
        s->stats[Longtail_BlockStoreAPI_StatU64_PutStoredBlock_FailCount]++;
        return ENOTSUP; // no device kernel for this codec; a non-zero return means OnComplete is not called (src/longtail.c:4747-4757)
    }
    {
        std::lock_guard<std::mutex> g(s->lock);
        ++s->pending;
        if (tag != 0)
        {
            s->queue.push_back({block, async});
            s->wake.notify_one();
            return 0;
        }
    }
    forward_put(s, block, nullptr, async); // tag 0: stored as is (longtail_compressblockstore.c:84-88)
    return 0;
}

void get_backing_complete(struct Longtail_AsyncGetStoredBlockAPI* api, struct Longtail_StoredBlock* block, int err)
{
    GetRequest* r = reinterpret_cast<GetRequest*>(api);
    B200CompressStore* s = r->store;
    struct Longtail_AsyncGetStoredBlockAPI* caller = r->caller;
    lt_free(r);
    if (err)
    {
        if (err != ENOENT) s->stats[Longtail_BlockStoreAPI_StatU64_GetStoredBlock_FailCount]++;
        caller->OnComplete(caller, block, err);
        store_complete_request(s);
        return;
    }
    const uint32_t n = *block->m_BlockIndex->m_ChunkCount;
    s->stats[Longtail_BlockStoreAPI_StatU64_GetStoredBlock_Chunk_Count] += n;
    s->stats[Longtail_BlockStoreAPI_StatU64_GetStoredBlock_Byte_Count] += block_index_data_size(n) + block->m_BlockChunksDataSize;
    const uint32_t tag = *block->m_BlockIndex->m_Tag;
    if (tag == 0)
    {
        caller->OnComplete(caller, block, 0);
        store_complete_request(s);
        return;
    }
    struct Longtail_StoredBlock* plain = nullptr;
    const bool known = tag == TYPE_LZ4 || is_zstd_type(tag);
    if (!known || block->m_BlockChunksDataSize < 8)
        err = !known ? ENOTSUP : EBADF;
    else
    {
        const uint32_t* header = static_cast<const uint32_t*>(block->m_BlockData); // longtail_compressblockstore.c:292-296
        const uint32_t raw = header[0], comp = header[1];
        if ((uint64_t)comp + 8 > block->m_BlockChunksDataSize) err = EBADF;
        if (!err)
        {
            plain = make_owned_block(block->m_BlockIndex, raw);
            if (!plain) err = ENOMEM;
        }
        if (!err)
        {
            size_t got = 0;
            err = (tag == TYPE_LZ4 ? lz4_decompress : zstd_decompress)(nullptr, reinterpret_cast<const char*>(header + 2),
                                                                       static_cast<char*>(plain->m_BlockData), comp, raw, &got);
            if (!err && got != raw) err = EBADF; // :323-327
            plain->m_BlockChunksDataSize = raw;
        }
    }
    if (err)
    {
        s->stats[Longtail_BlockStoreAPI_StatU64_GetStoredBlock_FailCount]++;
        if (plain) plain->Dispose(plain);
        if (block && block->Dispose) block->Dispose(block);
        caller->OnComplete(caller, 0, err);
    }
    else
    {
        if (block->Dispose) block->Dispose(block);
        caller->OnComplete(caller, plain, 0);
    }
    store_complete_request(s);
}

int store_get(struct Longtail_BlockStoreAPI* api, uint64_t block_hash, struct Longtail_AsyncGetStoredBlockAPI* async)
{
    B200CompressStore* s = reinterpret_cast<B200CompressStore*>(api);
    if (!async) return EINVAL;
    s->stats[Longtail_BlockStoreAPI_StatU64_GetStoredBlock_Count]++;
    GetRequest* r = static_cast<GetRequest*>(lt_alloc("B200CompressBlockStore", sizeof(GetRequest)));
    if (!r) return ENOMEM;
    r->api.m_API.Dispose = 0;
    r->api.OnComplete = get_backing_complete;
    r->store = s;
    r->caller = async;
    {
        std::lock_guard<std::mutex> g(s->lock);
        ++s->pending;
    }
    int err = s->backing->GetStoredBlock(s->backing, block_hash, &r->api);
    if (err)
    {
        if (err != ENOENT) s->stats[Longtail_BlockStoreAPI_StatU64_GetStoredBlock_FailCount]++;
        lt_free(r);
        store_complete_request(s);
    }
    return err;
}

int store_preflight(struct Longtail_BlockStoreAPI* api, uint32_t count, const TLongtail_Hash* hashes, struct Longtail_AsyncPreflightStartedAPI* async)
{
    B200CompressStore* s = reinterpret_cast<B200CompressStore*>(api);
    s->stats[Longtail_BlockStoreAPI_StatU64_PreflightGet_Count]++;
    int err = s->backing->PreflightGet(s->backing, count, hashes, async);
    if (err) s->stats[Longtail_BlockStoreAPI_StatU64_PreflightGet_FailCount]++;
    return err;
}

int store_existing(struct Longtail_BlockStoreAPI* api, uint32_t count, const TLongtail_Hash* hashes, uint32_t min_usage,
                   struct Longtail_AsyncGetExistingContentAPI* async)
{
    B200CompressStore* s = reinterpret_cast<B200CompressStore*>(api);
    s->stats[Longtail_BlockStoreAPI_StatU64_GetExistingContent_Count]++;
    int err = s->backing->GetExistingContent(s->backing, count, hashes, min_usage, async);
    if (err) s->stats[Longtail_BlockStoreAPI_StatU64_GetExistingContent_FailCount]++;
    return err;
}

int store_prune(struct Longtail_BlockStoreAPI*, uint32_t, const TLongtail_Hash*, struct Longtail_AsyncPruneBlocksAPI*) { return ENOTSUP; }

int store_stats(struct Longtail_BlockStoreAPI* api, struct Longtail_BlockStore_Stats* out)
{
    B200CompressStore* s = reinterpret_cast<B200CompressStore*>(api);
    if (!out) return EINVAL;
    s->stats[Longtail_BlockStoreAPI_StatU64_GetStats_Count]++;
    for (int i = 0; i < Longtail_BlockStoreAPI_StatU64_Count; ++i) out->m_StatU64[i] = s->stats[i].load();
    return 0;
}

int store_flush(struct Longtail_BlockStoreAPI* api, struct Longtail_AsyncFlushAPI* async)
{
    B200CompressStore* s = reinterpret_cast<B200CompressStore*>(api);
    if (!async) return EINVAL;
    s->stats[Longtail_BlockStoreAPI_StatU64_Flush_Count]++;
    {
        std::lock_guard<std::mutex> g(s->lock);
        if (s->pending > 0)
        {
            s->flushers.push_back(async); // fired by the request that brings the count to zero
            return 0;
        }
    }
    async->OnComplete(async, 0);
    return 0;
}

void store_dispose(struct Longtail_API* api)
{
    B200CompressStore* s = reinterpret_cast<B200CompressStore*>(api);
    for (;;) // like the reference, wait for requests in flight (longtail_compressblockstore.c:564-571)
    {
        {
            std::lock_guard<std::mutex> g(s->lock);
            if (s->pending == 0) break;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    {
        std::lock_guard<std::mutex> g(s->lock);
        s->stop = true;
        s->wake.notify_all();
    }
    s->worker.join();
    s->~B200CompressStore();
    lt_free(s);
}

} // namespace

// ---------------------------------------------------------------- WriteContent + compress block store in one verb
namespace {

struct PutWait
{
    struct Longtail_AsyncPutStoredBlockAPI api;
    std::mutex m;
    std::condition_variable cv;
    bool done = false;
    int err = 0;
};
void put_wait_done(struct Longtail_AsyncPutStoredBlockAPI* api, int err)
{
    PutWait* w = reinterpret_cast<PutWait*>(api);
    std::lock_guard<std::mutex> g(w->m);
    w->err = err;
    w->done = true;
    w->cv.notify_all();
}

bool is_b200_compress_store(const struct Longtail_BlockStoreAPI* api) { return api && api->PutStoredBlock == store_put; }

// the monitor table of the B200 verbs (Longtail_B200_SetMonitor): a copy, read without a lock by the verb that fires the events
struct Longtail_Monitor g_monitor;
bool g_monitor_on = false;

struct WriteSink
{
    struct Longtail_BlockStoreAPI* store;
    struct Longtail_ProgressAPI* progress;
    struct Longtail_CancelAPI* cancel;
    Longtail_CancelAPI_HCancelToken token;
    uint32_t total, done;
    const TLongtail_Hash* expected_hashes; // the store index's block hashes, in block order
    const struct Longtail_StoreIndex* store_index;
    uint32_t composed; // blocks whose BlockCompose event has been fired
};

// lt_b200_block_sink: the byte image becomes a Longtail_StoredBlock whose index points into it (Longtail_InitStoredBlockFromData,
// src/longtail.c:4056-4109); the image is only valid during the call, so the put is awaited here
int write_content_sink(void* user, const struct lt_b200_stored_block_view* v)
{
    WriteSink* ws = static_cast<WriteSink*>(user);
    if (ws->cancel && ws->cancel->IsCancelled(ws->cancel, ws->token) == ECANCELED) return ECANCELED;
    // the block is stored under the hash of ITS chunk hashes; a store index that lists another hash for it does not describe these blocks
    if (ws->expected_hashes && ws->done < ws->total && ws->expected_hashes[ws->done] != v->block_hash) return EINVAL;
    uint8_t* p = static_cast<uint8_t*>(const_cast<void*>(v->data));
    const uint32_t n = v->chunk_count;
    struct Longtail_BlockIndex bi;
    bi.m_BlockHash = reinterpret_cast<TLongtail_Hash*>(p);
    bi.m_HashIdentifier = reinterpret_cast<uint32_t*>(p + 8);
    bi.m_ChunkCount = reinterpret_cast<uint32_t*>(p + 12);
    bi.m_Tag = reinterpret_cast<uint32_t*>(p + 16);
    bi.m_ChunkHashes = reinterpret_cast<TLongtail_Hash*>(p + 20);
    bi.m_ChunkSizes = reinterpret_cast<uint32_t*>(p + 20 + 8 * (size_t)n);
    struct Longtail_StoredBlock sb;
    sb.Dispose = nullptr;
    sb.m_BlockIndex = &bi;
    sb.m_BlockData = p + block_index_data_size(n);
    sb.m_BlockChunksDataSize = (uint32_t)(v->size - block_index_data_size(n));
    PutWait w;
    w.api.m_API.Dispose = nullptr;
    w.api.OnComplete = put_wait_done;
    const uint32_t block_index = ws->done;
    if (g_monitor_on)
    {
        // the device composes a whole batch at once: a block counts as composed when the first block of its batch comes out
        if (g_monitor.BlockCompose && ws->composed <= block_index) g_monitor.BlockCompose(ws->store_index, ws->composed++);
        if (g_monitor.BlockSave) g_monitor.BlockSave(ws->store_index, block_index, v->size);
    }
    int err = ws->store->PutStoredBlock(ws->store, &sb, &w.api);
    if (err) // a non-zero return means OnComplete is not called (src/longtail.c:4747-4757)
    {
        if (g_monitor_on && g_monitor.BlockSaved) g_monitor.BlockSaved(ws->store_index, block_index, err);
        return err;
    }
    {
        std::unique_lock<std::mutex> g(w.m);
        w.cv.wait(g, [&w] { return w.done; });
        err = w.err;
    }
    if (g_monitor_on && g_monitor.BlockSaved) g_monitor.BlockSaved(ws->store_index, block_index, err);
    ++ws->done;
    if (!err && ws->progress) ws->progress->OnProgress(ws->progress, ws->total, ws->done);
    return err;
}

} // namespace

extern "C" int Longtail_B200_WriteContent(struct Longtail_StorageAPI* source_storage_api, struct Longtail_BlockStoreAPI* backing_block_store_api,
                                          struct Longtail_JobAPI* job_api, struct Longtail_ProgressAPI* progress_api,
                                          struct Longtail_CancelAPI* optional_cancel_api, Longtail_CancelAPI_HCancelToken optional_cancel_token,
                                          struct Longtail_StoreIndex* store_index, struct Longtail_VersionIndex* version_index, const char* assets_folder)
{
    // same argument validation as src/longtail.c:4781-4786
    if (!source_storage_api || !backing_block_store_api || !job_api || !version_index || !store_index || !assets_folder) return EINVAL;
    // this verb compresses the blocks itself (their tags say how): the store it writes to is the one BELOW the compress layer.  Chained onto
    // the B200 compress store the payloads would be compressed twice — refuse that outright
    if (is_b200_compress_store(backing_block_store_api)) return EINVAL;
    const uint32_t block_count = *store_index->m_BlockCount;
    if (block_count == 0) return 0; // :4788-4792
    const uint32_t hash_type = *version_index->m_HashIdentifier;
    const uint32_t A = *version_index->m_AssetCount, C = *version_index->m_ChunkCount, SC = *store_index->m_ChunkCount;

    // CreateAssetPartLookup (:4429-4500): a chunk's bytes come from its first occurrence, assets and chunks in version order
    std::unordered_map<uint64_t, std::pair<uint32_t, uint64_t>> where; // chunk hash -> (asset, offset inside the asset)
    where.reserve((size_t)C * 2);
    for (uint32_t a = 0; a < A; ++a)
    {
        uint64_t off = 0;
        const uint32_t first = version_index->m_AssetChunkIndexStarts[a];
        for (uint32_t k = 0; k < version_index->m_AssetChunkCounts[a]; ++k)
        {
            const uint32_t ci = version_index->m_AssetChunkIndexes[first + k];
            where.emplace(version_index->m_ChunkHashes[ci], std::make_pair(a, off));
            off += version_index->m_ChunkSizes[ci];
        }
    }
    std::unordered_map<uint64_t, uint32_t> version_chunk; // :4821-4824
    version_chunk.reserve((size_t)C * 2);
    for (uint32_t c = 0; c < C; ++c) version_chunk[version_index->m_ChunkHashes[c]] = c;

    // the chunks in store order with the sizes the VERSION gives them (:4826-4834), the tag of their block, and the assets they need
    std::vector<uint64_t> hashes(SC), offsets(SC);
    std::vector<uint32_t> sizes(SC), tags(SC), counts(block_count);
    std::vector<uint8_t> needed(A ? A : 1, 0);
    // A block is addressed through m_BlockChunksOffsets like WriteContentBlockJob does (:4596-4604); the chunk arrays handed to the device
    // verb are compacted into block order, so a store index whose offsets are not the running sum of the counts (a subset, a pruned or a
    // malformed index) is either composed correctly or refused, never read out of bounds
    uint32_t pos = 0;
    for (uint32_t b = 0; b < block_count; ++b)
    {
        counts[b] = store_index->m_BlockChunkCounts[b];
        if (counts[b] == 0) return EINVAL;
        const uint32_t first = store_index->m_BlockChunksOffsets[b];
        if (first > SC || counts[b] > SC - first || counts[b] > SC - pos) return EINVAL;
        const uint32_t tag = store_index->m_BlockTags[b];
        if (tag != 0 && tag != TYPE_LZ4 && !is_zstd_level3(tag)) return ENOTSUP;
        for (uint32_t k = 0; k < counts[b]; ++k, ++pos)
        {
            const uint64_t h = store_index->m_ChunkHashes[first + k];
            auto vc = version_chunk.find(h);
            auto w = where.find(h);
            if (vc == version_chunk.end() || w == where.end()) return EINVAL;
            hashes[pos] = h;
            sizes[pos] = version_index->m_ChunkSizes[vc->second];
            tags[pos] = tag;
            needed[w->second.first] = 1;
        }
    }
    const uint32_t written_chunks = pos;
    // arena layout: only the assets that hold a needed chunk
    std::vector<uint64_t> arena_off(A ? A : 1, 0);
    uint64_t total = 0;
    for (uint32_t a = 0; a < A; ++a)
        if (needed[a])
        {
            arena_off[a] = total;
            total += (version_index->m_AssetSizes[a] + 255u) & ~(uint64_t)255u;
        }
    for (uint32_t c = 0; c < written_chunks; ++c)
    {
        const auto& w = where[hashes[c]];
        offsets[c] = arena_off[w.first] + w.second;
    }

    std::lock_guard<std::mutex> g(g_verb);
    int err = ensure_verb_ctx();
    if (err) return err;
    const uint64_t arena_bytes = total + 4096;
    void* arena = nullptr;
    err = lt_b200_device_alloc(g_verb_ctx, arena_bytes, &arena);
    if (err) return err;
    // the reader of the index verb: StorageAPI reads fanned out over the caller's JobAPI, paths from the version index
    struct Longtail_FileInfos names;
    memset(&names, 0, sizeof(names));
    names.m_Count = A;
    names.m_PathStartOffsets = version_index->m_NameOffsets;
    names.m_PathData = version_index->m_NameData;
    ReadCtx rc = {source_storage_api, assets_folder, &names, job_api, nullptr, optional_cancel_api, optional_cancel_token, 0, 0};
    const uint64_t stage_bytes = 256ull << 20, piece = 8ull << 20;
    void* stage = nullptr;
    err = lt_b200_host_alloc_pinned(g_verb_ctx, stage_bytes, &stage);
    std::vector<lt_b200_read_job> jobs;
    std::vector<uint64_t> dst_off;
    uint64_t used = 0;
    auto flush = [&]() -> int {
        if (jobs.empty()) return 0;
        int e = read_batch(&rc, jobs.data(), (uint32_t)jobs.size());
        for (size_t i = 0; i < jobs.size() && !e; ++i)
            e = lt_b200_copy_to_device(g_verb_ctx, static_cast<uint8_t*>(arena) + dst_off[i], jobs[i].dst, jobs[i].size);
        jobs.clear();
        dst_off.clear();
        used = 0;
        return e;
    };
    for (uint32_t a = 0; a < A && !err; ++a)
    {
        if (!needed[a]) continue;
        if (g_monitor_on && g_monitor.AssetOpen) g_monitor.AssetOpen(version_index, a, 0);
        if (g_monitor_on && g_monitor.AssetClose) g_monitor.AssetClose(version_index, a);
        for (uint64_t o = 0; o < version_index->m_AssetSizes[a] && !err; o += piece)
        {
            const uint32_t n = (uint32_t)std::min<uint64_t>(piece, version_index->m_AssetSizes[a] - o);
            if (used + n > stage_bytes) err = flush();
            jobs.push_back({a, n, o, static_cast<uint8_t*>(stage) + used});
            dst_off.push_back(arena_off[a] + o);
            used += n;
        }
    }
    if (!err) err = flush();
    if (!err)
    {
        WriteSink ws = {backing_block_store_api, progress_api, optional_cancel_api, optional_cancel_token, block_count, 0, store_index->m_BlockHashes, store_index, 0};
        err = lt_b200_write_given_blocks_device(g_verb_ctx, static_cast<const uint8_t*>(arena), arena_bytes, written_chunks, hashes.data(), sizes.data(), tags.data(),
                                                offsets.data(), hash_type, block_count, counts.data(), write_content_sink, &ws);
    }
    if (stage) lt_b200_host_free_pinned(g_verb_ctx, stage);
    lt_b200_device_free(g_verb_ctx, arena);
    return err;
}

extern "C" struct Longtail_CompressionAPI* Longtail_CreateB200LZ4CompressionAPI(void)
{
    B200CompressionAPI* a = static_cast<B200CompressionAPI*>(lt_alloc("Longtail_CreateB200LZ4CompressionAPI", sizeof(B200CompressionAPI)));
    if (!a) return nullptr;
    a->api.m_API.Dispose = compression_api_dispose;
    a->api.GetMaxCompressedSize = lz4_max_size;
    a->api.Compress = lz4_compress;
    a->api.Decompress = lz4_decompress;
    return &a->api;
}

extern "C" struct Longtail_CompressionAPI* Longtail_CompressionRegistry_CreateForB200LZ4(uint32_t compression_type, uint32_t* out_settings)
{
    if (compression_type != TYPE_LZ4) return nullptr; // lib/lz4/longtail_lz4.c:104-118
    if (out_settings) *out_settings = TYPE_LZ4;
    return Longtail_CreateB200LZ4CompressionAPI();
}

extern "C" struct Longtail_CompressionAPI* Longtail_CreateB200ZStdCompressionAPI(void)
{
    B200CompressionAPI* a = static_cast<B200CompressionAPI*>(lt_alloc("Longtail_CreateB200ZStdCompressionAPI", sizeof(B200CompressionAPI)));
    if (!a) return nullptr;
    a->api.m_API.Dispose = compression_api_dispose;
    a->api.GetMaxCompressedSize = zstd_max_size;
    a->api.Compress = zstd_compress;
    a->api.Decompress = zstd_decompress;
    return &a->api;
}

extern "C" struct Longtail_CompressionAPI* Longtail_CompressionRegistry_CreateForB200ZStd(uint32_t compression_type, uint32_t* out_settings)
{
    if (!is_zstd_type(compression_type)) return nullptr; // lib/zstd/longtail_zstd.c:31-42: every 'ztd?' id maps to the one ZStd API
    if (out_settings) *out_settings = compression_type;
    return Longtail_CreateB200ZStdCompressionAPI();
}

extern "C" struct Longtail_BlockStoreAPI* Longtail_CreateB200CompressBlockStoreAPI(struct Longtail_BlockStoreAPI* backing,
                                                                                  struct Longtail_CompressionRegistryAPI* registry)
{
    (void)registry;
    if (!backing) return nullptr;
    void* mem = lt_alloc("Longtail_CreateB200CompressBlockStoreAPI", sizeof(B200CompressStore));
    if (!mem) return nullptr;
    B200CompressStore* s = new (mem) B200CompressStore();
    s->api.m_API.Dispose = store_dispose;
    s->api.PutStoredBlock = store_put;
    s->api.PreflightGet = store_preflight;
    s->api.GetStoredBlock = store_get;
    s->api.GetExistingContent = store_existing;
    s->api.PruneBlocks = store_prune;
    s->api.GetStats = store_stats;
    s->api.Flush = store_flush;
    s->backing = backing;
    for (auto& v : s->stats) v = 0;
    s->worker = std::thread(store_worker, s);
    return &s->api;
}

extern "C" void Longtail_B200_SetMonitor(const struct Longtail_Monitor* monitor)
{
    std::lock_guard<std::mutex> g(g_verb);
    g_monitor_on = monitor != nullptr;
    if (monitor) g_monitor = *monitor;
}

extern "C" int Longtail_B200_SetDevice(int device_ordinal)
{
    std::lock_guard<std::mutex> g(g_gpu);
    std::lock_guard<std::mutex> gv(g_verb);
    if (g_ctx || g_verb_ctx) return EBUSY;
    g_device = device_ordinal;
    return 0;
}

extern "C" struct Longtail_ChunkerAPI* Longtail_CreateB200ChunkerAPI(void)
{
    B200ChunkerAPI* a = static_cast<B200ChunkerAPI*>(lt_alloc("Longtail_CreateB200ChunkerAPI", sizeof(B200ChunkerAPI)));
    if (!a) return nullptr;
    a->api.m_API.Dispose = chunker_api_dispose;
    a->api.GetMinChunkSize = chunker_get_min;
    a->api.CreateChunker = chunker_create;
    a->api.NextChunk = chunker_next;
    a->api.DisposeChunker = chunker_dispose;
    a->api.NextChunkFromBuffer = chunker_next_from_buffer;
    return &a->api;
}

static struct Longtail_HashAPI* create_hash_api(uint32_t type);
extern "C" struct Longtail_HashAPI* Longtail_CreateB200Blake3HashAPI(void) { return create_hash_api(LT_B200_HASH_BLAKE3); }
extern "C" struct Longtail_HashAPI* Longtail_CreateB200Blake2HashAPI(void) { return create_hash_api(LT_B200_HASH_BLAKE2); }
extern "C" struct Longtail_HashAPI* Longtail_CreateB200MeowHashAPI(void) { return create_hash_api(LT_B200_HASH_MEOW); }

static struct Longtail_HashAPI* create_hash_api(uint32_t type)
{
    B200HashAPI* a = static_cast<B200HashAPI*>(lt_alloc("Longtail_CreateB200HashAPI", sizeof(B200HashAPI)));
    if (!a) return nullptr;
    a->type = type;
    a->api.m_API.Dispose = hash_api_dispose;
    a->api.GetIdentifier = hash_identifier;
    a->api.BeginContext = hash_begin;
    a->api.Hash = hash_update;
    a->api.EndContext = hash_end;
    a->api.HashBuffer = hash_buffer;
    return &a->api;
}

extern "C" int Longtail_B200_CreateVersionIndex(struct Longtail_StorageAPI* storage_api, struct Longtail_HashAPI* hash_api,
                                                struct Longtail_ChunkerAPI* chunker_api, struct Longtail_JobAPI* job_api,
                                                struct Longtail_ProgressAPI* progress_api, struct Longtail_CancelAPI* optional_cancel_api,
                                                Longtail_CancelAPI_HCancelToken optional_cancel_token, const char* root_path,
                                                const struct Longtail_FileInfos* file_infos, const uint32_t* optional_asset_tags,
                                                uint32_t target_chunk_size, int enable_file_map, struct Longtail_VersionIndex** out_version_index)
{
    (void)enable_file_map; // ignored by the reference as well (src/longtail.c:2453)
    // same argument validation as src/longtail.c:2826-2836
    if (!storage_api || !hash_api || !chunker_api || !root_path || !out_version_index || target_chunk_size == 0) return EINVAL;
    if (file_infos && file_infos->m_Count && !job_api) return EINVAL;
    const uint32_t hash_type = hash_api->GetIdentifier(hash_api);
    if (hash_type != LT_B200_HASH_BLAKE3 && hash_type != LT_B200_HASH_BLAKE2 && hash_type != LT_B200_HASH_MEOW) return ENOTSUP;
    uint32_t min_chunk = 0;
    int err = chunker_api->GetMinChunkSize(chunker_api, &min_chunk);
    if (err) return err;
    if (min_chunk != 48) return ENOTSUP;

    lt_b200_assets assets;
    memset(&assets, 0, sizeof(assets));
    if (file_infos)
    {
        assets.asset_count = file_infos->m_Count;
        assets.path_data_size = file_infos->m_PathDataSize;
        assets.sizes = file_infos->m_Sizes;
        assets.path_start_offsets = file_infos->m_PathStartOffsets;
        assets.permissions = file_infos->m_Permissions;
        assets.path_data = file_infos->m_PathData;
    }
    ReadCtx rc = {storage_api, root_path, file_infos, job_api, progress_api, optional_cancel_api, optional_cancel_token, 0, 0};
    const uint64_t part = (uint64_t)target_chunk_size * 1024;
    for (uint32_t i = 0; i < assets.asset_count; ++i)
        rc.total_jobs += (uint32_t)((assets.sizes[i] + part - 1) / part); // non-empty parts only

    std::lock_guard<std::mutex> g(g_verb);
    err = ensure_verb_ctx();
    if (err) return err;
    const void* data = nullptr;
    uint64_t size = 0;
    err = lt_b200_index_stream_assets(g_verb_ctx, &assets, optional_asset_tags, hash_type, target_chunk_size, read_batch, &rc, &data, &size);
    if (err) return err;
    struct Longtail_VersionIndex* v = wrap_version_index(data, size);
    if (!v) return ENOMEM;
    *out_version_index = v;
    return 0;
}
