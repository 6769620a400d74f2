// longtail_api.cpp — drop-in Longtail_*API objects over the C ABI (include/longtail_b200_api.h).
//
// Host code only.  It mirrors the reference's object conventions: callback structs whose first member is Longtail_API
// (src/longtail.h:43-46), errno returns, ESPIPE at end of stream, Dispose through the struct.  All data-path work is done
// by the kernels behind lt_b200_*; nothing here hashes or scans a byte on the CPU.
#include "../../include/longtail_b200_api.h"

#include <dlfcn.h>
#include <errno.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
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
};

struct HashStream
{
    std::vector<uint8_t> bytes;
};

uint32_t hash_identifier(struct Longtail_HashAPI*) { return HASH_BLK3; }

int hash_on_device(uint32_t length, const void* data, uint64_t* out)
{
    std::lock_guard<std::mutex> g(g_gpu);
    int err = ensure_ctx();
    if (!err) err = ensure_arena((uint64_t)length + 64);
    if (!err && length) err = lt_b200_copy_to_device(g_ctx, g_arena, data, length);
    if (err) return err;
    uint64_t off = 0;
    return lt_b200_hash_segments(g_ctx, HASH_BLK3, static_cast<const uint8_t*>(g_arena), g_arena_cap, &off, &length, 1, out);
}

int hash_buffer(struct Longtail_HashAPI*, uint32_t length, const void* data, uint64_t* out)
{
    if (!out || (!data && length)) return EINVAL;
    const uint8_t* p = static_cast<const uint8_t*>(data);
    {
        // a range handed out by one of our chunkers: its hash was computed in the chunker's GPU pass
        std::lock_guard<std::mutex> g(g_registry_lock);
        for (B200Chunker* c : g_live)
        {
            if (!c->buf || p < c->buf || p >= c->buf + c->size) continue;
            const uint64_t off = (uint64_t)(p - c->buf);
            auto it = std::lower_bound(c->chunks.begin(), c->chunks.end(), off, [](const ChunkRec& r, uint64_t o) { return r.offset < o; });
            if (it != c->chunks.end() && it->offset == off && it->len == length)
            {
                *out = it->hash;
                return 0;
            }
            break;
        }
    }
    return hash_on_device(length, data, out);
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

uint64_t hash_end(struct Longtail_HashAPI*, Longtail_HashAPI_HContext h)
{
    HashStream* s = reinterpret_cast<HashStream*>(h);
    uint64_t out = 0;
    hash_on_device((uint32_t)s->bytes.size(), s->bytes.data(), &out); // EndContext frees the context (lib/blake3/longtail_blake3.c:60-79)
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

} // namespace

extern "C" int Longtail_B200_SetDevice(int device_ordinal)
{
    std::lock_guard<std::mutex> g(g_gpu);
    if (g_ctx) return EBUSY;
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

extern "C" struct Longtail_HashAPI* Longtail_CreateB200Blake3HashAPI(void)
{
    B200HashAPI* a = static_cast<B200HashAPI*>(lt_alloc("Longtail_CreateB200Blake3HashAPI", sizeof(B200HashAPI)));
    if (!a) return nullptr;
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
    if (hash_api->GetIdentifier(hash_api) != HASH_BLK3) return ENOTSUP;
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

    std::lock_guard<std::mutex> g(g_gpu);
    err = ensure_ctx();
    if (err) return err;
    const void* data = nullptr;
    uint64_t size = 0;
    err = lt_b200_index_stream_assets(g_ctx, &assets, optional_asset_tags, HASH_BLK3, target_chunk_size, read_batch, &rc, &data, &size);
    if (err) return err;
    struct Longtail_VersionIndex* v = wrap_version_index(data, size);
    if (!v) return ENOMEM;
    *out_version_index = v;
    return 0;
}
