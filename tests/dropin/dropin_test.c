/* dropin_test.c — the B200 backends inside the UNMODIFIED reference pipeline.
 *
 * Built by oracle/Makefile (target `dropin`) against the reference's own src/longtail.h and liblongtail_ref.a, linked with
 * liblongtail_b200.so.  Needs a GPU to run.  Checks, on an in-memory tree written through the reference's memstorage:
 *   0. include/longtail_abi.h has the same struct layout as src/longtail.h (abi_table.c compiled both ways)
 *   1. reference Longtail_CreateVersionIndex with the reference backends                      -> baseline bytes
 *   2. reference Longtail_CreateVersionIndex with Longtail_CreateB200ChunkerAPI + B200 HashAPI -> identical bytes
 *   3. Longtail_B200_CreateVersionIndex (same parameter list)                                 -> identical bytes
 *   4. the reference's ChunkerLargeFile golden vector (test/test.cpp:3363-3465) through the B200 ChunkerAPI
 *   5. the compress half inside the reference's WriteContent; Longtail_B200_WriteContent
 *   6. fault injection: failing Read, a store that fills up, a cancel token flipped mid-verb, malformed store indexes — the B200 verbs
 *      return the reference's errno (test/test.cpp:4733, :4839, :5752 are the reference's own doubles for these)
 */
#define LONGTAIL_B200_USE_LONGTAIL_H
#include "longtail.h"
#include "longtail_b200_api.h"
#include "../lib/bikeshed/longtail_bikeshed.h"
#include "../lib/blake3/longtail_blake3.h"
#include "../lib/blake2/longtail_blake2.h"
#include "../lib/meowhash/longtail_meowhash.h"
#include "../lib/hpcdcchunker/longtail_hpcdcchunker.h"
#include "../lib/memstorage/longtail_memstorage.h"
#include "../lib/compressblockstore/longtail_compressblockstore.h"
#include "../lib/compressionregistry/longtail_compression_registry.h"
#include "../lib/compressionregistry/longtail_full_compression_registry.h"
#include "../lib/lz4/longtail_lz4.h"
#include <pthread.h>

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct abi_entry { const char* name; unsigned long value; };
extern const struct abi_entry abi_table_b200[];
extern const struct abi_entry abi_table_reference[];

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { ++failures; printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)

static uint64_t mix(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static void fill(uint8_t* p, size_t n, uint64_t seed, int low_entropy)
{
    for (size_t i = 0; i < n; ++i)
    {
        uint8_t b = (uint8_t)(mix(seed * 0x9E3779B97F4A7C15ull + i / 8) >> (8 * (i % 8)));
        p[i] = low_entropy ? (b & 0x0f) : b;
    }
}

static int write_file(struct Longtail_StorageAPI* st, const char* path, const uint8_t* data, size_t n)
{
    char* parent = st->GetParentPath(st, path);
    if (parent) { /* create parents bottom-up */
        char tmp[512]; size_t len = strlen(parent);
        for (size_t i = 1; i <= len; ++i) if (parent[i] == '/' || parent[i] == 0) { memcpy(tmp, parent, i); tmp[i] = 0; st->CreateDir(st, tmp); }
        Longtail_Free(parent);
    }
    Longtail_StorageAPI_HOpenFile f;
    int err = st->OpenWriteFile(st, path, n, &f);
    if (err) return err;
    if (n) err = st->Write(st, f, 0, n, data);
    st->CloseFile(st, f);
    return err;
}

static int serialise(struct Longtail_VersionIndex* v, void** buf, size_t* size) { return Longtail_WriteVersionIndexToBuffer(v, buf, size); }

struct feed { const uint8_t* data; uint64_t size, off; };
static int feed_func(void* ctx, Longtail_ChunkerAPI_HChunker c, uint32_t requested, char* buffer, uint32_t* out)
{
    (void)c;
    struct feed* f = (struct feed*)ctx;
    uint64_t n = f->size - f->off;
    if (n > requested) n = requested;
    memcpy(buffer, f->data + f->off, n);
    f->off += n;
    *out = (uint32_t)n;
    return 0;
}

/* ---- a backing block store that keeps every serialised block in memory (Put) and hands them back (Get) */
struct kept_block { uint64_t hash; void* data; size_t size; };
struct keep_store
{
    struct Longtail_BlockStoreAPI api;
    pthread_mutex_t lock;
    struct kept_block* blocks;
    uint32_t count, capacity;
};
static int keep_put(struct Longtail_BlockStoreAPI* api, struct Longtail_StoredBlock* b, struct Longtail_AsyncPutStoredBlockAPI* async)
{
    struct keep_store* s = (struct keep_store*)api;
    void* buf = 0; size_t size = 0;
    int err = Longtail_WriteStoredBlockToBuffer(b, &buf, &size);
    if (!err)
    {
        pthread_mutex_lock(&s->lock);
        if (s->count == s->capacity)
        {
            s->capacity = s->capacity ? s->capacity * 2 : 4096;
            s->blocks = (struct kept_block*)realloc(s->blocks, sizeof(struct kept_block) * s->capacity);
        }
        s->blocks[s->count].hash = *b->m_BlockIndex->m_BlockHash;
        s->blocks[s->count].data = buf;
        s->blocks[s->count].size = size;
        ++s->count;
        pthread_mutex_unlock(&s->lock);
    }
    async->OnComplete(async, err);
    return 0;
}
static int keep_get(struct Longtail_BlockStoreAPI* api, uint64_t hash, struct Longtail_AsyncGetStoredBlockAPI* async)
{
    struct keep_store* s = (struct keep_store*)api;
    for (uint32_t i = 0; i < s->count; ++i)
        if (s->blocks[i].hash == hash)
        {
            struct Longtail_StoredBlock* b = 0;
            int err = Longtail_ReadStoredBlockFromBuffer(s->blocks[i].data, s->blocks[i].size, &b);
            async->OnComplete(async, b, err);
            return 0;
        }
    return ENOENT;
}
static int keep_flush(struct Longtail_BlockStoreAPI* api, struct Longtail_AsyncFlushAPI* async) { (void)api; async->OnComplete(async, 0); return 0; }
static void keep_dispose(struct Longtail_API* api)
{
    struct keep_store* s = (struct keep_store*)api;
    for (uint32_t i = 0; i < s->count; ++i) Longtail_Free(s->blocks[i].data);
    free(s->blocks);
    free(s);
}
static struct keep_store* make_keep_store(void)
{
    struct keep_store* s = (struct keep_store*)calloc(1, sizeof(*s));
    s->api.m_API.Dispose = keep_dispose;
    s->api.PutStoredBlock = keep_put;
    s->api.GetStoredBlock = keep_get;
    s->api.Flush = keep_flush;
    pthread_mutex_init(&s->lock, 0);
    return s;
}
static const struct kept_block* find_block(const struct keep_store* s, uint64_t hash)
{
    for (uint32_t i = 0; i < s->count; ++i) if (s->blocks[i].hash == hash) return &s->blocks[i];
    return 0;
}
struct get_wait { struct Longtail_AsyncGetStoredBlockAPI api; struct Longtail_StoredBlock* block; int err; int done; };
static void get_done(struct Longtail_AsyncGetStoredBlockAPI* a, struct Longtail_StoredBlock* b, int err)
{
    struct get_wait* w = (struct get_wait*)a;
    w->block = b; w->err = err; w->done = 1;
}

/* upsync of a fresh store through `store`; the blocks land in whatever backing store it wraps */
static int write_content(struct Longtail_StorageAPI* storage, struct Longtail_BlockStoreAPI* store, struct Longtail_JobAPI* jobs,
                         struct Longtail_HashAPI* hash, struct Longtail_VersionIndex* vi, uint32_t block_size, uint32_t chunks_per_block,
                         struct Longtail_StoreIndex** out_missing)
{
    struct Longtail_StoreIndex* empty = 0;
    int err = Longtail_CreateStoreIndexFromBlocks(0, 0, &empty);
    if (!err) err = Longtail_CreateMissingContent(hash, empty, vi, block_size, chunks_per_block, out_missing);
    if (!err) err = Longtail_WriteContent(storage, store, jobs, 0, 0, 0, *out_missing, vi, "root");
    Longtail_Free(empty);
    return err;
}


/* ---------------------------------------------------------------- fault injection through the boundary
 * The reference's own doubles for this are a storage that runs out of space (test/test.cpp:5752), a cancel token flipped in the middle of an
 * operation (:4733, :4839) and a failing store; each fault goes through the reference verb AND the B200 verb, and both must answer with
 * the same errno, return (no hang) and leave the library usable for the next call. */
struct flaky_storage
{
    struct Longtail_StorageAPI api;
    struct Longtail_StorageAPI* inner;
    volatile long reads;     /* Read calls seen so far */
    long fail_at;            /* the call that fails (counted from 1); 0 = never */
    int error;
};
static void flaky_dispose(struct Longtail_API* api) { (void)api; }
static int flaky_open(struct Longtail_StorageAPI* a, const char* path, Longtail_StorageAPI_HOpenFile* out)
{ struct flaky_storage* f = (struct flaky_storage*)a; return f->inner->OpenReadFile(f->inner, path, out); }
static int flaky_size(struct Longtail_StorageAPI* a, Longtail_StorageAPI_HOpenFile h, uint64_t* out)
{ struct flaky_storage* f = (struct flaky_storage*)a; return f->inner->GetSize(f->inner, h, out); }
static int flaky_read(struct Longtail_StorageAPI* a, Longtail_StorageAPI_HOpenFile h, uint64_t offset, uint64_t length, void* output)
{
    struct flaky_storage* f = (struct flaky_storage*)a;
    const long n = __sync_add_and_fetch(&f->reads, 1);
    if (f->fail_at && n >= f->fail_at) return f->error;
    return f->inner->Read(f->inner, h, offset, length, output);
}
static void flaky_close(struct Longtail_StorageAPI* a, Longtail_StorageAPI_HOpenFile h) { struct flaky_storage* f = (struct flaky_storage*)a; f->inner->CloseFile(f->inner, h); }
static char* flaky_concat(struct Longtail_StorageAPI* a, const char* r, const char* s) { struct flaky_storage* f = (struct flaky_storage*)a; return f->inner->ConcatPath(f->inner, r, s); }
static void make_flaky(struct flaky_storage* f, struct Longtail_StorageAPI* inner, long fail_at, int error)
{
    memset(f, 0, sizeof(*f));
    f->api.m_API.Dispose = flaky_dispose;
    f->api.OpenReadFile = flaky_open;
    f->api.GetSize = flaky_size;
    f->api.Read = flaky_read;
    f->api.CloseFile = flaky_close;
    f->api.ConcatPath = flaky_concat;
    f->inner = inner;
    f->fail_at = fail_at;
    f->error = error;
}

/* a block store that accepts `good` blocks and then fails: by its return value (OnComplete is not called then, src/longtail.c:4747-4757)
 * or through OnComplete */
struct full_store
{
    struct Longtail_BlockStoreAPI api;
    volatile long puts;
    long good;
    int error, through_callback;
};
static void full_dispose(struct Longtail_API* api) { (void)api; }
static int full_put(struct Longtail_BlockStoreAPI* api, struct Longtail_StoredBlock* b, struct Longtail_AsyncPutStoredBlockAPI* async)
{
    struct full_store* s = (struct full_store*)api;
    (void)b;
    const long n = __sync_add_and_fetch(&s->puts, 1);
    if (n > s->good)
    {
        if (!s->through_callback) return s->error;
        async->OnComplete(async, s->error);
        return 0;
    }
    async->OnComplete(async, 0);
    return 0;
}
static int full_flush(struct Longtail_BlockStoreAPI* api, struct Longtail_AsyncFlushAPI* async) { (void)api; async->OnComplete(async, 0); return 0; }
static void make_full(struct full_store* s, long good, int error, int through_callback)
{
    memset(s, 0, sizeof(*s));
    s->api.m_API.Dispose = full_dispose;
    s->api.PutStoredBlock = full_put;
    s->api.Flush = full_flush;
    s->good = good;
    s->error = error;
    s->through_callback = through_callback;
}

/* a cancel API whose token reads "cancelled" after `after` polls */
struct late_cancel
{
    struct Longtail_CancelAPI api;
    volatile long polls;
    long after;
};
static void cancel_dispose(struct Longtail_API* api) { (void)api; }
static int cancel_create(struct Longtail_CancelAPI* a, Longtail_CancelAPI_HCancelToken* out) { *out = (Longtail_CancelAPI_HCancelToken)a; return 0; }
static int cancel_cancel(struct Longtail_CancelAPI* a, Longtail_CancelAPI_HCancelToken t) { (void)a; (void)t; return 0; }
static int cancel_is(struct Longtail_CancelAPI* a, Longtail_CancelAPI_HCancelToken t)
{
    struct late_cancel* c = (struct late_cancel*)a;
    (void)t;
    return __sync_add_and_fetch(&c->polls, 1) > c->after ? ECANCELED : 0;
}
static int cancel_dispose_token(struct Longtail_CancelAPI* a, Longtail_CancelAPI_HCancelToken t) { (void)a; (void)t; return 0; }
static void make_cancel(struct late_cancel* c, long after)
{
    memset(c, 0, sizeof(*c));
    c->api.m_API.Dispose = cancel_dispose;
    c->api.CreateToken = cancel_create;
    c->api.Cancel = cancel_cancel;
    c->api.IsCancelled = cancel_is;
    c->api.DisposeToken = cancel_dispose_token;
    c->after = after;
}

static volatile long mon_compose, mon_save, mon_saved, mon_saved_err, mon_open, mon_close;
static void mon_block_compose(const struct Longtail_StoreIndex* s, uint32_t b) { (void)s; (void)b; __sync_add_and_fetch(&mon_compose, 1); }
static void mon_block_save(const struct Longtail_StoreIndex* s, uint32_t b, uint64_t size) { (void)s; (void)b; (void)size; __sync_add_and_fetch(&mon_save, 1); }
static void mon_block_saved(const struct Longtail_StoreIndex* s, uint32_t b, int err) { (void)s; (void)b; __sync_add_and_fetch(&mon_saved, 1); if (err) __sync_add_and_fetch(&mon_saved_err, 1); }
static void mon_asset_open(const struct Longtail_VersionIndex* v, uint32_t a, int err) { (void)v; (void)a; (void)err; __sync_add_and_fetch(&mon_open, 1); }
static void mon_asset_close(const struct Longtail_VersionIndex* v, uint32_t a) { (void)v; (void)a; __sync_add_and_fetch(&mon_close, 1); }

static void fault_injection(struct Longtail_StorageAPI* storage, struct Longtail_JobAPI* jobs, struct Longtail_HashAPI* ref_hash,
                            struct Longtail_ChunkerAPI* ref_chunker, struct Longtail_FileInfos* infos, uint32_t* tags, uint32_t target,
                            const void* b_ref, size_t n_ref)
{
    const int before = failures;
    /* (a) a Read that fails in the middle of CreateVersionIndex */
    for (int k = 0; k < 3; ++k)
    {
        const long fail_at = k == 0 ? 1 : k == 1 ? 7 : 40;
        struct flaky_storage f_ref, f_b200;
        make_flaky(&f_ref, storage, fail_at, EIO);
        make_flaky(&f_b200, storage, fail_at, EIO);
        struct Longtail_VersionIndex *v1 = 0, *v2 = 0;
        const int e1 = Longtail_CreateVersionIndex(&f_ref.api, ref_hash, ref_chunker, jobs, 0, 0, 0, "root", infos, tags, target, 0, &v1);
        const int e2 = Longtail_B200_CreateVersionIndex(&f_b200.api, ref_hash, ref_chunker, jobs, 0, 0, 0, "root", infos, tags, target, 0, &v2);
        /* the reference answers EIO when the first job to fail reads a file of at most 48 bytes directly (src/longtail.c:2076) and ESPIPE when it
         * fails inside the chunker, which swallows the feeder's errno (longtail_hpcdcchunker.c:243-250 -> :414-423) — which one is latched depends
         * on the worker schedule.  The B200 verb reports the storage's own errno. */
        CHECK((e1 == EIO || e1 == ESPIPE) && e2 == EIO, "failing Read #%ld: reference %d, B200 verb %d", fail_at, e1, e2);
        CHECK(v1 == 0 && v2 == 0, "no index may come out of a failed call");
        Longtail_Free(v1); Longtail_Free(v2);
    }
    /* (b) a cancel token that flips while the verb runs */
    for (int k = 0; k < 2; ++k)
    {
        struct late_cancel c_ref, c_b200;
        make_cancel(&c_ref, k ? 5 : 0);
        make_cancel(&c_b200, k ? 5 : 0);
        struct Longtail_VersionIndex *v1 = 0, *v2 = 0;
        const int e1 = Longtail_CreateVersionIndex(storage, ref_hash, ref_chunker, jobs, 0, &c_ref.api, (Longtail_CancelAPI_HCancelToken)&c_ref, "root", infos, tags, target, 0, &v1);
        const int e2 = Longtail_B200_CreateVersionIndex(storage, ref_hash, ref_chunker, jobs, 0, &c_b200.api, (Longtail_CancelAPI_HCancelToken)&c_b200, "root", infos, tags, target, 0, &v2);
        CHECK(e1 == ECANCELED && e2 == e1, "cancel after %d polls in CreateVersionIndex: reference %d, B200 verb %d", k ? 5 : 0, e1, e2);
        Longtail_Free(v1); Longtail_Free(v2);
    }
    /* the library is still usable and still right (the tags have changed since section 1: compare with a fresh reference index) */
    {
        struct Longtail_VersionIndex *v = 0, *w = 0;
        const int e = Longtail_B200_CreateVersionIndex(storage, ref_hash, ref_chunker, jobs, 0, 0, 0, "root", infos, tags, target, 0, &v);
        CHECK(Longtail_CreateVersionIndex(storage, ref_hash, ref_chunker, jobs, 0, 0, 0, "root", infos, tags, target, 0, &w) == 0, "reference index");
        void *b = 0, *bw = 0; size_t n = 0, nw = 0;
        if (!e) serialise(v, &b, &n);
        serialise(w, &bw, &nw);
        CHECK(e == 0 && n == nw && b && memcmp(b, bw, n) == 0, "CreateVersionIndex after the injected faults: %d", e);
        Longtail_Free(b); Longtail_Free(bw); Longtail_Free(v); Longtail_Free(w);
        (void)b_ref; (void)n_ref;
    }
    /* (c) WriteContent: a store that fills up (by return value and through OnComplete), a failing Read, a cancel */
    struct Longtail_VersionIndex* vi = 0;
    CHECK(Longtail_CreateVersionIndex(storage, ref_hash, ref_chunker, jobs, 0, 0, 0, "root", infos, tags, target, 0, &vi) == 0, "index for the write faults");
    struct Longtail_StoreIndex *empty = 0, *missing = 0;
    Longtail_CreateStoreIndexFromBlocks(0, 0, &empty);
    CHECK(Longtail_CreateMissingContent(ref_hash, empty, vi, target * 16, 64, &missing) == 0, "missing content");
    struct Longtail_CompressionRegistryAPI* full = Longtail_CreateFullCompressionRegistry();
    for (int k = 0; k < 4; ++k)
    {
        struct full_store s_ref_inner, s_b200;
        make_full(&s_ref_inner, k < 2 ? 2 : 0, ENOSPC, k & 1);
        make_full(&s_b200, k < 2 ? 2 : 0, ENOSPC, k & 1);
        struct Longtail_BlockStoreAPI* s_ref = Longtail_CreateCompressBlockStoreAPI(&s_ref_inner.api, full);
        const int e1 = Longtail_WriteContent(storage, s_ref, jobs, 0, 0, 0, missing, vi, "root");
        const int e2 = Longtail_B200_WriteContent(storage, &s_b200.api, jobs, 0, 0, 0, missing, vi, "root");
        CHECK(e1 == ENOSPC && e2 == e1, "store full after %d blocks (%s): reference %d, B200 verb %d", k < 2 ? 2 : 0, (k & 1) ? "OnComplete" : "return value", e1, e2);
        SAFE_DISPOSE_API(s_ref);
    }
    {
        struct flaky_storage f_ref, f_b200;
        make_flaky(&f_ref, storage, 9, EIO);
        make_flaky(&f_b200, storage, 9, EIO);
        struct keep_store* k1 = make_keep_store();
        struct keep_store* k2 = make_keep_store();
        struct Longtail_BlockStoreAPI* s_ref = Longtail_CreateCompressBlockStoreAPI(&k1->api, full);
        const int e1 = Longtail_WriteContent(&f_ref.api, s_ref, jobs, 0, 0, 0, missing, vi, "root");
        const int e2 = Longtail_B200_WriteContent(&f_b200.api, &k2->api, jobs, 0, 0, 0, missing, vi, "root");
        CHECK(e1 == EIO && e2 == e1, "failing Read in WriteContent: reference %d, B200 verb %d", e1, e2);
        SAFE_DISPOSE_API(s_ref);
        SAFE_DISPOSE_API(&k1->api); SAFE_DISPOSE_API(&k2->api);
    }
    {
        struct late_cancel c_ref, c_b200;
        make_cancel(&c_ref, 3);
        make_cancel(&c_b200, 3);
        struct keep_store* k1 = make_keep_store();
        struct keep_store* k2 = make_keep_store();
        struct Longtail_BlockStoreAPI* s_ref = Longtail_CreateCompressBlockStoreAPI(&k1->api, full);
        const int e1 = Longtail_WriteContent(storage, s_ref, jobs, 0, &c_ref.api, (Longtail_CancelAPI_HCancelToken)&c_ref, missing, vi, "root");
        const int e2 = Longtail_B200_WriteContent(storage, &k2->api, jobs, 0, &c_b200.api, (Longtail_CancelAPI_HCancelToken)&c_b200, missing, vi, "root");
        CHECK(e1 == ECANCELED && e2 == e1, "cancel in WriteContent: reference %d, B200 verb %d", e1, e2);
        SAFE_DISPOSE_API(s_ref);
        SAFE_DISPOSE_API(&k1->api); SAFE_DISPOSE_API(&k2->api);
    }
    /* (d) the B200 compress block store under the B200 write verb would compress twice: refused; a store index whose offsets point
     *     outside its chunk array: refused, not read */
    {
        struct keep_store* k = make_keep_store();
        struct Longtail_BlockStoreAPI* chained = Longtail_CreateB200CompressBlockStoreAPI(&k->api, full);
        CHECK(Longtail_B200_WriteContent(storage, chained, jobs, 0, 0, 0, missing, vi, "root") == EINVAL, "B200 compress store as backing store must be EINVAL");
        SAFE_DISPOSE_API(chained);
        const uint32_t saved = missing->m_BlockChunksOffsets[*missing->m_BlockCount - 1];
        missing->m_BlockChunksOffsets[*missing->m_BlockCount - 1] = *missing->m_ChunkCount;
        CHECK(Longtail_B200_WriteContent(storage, &k->api, jobs, 0, 0, 0, missing, vi, "root") == EINVAL, "block offsets outside the chunk array must be EINVAL");
        missing->m_BlockChunksOffsets[*missing->m_BlockCount - 1] = saved;
        const uint64_t saved_hash = missing->m_BlockHashes[0];
        missing->m_BlockHashes[0] ^= 1;
        CHECK(Longtail_B200_WriteContent(storage, &k->api, jobs, 0, 0, 0, missing, vi, "root") == EINVAL, "a block hash the chunks do not hash to must be EINVAL");
        missing->m_BlockHashes[0] = saved_hash;
        /* and the verb still works afterwards — with the monitor events of WriteContentBlockJob (src/longtail.c:4586-4753) switched on */
        struct Longtail_Monitor mon;
        memset(&mon, 0, sizeof(mon));
        mon.StructSize = sizeof(mon);
        mon.BlockCompose = mon_block_compose; mon.BlockSave = mon_block_save; mon.BlockSaved = mon_block_saved;
        mon.AssetOpen = mon_asset_open; mon.AssetClose = mon_asset_close;
        Longtail_B200_SetMonitor(&mon);
        struct keep_store* k_ok = make_keep_store();
        struct keep_store* k_want = make_keep_store();
        struct Longtail_BlockStoreAPI* s_ref = Longtail_CreateCompressBlockStoreAPI(&k_want->api, full);
        CHECK(Longtail_WriteContent(storage, s_ref, jobs, 0, 0, 0, missing, vi, "root") == 0, "reference WriteContent after the faults");
        CHECK(Longtail_B200_WriteContent(storage, &k_ok->api, jobs, 0, 0, 0, missing, vi, "root") == 0, "Longtail_B200_WriteContent after the faults");
        uint32_t same = 0;
        for (uint32_t i = 0; i < k_want->count; ++i)
        {
            const struct kept_block* c = find_block(k_ok, k_want->blocks[i].hash);
            if (c && c->size == k_want->blocks[i].size && memcmp(c->data, k_want->blocks[i].data, c->size) == 0) ++same;
        }
        CHECK(k_ok->count == k_want->count && same == k_want->count, "after the faults: %u of %u blocks identical", same, k_want->count);
        Longtail_B200_SetMonitor(0);
        CHECK(mon_compose == (long)k_want->count && mon_save == mon_compose && mon_saved == mon_compose && mon_saved_err == 0 && mon_open > 0 && mon_open == mon_close,
              "monitor events: %ld compose, %ld save, %ld saved (%ld errors), %ld open, %ld close for %u blocks", mon_compose, mon_save, mon_saved, mon_saved_err,
              mon_open, mon_close, k_want->count);
        SAFE_DISPOSE_API(s_ref);
        SAFE_DISPOSE_API(&k->api); SAFE_DISPOSE_API(&k_ok->api); SAFE_DISPOSE_API(&k_want->api);
    }
    SAFE_DISPOSE_API(full);
    Longtail_Free(missing); Longtail_Free(empty); Longtail_Free(vi);
    printf("fault injection (failing Read, full store, cancel, malformed store index, chained compress store): %s\n",
           failures == before ? "same errno as the reference everywhere" : "see failures");
}

int main(int argc, char** argv)
{
    /* 0. ABI layout */
    for (int i = 0; abi_table_reference[i].name; ++i)
    {
        CHECK(strcmp(abi_table_reference[i].name, abi_table_b200[i].name) == 0, "abi table order");
        CHECK(abi_table_reference[i].value == abi_table_b200[i].value, "%s: reference %lu, longtail_abi.h %lu", abi_table_reference[i].name,
              abi_table_reference[i].value, abi_table_b200[i].value);
    }
    printf("abi: %s\n", failures ? "MISMATCH" : "identical layout");
    if (argc > 1 && strcmp(argv[1], "--abi-only") == 0) return failures ? 1 : 0;

    const uint32_t target = argc > 2 ? (uint32_t)atoi(argv[2]) : 4096;
    struct Longtail_StorageAPI* storage = Longtail_CreateInMemStorageAPI();
    struct Longtail_JobAPI* jobs = Longtail_CreateBikeshedJobAPI(3, 0);
    const size_t part = (size_t)target * 1024;
    const struct { const char* path; size_t size; int low; } files[] = {
        {"root/a/empty.bin", 0, 0}, {"root/a/one.bin", 1, 0}, {"root/a/48.bin", 48, 0}, {"root/a/49.bin", 49, 0},
        {"root/b/exact.bin", part, 0}, {"root/b/two.bin", part + 12345, 0}, {"root/b/three.bin", 2 * part + 777, 1},
        {"root/c/dup1.bin", 300000, 0}, {"root/d/dup2.bin", 300000, 0}, {"root/e/low.bin", 500000, 1}, {"root/z.txt", 100000, 1}};
    for (size_t i = 0; i < sizeof(files) / sizeof(files[0]); ++i)
    {
        uint8_t* d = (uint8_t*)malloc(files[i].size ? files[i].size : 1);
        fill(d, files[i].size, strstr(files[i].path, "dup") ? 99 : 7 + i, files[i].low);
        CHECK(write_file(storage, files[i].path, d, files[i].size) == 0, "write %s", files[i].path);
        free(d);
    }
    /* many assets whose parts end in the middle of a scan tile, so that one GPU batch holds dozens of ragged part tails */
    for (uint32_t i = 0; i < (target < 256 ? 24u : 120u); ++i) /* tiny targets mean one GPU round trip per ~50-byte block in section 5 */
    {
        char path[64];
        snprintf(path, sizeof(path), "root/r%u/f%03u.bin", i % 5, i);
        const size_t size = 1 + (size_t)(((uint64_t)(i + 1) * 2654435761u) % (part + 100000));
        uint8_t* d = (uint8_t*)malloc(size);
        fill(d, size, 1000 + i, (int)(i % 3 == 0));
        CHECK(write_file(storage, path, d, size) == 0, "write %s", path);
        free(d);
    }
    struct Longtail_FileInfos* infos = 0;
    CHECK(Longtail_GetFilesRecursively2(storage, jobs, 0, 0, 0, "root", &infos) == 0, "GetFilesRecursively2");
    uint32_t* tags = (uint32_t*)malloc(sizeof(uint32_t) * infos->m_Count);
    for (uint32_t i = 0; i < infos->m_Count; ++i) tags[i] = (i % 2) ? 0x6c7a3432u : 0u;

    struct Longtail_HashAPI* ref_hash = Longtail_CreateBlake3HashAPI();
    struct Longtail_ChunkerAPI* ref_chunker = Longtail_CreateHPCDCChunkerAPI();
    struct Longtail_HashAPI* b200_hash = Longtail_CreateB200Blake3HashAPI();
    struct Longtail_ChunkerAPI* b200_chunker = Longtail_CreateB200ChunkerAPI();
    CHECK(b200_hash && b200_chunker, "B200 API objects");

    /* 1. baseline */
    struct Longtail_VersionIndex* v_ref = 0;
    CHECK(Longtail_CreateVersionIndex(storage, ref_hash, ref_chunker, jobs, 0, 0, 0, "root", infos, tags, target, 0, &v_ref) == 0, "reference CreateVersionIndex");
    void* b_ref = 0; size_t n_ref = 0;
    serialise(v_ref, &b_ref, &n_ref);

    /* 2. the reference core driving the B200 ChunkerAPI / HashAPI objects */
    struct Longtail_VersionIndex* v_obj = 0;
    int err = Longtail_CreateVersionIndex(storage, b200_hash, b200_chunker, jobs, 0, 0, 0, "root", infos, tags, target, 0, &v_obj);
    CHECK(err == 0, "reference core + B200 objects: %d", err);
    if (!err)
    {
        void* b = 0; size_t n = 0;
        serialise(v_obj, &b, &n);
        CHECK(n == n_ref && memcmp(b, b_ref, n) == 0, "VersionIndex through B200 API objects differs (%zu vs %zu bytes)", n, n_ref);
        printf("reference core + B200 ChunkerAPI/HashAPI: %zu bytes %s\n", n, (n == n_ref && memcmp(b, b_ref, n) == 0) ? "identical" : "DIFFERENT");
        Longtail_Free(b);
        Longtail_Free(v_obj);
    }

    /* 3. the batched verb with the reference's parameter list */
    struct Longtail_VersionIndex* v_verb = 0;
    err = Longtail_B200_CreateVersionIndex(storage, ref_hash, ref_chunker, jobs, 0, 0, 0, "root", infos, tags, target, 0, &v_verb);
    CHECK(err == 0, "Longtail_B200_CreateVersionIndex: %d", err);
    if (!err)
    {
        void* b = 0; size_t n = 0;
        serialise(v_verb, &b, &n);
        CHECK(n == n_ref && memcmp(b, b_ref, n) == 0, "VersionIndex of Longtail_B200_CreateVersionIndex differs (%zu vs %zu bytes)", n, n_ref);
        printf("Longtail_B200_CreateVersionIndex: %zu bytes %s\n", n, (n == n_ref && memcmp(b, b_ref, n) == 0) ? "identical" : "DIFFERENT");
        Longtail_Free(b);
        Longtail_Free(v_verb);
    }

    /* 3b. BLAKE2s and Meow: the reference core with the B200 chunker + the B200 'blk2' / 'meow' HashAPI, and the batched verb
     *     selected by the reference's own HashAPI of that identifier */
    for (int alg = 0; alg < 2; ++alg)
    {
        const char* name = alg ? "Meow" : "BLAKE2s";
        struct Longtail_HashAPI* ref_h = alg ? Longtail_CreateMeowHashAPI() : Longtail_CreateBlake2HashAPI();
        struct Longtail_HashAPI* b200_h = alg ? Longtail_CreateB200MeowHashAPI() : Longtail_CreateB200Blake2HashAPI();
        struct Longtail_VersionIndex *v0 = 0, *v1 = 0, *v2 = 0;
        CHECK(ref_h && b200_h && ref_h->GetIdentifier(ref_h) == b200_h->GetIdentifier(b200_h), "%s identifiers differ", name);
        CHECK(Longtail_CreateVersionIndex(storage, ref_h, ref_chunker, jobs, 0, 0, 0, "root", infos, tags, target, 0, &v0) == 0, "reference %s index", name);
        int e1 = Longtail_CreateVersionIndex(storage, b200_h, b200_chunker, jobs, 0, 0, 0, "root", infos, tags, target, 0, &v1);
        int e2 = Longtail_B200_CreateVersionIndex(storage, ref_h, ref_chunker, jobs, 0, 0, 0, "root", infos, tags, target, 0, &v2);
        CHECK(e1 == 0 && e2 == 0, "%s through B200: %d %d", name, e1, e2);
        void *b0 = 0, *b1 = 0, *b2 = 0; size_t n0 = 0, n1 = 0, n2 = 0;
        serialise(v0, &b0, &n0);
        if (!e1) serialise(v1, &b1, &n1);
        if (!e2) serialise(v2, &b2, &n2);
        CHECK(n0 == n1 && b1 && memcmp(b0, b1, n0) == 0, "%s VersionIndex through the B200 API objects differs", name);
        CHECK(n0 == n2 && b2 && memcmp(b0, b2, n0) == 0, "%s VersionIndex of Longtail_B200_CreateVersionIndex differs", name);
        printf("%s: %zu bytes, objects %s, verb %s\n", name, n0, (b1 && n0 == n1 && !memcmp(b0, b1, n0)) ? "identical" : "DIFFERENT",
               (b2 && n0 == n2 && !memcmp(b0, b2, n0)) ? "identical" : "DIFFERENT");
        Longtail_Free(b0); Longtail_Free(b1); Longtail_Free(b2);
        Longtail_Free(v0); Longtail_Free(v1); Longtail_Free(v2);
        SAFE_DISPOSE_API(ref_h); SAFE_DISPOSE_API(b200_h);
    }

    /* 5. the compress half inside the reference's Longtail_WriteContent: reference compressblockstore vs
     *    (a) the B200 compress block store and (b) the reference compressblockstore with the B200 CompressionAPI in its registry;
     *    pass 0 = LZ4 ('lz42'), pass 1 = ZStd level 3 ('ztd2' and 'ztd1') */
    for (int pass = 0; pass < 2; ++pass)
    {
        const char* codec_name = pass ? "zstd" : "lz4";
        for (uint32_t i = 0; i < infos->m_Count; ++i)
            tags[i] = pass == 0 ? ((i % 3) ? 0x6c7a3432u : 0u) : ((i % 3) == 1 ? 0x7a746432u : (i % 3) == 2 ? 0x7a746431u : 0u);
        struct Longtail_VersionIndex* vi = 0;
        CHECK(Longtail_CreateVersionIndex(storage, ref_hash, ref_chunker, jobs, 0, 0, 0, "root", infos, tags, target, 0, &vi) == 0, "index for upsync");
        const uint32_t block_size = target * 16, per_block = 64;
        struct Longtail_CompressionRegistryAPI* full = Longtail_CreateFullCompressionRegistry();
        Longtail_CompressionRegistry_CreateForTypeFunc b200_funcs[1] = {pass ? Longtail_CompressionRegistry_CreateForB200ZStd : Longtail_CompressionRegistry_CreateForB200LZ4};
        struct Longtail_CompressionRegistryAPI* b200_registry = Longtail_CreateDefaultCompressionRegistry(1, b200_funcs);
        struct keep_store* k_ref = make_keep_store();
        struct keep_store* k_b200 = make_keep_store();
        struct keep_store* k_codec = make_keep_store();
        struct Longtail_BlockStoreAPI* s_ref = Longtail_CreateCompressBlockStoreAPI(&k_ref->api, full);
        struct Longtail_BlockStoreAPI* s_b200 = Longtail_CreateB200CompressBlockStoreAPI(&k_b200->api, full);
        struct Longtail_BlockStoreAPI* s_codec = Longtail_CreateCompressBlockStoreAPI(&k_codec->api, b200_registry);
        struct Longtail_StoreIndex *m_ref = 0, *m_b200 = 0, *m_codec = 0;
        CHECK(write_content(storage, s_ref, jobs, ref_hash, vi, block_size, per_block, &m_ref) == 0, "reference WriteContent");
        int e1 = write_content(storage, s_b200, jobs, ref_hash, vi, block_size, per_block, &m_b200);
        CHECK(e1 == 0, "WriteContent through Longtail_CreateB200CompressBlockStoreAPI: %d", e1);
        int e2 = write_content(storage, s_codec, jobs, ref_hash, vi, block_size, per_block, &m_codec);
        CHECK(e2 == 0, "WriteContent through the B200 %s CompressionAPI: %d", codec_name, e2);
        CHECK(k_ref->count > 3 && k_ref->count == k_b200->count && k_ref->count == k_codec->count, "block counts %u %u %u", k_ref->count, k_b200->count, k_codec->count);
        /* (c) Longtail_B200_WriteContent: WriteContent and the compress block store in one verb, into a plain store, for the store index
         *     the reference computed; and for a store index with a DIFFERENT block composition (two blocks merged) to show that the
         *     blocks are written as given, not re-packed */
        struct keep_store* k_verb = make_keep_store();
        int e3 = Longtail_B200_WriteContent(storage, &k_verb->api, jobs, 0, 0, 0, m_ref, vi, "root");
        CHECK(e3 == 0, "Longtail_B200_WriteContent: %d", e3);
        CHECK(k_verb->count == k_ref->count, "Longtail_B200_WriteContent block count %u vs %u", k_verb->count, k_ref->count);
        uint32_t same_c = 0;
        for (uint32_t i = 0; i < k_ref->count; ++i)
        {
            const struct kept_block* c = find_block(k_verb, k_ref->blocks[i].hash);
            if (c && c->size == k_ref->blocks[i].size && memcmp(c->data, k_ref->blocks[i].data, c->size) == 0) ++same_c;
        }
        CHECK(same_c == k_ref->count, "Longtail_B200_WriteContent: %u of %u stored blocks identical", same_c, k_ref->count);
        SAFE_DISPOSE_API(&k_verb->api);
        {
            struct Longtail_StoreIndex* regrouped = 0;
            struct Longtail_StoreIndex* empty = 0;
            Longtail_CreateStoreIndexFromBlocks(0, 0, &empty);
            CHECK(Longtail_CreateMissingContent(ref_hash, empty, vi, block_size * 2, per_block * 2, &regrouped) == 0, "larger blocks");
            struct keep_store* k_r2 = make_keep_store();
            struct keep_store* k_v2 = make_keep_store();
            struct Longtail_BlockStoreAPI* s_r2 = Longtail_CreateCompressBlockStoreAPI(&k_r2->api, full);
            CHECK(Longtail_WriteContent(storage, s_r2, jobs, 0, 0, 0, regrouped, vi, "root") == 0, "reference WriteContent, larger blocks");
            int e4 = Longtail_B200_WriteContent(storage, &k_v2->api, jobs, 0, 0, 0, regrouped, vi, "root");
            CHECK(e4 == 0, "Longtail_B200_WriteContent, larger blocks: %d", e4);
            uint32_t same_d = 0;
            for (uint32_t i = 0; i < k_r2->count; ++i)
            {
                const struct kept_block* d = find_block(k_v2, k_r2->blocks[i].hash);
                if (d && d->size == k_r2->blocks[i].size && memcmp(d->data, k_r2->blocks[i].data, d->size) == 0) ++same_d;
            }
            CHECK(k_r2->count == k_v2->count && same_d == k_r2->count && k_r2->count < k_ref->count, "larger blocks: %u of %u identical (%u before)", same_d, k_r2->count, k_ref->count);
            printf("Longtail_B200_WriteContent (%s): %u + %u blocks identical to WriteContent -> compressblockstore\n", codec_name, same_c, same_d);
            /* a store index that names a chunk the version does not hold: EINVAL, like src/longtail.c:4826-4832 */
            uint64_t saved = regrouped->m_ChunkHashes[0];
            regrouped->m_ChunkHashes[0] = 0x1234567890abcdefull;
            struct keep_store* k_bad = make_keep_store();
            CHECK(Longtail_B200_WriteContent(storage, &k_bad->api, jobs, 0, 0, 0, regrouped, vi, "root") == EINVAL, "unknown chunk must be EINVAL");
            regrouped->m_ChunkHashes[0] = saved;
            SAFE_DISPOSE_API(&k_bad->api);
            SAFE_DISPOSE_API(s_r2);
            SAFE_DISPOSE_API(&k_r2->api); SAFE_DISPOSE_API(&k_v2->api);
            Longtail_Free(regrouped); Longtail_Free(empty);
        }
        uint32_t same_a = 0, same_b = 0, lz4_blocks = 0;
        for (uint32_t i = 0; i < k_ref->count; ++i)
        {
            const struct kept_block* a = find_block(k_b200, k_ref->blocks[i].hash);
            const struct kept_block* b = find_block(k_codec, k_ref->blocks[i].hash);
            if (a && a->size == k_ref->blocks[i].size && memcmp(a->data, k_ref->blocks[i].data, a->size) == 0) ++same_a;
            if (b && b->size == k_ref->blocks[i].size && memcmp(b->data, k_ref->blocks[i].data, b->size) == 0) ++same_b;
            if (((uint32_t*)k_ref->blocks[i].data)[4] != 0) ++lz4_blocks;
        }
        CHECK(same_a == k_ref->count, "B200 compress block store: %u of %u stored blocks identical", same_a, k_ref->count);
        CHECK(same_b == k_ref->count, "B200 %s CompressionAPI: %u of %u stored blocks identical", codec_name, same_b, k_ref->count);
        CHECK(lz4_blocks > 0 && lz4_blocks < k_ref->count, "mix of raw and compressed blocks expected (%u of %u)", lz4_blocks, k_ref->count);
        struct Longtail_BlockStore_Stats st_ref, st_b200;
        s_ref->GetStats(s_ref, &st_ref);
        s_b200->GetStats(s_b200, &st_b200);
        for (int i = Longtail_BlockStoreAPI_StatU64_PutStoredBlock_Count; i <= Longtail_BlockStoreAPI_StatU64_PutStoredBlock_Byte_Count; ++i)
            CHECK(st_ref.m_StatU64[i] == st_b200.m_StatU64[i], "stat %d: %llu vs %llu", i, (unsigned long long)st_ref.m_StatU64[i], (unsigned long long)st_b200.m_StatU64[i]);
        /* read every block back through both stores: same uncompressed payloads (LZ4 and ZStd decode kernels) */
        uint32_t round_trips = 0;
        for (uint32_t i = 0; i < k_ref->count; ++i)
        {
            struct get_wait w1, w2;
            memset(&w1, 0, sizeof(w1)); memset(&w2, 0, sizeof(w2));
            w1.api.OnComplete = get_done; w2.api.OnComplete = get_done;
            CHECK(s_ref->GetStoredBlock(s_ref, k_ref->blocks[i].hash, &w1.api) == 0 && w1.done && w1.err == 0, "reference GetStoredBlock");
            CHECK(s_b200->GetStoredBlock(s_b200, k_ref->blocks[i].hash, &w2.api) == 0 && w2.done && w2.err == 0, "B200 GetStoredBlock %d", w2.err);
            if (w1.block && w2.block && w1.block->m_BlockChunksDataSize == w2.block->m_BlockChunksDataSize &&
                memcmp(w1.block->m_BlockData, w2.block->m_BlockData, w1.block->m_BlockChunksDataSize) == 0) ++round_trips;
            if (w1.block) w1.block->Dispose(w1.block);
            if (w2.block) w2.block->Dispose(w2.block);
        }
        CHECK(round_trips == k_ref->count, "GetStoredBlock round trips: %u of %u", round_trips, k_ref->count);
        printf("WriteContent: %u blocks (%u %s), B200 block store %u identical, B200 codec %u identical, %u read back\n", k_ref->count, lz4_blocks, codec_name, same_a, same_b, round_trips);
        Longtail_Free(m_ref); Longtail_Free(m_b200); Longtail_Free(m_codec);
        SAFE_DISPOSE_API(s_ref); SAFE_DISPOSE_API(s_b200); SAFE_DISPOSE_API(s_codec);
        SAFE_DISPOSE_API(&k_ref->api); SAFE_DISPOSE_API(&k_b200->api); SAFE_DISPOSE_API(&k_codec->api);
        SAFE_DISPOSE_API(full); SAFE_DISPOSE_API(b200_registry);
        Longtail_Free(vi);
    }

    /* 4. golden chunker vector through the B200 ChunkerAPI */
    if (argc > 1)
    {
        FILE* f = fopen(argv[1], "rb");
        CHECK(f != 0, "open %s", argv[1]);
        if (f)
        {
            static const uint32_t expected[20] = {81590, 46796, 36543, 83172, 76749, 79550, 41484, 20326, 31652, 19995, 103873, 38087, 38377, 23449,
                                                  47321, 86692, 28268, 65465, 33255, 65932};
            uint8_t* data = (uint8_t*)malloc(1048576);
            size_t got = fread(data, 1, 1048576, f);
            fclose(f);
            struct feed fd = {data, got, 0};
            Longtail_ChunkerAPI_HChunker c;
            CHECK(b200_chunker->CreateChunker(b200_chunker, 16384, 65536, 262144, &c) == 0, "CreateChunker");
            struct Longtail_Chunker_ChunkRange r;
            uint64_t off = 0;
            for (int i = 0; i < 20; ++i)
            {
                CHECK(b200_chunker->NextChunk(b200_chunker, c, feed_func, &fd, &r) == 0, "NextChunk %d", i);
                CHECK(r.offset == off && r.len == expected[i], "chunk %d: offset %llu len %u", i, (unsigned long long)r.offset, r.len);
                uint64_t h1 = 0, h2 = 0;
                ref_hash->HashBuffer(ref_hash, r.len, r.buf, &h1);
                b200_hash->HashBuffer(b200_hash, r.len, r.buf, &h2);
                CHECK(h1 == h2, "chunk %d hash", i);
                off += r.len;
            }
            CHECK(b200_chunker->NextChunk(b200_chunker, c, feed_func, &fd, &r) == ESPIPE, "ESPIPE at end");
            CHECK(r.buf == 0 && r.offset == 1048576 && r.len == 0, "end range");
            b200_chunker->DisposeChunker(b200_chunker, c);
            free(data);
            printf("golden chunker vector through Longtail_CreateB200ChunkerAPI: %s\n", failures ? "see failures" : "20/20");
        }
    }
    /* generic HashBuffer (not a chunker range) incl. the KAT of test/test.cpp:465-474 and the empty buffer */
    {
        const char* s = "This is the first test string which is fairly long and should - reconstructed properly, than you very much";
        uint64_t h = 0;
        CHECK(b200_hash->HashBuffer(b200_hash, (uint32_t)strlen(s) + 1, s, &h) == 0 && h == 0xd38bbe79f1f03fdaull, "blake3 KAT %llx", (unsigned long long)h);
        uint64_t e1 = 0, e2 = 0;
        ref_hash->HashBuffer(ref_hash, 0, s, &e1);
        b200_hash->HashBuffer(b200_hash, 0, s, &e2);
        CHECK(e1 == e2, "empty hash");
        CHECK(b200_hash->GetIdentifier(b200_hash) == ref_hash->GetIdentifier(ref_hash), "identifier");
    }

    /* 6. faults injected through the boundary */
    fault_injection(storage, jobs, ref_hash, ref_chunker, infos, tags, target, b_ref, n_ref);

    Longtail_Free(b_ref);
    Longtail_Free(v_ref);
    free(tags);
    Longtail_Free(infos);
    SAFE_DISPOSE_API(b200_chunker);
    SAFE_DISPOSE_API(b200_hash);
    SAFE_DISPOSE_API(ref_chunker);
    SAFE_DISPOSE_API(ref_hash);
    SAFE_DISPOSE_API(jobs);
    SAFE_DISPOSE_API(storage);
    printf("%s (%d failures)\n", failures ? "DROPIN FAILED" : "DROPIN OK", failures);
    return failures ? 1 : 0;
}
