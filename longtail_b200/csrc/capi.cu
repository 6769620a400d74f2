// capi.cu — the C ABI of liblongtail_b200.so (include/longtail_b200.h): context, workspace and the batched verbs.
//
// This file holds host logic only; every byte of asset data is touched by the kernels in hpcdc.cu / blake3.cu / util.cu.
// There is deliberately no CPU fallback: without a CUDA device every entry point fails with ENODEV.
#include "../../include/longtail_b200.h"
#include "lt_kernels.h"

#include <errno.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <vector>

using namespace ltb;

namespace {

const uint32_t k_hpcdc_table[256] = {
#include "hpcdc_table.inc"
};

enum WsSlot
{
    WS_PARTS, WS_TILE_DESC, WS_TILE_COUNT, WS_TILE_SLOTS, WS_CAND, WS_STAGE_OFF, WS_STAGE_LEN, WS_PART_COUNT, WS_PART_BASE,
    WS_SCAN_TMP, WS_CHUNK_OFF, WS_CHUNK_LEN, WS_CHUNK_TAG, WS_CHUNK_HASH, WS_LEAF_COUNT, WS_LEAF_PREFIX, WS_CVS,
    WS_SEG_OFF, WS_SEG_LEN, WS_SEG_HASH, WS_SEG_LEAF_COUNT, WS_SEG_LEAF_PREFIX, WS_SEG_CVS,
    WS_TAB_HASH, WS_TAB_LEN, WS_TAB_TAG,
    WS_DEDUP_KEYS, WS_DEDUP_VALS, WS_DEDUP_FIRST, WS_DEDUP_ISFIRST, WS_DEDUP_UIDX, WS_ACI, WS_UHASH, WS_ULEN, WS_UTAG,
    WS_PATHS, WS_INDEX_OUT, WS_ARENA_A, WS_ARENA_B, WS_ACC_HASH, WS_ACC_LEN, WS_ACC_TAG,
    WS_UOFF, WS_BLK_HASHES, WS_BLK_SEG_OFF, WS_BLK_SEG_LEN, WS_BLK_HASH_OUT, WS_BLK_SRC_OFF, WS_BLK_DST_OFF, WS_BLK_LEN, WS_BLK_RAW, WS_BLK_OUT,
    WS_BLK_RAW_OFF, WS_BLK_RAW_LEN, WS_BLK_OUT_OFF, WS_BLK_OUT_LEN, WS_BLK_JOBS, WS_BLK_JOB_START, WS_BLK_JOB_COUNT, WS_QUEUE_HEAD, WS_MERGE_A, WS_MERGE_B, WS_MERGE_COUNTS, WS_MEOW_TABLE,
    WS_LZ4_TABLES, WS_LZ4_V2, WS_LZ4_TABLES_B, WS_LZ4_V2_B, WS_BLK_RAW_B, WS_BLK_OUT_B, WS_BLK_JOBS_B, WS_BLK_TAB, WS_BLK_TAB_B, WS_BLK_RAW_C, WS_BLK_OUT_C, WS_BLK_JOBS_C, WS_BLK_TAB_C, WS_LZ4_TABLES_C, WS_LZ4_V2_C, WS_BLK_RAW_D, WS_BLK_OUT_D, WS_BLK_JOBS_D, WS_BLK_TAB_D, WS_LZ4_TABLES_D, WS_LZ4_V2_D, WS_BLK_CHUNK_SIZES, WS_UPLOAD_SEGS_A, WS_UPLOAD_SEGS_B, WS_UFIRST, WS_G_COUNTS, WS_G_HASH, WS_G_LEN, WS_G_TAG, WS_X_SEND, WS_X_RECV, WS_X_TAB, WS_PACK_A, WS_PACK_B, WS_PACK_C, WS_ZSTD_WORKERS, WS_ZSTD_RAW_OFF, WS_ZSTD_RAW_LEN, WS_ZSTD_OUT_OFF, WS_ZSTD_OUT_LEN, WS_ZSTD_DEC_WORKERS,
    WS_COUNT
};

enum HostSlot
{
    HS_PARTS, HS_SMALL, HS_RANGE_COUNTS, HS_CHUNK_HASH, HS_CHUNK_LEN, HS_CHUNK_TAG, HS_CHUNK_OFF, HS_SEG, HS_INDEX_OUT, HS_STAGE_A, HS_STAGE_B, HS_BLK_META, HS_BLK_OUT_LEN, HS_BLK_OUT_LEN_B, HS_BLK_TAB, HS_BLK_TAB_B, HS_BLK_OUT_LEN_C, HS_BLK_TAB_C, HS_BLK_OUT_LEN_D, HS_BLK_TAB_D, HS_BLK_STAGE, HS_UPLOAD_SEGS_A, HS_UPLOAD_SEGS_B, HS_COUNT
};

struct Buf
{
    void* p = nullptr;
    size_t cap = 0;
};

} // namespace

struct lt_b200_context
{
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr; // more compute streams: WriteContent keeps up to four batches of stored blocks in flight
    cudaStream_t aux_stream2[2] = {nullptr, nullptr};
    cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
    cudaEvent_t slot_done[2] = {nullptr, nullptr};
    cudaStream_t upload_stream = nullptr; // host -> device copies of the streaming upsync (its device -> host copies own copy_stream)
    cudaEvent_t upload_done[2] = {nullptr, nullptr}, arena_free[2] = {nullptr, nullptr};
    cudaEvent_t copy_done[2] = {nullptr, nullptr};
    cudaEvent_t compute_done[2] = {nullptr, nullptr};
    uint32_t* d_table = nullptr;
    ScanLayout scan_layout = {0, 0, 0};
    Buf ws[WS_COUNT];
    Buf hs[HS_COUNT];
    uint64_t launches = 0;
    char err[512] = {0};
    // chunk table left resident by the last lt_b200_chunk_ranges call
    uint32_t table_chunks = 0;
    uint32_t unique_chunks = 0; // entries of WS_UHASH/ULEN/UTAG/UOFF left by the last index build (UOFF only for resident arenas)
    bool unique_offsets_valid = false;
    // optional per-kernel CUDA-event timing (lt_b200_profile_*)
    bool prof_on = false;
    struct Span { cudaEvent_t a, b; int id; uint64_t bytes; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    double prof_ms[LT_B200_KERNEL_COUNT] = {0};
    uint64_t prof_launches[LT_B200_KERNEL_COUNT] = {0};
    uint64_t prof_bytes[LT_B200_KERNEL_COUNT] = {0};
};

// ---------------------------------------------------------------- NCCL, loaded on first use
// Only the multi-GPU verbs need NCCL; it is dlopen'ed (libnccl.so.2: the copy already in the process, e.g. torch's, wins) so that a
// single-GPU caller has no such dependency.
struct NcclApi
{
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_lock;

static int nccl_load()
{
    std::lock_guard<std::mutex> lock(g_nccl_lock);
    if (g_nccl.lib) return 0;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return ENOSYS;
    NcclApi a;
    a.lib = lib;
#define LT_NCCL_SYM(field, name) *reinterpret_cast<void**>(&a.field) = dlsym(lib, name); if (!a.field) { dlclose(lib); return ENOSYS; }
    LT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    LT_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    LT_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    LT_NCCL_SYM(Broadcast, "ncclBroadcast")
    LT_NCCL_SYM(AllReduce, "ncclAllReduce")
    LT_NCCL_SYM(AllGather, "ncclAllGather")
    LT_NCCL_SYM(Send, "ncclSend")
    LT_NCCL_SYM(Recv, "ncclRecv")
    LT_NCCL_SYM(GroupStart, "ncclGroupStart")
    LT_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    LT_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef LT_NCCL_SYM
    g_nccl = a;
    return 0;
}

// One communicator per context: `world` processes, one GPU each (the reference has no distributed runtime at all — SURVEY.md F1; the
// unit that is sharded is its own job, one (asset, part) pair, src/longtail.c:2396-2457)
struct lt_b200_comm
{
    lt_b200_context* ctx = nullptr;
    ncclComm_t comm = nullptr;
    uint32_t rank = 0, world = 1;
    // state left by the last lt_b200_index_sharded for lt_b200_write_blocks_sharded
    std::vector<uint32_t> rank_chunk_start; // [world + 1] first global chunk ordinal of every rank's slice
    std::vector<uint32_t> rank_unique_start; // [world + 1] first unique chunk whose first occurrence lies in every rank's slice
    uint32_t global_chunks = 0, global_unique = 0;
    const uint8_t* d_arena = nullptr;
    uint64_t arena_size = 0;
    uint32_t hash_type = 0;
};

namespace {

int fail(lt_b200_context* c, int code, const char* fmt, ...)
{
    if (c)
    {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

int cuda_fail(lt_b200_context* c, cudaError_t e, const char* what)
{
    return fail(c, e == cudaErrorMemoryAllocation ? ENOMEM : EIO, "%s: %s", what, cudaGetErrorString(e));
}

#define CU(call)                                                   \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(c, e__, #call);   \
    } while (0)

#define TRY(call)                  \
    do {                           \
        int r__ = (call);          \
        if (r__) return r__;       \
    } while (0)

#define NC(call)                                                                                      \
    do {                                                                                              \
        ncclResult_t n__ = (call);                                                                    \
        if (n__ != ncclSuccess) return fail(c, EIO, "%s: %s", #call, g_nccl.GetErrorString(n__));     \
    } while (0)

int ws_reserve(lt_b200_context* c, int slot, size_t bytes)
{
    Buf& b = c->ws[slot];
    if (b.cap >= bytes && b.p) return 0;
    if (b.p)
    {
        CU(cudaDeviceSynchronize()); // kernels of either compute stream may still use the old buffer
        CU(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        want = bytes ? bytes : 256;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess)
    {
        // the grow-only workspace of the OTHER phase is in the way (the block batches of WriteContent take whatever the arena leaves;
        // the chunk + hash scratch is ~5 % of the arena): give it back and try once more — it regrows when that phase runs again
        cudaGetLastError();
        static const int write_phase[] = {WS_BLK_RAW, WS_BLK_OUT, WS_BLK_JOBS, WS_BLK_RAW_B, WS_BLK_OUT_B, WS_BLK_JOBS_B, WS_BLK_TAB, WS_BLK_TAB_B,
                                          WS_BLK_RAW_C, WS_BLK_OUT_C, WS_BLK_JOBS_C, WS_BLK_TAB_C, WS_BLK_RAW_D, WS_BLK_OUT_D, WS_BLK_JOBS_D, WS_BLK_TAB_D,
                                          WS_LZ4_TABLES, WS_LZ4_TABLES_B, WS_LZ4_TABLES_C, WS_LZ4_TABLES_D, WS_X_SEND, WS_X_RECV, WS_PACK_A};
        bool is_write_slot = false;
        for (int w : write_phase) is_write_slot = is_write_slot || w == slot;
        if (!is_write_slot)
        {
            cudaDeviceSynchronize();
            for (int w : write_phase)
            {
                Buf& o = c->ws[w];
                if (o.p) cudaFree(o.p);
                o.p = nullptr;
                o.cap = 0;
            }
            e = cudaMalloc(&b.p, want);
        }
    }
    if (e != cudaSuccess) { b.p = nullptr; cudaGetLastError(); return cuda_fail(c, e, "cudaMalloc(workspace)"); }
    b.cap = want;
    return 0;
}

template <typename T>
T* ws(lt_b200_context* c, int slot) { return static_cast<T*>(c->ws[slot].p); }

int hs_reserve(lt_b200_context* c, int slot, size_t bytes)
{
    Buf& b = c->hs[slot];
    if (b.cap >= bytes && b.p) return 0;
    if (b.p)
    {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFreeHost(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    CU(cudaHostAlloc(&b.p, want, cudaHostAllocDefault));
    b.cap = want;
    return 0;
}

template <typename T>
T* hs(lt_b200_context* c, int slot) { return static_cast<T*>(c->hs[slot].p); }

cudaEvent_t prof_event(lt_b200_context* c)
{
    cudaEvent_t e = nullptr;
    if (!c->event_pool.empty()) { e = c->event_pool.back(); c->event_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

// brackets the next kernel launch(es) on the context stream with CUDA events when profiling is on
struct ProfScope
{
    lt_b200_context* c;
    size_t idx;
    bool on;
    cudaStream_t st;
    ProfScope(lt_b200_context* ctx, int id, uint64_t bytes, cudaStream_t stream = nullptr) : c(ctx), idx(0), on(ctx->prof_on), st(stream ? stream : ctx->stream)
    {
        if (!on) return;
        lt_b200_context::Span s = {prof_event(c), prof_event(c), id, bytes};
        cudaEventRecord(s.a, st);
        idx = c->spans.size();
        c->spans.push_back(s);
    }
    ~ProfScope() { if (on) cudaEventRecord(c->spans[idx].b, st); }
};

// Launches of one kernel may overlap in time (WriteContent keeps two batches in flight on two streams): a kernel's time is the length of
// the UNION of its launches' intervals, so that bytes / time stays a throughput of the device and never counts a moment twice.
void prof_collect(lt_b200_context* c)
{
    if (c->spans.empty()) return;
    struct Iv { float a, b; };
    std::vector<Iv> iv[LT_B200_KERNEL_COUNT];
    const cudaEvent_t ref = c->spans[0].a;
    for (auto& s : c->spans)
    {
        float ta = 0, tb = 0;
        if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&ta, ref, s.a) == cudaSuccess &&
            cudaEventElapsedTime(&tb, ref, s.b) == cudaSuccess)
        {
            iv[s.id].push_back({ta, tb});
            c->prof_launches[s.id] += 1;
            c->prof_bytes[s.id] += s.bytes;
        }
    }
    for (int k = 0; k < LT_B200_KERNEL_COUNT; ++k)
    {
        std::sort(iv[k].begin(), iv[k].end(), [](const Iv& x, const Iv& y) { return x.a < y.a; });
        float end = -1e30f;
        for (const Iv& x : iv[k])
        {
            if (x.b <= end) continue;
            c->prof_ms[k] += x.b - std::max(x.a, end);
            end = x.b;
        }
    }
    for (auto& s : c->spans)
    {
        c->event_pool.push_back(s.a);
        c->event_pool.push_back(s.b);
    }
    c->spans.clear();
    cudaGetLastError();
}

uint32_t inverse_mod_2_32(uint32_t odd)
{
    uint32_t x = odd; // correct to 3 bits; each Newton step doubles the precision
    for (int i = 0; i < 5; ++i) x *= 2u - odd * x;
    return x;
}

int make_chunk_params(lt_b200_context* c, uint32_t mn, uint32_t av, uint32_t mx, ChunkParams* cp)
{
    // lib/hpcdcchunker/longtail_hpcdcchunker.c:146-150
    if (mn < (uint32_t)SCAN_WINDOW || mn > mx || mn > av || av > mx) return fail(c, EINVAL, "invalid chunker parameters %u/%u/%u", mn, av, mx);
    const double a = (double)av;
    const uint32_t d = (uint32_t)(a / (-1.42888852e-7 * a + 1.33237515)); // :126-129, evaluated in double like the reference
    if (d == 0) return fail(c, EINVAL, "discriminator is zero for avg %u", av);
    uint32_t odd = d;
    while (!(odd & 1u)) odd >>= 1;
    cp->min = mn;
    cp->avg = av;
    cp->max = mx;
    cp->d = d;
    cp->d_odd_inv = inverse_mod_2_32(odd);
    cp->d_odd_thr = 0xffffffffu / odd;
    uint32_t expect = (uint32_t)SCAN_TILE / d + 1;
    uint32_t slots = (6 * expect + 2 + 7) & ~7u;
    if (slots > 4096) slots = 4096;
    cp->slots = slots;
    return 0;
}

// hash segments whose offsets / lengths already sit in device memory; result stays on the device
int hash_segments_device(lt_b200_context* c, uint32_t hash_type, const uint8_t* d_base, uint64_t base_size, const uint64_t* d_off,
                         const uint32_t* d_len, uint32_t count, uint64_t upper_leaves, int slot_leaf_count, int slot_leaf_prefix,
                         int slot_cvs, uint64_t* d_hash_out, uint64_t payload_bytes = 0, uint32_t max_segment_bytes = 0xffffffffu)
{
    if (hash_type != LT_B200_HASH_BLAKE3 && hash_type != LT_B200_HASH_BLAKE2 && hash_type != LT_B200_HASH_MEOW)
        return fail(c, ENOTSUP, "hash type 0x%08x has no device implementation", hash_type);
    if (!count) return 0;
    if (hash_type == LT_B200_HASH_BLAKE2)
    {
        TRY(ws_reserve(c, WS_QUEUE_HEAD, 256));
        ProfScope ps(c, LT_B200_KERNEL_BLAKE2S, payload_bytes);
        launch_blake2s_segments(d_base, d_off, d_len, count, ws<uint32_t>(c, WS_QUEUE_HEAD), d_hash_out, c->sm_count, c->stream);
        c->launches += 1;
        CU(cudaGetLastError());
        return 0;
    }
    if (hash_type == LT_B200_HASH_MEOW)
    {
        TRY(ws_reserve(c, WS_QUEUE_HEAD, 256));
        if (!c->ws[WS_MEOW_TABLE].p)
        {
            uint32_t td0[256];
            meow_build_table(td0);
            TRY(ws_reserve(c, WS_MEOW_TABLE, sizeof(td0)));
            CU(cudaMemcpyAsync(ws<void>(c, WS_MEOW_TABLE), td0, sizeof(td0), cudaMemcpyHostToDevice, c->stream));
            CU(cudaStreamSynchronize(c->stream)); // td0 lives on this stack frame
        }
        ProfScope ps(c, LT_B200_KERNEL_MEOW, payload_bytes);
        CU(launch_meow_segments(d_base, d_off, d_len, count, ws<uint32_t>(c, WS_QUEUE_HEAD), d_hash_out, ws<uint32_t>(c, WS_MEOW_TABLE), c->sm_count, c->stream));
        c->launches += 1;
        return 0;
    }
    TRY(ws_reserve(c, slot_leaf_count, sizeof(uint32_t) * (size_t)count));
    TRY(ws_reserve(c, slot_leaf_prefix, sizeof(uint32_t) * ((size_t)count + 1)));
    TRY(ws_reserve(c, WS_SCAN_TMP, sizeof(uint32_t) * scan_tmp_words(count)));
    TRY(ws_reserve(c, slot_cvs, 32 * (size_t)upper_leaves));
    TRY(ws_reserve(c, WS_MERGE_A, sizeof(uint4) * ((size_t)upper_leaves / 2 + 2)));
    TRY(ws_reserve(c, WS_MERGE_B, sizeof(uint4) * ((size_t)upper_leaves / 2 + 2)));
    TRY(ws_reserve(c, WS_MERGE_COUNTS, sizeof(uint32_t) * BLAKE3_MAX_LEVELS));
    launch_leaf_counts(d_len, count, ws<uint32_t>(c, slot_leaf_count), c->stream);
    launch_exclusive_scan(ws<uint32_t>(c, slot_leaf_count), count, ws<uint32_t>(c, slot_leaf_prefix), ws<uint32_t>(c, WS_SCAN_TMP), c->stream);
    c->launches += 4;
    uint32_t total_leaves = 0;
    CU(cudaMemcpyAsync(&total_leaves, ws<uint32_t>(c, slot_leaf_prefix) + count, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if ((uint64_t)total_leaves > upper_leaves) return fail(c, EFAULT, "leaf count %u exceeds bound %llu", total_leaves, (unsigned long long)upper_leaves);
    uint64_t seg_bytes = (uint64_t)total_leaves * 1024; // upper bound; callers that know the exact byte count pass it instead
    if (payload_bytes) seg_bytes = payload_bytes;
    {
        ProfScope ps(c, LT_B200_KERNEL_BLAKE3_LEAVES, seg_bytes);
        launch_blake3_leaves(d_base, base_size, d_off, d_len, ws<uint32_t>(c, slot_leaf_prefix), count, total_leaves,
                             ws<uint32_t>(c, slot_cvs), d_hash_out, ws<uint4>(c, WS_MERGE_A), ws<uint32_t>(c, WS_MERGE_COUNTS), c->stream);
    }
    {
        ProfScope ps(c, LT_B200_KERNEL_BLAKE3_MERGE, (uint64_t)total_leaves * 32);
        const uint32_t max_leaves = max_segment_bytes == 0xffffffffu ? 0x400000u : (max_segment_bytes + 1023u) / 1024u;
        c->launches += launch_blake3_merge(total_leaves, count, max_leaves, ws<uint32_t>(c, slot_cvs), d_hash_out, ws<uint4>(c, WS_MERGE_A),
                                           ws<uint4>(c, WS_MERGE_B), ws<uint32_t>(c, WS_MERGE_COUNTS), c->stream);
    }
    c->launches += 1;
    CU(cudaGetLastError());
    return 0;
}

} // namespace

// ================================================================ context

extern "C" int lt_b200_context_create(int device_ordinal, lt_b200_context** out_context)
{
    if (!out_context) return EINVAL;
    *out_context = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device_ordinal < 0 || device_ordinal >= n)
    {
        cudaGetLastError();
        return ENODEV; // no CUDA device: there is no CPU fallback by design
    }
    lt_b200_context* c = new (std::nothrow) lt_b200_context();
    if (!c) return ENOMEM;
    c->device = device_ordinal;
    cudaDeviceProp prop;
    if (cudaSetDevice(device_ordinal) != cudaSuccess || cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->aux_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->aux_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaMalloc(&c->d_table, sizeof(k_hpcdc_table)) != cudaSuccess ||
        cudaMemcpy(c->d_table, k_hpcdc_table, sizeof(k_hpcdc_table), cudaMemcpyHostToDevice) != cudaSuccess)
    {
        cudaGetLastError();
        delete c;
        return EIO;
    }
    for (int i = 0; i < 2; ++i)
    {
        cudaStreamCreateWithFlags(&c->aux_stream2[i], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&c->upload_done[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->arena_free[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->slot_done[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->copy_done[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->compute_done[i], cudaEventDisableTiming);
    }
    cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking);
    c->sm_count = prop.multiProcessorCount;
    if (make_scan_layout(&c->scan_layout, c->stream) != cudaSuccess)
    {
        cudaGetLastError();
        lt_b200_context_destroy(c);
        return ENOTSUP; // the device cannot hold the scan kernel's shared-memory layout (not a B200-class part)
    }
    *out_context = c;
    return 0;
}

extern "C" void lt_b200_context_destroy(lt_b200_context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    prof_collect(c);
    for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
    for (Buf& b : c->ws) if (b.p) cudaFree(b.p);
    for (Buf& b : c->hs) if (b.p) cudaFreeHost(b.p);
    if (c->d_table) cudaFree(c->d_table);
    for (int i = 0; i < 2; ++i)
    {
        if (c->copy_done[i]) cudaEventDestroy(c->copy_done[i]);
        if (c->slot_done[i]) cudaEventDestroy(c->slot_done[i]);
        if (c->upload_done[i]) cudaEventDestroy(c->upload_done[i]);
        if (c->arena_free[i]) cudaEventDestroy(c->arena_free[i]);
        if (c->aux_stream2[i]) { cudaStreamSynchronize(c->aux_stream2[i]); cudaStreamDestroy(c->aux_stream2[i]); }
        if (c->compute_done[i]) cudaEventDestroy(c->compute_done[i]);
    }
    cudaStreamDestroy(c->stream);
    if (c->aux_stream) { cudaStreamSynchronize(c->aux_stream); cudaStreamDestroy(c->aux_stream); }
    if (c->aux_fork) cudaEventDestroy(c->aux_fork);
    if (c->aux_join) cudaEventDestroy(c->aux_join);
    cudaStreamDestroy(c->copy_stream);
    if (c->upload_stream) { cudaStreamSynchronize(c->upload_stream); cudaStreamDestroy(c->upload_stream); }
    delete c;
}

extern "C" const char* lt_b200_last_error(const lt_b200_context* c) { return c ? c->err : "no context"; }
extern "C" uint64_t lt_b200_launch_count(const lt_b200_context* c) { return c ? c->launches : 0; }
extern "C" void* lt_b200_stream(lt_b200_context* c) { return c ? (void*)c->stream : nullptr; }

extern "C" int lt_b200_synchronize(lt_b200_context* c)
{
    if (!c) return EINVAL;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->copy_stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// frees the grow-only device workspace of the context (the buffers come back on demand): for callers that move on to a phase with a
// different memory picture, e.g. from a resident-arena upsync to one that brings its own arena
extern "C" int lt_b200_trim(lt_b200_context* c)
{
    if (!c) return EINVAL;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    for (Buf& b : c->ws)
    {
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    c->table_chunks = 0;
    c->unique_chunks = 0;
    c->unique_offsets_valid = false;
    return 0;
}

extern "C" int lt_b200_profile_enable(lt_b200_context* c, int on)
{
    if (!c) return EINVAL;
    c->prof_on = on != 0;
    return 0;
}

extern "C" int lt_b200_profile_reset(lt_b200_context* c)
{
    if (!c) return EINVAL;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    prof_collect(c);
    for (int i = 0; i < LT_B200_KERNEL_COUNT; ++i) { c->prof_ms[i] = 0; c->prof_launches[i] = 0; c->prof_bytes[i] = 0; }
    return 0;
}

extern "C" int lt_b200_profile_read(lt_b200_context* c, uint32_t kernel, double* out_ms, uint64_t* out_launches, uint64_t* out_bytes)
{
    if (!c || kernel >= LT_B200_KERNEL_COUNT) return EINVAL;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    prof_collect(c);
    if (out_ms) *out_ms = c->prof_ms[kernel];
    if (out_launches) *out_launches = c->prof_launches[kernel];
    if (out_bytes) *out_bytes = c->prof_bytes[kernel];
    return 0;
}

extern "C" int lt_b200_device_alloc(lt_b200_context* c, uint64_t bytes, void** out)
{
    if (!c || !out) return EINVAL;
    CU(cudaSetDevice(c->device));
    CU(cudaMalloc(out, bytes ? bytes : 1));
    return 0;
}

extern "C" int lt_b200_device_free(lt_b200_context* c, void* p)
{
    if (!c) return EINVAL;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(p));
    return 0;
}

extern "C" int lt_b200_host_alloc_pinned(lt_b200_context* c, uint64_t bytes, void** out)
{
    if (!c || !out) return EINVAL;
    CU(cudaSetDevice(c->device));
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return 0;
}

extern "C" int lt_b200_host_free_pinned(lt_b200_context* c, void* p)
{
    if (!c) return EINVAL;
    CU(cudaFreeHost(p));
    return 0;
}

extern "C" int lt_b200_copy_to_device(lt_b200_context* c, void* dst, const void* src, uint64_t bytes)
{
    if (!c) return EINVAL;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// queued on the context stream without waiting (the next verb on this context runs after it); host_src must stay valid until then
extern "C" int lt_b200_copy_to_device_async(lt_b200_context* c, void* dst, const void* src, uint64_t bytes)
{
    if (!c || (bytes && (!dst || !src))) return EINVAL;
    CU(cudaSetDevice(c->device));
    if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

extern "C" int lt_b200_copy_to_host(lt_b200_context* c, void* dst, const void* src, uint64_t bytes)
{
    if (!c) return EINVAL;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int lt_b200_synth_fill(lt_b200_context* c, void* dst, uint64_t bytes, const lt_b200_synth_spec* spec, uint64_t asset_id, uint64_t offset)
{
    if (!c || !spec || (offset & 15u)) return EINVAL;
    CU(cudaSetDevice(c->device));
    lt_synth_spec_dev s = {spec->seed, spec->shared_permille, spec->pool_segments, spec->class_mode, 0};
    launch_synth_fill(static_cast<uint8_t*>(dst), bytes, s, asset_id, offset, c->stream);
    CU(cudaGetLastError());
    return 0;
}

// ================================================================ layer 1

extern "C" int lt_b200_chunk_ranges(lt_b200_context* c, const uint8_t* d_arena, uint64_t arena_size, const lt_b200_range* ranges,
                                    uint32_t range_count, uint32_t mn, uint32_t av, uint32_t mx, uint32_t hash_type, int want_host,
                                    lt_b200_chunk_table* out)
{
    if (!c || !out || (range_count && !ranges)) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    memset(out, 0, sizeof(*out));
    c->table_chunks = 0;
    ChunkParams cp;
    TRY(make_chunk_params(c, mn, av, mx, &cp));
    if (hash_type != LT_B200_HASH_BLAKE3 && hash_type != LT_B200_HASH_BLAKE2 && hash_type != LT_B200_HASH_MEOW)
        return fail(c, ENOTSUP, "hash type 0x%08x has no device implementation", hash_type);
    if (((uintptr_t)d_arena) & 15u) return fail(c, EINVAL, "device arena must be 16-byte aligned");
    out->range_count = range_count;
    if (!range_count) return 0;

    // part descriptors (host) -> device
    TRY(hs_reserve(c, HS_PARTS, sizeof(PartDesc) * (size_t)range_count));
    PartDesc* h_parts = hs<PartDesc>(c, HS_PARTS);
    uint64_t tiles = 0, stage = 0, bytes = 0;
    for (uint32_t r = 0; r < range_count; ++r)
    {
        const lt_b200_range& rg = ranges[r];
        if ((rg.arena_offset & 15u) || rg.size >= 0x80000000u || rg.arena_offset + rg.size > arena_size)
            return fail(c, EINVAL, "range %u (offset %llu size %u) is misaligned or outside the arena", r, (unsigned long long)rg.arena_offset, rg.size);
        PartDesc& p = h_parts[r];
        p.data_off = rg.arena_offset;
        p.size = rg.size;
        p.tile_start = (uint32_t)tiles;
        p.chunk_start = (uint32_t)stage;
        p.asset = r;
        p.tag = rg.tag;
        p.pad = 0;
        tiles += (rg.size + (uint32_t)SCAN_TILE - 1) / (uint32_t)SCAN_TILE;
        stage += rg.size / mn + 2;
        bytes += rg.size;
    }
    if (tiles >= 0xffffffffull || stage >= 0xffffffffull) return fail(c, E2BIG, "batch too large: %llu tiles, %llu chunk slots", (unsigned long long)tiles, (unsigned long long)stage);
    const uint32_t num_tiles = (uint32_t)tiles;

    TRY(ws_reserve(c, WS_PARTS, sizeof(PartDesc) * (size_t)range_count));
    TRY(ws_reserve(c, WS_TILE_DESC, sizeof(uint32_t) * (size_t)num_tiles));
    TRY(ws_reserve(c, WS_TILE_COUNT, sizeof(uint32_t) * (size_t)num_tiles));
    TRY(ws_reserve(c, WS_TILE_SLOTS, sizeof(uint32_t) * (size_t)num_tiles * cp.slots));
    TRY(ws_reserve(c, WS_CAND, sizeof(uint32_t) * (size_t)num_tiles * cp.slots));
    TRY(ws_reserve(c, WS_STAGE_OFF, sizeof(uint64_t) * (size_t)stage));
    TRY(ws_reserve(c, WS_STAGE_LEN, sizeof(uint32_t) * (size_t)stage));
    TRY(ws_reserve(c, WS_PART_COUNT, sizeof(uint32_t) * (size_t)range_count));
    TRY(ws_reserve(c, WS_PART_BASE, sizeof(uint32_t) * ((size_t)range_count + 1)));
    TRY(ws_reserve(c, WS_SCAN_TMP, sizeof(uint32_t) * scan_tmp_words(range_count)));

    CU(cudaMemcpyAsync(ws<PartDesc>(c, WS_PARTS), h_parts, sizeof(PartDesc) * (size_t)range_count, cudaMemcpyHostToDevice, c->stream));
    launch_tile_part(ws<PartDesc>(c, WS_PARTS), range_count, num_tiles, ws<uint32_t>(c, WS_TILE_DESC), c->stream);
    {
        ProfScope ps(c, LT_B200_KERNEL_HPCDC_SCAN, bytes);
        CU(launch_hpcdc_scan(d_arena, ws<PartDesc>(c, WS_PARTS), ws<uint32_t>(c, WS_TILE_DESC), num_tiles, cp, c->d_table,
                             ws<uint32_t>(c, WS_TILE_COUNT), ws<uint32_t>(c, WS_TILE_SLOTS), c->scan_layout, c->sm_count, c->stream));
    }
    {
        ProfScope ps(c, LT_B200_KERNEL_HPCDC_WALK, (uint64_t)num_tiles * 4);
        launch_hpcdc_walk(d_arena, ws<PartDesc>(c, WS_PARTS), range_count, cp, c->d_table, ws<uint32_t>(c, WS_TILE_COUNT),
                          ws<uint32_t>(c, WS_TILE_SLOTS), ws<uint32_t>(c, WS_CAND), ws<uint64_t>(c, WS_STAGE_OFF),
                          ws<uint32_t>(c, WS_STAGE_LEN), ws<uint32_t>(c, WS_PART_COUNT), c->stream);
    }
    launch_exclusive_scan(ws<uint32_t>(c, WS_PART_COUNT), range_count, ws<uint32_t>(c, WS_PART_BASE), ws<uint32_t>(c, WS_SCAN_TMP), c->stream);
    c->launches += 6;

    TRY(hs_reserve(c, HS_RANGE_COUNTS, sizeof(uint32_t) * ((size_t)range_count + 1)));
    uint32_t* h_counts = hs<uint32_t>(c, HS_RANGE_COUNTS);
    CU(cudaMemcpyAsync(h_counts, ws<uint32_t>(c, WS_PART_COUNT), sizeof(uint32_t) * (size_t)range_count, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(h_counts + range_count, ws<uint32_t>(c, WS_PART_BASE) + range_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    const uint32_t chunk_count = h_counts[range_count];
    out->chunk_count = chunk_count;
    out->range_chunk_counts = h_counts;
    if (!chunk_count) return 0;

    TRY(ws_reserve(c, WS_CHUNK_OFF, sizeof(uint64_t) * (size_t)chunk_count));
    TRY(ws_reserve(c, WS_CHUNK_LEN, sizeof(uint32_t) * (size_t)chunk_count));
    TRY(ws_reserve(c, WS_CHUNK_TAG, sizeof(uint32_t) * (size_t)chunk_count));
    TRY(ws_reserve(c, WS_CHUNK_HASH, sizeof(uint64_t) * (size_t)chunk_count));
    launch_compact_chunks(ws<PartDesc>(c, WS_PARTS), range_count, ws<uint32_t>(c, WS_PART_COUNT), ws<uint32_t>(c, WS_PART_BASE),
                          ws<uint64_t>(c, WS_STAGE_OFF), ws<uint32_t>(c, WS_STAGE_LEN), ws<uint64_t>(c, WS_CHUNK_OFF),
                          ws<uint32_t>(c, WS_CHUNK_LEN), ws<uint32_t>(c, WS_CHUNK_TAG), c->stream);
    c->launches += 1;
    TRY(hash_segments_device(c, hash_type, d_arena, arena_size, ws<uint64_t>(c, WS_CHUNK_OFF), ws<uint32_t>(c, WS_CHUNK_LEN), chunk_count,
                             bytes / 1024 + chunk_count, WS_LEAF_COUNT, WS_LEAF_PREFIX, WS_CVS, ws<uint64_t>(c, WS_CHUNK_HASH), bytes, mx));
    c->table_chunks = chunk_count;
    if (want_host)
    {
        TRY(hs_reserve(c, HS_CHUNK_HASH, sizeof(uint64_t) * (size_t)chunk_count));
        TRY(hs_reserve(c, HS_CHUNK_LEN, sizeof(uint32_t) * (size_t)chunk_count));
        TRY(hs_reserve(c, HS_CHUNK_TAG, sizeof(uint32_t) * (size_t)chunk_count));
        TRY(hs_reserve(c, HS_CHUNK_OFF, sizeof(uint64_t) * (size_t)chunk_count));
        CU(cudaMemcpyAsync(hs<void>(c, HS_CHUNK_HASH), ws<void>(c, WS_CHUNK_HASH), sizeof(uint64_t) * (size_t)chunk_count, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(hs<void>(c, HS_CHUNK_LEN), ws<void>(c, WS_CHUNK_LEN), sizeof(uint32_t) * (size_t)chunk_count, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(hs<void>(c, HS_CHUNK_TAG), ws<void>(c, WS_CHUNK_TAG), sizeof(uint32_t) * (size_t)chunk_count, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(hs<void>(c, HS_CHUNK_OFF), ws<void>(c, WS_CHUNK_OFF), sizeof(uint64_t) * (size_t)chunk_count, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        out->chunk_hashes = hs<uint64_t>(c, HS_CHUNK_HASH);
        out->chunk_sizes = hs<uint32_t>(c, HS_CHUNK_LEN);
        out->chunk_tags = hs<uint32_t>(c, HS_CHUNK_TAG);
        out->chunk_offsets = hs<uint64_t>(c, HS_CHUNK_OFF);
    }
    return 0;
}

extern "C" int lt_b200_hash_segments(lt_b200_context* c, uint32_t hash_type, const uint8_t* d_base, uint64_t base_size,
                                     const uint64_t* offsets, const uint32_t* sizes, uint32_t count, uint64_t* out_hashes)
{
    if (!c || (count && (!offsets || !sizes || !out_hashes))) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    if (!count) return 0;
    if (((uintptr_t)d_base) & 15u) return fail(c, EINVAL, "device buffer must be 16-byte aligned");
    uint64_t upper = count;
    uint32_t max_size = 0;
    for (uint32_t i = 0; i < count; ++i)
    {
        if (offsets[i] + sizes[i] > base_size) return fail(c, EINVAL, "segment %u outside the buffer", i);
        upper += sizes[i] / 1024;
        if (sizes[i] > max_size) max_size = sizes[i];
    }
    TRY(ws_reserve(c, WS_SEG_OFF, sizeof(uint64_t) * (size_t)count));
    TRY(ws_reserve(c, WS_SEG_LEN, sizeof(uint32_t) * (size_t)count));
    TRY(ws_reserve(c, WS_SEG_HASH, sizeof(uint64_t) * (size_t)count));
    CU(cudaMemcpyAsync(ws<void>(c, WS_SEG_OFF), offsets, sizeof(uint64_t) * (size_t)count, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(ws<void>(c, WS_SEG_LEN), sizes, sizeof(uint32_t) * (size_t)count, cudaMemcpyHostToDevice, c->stream));
    TRY(hash_segments_device(c, hash_type, d_base, base_size, ws<uint64_t>(c, WS_SEG_OFF), ws<uint32_t>(c, WS_SEG_LEN), count, upper,
                             WS_SEG_LEAF_COUNT, WS_SEG_LEAF_PREFIX, WS_SEG_CVS, ws<uint64_t>(c, WS_SEG_HASH), 0, max_size));
    CU(cudaMemcpyAsync(out_hashes, ws<void>(c, WS_SEG_HASH), sizeof(uint64_t) * (size_t)count, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// ================================================================ layer 2

namespace {

// device-side table to index: [d_hash, d_len, d_tag] x chunk_count already resident
int build_index_from_device_table(lt_b200_context* c, const lt_b200_assets* a, const uint32_t* asset_chunk_counts, uint32_t chunk_count,
                                  const uint64_t* d_hash, const uint32_t* d_len, const uint32_t* d_tag, uint32_t hash_type,
                                  uint32_t target_chunk_size, const void** out_buffer, uint64_t* out_size, const uint64_t* d_chunk_off = nullptr,
                                  lt_b200_comm* comm = nullptr, bool want_host = true)
{
    c->unique_chunks = 0;
    c->unique_offsets_valid = false;
    const uint32_t A = a->asset_count;
    // ---- per-asset segments: content hash over the asset's chunk-hash array (src/longtail.c:2518-2537) and path hash over
    // strlen(path) bytes (src/longtail.c:1281-1297).  Both are HashBuffer calls -> same segment kernel.
    TRY(hs_reserve(c, HS_SEG, (sizeof(uint64_t) + sizeof(uint32_t)) * 2 * (size_t)(A + 1) + sizeof(uint32_t) * (size_t)(A + 1)));
    uint64_t* h_off = hs<uint64_t>(c, HS_SEG);
    uint32_t* h_len = reinterpret_cast<uint32_t*>(h_off + 2 * (size_t)(A + 1));
    uint32_t* h_starts = h_len + 2 * (size_t)(A + 1);
    uint64_t run = 0;
    uint64_t upper = 2ull * A;
    uint32_t max_content = 0, max_path = 0;
    for (uint32_t i = 0; i < A; ++i)
    {
        h_starts[i] = (uint32_t)run;
        h_off[i] = run * 8;
        h_len[i] = asset_chunk_counts[i] * 8u;
        if (h_len[i] > max_content) max_content = h_len[i];
        upper += h_len[i] / 1024;
        run += asset_chunk_counts[i];
        const char* path = a->path_data + a->path_start_offsets[i];
        h_off[A + i] = a->path_start_offsets[i];
        h_len[A + i] = (uint32_t)strlen(path);
        if (h_len[A + i] > max_path) max_path = h_len[A + i];
        upper += h_len[A + i] / 1024;
    }
    if (run != chunk_count) return fail(c, EINVAL, "asset chunk counts sum to %llu, table has %u", (unsigned long long)run, chunk_count);

    const size_t path_bytes = ((size_t)a->path_data_size + 15) & ~(size_t)15;
    TRY(ws_reserve(c, WS_PATHS, path_bytes + 16));
    TRY(ws_reserve(c, WS_SEG_OFF, sizeof(uint64_t) * 2 * (size_t)(A + 1)));
    TRY(ws_reserve(c, WS_SEG_LEN, sizeof(uint32_t) * 2 * (size_t)(A + 1)));
    TRY(ws_reserve(c, WS_SEG_HASH, sizeof(uint64_t) * 2 * (size_t)(A + 1)));
    uint64_t* d_seg_hash = ws<uint64_t>(c, WS_SEG_HASH);
    if (A)
    {
        CU(cudaMemcpyAsync(ws<void>(c, WS_PATHS), a->path_data, a->path_data_size, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(ws<void>(c, WS_SEG_OFF), h_off, sizeof(uint64_t) * 2 * (size_t)A, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(ws<void>(c, WS_SEG_LEN), h_len, sizeof(uint32_t) * 2 * (size_t)A, cudaMemcpyHostToDevice, c->stream));
        // content hashes: base = the chunk-hash array itself (any valid pointer when there are no chunks at all)
        const uint8_t* hash_bytes = chunk_count ? reinterpret_cast<const uint8_t*>(d_hash) : ws<uint8_t>(c, WS_PATHS);
        TRY(hash_segments_device(c, hash_type, hash_bytes, 8ull * chunk_count, ws<uint64_t>(c, WS_SEG_OFF),
                                 ws<uint32_t>(c, WS_SEG_LEN), A, upper, WS_SEG_LEAF_COUNT, WS_SEG_LEAF_PREFIX, WS_SEG_CVS, d_seg_hash, 0, max_content));
        TRY(hash_segments_device(c, hash_type, ws<uint8_t>(c, WS_PATHS), a->path_data_size, ws<uint64_t>(c, WS_SEG_OFF) + A,
                                 ws<uint32_t>(c, WS_SEG_LEN) + A, A, upper, WS_SEG_LEAF_COUNT, WS_SEG_LEAF_PREFIX, WS_SEG_CVS, d_seg_hash + A, 0, max_path));
    }

    // ---- first-occurrence dedup (src/longtail.c:2952-2970)
    uint32_t unique = 0;
    if (chunk_count)
    {
        uint32_t cap = 1024;
        const uint64_t keys_here = comm && comm->world > 1 ? 2ull * (chunk_count / comm->world) + 4096 : chunk_count;
        while (cap < 2ull * keys_here && cap < 0x80000000u) cap <<= 1;
        TRY(ws_reserve(c, WS_DEDUP_KEYS, sizeof(uint64_t) * (size_t)cap));
        TRY(ws_reserve(c, WS_DEDUP_VALS, sizeof(uint32_t) * ((size_t)cap + 1)));
        TRY(ws_reserve(c, WS_DEDUP_FIRST, sizeof(uint32_t) * (size_t)chunk_count));
        TRY(ws_reserve(c, WS_DEDUP_ISFIRST, sizeof(uint32_t) * (size_t)chunk_count));
        TRY(ws_reserve(c, WS_DEDUP_UIDX, sizeof(uint32_t) * ((size_t)chunk_count + 1)));
        TRY(ws_reserve(c, WS_SCAN_TMP, sizeof(uint32_t) * scan_tmp_words(chunk_count)));
        TRY(ws_reserve(c, WS_ACI, sizeof(uint32_t) * (size_t)chunk_count));
        TRY(ws_reserve(c, WS_UHASH, sizeof(uint64_t) * (size_t)chunk_count));
        TRY(ws_reserve(c, WS_ULEN, sizeof(uint32_t) * (size_t)chunk_count));
        TRY(ws_reserve(c, WS_UTAG, sizeof(uint32_t) * (size_t)chunk_count));
        TRY(ws_reserve(c, WS_UOFF, sizeof(uint64_t) * (size_t)chunk_count));
        DedupBuffers db;
        db.keys = ws<uint64_t>(c, WS_DEDUP_KEYS);
        db.vals = ws<uint32_t>(c, WS_DEDUP_VALS);
        db.capacity = cap;
        db.first = ws<uint32_t>(c, WS_DEDUP_FIRST);
        db.is_first = ws<uint32_t>(c, WS_DEDUP_ISFIRST);
        db.uidx = ws<uint32_t>(c, WS_DEDUP_UIDX);
        CU(cudaMemsetAsync(db.keys, 0xff, sizeof(uint64_t) * (size_t)cap, c->stream));
        CU(cudaMemsetAsync(db.vals, 0xff, sizeof(uint32_t) * ((size_t)cap + 1), c->stream));
        if (comm && comm->world > 1)
        {
            // every rank holds the whole table; the random-access part of the dedup is split by hash and one all-reduce merges the answers
            launch_dedup_insert_part(d_hash, chunk_count, db, comm->world, comm->rank, c->stream);
            launch_dedup_lookup_part(d_hash, chunk_count, db, comm->world, comm->rank, c->stream);
            NC(g_nccl.AllReduce(db.first, db.first, chunk_count, ncclUint32, ncclSum, comm->comm, c->stream));
            launch_mark_first(db, chunk_count, c->stream);
        }
        else
        {
            launch_dedup_insert(d_hash, chunk_count, db, c->stream);
            launch_dedup_lookup(d_hash, chunk_count, db, c->stream);
        }
        launch_exclusive_scan(db.is_first, chunk_count, db.uidx, ws<uint32_t>(c, WS_SCAN_TMP), c->stream);
        TRY(ws_reserve(c, WS_UFIRST, sizeof(uint32_t) * (size_t)chunk_count));
        launch_dedup_emit(d_hash, d_len, d_tag, chunk_count, db, ws<uint32_t>(c, WS_ACI), ws<uint64_t>(c, WS_UHASH), ws<uint32_t>(c, WS_ULEN),
                          ws<uint32_t>(c, WS_UTAG), d_chunk_off, ws<uint64_t>(c, WS_UOFF), c->stream, ws<uint32_t>(c, WS_UFIRST));
        c->launches += 6;
        CU(cudaMemcpyAsync(&unique, db.uidx + chunk_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
    }

    // ---- serialised layout, assembled in device memory then copied out once (src/longtail.c:2566-2584)
    const size_t total = 24 + (size_t)A * (8 + 8 + 8 + 4 + 4 + 4 + 2) + 4 * (size_t)chunk_count + 16 * (size_t)unique + a->path_data_size;
    TRY(ws_reserve(c, WS_INDEX_OUT, total + 16));
    if (want_host) TRY(hs_reserve(c, HS_INDEX_OUT, total + 16));
    TRY(hs_reserve(c, HS_SMALL, 64));
    uint8_t* d_out = ws<uint8_t>(c, WS_INDEX_OUT);
    uint32_t* hdr = hs<uint32_t>(c, HS_SMALL);
    hdr[0] = 2; // Longtail_CurrentVersionIndexVersion, src/longtail.c:16-22
    hdr[1] = hash_type;
    hdr[2] = target_chunk_size;
    hdr[3] = A;
    hdr[4] = unique;
    hdr[5] = chunk_count;
    size_t o = 0;
    auto put_h = [&](const void* src, size_t n) -> cudaError_t {
        cudaError_t e = n ? cudaMemcpyAsync(d_out + o, src, n, cudaMemcpyHostToDevice, c->stream) : cudaSuccess;
        o += n;
        return e;
    };
    auto put_d = [&](const void* src, size_t n) -> cudaError_t {
        cudaError_t e = n ? cudaMemcpyAsync(d_out + o, src, n, cudaMemcpyDeviceToDevice, c->stream) : cudaSuccess;
        o += n;
        return e;
    };
    CU(put_h(hdr, 24));
    CU(put_d(d_seg_hash + A, 8 * (size_t)A));                       // m_PathHashes
    CU(put_d(d_seg_hash, 8 * (size_t)A));                           // m_ContentHashes
    CU(put_h(a->sizes, 8 * (size_t)A));                             // m_AssetSizes
    CU(put_h(asset_chunk_counts, 4 * (size_t)A));                   // m_AssetChunkCounts
    CU(put_h(h_starts, 4 * (size_t)A));                             // m_AssetChunkIndexStarts
    CU(put_d(ws<void>(c, WS_ACI), 4 * (size_t)chunk_count));        // m_AssetChunkIndexes
    CU(put_d(ws<void>(c, WS_UHASH), 8 * (size_t)unique));           // m_ChunkHashes
    CU(put_d(ws<void>(c, WS_ULEN), 4 * (size_t)unique));            // m_ChunkSizes
    CU(put_d(ws<void>(c, WS_UTAG), 4 * (size_t)unique));            // m_ChunkTags
    CU(put_h(a->path_start_offsets, 4 * (size_t)A));                // m_NameOffsets
    CU(put_h(a->permissions, 2 * (size_t)A));                       // m_Permissions
    CU(put_h(a->path_data, a->path_data_size));                     // m_NameData
    if (o != total) return fail(c, EFAULT, "index layout mismatch %zu != %zu", o, total);
    if (want_host) CU(cudaMemcpyAsync(hs<void>(c, HS_INDEX_OUT), d_out, total, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *out_buffer = want_host ? hs<void>(c, HS_INDEX_OUT) : nullptr;
    *out_size = total;
    c->unique_chunks = unique;
    c->unique_offsets_valid = d_chunk_off != nullptr;
    return 0;
}

int validate_assets(lt_b200_context* c, const lt_b200_assets* a)
{
    if (!a) return EINVAL;
    if (a->asset_count && (!a->sizes || !a->path_start_offsets || !a->permissions || !a->path_data)) return fail(c, EINVAL, "incomplete asset description");
    for (uint32_t i = 0; i < a->asset_count; ++i)
        if (a->path_start_offsets[i] >= a->path_data_size) return fail(c, EINVAL, "asset %u path offset outside path data", i);
    return 0;
}

void target_to_params(uint32_t t, uint32_t* mn, uint32_t* av, uint32_t* mx)
{
    // src/longtail.c:1985-1987 with ChunkerAPI.GetMinChunkSize() == 48 (lib/hpcdcchunker/longtail_hpcdcchunker.c:332-346)
    *mn = t / 8 < 48 ? 48 : t / 8;
    *av = t / 2 < 48 ? 48 : t / 2;
    *mx = (uint64_t)t * 2 < 48 ? 48 : t * 2;
}

} // namespace

extern "C" int lt_b200_build_version_index(lt_b200_context* c, const lt_b200_assets* a, const uint32_t* asset_chunk_counts,
                                           uint32_t chunk_count, const uint64_t* chunk_hashes, const uint32_t* chunk_sizes,
                                           const uint32_t* chunk_tags, uint32_t hash_type, uint32_t target_chunk_size,
                                           const void** out_buffer, uint64_t* out_size)
{
    if (!c || !out_buffer || !out_size) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    TRY(validate_assets(c, a));
    if (a->asset_count && !asset_chunk_counts) return EINVAL;
    if (!chunk_hashes)
    {
        if (chunk_count != c->table_chunks) return fail(c, EINVAL, "no resident chunk table of %u chunks (have %u)", chunk_count, c->table_chunks);
        return build_index_from_device_table(c, a, asset_chunk_counts, chunk_count, ws<uint64_t>(c, WS_CHUNK_HASH), ws<uint32_t>(c, WS_CHUNK_LEN),
                                             ws<uint32_t>(c, WS_CHUNK_TAG), hash_type, target_chunk_size, out_buffer, out_size);
    }
    if (chunk_count && (!chunk_sizes || !chunk_tags)) return EINVAL;
    TRY(ws_reserve(c, WS_TAB_HASH, sizeof(uint64_t) * (size_t)chunk_count + 16));
    TRY(ws_reserve(c, WS_TAB_LEN, sizeof(uint32_t) * (size_t)chunk_count + 16));
    TRY(ws_reserve(c, WS_TAB_TAG, sizeof(uint32_t) * (size_t)chunk_count + 16));
    if (chunk_count)
    {
        CU(cudaMemcpyAsync(ws<void>(c, WS_TAB_HASH), chunk_hashes, sizeof(uint64_t) * (size_t)chunk_count, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(ws<void>(c, WS_TAB_LEN), chunk_sizes, sizeof(uint32_t) * (size_t)chunk_count, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(ws<void>(c, WS_TAB_TAG), chunk_tags, sizeof(uint32_t) * (size_t)chunk_count, cudaMemcpyHostToDevice, c->stream));
    }
    return build_index_from_device_table(c, a, asset_chunk_counts, chunk_count, ws<uint64_t>(c, WS_TAB_HASH), ws<uint32_t>(c, WS_TAB_LEN),
                                         ws<uint32_t>(c, WS_TAB_TAG), hash_type, target_chunk_size, out_buffer, out_size);
}

extern "C" int lt_b200_resident_table(lt_b200_context* c, void** out_hashes, void** out_sizes, void** out_tags, uint32_t* out_count)
{
    if (!c || !out_hashes || !out_sizes || !out_tags || !out_count) return EINVAL;
    *out_hashes = ws<void>(c, WS_CHUNK_HASH);
    *out_sizes = ws<void>(c, WS_CHUNK_LEN);
    *out_tags = ws<void>(c, WS_CHUNK_TAG);
    *out_count = c->table_chunks;
    return 0;
}

extern "C" int lt_b200_build_version_index_device(lt_b200_context* c, const lt_b200_assets* a, const uint32_t* asset_chunk_counts,
                                                  uint32_t chunk_count, const void* d_hashes, const void* d_sizes, const void* d_tags,
                                                  uint32_t hash_type, uint32_t target_chunk_size, const void** out_buffer, uint64_t* out_size)
{
    if (!c || !out_buffer || !out_size) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    TRY(validate_assets(c, a));
    if (a->asset_count && !asset_chunk_counts) return EINVAL;
    if (chunk_count && (!d_hashes || !d_sizes || !d_tags)) return EINVAL;
    if (((uintptr_t)d_hashes) & 15u) return fail(c, EINVAL, "device hash array must be 16-byte aligned");
    return build_index_from_device_table(c, a, asset_chunk_counts, chunk_count, static_cast<const uint64_t*>(d_hashes),
                                         static_cast<const uint32_t*>(d_sizes), static_cast<const uint32_t*>(d_tags), hash_type,
                                         target_chunk_size, out_buffer, out_size);
}

extern "C" int lt_b200_index_device_assets(lt_b200_context* c, const uint8_t* d_arena, uint64_t arena_size, const lt_b200_assets* a,
                                           const uint64_t* asset_arena_offsets, const uint32_t* asset_tags, uint32_t hash_type,
                                           uint32_t target_chunk_size, const void** out_buffer, uint64_t* out_size)
{
    if (!c || !out_buffer || !out_size) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    TRY(validate_assets(c, a));
    if (a->asset_count && !asset_arena_offsets) return EINVAL;
    if (target_chunk_size == 0 || target_chunk_size > (1u << 20)) return fail(c, EINVAL, "target_chunk_size %u outside (0, 1 MiB]", target_chunk_size);
    uint32_t mn, av, mx;
    target_to_params(target_chunk_size, &mn, &av, &mx);
    const uint64_t part_size = (uint64_t)target_chunk_size * 1024; // src/longtail.c:2396
    std::vector<lt_b200_range> ranges;
    std::vector<uint32_t> first_range(a->asset_count + 1);
    for (uint32_t i = 0; i < a->asset_count; ++i)
    {
        first_range[i] = (uint32_t)ranges.size();
        const uint64_t size = a->sizes[i];
        const uint64_t parts = 1 + size / part_size; // :2402 — a trailing empty part produces no chunks and is skipped here
        for (uint64_t p = 0; p < parts; ++p)
        {
            const uint64_t start = p * part_size;
            const uint64_t n = size - start > part_size ? part_size : size - start;
            if (!n) continue;
            lt_b200_range r = {asset_arena_offsets[i] + start, (uint32_t)n, asset_tags ? asset_tags[i] : 0u};
            ranges.push_back(r);
        }
    }
    first_range[a->asset_count] = (uint32_t)ranges.size();
    lt_b200_chunk_table table;
    TRY(lt_b200_chunk_ranges(c, d_arena, arena_size, ranges.data(), (uint32_t)ranges.size(), mn, av, mx, hash_type, 0, &table));
    std::vector<uint32_t> asset_chunks(a->asset_count);
    for (uint32_t i = 0; i < a->asset_count; ++i)
    {
        uint32_t n = 0;
        for (uint32_t r = first_range[i]; r < first_range[i + 1]; ++r) n += table.range_chunk_counts[r];
        asset_chunks[i] = n;
    }
    return build_index_from_device_table(c, a, asset_chunks.data(), table.chunk_count, ws<uint64_t>(c, WS_CHUNK_HASH), ws<uint32_t>(c, WS_CHUNK_LEN),
                                         ws<uint32_t>(c, WS_CHUNK_TAG), hash_type, target_chunk_size, out_buffer, out_size,
                                         ws<uint64_t>(c, WS_CHUNK_OFF));
}

extern "C" int lt_b200_index_host_assets(lt_b200_context* c, const lt_b200_assets* a, const uint8_t* const* asset_data,
                                         const uint32_t* asset_tags, uint32_t hash_type, uint32_t target_chunk_size,
                                         const void** out_buffer, uint64_t* out_size)
{
    if (!c || !out_buffer || !out_size) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    TRY(validate_assets(c, a));
    if (a->asset_count && !asset_data) return EINVAL;
    if (target_chunk_size == 0 || target_chunk_size > (1u << 20)) return fail(c, EINVAL, "target_chunk_size %u outside (0, 1 MiB]", target_chunk_size);
    uint32_t mn, av, mx;
    target_to_params(target_chunk_size, &mn, &av, &mx);
    const uint64_t part_size = (uint64_t)target_chunk_size * 1024;

    // the job list of ChunkAssets (src/longtail.c:2399-2457), empty parts dropped
    struct Job { uint32_t asset; uint64_t start; uint32_t size; };
    std::vector<Job> jobs;
    uint64_t total_bytes = 0;
    for (uint32_t i = 0; i < a->asset_count; ++i)
    {
        const uint64_t size = a->sizes[i];
        const uint64_t parts = 1 + size / part_size;
        for (uint64_t p = 0; p < parts; ++p)
        {
            const uint64_t start = p * part_size;
            const uint64_t n = size - start > part_size ? part_size : size - start;
            if (n) jobs.push_back({i, start, (uint32_t)n});
        }
        total_bytes += size;
    }
    // batches of whole jobs, double-buffered device arenas: copy of batch k+1 overlaps the kernels of batch k
    const uint64_t batch_bytes = 1ull << 30 > part_size ? 1ull << 30 : part_size;
    const uint64_t arena_cap = batch_bytes + (uint64_t)jobs.size() * 0 + 4096;
    uint64_t max_chunks = total_bytes / mn + 2 * (uint64_t)jobs.size() + 16;
    TRY(ws_reserve(c, WS_ACC_HASH, sizeof(uint64_t) * (size_t)max_chunks));
    TRY(ws_reserve(c, WS_ACC_LEN, sizeof(uint32_t) * (size_t)max_chunks));
    TRY(ws_reserve(c, WS_ACC_TAG, sizeof(uint32_t) * (size_t)max_chunks));
    std::vector<uint32_t> asset_chunks(a->asset_count, 0);

    struct Batch { size_t first, last; std::vector<lt_b200_range> ranges; uint64_t bytes; };
    auto plan = [&](size_t first) {
        Batch b;
        b.first = first;
        b.bytes = 0;
        size_t j = first;
        while (j < jobs.size())
        {
            uint64_t padded = ((uint64_t)jobs[j].size + 255) & ~255ull;
            if (b.bytes + padded > batch_bytes && j > first) break;
            lt_b200_range r = {b.bytes, jobs[j].size, asset_tags ? asset_tags[jobs[j].asset] : 0u};
            b.ranges.push_back(r);
            b.bytes += padded;
            ++j;
        }
        b.last = j;
        return b;
    };
    auto upload = [&](const Batch& b, int which) -> int {
        TRY(ws_reserve(c, which ? WS_ARENA_B : WS_ARENA_A, arena_cap > b.bytes + 4096 ? arena_cap : b.bytes + 4096));
        uint8_t* d = ws<uint8_t>(c, which ? WS_ARENA_B : WS_ARENA_A);
        CU(cudaStreamWaitEvent(c->copy_stream, c->compute_done[which], 0)); // the arena's previous batch has been consumed
        for (size_t j = b.first; j < b.last; ++j)
            CU(cudaMemcpyAsync(d + b.ranges[j - b.first].arena_offset, asset_data[jobs[j].asset] + jobs[j].start, jobs[j].size,
                               cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(c->copy_done[which], c->copy_stream));
        return 0;
    };

    uint64_t acc = 0;
    if (!jobs.empty())
    {
        // make the events "signalled" for the first use
        CU(cudaEventRecord(c->compute_done[0], c->stream));
        CU(cudaEventRecord(c->compute_done[1], c->stream));
        Batch cur = plan(0);
        int which = 0;
        TRY(upload(cur, which));
        while (true)
        {
            Batch next;
            const bool has_next = cur.last < jobs.size();
            if (has_next)
            {
                next = plan(cur.last);
                TRY(upload(next, which ^ 1));
            }
            CU(cudaStreamWaitEvent(c->stream, c->copy_done[which], 0));
            const uint8_t* d = ws<uint8_t>(c, which ? WS_ARENA_B : WS_ARENA_A);
            lt_b200_chunk_table table;
            TRY(lt_b200_chunk_ranges(c, d, c->ws[which ? WS_ARENA_B : WS_ARENA_A].cap, cur.ranges.data(), (uint32_t)cur.ranges.size(), mn, av, mx,
                                     hash_type, 0, &table));
            CU(cudaEventRecord(c->compute_done[which], c->stream));
            if (acc + table.chunk_count > max_chunks) return fail(c, EFAULT, "chunk accumulation overflow");
            if (table.chunk_count)
            {
                CU(cudaMemcpyAsync(ws<uint64_t>(c, WS_ACC_HASH) + acc, ws<void>(c, WS_CHUNK_HASH), sizeof(uint64_t) * (size_t)table.chunk_count, cudaMemcpyDeviceToDevice, c->stream));
                CU(cudaMemcpyAsync(ws<uint32_t>(c, WS_ACC_LEN) + acc, ws<void>(c, WS_CHUNK_LEN), sizeof(uint32_t) * (size_t)table.chunk_count, cudaMemcpyDeviceToDevice, c->stream));
                CU(cudaMemcpyAsync(ws<uint32_t>(c, WS_ACC_TAG) + acc, ws<void>(c, WS_CHUNK_TAG), sizeof(uint32_t) * (size_t)table.chunk_count, cudaMemcpyDeviceToDevice, c->stream));
            }
            for (size_t j = cur.first; j < cur.last; ++j) asset_chunks[jobs[j].asset] += table.range_chunk_counts[j - cur.first];
            acc += table.chunk_count;
            if (!has_next) break;
            cur = std::move(next);
            which ^= 1;
        }
    }
    if (acc > 0xffffffffull) return fail(c, E2BIG, "more than 2^32 chunks");
    return build_index_from_device_table(c, a, asset_chunks.data(), (uint32_t)acc, ws<uint64_t>(c, WS_ACC_HASH), ws<uint32_t>(c, WS_ACC_LEN),
                                         ws<uint32_t>(c, WS_ACC_TAG), hash_type, target_chunk_size, out_buffer, out_size);
}

extern "C" int lt_b200_index_stream_assets(lt_b200_context* c, const lt_b200_assets* a, const uint32_t* asset_tags, uint32_t hash_type,
                                           uint32_t target_chunk_size, lt_b200_read_batch_func read_batch, void* user,
                                           const void** out_buffer, uint64_t* out_size)
{
    if (!c || !out_buffer || !out_size || !read_batch) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    TRY(validate_assets(c, a));
    if (target_chunk_size == 0 || target_chunk_size > (1u << 20)) return fail(c, EINVAL, "target_chunk_size %u outside (0, 1 MiB]", target_chunk_size);
    uint32_t mn, av, mx;
    target_to_params(target_chunk_size, &mn, &av, &mx);
    const uint64_t part_size = (uint64_t)target_chunk_size * 1024;

    struct Job { uint32_t asset; uint64_t start; uint32_t size; };
    std::vector<Job> jobs;
    uint64_t total_bytes = 0;
    for (uint32_t i = 0; i < a->asset_count; ++i)
    {
        const uint64_t size = a->sizes[i];
        const uint64_t parts = 1 + size / part_size; // src/longtail.c:2402
        for (uint64_t p = 0; p < parts; ++p)
        {
            const uint64_t start = p * part_size;
            const uint64_t n = size - start > part_size ? part_size : size - start;
            if (n) jobs.push_back({i, start, (uint32_t)n});
        }
        total_bytes += size;
    }
    const uint64_t batch_bytes = 1ull << 30 > part_size ? 1ull << 30 : part_size;
    const uint64_t max_chunks = total_bytes / mn + 2 * (uint64_t)jobs.size() + 16;
    TRY(ws_reserve(c, WS_ACC_HASH, sizeof(uint64_t) * (size_t)max_chunks));
    TRY(ws_reserve(c, WS_ACC_LEN, sizeof(uint32_t) * (size_t)max_chunks));
    TRY(ws_reserve(c, WS_ACC_TAG, sizeof(uint32_t) * (size_t)max_chunks));
    std::vector<uint32_t> asset_chunks(a->asset_count, 0);

    struct Batch { size_t first = 0, last = 0; std::vector<lt_b200_range> ranges; uint64_t bytes = 0; };
    auto plan = [&](size_t first) {
        Batch b;
        b.first = first;
        size_t j = first;
        while (j < jobs.size())
        {
            const uint64_t padded = ((uint64_t)jobs[j].size + 255) & ~255ull;
            if (b.bytes + padded > batch_bytes && j > first) break;
            b.ranges.push_back({b.bytes, jobs[j].size, asset_tags ? asset_tags[jobs[j].asset] : 0u});
            b.bytes += padded;
            ++j;
        }
        b.last = j;
        return b;
    };
    // host reads into pinned staging `which`, then the batch is queued for the device on the copy stream
    std::vector<lt_b200_read_job> read_jobs;
    auto read_and_upload = [&](const Batch& b, int which) -> int {
        const size_t need = (size_t)(batch_bytes > b.bytes ? batch_bytes : b.bytes) + 4096;
        TRY(hs_reserve(c, which ? HS_STAGE_B : HS_STAGE_A, need));
        TRY(ws_reserve(c, which ? WS_ARENA_B : WS_ARENA_A, need));
        uint8_t* h = hs<uint8_t>(c, which ? HS_STAGE_B : HS_STAGE_A);
        read_jobs.clear();
        for (size_t j = b.first; j < b.last; ++j)
            read_jobs.push_back({jobs[j].asset, jobs[j].size, jobs[j].start, h + b.ranges[j - b.first].arena_offset});
        int err = read_batch(user, read_jobs.data(), (uint32_t)read_jobs.size());
        if (err) return fail(c, err, "read callback failed with %d", err);
        CU(cudaStreamWaitEvent(c->copy_stream, c->compute_done[which], 0));
        CU(cudaMemcpyAsync(ws<uint8_t>(c, which ? WS_ARENA_B : WS_ARENA_A), h, b.bytes, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(c->copy_done[which], c->copy_stream));
        return 0;
    };

    uint64_t acc = 0;
    if (!jobs.empty())
    {
        CU(cudaEventRecord(c->compute_done[0], c->stream));
        CU(cudaEventRecord(c->compute_done[1], c->stream));
        Batch cur = plan(0);
        int which = 0;
        TRY(read_and_upload(cur, which));
        while (true)
        {
            Batch next;
            const bool has_next = cur.last < jobs.size();
            if (has_next)
            {
                next = plan(cur.last);
                TRY(read_and_upload(next, which ^ 1)); // host reads batch k+1 while batch k travels over PCIe
            }
            CU(cudaStreamWaitEvent(c->stream, c->copy_done[which], 0));
            lt_b200_chunk_table table;
            TRY(lt_b200_chunk_ranges(c, ws<uint8_t>(c, which ? WS_ARENA_B : WS_ARENA_A), c->ws[which ? WS_ARENA_B : WS_ARENA_A].cap, cur.ranges.data(),
                                     (uint32_t)cur.ranges.size(), mn, av, mx, hash_type, 0, &table));
            CU(cudaEventRecord(c->compute_done[which], c->stream));
            if (acc + table.chunk_count > max_chunks) return fail(c, EFAULT, "chunk accumulation overflow");
            if (table.chunk_count)
            {
                CU(cudaMemcpyAsync(ws<uint64_t>(c, WS_ACC_HASH) + acc, ws<void>(c, WS_CHUNK_HASH), sizeof(uint64_t) * (size_t)table.chunk_count, cudaMemcpyDeviceToDevice, c->stream));
                CU(cudaMemcpyAsync(ws<uint32_t>(c, WS_ACC_LEN) + acc, ws<void>(c, WS_CHUNK_LEN), sizeof(uint32_t) * (size_t)table.chunk_count, cudaMemcpyDeviceToDevice, c->stream));
                CU(cudaMemcpyAsync(ws<uint32_t>(c, WS_ACC_TAG) + acc, ws<void>(c, WS_CHUNK_TAG), sizeof(uint32_t) * (size_t)table.chunk_count, cudaMemcpyDeviceToDevice, c->stream));
            }
            for (size_t j = cur.first; j < cur.last; ++j) asset_chunks[jobs[j].asset] += table.range_chunk_counts[j - cur.first];
            acc += table.chunk_count;
            if (!has_next) break;
            cur = std::move(next);
            which ^= 1;
        }
    }
    if (acc > 0xffffffffull) return fail(c, E2BIG, "more than 2^32 chunks");
    return build_index_from_device_table(c, a, asset_chunks.data(), (uint32_t)acc, ws<uint64_t>(c, WS_ACC_HASH), ws<uint32_t>(c, WS_ACC_LEN),
                                         ws<uint32_t>(c, WS_ACC_TAG), hash_type, target_chunk_size, out_buffer, out_size);
}

// ================================================================ block build + compress (the WriteContent half)

extern "C" int lt_b200_unique_chunk_offsets(lt_b200_context* c, uint64_t* out_offsets, uint32_t count)
{
    if (!c || (count && !out_offsets)) return EINVAL;
    CU(cudaSetDevice(c->device));
    if (!c->unique_offsets_valid || count != c->unique_chunks)
        return fail(c, EINVAL, "no resident unique-chunk offsets for %u chunks (have %u, valid %d)", count, c->unique_chunks, (int)c->unique_offsets_valid);
    if (!count) return 0;
    CU(cudaMemcpyAsync(out_offsets, ws<void>(c, WS_UOFF), sizeof(uint64_t) * (size_t)count, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" uint64_t lt_b200_zstd_bound(uint64_t n) { return n + (n >> 8) + (n < (128u << 10) ? ((128u << 10) - n) >> 11 : 0); }

namespace {

bool is_zstd_level3(uint32_t tag) { return tag == LT_B200_COMPRESSION_ZSTD_DEFAULT || tag == LT_B200_COMPRESSION_ZSTD_MIN; }

// launch the ZStd frame encoder over frames laid out in device buffers; results land in WS_ZSTD_OUT_LEN (device)
int zstd_launch(lt_b200_context* c, const uint8_t* d_raw, const std::vector<uint64_t>& raw_off, const std::vector<uint32_t>& raw_len, uint8_t* d_out,
                const std::vector<uint64_t>& out_off, uint64_t payload_bytes)
{
    const uint32_t n = (uint32_t)raw_len.size();
    const uint32_t workers = zstd_worker_count(n, c->sm_count);
    TRY(ws_reserve(c, WS_ZSTD_WORKERS, zstd_worker_bytes() * (size_t)workers));
    TRY(ws_reserve(c, WS_ZSTD_RAW_OFF, sizeof(uint64_t) * (size_t)n));
    TRY(ws_reserve(c, WS_ZSTD_OUT_OFF, sizeof(uint64_t) * (size_t)n));
    TRY(ws_reserve(c, WS_ZSTD_RAW_LEN, sizeof(uint32_t) * (size_t)n));
    TRY(ws_reserve(c, WS_ZSTD_OUT_LEN, sizeof(uint32_t) * (size_t)n));
    TRY(ws_reserve(c, WS_QUEUE_HEAD, 256));
    CU(cudaMemcpyAsync(ws<void>(c, WS_ZSTD_RAW_OFF), raw_off.data(), sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(ws<void>(c, WS_ZSTD_OUT_OFF), out_off.data(), sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(ws<void>(c, WS_ZSTD_RAW_LEN), raw_len.data(), sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream)); // the host vectors belong to the caller's scope
    ProfScope ps(c, LT_B200_KERNEL_ZSTD, payload_bytes);
    CU(launch_zstd_frames(d_raw, ws<uint64_t>(c, WS_ZSTD_RAW_OFF), ws<uint32_t>(c, WS_ZSTD_RAW_LEN), d_out, ws<uint64_t>(c, WS_ZSTD_OUT_OFF),
                          ws<uint32_t>(c, WS_ZSTD_OUT_LEN), n, ws<void>(c, WS_ZSTD_WORKERS), workers, ws<uint32_t>(c, WS_QUEUE_HEAD), c->stream));
    c->launches += 1;
    return 0;
}

} // namespace

// DiffHashes of Longtail_CreateMissingContent (src/longtail.c:6620-6743, :6882-6998) on the device: which of the version's unique chunks
// does the store NOT hold yet.  The reference sorts both hash lists and merges them; here the store's hashes go into the dedup hash
// table and every version chunk probes it.  out_missing[i] = 1 when chunk_hashes[i] is absent from existing_hashes; the caller keeps the
// flagged chunks in their (version) order, which is the order the reference restores after its sort (:6719-6739).
extern "C" int lt_b200_missing_chunks(lt_b200_context* c, uint32_t chunk_count, const uint64_t* chunk_hashes, uint32_t existing_count,
                                      const uint64_t* existing_hashes, uint8_t* out_missing)
{
    if (!c || (chunk_count && (!chunk_hashes || !out_missing)) || (existing_count && !existing_hashes)) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    if (!chunk_count) return 0;
    if (!existing_count) { memset(out_missing, 1, chunk_count); return 0; }
    uint32_t cap = 1024;
    while (cap < 2ull * existing_count && cap < 0x80000000u) cap <<= 1;
    TRY(ws_reserve(c, WS_DEDUP_KEYS, sizeof(uint64_t) * (size_t)cap));
    TRY(ws_reserve(c, WS_DEDUP_VALS, sizeof(uint32_t) * ((size_t)cap + 1)));
    TRY(ws_reserve(c, WS_SEG_OFF, sizeof(uint64_t) * (size_t)(chunk_count > existing_count ? chunk_count : existing_count)));
    TRY(ws_reserve(c, WS_DEDUP_FIRST, (size_t)chunk_count + 16));
    DedupBuffers db;
    db.keys = ws<uint64_t>(c, WS_DEDUP_KEYS);
    db.vals = ws<uint32_t>(c, WS_DEDUP_VALS);
    db.capacity = cap;
    db.first = nullptr; db.is_first = nullptr; db.uidx = nullptr;
    CU(cudaMemsetAsync(db.keys, 0xff, sizeof(uint64_t) * (size_t)cap, c->stream));
    CU(cudaMemsetAsync(db.vals, 0xff, sizeof(uint32_t) * ((size_t)cap + 1), c->stream));
    CU(cudaMemcpyAsync(ws<void>(c, WS_SEG_OFF), existing_hashes, sizeof(uint64_t) * (size_t)existing_count, cudaMemcpyHostToDevice, c->stream));
    launch_dedup_insert(ws<uint64_t>(c, WS_SEG_OFF), existing_count, db, c->stream);
    CU(cudaMemcpyAsync(ws<void>(c, WS_SEG_OFF), chunk_hashes, sizeof(uint64_t) * (size_t)chunk_count, cudaMemcpyHostToDevice, c->stream));
    launch_set_contains(ws<uint64_t>(c, WS_SEG_OFF), chunk_count, db, ws<uint8_t>(c, WS_DEDUP_FIRST), c->stream);
    c->launches += 2;
    CU(cudaMemcpyAsync(out_missing, ws<void>(c, WS_DEDUP_FIRST), chunk_count, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    for (uint32_t i = 0; i < chunk_count; ++i) out_missing[i] = out_missing[i] ? 0 : 1; // present -> missing
    return 0;
}

// Longtail_CreateStoreIndex's greedy packing (src/longtail.c:6796-6860) as a host helper for planners (multi-GPU sharding by block)
extern "C" int lt_b200_pack_blocks(uint32_t chunk_count, const uint32_t* chunk_sizes, const uint32_t* chunk_tags, uint32_t max_block_size,
                                   uint32_t max_chunks_per_block, uint32_t* out_block_first, uint32_t* out_block_count, uint32_t* out_blocks)
{
    if ((chunk_count && (!chunk_sizes || !out_block_first || !out_block_count)) || !out_blocks || max_chunks_per_block == 0) return EINVAL;
    const uint64_t limit = (uint64_t)max_block_size + max_block_size / 10;
    uint32_t nb = 0;
    for (uint32_t i = 0; i < chunk_count;)
    {
        uint32_t count = 1;
        uint64_t raw = chunk_sizes[i];
        const uint32_t tag = chunk_tags ? chunk_tags[i] : 0u;
        while (i + count < chunk_count)
        {
            const uint32_t j = i + count;
            if ((chunk_tags ? chunk_tags[j] : 0u) != tag || count == max_chunks_per_block || raw + chunk_sizes[j] > limit) break;
            raw += chunk_sizes[j];
            ++count;
        }
        out_block_first[nb] = i;
        out_block_count[nb] = count;
        ++nb;
        i += count;
    }
    *out_blocks = nb;
    return 0;
}

namespace {
int write_blocks_impl(lt_b200_context* c, const uint8_t* d_arena, uint64_t arena_size, uint32_t chunk_count,
                      const uint64_t* chunk_hashes, const uint32_t* chunk_sizes, const uint32_t* chunk_tags,
                      const uint64_t* chunk_arena_offsets, uint32_t hash_type, uint32_t max_block_size,
                      uint32_t max_chunks_per_block, uint32_t given_block_count, const uint32_t* given_block_chunk_counts,
                      uint32_t flags, lt_b200_block_sink sink, void* user);
}

extern "C" int lt_b200_write_blocks_device(lt_b200_context* c, const uint8_t* d_arena, uint64_t arena_size, uint32_t chunk_count,
                                           const uint64_t* chunk_hashes, const uint32_t* chunk_sizes, const uint32_t* chunk_tags,
                                           const uint64_t* chunk_arena_offsets, uint32_t hash_type, uint32_t max_block_size,
                                           uint32_t max_chunks_per_block, lt_b200_block_sink sink, void* user)
{
    return write_blocks_impl(c, d_arena, arena_size, chunk_count, chunk_hashes, chunk_sizes, chunk_tags, chunk_arena_offsets, hash_type, max_block_size,
                             max_chunks_per_block, 0, nullptr, 0, sink, user);
}

extern "C" int lt_b200_write_blocks_device_ex(lt_b200_context* c, const uint8_t* d_arena, uint64_t arena_size, uint32_t chunk_count,
                                              const uint64_t* chunk_hashes, const uint32_t* chunk_sizes, const uint32_t* chunk_tags,
                                              const uint64_t* chunk_arena_offsets, uint32_t hash_type, uint32_t max_block_size,
                                              uint32_t max_chunks_per_block, uint32_t flags, lt_b200_block_sink sink, void* user)
{
    return write_blocks_impl(c, d_arena, arena_size, chunk_count, chunk_hashes, chunk_sizes, chunk_tags, chunk_arena_offsets, hash_type, max_block_size,
                             max_chunks_per_block, 0, nullptr, flags, sink, user);
}

// a sink that only counts: user = uint64_t[4] {blocks, stored bytes, raw payload bytes, xor of the block hashes}
extern "C" int lt_b200_counting_sink(void* user, const lt_b200_stored_block_view* block)
{
    uint64_t* acc = static_cast<uint64_t*>(user);
    if (!acc || !block) return EINVAL;
    acc[0] += 1;
    acc[1] += block->size;
    acc[2] += block->raw_payload_size;
    acc[3] ^= block->block_hash;
    return 0;
}

extern "C" int lt_b200_write_given_blocks_device(lt_b200_context* c, const uint8_t* d_arena, uint64_t arena_size, uint32_t chunk_count,
                                                 const uint64_t* chunk_hashes, const uint32_t* chunk_sizes, const uint32_t* chunk_tags,
                                                 const uint64_t* chunk_arena_offsets, uint32_t hash_type, uint32_t block_count,
                                                 const uint32_t* block_chunk_counts, lt_b200_block_sink sink, void* user)
{
    if (block_count && !block_chunk_counts) return EINVAL;
    uint64_t sum = 0;
    uint32_t most = 1;
    for (uint32_t b = 0; b < block_count; ++b)
    {
        if (!block_chunk_counts[b]) return EINVAL;
        sum += block_chunk_counts[b];
        if (block_chunk_counts[b] > most) most = block_chunk_counts[b];
    }
    if (sum != chunk_count) return EINVAL;
    return write_blocks_impl(c, d_arena, arena_size, chunk_count, chunk_hashes, chunk_sizes, chunk_tags, chunk_arena_offsets, hash_type, 0xffffffffu, most,
                             block_count, block_chunk_counts, 0, sink, user);
}

namespace {

// One batch of stored blocks in flight: its device buffers, its launch tables (one pinned host image, one device image) and the
// event that marks its kernels done.  With room for two of them the device -> host copies of batch k overlap the kernels of k + 1.
constexpr uint32_t WRITE_PIPELINED = 0x80000000u; // internal flag of write_blocks_impl, not part of the C ABI

struct WriteSlot
{
    int ws_raw, ws_out, ws_tab, ws_jobs, hs_tab, hs_len, ws_lz4_queue, ws_lz4_tables;
    cudaEvent_t done;
    cudaStream_t st; // every slot launches on its own stream: the codec kernel of batch k + 1 fills the warp slots batch k's slowest blocks leave idle
    // filled by launch
    uint32_t b0 = 0, nb = 0, n_lz = 0, n_zs = 0;
    std::vector<uint64_t> img_off;  // offset of each block's serialised image (block index + payload) in the out buffer
    std::vector<uint32_t> hdr_size;
    std::vector<uint32_t> lz_idx, zs_idx;
    bool busy = false;
};

int write_blocks_impl(lt_b200_context* c, const uint8_t* d_arena, uint64_t arena_size, uint32_t chunk_count,
                      const uint64_t* chunk_hashes, const uint32_t* chunk_sizes, const uint32_t* chunk_tags,
                      const uint64_t* chunk_arena_offsets, uint32_t hash_type, uint32_t max_block_size,
                      uint32_t max_chunks_per_block, uint32_t given_block_count, const uint32_t* given_block_chunk_counts,
                      uint32_t flags, lt_b200_block_sink sink, void* user)
{
    if (!c || !sink || (chunk_count && (!chunk_hashes || !chunk_sizes || !chunk_arena_offsets))) return EINVAL;
    if (max_chunks_per_block == 0) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    if (hash_type != LT_B200_HASH_BLAKE3 && hash_type != LT_B200_HASH_BLAKE2 && hash_type != LT_B200_HASH_MEOW)
        return fail(c, ENOTSUP, "hash type 0x%08x has no device implementation", hash_type);
    if (!chunk_count) return 0;
    const bool device_sink = (flags & LT_B200_WRITE_DEVICE_SINK) != 0;

    // ---- Longtail_CreateStoreIndex's greedy packing (src/longtail.c:6796-6860): in order; a block closes on a tag change, at
    // max_chunks_per_block chunks, or when the next chunk would exceed max_block_size + max_block_size/10
    struct Block { uint32_t first, count, raw, tag; };
    std::vector<Block> blocks;
    const uint64_t limit = (uint64_t)max_block_size + max_block_size / 10;
    uint32_t given = 0;
    for (uint32_t i = 0; i < chunk_count;)
    {
        Block b = {i, 1, chunk_sizes[i], chunk_tags ? chunk_tags[i] : 0u};
        if (b.tag != 0 && b.tag != LT_B200_COMPRESSION_LZ4 && !is_zstd_level3(b.tag))
            return fail(c, ENOTSUP, "compression type 0x%08x has no device implementation", b.tag);
        if (chunk_arena_offsets[i] > arena_size || chunk_sizes[i] > arena_size - chunk_arena_offsets[i]) return fail(c, EINVAL, "chunk %u lies outside the arena", i);
        const uint32_t want = given_block_chunk_counts ? given_block_chunk_counts[given++] : 0u; // the caller's blocks, as given (a store index)
        while (i + b.count < chunk_count)
        {
            const uint32_t j = i + b.count;
            if (given_block_chunk_counts)
            {
                if (b.count == want) break;
            }
            else
            {
                if ((chunk_tags ? chunk_tags[j] : 0u) != b.tag) break;
                if (b.count == max_chunks_per_block) break;
                if ((uint64_t)b.raw + chunk_sizes[j] > limit) break;
            }
            if (chunk_arena_offsets[j] > arena_size || chunk_sizes[j] > arena_size - chunk_arena_offsets[j]) return fail(c, EINVAL, "chunk %u lies outside the arena", j);
            if ((uint64_t)b.raw + chunk_sizes[j] > 0x7E000000ull) return fail(c, E2BIG, "block %zu too large", blocks.size()); // LZ4_MAX_INPUT_SIZE, u32 fields
            b.raw += chunk_sizes[j];
            ++b.count;
        }
        blocks.push_back(b);
        i += b.count;
    }
    const uint32_t nblocks = (uint32_t)blocks.size();

    // ---- block hashes = HashBuffer over each block's chunk-hash array (Longtail_CreateBlockIndex, src/longtail.c:3712-3770); the chunk
    // hashes and sizes stay on the device for the block indexes written in front of every payload
    TRY(ws_reserve(c, WS_BLK_HASHES, sizeof(uint64_t) * (size_t)chunk_count + 16));
    TRY(ws_reserve(c, WS_BLK_CHUNK_SIZES, sizeof(uint32_t) * (size_t)chunk_count + 16));
    TRY(ws_reserve(c, WS_BLK_SEG_OFF, sizeof(uint64_t) * (size_t)nblocks));
    TRY(ws_reserve(c, WS_BLK_SEG_LEN, sizeof(uint32_t) * (size_t)nblocks));
    TRY(ws_reserve(c, WS_BLK_HASH_OUT, sizeof(uint64_t) * (size_t)nblocks));
    TRY(hs_reserve(c, HS_BLK_META, (sizeof(uint64_t) * 2 + sizeof(uint32_t)) * (size_t)nblocks + 64));
    uint64_t* h_seg_off = hs<uint64_t>(c, HS_BLK_META);
    uint64_t* h_blk_hash = h_seg_off + nblocks;
    uint32_t* h_seg_len = reinterpret_cast<uint32_t*>(h_blk_hash + nblocks);
    uint64_t upper = nblocks;
    for (uint32_t b = 0; b < nblocks; ++b)
    {
        h_seg_off[b] = 8ull * blocks[b].first;
        h_seg_len[b] = 8u * blocks[b].count;
        upper += h_seg_len[b] / 1024;
    }
    CU(cudaMemcpyAsync(ws<void>(c, WS_BLK_HASHES), chunk_hashes, sizeof(uint64_t) * (size_t)chunk_count, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(ws<void>(c, WS_BLK_CHUNK_SIZES), chunk_sizes, sizeof(uint32_t) * (size_t)chunk_count, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(ws<void>(c, WS_BLK_SEG_OFF), h_seg_off, sizeof(uint64_t) * (size_t)nblocks, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(ws<void>(c, WS_BLK_SEG_LEN), h_seg_len, sizeof(uint32_t) * (size_t)nblocks, cudaMemcpyHostToDevice, c->stream));
    TRY(hash_segments_device(c, hash_type, ws<uint8_t>(c, WS_BLK_HASHES), 8ull * chunk_count, ws<uint64_t>(c, WS_BLK_SEG_OFF),
                             ws<uint32_t>(c, WS_BLK_SEG_LEN), nblocks, upper, WS_SEG_LEAF_COUNT, WS_SEG_LEAF_PREFIX, WS_SEG_CVS,
                             ws<uint64_t>(c, WS_BLK_HASH_OUT), 0, 8u * max_chunks_per_block));
    CU(cudaMemcpyAsync(h_blk_hash, ws<void>(c, WS_BLK_HASH_OUT), sizeof(uint64_t) * (size_t)nblocks, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));

    // ---- batches of blocks bounded by a device-memory budget: gather -> codec -> block indexes, all on the device; then every block's
    // serialised image (Longtail_WriteStoredBlockToBuffer, src/longtail.c:4111-4150) is handed to the sink in store order — as a device
    // address (LT_B200_WRITE_DEVICE_SINK) or after one device -> host copy into pinned staging
    auto lz4_bound = [](uint64_t n) { return n + n / 255 + 16; }; // lib/lz4/ext/lz4.h:215
    auto in_place = [](const Block& bl) { return bl.tag == LT_B200_COMPRESSION_LZ4 && lz4_block_in_place(bl.raw); };
    auto block_need = [&](const Block& bl, uint64_t* raw_need, uint64_t* out_need, uint64_t* job_need) {
        const uint64_t hdr = (20 + 12ull * bl.count + 15) & ~15ull;
        if (in_place(bl))
        {
            // [block index][8][8 pad .. output grows from here ..][gap][raw bytes][16]: see lz4_block_in_place
            *raw_need = 0;
            *out_need = hdr + ((16 + lz4_in_place_offset(bl.raw) + bl.raw + 16 + 15) & ~15ull);
            *job_need = 0;
            return;
        }
        const uint64_t pay = bl.tag ? 8 + (is_zstd_level3(bl.tag) ? lt_b200_zstd_bound(bl.raw) : lz4_bound(bl.raw)) : bl.raw;
        *raw_need = bl.tag ? (((uint64_t)bl.raw + 16 + 15) & ~15ull) : 0;
        *out_need = hdr + ((pay + 16 + 15) & ~15ull);
        *job_need = bl.tag == LT_B200_COMPRESSION_LZ4 ? sizeof(uint3) * (uint64_t)lz4_copy_job_capacity(bl.raw) : 0;
    };
    uint64_t need_all = 0, need_max = 0;
    for (const Block& bl : blocks)
    {
        uint64_t r, o, j;
        block_need(bl, &r, &o, &j);
        need_all += r + o + j;
        if (r + o + j > need_max) need_max = r + o + j;
    }
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    // what the grow-only workspace already holds for these buffers counts as available
    uint64_t held = 0;
    for (int w : {WS_BLK_RAW, WS_BLK_OUT, WS_BLK_JOBS, WS_BLK_RAW_B, WS_BLK_OUT_B, WS_BLK_JOBS_B, WS_BLK_RAW_C, WS_BLK_OUT_C, WS_BLK_JOBS_C, WS_BLK_RAW_D,
                  WS_BLK_OUT_D, WS_BLK_JOBS_D})
        held += c->ws[w].cap;
    uint64_t avail = (uint64_t)((free_b + held) * 0.85); // the rest: launch tables, v1 hash tables, allocator granularity
    // the codec kernels want a batch to fill every resident warp (13 x 148 LZ4 blocks); two batches in flight only when both can be that large
    // One batch when everything fits (the codec's work queue then balances all blocks over the resident warps).  Otherwise two batches in
    // flight on two streams: a codec launch ends on its slowest block — stored blocks differ 2x in parse time with their content — and the
    // next batch's kernels take the warp slots that fall idle meanwhile; with a host sink the device -> host copies of batch k also overlap
    // the kernels of k + 1.  ZStd batches keep one stream (its launcher owns shared worker slabs).
    bool any_zstd = false;
    for (const Block& bl : blocks) any_zstd = any_zstd || is_zstd_level3(bl.tag);
    uint32_t many = 4;
    if (const char* e = getenv("LT_B200_WRITE_SLOTS")) many = (uint32_t)std::min(4, std::max(1, atoi(e))); // A/B knob
    // WRITE_PIPELINED (the streaming upsync, whose sink is a device -> host copy): split even a batch that fits, so that the copies of the
    // first quarter start when its blocks are done instead of when the slowest block of the whole batch is
    const bool pipelined = (flags & WRITE_PIPELINED) != 0 && !any_zstd && !device_sink && nblocks >= 64;
    const uint32_t nslots = (need_all + (64ull << 20) > avail || pipelined) && !any_zstd ? many : 1u;
    uint64_t budget = avail / nslots;
    if (pipelined && budget > need_all / nslots + need_max) budget = need_all / nslots + need_max;
    if (budget > (64ull << 30)) budget = 64ull << 30;
    if (budget < need_max + (1ull << 20)) budget = need_max + (1ull << 20);

    WriteSlot slots[4] = {{WS_BLK_RAW, WS_BLK_OUT, WS_BLK_TAB, WS_BLK_JOBS, HS_BLK_TAB, HS_BLK_OUT_LEN, WS_LZ4_V2, WS_LZ4_TABLES, c->copy_done[0], c->stream},
                          {WS_BLK_RAW_B, WS_BLK_OUT_B, WS_BLK_TAB_B, WS_BLK_JOBS_B, HS_BLK_TAB_B, HS_BLK_OUT_LEN_B, WS_LZ4_V2_B, WS_LZ4_TABLES_B, c->copy_done[1],
                           c->aux_stream},
                          {WS_BLK_RAW_C, WS_BLK_OUT_C, WS_BLK_TAB_C, WS_BLK_JOBS_C, HS_BLK_TAB_C, HS_BLK_OUT_LEN_C, WS_LZ4_V2_C, WS_LZ4_TABLES_C, c->slot_done[0],
                           c->aux_stream2[0]},
                          {WS_BLK_RAW_D, WS_BLK_OUT_D, WS_BLK_TAB_D, WS_BLK_JOBS_D, HS_BLK_TAB_D, HS_BLK_OUT_LEN_D, WS_LZ4_V2_D, WS_LZ4_TABLES_D, c->slot_done[1],
                           c->aux_stream2[1]}};
    // the other streams start behind the block hashes / chunk tables uploaded above
    CU(cudaEventRecord(c->aux_fork, c->stream));
    for (uint32_t k = 1; k < nslots; ++k) CU(cudaStreamWaitEvent(slots[k].st, c->aux_fork, 0));

    // launch the kernels of blocks [b0, b1) into slot s
    auto launch = [&](WriteSlot& s, uint32_t b0, uint32_t b1) -> int {
        const uint32_t nb = b1 - b0;
        s.b0 = b0; s.nb = nb;
        s.img_off.assign(nb, 0); s.hdr_size.assign(nb, 0); s.lz_idx.clear(); s.zs_idx.clear();
        uint64_t raw_bytes = 0, out_bytes = 0, job_count = 0, gather_bytes = 0;
        uint32_t nc = 0;
        std::vector<uint64_t> raw_off(nb), pay_off(nb);
        for (uint32_t i = 0; i < nb; ++i)
        {
            const Block& bl = blocks[b0 + i];
            uint64_t r, o, j;
            block_need(bl, &r, &o, &j);
            const uint64_t hdr = 20 + 12ull * bl.count, hdr_pad = (hdr + 15) & ~15ull;
            raw_off[i] = raw_bytes;
            pay_off[i] = out_bytes + hdr_pad;       // 16-byte aligned; the block index sits right in front of it
            s.img_off[i] = pay_off[i] - hdr;
            s.hdr_size[i] = (uint32_t)hdr;
            raw_bytes += r; out_bytes += o; nc += bl.count;
            if (bl.tag) gather_bytes += bl.raw;
            if (bl.tag == LT_B200_COMPRESSION_LZ4) { s.lz_idx.push_back(i); job_count += in_place(bl) ? 0u : lz4_copy_job_capacity(bl.raw); }
            else if (bl.tag) s.zs_idx.push_back(i);
        }
        s.n_lz = (uint32_t)s.lz_idx.size(); s.n_zs = (uint32_t)s.zs_idx.size();
        // the LZ4 encoder prefetches the source a few batches of probes ahead: keep that inside the allocation
        TRY(ws_reserve(c, s.ws_raw, raw_bytes + (128u << 10)));
        TRY(ws_reserve(c, s.ws_out, out_bytes + (128u << 10)));
        TRY(ws_reserve(c, s.ws_jobs, sizeof(uint3) * (size_t)(job_count + 1)));
        uint8_t* d_raw = ws<uint8_t>(c, s.ws_raw);
        uint8_t* d_out = ws<uint8_t>(c, s.ws_out);
        // one table image: u64 src_off[nc] dst_abs[nc] hdr_dst[nb] lz_raw_off[nlz] lz_out_off[nlz] | u32 len[nc] blk_first[nb] blk_count[nb] blk_tag[nb]
        //                  lz_raw_len[nlz] lz_job_start[nlz] | outputs: u32 lz_out_len[nlz] lz_job_count[nlz]
        const size_t n64 = 2 * (size_t)nc + nb + 2 * (size_t)s.n_lz;
        const size_t n32 = (size_t)nc + 3 * (size_t)nb + 4 * (size_t)s.n_lz;
        const size_t tab_bytes = 8 * n64 + 4 * n32 + 64;
        TRY(ws_reserve(c, s.ws_tab, tab_bytes));
        TRY(hs_reserve(c, s.hs_tab, tab_bytes));
        uint64_t* h64 = hs<uint64_t>(c, s.hs_tab);
        uint32_t* h32 = reinterpret_cast<uint32_t*>(h64 + n64);
        uint64_t* d64 = ws<uint64_t>(c, s.ws_tab);
        uint32_t* d32 = reinterpret_cast<uint32_t*>(d64 + n64);
        uint64_t *h_src = h64, *h_dst = h64 + nc, *h_hdr = h64 + 2 * (size_t)nc, *h_lro = h_hdr + nb, *h_loo = h_lro + s.n_lz;
        uint32_t *h_len = h32, *h_first = h32 + nc, *h_count = h_first + nb, *h_tag = h_count + nb, *h_lrl = h_tag + nb, *h_ljs = h_lrl + s.n_lz;
        uint32_t ci = 0;
        for (uint32_t i = 0; i < nb; ++i)
        {
            const Block& bl = blocks[b0 + i];
            // compressed blocks gather into the raw buffer, stored-raw blocks straight into their place in the image
            uint64_t w = in_place(bl) ? (uint64_t)(uintptr_t)(d_out + pay_off[i] + 16 + lz4_in_place_offset(bl.raw))
                         : bl.tag     ? (uint64_t)(uintptr_t)(d_raw + raw_off[i])
                                      : (uint64_t)(uintptr_t)(d_out + pay_off[i]);
            for (uint32_t k = 0; k < bl.count; ++k, ++ci)
            {
                h_src[ci] = chunk_arena_offsets[bl.first + k];
                h_dst[ci] = w;
                h_len[ci] = chunk_sizes[bl.first + k];
                w += chunk_sizes[bl.first + k];
            }
            h_hdr[i] = s.img_off[i];
            h_first[i] = bl.first; h_count[i] = bl.count; h_tag[i] = bl.tag;
        }
        uint32_t job_at = 0;
        for (uint32_t q = 0; q < s.n_lz; ++q)
        {
            const uint32_t i = s.lz_idx[q];
            const Block& bl = blocks[b0 + i];
            // source addresses are absolute (the encoder's source base is null): in-place blocks read from their own output slot
            h_lro[q] = in_place(bl) ? (uint64_t)(uintptr_t)(d_out + pay_off[i] + 16 + lz4_in_place_offset(bl.raw)) : (uint64_t)(uintptr_t)(d_raw + raw_off[i]);
            h_loo[q] = pay_off[i]; h_lrl[q] = bl.raw;
            h_ljs[q] = in_place(bl) ? LZ4_JOBS_INLINE : job_at;
            if (!in_place(bl)) job_at += lz4_copy_job_capacity(bl.raw);
        }
        CU(cudaMemcpyAsync(d64, h64, 8 * n64 + 4 * (n32 - 2 * (size_t)s.n_lz), cudaMemcpyHostToDevice, s.st));
        {
            ProfScope ps(c, LT_B200_KERNEL_GATHER, gather_bytes, s.st);
            launch_gather_chunks(d_arena, d64, d64 + nc, d32, nullptr, nc, s.st);
        }
        uint32_t* d_lz_out_len = d32 + nc + 3 * (size_t)nb + 2 * (size_t)s.n_lz;
        if (s.n_lz)
        {
            uint64_t lz_bytes = 0;
            for (uint32_t q = 0; q < s.n_lz; ++q) lz_bytes += h_lrl[q];
            TRY(ws_reserve(c, s.ws_lz4_tables, LZ4_TABLE_BYTES_PER_BLOCK * (size_t)s.n_lz));
            TRY(ws_reserve(c, s.ws_lz4_queue, lz4_v2_scratch_bytes()));
            ProfScope ps(c, LT_B200_KERNEL_LZ4, lz_bytes, s.st);
            CU(launch_lz4_blocks(nullptr, d64 + 2 * (size_t)nc + nb, d32 + nc + 3 * (size_t)nb, d_out, d64 + 2 * (size_t)nc + nb + s.n_lz, d_lz_out_len,
                                 ws<uint3>(c, s.ws_jobs), d32 + nc + 3 * (size_t)nb + s.n_lz, d_lz_out_len + s.n_lz, s.n_lz,
                                 ws<uint32_t>(c, s.ws_lz4_tables), ws<void>(c, s.ws_lz4_queue), c->sm_count, s.st));
        }
        if (s.n_zs) // ZStd level 3 over the 'ztd1' / 'ztd2' blocks of the batch: one frame per block
        {
            std::vector<uint64_t> zro(s.n_zs), zoo(s.n_zs);
            std::vector<uint32_t> zrl(s.n_zs);
            uint64_t zs_bytes = 0;
            for (uint32_t q = 0; q < s.n_zs; ++q)
            {
                const uint32_t i = s.zs_idx[q];
                zro[q] = raw_off[i]; zoo[q] = pay_off[i]; zrl[q] = blocks[b0 + i].raw;
                zs_bytes += zrl[q];
            }
            TRY(zstd_launch(c, d_raw, zro, zrl, d_out, zoo, zs_bytes));
        }
        launch_block_headers(d_out, d64 + 2 * (size_t)nc, d32 + nc, d32 + nc + nb, d32 + nc + 2 * (size_t)nb, ws<uint64_t>(c, WS_BLK_HASH_OUT) + b0,
                             ws<uint64_t>(c, WS_BLK_HASHES), ws<uint32_t>(c, WS_BLK_CHUNK_SIZES), hash_type, nb, s.st);
        c->launches += 5;
        TRY(hs_reserve(c, s.hs_len, sizeof(uint32_t) * 2 * (size_t)nb + 16));
        uint32_t* h_out_len = hs<uint32_t>(c, s.hs_len);
        if (s.n_lz) CU(cudaMemcpyAsync(h_out_len, d_lz_out_len, sizeof(uint32_t) * s.n_lz, cudaMemcpyDeviceToHost, s.st));
        if (s.n_zs) CU(cudaMemcpyAsync(h_out_len + nb, ws<void>(c, WS_ZSTD_OUT_LEN), sizeof(uint32_t) * s.n_zs, cudaMemcpyDeviceToHost, s.st));
        CU(cudaEventRecord(s.done, s.st));
        s.busy = true;
        return 0;
    };

    // hand the finished blocks of slot s to the sink, in store order
    auto drain = [&](WriteSlot& s) -> int {
        if (!s.busy) return 0;
        s.busy = false;
        CU(cudaEventSynchronize(s.done));
        CU(cudaGetLastError());
        const uint32_t nb = s.nb, b0 = s.b0;
        const uint32_t* h_out_len = hs<uint32_t>(c, s.hs_len);
        std::vector<uint32_t> payload_len(nb);
        for (uint32_t i = 0; i < nb; ++i) payload_len[i] = blocks[b0 + i].raw;
        for (uint32_t q = 0; q < s.n_lz; ++q) payload_len[s.lz_idx[q]] = h_out_len[q];
        for (uint32_t q = 0; q < s.n_zs; ++q)
        {
            if (h_out_len[nb + q] == 0xffffffffu) return fail(c, EINVAL, "ZStd encoder failed on block %u", b0 + s.zs_idx[q]);
            payload_len[s.zs_idx[q]] = h_out_len[nb + q];
        }
        const uint8_t* d_out = ws<uint8_t>(c, s.ws_out);
        auto view_of = [&](uint32_t k, const void* data) {
            const Block& bl = blocks[b0 + k];
            lt_b200_stored_block_view v;
            v.block_hash = h_blk_hash[b0 + k];
            v.data = data;
            v.size = (uint64_t)s.hdr_size[k] + payload_len[k];
            v.chunk_count = bl.count;
            v.tag = bl.tag;
            v.raw_payload_size = bl.raw;
            v.first_chunk = bl.first;
            return v;
        };
        if (device_sink)
        {
            for (uint32_t k = 0; k < nb; ++k)
            {
                const lt_b200_stored_block_view v = view_of(k, d_out + s.img_off[k]);
                const int err = sink(user, &v);
                if (err) return fail(c, err, "block sink failed with %d", err);
            }
            return 0;
        }
        // pinned staging, two halves: while the sink consumes one half the copies into the other are in flight on the copy stream
        const uint64_t half_cap = 256ull << 20;
        uint64_t biggest = 0;
        for (uint32_t k = 0; k < nb; ++k) biggest = std::max<uint64_t>(biggest, (uint64_t)s.hdr_size[k] + payload_len[k]);
        const uint64_t half = std::max<uint64_t>(half_cap, (biggest + 63) & ~63ull);
        TRY(hs_reserve(c, HS_BLK_STAGE, 2 * half + 64));
        uint8_t* stage = hs<uint8_t>(c, HS_BLK_STAGE);
        struct Group { uint32_t i, j; std::vector<uint64_t> at; };
        auto issue = [&](uint32_t i, int which, Group* g) -> int {
            g->i = i; g->at.clear();
            uint64_t used = 0;
            uint32_t j = i;
            while (j < nb)
            {
                const uint64_t need = (((uint64_t)s.hdr_size[j] + payload_len[j]) + 63) & ~63ull;
                if (j > i && used + need > half) break;
                g->at.push_back(used);
                CU(cudaMemcpyAsync(stage + (size_t)which * half + used, d_out + s.img_off[j], (size_t)s.hdr_size[j] + payload_len[j], cudaMemcpyDeviceToHost,
                                   c->copy_stream));
                used += need;
                ++j;
            }
            g->j = j;
            CU(cudaEventRecord(c->compute_done[which], c->copy_stream));
            return 0;
        };
        Group g[2];
        int which = 0;
        TRY(issue(0, 0, &g[0]));
        while (g[which].i < nb)
        {
            Group& cur = g[which];
            if (cur.j < nb) TRY(issue(cur.j, which ^ 1, &g[which ^ 1])); else g[which ^ 1].i = g[which ^ 1].j = nb;
            CU(cudaEventSynchronize(c->compute_done[which]));
            for (uint32_t k = cur.i; k < cur.j; ++k)
            {
                const lt_b200_stored_block_view v = view_of(k, stage + (size_t)which * half + cur.at[k - cur.i]);
                const int err = sink(user, &v);
                if (err)
                {
                    cudaStreamSynchronize(c->copy_stream);
                    return fail(c, err, "block sink failed with %d", err);
                }
            }
            which ^= 1;
        }
        return 0;
    };

    // the batches, and every slot's buffers at the size of the largest one up front: a buffer that had to grow in the middle would stall
    // both streams
    std::vector<uint32_t> batch_end;
    {
        uint64_t max_raw = 0, max_out = 0, max_jobs = 0, max_tab = 0, max_lz = 0;
        for (uint32_t b0 = 0; b0 < nblocks;)
        {
            uint64_t used = 0, raw = 0, out = 0, jobs = 0, nc = 0, nlz = 0;
            uint32_t b1 = b0;
            while (b1 < nblocks)
            {
                uint64_t r, o, j;
                block_need(blocks[b1], &r, &o, &j);
                if (b1 > b0 && used + r + o + j > budget) break;
                used += r + o + j;
                raw += r; out += o; jobs += j; nc += blocks[b1].count;
                nlz += blocks[b1].tag == LT_B200_COMPRESSION_LZ4;
                ++b1;
            }
            const uint64_t nb = b1 - b0;
            max_raw = std::max(max_raw, raw); max_out = std::max(max_out, out); max_jobs = std::max(max_jobs, jobs); max_lz = std::max(max_lz, nlz);
            max_tab = std::max(max_tab, 8 * (2 * nc + nb + 2 * nlz) + 4 * (nc + 3 * nb + 4 * nlz) + 64);
            batch_end.push_back(b1);
            b0 = b1;
        }
        const uint32_t use = std::min<uint32_t>(nslots, (uint32_t)batch_end.size());
        for (uint32_t k = 0; k < use; ++k)
        {
            TRY(ws_reserve(c, slots[k].ws_raw, max_raw + (128u << 10)));
            TRY(ws_reserve(c, slots[k].ws_out, max_out + (128u << 10)));
            TRY(ws_reserve(c, slots[k].ws_jobs, max_jobs + sizeof(uint3)));
            TRY(ws_reserve(c, slots[k].ws_tab, max_tab));
            TRY(hs_reserve(c, slots[k].hs_tab, max_tab));
            if (max_lz)
            {
                TRY(ws_reserve(c, slots[k].ws_lz4_tables, LZ4_TABLE_BYTES_PER_BLOCK * (size_t)max_lz));
                TRY(ws_reserve(c, slots[k].ws_lz4_queue, lz4_v2_scratch_bytes()));
            }
        }
    }
    // batch k goes into slot k % nslots once batch k - nslots has been handed to the sink: nslots batches are in flight, and while the
    // host waits for the oldest one the codec kernels of the younger ones take the warp slots its slowest blocks leave idle
    int rc = 0;
    const uint32_t nbatches = (uint32_t)batch_end.size();
    for (uint32_t batch = 0, b0 = 0; batch < nbatches && !rc; ++batch)
    {
        const uint32_t b1 = batch_end[batch];
        WriteSlot& s = slots[batch % nslots];
        rc = drain(s);
        if (!rc) rc = launch(s, b0, b1);
        b0 = b1;
    }
    for (uint32_t k = 0; k < nslots && !rc; ++k) rc = drain(slots[(nbatches + k) % nslots]); // oldest first: store order
    if (rc)
    {
        cudaStreamSynchronize(c->stream);
        for (uint32_t k = 1; k < 4; ++k) cudaStreamSynchronize(slots[k].st);
        cudaStreamSynchronize(c->copy_stream);
        for (WriteSlot& s : slots) s.busy = false;
    }
    // whatever the other streams did is done (their batches were drained); later work on the context stream is ordered behind it anyway
    for (uint32_t k = 1; k < nslots; ++k)
    {
        cudaEventRecord(c->aux_join, slots[k].st);
        cudaStreamWaitEvent(c->stream, c->aux_join, 0);
    }
    return rc;
}
} // namespace

// ================================================================ streaming upsync: one pass over host-resident assets of any total size
//
// cmd/main.c:UpSync (:972-1153) without holding the version on the device.  Batches of whole parts travel host -> device into two
// arenas; while batch k + 1 is on the bus, batch k is chunked + hashed, its chunks are looked up in a host-side set of the hashes seen so
// far (first occurrence wins, src/longtail.c:2952-2970; the set starts with `existing_hashes` = Longtail_CreateMissingContent :7257-7340),
// the new chunks join the pending list, Longtail_CreateStoreIndex's greedy packing (:6796-6860) runs over that list and every block
// that can no longer change — all but the last — is compressed and handed to the sink.  The last block's chunks are carried (copied
// device -> device) into the next batch's arena.  The blocks and their order are those of the resident verbs: the packing is a left to
// right scan, so cutting the chunk list at block boundaries does not change it.  The VersionIndex is laid out at the end from the
// accumulated chunk table.  PCIe runs in both directions at once: uploads on their own stream, stored blocks leave on the copy stream.
extern "C" int lt_b200_upsync_stream_host_assets(lt_b200_context* c, const lt_b200_assets* a, const uint8_t* const* asset_data,
                                                 const uint32_t* asset_tags, uint32_t hash_type, uint32_t target_chunk_size,
                                                 uint32_t max_block_size, uint32_t max_chunks_per_block, uint32_t existing_count,
                                                 const uint64_t* existing_hashes, uint32_t flags, uint64_t batch_bytes, lt_b200_block_sink sink,
                                                 void* user, const void** out_version_index, uint64_t* out_size, uint32_t* out_chunks_written)
{
    if (!c || !sink || !out_version_index || !out_size || (existing_count && !existing_hashes)) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    TRY(validate_assets(c, a));
    if (a->asset_count && !asset_data) return EINVAL;
    if (max_chunks_per_block == 0) return EINVAL;
    if (target_chunk_size == 0 || target_chunk_size > (1u << 20)) return fail(c, EINVAL, "target_chunk_size %u outside (0, 1 MiB]", target_chunk_size);
    uint32_t mn, av, mx;
    target_to_params(target_chunk_size, &mn, &av, &mx);
    const uint64_t part_size = (uint64_t)target_chunk_size * 1024;

    // the job list of ChunkAssets (src/longtail.c:2399-2457), empty parts dropped
    struct Job { uint32_t asset; uint64_t start; uint32_t size; };
    std::vector<Job> jobs;
    uint64_t total_bytes = 0;
    for (uint32_t i = 0; i < a->asset_count; ++i)
    {
        const uint64_t size = a->sizes[i];
        if (size && !asset_data[i]) return fail(c, EINVAL, "asset %u has no data", i);
        for (uint64_t start = 0; start < size; start += part_size) jobs.push_back({i, start, (uint32_t)std::min<uint64_t>(part_size, size - start)});
        total_bytes += size;
    }
    if (batch_bytes == 0) batch_bytes = 16ull << 30;
    if (batch_bytes < part_size) batch_bytes = part_size;
    if (batch_bytes > total_bytes + 256ull * jobs.size() + 4096) batch_bytes = total_bytes + 256ull * jobs.size() + 4096;
    // room in front of every arena for the chunks of the block still open when the previous batch ended
    const uint64_t block_limit = (uint64_t)max_block_size + max_block_size / 10;
    const uint64_t carry_cap = (std::max<uint64_t>(block_limit, mx) + 16ull * max_chunks_per_block + 4096 + 255) & ~255ull;
    const uint64_t max_chunks = total_bytes / mn + 2 * (uint64_t)jobs.size() + 16;
    if (max_chunks > 0xffffffffull) return fail(c, E2BIG, "more than 2^32 chunks");
    TRY(ws_reserve(c, WS_ACC_HASH, sizeof(uint64_t) * (size_t)max_chunks));
    TRY(ws_reserve(c, WS_ACC_LEN, sizeof(uint32_t) * (size_t)max_chunks));
    TRY(ws_reserve(c, WS_ACC_TAG, sizeof(uint32_t) * (size_t)max_chunks));
    std::vector<uint32_t> asset_chunks(a->asset_count, 0);

    // the hashes seen so far: open addressing, linear probing; 0 is kept aside
    size_t set_cap = 1024;
    while (set_cap < 2 * (max_chunks + existing_count)) set_cap <<= 1;
    std::vector<uint64_t> set(set_cap, 0);
    bool seen_zero = false;
    auto insert = [&](uint64_t h) -> bool { // true: first time
        if (h == 0) { const bool first = !seen_zero; seen_zero = true; return first; }
        size_t at = (size_t)((h * 0x9E3779B97F4A7C15ull) >> 20) & (set_cap - 1);
        while (set[at]) { if (set[at] == h) return false; at = (at + 1) & (set_cap - 1); }
        set[at] = h;
        return true;
    };
    for (uint32_t i = 0; i < existing_count; ++i) insert(existing_hashes[i]);

    struct Batch { size_t first = 0, last = 0; std::vector<lt_b200_range> ranges; uint64_t bytes = 0; };
    auto plan = [&](size_t first) {
        Batch b;
        b.first = first;
        size_t j = first;
        while (j < jobs.size())
        {
            const uint64_t padded = ((uint64_t)jobs[j].size + 255) & ~255ull;
            if (b.bytes + padded > batch_bytes && j > first) break;
            b.ranges.push_back({b.bytes, jobs[j].size, asset_tags ? asset_tags[jobs[j].asset] : 0u});
            b.bytes += padded;
            ++j;
        }
        b.last = j;
        return b;
    };
    const int arena_slot[2] = {WS_ARENA_A, WS_ARENA_B};
    // Assets in pinned (device-mapped) host memory are read by a kernel (launch_upload_segments: the copy engine stays free for the small
    // table uploads of the kernels working on the previous batch); anything else goes through cudaMemcpyAsync.
    std::vector<const uint8_t*> mapped(a->asset_count, nullptr);
    if (!getenv("LT_B200_UPLOAD_MEMCPY")) // A/B knob
        for (uint32_t i = 0; i < a->asset_count; ++i)
        {
            cudaPointerAttributes attr;
            if (a->sizes[i] && cudaPointerGetAttributes(&attr, asset_data[i]) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer &&
                (((uintptr_t)attr.devicePointer) & 15u) == 0)
                mapped[i] = static_cast<const uint8_t*>(attr.devicePointer);
            else
                cudaGetLastError();
        }
    auto upload = [&](const Batch& b, int which) -> int {
        uint8_t* d = ws<uint8_t>(c, arena_slot[which]) + carry_cap;
        const size_t nj = b.last - b.first;
        TRY(hs_reserve(c, which ? HS_UPLOAD_SEGS_B : HS_UPLOAD_SEGS_A, sizeof(UploadSeg) * (nj + 1)));
        TRY(ws_reserve(c, which ? WS_UPLOAD_SEGS_B : WS_UPLOAD_SEGS_A, sizeof(UploadSeg) * (nj + 1)));
        UploadSeg* h_segs = hs<UploadSeg>(c, which ? HS_UPLOAD_SEGS_B : HS_UPLOAD_SEGS_A);
        UploadSeg* d_segs = ws<UploadSeg>(c, which ? WS_UPLOAD_SEGS_B : WS_UPLOAD_SEGS_A);
        CU(cudaStreamWaitEvent(c->upload_stream, c->arena_free[which], 0)); // the arena's previous batch has been consumed
        uint32_t nsegs = 0;
        for (size_t j = b.first; j < b.last; ++j)
        {
            uint8_t* dst = d + b.ranges[j - b.first].arena_offset;
            const uint8_t* m = mapped[jobs[j].asset];
            if (m && (jobs[j].start & 15u) == 0)
                h_segs[nsegs++] = {m + jobs[j].start, dst, jobs[j].size};
            else
                CU(cudaMemcpyAsync(dst, asset_data[jobs[j].asset] + jobs[j].start, jobs[j].size, cudaMemcpyHostToDevice, c->upload_stream));
        }
        if (nsegs)
        {
            CU(cudaMemcpyAsync(d_segs, h_segs, sizeof(UploadSeg) * nsegs, cudaMemcpyHostToDevice, c->upload_stream));
            launch_upload_segments(d_segs, nsegs, c->upload_stream);
            CU(cudaGetLastError());
            c->launches += 1;
        }
        CU(cudaEventRecord(c->upload_done[which], c->upload_stream));
        return 0;
    };

    // whatever way this call ends, no upload may still be reading the caller's memory afterwards
    struct UploadGuard { lt_b200_context* c; ~UploadGuard() { cudaStreamSynchronize(c->upload_stream); } } upload_guard{c};
    // the pending chunks: new to the store, not yet in a closed block
    std::vector<uint64_t> p_hash, p_addr;
    std::vector<uint32_t> p_size, p_tag, blk_first, blk_count;
    uint64_t acc = 0, written = 0;
    // LT_B200_TRACE=1: where the host thread's wall clock goes (waiting for the bus, chunk + hash, host dedup, pack + compress + sink)
    const bool trace = getenv("LT_B200_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_wait = 0, t_chunk = 0, t_dedup = 0, t_write = 0;
    const double t_start = now();
    if (!jobs.empty())
    {
        const uint64_t arena_bytes = carry_cap + batch_bytes + 4096;
        for (int w = 0; w < 2; ++w)
        {
            if (w == 1 && batch_bytes >= total_bytes) break; // one batch: one arena
            TRY(ws_reserve(c, arena_slot[w], arena_bytes));
        }
        CU(cudaEventRecord(c->arena_free[0], c->stream));
        CU(cudaEventRecord(c->arena_free[1], c->stream));
        Batch cur = plan(0);
        int which = 0;
        TRY(upload(cur, which));
        while (true)
        {
            Batch next;
            const bool has_next = cur.last < jobs.size();
            if (has_next)
            {
                next = plan(cur.last);
                TRY(ws_reserve(c, arena_slot[which ^ 1], arena_bytes));
                TRY(upload(next, which ^ 1));
            }
            CU(cudaStreamWaitEvent(c->stream, c->upload_done[which], 0));
            if (trace) { const double t = now(); CU(cudaEventSynchronize(c->upload_done[which])); t_wait += now() - t; }
            uint8_t* d_batch = ws<uint8_t>(c, arena_slot[which]) + carry_cap;
            lt_b200_chunk_table table;
            double t_mark = now();
            TRY(lt_b200_chunk_ranges(c, d_batch, batch_bytes + 4096, cur.ranges.data(), (uint32_t)cur.ranges.size(), mn, av, mx, hash_type, 1, &table));
            t_chunk += now() - t_mark; t_mark = now();
            if (acc + table.chunk_count > max_chunks) return fail(c, EFAULT, "chunk accumulation overflow");
            if (table.chunk_count)
            {
                CU(cudaMemcpyAsync(ws<uint64_t>(c, WS_ACC_HASH) + acc, ws<void>(c, WS_CHUNK_HASH), sizeof(uint64_t) * (size_t)table.chunk_count, cudaMemcpyDeviceToDevice, c->stream));
                CU(cudaMemcpyAsync(ws<uint32_t>(c, WS_ACC_LEN) + acc, ws<void>(c, WS_CHUNK_LEN), sizeof(uint32_t) * (size_t)table.chunk_count, cudaMemcpyDeviceToDevice, c->stream));
                CU(cudaMemcpyAsync(ws<uint32_t>(c, WS_ACC_TAG) + acc, ws<void>(c, WS_CHUNK_TAG), sizeof(uint32_t) * (size_t)table.chunk_count, cudaMemcpyDeviceToDevice, c->stream));
            }
            for (size_t j = cur.first; j < cur.last; ++j) asset_chunks[jobs[j].asset] += table.range_chunk_counts[j - cur.first];
            acc += table.chunk_count;
            for (uint32_t i = 0; i < table.chunk_count; ++i)
                if (insert(table.chunk_hashes[i]))
                {
                    p_hash.push_back(table.chunk_hashes[i]);
                    p_size.push_back(table.chunk_sizes[i]);
                    p_tag.push_back(table.chunk_tags[i]);
                    p_addr.push_back((uint64_t)(uintptr_t)(d_batch + table.chunk_offsets[i]));
                }
            t_dedup += now() - t_mark; t_mark = now();
            // close what can be closed
            const uint32_t pending = (uint32_t)p_hash.size();
            uint32_t nblk = 0;
            blk_first.resize(pending ? pending : 1); blk_count.resize(pending ? pending : 1);
            if (pending) TRY(lt_b200_pack_blocks(pending, p_size.data(), p_tag.data(), max_block_size, max_chunks_per_block, blk_first.data(), blk_count.data(), &nblk));
            const uint32_t closed_blocks = has_next && nblk ? nblk - 1 : nblk;
            const uint32_t closed_chunks = closed_blocks == nblk ? pending : blk_first[closed_blocks];
            if (closed_chunks)
            {
                TRY(write_blocks_impl(c, nullptr, ~0ull, closed_chunks, p_hash.data(), p_size.data(), p_tag.data(), p_addr.data(), hash_type, max_block_size,
                                      max_chunks_per_block, closed_blocks, blk_count.data(), flags | WRITE_PIPELINED, sink, user));
                written += closed_chunks;
            }
            t_write += now() - t_mark;
            // carry the open block's chunks into the front of the other arena
            const uint32_t open_chunks = pending - closed_chunks;
            if (open_chunks)
            {
                uint8_t* d_carry = ws<uint8_t>(c, arena_slot[which ^ 1]);
                uint64_t at = 0;
                for (uint32_t k = 0; k < open_chunks; ++k)
                {
                    const uint32_t i = closed_chunks + k;
                    if (at + p_size[i] > carry_cap) return fail(c, EFAULT, "open block exceeds the carry room");
                    CU(cudaMemcpyAsync(d_carry + at, (const void*)(uintptr_t)p_addr[i], p_size[i], cudaMemcpyDeviceToDevice, c->stream));
                    p_hash[k] = p_hash[i]; p_size[k] = p_size[i]; p_tag[k] = p_tag[i];
                    p_addr[k] = (uint64_t)(uintptr_t)(d_carry + at);
                    at += ((uint64_t)p_size[i] + 15) & ~15ull;
                }
            }
            p_hash.resize(open_chunks); p_size.resize(open_chunks); p_tag.resize(open_chunks); p_addr.resize(open_chunks);
            CU(cudaEventRecord(c->arena_free[which], c->stream));
            if (!has_next) break;
            cur = std::move(next);
            which ^= 1;
        }
    }
    if (out_chunks_written) *out_chunks_written = (uint32_t)written;
    const double t_loop = now();
    const int rc = build_index_from_device_table(c, a, asset_chunks.data(), (uint32_t)acc, ws<uint64_t>(c, WS_ACC_HASH), ws<uint32_t>(c, WS_ACC_LEN),
                                                 ws<uint32_t>(c, WS_ACC_TAG), hash_type, target_chunk_size, out_version_index, out_size);
    if (trace)
        fprintf(stderr, "lt_b200_upsync_stream_host_assets: %.1f ms: upload wait %.1f chunk+hash %.1f dedup %.1f pack+compress+sink %.1f index %.1f\n",
                1e3 * (now() - t_start), 1e3 * t_wait, 1e3 * t_chunk, 1e3 * t_dedup, 1e3 * t_write, 1e3 * (now() - t_loop));
    return rc;
}

// ================================================================ multi-GPU: one process per GPU, NCCL over NVLink
//
// The path shards by the reference's own job unit — one (asset, part) pair, parts of target_chunk_size * 1024 bytes, every part its own
// chunker instance (src/longtail.c:2396-2457) — so ranks chunk + hash disjoint, contiguous slices of the job list with no data-path
// collective.  One exchange follows: the per-job chunk counts and the (hash, size, tag) tables are all-gathered (variable sizes: one
// grouped set of broadcasts, no padding) so that every rank holds the table in global order; the first-occurrence dedup
// (:2952-2970) is split by hash over the ranks and merged with one all-reduce; then every rank lays out the same VersionIndex.
// WriteContent shards by stored block, balanced by bytes; the few chunks whose first occurrence lives on another rank move point to point.

namespace {

// LT_B200_TRACE=1: wall-clock phase times of the sharded verbs on stderr (rank 0)
struct PhaseTrace
{
    bool on;
    const char* verb;
    double t0, last;
    char line[512];
    size_t at = 0;
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    PhaseTrace(const char* v, uint32_t rank) : on(rank == 0 && getenv("LT_B200_TRACE") != nullptr), verb(v), t0(now()), last(t0) { line[0] = 0; }
    void mark(lt_b200_context* c, const char* what)
    {
        if (!on) return;
        cudaStreamSynchronize(c->stream);
        const double t = now();
        at += (size_t)snprintf(line + at, sizeof(line) - at, " %s %.1f", what, 1e3 * (t - last));
        if (at > sizeof(line) - 64) at = sizeof(line) - 64;
        last = t;
    }
    ~PhaseTrace() { if (on) fprintf(stderr, "%s: %.1f ms:%s\n", verb, 1e3 * (now() - t0), line); }
};

struct ShardJob { uint32_t asset; uint64_t offset; uint32_t size; };

// the reference's job list without its empty parts, in asset / part order
void shard_jobs(const lt_b200_assets* a, uint32_t target_chunk_size, std::vector<ShardJob>& jobs)
{
    const uint64_t part_size = (uint64_t)target_chunk_size * 1024;
    for (uint32_t i = 0; i < a->asset_count; ++i)
        for (uint64_t start = 0; start < a->sizes[i]; start += part_size)
        {
            ShardJob j = {i, start, (uint32_t)std::min<uint64_t>(part_size, a->sizes[i] - start)};
            jobs.push_back(j);
        }
}

// contiguous slices of the job list with (nearly) equal bytes: slice r ends where the running total first reaches (r + 1) / world of all
void shard_plan(const std::vector<ShardJob>& jobs, uint32_t world, std::vector<uint32_t>& first_job)
{
    uint64_t total = 0;
    for (const ShardJob& j : jobs) total += j.size;
    first_job.assign(world + 1, (uint32_t)jobs.size());
    first_job[0] = 0;
    uint64_t run = 0;
    uint32_t r = 1;
    for (uint32_t i = 0; i < jobs.size() && r < world; ++i)
    {
        run += jobs[i].size;
        while (r < world && run * world >= total * r) first_job[r++] = i + 1;
    }
}

} // namespace

extern "C" int lt_b200_comm_unique_id(uint8_t out_id[LT_B200_COMM_ID_BYTES])
{
    if (!out_id) return EINVAL;
    if (nccl_load()) return ENOSYS;
    static_assert(sizeof(ncclUniqueId) <= LT_B200_COMM_ID_BYTES, "ncclUniqueId does not fit LT_B200_COMM_ID_BYTES");
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return EIO;
    memset(out_id, 0, LT_B200_COMM_ID_BYTES);
    memcpy(out_id, &id, sizeof(id));
    return 0;
}

extern "C" int lt_b200_comm_create(lt_b200_context* c, const uint8_t id_bytes[LT_B200_COMM_ID_BYTES], uint32_t rank, uint32_t world, lt_b200_comm** out)
{
    if (!c || !id_bytes || !out || world == 0 || rank >= world) return EINVAL;
    *out = nullptr;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    if (nccl_load()) return fail(c, ENOSYS, "libnccl.so.2 could not be loaded");
    lt_b200_comm* m = new (std::nothrow) lt_b200_comm();
    if (!m) return ENOMEM;
    m->ctx = c;
    m->rank = rank;
    m->world = world;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    const ncclResult_t r = g_nccl.CommInitRank(&m->comm, (int)world, id, (int)rank);
    if (r != ncclSuccess)
    {
        delete m;
        return fail(c, EIO, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    }
    *out = m;
    return 0;
}

extern "C" void lt_b200_comm_destroy(lt_b200_comm* m)
{
    if (!m) return;
    if (m->comm)
    {
        cudaSetDevice(m->ctx->device);
        cudaStreamSynchronize(m->ctx->stream);
        g_nccl.CommDestroy(m->comm);
    }
    delete m;
}

extern "C" int lt_b200_plan_shards(const lt_b200_assets* a, uint32_t target_chunk_size, uint32_t world, uint32_t* out_first_job, uint32_t* out_job_count)
{
    if (!a || !out_first_job || world == 0 || target_chunk_size == 0) return EINVAL;
    std::vector<ShardJob> jobs;
    shard_jobs(a, target_chunk_size, jobs);
    std::vector<uint32_t> first;
    shard_plan(jobs, world, first);
    memcpy(out_first_job, first.data(), sizeof(uint32_t) * (world + 1));
    if (out_job_count) *out_job_count = (uint32_t)jobs.size();
    return 0;
}

extern "C" int lt_b200_shard_jobs(const lt_b200_assets* a, uint32_t target_chunk_size, uint32_t first_job, uint32_t job_count, lt_b200_shard_job* out_jobs)
{
    if (!a || (job_count && !out_jobs) || target_chunk_size == 0) return EINVAL;
    std::vector<ShardJob> jobs;
    shard_jobs(a, target_chunk_size, jobs);
    if ((uint64_t)first_job + job_count > jobs.size()) return EINVAL;
    for (uint32_t i = 0; i < job_count; ++i)
    {
        out_jobs[i].asset_index = jobs[first_job + i].asset;
        out_jobs[i].size = jobs[first_job + i].size;
        out_jobs[i].offset = jobs[first_job + i].offset;
    }
    return 0;
}

extern "C" int lt_b200_index_sharded(lt_b200_context* c, lt_b200_comm* m, const uint8_t* d_arena, uint64_t arena_size, const lt_b200_assets* a,
                                     const uint32_t* asset_tags, const uint64_t* job_arena_offsets, uint32_t hash_type,
                                     uint32_t target_chunk_size, int want_host, const void** out_buffer, uint64_t* out_size)
{
    if (!c || !m || m->ctx != c || !out_buffer || !out_size) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    TRY(validate_assets(c, a));
    if (target_chunk_size == 0 || target_chunk_size > (1u << 20)) return fail(c, EINVAL, "target_chunk_size %u outside (0, 1 MiB]", target_chunk_size);
    uint32_t mn, av, mx;
    target_to_params(target_chunk_size, &mn, &av, &mx);
    const uint32_t world = m->world, rank = m->rank;
    std::vector<ShardJob> jobs;
    shard_jobs(a, target_chunk_size, jobs);
    std::vector<uint32_t> first_job;
    shard_plan(jobs, world, first_job);
    const uint32_t total_jobs = (uint32_t)jobs.size(), j0 = first_job[rank], j1 = first_job[rank + 1];
    if (j1 > j0 && !job_arena_offsets) return EINVAL;
    PhaseTrace trace("lt_b200_index_sharded", rank);
    trace.mark(c, "plan");

    // ---- this rank's slice: chunk + hash, table left resident
    std::vector<lt_b200_range> ranges(j1 - j0);
    for (uint32_t j = j0; j < j1; ++j)
    {
        lt_b200_range r = {job_arena_offsets[j - j0], jobs[j].size, asset_tags ? asset_tags[jobs[j].asset] : 0u};
        ranges[j - j0] = r;
    }
    lt_b200_chunk_table table;
    TRY(lt_b200_chunk_ranges(c, d_arena, arena_size, ranges.data(), (uint32_t)ranges.size(), mn, av, mx, hash_type, 0, &table));
    trace.mark(c, "chunk+hash");

    // ---- exchange 1: chunk count of every job (each rank contributes its slice)
    TRY(ws_reserve(c, WS_G_COUNTS, sizeof(uint32_t) * ((size_t)total_jobs + 1)));
    uint32_t* g_counts = ws<uint32_t>(c, WS_G_COUNTS);
    std::vector<uint32_t> all_counts(total_jobs + 1, 0);
    if (world > 1)
    {
        if (j1 > j0) CU(cudaMemcpyAsync(g_counts + j0, table.range_chunk_counts, sizeof(uint32_t) * (j1 - j0), cudaMemcpyHostToDevice, c->stream));
        NC(g_nccl.GroupStart());
        for (uint32_t r = 0; r < world; ++r)
            if (first_job[r + 1] > first_job[r])
                NC(g_nccl.Broadcast(g_counts + first_job[r], g_counts + first_job[r], first_job[r + 1] - first_job[r], ncclUint32, (int)r, m->comm, c->stream));
        NC(g_nccl.GroupEnd());
        CU(cudaMemcpyAsync(all_counts.data(), g_counts, sizeof(uint32_t) * total_jobs, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    else
        for (uint32_t j = 0; j < total_jobs; ++j) all_counts[j] = table.range_chunk_counts[j];
    std::vector<uint32_t> asset_chunks(a->asset_count, 0);
    m->rank_chunk_start.assign(world + 1, 0);
    uint64_t run = 0;
    for (uint32_t r = 0, j = 0; r < world; ++r)
    {
        m->rank_chunk_start[r] = (uint32_t)run;
        for (; j < first_job[r + 1]; ++j)
        {
            run += all_counts[j];
            asset_chunks[jobs[j].asset] += all_counts[j];
        }
    }
    if (run >= 0xffffffffull) return fail(c, E2BIG, "%llu chunks exceed the VersionIndex's 32-bit indexes", (unsigned long long)run);
    const uint32_t N = (uint32_t)run;
    m->rank_chunk_start[world] = N;
    if (m->rank_chunk_start[rank + 1] - m->rank_chunk_start[rank] != table.chunk_count) return fail(c, EFAULT, "chunk count exchange out of step");

    // ---- exchange 2: the (hash, size, tag) tables, every rank's slice broadcast into place
    const uint64_t* g_hash = ws<uint64_t>(c, WS_CHUNK_HASH);
    const uint32_t* g_len = ws<uint32_t>(c, WS_CHUNK_LEN);
    const uint32_t* g_tag = ws<uint32_t>(c, WS_CHUNK_TAG);
    if (world > 1)
    {
        TRY(ws_reserve(c, WS_G_HASH, sizeof(uint64_t) * (size_t)N + 16));
        TRY(ws_reserve(c, WS_G_LEN, sizeof(uint32_t) * (size_t)N + 16));
        TRY(ws_reserve(c, WS_G_TAG, sizeof(uint32_t) * (size_t)N + 16));
        NC(g_nccl.GroupStart());
        for (uint32_t r = 0; r < world; ++r)
        {
            const uint32_t s0 = m->rank_chunk_start[r], n = m->rank_chunk_start[r + 1] - s0;
            if (!n) continue;
            NC(g_nccl.Broadcast(ws<uint64_t>(c, WS_CHUNK_HASH), ws<uint64_t>(c, WS_G_HASH) + s0, n, ncclUint64, (int)r, m->comm, c->stream));
            NC(g_nccl.Broadcast(ws<uint32_t>(c, WS_CHUNK_LEN), ws<uint32_t>(c, WS_G_LEN) + s0, n, ncclUint32, (int)r, m->comm, c->stream));
            NC(g_nccl.Broadcast(ws<uint32_t>(c, WS_CHUNK_TAG), ws<uint32_t>(c, WS_G_TAG) + s0, n, ncclUint32, (int)r, m->comm, c->stream));
        }
        NC(g_nccl.GroupEnd());
        g_hash = ws<uint64_t>(c, WS_G_HASH);
        g_len = ws<uint32_t>(c, WS_G_LEN);
        g_tag = ws<uint32_t>(c, WS_G_TAG);
        c->launches += 2;
    }

    trace.mark(c, "allgather");
    // ---- the one VersionIndex, on every rank (the host copy only where it is wanted)
    TRY(build_index_from_device_table(c, a, asset_chunks.data(), N, g_hash, g_len, g_tag, hash_type, target_chunk_size, out_buffer, out_size, nullptr, m,
                                      want_host != 0));
    // where the unique chunks' first occurrences live: rank r holds unique chunks [rank_unique_start[r], rank_unique_start[r + 1])
    m->rank_unique_start.assign(world + 1, c->unique_chunks);
    for (uint32_t r = 0; r < world; ++r)
        if (m->rank_chunk_start[r] < N)
            CU(cudaMemcpyAsync(&m->rank_unique_start[r], ws<uint32_t>(c, WS_DEDUP_UIDX) + m->rank_chunk_start[r], sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    trace.mark(c, "dedup+layout");
    m->global_chunks = N;
    m->global_unique = c->unique_chunks;
    m->d_arena = d_arena;
    m->arena_size = arena_size;
    m->hash_type = hash_type;
    return 0;
}

extern "C" int lt_b200_write_blocks_sharded(lt_b200_context* c, lt_b200_comm* m, uint32_t max_block_size, uint32_t max_chunks_per_block, uint32_t flags,
                                            lt_b200_block_sink sink, void* user, uint32_t* out_my_blocks, uint32_t* out_total_blocks)
{
    if (!c || !m || m->ctx != c || !sink || max_chunks_per_block == 0) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    if (m->rank_unique_start.size() != m->world + 1) return fail(c, EINVAL, "lt_b200_index_sharded has not run on this communicator");
    const uint32_t world = m->world, rank = m->rank, U = m->global_unique;
    if (out_my_blocks) *out_my_blocks = 0;
    if (out_total_blocks) *out_total_blocks = 0;
    if (!U) return 0;

    // ---- the same plan on every rank: Longtail_CreateStoreIndex's packing over the unique chunks in VersionIndex order
    // (src/longtail.c:6796-6860) — on the device, from the unique table the index build left there — then contiguous runs of blocks per
    // rank with (nearly) equal payload bytes.  Only the block list (one entry per ~270 chunks) and this rank's own slice come to the host.
    PhaseTrace trace("lt_b200_write_blocks_sharded", rank);
    const uint64_t limit = (uint64_t)max_block_size + max_block_size / 10;
    const uint32_t* d_ulen = ws<uint32_t>(c, WS_ULEN);
    const uint32_t* d_utag = ws<uint32_t>(c, WS_UTAG);
    PackBuffers pb;
    {
        const size_t n1 = (size_t)U + 1, t32 = scan_tmp_words(U + 1) + 2, t64 = scan64_tmp_words(U) + 2;
        TRY(ws_reserve(c, WS_PACK_A, 8 * (2 * n1 + t64) + 4 * (9 * n1 + t32) + 256));
        uint64_t* q = ws<uint64_t>(c, WS_PACK_A);
        pb.prefix = q; q += n1;
        pb.blk_end = q; q += n1;
        pb.tmp64 = q; q += t64;
        uint32_t* w = reinterpret_cast<uint32_t*>(q);
        pb.flag = w; w += n1;
        pb.rscan = w; w += n1;
        pb.run_start = w; w += n1;
        pb.next = w; w += n1;
        pb.jump_a = w; w += n1;
        pb.jump_b = w; w += n1;
        pb.mark = w; w += n1;
        pb.mscan = w; w += n1;
        pb.blk_first = w; w += n1;
        pb.tmp32 = w;
    }
    launch_pack_blocks(d_ulen, d_utag, U, limit, max_chunks_per_block, pb, c->stream);
    c->launches += 40;
    uint32_t B = 0;
    CU(cudaMemcpyAsync(&B, pb.mscan + U, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    std::vector<uint32_t> blk_first((size_t)B + 1);
    std::vector<uint64_t> blk_end_bytes(B);
    CU(cudaMemcpyAsync(blk_first.data(), pb.blk_first, sizeof(uint32_t) * (size_t)B, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(blk_end_bytes.data(), pb.blk_end, sizeof(uint64_t) * (size_t)B, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    blk_first[B] = U;
    trace.mark(c, "pack");
    const uint64_t total_bytes = B ? blk_end_bytes[B - 1] : 0;
    std::vector<uint32_t> rank_blk(world + 1, B);
    rank_blk[0] = 0;
    for (uint32_t b = 0, r = 1; b < B && r < world; ++b)
        while (r < world && blk_end_bytes[b] * world >= total_bytes * r) rank_blk[r++] = b + 1;
    // A rank's blocks should mostly be made of chunks it holds: every byte that is not has to cross NVLink into a receive buffer that takes
    // device memory away from the block batches.  The first occurrences are not spread evenly (the rank that indexed the shared content
    // first holds more unique bytes), so the byte-balanced boundaries are pulled back towards the ownership boundaries until no rank
    // imports more than `cap` bytes per boundary: a little imbalance instead of gigabytes of exchange.
    {
        uint64_t cap = 1ull << 30;
        if (const char* e = getenv("LT_B200_EXCHANGE_CAP_MB")) cap = (uint64_t)strtoull(e, nullptr, 10) << 20;
        auto bytes_before = [&](uint32_t b) { return b ? blk_end_bytes[b - 1] : 0ull; }; // payload bytes of blocks 0 .. b-1
        for (uint32_t r = 1; r < world; ++r)
        {
            const uint32_t own_b = (uint32_t)(std::lower_bound(blk_first.begin(), blk_first.begin() + B, m->rank_unique_start[r]) - blk_first.begin());
            uint32_t b = rank_blk[r];
            if (b > own_b)
                while (b > own_b && bytes_before(b) - bytes_before(own_b) > cap) --b; // rank r-1 imports blocks own_b .. b-1 from rank r
            else
                while (b < own_b && bytes_before(own_b) - bytes_before(b) > cap) ++b; // rank r imports blocks b .. own_b-1 from rank r-1
            rank_blk[r] = std::max(b, rank_blk[r - 1]);
        }
    }
    auto chunk_lo = [&](uint32_t r) { return blk_first[rank_blk[r]]; }; // first unique chunk rank r writes (U past the last block)
    if (out_total_blocks) *out_total_blocks = B;
    const uint32_t my_b0 = rank_blk[rank], my_b1 = rank_blk[rank + 1];
    const uint32_t cs = chunk_lo(rank), ce = chunk_lo(rank + 1);

    // ---- exchange: owner o holds unique chunks [us_o, ue_o); writer w needs [cs_w, ce_w).  Local offsets of my own first occurrences:
    const uint32_t my_c0 = m->rank_chunk_start[rank], my_cn = m->rank_chunk_start[rank + 1] - my_c0;
    const uint32_t us = m->rank_unique_start[rank], ue = m->rank_unique_start[rank + 1];
    std::vector<uint64_t> local_off(my_cn);
    std::vector<uint32_t> u_first(ue > us ? ue - us : 0);
    std::vector<uint32_t> own_len(ue > us ? ue - us : 0);         // sizes of the unique chunks this rank holds   [us, ue)
    std::vector<uint32_t> my_len(ce > cs ? ce - cs : 0), my_tag(ce > cs ? ce - cs : 0); // ... and of the ones it writes [cs, ce)
    if (my_cn) CU(cudaMemcpyAsync(local_off.data(), ws<void>(c, WS_CHUNK_OFF), sizeof(uint64_t) * (size_t)my_cn, cudaMemcpyDeviceToHost, c->stream));
    if (ue > us)
    {
        CU(cudaMemcpyAsync(u_first.data(), ws<uint32_t>(c, WS_UFIRST) + us, sizeof(uint32_t) * (size_t)(ue - us), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(own_len.data(), d_ulen + us, sizeof(uint32_t) * (size_t)(ue - us), cudaMemcpyDeviceToHost, c->stream));
    }
    if (ce > cs)
    {
        CU(cudaMemcpyAsync(my_len.data(), d_ulen + cs, sizeof(uint32_t) * (size_t)(ce - cs), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(my_tag.data(), d_utag + cs, sizeof(uint32_t) * (size_t)(ce - cs), cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    auto off_of = [&](uint32_t u) { return local_off[u_first[u - us] - my_c0]; }; // arena offset of unique chunk u (mine)

    struct Seg { uint32_t peer, a, b; uint64_t bytes, at; };
    std::vector<Seg> sends, recvs;
    uint64_t send_bytes = 0, recv_bytes = 0;
    for (uint32_t w = 0; w < world; ++w)
    {
        if (w == rank) continue;
        const uint32_t a = std::max(chunk_lo(w), us), b = std::min(chunk_lo(w + 1), ue); // what writer w needs from me
        if (a < b)
        {
            Seg s = {w, a, b, 0, send_bytes};
            for (uint32_t u = a; u < b; ++u) s.bytes += own_len[u - us];
            send_bytes += (s.bytes + 15) & ~15ull;
            sends.push_back(s);
        }
        const uint32_t ra = std::max(cs, m->rank_unique_start[w]), rb = std::min(ce, m->rank_unique_start[w + 1]); // what I need from owner w
        if (ra < rb)
        {
            Seg s = {w, ra, rb, 0, recv_bytes};
            for (uint32_t u = ra; u < rb; ++u) s.bytes += my_len[u - cs];
            recv_bytes += (s.bytes + 15) & ~15ull;
            recvs.push_back(s);
        }
    }
    trace.mark(c, "plan");
    if (world > 1)
    {
        TRY(ws_reserve(c, WS_X_SEND, send_bytes + 64));
        TRY(ws_reserve(c, WS_X_RECV, recv_bytes + 64));
        if (!sends.empty())
        {
            size_t n = 0;
            for (const Seg& s : sends) n += s.b - s.a;
            std::vector<uint64_t> src(n), dst(n);
            std::vector<uint32_t> len(n);
            size_t i = 0;
            for (const Seg& s : sends)
            {
                uint64_t w = s.at;
                for (uint32_t u = s.a; u < s.b; ++u, ++i)
                {
                    src[i] = off_of(u);
                    dst[i] = w;
                    len[i] = own_len[u - us];
                    w += own_len[u - us];
                }
            }
            TRY(ws_reserve(c, WS_X_TAB, 20 * n + 64));
            uint64_t* d_src = ws<uint64_t>(c, WS_X_TAB);
            uint64_t* d_dst = d_src + n;
            uint32_t* d_len = reinterpret_cast<uint32_t*>(d_dst + n);
            CU(cudaMemcpyAsync(d_src, src.data(), 8 * n, cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(d_dst, dst.data(), 8 * n, cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(d_len, len.data(), 4 * n, cudaMemcpyHostToDevice, c->stream));
            launch_gather_chunks(m->d_arena, d_src, d_dst, d_len, ws<uint8_t>(c, WS_X_SEND), (uint32_t)n, c->stream);
            CU(cudaStreamSynchronize(c->stream)); // the host vectors go out of scope
        }
        NC(g_nccl.GroupStart());
        for (const Seg& s : sends) NC(g_nccl.Send(ws<uint8_t>(c, WS_X_SEND) + s.at, s.bytes, ncclUint8, (int)s.peer, m->comm, c->stream));
        for (const Seg& s : recvs) NC(g_nccl.Recv(ws<uint8_t>(c, WS_X_RECV) + s.at, s.bytes, ncclUint8, (int)s.peer, m->comm, c->stream));
        NC(g_nccl.GroupEnd());
        c->launches += 2;
    }

    trace.mark(c, "exchange");
    // ---- this rank's blocks: chunks [cs, ce) with their bytes at absolute device addresses (own arena or the receive buffer)
    if (out_my_blocks) *out_my_blocks = my_b1 - my_b0;
    if (cs >= ce) return 0;
    const uint32_t n = ce - cs;
    std::vector<uint64_t> hashes(n), addr(n);
    CU(cudaMemcpyAsync(hashes.data(), ws<uint64_t>(c, WS_UHASH) + cs, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    for (uint32_t u = std::max(cs, us); u < std::min(ce, ue); ++u) addr[u - cs] = (uint64_t)(uintptr_t)m->d_arena + off_of(u);
    for (const Seg& s : recvs)
    {
        uint64_t w = (uint64_t)(uintptr_t)ws<uint8_t>(c, WS_X_RECV) + s.at;
        for (uint32_t u = s.a; u < s.b; ++u)
        {
            addr[u - cs] = w;
            w += my_len[u - cs];
        }
    }
    CU(cudaStreamSynchronize(c->stream));
    std::vector<uint32_t> counts(my_b1 - my_b0);
    for (uint32_t b = my_b0; b < my_b1; ++b) counts[b - my_b0] = blk_first[b + 1] - blk_first[b];
    uint32_t most = 1;
    for (uint32_t v : counts) most = std::max(most, v);
    const int rc = write_blocks_impl(c, nullptr, ~0ull, n, hashes.data(), my_len.data(), my_tag.data(), addr.data(), m->hash_type, 0xffffffffu, most,
                                     my_b1 - my_b0, counts.data(), flags, sink, user);
    trace.mark(c, "write");
    return rc;
}

// ================================================================ CompressionAPI batch entry points (host buffers)

extern "C" uint64_t lt_b200_lz4_bound(uint64_t size) { return size + size / 255 + 16; }

namespace {

enum CodecMode { CODEC_LZ4_ENCODE, CODEC_LZ4_DECODE, CODEC_ZSTD_DECODE };

// shared driver of LZ4 encode and the two decoders: stage inputs at 16-byte aligned offsets, run the kernel, copy results out
int codec_host_batch(lt_b200_context* c, uint32_t count, const void* const* src, const uint32_t* src_size, void* const* dst,
                     const uint64_t* dst_capacity, uint64_t* out_size, CodecMode mode)
{
    const bool compress = mode == CODEC_LZ4_ENCODE;
    if (!c || (count && (!src || !src_size || !dst || !dst_capacity || !out_size))) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    uint64_t budget = (uint64_t)(free_b * 0.4);
    if (budget > (16ull << 30)) budget = 16ull << 30;
    std::vector<uint64_t> in_off, out_off;
    std::vector<uint32_t> in_len, out_cap, job_start;
    for (uint32_t i0 = 0; i0 < count;)
    {
        uint64_t in_bytes = 0, out_bytes = 0;
        uint32_t job_cap = 0;
        uint32_t i1 = i0;
        in_off.clear(); out_off.clear(); in_len.clear(); out_cap.clear(); job_start.clear();
        while (i1 < count)
        {
            const uint64_t need_out = compress ? 8 + lt_b200_lz4_bound(src_size[i1]) : dst_capacity[i1];
            if (compress && dst_capacity[i1] < lt_b200_lz4_bound(src_size[i1])) return fail(c, ENOMEM, "buffer %u: capacity below LZ4_COMPRESSBOUND", i1);
            if (need_out >= 0xffffffffull) return fail(c, E2BIG, "buffer %u too large", i1);
            const uint64_t a = ((uint64_t)src_size[i1] + 16 + 15) & ~15ull, b = (need_out + 16 + 15) & ~15ull;
            if (i1 > i0 && in_bytes + out_bytes + a + b > budget) break;
            in_off.push_back(in_bytes); out_off.push_back(out_bytes); in_len.push_back(src_size[i1]); out_cap.push_back((uint32_t)need_out);
            job_start.push_back(job_cap);
            job_cap += lz4_copy_job_capacity(src_size[i1]);
            in_bytes += a; out_bytes += b;
            ++i1;
        }
        const uint32_t nb = i1 - i0;
        TRY(ws_reserve(c, WS_BLK_RAW, in_bytes + 64));
        TRY(ws_reserve(c, WS_BLK_OUT, out_bytes + 64));
        TRY(ws_reserve(c, WS_BLK_RAW_OFF, sizeof(uint64_t) * (size_t)nb));
        TRY(ws_reserve(c, WS_BLK_RAW_LEN, sizeof(uint32_t) * (size_t)nb));
        TRY(ws_reserve(c, WS_BLK_OUT_OFF, sizeof(uint64_t) * (size_t)nb));
        TRY(ws_reserve(c, WS_BLK_OUT_LEN, sizeof(uint32_t) * (size_t)nb));
        TRY(ws_reserve(c, WS_BLK_LEN, sizeof(uint32_t) * (size_t)nb));
        TRY(ws_reserve(c, WS_BLK_JOBS, sizeof(uint3) * (size_t)(job_cap + 1)));
        TRY(ws_reserve(c, WS_BLK_JOB_START, sizeof(uint32_t) * (size_t)nb));
        TRY(ws_reserve(c, WS_BLK_JOB_COUNT, sizeof(uint32_t) * (size_t)nb));
        for (uint32_t i = 0; i < nb; ++i)
            if (in_len[i]) CU(cudaMemcpyAsync(ws<uint8_t>(c, WS_BLK_RAW) + in_off[i], src[i0 + i], in_len[i], cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(ws<void>(c, WS_BLK_RAW_OFF), in_off.data(), sizeof(uint64_t) * (size_t)nb, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(ws<void>(c, WS_BLK_RAW_LEN), in_len.data(), sizeof(uint32_t) * (size_t)nb, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(ws<void>(c, WS_BLK_OUT_OFF), out_off.data(), sizeof(uint64_t) * (size_t)nb, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(ws<void>(c, WS_BLK_LEN), out_cap.data(), sizeof(uint32_t) * (size_t)nb, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(ws<void>(c, WS_BLK_JOB_START), job_start.data(), sizeof(uint32_t) * (size_t)nb, cudaMemcpyHostToDevice, c->stream));
        if (compress)
        {
            TRY(ws_reserve(c, WS_LZ4_TABLES, LZ4_TABLE_BYTES_PER_BLOCK * (size_t)nb));
            TRY(ws_reserve(c, WS_LZ4_V2, lz4_v2_scratch_bytes()));
            ProfScope ps(c, LT_B200_KERNEL_LZ4, in_bytes);
            CU(launch_lz4_blocks(ws<uint8_t>(c, WS_BLK_RAW), ws<uint64_t>(c, WS_BLK_RAW_OFF), ws<uint32_t>(c, WS_BLK_RAW_LEN), ws<uint8_t>(c, WS_BLK_OUT),
                                 ws<uint64_t>(c, WS_BLK_OUT_OFF), ws<uint32_t>(c, WS_BLK_OUT_LEN), ws<uint3>(c, WS_BLK_JOBS),
                                 ws<uint32_t>(c, WS_BLK_JOB_START), ws<uint32_t>(c, WS_BLK_JOB_COUNT), nb, ws<uint32_t>(c, WS_LZ4_TABLES),
                                 ws<void>(c, WS_LZ4_V2), c->sm_count, c->stream));
            c->launches += 2;
        }
        else if (mode == CODEC_LZ4_DECODE)
        {
            ProfScope ps(c, LT_B200_KERNEL_LZ4_DECODE, out_bytes);
            CU(launch_lz4_decode(ws<uint8_t>(c, WS_BLK_RAW), ws<uint64_t>(c, WS_BLK_RAW_OFF), ws<uint32_t>(c, WS_BLK_RAW_LEN), ws<uint8_t>(c, WS_BLK_OUT),
                                 ws<uint64_t>(c, WS_BLK_OUT_OFF), ws<uint32_t>(c, WS_BLK_LEN), ws<uint32_t>(c, WS_BLK_OUT_LEN), nb, c->stream));
            c->launches += 1;
        }
        else
        {
            const uint32_t workers = zstd_dec_worker_count(nb, c->sm_count);
            TRY(ws_reserve(c, WS_ZSTD_DEC_WORKERS, zstd_dec_worker_bytes() * (size_t)workers));
            TRY(ws_reserve(c, WS_QUEUE_HEAD, 64));
            ProfScope ps(c, LT_B200_KERNEL_ZSTD_DECODE, out_bytes);
            CU(launch_zstd_decode(ws<uint8_t>(c, WS_BLK_RAW), ws<uint64_t>(c, WS_BLK_RAW_OFF), ws<uint32_t>(c, WS_BLK_RAW_LEN), ws<uint8_t>(c, WS_BLK_OUT),
                                  ws<uint64_t>(c, WS_BLK_OUT_OFF), ws<uint32_t>(c, WS_BLK_LEN), ws<uint32_t>(c, WS_BLK_OUT_LEN), nb,
                                  ws<void>(c, WS_ZSTD_DEC_WORKERS), workers, ws<uint32_t>(c, WS_QUEUE_HEAD), c->stream));
            c->launches += 1;
        }
        TRY(hs_reserve(c, HS_BLK_OUT_LEN, sizeof(uint32_t) * (size_t)nb + 16));
        uint32_t* h_len = hs<uint32_t>(c, HS_BLK_OUT_LEN);
        CU(cudaMemcpyAsync(h_len, ws<void>(c, WS_BLK_OUT_LEN), sizeof(uint32_t) * (size_t)nb, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        for (uint32_t i = 0; i < nb; ++i)
        {
            if (h_len[i] == 0xffffffffu) return fail(c, EBADF, "buffer %u: malformed %s", i0 + i, mode == CODEC_ZSTD_DECODE ? "ZStd frame" : "LZ4 stream");
            const uint64_t n = compress ? h_len[i] - 8u : h_len[i];                       // the kernel writes the block-store header first
            const uint8_t* d = ws<uint8_t>(c, WS_BLK_OUT) + out_off[i] + (compress ? 8 : 0);
            if (n > dst_capacity[i0 + i]) return fail(c, ENOMEM, "buffer %u: output does not fit", i0 + i);
            if (n) CU(cudaMemcpyAsync(dst[i0 + i], d, n, cudaMemcpyDeviceToHost, c->stream));
            out_size[i0 + i] = n;
        }
        CU(cudaStreamSynchronize(c->stream));
        i0 = i1;
    }
    return 0;
}

} // namespace

// diagnostic: cycles the ZStd workers spent per phase since the worker slabs were (re)allocated
extern "C" int lt_b200_zstd_phase_cycles(lt_b200_context* c, uint64_t out_cycles[4])
{
    if (!c || !out_cycles) return EINVAL;
    CU(cudaSetDevice(c->device));
    for (int i = 0; i < 4; ++i) out_cycles[i] = 0;
    const size_t slab = zstd_worker_bytes();
    const size_t n = c->ws[WS_ZSTD_WORKERS].cap / slab;
    for (size_t w = 0; w < n; ++w)
    {
        unsigned long long t[4];
        CU(cudaMemcpy(t, ws<uint8_t>(c, WS_ZSTD_WORKERS) + w * slab + zstd_worker_phase_offset(), sizeof(t), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 4; ++i) out_cycles[i] += t[i];
    }
    return 0;
}

extern "C" int lt_b200_zstd_compress_host(lt_b200_context* c, uint32_t compression_type, uint32_t count, const void* const* src, const uint32_t* src_size,
                                          void* const* dst, const uint64_t* dst_capacity, uint64_t* out_size)
{
    if (!c || (count && (!src || !src_size || !dst || !dst_capacity || !out_size))) return EINVAL;
    CU(cudaSetDevice(c->device));
    c->err[0] = 0;
    if (!is_zstd_level3(compression_type)) return fail(c, ENOTSUP, "compression type 0x%08x has no device encoder (only 'ztd1' / 'ztd2' = level 3)", compression_type);
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    uint64_t budget = (uint64_t)(free_b * 0.3);
    if (budget > (16ull << 30)) budget = 16ull << 30;
    std::vector<uint64_t> in_off, out_off;
    std::vector<uint32_t> in_len;
    for (uint32_t i0 = 0; i0 < count;)
    {
        uint64_t in_bytes = 0, out_bytes = 0;
        uint32_t i1 = i0;
        in_off.clear(); out_off.clear(); in_len.clear();
        while (i1 < count)
        {
            if (src_size[i1] >= 0x7E000000u) return fail(c, E2BIG, "buffer %u too large", i1);
            if (dst_capacity[i1] < lt_b200_zstd_bound(src_size[i1])) return fail(c, EINVAL, "buffer %u: capacity below ZSTD_COMPRESSBOUND", i1);
            const uint64_t a = ((uint64_t)src_size[i1] + 16 + 15) & ~15ull, b = (8 + lt_b200_zstd_bound(src_size[i1]) + 16 + 15) & ~15ull;
            if (i1 > i0 && in_bytes + out_bytes + a + b > budget) break;
            in_off.push_back(in_bytes); out_off.push_back(out_bytes); in_len.push_back(src_size[i1]);
            in_bytes += a; out_bytes += b;
            ++i1;
        }
        const uint32_t nb = i1 - i0;
        TRY(ws_reserve(c, WS_BLK_RAW, in_bytes + 64));
        TRY(ws_reserve(c, WS_BLK_OUT, out_bytes + 64));
        uint64_t payload = 0;
        for (uint32_t i = 0; i < nb; ++i)
        {
            payload += in_len[i];
            if (in_len[i]) CU(cudaMemcpyAsync(ws<uint8_t>(c, WS_BLK_RAW) + in_off[i], src[i0 + i], in_len[i], cudaMemcpyHostToDevice, c->stream));
        }
        TRY(zstd_launch(c, ws<uint8_t>(c, WS_BLK_RAW), in_off, in_len, ws<uint8_t>(c, WS_BLK_OUT), out_off, payload));
        TRY(hs_reserve(c, HS_BLK_OUT_LEN, sizeof(uint32_t) * (size_t)nb + 16));
        uint32_t* h_len = hs<uint32_t>(c, HS_BLK_OUT_LEN);
        CU(cudaMemcpyAsync(h_len, ws<void>(c, WS_ZSTD_OUT_LEN), sizeof(uint32_t) * (size_t)nb, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaGetLastError());
        for (uint32_t i = 0; i < nb; ++i)
        {
            if (h_len[i] == 0xffffffffu) return fail(c, EINVAL, "buffer %u: ZStd encoder error", i0 + i);
            const uint64_t n = h_len[i] - 8u; // the kernel writes the block-store header first
            CU(cudaMemcpyAsync(dst[i0 + i], ws<uint8_t>(c, WS_BLK_OUT) + out_off[i] + 8, n, cudaMemcpyDeviceToHost, c->stream));
            out_size[i0 + i] = n;
        }
        CU(cudaStreamSynchronize(c->stream));
        i0 = i1;
    }
    return 0;
}

extern "C" int lt_b200_lz4_compress_host(lt_b200_context* c, uint32_t count, const void* const* src, const uint32_t* src_size, void* const* dst,
                                         const uint64_t* dst_capacity, uint64_t* out_size)
{
    return codec_host_batch(c, count, src, src_size, dst, dst_capacity, out_size, CODEC_LZ4_ENCODE);
}

extern "C" int lt_b200_lz4_decompress_host(lt_b200_context* c, uint32_t count, const void* const* src, const uint32_t* src_size, void* const* dst,
                                           const uint64_t* dst_capacity, uint64_t* out_size)
{
    return codec_host_batch(c, count, src, src_size, dst, dst_capacity, out_size, CODEC_LZ4_DECODE);
}

extern "C" int lt_b200_zstd_decompress_host(lt_b200_context* c, uint32_t count, const void* const* src, const uint32_t* src_size, void* const* dst,
                                            const uint64_t* dst_capacity, uint64_t* out_size)
{
    return codec_host_batch(c, count, src, src_size, dst, dst_capacity, out_size, CODEC_ZSTD_DECODE);
}
