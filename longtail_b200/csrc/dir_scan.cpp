// dir_scan.cpp — the step BEFORE the hot path: what Longtail_GetFilesRecursively2 (src/longtail.c:1656-1893)
// does for cmd/main.c:UpSync, and a reader that keeps lt_b200_index_stream_assets fed.
//
//  * scan: every directory is listed with readdir + stat like the reference's file storage (lib/longtail_platform.c:2025-2180: "." and ".."
//    skipped, permissions = st_mode & 0x1FF, directories have size 0); an entry is named by its path relative to the root; the result is
//    sorted with strcmp over those names BEFORE directories receive their trailing '/' (SortScannedPaths, src/longtail.c:1600-1620, then
//    :1866-1872) — so "a.b" sorts before "a/" 's children but the directory "a" itself sorts by the name "a";
//  * read: the streaming verb hands out batches of {asset, offset, size, destination in pinned staging}; a pool of threads preads them
//    (the reference issues one StorageAPI.Read per chunker refill from its job threads, src/longtail.c:1923-1960).
//
// Host code only.  Entries that are neither regular files nor directories (the reference's iterator hands them on with a NULL name and
// fails) are rejected with ENOTSUP.
#include "../../include/longtail_b200.h"

#include <dirent.h>
#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <stdio.h>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Entry
{
    std::string name; // relative to the root, no trailing '/'
    uint64_t size;
    uint16_t permissions;
    bool is_dir;
};

// one directory: ScanFolder (src/longtail.c:1440-1570)
int list_directory(const std::string& root, const std::string& sub, std::vector<Entry>* out)
{
    const std::string full = sub.empty() ? root : root + "/" + sub;
    DIR* d = opendir(full.c_str());
    if (!d) return errno == ENOENT ? 0 : errno; // a folder that vanished lists as empty (:1465-1482)
    int err = 0;
    for (;;)
    {
        errno = 0;
        struct dirent* e = readdir(d);
        if (!e)
        {
            err = errno;
            break;
        }
        const char* n = e->d_name;
        if (n[0] == '.' && (n[1] == 0 || (n[1] == '.' && n[2] == 0))) continue;
        const std::string path = full + "/" + n;
        struct stat st;
        if (stat(path.c_str(), &st) != 0)
        {
            err = errno;
            break;
        }
        const bool is_dir = S_ISDIR(st.st_mode);
        if (!is_dir && !S_ISREG(st.st_mode)) { err = ENOTSUP; break; }
        if (e->d_type != DT_UNKNOWN && e->d_type != (is_dir ? DT_DIR : DT_REG)) { err = ENOTSUP; break; } // a symbolic link: the reference has no name for it
        Entry en;
        en.name = sub.empty() ? std::string(n) : sub + "/" + n;
        en.is_dir = is_dir;
        en.size = is_dir ? 0 : (uint64_t)st.st_size;
        en.permissions = (uint16_t)(st.st_mode & 0x1FF);
        out->push_back(std::move(en));
    }
    closedir(d);
    return err;
}

} // namespace

struct lt_b200_file_list
{
    struct lt_b200_assets assets;
    std::vector<uint64_t> sizes;
    std::vector<uint32_t> offsets;
    std::vector<uint16_t> permissions;
    std::string path_data;
    std::string root;
};

extern "C" int lt_b200_scan_directory(const char* root_path, uint32_t threads, lt_b200_file_list** out_list)
{
    if (!root_path || !out_list) return EINVAL;
    std::string root = root_path;
    while (root.size() > 1 && root.back() == '/') root.pop_back();
    std::vector<Entry> all;
    std::vector<std::string> level(1, std::string());
    if (threads == 0) threads = 1;
    // level by level, the folders of one level in parallel (the reference: one ScanFolder job per folder, :1681-1747)
    while (!level.empty())
    {
        std::vector<std::vector<Entry>> found(level.size());
        std::atomic<size_t> next(0);
        std::atomic<int> first_err(0);
        auto work = [&]() {
            for (;;)
            {
                const size_t i = next.fetch_add(1);
                if (i >= level.size()) return;
                const int err = list_directory(root, level[i], &found[i]);
                if (err)
                {
                    int expected = 0;
                    first_err.compare_exchange_strong(expected, err);
                }
            }
        };
        const uint32_t n = (uint32_t)std::min<size_t>(threads, level.size());
        std::vector<std::thread> pool;
        for (uint32_t t = 1; t < n; ++t) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
        if (first_err.load()) return first_err.load();
        std::vector<std::string> deeper;
        for (auto& f : found)
            for (auto& e : f)
            {
                if (e.is_dir) deeper.push_back(e.name);
                all.push_back(std::move(e));
            }
        level.swap(deeper);
    }
    std::sort(all.begin(), all.end(), [](const Entry& a, const Entry& b) { return strcmp(a.name.c_str(), b.name.c_str()) < 0; });
    if (all.size() > 0xfffffffeull) return E2BIG;
    lt_b200_file_list* l = new (std::nothrow) lt_b200_file_list();
    if (!l) return ENOMEM;
    l->root = root;
    for (const Entry& e : all)
    {
        l->offsets.push_back((uint32_t)l->path_data.size());
        l->path_data += e.name;
        if (e.is_dir) l->path_data += '/';
        l->path_data += '\0';
        l->sizes.push_back(e.size);
        l->permissions.push_back(e.permissions);
    }
    if (l->path_data.size() > 0xffffffffull) { delete l; return E2BIG; }
    l->assets.asset_count = (uint32_t)all.size();
    l->assets.path_data_size = (uint32_t)l->path_data.size();
    l->assets.sizes = l->sizes.data();
    l->assets.path_start_offsets = l->offsets.data();
    l->assets.permissions = l->permissions.data();
    l->assets.path_data = l->path_data.data();
    *out_list = l;
    return 0;
}

extern "C" const struct lt_b200_assets* lt_b200_file_list_assets(const lt_b200_file_list* list) { return list ? &list->assets : nullptr; }

extern "C" void lt_b200_file_list_free(lt_b200_file_list* list) { delete list; }

namespace {

struct ReadContext
{
    const lt_b200_file_list* list;
    uint32_t threads;
};

// lt_b200_read_batch_func: jobs are handed to `threads` readers; a reader keeps the file of consecutive jobs of one asset open
int read_batch(void* user, const struct lt_b200_read_job* jobs, uint32_t job_count)
{
    const ReadContext* rc = static_cast<const ReadContext*>(user);
    std::atomic<uint32_t> next(0);
    std::atomic<int> first_err(0);
    auto work = [&]() {
        int fd = -1;
        uint32_t open_asset = 0xffffffffu;
        for (;;)
        {
            // a few jobs at a time, so that one reader works through neighbouring parts of the same file
            const uint32_t i0 = next.fetch_add(4), i1 = std::min(job_count, i0 + 4);
            if (i0 >= job_count) break;
            for (uint32_t i = i0; i < i1 && !first_err.load(); ++i)
            {
                const lt_b200_read_job& j = jobs[i];
                if (j.asset_index != open_asset)
                {
                    if (fd >= 0) close(fd);
                    const std::string path = rc->list->root + "/" + (rc->list->path_data.data() + rc->list->offsets[j.asset_index]);
                    fd = open(path.c_str(), O_RDONLY);
                    open_asset = j.asset_index;
                    if (fd < 0)
                    {
                        int expected = 0;
                        first_err.compare_exchange_strong(expected, errno ? errno : EIO);
                        open_asset = 0xffffffffu;
                        break;
                    }
                }
                uint8_t* dst = static_cast<uint8_t*>(j.dst);
                uint64_t off = j.offset;
                uint32_t left = j.size;
                while (left)
                {
                    const ssize_t n = pread(fd, dst, left, (off_t)off);
                    if (n < 0 && errno == EINTR) continue;
                    if (n <= 0)
                    {
                        int expected = 0;
                        first_err.compare_exchange_strong(expected, n < 0 ? errno : EIO); // the file shrank since the scan
                        break;
                    }
                    dst += n;
                    off += (uint64_t)n;
                    left -= (uint32_t)n;
                }
            }
        }
        if (fd >= 0) close(fd);
    };
    const uint32_t n = std::max(1u, std::min(rc->threads, (job_count + 3) / 4));
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < n; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    return first_err.load();
}

} // namespace

extern "C" int lt_b200_index_file_list(lt_b200_context* context, const lt_b200_file_list* list, const uint32_t* asset_tags, uint32_t hash_type,
                                       uint32_t target_chunk_size, uint32_t reader_threads, const void** out_buffer, uint64_t* out_size)
{
    if (!context || !list) return EINVAL;
    ReadContext rc = {list, reader_threads ? reader_threads : 1};
    return lt_b200_index_stream_assets(context, &list->assets, asset_tags, hash_type, target_chunk_size, read_batch, &rc, out_buffer, out_size);
}

// ---------------------------------------------------------------- a whole upsync of a scanned tree (cmd/main.c:UpSync, :972-1153)
// The tree is loaded ONCE into a device arena (180 GB of HBM hold most asset trees whole): readers pread into two pinned staging
// buffers while the previous one crosses PCIe.  Then everything stays on the device: CreateVersionIndex, DiffHashes against the chunks the
// store already lists, block packing, payload gather, compression, and the stored blocks leave through the fsblockstore-layout sink.
extern "C" int lt_b200_upsync_file_list(lt_b200_context* context, const lt_b200_file_list* list, const uint32_t* asset_tags, uint32_t hash_type,
                                        uint32_t target_chunk_size, uint32_t max_block_size, uint32_t max_chunks_per_block, uint32_t reader_threads,
                                        lt_b200_fs_store* store, const void** out_version_index, uint64_t* out_size, uint32_t* out_blocks_written)
{
    if (!context || !list || !store || !out_version_index || !out_size) return EINVAL;
    const bool trace = getenv("LT_B200_TRACE") != nullptr; // phase times on stderr
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    const uint32_t n = list->assets.asset_count;
    std::vector<uint64_t> arena_off(n ? n : 1);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        arena_off[i] = total;
        total += (list->sizes[i] + 255u) & ~(uint64_t)255u;
    }
    const uint64_t arena_bytes = total + 4096;
    void* arena = nullptr;
    int err = lt_b200_device_alloc(context, arena_bytes, &arena);
    if (err) return err; // ENOMEM: the tree does not fit this GPU; shard the file list across ranks (longtail_b200/distributed.py)
    const uint64_t stage_bytes = 256ull << 20;
    void* stage[2] = {nullptr, nullptr};
    err = lt_b200_host_alloc_pinned(context, stage_bytes, &stage[0]);
    if (!err) err = lt_b200_host_alloc_pinned(context, stage_bytes, &stage[1]);
    ReadContext rc = {list, reader_threads ? reader_threads : 1};
    // pieces of at most 8 MiB, packed into staging batches; the arena keeps each asset contiguous at arena_off
    struct Piece { uint32_t asset; uint64_t offset; uint32_t size; };
    std::vector<Piece> pieces;
    for (uint32_t i = 0; i < n; ++i)
        for (uint64_t o = 0; o < list->sizes[i]; o += 8u << 20)
            pieces.push_back({i, o, (uint32_t)std::min<uint64_t>(8u << 20, list->sizes[i] - o)});
    struct Batch { size_t first, last; uint64_t bytes; };
    auto plan = [&](size_t first) {
        Batch b = {first, first, 0};
        while (b.last < pieces.size() && (b.bytes + pieces[b.last].size <= stage_bytes || b.last == first)) b.bytes += pieces[b.last++].size;
        return b;
    };
    auto read_into = [&](const Batch& b, void* buf) -> int {
        std::vector<lt_b200_read_job> jobs;
        uint64_t at = 0;
        for (size_t k = b.first; k < b.last; ++k)
        {
            jobs.push_back({pieces[k].asset, pieces[k].size, pieces[k].offset, static_cast<uint8_t*>(buf) + at});
            at += pieces[k].size;
        }
        return jobs.empty() ? 0 : read_batch(&rc, jobs.data(), (uint32_t)jobs.size());
    };
    if (!err && !pieces.empty())
    {
        Batch cur = plan(0);
        int which = 0;
        err = read_into(cur, stage[0]);
        while (!err)
        {
            const bool has_next = cur.last < pieces.size();
            Batch next = has_next ? plan(cur.last) : cur;
            int read_err = 0;
            std::thread reader;
            if (has_next) reader = std::thread([&]() { read_err = read_into(next, stage[which ^ 1]); });
            uint64_t at = 0;
            for (size_t k = cur.first; k < cur.last && !err; ++k) // consecutive pieces of one asset are consecutive in both buffers: one copy per run
            {
                size_t e = k;
                uint64_t run = 0;
                while (e < cur.last && pieces[e].asset == pieces[k].asset) run += pieces[e++].size;
                err = lt_b200_copy_to_device(context, static_cast<uint8_t*>(arena) + arena_off[pieces[k].asset] + pieces[k].offset,
                                             static_cast<uint8_t*>(stage[which]) + at, run);
                at += run;
                k = e - 1;
            }
            if (has_next) reader.join();
            if (!err) err = read_err;
            if (!has_next) break;
            cur = next;
            which ^= 1;
        }
    }
    const double t_loaded = now();
    const void* vi = nullptr;
    uint64_t vi_size = 0;
    if (!err) err = lt_b200_index_device_assets(context, static_cast<const uint8_t*>(arena), arena_bytes, &list->assets, arena_off.data(), asset_tags, hash_type,
                                                target_chunk_size, &vi, &vi_size);
    const double t_indexed = now();
    if (!err)
    {
        // the unique chunks of the version, in version order, from the serialised index (src/longtail.c:2566-2584): 6 x u32, then
        // u64[A] x 3, u32[A] x 2, u32[I], u64 chunk_hash[C], u32 chunk_size[C], u32 chunk_tag[C]
        const uint8_t* p = static_cast<const uint8_t*>(vi);
        uint32_t head[6];
        memcpy(head, p, sizeof(head));
        const uint64_t A = head[3], C = head[4], I = head[5];
        const uint8_t* q = p + 24 + 24 * A + 8 * A + 4 * I;
        std::vector<uint64_t> hashes(C ? C : 1), offsets(C ? C : 1);
        std::vector<uint32_t> sizes(C ? C : 1), tags(C ? C : 1);
        memcpy(hashes.data(), q, 8 * C);
        memcpy(sizes.data(), q + 8 * C, 4 * C);
        memcpy(tags.data(), q + 12 * C, 4 * C);
        if (C) err = lt_b200_unique_chunk_offsets(context, offsets.data(), (uint32_t)C);
        // Longtail_CreateMissingContent: only what the store does not list yet, in version order
        uint32_t have = 0;
        std::vector<uint64_t> existing;
        if (!err) err = lt_b200_fs_store_existing_chunks(store, nullptr, 0, &have);
        if (!err && have)
        {
            existing.resize(have);
            err = lt_b200_fs_store_existing_chunks(store, existing.data(), have, &have);
        }
        std::vector<uint8_t> missing(C ? C : 1, 1);
        if (!err && C) err = lt_b200_missing_chunks(context, (uint32_t)C, hashes.data(), have, have ? existing.data() : nullptr, missing.data());
        uint32_t m = 0;
        for (uint64_t i = 0; i < C && !err; ++i)
            if (missing[i])
            {
                hashes[m] = hashes[i]; sizes[m] = sizes[i]; tags[m] = tags[i]; offsets[m] = offsets[i];
                ++m;
            }
        uint64_t before = 0, after = 0;
        lt_b200_fs_store_stats(store, &before, nullptr, nullptr);
        if (!err && m)
            err = lt_b200_write_blocks_device(context, static_cast<const uint8_t*>(arena), arena_bytes, m, hashes.data(), sizes.data(), tags.data(), offsets.data(),
                                              hash_type, max_block_size, max_chunks_per_block, lt_b200_fs_store_sink, store);
        if (!err) err = lt_b200_fs_store_flush(store);
        lt_b200_fs_store_stats(store, &after, nullptr, nullptr);
        if (out_blocks_written) *out_blocks_written = (uint32_t)(after - before);
        // write_blocks_device reuses pinned staging of the context, the index buffer is a different one and stays valid (see the header)
        *out_version_index = vi;
        *out_size = vi_size;
    }
    if (trace)
        fprintf(stderr, "lt_b200_upsync_file_list: %.2f GiB: read + H2D %.3f s, index %.3f s, missing + blocks + sink + flush %.3f s\n", total / 1073741824.0,
                t_loaded - t_start, t_indexed - t_loaded, now() - t_indexed);
    if (stage[0]) lt_b200_host_free_pinned(context, stage[0]);
    if (stage[1]) lt_b200_host_free_pinned(context, stage[1]);
    lt_b200_device_free(context, arena);
    return err;
}

// cmd/main.c:UpSync (:972-1153) for assets that sit in HOST memory (pinned for full PCIe speed) and fit one GPU: every asset is copied once
// into a device arena, then CreateVersionIndex, CreateMissingContent against `existing_hashes` (may be NULL: a fresh store) and WriteContent
// run on the resident bytes; stored blocks leave through `sink`.  The reference reads every file twice (once to index it, once to compose
// the blocks, src/longtail.c:2076, :4675); here the bytes cross PCIe once.
extern "C" int lt_b200_upsync_host_assets(lt_b200_context* context, const lt_b200_assets* assets, const uint8_t* const* asset_data,
                                          const uint32_t* asset_tags, uint32_t hash_type, uint32_t target_chunk_size, uint32_t max_block_size,
                                          uint32_t max_chunks_per_block, uint32_t existing_count, const uint64_t* existing_hashes, uint32_t flags,
                                          lt_b200_block_sink sink, void* user, const void** out_version_index, uint64_t* out_size,
                                          uint32_t* out_chunks_written)
{
    if (!context || !assets || !sink || !out_version_index || !out_size || (assets->asset_count && !asset_data)) return EINVAL;
    const uint32_t n = assets->asset_count;
    std::vector<uint64_t> arena_off(n ? n : 1);
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; ++i)
    {
        arena_off[i] = total;
        total += (assets->sizes[i] + 255u) & ~(uint64_t)255u;
    }
    const uint64_t arena_bytes = total + 4096;
    void* arena = nullptr;
    int err = lt_b200_device_alloc(context, arena_bytes, &arena);
    if (err == ENOMEM && lt_b200_trim(context) == 0) err = lt_b200_device_alloc(context, arena_bytes, &arena); // workspace of earlier verbs in the way
    if (err) return err; // ENOMEM: shard the asset list across GPUs (lt_b200_index_sharded)
    for (uint32_t i = 0; i < n && !err; ++i)
        if (assets->sizes[i])
        {
            if (!asset_data[i]) err = EINVAL;
            else err = lt_b200_copy_to_device_async(context, static_cast<uint8_t*>(arena) + arena_off[i], asset_data[i], assets->sizes[i]);
        }
    const void* vi = nullptr;
    uint64_t vi_size = 0;
    if (!err) err = lt_b200_index_device_assets(context, static_cast<const uint8_t*>(arena), arena_bytes, assets, arena_off.data(), asset_tags, hash_type,
                                                target_chunk_size, &vi, &vi_size);
    if (!err)
    {
        const uint8_t* p = static_cast<const uint8_t*>(vi);
        uint32_t head[6];
        memcpy(head, p, sizeof(head));
        const uint64_t A = head[3], C = head[4], I = head[5];
        const uint8_t* q = p + 24 + 24 * A + 8 * A + 4 * I; // the unique chunks of the version, in version order (src/longtail.c:2566-2584)
        std::vector<uint64_t> hashes(C ? C : 1), offsets(C ? C : 1);
        std::vector<uint32_t> sizes(C ? C : 1), tags(C ? C : 1);
        memcpy(hashes.data(), q, 8 * C);
        memcpy(sizes.data(), q + 8 * C, 4 * C);
        memcpy(tags.data(), q + 12 * C, 4 * C);
        if (C) err = lt_b200_unique_chunk_offsets(context, offsets.data(), (uint32_t)C);
        uint32_t m = (uint32_t)C;
        if (!err && C && existing_count)
        {
            std::vector<uint8_t> missing(C, 1);
            err = lt_b200_missing_chunks(context, (uint32_t)C, hashes.data(), existing_count, existing_hashes, missing.data());
            m = 0;
            for (uint64_t i = 0; i < C && !err; ++i)
                if (missing[i])
                {
                    hashes[m] = hashes[i]; sizes[m] = sizes[i]; tags[m] = tags[i]; offsets[m] = offsets[i];
                    ++m;
                }
        }
        if (!err && m)
            err = lt_b200_write_blocks_device_ex(context, static_cast<const uint8_t*>(arena), arena_bytes, m, hashes.data(), sizes.data(), tags.data(),
                                                 offsets.data(), hash_type, max_block_size, max_chunks_per_block, flags, sink, user);
        if (out_chunks_written) *out_chunks_written = m;
        *out_version_index = vi;
        *out_size = vi_size;
    }
    lt_b200_device_free(context, arena);
    return err;
}
