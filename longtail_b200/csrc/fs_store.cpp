// fs_store.cpp — the step after PutStoredBlock: a batched on-disk sink for the stored blocks that
// lt_b200_write_blocks_device produces, in the layout of the reference's fsblockstore so that an unmodified longtail opens the
// directory as a block store.
//
//   <store>/chunks/<first 4 hex digits>/0x<16 hex digits>.lrb    one file per block = the serialised stored block
//                                                                (lib/fsblockstore/longtail_fsblockstore.c:66-129, src/longtail.c:4111-4150)
//   <store>/store.lsi                                            the serialised StoreIndex (src/longtail.c:8913-8977, :9064-9121)
//   <store>/store.lsi.sync                                       flock()ed while store.lsi is replaced (longtail_fsblockstore.c:1443, lib/longtail_platform.c:2394)
//
// Semantics mirrored from the reference:
//  * a block whose file exists is not written again (SafeWriteStoredBlock, longtail_fsblockstore.c:243-320); a block is written to a
//    temporary name and renamed (:273-292);
//  * a block hash is accepted once per store object (m_BlockState, :791-803);
//  * Flush: index of the added blocks merged IN FRONT of the index found on disk — "added first as it has precedence"
//    (UpdateStoreIndex :330-352, FSBlockStore_UpdateStoreIndex + WriteStoreIndex :150-241, Longtail_MergeStoreIndex src/longtail.c:9155-9290),
//    written to a temporary name, the old file removed, the new one renamed.
//
// Host code only (no CUDA): the hot path hands over finished byte images; what is left is file IO, spread over writer threads
// because one synchronous write per block (what the reference does inside each WriteContent job) would serialise the block sink.
#include "../../include/longtail_b200.h"

#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/file.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

struct BlockRec
{
    uint64_t hash;
    uint32_t hash_id, tag;
    std::vector<uint64_t> chunk_hashes;
    std::vector<uint32_t> chunk_sizes;
};

struct Job
{
    uint64_t hash;
    void* data;
    size_t size;
};

// parsed store index (also the shape that is serialised)
struct Index
{
    uint32_t version = 0, hash_id = 0;
    std::vector<uint64_t> block_hashes, chunk_hashes;
    std::vector<uint32_t> block_offsets, block_counts, block_tags, chunk_sizes;
};

void hex16(uint64_t v, char* out) // lower case, 16 digits (HashLUT, longtail_fsblockstore.c:40)
{
    static const char* lut = "0123456789abcdef";
    for (int i = 0; i < 16; ++i) out[i] = lut[(v >> (60 - 4 * i)) & 15u];
}

int mkdir_p(const std::string& path)
{
    if (mkdir(path.c_str(), 0777) == 0 || errno == EEXIST) return 0;
    if (errno != ENOENT) return errno;
    const size_t slash = path.find_last_of('/');
    if (slash == std::string::npos || slash == 0) return ENOENT;
    int err = mkdir_p(path.substr(0, slash));
    if (err) return err;
    return (mkdir(path.c_str(), 0777) == 0 || errno == EEXIST) ? 0 : errno;
}

bool is_file(const std::string& path)
{
    struct stat st;
    return stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

int write_all(const std::string& path, const void* data, size_t size)
{
    const int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) return errno;
    const uint8_t* p = static_cast<const uint8_t*>(data);
    size_t left = size;
    while (left)
    {
        const ssize_t n = write(fd, p, left);
        if (n < 0)
        {
            if (errno == EINTR) continue;
            const int e = errno;
            close(fd);
            return e;
        }
        p += n;
        left -= (size_t)n;
    }
    return close(fd) == 0 ? 0 : errno;
}

int read_all(const std::string& path, std::vector<uint8_t>* out)
{
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return errno;
    struct stat st;
    if (fstat(fd, &st) != 0) { const int e = errno; close(fd); return e; }
    out->resize((size_t)st.st_size);
    size_t got = 0;
    while (got < out->size())
    {
        const ssize_t n = read(fd, out->data() + got, out->size() - got);
        if (n < 0) { if (errno == EINTR) continue; const int e = errno; close(fd); return e; }
        if (n == 0) break;
        got += (size_t)n;
    }
    close(fd);
    return got == out->size() ? 0 : EIO;
}

// Longtail_StoreIndex wire format (src/longtail.c:8913-8977): u32 version, hash id, block count, chunk count, then
// u64 block_hash[B], u64 chunk_hash[C], u32 block_chunks_offset[B], u32 block_chunk_count[B], u32 block_tag[B], u32 chunk_size[C]
int parse_index(const std::vector<uint8_t>& buf, Index* ix)
{
    if (buf.size() < 16) return EBADF;
    const uint32_t* h = reinterpret_cast<const uint32_t*>(buf.data());
    ix->version = h[0];
    ix->hash_id = h[1];
    const uint64_t B = h[2], C = h[3];
    if (ix->version != LT_B200_STORE_INDEX_VERSION) return EBADF; // InitStoreIndexFromData, src/longtail.c:9004
    if (16 + 8 * B + 8 * C + 12 * B + 4 * C > buf.size()) return EBADF;
    const uint8_t* p = buf.data() + 16;
    ix->block_hashes.assign(reinterpret_cast<const uint64_t*>(p), reinterpret_cast<const uint64_t*>(p) + B); p += 8 * B;
    ix->chunk_hashes.assign(reinterpret_cast<const uint64_t*>(p), reinterpret_cast<const uint64_t*>(p) + C); p += 8 * C;
    ix->block_offsets.assign(reinterpret_cast<const uint32_t*>(p), reinterpret_cast<const uint32_t*>(p) + B); p += 4 * B;
    ix->block_counts.assign(reinterpret_cast<const uint32_t*>(p), reinterpret_cast<const uint32_t*>(p) + B); p += 4 * B;
    ix->block_tags.assign(reinterpret_cast<const uint32_t*>(p), reinterpret_cast<const uint32_t*>(p) + B); p += 4 * B;
    ix->chunk_sizes.assign(reinterpret_cast<const uint32_t*>(p), reinterpret_cast<const uint32_t*>(p) + C);
    for (uint64_t b = 0; b < B; ++b)
        if ((uint64_t)ix->block_offsets[b] + ix->block_counts[b] > C) return EBADF;
    return 0;
}

void serialise_index(const Index& ix, std::vector<uint8_t>* out)
{
    const size_t B = ix.block_hashes.size(), C = ix.chunk_hashes.size();
    out->resize(16 + 8 * B + 8 * C + 12 * B + 4 * C);
    uint32_t* h = reinterpret_cast<uint32_t*>(out->data());
    h[0] = LT_B200_STORE_INDEX_VERSION;
    h[1] = ix.hash_id;
    h[2] = (uint32_t)B;
    h[3] = (uint32_t)C;
    uint8_t* p = out->data() + 16;
    memcpy(p, ix.block_hashes.data(), 8 * B); p += 8 * B;
    memcpy(p, ix.chunk_hashes.data(), 8 * C); p += 8 * C;
    memcpy(p, ix.block_offsets.data(), 4 * B); p += 4 * B;
    memcpy(p, ix.block_counts.data(), 4 * B); p += 4 * B;
    memcpy(p, ix.block_tags.data(), 4 * B); p += 4 * B;
    memcpy(p, ix.chunk_sizes.data(), 4 * C);
}

void append_block(Index* dst, uint64_t hash, uint32_t tag, const uint64_t* chunk_hashes, const uint32_t* chunk_sizes, uint32_t n)
{
    dst->block_hashes.push_back(hash);
    dst->block_tags.push_back(tag);
    dst->block_counts.push_back(n);
    dst->block_offsets.push_back((uint32_t)dst->chunk_hashes.size());
    dst->chunk_hashes.insert(dst->chunk_hashes.end(), chunk_hashes, chunk_hashes + n);
    dst->chunk_sizes.insert(dst->chunk_sizes.end(), chunk_sizes, chunk_sizes + n);
}

// Longtail_MergeStoreIndex(local, remote) (src/longtail.c:9155-9290): local blocks in their order (first occurrence of a hash), then
// the remote blocks local does not hold; EINVAL when both are non-empty and disagree about the hash identifier
int merge_index(const Index& local, const Index& remote, Index* out)
{
    const size_t lb = local.block_hashes.size(), rb = remote.block_hashes.size();
    out->version = LT_B200_STORE_INDEX_VERSION;
    if (lb == 0) out->hash_id = rb ? remote.hash_id : 0;
    else
    {
        out->hash_id = local.hash_id;
        if (rb && remote.hash_id != local.hash_id) return EINVAL;
    }
    std::unordered_set<uint64_t> seen;
    for (size_t b = 0; b < lb; ++b)
        if (seen.insert(local.block_hashes[b]).second)
            append_block(out, local.block_hashes[b], local.block_tags[b], local.chunk_hashes.data() + local.block_offsets[b],
                         local.chunk_sizes.data() + local.block_offsets[b], local.block_counts[b]);
    for (size_t b = 0; b < rb; ++b)
        if (seen.insert(remote.block_hashes[b]).second)
            append_block(out, remote.block_hashes[b], remote.block_tags[b], remote.chunk_hashes.data() + remote.block_offsets[b],
                         remote.chunk_sizes.data() + remote.block_offsets[b], remote.block_counts[b]);
    return 0;
}

} // namespace

struct lt_b200_fs_store
{
    std::string root;
    char tmp_ext[18]; // "." + 16 hex digits (GetUniqueExtension, longtail_fsblockstore.c:42-64)
    std::mutex lock;
    std::condition_variable wake_writer, wake_producer;
    std::deque<Job> queue;
    std::vector<std::thread> writers;
    std::unordered_map<uint64_t, int> block_state; // 0 = being written, 1 = stored (m_BlockState)
    std::vector<BlockRec> added;                   // in put order
    uint32_t in_flight = 0;
    std::atomic<bool> stop{false};
    int first_error = 0;
    uint64_t blocks_written = 0, bytes_written = 0, blocks_skipped = 0;
    size_t queue_limit = 32;
    // the copy of a block image out of the caller's staging memory is shared between the sink thread and a few copier threads
    struct CopyPart { uint8_t* dst; const uint8_t* src; size_t size; std::atomic<uint32_t>* left; };
    std::mutex copy_lock;
    std::condition_variable wake_copier;
    std::deque<CopyPart> copy_queue;
    std::vector<std::thread> copiers;
};

namespace {

std::string block_path(const lt_b200_fs_store* s, uint64_t hash, const char* ext)
{
    char name[16];
    hex16(hash, name);
    std::string p = s->root + "/chunks/";
    p.append(name, 4);
    p += "/0x";
    p.append(name, 16);
    p += ext;
    return p;
}

// SafeWriteStoredBlock (longtail_fsblockstore.c:243-320)
int write_block_file(lt_b200_fs_store* s, uint64_t hash, const void* data, size_t size, bool* skipped)
{
    const std::string final_path = block_path(s, hash, ".lrb");
    *skipped = false;
    if (is_file(final_path))
    {
        *skipped = true; // the block exists, only the store index was out of sync
        return 0;
    }
    const std::string tmp_path = block_path(s, hash, s->tmp_ext);
    int err = mkdir_p(tmp_path.substr(0, tmp_path.find_last_of('/')));
    if (err) return err;
    err = write_all(tmp_path, data, size);
    if (err)
    {
        unlink(tmp_path.c_str()); // a short write (disk full) must not leave the temporary file behind
        return err;
    }
    if (rename(tmp_path.c_str(), final_path.c_str()) != 0)
    {
        err = errno;
        unlink(tmp_path.c_str());
        if (is_file(final_path)) return 0; // someone beat us to it
        return err;
    }
    return 0;
}

void finish_job(lt_b200_fs_store* s, const Job& j, int err, bool skipped)
{
    std::lock_guard<std::mutex> g(s->lock);
    if (err)
    {
        if (!s->first_error) s->first_error = err;
        s->block_state.erase(j.hash);
    }
    else
    {
        s->block_state[j.hash] = 1;
        if (skipped) ++s->blocks_skipped;
        else { ++s->blocks_written; s->bytes_written += j.size; }
    }
    --s->in_flight;
    s->wake_producer.notify_all();
}

void writer_main(lt_b200_fs_store* s)
{
    for (;;)
    {
        Job j;
        {
            std::unique_lock<std::mutex> g(s->lock);
            s->wake_writer.wait(g, [s] { return s->stop || !s->queue.empty(); });
            if (s->queue.empty()) return;
            j = s->queue.front();
            s->queue.pop_front();
            s->wake_producer.notify_all();
        }
        bool skipped = false;
        const int err = write_block_file(s, j.hash, j.data, j.size, &skipped);
        free(j.data);
        finish_job(s, j, err, skipped);
    }
}

void copier_main(lt_b200_fs_store* s)
{
    for (;;)
    {
        lt_b200_fs_store::CopyPart c;
        {
            std::unique_lock<std::mutex> g(s->copy_lock);
            s->wake_copier.wait(g, [s] { return s->stop || !s->copy_queue.empty(); });
            if (s->copy_queue.empty()) return;
            c = s->copy_queue.front();
            s->copy_queue.pop_front();
        }
        memcpy(c.dst, c.src, c.size);
        c.left->fetch_sub(1, std::memory_order_acq_rel);
    }
}

// dst[0..size) = src[0..size) with the copier threads' help; returns when the whole image has been copied
void copy_shared(lt_b200_fs_store* s, uint8_t* dst, const uint8_t* src, size_t size)
{
    const size_t parts = s->copiers.empty() || size < (2u << 20) ? 1 : s->copiers.size() + 1;
    if (parts == 1)
    {
        memcpy(dst, src, size);
        return;
    }
    const size_t per = ((size + parts - 1) / parts + 4095) & ~(size_t)4095;
    std::atomic<uint32_t> left(0);
    uint32_t queued = 0;
    {
        std::lock_guard<std::mutex> g(s->copy_lock);
        for (size_t off = per; off < size; off += per)
        {
            s->copy_queue.push_back({dst + off, src + off, size - off < per ? size - off : per, &left});
            ++queued;
        }
        left.store(queued);
    }
    s->wake_copier.notify_all();
    memcpy(dst, src, per < size ? per : size);
    while (left.load(std::memory_order_acquire)) std::this_thread::yield();
}

void drain(lt_b200_fs_store* s)
{
    std::unique_lock<std::mutex> g(s->lock);
    s->wake_producer.wait(g, [s] { return s->in_flight == 0; });
}

} // namespace

extern "C" int lt_b200_fs_store_open(const char* store_path, uint32_t writer_threads, lt_b200_fs_store** out_store)
{
    if (!store_path || !out_store) return EINVAL;
    lt_b200_fs_store* s = new (std::nothrow) lt_b200_fs_store();
    if (!s) return ENOMEM;
    s->root = store_path;
    while (s->root.size() > 1 && s->root.back() == '/') s->root.pop_back();
    int err = mkdir_p(s->root);
    if (err) { delete s; return err; }
    // unique per store object: two processes writing the same block never share a temporary name
    uint64_t id = ((uint64_t)getpid() << 32) ^ (uint64_t)reinterpret_cast<uintptr_t>(s) ^ ((uint64_t)time(nullptr) << 20);
    s->tmp_ext[0] = '.';
    hex16(id, s->tmp_ext + 1);
    s->tmp_ext[17] = 0;
    s->queue_limit = writer_threads ? 4 * (size_t)writer_threads : 0;
    for (uint32_t i = 0; i < writer_threads; ++i) s->writers.emplace_back(writer_main, s);
    for (uint32_t i = 0; i < (writer_threads >= 4 ? 3u : 0u); ++i) s->copiers.emplace_back(copier_main, s);
    *out_store = s;
    return 0;
}

// lt_b200_block_sink: `user` is the store.  The byte image is only valid during the call, so it is copied before it is queued.
extern "C" int lt_b200_fs_store_sink(void* user, const struct lt_b200_stored_block_view* block)
{
    lt_b200_fs_store* s = static_cast<lt_b200_fs_store*>(user);
    if (!s || !block || !block->data) return EINVAL;
    const uint32_t n = block->chunk_count;
    if (block->size < 20 + 12 * (uint64_t)n) return EBADF;
    const uint8_t* p = static_cast<const uint8_t*>(block->data);
    BlockRec rec; // the block index at the front of the image: u64 hash, u32 hash id, u32 chunk count, u32 tag, u64[n], u32[n] (src/longtail.c:3585-3637)
    memcpy(&rec.hash, p, 8);
    memcpy(&rec.hash_id, p + 8, 4);
    uint32_t n_in = 0;
    memcpy(&n_in, p + 12, 4);
    memcpy(&rec.tag, p + 16, 4);
    if (rec.hash != block->block_hash || n_in != n) return EBADF;
    rec.chunk_hashes.resize(n);
    rec.chunk_sizes.resize(n);
    memcpy(rec.chunk_hashes.data(), p + 20, 8 * (size_t)n);
    memcpy(rec.chunk_sizes.data(), p + 20 + 8 * (size_t)n, 4 * (size_t)n);
    {
        std::unique_lock<std::mutex> g(s->lock);
        if (s->first_error) return s->first_error;
        if (s->block_state.count(rec.hash)) return 0; // already being written or stored by this store object (:791-800)
        s->block_state[rec.hash] = 0;
        s->added.push_back(std::move(rec));
        ++s->in_flight;
    }
    Job j = {block->block_hash, nullptr, (size_t)block->size};
    if (s->writers.empty())
    {
        bool skipped = false;
        const int err = write_block_file(s, j.hash, block->data, j.size, &skipped);
        finish_job(s, j, err, skipped);
        return err;
    }
    j.data = malloc(j.size ? j.size : 1);
    if (!j.data)
    {
        finish_job(s, j, ENOMEM, false);
        return ENOMEM;
    }
    copy_shared(s, static_cast<uint8_t*>(j.data), static_cast<const uint8_t*>(block->data), j.size);
    {
        std::unique_lock<std::mutex> g(s->lock);
        s->wake_producer.wait(g, [s] { return s->queue.size() < s->queue_limit; });
        s->queue.push_back(j);
    }
    s->wake_writer.notify_one();
    return 0;
}

extern "C" int lt_b200_fs_store_flush(lt_b200_fs_store* s)
{
    if (!s) return EINVAL;
    drain(s);
    // The records leave s->added only once store.lsi holds them: taken out here, put back in front when anything below fails, so that a
    // retried flush indexes them.  Records of blocks that have not reached the disk yet (a sink that ran after drain) stay for the next flush.
    std::vector<BlockRec> added, later;
    {
        std::lock_guard<std::mutex> g(s->lock);
        if (s->first_error) return s->first_error;
        for (auto& r : s->added)
        {
            auto it = s->block_state.find(r.hash);
            if (it != s->block_state.end() && it->second == 1) added.push_back(std::move(r)); // only blocks that reached the disk enter the index
            else later.push_back(std::move(r));
        }
        s->added.swap(later);
    }
    auto give_back = [&](int e) {
        std::lock_guard<std::mutex> g(s->lock);
        s->added.insert(s->added.begin(), std::make_move_iterator(added.begin()), std::make_move_iterator(added.end()));
        return e;
    };
    const std::string index_path = s->root + "/store.lsi";
    if (added.empty() && is_file(index_path)) return 0;
    const std::string lock_path = s->root + "/store.lsi.sync";
    const int lock_fd = open(lock_path.c_str(), O_RDWR | O_CREAT, 0666);
    if (lock_fd < 0) return give_back(errno);
    if (flock(lock_fd, LOCK_EX) != 0) { const int e = errno; close(lock_fd); return give_back(e); }
    int err = 0;
    Index add_ix, disk_ix, merged;
    add_ix.version = LT_B200_STORE_INDEX_VERSION;
    for (const auto& r : added)
    {
        if (!add_ix.hash_id) add_ix.hash_id = r.hash_id; // Longtail_CreateStoreIndexFromBlocks, src/longtail.c:9080-9085
        append_block(&add_ix, r.hash, r.tag, r.chunk_hashes.data(), r.chunk_sizes.data(), (uint32_t)r.chunk_hashes.size());
    }
    if (is_file(index_path))
    {
        std::vector<uint8_t> buf;
        err = read_all(index_path, &buf);
        if (!err) err = parse_index(buf, &disk_ix);
    }
    if (!err) err = merge_index(add_ix, disk_ix, &merged);
    if (!err)
    {
        std::vector<uint8_t> image;
        serialise_index(merged, &image);
        const std::string tmp_path = s->root + "/store" + s->tmp_ext;
        err = write_all(tmp_path, image.data(), image.size());
        if (!err && is_file(index_path) && unlink(index_path.c_str()) != 0) err = errno;
        if (!err && rename(tmp_path.c_str(), index_path.c_str()) != 0) err = errno;
        if (err) unlink(tmp_path.c_str());
    }
    flock(lock_fd, LOCK_UN);
    close(lock_fd);
    return err ? give_back(err) : 0;
}

extern "C" int lt_b200_fs_store_existing_chunks(lt_b200_fs_store* s, uint64_t* out_hashes, uint32_t capacity, uint32_t* out_count)
{
    if (!s || !out_count) return EINVAL;
    *out_count = 0;
    const std::string index_path = s->root + "/store.lsi";
    if (!is_file(index_path)) return 0;
    std::vector<uint8_t> buf;
    int err = read_all(index_path, &buf);
    if (err) return err;
    Index ix;
    err = parse_index(buf, &ix);
    if (err) return err;
    *out_count = (uint32_t)ix.chunk_hashes.size();
    if (out_hashes)
    {
        if (capacity < ix.chunk_hashes.size()) return ENOMEM;
        memcpy(out_hashes, ix.chunk_hashes.data(), 8 * ix.chunk_hashes.size());
    }
    return 0;
}

extern "C" int lt_b200_fs_store_stats(lt_b200_fs_store* s, uint64_t* out_blocks_written, uint64_t* out_bytes_written, uint64_t* out_blocks_skipped)
{
    if (!s) return EINVAL;
    std::lock_guard<std::mutex> g(s->lock);
    if (out_blocks_written) *out_blocks_written = s->blocks_written;
    if (out_bytes_written) *out_bytes_written = s->bytes_written;
    if (out_blocks_skipped) *out_blocks_skipped = s->blocks_skipped;
    return 0;
}

extern "C" int lt_b200_fs_store_close(lt_b200_fs_store* s)
{
    if (!s) return EINVAL;
    int err = lt_b200_fs_store_flush(s);
    {
        std::lock_guard<std::mutex> g(s->lock);
        s->stop = true;
    }
    s->wake_writer.notify_all();
    for (auto& t : s->writers) t.join();
    {
        std::lock_guard<std::mutex> g(s->copy_lock); // stop is read under this lock by the copiers
    }
    s->wake_copier.notify_all();
    for (auto& t : s->copiers) t.join();
    delete s;
    return err;
}
