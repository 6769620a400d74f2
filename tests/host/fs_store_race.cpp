// fs_store_race.cpp — test infrastructure: two store objects (two uploading processes, here two threads) put overlapping blocks into one
// directory through lt_b200_fs_store_*; built by tests/test_fs_store.py with -fsanitize=thread and with -fsanitize=address,undefined.
#include "../../include/longtail_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

static std::vector<uint8_t> image(uint64_t hash, uint32_t chunks, size_t payload)
{
    std::vector<uint8_t> v(20 + 12 * (size_t)chunks + payload, 7);
    const uint32_t hash_id = 0x626c6b33u, tag = 0;
    memcpy(&v[0], &hash, 8);
    memcpy(&v[8], &hash_id, 4);
    memcpy(&v[12], &chunks, 4);
    memcpy(&v[16], &tag, 4);
    for (uint32_t i = 0; i < chunks; ++i)
    {
        const uint64_t chunk_hash = hash * 131 + i;
        const uint32_t size = 100;
        memcpy(&v[20 + 8 * (size_t)i], &chunk_hash, 8);
        memcpy(&v[20 + 8 * (size_t)chunks + 4 * (size_t)i], &size, 4);
    }
    return v;
}

int main(int argc, char** argv)
{
    if (argc < 2) return 64;
    const char* root = argv[1];
    auto work = [&](int base) {
        lt_b200_fs_store* s = nullptr;
        if (lt_b200_fs_store_open(root, 4, &s)) exit(3);
        for (int i = 0; i < 60; ++i)
        {
            const std::vector<uint8_t> img = image(1000ull + base + i, 5, 3u << 20);
            lt_b200_stored_block_view v = {1000ull + base + i, img.data(), img.size(), 5, 0, 0, 0};
            if (lt_b200_fs_store_sink(s, &v)) exit(4);
            if (i % 20 == 19 && lt_b200_fs_store_flush(s)) exit(5);
        }
        if (lt_b200_fs_store_close(s)) exit(6);
    };
    std::thread a(work, 0), b(work, 30); // blocks 1030..1059 are put by both
    a.join();
    b.join();
    lt_b200_fs_store* s = nullptr;
    if (lt_b200_fs_store_open(root, 0, &s)) return 7;
    uint32_t n = 0;
    lt_b200_fs_store_existing_chunks(s, nullptr, 0, &n);
    lt_b200_fs_store_close(s);
    printf("chunks listed: %u\n", n);
    return n == 90 * 5 ? 0 : 2;
}
