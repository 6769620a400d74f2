"""upsync a directory into an fsblockstore-compatible store on the GPU (what `longtail upsync` does, cmd/main.c:972-1153):
    python tools_upsync_dir.py <source dir> <store dir> <version.lvi> [--tag lz42|ztd2|none] [--target-chunk-size 32768]
The store directory can then be opened by an unmodified longtail (Longtail_CreateFSBlockStoreAPI); a second run with a changed source
directory writes only the blocks of the chunks the store does not hold yet."""
import argparse
import time

import longtail_b200

TAGS = {"lz42": longtail_b200.COMPRESSION_LZ4, "ztd2": longtail_b200.COMPRESSION_ZSTD_DEFAULT, "none": 0}


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("source")
    ap.add_argument("store")
    ap.add_argument("version_index")
    ap.add_argument("--tag", default="lz42", choices=sorted(TAGS))
    ap.add_argument("--target-chunk-size", type=int, default=32768)
    ap.add_argument("--max-block-size", type=int, default=8388608)
    ap.add_argument("--max-chunks-per-block", type=int, default=1024)
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--gpu", type=int, default=0)
    a = ap.parse_args()
    t0 = time.perf_counter()
    files = longtail_b200.FileList(a.source, threads=a.threads)
    t1 = time.perf_counter()
    ctx = longtail_b200.Context(a.gpu)
    store = longtail_b200.FsStore(a.store, writer_threads=max(1, a.threads // 2))
    vi, blocks = ctx.upsync_file_list(files, store, [TAGS[a.tag]] * len(files.paths), target_chunk_size=a.target_chunk_size,
                                      max_block_size=a.max_block_size, max_chunks_per_block=a.max_chunks_per_block, reader_threads=a.threads)
    stats = store.stats()
    store.close()
    with open(a.version_index, "wb") as f:
        f.write(vi)
    t2 = time.perf_counter()
    total = sum(files.sizes)
    v = longtail_b200.parse_version_index(vi)
    print("%d entries, %.2f GiB: scan %.2f s, upsync %.2f s (%.2f GiB/s); %d unique chunks, %d blocks written (%.2f GiB), version index %d bytes"
          % (len(files.paths), total / 2**30, t1 - t0, t2 - t1, total / max(t2 - t1, 1e-9) / 2**30, int(v["chunk_count"]), blocks,
             stats["bytes_written"] / 2**30, len(vi)))
    files.close()
    ctx.close()


if __name__ == "__main__":
    main()
