"""end-to-end upsync of a real directory tree on a RAM-backed file system, this repository vs the unmodified reference on the same box:
    python tests/tools_bench_upsync_dir.py [GiB] [files]
scan -> read -> CreateVersionIndex -> CreateMissingContent -> WriteContent (LZ4) -> fsblockstore directory + store.lsi.
Both sides read the same files from /dev/shm and write their store there, so the number compares the pipelines, not a disk."""
import os
import shutil
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import longtail_b200  # noqa: E402
import oracle_lib as ol  # noqa: E402  (the reference arm of this measurement; it lives under tests/ because only tests may run the checker)

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
nfiles = int(sys.argv[2]) if len(sys.argv) > 2 else 64
base = tempfile.mkdtemp(prefix="lt_b200_upsync_", dir="/dev/shm")
src, ours, theirs = os.path.join(base, "src"), os.path.join(base, "ours"), os.path.join(base, "ref")
try:
    ctx = longtail_b200.Context(0)
    per = int(gib * (1 << 30) / nfiles) // 256 * 256
    dev = ctx.device_alloc(per + 4096)
    t0 = time.perf_counter()
    for i in range(nfiles):
        d = os.path.join(src, "d%02d" % (i % 8))
        os.makedirs(d, exist_ok=True)
        # the configs[2] generator: half of the 1 MiB segments shared between files, random / 4-bit / text-like classes
        ctx.synth_fill(dev, per, seed=5, asset_id=i, class_mode=1, shared_permille=500, pool_segments=256)
        ctx.to_host(dev, per).tofile(os.path.join(d, "f%04d.bin" % i))
    ctx.device_free(dev)
    total = per * nfiles
    print("tree: %d files, %.2f GiB in %.1f s" % (nfiles, total / 2**30, time.perf_counter() - t0))
    for run in range(2):  # run 1 is the reported one (workspace growth and page cache warm)
        shutil.rmtree(ours, ignore_errors=True)
        t0 = time.perf_counter()
        files = longtail_b200.FileList(src, threads=16)
        t1 = time.perf_counter()
        store = longtail_b200.FsStore(ours, writer_threads=12)
        vi, blocks = ctx.upsync_file_list(files, store, [longtail_b200.COMPRESSION_LZ4] * len(files.paths), target_chunk_size=65536, reader_threads=16)
        store.close()
        files.close()
        t2 = time.perf_counter()
    print("b200     : scan %.3f s, upsync %.2f s -> %.2f GiB/s (%d blocks)" % (t1 - t0, t2 - t1, total / (t2 - t0) / 2**30, blocks))
    ref = ol.Reference()
    if ref.available:
        cores = ref.cpu_count()
        t0 = time.perf_counter()
        want_vi, want_blocks = ol.ref_upsync_dir_to_dir(ref, src, theirs, 65536, workers=cores, tag=ol.COMP_LZ4)
        t1 = time.perf_counter()
        print("reference: %.2f s -> %.2f GiB/s (%d blocks, %d workers)" % (t1 - t0, total / (t1 - t0) / 2**30, want_blocks, cores))
        same_vi = vi == want_vi
        a = {os.path.relpath(os.path.join(d, f), ours) for d, _, fs in os.walk(ours) for f in fs if f.endswith(".lrb")}
        b = {os.path.relpath(os.path.join(d, f), theirs) for d, _, fs in os.walk(theirs) for f in fs if f.endswith(".lrb")}
        sizes_equal = all(os.path.getsize(os.path.join(ours, k)) == os.path.getsize(os.path.join(theirs, k)) for k in a & b)
        print("parity   : VersionIndex %s, block files %s (%d), sizes %s" % ("identical" if same_vi else "DIFFERENT", "same set" if a == b else "DIFFERENT", len(a),
                                                                              "equal" if sizes_equal else "DIFFERENT"))
    ctx.close()
finally:
    shutil.rmtree(base, ignore_errors=True)
