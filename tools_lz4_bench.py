"""quick throughput probe of the device WriteContent half (not the headline bench): python tools_lz4_bench.py [gib] [class_mode] [permille] [lz4|zstd]"""
import sys, time
import numpy as np
import longtail_b200

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
perm = int(sys.argv[3]) if len(sys.argv) > 3 else 0
codec = sys.argv[4] if len(sys.argv) > 4 else "lz4"
TAG = longtail_b200.COMPRESSION_LZ4 if codec == "lz4" else longtail_b200.COMPRESSION_ZSTD_DEFAULT
KERNEL = "k_lz4_blocks" if codec == "lz4" else "k_zstd_frames"
n = int(gib * (1 << 30))
ctx = longtail_b200.Context(0)
ptr = ctx.device_alloc(n + 4096)
ctx.synth_fill(ptr, n, seed=2, class_mode=mode, shared_permille=perm, pool_segments=64)
ctx.synchronize()
nassets = 16
sz = n // nassets // 256 * 256
al = longtail_b200.AssetList(["a%03d.bin" % i for i in range(nassets)], [sz] * nassets)
offs = [i * sz for i in range(nassets)]
tags = [TAG] * nassets
for it in range(2):
    ctx.profile_reset(); ctx.profile_enable(True)
    t0 = time.perf_counter()
    v = ctx.index_device_assets(ptr, n + 4096, al, offs, tags, target_chunk_size=65536)
    t1 = time.perf_counter()
    ctx.profile_enable(False)
    print("  index kernels:", {k: round(x[0], 2) for k, x in ctx.profile_read().items() if x[0]})
    vi = longtail_b200.parse_version_index(v)
    uoff = ctx.unique_chunk_offsets(vi["chunk_count"])
    ctx.profile_reset(); ctx.profile_enable(True)
    t2 = time.perf_counter()
    blocks = ctx.write_blocks_device(ptr, n + 4096, vi["chunk_hashes"], vi["chunk_sizes"], vi["chunk_tags"], uoff, keep_bytes=False)
    t3 = time.perf_counter()
    ctx.profile_enable(False)
    prof = ctx.profile_read()
    uniq = int(vi["chunk_sizes"].astype(np.uint64).sum())
    stored = sum(s for _, s in blocks)
    print("iter %d: index %.1f ms (%.1f GiB/s); write %d blocks %.1f ms (%.2f GiB/s of unique %.2f GiB, ratio %.3f); codec kernel %.1f ms (%.2f GB/s), gather %.1f ms" % (
        it, 1e3 * (t1 - t0), n / (t1 - t0) / 2**30, len(blocks), 1e3 * (t3 - t2), uniq / (t3 - t2) / 2**30, uniq / 2**30, stored / max(uniq, 1),
        prof[KERNEL][0], prof[KERNEL][2] / max(prof[KERNEL][0], 1e-9) / 1e6, prof["k_gather_chunks"][0]))

if codec == "zstd":
    import ctypes as C
    t = (C.c_uint64 * 4)()
    ctx.lib.lt_b200_zstd_phase_cycles(ctx.handle, t)
    tot = float(sum(t)) or 1.0
    print("zstd phase cycles: matcher %.1f%%, literals %.1f%%, sequences %.1f%%, copy-out %.1f%%" % tuple(100.0 * x / tot for x in t))
