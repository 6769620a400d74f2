#!/usr/bin/env python
"""The multi-GPU verbs of the C ABI under torchrun, one process per GPU (Python is only the launcher glue):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/tools_multi_gpu_upsync.py --gib 0.25 --codec lz4

Every rank holds the bytes of its slice of the job list of ONE version (configs[2] generator: random / 4-bit / text-like segments, half of them from a pool all
ranks share, so chunks deduplicate ACROSS ranks).  lt_b200_index_sharded (chunk + hash of the rank's jobs, NCCL all-gather of the tables,
dedup split by hash) and lt_b200_write_blocks_sharded (global block plan, chunk exchange, per-rank WriteContent) run on all ranks; rank 0
then runs the unmodified reference's single-process upsync over all ranks' bytes: the VersionIndex must be memcmp-identical and the
StoredBlocks of all ranks, concatenated in rank order, byte-identical to the reference's block list.  Prints VERIFY OK on stderr."""
import argparse
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
GIB = 1 << 30
TARGET = 65536


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=0.25, help="asset bytes per GPU")
    ap.add_argument("--codec", default="lz4", choices=["lz4", "zstd", "none"])
    ap.add_argument("--single-file", action="store_true", help="configs[3] shape: ONE PAK-like file, byte ranges of whole parts per rank")
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist

    import longtail_b200
    import oracle_lib as ol

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = longtail_b200.Context(local_rank)
    box = [ctx.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    comm = ctx.comm_create(box[0], rank, world)
    tag = {"lz4": longtail_b200.COMPRESSION_LZ4, "zstd": longtail_b200.COMPRESSION_ZSTD_DEFAULT, "none": 0}[args.codec]
    part = TARGET * 1024

    # every rank knows the whole version (names, sizes, generator ids) and fills the bytes of the jobs the plan hands it
    if args.single_file:
        per = max(part, int(args.gib * GIB) // part * part)
        names, sizes, ids = ["pak/data.pak"], [per * world + 12345], [0]
        spec = dict(seed=3, class_mode=2)
    else:
        count = 6
        sizes = [int(args.gib * GIB / count) // 256 * 256 + 4096 * (r + 1) + 77 * i for r in range(world) for i in range(count)]  # ragged
        names = ["r%d/a%02d.bin" % (r, i) for r in range(world) for i in range(count)]
        ids = list(range(world * count))
        spec = dict(seed=2, class_mode=1, shared_permille=500, pool_segments=max(8, int(args.gib * 256)))
    al = longtail_b200.AssetList(names, sizes)
    first, n = ctx.plan_shards(al, TARGET, world)
    jobs = ctx.shard_jobs(al, TARGET, int(first[rank]), int(first[rank + 1] - first[rank]))
    job_offs, off = [], 0
    for j in jobs:
        job_offs.append(off)
        off += (int(j["size"]) + 255) & ~255
    arena_bytes = off + 4096
    arena = ctx.device_alloc(arena_bytes)
    for j, o in zip(jobs, job_offs):
        ctx.synth_fill(arena + o, int(j["size"]), asset_id=ids[int(j["asset_index"])], offset=int(j["offset"]), **spec)
    job_offs = np.asarray(job_offs, dtype=np.uint64)
    ctx.synchronize()

    tags = [tag] * len(names)
    v = ctx.index_sharded(comm, arena, arena_bytes, al, tags, job_offs, TARGET, want_host=True)
    blocks = []

    def sink(_user, view):
        b = view.contents
        blocks.append((int(b.block_hash), C.string_at(b.data, b.size)))
        return 0

    cb = longtail_b200.BLOCK_SINK(sink)
    n_mine, n_total = ctx.write_blocks_sharded(comm, (C.cast(cb, C.c_void_p), None))
    assert n_mine == len(blocks)
    if world > 1:
        gathered_blocks = [None] * world if rank == 0 else None
        dist.gather_object(blocks, gathered_blocks, dst=0)
        vs = [None] * world if rank == 0 else None
        dist.gather_object(bytes(v), vs, dst=0)
    else:
        gathered_blocks, vs = [blocks], [bytes(v)]
    if rank == 0:
        ref = ol.Reference()
        checker = ref if ref.available else ol.Oracle()
        # the same bytes once more, made on the host by the same generator (include/lt_synth.h)
        import subprocess
        lib_path = os.path.join(ROOT, "oracle", "_ref", "libsynth_host.so")
        if not os.path.exists(lib_path):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        lib = C.CDLL(lib_path)
        hspec = longtail_b200.SynthSpec(spec["seed"], spec.get("shared_permille", 0), spec.get("pool_segments", 1), spec["class_mode"], 0)
        datas = []
        for sz, i in zip(sizes, ids):
            buf = np.empty(sz, dtype=np.uint8)
            lib.synth_fill_mt(C.byref(hspec), C.c_uint64(i), C.c_uint64(0), buf.ctypes.data_as(C.c_void_p), C.c_uint64(sz), C.c_uint32(8))
            datas.append(buf)
        want_blocks, want_v = checker.upsync(list(zip(names, datas)), TARGET, tags=tags)
        assert all(x == want_v for x in vs), "VersionIndex differs from the %s (or between ranks)" % ("reference" if ref.available else "oracle")
        got = [b for r in gathered_blocks for b in r]
        assert n_total == len(want_blocks), "%d blocks planned, the checker wrote %d" % (n_total, len(want_blocks))
        assert [h for h, _ in got] == [h for h, _ in want_blocks], "block hashes / order differ"
        assert got == want_blocks, "StoredBlock bytes differ"
        print("VERIFY OK: %d ranks, %d-byte VersionIndex and %d StoredBlocks (%s; %s per rank) identical to the %s" % (
            world, len(want_v), len(got), args.codec, [len(r) for r in gathered_blocks], "reference" if ref.available else "oracle"), file=sys.stderr)
    ctx.comm_destroy(comm)
    ctx.device_free(arena)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
