#!/usr/bin/env python
"""The multi-GPU verbs of the C ABI under torchrun, one process per GPU (Python is only the launcher glue):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/tools_multi_gpu_upsync.py --gib 0.25 --codec lz4

Every rank holds its own assets of ONE version (configs[2] generator: random / 4-bit / text-like segments, half of them from a pool all
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

    if args.single_file:
        per = max(part, int(args.gib * GIB) // part * part)
        names, sizes = ["pak/data.pak"], [per * world]
        arena_bytes = per + 4096
        arena = ctx.device_alloc(arena_bytes)
        ctx.synth_fill(arena, per, seed=3, asset_id=0, offset=rank * per, class_mode=2)
        al = longtail_b200.AssetList(names, sizes)
        first, n = ctx.plan_shards(al, TARGET, world)
        jobs = ctx.shard_jobs(al, TARGET, int(first[rank]), int(first[rank + 1] - first[rank]))
        job_offs = (jobs["offset"].astype(np.int64) - rank * per).astype(np.uint64)
        mine = [ctx.to_host(arena, per)]
    else:
        count = 6
        my_sizes = [int(args.gib * GIB / count) // 256 * 256 + 4096 * (rank + 1) + 77 * i for i in range(count)]  # ragged, different per rank
        all_sizes = [int(args.gib * GIB / count) // 256 * 256 + 4096 * (r + 1) + 77 * i for r in range(world) for i in range(count)]
        names = ["r%d/a%02d.bin" % (r, i) for r in range(world) for i in range(count)]
        sizes = all_sizes
        offs, off = [], 0
        for s in my_sizes:
            offs.append(off)
            off += (s + 255) & ~255
        arena_bytes = off + 4096
        arena = ctx.device_alloc(arena_bytes)
        pool = max(8, int(args.gib * 256))
        for i, (o, s) in enumerate(zip(offs, my_sizes)):
            ctx.synth_fill(arena + o, s, seed=2, asset_id=rank * count + i, class_mode=1, shared_permille=500, pool_segments=pool)
        al = longtail_b200.AssetList(names, sizes)
        first, n = ctx.plan_shards(al, TARGET, world)
        jobs = ctx.shard_jobs(al, TARGET, int(first[rank]), int(first[rank + 1] - first[rank]))
        local = jobs["asset_index"].astype(np.int64) - rank * count
        # the byte-balanced plan may hand a rank jobs of a neighbour's assets: this tool keeps it simple and requires asset-aligned slices
        if local.size and (local.min() < 0 or local.max() >= count):
            raise SystemExit("rank %d: the shard plan crosses the asset ownership of this test; use equal per-rank sizes" % rank)
        job_offs = np.asarray(offs, dtype=np.uint64)[local] + jobs["offset"]
        mine = [ctx.to_host(arena + o, s) for o, s in zip(offs, my_sizes)]
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
        gathered_data = [None] * world if rank == 0 else None
        dist.gather_object(blocks, gathered_blocks, dst=0)
        dist.gather_object(mine, gathered_data, dst=0)
        vs = [None] * world if rank == 0 else None
        dist.gather_object(bytes(v), vs, dst=0)
    else:
        gathered_blocks, gathered_data, vs = [blocks], [mine], [bytes(v)]
    if rank == 0:
        ref = ol.Reference()
        checker = ref if ref.available else ol.Oracle()
        if args.single_file:
            datas = [np.concatenate([d[0] for d in gathered_data])]
        else:
            datas = [d for r in gathered_data for d in r]
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
