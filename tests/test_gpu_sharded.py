"""GPU parity (-m gpu) of the multi-GPU verbs on ONE GPU: a communicator of world 1 runs the whole sharded code path (job plan, table
exchange skipped, split dedup degenerate, block plan, no chunk exchange) — lt_b200_index_sharded + lt_b200_write_blocks_sharded must give
the reference's VersionIndex and StoredBlocks.  The 2-GPU run of the same verbs is tests/test_gpu_multi.py."""
import numpy as np
import pytest

import oracle_lib as ol
from synth import synth_bytes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", [ol.COMP_LZ4, ol.COMP_ZSTD_DEFAULT if hasattr(ol, "COMP_ZSTD_DEFAULT") else 0x7a746432, 0])
def test_sharded_verbs_world_1(oracle, reference, tag):
    import longtail_b200
    ctx = longtail_b200.Context(0)
    comm = ctx.comm_create(ctx.comm_unique_id(), 0, 1)
    try:
        target = 4096
        assets = [("s/%02d.bin" % i, synth_bytes(40 + 5 * i, 200000 + 1230000 * i, ["rand", "nib", "text", "rec"][i % 4])) for i in range(7)]
        assets.append(("s/dup.bin", assets[2][1].copy()))
        assets.append(("s/empty.bin", synth_bytes(1, 0)))
        tags = [tag] * len(assets)
        al = longtail_b200.AssetList([p for p, _ in assets], [d.size for _, d in assets])
        offs, off = [], 0
        for _, d in assets:
            offs.append(off)
            off = (off + d.size + 255) & ~255
        arena = np.zeros(off + 4096, np.uint8)
        for o, (_, d) in zip(offs, assets):
            arena[o:o + d.size] = d
        ptr = ctx.device_alloc(arena.size)
        ctx.to_device(ptr, arena)
        first, n = ctx.plan_shards(al, target, 1)
        jobs = ctx.shard_jobs(al, target, 0, n)
        job_offs = np.asarray(offs, dtype=np.uint64)[jobs["asset_index"]] + jobs["offset"]
        v = ctx.index_sharded(comm, ptr, arena.size, al, tags, job_offs, target)
        checker = reference if reference is not None else oracle
        want_blocks, want_v = checker.upsync(assets, target, tags=tags)
        assert v == want_v
        blocks = []
        import ctypes as C

        def sink(_user, view):
            b = view.contents
            blocks.append((int(b.block_hash), C.string_at(b.data, b.size)))
            return 0

        cb = longtail_b200.BLOCK_SINK(sink)
        mine, total = ctx.write_blocks_sharded(comm, (C.cast(cb, C.c_void_p), None))
        ctx.device_free(ptr)
        assert mine == total == len(want_blocks)
        assert blocks == want_blocks
    finally:
        ctx.comm_destroy(comm)
        ctx.close()


@pytest.mark.parametrize("max_block_size,max_chunks", [(65536, 7), (262144, 1024), (8388608, 3), (1, 1024)])
def test_device_block_packing_matches_reference(oracle, reference, max_block_size, max_chunks):
    """lt_b200_write_blocks_sharded packs on the device (next() by binary search + pointer doubling): block composition for small block
    sizes / chunk limits, tag changes and single-chunk blocks must equal Longtail_CreateStoreIndex's greedy loop (src/longtail.c:6796-6860)"""
    import ctypes as C

    import longtail_b200
    ctx = longtail_b200.Context(0)
    comm = ctx.comm_create(ctx.comm_unique_id(), 0, 1)
    try:
        target = 1024
        kinds = ["rand", "nib", "text", "rec"]
        assets = [("p/%02d.bin" % i, synth_bytes(140 + 7 * i, 30000 + 91000 * i, kinds[i % 4])) for i in range(10)]
        tags = [[ol.COMP_LZ4, 0, ol.COMP_ZSTD_DEFAULT][(i // 2) % 3] for i in range(len(assets))]  # tag runs of two assets
        al = longtail_b200.AssetList([p for p, _ in assets], [d.size for _, d in assets])
        offs, off = [], 0
        for _, d in assets:
            offs.append(off)
            off = (off + d.size + 255) & ~255
        arena = np.zeros(off + 4096, np.uint8)
        for o, (_, d) in zip(offs, assets):
            arena[o:o + d.size] = d
        ptr = ctx.device_alloc(arena.size)
        ctx.to_device(ptr, arena)
        first, n = ctx.plan_shards(al, target, 1)
        jobs = ctx.shard_jobs(al, target, 0, n)
        job_offs = np.asarray(offs, dtype=np.uint64)[jobs["asset_index"]] + jobs["offset"]
        v = ctx.index_sharded(comm, ptr, arena.size, al, tags, job_offs, target)
        checker = reference if reference is not None else oracle
        want_blocks, want_v = checker.upsync(assets, target, max_block_size=max_block_size, max_chunks_per_block=max_chunks, tags=tags)
        assert v == want_v
        blocks = []

        def sink(_user, view):
            b = view.contents
            blocks.append((int(b.block_hash), C.string_at(b.data, b.size)))
            return 0

        cb = longtail_b200.BLOCK_SINK(sink)
        mine, total = ctx.write_blocks_sharded(comm, (C.cast(cb, C.c_void_p), None), max_block_size=max_block_size, max_chunks_per_block=max_chunks)
        ctx.device_free(ptr)
        assert mine == total == len(want_blocks)
        assert blocks == want_blocks
    finally:
        ctx.comm_destroy(comm)
        ctx.close()
