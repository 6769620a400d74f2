"""CPU (-m "not gpu"): the C planners of the multi-GPU verbs — lt_b200_plan_shards / lt_b200_shard_jobs (host code of liblongtail_b200.so, no
GPU needed) follow the reference's part rule (src/longtail.c:2396-2457) and cut the job list into contiguous, byte-balanced slices; and the
greedy block packer lt_b200_pack_blocks that every rank runs over the global unique-chunk list equals the oracle's packing."""
import numpy as np
import pytest

import longtail_b200
from dist_model import plan_jobs


def _assets(sizes):
    return longtail_b200.AssetList(["d/%04d.bin" % i for i in range(len(sizes))], sizes)


@pytest.mark.parametrize("target", [16, 64, 65536])
def test_jobs_follow_the_reference_part_rule(target):
    part = target * 1024
    sizes = [0, 1, part - 1, part, part + 1, 3 * part, 5 * part + 7, 0, 2 * part]
    al = _assets(sizes)
    first, n = longtail_b200.plan_shards(al, target, 1)
    want = plan_jobs(sizes, target)  # 1 + size/part parts per asset, empty ones dropped
    assert n == len(want) and first.tolist() == [0, n]
    got = longtail_b200.shard_jobs(al, target, 0, n)
    assert [(int(j["asset_index"]), int(j["offset"]), int(j["size"])) for j in got] == want


@pytest.mark.parametrize("world", [1, 2, 3, 8, 64])
def test_slices_are_contiguous_and_balanced(world):
    rng = np.random.RandomState(7)
    target = 64
    part = target * 1024
    sizes = [int(x) for x in rng.randint(0, 40 * part, size=200)]
    al = _assets(sizes)
    first, n = longtail_b200.plan_shards(al, target, world)
    assert first[0] == 0 and first[-1] == n and all(first[i] <= first[i + 1] for i in range(world))
    jobs = longtail_b200.shard_jobs(al, target, 0, n)
    total = int(jobs["size"].astype(np.uint64).sum())
    assert total == sum(sizes)
    per = [int(jobs["size"][first[r]:first[r + 1]].astype(np.uint64).sum()) for r in range(world)]
    assert sum(per) == total
    assert max(per) - total / world <= part  # no slice is more than one part above its share


def test_single_file_splits_on_part_boundaries():
    """configs[3]: ONE file sharded by byte range — slices of whole parts, so no overlap stitch is needed"""
    target = 65536
    part = target * 1024
    al = _assets([8 * 4 * part])
    first, n = longtail_b200.plan_shards(al, target, 8)
    assert n == 32 and first.tolist() == [4 * r for r in range(9)]
    jobs = longtail_b200.shard_jobs(al, target, int(first[3]), 4)
    assert jobs["offset"].tolist() == [(12 + k) * part for k in range(4)] and set(jobs["size"].tolist()) == {part}


def test_pack_blocks_matches_oracle_upsync(oracle):
    import ctypes as C

    from synth import synth_bytes
    lib = longtail_b200.load_library()
    assets = [("p/%02d.bin" % i, synth_bytes(700 + 3 * i, 90000 + 41000 * i, ["rand", "nib", "text"][i % 3])) for i in range(9)]
    tags = [0x6c7a3432 if i % 4 else 0 for i in range(9)]
    blocks, v = oracle.upsync(assets, 512, max_block_size=65536, max_chunks_per_block=48, tags=tags)
    vi = longtail_b200.parse_version_index(v)
    n = vi["chunk_count"]
    sizes = np.ascontiguousarray(vi["chunk_sizes"], dtype=np.uint32)
    ctags = np.ascontiguousarray(vi["chunk_tags"], dtype=np.uint32)
    first, count = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    nb = C.c_uint32(0)
    assert lib.lt_b200_pack_blocks(C.c_uint32(n), sizes.ctypes.data_as(C.c_void_p), ctags.ctypes.data_as(C.c_void_p), C.c_uint32(65536), C.c_uint32(48),
                                   first.ctypes.data_as(C.c_void_p), count.ctypes.data_as(C.c_void_p), C.byref(nb)) == 0
    assert nb.value == len(blocks)
    import struct
    for b, (h, blob) in enumerate(blocks):  # block index: u64 hash, u32 hash id, u32 chunk count, u32 tag
        _, _, cnt, tag = struct.unpack_from("<QIII", blob, 0)
        assert cnt == count[b] and tag == ctags[first[b]]


def _pack(lib, sizes, tags, max_block, per_block):
    import ctypes as C
    n = len(sizes)
    s = np.ascontiguousarray(sizes, dtype=np.uint32)
    t = np.ascontiguousarray(tags, dtype=np.uint32)
    first, count = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.uint32)
    nb = C.c_uint32(0)
    assert lib.lt_b200_pack_blocks(C.c_uint32(n), s.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), C.c_uint32(max_block), C.c_uint32(per_block),
                                   first.ctypes.data_as(C.c_void_p), count.ctypes.data_as(C.c_void_p), C.byref(nb)) == 0
    return count[:nb.value].tolist()


@pytest.mark.parametrize("seed,max_block,per_block", [(1, 65536, 48), (2, 100000, 7), (3, 8388608, 1024), (4, 40000, 3)])
def test_packing_is_stable_under_streaming(seed, max_block, per_block):
    """what lt_b200_upsync_stream_host_assets relies on: the greedy packing (src/longtail.c:6796-6860) is a left to right scan, so packing
    the chunks as they arrive, closing every block but the last of each round and carrying that one's chunks into the next round, gives
    the blocks of ONE packing over the whole list — whatever the batch cuts are (empty rounds and one-chunk rounds included)"""
    lib = longtail_b200.load_library()
    rng = np.random.RandomState(seed)
    n = 5000
    sizes = rng.randint(1, 30000, size=n).astype(np.uint32)
    tags = np.repeat(rng.choice(np.array([0, 0x6c7a3432, 0x7a746432], np.uint32), size=n // 50 + 1), 50)[:n]
    want = _pack(lib, sizes, tags, max_block, per_block)
    cuts = sorted(set([0, n] + rng.randint(0, n, size=40).tolist() + [17, 17, 18]))
    got, pending = [], []  # pending = indices of the chunks of the open block
    for a, b in zip(cuts[:-1] + [n], cuts[1:] + [n]):  # the last round (n, n) adds nothing and closes everything
        pending += list(range(a, b))
        last = (a, b) == (n, n)
        counts = _pack(lib, sizes[pending], tags[pending], max_block, per_block) if pending else []
        closed = counts if last else counts[:-1]
        got += closed
        pending = pending[sum(closed):]
        assert sum(sizes[pending].astype(np.uint64)) <= max_block + max_block // 10 or len(pending) == 1
    assert not pending and got == want
