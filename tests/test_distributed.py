"""The N>1 host logic (tests/dist_model.py) on CPU: world_size 2, gloo.  Each rank produces the chunk table of its
slice of the job list with the CPU oracle (standing in for the GPU), the tables are merged with allgather_tables, and the
merged table must equal the single-process table in global job order."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_job_plan_matches_reference_part_rule():
    from dist_model import plan_jobs, shard_jobs
    part = 16 * 1024
    sizes = [0, 1, part - 1, part, part + 1, 3 * part, 5 * part + 7]
    jobs = plan_jobs(sizes, 16)
    # src/longtail.c:2402: 1 + size/part parts, the trailing empty one produces no chunks and is dropped
    assert [j for j in jobs if j[0] == 3] == [(3, 0, part)]
    assert [j[2] for j in jobs if j[0] == 6] == [part] * 5 + [7]
    assert sum(j[2] for j in jobs) == sum(sizes)
    for world in (1, 2, 3, 8, 64):
        shards = shard_jobs(jobs, world)
        assert len(shards) == world and shards[0][0] == 0 and shards[-1][1] == len(jobs)
        assert all(shards[i][1] == shards[i + 1][0] for i in range(world - 1))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, HERE)
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle_lib as ol
    import dist_model as ltd
    from synth import chunker_params, synth_bytes

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    target = 64
    part = target * 1024
    mn, av, mx = chunker_params(target)
    oracle = ol.Oracle()
    datas = [synth_bytes(40 + i, n, k) for i, (n, k) in enumerate([(3 * part + 5, "rand"), (0, "rand"), (part, "nib"), (200000, "text"), (2 * part, "rand")])]
    tags = [0, 7, 0, 9, 9]
    jobs = ltd.plan_jobs([d.size for d in datas], target)
    first, last = ltd.shard_jobs(jobs, world)[rank]
    counts, hashes, sizes, ctags = [], [], [], []
    for a, start, n in jobs[first:last]:
        chunk = datas[a][start:start + n]
        lens = oracle.chunk(chunk, mn, av, mx)
        counts.append(lens.size)
        offs = np.concatenate([[0], np.cumsum(lens.astype(np.uint64))[:-1]]).astype(np.uint64)
        hashes.extend(oracle.hash_segments(ol.HASH_BLAKE3, chunk, offs, lens).tolist())
        sizes.extend(lens.tolist())
        ctags.extend([tags[a]] * lens.size)
    t = lambda x, dt: torch.from_numpy(np.asarray(x, dtype=dt))
    jc, gh, gs, gt = ltd.allgather_tables(t(counts, np.int64), t(np.asarray(hashes, dtype=np.uint64).view(np.int64), np.int64),
                                          t(sizes, np.int32), t(ctags, np.int32))
    if rank == 0:
        acc = ltd.asset_chunk_counts(jobs, jc.numpy(), len(datas))
        np.savez(out_path, job_counts=jc.numpy(), hashes=gh.numpy().view(np.uint64), sizes=gs.numpy(), tags=gt.numpy(), asset_counts=acc)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_allgather_merge_equals_single_process(world, tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000
    merged = str(tmp_path / "merged.npz")
    single = str(tmp_path / "single.npz")
    mp.spawn(_worker, args=(world, port, merged), nprocs=world, join=True)
    mp.spawn(_worker, args=(1, port + 1, single), nprocs=1, join=True)
    a, b = np.load(merged), np.load(single)
    for k in ("job_counts", "hashes", "sizes", "tags", "asset_counts"):
        assert a[k].tolist() == b[k].tolist(), k
    assert a["hashes"].size > 20


# ---------------------------------------------------------------------------------------------------------------------
# WriteContent across ranks: plan_write + exchange_chunks.  Every rank serialises the blocks it owns (payloads assembled from its
# own chunks plus the foreign chunks it received); the union over ranks must equal the single-process upsync of the oracle.


def _serialise_block(oracle, ol, hashes, sizes, tag, payload):
    import struct
    hs = np.asarray(hashes, dtype="<u8")
    block_hash = oracle.hash(ol.HASH_BLAKE3, hs.tobytes())
    if tag:
        comp = oracle.lz4_compress(payload)
        body = struct.pack("<II", len(payload), len(comp)) + comp  # compressblockstore header, :127-131
    else:
        body = bytes(payload)
    head = struct.pack("<QIII", block_hash, ol.HASH_BLAKE3, len(hs), tag)
    return block_hash, head + hs.tobytes() + np.asarray(sizes, dtype="<u4").tobytes() + body


def _write_worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, ROOT)
    import pickle

    import torch
    import torch.distributed as dist

    import oracle_lib as ol
    import dist_model as ltd
    from synth import chunker_params, synth_bytes

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    target = 16
    part = target * 1024
    mn, av, mx = chunker_params(target)
    oracle = ol.Oracle()
    dup = synth_bytes(60, 5 * part + 11, "rand")
    datas = [synth_bytes(50, 7 * part + 5, "rand"), dup, synth_bytes(51, 3 * part, "nib"), dup.copy(), synth_bytes(52, 9 * part + 1, "text"),
             synth_bytes(53, 0, "rand"), synth_bytes(54, 4 * part + 77, "rand")]
    tags = [0, ol.COMP_LZ4, ol.COMP_LZ4, 0, ol.COMP_LZ4, 0, 0]
    jobs = ltd.plan_jobs([d.size for d in datas], target)
    first, last = ltd.shard_jobs(jobs, world)[rank]
    counts, hashes, sizes, ctags, local_chunks = [], [], [], [], []
    for a, start, n in jobs[first:last]:
        chunk = datas[a][start:start + n]
        lens = oracle.chunk(chunk, mn, av, mx)
        counts.append(lens.size)
        offs = np.concatenate([[0], np.cumsum(lens.astype(np.uint64))[:-1]]).astype(np.uint64)
        hashes.extend(oracle.hash_segments(ol.HASH_BLAKE3, chunk, offs, lens).tolist())
        sizes.extend(lens.tolist())
        ctags.extend([tags[a]] * lens.size)
        local_chunks.extend((a, start + int(o), int(l)) for o, l in zip(offs, lens))
    t = lambda x, dt: torch.from_numpy(np.asarray(x, dtype=dt))
    jc, gh, gs, gt = ltd.allgather_tables(t(counts, np.int64), t(np.asarray(hashes, dtype=np.uint64).view(np.int64), np.int64),
                                          t(sizes, np.int32), t(ctags, np.int32))
    # chunks per rank: every rank knows the shard boundaries and the per-job counts
    shards = ltd.shard_jobs(jobs, world)
    rank_counts = [int(jc[f:l].sum()) for f, l in shards]
    plan = ltd.plan_write(gh.numpy().view(np.uint64), gs.numpy(), gt.numpy(), rank_counts, max_block_size=4 * part, max_chunks_per_block=6)
    my_start = int(plan["rank_starts"][rank])

    def read_local(u):
        a, off, n = local_chunks[int(plan["first"][u]) - my_start]
        return torch.from_numpy(datas[a][off:off + n].copy())

    foreign = ltd.exchange_chunks(plan, rank, read_local, torch.device("cpu"))
    g_hashes = gh.numpy().view(np.uint64)
    blocks = []
    for (f, c), o in zip(plan["blocks"], plan["owner"]):
        if o != rank:
            continue
        payload = b"".join((foreign[u] if plan["chunk_owner"][u] != rank else read_local(u)).numpy().tobytes() for u in range(f, f + c))
        blocks.append(_serialise_block(oracle, ol, g_hashes[plan["first"][f:f + c]], plan["sizes"][f:f + c], int(plan["tags"][f]), payload))
    with open(os.path.join(out_dir, "blocks_%d_of_%d.pkl" % (rank, world)), "wb") as fh:
        pickle.dump({"blocks": blocks, "straddlers": len(plan["fetch"]), "fetched": sum(len(v) for v in plan["fetch"].values())}, fh)
    if rank == 0:
        assets = [("f%02d.bin" % i, d) for i, d in enumerate(datas)]
        want_blocks, _ = oracle.upsync(assets, target, max_block_size=4 * part, max_chunks_per_block=6, tags=tags)
        with open(os.path.join(out_dir, "want_%d.pkl" % world), "wb") as fh:
            pickle.dump(want_blocks, fh)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_write_content_sharded_by_block(world, tmp_path):
    import pickle

    import torch.multiprocessing as mp
    port = 31500 + os.getpid() % 2000 + world
    mp.spawn(_write_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    want = pickle.load(open(tmp_path / ("want_%d.pkl" % world), "rb"))
    got, fetched = {}, 0
    for r in range(world):
        d = pickle.load(open(tmp_path / ("blocks_%d_of_%d.pkl" % (r, world)), "rb"))
        fetched += d["fetched"] if r == 0 else 0
        for h, b in d["blocks"]:
            assert h not in got, "block stored twice"
            got[h] = b
    assert len(got) == len(want) and len(want) > 8
    for h, b in want:
        assert got[h] == b, "block %016x differs" % h
    assert fetched > 0, "the test data must contain a block that straddles a rank boundary"
