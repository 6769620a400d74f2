"""The N>1 host logic (longtail_b200/distributed.py) on CPU: world_size 2, gloo.  Each rank produces the chunk table of its
slice of the job list with the CPU oracle (standing in for the GPU), the tables are merged with allgather_tables, and the
merged table must equal the single-process table in global job order."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_job_plan_matches_reference_part_rule():
    from longtail_b200.distributed import plan_jobs, shard_jobs
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
    from longtail_b200 import distributed as ltd
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
