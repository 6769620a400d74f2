"""Host-side sharding of the indexing path across GPUs (one process per GPU, torch.distributed for the plumbing).

The unit of work is the reference's own job: one (asset, part) pair, part = target_chunk_size * 1024 bytes
(src/longtail.c:2396-2457).  Parts are independent chunker instances, so the global job list is cut into `world`
contiguous slices balanced by bytes; every rank chunks + hashes its slice on its GPU with no data-path collective, and a
single allgather of the per-rank chunk tables (u64 hash, u32 size, u32 tag per chunk) — NCCL over NVLink when the tensors
live in HBM, gloo in the CPU tests — rebuilds the table in global job order, because the slices are contiguous.
Content hashes, first-occurrence dedup and the VersionIndex layout then run once on the merged table
(lt_b200_build_version_index[_device]).
"""
import numpy as np


def plan_jobs(asset_sizes, target_chunk_size):
    """-> list of (asset_index, start, size) in the reference's job order; empty parts are dropped (they produce no chunks)"""
    part = int(target_chunk_size) * 1024
    jobs = []
    for a, size in enumerate(asset_sizes):
        size = int(size)
        for p in range(1 + size // part):
            start = p * part
            n = min(part, size - start)
            if n > 0:
                jobs.append((a, start, n))
    return jobs


def shard_jobs(jobs, world):
    """-> list of `world` (first, last) index pairs: contiguous slices of `jobs` with near-equal byte counts"""
    total = sum(j[2] for j in jobs)
    bounds = [0]
    acc = 0
    k = 1
    for i, j in enumerate(jobs):
        acc += j[2]
        while k < world and acc >= total * k / world:
            bounds.append(i + 1)
            k += 1
    while len(bounds) < world:
        bounds.append(len(jobs))
    bounds.append(len(jobs))
    return [(bounds[r], max(bounds[r], bounds[r + 1])) for r in range(world)]


def asset_chunk_counts(jobs, job_chunk_counts, asset_count):
    """sum the per-job chunk counts per asset (src/longtail.c:2499-2517)"""
    out = np.zeros(asset_count, dtype=np.uint32)
    for (a, _, _), n in zip(jobs, job_chunk_counts):
        out[a] += int(n)
    return out


def allgather_tables(local_job_counts, hashes, sizes, tags, group=None):
    """Merge per-rank chunk tables into global job order on every rank.

    local_job_counts: int64 torch tensor [local jobs] (chunks per job); hashes/sizes/tags: torch tensors
    (int64 view of the u64 hashes, int32 views of sizes and tags) on the same device, length = local chunk count.
    Returns (job_counts [all jobs], hashes, sizes, tags) as torch tensors on that device.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = hashes.device
    # 1. how many jobs / chunks each rank holds
    meta = torch.tensor([local_job_counts.numel(), hashes.numel()], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    njobs = [int(m[0]) for m in metas]
    nchunks = [int(m[1]) for m in metas]
    max_jobs, max_chunks = max(njobs + [1]), max(nchunks + [1])

    def gather(t, n_max, lens, dtype):
        pad = torch.zeros(n_max, dtype=dtype, device=dev)
        pad[:t.numel()] = t
        outs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(outs, pad, group=group)
        return torch.cat([o[:n] for o, n in zip(outs, lens)])

    return (gather(local_job_counts.to(torch.int64), max_jobs, njobs, torch.int64), gather(hashes, max_chunks, nchunks, torch.int64),
            gather(sizes, max_chunks, nchunks, torch.int32), gather(tags, max_chunks, nchunks, torch.int32))


class DeviceArray:
    """zero-copy view of device memory for torch.as_tensor (CUDA array interface v2)"""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2, "strides": None}
