"""TEST INFRASTRUCTURE — a Python model of the sharded indexing path used by the gloo (CPU, world_size 2 / 3) tests; the product's multi-GPU path
is C + NCCL behind the C ABI (lt_b200_index_sharded / lt_b200_write_blocks_sharded, longtail_b200/csrc/capi.cu).

Host-side sharding of the indexing path across GPUs (one process per GPU, torch.distributed for the plumbing).

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


def job_assets(jobs):
    """the asset index of every job as an array — computed once per job list, so that the per-step merge below is vectorised"""
    return np.fromiter((j[0] for j in jobs), dtype=np.int64, count=len(jobs))


def asset_chunk_counts(jobs, job_chunk_counts, asset_count):
    """sum the per-job chunk counts per asset (src/longtail.c:2499-2517); `jobs` is the job list or job_assets(jobs)"""
    assets = jobs if isinstance(jobs, np.ndarray) else job_assets(jobs)
    counts = np.asarray(job_chunk_counts, dtype=np.int64)[:assets.size]
    return np.bincount(assets, weights=counts, minlength=int(asset_count)).astype(np.uint32)


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


# ------------------------------------------------------------------------------------------------------------------
# WriteContent across GPUs: stored blocks are independent, so they shard by block.  Block composition
# needs the merged unique-chunk order, which every rank derives from the allgathered chunk table (no extra collective); a
# block goes to the rank that holds the first occurrence of its first chunk.  Ranks own contiguous ordinal ranges and unique
# chunks are ordered by first occurrence, so a block's chunks live on a contiguous run of ranks: at most world-1 blocks
# straddle a rank boundary, and only their foreign chunks (a few MiB) cross NVLink, point to point.


def first_occurrences(hashes):
    """global ordinals of the first occurrence of every distinct chunk hash, ascending = the VersionIndex's unique-chunk order
    (first-occurrence dedup, src/longtail.c:2952-2970)"""
    _, idx = np.unique(np.asarray(hashes, dtype=np.uint64), return_index=True)
    return np.sort(idx).astype(np.int64)


def pack_blocks(sizes, tags, max_block_size=8388608, max_chunks_per_block=1024):
    """Longtail_CreateStoreIndex's greedy packing over the unique chunks in order (src/longtail.c:6796-6860): a block closes on a tag
    change, at max_chunks_per_block chunks, or when the next chunk would exceed max_block_size + max_block_size/10.
    -> list of (first_unique_chunk, chunk_count).  Runs in the C library (lt_b200_pack_blocks, a host-only helper)."""
    import ctypes as C

    from longtail_b200 import load_library
    lib = load_library()
    sz = np.ascontiguousarray(sizes, dtype=np.uint32)
    tg = np.ascontiguousarray(tags, dtype=np.uint32)
    first = np.zeros(max(sz.size, 1), dtype=np.uint32)
    count = np.zeros(max(sz.size, 1), dtype=np.uint32)
    nb = C.c_uint32(0)
    err = lib.lt_b200_pack_blocks(C.c_uint32(sz.size), sz.ctypes.data_as(C.c_void_p), tg.ctypes.data_as(C.c_void_p), C.c_uint32(int(max_block_size)),
                                  C.c_uint32(int(max_chunks_per_block)), first.ctypes.data_as(C.c_void_p), count.ctypes.data_as(C.c_void_p), C.byref(nb))
    if err:
        raise RuntimeError("lt_b200_pack_blocks failed with %d" % err)
    return list(zip(first[:nb.value].tolist(), count[:nb.value].tolist()))


def plan_write(hashes, sizes, tags, rank_chunk_counts, max_block_size=8388608, max_chunks_per_block=1024):
    """Every rank calls this with the same allgathered table and gets the same plan.

    -> dict(first=ordinals of the unique chunks, blocks=[(first_unique, count)], owner=rank per block,
            chunk_owner=rank holding each unique chunk's first occurrence,
            fetch={(needer, holder): [unique chunk indices]} for the chunks of straddling blocks)"""
    first = first_occurrences(hashes)
    sizes_u = np.asarray(sizes)[first]
    tags_u = np.asarray(tags)[first]
    blocks = pack_blocks(sizes_u, tags_u, max_block_size, max_chunks_per_block)
    starts = np.concatenate([[0], np.cumsum(np.asarray(rank_chunk_counts, dtype=np.int64))])
    chunk_owner = (np.searchsorted(starts, first, side="right") - 1).astype(np.int64)
    bf = np.array([b[0] for b in blocks], dtype=np.int64)
    bl = np.array([b[0] + b[1] - 1 for b in blocks], dtype=np.int64)
    owner = chunk_owner[bf] if blocks else np.zeros(0, np.int64)
    fetch = {}
    # owners are non-decreasing along the unique-chunk order, so a block straddles iff its last chunk lives elsewhere
    for b in np.nonzero(chunk_owner[bl] != owner)[0] if blocks else []:
        f, c = blocks[int(b)]
        o = int(owner[b])
        for u in range(f, f + c):
            if chunk_owner[u] != o:
                fetch.setdefault((o, int(chunk_owner[u])), []).append(u)
    return {"first": first, "sizes": sizes_u, "tags": tags_u, "blocks": blocks, "owner": owner, "chunk_owner": chunk_owner,
            "fetch": fetch, "rank_starts": starts}


def exchange_chunks(plan, rank, read_local, device, group=None):
    """Move the foreign chunks of straddling blocks to the ranks that need them (torch.distributed point to point: NCCL over
    NVLink for device tensors, gloo on CPU).  read_local(unique_chunk_index) -> 1-D uint8 torch tensor on `device` holding that
    chunk's bytes (only called for chunks this rank owns).  -> {unique chunk index: uint8 tensor} for the chunks this rank needs."""
    import torch
    import torch.distributed as dist

    ops, recv_bufs = [], []
    for (needer, holder), chunks in sorted(plan["fetch"].items()):
        total = int(sum(int(plan["sizes"][u]) for u in chunks))
        if holder == rank:
            buf = torch.cat([read_local(u) for u in chunks]) if chunks else torch.zeros(0, dtype=torch.uint8, device=device)
            ops.append(dist.P2POp(dist.isend, buf.contiguous(), needer, group=group))
        elif needer == rank:
            buf = torch.empty(total, dtype=torch.uint8, device=device)
            recv_bufs.append((chunks, buf))
            ops.append(dist.P2POp(dist.irecv, buf, holder, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    got = {}
    for chunks, buf in recv_bufs:
        off = 0
        for u in chunks:
            n = int(plan["sizes"][u])
            got[u] = buf[off:off + n]
            off += n
    return got
