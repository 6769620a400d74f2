#!/usr/bin/env python
"""CreateVersionIndex + WriteContent sharded over the GPUs of one box, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools_multi_gpu_write.py --gib 2 --codec lz4 [--verify]

Every rank holds one synthetic asset (configs[2] data: random / 4-bit / text-like segments, half of them drawn from a pool that
all ranks share, so chunks deduplicate ACROSS ranks) in its own HBM arena.  Steps: chunk + hash the own slice
(lt_b200_chunk_ranges) -> allgather of the chunk tables over NCCL -> every rank derives the same plan (first-occurrence dedup,
greedy block packing, block -> rank; longtail_b200/distributed.py) -> the few foreign chunks of straddling blocks move point to
point over NVLink -> every rank gathers, compresses and serialises the blocks it owns (lt_b200_write_blocks_device).
--verify: rank 0 runs the unmodified reference's upsync over all assets on the CPU and compares every StoredBlock byte for byte.
"""
import argparse
import json
import os
import pickle
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
GIB = 1 << 30
TARGET = 65536


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=2.0, help="asset bytes per GPU")
    ap.add_argument("--codec", default="lz4", choices=["lz4", "zstd", "none"])
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--scratch", default=None, help="directory shared by the ranks for --verify (default: a temp dir under /dev/shm)")
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist

    import longtail_b200
    from longtail_b200 import distributed as ltd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = longtail_b200.Context(local_rank)
    tag = {"lz4": longtail_b200.COMPRESSION_LZ4, "zstd": longtail_b200.COMPRESSION_ZSTD_DEFAULT, "none": 0}[args.codec]

    nbytes = int(args.gib * GIB) // 256 * 256 + (rank * 4096 + 77)  # ragged, different per rank
    slack = 64 << 20                                                # room for the foreign chunks of straddling blocks
    arena_bytes = ((nbytes + 255) & ~255) + slack + 4096
    arena = ctx.device_alloc(arena_bytes)
    ctx.synth_fill(arena, nbytes, seed=2, asset_id=rank, class_mode=1, shared_permille=500, pool_segments=max(8, int(args.gib * 256)))
    ctx.synchronize()
    foreign_base = (nbytes + 255) & ~255

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sizes_all = [int(args.gib * GIB) // 256 * 256 + (r * 4096 + 77) for r in range(world)]
    jobs = ltd.plan_jobs(sizes_all, TARGET)
    my_jobs = [j for j in jobs if j[0] == rank]
    mn, av, mx = longtail_b200.chunker_params(TARGET)
    times = {}
    barrier()
    t0 = time.perf_counter()
    t = ctx.chunk_ranges(arena, arena_bytes, [(start, size, tag) for _, start, size in my_jobs], mn, av, mx)
    barrier()
    times["chunk_hash_s"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    if world > 1:
        jc, gh, gs, gt = ltd.allgather_tables(torch.as_tensor(t["range_chunk_counts"].astype(np.int64), device=dev),
                                              torch.as_tensor(t["hashes"].view(np.int64), device=dev),
                                              torch.as_tensor(t["sizes"].view(np.int32), device=dev), torch.as_tensor(t["tags"].view(np.int32), device=dev))
        jc, gh, gs, gt = jc.cpu().numpy(), gh.cpu().numpy().view(np.uint64), gs.cpu().numpy().view(np.uint32), gt.cpu().numpy().view(np.uint32)
    else:
        jc, gh, gs, gt = t["range_chunk_counts"].astype(np.int64), t["hashes"], t["sizes"], t["tags"]
    barrier()
    times["allgather_s"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    per_rank = np.zeros(world, dtype=np.int64)
    for (a, _, _), n in zip(jobs, jc):
        per_rank[a] += int(n)  # asset a lives on rank a
    plan = ltd.plan_write(gh, gs, gt, per_rank.tolist())
    my_start = int(plan["rank_starts"][rank])
    times["plan_s"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    arena_t = torch.as_tensor(ltd.DeviceArray(arena, arena_bytes, "|u1"), device=dev)

    def read_local(u):
        k = int(plan["first"][u]) - my_start
        off, n = int(t["offsets"][k]), int(t["sizes"][k])
        return arena_t[off:off + n].clone()

    foreign = ltd.exchange_chunks(plan, rank, read_local, dev) if world > 1 else {}
    foreign_off, cursor = {}, foreign_base
    for u, buf in foreign.items():
        if cursor + buf.numel() > foreign_base + slack:
            raise SystemExit("foreign-chunk slack exhausted")
        arena_t[cursor:cursor + buf.numel()].copy_(buf)
        foreign_off[u] = cursor
        cursor = (cursor + buf.numel() + 15) & ~15
    barrier()
    times["exchange_s"] = time.perf_counter() - t0

    # my blocks, whole: the greedy packer inside lt_b200_write_blocks_device re-derives the same boundaries from this list
    mine = [u for (f, c), o in zip(plan["blocks"], plan["owner"]) if o == rank for u in range(f, f + c)]
    mine = np.asarray(mine, dtype=np.int64)
    offs = np.array([foreign_off[u] if plan["chunk_owner"][u] != rank else int(t["offsets"][int(plan["first"][u]) - my_start]) for u in mine],
                    dtype=np.uint64)
    t0 = time.perf_counter()
    blocks = ctx.write_blocks_device(arena, arena_bytes, gh[plan["first"][mine]], plan["sizes"][mine], plan["tags"][mine], offs,
                                     keep_bytes=args.verify) if mine.size else []
    barrier()
    times["write_s"] = time.perf_counter() - t0

    stats = {"rank": rank, "blocks": len(blocks), "stored": sum(len(b) if args.verify else b for _, b in blocks), "foreign_chunks": len(foreign),
             "foreign_bytes": int(sum(b.numel() for b in foreign.values()))}
    ok = True
    if args.verify:
        scratch = args.scratch or os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(), "lt_b200_mgw_%s" % os.environ.get("MASTER_PORT", "0"))
        os.makedirs(scratch, exist_ok=True)
        with open(os.path.join(scratch, "r%d.pkl" % rank), "wb") as fh:
            pickle.dump({"asset": ctx.to_host(arena, nbytes), "blocks": blocks}, fh, protocol=4)
        barrier()
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib as ol
            ref = ol.Reference()
            checker = ref if ref.available else ol.Oracle()
            parts = [pickle.load(open(os.path.join(scratch, "r%d.pkl" % r), "rb")) for r in range(world)]
            assets = [("f%05d.bin" % r, p["asset"]) for r, p in enumerate(parts)]
            want_blocks, _ = checker.upsync(assets, TARGET, tags=[tag] * world)
            got = {}
            for p in parts:
                for h, b in p["blocks"]:
                    ok = ok and h not in got
                    got[h] = b
            ok = ok and len(got) == len(want_blocks) and all(got.get(h) == b for h, b in want_blocks)
            print("VERIFY %s: %d StoredBlocks from %d GPUs %s the %s's upsync" % ("OK" if ok else "FAILED", len(got), world,
                  "identical to" if ok else "DIFFER from", "reference" if ref.available else "oracle"), file=sys.stderr)
            for r in range(world):
                os.remove(os.path.join(scratch, "r%d.pkl" % r))
        barrier()
    all_stats = [None] * world
    if world > 1:
        dist.all_gather_object(all_stats, stats)
    else:
        all_stats = [stats]
    if rank == 0:
        total = sum(sizes_all)
        wall = sum(times.values())
        print(json.dumps({"n_gpus": world, "codec": args.codec, "bytes": total, "unique_chunks": int(plan["first"].size), "blocks": len(plan["blocks"]),
                          "straddling_pairs": len(plan["fetch"]), "seconds": {k: round(v, 4) for k, v in times.items()},
                          "GiBps_wall": round(total / wall / GIB, 3), "per_rank": all_stats, "verify": ("ok" if ok else "FAILED") if args.verify else None}))
    ctx.device_free(arena)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
