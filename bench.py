#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric: GiB/s of the chunk -> hash -> compress indexing path (CreateVersionIndex +
CreateMissingContent + WriteContent through the compress block store) and the fraction of the HBM roofline.

Workloads (config.workload names the one `value` is measured on):
  headline, every N   BASELINE.json configs[2] per GPU: 10 000 synthetic assets totalling 128 GiB with 50 % byte redundancy
                      (sizes log-uniform in [64 KiB, 256 MiB], 1 MiB segments half of which come from a
                      shared 8 GiB pool, random / 4-bit / text-like segments), target chunk 65 536, BLAKE3, tag 'lz42', blocks of
                      8 MiB / 1 024 chunks.  With N > 1 every rank holds its own 10 000 assets of ONE version (the pool is shared, so
                      chunks repeat across ranks): the job list is sharded by (asset, part), the chunk tables are merged with NCCL, the
                      dedup is global, the blocks are packed over the global unique-chunk list and written by the ranks — configs[4]'s
                      full upsync (CreateVersionIndex + CreateMissingContent + WriteContent, NCCL all-gather of the chunk-hash table) at
                      N x 128 GiB.  Weak scaling.
  index_only, N = 1   configs[1]: chunk + BLAKE3 only on one 64 GiB uniform-random file (the chunker / hash kernels' roofline numbers)
  zstd_pak, every N   configs[3]'s shape per GPU: a 32 GiB slice of ONE PAK-like file (64 KiB - 64 MiB members, 2 KiB zero padding), parts
                      sharded across the ranks, tag 'ztd2' (ZStd level 3) — 8 x 32 GiB = the 256 GiB file of configs[3]

A step = one full pass over the workload.
  value  inputs resident in HBM, stored blocks left in HBM (device sink): CUDA events on the context stream, max over ranks
  e2e    the same work with the asset bytes in pinned HOST memory and every stored block copied to pinned host memory:
         lt_b200_upsync_stream_host_assets (one streaming pass; --e2e-mode resident = lt_b200_upsync_host_assets), wall clock
  roofline / cpu_baseline  as the task contract describes; see DESIGN.md §Measurement

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libref_shim.so: Longtail_CreateVersionIndex + CreateMissingContent +
WriteContent through its compressblockstore with the bikeshed JobAPI on all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GIB = 1 << 30
TARGET_CHUNK_SIZE = 65536
MAX_BLOCK_SIZE = 8388608
MAX_CHUNKS_PER_BLOCK = 1024
SEED_CONFIG2 = 2
SEED_RANDOM = 1
SEED_PAK = 3
METRIC = "GiB/s end-to-end chunk+hash+compress (CreateVersionIndex); % HBM roofline"
try:
    with open(os.path.join(ROOT, "BASELINE.json")) as _f:
        METRIC = json.load(_f).get("metric", METRIC)
except (OSError, ValueError):
    pass


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gib", type=float, default=128.0, help="size of the configs[2] asset set per GPU (GiB); 128 = the full config")
    ap.add_argument("--index-gib", type=float, default=64.0, help="size of the configs[1] file of the index_only leg (0 = skip)")
    ap.add_argument("--zstd-gib", type=float, default=32.0, help="size of the PAK-like slice per GPU of the zstd_pak leg (0 = skip)")
    ap.add_argument("--e2e-gib", type=float, default=None, help="host-resident part of the asset set for the e2e leg (default: all that fits host RAM)")
    ap.add_argument("--cpu-gib", type=float, default=16.0, help="bounded sample of the reference arm per step (BASELINE.md: a fixed 16 GiB prefix)")
    ap.add_argument("--verify-gib", type=float, default=1.0, help="N > 1: per-rank size of the multi-GPU parity run against the reference (0 = skip)")
    ap.add_argument("--e2e-mode", default="stream", choices=["stream", "resident"], help="e2e verb: one streaming pass (default) or H2D of everything first")
    ap.add_argument("--e2e-batch-gib", type=float, default=16.0, help="batch size of the streaming e2e pass")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the reference runs (parity and cpu_baseline)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def config3_asset_sizes(total_bytes, count):
    """sizes log-uniform in [64 KiB, 256 MiB], rescaled to the total, 256-byte granules"""
    x, sizes = 12345, []
    for _ in range(count):
        x = (x * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
        u = (x >> 11) / float(1 << 53)
        sizes.append(math.exp(math.log(65536.0) + u * (math.log(268435456.0) - math.log(65536.0))))
    scale = total_bytes / sum(sizes)
    return [max(4096, int(v * scale) // 256 * 256) for v in sizes]


class Config2Set:
    """the configs[2] asset set of one rank: names, sizes, arena offsets, generator parameters"""

    def __init__(self, gib, rank, world):
        self.total_target = int(gib * GIB)
        self.count = max(8, int(round(10000 * gib / 128.0)))
        self.sizes = config3_asset_sizes(self.total_target, self.count)
        self.offsets, off = [], 0
        for sz in self.sizes:
            self.offsets.append(off)
            off += (sz + 255) & ~255
        self.arena_bytes = off + 4096
        self.nbytes = sum(self.sizes)
        self.pool = max(8, int(8192 * gib / 128.0))  # the shared pool: 8 GiB at full size (1 MiB segments)
        self.first_id = rank * self.count
        self.names = ["a/%06d.bin" % (self.first_id + i) for i in range(self.count)]
        self.all_names = ["a/%06d.bin" % i for i in range(self.count * world)]
        self.all_sizes = self.sizes * world

    def fill(self, ctx, arena):
        for i, (o, sz) in enumerate(zip(self.offsets, self.sizes)):
            ctx.synth_fill(arena + o, sz, seed=SEED_CONFIG2, asset_id=self.first_id + i, class_mode=1, shared_permille=500, pool_segments=self.pool)
        ctx.synchronize()


def synth_host_lib():
    import ctypes as C
    path = os.path.join(ROOT, "oracle", "_ref", "libsynth_host.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    return C.CDLL(path)


def host_assets(cs, limit_bytes, threads, first_id=0):
    """the first assets of a Config2Set (up to limit_bytes), generated on the HOST (oracle/_ref/libsynth_host.so) -> list of uint8 arrays"""
    import ctypes as C

    import numpy as np

    import longtail_b200
    lib = synth_host_lib()
    out, acc = [], 0
    spec = longtail_b200.SynthSpec(SEED_CONFIG2, 500, cs.pool, 1, 0)
    for i, sz in enumerate(cs.sizes):
        if acc + sz > limit_bytes and out:
            break
        buf = np.empty(sz, dtype=np.uint8)
        lib.synth_fill_mt(C.byref(spec), C.c_uint64(first_id + i), C.c_uint64(0), buf.ctypes.data_as(C.c_void_p), C.c_uint64(sz), C.c_uint32(threads))
        out.append(buf)
        acc += sz
    return out


# ---------------------------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    cs = Config2Set(args.gib, 0, 1)
    ref = ol.Reference()
    workload = "configs[2]: chunk+BLAKE3+LZ4 (CreateVersionIndex + CreateMissingContent + WriteContent), %d synthetic assets / %.0f GiB, 50 %% byte " \
               "redundancy, target_chunk_size 65536" % (cs.count, args.gib)
    if not ref.available:
        # the reference was not compiled on this box: the oracle's single-threaded C restatement of the same path stands in (kind "port")
        o = ol.Oracle()
        datas = host_assets(cs, int(min(args.cpu_gib, 0.5) * GIB), 4)
        assets = list(zip(cs.names, datas))
        nbytes = sum(d.size for d in datas)
        t = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            o.upsync(assets, TARGET_CHUNK_SIZE, tags=[ol.COMP_LZ4] * len(assets))
            if i >= args.warmup:
                t.append(time.perf_counter() - t0)
        total = sum(t)
        value = nbytes * args.steps / total / GIB
        sample = "first %d assets (%.2f GiB), oracle/lt_oracle.c (C restatement), 1 thread" % (len(assets), nbytes / GIB)
        cores, kind = 1, "port"
    else:
        cores = ref.cpu_count()
        datas = host_assets(cs, int(args.cpu_gib * GIB), cores)
        assets = list(zip(cs.names, datas))
        nbytes = sum(d.size for d in datas)
        tags = [ol.COMP_LZ4] * len(assets)

        def one_pass():
            secs, _stored = ref.upsync(assets, TARGET_CHUNK_SIZE, MAX_BLOCK_SIZE, MAX_CHUNKS_PER_BLOCK, tags=tags, workers=cores, keep_bytes=False)
            return sum(secs)

        for _ in range(args.warmup):
            one_pass()
        t = [one_pass() for _ in range(args.steps)]
        total = sum(t)
        value = nbytes * args.steps / total / GIB
        sample = "first %d assets (%.2f GiB) of the set: CreateVersionIndex + CreateMissingContent + WriteContent through compressblockstore " \
                 "(LZ4), bikeshed %d workers + caller" % (len(assets), nbytes / GIB, cores)
        kind = "reference"
    emit({
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": "GiB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1e3 * total / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload + " (CPU sample: %s)" % sample},
        "cpu_baseline": {"value": round(value, 4), "unit": "GiB/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 4), "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ---------------------------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import longtail_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = longtail_b200.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    comm = None
    if world > 1:
        # the library's own NCCL communicator: rank 0 makes the id, torch.distributed only carries its 128 bytes
        box = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = ctx.comm_create(box[0], rank, world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    LZ4 = longtail_b200.COMPRESSION_LZ4
    ZSTD = longtail_b200.COMPRESSION_ZSTD_DEFAULT
    part = TARGET_CHUNK_SIZE * 1024

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def timed(fn, steps):
        """-> device ms over `steps` calls of fn (CUDA events on the context stream), max over ranks"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ------------------------------------------------------------------ the sharded upsync of one version (any N)
    class Upsync:
        """CreateVersionIndex + CreateMissingContent (fresh store) + WriteContent of the version made of every rank's assets"""

        def __init__(self, names, sizes, tags, my_first_asset, my_offsets, arena, arena_bytes, byte_range_base=None):
            self.assets = longtail_b200.AssetList(names, sizes)
            self.tags = tags
            self.arena, self.arena_bytes = arena, arena_bytes
            self.my_first_asset, self.my_offsets = my_first_asset, my_offsets
            if comm is not None:
                first, njobs = ctx.plan_shards(self.assets, TARGET_CHUNK_SIZE, world)
                jobs = ctx.shard_jobs(self.assets, TARGET_CHUNK_SIZE, int(first[rank]), int(first[rank + 1] - first[rank]))
                if byte_range_base is not None:
                    # ONE asset sharded by byte range: this rank's arena holds its bytes [byte_range_base, byte_range_base + arena size)
                    if jobs.size and (int(jobs["offset"].min()) < byte_range_base or
                                      int((jobs["offset"] + jobs["size"]).max()) > byte_range_base + arena_bytes):
                        raise SystemExit("shard plan does not follow the byte ranges the ranks hold")
                    self.job_offsets = (jobs["offset"].astype(np.int64) - byte_range_base).astype(np.uint64)
                else:
                    local = jobs["asset_index"].astype(np.int64) - my_first_asset
                    if local.size and (local.min() < 0 or local.max() >= len(my_offsets)):
                        raise SystemExit("shard plan hands rank %d a job of an asset it does not hold" % rank)
                    self.job_offsets = np.asarray(my_offsets, dtype=np.uint64)[local] + jobs["offset"]
            self.index_bytes = 0

        def index(self, want_host=True, copy=False):
            if comm is None:
                v = ctx.index_device_assets(self.arena, self.arena_bytes, self.assets, self.my_offsets, self.tags, target_chunk_size=TARGET_CHUNK_SIZE, copy=copy)
                self.index_bytes = len(v)
                return v
            v = ctx.index_sharded(comm, self.arena, self.arena_bytes, self.assets, self.tags, self.job_offsets, TARGET_CHUNK_SIZE,
                                  want_host=want_host, copy=copy)
            self.index_bytes = v if isinstance(v, int) else len(v)
            return v

        def write(self, c_sink, device_sink, vi=None):
            if comm is None:
                if vi is None:
                    raise ValueError("single-GPU write needs the parsed VersionIndex")
                ctx.write_blocks_device(self.arena, self.arena_bytes, vi["chunk_hashes"], vi["chunk_sizes"], vi["chunk_tags"],
                                        ctx.unique_chunk_offsets(vi["chunk_count"]), MAX_BLOCK_SIZE, MAX_CHUNKS_PER_BLOCK, c_sink=c_sink,
                                        device_sink=device_sink)
                return None
            return ctx.write_blocks_sharded(comm, c_sink, MAX_BLOCK_SIZE, MAX_CHUNKS_PER_BLOCK, device_sink=device_sink)

    def parity_against_reference(up, host_datas_of_all_ranks, names, tags, label):
        """rank 0 runs the unmodified reference over the whole version; the VersionIndex must be byte-identical and every StoredBlock
        (block hash, size, 64-bit digest of its serialised bytes, in store order) identical — the blocks of all ranks concatenated in
        rank order.  -> (text for the JSON line, reference seconds or None)"""
        import oracle_lib as ol
        ref = ol.Reference()
        if not ref.available:
            return "unpinned: oracle/_ref/libref_shim.so is not built on this box", None
        sink = ol.DigestSink(ref)
        v = up.index(want_host=True, copy=True)
        vi = longtail_b200.parse_version_index(v) if comm is None or rank == 0 else None
        up.write((sink.fn, sink.user), False, vi)
        mine = sink.records()
        sink.close()
        if world > 1:
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(mine, gathered, dst=0)
        else:
            gathered = [mine]
        secs = None
        if rank == 0:
            got = np.concatenate(gathered, axis=0)
            rec, want_v, secs = ref.upsync(list(zip(names, host_datas_of_all_ranks)), TARGET_CHUNK_SIZE, MAX_BLOCK_SIZE, MAX_CHUNKS_PER_BLOCK,
                                           tags=tags, workers=ref.cpu_count(), keep_bytes=2)
            if bytes(v) != want_v:
                raise SystemExit("PARITY FAILED (%s): the VersionIndex differs from the reference's" % label)
            if got.shape != rec.shape or not np.array_equal(got, rec):
                raise SystemExit("PARITY FAILED (%s): StoredBlocks differ from the reference's (%d vs %d blocks)" % (label, got.shape[0], rec.shape[0]))
            text = "full: %d-byte VersionIndex memcmp-identical and all %d StoredBlocks (hash, size, 64-bit digest of the serialised bytes, store " \
                   "order) identical to the unmodified reference on %s" % (len(want_v), rec.shape[0], label)
        else:
            text = ""
        return text, secs

    # ================================================================== headline: configs[2] per GPU
    cs = Config2Set(args.gib, rank, world)
    arena = ctx.device_alloc(cs.arena_bytes)
    cs.fill(ctx, arena)
    tags_all = [LZ4] * (cs.count * world)
    up = Upsync(cs.all_names, cs.all_sizes, tags_all, cs.first_id, cs.offsets, arena, cs.arena_bytes)

    # ---- host copy of this rank's assets in pinned memory: the e2e source and (N = 1) the reference's input
    import psutil
    avail = psutil.virtual_memory().available
    want_host = cs.nbytes if args.e2e_gib is None else int(args.e2e_gib * GIB)
    host_limit = int(min(want_host, 0.72 * avail / max(world, 1)))
    host_buf, host_datas, host_count = None, [], 0
    if not (args.no_e2e and args.no_cpu):
        acc = 0
        while host_count < cs.count and acc + cs.sizes[host_count] <= host_limit:
            acc += cs.sizes[host_count]
            host_count += 1
        host_count = max(host_count, 1)
        span = cs.offsets[host_count - 1] + cs.sizes[host_count - 1]
        host_buf = ctx.pinned_alloc(span)
        ctx.lib.lt_b200_copy_to_host(ctx.handle, host_buf.ctypes.data, arena, span)
        host_datas = [host_buf[o:o + s] for o, s in zip(cs.offsets[:host_count], cs.sizes[:host_count])]
    host_bytes = sum(d.size for d in host_datas)

    # ---- parity before anything is timed
    parity, cpu = "skipped (--no-cpu)", None
    if not args.no_cpu:
        if world == 1 and host_count == cs.count:
            parity, secs = parity_against_reference(up, host_datas, cs.names, [LZ4] * cs.count, "the whole %.0f GiB asset set" % (cs.nbytes / GIB))
            if secs is not None:
                import oracle_lib as ol
                cores = ol.Reference().cpu_count()
                cpu = {"value": round(cs.nbytes / sum(secs) / GIB, 4), "unit": "GiB/s", "cores": cores, "kind": "reference",
                       "sample": "the whole set (%d assets, %.1f GiB), one pass: CreateVersionIndex + CreateMissingContent + WriteContent through "
                                 "compressblockstore (LZ4), bikeshed %d workers + caller" % (cs.count, cs.nbytes / GIB, cores),
                       "seconds_index_missing_write": [round(x, 3) for x in secs]}
        elif world == 1:
            # host RAM does not hold the whole set: the reference runs over the first host_count assets, the B200 path over the same ones
            sub = Upsync(cs.names[:host_count], cs.sizes[:host_count], [LZ4] * host_count, 0, cs.offsets[:host_count], arena, cs.arena_bytes)
            parity, secs = parity_against_reference(sub, host_datas, cs.names[:host_count], [LZ4] * host_count,
                                                    "the first %d assets (%.1f GiB; host RAM bounds the reference's input)" % (host_count, host_bytes / GIB))
            if secs is not None:
                import oracle_lib as ol
                cores = ol.Reference().cpu_count()
                cpu = {"value": round(host_bytes / sum(secs) / GIB, 4), "unit": "GiB/s", "cores": cores, "kind": "reference",
                       "sample": "first %d assets (%.1f GiB), one pass of the reference upsync (LZ4), bikeshed %d workers + caller" % (host_count, host_bytes / GIB, cores)}
        elif args.verify_gib > 0:
            # N > 1: the whole multi-GPU path (sharded index, NCCL merge, split dedup, sharded WriteContent with chunk exchange) on a small
            # version first, against the reference over all ranks' bytes on rank 0
            vs = Config2Set(args.verify_gib, rank, world)
            varena = ctx.device_alloc(vs.arena_bytes)
            vs.fill(ctx, varena)
            vup = Upsync(vs.all_names, vs.all_sizes, [LZ4] * (vs.count * world), vs.first_id, vs.offsets, varena, vs.arena_bytes)
            # rank 0 makes every rank's assets again on the host (the generator is the same function on both sides, include/lt_synth.h)
            everything = None
            if rank == 0:
                everything = []
                for r in range(world):
                    everything += host_assets(Config2Set(args.verify_gib, r, world), 1 << 62, os.cpu_count() or 8, first_id=r * vs.count)
            parity, _ = parity_against_reference(vup, everything, vs.all_names, [LZ4] * (vs.count * world),
                                                 "a %d x %.1f GiB version sharded over %d GPUs (the full-size run is checked by its invariants)" % (world, args.verify_gib, world))
            ctx.device_free(varena)

    # ---- resident-in-HBM measurement: index + write, blocks left in HBM
    count_fn, count_user, acc = ctx.counting_sink()
    state = {"vi": None}

    def step_resident():
        v = up.index(want_host=(rank == 0), copy=False)
        vi = longtail_b200.parse_version_index(v) if rank == 0 else None
        up.write((count_fn, count_user), True, vi)
        state["vi"] = vi
        state["index_bytes"] = len(v) if rank == 0 else 0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    one_step = None
    for _ in range(max(args.warmup, 3)):
        step_resident()
        if one_step is None:
            one_step = acc.copy()  # {blocks, stored bytes, raw bytes, xor of block hashes} of ONE pass: what the e2e pass must reproduce
    acc[:] = 0
    launches0 = ctx.launch_count
    ctx.profile_reset()
    ctx.profile_enable(True)
    ms = timed(step_resident, args.steps)
    ctx.profile_enable(False)
    prof = ctx.profile_read()
    launches = ctx.launch_count - launches0
    clocks = sampler.finish() if rank == 0 else None
    blocks_mine, stored_mine, raw_mine = int(acc[0]) // args.steps, int(acc[1]) // args.steps, int(acc[2]) // args.steps
    blocks_all, stored_all, raw_all = int(sum_over_ranks(blocks_mine)), int(sum_over_ranks(stored_mine)), int(sum_over_ranks(raw_mine))
    value = world * cs.nbytes * args.steps / (ms / 1e3) / GIB
    # invariants of the full-size run: every unique byte lands in exactly one block, on every rank's count
    if rank == 0 and state["vi"] is not None:
        uniq = int(state["vi"]["chunk_sizes"].astype(np.uint64).sum())
        if raw_all != uniq:
            raise SystemExit("invariant failed: %d payload bytes in blocks, %d unique chunk bytes in the VersionIndex" % (raw_all, uniq))

    # ---- end to end from pinned host memory (this rank's assets; every stored block copied to pinned host memory)
    e2e = None
    if not args.no_e2e and host_count:
        e2e_assets = longtail_b200.AssetList(cs.names[:host_count], cs.sizes[:host_count])
        e2e_tags = [LZ4] * host_count
        efn, euser, eacc = ctx.counting_sink()
        ibytes = [0]

        streaming = args.e2e_mode == "stream"

        def step_e2e():
            if streaming:
                v, _ = ctx.upsync_stream_host_assets(e2e_assets, host_datas, e2e_tags, (efn, euser), TARGET_CHUNK_SIZE, MAX_BLOCK_SIZE, MAX_CHUNKS_PER_BLOCK,
                                                     batch_bytes=int(args.e2e_batch_gib * GIB), copy=False)
            else:
                v, _ = ctx.upsync_host_assets(e2e_assets, host_datas, e2e_tags, (efn, euser), TARGET_CHUNK_SIZE, MAX_BLOCK_SIZE, MAX_CHUNKS_PER_BLOCK, copy=False)
            ibytes[0] = len(v)

        ctx.device_free(arena)  # the verb brings its own arena
        arena = None
        ctx.trim()              # ... and the workspace of the resident run (write batches, merged tables) is not needed next to it
        step_e2e()
        same = None
        if world == 1 and host_count == cs.count:
            # the whole version went through: the pass from host memory must have written the resident pass's blocks
            same = bool((eacc == one_step).all()) and ibytes[0] == int(state["index_bytes"])
            if not same:
                raise SystemExit("e2e pass differs from the resident pass: %s vs %s, index %d vs %d bytes" % (eacc, one_step, ibytes[0], state["index_bytes"]))
        eacc[:] = 0
        e2e_steps = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        barrier()
        wall = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": round(world * host_bytes * e2e_steps / wall / GIB, 3), "unit": "GiB/s", "h2d_bytes_per_step": host_bytes,
               "d2h_bytes_per_step": int(eacc[1]) // e2e_steps + ibytes[0], "steps": e2e_steps,
               "verb": "lt_b200_upsync_stream_host_assets" if streaming else "lt_b200_upsync_host_assets",
               "same_blocks_as_resident_pass": same,
               "note": ("one streaming pass per GPU over %d assets (%.1f GiB) in pinned host memory: batches of %.0f GiB travel host -> device while the "
                        "previous batch is chunked, hashed, deduplicated, packed and compressed and its StoredBlocks are copied to pinned host staging; "
                        "wall clock, max over ranks%s" % (host_count, host_bytes / GIB, args.e2e_batch_gib,
                                                          "" if host_count == cs.count else " (host RAM bounds the host-resident part of the set)"))
               if streaming else
               ("lt_b200_upsync_host_assets per GPU over %d assets (%.1f GiB) in pinned host memory: one H2D copy per asset, CreateVersionIndex, "
                "WriteContent with every StoredBlock copied to pinned host staging; wall clock, max over ranks%s"
                % (host_count, host_bytes / GIB, "" if host_count == cs.count else " (host RAM bounds the host-resident part of the set)"))}
    if host_buf is not None:
        ctx.pinned_free(host_buf)
        host_buf, host_datas = None, []
    if arena is not None:
        ctx.device_free(arena)
        arena = None

    ctx.trim()

    # ================================================================== index_only: configs[1] (N = 1)
    index_only = None
    if world == 1 and args.index_gib > 0:
        nbytes = int(args.index_gib * GIB)
        a1 = ctx.device_alloc(nbytes + 4096)
        ctx.synth_fill(a1, nbytes, seed=SEED_RANDOM, asset_id=0)
        ctx.synchronize()
        al1 = longtail_b200.AssetList(["f00000.bin"], [nbytes])

        def step_index():
            ctx.index_device_assets(a1, nbytes + 4096, al1, [0], None, target_chunk_size=TARGET_CHUNK_SIZE, copy=False)

        for _ in range(3):
            step_index()
        ctx.profile_reset()
        ctx.profile_enable(True)
        ms1 = timed(step_index, args.steps)
        ctx.profile_enable(False)
        p1 = ctx.profile_read()
        index_only = {"workload": "configs[1]: chunk+BLAKE3 only, one %.0f GiB uniform-random file, target_chunk_size 65536" % args.index_gib,
                      "value_GiBps": round(nbytes * args.steps / (ms1 / 1e3) / GIB, 3), "ms_per_step": round(ms1 / args.steps, 3),
                      "per_kernel": {k: {"ms_per_step": round(v[0] / args.steps, 3), "GBps": round(v[2] / max(v[0], 1e-9) / 1e6, 1)}
                                     for k, v in p1.items() if v[0] > 0}}
        ctx.device_free(a1)

    # ================================================================== zstd_pak: configs[3]'s shape, a slice of one PAK-like file per GPU
    zstd_pak = None
    if args.zstd_gib > 0:
        zbytes = int(args.zstd_gib * GIB) // part * part
        za = ctx.device_alloc(zbytes + 4096)
        ctx.synth_fill(za, zbytes, seed=SEED_PAK, asset_id=0, offset=rank * zbytes, class_mode=2)
        ctx.synchronize()
        # ONE file of world x zbytes; rank r holds bytes [r * zbytes, (r + 1) * zbytes) — whole parts, so the byte-range shards need no stitch
        zup = Upsync(["pak/data.pak"], [zbytes * world], [ZSTD], 0, [0], za, zbytes + 4096, byte_range_base=rank * zbytes)
        zfn, zuser, zacc = ctx.counting_sink()

        def step_zstd():
            v = zup.index(want_host=(rank == 0), copy=False)
            zup.write((zfn, zuser), True, longtail_b200.parse_version_index(v) if rank == 0 else None)

        step_zstd()
        zacc[:] = 0
        ctx.profile_reset()
        ctx.profile_enable(True)
        zsteps = max(1, min(args.steps, 2))
        msz = timed(step_zstd, zsteps)
        ctx.profile_enable(False)
        pz = ctx.profile_read()
        zraw, zstored = sum_over_ranks(int(zacc[2]) // zsteps), sum_over_ranks(int(zacc[1]) // zsteps)
        zstd_pak = {"workload": "configs[3] shape: chunk+BLAKE3+ZStd-level-3 on one %.0f GiB PAK-like file, %.0f GiB of whole parts per GPU" % (world * zbytes / GIB, zbytes / GIB),
                    "value_GiBps": round(world * zbytes * zsteps / (msz / 1e3) / GIB, 3), "ms_per_step": round(msz / zsteps, 1),
                    "unique_bytes": int(zraw), "stored_bytes": int(zstored), "ratio": round(zstored / max(zraw, 1), 4),
                    "codec_kernel_ms_rank0": round(pz["k_zstd_frames"][0] / zsteps, 1),
                    "codec_kernel_GBps_rank0": round(pz["k_zstd_frames"][2] / max(pz["k_zstd_frames"][0], 1e-9) / 1e6, 2)}
        ctx.device_free(za)

    if rank == 0:
        peak, peak_src = peaks()
        step_ms = ms / args.steps
        per_kernel = {k: {"ms_per_step": round(v[0] / args.steps, 3), "launches_per_step": v[1] // args.steps,
                          "GBps": round(v[2] / max(v[0], 1e-9) / 1e6, 1), "share_of_step": round(v[0] / ms, 4)} for k, v in prof.items() if v[0] > 0}
        name = max(prof, key=lambda k: prof[k][0])
        kms, kn, kbytes = prof[name]
        achieved = kbytes / max(kms, 1e-9) / 1e6
        # algorithmic traffic of the whole step per asset byte: 1 (chunk + hash read) + u (gather read) + u (gather write)
        # + u (codec read) + c*u (codec write), u = unique fraction, c = compression ratio
        u = raw_all / float(world * cs.nbytes)
        c = stored_all / float(max(raw_all, 1))
        # DRAM traffic of the dominant kernel per launch: (dram__bytes_read + dram__bytes_write) / algorithmic bytes from the `ncu --set full`
        # captures under profiles/ (k_lz4_blocks_v2, in-place layout: 2.38 GB read + 2.09 GB written for 2.147 GB of block payload,
        # profiles/r03q_lz4v2_inplace_ncu.txt; scan and leaves: profiles/r01p_*_ncu.txt), applied to this run's algorithmic bytes per launch
        ncu_traffic_ratio = {"k_lz4_blocks": (2.382096e9 + 2.089249e9) / 2.147483648e9, "k_blake3_leaves": (8.963703e9 + 0.340567e9) / 8.589934592e9,
                             "k_hpcdc_scan": (8.609740e9 + 0.019096e9) / 8.589934592e9}
        traffic = round(kbytes / max(kn, 1) * ncu_traffic_ratio[name]) if name in ncu_traffic_ratio else None
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": "GiB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(step_ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "configs[2] per GPU: chunk+BLAKE3+LZ4 (CreateVersionIndex + CreateMissingContent + WriteContent through the compress "
                                   "block store) on %d synthetic assets totalling %.1f GiB per GPU with 50 %% byte redundancy, target_chunk_size 65536, "
                                   "blocks 8 MiB / 1024 chunks%s" % (cs.count, cs.nbytes / GIB, "" if world == 1 else
                                                                      "; %d GPUs = ONE version of %d assets / %.0f GiB: configs[4]'s full upsync with the NCCL "
                                                                      "all-gather of the chunk-hash table" % (world, cs.count * world, world * cs.nbytes / GIB)),
                       "bytes_per_gpu": cs.nbytes, "assets_per_gpu": cs.count, "l2": "inputs (%.0f GiB per GPU) far larger than L2; no flush needed" % (cs.nbytes / GIB),
                       "parity": parity, "index_bytes": up.index_bytes, "unique_frac": round(u, 4), "blocks": blocks_all, "stored_bytes": stored_all,
                       "compression_ratio": round(c, 4), "blocks_rank0": blocks_mine},
            "hbm_roofline_frac_whole_step": round(value * GIB / 1e9 / world / peak, 4),
            "hbm_roofline_frac_whole_step_all_traffic": round(value * GIB / 1e9 / world / peak * (1 + 3 * u + c * u), 4),
            "roofline": {"bound": "hbm", "kernel": name, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "traffic_source": "ncu --set full (profiles/r03q_lz4v2_inplace_ncu.txt): DRAM read + write / algorithmic bytes = 2.08 for the "
                                                                "in-place LZ4 encoder (the payload is read once, 1.11x with re-reads that miss L2; the whole "
                                                                "LZ4 stream, 0.97x, is written by the same kernel), scaled to this run's bytes per launch",
                         "peak_source": peak_src, "per_kernel": per_kernel,
                         "note": "algorithmic bytes of the codec kernel = the unique payload bytes it reads (SURVEY.md §8d); rank 0's kernels; a kernel's "
                                 "time = the union of its launches' intervals (two write batches are in flight on two streams); the LZ4 parse is a "
                                 "latency-bound serial chain per stored block, see DESIGN.md §4.6"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "index_only": index_only, "zstd_pak": zstd_pak,
        }
        emit(line)
    if comm is not None:
        ctx.comm_destroy(comm)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


_RESULT_OUT = None


def emit(line):
    """the ONE JSON line of the contract, on the process's original stdout"""
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _RESULT_OUT
    args = parse_args()
    # Libraries write banners to fd 1 (NCCL prints "NCCL version ..." there when NCCL_DEBUG is set): keep the real stdout for the result
    # line only and send everything else to stderr.
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
