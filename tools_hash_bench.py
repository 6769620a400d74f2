"""quick per-hash-backend throughput probe: python tools_hash_bench.py [gib] [meow|blk2|blk3 ...]"""
import sys
import longtail_b200
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
names = sys.argv[2:] or ["meow", "blk2", "blk3"]
types = {"meow": longtail_b200.HASH_MEOW, "blk2": longtail_b200.HASH_BLAKE2, "blk3": longtail_b200.HASH_BLAKE3}
ctx = longtail_b200.Context(0)
n = int(gib * (1 << 30))
ptr = ctx.device_alloc(n + 4096)
ctx.synth_fill(ptr, n, seed=1)
ctx.synchronize()
al = longtail_b200.AssetList(["f"], [n])
for name in names:
    for it in range(2):
        ctx.profile_reset(); ctx.profile_enable(True)
        ctx.index_device_assets(ptr, n + 4096, al, [0], None, hash_type=types[name], target_chunk_size=65536, copy=False)
        ctx.profile_enable(False)
    print(name, {k: "%.2f ms %.1f GB/s" % (x[0], x[2] / max(x[0], 1e-9) / 1e6) for k, x in ctx.profile_read().items() if x[0]})
