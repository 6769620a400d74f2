"""small codec workload for ncu: python tools_zstd_profile.py [frames] [mib] [kind] [zstd|lz4]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
import numpy as np
import longtail_b200
from synth import synth_bytes
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mib = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
kind = sys.argv[3] if len(sys.argv) > 3 else "nib"
codec = sys.argv[4] if len(sys.argv) > 4 else "zstd"
ctx = longtail_b200.Context(0)
bufs = [synth_bytes(40 + i, int(mib * (1 << 20)), kind) for i in range(frames)]
for it in range(2):
    t0 = time.perf_counter()
    out = ctx.zstd_compress_host(bufs) if codec == "zstd" else ctx.lz4_compress_host(bufs)
    t1 = time.perf_counter()
    print("pass %d: %d frames x %.1f MiB %s -> ratio %.3f in %.1f ms" % (it, frames, mib, kind, sum(map(len, out)) / sum(b.size for b in bufs), 1e3 * (t1 - t0)))
