"""debug: compare the GPU WriteContent with the reference upsync block by block on the bench's configs[2] sample"""
import sys, os
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
import numpy as np
import longtail_b200, oracle_lib as ol
import bench
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 32.0
sample_gib = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
codec = sys.argv[3] if len(sys.argv) > 3 else "lz4"
tag = longtail_b200.COMPRESSION_LZ4 if codec == "lz4" else longtail_b200.COMPRESSION_ZSTD_DEFAULT
GIB = 1 << 30
total = int(gib * GIB)
count = max(8, int(round(10000 * gib / 128.0)))
sizes = bench.config3_asset_sizes(total, count)
ctx = longtail_b200.Context(0)
pool = max(8, int(8192 * gib / 128.0))
assets, offs, off, acc = [], [], 0, 0
k = 0
for i, sz in enumerate(sizes):
    if acc + sz > sample_gib * GIB and k:
        break
    offs.append(off); off += (sz + 255) & ~255; acc += sz; k += 1
arena_bytes = off + 4096
arena = ctx.device_alloc(arena_bytes)
for i in range(k):
    ctx.synth_fill(arena + offs[i], sizes[i], seed=2, asset_id=i, class_mode=1, shared_permille=500, pool_segments=pool)
ctx.synchronize()
assets = [("a/%05d.bin" % i, ctx.to_host(arena + offs[i], sizes[i])) for i in range(k)]
al = longtail_b200.AssetList([p for p, _ in assets], sizes[:k])
v = ctx.index_device_assets(arena, arena_bytes, al, offs, [tag] * k, target_chunk_size=65536)
vi = longtail_b200.parse_version_index(v)
blocks = ctx.write_blocks_device(arena, arena_bytes, vi["chunk_hashes"], vi["chunk_sizes"], vi["chunk_tags"], ctx.unique_chunk_offsets(vi["chunk_count"]))
ref = ol.Reference()
want_blocks, want_v = ref.upsync(assets, 65536, tags=[tag] * k, workers=16)
print("index identical:", v == want_v, "blocks gpu %d ref %d" % (len(blocks), len(want_blocks)))
got = dict(blocks)
nd = 0
for h, b in want_blocks:
    g = got.get(h)
    if g is None:
        print("missing block %016x" % h); nd += 1; continue
    if g != b:
        nd += 1
        n = min(len(g), len(b)); first = next((i for i in range(n) if g[i] != b[i]), n)
        cnt = int(np.frombuffer(b[12:16], "<u4")[0])
        print("block %016x differs: gpu %d bytes ref %d bytes, first diff at %d (index is %d bytes), chunks %d" % (h, len(g), len(b), first, 20 + 12 * cnt, cnt))
        if nd < 3:
            raw = np.frombuffer(b[20 + 12 * cnt:20 + 12 * cnt + 8], "<u4")
            print("   ref header raw %d comp %d; gpu header %s" % (raw[0], raw[1], np.frombuffer(g[20 + 12 * cnt:20 + 12 * cnt + 8], "<u4")))
            # decode both and compare payloads; save the raw block for offline repro
            o = ol.Oracle()
            payload = o.lz4_decompress(b[20 + 12 * cnt + 8:], int(raw[0])) if codec == "lz4" else None
            if payload is not None:
                np.frombuffer(payload, np.uint8).tofile(os.path.join(ROOT, "gpurun_out", "diff_block_%d.bin" % nd))
                mine = o.lz4_compress(np.frombuffer(payload, np.uint8))
                print("   oracle(lz4) of the payload == ref: %s, == gpu: %s" % (mine == b[20 + 12 * cnt + 8:], mine == g[20 + 12 * cnt + 8:]))
print("differing blocks:", nd)

a, b = longtail_b200.parse_version_index(v), longtail_b200.parse_version_index(want_v)
print("chunk_count gpu %d ref %d; index_count %d %d" % (a["chunk_count"], b["chunk_count"], a["asset_chunk_index_count"], b["asset_chunk_index_count"]))
bad_assets = [i for i in range(k) if a["asset_chunk_counts"][i] != b["asset_chunk_counts"][i] or a["content_hashes"][i] != b["content_hashes"][i]]
print("assets with different chunk count / content hash:", [(i, sizes[i], int(a["asset_chunk_counts"][i]), int(b["asset_chunk_counts"][i])) for i in bad_assets][:10])
mn, av, mx = longtail_b200.chunker_params(65536)
part = 65536 * 1024
for i in bad_assets[:3]:
    data = assets[i][1]
    # per part: reference chunker vs GPU chunk_ranges
    for p0 in range(0, sizes[i], part):
        n = min(part, sizes[i] - p0)
        want = ref.chunk(data[p0:p0 + n], mn, av, mx)
        got = ctx.chunk_ranges(arena, arena_bytes, [(offs[i] + p0, n, 0)], mn, av, mx)["sizes"]
        if want.tolist() != got.tolist():
            j = next(x for x in range(min(len(want), len(got))) if want[x] != got[x]) if len(want) and len(got) else 0
            print("asset %d size %d part at %d (n=%d, n%%8192=%d): chunk #%d ref %s gpu %s (of %d / %d); offset in part %d" % (
                i, sizes[i], p0, n, n % 8192, j, want[j:j + 3].tolist(), got[j:j + 3].tolist(), len(want), len(got), int(want[:j].sum())))
            np.asarray(data[p0:p0 + n]).tofile(os.path.join(ROOT, "gpurun_out", "diff_part_%d.bin" % i))
            break

for i in bad_assets[:2]:
    data = assets[i][1]
    lens = np.concatenate([ref.chunk(data[p0:p0 + min(part, sizes[i] - p0)], mn, av, mx) for p0 in range(0, sizes[i], part)])
    co = np.concatenate([[0], np.cumsum(lens.astype(np.uint64))[:-1]]).astype(np.uint64)
    want_h = ref.hash_segments(ol.HASH_BLAKE3, data, co, lens)
    got_seg = ctx.hash_segments(arena, arena_bytes, co + np.uint64(offs[i]), lens)
    got_rng = ctx.chunk_ranges(arena, arena_bytes, [(offs[i] + p0, min(part, sizes[i] - p0), 0) for p0 in range(0, sizes[i], part)], mn, av, mx)
    bad_seg = np.nonzero(want_h != got_seg)[0]
    bad_rng = np.nonzero(want_h != got_rng["hashes"])[0]
    print("asset %d: hash_segments mismatches %s ; chunk_ranges mismatches %s" % (i, bad_seg[:8].tolist(), bad_rng[:8].tolist()))
    for j in list(bad_rng[:4]) + list(bad_seg[:4]):
        print("   chunk %d: offset %d (mod 16 = %d, arena mod 1024 = %d) len %d (leaves %d, len mod 1024 = %d, mod 64 = %d) ref %016x seg %016x rng %016x" % (
            j, int(co[j]), int(co[j]) % 16, (int(co[j]) + offs[i]) % 1024, int(lens[j]), (int(lens[j]) + 1023) // 1024, int(lens[j]) % 1024, int(lens[j]) % 64,
            int(want_h[j]), int(got_seg[j]), int(got_rng["hashes"][j])))
    # is it deterministic, and does it depend on the batch?
    again = ctx.chunk_ranges(arena, arena_bytes, [(offs[i] + p0, min(part, sizes[i] - p0), 0) for p0 in range(0, sizes[i], part)], mn, av, mx)["hashes"]
    print("   second run identical to the first:", again.tolist() == got_rng["hashes"].tolist())
    for j in bad_rng[:2]:
        one = ctx.hash_segments(arena, arena_bytes, [int(co[j]) + offs[i]], [int(lens[j])])
        print("   chunk %d hashed alone: %016x (ref %016x)" % (j, int(one[0]), int(want_h[j])))
        np.asarray(data[int(co[j]):int(co[j]) + int(lens[j])]).tofile(os.path.join(ROOT, "gpurun_out", "diff_chunk_%d.bin" % j))

for i in bad_assets[:2]:
    sa, sb = int(a["asset_chunk_index_starts"][i]), int(b["asset_chunk_index_starts"][i])
    n = int(a["asset_chunk_counts"][i])
    ia, ib = a["asset_chunk_indexes"][sa:sa + n], b["asset_chunk_indexes"][sb:sb + n]
    ha, hb = a["chunk_hashes"][ia], b["chunk_hashes"][ib]
    za, zb = a["chunk_sizes"][ia], b["chunk_sizes"][ib]
    bad = np.nonzero((ha != hb) | (za != zb))[0]
    print("full-run asset %d: %d chunks, differing positions %s" % (i, n, bad[:10].tolist()))
    lens = zb
    co = np.concatenate([[0], np.cumsum(lens.astype(np.uint64))[:-1]]).astype(np.uint64)
    for j in bad[:6]:
        print("   pos %d: offset %d (part offset %d) size gpu %d ref %d hash gpu %016x ref %016x" % (j, int(co[j]), int(co[j]) % part, int(za[j]), int(zb[j]), int(ha[j]), int(hb[j])))
# which assets precede asset 62 in the batch and how are parts laid out
print("asset 61 size %d, asset 62 size %d offs %d (mod 8192 = %d), asset 63 size %d" % (sizes[61], sizes[62], offs[62], offs[62] % 8192, sizes[63]))
