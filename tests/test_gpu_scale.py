"""-m gpu: a configs[2]-shaped asset set (~1 GiB: 90 assets, log-uniform sizes up to 150 MiB so that several span parts, half of the
1 MiB segments shared between assets, random / 4-bit / text-like data, tags cycling raw / LZ4 / ZStd) through every entry path,
against the unmodified reference's upsync: VersionIndex and every StoredBlock byte for byte.  Sizes the unit tests do not reach
(this shape exposed a stale-candidate bug in the scan kernel that 15-asset trees never triggered)."""
import ctypes as C
import math
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = 65536


def _sizes(total, count, lo=4096.0, hi=150e6):
    x, out = 987654321, []
    for _ in range(count):
        x = (x * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
        u = (x >> 11) / float(1 << 53)
        out.append(math.exp(math.log(lo) + u * (math.log(hi) - math.log(lo))))
    scale = total / sum(out)
    return [max(1, int(v * scale)) + (i % 7) for i, v in enumerate(out)]  # ragged: not multiples of anything


@pytest.fixture(scope="module")
def workload():
    import longtail_b200
    ctx = longtail_b200.Context(0)
    sizes = _sizes(1 << 30, 90)
    sizes[5] = 0
    sizes[17] = TARGET * 1024          # exactly one part
    sizes[18] = 2 * TARGET * 1024 + 1  # two parts and one byte
    offs, off = [], 0
    for sz in sizes:
        offs.append(off)
        off += (sz + 255) & ~255
    arena_bytes = off + 4096
    arena = ctx.device_alloc(arena_bytes)
    for i, (o, sz) in enumerate(zip(offs, sizes)):
        if sz:
            ctx.synth_fill(arena + o, sz, seed=9, asset_id=i, class_mode=1, shared_permille=500, pool_segments=64)
    ctx.synchronize()
    assets = [("d%d/f%03d.bin" % (i % 4, i), ctx.to_host(arena + o, sz) if sz else np.zeros(0, np.uint8)) for i, (o, sz) in enumerate(zip(offs, sizes))]
    tags = [(0, ol.COMP_LZ4, ol.COMP_ZSTD_DEFAULT)[i % 3] for i in range(len(assets))]
    ref = ol.Reference()
    if not ref.available:
        pytest.skip("reference not built")
    want_blocks, want_index = ref.upsync(assets, TARGET, tags=tags, workers=8)
    yield dict(ctx=ctx, arena=arena, arena_bytes=arena_bytes, offs=offs, sizes=sizes, assets=assets, tags=tags, want_blocks=want_blocks, want_index=want_index)
    ctx.device_free(arena)
    ctx.close()


def test_index_device_and_write_blocks(workload):
    import longtail_b200
    w = workload
    ctx = w["ctx"]
    al = longtail_b200.AssetList([p for p, _ in w["assets"]], w["sizes"])
    v = ctx.index_device_assets(w["arena"], w["arena_bytes"], al, w["offs"], w["tags"], target_chunk_size=TARGET)
    assert v == w["want_index"]
    vi = longtail_b200.parse_version_index(v)
    blocks = ctx.write_blocks_device(w["arena"], w["arena_bytes"], vi["chunk_hashes"], vi["chunk_sizes"], vi["chunk_tags"],
                                     ctx.unique_chunk_offsets(vi["chunk_count"]))
    assert len(blocks) == len(w["want_blocks"]) and len(blocks) > 60
    assert [h for h, _ in blocks] == [h for h, _ in w["want_blocks"]]
    for (h, got), (_, want) in zip(blocks, w["want_blocks"]):
        assert got == want, "block %016x differs" % h
    kinds = {int(np.frombuffer(b[16:20], "<u4")[0]) for _, b in blocks}
    assert kinds == {0, ol.COMP_LZ4, ol.COMP_ZSTD_DEFAULT}


def test_index_host_assets(workload):
    import longtail_b200
    w = workload
    al = longtail_b200.AssetList([p for p, _ in w["assets"]], w["sizes"])
    v = w["ctx"].index_host_assets(al, [d for _, d in w["assets"]], w["tags"], target_chunk_size=TARGET)
    assert v == w["want_index"]


@pytest.mark.parametrize("hash_type", ["blk2", "meow"])
def test_other_hashes_at_scale(workload, hash_type):
    """first 24 assets (incl. the empty and the multi-part ones) with the BLAKE2s / Meow identifiers"""
    import longtail_b200
    w = workload
    ht_b, ht_o = {"blk2": (longtail_b200.HASH_BLAKE2, ol.HASH_BLAKE2), "meow": (longtail_b200.HASH_MEOW, ol.HASH_MEOW)}[hash_type]
    n = 24
    al = longtail_b200.AssetList([p for p, _ in w["assets"][:n]], w["sizes"][:n])
    v = w["ctx"].index_device_assets(w["arena"], w["arena_bytes"], al, w["offs"][:n], w["tags"][:n], hash_type=ht_b, target_chunk_size=TARGET)
    ref = ol.Reference()
    assert v == ref.create_version_index(w["assets"][:n], TARGET, hash_type=ht_o, tags=w["tags"][:n], workers=8)


def test_incremental_upsync_missing_chunks(workload):
    """upsync into a store that already holds part of the content (CreateMissingContent's DiffHashes on
    the device): the store knows every third unique chunk plus some foreign hashes; only the missing chunks are packed and written,
    in version order — blocks byte-identical to the reference's CreateMissingContent + WriteContent against the same store index"""
    import longtail_b200
    w = workload
    ctx = w["ctx"]
    al = longtail_b200.AssetList([p for p, _ in w["assets"]], w["sizes"])
    v = ctx.index_device_assets(w["arena"], w["arena_bytes"], al, w["offs"], w["tags"], target_chunk_size=TARGET)
    vi = longtail_b200.parse_version_index(v)
    uoff = ctx.unique_chunk_offsets(vi["chunk_count"])
    existing = np.concatenate([vi["chunk_hashes"][::3], np.arange(1, 5000, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)])
    missing = ctx.missing_chunks(vi["chunk_hashes"], existing)
    assert missing.tolist() == (~np.isin(vi["chunk_hashes"], existing)).tolist()
    assert 0 < missing.sum() < missing.size
    blocks = ctx.write_blocks_device(w["arena"], w["arena_bytes"], vi["chunk_hashes"][missing], vi["chunk_sizes"][missing], vi["chunk_tags"][missing],
                                     uoff[missing])
    ref = ol.Reference()
    want_blocks, want_index = ref.upsync(w["assets"], TARGET, tags=w["tags"], workers=8, existing_hashes=existing)
    assert want_index == v
    assert len(blocks) == len(want_blocks) and len(blocks) > 30
    for (h, got), (hw, want) in zip(blocks, want_blocks):
        assert h == hw and got == want, "block %016x differs" % hw


def test_upsync_to_disk_matches_reference_fsblockstore(workload, tmp_path):
    """lt_b200_write_blocks_device -> lt_b200_fs_store_sink (C to C, writer threads) leaves the same directory as
    the reference's compressblockstore -> fsblockstore: every chunks/xxxx/0x....lrb and store.lsi byte for byte; then an incremental
    upsync of a changed version (existing chunks from store.lsi -> lt_b200_missing_chunks -> only the new blocks) does too, and the
    unmodified reference reads both stores back identically"""
    import longtail_b200
    w = workload
    ctx = w["ctx"]
    n = 40  # ~450 MiB of the set
    assets, sizes, offs, tags = w["assets"][:n], w["sizes"][:n], w["offs"][:n], w["tags"][:n]
    ours, theirs = str(tmp_path / "ours"), str(tmp_path / "ref")
    ref = ol.Reference()

    def tree(root):
        out = {}
        for d, _, files in os.walk(root):
            for f in files:
                if f != "store.lsi.sync":
                    out[os.path.relpath(os.path.join(d, f), root)] = open(os.path.join(d, f), "rb").read()
        return out

    def upsync_ours(count):
        al = longtail_b200.AssetList([p for p, _ in assets[:count]], sizes[:count])
        v = ctx.index_device_assets(w["arena"], w["arena_bytes"], al, offs[:count], tags[:count], target_chunk_size=TARGET)
        vi = longtail_b200.parse_version_index(v)
        uoff = ctx.unique_chunk_offsets(vi["chunk_count"])
        st = longtail_b200.FsStore(ours, writer_threads=4)
        missing = ctx.missing_chunks(vi["chunk_hashes"], st.existing_chunks())
        ctx.write_blocks_device(w["arena"], w["arena_bytes"], vi["chunk_hashes"][missing], vi["chunk_sizes"][missing], vi["chunk_tags"][missing],
                                uoff[missing], fs_store=st)
        st.flush()
        stats = st.stats()
        st.close()
        return stats

    s1 = upsync_ours(25)
    n1 = ol.ref_upsync_to_dir(ref, assets[:25], TARGET, theirs, tags=tags[:25], workers=0)
    assert s1["blocks_written"] == n1 and n1 > 10
    a, b = tree(ours), tree(theirs)
    assert sorted(a) == sorted(b) and len(a) == n1 + 1
    assert all(a[k] == b[k] for k in a), [k for k in a if a[k] != b[k]][:3]
    # the version grows by 15 assets (half of their segments are shared with what the store holds)
    s2 = upsync_ours(n)
    n2 = ol.ref_upsync_to_dir(ref, assets, TARGET, theirs, tags=tags, workers=0)
    assert s2["blocks_written"] == n2 and n2 > 5 and s2["blocks_skipped"] == 0
    a, b = tree(ours), tree(theirs)
    assert sorted(a) == sorted(b) and len(a) == n1 + n2 + 1
    assert all(a[k] == b[k] for k in a), [k for k in a if a[k] != b[k]][:3]
    assert ol.ref_read_store_dir(ref, ours) == ol.ref_read_store_dir(ref, theirs)
