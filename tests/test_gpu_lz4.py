"""GPU parity (-m gpu) of the WriteContent half: LZ4 block bytes, block packing, block hashes and serialised StoredBlocks
against the CPU oracle / reference and the golden fixtures."""
import hashlib
import json
import os
import struct

import numpy as np
import pytest

import oracle_lib as ol
from synth import small_tree, synth_bytes

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


@pytest.fixture(scope="module")
def ctx():
    import longtail_b200
    c = longtail_b200.Context(0)
    yield c
    c.close()


def gpu_lz4(ctx, data):
    """compress one buffer as a single one-chunk block; -> (lz4 bytes, whole serialised block)"""
    n = int(data.size)
    ptr = ctx.device_alloc(n + 64)
    try:
        ctx.to_device(ptr, data)
        blocks = ctx.write_blocks_device(ptr, n + 64, [0x1234], [n], [ol.COMP_LZ4], [0], max_block_size=max(n, 1) * 2, max_chunks_per_block=4)
    finally:
        ctx.device_free(ptr)
    assert len(blocks) == 1
    blob = blocks[0][1]
    # block index: u64 hash, u32 hash id, u32 count(=1), u32 tag, u64 chunk hash, u32 chunk size = 32 bytes, then payload
    raw, comp = struct.unpack_from("<II", blob, 32)
    assert raw == n and comp == len(blob) - 40
    return blob[40:], blob


@pytest.mark.parametrize("case", [c for c in GOLDEN["lz4"] if c["n"] > 0], ids=lambda c: "%s-%d" % (c["kind"], c["n"]))
def test_lz4_golden_fixtures(ctx, case):
    x = synth_bytes(200 + case["n"], case["n"], case["kind"])
    comp, _ = gpu_lz4(ctx, x)
    assert len(comp) == case["size"]
    assert sha(comp) == case["sha256"]


@pytest.mark.parametrize("n,kind", [(1, "rand"), (12, "zero"), (13, "zero"), (14, "text"), (64, "p3"), (4096, "text"), (65546, "nib"), (65547, "nib"),
                                    (70000, "rand"), (300001, "bit"), (1 << 20, "text"), (3 << 20, "p48"), (9 << 20, "nib"), (9 << 20, "zero"),
                                    ((8 << 20) + 12345, "rand"), (2 << 20, "p1")])
def test_lz4_vs_oracle(ctx, oracle, n, kind):
    x = synth_bytes(500 + n, n, kind)
    comp, _ = gpu_lz4(ctx, x)
    want = oracle.lz4_compress(x)
    assert len(comp) == len(want)
    assert comp == want
    assert oracle.lz4_decompress(comp, n) == x.tobytes()


def _with_repeats(n, seed, dist, length, every):
    """random bytes with copies of `length` bytes taken `dist` bytes back, every `every` bytes"""
    x = synth_bytes(seed, n, "rand")
    for i in range(max(dist, 100000), n - length - 16, every):
        x[i:i + length] = x[i - dist:i - dist + length]
    return x


# the shared-memory-table encoder (k_lz4_blocks_v2) keeps 17 position bits + a 14-bit tag per slot and sweeps stale entries:
# matches at the edge of the 65 535-byte window, matches longer than the window and longer than 2^17 (table cleared / multi-step
# sweep), the v1 <-> v2 hand-over sizes (65 546 / 65 547 bytes, 16 MiB / 16 MiB + 1) and every synthetic class at block size
@pytest.mark.parametrize("name,make", [
    ("rep65535", lambda: _with_repeats(3 << 20, 901, 65535, 5000, 200001)),
    ("rep65536", lambda: _with_repeats(3 << 20, 902, 65536, 5000, 200003)),
    ("rep60000x200k", lambda: _with_repeats(4 << 20, 903, 60000, 200000, 700001)),
    ("rep70000x100k", lambda: _with_repeats(4 << 20, 904, 70000, 100000, 500009)),
    ("rep300x70000", lambda: _with_repeats(2 << 20, 905, 300, 70000, 150001)),
    ("zero_then_rand", lambda: np.concatenate([np.zeros(200000, np.uint8), synth_bytes(906, 300000, "rand"), np.zeros(140000, np.uint8),
                                               synth_bytes(907, 100000, "text")])),
    ("text_9m", lambda: synth_bytes(908, 9200000, "text")),
    ("nib_9m", lambda: synth_bytes(909, 9200000, "nib")),
    ("rec_9m", lambda: synth_bytes(910, 9200000, "rec")),
    ("bit_2m", lambda: synth_bytes(911, 2 << 20, "bit")),
    ("rand_16m", lambda: synth_bytes(912, 16 << 20, "rand")),
    ("text_16m_plus1", lambda: synth_bytes(913, (16 << 20) + 1, "text")),
    ("mixed_9m", lambda: np.concatenate([synth_bytes(914 + i, 1 << 20, k) for i, k in enumerate(["rand", "text", "nib", "text", "rec", "rand", "nib", "text", "zero"])])),
], ids=lambda v: v if isinstance(v, str) else "")
def test_lz4_shared_table_encoder_vs_oracle(ctx, oracle, name, make):
    x = make()
    comp, _ = gpu_lz4(ctx, x)
    want = oracle.lz4_compress(x)
    assert len(comp) == len(want)
    assert comp == want


def test_lz4_many_blocks_one_launch(ctx, oracle):
    """more blocks than resident warps of one wave, of mixed kinds and sizes, through the CompressionAPI batch entry point"""
    kinds = ["text", "nib", "rand", "rec", "zero", "p3", "bit"]
    bufs = [synth_bytes(3000 + i, 66000 + 37 * i + (i % 5) * 40000, kinds[i % len(kinds)]) for i in range(300)]
    bufs += [synth_bytes(4000 + i, 1000 + 977 * i, kinds[i % len(kinds)]) for i in range(40)]
    got = ctx.lz4_compress_host(bufs)
    for b, g in zip(bufs, got):
        assert g == oracle.lz4_compress(b)


def test_lz4_size_pin(ctx):
    # reference test/test.cpp:2185-2192: this input compresses to 38 bytes of LZ4 payload
    data = np.concatenate([np.full(1147, 0x0D, np.uint8), np.full(4711, 0x4D, np.uint8)])
    comp, _ = gpu_lz4(ctx, data)
    assert len(comp) == 38


def _tree(target):
    assets = small_tree(target)
    tags = [ol.COMP_LZ4 if i % 3 else 0 for i in range(len(assets))]
    perms = [0o644 + i for i in range(len(assets))]
    return assets, tags, perms


@pytest.mark.parametrize("case", GOLDEN["upsync"], ids=lambda c: "t%d" % c["target"])
def test_upsync_blocks_match_reference(ctx, oracle, case):
    """index on the device, then pack + hash + gather + LZ4 + serialise every block: bytes equal to the reference's upsync"""
    import longtail_b200
    target = case["target"]
    assets, tags, perms = _tree(target)
    al = longtail_b200.AssetList([p for p, _ in assets], [d.size for _, d in assets], perms)
    offs, off = [], 0
    for _, d in assets:
        offs.append(off)
        off = (off + d.size + 255) & ~255
    arena = np.zeros(off + 256, np.uint8)
    for o, (_, d) in zip(offs, assets):
        arena[o:o + d.size] = d
    ptr = ctx.device_alloc(arena.size)
    try:
        ctx.to_device(ptr, arena)
        v = ctx.index_device_assets(ptr, arena.size, al, offs, tags, target_chunk_size=target)
        vi = longtail_b200.parse_version_index(v)
        uoff = ctx.unique_chunk_offsets(vi["chunk_count"])
        blocks = ctx.write_blocks_device(ptr, arena.size, vi["chunk_hashes"], vi["chunk_sizes"], vi["chunk_tags"], uoff,
                                         max_block_size=case["max_block_size"], max_chunks_per_block=case["max_chunks_per_block"])
    finally:
        ctx.device_free(ptr)
    assert len(blocks) == case["blocks"]
    assert sha(v) == case["version_sha256"]
    assert sha(b"".join(h.to_bytes(8, "little") + b for h, b in blocks)) == case["blocks_sha256"]
    want_blocks, want_v = oracle.upsync(assets, target, max_block_size=case["max_block_size"], max_chunks_per_block=case["max_chunks_per_block"],
                                        tags=tags, perms=perms)
    assert v == want_v
    assert blocks == want_blocks


def test_upsync_default_parameters_against_reference(ctx, oracle, reference):
    """8 MiB blocks / 1024 chunks (CLI defaults, cmd/main.c:2985-3021), target 32768, duplicates across assets"""
    import longtail_b200
    assets = [("big/f%03d.bin" % i, synth_bytes(400 + i, 150000 + 1310000 * i, ["rand", "nib", "text"][i % 3])) for i in range(12)]
    assets.append(("big/f003_copy.bin", assets[3][1].copy()))
    assets.sort(key=lambda a: a[0].encode())
    tags = [ol.COMP_LZ4] * len(assets)
    al = longtail_b200.AssetList([p for p, _ in assets], [d.size for _, d in assets])
    offs, off = [], 0
    for _, d in assets:
        offs.append(off)
        off = (off + d.size + 255) & ~255
    arena = np.zeros(off + 256, np.uint8)
    for o, (_, d) in zip(offs, assets):
        arena[o:o + d.size] = d
    ptr = ctx.device_alloc(arena.size)
    try:
        ctx.to_device(ptr, arena)
        v = ctx.index_device_assets(ptr, arena.size, al, offs, tags, target_chunk_size=32768)
        vi = longtail_b200.parse_version_index(v)
        uoff = ctx.unique_chunk_offsets(vi["chunk_count"])
        blocks = ctx.write_blocks_device(ptr, arena.size, vi["chunk_hashes"], vi["chunk_sizes"], vi["chunk_tags"], uoff)
    finally:
        ctx.device_free(ptr)
    checker = reference if reference is not None else oracle
    want_blocks, want_v = checker.upsync(assets, 32768, tags=tags)
    assert v == want_v
    assert [h for h, _ in blocks] == [h for h, _ in want_blocks]
    assert blocks == want_blocks
