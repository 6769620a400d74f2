"""Pins the CPU restatement (oracle/lt_oracle.c) — against the reference's own golden vectors and
KATs, against fixtures generated from the unmodified reference (tests/golden/golden.json), and,
when oracle/_ref/libref_shim.so is present, byte-for-byte against the reference itself."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from synth import chunker_params, small_tree, synth_bytes

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))

# reference test/test.cpp:3422-3445 (ChunkerLargeFile): min/avg/max 16384/65536/262144 over testdata/chunker.input
GOLDEN_CHUNKS = [81590, 46796, 36543, 83172, 76749, 79550, 41484, 20326, 31652, 19995, 103873, 38087, 38377, 23449,
                 47321, 86692, 28268, 65465, 33255, 65932]
KAT_STRING = b"This is the first test string which is fairly long and should - reconstructed properly, than you very much\0"


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def test_chunker_golden_vector(oracle):
    data = np.fromfile(os.path.join(HERE, "golden", "chunker.input"), dtype=np.uint8)
    assert data.size == 1048576
    assert oracle.chunk(data, 16384, 65536, 262144).tolist() == GOLDEN_CHUNKS


def test_hash_kats(oracle):
    assert oracle.hash(ol.HASH_BLAKE3, KAT_STRING) == 0xD38BBE79F1F03FDA  # test.cpp:472
    assert oracle.hash(ol.HASH_BLAKE2, KAT_STRING) == 0xD336E5AFA4FA1F4D  # test.cpp:460
    assert oracle.hash(ol.HASH_MEOW, KAT_STRING) == 0x4EDC68DAC105C4EE  # test.cpp:484


def test_discriminator(oracle):
    assert oracle.discriminator(32768) == 24680
    assert oracle.discriminator(65536) == 49535


def test_lz4_size_pin(oracle):
    # test.cpp:2185-2192: 1147 x 0x0d + 4711 x 0x4d compresses to 38 bytes of LZ4 payload
    data = np.concatenate([np.full(1147, 0x0D, np.uint8), np.full(4711, 0x4D, np.uint8)])
    comp = oracle.lz4_compress(data)
    assert len(comp) == 38
    assert oracle.lz4_decompress(comp, data.size) == data.tobytes()


@pytest.mark.parametrize("case", GOLDEN["chunker"], ids=lambda c: "%s-%d-t%d" % (c["kind"], c["n"], c["target"]))
def test_chunker_fixtures(oracle, case):
    mn, av, mx = chunker_params(case["target"])
    lens = oracle.chunk(synth_bytes(case["seed"], case["n"], case["kind"]), mn, av, mx)
    assert lens.size == case["count"]
    assert lens[:8].tolist() == case["first"]
    assert sha(lens.astype("<u4").tobytes()) == case["sha256"]
    assert int(lens.sum()) == case["n"]


@pytest.mark.parametrize("case", GOLDEN["hash"], ids=lambda c: str(c["n"]))
def test_hash_fixtures(oracle, case):
    x = synth_bytes(100 + case["n"], case["n"])
    assert "%016x" % oracle.hash(ol.HASH_BLAKE3, x) == case["blk3"]
    assert "%016x" % oracle.hash(ol.HASH_BLAKE2, x) == case["blk2"]
    assert "%016x" % oracle.hash(ol.HASH_MEOW, x) == case["meow"]


@pytest.mark.parametrize("case", GOLDEN["lz4"], ids=lambda c: "%s-%d" % (c["kind"], c["n"]))
def test_lz4_fixtures(oracle, case):
    x = synth_bytes(200 + case["n"], case["n"], case["kind"])
    comp = oracle.lz4_compress(x)
    assert len(comp) == case["size"]
    assert sha(comp) == case["sha256"]
    assert oracle.lz4_decompress(comp, x.size) == x.tobytes()


def _tree(target):
    assets = small_tree(target)
    tags = [ol.COMP_LZ4 if i % 3 else 0 for i in range(len(assets))]
    perms = [0o644 + i for i in range(len(assets))]
    return assets, tags, perms


@pytest.mark.parametrize("case", GOLDEN["version_index"], ids=lambda c: "t%d-%s" % (c["target"], c["hash"]))
def test_version_index_fixtures(oracle, case):
    assets, tags, perms = _tree(case["target"])
    ht = {"blk3": ol.HASH_BLAKE3, "blk2": ol.HASH_BLAKE2}[case["hash"]]
    v = oracle.create_version_index(assets, case["target"], hash_type=ht, tags=tags, perms=perms)
    assert len(v) == case["size"]
    assert sha(v) == case["sha256"]


@pytest.mark.parametrize("case", GOLDEN["upsync"], ids=lambda c: "t%d" % c["target"])
def test_upsync_fixtures(oracle, case):
    assets, tags, perms = _tree(case["target"])
    blocks, v = oracle.upsync(assets, case["target"], max_block_size=case["max_block_size"],
                              max_chunks_per_block=case["max_chunks_per_block"], tags=tags, perms=perms)
    assert len(blocks) == case["blocks"]
    assert sha(v) == case["version_sha256"]
    assert sha(b"".join(h.to_bytes(8, "little") + b for h, b in blocks)) == case["blocks_sha256"]


@pytest.mark.parametrize("case", GOLDEN["zstd"], ids=lambda c: "%s-%d" % (c["kind"], c["n"]))
def test_zstd_fixtures(oracle, case):
    """ZStd level 3 frames ('ztd2', lib/zstd/longtail_zstd.c:107-140) == the fixtures generated from the reference"""
    c = oracle.zstd_compress(synth_bytes(500 + case["n"], case["n"], case["kind"]))
    assert len(c) == case["size"]
    assert sha(c) == case["sha256"]


# ---------------------------------------------------------------- live differential vs the unmodified reference


def test_reference_matches_its_own_vectors(reference):
    if reference is None:
        pytest.skip("oracle/_ref/libref_shim.so not built (needs /root/reference)")
    data = np.fromfile(os.path.join(HERE, "golden", "chunker.input"), dtype=np.uint8)
    assert reference.chunk(data, 16384, 65536, 262144).tolist() == GOLDEN_CHUNKS
    assert reference.hash(ol.HASH_BLAKE3, KAT_STRING) == 0xD38BBE79F1F03FDA
    assert reference.hash(ol.HASH_BLAKE2, KAT_STRING) == 0xD336E5AFA4FA1F4D
    assert reference.hash(ol.HASH_MEOW, KAT_STRING) == 0x4EDC68DAC105C4EE  # test.cpp:484


@pytest.mark.parametrize("seed,n,target,kind", [(21, 6 << 20, 65536, "rand"), (22, 3 << 20, 32768, "text"), (23, 1 << 20, 2048, "nib"),
                                                (24, 700001, 512, "p7"), (25, 65536 * 3, 64, "rand")])
def test_chunker_vs_reference(oracle, reference, seed, n, target, kind):
    if reference is None:
        pytest.skip("reference not built")
    x = synth_bytes(seed, n, kind)
    mn, av, mx = chunker_params(target)
    assert oracle.chunk(x, mn, av, mx).tolist() == reference.chunk(x, mn, av, mx).tolist()


def test_meow_vs_reference_every_small_length(oracle, reference):
    """MeowEnd's residual / lane / 16-byte branches (meow_hash_x64_aesni.h:583-700): all lengths 0..1100 and a few large ones"""
    if reference is None:
        pytest.skip("reference not built")
    x = synth_bytes(77, 300000)
    for n in list(range(0, 1101)) + [4096, 65535, 65536, 131072, 299999]:
        assert oracle.hash(ol.HASH_MEOW, x[3:3 + n]) == reference.hash(ol.HASH_MEOW, x[3:3 + n]), n


@pytest.mark.parametrize("n,kind", [(70000, "text"), (65546, "nib"), (65547, "nib"), (2 << 20, "text"), (9 << 20, "nib"), (1 << 20, "rand")])
def test_lz4_vs_reference(oracle, reference, n, kind):
    if reference is None:
        pytest.skip("reference not built")
    x = synth_bytes(300 + n, n, kind)
    assert oracle.lz4_compress(x) == reference.compress(ol.COMP_LZ4, x)


def _in_place_offset(n):
    """longtail_b200/csrc/lz4.cu lz4_in_place_offset: how far behind the output start the source of an in-place block sits"""
    return (n // 255 + 16 + 65536 + 256 + 15) & ~15


@pytest.mark.parametrize("n,kind", [(65547, "nib"), (70000, "text"), (300000, "rec"), (2 << 20, "text"), (3 << 20, "nib"), (2 << 20, "rand"),
                                    (1 << 20, "zero"), (5 << 20, "mix")])
def test_lz4_in_place_margin(oracle, n, kind):
    """the layout of the device write path: source at the END of the block's own output slot, lz4_in_place_offset(n)
    bytes behind the output start.  The reference's parse only reads source bytes >= anchor - 65 536 and has written at most
    consumed * (1 + 1/255) + 16 bytes by then, so the output never reaches a byte it still reads: compressing in ONE buffer gives the
    bytes of the ordinary call."""
    if kind == "mix":
        x = np.concatenate([synth_bytes(5, n // 4, k) for k in ("rand", "text", "nib", "rec")])
    else:
        x = synth_bytes(1200 + n, n, kind)
    want = oracle.lz4_compress(x)
    assert oracle.lz4_compress_in_place(x, 8 + _in_place_offset(x.size)) == want
    assert oracle.lz4_decompress(want, x.size) == x.tobytes()
    if kind in ("text", "nib", "rec") and n >= 300000:
        # the 64 KiB of the margin are needed, not a safety habit: with 4 KiB instead the parse reads bytes its own output has overwritten
        assert oracle.lz4_compress_in_place(x, 8 + x.size // 255 + 16 + 4096) != want


@pytest.mark.parametrize("n,kind", [(100, "rec"), (5000, "rec"), (70000, "rec"), (140000, "text"), (270000, "rec"), (3 << 20, "rec"), (1 << 20, "nib"),
                                    (600000, "zero"), (9 << 20, "rec")])
def test_zstd_vs_reference(oracle, reference, n, kind):
    if reference is None:
        pytest.skip("reference not built")
    x = synth_bytes(900 + n, n, kind)
    c = oracle.zstd_compress(x)
    assert c == reference.compress(ol.COMP_ZSTD_DEFAULT, x)
    assert c == reference.compress(ol.COMP_ZSTD_MIN, x)  # 'ztd1' -> level 0 == level 3
    assert len(c) < n or kind == "rand"


def test_zstd_vs_reference_adversarial(oracle, reference):
    """the same rare-path inputs the GPU test uses, oracle vs the unmodified reference"""
    if reference is None:
        pytest.skip("reference not built")
    rec = synth_bytes(830, 1 << 20, "rec")
    rnd = synth_bytes(831, 1 << 20, "rand")
    bufs = [
        np.concatenate([np.zeros(200000, np.uint8), rec[:70000], np.full(131072 * 2, 7, np.uint8), rec[:1000]]),
        np.tile(rec[:131072 - 5], 9),
        np.concatenate([rec[:131072], rnd[:131072]] * 6),
        np.concatenate([rec[:300000], np.zeros(2200000, np.uint8), rec[:300000]]),
        np.concatenate([rec[:300000], np.zeros(1700000, np.uint8), rec[:300000]]),
        np.tile(np.arange(256, dtype=np.uint8), 3000),
        np.concatenate([synth_bytes(832 + i, 37 + 11 * i, "text") for i in range(400)] * 3),
        synth_bytes(833, 131072 * 3 + 6, "rec"),
        synth_bytes(834, 131072 * 3 + 7, "rec"),
    ]
    for i, b in enumerate(bufs):
        assert oracle.zstd_compress(b) == reference.compress(ol.COMP_ZSTD_DEFAULT, b), "case %d" % i


def test_zstd_vs_reference_real_files(oracle, reference):
    """source text and machine code of the reference tree itself: real match / literal / sequence statistics"""
    if reference is None or not os.path.isdir("/root/reference"):
        pytest.skip("reference not available")
    import glob
    files = sorted(glob.glob("/root/reference/src/*.c") + glob.glob("/root/reference/lib/zstd/ext/compress/*.c"))[:12]
    files.append(os.path.join(HERE, "..", "oracle", "_ref", "libref_shim.so"))
    for f in files:
        x = np.fromfile(f, dtype=np.uint8)[:6 << 20]
        assert oracle.zstd_compress(x) == reference.compress(ol.COMP_ZSTD_DEFAULT, x), f


def test_upsync_vs_reference_default_params(oracle, reference):
    if reference is None:
        pytest.skip("reference not built")
    assets = [("big/f%03d.bin" % i, synth_bytes(400 + i, 150000 + 410000 * i, "rand" if i % 2 else "nib")) for i in range(10)]
    assets.append(("big/f003_copy.bin", assets[3][1].copy()))
    assets.sort(key=lambda a: a[0].encode())
    tags = [ol.COMP_LZ4] * len(assets)
    ours = oracle.upsync(assets, 32768, tags=tags)
    theirs = reference.upsync(assets, 32768, tags=tags, workers=3)
    assert ours[1] == theirs[1]
    assert ours[0] == theirs[0]
