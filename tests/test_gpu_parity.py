"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle / the unmodified
reference / the committed golden fixtures — bit-exact chunk boundaries, chunk hashes and serialised VersionIndex."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from synth import chunker_params, small_tree, synth_bytes

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))
GOLDEN_CHUNKS = [81590, 46796, 36543, 83172, 76749, 79550, 41484, 20326, 31652, 19995, 103873, 38087, 38377, 23449,
                 47321, 86692, 28268, 65465, 33255, 65932]  # reference test/test.cpp:3422-3445


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


@pytest.fixture(scope="module")
def ctx():
    import longtail_b200
    c = longtail_b200.Context(0)
    yield c
    c.close()


class DeviceBytes:
    def __init__(self, ctx, data, pad=0):
        self.ctx = ctx
        self.size = int(data.size)
        self.ptr = ctx.device_alloc(self.size + pad + 16)
        if self.size:
            ctx.to_device(self.ptr, data)

    def free(self):
        self.ctx.device_free(self.ptr)


def check_ranges(ctx, oracle, data, ranges, mn, av, mx):
    """chunk_ranges over `ranges` of `data` == oracle chunker + oracle hash per range"""
    dev = DeviceBytes(ctx, data)
    try:
        got = ctx.chunk_ranges(dev.ptr, dev.size, ranges, mn, av, mx)
    finally:
        dev.free()
    exp_counts, exp_sizes, exp_offsets = [], [], []
    for off, size in [(r[0], r[1]) for r in ranges]:
        lens = oracle.chunk(data[off:off + size], mn, av, mx) if size else np.zeros(0, np.uint32)
        exp_counts.append(lens.size)
        o = off
        for n in lens:
            exp_sizes.append(int(n))
            exp_offsets.append(o)
            o += int(n)
    assert got["range_chunk_counts"].tolist() == exp_counts
    assert got["sizes"].tolist() == exp_sizes
    assert got["offsets"].tolist() == exp_offsets
    exp_hashes = oracle.hash_segments(ol.HASH_BLAKE3, data, np.array(exp_offsets, np.uint64), np.array(exp_sizes, np.uint32))
    assert got["hashes"].tolist() == exp_hashes.tolist()
    return got


def test_golden_chunker_vector(ctx, oracle):
    data = np.fromfile(os.path.join(HERE, "golden", "chunker.input"), dtype=np.uint8)
    got = check_ranges(ctx, oracle, data, [(0, data.size)], 16384, 65536, 262144)
    assert got["sizes"].tolist() == GOLDEN_CHUNKS


@pytest.mark.parametrize("case", GOLDEN["chunker"], ids=lambda c: "%s-%d-t%d" % (c["kind"], c["n"], c["target"]))
def test_chunker_fixtures(ctx, oracle, case):
    data = synth_bytes(case["seed"], case["n"], case["kind"])
    mn, av, mx = chunker_params(case["target"])
    got = check_ranges(ctx, oracle, data, [(0, data.size)], mn, av, mx)
    assert got["chunk_count"] == case["count"]
    assert sha(got["sizes"].astype("<u4").tobytes()) == case["sha256"]


def test_many_ragged_ranges(ctx, oracle):
    """ranges of every awkward size, including empty ones, packed at 16-byte aligned offsets"""
    sizes = [0, 1, 47, 48, 49, 50, 300, 4095, 4096, 4097, 65535, 65536, 65537, 65536 * 2 - 1, 65536 * 2, 65536 * 2 + 1, 200001, 0, 17]
    data = synth_bytes(77, 2 << 20)
    ranges, off = [], 0
    for s in sizes:
        ranges.append((off, s, 0))
        off = (off + s + 15) & ~15
    assert off <= data.size
    for target in (64, 512, 4096):
        check_ranges(ctx, oracle, data, ranges, *chunker_params(target))


@pytest.mark.parametrize("kind", ["zero", "p1", "p3", "p48", "p49", "bit"])
def test_degenerate_content(ctx, oracle, kind):
    data = synth_bytes(5, 700001, kind)
    for target in (16, 384, 8192):
        check_ranges(ctx, oracle, data, [(0, data.size), (16 * 1000, 300000)], *chunker_params(target))


def test_dense_candidates_force_overflow_path(ctx, oracle):
    """tiny discriminator: nearly every tile overflows its slot list and the walker's exact fallback runs"""
    data = synth_bytes(6, 1 << 20)
    check_ranges(ctx, oracle, data, [(0, data.size)], 48, 48, 4096)   # d = 36, cuts far apart relative to candidates
    check_ranges(ctx, oracle, data, [(0, data.size)], 4096, 4096, 65536)


@pytest.mark.parametrize("n", [0, 1, 2, 63, 64, 65, 127, 128, 1023, 1024, 1025, 2047, 2048, 2049, 3072, 4097, 65536, 100000, 131072, (1 << 20) + 3])
def test_hash_segments_sizes(ctx, oracle, n):
    data = synth_bytes(100 + n, n + 64)
    dev = DeviceBytes(ctx, data)
    try:
        offs = np.array([0, 1, 3, 16, 31], np.uint64)
        got = ctx.hash_segments(dev.ptr, dev.size, offs, np.full(offs.size, n, np.uint32))
    finally:
        dev.free()
    exp = oracle.hash_segments(ol.HASH_BLAKE3, data, offs, np.full(offs.size, n, np.uint32))
    assert got.tolist() == exp.tolist()
    case = [c for c in GOLDEN["hash"] if c["n"] == n]
    if case:
        assert "%016x" % got[0] == case[0]["blk3"]


@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 127, 128, 129, 1023, 1024, 4097, 65536, 100000, (1 << 20) + 3])
def test_blake2s_segments_sizes(ctx, oracle, n):
    """BLAKE2s with digest length 8 (lib/blake2/longtail_blake2.c:95-112), every block-boundary case, unaligned starts"""
    import longtail_b200
    data = synth_bytes(100 + n, n + 64)
    dev = DeviceBytes(ctx, data)
    try:
        offs = np.array([0, 1, 3, 16, 31, 7, 0], np.uint64)
        lens = np.array([n, n, n, n, n, max(n - 5, 0), min(n, 64)], np.uint32)
        got = ctx.hash_segments(dev.ptr, dev.size, offs, lens, hash_type=longtail_b200.HASH_BLAKE2)
    finally:
        dev.free()
    exp = oracle.hash_segments(ol.HASH_BLAKE2, data, offs, lens)
    assert got.tolist() == exp.tolist()
    case = [c for c in GOLDEN["hash"] if c["n"] == n]
    if case:
        assert "%016x" % got[0] == case[0]["blk2"]


def test_blake2s_many_ragged_segments(ctx, oracle):
    """thousands of segments of very different lengths: the per-lane work queue must hand out every one exactly once"""
    import longtail_b200
    data = synth_bytes(9, 4 << 20)
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 40000, 5000).astype(np.uint32)
    lens[::97] = 0
    offs = rng.integers(0, data.size - 40000, 5000).astype(np.uint64)
    dev = DeviceBytes(ctx, data)
    try:
        got = ctx.hash_segments(dev.ptr, dev.size, offs, lens, hash_type=longtail_b200.HASH_BLAKE2)
    finally:
        dev.free()
    assert got.tolist() == oracle.hash_segments(ol.HASH_BLAKE2, data, offs, lens).tolist()


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 255, 256, 257, 287, 288, 511, 512, 513, 1023, 4097, 65536, 100000, (1 << 20) + 19])
def test_meow_segments_sizes(ctx, oracle, n):
    """Meow 0.5 low 64 bits (lib/meowhash/longtail_meowhash.c:43-50): every residual / lane-count / 16-byte case of MeowEnd
    (meow_hash_x64_aesni.h:583-700), unaligned starts"""
    import longtail_b200
    data = synth_bytes(300 + n, n + 64)
    dev = DeviceBytes(ctx, data)
    try:
        offs = np.array([0, 1, 3, 16, 31, 7, 0, 13], np.uint64)
        lens = np.array([n, n, n, n, n, max(n - 5, 0), min(n, 64), max(n - 17, 0)], np.uint32)
        got = ctx.hash_segments(dev.ptr, dev.size, offs, lens, hash_type=longtail_b200.HASH_MEOW)
    finally:
        dev.free()
    exp = oracle.hash_segments(ol.HASH_MEOW, data, offs, lens)
    assert got.tolist() == exp.tolist()


def test_meow_many_ragged_segments(ctx, oracle):
    """thousands of segments of very different lengths (work queue + deferred finalisation), all lengths 0..600 included"""
    import longtail_b200
    data = synth_bytes(19, 4 << 20)
    rng = np.random.default_rng(5)
    lens = rng.integers(0, 40000, 6000).astype(np.uint32)
    lens[:601] = np.arange(601)
    lens[601::97] = 0
    offs = rng.integers(0, data.size - 40000, 6000).astype(np.uint64)
    dev = DeviceBytes(ctx, data)
    try:
        got = ctx.hash_segments(dev.ptr, dev.size, offs, lens, hash_type=longtail_b200.HASH_MEOW)
    finally:
        dev.free()
    assert got.tolist() == oracle.hash_segments(ol.HASH_MEOW, data, offs, lens).tolist()


def test_version_index_meow(ctx, oracle):
    """whole CreateVersionIndex with the 'meow' identifier == the oracle (which is pinned to the reference in test_oracle.py)"""
    import longtail_b200
    target = 4096
    assets = [("a/one.bin", synth_bytes(1, 3 * target * 1024 + 777)), ("a/two.bin", synth_bytes(2, 500000, "nib")),
              ("b/", synth_bytes(3, 0)), ("b/dup.bin", synth_bytes(2, 500000, "nib")), ("c.txt", synth_bytes(4, 123456, "text")), ("e", synth_bytes(5, 0))]
    tags = [0, ol.COMP_LZ4, 0, ol.COMP_LZ4, 0, 0]
    al = longtail_b200.AssetList([p for p, _ in assets], [d.size for _, d in assets])
    v = ctx.index_host_assets(al, [d for _, d in assets], tags, hash_type=longtail_b200.HASH_MEOW, target_chunk_size=target)
    assert v == oracle.create_version_index(assets, target, hash_type=ol.HASH_MEOW, tags=tags)


def test_hash_kats_both_algorithms(ctx):
    import longtail_b200
    s = np.frombuffer(b"This is the first test string which is fairly long and should - reconstructed properly, than you very much\0", dtype=np.uint8)
    dev = DeviceBytes(ctx, s)
    try:
        assert ctx.hash_segments(dev.ptr, dev.size, [0], [s.size], hash_type=longtail_b200.HASH_BLAKE2)[0] == 0xD336E5AFA4FA1F4D  # test.cpp:460
        assert ctx.hash_segments(dev.ptr, dev.size, [0], [s.size])[0] == 0xD38BBE79F1F03FDA  # test.cpp:472
        assert ctx.hash_segments(dev.ptr, dev.size, [0], [s.size], hash_type=longtail_b200.HASH_MEOW)[0] == 0x4EDC68DAC105C4EE  # test.cpp:484
    finally:
        dev.free()


@pytest.mark.parametrize("case", [c for c in GOLDEN["version_index"] if c["hash"] == "blk2"], ids=lambda c: "t%d" % c["target"])
def test_version_index_blake2(ctx, oracle, case):
    import longtail_b200
    target = case["target"]
    assets, tags, perms = _tree(target)
    al = longtail_b200.AssetList([p for p, _ in assets], [d.size for _, d in assets], perms)
    v = ctx.index_host_assets(al, [d for _, d in assets], tags, hash_type=longtail_b200.HASH_BLAKE2, target_chunk_size=target)
    assert len(v) == case["size"]
    assert sha(v) == case["sha256"]
    assert v == oracle.create_version_index(assets, target, hash_type=ol.HASH_BLAKE2, tags=tags, perms=perms)


def test_hash_kat(ctx):
    s = np.frombuffer(b"This is the first test string which is fairly long and should - reconstructed properly, than you very much\0", dtype=np.uint8)
    dev = DeviceBytes(ctx, s)
    try:
        got = ctx.hash_segments(dev.ptr, dev.size, [0], [s.size])
    finally:
        dev.free()
    assert got[0] == 0xD38BBE79F1F03FDA  # reference test/test.cpp:472


def _tree(target):
    assets = small_tree(target)
    tags = [ol.COMP_LZ4 if i % 3 else 0 for i in range(len(assets))]
    perms = [0o644 + i for i in range(len(assets))]
    return assets, tags, perms


@pytest.mark.parametrize("case", [c for c in GOLDEN["version_index"] if c["hash"] == "blk3"], ids=lambda c: "t%d" % c["target"])
def test_version_index_device_and_host_paths(ctx, oracle, case):
    import longtail_b200
    target = case["target"]
    assets, tags, perms = _tree(target)
    al = longtail_b200.AssetList([p for p, _ in assets], [d.size for _, d in assets], perms)
    # device-resident arena
    offs, off = [], 0
    for _, d in assets:
        offs.append(off)
        off = (off + d.size + 255) & ~255
    arena = np.zeros(off + 256, np.uint8)
    for o, (_, d) in zip(offs, assets):
        arena[o:o + d.size] = d
    dev = DeviceBytes(ctx, arena)
    try:
        v_dev = ctx.index_device_assets(dev.ptr, dev.size, al, offs, tags, target_chunk_size=target)
    finally:
        dev.free()
    assert len(v_dev) == case["size"]
    assert sha(v_dev) == case["sha256"]
    assert v_dev == oracle.create_version_index(assets, target, tags=tags, perms=perms)
    # host buffers through the copy pipeline
    v_host = ctx.index_host_assets(al, [d for _, d in assets], tags, target_chunk_size=target)
    assert v_host == v_dev


def test_build_version_index_from_host_table(ctx, oracle):
    """the multi-GPU merge entry point: a chunk table that lives on the host (as after an allgather)"""
    import longtail_b200
    target = 256
    assets, tags, perms = _tree(target)
    mn, av, mx = chunker_params(target)
    part = target * 1024
    hashes, sizes, ctags, counts = [], [], [], []
    for (path, d), tag in zip(assets, tags):
        n = 0
        for p in range(1 + d.size // part):
            chunk = d[p * part:(p + 1) * part]
            if not chunk.size:
                continue
            lens = oracle.chunk(chunk, mn, av, mx)
            o = 0
            for ln in lens:
                hashes.append(oracle.hash(ol.HASH_BLAKE3, chunk[o:o + int(ln)]))
                sizes.append(int(ln))
                ctags.append(tag)
                o += int(ln)
            n += lens.size
        counts.append(n)
    al = longtail_b200.AssetList([p for p, _ in assets], [d.size for _, d in assets], perms)
    v = ctx.build_version_index(al, counts, np.array(hashes, np.uint64), np.array(sizes, np.uint32), np.array(ctags, np.uint32), target_chunk_size=target)
    assert v == oracle.create_version_index(assets, target, tags=tags, perms=perms)


def test_empty_inputs(ctx, oracle):
    import longtail_b200
    al = longtail_b200.AssetList([], [], [])
    v = ctx.index_host_assets(al, [], None, target_chunk_size=32768)
    assert v == oracle.create_version_index([], 32768)
    assets = [("a/", np.zeros(0, np.uint8)), ("a/e.bin", np.zeros(0, np.uint8))]
    al = longtail_b200.AssetList([p for p, _ in assets], [0, 0])
    v = ctx.index_host_assets(al, [d for _, d in assets], None, target_chunk_size=32768)
    assert v == oracle.create_version_index(assets, 32768)


def test_large_random_against_reference(ctx, oracle, reference):
    """256 MiB generated in HBM, default CLI target (32768) and the config-2 target (65536): every boundary and hash"""
    n = 256 << 20
    ptr = ctx.device_alloc(n + 64)
    try:
        ctx.synth_fill(ptr, n, seed=1)
        ctx.synchronize()
        host = ctx.to_host(ptr, n)
        for target in (32768, 65536):
            mn, av, mx = chunker_params(target)
            part = target * 1024
            ranges = [(o, min(part, n - o), 0) for o in range(0, n, part)]
            got = ctx.chunk_ranges(ptr, n, ranges, mn, av, mx)
            checker = reference if reference is not None else oracle
            exp_sizes = np.concatenate([checker.chunk(host[o:o + s], mn, av, mx) for o, s, _ in ranges])
            assert got["sizes"].tolist() == exp_sizes.tolist()
            offs = np.concatenate([[0], np.cumsum(exp_sizes.astype(np.uint64))[:-1]]).astype(np.uint64)
            exp_hashes = checker.hash_segments(ol.HASH_BLAKE3, host, offs, exp_sizes)
            assert got["hashes"].tolist() == exp_hashes.tolist()
    finally:
        ctx.device_free(ptr)


def test_synth_matches_host_generator(ctx):
    """the device generator and include/lt_synth.h's host loop produce the same bytes (the CPU baseline depends on it)"""
    import ctypes as C
    import subprocess
    import tempfile
    root = os.path.dirname(HERE)
    src = '#include "lt_synth.h"\nvoid fill(const struct lt_synth_spec* s, uint64_t a, uint64_t o, uint8_t* d, uint64_t n){lt_synth_fill(s,a,o,d,n);}\n'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(root, "include"), "-o", os.path.join(td, "s.so"), os.path.join(td, "s.c")], check=True)
        lib = C.CDLL(os.path.join(td, "s.so"))
        import longtail_b200
        n = (3 << 20) + 123
        for mode, perm in ((0, 0), (1, 500)):
            spec = longtail_b200.SynthSpec(9, perm, 4, mode, 0)
            host = np.zeros(n, np.uint8)
            lib.fill(C.byref(spec), C.c_uint64(7), C.c_uint64(1 << 20), host.ctypes.data_as(C.c_void_p), C.c_uint64(n))
            ptr = ctx.device_alloc(n + 16)
            try:
                ctx.synth_fill(ptr, n, seed=9, asset_id=7, offset=1 << 20, shared_permille=perm, pool_segments=4, class_mode=mode)
                ctx.synchronize()
                dev = ctx.to_host(ptr, n)
            finally:
                ctx.device_free(ptr)
            assert (dev == host).all()


def test_blake3_tree_shapes_many_small_segments(ctx, oracle):
    """segments of 1..40 leaves (every tree shape incl. 3, 5, 6, 7 leaves) in bulk: the per-level work lists must cover every parent"""
    data = synth_bytes(21, 8 << 20)
    rng = np.random.default_rng(5)
    leaves = np.concatenate([np.full(4000, 3), rng.integers(1, 41, 4000)])
    lens = (leaves * 1024 - rng.integers(0, 1024, leaves.size)).astype(np.uint32)
    offs = rng.integers(0, data.size - 41 * 1024, leaves.size).astype(np.uint64)
    dev = DeviceBytes(ctx, data)
    try:
        got = ctx.hash_segments(dev.ptr, dev.size, offs, lens)
    finally:
        dev.free()
    assert got.tolist() == oracle.hash_segments(ol.HASH_BLAKE3, data, offs, lens).tolist()


@pytest.mark.parametrize("target", [64, 2048])
def test_many_ragged_parts_in_one_batch(ctx, oracle, target):
    """hundreds of assets whose parts end in the middle of a scan tile, all in ONE batch: candidate bits that a ragged tail leaves
    beyond the end of a part (zero-filled staging) must not leak into the next tile the same warp scans (regression: they did,
    found by bench.py's stored-bytes check against the reference on the configs[2] sample)"""
    import longtail_b200
    part = target * 1024
    sizes = [(int(x) % (3 * part + 5000)) + 1 for x in (np.arange(1, 401, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(1 << 31)]
    assets = [("r/%04d.bin" % i, synth_bytes(3000 + i, n, "rand" if i % 3 else "nib")) for i, n in enumerate(sizes)]
    al = longtail_b200.AssetList([p for p, _ in assets], [d.size for _, d in assets])
    v = ctx.index_host_assets(al, [d for _, d in assets], None, target_chunk_size=target)
    assert v == oracle.create_version_index(assets, target)
