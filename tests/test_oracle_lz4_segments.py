"""CPU suite: the model of the speculative-segment LZ4 encoder (oracle/lt_lz4_segments.c, the design of DESIGN.md section 8 item 1).
Whatever the hand-over decisions are, the bytes must equal the sequential encoder's (itself pinned to the reference in test_oracle.py);
the acceptance rates show on which data the design pays."""
import ctypes as C

import numpy as np
import pytest

from synth import synth_bytes


def run(oracle, data, segments, warm):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    cap = oracle.lib.lto_lz4_bound(data.size)
    dst = np.zeros(cap, dtype=np.uint8)
    n = C.c_uint64(0)
    stats = (C.c_uint64 * 5)()
    err = oracle.lib.lto_lz4_compress_segments(data.ctypes.data_as(C.c_void_p), C.c_uint64(data.size), C.c_uint32(segments), C.c_uint64(warm),
                                               dst.ctypes.data_as(C.c_void_p), C.c_uint64(cap), C.byref(n), stats)
    assert err == 0, err
    return dst[:n.value].tobytes(), [int(x) for x in stats]


def mixed(seed, n):
    """1 MiB stretches of random / 4-bit / text-like bytes, like the benchmark's stored blocks"""
    parts = [synth_bytes(seed + i, min(1 << 20, n - i * (1 << 20)), ("rand", "nib", "text")[(seed + i) % 3]) for i in range((n + (1 << 20) - 1) >> 20)]
    return np.concatenate(parts)[:n]


@pytest.fixture(scope="module")
def lib(oracle):
    oracle.lib.lto_lz4_bound.restype = C.c_uint64
    oracle.lib.lto_lz4_bound.argtypes = [C.c_uint64]
    return oracle


@pytest.mark.parametrize("kind", ["text", "nib", "rec", "rand", "mixed", "zero", "p7"])
def test_bytes_identical_whatever_the_handovers(lib, kind):
    n = (6 << 20) + 12345
    data = mixed(7, n) if kind == "mixed" else synth_bytes(11, n, kind)
    want = lib.lz4_compress(data)
    for segments, warm in [(1, 1 << 18), (4, 192 << 10), (8, 64 << 10), (4, 4096), (16, 1 << 20), (3, 0)]:
        got, stats = run(lib, data, segments, warm)
        assert got == want, (kind, segments, warm, stats)
        assert stats[1] + stats[2] + stats[3] <= stats[0]


def test_acceptance_rates_on_the_benchmark_classes(lib, capsys):
    """documents WHERE speculation pays (printed, asserted loosely).  Finding of this model: two parses only resynchronise on data with long
    matches (records: the same long match ends both of them at the same place, after which they insert the same positions); on the
    benchmark's order-0 text and 4-bit classes every probe has many short candidates, the tables hold different positions and the parses
    never meet again, however long the warm-up — there the design would fall back to sequential parsing every time."""
    n = 9 << 20
    rates = {}
    for kind in ["text", "nib", "rec", "mixed", "rand"]:
        data = mixed(3, n) if kind == "mixed" else synth_bytes(5, n, kind)
        for warm in (192 << 10, 1 << 20):
            _, stats = run(lib, data, 8, warm)
            rates[(kind, warm >> 10)] = (stats[1], stats[0], stats[2], stats[3])
    with capsys.disabled():
        print("\nspeculative LZ4 segments, 9 MiB blocks, 8 segments: accepted / attempted (boundary differs, table differs)")
        for (k, w), (a, t, b, tb) in rates.items():
            print("  %-6s warm-up %4d KiB: %d / %d  (%d, %d)" % (k, w, a, t, b, tb))
    assert rates[("rec", 1024)][0] >= 5                      # long matches: the parses meet and stay together
    assert rates[("nib", 1024)][0] == 0 and rates[("text", 1024)][0] <= 1  # short, dense matches: they do not
    assert rates[("rand", 1024)][1] == 0                     # no match, no sequence boundary, nothing to hand over (and nothing to gain)
