"""CPU suite: the ZStd frame DECODER of longtail_b200/csrc/zstd_dec.cu, compiled for the host with one lane (tests/host/zstd_dec_host.cpp),
against frames written by the unmodified reference at every level longtail maps ('ztd1'/'ztd2' = 3, 'ztd4' = 8, 'ztd3'/'ztd5' = 22,
lib/zstd/longtail_zstd.c:43-62) and by the oracle's level-3 restatement.  The same source runs as a warp in tests/test_gpu_zstd_dec.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from synth import synth_bytes

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BAD = 0xFFFFFFFF


@pytest.fixture(scope="module")
def dec(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("zd") / "zd_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "host", "zstd_dec_host.cpp")], check=True)
    lib = C.CDLL(so)
    lib.zd_host_decode.restype = C.c_uint32
    lib.zd_host_worker_bytes.restype = C.c_uint32
    worker = np.zeros(lib.zd_host_worker_bytes() + 64, dtype=np.uint8)

    def decode(frame, cap):
        n = len(frame)
        src = np.zeros(n + 32, dtype=np.uint8)  # the decoder reads whole aligned words: slack behind the frame
        src[:n] = np.frombuffer(frame, dtype=np.uint8)
        dst = np.zeros(cap + 32, dtype=np.uint8)
        got = lib.zd_host_decode(worker.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), C.c_uint32(n),
                                 dst.ctypes.data_as(C.c_void_p), C.c_uint32(cap))
        return got, dst
    return decode


KINDS = ["rec", "text", "nib", "rand", "zero", "p7", "bit"]
SIZES = [0, 1, 2, 7, 63, 64, 255, 256, 257, 1023, 1025, 4096, 16385, 65536, 131071, 131072, 131073, 262145, 700001, (1 << 20) + 3]


@pytest.mark.parametrize("kind", KINDS)
def test_decodes_oracle_level3_frames(dec, oracle, kind):
    for n in SIZES:
        x = synth_bytes(900 + n, n, kind)
        f = oracle.zstd_compress(x)
        got, out = dec(f, n)
        assert got == n, (kind, n, got)
        assert out[:n].tobytes() == x.tobytes(), (kind, n)


@pytest.mark.parametrize("comp", [ol.COMP_ZSTD_DEFAULT, 0x7A746434, 0x7A746433])  # 'ztd2' (3), 'ztd4' (8), 'ztd3' (22)
def test_decodes_reference_frames_all_levels(dec, reference, comp):
    if reference is None:
        pytest.skip("oracle/_ref/libref_shim.so not built")
    for kind in KINDS:
        for n in [0, 5, 300, 5000, 70000, 131072, 300000, (1 << 20) + 17, 3 << 20]:
            x = synth_bytes(950 + n, n, kind)
            f = reference.compress(comp, x)
            got, out = dec(f, n)
            assert got == n, (hex(comp), kind, n, got)
            assert out[:n].tobytes() == x.tobytes(), (hex(comp), kind, n)


def test_capacity_and_malformed(dec, oracle):
    x = synth_bytes(77, 200000, "rec")
    f = oracle.zstd_compress(x)
    assert dec(f, len(x) - 1)[0] == BAD          # declared content size exceeds the capacity
    assert dec(f, len(x) + 100)[0] == len(x)     # a larger buffer is fine
    assert dec(f[:-1], len(x))[0] == BAD         # truncated
    assert dec(f[:len(f) // 2], len(x))[0] == BAD
    assert dec(b"\x00" * 20, 100)[0] == BAD      # no magic
    assert dec(b"", 100)[0] == BAD
    # every single-byte corruption either decodes to something of the declared size or is rejected; it never over-runs the buffers
    bad = 0
    for i in range(4, min(len(f), 400), 3):
        g = bytearray(f)
        g[i] ^= 0x5A
        got, out = dec(bytes(g), len(x))
        assert got == BAD or got == len(x)
        assert not out[len(x):].any()
        bad += got == BAD
    assert bad > 0
