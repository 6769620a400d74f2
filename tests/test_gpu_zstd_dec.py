"""GPU parity (-m gpu) of the ZStd frame decoder (k_zstd_decode, one warp per frame): frames of the reference at every level longtail maps
and frames of our own device encoder decode to the original bytes; malformed frames are rejected.  The serial logic of the same source is
covered on the CPU by tests/test_zstd_dec_host.py."""
import errno

import numpy as np
import pytest

import oracle_lib as ol
from synth import synth_bytes

pytestmark = pytest.mark.gpu

KINDS = ["rec", "text", "nib", "rand", "zero", "p7", "bit"]


@pytest.fixture(scope="module")
def ctx():
    import longtail_b200
    c = longtail_b200.Context(0)
    yield c
    c.close()


def test_decodes_oracle_frames_one_launch(ctx, oracle):
    sizes = [0, 1, 2, 7, 63, 64, 255, 256, 257, 1023, 1025, 4096, 16385, 65536, 131071, 131072, 131073, 262145, 700001, (1 << 20) + 3]
    bufs = [synth_bytes(900 + n, n, k) for k in KINDS for n in sizes]
    frames = [oracle.zstd_compress(b) for b in bufs]
    outs = ctx.zstd_decompress_host(frames, [b.size for b in bufs])
    for b, o in zip(bufs, outs):
        assert o == b.tobytes(), b.size


@pytest.mark.parametrize("comp", [ol.COMP_ZSTD_DEFAULT, 0x7A746434, 0x7A746433])  # 'ztd2' (3), 'ztd4' (8), 'ztd3' (22)
def test_decodes_reference_frames_all_levels(ctx, reference, comp):
    if reference is None:
        pytest.skip("oracle/_ref/libref_shim.so not built")
    bufs = [synth_bytes(950 + n, n, k) for k in KINDS for n in [0, 5, 300, 5000, 70000, 131072, 300000, (1 << 20) + 17, 3 << 20]]
    frames = [reference.compress(comp, b) for b in bufs]
    outs = ctx.zstd_decompress_host(frames, [b.size for b in bufs])
    for b, f, o in zip(bufs, frames, outs):
        assert o == b.tobytes(), (hex(comp), b.size)
        assert reference.decompress(comp, f, b.size) == o


def test_round_trip_of_device_encoder_at_block_size(ctx):
    """stored-block sized payloads through k_zstd_frames then k_zstd_decode"""
    bufs = [synth_bytes(800, 9 << 20, "rec"), synth_bytes(801, (8 << 20) + 12345, "nib"), synth_bytes(802, 5 << 20, "text"),
            synth_bytes(805, 8 << 20, "rand")]
    frames = ctx.zstd_compress_host(bufs)
    outs = ctx.zstd_decompress_host(frames, [b.size for b in bufs])
    for b, o in zip(bufs, outs):
        assert o == b.tobytes()
    # oversized capacity is fine (the frame header carries the content size)
    outs = ctx.zstd_decompress_host(frames[:1], [bufs[0].size + 4096])
    assert outs[0] == bufs[0].tobytes()


def test_many_frames_more_than_resident_warps(ctx, oracle):
    """the frame queue: 6 000 small frames > 148 * 16 resident warps"""
    bufs = [synth_bytes(3000 + i, 3000 + 7 * (i % 900), KINDS[i % len(KINDS)]) for i in range(6000)]
    frames = ctx.zstd_compress_host(bufs)
    outs = ctx.zstd_decompress_host(frames, [b.size for b in bufs])
    assert all(o == b.tobytes() for b, o in zip(bufs, outs))


def test_malformed_frames_are_rejected(ctx, oracle):
    import longtail_b200
    x = synth_bytes(77, 200000, "rec")
    f = oracle.zstd_compress(x)
    for bad, cap in [(f, x.size - 1), (f[:-1], x.size), (f[:len(f) // 2], x.size), (b"\x00" * 20, 100)]:
        with pytest.raises(longtail_b200.LongtailB200Error) as e:
            ctx.zstd_decompress_host([bad], [cap])
        assert e.value.errno == errno.EBADF
    # single-byte corruptions: rejected or decoded to the declared size, never a fault (the context stays usable)
    for i in range(4, 300, 7):
        g = bytearray(f)
        g[i] ^= 0x5A
        try:
            out = ctx.zstd_decompress_host([bytes(g)], [x.size])
            assert len(out[0]) == x.size
        except longtail_b200.LongtailB200Error as e:
            assert e.errno == errno.EBADF
    assert ctx.zstd_decompress_host([f], [x.size])[0] == x.tobytes()


def test_lz4_decompress_binding(ctx, oracle):
    bufs = [synth_bytes(60 + i, n, k) for i, (n, k) in enumerate([(0, "rand"), (13, "text"), (70000, "rec"), (1 << 20, "nib")])]
    comp = ctx.lz4_compress_host(bufs)
    outs = ctx.lz4_decompress_host(comp, [b.size for b in bufs])
    assert all(o == b.tobytes() for b, o in zip(bufs, outs))
