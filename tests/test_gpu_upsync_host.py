"""GPU parity (-m gpu): lt_b200_upsync_host_assets (cmd/main.c:UpSync for assets in host memory: H2D once, CreateVersionIndex,
CreateMissingContent, WriteContent) and the device sink of lt_b200_write_blocks_device_ex against the reference's upsync."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from synth import synth_bytes

pytestmark = pytest.mark.gpu


def _assets():
    kinds = ["rand", "nib", "text", "rec", "zero"]
    assets = [("u/%02d.bin" % i, synth_bytes(900 + 11 * i, 70000 + 830000 * i, kinds[i % 5])) for i in range(9)]
    assets.append(("u/copy.bin", assets[4][1].copy()))
    assets.append(("u/empty.bin", synth_bytes(1, 0)))
    return assets


def _collect():
    blocks = []

    def sink(_user, view):
        b = view.contents
        blocks.append((int(b.block_hash), C.string_at(b.data, b.size)))
        return 0

    return blocks, sink


@pytest.mark.parametrize("tag", [ol.COMP_LZ4, ol.COMP_ZSTD_DEFAULT])
def test_upsync_host_assets_fresh_and_incremental(oracle, reference, tag):
    import longtail_b200
    ctx = longtail_b200.Context(0)
    try:
        assets = _assets()
        tags = [tag if i % 4 else 0 for i in range(len(assets))]
        al = longtail_b200.AssetList([p for p, _ in assets], [d.size for _, d in assets])
        checker = reference if reference is not None else oracle
        want_blocks, want_v = checker.upsync(assets, 32768, tags=tags)
        blocks, sink = _collect()
        cb = longtail_b200.BLOCK_SINK(sink)
        v, written = ctx.upsync_host_assets(al, [d for _, d in assets], tags, (C.cast(cb, C.c_void_p), None), target_chunk_size=32768)
        assert v == want_v
        assert blocks == want_blocks
        vi = longtail_b200.parse_version_index(v)
        assert written == vi["chunk_count"]
        if reference is not None:
            # a store that already holds every third chunk: only the missing ones are packed and written (CreateMissingContent)
            have = vi["chunk_hashes"][::3].copy()
            want2, _ = reference.upsync(assets, 32768, tags=tags, existing_hashes=have)
            blocks2, sink2 = _collect()
            cb2 = longtail_b200.BLOCK_SINK(sink2)
            v2, written2 = ctx.upsync_host_assets(al, [d for _, d in assets], tags, (C.cast(cb2, C.c_void_p), None), target_chunk_size=32768,
                                                  existing_hashes=have)
            assert v2 == want_v and written2 == vi["chunk_count"] - have.size
            assert blocks2 == want2
    finally:
        ctx.close()


@pytest.mark.parametrize("target,block,per_block,batch,pinned", [(32768, 8388608, 1024, 0, False), (4096, 262144, 64, 4 << 20, False),
                                                                 (4096, 100000, 7, 8 << 20, True), (8192, 1 << 20, 1024, 8 << 20, True)])
def test_upsync_stream_host_assets(oracle, reference, target, block, per_block, batch, pinned):
    """the streaming pass (batches of `batch` bytes, open block carried from arena to arena) writes the reference's blocks in its order;
    pinned: the assets sit in pinned host memory (every other one 16-byte aligned: read by the upload kernel, the rest by cudaMemcpyAsync)"""
    import longtail_b200
    ctx = longtail_b200.Context(0)
    host = None
    try:
        assets = _assets()
        if pinned:
            host = ctx.pinned_alloc(sum(d.size + 64 for _, d in assets) + 64)
            at, placed = 0, []
            for i, (p, d) in enumerate(assets):
                at = (at + 15) & ~15
                if i % 2:
                    at += 3
                host[at:at + d.size] = d
                placed.append((p, host[at:at + d.size]))
                at += d.size
            assets = placed
        tags = [(ol.COMP_LZ4 if i % 4 else 0) if i % 5 else ol.COMP_ZSTD_DEFAULT for i in range(len(assets))]
        al = longtail_b200.AssetList([p for p, _ in assets], [d.size for _, d in assets])
        checker = reference if reference is not None else oracle
        want_blocks, want_v = checker.upsync(assets, target, max_block_size=block, max_chunks_per_block=per_block, tags=tags)
        blocks, sink = _collect()
        cb = longtail_b200.BLOCK_SINK(sink)
        v, written = ctx.upsync_stream_host_assets(al, [d for _, d in assets], tags, (C.cast(cb, C.c_void_p), None), target_chunk_size=target,
                                                   max_block_size=block, max_chunks_per_block=per_block, batch_bytes=batch)
        assert v == want_v
        assert [h for h, _ in blocks] == [h for h, _ in want_blocks]
        assert blocks == want_blocks
        vi = longtail_b200.parse_version_index(v)
        assert written == vi["chunk_count"]
        if reference is not None:
            have = vi["chunk_hashes"][1::3].copy()
            want2, _ = reference.upsync(assets, target, max_block_size=block, max_chunks_per_block=per_block, tags=tags, existing_hashes=have)
            blocks2, sink2 = _collect()
            cb2 = longtail_b200.BLOCK_SINK(sink2)
            v2, written2 = ctx.upsync_stream_host_assets(al, [d for _, d in assets], tags, (C.cast(cb2, C.c_void_p), None), target_chunk_size=target,
                                                         max_block_size=block, max_chunks_per_block=per_block, batch_bytes=batch, existing_hashes=have)
            assert v2 == want_v and written2 == vi["chunk_count"] - have.size
            assert blocks2 == want2
    finally:
        if host is not None:
            ctx.pinned_free(host)
        ctx.close()


def test_device_sink_leaves_identical_images_in_hbm(oracle, reference):
    """LT_B200_WRITE_DEVICE_SINK: `data` of every view is a device address; copied back, the images are the reference's StoredBlocks"""
    import longtail_b200
    ctx = longtail_b200.Context(0)
    try:
        assets = _assets()
        tags = [ol.COMP_LZ4 if i % 3 else 0 for i in range(len(assets))]
        al = longtail_b200.AssetList([p for p, _ in assets], [d.size for _, d in assets])
        offs, off = [], 0
        for _, d in assets:
            offs.append(off)
            off = (off + d.size + 255) & ~255
        arena = np.zeros(off + 4096, np.uint8)
        for o, (_, d) in zip(offs, assets):
            arena[o:o + d.size] = d
        ptr = ctx.device_alloc(arena.size)
        ctx.to_device(ptr, arena)
        v = ctx.index_device_assets(ptr, arena.size, al, offs, tags, target_chunk_size=32768)
        vi = longtail_b200.parse_version_index(v)
        got = []

        def sink(_user, view):
            b = view.contents
            got.append((int(b.block_hash), ctx.to_host(b.data, b.size).tobytes()))  # a device -> host copy inside the sink: the image lives in HBM
            return 0

        cb = longtail_b200.BLOCK_SINK(sink)
        ctx.write_blocks_device(ptr, arena.size, vi["chunk_hashes"], vi["chunk_sizes"], vi["chunk_tags"], ctx.unique_chunk_offsets(vi["chunk_count"]),
                                c_sink=(C.cast(cb, C.c_void_p), None), device_sink=True)
        ctx.device_free(ptr)
        checker = reference if reference is not None else oracle
        want_blocks, want_v = checker.upsync(assets, 32768, tags=tags)
        assert v == want_v
        assert got == want_blocks
    finally:
        ctx.close()
