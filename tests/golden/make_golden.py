"""Generates tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/libref_shim.so,
built by `make -C oracle ref` in a container that has /root/reference).  Commit the output.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
from synth import chunker_params, small_tree, synth_bytes  # noqa: E402

CHUNK_CASES = [(1, 8 << 20, 65536, "rand"), (2, 4 << 20, 32768, "bit"), (3, 4194311, 256, "rand"), (4, 1 << 20, 16, "rand"),
               (5, 300000, 2048, "zero"), (6, 300000, 2048, "p3"), (7, 300000, 512, "p49"), (8, 300000, 512, "p48"),
               (9, 3 << 20, 4096, "text"), (10, 2 << 20, 1024, "nib")]
HASH_SIZES = [0, 1, 63, 64, 65, 1023, 1024, 1025, 2048, 2049, 3072, 5000, 65536, 100000, 131072, 1 << 20]
LZ4_CASES = [(0, "rand"), (5, "rand"), (12, "zero"), (13, "zero"), (41, "text"), (40000, "text"), (65546, "text"), (65547, "text"),
             (200000, "rand"), (300000, "zero"), (500000, "nib"), (720000, "text"), (1 << 20, "p3"), (3 << 20, "nib")]


# ZStd level 3 ('ztd2'): every parameter row of clevels.h (<=16 KiB, <=128 KiB, <=256 KiB, above), block-size edges, inputs larger
# than the 2 MiB window, incompressible / RLE / tiny inputs
ZSTD_CASES = [(0, "rand"), (1, "rand"), (6, "zero"), (7, "zero"), (40, "text"), (63, "rec"), (64, "rec"), (300, "rec"), (1024, "rec"), (1025, "text"),
              (16384, "rec"), (16385, "rec"), (65536, "nib"), (131072, "rec"), (131073, "rec"), (200000, "zero"), (262144, "rec"),
              (262145, "rec"), (300000, "p7"), (500000, "text"), (1 << 20, "rand"), (1 << 20, "rec"), (2500000, "nib"), (7 << 20, "rec")]


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def main():
    r = ol.Reference()
    assert r.available, "build the reference first: make -C oracle ref"
    g = {"chunker": [], "hash": [], "lz4": [], "zstd": [], "version_index": [], "upsync": []}
    for n, kind in ZSTD_CASES:
        c = r.compress(ol.COMP_ZSTD_DEFAULT, synth_bytes(500 + n, n, kind))
        g["zstd"].append({"n": n, "kind": kind, "size": len(c), "sha256": sha(c)})
    for seed, n, target, kind in CHUNK_CASES:
        mn, av, mx = chunker_params(target)
        lens = r.chunk(synth_bytes(seed, n, kind), mn, av, mx)
        g["chunker"].append({"seed": seed, "n": n, "target": target, "kind": kind, "count": int(lens.size),
                             "first": lens[:8].tolist(), "sha256": sha(lens.astype("<u4").tobytes())})
    for n in HASH_SIZES:
        x = synth_bytes(100 + n, n)
        g["hash"].append({"n": n, "blk3": "%016x" % r.hash(ol.HASH_BLAKE3, x), "blk2": "%016x" % r.hash(ol.HASH_BLAKE2, x),
                          "meow": "%016x" % r.hash(ol.HASH_MEOW, x)})
    for n, kind in LZ4_CASES:
        c = r.compress(ol.COMP_LZ4, synth_bytes(200 + n, n, kind))
        g["lz4"].append({"n": n, "kind": kind, "size": len(c), "sha256": sha(c)})
    for target in (16, 256):
        assets = small_tree(target)
        tags = [ol.COMP_LZ4 if i % 3 else 0 for i in range(len(assets))]
        perms = [0o644 + i for i in range(len(assets))]
        for name, ht in (("blk3", ol.HASH_BLAKE3), ("blk2", ol.HASH_BLAKE2)):
            v = r.create_version_index(assets, target, hash_type=ht, tags=tags, perms=perms, workers=3)
            g["version_index"].append({"target": target, "hash": name, "size": len(v), "sha256": sha(v)})
        blocks, v = r.upsync(assets, target, max_block_size=65536, max_chunks_per_block=64, tags=tags, perms=perms, workers=3)
        g["upsync"].append({"target": target, "max_block_size": 65536, "max_chunks_per_block": 64, "blocks": len(blocks),
                            "version_sha256": sha(v),
                            "blocks_sha256": sha(b"".join(h.to_bytes(8, "little") + b for h, b in blocks))})
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote golden.json:", {k: len(v) for k, v in g.items()})


if __name__ == "__main__":
    main()
