"""CPU suite: the on-disk block sink (lt_b200_fs_store_*, longtail_b200/csrc/fs_store.cpp — SURVEY.md section 8f row 2) against the unmodified
reference's fsblockstore: same files, same bytes, same store.lsi, readable by the reference, incremental upsync included.  The blocks come
from the reference's capturing sink here (the sink under test is host code); tests/test_gpu_fs_store.py feeds it from lt_b200_write_blocks_device."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from synth import synth_bytes


def tree(root):
    out = {}
    for d, _, files in os.walk(root):
        for f in files:
            if f == "store.lsi.sync":
                continue  # the lock file
            p = os.path.join(d, f)
            out[os.path.relpath(p, root)] = open(p, "rb").read()
    return out


def version(seed, extra=()):
    assets = [("a/one.bin", synth_bytes(seed, 700000, "rec")), ("a/two.bin", synth_bytes(seed + 1, 300000, "nib")),
              ("b/three.txt", synth_bytes(seed + 2, 123456, "text")), ("c.bin", synth_bytes(seed + 3, 400000))]
    return assets + list(extra)


TAGS4 = [ol.COMP_LZ4, ol.COMP_ZSTD_DEFAULT, 0, ol.COMP_LZ4]


@pytest.fixture(scope="module")
def fs():
    import longtail_b200
    return longtail_b200


@pytest.mark.parametrize("writers", [0, 4])
def test_fresh_store_identical_to_reference(fs, oracle, reference, tmp_path, writers):
    if reference is None:
        pytest.skip("oracle/_ref/libref_shim.so not built")
    assets = version(10)
    ours, theirs = str(tmp_path / "ours"), str(tmp_path / "ref")
    # the blocks are produced by the reference with a capturing sink: the producer is not under test here
    blocks, _ = reference.upsync(assets, 16384, max_block_size=262144, max_chunks_per_block=64, tags=TAGS4)
    st = fs.FsStore(ours, writer_threads=writers)
    for h, image in blocks:
        st.put(h, image)
    st.close()
    n_ref = ol.ref_upsync_to_dir(reference, assets, 16384, theirs, max_block_size=262144, max_chunks_per_block=64, tags=TAGS4, workers=0)
    assert n_ref == len(blocks) and len(blocks) > 5
    a, b = tree(ours), tree(theirs)
    assert sorted(a) == sorted(b)
    assert all(k.startswith("chunks/") and k.endswith(".lrb") for k in a if k != "store.lsi")
    for k in a:
        assert a[k] == b[k], k  # every .lrb and, with one reference worker (deterministic block order), store.lsi itself
    assert ol.ref_read_store_dir(reference, ours) == ol.ref_read_store_dir(reference, theirs)


def test_incremental_upsync_and_skip_existing(fs, oracle, reference, tmp_path):
    if reference is None:
        pytest.skip("oracle/_ref/libref_shim.so not built")
    v1 = version(20)
    v2 = version(20, extra=[("d/new.bin", synth_bytes(99, 500000, "rec"))])
    v2[1] = ("a/two.bin", synth_bytes(77, 310000, "nib"))  # one asset replaced, one added
    tags5 = TAGS4 + [ol.COMP_LZ4]
    ours, theirs = str(tmp_path / "ours"), str(tmp_path / "ref")
    kw = dict(max_block_size=262144, max_chunks_per_block=64)
    # version 1 into both
    b1, _ = reference.upsync(v1, 16384, tags=TAGS4, **kw)
    st = fs.FsStore(ours, writer_threads=2)
    for h, image in b1:
        st.put(h, image)
    st.close()
    ol.ref_upsync_to_dir(reference, v1, 16384, theirs, tags=TAGS4, workers=0, **kw)
    # version 2: only the chunks the store lacks (DiffHashes against the chunk hashes of store.lsi), packed the reference's way
    st = fs.FsStore(ours, writer_threads=2)
    existing = st.existing_chunks()
    assert existing.size == sum(int(np.frombuffer(img, dtype=np.uint32, count=1, offset=12)[0]) for _, img in b1)
    b2, _ = reference.upsync(v2, 16384, tags=tags5, existing_hashes=existing, **kw)
    assert 0 < len(b2) < len(b1)
    for h, image in b2:
        st.put(h, image)
    st.put(*b2[0])  # the same block again: accepted once per store object
    st.flush()
    s = st.stats()
    assert s["blocks_written"] == len(b2) and s["blocks_skipped"] == 0
    st.close()
    n_ref = ol.ref_upsync_to_dir(reference, v2, 16384, theirs, tags=tags5, workers=0, **kw)
    assert n_ref == len(b2)
    a, b = tree(ours), tree(theirs)
    assert sorted(a) == sorted(b)
    for k in a:
        assert a[k] == b[k], k  # store.lsi: new blocks in front of the old ones, as Longtail_MergeStoreIndex(added, existing) orders them
    assert ol.ref_read_store_dir(reference, ours) == ol.ref_read_store_dir(reference, theirs)
    # a block whose file already exists is not rewritten (SafeWriteStoredBlock): a fresh store object over the same directory
    st = fs.FsStore(ours, writer_threads=0)
    st.put(*b1[0])
    st.flush()
    assert st.stats() == {"blocks_written": 0, "bytes_written": 0, "blocks_skipped": 1}
    st.close()
    # like the reference (the block index is added even when the file was there, longtail_fsblockstore.c:805-842, and added blocks are
    # merged in front), the re-added block now leads store.lsi; nothing else changed
    lsi = tree(ours)["store.lsi"]
    nb = int(np.frombuffer(lsi, dtype=np.uint32, count=1, offset=8)[0])
    hashes = np.frombuffer(lsi, dtype=np.uint64, count=nb, offset=16)
    old = np.frombuffer(a["store.lsi"], dtype=np.uint64, count=nb, offset=16)
    assert int(hashes[0]) == b1[0][0] and sorted(hashes.tolist()) == sorted(old.tolist()) and len(lsi) == len(a["store.lsi"])
    assert ol.ref_read_store_dir(reference, ours) == ol.ref_read_store_dir(reference, theirs)


def test_rejects_foreign_store_index_and_bad_image(fs, tmp_path):
    root = tmp_path / "bad"
    root.mkdir()
    (root / "store.lsi").write_bytes(b"\x02\x00\x00\x00" + b"\x00" * 12)  # wrong version
    st = fs.FsStore(str(root), writer_threads=0)
    with pytest.raises(fs.LongtailB200Error):
        st.existing_chunks()
    with pytest.raises(fs.LongtailB200Error):
        st.put(1, b"\x00" * 24 + b"\x07" * 8)  # hash in the image differs from the view's
    st.close()  # nothing was added: the foreign index is left alone
    assert (root / "store.lsi").read_bytes()[:4] == b"\x02\x00\x00\x00"


def test_two_stores_one_directory_concurrently(fs, reference, tmp_path):
    """two store objects (as two uploading processes would hold) put overlapping block sets into one directory from two threads and flush:
    every block file exists once, complete; store.lsi lists the union exactly once (flock + merge with the index on disk), and the
    unmodified reference reads the whole store back"""
    if reference is None:
        pytest.skip("oracle/_ref/libref_shim.so not built")
    import threading
    assets = version(30, extra=[("e/more.bin", synth_bytes(31, 900000, "rec"))])
    blocks, _ = reference.upsync(assets, 16384, max_block_size=131072, max_chunks_per_block=32, tags=TAGS4 + [ol.COMP_LZ4])
    assert len(blocks) > 12
    root = str(tmp_path / "shared")
    halves = [blocks[:2 * len(blocks) // 3], blocks[len(blocks) // 3:]]  # the middle third is put by both
    errors = []

    def upload(part):
        try:
            st = fs.FsStore(root, writer_threads=3)
            for h, image in part:
                st.put(h, image)
            st.close()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=upload, args=(p,)) for p in halves]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    files = tree(root)
    lrb = {k: v for k, v in files.items() if k.endswith(".lrb")}
    assert len(lrb) == len(blocks)
    by_hash = {"chunks/%s/0x%016x.lrb" % (("%016x" % h)[:4], h): img for h, img in blocks}
    assert lrb == by_hash
    assert not [k for k in files if k != "store.lsi" and not k.endswith(".lrb")]  # no temporary file left behind
    lsi = files["store.lsi"]
    nb = int(np.frombuffer(lsi, dtype=np.uint32, count=1, offset=8)[0])
    hashes = np.frombuffer(lsi, dtype=np.uint64, count=nb, offset=16)
    assert sorted(hashes.tolist()) == sorted(h for h, _ in blocks)
    got = ol.ref_read_store_dir(reference, root)
    assert got[0] == len(blocks)


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_fs_store_under_sanitizers(tmp_path, sanitizer):
    """fs_store.cpp compiled with ThreadSanitizer / AddressSanitizer + UBSan: two stores, 60 blocks of 3 MiB each, 30 of them put by both,
    flushes in between — no report, and store.lsi lists the union"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "race")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=" + sanitizer, "-o", exe, os.path.join(root, "tests", "host", "fs_store_race.cpp"),
                        os.path.join(root, "longtail_b200", "csrc", "fs_store.cpp"), "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr:
        pytest.skip("this toolchain lacks -fsanitize=" + sanitizer)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe, str(tmp_path / "store")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr[-3000:]
    assert "chunks listed: 450" in r.stdout and "Sanitizer" not in r.stderr
