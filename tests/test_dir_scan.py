"""lt_b200_scan_directory against the unmodified reference's Longtail_GetFilesRecursively2 over its own file
storage (same entries, order, sizes, permissions), and — on the GPU — CreateVersionIndex over the scanned tree with the threaded reader
against the reference indexing the same directory."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from synth import synth_bytes


def make_tree(root):
    """names chosen to exercise the ordering rule: strcmp over relative names BEFORE directories get their '/' ("a.b" < "a/x" but "a" < "a.b"),
    upper/lower case, empty directories and files, nested empties, files spanning several parts at a small target"""
    files = {"a/x.bin": 300000, "a/y/z.txt": 1234, "a.b": 77, "a-b/q": 5000, "A/upper.bin": 70000, "b": 0, "c/empty.dat": 0,
             "c/d/e/f/deep.bin": 150000, "zz/big.bin": 3 * 1024 * 1024 + 17, "a/x.bin.bak": 300000, "0": 48, "_under/score": 49}
    for i, (rel, n) in enumerate(sorted(files.items())):
        p = os.path.join(root, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        synth_bytes(40 + (0 if rel.endswith(".bak") else i), n, ("rand", "text", "nib", "rec")[i % 4] if not rel.endswith(".bak") else "rand").tofile(p) if n else open(p, "wb").close()
    # the .bak file must duplicate a/x.bin byte for byte (dedup across assets)
    open(os.path.join(root, "a/x.bin.bak"), "wb").write(open(os.path.join(root, "a/x.bin"), "rb").read())
    for d in ("empty_dir", "c/d/also_empty", "a/y/w"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    os.chmod(os.path.join(root, "a/x.bin"), 0o640)
    os.chmod(os.path.join(root, "zz/big.bin"), 0o755)
    os.chmod(os.path.join(root, "empty_dir"), 0o700)
    return files


def test_scan_matches_reference(reference, tmp_path):
    if reference is None:
        pytest.skip("oracle/_ref/libref_shim.so not built")
    import longtail_b200
    root = str(tmp_path / "tree")
    os.makedirs(root)
    make_tree(root)
    want = ol.ref_scan_directory(reference, root, workers=0)
    assert want == ol.ref_scan_directory(reference, root, workers=4)
    for threads in (1, 8):
        fl = longtail_b200.FileList(root, threads=threads)
        got = list(zip(fl.paths, fl.sizes, fl.permissions))
        fl.close()
        assert got == want
    names = [p for p, _, _ in want]
    assert names.index("a/") < names.index("a-b/") < names.index("a.b") < names.index("a/x.bin")  # "a" < "a-b" < "a.b" < "a/x.bin"
    assert "empty_dir/" in names and "c/d/also_empty/" in names and dict((p, s) for p, s, _ in want)["b"] == 0


def test_scan_errors(tmp_path):
    import longtail_b200
    root = tmp_path / "t"
    root.mkdir()
    (root / "f").write_bytes(b"x")
    os.symlink(str(root / "f"), str(root / "link"))
    with pytest.raises(longtail_b200.LongtailB200Error):
        longtail_b200.FileList(str(root))  # the reference's iterator has no name for a symbolic link; rejected instead of guessed
    fl = longtail_b200.FileList(str(tmp_path / "missing"))  # a root that does not exist lists as empty, like ScanFolder's ENOENT branch
    assert fl.paths == []
    fl.close()


@pytest.mark.gpu
@pytest.mark.parametrize("target,readers", [(64, 1), (64, 8), (32768, 4)])
def test_index_directory_matches_reference(reference, tmp_path, target, readers):
    if reference is None:
        pytest.skip("oracle/_ref/libref_shim.so not built")
    import longtail_b200
    root = str(tmp_path / "tree")
    os.makedirs(root)
    make_tree(root)
    ctx = longtail_b200.Context(0)
    fl = longtail_b200.FileList(root)
    tags = [ol.COMP_LZ4] * len(fl.paths)
    got = ctx.index_file_list(fl, tags, target_chunk_size=target, reader_threads=readers)
    fl.close()
    ctx.close()
    assert got == ol.ref_index_directory(reference, root, target, workers=4, tag=ol.COMP_LZ4)


@pytest.mark.gpu
def test_upsync_directory_to_directory_matches_reference(reference, tmp_path):
    """cmd/main.c:UpSync end to end on the device: scan -> load the tree into HBM -> index -> missing chunks -> blocks -> fsblockstore
    directory; VersionIndex, every .lrb and store.lsi identical to the unmodified reference doing the same, then the tree changes and both
    upsync again into their stores (incremental)"""
    if reference is None:
        pytest.skip("oracle/_ref/libref_shim.so not built")
    import longtail_b200
    src = str(tmp_path / "tree")
    os.makedirs(src)
    make_tree(src)
    ours, theirs = str(tmp_path / "ours"), str(tmp_path / "ref")
    kw = dict(max_block_size=262144, max_chunks_per_block=64)
    ctx = longtail_b200.Context(0)

    def tree(root):
        return {os.path.relpath(os.path.join(d, f), root): open(os.path.join(d, f), "rb").read()
                for d, _, files in os.walk(root) for f in files if f != "store.lsi.sync"}

    def both():
        fl = longtail_b200.FileList(src)
        st = longtail_b200.FsStore(ours, writer_threads=4)
        vi, n = ctx.upsync_file_list(fl, st, [ol.COMP_ZSTD_DEFAULT] * len(fl.paths), target_chunk_size=64, reader_threads=4, **kw)
        st.close()
        fl.close()
        want_vi, want_n = ol.ref_upsync_dir_to_dir(reference, src, theirs, 64, workers=0, tag=ol.COMP_ZSTD_DEFAULT, **kw)
        assert vi == want_vi and n == want_n and n > 0
        a, b = tree(ours), tree(theirs)
        assert sorted(a) == sorted(b)
        assert all(a[k] == b[k] for k in a), [k for k in a if a[k] != b[k]][:3]
        return n

    n1 = both()
    synth_bytes(555, 400000, "rec").tofile(os.path.join(src, "a/new.bin"))
    os.remove(os.path.join(src, "a.b"))
    with open(os.path.join(src, "zz/big.bin"), "r+b") as f:  # an edit in the middle of a multi-part file
        f.seek(1500000)
        f.write(b"edited" * 100)
    n2 = both()
    assert n2 < n1
    assert ol.ref_read_store_dir(reference, ours) == ol.ref_read_store_dir(reference, theirs)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("sub", ["longtail_b200/csrc", "longtail_b200/_lib", "tests/golden"])
def test_index_real_directories_of_this_repository(reference, sub):
    """real files (sources, object files, the shared library, fixtures) instead of synthetic bytes: the VersionIndex of the tree equals the
    reference's, default target chunk size of the CLI"""
    if reference is None:
        pytest.skip("oracle/_ref/libref_shim.so not built")
    import longtail_b200
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), sub)
    ctx = longtail_b200.Context(0)
    fl = longtail_b200.FileList(root)
    assert len(fl.paths) >= 3
    got = ctx.index_file_list(fl, None, target_chunk_size=32768, reader_threads=8)
    fl.close()
    ctx.close()
    assert got == ol.ref_index_directory(reference, root, 32768, workers=8, tag=0)
