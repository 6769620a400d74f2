"""ctypes bindings for the two CPU checkers (test infrastructure only).

* ``oracle``  — oracle/_ref/liblt_oracle.so: this repo's C restatement (oracle/lt_oracle.c).
* ``ref``     — oracle/_ref/libref_shim.so: the UNMODIFIED reference compiled from
  /root/reference by ``make -C oracle ref`` (absent when that build was never run).

Both expose the same shapes so tests can run either against the CUDA path.
"""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF_DIR = os.path.join(ROOT, "oracle", "_ref")

HASH_BLAKE3 = 0x626C6B33
HASH_BLAKE2 = 0x626C6B32
HASH_MEOW = 0x6D656F77
COMP_LZ4 = 0x6C7A3432
COMP_ZSTD_DEFAULT = 0x7A746432  # 'ztd2'
COMP_ZSTD_MIN = 0x7A746431  # 'ztd1' (level 0 -> default level 3)

_u8p = C.POINTER(C.c_uint8)


def _ptr(a, t=C.c_uint8):
    return a.ctypes.data_as(C.POINTER(t))


def build_oracle():
    """(re)build the restatement; and the reference shim when /root/reference exists."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)


def _load(name):
    path = os.path.join(_REF_DIR, name)
    if not os.path.exists(path):
        return None
    return C.CDLL(path)


class _AssetArgs:
    """marshals (path, bytes-like) assets into the parallel arrays both shims take"""

    def __init__(self, assets, tags=None, perms=None):
        self.n = len(assets)
        self.paths = (C.c_char_p * max(self.n, 1))(*[p.encode() for p, _ in assets])
        self.bufs = [np.ascontiguousarray(np.frombuffer(d, dtype=np.uint8)) if not isinstance(d, np.ndarray) else np.ascontiguousarray(d)
                     for _, d in assets]
        self.datas = (_u8p * max(self.n, 1))(*[_ptr(b) for b in self.bufs])
        self.sizes = (C.c_uint64 * max(self.n, 1))(*[b.size for b in self.bufs])
        self.tags = None if tags is None else (C.c_uint32 * max(self.n, 1))(*tags)
        self.perms = None if perms is None else (C.c_uint16 * max(self.n, 1))(*perms)


def parse_upsync(buf):
    """-> (list of (block_hash, serialised_block_bytes), version_index_bytes)"""
    (count,) = struct.unpack_from("<I", buf, 0)
    off = 4
    blocks = []
    for _ in range(count):
        h, size = struct.unpack_from("<QQ", buf, off)
        off += 16
        blocks.append((h, bytes(buf[off:off + size])))
        off += size
    (vsize,) = struct.unpack_from("<Q", buf, off)
    off += 8
    return blocks, bytes(buf[off:off + vsize])


class Oracle:
    """liblt_oracle.so"""

    def __init__(self):
        lib = _load("liblt_oracle.so")
        if lib is None:
            build_oracle()
            lib = _load("liblt_oracle.so")
        self.lib = lib
        lib.lto_hpcdc_discriminator.restype = C.c_uint32
        lib.lto_hpcdc_window_hash.restype = C.c_uint32
        lib.lto_blake3_64.restype = C.c_uint64
        lib.lto_blake3_64.argtypes = [C.c_void_p, C.c_uint64]
        lib.lto_blake2s_64.restype = C.c_uint64
        lib.lto_blake2s_64.argtypes = [C.c_void_p, C.c_uint64]
        lib.lto_lz4_bound.restype = C.c_uint64
        lib.lto_lz4_bound.argtypes = [C.c_uint64]
        lib.lto_free.argtypes = [C.c_void_p]
        lib.lto_zstd_bound.restype = C.c_uint64
        lib.lto_zstd_bound.argtypes = [C.c_uint64]
        lib.lto_meow_64.restype = C.c_uint64
        lib.lto_meow_64.argtypes = [C.c_void_p, C.c_uint64]

    def discriminator(self, avg):
        return self.lib.lto_hpcdc_discriminator(C.c_uint32(avg))

    def chunk(self, data, mn, avg, mx):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        cap = data.size // mn + 2
        lens = np.zeros(cap, dtype=np.uint32)
        n = C.c_uint64(0)
        err = self.lib.lto_hpcdc_chunk(_ptr(data), C.c_uint64(data.size), C.c_uint32(mn), C.c_uint32(avg), C.c_uint32(mx),
                                       _ptr(lens, C.c_uint32), C.c_uint64(cap), C.byref(n))
        assert err == 0, err
        return lens[:n.value].copy()

    def hash(self, hash_type, data):
        data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        out = C.c_uint64(0)
        err = self.lib.lto_hash_buffer(C.c_uint32(hash_type), _ptr(data) if data.size else None, C.c_uint64(data.size), C.byref(out))
        assert err == 0, err
        return out.value

    def hash_segments(self, hash_type, base, offsets, lens):
        base = np.ascontiguousarray(base, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        out = np.zeros(offsets.size, dtype=np.uint64)
        err = self.lib.lto_hash_segments(C.c_uint32(hash_type), _ptr(base), C.c_uint64(offsets.size), _ptr(offsets, C.c_uint64),
                                         _ptr(lens, C.c_uint32), _ptr(out, C.c_uint64))
        assert err == 0, err
        return out

    def zstd_compress(self, data):
        """'ztd2' / 'ztd1' (ZStd level 3), one frame"""
        data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        cap = self.lib.lto_zstd_bound(data.size)
        dst = np.zeros(cap + 8, dtype=np.uint8)
        n = C.c_uint64(0)
        dummy = np.zeros(1, dtype=np.uint8)
        err = self.lib.lto_zstd_compress(_ptr(data) if data.size else _ptr(dummy), C.c_uint64(data.size), _ptr(dst), C.c_uint64(cap), C.byref(n))
        assert err == 0, err
        return dst[:n.value].tobytes()

    def lz4_compress(self, data):
        data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        cap = self.lib.lto_lz4_bound(data.size)
        dst = np.zeros(cap, dtype=np.uint8)
        n = C.c_uint64(0)
        err = self.lib.lto_lz4_compress(_ptr(data), C.c_uint64(data.size), _ptr(dst), C.c_uint64(cap), C.byref(n))
        assert err == 0, err
        return dst[:n.value].tobytes()

    def lz4_compress_in_place(self, data, src_offset):
        """the same encoder with source and destination in ONE buffer: output from byte 0, source at src_offset (the layout of the device
        write path, longtail_b200/csrc/lz4.cu lz4_in_place_offset) -> compressed bytes"""
        data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        buf = np.zeros(src_offset + data.size + 16, dtype=np.uint8)
        buf[src_offset:src_offset + data.size] = data
        n = C.c_uint64(0)
        base = buf.ctypes.data
        err = self.lib.lto_lz4_compress(C.cast(base + src_offset, _u8p), C.c_uint64(data.size), C.cast(base, _u8p),
                                        C.c_uint64(self.lib.lto_lz4_bound(data.size)), C.byref(n))
        assert err == 0, err
        return buf[:n.value].tobytes()

    def lz4_decompress(self, comp, raw_size):
        comp = np.frombuffer(comp, dtype=np.uint8)
        dst = np.zeros(max(raw_size, 1), dtype=np.uint8)
        n = C.c_uint64(0)
        err = self.lib.lto_lz4_decompress(_ptr(comp), C.c_uint64(comp.size), _ptr(dst), C.c_uint64(raw_size), C.byref(n))
        assert err == 0, err
        return dst[:n.value].tobytes()

    def create_version_index(self, assets, target_chunk_size, hash_type=HASH_BLAKE3, tags=None, perms=None):
        a = _AssetArgs(assets, tags, perms)
        buf = C.c_void_p()
        size = C.c_uint64(0)
        err = self.lib.lto_create_version_index(C.c_uint32(a.n), a.paths, a.datas, a.sizes, a.perms, a.tags, C.c_uint32(hash_type),
                                                C.c_uint32(target_chunk_size), C.byref(buf), C.byref(size))
        assert err == 0, err
        out = C.string_at(buf, size.value)
        self.lib.lto_free(buf)
        return out

    def upsync(self, assets, target_chunk_size, max_block_size=8388608, max_chunks_per_block=1024, hash_type=HASH_BLAKE3,
               tags=None, perms=None):
        a = _AssetArgs(assets, tags, perms)
        buf = C.c_void_p()
        size = C.c_uint64(0)
        err = self.lib.lto_upsync(C.c_uint32(a.n), a.paths, a.datas, a.sizes, a.perms, a.tags, C.c_uint32(hash_type),
                                  C.c_uint32(target_chunk_size), C.c_uint32(max_block_size), C.c_uint32(max_chunks_per_block),
                                  C.byref(buf), C.byref(size))
        assert err == 0, err
        out = C.string_at(buf, size.value)
        self.lib.lto_free(buf)
        return parse_upsync(out)


class Reference:
    """libref_shim.so — the unmodified reference; ``available`` is False when it was not built"""

    def __init__(self):
        self.lib = _load("libref_shim.so")
        self.available = self.lib is not None
        if self.available:
            self.lib.ref_compress_bound.restype = C.c_uint64
            self.lib.ref_compress_bound.argtypes = [C.c_uint32, C.c_uint64]
            self.lib.ref_free.argtypes = [C.c_void_p]
            self.lib.ref_cpu_count.restype = C.c_uint32

    def cpu_count(self):
        return self.lib.ref_cpu_count()

    def chunk(self, data, mn, avg, mx):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        cap = data.size // mn + 2
        lens = np.zeros(cap, dtype=np.uint32)
        n = C.c_uint64(0)
        err = self.lib.ref_hpcdc_chunk(_ptr(data), C.c_uint64(data.size), C.c_uint32(mn), C.c_uint32(avg), C.c_uint32(mx),
                                       _ptr(lens, C.c_uint32), C.c_uint64(cap), C.byref(n))
        assert err == 0, err
        return lens[:n.value].copy()

    def hash(self, hash_type, data):
        data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        out = C.c_uint64(0)
        dummy = np.zeros(1, dtype=np.uint8)  # the reference validates data != NULL even for length 0
        err = self.lib.ref_hash_buffer(C.c_uint32(hash_type), _ptr(data) if data.size else _ptr(dummy), C.c_uint32(data.size), C.byref(out))
        assert err == 0, err
        return out.value

    def hash_segments(self, hash_type, base, offsets, lens):
        base = np.ascontiguousarray(base, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        out = np.zeros(offsets.size, dtype=np.uint64)
        err = self.lib.ref_hash_segments(C.c_uint32(hash_type), _ptr(base), C.c_uint64(offsets.size), _ptr(offsets, C.c_uint64),
                                         _ptr(lens, C.c_uint32), _ptr(out, C.c_uint64))
        assert err == 0, err
        return out

    def compress(self, comp_type, data):
        data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        cap = self.lib.ref_compress_bound(comp_type, data.size)
        dst = np.zeros(cap, dtype=np.uint8)
        n = C.c_uint64(0)
        err = self.lib.ref_compress(C.c_uint32(comp_type), _ptr(data), C.c_uint64(data.size), _ptr(dst), C.c_uint64(cap), C.byref(n))
        assert err == 0, err
        return dst[:n.value].tobytes()

    def decompress(self, comp_type, data, raw_size):
        data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        dst = np.zeros(max(raw_size, 1), dtype=np.uint8)
        n = C.c_uint64(0)
        err = self.lib.ref_decompress(C.c_uint32(comp_type), _ptr(data), C.c_uint64(data.size), _ptr(dst), C.c_uint64(raw_size), C.byref(n))
        assert err == 0, err
        return dst[:n.value].tobytes()

    def create_version_index(self, assets, target_chunk_size, hash_type=HASH_BLAKE3, tags=None, perms=None, workers=0, want_seconds=False):
        a = _AssetArgs(assets, tags, perms)
        buf = C.c_void_p()
        size = C.c_uint64(0)
        secs = C.c_double(0)
        err = self.lib.ref_create_version_index(C.c_uint32(a.n), a.paths, a.datas, a.sizes, a.perms, a.tags, C.c_uint32(hash_type),
                                                C.c_uint32(target_chunk_size), C.c_uint32(workers), C.byref(buf), C.byref(size), C.byref(secs))
        assert err == 0, err
        out = C.string_at(buf, size.value)
        self.lib.ref_free(buf)
        return (out, secs.value) if want_seconds else out

    def upsync(self, assets, target_chunk_size, max_block_size=8388608, max_chunks_per_block=1024, hash_type=HASH_BLAKE3,
               tags=None, perms=None, workers=0, keep_bytes=True, want_seconds=False, existing_hashes=None):
        """existing_hashes: chunk hashes a (non-empty) store already holds -> only the missing chunks are packed and written"""
        a = _AssetArgs(assets, tags, perms)
        buf = C.c_void_p()
        size = C.c_uint64(0)
        secs = (C.c_double * 3)()
        stored = C.c_uint64(0)
        if existing_hashes is not None and len(existing_hashes):
            eh = np.ascontiguousarray(existing_hashes, dtype=np.uint64)
            err = self.lib.ref_upsync_existing(C.c_uint32(a.n), a.paths, a.datas, a.sizes, a.perms, a.tags, C.c_uint32(hash_type),
                                               C.c_uint32(target_chunk_size), C.c_uint32(max_block_size), C.c_uint32(max_chunks_per_block),
                                               C.c_uint32(workers), C.c_int(int(keep_bytes)), C.byref(buf), C.byref(size), secs, C.byref(stored),
                                               C.c_uint32(eh.size), eh.ctypes.data_as(C.c_void_p))
        else:
            err = self.lib.ref_upsync(C.c_uint32(a.n), a.paths, a.datas, a.sizes, a.perms, a.tags, C.c_uint32(hash_type),
                                      C.c_uint32(target_chunk_size), C.c_uint32(max_block_size), C.c_uint32(max_chunks_per_block),
                                      C.c_uint32(workers), C.c_int(int(keep_bytes)), C.byref(buf), C.byref(size), secs, C.byref(stored))
        assert err == 0, err
        out = C.string_at(buf, size.value)
        self.lib.ref_free(buf)
        if not keep_bytes:
            return list(secs), stored.value
        if keep_bytes == 2:
            # -> (u64 array [blocks, 3] of {block hash, size, ref_digest64}, version index bytes, seconds)
            (count,) = struct.unpack_from("<I", out, 0)
            rec = np.frombuffer(out, dtype="<u8", count=3 * count, offset=4).reshape(count, 3).copy()
            (vsize,) = struct.unpack_from("<Q", out, 4 + 24 * count)
            return rec, out[12 + 24 * count:12 + 24 * count + vsize], list(secs)
        res = parse_upsync(out)
        return (res, list(secs)) if want_seconds else res


class DigestSink:
    """C block sink of the checker library (oracle/ref_shim.c: ref_digest_sink): records {block hash, size, ref_digest64} of every block
    the B200 path hands over — pass .fn / .user to Context.write_blocks_device(c_sink=...)"""

    def __init__(self, ref):
        self.lib = ref.lib
        self.lib.ref_digest_list_create.restype = C.c_void_p
        self.lib.ref_digest_list_free.argtypes = [C.c_void_p]
        self.lib.ref_digest_list_data.restype = C.c_void_p
        self.lib.ref_digest_list_data.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        self.user = C.c_void_p(self.lib.ref_digest_list_create())
        self.fn = C.cast(self.lib.ref_digest_sink, C.c_void_p)

    def records(self):
        n = C.c_uint64(0)
        ptr = self.lib.ref_digest_list_data(self.user, C.byref(n))
        if not n.value:
            return np.zeros((0, 3), dtype="<u8")
        return np.ctypeslib.as_array((C.c_uint64 * (3 * n.value)).from_address(ptr)).reshape(n.value, 3).copy()

    def close(self):
        if self.user:
            self.lib.ref_digest_list_free(self.user)
            self.user = None


def ref_upsync_to_dir(ref, assets, target_chunk_size, directory, max_block_size=8388608, max_chunks_per_block=1024, hash_type=HASH_BLAKE3,
                      tags=None, perms=None, workers=0):
    """the unmodified reference's upsync into an fsblockstore directory (incremental when the directory already holds a store);
    -> number of blocks it wrote"""
    a = _AssetArgs(assets, tags, perms)
    written = C.c_uint32(0)
    err = ref.lib.ref_upsync_to_dir(C.c_uint32(a.n), a.paths, a.datas, a.sizes, a.perms, a.tags, C.c_uint32(hash_type), C.c_uint32(target_chunk_size),
                                    C.c_uint32(max_block_size), C.c_uint32(max_chunks_per_block), C.c_uint32(workers), directory.encode(),
                                    C.byref(written))
    assert err == 0, err
    return written.value


def ref_read_store_dir(ref, directory):
    """the unmodified reference opens the directory as a block store and reads every block of store.lsi back (decompressed)
    -> (blocks, chunks, payload bytes, order-independent digest)"""
    out = (C.c_uint64 * 4)()
    err = ref.lib.ref_read_store_dir(directory.encode(), out)
    assert err == 0, err
    return tuple(int(x) for x in out)


def ref_scan_directory(ref, root, workers=0):
    """Longtail_GetFilesRecursively2 of the unmodified reference over its own file storage -> [(path, size, permissions)]"""
    buf, size = C.c_void_p(), C.c_uint64(0)
    err = ref.lib.ref_scan_directory(root.encode(), C.c_uint32(workers), C.byref(buf), C.byref(size))
    assert err == 0, err
    raw = C.string_at(buf, size.value)
    ref.lib.ref_free(buf)
    n = struct.unpack_from("<I", raw, 0)[0]
    out, p = [], 4
    for _ in range(n):
        sz, perm, ln = struct.unpack_from("<QHI", raw, p)
        p += 14
        out.append((raw[p:p + ln].decode(), sz, perm))
        p += ln
    return out


def ref_index_directory(ref, root, target_chunk_size, hash_type=HASH_BLAKE3, workers=4, tag=0):
    """GetFilesRecursively2 + CreateVersionIndex of the unmodified reference over a real directory -> serialised VersionIndex"""
    buf, size = C.c_void_p(), C.c_uint64(0)
    err = ref.lib.ref_index_directory(root.encode(), C.c_uint32(hash_type), C.c_uint32(target_chunk_size), C.c_uint32(workers), C.c_uint32(tag),
                                      C.byref(buf), C.byref(size))
    assert err == 0, err
    out = C.string_at(buf, size.value)
    ref.lib.ref_free(buf)
    return out


def ref_upsync_dir_to_dir(ref, source_root, store_dir, target_chunk_size, max_block_size=8388608, max_chunks_per_block=1024,
                          hash_type=HASH_BLAKE3, workers=0, tag=0):
    """cmd/main.c:UpSync of the unmodified reference: real source directory -> fsblockstore directory -> (VersionIndex bytes, blocks written)"""
    buf, size, written = C.c_void_p(), C.c_uint64(0), C.c_uint32(0)
    err = ref.lib.ref_upsync_dir_to_dir(source_root.encode(), store_dir.encode(), C.c_uint32(hash_type), C.c_uint32(target_chunk_size),
                                        C.c_uint32(max_block_size), C.c_uint32(max_chunks_per_block), C.c_uint32(workers), C.c_uint32(tag),
                                        C.byref(buf), C.byref(size), C.byref(written))
    assert err == 0, err
    out = C.string_at(buf, size.value)
    ref.lib.ref_free(buf)
    return out, written.value
