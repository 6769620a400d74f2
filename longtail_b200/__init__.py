"""longtail_b200 — B200 (sm_100a) implementation of longtail's chunk -> hash -> compress indexing path.

This package is a thin ctypes mirror of the C ABI in include/longtail_b200.h (the product is the CUDA
library ``longtail_b200/_lib/liblongtail_b200.so``).  There is no CPU fallback: importing works anywhere,
but creating a :class:`Context` without the built library or without a CUDA device raises.
"""
import ctypes as C
import errno
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "liblongtail_b200.so")

HASH_BLAKE3 = 0x626C6B33  # 'blk3'
HASH_BLAKE2 = 0x626C6B32  # 'blk2'
HASH_MEOW = 0x6D656F77  # 'meow'
COMPRESSION_ZSTD_MIN = 0x7A746431  # 'ztd1' (level 0 -> 3)
COMPRESSION_ZSTD_DEFAULT = 0x7A746432  # 'ztd2' (level 3)
COMPRESSION_LZ4 = 0x6C7A3432  # 'lz42'

_lib = None


class LongtailB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s (errno %d %s)" % (message, code, errno.errorcode.get(code, "?")))
        self.errno = code


class FileList:
    """Longtail_GetFilesRecursively2 over a real directory (lt_b200_scan_directory): entries in the reference's order; host code."""

    def __init__(self, root, threads=8):
        self.lib = load_library()
        self.handle = C.c_void_p()
        err = self.lib.lt_b200_scan_directory(os.fsencode(root), C.c_uint32(threads), C.byref(self.handle))
        if err:
            raise LongtailB200Error(err, "scan_directory(%s)" % root)
        class _Raw(C.Structure):  # Assets with path_data as a plain address (c_char_p fields stop at the first NUL)
            _fields_ = [("asset_count", C.c_uint32), ("path_data_size", C.c_uint32), ("sizes", C.POINTER(C.c_uint64)),
                        ("path_start_offsets", C.POINTER(C.c_uint32)), ("permissions", C.POINTER(C.c_uint16)), ("path_data", C.c_void_p)]
        a = _Raw.from_address(self.lib.lt_b200_file_list_assets(self.handle))
        n = a.asset_count
        self.sizes = [int(a.sizes[i]) for i in range(n)]
        self.permissions = [int(a.permissions[i]) for i in range(n)]
        data = C.string_at(a.path_data, a.path_data_size)
        self.paths = [data[a.path_start_offsets[i]:data.index(b"\0", a.path_start_offsets[i])].decode() for i in range(n)]

    def close(self):
        if self.handle:
            self.lib.lt_b200_file_list_free(self.handle)
            self.handle = C.c_void_p()


class FsStore:
    """On-disk block sink in the reference's fsblockstore layout (include/longtail_b200.h, lt_b200_fs_store_*): host code, no GPU."""

    def __init__(self, path, writer_threads=4):
        self.lib = load_library()
        self.handle = C.c_void_p()
        err = self.lib.lt_b200_fs_store_open(os.fsencode(path), C.c_uint32(writer_threads), C.byref(self.handle))
        if err:
            raise LongtailB200Error(err, "fs_store_open(%s)" % path)

    def put(self, block_hash, image):
        """hand one serialised stored block (bytes) to the sink, as lt_b200_write_blocks_device would"""
        buf = np.frombuffer(image, dtype=np.uint8)
        n = int(np.frombuffer(image, dtype=np.uint32, count=1, offset=12)[0])
        tag = int(np.frombuffer(image, dtype=np.uint32, count=1, offset=16)[0])
        v = StoredBlockView(int(block_hash), buf.ctypes.data, buf.size, n, tag, 0, 0)
        err = self.lib.lt_b200_fs_store_sink(self.handle, C.byref(v))
        if err:
            raise LongtailB200Error(err, "fs_store_sink")

    def flush(self):
        err = self.lib.lt_b200_fs_store_flush(self.handle)
        if err:
            raise LongtailB200Error(err, "fs_store_flush")

    def existing_chunks(self):
        n = C.c_uint32(0)
        err = self.lib.lt_b200_fs_store_existing_chunks(self.handle, None, 0, C.byref(n))
        if err:
            raise LongtailB200Error(err, "fs_store_existing_chunks")
        out = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            err = self.lib.lt_b200_fs_store_existing_chunks(self.handle, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n))
            if err:
                raise LongtailB200Error(err, "fs_store_existing_chunks")
        return out

    def stats(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self.lib.lt_b200_fs_store_stats(self.handle, C.byref(a), C.byref(b), C.byref(c))
        return {"blocks_written": a.value, "bytes_written": b.value, "blocks_skipped": c.value}

    def close(self):
        if self.handle:
            err = self.lib.lt_b200_fs_store_close(self.handle)
            self.handle = C.c_void_p()
            if err:
                raise LongtailB200Error(err, "fs_store_close")


class Range(C.Structure):
    _fields_ = [("arena_offset", C.c_uint64), ("size", C.c_uint32), ("tag", C.c_uint32)]


class ChunkTable(C.Structure):
    _fields_ = [("range_count", C.c_uint32), ("chunk_count", C.c_uint32), ("range_chunk_counts", C.POINTER(C.c_uint32)),
                ("chunk_hashes", C.POINTER(C.c_uint64)), ("chunk_sizes", C.POINTER(C.c_uint32)),
                ("chunk_tags", C.POINTER(C.c_uint32)), ("chunk_offsets", C.POINTER(C.c_uint64))]


class Assets(C.Structure):
    _fields_ = [("asset_count", C.c_uint32), ("path_data_size", C.c_uint32), ("sizes", C.POINTER(C.c_uint64)),
                ("path_start_offsets", C.POINTER(C.c_uint32)), ("permissions", C.POINTER(C.c_uint16)), ("path_data", C.c_char_p)]


class StoredBlockView(C.Structure):
    _fields_ = [("block_hash", C.c_uint64), ("data", C.c_void_p), ("size", C.c_uint64), ("chunk_count", C.c_uint32), ("tag", C.c_uint32),
                ("raw_payload_size", C.c_uint32), ("first_chunk", C.c_uint32)]


BLOCK_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(StoredBlockView))


class SynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("shared_permille", C.c_uint32), ("pool_segments", C.c_uint32),
                ("class_mode", C.c_uint32), ("reserved", C.c_uint32)]


def load_library():
    """dlopen the CUDA library; raises when it has not been built (python -m longtail_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LT_B200_LIB", LIB_PATH)  # development: A/B a differently built library
    if not os.path.exists(path):
        raise LongtailB200Error(errno.ENOENT, "liblongtail_b200.so is not built: run `python -m longtail_b200.build` "
                                              "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(path)
    lib.lt_b200_last_error.restype = C.c_char_p
    lib.lt_b200_last_error.argtypes = [C.c_void_p]
    lib.lt_b200_launch_count.restype = C.c_uint64
    lib.lt_b200_launch_count.argtypes = [C.c_void_p]
    lib.lt_b200_stream.restype = C.c_void_p
    lib.lt_b200_stream.argtypes = [C.c_void_p]
    lib.lt_b200_context_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.lt_b200_context_destroy.argtypes = [C.c_void_p]
    lib.lt_b200_synchronize.argtypes = [C.c_void_p]
    lib.lt_b200_trim.argtypes = [C.c_void_p]
    lib.lt_b200_device_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.lt_b200_device_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.lt_b200_host_alloc_pinned.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.lt_b200_host_free_pinned.argtypes = [C.c_void_p, C.c_void_p]
    lib.lt_b200_copy_to_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    lib.lt_b200_copy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    lib.lt_b200_synth_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(SynthSpec), C.c_uint64, C.c_uint64]
    lib.lt_b200_chunk_ranges.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(Range), C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_uint32, C.c_uint32, C.c_int, C.POINTER(ChunkTable)]
    lib.lt_b200_hash_segments.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.lt_b200_build_version_index.argtypes = [C.c_void_p, C.POINTER(Assets), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.lt_b200_index_device_assets.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(Assets), C.c_void_p, C.c_void_p, C.c_uint32,
                                                C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.lt_b200_index_host_assets.argtypes = [C.c_void_p, C.POINTER(Assets), C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                              C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.lt_b200_resident_table.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
    lib.lt_b200_build_version_index_device.argtypes = [C.c_void_p, C.POINTER(Assets), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                       C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.lt_b200_write_blocks_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_uint32, C.c_uint32, C.c_uint32, BLOCK_SINK, C.c_void_p]
    lib.lt_b200_write_blocks_device_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.lt_b200_upsync_host_assets.argtypes = [C.c_void_p, C.POINTER(Assets), C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                               C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                               C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    lib.lt_b200_upsync_stream_host_assets.argtypes = [C.c_void_p, C.POINTER(Assets), C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                      C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p,
                                                      C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    lib.lt_b200_comm_unique_id.argtypes = [C.c_void_p]
    lib.lt_b200_comm_create.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.lt_b200_comm_destroy.argtypes = [C.c_void_p]
    lib.lt_b200_plan_shards.argtypes = [C.POINTER(Assets), C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
    lib.lt_b200_shard_jobs.argtypes = [C.POINTER(Assets), C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.lt_b200_index_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(Assets), C.c_void_p, C.c_void_p, C.c_uint32,
                                          C.c_uint32, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.lt_b200_write_blocks_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                                 C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.lt_b200_unique_chunk_offsets.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.lt_b200_scan_directory.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.lt_b200_file_list_assets.restype = C.c_void_p
    lib.lt_b200_file_list_assets.argtypes = [C.c_void_p]
    lib.lt_b200_file_list_free.argtypes = [C.c_void_p]
    lib.lt_b200_index_file_list.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_uint64)]
    lib.lt_b200_upsync_file_list.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                             C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    lib.lt_b200_fs_store_open.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.lt_b200_fs_store_sink.argtypes = [C.c_void_p, C.POINTER(StoredBlockView)]
    lib.lt_b200_fs_store_flush.argtypes = [C.c_void_p]
    lib.lt_b200_fs_store_close.argtypes = [C.c_void_p]
    lib.lt_b200_fs_store_existing_chunks.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    lib.lt_b200_fs_store_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.lt_b200_profile_enable.argtypes = [C.c_void_p, C.c_int]
    lib.lt_b200_profile_reset.argtypes = [C.c_void_p]
    lib.lt_b200_profile_read.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    _lib = lib
    return lib


WRITE_DEVICE_SINK = 1
COMM_ID_BYTES = 128
KERNEL_NAMES = {0: "k_hpcdc_scan", 1: "k_hpcdc_walk", 2: "k_blake3_leaves", 3: "k_blake3_merge", 4: "k_gather_chunks", 5: "k_lz4_blocks", 6: "k_blake2s_segments", 7: "k_meow_segments", 8: "k_zstd_frames", 9: "k_lz4_decode", 10: "k_zstd_decode"}


def parse_version_index(buf):
    """-> dict of numpy views over a serialised VersionIndex (src/longtail.c:2566-2584 layout)"""
    b = np.frombuffer(buf, dtype=np.uint8)
    hdr = b[:24].view("<u4")
    A, Cn, I = int(hdr[3]), int(hdr[4]), int(hdr[5])
    o = 24
    out = {"version": int(hdr[0]), "hash_id": int(hdr[1]), "target_chunk_size": int(hdr[2]), "asset_count": A, "chunk_count": Cn,
           "asset_chunk_index_count": I}

    def take(name, n, dt):
        nonlocal o
        size = n * np.dtype(dt).itemsize
        out[name] = b[o:o + size].view(dt)
        o += size

    take("path_hashes", A, "<u8")
    take("content_hashes", A, "<u8")
    take("asset_sizes", A, "<u8")
    take("asset_chunk_counts", A, "<u4")
    take("asset_chunk_index_starts", A, "<u4")
    take("asset_chunk_indexes", I, "<u4")
    take("chunk_hashes", Cn, "<u8")
    take("chunk_sizes", Cn, "<u4")
    take("chunk_tags", Cn, "<u4")
    take("name_offsets", A, "<u4")
    take("permissions", A, "<u2")
    out["name_data"] = bytes(b[o:])
    return out


def chunker_params(target_chunk_size):
    """min/avg/max of the reference driver (src/longtail.c:1985-1987, GetMinChunkSize() == 48)"""
    t = int(target_chunk_size)
    return max(48, t // 8), max(48, t // 2), max(48, t * 2)


class AssetList:
    """what Longtail_FileInfos carries (src/longtail.h:1684-1692): relative paths (dirs end in '/'), sizes, permissions"""

    def __init__(self, paths, sizes, permissions=None):
        self.paths = [p if isinstance(p, bytes) else p.encode() for p in paths]
        self.count = len(self.paths)
        self.sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
        self.permissions = np.ascontiguousarray(permissions if permissions is not None else [0o644] * self.count, dtype=np.uint16)
        offs, blob, o = [], bytearray(), 0
        for p in self.paths:
            offs.append(o)
            blob += p + b"\0"
            o += len(p) + 1
        self.path_data = bytes(blob)
        self.path_start_offsets = np.ascontiguousarray(offs, dtype=np.uint32)
        assert self.sizes.size == self.count and self.permissions.size == self.count

    def as_struct(self):
        a = Assets()
        a.asset_count = self.count
        a.path_data_size = len(self.path_data)
        a.sizes = self.sizes.ctypes.data_as(C.POINTER(C.c_uint64))
        a.path_start_offsets = self.path_start_offsets.ctypes.data_as(C.POINTER(C.c_uint32))
        a.permissions = self.permissions.ctypes.data_as(C.POINTER(C.c_uint16))
        a.path_data = self.path_data
        return a


def plan_shards(assets, target_chunk_size, world):
    """host helper (no GPU): -> (first_job[world + 1], job count) — contiguous slices of the reference's (asset, part) job list, balanced by bytes"""
    lib = load_library()
    first = np.zeros(world + 1, dtype=np.uint32)
    n = C.c_uint32(0)
    a = assets.as_struct()
    err = lib.lt_b200_plan_shards(C.byref(a), int(target_chunk_size), int(world), first.ctypes.data_as(C.c_void_p), C.byref(n))
    if err:
        raise LongtailB200Error(err, "lt_b200_plan_shards")
    return first, n.value


def shard_jobs(assets, target_chunk_size, first_job, job_count):
    """host helper (no GPU): -> structured array (asset_index, size, offset) of jobs [first_job, first_job + job_count)"""
    lib = load_library()
    out = np.zeros(max(int(job_count), 1), dtype=np.dtype([("asset_index", "<u4"), ("size", "<u4"), ("offset", "<u8")]))
    a = assets.as_struct()
    err = lib.lt_b200_shard_jobs(C.byref(a), int(target_chunk_size), int(first_job), int(job_count), out.ctypes.data_as(C.c_void_p))
    if err:
        raise LongtailB200Error(err, "lt_b200_shard_jobs")
    return out[:int(job_count)]


class Context:
    """one per GPU; mirrors lt_b200_context"""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        err = self.lib.lt_b200_context_create(int(device), C.byref(h))
        if err:
            raise LongtailB200Error(err, "lt_b200_context_create(device %d) failed: no usable CUDA device (no CPU fallback)" % device)
        self.handle = h
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.lib.lt_b200_context_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, err, what):
        if err:
            raise LongtailB200Error(err, "%s: %s" % (what, self.lib.lt_b200_last_error(self.handle).decode()))

    # ---- plumbing
    @property
    def stream(self):
        return self.lib.lt_b200_stream(self.handle)

    @property
    def launch_count(self):
        return int(self.lib.lt_b200_launch_count(self.handle))

    def trim(self):
        """free the grow-only device workspace (it comes back on demand)"""
        self._check(self.lib.lt_b200_trim(self.handle), "trim")

    def synchronize(self):
        self._check(self.lib.lt_b200_synchronize(self.handle), "synchronize")

    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.lib.lt_b200_device_alloc(self.handle, int(nbytes), C.byref(p)), "device_alloc")
        return p.value

    def device_free(self, ptr):
        self._check(self.lib.lt_b200_device_free(self.handle, C.c_void_p(ptr)), "device_free")

    def pinned_alloc(self, nbytes):
        """-> numpy uint8 view over pinned host memory (keep the returned array alive; free with pinned_free(arr))"""
        p = C.c_void_p()
        self._check(self.lib.lt_b200_host_alloc_pinned(self.handle, int(nbytes), C.byref(p)), "host_alloc_pinned")
        arr = np.ctypeslib.as_array((C.c_uint8 * max(int(nbytes), 1)).from_address(p.value))[:int(nbytes)]
        arr.flags.writeable = True
        return arr

    def pinned_free(self, arr):
        self._check(self.lib.lt_b200_host_free_pinned(self.handle, C.c_void_p(arr.ctypes.data)), "host_free_pinned")

    def to_device(self, dptr, host_array):
        a = np.ascontiguousarray(host_array)
        self._check(self.lib.lt_b200_copy_to_device(self.handle, C.c_void_p(dptr), a.ctypes.data_as(C.c_void_p), a.nbytes), "copy_to_device")

    def to_host(self, dptr, nbytes):
        out = np.empty(int(nbytes), dtype=np.uint8)
        self._check(self.lib.lt_b200_copy_to_host(self.handle, out.ctypes.data_as(C.c_void_p), C.c_void_p(dptr), int(nbytes)), "copy_to_host")
        return out

    def synth_fill(self, dptr, nbytes, seed, asset_id=0, offset=0, shared_permille=0, pool_segments=1, class_mode=0):
        spec = SynthSpec(int(seed), int(shared_permille), int(pool_segments), int(class_mode), 0)
        self._check(self.lib.lt_b200_synth_fill(self.handle, C.c_void_p(dptr), int(nbytes), C.byref(spec), int(asset_id), int(offset)), "synth_fill")

    def profile_enable(self, on=True):
        self._check(self.lib.lt_b200_profile_enable(self.handle, 1 if on else 0), "profile_enable")

    def profile_reset(self):
        self._check(self.lib.lt_b200_profile_reset(self.handle), "profile_reset")

    def profile_read(self):
        """-> {kernel name: (total ms, launches, algorithmic bytes)} accumulated since the last reset"""
        out = {}
        for k, name in KERNEL_NAMES.items():
            ms, n, b = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
            self._check(self.lib.lt_b200_profile_read(self.handle, k, C.byref(ms), C.byref(n), C.byref(b)), "profile_read")
            out[name] = (ms.value, int(n.value), int(b.value))
        return out

    def resident_table(self):
        """device addresses (hashes u64, sizes u32, tags u32) and length of the table left by the last chunk_ranges call"""
        h, s, t, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint32(0)
        self._check(self.lib.lt_b200_resident_table(self.handle, C.byref(h), C.byref(s), C.byref(t), C.byref(n)), "resident_table")
        return h.value, s.value, t.value, int(n.value)

    def build_version_index_device(self, assets, asset_chunk_counts, chunk_count, d_hashes, d_sizes, d_tags, hash_type=HASH_BLAKE3,
                                   target_chunk_size=32768, copy=True):
        st = assets.as_struct()
        counts = np.ascontiguousarray(asset_chunk_counts, dtype=np.uint32)
        buf, size = C.c_void_p(), C.c_uint64(0)
        self._check(self.lib.lt_b200_build_version_index_device(self.handle, C.byref(st), counts.ctypes.data_as(C.c_void_p), int(chunk_count),
                                                                C.c_void_p(d_hashes), C.c_void_p(d_sizes), C.c_void_p(d_tags), int(hash_type),
                                                                int(target_chunk_size), C.byref(buf), C.byref(size)), "build_version_index_device")
        return self._result(buf, size, copy)

    # ---- layer 1
    def chunk_ranges(self, dptr, arena_size, ranges, min_size, avg_size, max_size, hash_type=HASH_BLAKE3, want_host=True):
        """ranges: iterable of (arena_offset, size[, tag]).  Returns dict of numpy arrays (copies)."""
        rs = (Range * max(len(ranges), 1))()
        for i, r in enumerate(ranges):
            rs[i].arena_offset, rs[i].size, rs[i].tag = int(r[0]), int(r[1]), int(r[2]) if len(r) > 2 else 0
        t = ChunkTable()
        self._check(self.lib.lt_b200_chunk_ranges(self.handle, C.c_void_p(dptr), int(arena_size), rs, len(ranges), int(min_size), int(avg_size),
                                                  int(max_size), int(hash_type), 1 if want_host else 0, C.byref(t)), "chunk_ranges")
        out = {"chunk_count": int(t.chunk_count),
               "range_chunk_counts": np.ctypeslib.as_array(t.range_chunk_counts, (max(t.range_count, 1),))[:t.range_count].copy()
               if t.range_count else np.zeros(0, np.uint32)}
        if want_host and t.chunk_count:
            n = t.chunk_count
            out["hashes"] = np.ctypeslib.as_array(t.chunk_hashes, (n,)).copy()
            out["sizes"] = np.ctypeslib.as_array(t.chunk_sizes, (n,)).copy()
            out["tags"] = np.ctypeslib.as_array(t.chunk_tags, (n,)).copy()
            out["offsets"] = np.ctypeslib.as_array(t.chunk_offsets, (n,)).copy()
        elif want_host:
            out.update(hashes=np.zeros(0, np.uint64), sizes=np.zeros(0, np.uint32), tags=np.zeros(0, np.uint32), offsets=np.zeros(0, np.uint64))
        return out

    def hash_segments(self, dptr, base_size, offsets, sizes, hash_type=HASH_BLAKE3):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
        out = np.zeros(offsets.size, dtype=np.uint64)
        self._check(self.lib.lt_b200_hash_segments(self.handle, int(hash_type), C.c_void_p(dptr), int(base_size), offsets.ctypes.data_as(C.c_void_p),
                                                   sizes.ctypes.data_as(C.c_void_p), offsets.size, out.ctypes.data_as(C.c_void_p)), "hash_segments")
        return out

    # ---- the WriteContent half
    def unique_chunk_offsets(self, count):
        out = np.zeros(int(count), dtype=np.uint64)
        self._check(self.lib.lt_b200_unique_chunk_offsets(self.handle, out.ctypes.data_as(C.c_void_p), int(count)), "unique_chunk_offsets")
        return out

    def missing_chunks(self, chunk_hashes, existing_hashes):
        """-> bool mask over chunk_hashes: True where the store (existing_hashes) does not hold the chunk yet (DiffHashes)"""
        h = np.ascontiguousarray(chunk_hashes, dtype=np.uint64)
        e = np.ascontiguousarray(existing_hashes, dtype=np.uint64)
        out = np.zeros(max(h.size, 1), dtype=np.uint8)
        self._check(self.lib.lt_b200_missing_chunks(self.handle, C.c_uint32(h.size), h.ctypes.data_as(C.c_void_p), C.c_uint32(e.size),
                                                    e.ctypes.data_as(C.c_void_p) if e.size else None, out.ctypes.data_as(C.c_void_p)), "missing_chunks")
        return out[:h.size].astype(bool)

    def counting_sink(self):
        """-> (fn, user, acc): the library's counting sink and its uint64[4] accumulator {blocks, stored bytes, raw bytes, xor of hashes}"""
        acc = np.zeros(4, dtype=np.uint64)
        return C.cast(self.lib.lt_b200_counting_sink, C.c_void_p), acc.ctypes.data_as(C.c_void_p), acc

    def write_blocks_device(self, dptr, arena_size, chunk_hashes, chunk_sizes, chunk_tags, chunk_offsets, max_block_size=8388608,
                            max_chunks_per_block=1024, hash_type=HASH_BLAKE3, keep_bytes=True, fs_store=None, c_sink=None, device_sink=False):
        """-> list of (block_hash, serialised stored block bytes | size) in store order; with fs_store (an FsStore) the blocks go
        straight to its C sink (no Python in the loop) and the result is None; c_sink = (function pointer, user pointer) of any C
        lt_b200_block_sink; device_sink leaves the block images in HBM (LT_B200_WRITE_DEVICE_SINK)"""
        h = np.ascontiguousarray(chunk_hashes, dtype=np.uint64)
        s = np.ascontiguousarray(chunk_sizes, dtype=np.uint32)
        t = np.ascontiguousarray(chunk_tags, dtype=np.uint32)
        o = np.ascontiguousarray(chunk_offsets, dtype=np.uint64)
        if c_sink is not None:
            self._check(self.lib.lt_b200_write_blocks_device_ex(self.handle, C.c_void_p(dptr), int(arena_size), h.size, h.ctypes.data_as(C.c_void_p),
                                                                s.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p),
                                                                int(hash_type), int(max_block_size), int(max_chunks_per_block),
                                                                WRITE_DEVICE_SINK if device_sink else 0, c_sink[0], c_sink[1]), "write_blocks_device_ex")
            return None
        if fs_store is not None:
            cb = C.cast(self.lib.lt_b200_fs_store_sink, BLOCK_SINK)
            self._check(self.lib.lt_b200_write_blocks_device(self.handle, C.c_void_p(dptr), int(arena_size), h.size, h.ctypes.data_as(C.c_void_p),
                                                             s.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p),
                                                             int(hash_type), int(max_block_size), int(max_chunks_per_block), cb, fs_store.handle),
                        "write_blocks_device")
            return None
        blocks = []

        def sink(_user, view):
            v = view.contents
            blocks.append((int(v.block_hash), C.string_at(v.data, v.size) if keep_bytes else int(v.size)))
            return 0

        cb = BLOCK_SINK(sink)
        self._check(self.lib.lt_b200_write_blocks_device(self.handle, C.c_void_p(dptr), int(arena_size), h.size, h.ctypes.data_as(C.c_void_p),
                                                         s.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p),
                                                         int(hash_type), int(max_block_size), int(max_chunks_per_block), cb, None), "write_blocks_device")
        return blocks

    def upsync_host_assets(self, assets, datas, tags, c_sink, target_chunk_size=32768, max_block_size=8388608, max_chunks_per_block=1024,
                           hash_type=HASH_BLAKE3, existing_hashes=None, device_sink=False, copy=True):
        """the whole upsync for assets in host memory (list of uint8 arrays, pinned for full PCIe speed): H2D once, CreateVersionIndex,
        CreateMissingContent, WriteContent into c_sink = (fn, user); -> (serialised VersionIndex, chunks written)"""
        st = assets.as_struct()
        ptrs = (C.c_void_p * max(len(datas), 1))(*[d.ctypes.data if d.size else None for d in datas])
        tg = None if tags is None else np.ascontiguousarray(tags, dtype=np.uint32)
        eh = None if existing_hashes is None else np.ascontiguousarray(existing_hashes, dtype=np.uint64)
        buf, size, written = C.c_void_p(), C.c_uint64(0), C.c_uint32(0)
        self._check(self.lib.lt_b200_upsync_host_assets(self.handle, C.byref(st), ptrs, None if tg is None else tg.ctypes.data_as(C.c_void_p),
                                                        int(hash_type), int(target_chunk_size), int(max_block_size), int(max_chunks_per_block),
                                                        0 if eh is None else int(eh.size), None if eh is None else eh.ctypes.data_as(C.c_void_p),
                                                        WRITE_DEVICE_SINK if device_sink else 0, c_sink[0], c_sink[1], C.byref(buf), C.byref(size),
                                                        C.byref(written)), "upsync_host_assets")
        return self._result(buf, size, copy), written.value

    def upsync_stream_host_assets(self, assets, datas, tags, c_sink, target_chunk_size=32768, max_block_size=8388608, max_chunks_per_block=1024,
                                  hash_type=HASH_BLAKE3, existing_hashes=None, device_sink=False, batch_bytes=0, copy=True):
        """upsync_host_assets as one streaming pass (batches of batch_bytes; the version need not fit the device); same results"""
        st = assets.as_struct()
        ptrs = (C.c_void_p * max(len(datas), 1))(*[d.ctypes.data if d.size else None for d in datas])
        tg = None if tags is None else np.ascontiguousarray(tags, dtype=np.uint32)
        eh = None if existing_hashes is None else np.ascontiguousarray(existing_hashes, dtype=np.uint64)
        buf, size, written = C.c_void_p(), C.c_uint64(0), C.c_uint32(0)
        self._check(self.lib.lt_b200_upsync_stream_host_assets(self.handle, C.byref(st), ptrs, None if tg is None else tg.ctypes.data_as(C.c_void_p),
                                                               int(hash_type), int(target_chunk_size), int(max_block_size), int(max_chunks_per_block),
                                                               0 if eh is None else int(eh.size), None if eh is None else eh.ctypes.data_as(C.c_void_p),
                                                               WRITE_DEVICE_SINK if device_sink else 0, C.c_uint64(int(batch_bytes)), c_sink[0], c_sink[1],
                                                               C.byref(buf), C.byref(size), C.byref(written)), "upsync_stream_host_assets")
        return self._result(buf, size, copy), written.value

    # ---- multi-GPU (one process per GPU; include/longtail_b200.h "multi-GPU")
    def comm_unique_id(self):
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        err = self.lib.lt_b200_comm_unique_id(buf)
        if err:
            raise LongtailB200Error(err, "lt_b200_comm_unique_id failed (NCCL not loadable?)")
        return bytes(buf)

    def comm_create(self, unique_id, rank, world):
        h = C.c_void_p()
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        self._check(self.lib.lt_b200_comm_create(self.handle, buf, int(rank), int(world), C.byref(h)), "comm_create")
        return h

    def comm_destroy(self, comm):
        self.lib.lt_b200_comm_destroy(comm)

    def plan_shards(self, assets, target_chunk_size, world):
        return plan_shards(assets, target_chunk_size, world)

    def shard_jobs(self, assets, target_chunk_size, first_job, job_count):
        return shard_jobs(assets, target_chunk_size, first_job, job_count)

    def index_sharded(self, comm, dptr, arena_size, assets, asset_tags, job_arena_offsets, target_chunk_size, hash_type=HASH_BLAKE3,
                      want_host=True, copy=True):
        """collective CreateVersionIndex; -> serialised VersionIndex (bytes / view) where want_host, else its size"""
        a = assets.as_struct()
        tags = None if asset_tags is None else np.ascontiguousarray(asset_tags, dtype=np.uint32)
        offs = np.ascontiguousarray(job_arena_offsets, dtype=np.uint64)
        buf, size = C.c_void_p(), C.c_uint64(0)
        self._check(self.lib.lt_b200_index_sharded(self.handle, comm, C.c_void_p(dptr), int(arena_size), C.byref(a),
                                                   tags.ctypes.data_as(C.c_void_p) if tags is not None else None,
                                                   offs.ctypes.data_as(C.c_void_p) if offs.size else None, int(hash_type),
                                                   int(target_chunk_size), 1 if want_host else 0, C.byref(buf), C.byref(size)), "index_sharded")
        if not want_host:
            return int(size.value)
        return self._result(buf, size, copy)

    def write_blocks_sharded(self, comm, c_sink, max_block_size=8388608, max_chunks_per_block=1024, device_sink=False):
        """collective WriteContent of a fresh store after index_sharded; this rank's blocks go to c_sink = (fn, user);
        -> (blocks written by this rank, blocks of the whole store)"""
        mine, total = C.c_uint32(0), C.c_uint32(0)
        self._check(self.lib.lt_b200_write_blocks_sharded(self.handle, comm, int(max_block_size), int(max_chunks_per_block),
                                                          WRITE_DEVICE_SINK if device_sink else 0, c_sink[0], c_sink[1], C.byref(mine), C.byref(total)),
                    "write_blocks_sharded")
        return mine.value, total.value

    def lz4_compress_host(self, buffers):
        """CompressionAPI.Compress for 'lz42' over a list of host buffers in one launch -> list of LZ4 blocks (bytes)"""
        keep = [np.ascontiguousarray(b, dtype=np.uint8) for b in buffers]
        n = len(keep)
        self.lib.lt_b200_lz4_bound.restype = C.c_uint64
        self.lib.lt_b200_lz4_bound.argtypes = [C.c_uint64]
        caps = [int(self.lib.lt_b200_lz4_bound(b.size)) for b in keep]
        outs = [np.empty(max(c, 1), dtype=np.uint8) for c in caps]
        src = (C.c_void_p * max(n, 1))(*[b.ctypes.data if b.size else None for b in keep])
        dst = (C.c_void_p * max(n, 1))(*[o.ctypes.data for o in outs])
        sizes = (C.c_uint32 * max(n, 1))(*[b.size for b in keep])
        cap = (C.c_uint64 * max(n, 1))(*caps)
        got = (C.c_uint64 * max(n, 1))()
        self._check(self.lib.lt_b200_lz4_compress_host(self.handle, C.c_uint32(n), src, sizes, dst, cap, got), "lz4_compress_host")
        return [outs[i][:got[i]].tobytes() for i in range(n)]

    def zstd_compress_host(self, buffers, compression_type=None):
        """CompressionAPI.Compress for 'ztd2' / 'ztd1' over a list of host buffers in one launch -> list of frames (bytes)"""
        ctype = COMPRESSION_ZSTD_DEFAULT if compression_type is None else compression_type
        keep = [np.ascontiguousarray(b, dtype=np.uint8) for b in buffers]
        n = len(keep)
        self.lib.lt_b200_zstd_bound.restype = C.c_uint64
        self.lib.lt_b200_zstd_bound.argtypes = [C.c_uint64]
        caps = [int(self.lib.lt_b200_zstd_bound(b.size)) for b in keep]
        outs = [np.empty(max(c, 1), dtype=np.uint8) for c in caps]
        src = (C.c_void_p * max(n, 1))(*[b.ctypes.data if b.size else None for b in keep])
        dst = (C.c_void_p * max(n, 1))(*[o.ctypes.data for o in outs])
        sizes = (C.c_uint32 * max(n, 1))(*[b.size for b in keep])
        cap = (C.c_uint64 * max(n, 1))(*caps)
        got = (C.c_uint64 * max(n, 1))()
        self._check(self.lib.lt_b200_zstd_compress_host(self.handle, C.c_uint32(ctype), C.c_uint32(n), src, sizes, dst, cap, got), "zstd_compress_host")
        return [outs[i][:got[i]].tobytes() for i in range(n)]

    def _decompress_host(self, fn, name, buffers, raw_sizes):
        keep = [np.ascontiguousarray(np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b, dtype=np.uint8) for b in buffers]
        n = len(keep)
        outs = [np.empty(max(int(c), 1), dtype=np.uint8) for c in raw_sizes]
        src = (C.c_void_p * max(n, 1))(*[b.ctypes.data if b.size else None for b in keep])
        dst = (C.c_void_p * max(n, 1))(*[o.ctypes.data for o in outs])
        sizes = (C.c_uint32 * max(n, 1))(*[b.size for b in keep])
        cap = (C.c_uint64 * max(n, 1))(*[int(c) for c in raw_sizes])
        got = (C.c_uint64 * max(n, 1))()
        self._check(fn(self.handle, C.c_uint32(n), src, sizes, dst, cap, got), name)
        return [outs[i][:got[i]].tobytes() for i in range(n)]

    def lz4_decompress_host(self, buffers, raw_sizes):
        """CompressionAPI.Decompress for 'lz42' over a list of LZ4 blocks in one launch; raw_sizes = capacities of the outputs"""
        return self._decompress_host(self.lib.lt_b200_lz4_decompress_host, "lz4_decompress_host", buffers, raw_sizes)

    def zstd_decompress_host(self, buffers, raw_sizes):
        """CompressionAPI.Decompress for 'ztd1'..'ztd5' over a list of frames in one launch (one warp per frame)"""
        return self._decompress_host(self.lib.lt_b200_zstd_decompress_host, "zstd_decompress_host", buffers, raw_sizes)

    # ---- layer 2
    def _result(self, buf, size, copy):
        if copy:
            return C.string_at(buf.value, size.value)
        raw = (C.c_uint8 * max(size.value, 1)).from_address(buf.value)
        return memoryview(raw)[:size.value]

    def build_version_index(self, assets, asset_chunk_counts, chunk_hashes=None, chunk_sizes=None, chunk_tags=None, chunk_count=None,
                            hash_type=HASH_BLAKE3, target_chunk_size=32768, copy=True):
        st = assets.as_struct()
        counts = np.ascontiguousarray(asset_chunk_counts, dtype=np.uint32)
        buf, size = C.c_void_p(), C.c_uint64(0)
        if chunk_hashes is None:
            n = int(chunk_count if chunk_count is not None else counts.sum())
            args = (None, None, None)
        else:
            h = np.ascontiguousarray(chunk_hashes, dtype=np.uint64)
            s = np.ascontiguousarray(chunk_sizes, dtype=np.uint32)
            t = np.ascontiguousarray(chunk_tags, dtype=np.uint32)
            n = h.size
            args = (h.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
        self._check(self.lib.lt_b200_build_version_index(self.handle, C.byref(st), counts.ctypes.data_as(C.c_void_p), n, args[0], args[1], args[2],
                                                         int(hash_type), int(target_chunk_size), C.byref(buf), C.byref(size)), "build_version_index")
        return self._result(buf, size, copy)

    def index_device_assets(self, dptr, arena_size, assets, arena_offsets, tags=None, hash_type=HASH_BLAKE3, target_chunk_size=32768, copy=True):
        st = assets.as_struct()
        offs = np.ascontiguousarray(arena_offsets, dtype=np.uint64)
        tg = None if tags is None else np.ascontiguousarray(tags, dtype=np.uint32)
        buf, size = C.c_void_p(), C.c_uint64(0)
        self._check(self.lib.lt_b200_index_device_assets(self.handle, C.c_void_p(dptr), int(arena_size), C.byref(st), offs.ctypes.data_as(C.c_void_p),
                                                         None if tg is None else tg.ctypes.data_as(C.c_void_p), int(hash_type),
                                                         int(target_chunk_size), C.byref(buf), C.byref(size)), "index_device_assets")
        return self._result(buf, size, copy)

    def index_file_list(self, file_list, tags=None, hash_type=HASH_BLAKE3, target_chunk_size=32768, reader_threads=8, copy=True):
        """CreateVersionIndex over a scanned directory: reader threads pread the parts into pinned staging behind the streaming verb"""
        tg = None if tags is None else np.ascontiguousarray(tags, dtype=np.uint32)
        buf, size = C.c_void_p(), C.c_uint64(0)
        self._check(self.lib.lt_b200_index_file_list(self.handle, file_list.handle, None if tg is None else tg.ctypes.data_as(C.c_void_p),
                                                     int(hash_type), int(target_chunk_size), int(reader_threads), C.byref(buf), C.byref(size)),
                    "index_file_list")
        return self._result(buf, size, copy)

    def upsync_file_list(self, file_list, fs_store, tags=None, hash_type=HASH_BLAKE3, target_chunk_size=32768, max_block_size=8388608,
                         max_chunks_per_block=1024, reader_threads=8):
        """cmd/main.c:UpSync for a scanned tree that fits the GPU -> (serialised VersionIndex, stored blocks written)"""
        tg = None if tags is None else np.ascontiguousarray(tags, dtype=np.uint32)
        buf, size, written = C.c_void_p(), C.c_uint64(0), C.c_uint32(0)
        self._check(self.lib.lt_b200_upsync_file_list(self.handle, file_list.handle, None if tg is None else tg.ctypes.data_as(C.c_void_p),
                                                      int(hash_type), int(target_chunk_size), int(max_block_size), int(max_chunks_per_block),
                                                      int(reader_threads), fs_store.handle, C.byref(buf), C.byref(size), C.byref(written)),
                    "upsync_file_list")
        return self._result(buf, size, True), written.value

    def index_host_assets(self, assets, datas, tags=None, hash_type=HASH_BLAKE3, target_chunk_size=32768, copy=True):
        """datas: list of contiguous uint8 numpy arrays (pinned for full PCIe speed), one per asset"""
        st = assets.as_struct()
        keep = [np.ascontiguousarray(d, dtype=np.uint8) for d in datas]
        ptrs = (C.c_void_p * max(len(keep), 1))(*[d.ctypes.data if d.size else None for d in keep])
        tg = None if tags is None else np.ascontiguousarray(tags, dtype=np.uint32)
        buf, size = C.c_void_p(), C.c_uint64(0)
        self._check(self.lib.lt_b200_index_host_assets(self.handle, C.byref(st), ptrs, None if tg is None else tg.ctypes.data_as(C.c_void_p),
                                                       int(hash_type), int(target_chunk_size), C.byref(buf), C.byref(size)), "index_host_assets")
        return self._result(buf, size, copy)
