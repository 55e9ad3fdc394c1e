"""ctypes binding of the C ABI in include/vrod_knn.h (vrod_b200/libvrod_knn.so).

This is the same shape of binding a vRod maintainer would write in Rust; the
tests and bench.py drive the library through it.  It has no CPU path: every compute call needs the
CUDA library and a B200, and raises VrodError otherwise.  Nothing here imports oracle/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# VROD_LIB: development override (e.g. the `make debug` build with the kernels' timing stamps compiled in)
LIB_PATH = os.environ.get("VROD_LIB") or os.path.join(_HERE, "libvrod_knn.so")

OK, EINVAL, ENOTFOUND, EEXISTS, ENOMEM, ECUDA, ENCCL, ENOGPU = range(8)
EUCLIDEAN, COSINE = 0, 1
MAX_K = 1024
COMM_ID_BYTES = 128
PAD_ID = 0xFFFFFFFFFFFFFFFF
PATH_AUTO, PATH_SCAN, PATH_EXACT, PATH_BATCHED = 0, 1, 2, 3

# every symbol include/vrod_knn.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "vrod_ctx_create", "vrod_comm_unique_id", "vrod_ctx_create_sharded", "vrod_ctx_create_multi", "vrod_ctx_devices",
    "vrod_ctx_destroy",
    "vrod_ctx_synchronize", "vrod_ctx_stream", "vrod_ctx_stats", "vrod_ctx_profile", "vrod_ctx_profile_read",
    "vrod_ctx_rank", "vrod_ctx_world",
    "vrod_collection_create", "vrod_collection_get", "vrod_collection_drop", "vrod_collection_list",
    "vrod_collection_info", "vrod_collection_insert", "vrod_collection_fill_synthetic",
    "vrod_collection_read_rows", "vrod_collection_shard", "vrod_collection_save", "vrod_collection_load",
    "vrod_collection_search",
    "vrod_collection_search_device", "vrod_collection_set_path", "vrod_last_error", "vrod_version",
]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("searches", "kernel_launches", "fast_scans", "exact_rescans",
                                          "batched_tiles", "h2d_bytes", "d2h_bytes")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class VrodError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"vrod status {status}: {msg}")
        self.status = status


_lib = None


def lib():
    """Load the CUDA library.  There is no fallback: a missing .so is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VrodError(ENOGPU, f"{LIB_PATH} is not built (run `make` or __graft_entry__.build())")
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.vrod_ctx_create.argtypes = [i32, C.POINTER(vp)]
        L.vrod_comm_unique_id.argtypes = [vp]
        L.vrod_ctx_create_sharded.argtypes = [i32, i32, i32, vp, C.POINTER(vp)]
        L.vrod_ctx_create_multi.argtypes = [C.POINTER(i32), i32, C.POINTER(vp)]
        L.vrod_ctx_devices.argtypes = [vp]
        L.vrod_ctx_destroy.argtypes = [vp]
        L.vrod_ctx_destroy.restype = None
        L.vrod_ctx_synchronize.argtypes = [vp]
        L.vrod_ctx_stream.argtypes = [vp]
        L.vrod_ctx_stream.restype = vp
        L.vrod_ctx_stats.argtypes = [vp, C.POINTER(Stats)]
        L.vrod_ctx_profile.argtypes = [vp, i32]
        L.vrod_ctx_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64)]
        L.vrod_ctx_rank.argtypes = [vp]
        L.vrod_ctx_world.argtypes = [vp]
        L.vrod_collection_create.argtypes = [vp, C.c_char_p, u32, i32, u64, C.POINTER(vp)]
        L.vrod_collection_get.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
        L.vrod_collection_drop.argtypes = [vp, C.c_char_p]
        L.vrod_collection_list.argtypes = [vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.vrod_collection_info.argtypes = [vp, C.POINTER(u32), C.POINTER(i32), C.POINTER(u64), C.POINTER(u64)]
        L.vrod_collection_insert.argtypes = [vp, vp, u64, C.POINTER(u64)]
        L.vrod_collection_fill_synthetic.argtypes = [vp, u64, u64]
        L.vrod_collection_read_rows.argtypes = [vp, u64, u64, vp]
        L.vrod_collection_shard.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
        L.vrod_collection_save.argtypes = [vp, C.c_char_p]
        L.vrod_collection_load.argtypes = [vp, C.c_char_p, C.c_char_p, u64, C.POINTER(vp)]
        L.vrod_collection_search.argtypes = [vp, vp, u32, u32, vp, vp]
        L.vrod_collection_search_device.argtypes = [vp, vp, u32, u32, vp, vp]
        L.vrod_collection_set_path.argtypes = [vp, i32]
        L.vrod_last_error.restype = C.c_char_p
        L.vrod_version.restype = C.c_char_p
        _lib = L
    return _lib


def _check(st):
    if st != OK:
        raise VrodError(st, lib().vrod_last_error().decode())


def version():
    return lib().vrod_version().decode()


def comm_unique_id():
    buf = (C.c_ubyte * COMM_ID_BYTES)()
    _check(lib().vrod_comm_unique_id(buf))
    return bytes(buf)


class Collection:
    """A collection handle (this rank's shard)."""

    def __init__(self, ctx, handle, name):
        self.ctx, self.h, self.name = ctx, handle, name

    def info(self):
        dim, metric, count, cap = C.c_uint32(), C.c_int(), C.c_uint64(), C.c_uint64()
        _check(lib().vrod_collection_info(self.h, C.byref(dim), C.byref(metric), C.byref(count), C.byref(cap)))
        return {"dim": dim.value, "metric": metric.value, "count": count.value, "capacity": cap.value}

    def shard(self):
        base, local = C.c_uint64(), C.c_uint64()
        _check(lib().vrod_collection_shard(self.h, C.byref(base), C.byref(local)))
        return base.value, local.value

    def insert(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        n = rows.shape[0] if rows.ndim == 2 else 0
        first = C.c_uint64()
        _check(lib().vrod_collection_insert(self.h, rows.ctypes.data_as(C.c_void_p), n, C.byref(first)))
        return first.value

    def fill_synthetic(self, n, seed):
        _check(lib().vrod_collection_fill_synthetic(self.h, n, seed))

    def read_rows(self, row0, n):
        out = np.empty((n, self.info()["dim"]), dtype=np.float32)
        _check(lib().vrod_collection_read_rows(self.h, row0, n, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_path(self, path):
        _check(lib().vrod_collection_set_path(self.h, path))

    def save(self, path):
        _check(lib().vrod_collection_save(self.h, str(path).encode()))

    def search(self, queries, k):
        """Host buffers in, host buffers out: (ids [b,k] uint64, dist [b,k] float32)."""
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        if queries.ndim == 1:
            queries = queries[None, :]
        b = queries.shape[0]
        ids = np.empty((b, k), dtype=np.uint64)
        dist = np.empty((b, k), dtype=np.float32)
        _check(lib().vrod_collection_search(self.h, queries.ctypes.data_as(C.c_void_p), b, k,
                                            ids.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p)))
        return ids, dist

    def search_device(self, q_ptr, b, k, ids_ptr, dist_ptr):
        """Device pointers (ints); enqueues on the context's stream and returns."""
        _check(lib().vrod_collection_search_device(self.h, q_ptr, b, k, ids_ptr, dist_ptr))


class Context:
    """device: a CUDA ordinal, or a list of ordinals for ONE process driving several GPUs (vrod_ctx_create_multi)."""

    def __init__(self, device=0, rank=0, world=1, comm_id=None):
        self.h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            arr = (C.c_int * len(device))(*device)
            _check(lib().vrod_ctx_create_multi(arr, len(device), C.byref(self.h)))
        elif world > 1:
            cid = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(comm_id)
            _check(lib().vrod_ctx_create_sharded(device, rank, world, cid, C.byref(self.h)))
        else:
            _check(lib().vrod_ctx_create(device, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().vrod_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def rank(self):
        return lib().vrod_ctx_rank(self.h)

    @property
    def world(self):
        return lib().vrod_ctx_world(self.h)

    @property
    def devices(self):
        return lib().vrod_ctx_devices(self.h)

    def stream(self):
        return lib().vrod_ctx_stream(self.h)

    def synchronize(self):
        _check(lib().vrod_ctx_synchronize(self.h))

    def stats(self):
        s = Stats()
        _check(lib().vrod_ctx_stats(self.h, C.byref(s)))
        return s.as_dict()

    def profile(self, enable):
        _check(lib().vrod_ctx_profile(self.h, 1 if enable else 0))

    def profile_read(self):
        """(summed kernel ms, bracketed launches) since the last read."""
        ms, n = C.c_double(), C.c_uint64()
        _check(lib().vrod_ctx_profile_read(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def create(self, name, dim, metric, capacity):
        h = C.c_void_p()
        _check(lib().vrod_collection_create(self.h, name.encode(), dim, metric, capacity, C.byref(h)))
        return Collection(self, h, name)

    def get(self, name):
        h = C.c_void_p()
        _check(lib().vrod_collection_get(self.h, name.encode(), C.byref(h)))
        return Collection(self, h, name)

    def load(self, name, path, capacity=0):
        h = C.c_void_p()
        _check(lib().vrod_collection_load(self.h, name.encode(), str(path).encode(), capacity, C.byref(h)))
        return Collection(self, h, name)

    def drop(self, name):
        _check(lib().vrod_collection_drop(self.h, name.encode()))

    def list(self):
        need = C.c_size_t()
        _check(lib().vrod_collection_list(self.h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _check(lib().vrod_collection_list(self.h, buf, need.value, C.byref(need)))
        s = buf.value.decode()
        return s.split("\n") if s else []
