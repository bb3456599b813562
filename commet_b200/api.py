"""ctypes binding of include/commet_b200.h, named after the reference's operators.

    ctx = Context(device=0)
    idx = ctx.stage(bases, offs)                  # Alphabet/HashKey per-base coding on device
    ctx.index_reads(idx, k)                       # include/index_reads.h:41
    found, searched = ctx.search_reads(q, k, t, tags)   # include/search_reads.h:34
    tags, info = ctx.index_and_search(k, t, (ibases, ioffs), [(qbases, qoffs), ...])
    bv, counters = ctx.filter_reads(bases, offs, min_len, max_N, min_shannon, max_reads)
    out = ctx.bvop(BV_AND, a, b); ones = ctx.nb_one(bv, n_bits)

Everything here is plumbing (numpy buffers in, numpy buffers out); the compute
is in commet_b200/csrc.  No function in this module falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import weakref
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
_LIB = PKG / "lib" / "libcommet_b200.so"

BV_AND, BV_OR, BV_ANDNOT, BV_NOT = 0, 1, 2, 3

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)

# every symbol include/commet_b200.h declares: (restype, argtypes)
_SIGS = {
    "commet_last_error": (C.c_char_p, []),
    "commet_abi_version": (C.c_int, []),
    "commet_device_count": (C.c_int, []),
    "commet_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "commet_ctx_destroy": (None, [C.c_void_p]),
    "commet_ctx_sync": (C.c_int, [C.c_void_p]),
    "commet_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "commet_ctx_launches": (C.c_uint64, [C.c_void_p]),
    "commet_ctx_count_probes": (C.c_int, [C.c_void_p, C.c_int]),
    "commet_ctx_binned_index": (C.c_int, [C.c_void_p, C.c_int]),
    "commet_host_alloc": (C.c_void_p, [C.c_size_t]),
    "commet_host_free": (None, [C.c_void_p]),
    "commet_filter_bytes": (C.c_uint64, [C.c_int]),
    "commet_max_kmer": (C.c_uint64, [C.c_int]),
    "commet_reads_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "commet_reads_upload_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "commet_reads_from_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                                           C.POINTER(C.c_void_p)]),
    "commet_reads_clone": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "commet_reads_free": (None, [C.c_void_p]),
    "commet_reads_count": (C.c_uint64, [C.c_void_p]),
    "commet_reads_bases": (C.c_uint64, [C.c_void_p]),
    "commet_reads_select": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "commet_reads_selected": (C.c_uint64, [C.c_void_p]),
    "commet_reads_kmer_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "commet_reads_kmer_total": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _u64p]),
    "commet_chunk_plan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_uint64, _u64p, _u64p]),
    "commet_index_begin": (C.c_int, [C.c_void_p, C.c_int]),
    "commet_index_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "commet_index_filter_ptr": (C.c_void_p, [C.c_void_p]),
    "commet_index_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "commet_index_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]),
    "commet_index_or": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "commet_index_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "commet_peer_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "commet_peer_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "commet_index_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "commet_dist_open": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "commet_dist_index_and_search": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "commet_dist_close": (None, [C.c_void_p]),
    "commet_dist_plan_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                        C.c_uint64, C.c_void_p, C.c_void_p]),
    "commet_dist_deal_regions": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "commet_group_create": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "commet_group_destroy": (None, [C.c_void_p]),
    "commet_group_size": (C.c_int, [C.c_void_p]),
    "commet_group_index_and_search": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                                C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p]),
    "commet_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, _u64p, _u64p]),
    "commet_search_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "commet_index_and_search": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "commet_index_and_search_staged": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "commet_index_and_search_resident": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_int,
                                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_void_p]),
    "commet_filter_reads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_float,
                                      C.c_int64, C.c_void_p, C.c_void_p]),
    "commet_filter_reads_staged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_int64,
                                             C.c_void_p, C.c_void_p]),
    "commet_filter_reads_range": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_float,
                                            C.c_int64, C.c_void_p, C.c_void_p]),
    "commet_filter_reads_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_float,
                                          C.c_int64, C.c_void_p, C.c_void_p]),
    "commet_reads_from_device_filtered": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64,
                                                    C.c_float, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "commet_bvop": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "commet_bv_popcount": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, _u64p]),
    "commet_bvop_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "commet_bv_popcount_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, _u64p]),
    "commet_bv_popcount_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "commet_bench_random_sectors": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


class CommetError(RuntimeError):
    pass


_lib = None


def lib_path() -> Path:
    return _LIB


def load_library() -> C.CDLL:
    """dlopen the CUDA extension; raises if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not _LIB.exists():
            raise CommetError(f"{_LIB} is missing: build it with `python -m commet_b200.build` "
                              "(nvcc, sm_100a). commet_b200 has no CPU fallback.")
        lib = C.CDLL(str(_LIB))
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def filter_bytes(k: int) -> int:
    return int(load_library().commet_filter_bytes(k))


def max_kmer(k: int) -> int:
    return int(load_library().commet_max_kmer(k))


def _ptr(a):
    """host numpy array, raw int (device address) or None -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


def _as_stream(bases, offs):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    if offs.ndim != 1 or offs.size < 1:
        raise ValueError("offs must hold n_reads+1 offsets")
    return bases, offs


_BARRIER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)
_ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)


class _CommStruct(C.Structure):            # commet_comm (include/commet_b200.h)
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("user", C.c_void_p), ("barrier", _BARRIER_FN),
                ("all_gather", _ALLGATHER_FN)]


class Dist:
    """commet_dist*: this rank's end of the multi-GPU chunk loop.  barrier() and all_gather_bytes(b) -> [bytes of every
    rank] are the host language's collectives (torch.distributed in bench.py and the tests)."""

    def __init__(self, ctx: "Context", world: int, rank: int, k: int, barrier, all_gather_bytes):
        self.ctx, self.world, self.rank, self.k = ctx, world, rank, k
        self._err = None

        def _barrier(_user):
            try:
                barrier()
                return 0
            except Exception as e:              # an exception must not unwind through the C frames
                self._err = e
                return -1

        def _all_gather(_user, p_in, p_out, nbytes):
            try:
                parts = all_gather_bytes(C.string_at(p_in, nbytes))
                C.memmove(p_out, b"".join(parts), nbytes * world)
                return 0
            except Exception as e:
                self._err = e
                return -1

        self._cbs = (_BARRIER_FN(_barrier), _ALLGATHER_FN(_all_gather))      # kept alive with the object
        self._comm = _CommStruct(world, rank, None, *self._cbs)
        h = C.c_void_p()
        self._ck(ctx.lib.commet_dist_open(ctx.handle, C.byref(self._comm), k, C.byref(h)))
        self.handle = h.value

    def _ck(self, rc):
        if rc != 0:
            err, self._err = self._err, None
            msg = self.ctx.lib.commet_last_error().decode()
            raise CommetError(msg) from err

    def index_and_search(self, t: int, shard: "ReadStream", n_global: int, queries, d_tags, block: int = 1 << 16,
                         maxk: int | None = None) -> dict:
        """collective: src/index_and_search.cpp:255-277 with the index set dealt block-cyclically over the ranks"""
        maxk = max_kmer(self.k) if maxk is None else maxk
        ns = len(queries)
        qh = (C.c_void_p * max(ns, 1))(*[q.handle for q in queries])
        th = (C.c_void_p * max(ns, 1))(*d_tags)
        searched = np.zeros(max(ns, 1), dtype=np.uint64)
        shared = np.zeros(max(ns, 1), dtype=np.uint64)
        stats = np.zeros(12, dtype=np.uint64)
        self._ck(self.ctx.lib.commet_dist_index_and_search(self.handle, t, maxk, shard.handle, n_global, block, ns,
                                                           C.cast(qh, C.c_void_p), C.cast(th, C.c_void_p), _ptr(searched),
                                                           _ptr(shared), _ptr(stats)))
        return dict(chunks=int(stats[0]), indexed_here=int(stats[1]), plan_s=stats[2] * 1e-9, index_s=stats[3] * 1e-9,
                    search_ns=int(stats[4]), merge_s=stats[5] * 1e-9, barrier_s=stats[6] * 1e-9, tests=int(stats[7]),
                    lookups=int(stats[8]), last_chunk=(int(stats[9]), int(stats[10])),
                    mode="owner-applied records + slice all-gather" if stats[11] else "partial filters + merge",
                    searched=[int(x) for x in searched[:ns]], shared=[int(x) for x in shared[:ns]])

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.commet_dist_close(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _comm_struct(world: int, rank: int, barrier, all_gather_bytes, errbox: list):
    """a commet_comm over two Python callables (kept alive by the returned tuple)"""

    def _barrier(_user):
        try:
            barrier()
            return 0
        except Exception as e:                  # an exception must not unwind through the C frames
            errbox.append(e)
            return -1

    def _all_gather(_user, p_in, p_out, nbytes):
        try:
            parts = all_gather_bytes(C.string_at(p_in, nbytes))
            C.memmove(p_out, b"".join(parts), nbytes * world)
            return 0
        except Exception as e:
            errbox.append(e)
            return -1

    cbs = (_BARRIER_FN(_barrier), _ALLGATHER_FN(_all_gather))
    return _CommStruct(world, rank, None, *cbs), cbs


def dist_plan_host(world: int, rank: int, counts_local, n_global: int, block: int, maxk: int, barrier, all_gather_bytes,
                   cap_chunks: int = 1 << 16):
    """commet_dist_plan_host: the library's multi-rank chunk plan from host-side counts (no device).  Collective.
    Returns ([(first, end), ...], chunk_kmers[world][n_chunks])."""
    lib = load_library()
    counts_local = np.ascontiguousarray(counts_local, dtype=np.uint32)
    errbox: list = []
    comm, _keep = _comm_struct(world, rank, barrier, all_gather_bytes, errbox)
    bounds = np.zeros(2 * cap_chunks, dtype=np.uint64)
    ck = np.zeros(world * cap_chunks, dtype=np.uint64)
    n = C.c_uint64(0)
    rc = lib.commet_dist_plan_host(C.byref(comm), _ptr(counts_local) if counts_local.size else None, counts_local.size, n_global,
                                   block, maxk, _ptr(bounds), cap_chunks, C.byref(n), _ptr(ck))
    if rc != 0:
        raise CommetError(lib.commet_last_error().decode()) from (errbox[0] if errbox else None)
    nc = int(n.value)
    return ([(int(bounds[2 * i]), int(bounds[2 * i + 1])) for i in range(nc)],
            ck[:world * nc].reshape(world, nc).astype(np.int64) if nc else np.zeros((world, 0), dtype=np.int64))


def deal_regions(fills) -> np.ndarray:
    """commet_dist_deal_regions: owner rank of every filter region from fills[world][n_bins] (records per rank and region)"""
    lib = load_library()
    fills = np.ascontiguousarray(fills, dtype=np.uint32)
    world, n_bins = fills.shape
    owner = np.zeros(n_bins, dtype=np.int32)
    if lib.commet_dist_deal_regions(_ptr(fills), world, n_bins, _ptr(owner)) != 0:
        raise CommetError(lib.commet_last_error().decode())
    return owner


class Group:
    """commet_group*: several GPUs of THIS process behind the signature of Context.index_and_search"""

    def __init__(self, devices):
        self.lib = load_library()
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        if self.lib.commet_group_create(C.cast(devs, C.c_void_p), len(devices), C.byref(h)) != 0:
            raise CommetError(self.lib.commet_last_error().decode())
        self.handle = h.value

    def index_and_search(self, k: int, t: int, index_stream, query_streams, maxk: int | None = None):
        maxk = max_kmer(k) if maxk is None else maxk
        ib, io = _as_stream(*index_stream)
        qs = [_as_stream(b, o) for b, o in query_streams]
        ns = len(qs)
        tags = [np.zeros((o.size - 1) // 8 + 1, dtype=np.uint8) for _, o in qs]
        qb = (C.c_void_p * max(ns, 1))(*[b.ctypes.data for b, _ in qs])
        qo = (C.c_void_p * max(ns, 1))(*[o.ctypes.data for _, o in qs])
        tg = (C.c_void_p * max(ns, 1))(*[t_.ctypes.data for t_ in tags])
        nq = np.array([o.size - 1 for _, o in qs] or [0], dtype=np.uint64)
        searched = np.zeros(max(ns, 1), dtype=np.uint64)
        shared = np.zeros(max(ns, 1), dtype=np.uint64)
        stats = np.zeros(8, dtype=np.uint64)
        if self.lib.commet_group_index_and_search(self.handle, k, t, maxk, _ptr(ib), _ptr(io), io.size - 1, ns,
                                                  C.cast(qb, C.c_void_p), C.cast(qo, C.c_void_p), _ptr(nq),
                                                  C.cast(tg, C.c_void_p), _ptr(searched), _ptr(shared), _ptr(stats)) != 0:
            raise CommetError(self.lib.commet_last_error().decode())
        info = dict(chunks=int(stats[0]), indexed=int(stats[1]), index_ns=int(stats[3]), search_ns=int(stats[4]),
                    tests=int(stats[5]), lookups=int(stats[6]), gpus=int(stats[7]),
                    searched=[int(x) for x in searched[:ns]], shared=[int(x) for x in shared[:ns]])
        return tags, info

    def close(self):
        if getattr(self, "handle", None):
            self.lib.commet_group_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ReadStream:
    """A device-resident, 2-bit encoded valid-read stream (commet_reads*)."""

    def __init__(self, ctx: "Context", handle: int):
        self.ctx = ctx
        self.handle = handle
        ctx._streams.add(self)

    @property
    def n_reads(self) -> int:
        return int(self.ctx.lib.commet_reads_count(self.handle))

    @property
    def n_bases(self) -> int:
        return int(self.ctx.lib.commet_reads_bases(self.handle))

    def free(self):
        if self.handle and self.ctx.handle:       # a closed context has already released its streams
            self.ctx.lib.commet_reads_free(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU, one stream (commet_ctx*)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        if self.lib.commet_ctx_create(device, C.byref(h)) != 0:
            raise CommetError(self.lib.commet_last_error().decode())
        self.handle = h.value
        self.device = device
        self._streams = weakref.WeakSet()

    def close(self):
        if getattr(self, "handle", None):
            for r in list(self._streams):         # commet_reads hold a pointer to their context
                r.free()
            self.lib.commet_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != 0:
            raise CommetError(self.lib.commet_last_error().decode())

    # -- plumbing ---------------------------------------------------------
    def sync(self):
        self._ck(self.lib.commet_ctx_sync(self.handle))

    @property
    def stream(self) -> int:
        return int(self.lib.commet_ctx_stream(self.handle) or 0)

    def count_probes(self, on: bool = True):
        """instrumented search: info["tests"], info["lookups"] = the reference's probe counts"""
        self._ck(self.lib.commet_ctx_count_probes(self.handle, int(on)))

    def binned_index(self, mode=1):
        """how filters larger than L2 are fed: 0 direct RED.OR, 1 keys sorted by region first (default), 16..30 region
        passes with 2^mode-byte regions (commet_ctx_binned_index)"""
        self._ck(self.lib.commet_ctx_binned_index(self.handle, int(mode)))

    @property
    def launches(self) -> int:
        return int(self.lib.commet_ctx_launches(self.handle))

    # -- staging ------------------------------------------------------------
    def stage(self, bases, offs) -> ReadStream:
        """Upload + encode a host read stream."""
        bases, offs = _as_stream(bases, offs)
        h = C.c_void_p()
        self._ck(self.lib.commet_reads_upload(self.handle, _ptr(bases), _ptr(offs), offs.size - 1, C.byref(h)))
        return ReadStream(self, h.value)

    def stage_async(self, bases, offs) -> ReadStream:
        """Queue the upload of a host read stream (commet_reads_upload_async): returns at once, the encode happens at
        first use.  The arrays are kept alive by the returned stream object until it is freed."""
        bases, offs = _as_stream(bases, offs)
        h = C.c_void_p()
        self._ck(self.lib.commet_reads_upload_async(self.handle, _ptr(bases), _ptr(offs), offs.size - 1, C.byref(h)))
        rs = ReadStream(self, h.value)
        rs._keep = (bases, offs)
        return rs

    def stage_device(self, d_bases: int, d_offs: int, n_reads: int, n_bases: int) -> ReadStream:
        """Encode a stream whose ASCII bases/offsets already sit in device memory."""
        h = C.c_void_p()
        self._ck(self.lib.commet_reads_from_device(self.handle, _ptr(d_bases), _ptr(d_offs), n_reads, n_bases,
                                                   C.byref(h)))
        return ReadStream(self, h.value)

    def select(self, reads: ReadStream, bv: np.ndarray | None):
        """restrict a staged stream to the reads whose bit is set in `bv` (.bv payload over its records; None = all):
        the input boolean vector of ReadFile (fasta_file.h:143-152)"""
        if bv is not None:
            bv = np.ascontiguousarray(bv, dtype=np.uint8)
            assert bv.size >= reads.n_reads // 8 + 1
        self._ck(self.lib.commet_reads_select(self.handle, reads.handle, _ptr(bv)))

    def kmer_counts(self, reads: ReadStream, k: int) -> np.ndarray:
        out = np.zeros(max(reads.n_reads, 1), dtype=np.uint32)
        self._ck(self.lib.commet_reads_kmer_counts(self.handle, reads.handle, k, _ptr(out)))
        return out[:reads.n_reads]

    def kmer_total(self, reads: ReadStream, k: int) -> int:
        tot = C.c_uint64(0)
        self._ck(self.lib.commet_reads_kmer_total(self.handle, reads.handle, k, C.byref(tot)))
        return int(tot.value)

    def chunk_plan(self, reads: ReadStream, k: int, maxk: int | None = None):
        """[(first, end), ...] read ranges of the index chunks, and reads indexed."""
        maxk = max_kmer(k) if maxk is None else maxk
        cap = 1024
        while True:
            b = np.zeros(2 * cap, dtype=np.uint64)
            nc, ni = C.c_uint64(0), C.c_uint64(0)
            self._ck(self.lib.commet_chunk_plan(self.handle, reads.handle, k, maxk, _ptr(b), cap, C.byref(nc),
                                                C.byref(ni)))
            if nc.value <= cap:
                return [(int(b[2 * i]), int(b[2 * i + 1])) for i in range(nc.value)], int(ni.value)
            cap = int(nc.value)

    # -- stage 1 --------------------------------------------------------------
    def index_begin(self, k: int):
        self._ck(self.lib.commet_index_begin(self.handle, k))

    def index_add(self, reads: ReadStream, first: int = 0, count: int | None = None):
        count = reads.n_reads - first if count is None else count
        self._ck(self.lib.commet_index_add(self.handle, reads.handle, first, count))

    def index_reads(self, reads: ReadStream, k: int, first: int = 0, count: int | None = None):
        """BloomFilter(k) + feed every k-mer of reads[first:first+count] (index_reads.h:41-63 without the stop rule;
        the stop rule is chunk_plan)."""
        self.index_begin(k)
        self.index_add(reads, first, count)

    def filter_download(self, k: int) -> np.ndarray:
        out = np.zeros(filter_bytes(k), dtype=np.uint8)
        self._ck(self.lib.commet_index_download(self.handle, _ptr(out), out.size))
        return out

    def filter_upload(self, k: int, filt: np.ndarray):
        filt = np.ascontiguousarray(filt, dtype=np.uint8)
        self._ck(self.lib.commet_index_upload(self.handle, k, _ptr(filt), filt.size))

    @property
    def filter_ptr(self) -> int:
        return int(self.lib.commet_index_filter_ptr(self.handle) or 0)

    def index_or(self, d_other: int, offset: int, nbytes: int):
        self._ck(self.lib.commet_index_or(self.handle, _ptr(d_other), offset, nbytes))

    # -- multi-GPU merge --------------------------------------------------------
    def index_export(self) -> bytes:
        """CUDA IPC handle (64 bytes) of this context's filter"""
        buf = (C.c_uint8 * 64)()
        self._ck(self.lib.commet_index_export(self.handle, buf))
        return bytes(buf)

    def peer_open(self, handle: bytes) -> int:
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        out = C.c_void_p()
        self._ck(self.lib.commet_peer_open(self.handle, buf, C.byref(out)))
        return int(out.value)

    def peer_close(self, d_filter: int):
        self._ck(self.lib.commet_peer_close(self.handle, _ptr(d_filter)))

    def index_merge(self, d_filters, rank: int):
        """one-kernel OR all-reduce of the partial filters over peer memory (see commet_index_merge)"""
        n = len(d_filters)
        arr = (C.c_void_p * n)(*[C.c_void_p(p or 0) for p in d_filters])
        self._ck(self.lib.commet_index_merge(self.handle, C.cast(arr, C.c_void_p), n, rank))

    # -- stage 2 --------------------------------------------------------------
    def search_reads(self, reads: ReadStream, k: int, t: int, tags: np.ndarray):
        """search_reads.h:34-87 against the current filter; tags (n/8+1 bytes) updated in place.
        Returns (newly found, searched)."""
        assert tags.dtype == np.uint8 and tags.size >= reads.n_reads // 8 + 1
        nf, ns = C.c_uint64(0), C.c_uint64(0)
        self._ck(self.lib.commet_search(self.handle, reads.handle, k, t, _ptr(tags), C.byref(nf), C.byref(ns)))
        return int(nf.value), int(ns.value)

    def search_reads_device(self, reads: ReadStream, k: int, t: int, d_tags: int, d_counters: int):
        self._ck(self.lib.commet_search_dev(self.handle, reads.handle, k, t, _ptr(d_tags), _ptr(d_counters)))

    # -- the chunk loop ---------------------------------------------------------
    def index_and_search(self, k: int, t: int, index_stream, query_streams, maxk: int | None = None):
        """src/index_and_search.cpp:241-277 on host streams.
        Returns (tags: list of .bv payloads, info dict)."""
        maxk = max_kmer(k) if maxk is None else maxk
        ib, io = _as_stream(*index_stream)
        qs = [_as_stream(b, o) for b, o in query_streams]
        ns = len(qs)
        tags = [np.zeros((o.size - 1) // 8 + 1, dtype=np.uint8) for _, o in qs]
        qb = (C.c_void_p * max(ns, 1))(*[b.ctypes.data for b, _ in qs])
        qo = (C.c_void_p * max(ns, 1))(*[o.ctypes.data for _, o in qs])
        tg = (C.c_void_p * max(ns, 1))(*[t_.ctypes.data for t_ in tags])
        nq = np.array([o.size - 1 for _, o in qs] or [0], dtype=np.uint64)
        searched = np.zeros(max(ns, 1), dtype=np.uint64)
        shared = np.zeros(max(ns, 1), dtype=np.uint64)
        stats = np.zeros(8, dtype=np.uint64)
        self._ck(self.lib.commet_index_and_search(self.handle, k, t, maxk, _ptr(ib), _ptr(io), io.size - 1, ns,
                                                  C.cast(qb, C.c_void_p), C.cast(qo, C.c_void_p), _ptr(nq),
                                                  C.cast(tg, C.c_void_p), _ptr(searched), _ptr(shared), _ptr(stats)))
        info = dict(chunks=int(stats[0]), indexed=int(stats[1]), kmers=int(stats[2]), index_ns=int(stats[3]),
                    search_ns=int(stats[4]), tests=int(stats[5]), lookups=int(stats[6]), parts=int(stats[7]),
                    searched=[int(x) for x in searched[:ns]], shared=[int(x) for x in shared[:ns]])
        return tags, info

    def index_and_search_staged(self, k: int, t: int, index: ReadStream, queries, d_tags, maxk: int | None = None):
        """Same loop on staged streams; d_tags: device addresses of zeroed u32 tag words."""
        maxk = max_kmer(k) if maxk is None else maxk
        ns = len(queries)
        qh = (C.c_void_p * max(ns, 1))(*[q.handle for q in queries])
        th = (C.c_void_p * max(ns, 1))(*d_tags)
        searched = np.zeros(max(ns, 1), dtype=np.uint64)
        shared = np.zeros(max(ns, 1), dtype=np.uint64)
        stats = np.zeros(8, dtype=np.uint64)
        self._ck(self.lib.commet_index_and_search_staged(self.handle, k, t, maxk, index.handle, ns,
                                                         C.cast(qh, C.c_void_p), C.cast(th, C.c_void_p),
                                                         _ptr(searched), _ptr(shared), _ptr(stats)))
        return dict(chunks=int(stats[0]), indexed=int(stats[1]), kmers=int(stats[2]), index_ns=int(stats[3]),
                    search_ns=int(stats[4]), tests=int(stats[5]), lookups=int(stats[6]),
                    searched=[int(x) for x in searched[:ns]], shared=[int(x) for x in shared[:ns]])

    # -- stage 3 --------------------------------------------------------------
    def filter_reads(self, bases, offs, min_len=0, max_N=-1, min_shannon=0.0, max_reads=-1):
        """src/filter_reads.cpp:181-205. Returns (.bv payload, counters dict)."""
        bases, offs = _as_stream(bases, offs)
        n = offs.size - 1
        bv = np.zeros(n // 8 + 1, dtype=np.uint8)
        cnt = np.zeros(4, dtype=np.uint64)
        self._ck(self.lib.commet_filter_reads(self.handle, _ptr(bases), _ptr(offs), n, min_len, max_N,
                                              C.c_float(min_shannon), max_reads, _ptr(bv), _ptr(cnt)))
        return bv, dict(rm_length=int(cnt[0]), rm_N=int(cnt[1]), rm_shannon=int(cnt[2]), selected=int(cnt[3]))

    def filter_reads_staged(self, reads: ReadStream, d_bv: int, min_len=0, max_N=-1, min_shannon=0.0, max_reads=-1):
        cnt = np.zeros(4, dtype=np.uint64)
        self._ck(self.lib.commet_filter_reads_staged(self.handle, reads.handle, min_len, max_N, C.c_float(min_shannon),
                                                     max_reads, _ptr(d_bv), _ptr(cnt)))
        return dict(rm_length=int(cnt[0]), rm_N=int(cnt[1]), rm_shannon=int(cnt[2]), selected=int(cnt[3]))

    def filter_reads_range(self, reads: ReadStream, first: int, count: int, min_len=0, max_N=-1, min_shannon=0.0,
                           max_reads=-1):
        """filter_reads on records [first, first+count) of a staged stream (one file of a set staged as a whole)"""
        bv = np.zeros(count // 8 + 1, dtype=np.uint8)
        cnt = np.zeros(4, dtype=np.uint64)
        self._ck(self.lib.commet_filter_reads_range(self.handle, reads.handle, first, count, min_len, max_N,
                                                    C.c_float(min_shannon), max_reads, _ptr(bv), _ptr(cnt)))
        return bv, dict(rm_length=int(cnt[0]), rm_N=int(cnt[1]), rm_shannon=int(cnt[2]), selected=int(cnt[3]))

    def filter_reads_device(self, d_bases: int, d_offs: int, n_reads: int, d_bv: int, min_len=0, max_N=-1, min_shannon=0.0,
                            max_reads=-1):
        """the selection fused into the staging pass, on ASCII bases already resident on the device"""
        cnt = np.zeros(4, dtype=np.uint64)
        self._ck(self.lib.commet_filter_reads_dev(self.handle, _ptr(d_bases), _ptr(d_offs), n_reads, min_len, max_N,
                                                  C.c_float(min_shannon), max_reads, _ptr(d_bv), _ptr(cnt)))
        return dict(rm_length=int(cnt[0]), rm_N=int(cnt[1]), rm_shannon=int(cnt[2]), selected=int(cnt[3]))

    def stage_device_filtered(self, d_bases: int, d_offs: int, n_reads: int, n_bases: int, d_bv: int, min_len=0, max_N=-1,
                              min_shannon=0.0, max_reads=-1):
        """staging and selection fused: (staged stream, counters) from one pass over the device-resident ASCII bases"""
        cnt = np.zeros(4, dtype=np.uint64)
        h = C.c_void_p()
        self._ck(self.lib.commet_reads_from_device_filtered(self.handle, _ptr(d_bases), _ptr(d_offs), n_reads, n_bases, min_len,
                                                            max_N, C.c_float(min_shannon), max_reads, _ptr(d_bv), _ptr(cnt),
                                                            C.byref(h)))
        return ReadStream(self, h.value), dict(rm_length=int(cnt[0]), rm_N=int(cnt[1]), rm_shannon=int(cnt[2]), selected=int(cnt[3]))

    # -- stage 4 --------------------------------------------------------------
    def bvop(self, op: int, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
        """BooleanVector::full_and/or/and_not/not over all payload bytes."""
        a = np.ascontiguousarray(a, dtype=np.uint8)
        if op != BV_NOT:
            b = np.ascontiguousarray(b, dtype=np.uint8)
            if b.size != a.size:
                raise CommetError("Error: the two vectors are not the same size")
        out = np.empty_like(a)
        self._ck(self.lib.commet_bvop(self.handle, op, _ptr(a), _ptr(b), _ptr(out), a.size))
        return out

    def nb_one(self, bv: np.ndarray, n_bits: int) -> int:
        bv = np.ascontiguousarray(bv, dtype=np.uint8)
        assert bv.size >= n_bits // 8 + 1
        ones = C.c_uint64(0)
        self._ck(self.lib.commet_bv_popcount(self.handle, _ptr(bv), n_bits, C.byref(ones)))
        return int(ones.value)

    def bvop_device(self, op: int, d_a: int, d_b: int | None, d_out: int, n_bytes: int):
        self._ck(self.lib.commet_bvop_dev(self.handle, op, _ptr(d_a), _ptr(d_b), _ptr(d_out), n_bytes))

    def nb_one_device(self, d_bv: int, n_bits: int) -> int:
        ones = C.c_uint64(0)
        self._ck(self.lib.commet_bv_popcount_dev(self.handle, _ptr(d_bv), n_bits, C.byref(ones)))
        return int(ones.value)

    def nb_one_device_batch(self, d_bvs, n_bits) -> list:
        """nb_one of several device-resident vectors: one read-back and one synchronisation for all"""
        n = len(d_bvs)
        ptrs = (C.c_void_p * max(n, 1))(*d_bvs)
        bits = np.asarray(n_bits, dtype=np.uint64)
        ones = np.zeros(max(n, 1), dtype=np.uint64)
        self._ck(self.lib.commet_bv_popcount_batch_dev(self.handle, C.cast(ptrs, C.c_void_p), _ptr(bits), n, _ptr(ones)))
        return [int(x) for x in ones[:n]]

    # -- measurement ------------------------------------------------------------
    def random_sector_rate(self, nbytes: int, n_ops: int, atomic: bool = False) -> float:
        """G sectors/s of independent random 32-byte-sector loads (or RED.OR) over a buffer."""
        ns = C.c_double(0)
        self._ck(self.lib.commet_bench_random_sectors(self.handle, nbytes, n_ops, int(atomic), C.byref(ns)))
        return n_ops / ns.value
