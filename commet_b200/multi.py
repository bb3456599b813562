"""One process per GPU: the index_and_search chunk loop with the index set sharded over the ranks.

For every index chunk (same chunk plan on every rank -- it is a pure function of the reference set):

    1. each rank zeroes its filter and inserts ITS share of the chunk's reads       (commet_index_add)
    2. barrier                                                                      (all partials complete)
    3. merge: one kernel per rank ORs slice `rank` of every peer's partial over
       NVLink peer memory and pushes the merged slice into every rank's filter      (commet_index_merge)
    4. barrier                                                                      (all pushes complete)
    5. each rank searches its own query sets against the now complete filter        (commet_search_dev)

Bloom insertion is commutative and idempotent, so the merged filter is bit-identical to the single-GPU
filter of the chunk, and every query bit is the single-GPU bit (axes 1 and 2).  No data-path
collective is needed besides the merge; torch.distributed only carries the 64-byte IPC handles and the
barriers.

The loop is written against a small backend protocol so that the sharding/merge logic is testable on CPU
(tests/test_multi_gloo.py runs it with world_size 2 over gloo and an oracle-backed stand-in); the product
backend is `DeviceBackend` over the C-ABI.
"""
from __future__ import annotations

import time
from typing import Protocol, Sequence


def shard_range(first: int, end: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous share of reads [first, end) owned by `rank`: sizes differ by at most one, union = range"""
    n = end - first
    return first + n * rank // world, first + n * (rank + 1) // world


def slice_range(n_vec: int, world: int, rank: int) -> tuple[int, int]:
    """the 16-byte vectors of the filter that `rank` reduces in the merge (same formula as commet_index_merge)"""
    return n_vec * rank // world, n_vec * (rank + 1) // world


class Backend(Protocol):
    def chunk_plan(self, k: int, maxk: int | None) -> list[tuple[int, int]]: ...
    def begin(self, k: int) -> None: ...
    def clear(self) -> None: ...
    def index(self, first: int, count: int) -> None: ...
    def flush(self) -> None: ...
    def merge(self) -> None: ...
    def search(self, k: int, t: int) -> None: ...


class Barrier(Protocol):
    def __call__(self) -> None: ...


def sharded_index_and_search(backend: Backend, barrier: Barrier, world: int, rank: int, k: int, t: int,
                             maxk: int | None = None) -> dict:
    """src/index_and_search.cpp:255-277 with every chunk's reads split over `world` ranks.
    Returns {"chunks": n, "indexed_here": reads this rank inserted, and host-clock seconds spent in the
    (synchronised) index, merge and barrier phases}."""
    plan = backend.chunk_plan(k, maxk)
    backend.begin(k)
    indexed = 0
    t_index = t_merge = t_wait = 0.0
    for ci, (c0, c1) in enumerate(plan):
        if ci:
            backend.clear()
        lo, hi = shard_range(c0, c1, world, rank)
        t0 = time.perf_counter()
        if hi > lo:
            backend.index(lo, hi - lo)
            indexed += hi - lo
        if world > 1:
            backend.flush()    # my partial filter is complete on the device ...
            t1 = time.perf_counter()
            barrier()          # ... and so is everybody else's
            t2 = time.perf_counter()
            backend.merge()
            backend.flush()    # my merged slice has landed in every filter ...
            t3 = time.perf_counter()
            barrier()          # ... and so has everybody else's
            t4 = time.perf_counter()
            t_index += t1 - t0
            t_merge += t3 - t2
            t_wait += (t2 - t1) + (t4 - t3)
        backend.search(k, t)
    return {"chunks": len(plan), "indexed_here": indexed, "index_s": t_index, "merge_s": t_merge, "barrier_s": t_wait}


class DeviceBackend:
    """The product backend: commet_b200.Context + staged streams + peer-mapped filters."""

    def __init__(self, ctx, index_stream, query_streams: Sequence, d_tags: Sequence[int], d_counters: Sequence[int]):
        self.ctx, self.index_stream = ctx, index_stream
        self.queries, self.d_tags, self.d_counters = list(query_streams), list(d_tags), list(d_counters)
        self.peers: list[int] | None = None
        self.rank = 0
        self.k = 0

    # -- peer mapping: exchange the IPC handles of the filters once per k -------------------------
    def connect(self, k: int, world: int, rank: int, all_gather_bytes):
        """all_gather_bytes(b: bytes) -> list[bytes] over the ranks (torch.distributed.all_gather_object)"""
        self.ctx.index_begin(k)
        self.k, self.rank = k, rank
        handles = all_gather_bytes(self.ctx.index_export())
        self.peers = [0 if p == rank else self.ctx.peer_open(handles[p]) for p in range(world)]

    def disconnect(self):
        for p in self.peers or []:
            if p:
                self.ctx.peer_close(p)
        self.peers = None

    # -- Backend ----------------------------------------------------------------------------------
    def chunk_plan(self, k, maxk):
        return self.ctx.chunk_plan(self.index_stream, k, maxk)[0]

    def begin(self, k):
        if self.peers is not None and k != self.k:
            raise ValueError(f"peers were connected for k={self.k}, not k={k}")
        self.ctx.index_begin(k)            # same k: the allocation (and its IPC handle) is kept, only zeroed
        self.k = k

    def clear(self):
        self.ctx.index_begin(self.k)

    def index(self, first, count):
        self.ctx.index_add(self.index_stream, first, count)

    def flush(self):
        self.ctx.sync()

    def merge(self):
        self.ctx.index_merge(self.peers, self.rank)

    def search(self, k, t):
        for q, tg, cn in zip(self.queries, self.d_tags, self.d_counters):
            self.ctx.search_reads_device(q, k, t, tg, cn)
