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

Two placements of the reference set are supported:

    replicated    every rank holds the whole set and inserts its share of each chunk  (sharded_index_and_search)
    distributed   every rank holds (parsed / uploaded / encoded) only ITS blocks of the set, dealt block-cyclically:
                  block b of `block` consecutive reads lives on rank b % world.  A chunk is a contiguous range of
                  global reads, so on every rank it is a contiguous range of LOCAL reads, evenly loaded whatever
                  the chunk.  The stop rule of index_reads needs the k-mer counts of all reads in global order:
                  the ranks exchange their local totals (one number each) and, only if the limit is reached at
                  all, their per-block totals; the read that closes a chunk is resolved by the rank owning its
                  block (distributed_plan, distributed_index_and_search).  Host-to-device traffic and device memory for the reference
                  set are 1/world of the replicated placement.

The loops are written against a small backend protocol so that the sharding/merge logic is testable on CPU
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
    # distributed placement only:
    def local_kmer_total(self, k: int) -> int: ...
    def local_kmer_counts(self, k: int): ...
    def max_kmer(self, k: int) -> int: ...


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


# ----------------------------------------------------------------------------------------------------------------
# distributed placement: block-cyclic shards of the reference set
# ----------------------------------------------------------------------------------------------------------------
DEFAULT_BLOCK = 1 << 16


def local_index(g: int, world: int, rank: int, block: int) -> int:
    """number of reads with global index < g that live on `rank` (= local index of global read g if it is local)"""
    cycle = block * world
    full, o = divmod(g, cycle)
    return full * block + min(max(o - rank * block, 0), block)


def owned_mask(n_global: int, world: int, rank: int, block: int):
    """boolean numpy mask over the global reads: True where the read lives on `rank`"""
    import numpy as np
    return (np.arange(n_global, dtype=np.int64) // block) % world == rank


def shard_stream(bases, offs, world: int, rank: int, block: int):
    """(bases, offs) of the reads of a host stream that live on `rank`, in global order (numpy; what a rank's
    loader delivers: it parses only its own blocks of the files)"""
    import numpy as np
    offs = np.asarray(offs, dtype=np.uint64)
    n = offs.size - 1
    keep = np.nonzero(owned_mask(n, world, rank, block))[0]
    lens = (offs[1:] - offs[:-1])[keep]
    out_offs = np.zeros(keep.size + 1, dtype=np.uint64)
    np.cumsum(lens, out=out_offs[1:])
    out = np.empty(int(out_offs[-1]), dtype=np.uint8)
    # block-wise copies: consecutive owned reads are contiguous in the source
    b = 0
    while b * block < n:
        if b % world == rank:
            g0, g1 = b * block, min(n, (b + 1) * block)
            l0 = local_index(g0, world, rank, block)
            out[int(out_offs[l0]):int(out_offs[l0 + (g1 - g0)])] = bases[int(offs[g0]):int(offs[g1])]
        b += 1
    return out, out_offs


def chunk_bounds(counts, max_kmer: int) -> list[tuple[int, int]]:
    """The stop rule of index_reads on per-read k-mer counts in stream order (include/index_reads.h:48-49,60;
    src/index_and_search.cpp:255): a chunk closes after the read at which the cumulative count reaches max_kmer,
    and the next read is fetched and lost.  Same walk as commet_chunk_plan, vectorised over the prefix sums."""
    import numpy as np
    n = len(counts)
    if n == 0 or max_kmer <= 0:
        return []
    cs = np.cumsum(np.asarray(counts, dtype=np.uint64), dtype=np.uint64)
    if int(cs[-1]) < max_kmer:
        return [(0, n)]
    out, i = [], 0
    while i < n:
        base = int(cs[i - 1]) if i else 0
        j = int(np.searchsorted(cs, np.uint64(base + max_kmer), side="left"))   # first read where cum >= max_kmer
        if j >= n:
            out.append((i, n))
            break
        out.append((i, j + 1))
        i = j + 2                                                               # read j+1 is fetched and lost
    return out


def distributed_plan(local_total: int, local_counts, n_global: int, world: int, rank: int, block: int, max_kmer: int,
                     all_gather):
    """Chunk plan of the whole reference set from every rank's local k-mer counts.  all_gather(obj) -> list of every
    rank's obj.  Fast path: the total never reaches max_kmer (one chunk) and only the totals travel; local_counts()
    (per-read u32 counts of the local shard) is called only otherwise.  Then the ranks exchange their per-BLOCK
    totals (n_global / block numbers in all) and walk them together; the read at which a chunk closes is resolved by
    the rank that owns the block it falls in, from its own per-read counts, and announced to the others -- one or
    two small exchanges per chunk, nothing per read.  Same plan as chunk_bounds() on the global count array."""
    import numpy as np
    totals = all_gather(int(local_total))
    if n_global == 0:
        return []
    if sum(totals) < max_kmer:
        return [(0, n_global)]
    if max_kmer <= 0:
        return []
    counts = np.ascontiguousarray(local_counts(), dtype=np.uint64)
    n_local = local_index(n_global, world, rank, block)
    if len(counts) != n_local:
        raise ValueError(f"rank {rank} holds {len(counts)} reads, its blocks of {n_global} reads are {n_local}")
    n_blocks = (n_global + block - 1) // block
    mine = np.add.reduceat(counts, np.arange(0, n_local, block)) if n_local else np.zeros(0, dtype=np.uint64)
    parts = all_gather(np.ascontiguousarray(mine, dtype=np.uint64))
    T = np.zeros(n_blocks, dtype=np.uint64)
    for r in range(world):
        T[r::world] = parts[r]
    csT = np.cumsum(T, dtype=np.uint64)                    # inclusive prefix over blocks

    def resolve(g0: int, g1: int, cum: int):
        """reads [g0, g1) of ONE block: (index of the read where cum reaches max_kmer or None, k-mers of the range)"""
        if (g0 // block) % world != rank:
            return None
        l0, l1 = local_index(g0, world, rank, block), local_index(g1, world, rank, block)
        cs = np.cumsum(counts[l0:l1], dtype=np.uint64)
        if len(cs) == 0:
            return (None, 0)
        idx = int(np.searchsorted(cs, np.uint64(max_kmer - cum), side="left"))
        return (g0 + idx if idx < len(cs) else None, int(cs[-1]))

    def ask(g0: int, g1: int, cum: int):
        return next(a for a in all_gather(resolve(g0, g1, cum)) if a is not None)

    plan, i = [], 0
    while i < n_global:
        b = i // block
        j, head = ask(i, min((b + 1) * block, n_global), 0)          # from read i to the end of its block
        if j is None:
            # whole blocks after b: the first one in which the running count reaches max_kmer
            target = np.uint64(max_kmer - head) + csT[b]
            b2 = int(np.searchsorted(csT, target, side="left"))
            if b2 >= n_blocks:
                plan.append((i, n_global))                            # the limit is not reached again
                break
            cum = head + int(csT[b2 - 1] - csT[b])
            j, _ = ask(b2 * block, min((b2 + 1) * block, n_global), cum)
        plan.append((i, j + 1))
        i = j + 2                                                     # read j+1 is fetched and lost (index_reads.h:60)
    return plan


def distributed_index_and_search(backend, barrier: Barrier, all_gather, world: int, rank: int, k: int, t: int,
                                 n_global: int, block: int = DEFAULT_BLOCK, maxk: int | None = None) -> dict:
    """src/index_and_search.cpp:255-277 with the reference set dealt block-cyclically over the ranks (the
    backend's index stream is this rank's shard) and every rank searching its own query sets."""
    maxk = backend.max_kmer(k) if maxk is None else maxk
    t0 = time.perf_counter()
    plan = distributed_plan(backend.local_kmer_total(k), lambda: backend.local_kmer_counts(k), n_global, world, rank,
                            block, maxk, all_gather)
    t_plan = time.perf_counter() - t0
    backend.begin(k)
    indexed = 0
    t_index = t_merge = t_wait = 0.0
    for ci, (c0, c1) in enumerate(plan):
        if ci:
            backend.clear()
        lo, hi = local_index(c0, world, rank, block), local_index(c1, world, rank, block)
        t0 = time.perf_counter()
        if hi > lo:
            backend.index(lo, hi - lo)
            indexed += hi - lo
        if world > 1:
            backend.flush()
            t1 = time.perf_counter()
            barrier()
            t2 = time.perf_counter()
            backend.merge()
            backend.flush()
            t3 = time.perf_counter()
            barrier()
            t4 = time.perf_counter()
            t_index += t1 - t0
            t_merge += t3 - t2
            t_wait += (t2 - t1) + (t4 - t3)
        backend.search(k, t)
    return {"chunks": len(plan), "indexed_here": indexed, "plan_s": t_plan, "index_s": t_index, "merge_s": t_merge,
            "barrier_s": t_wait, "plan": plan}


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

    # -- distributed placement: index_stream is this rank's shard -----------------------------------
    def local_kmer_total(self, k):
        return self.ctx.kmer_total(self.index_stream, k)

    def local_kmer_counts(self, k):
        return self.ctx.kmer_counts(self.index_stream, k)

    def max_kmer(self, k):
        from . import api
        return api.max_kmer(k)
