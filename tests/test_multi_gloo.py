"""CPU, world_size 2 over gloo: the multi-GPU chunk loop of commet_b200.multi (shard every chunk's reads over
the ranks, merge the partial filters as reduce-slice + push, search locally) gives the single-process oracle's
bits.  The device operations are stood in for by the oracle; the orchestration under test is the product's."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from commet_b200 import multi  # noqa: E402
from oracle import oracle  # noqa: E402
from tests import helpers as H  # noqa: E402


class OracleBackend:
    """multi.Backend over the CPU oracle; merge() mimics commet_index_merge: every rank ORs ITS slice of all
    partials, then the merged slices are exchanged (gloo all_gather stands in for the NVLink loads/stores)."""

    def __init__(self, k, index_stream, query_stream, world, rank):
        self.k, self.index_stream, self.query = k, index_stream, query_stream
        self.world, self.rank = world, rank
        self.filt = None
        self.tags = np.zeros(max(len(query_stream[1]) - 1, 1), dtype=np.uint8)
        self.shared = 0
        self.searched = 0

    def chunk_plan(self, k, maxk):
        maxk = oracle.max_kmer(k) if maxk is None else maxk
        n = len(self.index_stream[1]) - 1
        scratch = np.zeros(oracle.filter_bytes(k), dtype=np.uint8)
        pos, plan = 0, []
        while pos < n:
            nxt, ni, _ = oracle.index_chunk(scratch, k, *self.index_stream, pos, maxk)
            plan.append((pos, pos + ni))
            pos = nxt
        return plan

    def begin(self, k):
        self.filt = np.zeros(oracle.filter_bytes(k), dtype=np.uint8)

    def clear(self):
        self.filt[:] = 0

    def index(self, first, count):
        bases, offs = self.index_stream
        sub = offs[first:first + count + 1]
        oracle.index_chunk(self.filt, self.k, bases[int(sub[0]):int(sub[-1])], (sub - sub[0]).astype(np.uint64), 0, 1 << 62)

    def flush(self):
        pass

    def merge(self):
        mine = torch.from_numpy(self.filt.copy())
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine)                       # "peer loads"
        n = self.filt.size
        # the kernel works on 16-byte vectors; tiny filters (k < 5) are a single padded vector
        n_vec = max(n // 16, 1)
        v0, v1 = multi.slice_range(n_vec, self.world, self.rank)
        lo, hi = min(v0 * 16, n), (n if self.rank == self.world - 1 else min(v1 * 16, n))
        merged = np.zeros(hi - lo, dtype=np.uint8)
        for p in parts:
            merged |= p.numpy()[lo:hi]
        objs = [None] * self.world
        dist.all_gather_object(objs, (lo, hi, merged.tobytes()))
        for l, h, b in objs:                               # "peer stores"
            self.filt[l:h] = np.frombuffer(b, dtype=np.uint8)

    def search(self, k, t):
        r = oracle.search(self.filt, k, t, *self.query, self.tags)
        self.shared += r["found"]
        self.searched = r["searched"]

    # distributed placement: index_stream is this rank's block-cyclic shard
    def local_kmer_counts(self, k):
        bases, offs = self.index_stream
        return np.array([_kmers(bases[int(offs[i]):int(offs[i + 1])], k) for i in range(len(offs) - 1)], dtype=np.uint32)

    def local_kmer_total(self, k):
        return int(self.local_kmer_counts(k).sum())

    def max_kmer(self, k):
        return oracle.max_kmer(k)


def _kmers(seq, k):
    """k-mers index_reads feeds for one read: windows of k consecutive ACGTacgt chars (index_reads.h:52-58)"""
    valid = np.isin(seq, np.frombuffer(b"ACGTacgt", dtype=np.uint8))
    n = run = 0
    for v in valid:
        run = run + 1 if v else 0
        n += run >= k
    return n


def _worker(rank, world, port, k, t, maxk, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        ref = H.make_ref_set(rng, 900, 30, 90, p_N=0.01)
        queries = [H.make_query_set(rng, ref, 500, 30, 90, p_N=0.01) for _ in range(world)]
        be = OracleBackend(k, H.to_stream(ref), H.to_stream(queries[rank]), world, rank)
        info = multi.sharded_index_and_search(be, dist.barrier, world, rank, k, t, maxk)
        np.save(Path(out_dir) / f"tags{rank}.npy", be.tags)
        np.save(Path(out_dir) / f"meta{rank}.npy", np.array([info["chunks"], info["indexed_here"], be.shared, be.searched]))
    finally:
        dist.destroy_process_group()


def _worker_distributed(rank, world, port, k, t, maxk, seed, block, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        ref = H.make_ref_set(rng, 700, 30, 90, p_N=0.01)
        queries = [H.make_query_set(rng, ref, 300, 30, 90, p_N=0.01) for _ in range(world)]
        shard = multi.shard_stream(*H.to_stream(ref), world, rank, block)      # this rank never sees the other blocks
        be = OracleBackend(k, shard, H.to_stream(queries[rank]), world, rank)

        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        info = multi.distributed_index_and_search(be, dist.barrier, all_gather, world, rank, k, t, len(ref), block, maxk)
        np.save(Path(out_dir) / f"tags{rank}.npy", be.tags)
        np.save(Path(out_dir) / f"meta{rank}.npy", np.array([info["chunks"], info["indexed_here"], be.shared, be.searched,
                                                            len(shard[1]) - 1]))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("k,t,maxk,seed", [(13, 2, None, 1), (16, 2, 9000, 2), (11, 1, 4000, 3)])
def test_sharded_chunk_loop_matches_single_process_oracle(tmp_path, k, t, maxk, seed):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), k, t, maxk, seed, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(seed)
    ref = H.make_ref_set(rng, 900, 30, 90, p_N=0.01)
    queries = [H.make_query_set(rng, ref, 500, 30, 90, p_N=0.01) for _ in range(world)]
    e_tags, e = oracle.index_and_search(k, t, H.to_stream(ref), [H.to_stream(q) for q in queries], maxk)
    indexed = 0
    for r in range(world):
        tags = np.load(tmp_path / f"tags{r}.npy")
        chunks, indexed_here, shared, searched = np.load(tmp_path / f"meta{r}.npy").tolist()
        assert np.array_equal(tags[:len(queries[r])], e_tags[r]), f"rank {r}: tags differ from the single-process oracle"
        assert chunks == e["chunks"] and shared == e["shared"][r] and searched == e["searched"][r]
        indexed += indexed_here
    assert indexed == e["indexed"]
    if maxk:
        assert e["chunks"] > 2          # the sharding really was exercised across chunk boundaries


def test_shard_and_slice_ranges_partition():
    for n in (0, 1, 7, 8, 1000, 12345):
        for world in (1, 2, 3, 4, 8):
            got = [multi.shard_range(5, 5 + n, world, r) for r in range(world)]
            assert got[0][0] == 5 and got[-1][1] == 5 + n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            assert max(h - l for l, h in got) - min(h - l for l, h in got) <= 1
            sl = [multi.slice_range(n, world, r) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == n and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))


@pytest.mark.parametrize("world,k,t,maxk,block,seed", [(2, 13, 2, None, 64, 1), (2, 16, 2, 9000, 7, 2), (3, 11, 1, 4000, 50, 3),
                                                      (2, 12, 2, 2500, 1, 4)])
def test_distributed_reference_set_matches_single_process_oracle(tmp_path, world, k, t, maxk, block, seed):
    """The reference set dealt block-cyclically over the ranks (every rank holds only its blocks): global chunk
    boundaries from exchanged k-mer counts, contiguous local ranges per chunk, same bits as one process."""
    mp.spawn(_worker_distributed, args=(world, _free_port(), k, t, maxk, seed, block, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(seed)
    ref = H.make_ref_set(rng, 700, 30, 90, p_N=0.01)
    queries = [H.make_query_set(rng, ref, 300, 30, 90, p_N=0.01) for _ in range(world)]
    e_tags, e = oracle.index_and_search(k, t, H.to_stream(ref), [H.to_stream(q) for q in queries], maxk)
    indexed = held = 0
    for r in range(world):
        tags = np.load(tmp_path / f"tags{r}.npy")
        chunks, indexed_here, shared, searched, n_local = np.load(tmp_path / f"meta{r}.npy").tolist()
        assert np.array_equal(tags[:len(queries[r])], e_tags[r]), f"rank {r}: tags differ from the single-process oracle"
        assert chunks == e["chunks"] and shared == e["shared"][r] and searched == e["searched"][r]
        indexed += indexed_here
        held += n_local
    assert indexed == e["indexed"] and held == len(ref)
    if maxk:
        assert e["chunks"] > 2


def test_block_cyclic_helpers():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        for block in (1, 3, 8, 1000):
            n = 53
            masks = [multi.owned_mask(n, world, r, block) for r in range(world)]
            assert np.array_equal(np.sum(masks, axis=0), np.ones(n))                     # a partition of the reads
            for r in range(world):
                for g in range(n + 1):
                    assert multi.local_index(g, world, r, block) == int(masks[r][:g].sum())
            reads = H.make_ref_set(rng, n, 1, 20)
            bases, offs = H.to_stream(reads)
            for r in range(world):
                b, o = multi.shard_stream(bases, offs, world, r, block)
                got = [bytes(b[int(o[i]):int(o[i + 1])]) for i in range(len(o) - 1)]
                assert got == [x for x, m in zip(reads, masks[r]) if m]


def test_chunk_bounds_is_the_stop_rule_of_index_reads():
    """multi.chunk_bounds against the oracle's chunk walk on the same per-read counts"""
    rng = np.random.default_rng(3)
    for trial in range(40):
        k = int(rng.integers(8, 16))
        reads = H.make_ref_set(rng, int(rng.integers(1, 120)), 5, 60, p_N=0.03)
        counts = [_kmers(np.frombuffer(r, dtype=np.uint8), k) for r in reads]
        maxk = int(rng.choice([1, 7, 50, 400, 10**9]))
        stream = H.to_stream(reads)
        scratch = np.zeros(oracle.filter_bytes(k), dtype=np.uint8)
        pos, plan = 0, []
        while pos < len(reads):
            nxt, ni, _ = oracle.index_chunk(scratch, k, *stream, pos, maxk)
            plan.append((pos, pos + ni))
            pos = nxt
        assert multi.chunk_bounds(counts, maxk) == plan, (trial, k, maxk)


def _plan_in_lockstep(counts, world, block, maxk):
    """multi.distributed_plan on `world` threads exchanging through a barrier (no process group needed)"""
    import threading
    n = len(counts)
    barrier = threading.Barrier(world)
    slots, results, errs = [None] * world, [None] * world, []

    def gather_for(rank):
        def all_gather(obj):
            slots[rank] = obj
            barrier.wait()
            out = list(slots)
            barrier.wait()
            return out
        return all_gather

    def run(rank):
        try:
            loc = np.asarray(counts, dtype=np.uint32)[multi.owned_mask(n, world, rank, block)]
            results[rank] = multi.distributed_plan(int(loc.sum()), lambda: loc, n, world, rank, block, maxk, gather_for(rank))
        except Exception as e:              # pragma: no cover
            errs.append(e)
            barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs
    assert all(r == results[0] for r in results)        # every rank arrives at the same plan
    return results[0]


def test_two_level_distributed_plan_equals_the_global_walk():
    """per-block totals + owner-resolved boundary reads give chunk_bounds() of the global per-read counts, wherever
    the boundaries and the lost reads fall relative to the blocks (block sizes 1..100, 1..4 ranks, zero-count reads)"""
    rng = np.random.default_rng(0)
    for trial in range(400):
        n = int(rng.integers(0, 80))
        world = int(rng.integers(1, 5))
        block = int(rng.choice([1, 2, 3, 7, 16, 100]))
        counts = rng.integers(0, int(rng.choice([1, 3, 30])), size=n)
        if rng.random() < 0.3 and n:
            counts[rng.integers(0, n, size=max(1, n // 4))] = 0
        maxk = int(rng.choice([1, 2, 5, 17, 60, 10 ** 6]))
        assert _plan_in_lockstep(counts, world, block, maxk) == multi.chunk_bounds(counts, maxk), (trial, n, world, block, maxk)


# ----------------------------------------------------------------------------------------------------------------
# the LIBRARY's multi-rank host logic (commet_b200/csrc/capi/dist.inl), not its Python mirror: the chunk plan from counts
# dealt over the ranks (commet_dist_plan_host = the walk commet_dist_index_and_search runs, on host counts) with its
# collectives served by gloo, and the dealing of the filter regions to owner ranks
# ----------------------------------------------------------------------------------------------------------------
def _plan_counts(seed, n):
    """per-read k-mer counts with everything the stop rule trips over: zeros, runs of zeros, a few huge reads"""
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 120, n).astype(np.uint32)
    c[rng.random(n) < 0.15] = 0
    if n > 40:
        z = int(rng.integers(0, n - 30))
        c[z:z + 25] = 0
        c[rng.integers(0, n, 3)] = 5000
    return c


def _worker_plan(rank, world, port, cases, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from commet_b200 import api

        def all_gather_bytes(b):
            t_in = torch.frombuffer(bytearray(b), dtype=torch.uint8)
            outs = [torch.empty(len(b), dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(outs, t_in)
            return [o.numpy().tobytes() for o in outs]

        res = []
        for seed, n, block, maxk in cases:
            counts = _plan_counts(seed, n)
            mine = counts[multi.owned_mask(n, world, rank, block)]
            plan, ck = api.dist_plan_host(world, rank, mine, n, block, maxk, dist.barrier, all_gather_bytes)
            res.append((plan, ck.tolist()))
        import pickle
        (Path(out_dir) / f"plan{rank}.pkl").write_bytes(pickle.dumps(res))
    finally:
        dist.destroy_process_group()


PLAN_CASES = [(1, 900, 64, 10 ** 9), (2, 900, 64, 4000), (3, 700, 7, 9000), (4, 500, 1, 2500), (5, 1000, 50, 1), (6, 64, 64, 300),
              (7, 333, 16, 5000), (8, 2000, 100, 20000), (9, 10, 4, 10 ** 6), (10, 0, 8, 100), (11, 1, 8, 1), (12, 129, 64, 5001)]


@pytest.mark.parametrize("world", [2, 3])
def test_library_chunk_plan_over_gloo_equals_the_global_walk(tmp_path, world):
    """commet_dist_plan_host on every rank's block-cyclic share of the counts = the stop rule walked over all counts
    (index_reads.h:48-49,60), the read after every chunk lost; every rank gets the same plan; the per-rank k-mer
    shares of every chunk are the sums of that rank's counts inside it"""
    import pickle
    mp.spawn(_worker_plan, args=(world, _free_port(), PLAN_CASES, str(tmp_path)), nprocs=world, join=True)
    got = [pickle.loads((tmp_path / f"plan{r}.pkl").read_bytes()) for r in range(world)]
    multi_chunk = 0
    for ci, (seed, n, block, maxk) in enumerate(PLAN_CASES):
        counts = _plan_counts(seed, n)
        want = multi.chunk_bounds(counts, maxk)
        for r in range(world):
            plan, ck = got[r][ci]
            assert plan == want, (world, r, PLAN_CASES[ci], plan[:4], want[:4])
            mask = multi.owned_mask(n, world, r, block)
            for p in range(world):
                pm = multi.owned_mask(n, world, p, block)
                assert ck[p] == [int(counts[a:b][pm[a:b]].sum()) for a, b in want], (world, r, p, PLAN_CASES[ci])
            del mask
        multi_chunk += len(want) > 2
    assert multi_chunk >= 6             # chunk boundaries inside blocks, on block edges, lost reads at block starts


def test_library_plan_single_rank_and_argument_checks():
    from commet_b200 import api
    for seed, n, block, maxk in PLAN_CASES:
        counts = _plan_counts(seed, n)
        plan, ck = api.dist_plan_host(1, 0, counts, n, block, maxk, lambda: None, lambda b: [b])
        assert plan == multi.chunk_bounds(counts, maxk)
        assert ck.tolist() == [[int(counts[a:b].sum()) for a, b in plan]]
    with pytest.raises(api.CommetError, match="holds"):            # a rank that does not hold its blocks
        api.dist_plan_host(2, 0, np.zeros(10, dtype=np.uint32), 100, 8, 50, lambda: None, lambda b: [b, b])
    with pytest.raises(api.CommetError, match="room for"):
        api.dist_plan_host(1, 0, np.full(100, 10, dtype=np.uint32), 100, 8, 10, lambda: None, lambda b: [b], cap_chunks=3)


def test_library_region_dealing_is_a_balanced_partition():
    from commet_b200 import api
    rng = np.random.default_rng(7)
    for world in (1, 2, 3, 4, 8):
        for n_bins in (world, 16, 128, 512):
            if n_bins < world:
                continue
            # the d-key regions are hot (3/4 of its bits are ones): a few regions hold several times the mean
            tot = rng.integers(1000, 3000, n_bins).astype(np.int64)
            tot[rng.integers(0, n_bins, max(1, n_bins // 16))] *= 5
            fills = np.zeros((world, n_bins), dtype=np.uint32)
            for b in range(n_bins):
                fills[:, b] = rng.multinomial(int(tot[b]), np.ones(world) / world)
            owner = api.deal_regions(fills)
            assert owner.shape == (n_bins,) and owner.min() >= 0 and owner.max() < world
            load = np.array([int(tot[owner == p].sum()) for p in range(world)])
            # greedy largest-first: no rank exceeds the mean by more than the largest region
            assert load.max() <= tot.sum() / world + tot.max(), (world, n_bins, load)
            if n_bins >= 8 * world:
                assert load.max() <= 1.1 * tot.sum() / world, (world, n_bins, load)
            again = api.deal_regions(fills)
            assert np.array_equal(owner, again)             # deterministic: every rank computes the same owners
    empty = api.deal_regions(np.zeros((4, 128), dtype=np.uint32))
    assert np.bincount(empty, minlength=4).tolist() == [32, 32, 32, 32]     # empty regions are dealt evenly too
