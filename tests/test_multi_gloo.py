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
