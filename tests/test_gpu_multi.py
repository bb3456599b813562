"""GPU, 2 ranks (needs >= 2 devices; run with `gpurun --gpus 2`): the sharded chunk loop with the one-kernel
OR all-reduce over CUDA-IPC peer memory gives the single-process oracle's tags, chunk by chunk."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import oracle  # noqa: E402
from tests import helpers as H  # noqa: E402

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, k, t, maxk, seed, out_dir):
    import faulthandler
    faulthandler.enable()
    import torch
    import torch.distributed as dist
    import commet_b200
    from commet_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(rank)
        ctx = commet_b200.Context(rank)
        rng = np.random.default_rng(seed)
        ref = H.make_ref_set(rng, 3000, 40, 120, p_N=0.01)
        queries = [H.make_query_set(rng, ref, 1500, 40, 120, p_N=0.01) for _ in range(world)]
        idx = ctx.stage(*H.to_stream(ref))
        q = ctx.stage(*H.to_stream(queries[rank]))
        nq = len(queries[rank])
        tags = torch.zeros((nq // 8 + 1 + 3) // 4, dtype=torch.int32, device=f"cuda:{rank}")
        counters = torch.zeros(4, dtype=torch.int64, device=f"cuda:{rank}")
        torch.cuda.synchronize()
        be = multi.DeviceBackend(ctx, idx, [q], [tags.data_ptr()], [counters.data_ptr()])

        def all_gather_bytes(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        be.connect(k, world, rank, all_gather_bytes)
        info = multi.sharded_index_and_search(be, dist.barrier, world, rank, k, t, maxk)
        ctx.sync()
        filt = ctx.filter_download(k)           # the LAST chunk's merged filter: must be identical on every rank
        dist.barrier()
        be.disconnect()
        np.save(Path(out_dir) / f"tags{rank}.npy", tags.cpu().numpy().view(np.uint8)[:nq // 8 + 1])
        np.save(Path(out_dir) / f"meta{rank}.npy", np.array([info["chunks"], info["indexed_here"], int(counters[0]), int(counters[1])]))
        np.save(Path(out_dir) / f"filt{rank}.npy", filt)
        ctx.close()
    finally:
        dist.destroy_process_group()


def _worker_distributed(rank, world, port, k, t, maxk, seed, block, n_dev, out_dir):
    """reference set dealt block-cyclically: every rank uploads (asynchronously) only its shard, then its query set"""
    import faulthandler
    faulthandler.enable()
    import torch
    import torch.distributed as dist
    import commet_b200
    from commet_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = rank % n_dev                       # n_dev == 1: both ranks share the GPU (IPC mapping within one device)
        torch.cuda.set_device(dev)
        ctx = commet_b200.Context(dev)
        rng = np.random.default_rng(seed)
        ref = H.make_ref_set(rng, 3000, 40, 120, p_N=0.01)
        queries = [H.make_query_set(rng, ref, 1500, 40, 120, p_N=0.01) for _ in range(world)]
        shard = multi.shard_stream(*H.to_stream(ref), world, rank, block)
        idx = ctx.stage_async(*shard)
        q = ctx.stage_async(*H.to_stream(queries[rank]))
        nq = len(queries[rank])
        tags = torch.zeros((nq // 8 + 1 + 3) // 4, dtype=torch.int32, device=f"cuda:{dev}")
        counters = torch.zeros(4, dtype=torch.int64, device=f"cuda:{dev}")
        torch.cuda.synchronize()

        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        # the product loop: commet_dist_* in the library, torch.distributed only behind its two callbacks
        d = commet_b200.Dist(ctx, world, rank, k, dist.barrier, all_gather)
        info = d.index_and_search(t, idx, len(ref), [q], [tags.data_ptr()], block=block, maxk=maxk)
        counters[0], counters[1] = info["shared"][0], info["searched"][0]
        info["plan"] = [info["last_chunk"]]
        ctx.sync()
        filt = ctx.filter_download(k)
        dist.barrier()
        d.close()
        np.save(Path(out_dir) / f"tags{rank}.npy", tags.cpu().numpy().view(np.uint8)[:nq // 8 + 1])
        np.save(Path(out_dir) / f"meta{rank}.npy", np.array([info["chunks"], info["indexed_here"], int(counters[0]), int(counters[1]),
                                                            idx.n_reads, info["plan"][-1][0], info["plan"][-1][1]]))
        np.save(Path(out_dir) / f"filt{rank}.npy", filt)
        ctx.close()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("k,t,maxk,seed", [(16, 2, None, 1), (20, 2, 60000, 2), (29, 2, 90000, 3)])
def test_sharded_index_merge_search_two_gpus(tmp_path, k, t, maxk, seed):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), k, t, maxk, seed, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(seed)
    ref = H.make_ref_set(rng, 3000, 40, 120, p_N=0.01)
    queries = [H.make_query_set(rng, ref, 1500, 40, 120, p_N=0.01) for _ in range(world)]
    e_tags, e = oracle.index_and_search(k, t, H.to_stream(ref), [H.to_stream(q) for q in queries], maxk)
    indexed = 0
    for r in range(world):
        tags = np.load(tmp_path / f"tags{r}.npy")
        chunks, indexed_here, shared, searched = np.load(tmp_path / f"meta{r}.npy").tolist()
        assert np.array_equal(tags, oracle.tags_to_bv(e_tags[r])), f"rank {r}: tags differ from the oracle"
        assert chunks == e["chunks"] and shared == e["shared"][r] and searched == e["searched"][r]
        indexed += indexed_here
    assert indexed == e["indexed"]
    assert np.array_equal(np.load(tmp_path / "filt0.npy"), np.load(tmp_path / "filt1.npy"))
    if maxk:
        assert e["chunks"] >= 2


@pytest.mark.skipif(_n_gpus() < 1, reason="needs a GPU")
@pytest.mark.parametrize("world,k,t,maxk,block,seed", [(2, 16, 2, None, 64, 1), (2, 20, 2, 60000, 100, 2), (2, 29, 2, 90000, 7, 3),
                                                       (3, 20, 2, 60000, 33, 4), (4, 20, 1, 50000, 50, 5), (4, 29, 2, 90000, 16, 6),
                                                       (8, 16, 2, None, 64, 7), (8, 20, 2, 60000, 25, 8), (8, 29, 2, 90000, 16, 9), (2, 31, 1, 70000, 40, 10),
                                                       (4, 29, 2, 30000, 1000, 19)])      # blocks larger than chunks: ranks without reads in a chunk, one without any
def test_distributed_reference_set_ranks(tmp_path, world, k, t, maxk, block, seed):
    """2, 3, 4 and 8 ranks (one GPU each while the box has them, else sharing GPUs: the IPC mapping is per process), each
    holding only its block-cyclic shard of the reference set, uploaded with commet_reads_upload_async: global chunk
    plan from exchanged k-mer counts, one-kernel merge k_merge_peers<world> over IPC-mapped filters, tags of the
    single-process oracle; the merged filter of the last chunk is the same on every rank and equal to the oracle's
    filter of that chunk."""
    import torch.multiprocessing as mp
    mp.spawn(_worker_distributed, args=(world, _free_port(), k, t, maxk, seed, block, min(_n_gpus(), world), str(tmp_path)),
             nprocs=world, join=True)
    rng = np.random.default_rng(seed)
    ref = H.make_ref_set(rng, 3000, 40, 120, p_N=0.01)
    queries = [H.make_query_set(rng, ref, 1500, 40, 120, p_N=0.01) for _ in range(world)]
    e_tags, e = oracle.index_and_search(k, t, H.to_stream(ref), [H.to_stream(q) for q in queries], maxk)
    indexed = held = 0
    for r in range(world):
        tags = np.load(tmp_path / f"tags{r}.npy")
        chunks, indexed_here, shared, searched, n_local, c0, c1 = np.load(tmp_path / f"meta{r}.npy").tolist()
        assert np.array_equal(tags, oracle.tags_to_bv(e_tags[r])), f"rank {r}: tags differ from the oracle"
        assert chunks == e["chunks"] and shared == e["shared"][r] and searched == e["searched"][r]
        indexed += indexed_here
        held += n_local
    assert indexed == e["indexed"] and held == len(ref)
    f0 = np.load(tmp_path / "filt0.npy")
    for r in range(1, world):
        assert np.array_equal(f0, np.load(tmp_path / f"filt{r}.npy")), f"rank {r}: merged filter differs from rank 0's"
    # the single-rank filter of the last chunk (reads [c0, c1) of the global stream), built by the oracle
    bases, offs = H.to_stream(ref)
    offs = np.asarray(offs, dtype=np.uint64)
    exp = np.zeros(oracle.filter_bytes(k), dtype=np.uint8)
    sub = offs[c0:c1 + 1]
    oracle.index_chunk(exp, k, np.asarray(bases)[int(sub[0]):int(sub[-1])], (sub - sub[0]).astype(np.uint64), 0, 1 << 62)
    assert np.array_equal(f0, exp)
    if maxk:
        assert e["chunks"] >= 2


@pytest.mark.skipif(_n_gpus() < 1, reason="needs a GPU")
@pytest.mark.parametrize("world,k,t,maxk,block,seed", [(2, 16, 2, None, 64, 11), (2, 20, 2, 60000, 100, 12), (3, 29, 2, 90000, 7, 13),
                                                       (4, 20, 0, 50000, 33, 14), (8, 18, 3, 40000, 16, 15), (5, 3, 1, None, 9, 16), (8, 30, 2, 80000, 21, 17),
                                                       (4, 28, 2, None, 64, 18)])
def test_group_of_one_process_equals_the_oracle(monkeypatch, world, k, t, maxk, block, seed):
    """commet_group_index_and_search: a thread per device (the ranks share GPUs when the box has fewer), host buffers in,
    tag vectors out -- the signature of commet_index_and_search.  Several query sets of different sizes (slices cut at
    multiples of 32 reads, sets smaller than that land on the last rank), chunk boundaries inside blocks."""
    import commet_b200
    monkeypatch.setenv("COMMET_B200_DIST_BLOCK", str(block))
    rng = np.random.default_rng(seed)
    ref = H.make_ref_set(rng, 3000, max(1, k - 5), 120, p_N=0.01)
    queries = [H.make_query_set(rng, ref, n, max(1, k - 5), 120, p_N=0.01) for n in (1500, 17, 333, 64)]
    e_tags, e = oracle.index_and_search(k, t, H.to_stream(ref), [H.to_stream(q) for q in queries], maxk)
    n_dev = _n_gpus()
    g = commet_b200.Group([r % n_dev for r in range(world)])
    try:
        for _ in range(2):                       # the group's contexts are reused by the next call
            tags, info = g.index_and_search(k, t, H.to_stream(ref), [H.to_stream(q) for q in queries], maxk)
            for s in range(len(queries)):
                assert np.array_equal(tags[s], oracle.tags_to_bv(e_tags[s])), (s, world)
            assert info["chunks"] == e["chunks"] and info["indexed"] == e["indexed"]
            assert info["shared"] == e["shared"] and info["searched"] == e["searched"]
            assert info["gpus"] == world
    finally:
        g.close()
    if maxk:
        assert e["chunks"] >= 2


# ---- C4 twin: BASELINE.json configs[3] at a size the oracle can follow ----------------------------------------
C4_TWIN = dict(n_ref=2_000_000, n_sets=8, nq=50_000, L=150, k=27, t=2)      # 2.5e8 k-mers = 16 chunks at k=27, 15 reads lost


def _c4_worker(rank, world, port, n_dev, out_dir):
    import faulthandler
    faulthandler.enable()
    import importlib.util
    import torch
    import torch.distributed as dist
    import commet_b200
    spec = importlib.util.spec_from_file_location("bench_c4", ROOT / "scripts" / "bench_c4.py")
    bench_c4 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench_c4)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev_i = rank % n_dev
        torch.cuda.set_device(dev_i)
        dev = torch.device("cuda", dev_i)
        ctx = commet_b200.Context(dev_i)

        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        p = C4_TWIN
        keep = {}
        info, res = bench_c4.run_c4(torch, ctx, dev, world, rank, dist.barrier, all_gather, p["n_ref"], p["n_sets"], p["nq"], p["L"],
                                    p["k"], p["t"], keep=keep)
        np.save(Path(out_dir) / f"shard{rank}.npy", keep["shard"])
        for s, (shared, tg) in res.items():
            np.save(Path(out_dir) / f"query{s}.npy", keep[f"query{s}"])
            np.save(Path(out_dir) / f"tags{s}.npy", tg.cpu().numpy().view(np.uint8)[:p["nq"] // 8 + 1])
            np.save(Path(out_dir) / f"meta{s}.npy", np.array([info["chunks"], shared, info["searched"][list(res).index(s)]]))
        np.save(Path(out_dir) / f"indexed{rank}.npy", np.array([info["indexed_here"]]))
        ctx.close()
    finally:
        dist.destroy_process_group()


_c4_oracle_cache = {}


@pytest.mark.skipif(_n_gpus() < 1, reason="needs a GPU")
@pytest.mark.parametrize("world", [1, 2, 4])
def test_c4_twin_every_vector_equals_the_oracle(tmp_path, world):
    """The C4 workload of scripts/bench_c4.py (same generator, same library loop) at 1/250 of the reference set and
    k=27, so that the stop rule still cuts it into 16 chunks: every query set's vector, the chunk count, the lost reads
    and the counters equal the oracle's, with the reference set on one rank and dealt over 2 and 4 ranks."""
    import torch.multiprocessing as mp
    from commet_b200 import multi
    p = C4_TWIN
    mp.spawn(_c4_worker, args=(world, _free_port(), min(_n_gpus(), world), str(tmp_path)), nprocs=world, join=True)
    # the global reference stream from the ranks' shards (block b lives on rank b % world)
    L, n_ref, block = p["L"], p["n_ref"], multi.DEFAULT_BLOCK
    ref = np.empty(n_ref * L, dtype=np.uint8)
    shards = [np.load(tmp_path / f"shard{r}.npy") for r in range(world)]
    for b in range((n_ref + block - 1) // block):
        g0, g1 = b * block, min(n_ref, (b + 1) * block)
        r = b % world
        l0 = multi.local_index(g0, world, r, block)
        ref[g0 * L:g1 * L] = shards[r][l0 * L:(l0 + g1 - g0) * L]
    queries = [np.load(tmp_path / f"query{s}.npy").reshape(-1) for s in range(p["n_sets"])]
    key = hash(ref[::997].tobytes())
    if key not in _c4_oracle_cache:                  # the generator does not depend on the number of ranks
        offs = np.arange(n_ref + 1, dtype=np.uint64) * L
        qoffs = np.arange(p["nq"] + 1, dtype=np.uint64) * L
        _c4_oracle_cache[key] = oracle.index_and_search(p["k"], p["t"], (ref, offs), [(q, qoffs) for q in queries])
    e_tags, e = _c4_oracle_cache[key]
    assert e["chunks"] == 16
    indexed = sum(int(np.load(tmp_path / f"indexed{r}.npy")[0]) for r in range(world))
    assert indexed == e["indexed"] == n_ref - 15
    for s in range(p["n_sets"]):
        chunks, shared, searched = np.load(tmp_path / f"meta{s}.npy").tolist()
        assert np.array_equal(np.load(tmp_path / f"tags{s}.npy"), oracle.tags_to_bv(e_tags[s])), f"set {s}: vector differs from the oracle's"
        assert chunks == e["chunks"] and shared == e["shared"][s] and searched == e["searched"][s]
        assert 0.4 < shared / p["nq"] < 0.8
