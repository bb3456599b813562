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
                                                       (8, 16, 2, None, 64, 7), (8, 20, 2, 60000, 25, 8)])
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
                                                       (4, 20, 0, 50000, 33, 14), (8, 18, 3, 40000, 16, 15), (5, 3, 1, None, 9, 16)])
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
