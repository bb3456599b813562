"""C4-shaped run (BASELINE.json configs[3]): ONE reference set of many chunks against 8 query sets, at 1/2/4/8 GPUs.

    python scripts/bench_c4.py [--ref-reads 500000000] [--query-sets 8] [--query-reads 20000000] [--len 150] [-k 33] [-t 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_c4.py ...

One process per GPU, distributed placement (commet_b200/multi.py): the reference set is dealt block-cyclically, every
rank GENERATES only its own blocks on its device (counter-based: block b is drawn from a generator seeded with b, so
the set is the same whatever N), stages them, and searches its share of the query sets (set s on rank s mod N).  Query
reads: half are copies (half of those reverse-complemented, 1 % substitutions) of reads from a pool of 16 reference
blocks that every rank can regenerate, half are fresh random reads (SURVEY 8d recipe).  At k=33 the 500 M x 150 bp
reference is 5.9e10 k-mers = 59 chunks, 58 reads fetched and lost.

Prints one JSON line on rank 0: chunks, seconds (max over ranks, CUDA-synchronised wall clock around the loop), query
reads/s, per-set shared counts (identical at every N: the cross-N parity check), rank-0 phases.  GPU box only; parity
against the oracle at a size it can run is tests/test_gpu_fullsize.py::test_c4_shape_many_chunks_eight_query_sets."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-reads", type=int, default=500_000_000)
    ap.add_argument("--query-sets", type=int, default=8)
    ap.add_argument("--query-reads", type=int, default=20_000_000)
    ap.add_argument("--len", type=int, default=150, dest="length")
    ap.add_argument("-k", type=int, default=33)
    ap.add_argument("-t", type=int, default=2)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import commet_b200
    from commet_b200 import build, multi
    build.build_lib()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    ctx = commet_b200.Context(local_rank)
    n_ref, L, k, t = args.ref_reads, args.length, args.k, args.t
    BLOCK = multi.DEFAULT_BLOCK
    n_blocks = (n_ref + BLOCK - 1) // BLOCK
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    comp = torch.zeros(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    g = torch.Generator(device=dev)

    def ref_block(b):
        """reads of reference block b (the same on every rank and at every N)"""
        m = min(BLOCK, n_ref - b * BLOCK)
        g.manual_seed(1_000_000 + b)
        return acgt[torch.randint(0, 4, (m, L), generator=g, device=dev)]

    t_gen = time.perf_counter()
    # ---- this rank's shard of the reference set -------------------------------------------------------------
    n_loc = multi.local_index(n_ref, world, rank, BLOCK)
    ref_loc = torch.empty(n_loc * L, dtype=torch.uint8, device=dev)
    pos = 0
    for b in range(rank, n_blocks, world):
        blk = ref_block(b)
        ref_loc[pos * L:(pos + blk.shape[0]) * L] = blk.reshape(-1)
        pos += blk.shape[0]
    assert pos == n_loc
    offs_loc = torch.arange(0, n_loc + 1, dtype=torch.int64, device=dev) * L
    torch.cuda.synchronize()
    idx = ctx.stage_device(ref_loc.data_ptr(), offs_loc.data_ptr(), n_loc, n_loc * L)
    del ref_loc
    torch.cuda.empty_cache()
    # ---- query sets of this rank ----------------------------------------------------------------------------------
    pool_blocks = [int(x) for x in torch.linspace(0, n_blocks - 1, steps=min(16, n_blocks)).round().tolist()]
    pool = torch.cat([ref_block(b) for b in sorted(set(pool_blocks))])          # reads that exist in the reference
    nq = args.query_reads
    offs_q = torch.arange(0, nq + 1, dtype=torch.int64, device=dev) * L
    my_sets = list(range(rank, args.query_sets, world))
    queries, tags, counters = [], [], []
    for s in my_sets:
        g.manual_seed(2000 + s)
        q = torch.empty((nq, L), dtype=torch.uint8, device=dev)
        step = 1 << 20
        for s0 in range(0, nq, step):
            m = min(step, nq - s0)
            cp = pool[torch.randint(0, pool.shape[0], (m,), generator=g, device=dev)]
            rc = torch.rand(m, generator=g, device=dev) < 0.5
            cp = torch.where(rc[:, None], comp[cp.flip(1).long()], cp)
            mut = torch.rand((m, L), generator=g, device=dev) < 0.01
            cp = torch.where(mut, acgt[torch.randint(0, 4, (m, L), generator=g, device=dev)], cp)
            shared = torch.rand(m, generator=g, device=dev) < 0.5
            fresh = acgt[torch.randint(0, 4, (m, L), generator=g, device=dev)]
            q[s0:s0 + m] = torch.where(shared[:, None], cp, fresh)
        torch.cuda.synchronize()
        queries.append(ctx.stage_device(q.data_ptr(), offs_q.data_ptr(), nq, nq * L))
        del q
        tags.append(torch.zeros((nq // 8 + 1 + 3) // 4, dtype=torch.int32, device=dev))
        counters.append(torch.zeros(4, dtype=torch.int64, device=dev))
    del pool
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen

    be = multi.DeviceBackend(ctx, idx, queries, [x.data_ptr() for x in tags], [x.data_ptr() for x in counters])

    def all_gather(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:
        be.connect(k, world, rank, all_gather)
    barrier()
    t0 = time.perf_counter()
    info = multi.distributed_index_and_search(be, barrier, all_gather, world, rank, k, t, n_ref, BLOCK)
    ctx.sync()
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt)
    mine = {s: [int(c[0]), ctx.nb_one_device(tg.data_ptr(), nq)] for s, c, tg in zip(my_sets, counters, tags)}
    allsets = {}
    for part in all_gather(mine):
        allsets.update(part)
    indexed = sum(all_gather(info["indexed_here"]))
    if rank == 0:
        res = {"workload": f"C4 shape: 1 reference set of {n_ref} reads x {L} bp against {args.query_sets} query sets of {nq} reads, "
                           f"k={k} t={t}, reference dealt block-cyclically over {world} GPU(s)",
               "n_gpus": world, "chunks": info["chunks"], "reads_lost_at_chunk_boundaries": n_ref - indexed, "seconds": round(dt, 3),
               "query_reads_per_s": args.query_sets * nq / dt, "generate_s": round(t_gen, 2),
               "phases_s_rank0": {key: round(info[key], 3) for key in ("plan_s", "index_s", "merge_s", "barrier_s")},
               "shared_per_set": {str(s): allsets[s][0] for s in sorted(allsets)},
               "ones_per_set": {str(s): allsets[s][1] for s in sorted(allsets)}}
        assert all(v[0] == v[1] for v in allsets.values()), "shared counters and tag popcounts disagree"
        print(json.dumps(res))
        if args.out:
            Path(args.out).write_text(json.dumps(res, indent=1) + "\n")
    if world > 1:
        be.disconnect()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
