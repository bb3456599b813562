"""C4-shaped run (BASELINE.json configs[3]): ONE reference set of many chunks against 8 query sets, at 1/2/4/8 GPUs.

    python scripts/bench_c4.py [--ref-reads 500000000] [--query-sets 8] [--query-reads 20000000] [--len 150] [-k 33] [-t 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_c4.py ...

One process per GPU, the library's multi-GPU chunk loop (commet_dist_index_and_search): the reference set is dealt
block-cyclically, every rank GENERATES only its own blocks on its device (counter-based: block b is drawn from a
generator seeded with b, so the set is the same whatever N), stages them, and searches its share of the query sets
(set s on rank s mod N).  Query reads: half are copies (half of those reverse-complemented, 1 % substitutions) of reads
from a pool of 16 reference blocks that every rank can regenerate, half are fresh random reads (SURVEY 8d recipe).  At
k=33 the 500 M x 150 bp reference is 5.9e10 k-mers = 59 chunks, 58 reads fetched and lost.

Prints one JSON line on rank 0: chunks, seconds (max over ranks, CUDA-synchronised wall clock around the loop), query
reads/s, per-set shared counts (identical at every N: the cross-N parity check), rank-0 phases.  GPU box only; parity
against the oracle at a size it can run: tests/test_gpu_fullsize.py::test_c4_twin_* (same generator, same loop)."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

BLOCK = 1 << 16


class C4Generator:
    """the synthetic C4 sets, block by block, on one device"""

    def __init__(self, torch, dev, n_ref, length, block=BLOCK):
        self.torch, self.dev, self.n_ref, self.L, self.block = torch, dev, n_ref, length, block
        self.n_blocks = (n_ref + block - 1) // block
        self.acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
        self.comp = torch.zeros(256, dtype=torch.uint8, device=dev)
        for a, b in zip(b"ACGT", b"TGCA"):
            self.comp[a] = b
        self.g = torch.Generator(device=dev)

    def ref_block(self, b):
        """reads of reference block b (the same on every rank and at every N)"""
        m = min(self.block, self.n_ref - b * self.block)
        self.g.manual_seed(1_000_000 + b)
        return self.acgt[self.torch.randint(0, 4, (m, self.L), generator=self.g, device=self.dev)]

    def shard(self, world, rank):
        """(ASCII bases, offsets, n_local) of the blocks b = rank (mod world), in global order"""
        from commet_b200 import multi
        torch = self.torch
        n_loc = multi.local_index(self.n_ref, world, rank, self.block)
        out = torch.empty(n_loc * self.L, dtype=torch.uint8, device=self.dev)
        pos = 0
        for b in range(rank, self.n_blocks, world):
            blk = self.ref_block(b)
            out[pos * self.L:(pos + blk.shape[0]) * self.L] = blk.reshape(-1)
            pos += blk.shape[0]
        assert pos == n_loc
        offs = torch.arange(0, n_loc + 1, dtype=torch.int64, device=self.dev) * self.L
        return out, offs, n_loc

    def pool(self):
        torch = self.torch
        blocks = [int(x) for x in torch.linspace(0, self.n_blocks - 1, steps=min(16, self.n_blocks)).round().tolist()]
        return torch.cat([self.ref_block(b) for b in sorted(set(blocks))])          # reads that exist in the reference

    def query_set(self, s, nq, pool):
        torch, g, L = self.torch, self.g, self.L
        g.manual_seed(2000 + s)
        q = torch.empty((nq, L), dtype=torch.uint8, device=self.dev)
        step = 1 << 20
        for s0 in range(0, nq, step):
            m = min(step, nq - s0)
            cp = pool[torch.randint(0, pool.shape[0], (m,), generator=g, device=self.dev)]
            rc = torch.rand(m, generator=g, device=self.dev) < 0.5
            cp = torch.where(rc[:, None], self.comp[cp.flip(1).long()], cp)
            mut = torch.rand((m, L), generator=g, device=self.dev) < 0.01
            cp = torch.where(mut, self.acgt[torch.randint(0, 4, (m, L), generator=g, device=self.dev)], cp)
            shared = torch.rand(m, generator=g, device=self.dev) < 0.5
            fresh = self.acgt[torch.randint(0, 4, (m, L), generator=g, device=self.dev)]
            q[s0:s0 + m] = torch.where(shared[:, None], cp, fresh)
        return q


def run_c4(torch, ctx, dev, world, rank, barrier, all_gather_bytes, n_ref, n_sets, nq, L, k, t, block=BLOCK, maxk=None,
           keep=None):
    """stage this rank's part of the workload, run the library's loop, return (info, {set: (shared, tag words tensor)})"""
    import commet_b200
    gen = C4Generator(torch, dev, n_ref, L, block)
    t_gen = time.perf_counter()
    ref_loc, offs_loc, n_loc = gen.shard(world, rank)
    torch.cuda.synchronize()
    idx = ctx.stage_device(ref_loc.data_ptr(), offs_loc.data_ptr(), n_loc, n_loc * L)
    if keep is not None:
        keep["shard"] = ref_loc.cpu().numpy()
    del ref_loc
    torch.cuda.empty_cache()
    pool = gen.pool()
    offs_q = torch.arange(0, nq + 1, dtype=torch.int64, device=dev) * L
    my_sets = list(range(rank, n_sets, world))
    queries, tags = [], []
    for s in my_sets:
        q = gen.query_set(s, nq, pool)
        torch.cuda.synchronize()
        queries.append(ctx.stage_device(q.data_ptr(), offs_q.data_ptr(), nq, nq * L))
        if keep is not None:
            keep[f"query{s}"] = q.cpu().numpy()
        del q
        tags.append(torch.zeros((nq // 8 + 1 + 3) // 4, dtype=torch.int32, device=dev))
    del pool
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    d = commet_b200.Dist(ctx, world, rank, k, barrier, all_gather_bytes)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = d.index_and_search(t, idx, n_ref, queries, [x.data_ptr() for x in tags], block=block, maxk=maxk)
    ctx.sync()
    barrier()
    info["seconds"] = time.perf_counter() - t0
    info["generate_s"] = t_gen
    d.close()
    return info, {s: (info["shared"][i], tags[i]) for i, s in enumerate(my_sets)}


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

    import pickle
    import torch
    import torch.distributed as dist
    import commet_b200
    from commet_b200 import build
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
    n_ref, L, k, t, nq = args.ref_reads, args.length, args.k, args.t, args.query_reads

    ctl = dist.new_group(backend="gloo") if world > 1 else None      # host-side control messages of a few bytes

    def all_gather_bytes(b):
        if world == 1:
            return [b]
        t_in = torch.frombuffer(bytearray(b), dtype=torch.uint8)
        outs = [torch.empty(len(b), dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(outs, t_in, group=ctl)
        return [o.numpy().tobytes() for o in outs]

    def all_gather_obj(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    def barrier():
        if world > 1:
            dist.barrier(group=ctl)

    info, res = run_c4(torch, ctx, dev, world, rank, barrier, all_gather_bytes, n_ref, args.query_sets, nq, L, k, t)
    dt = info["seconds"]
    if world > 1:
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt)
    mine = {s: [int(sh), ctx.nb_one_device(tg.data_ptr(), nq)] for s, (sh, tg) in res.items()}
    allsets = {}
    for part in all_gather_obj(mine):
        allsets.update(part)
    indexed = sum(all_gather_obj(info["indexed_here"]))
    phases = all_gather_obj({key: round(info[key], 3) for key in ("plan_s", "index_s", "merge_s", "barrier_s")} |
                            {"search_s": round(info["search_ns"] * 1e-9, 3)})
    if rank == 0:
        out = {"workload": f"C4 shape: 1 reference set of {n_ref} reads x {L} bp against {args.query_sets} query sets of {nq} reads, "
                           f"k={k} t={t}, reference dealt block-cyclically over {world} GPU(s)",
               "n_gpus": world, "chunks": info["chunks"], "reads_lost_at_chunk_boundaries": n_ref - indexed, "seconds": round(dt, 3),
               "query_reads_per_s": args.query_sets * nq / dt, "generate_s": round(info["generate_s"], 2),
               "phases_s_rank0": phases[0], "phases_s_slowest_search_rank": max(phases, key=lambda p: p["search_s"]),
               "shared_per_set": {str(s): allsets[s][0] for s in sorted(allsets)},
               "ones_per_set": {str(s): allsets[s][1] for s in sorted(allsets)}}
        assert all(v[0] == v[1] for v in allsets.values()), "shared counters and tag popcounts disagree"
        print(json.dumps(out))
        if args.out:
            Path(args.out).write_text(json.dumps(out, indent=1) + "\n")
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
