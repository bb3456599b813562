"""C2 through commet_group_index_and_search: ONE process, a thread per GPU (what the drop-in index_and_search tool runs with
COMMET_B200_GPUS=N): pinned host buffers in, tag vector out.  GPU box only.

    python scripts/group_bench.py [--gpus 2] [--reads 10000000] [--steps 3]

Prints one JSON line: ms per call (wall), query reads/s, the result's shared count (equal at every N) and the mode."""
import argparse, json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import commet_b200
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=2)
ap.add_argument("--reads", type=int, default=10_000_000)
ap.add_argument("--len", type=int, default=100, dest="length")
ap.add_argument("-k", type=int, default=33)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
args = ap.parse_args()
n, L, k, t = args.reads, args.length, args.k, 2
dev = torch.device("cuda", 0)
ref_d, qry_d, offs_d = bench.make_sets_torch(n, L, 0, dev)
ref_h = torch.empty(n * L, dtype=torch.uint8).pin_memory(); ref_h.copy_(ref_d)
qry_h = torch.empty(n * L, dtype=torch.uint8).pin_memory(); qry_h.copy_(qry_d)
offs_h = (np.arange(n + 1, dtype=np.uint64) * L)
del ref_d, qry_d, offs_d
torch.cuda.empty_cache()
g = commet_b200.Group(list(range(args.gpus)))
times = []
for it in range(args.warmup + args.steps):
    t0 = time.perf_counter()
    tags, info = g.index_and_search(k, t, (ref_h.numpy(), offs_h), [(qry_h.numpy(), offs_h)])
    dt = time.perf_counter() - t0
    if it >= args.warmup:
        times.append(dt)
print(json.dumps({"workload": f"C2 through commet_group_index_and_search, {args.gpus} GPU(s) of one process, {n} reads x {L} bp, k={k}",
                  "n_gpus": args.gpus, "ms_per_call": 1e3 * min(times), "ms_per_call_all": [round(1e3 * x, 2) for x in times],
                  "query_reads_per_s": n / min(times), "shared": info["shared"][0], "chunks": info["chunks"], "gpus": info["gpus"],
                  "index_ms_slowest_rank": info["index_ns"] / 1e6, "search_ms_slowest_rank": info["search_ns"] / 1e6}))
g.close()
