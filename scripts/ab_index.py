"""A/B of the insert path knobs on the C2 reference set (device-resident and through the host entry). GPU box only."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import commet_b200
import bench

n, L, k, t = 10_000_000, 100, 33, 2
dev = torch.device("cuda", 0)
ctx = commet_b200.Context(0)
ref_d, qry_d, offs_d = bench.make_sets_torch(n, L, 0, dev)
tags = torch.zeros((n // 8 + 4) // 4 + 1, dtype=torch.int32, device=dev)
torch.cuda.synchronize()


def step():
    q = ctx.stage_device(qry_d.data_ptr(), offs_d.data_ptr(), n, n * L)
    idx = ctx.stage_device(ref_d.data_ptr(), offs_d.data_ptr(), n, n * L)
    tags.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = ctx.index_and_search_staged(k, t, idx, [q], [tags.data_ptr()])
    ctx.sync()
    ms = 1e3 * (time.perf_counter() - t0)
    q.free(); idx.free()
    return ms, info["index_ns"] / 1e6, info["search_ns"] / 1e6, info["shared"][0]


ref_h = torch.empty(n * L, dtype=torch.uint8).pin_memory(); ref_h.copy_(ref_d)
qry_h = torch.empty(n * L, dtype=torch.uint8).pin_memory(); qry_h.copy_(qry_d)
offs_ht = torch.empty(n + 1, dtype=torch.int64).pin_memory(); offs_ht.copy_(offs_d)
offs_h = offs_ht.numpy().view(np.uint64)


def e2e():
    t0 = time.perf_counter()
    _, inf = ctx.index_and_search(k, t, (ref_h.numpy(), offs_h), [(qry_h.numpy(), offs_h)])
    return 1e3 * (time.perf_counter() - t0), inf["index_ns"] / 1e6, inf["search_ns"] / 1e6, inf["shared"][0]


for _ in range(3):
    step(); e2e()
for env in sys.argv[1:] or [""]:
    for kv in env.split(","):
        if "=" in kv:
            a, b = kv.split("=")
            os.environ[a] = b
    r = [step() for _ in range(4)]
    e = [e2e() for _ in range(4)]
    print(f"{env or 'default':50s} dev {min(x[0] for x in r):6.2f} ms (index {min(x[1] for x in r):6.2f} search {min(x[2] for x in r):6.2f})"
          f"   e2e {min(x[0] for x in e):6.2f} ms (index {min(x[1] for x in e):6.2f})  shared {r[0][3]} {e[0][3]}", flush=True)
