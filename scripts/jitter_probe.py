"""Is the step time disturbed by clock sampling / a busy host?  GPU box only (diagnostic)."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import commet_b200
import bench

n, L, k, t = 10_000_000, 100, 33, 2
dev = torch.device("cuda", 0)
print("loadavg", os.getloadavg(), "cpus", os.cpu_count(), flush=True)
ctx = commet_b200.Context(0)
ref_d, qry_d, offs_d = bench.make_sets_torch(n, L, 0, dev)
tags = torch.zeros((n // 8 + 4) // 4 + 1, dtype=torch.int32, device=dev)
torch.cuda.synchronize()


def step():
    t0 = time.perf_counter()
    q = ctx.stage_device(qry_d.data_ptr(), offs_d.data_ptr(), n, n * L)
    idx = ctx.stage_device(ref_d.data_ptr(), offs_d.data_ptr(), n, n * L)
    tags.zero_()
    torch.cuda.synchronize()
    info = ctx.index_and_search_staged(k, t, idx, [q], [tags.data_ptr()])
    q.free(); idx.free(); ctx.sync()
    return 1e3 * (time.perf_counter() - t0), info["index_ns"] / 1e6, info["search_ns"] / 1e6


def loop(label, m=8):
    rows = [step() for _ in range(m)]
    print(label, " ".join(f"{a:.1f}/{b:.1f}/{c:.1f}" for a, b, c in rows), flush=True)


loop("warmup      ", 3)
loop("no sampler  ")
with bench.ClockSampler(0) as cs:
    loop("nvml 20ms   ")
print("   ", cs.summary())
cs2 = bench.ClockSampler(0); cs2._nvml = None; cs2.source = "nvidia-smi"
with cs2:
    loop("nvidia-smi  ")
print("   ", cs2.summary())
loop("no sampler  ")

ref_h = torch.empty(n * L, dtype=torch.uint8).pin_memory(); ref_h.copy_(ref_d)
qry_h = torch.empty(n * L, dtype=torch.uint8).pin_memory(); qry_h.copy_(qry_d)
offs_ht = torch.empty(n + 1, dtype=torch.int64).pin_memory(); offs_ht.copy_(offs_d)
offs_h = offs_ht.numpy().view(np.uint64)
torch.cuda.synchronize()
for rep in range(8):
    t0 = time.perf_counter()
    _, inf = ctx.index_and_search(k, t, (ref_h.numpy(), offs_h), [(qry_h.numpy(), offs_h)])
    print(f"e2e {1e3 * (time.perf_counter() - t0):.1f} ms  index {inf['index_ns']/1e6:.1f} search {inf['search_ns']/1e6:.1f}", flush=True)
# raw PCIe: one 1 GB pinned H2D copy
d = torch.empty(n * L, dtype=torch.uint8, device=dev)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d.copy_(ref_h, non_blocking=True); torch.cuda.synchronize()
    print(f"H2D 1 GB: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
print("loadavg", os.getloadavg())
os.environ["COMMET_B200_TRACE"] = "1"
for rep in range(4):
    t0 = time.perf_counter()
    _, inf = ctx.index_and_search(k, t, (ref_h.numpy(), offs_h), [(qry_h.numpy(), offs_h)])
    print(f"e2e {1e3 * (time.perf_counter() - t0):.1f} ms  index {inf['index_ns']/1e6:.1f} search {inf['search_ns']/1e6:.1f}", flush=True)
