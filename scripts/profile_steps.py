"""Host-side wall-clock breakdown of one C2 step (every call followed by a sync). GPU box only."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import commet_b200
from bench import make_sets_torch

n, L, k, t = 10_000_000, 100, 33, 2
dev = torch.device("cuda", 0)
ctx = commet_b200.Context(0)
ref_d, qry_d, offs_d = make_sets_torch(n, L, 0, dev)
tags = torch.zeros((n // 8 + 4) // 4 + 1, dtype=torch.int32, device=dev)
torch.cuda.synchronize()


def T(label, fn):
    t0 = time.perf_counter(); r = fn(); ctx.sync(); torch.cuda.synchronize()
    print(f"  {label:28s} {1e3 * (time.perf_counter() - t0):8.2f} ms"); return r


for it in range(3):
    print("step", it)
    t0 = time.perf_counter()
    q = T("stage query", lambda: ctx.stage_device(qry_d.data_ptr(), offs_d.data_ptr(), n, n * L))
    idx = T("stage index", lambda: ctx.stage_device(ref_d.data_ptr(), offs_d.data_ptr(), n, n * L))
    T("chunk_plan", lambda: ctx.chunk_plan(idx, k))
    T("kmer_counts(query)->prepare", lambda: ctx.chunk_plan(q, k))
    T("index_begin (memset)", lambda: ctx.index_begin(k))
    T("index_add", lambda: ctx.index_add(idx))
    tags.zero_()
    cnt = torch.zeros(4, dtype=torch.int64, device=dev)
    T("search", lambda: ctx.search_reads_device(q, k, t, tags.data_ptr(), cnt.data_ptr()))
    T("free", lambda: (q.free(), idx.free()))
    print(f"  total {1e3 * (time.perf_counter() - t0):8.2f} ms   shared={int(cnt[0])}")
    t0 = time.perf_counter()
    q = ctx.stage_device(qry_d.data_ptr(), offs_d.data_ptr(), n, n * L)
    idx = ctx.stage_device(ref_d.data_ptr(), offs_d.data_ptr(), n, n * L)
    tags.zero_()
    info = ctx.index_and_search_staged(k, t, idx, [q], [tags.data_ptr()])
    q.free(); idx.free(); ctx.sync()
    print(f"  fused call total {1e3 * (time.perf_counter() - t0):8.2f} ms  index {info['index_ns']/1e6:.2f} search {info['search_ns']/1e6:.2f}")
