"""A/B of the host-side pipelining of commet_index_and_search on C2 (2 x 10 M x 100 bp, k=33, t=2, pinned host buffers):
where the index set is cut into parts (COMMET_B200_PART_FRACS) and in how many parts a query set goes up
(COMMET_B200_QUERY_PARTS).  Both are read at call time.  One JSON line per setting; GPU box only.

    python scripts/sweep_e2e_parts.py [--reads 10000000] [--reps 5]
"""
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
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--len", type=int, default=100, dest="length")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    import commet_b200
    from commet_b200 import build
    build.build_lib()
    dev = torch.device("cuda", 0)
    n, L, k, t = args.reads, args.length, 33, 2
    ref_d, qry_d = bench.make_sets_torch(n, L, 0, dev)[:2]
    ref_h = torch.empty(n * L, dtype=torch.uint8).pin_memory()
    qry_h = torch.empty(n * L, dtype=torch.uint8).pin_memory()
    ref_h.copy_(ref_d.reshape(-1)); qry_h.copy_(qry_d.reshape(-1))
    del ref_d, qry_d
    offs_t = torch.empty(n + 1, dtype=torch.int64).pin_memory()
    offs_t.copy_(torch.arange(0, n + 1, dtype=torch.int64) * L)
    offs_h = offs_t.numpy().view(np.uint64)
    ctx = commet_b200.Context(0)
    settings = [
        ("default 0.2,0.5 / 4", None, None),
        ("0.1,0.4 / 4", "0.1,0.4", None),
        ("0.05,0.2,0.5 / 4", "0.05,0.2,0.5", None),
        ("0.05,0.2,0.45,0.75 / 4", "0.05,0.2,0.45,0.75", None),
        ("0.1,0.3,0.6 / 4", "0.1,0.3,0.6", None),
        ("default / 8", None, "8"),
        ("0.05,0.2,0.5 / 8", "0.05,0.2,0.5", "8"),
        ("0.1,0.3,0.6 / 8", "0.1,0.3,0.6", "8"),
        ("0.1,0.3,0.6 / 6", "0.1,0.3,0.6", "6"),
        ("one part / 1", "", "1"),
    ]
    first = None
    for name, fr, qp in settings:
        for key, val in (("COMMET_B200_PART_FRACS", fr), ("COMMET_B200_QUERY_PARTS", qp)):
            if val is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = val
        if qp == "8" or qp == "6":
            os.environ["COMMET_B200_QUERY_PART_BYTES"] = str(100 << 20)
        else:
            os.environ.pop("COMMET_B200_QUERY_PART_BYTES", None)
        ts = []
        for i in range(2 + args.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, inf = ctx.index_and_search(k, t, (ref_h.numpy(), offs_h), [(qry_h.numpy(), offs_h)])
            ts.append((time.perf_counter() - t0) * 1e3)
        sh = int(inf["shared"][0])
        first = sh if first is None else first
        assert sh == first, (sh, first)
        ts = sorted(ts[2:])
        print(json.dumps({"setting": name, "ms_median": round(ts[len(ts) // 2], 2), "ms_min": round(ts[0], 2), "ms_max": round(ts[-1], 2),
                          "index_ms": round(inf["index_ns"] / 1e6, 2), "search_ms": round(inf["search_ns"] / 1e6, 2), "shared": sh}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
