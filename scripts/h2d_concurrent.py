"""Host-to-device bandwidth of one box, rank by rank and all ranks at once (what bounds the end-to-end leg at N = 8).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/h2d_concurrent.py [--mb 1024]

Every rank copies a page-locked buffer to its GPU: first alone (the others wait), then all together.  Rank 0 prints one JSON
line: GB/s per rank in both situations, the NUMA facts the ranks can see (sysfs node of the GPU, nodes online, the driver's
CPU affinity for the GPU).  GPU box only."""
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
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--bind", type=int, default=0, help="1: bind to the GPU's NUMA node first (bench.bind_to_gpu_numa_node)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    if args.bind:
        os.environ["BENCH_NUMA_BIND"] = "1"
        numa = bench.bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("gloo")
    n = args.mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(rank + 1)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    def timed():
        best = 1e9
        for _ in range(args.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            d.copy_(h, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return n / best / 1e9

    solo = 0.0
    for r in range(world):
        barrier()
        if r == rank:
            solo = timed()
    barrier()
    together = timed()
    barrier()
    facts = {"rank": rank, "solo_GBps": round(solo, 1), "together_GBps": round(together, 1), "cpu_now": sorted(os.sched_getaffinity(0))[:1],
             "n_cpus_allowed": len(os.sched_getaffinity(0)), "numa": numa}
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{getattr(p, 'pci_device_id', 0):02x}.0"
        facts["bdf"] = bdf
        facts["sysfs_node"] = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        import pynvml
        pynvml.nvmlInit()
        hd = pynvml.nvmlDeviceGetHandleByPciBusId(bdf.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(hd, (os.cpu_count() + 63) // 64)
        near = sorted(64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1)
        facts["nvml_near_cpus"] = [near[0], near[-1], len(near)] if near else []
    except Exception as e:
        facts["facts_error"] = str(e)[:120]
    out = [None] * world
    if world > 1:
        dist.all_gather_object(out, facts)
    else:
        out = [facts]
    if rank == 0:
        line = {"mb": args.mb, "n_ranks": world, "cpu_count": os.cpu_count(), "bind": args.bind,
                "nodes_online": open("/sys/devices/system/node/online").read().strip() if os.path.exists("/sys/devices/system/node/online") else None,
                "sum_solo_GBps": round(sum(f["solo_GBps"] for f in out), 1), "sum_together_GBps": round(sum(f["together_GBps"] for f in out), 1),
                "ranks": out}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
