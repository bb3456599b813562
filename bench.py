#!/usr/bin/env python
"""bench.py -- Commet index_and_search hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--len L] [-k 33] [-t 2]
    python bench.py --impl reference ...       # the reference's own CPU tool on this box's cores

One "step" = one index_and_search pass: stage (2-bit encode) the reference set
and the query set, index every k-mer of the reference set into the
bloom_filter.h-layout bit array, then search every query read (forward +
reverse-complement greedy scan).  Workload at N=1 = BASELINE.json configs[1]:
2 synthetic sets x 10M reads x 100 bp, k=33, t=2.

At N>1 (weak scaling): every rank owns its own 10M-read query set; the
reference set is dealt block-cyclically over the ranks (each rank holds,
uploads and encodes only its 1/N of it), each rank builds a partial filter
from its shard, the partials are merged by ONE kernel per rank over NVLink
peer memory (reduce-scatter + all-gather of the OR, commet_index_merge), then
each rank probes locally (commet_b200/multi.py, distributed placement).
value = total query reads / max-over-ranks time.

`value`   : inputs (ASCII bases + offsets) already resident in HBM.
`e2e`     : same pass through Context.index_and_search with HOST buffers
            (H2D of both sets and D2H of the tag vector inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "query_reads_per_s"
UNIT = "reads/s"


# ----------------------------------------------------------------------------
# synthetic sets: i.i.d. ACGT reference; queries = 50 % copies of
# reference reads (half reverse-complemented, 1 % substitutions) + 50 % random
# ----------------------------------------------------------------------------
def make_sets_torch(n_reads: int, length: int, seed: int, device, qseed: int | None = None):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(1000 + seed)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    comp = torch.zeros(256, dtype=torch.uint8, device=device)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    ref = torch.empty((n_reads, length), dtype=torch.uint8, device=device)
    qry = torch.empty((n_reads, length), dtype=torch.uint8, device=device)
    step = 1 << 20
    for s in range(0, n_reads, step):
        e = min(n_reads, s + step)
        ref[s:e] = acgt[torch.randint(0, 4, (e - s, length), generator=g, device=device)]
    g.manual_seed(2000 + (seed if qseed is None else qseed))
    for s in range(0, n_reads, step):
        e = min(n_reads, s + step)
        m = e - s
        src = torch.randint(0, n_reads, (m,), generator=g, device=device)
        cp = ref[src]
        rc = torch.rand(m, generator=g, device=device) < 0.5
        cp = torch.where(rc[:, None], comp[cp.flip(1).long()], cp)
        mut = torch.rand((m, length), generator=g, device=device) < 0.01
        rnd = acgt[torch.randint(0, 4, (m, length), generator=g, device=device)]
        cp = torch.where(mut, rnd, cp)
        shared = torch.rand(m, generator=g, device=device) < 0.5
        fresh = acgt[torch.randint(0, 4, (m, length), generator=g, device=device)]
        qry[s:e] = torch.where(shared[:, None], cp, fresh)
    offs = torch.arange(0, n_reads + 1, dtype=torch.int64, device=device) * length
    return ref.reshape(-1), qry.reshape(-1), offs


class ClockSampler:
    """SM clocks and throttle reasons during the timed region (B200_PROFILING.md), sampled in-process through
    NVML every 20 ms -- spawning nvidia-smi takes ~100 ms per sample and holds driver locks the timed CUDA calls
    wait on; nvidia-smi is only the fallback when NVML cannot be loaded."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index = index
        self.rows = []          # (sm_mhz, max_mhz, [active reason names])
        self.source = "nvml"
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remaps indices: find the device through its PCI bus id
            import torch
            bus = getattr(torch.cuda.get_device_properties(index), "pci_bus_id", None)
            h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(hi).bus) == int(bus):
                        h = hi
                        break
            self._h = h if h is not None else pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml = pynvml
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
            self.source = "nvidia-smi"

    def _sample_nvml(self):
        nv = self._nvml
        sm = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]
        self.rows.append((sm, self._max, [nm for nm, b in zip(self.NAMES, bits) if mask & b]))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        r = [x.strip() for x in out.strip().split(",")]
        self.rows.append((float(r[0]), float(r[1]), [nm for nm, v in zip(self.NAMES, r[3:7]) if v.lower().startswith("active")]))

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.02 if self._nvml is not None else 0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [r[0] for r in self.rows]
        mx = max([r[1] for r in self.rows], default=0)
        reasons = sorted({nm for r in self.rows for nm in r[2]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm), "source": self.source}


def bind_to_gpu_numa_node(index: int):
    """Run this rank on the CPUs of its GPU's NUMA node and prefer that node's memory, so that the pinned buffers its H2D
    copies read from are local to the GPU's PCIe root.  On by default when there is more than one rank (BENCH_NUMA_BIND=0
    turns it off; A/B in profiles/).  Returns what was done (for the JSON line) or None."""
    dflt = "1" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else "0"
    if os.environ.get("BENCH_NUMA_BIND", dflt) in ("", "0"):
        return None
    try:
        import ctypes
        import torch
        p = torch.cuda.get_device_properties(index)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{getattr(p, 'pci_device_id', 0):02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        info = {"bdf": bdf, "node": node}
        if node < 0:
            # sysfs does not say (virtualised topology): ask the driver which CPUs are close to this GPU and take the
            # node of the first of them -- unless the answer is "all of them", which carries no information
            info["nodes_online"] = open("/sys/devices/system/node/online").read().strip()
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bdf.encode())
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            near = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
            info["nvml_cpus"] = len(near)
            if not near or len(near) >= os.cpu_count():
                return info
            import glob
            for d in glob.glob("/sys/devices/system/node/node[0-9]*"):
                lst = set()
                for part in open(d + "/cpulist").read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    if lo:
                        lst.update(range(int(lo), int(hi or lo) + 1))
                if min(near) in lst:
                    node = int(d.rsplit("node", 1)[1])
            if node < 0:
                return info
            info["node_from_nvml"] = node
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        use = cpus & os.sched_getaffinity(0)
        if use:
            os.sched_setaffinity(0, use)
        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(node + 2))     # x86-64 set_mempolicy(MPOL_PREFERRED)
        info.update({"node": node, "cpus": len(use), "set_mempolicy": int(rc)})
        return info
    except Exception as e:                                                           # never fatal: it is an A/B switch
        return {"error": str(e)[:200]}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------
# CPU arms
# ----------------------------------------------------------------------------
def write_fasta(path, bases: np.ndarray, n: int, length: int):
    arr = bases[:n * length].reshape(n, length)
    hdr = np.char.add(np.char.add(">", np.arange(n).astype(str)), "\n").astype("S")
    with open(path, "wb") as f:
        step = 100000
        for s in range(0, n, step):
            e = min(n, s + step)
            rows = [hdr[i] + arr[i].tobytes() + b"\n" for i in range(s, e)]
            f.write(b"".join(rows))


def cpu_reference_run(ref_b, qry_b, length, n_ref, n_qry, k, t, procs: int):
    """The reference's own index_and_search (oracle/_ref, compiled from /root/reference) as `procs`
    independent processes -- the only parallelism it supports: each indexes the reference sample and
    searches its share of the queries.  Returns (wall seconds, kind, cores)."""
    from oracle import oracle
    tool = oracle.REF_DIR / "index_and_search"
    if not tool.exists():
        return None
    with tempfile.TemporaryDirectory(prefix="commet_cpu_") as td:
        td = Path(td)
        write_fasta(td / "ref.fa", ref_b, n_ref, length)
        (td / "ref.txt").write_text(f"ref:{td}/ref.fa\n")
        per = (n_qry + procs - 1) // procs
        cmds = []
        for p in range(procs):
            s, e = p * per, min(n_qry, (p + 1) * per)
            if s >= e:
                break
            write_fasta(td / f"q{p}.fa", qry_b[s * length:e * length], e - s, length)
            (td / f"q{p}.txt").write_text(f"q{p}:{td}/q{p}.fa\n")
            cmds.append([str(tool), "-i", str(td / "ref.txt"), "-s", str(td / f"q{p}.txt"), "-o", str(td / f"o{p}"),
                         "-l", str(td / f"o{p}"), "-k", str(k), "-t", str(t)])
        t0 = time.perf_counter()
        ps = [subprocess.Popen(c, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in cmds]
        rcs = [p.wait() for p in ps]
        wall = time.perf_counter() - t0
        if any(rcs):
            return None
        return wall, "reference", len(cmds)


def cpu_port_run(ref_b, qry_b, length, n_ref, n_qry, k, t):
    from oracle import oracle
    io = np.arange(n_ref + 1, dtype=np.uint64) * length
    qo = np.arange(n_qry + 1, dtype=np.uint64) * length
    t0 = time.perf_counter()
    oracle.index_and_search(k, t, (ref_b[:n_ref * length], io), [(qry_b[:n_qry * length], qo)])
    return time.perf_counter() - t0, "port", 1


def cpu_baseline(ref_b, qry_b, length, k, t, sample: int, procs: int = 1):
    r = cpu_reference_run(ref_b, qry_b, length, sample, sample, k, t, procs)
    if r is None:
        r = cpu_port_run(ref_b, qry_b, length, sample, sample, k, t)
    wall, kind, cores = r
    return {"value": sample / wall, "unit": UNIT, "cores": cores, "kind": kind, "seconds": round(wall, 3),
            "sample": f"first {sample} reference reads indexed + first {sample} query reads searched "
                      f"(same generator, {length} bp, k={k}, t={t}); value = query reads / wall"}


# ----------------------------------------------------------------------------
# cheap extra legs of the N=1 line (BASELINE.json configs other than the headline one): each is a few seconds, parity for
# each is in tests/ (named per leg); a leg that fails reports its error instead of taking the line down
# ----------------------------------------------------------------------------
def extra_legs(torch, ctx, dev, ext, args, ref_d, qry_d, offs_d, ref_h, qry_h, shared_c2):
    import commet_b200
    from commet_b200 import build
    n, L, k, t = args.reads, args.length, args.k, args.t
    peak, _ = measured_peaks()
    out = {}

    def leg(name, fn):
        try:
            out[name] = fn()
        except Exception as e:                       # noqa: BLE001
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}

    def tool_e2e():
        """C2 through the drop-in executable: FASTA parse, H2D, kernels, D2H and the .bv write inside the clock (SURVEY 8d's
        end-to-end scope); the FASTA files are written to a tmpfs first (not timed).  Parity: tests/test_gpu_fullsize.py::
        test_c2_whole_vector_equals_the_reference_binary."""
        build.build_tools()
        base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
        with tempfile.TemporaryDirectory(prefix="commet_tool_", dir=base) as td:
            td = Path(td)
            for name, arr in (("ref", ref_h), ("qry", qry_h)):
                rows = np.empty((n, 1 + 8 + 1 + L + 1), dtype=np.uint8)
                rows[:, 0] = ord(">")
                idx = np.arange(n, dtype=np.int64)
                for d in range(8):
                    rows[:, 1 + d] = (idx // 10 ** (7 - d)) % 10 + 48
                rows[:, 9] = 10
                rows[:, 10:10 + L] = arr.numpy().reshape(n, L)
                rows[:, -1] = 10
                rows.tofile(td / f"{name}.fa")
                (td / f"{name}.txt").write_text(f"{name}:{td}/{name}.fa\n")
                del rows
            best = None
            for _ in range(2):
                t0 = time.perf_counter()
                r = subprocess.run([str(build.BIN / "index_and_search"), "-i", str(td / "ref.txt"), "-s", str(td / "qry.txt"), "-o",
                                    str(td / "out"), "-l", str(td / "out"), "-k", str(k), "-t", str(t)], capture_output=True, text=True)
                wall = time.perf_counter() - t0
                if r.returncode != 0:
                    raise RuntimeError(r.stderr[-300:])
                best = wall if best is None else min(best, wall)
            log = (td / "out" / "qry_in_ref.log").read_text()
            assert f"shared {shared_c2}]" in log, log
            return {"value": n / best, "unit": UNIT, "seconds": round(best, 3), "fasta_bytes": int(2 * n * (L + 11)),
                    "scope": "process start + CUDA context + FASTA parse (2 files) + H2D + kernels + D2H + .bv and .log write; best of 2"}

    def k27():
        """the L2-resident variant of C2: k=27, 64 MiB filter, 48 chunks.  Parity: test_search_small_and_large_k,
        test_index_and_search_chunk_loop (multi-chunk), the C4 twin at k=27."""
        tags = torch.zeros((n // 8 + 1 + 3) // 4, dtype=torch.int32, device=dev)
        res = None
        for it in range(2):
            ctx.count_probes(it == 1)
            tags.zero_()
            torch.cuda.synchronize()
            q = ctx.stage_device(qry_d.data_ptr(), offs_d.data_ptr(), n, n * L)
            idx = ctx.stage_device(ref_d.data_ptr(), offs_d.data_ptr(), n, n * L)
            info = ctx.index_and_search_staged(27, t, idx, [q], [tags.data_ptr()])
            ctx.sync()
            idx.free(); q.free()
            if it == 0:
                res = info
        ctx.count_probes(False)
        ceil = random_sector_ceiling(1 << 26)
        ach = info["tests"] * 32 / (res["search_ns"] / 1e9) / 1e9
        return {"chunks": res["chunks"], "index_ms": res["index_ns"] / 1e6, "search_ms": res["search_ns"] / 1e6,
                "query_reads_per_s": n / ((res["index_ns"] + res["search_ns"]) / 1e9), "n_probes": info["tests"],
                "probes_per_s": info["tests"] / (res["search_ns"] / 1e9), "achieved_GBps": ach, "l2_random_sector_ceiling_GBps": ceil,
                "frac_of_l2_random_sector_ceiling": ach / ceil if ceil else None}

    def c5():
        """bandwidth-bound vector operators at 1e9 bits and the selection over the C2 query set (-l 66 -n 2 -e 1.5).
        Parity: test_c5_vector_identities_at_1e9_bits, test_c5_filter_reads_constructed_classes_10m, test_filter_reads."""
        bits = 1_000_000_000
        nbytes = bits // 8 + 1
        g = torch.Generator(device=dev)
        g.manual_seed(7)
        pad = (nbytes + 15) // 16 * 16
        a = torch.randint(0, 256, (pad,), dtype=torch.uint8, generator=g, device=dev)
        b = torch.randint(0, 256, (pad,), dtype=torch.uint8, generator=g, device=dev)
        o = torch.empty_like(a)
        torch.cuda.synchronize()
        ctx.bvop_device(commet_b200.BV_AND, a.data_ptr(), b.data_ptr(), o.data_ptr(), nbytes)
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
        for _ in range(5):
            ctx.bvop_device(commet_b200.BV_AND, a.data_ptr(), b.data_ptr(), o.data_ptr(), nbytes)
        with torch.cuda.stream(ext):
            e1.record()
        ctx.sync()
        ms_and = e0.elapsed_time(e1) / 5
        t0 = time.perf_counter()
        ones = ctx.nb_one_device_batch([o.data_ptr()] * 8, [bits] * 8)
        ms_pop = (time.perf_counter() - t0) / 8 * 1e3
        d_bv = torch.zeros((n // 8 + 1 + 3) // 4, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        rs = ctx.stage_device(qry_d.data_ptr(), offs_d.data_ptr(), n, n * L)
        ctx.filter_reads_staged(rs, d_bv.data_ptr(), min_len=66, max_N=2, min_shannon=1.5)
        rs.free()
        ctx.sync()
        with torch.cuda.stream(ext):
            e0.record()
        rs = ctx.stage_device(qry_d.data_ptr(), offs_d.data_ptr(), n, n * L)
        cnt = ctx.filter_reads_staged(rs, d_bv.data_ptr(), min_len=66, max_N=2, min_shannon=1.5)
        with torch.cuda.stream(ext):
            e1.record()
        ctx.sync()
        rs.free()
        ms_f = e0.elapsed_time(e1)
        return {"bvop_and_1e9_bits": {"ms": ms_and, "GBps": 3 * nbytes / ms_and / 1e6, "frac_of_hbm_peak": 3 * nbytes / ms_and / 1e6 / peak},
                "popcount_1e9_bits_batched": {"ms": ms_pop, "GBps": nbytes / ms_pop / 1e6, "frac_of_hbm_peak": nbytes / ms_pop / 1e6 / peak,
                                               "ones": ones[0]},
                "stage_and_filter_reads": {"reads": n, "bases": n * L, "ms": ms_f, "reads_per_s": n / (ms_f / 1e3),
                                           "GBps_of_ascii": n * L / ms_f / 1e6, "frac_of_hbm_peak": n * L / ms_f / 1e6 / peak,
                                           "selected": cnt["selected"],
                                           "scope": "k_encode + k_filter over the planes + cutoff, the call with its read-backs (CUDA events)"}}

    def c4_small():
        """C4 at 1/25 of the reference set (20 M reads x 150 bp, 3 chunks at k=33) against 8 query sets of 2 M reads: the multi-chunk,
        multi-set loop in this line; full size (500 M reads, 59 chunks) at 1/2/4/8 GPUs is profiles/r02_c4_full_n*.json.
        Parity: tests/test_gpu_multi.py::test_c4_twin_every_vector_equals_the_oracle."""
        import importlib.util
        spec = importlib.util.spec_from_file_location("bench_c4", ROOT / "scripts" / "bench_c4.py")
        bench_c4 = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench_c4)
        info, res = bench_c4.run_c4(torch, ctx, dev, 1, 0, torch.cuda.synchronize, lambda b: [b], 20_000_000, 8, 2_000_000, 150, 33, 2)
        return {"ref_reads": 20_000_000, "query_sets": 8, "query_reads_per_set": 2_000_000, "read_len": 150, "chunks": info["chunks"],
                "seconds": round(info["seconds"], 3), "query_reads_per_s": 8 * 2_000_000 / info["seconds"],
                "index_s": round(info["index_s"], 3), "search_s": round(info["search_ns"] * 1e-9, 3), "shared_set0": int(res[0][0])}

    leg("tool_e2e_c2", tool_e2e)
    leg("k27_l2_resident", k27)
    leg("c5", c5)
    leg("c4_small", c4_small)
    return out


# ----------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--len", type=int, default=100, dest="length")
    ap.add_argument("-k", type=int, default=33)
    ap.add_argument("-t", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=2_000_000,
                    help="reads per set of the bounded CPU sample (2 M: ~20 s for one reference process at C2, so that the "
                         "fixed cost of zeroing the 4 GiB filter is amortised as it is at full size)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (tool-level end to end, k=27, C5, small C4) of the N=1 line")
    ap.add_argument("--direct-index", action="store_true", help="disable the L2-blocked insert (A/B)")
    ap.add_argument("--index-mode", type=int, default=None, help="commet_ctx_binned_index mode (A/B): 0 direct, 1 "
                    "sorted records (default), 16..30 region passes with 2^mode-byte regions")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"C2: 2 synthetic sets x {args.reads} reads x {args.length} bp, k={args.k} t={args.t}, "
                "single index_and_search" + (f", reference set dealt block-cyclically over {world} GPUs (each holds 1/{world}), one "
                                             f"query set per GPU" if world > 1 else ""))
    config = {"workload": workload, "reads_per_set": args.reads, "read_len": args.length, "k": args.k, "t": args.t,
              "filter_bytes": 1 << (args.k - 1),
              "l2_policy": f"inputs larger than L2 ({2 * args.reads * args.length / 1e9:.1f} GB of bases + "
                           f"{(1 << (args.k - 1)) / 2**20:.0f} MiB filter per step, 126 MB L2)"}

    if args.impl == "reference":
        return reference_arm(args, rank, world, config)

    # stdout carries ONE JSON line: everything libraries print there (NCCL's version banner, warnings) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    import torch
    import torch.distributed as dist
    import commet_b200
    from commet_b200 import build
    build.build_lib()

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        # stdout carries ONE JSON line: NCCL writes its version banner / warnings to stdout unless told otherwise
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    ctx = commet_b200.Context(local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    if args.direct_index:
        ctx.binned_index(0)
    if args.index_mode is not None:
        ctx.binned_index(args.index_mode)

    n, L, k, t = args.reads, args.length, args.k, args.t
    k_arg = k
    ref_d, qry_d, offs_d = make_sets_torch(n, L, seed=0, device=dev, qseed=rank)
    torch.cuda.synchronize()
    n_tag_words = (n // 8 + 1 + 3) // 4
    tags_d = torch.zeros(n_tag_words, dtype=torch.int32, device=dev)

    dd = None
    if world > 1:
        # one-time: every rank's filter is mapped into every process (CUDA IPC over NVLink peer memory) by
        # commet_dist_open; torch.distributed only serves the two callbacks of the library's loop (barrier, all-gather of
        # a few bytes: the k-mer totals of the chunk plan)
        from commet_b200 import multi

        # (host-side control messages of a few bytes: a gloo group -- an NCCL collective would cost a launch, two copies and
        # a device synchronisation each; NCCL stays the backend of the timing all-reduces and of the barriers around them)
        ctl = dist.new_group(backend="gloo")

        def all_gather_bytes(b):
            t_in = torch.frombuffer(bytearray(b), dtype=torch.uint8)
            outs = [torch.empty(len(b), dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(outs, t_in, group=ctl)
            return [o.numpy().tobytes() for o in outs]

        def ctl_barrier():
            dist.barrier(group=ctl)
        dd = commet_b200.Dist(ctx, world, rank, k_arg, ctl_barrier, all_gather_bytes)
        # this rank's shard of the reference set: blocks b = rank (mod world) of BLOCK consecutive reads -- what a
        # rank's loader delivers when every process parses only its own blocks of the files
        BLOCK = multi.DEFAULT_BLOCK
        own = torch.as_tensor(multi.owned_mask(n, world, rank, BLOCK), device=dev)
        n_loc = int(own.sum())
        ref_loc_d = ref_d.view(n, L)[own].reshape(-1).contiguous()
        offs_loc_d = torch.arange(0, n_loc + 1, dtype=torch.int64, device=dev) * L

    def step_device():
        """one pass with inputs resident in HBM; returns info dict"""
        tags_d.zero_()
        ext.wait_stream(torch.cuda.current_stream())
        q = ctx.stage_device(qry_d.data_ptr(), offs_d.data_ptr(), n, n * L)
        if world == 1:
            idx = ctx.stage_device(ref_d.data_ptr(), offs_d.data_ptr(), n, n * L)
            info = ctx.index_and_search_staged(k, t, idx, [q], [tags_d.data_ptr()])
        else:
            # the rank stages ITS shard of the reference set; every chunk's partial filters are merged by the
            # one-kernel OR all-reduce over peer memory; every rank probes its own query set
            idx = ctx.stage_device(ref_loc_d.data_ptr(), offs_loc_d.data_ptr(), n_loc, n_loc * L)
            torch.cuda.current_stream().synchronize()
            r = dd.index_and_search(t, idx, n, [q], [tags_d.data_ptr()], block=BLOCK)
            info = {"shared": r["shared"], "searched": r["searched"], "chunks": r["chunks"], "index_ns": 0, "search_ns": r["search_ns"],
                    "kmers": 0, "phases_ms": {key[:-2]: round(r[key] * 1e3, 3) for key in ("plan_s", "index_s", "merge_s", "barrier_s")}}
            info["phases_ms"]["search"] = round(r["search_ns"] / 1e6, 3)
            info["dist_mode"] = r["mode"]
        idx.free()
        q.free()
        return info

    # pinned host copies for the end-to-end leg
    ref_h = torch.empty(n * L, dtype=torch.uint8).pin_memory()
    qry_h = torch.empty(n * L, dtype=torch.uint8).pin_memory()
    ref_h.copy_(ref_d); qry_h.copy_(qry_d)
    offs_ht = torch.empty(n + 1, dtype=torch.int64).pin_memory()       # page-locked like the bases: H2D copies queue ahead
    offs_ht.copy_(offs_d)
    offs_h = offs_ht.numpy().view(np.uint64)
    if world > 1:
        ref_loc_h = torch.empty(n_loc * L, dtype=torch.uint8).pin_memory()
        ref_loc_h.copy_(ref_loc_d)
        offs_loc_ht = torch.empty(n_loc + 1, dtype=torch.int64).pin_memory()
        offs_loc_ht.copy_(offs_loc_d)
        offs_loc_h = offs_loc_ht.numpy().view(np.uint64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record()
        infos = [fn() for _ in range(steps)]
        with torch.cuda.stream(ext):
            e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt)
        return ms / steps, infos

    for _ in range(args.warmup):
        info = step_device()
    launches0 = ctx.launches
    with ClockSampler(local_rank) as cs:
        ms_dev, infos = timed(step_device, args.steps)
    launches = (ctx.launches - launches0) // args.steps
    info = infos[-1]
    shared = info["shared"][0]
    probes = None
    if world == 1:
        # untimed instrumented pass: the number of filter byte tests the REFERENCE semantics perform (N_probes)
        ctx.count_probes(True)
        probes = step_device()
        ctx.count_probes(False)
        assert probes["shared"][0] == shared

    # end-to-end leg: the same pass through the public API with HOST (pinned) buffers: H2D of the sets and D2H
    # of the tag vector inside the timed region.  N=1: Context.index_and_search (= the C-ABI call the drop-in
    # index_and_search tool makes).  N>1: every rank queues the upload of its shard of the reference set, then of
    # its own query set (commet_reads_upload_async: the query bytes cross PCIe during the insert and the merge),
    # then runs the distributed loop of commet_b200/multi.py; wall clock, max over ranks.
    tags_h = torch.empty(n // 8 + 1, dtype=torch.uint8).pin_memory()

    def step_e2e():
        if world == 1:
            _, inf = ctx.index_and_search(k, t, (ref_h.numpy(), offs_h), [(qry_h.numpy(), offs_h)])
            return inf["shared"][0]
        tags_d.zero_()
        torch.cuda.current_stream().synchronize()
        idx = ctx.stage_async(ref_loc_h.numpy(), offs_loc_h)
        # the query set goes up in four parts cut at multiples of 32 reads (disjoint tag words), as commet_index_and_search
        # does with host sets: the search of a part runs while the next one crosses PCIe
        cuts = [(n * i // 4) & ~31 for i in range(4)] + [n]
        qs = [ctx.stage_async(qry_h.numpy()[a * L:b * L], offs_h[:b - a + 1]) for a, b in zip(cuts[:-1], cuts[1:])]
        r = dd.index_and_search(t, idx, n, qs, [tags_d.data_ptr() + 4 * (a // 32) for a in cuts[:-1]], block=BLOCK)
        tags_h.copy_(tags_d.view(torch.uint8)[:n // 8 + 1])
        idx.free()
        for q in qs:
            q.free()
        return sum(r["shared"])

    for _ in range(max(1, args.warmup)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        ones = step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert ones == shared, (ones, shared)
    if world > 1:
        tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt)
    e2e = {"value": n * world / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
           "h2d_bytes_per_step": int(world * (n * L + 8 * (n + 1)) + n * L + 8 * (n + world)),
           "d2h_bytes_per_step": int(world * (n // 8 + 1))}

    if dd is not None:
        dd.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peaks()
    kmers = info["kmers"] or n * max(0, L - k + 1)
    idx_ms, srch_ms = info["index_ns"] / 1e6, info["search_ns"] / 1e6
    line = {
        "metric": METRIC, "value": n * world / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": config, "gpu_launches": int(launches),
        "clocks": cs.summary(), "shared_reads": int(shared), "chunks": info["chunks"], "e2e": e2e,
    }
    if "phases_ms" in info:
        line["phases_ms_rank0"] = info["phases_ms"]
        line["dist_mode"] = info.get("dist_mode")
    if numa is not None:
        line["numa_bind_rank0"] = numa
    if world == 1 and srch_ms > 0 and probes:
        # dominant kernel = k_search (one launch per query set and chunk).  Algorithmic bytes (SURVEY 8d,
        # DESIGN.md 5): one Bloom bit test = one 32-byte sector, counted with the REFERENCE's semantics
        # (short-circuit a,b,c,d; greedy k-jump; reverse strand only when the forward scan failed).
        ncu = ncu_traffic("k_search", config)
        ach = probes["tests"] * 32 / (srch_ms / 1e3) / 1e9
        ceil = random_sector_ceiling(1 << (k - 1))
        line["roofline"] = {"bound": "hbm", "kernel": "k_search", "achieved": ach, "peak": peak, "unit": "GB/s",
                            "frac": ach / peak, "traffic": ncu, "peak_source": peak_src,
                            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one k_search launch of this workload in the committed "
                                              "ncu --set full capture (profiles/ncu_traffic.json), not measured in this run",
                            "algorithmic": f"{probes['tests']} filter bit tests x 32 B sector per launch",
                            "ms_per_launch": srch_ms,
                            "random_sector_ceiling_GBps": ceil, "frac_of_random_sector_ceiling": (ach / ceil) if ceil else None,
                            "random_sector_ceiling_residency": "dram" if (1 << (k - 1)) > (64 << 20) else "l2"}
        # the insert is L2-blocked: its bound is the L2-resident RED.OR rate (profiles/r01_ceilings.json, measured with
        # scripts/ubench_sectors.cu), not DRAM bytes.  SURVEY 8(d)'s model (64 B per key: a sector read and its write-back)
        # is kept beside it -- it can exceed the copy peak precisely because no key pays that DRAM round trip any more.
        red_ceiling = None
        try:
            red_ceiling = float(json.loads((ROOT / "profiles" / "r01_ceilings.json").read_text())["l2_64MiB_redor_Gsectors_s"])
        except Exception:
            pass
        reds = 4 * kmers / (idx_ms / 1e3) / 1e9
        ins_bytes = kmers * 4 * 64
        line["roofline_index"] = {"bound": "l2-atomic", "kernel": "k_bin_scatter2+k_bin_plan2+k_bin_apply2", "achieved": reds,
                                  "peak": red_ceiling, "unit": "G RED.OR/s", "frac": (reds / red_ceiling) if red_ceiling else None,
                                  "peak_source": "measured L2-resident RED.OR ceiling, 64 MiB buffer (profiles/r01_ceilings.json)",
                                  "algorithmic": f"{kmers} k-mers x 4 keys = {4 * kmers} RED.OR per step; the time is the whole insert "
                                                 "(record scatter + apply), the apply alone is in the launch list under profiles/",
                                  "ms": idx_ms,
                                  "survey_model": {"bytes_per_key": 64, "GBps": ins_bytes / (idx_ms / 1e3) / 1e9,
                                                   "frac_of_hbm_peak": ins_bytes / (idx_ms / 1e3) / 1e9 / peak}}
        line["kernels"] = {"index_ms": idx_ms, "search_ms": srch_ms, "kmers_per_s": kmers / (idx_ms / 1e3),
                           "key_inserts_per_s": 4 * kmers / (idx_ms / 1e3), "n_probes": probes["tests"],
                           "n_lookups": probes["lookups"], "probes_per_s": probes["tests"] / (srch_ms / 1e3)}
    if not args.no_cpu and world == 1:
        line["cpu_baseline"] = cpu_baseline(ref_h.numpy(), qry_h.numpy(), L, k, t, min(args.cpu_sample, n))
    if world == 1 and not args.no_extra and k == 33:
        line["extra"] = extra_legs(torch, ctx, dev, ext, args, ref_d, qry_d, offs_d, ref_h, qry_h, int(shared))
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def ncu_traffic(kernel: str, config: dict):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full
    capture of this workload (profiles/ncu_traffic.json); None when the workload differs."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    try:
        d = json.loads(p.read_text())
        w = d["workload"]
        if all(config.get(key) == w[key] for key in ("reads_per_set", "read_len", "k", "t")):
            return d["kernels"][kernel]["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def random_sector_ceiling(filter_bytes: int = 1 << 32):
    """measured random 32-byte-sector load ceiling for the filter's residency (profiles/r01_ceilings.json), in GB/s:
    over a 4 GiB buffer (DRAM) for filters larger than L2, over a 64 MiB buffer (L2-resident) otherwise"""
    try:
        d = json.loads((ROOT / "profiles" / "r01_ceilings.json").read_text())
        key = "dram_4GiB_load_Gsectors_s" if filter_bytes > (64 << 20) else "l2_64MiB_load_Gsectors_s"
        return d[key] * 32
    except Exception:
        return None


def _mem_procs(filter_bytes: int) -> int:
    """how many reference processes fit in RAM: each callocs (and touches) the whole filter"""
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                avail = int(ln.split()[1]) * 1024
                return max(1, int(avail * 0.5 // (filter_bytes + (1 << 30))))
    except Exception:
        pass
    return 1


def reference_arm(args, rank, world, config):
    """bench.py --impl reference: the reference's CPU index_and_search on the host cores."""
    if rank != 0:
        return 0
    n, L, k, t = args.reads, args.length, args.k, args.t
    sample = min(args.cpu_sample, n)
    rng = np.random.default_rng(1000)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    ref = acgt[rng.integers(0, 4, size=sample * L)]
    # same query recipe as make_sets_torch, on the sample
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    r2 = ref.reshape(sample, L)
    src = rng.integers(0, sample, size=sample)
    cp = r2[src]
    rc = rng.random(sample) < 0.5
    cp = np.where(rc[:, None], comp[cp[:, ::-1]], cp)
    mut = rng.random((sample, L)) < 0.01
    cp = np.where(mut, acgt[rng.integers(0, 4, size=(sample, L))], cp)
    shared = rng.random(sample) < 0.5
    qry = np.where(shared[:, None], cp, acgt[rng.integers(0, 4, size=(sample, L))]).astype(np.uint8).reshape(-1)
    # The reference is single-threaded; its only parallelism is independent processes.  Every process
    # must index the whole reference sample (and calloc + touch its own 2^(k-1)-byte filter), so more processes
    # only split the search half of the job and contend for DRAM during the index half: the process count that
    # is fastest on this box is found by trying 1, 4 and the most that fit, and the best one is the arm's value.
    most = max(1, min(os.cpu_count() or 1, 16, _mem_procs(1 << (k - 1))))
    tried = {}
    for procs in sorted({1, min(4, most), most}):
        r = cpu_reference_run(ref, qry, L, sample, sample, k, t, procs)
        if r is None:
            r = cpu_port_run(ref, qry, L, sample, sample, k, t)
        tried[r[2]] = r
        if r[1] == "port":
            break
    best = min(tried.values())
    runs = 1
    for _ in range(max(0, min(args.steps, 3) - 1)):      # repeat the best configuration, keep its fastest run
        r = cpu_reference_run(ref, qry, L, sample, sample, k, t, best[2]) if best[1] == "reference" else None
        if r is not None:
            best = min(best, r)
            runs += 1
    wall, kind, cores = best
    v = sample / wall
    # `steps` / `warmup` are what this arm really ran: `runs` timed passes of the fastest process count over the SAMPLE
    # (ms_per_step = one such pass), after len(tried) - 1 exploratory passes with other process counts
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": runs,
            "warmup": len(tried) - 1, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "sample_of": {"reads_per_set": n, "sample_reads_per_set": sample, "fraction": sample / n,
                          "note": "the filter of the sample holds sample/reads_per_set of the k-mers of the full set (5x emptier at "
                                  "the default 2 M of 10 M): fewer b/c/d probes per lookup than at full size, i.e. the CPU rate is, if "
                                  "anything, flattered"},
            "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                             "host_cpus": os.cpu_count(),
                             "tried_reads_per_s": {str(c): round(sample / r[0], 1) for c, r in sorted(tried.items())},
                             "sample": f"{sample} reference reads indexed by each of {cores} process(es), {sample} query "
                                       f"reads split across them ({L} bp, k={k}, t={t}); value = query reads / wall of the "
                                       "fastest process count tried"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
