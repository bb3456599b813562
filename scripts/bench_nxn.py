"""C3-shaped run (BASELINE.json configs[2]): S synthetic sets x R reads x L bp, full N x N comparison.

    python scripts/bench_nxn.py [--sets 10] [--reads 2000000] [--len 150] [-k 33] [--gpus N] [--flow] [--ref-sample R]

  commet_nxn : one process, resident sets, rounds spread over the GPUs          (commet_b200/bin/commet_nxn)
  --flow     : the same comparison as Commet.py drives it, one drop-in tool process per round (tests/commet_flow.py),
               checked byte-for-byte against commet_nxn's CSVs
  --ref-sample R : the reference's own binaries (oracle/_ref) on the first R reads of the first 3 sets, extrapolated
               per round, as the CPU yardstick (full scale is ~20 CPU-hours)
Sets: a pool of R random reads; every read of a set is, with probability 1/2, a copy of a pool read (half of them
reverse-complemented, 1 % substitutions), else a fresh random read (SURVEY 8d recipe).  GPU box only."""
import argparse, json, os, shutil, subprocess, sys, tempfile, time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def write_fasta_fixed(path, arr, first=0, append=False):
    """arr: (n, L) uint8 ACGT -> '>%09d\\n' + L bases + '\\n' per record (numbered from `first`), vectorised"""
    n, L = arr.shape
    rec = np.empty((n, 11 + L + 1), dtype=np.uint8)
    rec[:, 0] = ord(">")
    idx = np.arange(first, first + n)
    for d in range(9):
        rec[:, 9 - d] = ord("0") + (idx // 10 ** d) % 10
    rec[:, 10] = ord("\n")
    rec[:, 11:11 + L] = arr
    rec[:, 11 + L] = ord("\n")
    with open(path, "ab" if append else "wb") as f:
        rec.tofile(f)


def write_fasta_blocks(path, blocks):
    first = 0
    for i, arr in enumerate(blocks):
        write_fasta_fixed(path, arr, first, append=i > 0)
        first += arr.shape[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", type=int, default=10)
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--len", type=int, default=150, dest="length")
    ap.add_argument("-k", type=int, default=33)
    ap.add_argument("-t", type=int, default=2)
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--flow", action="store_true")
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import torch
    from commet_b200 import build
    build.build_all()
    dev = torch.device("cuda", 0)
    S, R, L = args.sets, args.reads, args.length
    need = S * R * (L + 12) * 1.1                  # FASTA bytes + outputs
    base = None
    for cand in ("/dev/shm", tempfile.gettempdir()):
        if os.path.isdir(cand) and shutil.disk_usage(cand).free > need:
            base = cand
            break
    work = Path(tempfile.mkdtemp(prefix="nxn_", dir=base))
    g = torch.Generator(device=dev)
    g.manual_seed(1000)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    comp = torch.zeros(256, dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    pool = torch.empty((R, L), dtype=torch.uint8, device=dev)
    for b0 in range(0, R, 2_000_000):              # in blocks: the int64 index tensor of a 20 M x 150 draw alone is 24 GB
        m = min(2_000_000, R - b0)
        pool[b0:b0 + m] = acgt[torch.randint(0, 4, (m, L), generator=g, device=dev)]
    t0 = time.perf_counter()
    lines = []
    from concurrent.futures import ThreadPoolExecutor
    pool_w = ThreadPoolExecutor(max_workers=6)     # formatting + writing a set overlaps the generation of the next ones
    pending = []
    BLOCK = 2_000_000                              # reads generated at a time: bounds the device memory of the recipe
    for s in range(S):
        g.manual_seed(2000 + s)
        blocks = []
        for b0 in range(0, R, BLOCK):
            m = min(BLOCK, R - b0)
            src = torch.randint(0, R, (m,), generator=g, device=dev)
            cp = pool[src]
            rc = torch.rand(m, generator=g, device=dev) < 0.5
            cp = torch.where(rc[:, None], comp[cp.flip(1).long()], cp)
            mut = torch.rand((m, L), generator=g, device=dev) < 0.01
            rnd = acgt[torch.randint(0, 4, (m, L), generator=g, device=dev)]
            cp = torch.where(mut, rnd, cp)
            shared = torch.rand(m, generator=g, device=dev) < 0.5
            fresh = acgt[torch.randint(0, 4, (m, L), generator=g, device=dev)]
            blocks.append(torch.where(shared[:, None], cp, fresh).cpu().numpy())
        pending.append(pool_w.submit(write_fasta_blocks, work / f"set{s}.fa", blocks))
        lines.append(f"set{s}:set{s}.fa")
        if s < 3 and args.ref_sample:
            write_fasta_fixed(work / f"sample{s}.fa", blocks[0][:args.ref_sample])
        del blocks
    for f in pending:
        f.result()
    (work / "cfg.txt").write_text("\n".join(lines) + "\n")
    del pool
    torch.cuda.empty_cache()
    gen_s = time.perf_counter() - t0
    res = {"workload": f"C3 shape: {S} sets x {R} reads x {L} bp, k={args.k} t={args.t}, full N x N ({S * S - 1} index_and_search rounds)",
           "fasta_bytes": sum((work / f"set{s}.fa").stat().st_size for s in range(S)), "generate_s": round(gen_s, 2)}
    kt = ["-k", str(args.k), "-t", str(args.t)]
    for rep in range(args.reps):
        t0 = time.perf_counter()
        cmd = [str(build.BIN / "commet_nxn"), "cfg.txt", "-o", "nxn_out/", "-q", "--report", "report.json", *kt]
        if args.gpus:
            cmd += ["--gpus", str(args.gpus)]
        r = subprocess.run(cmd, cwd=work, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        wall = time.perf_counter() - t0
        rep_j = json.loads((work / "report.json").read_text())
        res[f"commet_nxn_run{rep}"] = {"wall_s": round(wall, 3), **rep_j}
    last = f"commet_nxn_run{args.reps - 1}"
    ph = res[last]["seconds_at_end_of"]
    rounds_s = ph["rounds"] - ph["stage_and_filter"]
    res["rounds_per_s"] = (S * S - 1) / rounds_s
    # every round searches the reads of its query sets: "all in Si" N-1-i sets, the two refinement rounds one each
    searched = sum((S - 1 - i) * R + 2 * (S - 1 - i) * R for i in range(S - 1))
    res["query_reads_per_s_rounds_only"] = searched / rounds_s
    res["query_reads_per_s_whole_run"] = searched / res[last]["wall_s"]
    res["matrix_plain"] = (work / "nxn_out" / "matrix_plain.csv").read_text()
    if args.flow:
        from tests import commet_flow
        t0 = time.perf_counter()
        csv = commet_flow.run("cfg.txt", build.BIN, work, out_dir="flow_out/", k=args.k, t=args.t)
        res["tool_flow_wall_s"] = round(time.perf_counter() - t0, 3)
        same = all(csv[n] == (work / "nxn_out" / f"matrix_{n}.csv").read_text() for n in ("plain", "percentage", "normalized"))
        bv_same = all((work / "flow_out" / p.name).read_bytes() == p.read_bytes() for p in (work / "nxn_out").glob("*_in_*.bv"))
        res["tool_flow_identical_csv_and_bv"] = bool(same and bv_same)
    if args.ref_sample:
        from oracle import oracle
        if oracle.have_ref():
            (work / "a.txt").write_text("a:sample0.fa\n")
            (work / "q.txt").write_text("b:sample1.fa\nc:sample2.fa\n")
            t0 = time.perf_counter()
            r = subprocess.run([str(oracle.REF_DIR / "index_and_search"), "-i", "a.txt", "-s", "q.txt", "-o", "ref_out", "-l", "ref_out",
                                *kt], cwd=work, capture_output=True, text=True)
            dt = time.perf_counter() - t0
            res["reference_cpu"] = {"sample_reads_per_set": args.ref_sample, "one_round_1_index_2_query_sets_s": round(dt, 2),
                                    "query_reads_per_s_1_core": 2 * args.ref_sample / dt, "returncode": r.returncode}
    print(json.dumps(res, indent=1))
    if args.out:
        Path(args.out).write_text(json.dumps(res, indent=1) + "\n")
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
