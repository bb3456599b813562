"""Per-phase breakdown of one kernel from an ncu report's source page (needs --import-source on / -lineinfo):
the SASS listing is cut at every BAR.SYNC; for each segment: share of stall samples, of executed warp instructions,
shared-memory wavefronts (actual / ideal); then the opcode histogram.  Runs here, no GPU.

    python scripts/ncu_phases.py gpurun_out/r01b_full.ncu-rep k_bin_scatter
"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, pat):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not heads:
        sys.exit("no source page for " + pat)
    h = rows[heads[0]]
    d = rows[heads[0] + 1:(heads[1] - 1 if len(heads) > 1 else len(rows))]
    print(rows[heads[0] - 1][1][:140])
    col = {n: (h.index(n) if n in h else None)
           for n in ("Warp Stall Sampling (All Samples)", "Source", "Instructions Executed", "L1 Wavefronts Shared",
                     "L1 Wavefronts Shared Ideal", "Thread Instructions Executed")}      # no shared-memory columns without smem
    val = lambda r, n: int(r[col[n]] or 0) if col[n] is not None else 0
    tot_s = sum(val(r, "Warp Stall Sampling (All Samples)") for r in d) or 1
    tot_x = sum(val(r, "Instructions Executed") for r in d) or 1
    tot_t = sum(val(r, "Thread Instructions Executed") for r in d)
    print(f"{len(d)} SASS instructions, {tot_x / 1e6:.0f} M warp instructions executed, {tot_t / tot_x:.1f} active lanes on average, {tot_s} stall samples")
    cur, first = collections.Counter(), 0
    print("segments between BAR.SYNC:")
    for i, r in enumerate(d):
        src = r[col["Source"]].strip()
        cur["s"] += val(r, "Warp Stall Sampling (All Samples)")
        cur["x"] += val(r, "Instructions Executed")
        cur["w"] += val(r, "L1 Wavefronts Shared")
        cur["wi"] += val(r, "L1 Wavefronts Shared Ideal")
        if src.startswith("BAR.SYNC") or i == len(d) - 1:
            print(f"  SASS #{first:4d}..{i:4d}  samples {100 * cur['s'] / tot_s:5.1f} %  warp instr {100 * cur['x'] / tot_x:5.1f} % ({cur['x'] / 1e6:6.0f} M)"
                  f"  smem wavefronts {cur['w'] / 1e6:7.1f} M (ideal {cur['wi'] / 1e6:7.1f} M)")
            cur, first = collections.Counter(), i + 1
    ops, samp = collections.Counter(), collections.Counter()
    for r in d:
        parts = r[col["Source"]].strip().split()
        o = parts[1] if parts[0].startswith("@") else parts[0]
        o = o.split(".")[0]
        ops[o] += val(r, "Instructions Executed")
        samp[o] += val(r, "Warp Stall Sampling (All Samples)")
    print("opcodes:")
    for o, c in ops.most_common(14):
        print(f"  {o:8s} {c / 1e6:7.0f} M  {100 * c / tot_x:5.1f} % of instructions  {100 * samp[o] / tot_s:5.1f} % of samples")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
