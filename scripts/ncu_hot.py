"""Top stall-sampled SASS instructions of one kernel in an ncu report (source page). python scripts/ncu_hot.py rep kernel_regex [N]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
lines = raw.splitlines()
# first line: kernel name; second: header
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
iS, iSrc, iEx = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
data = [(int(r[iS] or 0), r[iSrc].strip(), int(r[iEx] or 0), i) for i, r in enumerate(rows[1:]) if len(r) > iS and r[iS].isdigit()]
tot = sum(d[0] for d in data)
print(lines[0][:150], "total samples", tot, "instr", len(data))
for s, src, ex, i in sorted(data, reverse=True)[:n]:
    print(f"{100*s/tot:5.1f}%  #{i:4d} exec={ex:>10d}  {src[:110]}")
