"""One device step of bench.py from an ncu launch list (gpu__time_duration.sum per launch): kernel, grid, ms, share.

    python scripts/launch_summary.py profiles/r01_launches.csv > profiles/r01_launches_summary.txt

The list holds every launch of `bench.py --steps 2 --warmup 1 --no-cpu` (torch's generator kernels first, then the
warm-up, the timed device steps, the probe-counting pass and the first end-to-end steps); a device step starts at
the first k_encode after a k_search.  ncu serialises and cold-starts each launch: shares, not absolutes, are the
comparable figure.
"""
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi, gi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Metric Unit")
    seq = []
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("commet::", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui].startswith("n") else v / 1e3 if r[ui].startswith("u") else v
        seq.append((name, r[gi], v))
    ours = [s for s in seq if s[0].startswith("k_")]
    # steps: split after every k_search
    steps, cur = [], []
    for s in ours:
        cur.append(s)
        if s[0].startswith("k_search"):
            steps.append(cur)
            cur = []
    print(f"{len(seq)} launches captured, {len(ours)} of them commet kernels, {len(steps)} complete device steps")
    step = steps[1] if len(steps) > 1 else steps[0]      # the second one: first timed step (after the warm-up)
    tot = sum(s[2] for s in step)
    print(f"device step (second in the list): {len(step)} launches, {tot:.3f} ms of kernel time")
    for name, grid, ms in step:
        print(f"  {ms:9.3f} ms  {100 * ms / tot:5.1f} %  {name:28s} grid {grid}")
    agg = {}
    for name, grid, ms in step:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    print("by kernel:")
    for name, (c, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {ms:9.3f} ms  {100 * ms / tot:5.1f} %  x{c}  {name}")


if __name__ == "__main__":
    main(sys.argv[1])
