"""A minimal stand-in for the reference orchestrator, for the GPU box where /root/reference is absent.

It issues the same tool invocations Commet.py does for a non-SGE run (Commet.py:103-121 filtering,
:186-240 the three index_and_search rounds per reference set, :245-317 the matrices read back through
`bvop -i`) against any directory of binaries.  tests/test_flow_cpu.py pins it against the real Commet.py
(both driving the reference binaries) so the GPU test can trust it.
"""
from __future__ import annotations

import os
import subprocess
from pathlib import Path


def parse_config(path):
    """-> (set names, files per set, bvs per set or None).  Commet.py:42-95."""
    names, files, bvs = [], [], []
    lines = [ln for ln in Path(path).read_text().split("\n") if ln.strip()]
    with_bv = "," in Path(path).read_text().split("\n")[0]
    for ln in lines:
        name, rest = ln.split(":", 1)[0], ln.split(":")[1]
        items = [x.strip() for x in rest.split(";")]
        names.append(name.strip())
        files.append([x.split(",")[0] for x in items])
        if with_bv:
            bvs.append([x.split(",")[1] for x in items])
    return names, files, (bvs if with_bv else None)


def _run(cmd, cwd, log):
    r = subprocess.run(cmd, shell=True, cwd=cwd, capture_output=True, text=True)
    log.append((cmd, r.returncode))
    return r.stdout


def run(config, bin_dir, cwd, out_dir="output_commet/", k=33, t=2, l=0, n=-1, e=0, m=-1, tmp="tmpflow"):
    """Returns dict(plain=..., percentage=..., normalized=...) CSV texts; .bv files are left in out_dir."""
    cwd = Path(cwd)
    bin_dir = str(bin_dir).rstrip("/") + "/"
    if not out_dir.endswith("/"):
        out_dir += "/"
    (cwd / out_dir).mkdir(parents=True, exist_ok=True)
    log = []
    names, files, bvs = parse_config(cwd / config if not os.path.isabs(config) else config)
    if l < k * t:
        if l != 0:
            l = k * t
    if bvs is None:
        opts = f" -l {l} -e {e}" + (f" -n {n}" if n >= 0 else "")
        for fl in files:
            mopt = f" -m {m / len(fl)}" if m >= 0 else ""
            for f in fl:
                _run(f"{bin_dir}filter_reads {f}{opts}{mopt} -o {out_dir}{os.path.basename(f)}.bv", cwd, log)
        bvs = [[out_dir + os.path.basename(f) + ".bv" for f in fl] for fl in files]

    def fof(i, bv_of=None):
        items = [f + "," + (b if bv_of is None else bv_of(f)) for f, b in zip(files[i], bvs[i])]
        return names[i] + ":" + ";".join(items)

    for i, nm in enumerate(names):
        (cwd / f"{nm}_{tmp}.txt").write_text(fof(i))
    kt = f" -t {t} -k {k} "
    N = len(names)
    for ref in range(N - 1):
        qname = f"queries_for_index_{names[ref]}_{tmp}.txt"
        (cwd / qname).write_text("".join(fof(j) + "\n" for j in range(ref + 1, N)))
        _run(f"{bin_dir}index_and_search -i {names[ref]}_{tmp}.txt -s {qname} -o {out_dir}{kt} -l {out_dir}", cwd, log)
        for i in range(ref + 1, N):
            # X restricted to (X in Sref) indexed, Sref searched
            iname = f"index_{names[i]}_previous_{names[ref]}_{tmp}.txt"
            (cwd / iname).write_text(fof(i, lambda f: out_dir + os.path.basename(f) + "_in_" + os.path.basename(names[ref]) + ".bv"))
            _run(f"{bin_dir}index_and_search -i {iname} -s {names[ref]}_{tmp}.txt -o {out_dir}{kt} -l {out_dir}", cwd, log)
            # Sref restricted to (Sref in X) indexed, X searched: overwrites X_in_Sref
            iname = f"index_{names[ref]}_previous_{names[i]}_{tmp}.txt"
            (cwd / iname).write_text(fof(ref, lambda f: out_dir + os.path.basename(f) + "_in_" + os.path.basename(names[i]) + ".bv"))
            _run(f"{bin_dir}index_and_search -i {iname} -s {names[i]}_{tmp}.txt -o {out_dir}{kt} -l {out_dir}", cwd, log)

    def selected(bv):
        return int(_run(f"{bin_dir}bvop {bv} -i", cwd, log).split("\n")[-2].split()[0])

    totals = [sum(selected(b) for b in bvs[i]) for i in range(N)]
    shared = [[totals[i] if i == j else sum(selected(f"{out_dir}{os.path.basename(f)}_in_{names[j]}.bv") for f in files[i])
               for j in range(N)] for i in range(N)]
    head = "".join(";" + nm for nm in names) + "\n"
    plain = head + "".join(names[i] + "".join(";" + str(shared[i][j]) for j in range(N)) + "\n" for i in range(N))
    pct = head + "".join(names[i] + "".join(";" + str(100 * shared[i][j] / float(totals[i])) for j in range(N)) + "\n"
                         for i in range(N))
    norm = head + "".join(names[i] + "".join(";" + str(100 * (shared[i][j] + shared[j][i]) / float(totals[i] + totals[j]))
                                             for j in range(N)) + "\n" for i in range(N))
    for p in cwd.glob(f"*{tmp}*"):
        p.unlink()
    bad = [c for c, rc in log if rc != 0]
    if bad:
        raise RuntimeError(f"tool failed: {bad[0]}")
    return dict(plain=plain, percentage=pct, normalized=norm)
