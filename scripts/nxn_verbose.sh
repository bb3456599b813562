#!/bin/bash
# verbose commet_nxn run on synthetic sets: $1 sets, $2 reads per set (diagnostic)
set -e
W=/dev/shm/nxn_v; rm -rf $W; mkdir -p $W; cd $W
python - $1 $2 <<PY
import sys, numpy as np
sys.path.insert(0, "/root/repo/scripts")
from bench_nxn import write_fasta_fixed
S, R = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(0)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
pool = acgt[rng.integers(0, 4, size=(R, 150))]
for s in range(S):
    a = acgt[rng.integers(0, 4, size=(R, 150))]
    m = rng.random(R) < 0.3
    a[m] = pool[rng.integers(0, R, size=int(m.sum()))]
    write_fasta_fixed(f"set{s}.fa", a)
open("cfg.txt", "w").write("".join(f"set{s}:set{s}.fa\n" for s in range(S)))
PY
/root/repo/commet_b200/bin/commet_nxn cfg.txt -k 33 --report rep.json -o out/
cat rep.json
rm -rf $W
