#!/bin/bash
# how long does the process take to go away?  (2-GPU diagnostic)
set -e
W=/dev/shm/nxn_ab; rm -rf $W; mkdir -p $W; cd $W
python - <<PY
import sys, numpy as np
sys.path.insert(0, "/root/repo/scripts")
from bench_nxn import write_fasta_fixed
rng = np.random.default_rng(0)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
for s in range(6):
    write_fasta_fixed(f"set{s}.fa", acgt[rng.integers(0, 4, size=(2_000_000, 150))])
open("cfg.txt", "w").write("".join(f"set{s}:set{s}.fa\n" for s in range(6)))
PY
for mode in quick return destroy quick; do
  for g in 1 2; do
    s=$(date +%s.%N)
    COMMET_NXN_EXIT=$mode /root/repo/commet_b200/bin/commet_nxn cfg.txt -k 33 -q --gpus $g --report rep.json -o out_$mode$g/
    e=$(date +%s.%N)
    python -c "import json;d=json.load(open('rep.json'));print('$mode gpus=$g wall', round($e-$s,3), 'phases_end', round(d['seconds_at_end_of']['vectors_and_matrices'],3))"
  done
done
rm -rf $W
