#!/bin/bash
# round 2, call w: host pipelining A/B of commet_index_and_search on C2 (index-set cut positions, query parts);
# full ncu captures of the C5 kernels (fused staging + selection, k_filter, k_bvop, k_popcount)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python scripts/sweep_e2e_parts.py > gpurun_out/r02w_e2e_parts.txt 2> gpurun_out/r02w_e2e_parts.err; echo "e2e sweep rc=$?"; cat gpurun_out/r02w_e2e_parts.txt; tail -3 gpurun_out/r02w_e2e_parts.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_stage_filter|k_filter|k_bvop|k_popcount|k_encode' -c 12 -f -o gpurun_out/r02w_c5_full \
    python scripts/sweep_c5.py --reads 16000000 --batches 1 --reps 1 > gpurun_out/r02w_c5_full.log 2>&1; echo "c5 full capture rc=$?"; ls -la gpurun_out/r02w_c5_full.ncu-rep
