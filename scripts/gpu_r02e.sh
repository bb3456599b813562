#!/bin/bash
# round 2, call e: scatter3 / apply3 parity, then A/B on C2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -m gpu -x -k "l2_blocked or k33 or sorted_insert or chunk_loop" > gpurun_out/r02e_insert_tests.txt 2>&1; echo "insert tests rc=$?"; tail -5 gpurun_out/r02e_insert_tests.txt
timeout 900 python scripts/ab_index.py COMMET_B200_SCATTER_FORM=2,COMMET_B200_APPLY_FORM=2,COMMET_B200_S2_TW=128,COMMET_B200_APPLY_BPS=8 \
   COMMET_B200_SCATTER_FORM=3,COMMET_B200_S2_TW=128,COMMET_B200_SCATTER_BPS=2 COMMET_B200_S2_TW=96,COMMET_B200_SCATTER_BPS=3 COMMET_B200_S2_TW=64,COMMET_B200_SCATTER_BPS=4 \
   COMMET_B200_S2_TW=96,COMMET_B200_SCATTER_BPS=3,COMMET_B200_APPLY_FORM=3,COMMET_B200_APPLY_BPS=6 COMMET_B200_APPLY_BPS=4 COMMET_B200_APPLY_BPS=5 \
   COMMET_B200_APPLY_TILE=4096,COMMET_B200_APPLY_BPS=4 COMMET_B200_APPLY_TILE=4096,COMMET_B200_APPLY_BPS=3 COMMET_B200_APPLY_TILE=2048,COMMET_B200_APPLY_BPS=6,COMMET_B200_APPLY_PREFETCH=0 > gpurun_out/r02e_ab.txt 2>&1; echo "ab rc=$?"; cat gpurun_out/r02e_ab.txt
