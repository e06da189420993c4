#!/bin/bash
# Round-2 single-GPU session H: ba_solve_cam_dsm variants (bit 0: mbarrier / st.async hand-over, bit 1: trailing update before the factorisation):
# stage times (product library), phase clocks (profiling library), split-pipeline parity tests per variant.
set -u
T=${1:-r2h}
O=gpurun_out
mkdir -p $O
for v in 0 1 2 3; do
  ICG_BA_DSM_VARIANT=$v ICG_BA_PROFILE=1 timeout 120 python scripts/prof_ba.py 32 3 20 2000 > $O/${T}_stages_v$v.log 2>&1
  ICG_BA_DSM_VARIANT=$v ICG_LIB_VARIANT=prof ICG_BA_PROFILE=1 timeout 120 python scripts/prof_ba.py 32 3 20 2000 > $O/${T}_clocks_v$v.log 2>&1
  echo "== variant $v"; grep -E "^rep|  solve|final_cost" $O/${T}_stages_v$v.log | tail -5; sed -n '/ba_solve_cam_dsm phases/,$p' $O/${T}_clocks_v$v.log
done
for v in 1 3; do
  ICG_BA_DSM_VARIANT=$v timeout 300 python -m pytest tests/test_ba_gpu.py -q -m gpu -x -k "cfg4 or sharded" > $O/${T}_pytest_split_v$v.log 2>&1; echo "pytest split v$v rc=$?"; tail -3 $O/${T}_pytest_split_v$v.log
done
