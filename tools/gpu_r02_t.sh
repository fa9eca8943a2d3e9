#!/bin/bash
# round 2, GPU call T: source-level ncu captures of the committed pair kernel (four staged targets per lane)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -c 1 -f -o gpurun_out/r02_pair_cfg2_v14 python tools/profile_case.py cfg2 > gpurun_out/t_ncu_pair.log 2>&1; tail -2 gpurun_out/t_ncu_pair.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -c 1 -f -o gpurun_out/r02_pair_3d_v14 python tools/profile_case.py 3d > gpurun_out/t_ncu_pair3d.log 2>&1; tail -2 gpurun_out/t_ncu_pair3d.log
