#!/bin/bash
# round 2, GPU call E: source-level ncu capture of the pair kernel (cfg2, 4 frames)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -c 1 -f -o gpurun_out/r02_pair_cfg2_v12 python tools/profile_case.py cfg2 > gpurun_out/e_ncu_pair.log 2>&1; tail -3 gpurun_out/e_ncu_pair.log
