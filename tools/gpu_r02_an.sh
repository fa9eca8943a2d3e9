#!/bin/bash
# round 2, GPU call AN: ncu --set full of the final pair kernel on configs[1] ((x, y) records)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -c 1 -f -o gpurun_out/r02_pair_cfg2_v15 python tools/profile_case.py cfg2 > gpurun_out/an_ncu_pair.log 2>&1; tail -2 gpurun_out/an_ncu_pair.log
