#!/bin/bash
# round 2, GPU call Y: ncu of the pcf_angle kernel (default rmax) and the spatialcor pair kernel, N = 100k
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pcf_angle_kernel -s 2 -c 1 -f -o gpurun_out/r02_pcfa_default_v2 python tools/profile_paircor.py 100000 pcfa > gpurun_out/y_ncu_pcfa.log 2>&1; tail -2 gpurun_out/y_ncu_pcfa.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scor_pair_kernel -c 1 -f -o gpurun_out/r02_scor_v1 python tools/profile_paircor.py 100000 scor > gpurun_out/y_ncu_scor.log 2>&1; tail -2 gpurun_out/y_ncu_scor.log
