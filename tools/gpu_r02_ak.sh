#!/bin/bash
# round 2, GPU call AK: source-level ncu of the strip kernels (tools/build_probe.py)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"strip_tile|strip_sort" -s 6 -c 3 -f -o gpurun_out/r02_strip_v1 python tools/build_probe.py > gpurun_out/ak.log 2>&1; tail -1 gpurun_out/ak.log
