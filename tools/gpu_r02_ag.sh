#!/bin/bash
# round 2, GPU call AG: per-kernel times of the partitioned build (ncu launch list of tools/build_probe.py)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:"strip|scan|place|stats" -s 30 -c 30 --csv --log-file gpurun_out/ag_launches.csv python tools/build_probe.py > gpurun_out/ag.log 2>&1; tail -1 gpurun_out/ag.log
