#!/bin/bash
# round 2, GPU call S: streamed self placement: timing + ncu capture
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python tools/build_probe.py 2>&1 | tail -1
AMEPB_PLACE_OLD=1 timeout 200 python tools/build_probe.py 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:place_self -c 2 -f -o gpurun_out/r02_place_stream_v2 python tools/build_probe.py > gpurun_out/s_ncu.log 2>&1; tail -2 gpurun_out/s_ncu.log
timeout 900 python -m pytest tests/test_rdf_gpu.py tests/test_baseline_configs_gpu.py -x -q > gpurun_out/s_pytest.log 2>&1; tail -3 gpurun_out/s_pytest.log
