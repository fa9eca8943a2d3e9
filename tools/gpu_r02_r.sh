#!/bin/bash
# round 2, GPU call R: streamed self placement vs the old kernels
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python tools/build_probe.py 2>&1 | tail -1
AMEPB_PLACE_OLD=1 timeout 200 python tools/build_probe.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_rdf_gpu.py tests/test_baseline_configs_gpu.py -x -q > gpurun_out/r_pytest.log 2>&1; tail -3 gpurun_out/r_pytest.log
timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered cfg5 cfg1 2>&1 | cut -c1-260
