#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python tools/build_probe.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_rdf_gpu.py tests/test_baseline_configs_gpu.py tests/test_sq_gpu.py -x -q > gpurun_out/q_pytest.log 2>&1; tail -2 gpurun_out/q_pytest.log
timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered cfg5 cfg1 2>&1 | cut -c1-300
