#!/bin/bash
# round 2, GPU call D: flattened gather in the pair kernel -- parity and timing
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rdf_gpu.py tests/test_baseline_configs_gpu.py tests/test_order_gpu.py -x -q > gpurun_out/d_pytest.log 2>&1; tail -4 gpurun_out/d_pytest.log
timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered cfg5 cfg1 > gpurun_out/d_quick.jsonl 2>&1; cut -c1-420 gpurun_out/d_quick.jsonl
