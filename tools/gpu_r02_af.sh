#!/bin/bash
# round 2, GPU call AF: partitioned build of the query array (strip kernels) against place_self_kernel
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rdf_gpu.py tests/test_baseline_configs_gpu.py -x -q > gpurun_out/af_pytest.log 2>&1; tail -3 gpurun_out/af_pytest.log
timeout 200 python tools/build_probe.py | tail -1
AMEPB_NO_PARTITIONED_BUILD=1 timeout 200 python tools/build_probe.py | tail -1
timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered cfg5 2>&1 | cut -c1-330 | tee gpurun_out/af_quick.jsonl
