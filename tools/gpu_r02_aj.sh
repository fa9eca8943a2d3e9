#!/bin/bash
# round 2, GPU call AJ: partitioned build against the placement kernels (new tests), random configurations, full RDF suite
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rdf_gpu.py tests/test_baseline_configs_gpu.py -x -q > gpurun_out/aj_pytest.log 2>&1; tail -3 gpurun_out/aj_pytest.log
