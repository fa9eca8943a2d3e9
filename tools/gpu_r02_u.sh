#!/bin/bash
# round 2, GPU call U: threshold table copies adjacent to the histogram halves (one address per pair) against the separate table
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rdf_gpu.py tests/test_baseline_configs_gpu.py -x -q > gpurun_out/u_pytest.log 2>&1; tail -3 gpurun_out/u_pytest.log
timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered cfg5 cfg1 2>&1 | cut -c1-330 | tee gpurun_out/u_quick_ab.jsonl
AMEPB_NO_ADJACENT_TABLES=1 timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered cfg5 cfg1 2>&1 | cut -c1-330 | tee gpurun_out/u_quick_table.jsonl
