#!/bin/bash
# round 2, GPU call Z: compact (x, y) records of flat float32 frames, statistics in the element type
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rdf_gpu.py tests/test_baseline_configs_gpu.py -x -q > gpurun_out/z_pytest.log 2>&1; tail -2 gpurun_out/z_pytest.log
timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered 2>&1 | cut -c1-330 | tee gpurun_out/z_quick_compact.jsonl
AMEPB_NO_COMPACT_RECORDS=1 timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered 2>&1 | cut -c1-330 | tee gpurun_out/z_quick_vec4.jsonl
timeout 200 python tools/build_probe.py 2>&1 | tail -1
AMEPB_NO_COMPACT_RECORDS=1 timeout 200 python tools/build_probe.py 2>&1 | tail -1
