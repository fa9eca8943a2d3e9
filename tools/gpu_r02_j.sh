#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rdf_gpu.py tests/test_baseline_configs_gpu.py -x -q > gpurun_out/j_pytest.log 2>&1; tail -3 gpurun_out/j_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-strong --no-sq > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc=$?"; tail -c 800 gpurun_out/j_bench.err
python tools/print_bench.py gpurun_out/j_bench.json 2>/dev/null | head -30
