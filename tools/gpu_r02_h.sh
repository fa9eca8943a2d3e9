#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rdf_gpu.py -x -q > gpurun_out/h_pytest.log 2>&1; tail -3 gpurun_out/h_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-strong > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?"; tail -c 800 gpurun_out/h_bench.err
python tools/print_bench.py gpurun_out/h_bench.json 2>/dev/null | head -30
