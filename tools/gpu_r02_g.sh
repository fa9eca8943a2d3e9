#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_paircor_gpu.py tests/test_order_gpu.py -x -q > gpurun_out/g_pytest.log 2>&1; tail -40 gpurun_out/g_pytest.log
