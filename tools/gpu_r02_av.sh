#!/bin/bash
# round 2, GPU call AV: S(q) suite with the q-shard test
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sq_gpu.py tests/test_abi_host.py tests/test_order_gpu.py -x -q > gpurun_out/av_pytest.log 2>&1; tail -2 gpurun_out/av_pytest.log
