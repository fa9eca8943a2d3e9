#!/bin/bash
# round 2, GPU call X: pcf_angle with the (r, theta) grid in shared memory
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_paircor_gpu.py -x -q > gpurun_out/x_pytest.log 2>&1; tail -2 gpurun_out/x_pytest.log
timeout 300 python tools/profile_paircor.py 100000 all
AMEPB_PCFA_GLOBAL_ATOMICS=1 timeout 300 python tools/profile_paircor.py 100000 pcfa
