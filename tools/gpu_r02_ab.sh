#!/bin/bash
# round 2, GPU call AB: count pass with the frame as the fast block index
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rdf_gpu.py tests/test_baseline_configs_gpu.py -x -q > gpurun_out/ab_pytest.log 2>&1; tail -1 gpurun_out/ab_pytest.log
timeout 200 python tools/build_probe.py | tail -1
AMEPB_NO_FRAME_FAST=1 timeout 200 python tools/build_probe.py | tail -1
timeout 200 python tools/build_probe.py | tail -1
AMEPB_NO_FRAME_FAST=1 timeout 200 python tools/build_probe.py | tail -1
