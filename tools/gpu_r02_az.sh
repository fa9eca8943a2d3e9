#!/bin/bash
# round 2, GPU call AZ (2 GPUs): intra-frame sharding of the spatialcor classes
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_paircor_gpu.py -x -q > gpurun_out/az_pytest.log 2>&1; tail -2 gpurun_out/az_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dist_check.py > gpurun_out/az_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -v "^W\|^\[W" gpurun_out/az_dist_check.log | tail -18
