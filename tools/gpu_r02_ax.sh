#!/bin/bash
# round 2, GPU call AX (2 GPUs): intra-frame sharding of pcf2d over the queries (amepb_pcf2d_set_query_offset)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pcf2d_gpu.py tests/test_abi_host.py -x -q > gpurun_out/ax_pytest.log 2>&1; tail -2 gpurun_out/ax_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dist_check.py > gpurun_out/ax_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -v "^W\|^\[W" gpurun_out/ax_dist_check.log | tail -15
