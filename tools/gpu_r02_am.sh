#!/bin/bash
# round 2, GPU call AM (2 GPUs, final code): sharded evaluate classes vs single rank, bench line at N=2
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/am_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -v "^W\|^\[W" gpurun_out/am_dist_check.log | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/am_bench2.json 2> gpurun_out/am_bench2.err; echo "bench rc=$?"; tail -c 600 gpurun_out/am_bench2.err; head -c 1500 gpurun_out/am_bench2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/am_ref2.json 2>&1; echo "ref rc=$?"; head -c 400 gpurun_out/am_ref2.json
