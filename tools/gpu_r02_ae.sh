#!/bin/bash
# round 2, GPU call AE (8 GPUs; rerun with the final code as call AR): sharded evaluate classes vs single rank, bench line at N=8
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/ae_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -v "^W\|^\[W" gpurun_out/ae_dist_check.log | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/ae_bench8.json 2> gpurun_out/ae_bench8.err; echo "bench rc=$?"; tail -c 600 gpurun_out/ae_bench8.err; head -c 1500 gpurun_out/ae_bench8.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > gpurun_out/ae_ref8.json 2>&1; echo "ref rc=$?"; head -c 400 gpurun_out/ae_ref8.json
