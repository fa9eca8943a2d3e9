#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for mb in 384 768 1536 3072; do
  AMEPB_STAGE_MB=$mb timeout 600 python bench.py --steps 3 --warmup 3 --no-strong --no-sq --no-cpu > gpurun_out/k_bench_$mb.json 2> gpurun_out/k_bench_$mb.err; echo "stage $mb rc=$?"
  python tools/print_bench.py gpurun_out/k_bench_$mb.json 2>/dev/null | head -3
done
