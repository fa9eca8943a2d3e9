#!/bin/bash
# round 2, GPU call C: anisotropic cell experiments, dispatch probe, ncu of the placement kernels
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_order_gpu.py -x -q > gpurun_out/c_order.log 2>&1; tail -3 gpurun_out/c_order.log
: > gpurun_out/c_quick.jsonl
for cfg in "- -" "2 1" "3 1" "4 1" "1 2" "2 2" "3 2" "1 1"; do
  set -- $cfg
  if [ "$1" = "-" ]; then env -u AMEPB_KX -u AMEPB_KY timeout 200 python tools/quick_rdf_bench.py cfg2fast cfg2clustered >> gpurun_out/c_quick.jsonl 2>&1
  else AMEPB_KX=$1 AMEPB_KY=$2 timeout 200 python tools/quick_rdf_bench.py cfg2fast cfg2clustered >> gpurun_out/c_quick.jsonl 2>&1; fi
done
cut -c1-400 gpurun_out/c_quick.jsonl
timeout 120 ./tools/dispatch_probe.bin > gpurun_out/c_dispatch_probe.json 2>&1; cat gpurun_out/c_dispatch_probe.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:place_self_kernel -c 2 -f -o gpurun_out/r02_place_v9 python tools/profile_case.py cfg2 > gpurun_out/c_ncu_place.log 2>&1; tail -3 gpurun_out/c_ncu_place.log
