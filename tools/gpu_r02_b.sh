#!/bin/bash
# round 2, GPU call B: new order tests + whole suite, quick RDF bench per k (ILP placement), S(q) timing, sqrt probe
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_order_gpu.py -x -q > gpurun_out/b_order.log 2>&1; echo "order rc=$?" >> gpurun_out/b_order.log
tail -25 gpurun_out/b_order.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_order_gpu.py > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -8 gpurun_out/b_pytest.log
timeout 300 python tools/quick_rdf_bench.py cfg2 cfg5 > gpurun_out/b_quick.jsonl 2>&1
cat gpurun_out/b_quick.jsonl | cut -c1-330
timeout 300 python tools/profile_sq.py sf2d_c64 1 > gpurun_out/b_sq.log 2>&1; timeout 300 python tools/profile_sq.py sf2d_c64 8 >> gpurun_out/b_sq.log 2>&1; cat gpurun_out/b_sq.log
timeout 120 ./tools/sqrt_probe.bin > gpurun_out/b_sqrt_probe.json 2>&1; cat gpurun_out/b_sqrt_probe.json
