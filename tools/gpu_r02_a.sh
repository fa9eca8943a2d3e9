#!/bin/bash
# round 2, GPU call A: full GPU test suite, default bench line, ncu captures of the two sf2d kernels
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
free -g > gpurun_out/a_mem.txt; nproc >> gpurun_out/a_mem.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/a_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sf2d_c64_kernel -c 1 -f -o gpurun_out/r02_sf2d_c64_v6 python tools/profile_sq.py sf2d_c64 1 > gpurun_out/a_ncu_c64.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sf2d_f64_kernel -c 1 -f -o gpurun_out/r02_sf2d_f64_v3 python tools/profile_sq.py sf2d_f64 1 > gpurun_out/a_ncu_f64.log 2>&1
ls -la gpurun_out
