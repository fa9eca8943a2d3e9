#!/bin/bash
# round 2, GPU call N: two targets per lane (experiment) vs the kept kernel; ncu of the kept kernel
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered cfg5 > gpurun_out/n_quick_base.jsonl 2>&1; cut -c1-330 gpurun_out/n_quick_base.jsonl
AMEPB_LIB=$PWD/amep_b200/lib/libamepb_two.so timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered cfg5 > gpurun_out/n_quick_two.jsonl 2>&1; cut -c1-330 gpurun_out/n_quick_two.jsonl
AMEPB_LIB=$PWD/amep_b200/lib/libamepb_two.so timeout 600 python -m pytest tests/test_rdf_gpu.py -x -q > gpurun_out/n_pytest_two.log 2>&1; tail -2 gpurun_out/n_pytest_two.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -c 1 -f -o gpurun_out/r02_pair_cfg2_v13 python tools/profile_case.py cfg2 > gpurun_out/m_ncu_pair.log 2>&1; tail -2 gpurun_out/m_ncu_pair.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -c 1 -f -o gpurun_out/r02_pair_3d_v13 python tools/profile_case.py 3d > gpurun_out/m_ncu_pair3d.log 2>&1; tail -2 gpurun_out/m_ncu_pair3d.log
