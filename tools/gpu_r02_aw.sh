#!/bin/bash
# round 2, GPU call AW: full GPU suite, smoke, default bench line, reference arm, ncu launch list of a short bench
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/aw_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/aw_pytest.log; tail -5 gpurun_out/aw_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/aw_smoke.log 2>&1; tail -2 gpurun_out/aw_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/aw_bench.json 2> gpurun_out/aw_bench.err; echo "bench rc=$?"; tail -c 500 gpurun_out/aw_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/aw_ref.json 2> gpurun_out/aw_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_v7.csv python bench.py --steps 2 --warmup 3 --frames 64 --e2e-frames 32 --no-cpu --no-strong --no-sq > gpurun_out/aw_ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/print_bench.py gpurun_out/aw_bench.json | head
