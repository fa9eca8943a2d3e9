#!/bin/bash
# round 2, GPU call AY: the whole GPU suite and smoke() at the end of the round
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/ay_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ay_pytest.log; tail -4 gpurun_out/ay_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ay_smoke.log 2>&1; tail -2 gpurun_out/ay_smoke.log
