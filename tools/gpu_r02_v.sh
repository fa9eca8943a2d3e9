#!/bin/bash
# round 2, GPU call V: sfiso('std') in the reference's float32 summation order (amepb_sq_debye_ref)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sq_gpu.py tests/test_abi_host.py -x -q -s > gpurun_out/v_pytest.log 2>&1; grep -E "reference order|passed|failed|Error" gpurun_out/v_pytest.log | tail -20
