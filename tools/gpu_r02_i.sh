#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
AMEPB_TRACE=1 timeout 300 python tools/e2e_trace.py > gpurun_out/i_trace.log 2>&1; grep -c "sub " gpurun_out/i_trace.log; grep "^rep\|wall" gpurun_out/i_trace.log | tail -5; grep "\[amepb\] sub" gpurun_out/i_trace.log | grep "copy" | tail -16
