#!/bin/bash
# round 2, GPU call AA: placement kernels -- near test on the held values; 2 / 4 / 8 particles per thread
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rdf_gpu.py -x -q > gpurun_out/aa_pytest.log 2>&1; tail -1 gpurun_out/aa_pytest.log
for v in default ilp1 ilp3; do
  echo "== $v"
  if [ "$v" = default ]; then timeout 200 python tools/build_probe.py; else AMEPB_LIB=$PWD/amep_b200/lib/libamepb_$v.so timeout 200 python tools/build_probe.py; fi
done 2>&1 | grep -v "^+"
