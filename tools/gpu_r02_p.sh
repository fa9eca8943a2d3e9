#!/bin/bash
# round 2, GPU call P: which part of the placement kernels costs what (results of e1..e4 are wrong on purpose)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in default e1 e2 e3 e4; do
  echo "== $v"
  if [ "$v" = default ]; then timeout 200 python tools/build_probe.py; else AMEPB_LIB=$PWD/amep_b200/lib/libamepb_$v.so timeout 200 python tools/build_probe.py; fi
done 2>&1 | grep -v "^+" | tee gpurun_out/p_build_probe.txt
