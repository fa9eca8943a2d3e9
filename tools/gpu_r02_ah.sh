#!/bin/bash
# round 2, GPU call AH: particles per thread of the strip tile kernels
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
for v in default sb5 sb6; do
  echo "== $v"
  if [ "$v" = default ]; then timeout 200 python tools/build_probe.py; else AMEPB_LIB=$PWD/amep_b200/lib/libamepb_$v.so timeout 200 python tools/build_probe.py; fi
done 2>&1 | grep -v "^+"
