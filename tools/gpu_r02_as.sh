#!/bin/bash
# round 2, GPU call AS: threads and cached records of strip_sort_kernel
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
for v in default so512x8 so256x16 so128x16 so512x4; do
  echo "== $v"
  if [ "$v" = default ]; then timeout 200 python tools/build_probe.py; else AMEPB_LIB=$PWD/amep_b200/lib/libamepb_$v.so timeout 200 python tools/build_probe.py; fi
done 2>&1 | grep -v "^+"
