#!/bin/bash
# round 2, GPU call AO: query-loop unroll of the multi-target pair loop (2 / 4 / 6 / 8), 2D uniform, clustered, 3D
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
for v in default qu4 qu6 qu8; do
  echo "== $v"
  if [ "$v" = default ]; then timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered cfg5fast; else AMEPB_LIB=$PWD/amep_b200/lib/libamepb_$v.so timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered cfg5fast; fi
done 2>&1 | grep -v "^+" | cut -c1-300
