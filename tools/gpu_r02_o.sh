#!/bin/bash
# round 2, GPU call O: targets per lane 2 (default build) vs 3 / 4 (5 blocks per SM) vs 4 (4 blocks per SM)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
: > gpurun_out/o_quick.jsonl
for v in default u1 u4; do
  echo "{\"variant\": \"$v\"}" >> gpurun_out/o_quick.jsonl
  if [ "$v" = default ]; then timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered cfg5 cfg1 >> gpurun_out/o_quick.jsonl 2>&1
  else AMEPB_LIB=$PWD/amep_b200/lib/libamepb_$v.so timeout 300 python tools/quick_rdf_bench.py cfg2fast cfg2clustered cfg5 cfg1 >> gpurun_out/o_quick.jsonl 2>&1; fi
done
python - <<'PY'
import json
for line in open('gpurun_out/o_quick.jsonl'):
    try: d=json.loads(line)
    except Exception: continue
    if 'variant' in d: print('==', d['variant'])
    elif 'k' in d and d.get('k')==0: print(d['case'][:32], 'ms/frame', d['ms_per_frame'], 'pair_ms', d['pair_ms'])
    elif 'k' not in d: print(d['case'][:32], 'ms/frame', d['ms_per_frame'], 'pair_ms', d['pair_ms'])
PY

timeout 900 python -m pytest tests/test_rdf_gpu.py tests/test_baseline_configs_gpu.py -x -q > gpurun_out/o_pytest.log 2>&1; tail -2 gpurun_out/o_pytest.log
