#!/bin/bash
# round 2, GPU call AQ: the whole RDF / evaluate test suites with the partitioned build forced for every eligible call
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
AMEPB_FORCE_PARTITIONED_BUILD=1 timeout 900 python -m pytest tests/test_rdf_gpu.py tests/test_baseline_configs_gpu.py tests/test_abi_host.py -q -k "not build_equals and not random_configurations and not per_frame_boxes_part" > gpurun_out/aq_pytest.log 2>&1; tail -3 gpurun_out/aq_pytest.log
