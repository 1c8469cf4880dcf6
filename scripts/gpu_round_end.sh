#!/bin/bash
# Round evidence in one call: the whole -m gpu suite, then (scripts/gpu_final_profile.sh) the shim / other-config probes, the full
# hour of configs[4], the bench line + launch list + one ncu --set full capture, then configs 0, 1 and 3 on one GPU.
#   usage: scripts/gpu_round_end.sh <tag>
set -u
tag=$1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; rc=$?
tail -4 gpurun_out/${tag}_pytest.log; echo "pytest rc=$rc"
[ $rc -ne 0 ] && exit $rc
bash scripts/gpu_final_profile.sh ${tag}
bash scripts/gpu_configs.sh ${tag} 0 1 3
