#!/bin/bash
# GPU pass: parity tests, then (only if green) the profile script.  usage: scripts/gpu_check2.sh <tag> [bench args...]
set -u
tag=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; rc=$?
tail -15 gpurun_out/${tag}_pytest.log; echo "pytest rc=$rc"
if [ $rc -ne 0 ]; then exit $rc; fi
bash scripts/gpu_profile_all.sh $tag "$@"
