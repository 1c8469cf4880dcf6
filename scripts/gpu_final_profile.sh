#!/bin/bash
# Round evidence in one call: shim probe, other-config probes, cfg4 full hour, then the bench line + launch list + one
# ncu --set full capture of the first time chunk's kernels.   usage: scripts/gpu_final_profile.sh <tag>
set -u
tag=$1
mkdir -p gpurun_out
python scripts/probe_shim.py > gpurun_out/${tag}_shim.txt 2>&1; cat gpurun_out/${tag}_shim.txt | tail -4
python scripts/probe_configs.py > gpurun_out/${tag}_other_configs.txt 2>&1; tail -5 gpurun_out/${tag}_other_configs.txt | cut -c1-200
python scripts/probe_relay.py > gpurun_out/${tag}_relay_probe.txt 2>&1
CFG4_SECONDS=3600 bash scripts/gpu_configs.sh ${tag} 4
bash scripts/gpu_profile_all.sh ${tag}
