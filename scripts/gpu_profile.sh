#!/bin/bash
# ncu evidence for profiles/: launch list (per-launch device time of OUR kernels) and one --set full capture of the
# named kernel.   usage: scripts/gpu_profile.sh <tag> <kernel-regex> [bench args...]
set -u
tag=$1; kern=$2; shift 2
mkdir -p gpurun_out
CMD="python bench.py --no-cpu-baseline $*"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(analysis|map_energy|map_smooth|map_peaks|map_fmapply|freqest|preterms|carry|chain|isynth|ola)_kernel' -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$kern -s 3 -c 2 -f -o gpurun_out/${tag}_prof $CMD > gpurun_out/${tag}_ncu2.log 2>&1
echo "full capture rc=$?"
tail -n 3 gpurun_out/${tag}_ncu1.log gpurun_out/${tag}_ncu2.log
