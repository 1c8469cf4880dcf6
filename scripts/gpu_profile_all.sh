#!/bin/bash
# ncu evidence for profiles/: a plain bench line, the launch list of the same command, and ONE --set full capture that
# holds the first launch of every major kernel (the first time chunk: every stream live, every slot used).
#   usage: scripts/gpu_profile_all.sh <tag> [bench args...]      (run under gpurun from the repo root)
set -u
tag=$1; shift
mkdir -p gpurun_out
CMD="python bench.py --no-cpu-baseline --steps 2 --warmup 1 $*"
K='^(analysis|map_energy|map_smooth|map_peaks|map_fmapply|freqest|preterms|carry|chain|isynth|ola|synth)'
$CMD > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err; echo "plain rc=$?"
tail -c 1500 gpurun_out/${tag}_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"$K" -c 9 -f -o gpurun_out/${tag}_prof $CMD > gpurun_out/${tag}_ncu2.log 2>&1
echo "full capture rc=$?"
tail -n 3 gpurun_out/${tag}_ncu1.log gpurun_out/${tag}_ncu2.log
ls -la gpurun_out/${tag}_prof.ncu-rep
