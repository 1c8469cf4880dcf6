#!/bin/bash
# one ncu --set full capture of one kernel of a bench.py command.  usage: scripts/gpu_ncu_one.sh <tag> <kernel-regex> <launch-skip> -- bench args...
set -u
tag=$1; kern=$2; skip=$3; shift 4
mkdir -p gpurun_out
CMD="python bench.py --no-cpu-baseline --no-parity-check --steps 1 --warmup 0 $*"
$CMD > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -f -o gpurun_out/${tag}_prof $CMD > gpurun_out/${tag}_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/${tag}_ncu.log | cut -c1-200
