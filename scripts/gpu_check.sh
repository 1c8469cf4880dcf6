#!/bin/bash
# One GPU-box pass: parity tests, a bench line, the ncu launch list and one full capture of the top kernel.
# usage: scripts/gpu_check.sh <tag> [bench args...]     (run under gpurun from the repo root)
set -u
tag=${1:-r1}; shift || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
(node --version || echo "node: absent") > gpurun_out/${tag}_node.txt 2>&1
nproc >> gpurun_out/${tag}_node.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/${tag}_node.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 1200 python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
