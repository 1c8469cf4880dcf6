#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r5g_pytest.log 2>&1; rc=$?
tail -6 gpurun_out/r5g_pytest.log; echo "pytest rc=$rc"
[ $rc -ne 0 ] && exit $rc
bash scripts/gpu_profile_all.sh r5g
CFG4_SECONDS=120 bash scripts/gpu_configs.sh r5g 4
SECONDS=0; timeout 900 python bench.py --config 3 --seconds 8 --steps 2 --warmup 1 > gpurun_out/r5g_cfg3.json 2> gpurun_out/r5g_cfg3.err; echo "cfg3 rc=$? wall ${SECONDS}s"; tail -3 gpurun_out/r5g_cfg3.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r5g_cfg3.json").read().strip().split("\n")[-1])
    print("cfg3 value %.0f ms/step %.2f e2e %.0f parity %s plan %.1fs cpu %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_check"] and d["parity_check"]["ok"], d["plan_seconds"], d["cpu_baseline"] and round(d["cpu_baseline"]["value"],1)))
except Exception as e: print("cfg3 no line", e)
PY
