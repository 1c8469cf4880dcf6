#!/bin/bash
# bench.py on every BASELINE config (1 GPU), JSON lines into gpurun_out/<tag>_cfgN.json.  usage: scripts/gpu_configs.sh <tag> [configs...]
set -u
tag=$1; shift
cfgs=${*:-"2 0 1 3 4"}
mkdir -p gpurun_out
for c in $cfgs; do
  extra=""
  [ "$c" = "4" ] && extra="--seconds ${CFG4_SECONDS:-600}"
  SECONDS=0; timeout 900 python bench.py --config $c $extra > gpurun_out/${tag}_cfg$c.json 2> gpurun_out/${tag}_cfg$c.err; echo "cfg$c rc=$? wall ${SECONDS}s"
  tail -2 gpurun_out/${tag}_cfg$c.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_cfg$c.json").read().strip().split("\n")[-1])
    print("cfg$c value %.0f  ms/step %.2f  e2e %.0f  e2e(out on dev) %.0f  parity %s  cpu %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["outputs_left_on_device"]["value"], d["parity_check"] and d["parity_check"]["ok"], d["cpu_baseline"] and (round(d["cpu_baseline"]["value"],1), round(d["cpu_baseline"]["per_core_x_realtime"],1), d["cpu_baseline"]["cores"])))
except Exception as e:
    print("cfg$c: no line", e)
PY
done
