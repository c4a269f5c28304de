#!/bin/bash
# usage: scripts/sweep_env.sh VAR v1 v2 ...   -> one short bench run per value, prints value / samples/s / ms per step
var=$1; shift
for v in "$@"; do
  env $var=$v MGNNS_BENCH_WATCHDOG_S=150 timeout 200 python bench.py --no-cpu-baseline --no-cfg2 --steps 20 2> gpurun_out/sweep_$v.err > gpurun_out/sweep_$v.json
  python - "$var" "$v" <<PY
import json, sys
try:
    d = json.load(open("gpurun_out/sweep_%s.json" % sys.argv[2]))
    print(sys.argv[1], sys.argv[2], round(d["value"], 1), round(d["ms_per_step"], 3))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed", e)
PY
done
