#!/bin/bash
# round-2 GPU call 43: final verification: whole GPU suite, smoke(), default bench run (all legs) and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c43_all.log 2>&1; echo "all rc=$?"; tail -3 gpurun_out/r2c43_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c43_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c43_smoke.log
t0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r2c43_bench_default.json 2> gpurun_out/r2c43_bench_default.err; echo "default bench rc=$? in $(( $(date +%s) - t0 )) s"; tail -c 200 gpurun_out/r2c43_bench_default.err
python -c "
import json
d=json.load(open('gpurun_out/r2c43_bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','steps')}, 'e2e', d['e2e']['value'])
r=d['roofline']; print('dominant', r['kernel'], round(r['frac'],3), r['ms_per_launch'])
for k,v in r['others'].items(): print(' ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('frac','ms_per_launch')})
print(d['cpu_baseline']['value'], d['clocks'])
"
