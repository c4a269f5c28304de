#!/bin/bash
# round-2 GPU call 37: cfg 5 at one GPU with the final code; smoke(); default bench timing
mkdir -p gpurun_out
timeout 900 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c37_cfg5_n1.json 2> gpurun_out/r2c37_cfg5_n1.err; echo "cfg5 rc=$?"; tail -c 200 gpurun_out/r2c37_cfg5_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r2c37_cfg5_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c37_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2c37_smoke.log
