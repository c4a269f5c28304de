#!/bin/bash
# round-2 GPU call 11: whole suite (no -x), gated heavy deferral A/B, timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c11_all.log 2>&1; echo "all rc=$?"; tail -8 gpurun_out/r2c11_all.log
for cfg in "1 1" "0 0" "1 0" "0 1"; do
  set -- $cfg
  MGNNS_GATE_HEAVY=$1 MGNNS_PLACE_AFTER_LAST_LAYER=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c11_bench_$1$2.json 2> gpurun_out/r2c11_bench_$1$2.err; echo "bench gate=$1 place=$2 rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r2c11_bench_$1$2.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
done
timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c11_timeline.txt 2>&1; echo "timeline rc=$?"; sed -n 3,12p gpurun_out/r2c11_timeline.txt | cut -c1-110; grep "lstm_rec\|tc_gemm" gpurun_out/r2c11_timeline.txt | tail -12
