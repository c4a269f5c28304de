#!/bin/bash
# round-2 GPU call 12: optimizer / graph-step tests, heavy-job scheduling sweep (gated vs early with a capped grid), timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "flat_clip or graphed or deferred or defer" > gpurun_out/r2c12_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c12_tests.log
for c in 0 64 80 96 112; do
  MGNNS_HEAVY_CTAS=$c timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c12_bench_$c.json 2> gpurun_out/r2c12_bench_$c.err; echo "bench heavy_ctas=$c rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r2c12_bench_$c.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
done
MGNNS_HEAVY_CTAS=96 timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c12_timeline96.txt 2>&1; echo "timeline rc=$?"; sed -n 3,12p gpurun_out/r2c12_timeline96.txt | cut -c1-110; grep "lstm_rec\|tc_gemm" gpurun_out/r2c12_timeline96.txt | tail -8
