#!/bin/bash
# round-2 GPU call 14: label channels on their own streams; heavy-job plans; timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -q -k "flat_clip or graphed or deferred or defer or stream or engine or full_model" > gpurun_out/r2c14_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c14_tests.log
for plan in g0,g1 c100,g0 c110,c110 g0,g0 c100g0,g1; do
  MGNNS_HEAVY_PLAN=$plan timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c14_bench_$plan.json 2> gpurun_out/r2c14_bench_$plan.err; echo "bench plan=$plan rc=$?"; tail -c 200 gpurun_out/r2c14_bench_$plan.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c14_bench_$plan.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
done
timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c14_timeline.txt 2>&1; echo "timeline rc=$?"; sed -n 3,12p gpurun_out/r2c14_timeline.txt | cut -c1-110; grep "lstm_rec\|tc_gemm" gpurun_out/r2c14_timeline.txt | tail -8
