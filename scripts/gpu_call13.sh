#!/bin/bash
# round-2 GPU call 13: tests, deferral of the small weight gradients A/B, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "flat_clip or graphed or deferred or defer or stream" > gpurun_out/r2c13_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c13_tests.log
for c in 1 0; do
  MGNNS_DEFER_SMALL=$c timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c13_bench_$c.json 2> gpurun_out/r2c13_bench_$c.err; echo "bench defer_small=$c rc=$?"; tail -c 300 gpurun_out/r2c13_bench_$c.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c13_bench_$c.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
done
timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c13_timeline.txt 2>&1; echo "timeline rc=$?"; sed -n 3,12p gpurun_out/r2c13_timeline.txt | cut -c1-110; grep "lstm_rec\|tc_gemm" gpurun_out/r2c13_timeline.txt | tail -8
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c13_all.log 2>&1; echo "all rc=$?"; tail -5 gpurun_out/r2c13_all.log
