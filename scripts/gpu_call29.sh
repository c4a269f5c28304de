#!/bin/bash
# round-2 GPU call 29: stream priorities for the critical chain (A/B), tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -q -k "graphed or golden or whole_model or model or stream or engine" > gpurun_out/r2c29_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2c29_tests.log
for pr in 1; do
  MGNNS_STREAM_PRIORITIES=$pr timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c29_bench_$pr.json 2> gpurun_out/r2c29_bench_$pr.err; echo "bench priorities=$pr rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r2c29_bench_$pr.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
done
MGNNS_TIMELINE_ALL=1 timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c29_timeline_all.txt 2>&1; echo "timeline rc=$?"; sed -n 3,5p gpurun_out/r2c29_timeline_all.txt | cut -c1-110; grep "lstm_rec\|tc_gemm" gpurun_out/r2c29_timeline_all.txt | tail -9
