#!/bin/bash
# round-2 GPU call 26: fast gate math, reserved SMs for the image-bank kernels (sweep), tests, timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "lstm or graphed or golden or whole_model or model or stream or imgbank" > gpurun_out/r2c26_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2c26_tests.log
timeout 300 python scripts/lstm_phases.py 2>&1 | tail -2
for r in 8 0 4 16 24; do
  MGNNS_IMGBANK_RESERVE_SMS=$r timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c26_bench_$r.json 2> gpurun_out/r2c26_bench_$r.err; echo "bench reserve=$r rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r2c26_bench_$r.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
done
MGNNS_TIMELINE_ALL=1 timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c26_timeline_all.txt 2>&1; echo "timeline rc=$?"; sed -n 3,5p gpurun_out/r2c26_timeline_all.txt | cut -c1-110; grep "lstm_rec\|tc_gemm" gpurun_out/r2c26_timeline_all.txt | tail -9
