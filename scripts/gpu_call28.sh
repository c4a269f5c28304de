#!/bin/bash
# round-2 GPU call 28: hardware queue count (CUDA_DEVICE_MAX_CONNECTIONS) vs the 18-branch step graph
mkdir -p gpurun_out
for c in 32 8 16; do
  CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c28_bench_$c.json 2> gpurun_out/r2c28_bench_$c.err; echo "bench connections=$c rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r2c28_bench_$c.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
done
CUDA_DEVICE_MAX_CONNECTIONS=32 MGNNS_TIMELINE_ALL=1 timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c28_timeline_all.txt 2>&1; echo "timeline rc=$?"; sed -n 3,5p gpurun_out/r2c28_timeline_all.txt | cut -c1-110; grep "lstm_rec\|tc_gemm" gpurun_out/r2c28_timeline_all.txt | tail -9
