#!/bin/bash
# round-2 GPU call 24: FFMA2 in the LSTM recurrence and the CUDA-core GEMM: whole suite, bench, timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c24_all.log 2>&1; echo "all rc=$?"; tail -5 gpurun_out/r2c24_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c24_bench.json 2> gpurun_out/r2c24_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2c24_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2c24_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
for k,v in d['roofline']['others'].items():
    if 'lstm' in k: print(k, {a:b for a,b in v.items() if a in ('ms_per_launch','share_of_step')})
"
MGNNS_TIMELINE_ALL=1 timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c24_timeline_all.txt 2>&1; echo "timeline rc=$?"; sed -n 3,5p gpurun_out/r2c24_timeline_all.txt | cut -c1-110; grep "lstm_rec" gpurun_out/r2c24_timeline_all.txt | tail -5; sed -n 6,30p gpurun_out/r2c24_timeline_all.txt | cut -c1-100
