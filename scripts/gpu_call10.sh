#!/bin/bash
# round-2 GPU call 10: LSTM barrier split / prefetch, optimizer test, whole suite, N=1 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lstm or flat_clip or graphed or full_model" > gpurun_out/r2c10_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2c10_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2c10_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2c10_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
for k,v in d['roofline']['others'].items(): print(k, {a:b for a,b in v.items() if a in ('frac','ms_per_launch','share_of_step')})
print('dom', d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['ms_per_launch'])
"
timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c10_timeline.txt 2>&1; echo "timeline rc=$?"; sed -n 3,12p gpurun_out/r2c10_timeline.txt | cut -c1-110; grep "lstm_rec" gpurun_out/r2c10_timeline.txt | tail -6
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c10_all.log 2>&1; echo "all rc=$?"; tail -4 gpurun_out/r2c10_all.log
