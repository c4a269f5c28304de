#!/bin/bash
# round-2 GPU call 41: persistent LSTM clusters + delayed release of the image-bank kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "graphed or stream or whole_model" > gpurun_out/r2c41_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c41_tests.log
for v in 20 37; do
  MGNNS_LSTM_CLUSTERS=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c41_bench_$v.json 2> gpurun_out/r2c41_bench_$v.err; echo "bench lstm_clusters=$v rc=$?"; tail -c 200 gpurun_out/r2c41_bench_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c41_bench_$v.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
done
MGNNS_TIMELINE_ALL=1 timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c41_timeline_all.txt 2>&1; echo "timeline rc=$?"; sed -n 3,5p gpurun_out/r2c41_timeline_all.txt | cut -c1-110; grep "lstm_rec\|tc_gemm_kernel\|text_maxagg_bwd\|clip_adam" gpurun_out/r2c41_timeline_all.txt | tail -12
