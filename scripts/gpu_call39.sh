#!/bin/bash
# round-2 GPU call 39: top-layer LSTM backward held until the image banks' gradients are ready (A/B)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "graphed or deferred or defer" > gpurun_out/r2c39_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c39_tests.log
for v in 1 0; do
  MGNNS_LSTM_BWD_AFTER_IMAGE=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c39_bench_$v.json 2> gpurun_out/r2c39_bench_$v.err; echo "bench lstm_bwd_after_image=$v rc=$?"; tail -c 200 gpurun_out/r2c39_bench_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c39_bench_$v.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
done
MGNNS_TIMELINE_ALL=1 timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c39_timeline_all.txt 2>&1; echo "timeline rc=$?"; sed -n 3,5p gpurun_out/r2c39_timeline_all.txt | cut -c1-110; grep "lstm_rec_bwd\|tc_gemm_kernel<1\|text_maxagg_bwd\|clip_adam\|attn_q1_tc_bwd" gpurun_out/r2c39_timeline_all.txt | tail -16
