#!/bin/bash
# round-2 GPU call 35: text-bank glue kernels: tests, bench A/B not needed (single path), timeline tail
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py -q -k "glue or lstm or graphed or golden or whole_model or model or engine or stream or deferred" > gpurun_out/r2c35_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c35_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c35_bench.json 2> gpurun_out/r2c35_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2c35_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2c35_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
MGNNS_TIMELINE_ALL=1 timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c35_timeline_all.txt 2>&1; echo "timeline rc=$?"; sed -n 3,5p gpurun_out/r2c35_timeline_all.txt | cut -c1-110; grep "lstm_rec\|tc_gemm_kernel\|text_maxagg_bwd\|clip_adam\|pad_rows\|embedding_bwd" gpurun_out/r2c35_timeline_all.txt | tail -14
