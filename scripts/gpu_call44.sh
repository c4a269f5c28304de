#!/bin/bash
# round-2 GPU call 44: batched accumulation of the deferred gradients: tests + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "graphed or deferred or defer or flat_clip or stream" > gpurun_out/r2c44_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c44_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c44_bench.json 2> gpurun_out/r2c44_bench.err; echo "bench rc=$?"; tail -c 200 gpurun_out/r2c44_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2c44_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
