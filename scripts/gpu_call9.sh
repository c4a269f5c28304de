#!/bin/bash
# round-2 GPU call 9: flat optimizer tests, whole suite, N=1 bench (both optimizers), timeline, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "flat_clip or graphed" > gpurun_out/r2c9_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2c9_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --torch-optimizer > gpurun_out/r2c9_bench_torchopt.json 2> gpurun_out/r2c9_bench_torchopt.err; echo "bench torch-opt rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c9_bench_flat.json 2> gpurun_out/r2c9_bench_flat.err; echo "bench flat rc=$?"; tail -c 500 gpurun_out/r2c9_bench_flat.err
python -c "
import json
for f in ('torchopt','flat'):
    d=json.load(open('gpurun_out/r2c9_bench_%s.json'%f))
    print(f, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['config'].get('optimizer'))
"
timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c9_timeline.txt 2>&1; echo "timeline rc=$?"; head -40 gpurun_out/r2c9_timeline.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c9_all.log 2>&1; echo "all rc=$?"; tail -4 gpurun_out/r2c9_all.log
