#!/bin/bash
# round-2 GPU call 36: row threshold from which a product goes to the tcgen05 kernels (M = 512 projections on tensor cores?)
mkdir -p gpurun_out
for v in 512 1024; do
  MGNNS_TC_MIN_ROWS=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c36_bench_$v.json 2> gpurun_out/r2c36_bench_$v.err; echo "bench tc_min_rows=$v rc=$?"; tail -c 200 gpurun_out/r2c36_bench_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c36_bench_$v.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
done
