#!/bin/bash
# round-2 GPU call 30: which chains get the high-priority streams (0 LSTM, 1/2 image channels + image-query stacks, 3/4 label channels, 5/6 text-bank stacks)
mkdir -p gpurun_out
i=0
for np in "1,2" "1,2,3,4,5,6" "3,4,5,6" "5,6" ""; do
  i=$((i+1))
  MGNNS_NORMAL_PRIORITY_STREAMS="$np" timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c30_bench_$i.json 2> gpurun_out/r2c30_bench_$i.err; echo "bench normal='$np' rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r2c30_bench_$i.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
"
done
