#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29873 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c45_n2.json 2> gpurun_out/r2c45_n2.err; echo "bench rc=$?"; tail -c 200 gpurun_out/r2c45_n2.err
python -c "
import json
d=json.load(open('gpurun_out/r2c45_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'ar_ms', d['e2e'].get('allreduce_exposed_ms'))
"
