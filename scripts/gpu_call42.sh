#!/bin/bash
# round-2 GPU call 42 (2 GPUs): multi-GPU tests and a 2-GPU bench with the final code
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2c42_multi.log 2>&1; echo "multi tests rc=$?"; tail -4 gpurun_out/r2c42_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29871 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c42_n2.json 2> gpurun_out/r2c42_n2.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2c42_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'ar_ms', d['e2e'].get('allreduce_exposed_ms'))
"
