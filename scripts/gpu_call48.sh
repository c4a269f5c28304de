#!/bin/bash
# round-2 GPU call 48 (2 GPUs): forced peer-memory failure on rank 1 -> every rank falls back to NCCL; normal path still green
mkdir -p gpurun_out
MGNNS_P2P_FAIL_RANK=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29875 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c48_fallback.json 2> gpurun_out/r2c48_fallback.err; echo "fallback bench rc=$?"; grep -i "unavailable" gpurun_out/r2c48_fallback.err | head -2 | cut -c1-200
python -c "
import json
d=json.load(open('gpurun_out/r2c48_fallback.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config'].get('allreduce'), d['config'].get('launch'), d['config'].get('launch_note'))
"
timeout 200 python -m pytest tests/test_gpu_multi.py -q -m gpu -k peer 2>&1 | tail -2
