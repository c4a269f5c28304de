#!/bin/bash
# round-2 GPU call 15 (2 GPUs): peer-memory all-reduce tests, 2-GPU bench with P2P vs NCCL
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2c15_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2c15_tests.log
run() { # name nproc p2p
  MGNNS_P2P_ALLREDUCE=$3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) bench.py --gpus $2 --steps 20 --warmup 5 > gpurun_out/r2c15_$1.json 2> gpurun_out/r2c15_$1.err
  echo "$1 rc=$?"; tail -c 300 gpurun_out/r2c15_$1.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c15_$1.json'))
print('$1', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'ar_ms', d['e2e'].get('allreduce_exposed_ms'), d['config'].get('allreduce'))
"
}
run n2_p2p 2 1
run n2_nccl 2 0
