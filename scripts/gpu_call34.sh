#!/bin/bash
# round-2 GPU call 34 (N GPUs = $1): final cfg-4 multi-GPU number with the peer-memory all-reduce in the graph
mkdir -p gpurun_out
N=$1
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2c34_cfg4_n$N.json 2> gpurun_out/r2c34_cfg4_n$N.err
echo "cfg4 n$N rc=$?"; tail -c 300 gpurun_out/r2c34_cfg4_n$N.err
python -c "
import json
d=json.load(open('gpurun_out/r2c34_cfg4_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'h2d', d['e2e'].get('h2d_gbs_per_gpu'), 'ar_ms', d['e2e'].get('allreduce_exposed_ms'), d['clocks'])
"
