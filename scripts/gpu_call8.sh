#!/bin/bash
# round-2 GPU call 8 (8 GPUs): cfg 5 at size (B=4096 over 8 GPUs), cfg 4 at 8 and 4 GPUs
mkdir -p gpurun_out
nvidia-smi -L | wc -l
nvidia-smi topo -m > gpurun_out/r2c8_topo.txt 2>&1
run() { # name nproc workload
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $2 --workload $3 --steps 20 --warmup 5 > gpurun_out/r2c8_$1.json 2> gpurun_out/r2c8_$1.err
  echo "$1 rc=$?"; tail -c 300 gpurun_out/r2c8_$1.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c8_$1.json'))
print('$1', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'h2d', d['e2e'].get('h2d_gbs_per_gpu'), 'ceil', d['e2e'].get('h2d_ceiling_gbs_per_gpu'), 'ar_ms', d['e2e'].get('allreduce_exposed_ms'), d['config'].get('numa'), d['clocks'])
"
}
run cfg5_n8 8 cfg5
run cfg4_n8 8 cfg4
run cfg4_n4 4 cfg4
