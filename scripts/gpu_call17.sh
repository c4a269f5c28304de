#!/bin/bash
# round-2 GPU call 17 (8 GPUs): all-reduce microbench, cfg 4 and cfg 5 at 8 GPUs with the peer-memory all-reduce in the graph
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29821 scripts/p2p_bench.py > gpurun_out/r2c17_p2p_bench.log 2>&1; echo "p2p bench rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2c17_p2p_bench.log | tail -14
run() { # name nproc workload p2p
  MGNNS_P2P_ALLREDUCE=$4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $2 --workload $3 --steps 20 --warmup 5 > gpurun_out/r2c17_$1.json 2> gpurun_out/r2c17_$1.err
  echo "$1 rc=$?"; tail -c 300 gpurun_out/r2c17_$1.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c17_$1.json'))
print('$1', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'h2d', d['e2e'].get('h2d_gbs_per_gpu'), 'ar_ms', d['e2e'].get('allreduce_exposed_ms'), d['clocks'])
"
}
run cfg4_n8 8 cfg4 1
run cfg4_n8_nccl 8 cfg4 0
run cfg5_n8 8 cfg5 1
