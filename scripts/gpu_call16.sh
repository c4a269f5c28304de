#!/bin/bash
# round-2 GPU call 16 (2 GPUs): all-reduce microbench (CTA sweep vs NCCL), P2P test, 2-GPU bench
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 scripts/p2p_bench.py > gpurun_out/r2c16_p2p_bench.log 2>&1; echo "p2p bench rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2c16_p2p_bench.log | tail -14
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -k peer > gpurun_out/r2c16_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c16_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29813 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c16_n2.json 2> gpurun_out/r2c16_n2.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2c16_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'ar_ms', d['e2e'].get('allreduce_exposed_ms'))
"
