#!/bin/bash
# round-2 GPU call 6 (2 GPUs): 2-GPU DDP test, attention timing, cfg5 at N=1 and N=2, cfg4 at N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2c6_multi.log 2>&1; echo "multi rc=$?"; tail -5 gpurun_out/r2c6_multi.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "attn or attention or cfg5 or spmm" > gpurun_out/r2c6_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c6_tests.log
MGNNS_ATTN=tc timeout 300 python scripts/attn_bench.py 2>&1 | tail -2
timeout 900 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-extras > gpurun_out/r2c6_cfg5_n1.json 2> gpurun_out/r2c6_cfg5_n1.err; echo "cfg5 n1 rc=$?"; tail -c 600 gpurun_out/r2c6_cfg5_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r2c6_cfg5_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d.get('cpu_baseline'))
"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --workload cfg5 --steps 10 --warmup 3 > gpurun_out/r2c6_cfg5_n2.json 2> gpurun_out/r2c6_cfg5_n2.err; echo "cfg5 n2 rc=$?"; tail -c 400 gpurun_out/r2c6_cfg5_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c6_cfg4_n2.json 2> gpurun_out/r2c6_cfg4_n2.err; echo "cfg4 n2 rc=$?"; tail -c 400 gpurun_out/r2c6_cfg4_n2.err
python -c "
import json
for f in ('r2c6_cfg5_n2','r2c6_cfg4_n2'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['e2e'].get('h2d_ceiling_gbs_per_gpu'), d['e2e'].get('h2d_gbs_per_gpu'), 'ar_ms', d['e2e'].get('allreduce_exposed_ms'), d['config'].get('numa'))
"
