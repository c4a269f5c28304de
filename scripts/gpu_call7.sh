#!/bin/bash
# round-2 GPU call 7 (2 GPUs): 2-GPU DDP test, cfg5 at N=1 and N=2, step timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2c7_multi.log 2>&1; echo "multi rc=$?"; tail -3 gpurun_out/r2c7_multi.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "cfg5 or spmm" > gpurun_out/r2c7_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c7_tests.log
timeout 900 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-extras > gpurun_out/r2c7_cfg5_n1.json 2> gpurun_out/r2c7_cfg5_n1.err; echo "cfg5 n1 rc=$?"; tail -c 600 gpurun_out/r2c7_cfg5_n1.err
python -c "
import json
d=json.load(open('gpurun_out/r2c7_cfg5_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d.get('cpu_baseline'))
for k,v in d['roofline']['others'].items(): print(k, {a:b for a,b in v.items() if a in ('frac','ms_per_launch','share_of_step','launches_timed')})
"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --workload cfg5 --steps 10 --warmup 3 > gpurun_out/r2c7_cfg5_n2.json 2> gpurun_out/r2c7_cfg5_n2.err; echo "cfg5 n2 rc=$?"; tail -c 400 gpurun_out/r2c7_cfg5_n2.err
python -c "
import json
d=json.load(open('gpurun_out/r2c7_cfg5_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'ar_ms', d['e2e'].get('allreduce_exposed_ms'))
"
timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c7_timeline.txt 2>&1; echo "timeline rc=$?"; head -45 gpurun_out/r2c7_timeline.txt
