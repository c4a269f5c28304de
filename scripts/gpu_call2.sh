#!/bin/bash
# round-2 GPU call 2: whole GPU suite (new PMI / engine / fused tests), N=1 bench line with the new legs, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c2_tests.log
tail -15 gpurun_out/r2c2_tests.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2c2_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2c2_bench.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches')}))
print(json.dumps(d['e2e']))
for k,v in d['roofline']['others'].items(): print(k, json.dumps({a:b for a,b in v.items() if a in ('frac','ms_per_launch','achieved','unit','error','docs_per_s','sample_bit_exact','share_of_step','gather_tbs')}))
print('dom', d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['ms_per_launch'])
print(json.dumps(d['config'].get('other_configs'), indent=1))
print(d.get('cpu_baseline'))
"
