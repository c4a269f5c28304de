#!/bin/bash
# round-2 GPU call 40: persistent LSTM clusters with a tile queue: tests, sanitizer, bench (cluster-count sweep), timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "lstm or graphed or golden or whole_model or model or glue" > gpurun_out/r2c40_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2c40_tests.log
timeout 300 python scripts/lstm_phases.py 2>&1 | tail -2
for v in 20 16 28 37; do
  MGNNS_LSTM_CLUSTERS=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c40_bench_$v.json 2> gpurun_out/r2c40_bench_$v.err; echo "bench lstm_clusters=$v rc=$?"; tail -c 200 gpurun_out/r2c40_bench_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c40_bench_$v.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], [ (k, round(v['ms_per_launch'],3)) for k,v in d['roofline']['others'].items() if 'lstm' in k])
"
done
MGNNS_TIMELINE_ALL=1 timeout 600 python scripts/timeline_step.py 512 > gpurun_out/r2c40_timeline_all.txt 2>&1; echo "timeline rc=$?"; sed -n 3,5p gpurun_out/r2c40_timeline_all.txt | cut -c1-110; grep "lstm_rec\|tc_gemm_kernel\|text_maxagg_bwd\|clip_adam" gpurun_out/r2c40_timeline_all.txt | tail -12
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_targets.py > gpurun_out/r2c40_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r2c40_racecheck.log
