#!/bin/bash
# round-2 GPU call 3: tensor-core attention correctness + timing, PMI phase timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "attn or attention or pmi" > gpurun_out/r2c3_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c3_tests.log
tail -25 gpurun_out/r2c3_tests.log
for impl in tc scalar; do
  MGNNS_ATTN=$impl timeout 300 python scripts/attn_bench.py > gpurun_out/r2c3_attn_$impl.log 2>&1
  echo "== $impl"; cat gpurun_out/r2c3_attn_$impl.log | tail -8
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c3_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r2c3_all.log
tail -5 gpurun_out/r2c3_all.log
B=512 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_q1_tc" -s 2 -c 4 -o gpurun_out/r2c3_ncu_attn python scripts/ncu_targets.py attn > gpurun_out/r2c3_ncu.log 2>&1
echo "ncu rc=$?"
timeout 300 python scripts/pmi_phases.py > gpurun_out/r2c3_pmi.log 2>&1; cat gpurun_out/r2c3_pmi.log | tail -12
