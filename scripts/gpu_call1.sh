#!/bin/bash
# round-2 GPU call 1: fused GCN correctness, cfg-2 timings, whole GPU suite, ncu of the fused kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "gcn_fused or cfg2_bench_shape" > gpurun_out/r2c1_fused_tests.log 2>&1
echo "fused tests rc=$?" >> gpurun_out/r2c1_fused_tests.log
tail -5 gpurun_out/r2c1_fused_tests.log
timeout 600 python scripts/cfg2_bench.py 256 3 > gpurun_out/r2c1_cfg2.log 2>&1
echo "cfg2 rc=$?" >> gpurun_out/r2c1_cfg2.log
cat gpurun_out/r2c1_cfg2.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_all_tests.log 2>&1
echo "all tests rc=$?" >> gpurun_out/r2c1_all_tests.log
tail -5 gpurun_out/r2c1_all_tests.log
BS=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gcn_fused" -s 1 -c 1 -o gpurun_out/r2c1_ncu_fused python scripts/ncu_targets.py gcn > gpurun_out/r2c1_ncu.log 2>&1
echo "ncu rc=$?"
