#!/bin/bash
# round-2 GPU call 4: hub-staged SpMM correctness + cfg-2 timing + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "spmm or cfg2 or gcn_fused or graph_conv" > gpurun_out/r2c4_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c4_tests.log
tail -25 gpurun_out/r2c4_tests.log
timeout 600 python scripts/cfg2_bench.py 256 3 > gpurun_out/r2c4_cfg2.log 2>&1; cat gpurun_out/r2c4_cfg2.log
MGNNS_SPMM_HUB=0 timeout 600 python scripts/spmm_bench.py 256 > gpurun_out/r2c4_spmm_plain.log 2>&1; cat gpurun_out/r2c4_spmm_plain.log
timeout 600 python scripts/spmm_bench.py 256 > gpurun_out/r2c4_spmm_hub.log 2>&1; cat gpurun_out/r2c4_spmm_hub.log
BS=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_hub_kernel" -s 1 -c 1 -o gpurun_out/r2c4_ncu_spmm_hub python scripts/ncu_targets.py spmm > gpurun_out/r2c4_ncu.log 2>&1
echo "ncu rc=$?"
