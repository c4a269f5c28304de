#!/bin/bash
# round-2 GPU call 5: packed tensor-core attention (H<=4), padded-row hub SpMM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "attn or attention or spmm or cfg2" > gpurun_out/r2c5_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c5_tests.log
tail -25 gpurun_out/r2c5_tests.log
MGNNS_ATTN=tc timeout 300 python scripts/attn_bench.py > gpurun_out/r2c5_attn_tc.log 2>&1; tail -3 gpurun_out/r2c5_attn_tc.log
for pad in 0 1; do
  for hub in 0 1; do
    echo "== pad=$pad hub=$hub"
    MGNNS_SPMM_PAD=$pad MGNNS_SPMM_HUB=$hub timeout 600 python scripts/spmm_bench.py 256 2>&1 | tail -2
  done
done
timeout 600 python scripts/cfg2_bench.py 256 3 2>&1 | head -3
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c5_all.log 2>&1
echo "all rc=$?" >> gpurun_out/r2c5_all.log
tail -5 gpurun_out/r2c5_all.log
