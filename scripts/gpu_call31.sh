#!/bin/bash
# round-2 GPU call 31: default bench run (all legs), reference arm, launch list, ncu full of the judged kernels, whole suite
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r2c31_bench_default.json 2> gpurun_out/r2c31_bench_default.err; echo "default bench rc=$? in $(( $(date +%s) - t0 )) s"; tail -c 300 gpurun_out/r2c31_bench_default.err
t0=$(date +%s)
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c31_bench_reference.json 2> gpurun_out/r2c31_bench_reference.err; echo "reference arm rc=$? in $(( $(date +%s) - t0 )) s"; cat gpurun_out/r2c31_bench_reference.json | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2c31_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --eager --no-branch-streams > gpurun_out/r2c31_launches.log 2>&1; echo "launch list rc=$?"
python scripts/summarize_launches.py gpurun_out/r2c31_launches.csv --top 30 > gpurun_out/r2c31_launches.md 2>&1; head -24 gpurun_out/r2c31_launches.md
B=512 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel" -s 2 -c 2 -f -o gpurun_out/r2c31_ncu_imgbank python scripts/ncu_targets.py imgbank > gpurun_out/r2c31_ncu1.log 2>&1; echo "ncu imgbank rc=$?"
B=512 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_q1_tc" -s 2 -c 4 -f -o gpurun_out/r2c31_ncu_attn python scripts/ncu_targets.py attn > gpurun_out/r2c31_ncu2.log 2>&1; echo "ncu attn rc=$?"
BS=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_csr" -s 1 -c 1 -f -o gpurun_out/r2c31_ncu_spmm python scripts/ncu_targets.py spmm > gpurun_out/r2c31_ncu3.log 2>&1; echo "ncu spmm rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c31_all.log 2>&1; echo "all rc=$?"; tail -3 gpurun_out/r2c31_all.log
