#!/bin/bash
# round-2 GPU call 22: ncu full-set capture of the LSTM cluster kernels (source-level stalls), per-launch time list
mkdir -p gpurun_out
timeout 300 python scripts/ncu_targets.py lstm > gpurun_out/r2c22_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/r2c22_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_rec -s 4 -c 4 -f -o gpurun_out/r2c22_ncu_lstm python scripts/ncu_targets.py lstm > gpurun_out/r2c22_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2c22_ncu.log
ls -la gpurun_out/r2c22_ncu_lstm.ncu-rep
