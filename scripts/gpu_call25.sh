#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/lstm_phases.py > gpurun_out/r2c25_phases.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2c25_phases.log
