#!/bin/bash
# ncu launch list (durations only, one pass per launch) of ONE energy + gradient evaluation at n = 32768
mkdir -p gpurun_out
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02g_launches_egrad32768.csv \
  python tools/prof_targets.py egrad 32768 > gpurun_out/r02g_launches.log 2>&1
ls -la gpurun_out | tail -3
