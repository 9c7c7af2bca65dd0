#!/bin/bash
# Final validation + measurement call of round 2 (about 7 GPU-minutes were left): GPU suite, ncu of the n = 50000
# GEMV (config 3's f = K alpha) and of the EP site sweep, then the driver-style bench at N = 1.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 120 python -m pytest tests -m gpu -q ) > $O/r02f_pytest_gpu.log 2>&1
timeout 60 ncu --set full --clock-control none -k regex:gemv -c 4 -o $O/r02f_ncu_gemv50000 -f \
  python tools/prof_targets.py gemv 50000 > $O/r02f_ncu_gemv.log 2>&1
python tools/ncu_select.py $O/r02f_ncu_gemv50000.ncu-rep $O/r02f_ncu_gemv50000_selected.csv >> $O/r02f_ncu_gemv.log 2>&1
rm -f $O/r02f_ncu_gemv50000.ncu-rep
timeout 90 ncu --set full --clock-control none -k regex:ep_sites -c 2 -o $O/r02f_ncu_ep_sites50000 -f \
  python tools/prof_targets.py sweeps 50000 > $O/r02f_ncu_ep_sites.log 2>&1
python tools/ncu_select.py $O/r02f_ncu_ep_sites50000.ncu-rep $O/r02f_ncu_ep_sites50000_selected.csv >> $O/r02f_ncu_ep_sites.log 2>&1
rm -f $O/r02f_ncu_ep_sites50000.ncu-rep
timeout 240 python bench.py --steps 5 --warmup 3 > $O/r02f_bench_n1.json 2> $O/r02f_bench_n1.err
du -sm $O
