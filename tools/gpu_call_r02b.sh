#!/bin/bash
# Round-2 consolidated evidence run (one gpurun call): GPU tests, K-build probe of every config kernel with the
# tensor-path kernel (+ variants), ncu of the K-build cases and of the gradient trace, driver-style bench at N=1.
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -x -q ) > $O/r02b_pytest_gpu.log 2>&1
python tools/kbuild_probe.py 32768 > $O/r02b_kbuild_v2.jsonl 2> $O/r02b_kbuild_v2.err
for v in 0 1 3; do
  GPL_GRAM_VARIANT=$v python tools/kbuild_probe.py 32768 "matern52 d8,sexp*lin,nn" >> $O/r02b_kbuild_variants.jsonl 2>> $O/r02b_kbuild_v2.err
done
python bench.py --steps 5 --warmup 3 > $O/r02b_bench_n1.json 2> $O/r02b_bench_n1.err
GPL_PROBE_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel \
  -o $O/r02b_ncu_kbuild_cases -f python tools/kbuild_probe.py 32768 > $O/r02b_ncu_kbuild.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grad_trace \
  -o $O/r02b_ncu_gradtrace -f python tools/prof_targets.py egrad 32768 > $O/r02b_ncu_gradtrace.log 2>&1
GPL_TIMING=1 python tools/prof_targets.py egrad 32768 > $O/r02b_phase_timing_egrad32768.log 2>&1
ls -la $O
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file $O/r02b_launches_egrad32768.csv python tools/prof_targets.py egrad 32768 > $O/r02b_launches.log 2>&1
