#!/bin/bash
# Round-2 evidence run, lean version (gpurun_out must stay below 64 MiB: ncu reports are reduced to CSV on the
# box and deleted).  GPU tests, K-build probe of every config kernel (+ tile / occupancy variants), gradient-trace
# occupancy variants, bench at N=1, ncu --set full of the K-build cases and the gradient trace.
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -q ) > $O/r02c_pytest_gpu.log 2>&1
python tools/kbuild_probe.py 32768 > $O/r02c_kbuild_v2.jsonl 2> $O/r02c_kbuild.err
for v in 0 1 3; do
  GPL_GRAM_VARIANT=$v python tools/kbuild_probe.py 32768 "matern52 d8,sexp*lin,nn,sexp d6" >> $O/r02c_kbuild_variants.jsonl 2>> $O/r02c_kbuild.err
done
for v in 4 0 1 3; do
  GPL_GRAM_ONE_TERM=1 GPL_GRAM_VARIANT=$v python tools/kbuild_probe.py 32768 "sexp*lin" >> $O/r02c_kbuild_one_term.jsonl 2>> $O/r02c_kbuild.err
done
for m in 2 3 4; do
  echo "GPL_GRAD_MINB=$m" >> $O/r02c_phase_timing_egrad32768.log
  GPL_GRAD_MINB=$m GPL_TIMING=1 python tools/prof_targets.py egrad 32768 >> $O/r02c_phase_timing_egrad32768.log 2>&1
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r02c_bench_n1.json 2> $O/r02c_bench_n1.err
GPL_PROBE_REPS=1 timeout 300 ncu --set full --clock-control none -k regex:gram_kernel \
  -o $O/r02c_ncu_kbuild_cases -f python tools/kbuild_probe.py 32768 "matern52 d8,sexp*lin" > $O/r02c_ncu_kbuild.log 2>&1
python tools/ncu_select.py $O/r02c_ncu_kbuild_cases.ncu-rep $O/r02c_ncu_kbuild_cases_selected.csv >> $O/r02c_ncu_kbuild.log 2>&1
rm -f $O/r02c_ncu_kbuild_cases.ncu-rep
timeout 300 ncu --set full --clock-control none -k regex:grad_trace \
  -o $O/r02c_ncu_gradtrace -f python tools/prof_targets.py egrad 32768 > $O/r02c_ncu_gradtrace.log 2>&1
python tools/ncu_select.py $O/r02c_ncu_gradtrace.ncu-rep $O/r02c_ncu_gradtrace_selected.csv >> $O/r02c_ncu_gradtrace.log 2>&1
ncu -i $O/r02c_ncu_gradtrace.ncu-rep --page details --csv > $O/r02c_ncu_gradtrace_details.csv 2>> $O/r02c_ncu_gradtrace.log
rm -f $O/r02c_ncu_gradtrace.ncu-rep
du -sm $O; ls -la $O
