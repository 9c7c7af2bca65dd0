#!/bin/bash
# Validation call: the full GPU suite with the committed defaults, then the opt-in 128x64 two-CTA GEMM tile
# (lower-triangular mode fixed for non-square tiles) on the linear-algebra / fit suites plus its phase timings.
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -q ) > $O/r02e_pytest_gpu.log 2>&1
GPL_GEMM_L2X=1 python -m pytest tests/test_gpu_linalg.py tests/test_gpu_fit.py -q > $O/r02e_pytest_l2x.log 2>&1
for i in 1 2; do
GPL_GEMM_L2X=1 GPL_TIMING=1 python tools/prof_targets.py egrad 32768 2>&1 | grep timing >> $O/r02e_phase_timing_l2x.log
GPL_TIMING=1 python tools/prof_targets.py egrad 32768 2>&1 | grep timing >> $O/r02e_phase_timing_default.log
done
du -sm $O
