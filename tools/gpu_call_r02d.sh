#!/bin/bash
# Short diagnostic call: which K-build path restores the 1e-8 FITC parity on the ill-conditioned case; correctness
# of the opt-in variants (one-term K build, 128x64 two-CTA GEMM tile) and their timings.
mkdir -p gpurun_out
O=gpurun_out
python tools/diag_fitc_parity.py > $O/r02d_fitc_diag.jsonl 2> $O/r02d_fitc_diag.err
GPL_FITC_STREAM_GB=0 GPL_FITC_CHUNK=512 python tools/diag_fitc_parity.py >> $O/r02d_fitc_diag.jsonl 2>> $O/r02d_fitc_diag.err
GPL_KBUILD_V1=1 python tools/diag_fitc_parity.py >> $O/r02d_fitc_diag.jsonl 2>> $O/r02d_fitc_diag.err
GPL_KBUILD_V1=1 GPL_FITC_STREAM_GB=0 GPL_FITC_CHUNK=512 python tools/diag_fitc_parity.py >> $O/r02d_fitc_diag.jsonl 2>> $O/r02d_fitc_diag.err
GPL_KBUILD_V1=1 GPL_KBUILD_DIRECT=1 python tools/diag_fitc_parity.py >> $O/r02d_fitc_diag.jsonl 2>> $O/r02d_fitc_diag.err
GPL_FITC_DIRECT_K=1 python tools/diag_fitc_parity.py >> $O/r02d_fitc_diag.jsonl 2>> $O/r02d_fitc_diag.err
GPL_FITC_DIRECT_K=1 GPL_FITC_STREAM_GB=0 GPL_FITC_CHUNK=512 python tools/diag_fitc_parity.py >> $O/r02d_fitc_diag.jsonl 2>> $O/r02d_fitc_diag.err
GPL_GRAM_ONE_TERM=1 python -m pytest tests/test_gpu_cf.py -q > $O/r02d_pytest_one_term.log 2>&1
GPL_GEMM_L2X=1 python -m pytest tests/test_gpu_linalg.py tests/test_gpu_fit.py -q -x > $O/r02d_pytest_l2x.log 2>&1
GPL_GEMM_L2X=1 GPL_TIMING=1 python tools/prof_targets.py egrad 32768 > $O/r02d_phase_timing_l2x.log 2>&1
GPL_TIMING=1 python tools/prof_targets.py egrad 32768 > $O/r02d_phase_timing_default.log 2>&1
du -sm $O
