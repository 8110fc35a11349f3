#!/bin/bash
# round 2, call 19: two-phase (hi first, lo later) operand hand-over: parity, trace, timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_gpu_round2.py tests/test_gpu_fp16_split.py tests/test_gpu_adaptive.py -q -x > gpurun_out/r02_tests_c19.txt 2>&1
tail -5 gpurun_out/r02_tests_c19.txt
timeout 300 python scripts/trace_tc.py 4 0,36,1028 > gpurun_out/r02_trace_tc7.txt 2>&1
cat gpurun_out/r02_trace_tc7.txt
TCDE_REPS=4 timeout 300 python scripts/adjoint_bench.py 65536 0 > gpurun_out/r02_adjoint_bench8.txt 2>&1
cat gpurun_out/r02_adjoint_bench8.txt
