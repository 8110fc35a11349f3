#!/bin/bash
# round 2, call 13: consumer-side proxy fence (solve + param-grad), 16 producer warps: parity, timing, trace
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_gpu_adaptive.py tests/test_gpu_round2.py tests/test_gpu_fp16_split.py -q > gpurun_out/r02_tests_c13.txt 2>&1
tail -6 gpurun_out/r02_tests_c13.txt
timeout 300 python scripts/trace_tc.py 4 0,36,1028 > gpurun_out/r02_trace_tc6.txt 2>&1
cat gpurun_out/r02_trace_tc6.txt
TCDE_VERBOSE=1 TCDE_REPS=5 timeout 300 python scripts/adjoint_bench.py 65536 0 > gpurun_out/r02_adjoint_bench5.txt 2>&1
cat gpurun_out/r02_adjoint_bench5.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train3.csv python scripts/adjoint_bench.py 65536 0 > gpurun_out/r02_train_under_ncu3.txt 2>&1
grep -E "cdeint_tc_kernel|param_grad" gpurun_out/r02_launches_train3.csv | tail -4 | rev | cut -c1-40 | rev
