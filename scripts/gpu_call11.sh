#!/bin/bash
# round 2, call 11: staged stage dump + 8-warp param-grad kernel: tests, training-step launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_adaptive.py tests/test_gpu_solve.py tests/test_gpu_builders.py -q > gpurun_out/r02_tests_c11.txt 2>&1
tail -8 gpurun_out/r02_tests_c11.txt
timeout 300 python scripts/adjoint_bench.py 65536 0 > gpurun_out/r02_adjoint_bench3.txt 2>&1
cat gpurun_out/r02_adjoint_bench3.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train2.csv python scripts/adjoint_bench.py 65536 0 > gpurun_out/r02_train_under_ncu2.txt 2>&1
grep -E "cdeint_tc_kernel|param_grad" gpurun_out/r02_launches_train2.csv | tail -6 | cut -c1-60,150-
