#!/bin/bash
# round 2, call 18: stage-dump store moved off the commit chain: parity + training-step launches
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_adaptive.py tests/test_gpu_builders.py -q -k "trajectory or parameter_gradient or adjoint or fill" > gpurun_out/r02_tests_c18.txt 2>&1
tail -4 gpurun_out/r02_tests_c18.txt
TCDE_VERBOSE=1 TCDE_REPS=4 timeout 300 python scripts/adjoint_bench.py 65536 0 > gpurun_out/r02_adjoint_bench7.txt 2>&1
cat gpurun_out/r02_adjoint_bench7.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train5.csv python scripts/adjoint_bench.py 65536 0 > /dev/null 2>&1
grep -E "cdeint_tc_kernel|param_grad" gpurun_out/r02_launches_train5.csv | tail -4 | rev | cut -c1-20 | rev
