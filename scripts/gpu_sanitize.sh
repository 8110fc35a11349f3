#!/bin/bash
# compute-sanitizer memcheck over small instances of every new kernel path of round 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SAN="compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0"
run() { echo "== $*" >> gpurun_out/r02_sanitizer.txt; timeout 900 $SAN python -m pytest -q -x "$@" >> gpurun_out/r02_sanitizer.txt 2>&1; echo "exit $?" >> gpurun_out/r02_sanitizer.txt; }
rm -f gpurun_out/r02_sanitizer.txt
run tests/test_gpu_solve.py -k "matches_cuda_core_and_oracle and (129 or 300) and (variant3 or variant4)"
run tests/test_gpu_adaptive.py -k "trajectory or parameter_gradient"
run tests/test_gpu_round2.py -k "fused_fill or contract or generic"
run tests/test_gpu_dopri5_device.py -k "matches_the_host_driver"
run tests/test_gpu_logsig.py
grep -E "^==|exit|ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitizer.txt | tail -30
