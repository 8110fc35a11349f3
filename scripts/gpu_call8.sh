mkdir -p gpurun_out
scripts/ubench/build/mn_major_test > gpurun_out/r02_mn_major.txt 2>&1
timeout 600 python scripts/adjoint_bench.py 65536 64 > gpurun_out/r02_adjoint_bench.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_logsig.py tests/test_gpu_round2.py -q 2>&1 | tail -15 > gpurun_out/r02_tests_logsig.txt
cat gpurun_out/r02_mn_major.txt; tail -4 gpurun_out/r02_adjoint_bench.txt; cat gpurun_out/r02_tests_logsig.txt
