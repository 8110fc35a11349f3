mkdir -p gpurun_out
timeout 300 python scripts/trace_tc.py > gpurun_out/r02_trace_tc.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dopri5_device.py tests/test_gpu_tricks.py -q -x -s 2>&1 | tail -30 > gpurun_out/r02_tests_dopri5.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cdeint_tc -s 2 -c 1 -o gpurun_out/r02_tc_fp16 python scripts/time_variants.py 4 > gpurun_out/r02_ncu_log.txt 2>&1
cat gpurun_out/r02_trace_tc.txt gpurun_out/r02_tests_dopri5.txt; tail -3 gpurun_out/r02_ncu_log.txt
