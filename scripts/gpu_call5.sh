mkdir -p gpurun_out
timeout 300 python scripts/trace_tc.py > gpurun_out/r02_trace_tc4.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_gpu_fp16_split.py -q -x -k "tensor_core or decreasing or split" 2>&1 | tail -8 > gpurun_out/r02_tests_tc4.txt
timeout 900 python -m pytest tests/test_gpu_dopri5_device.py tests/test_gpu_tricks.py tests/test_gpu_round2.py -q -x 2>&1 | tail -15 > gpurun_out/r02_tests_new4.txt
cat gpurun_out/r02_trace_tc4.txt gpurun_out/r02_tests_tc4.txt gpurun_out/r02_tests_new4.txt
