mkdir -p gpurun_out
timeout 300 python scripts/trace_tc.py > gpurun_out/r02_trace_tc3.txt 2>&1
timeout 600 python scripts/debug_dopri5.py > gpurun_out/r02_dopri5_debug.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_gpu_fp16_split.py -q -x -k "tensor_core or decreasing or split" 2>&1 | tail -8 > gpurun_out/r02_tests_tc3.txt
cat gpurun_out/r02_trace_tc3.txt gpurun_out/r02_dopri5_debug.txt gpurun_out/r02_tests_tc3.txt
