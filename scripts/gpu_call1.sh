mkdir -p gpurun_out
python scripts/probe_torchdiffeq.py > gpurun_out/r02_torchdiffeq_probe.txt 2>&1
scripts/ubench/build/mma_rates > gpurun_out/r02_ubench_mma.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> gpurun_out/r02_ubench_mma.txt
timeout 300 python scripts/time_variants.py 2,3,4 > gpurun_out/r02_variants.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_solve.py tests/test_gpu_fp16_split.py -q -x -k "tensor_core or decreasing or split" 2>&1 | tail -15 > gpurun_out/r02_tests_tc.txt
timeout 600 python -m pytest tests/test_gpu_tricks.py tests/test_gpu_round2.py -q 2>&1 | tail -25 > gpurun_out/r02_tests_new.txt
cat gpurun_out/r02_ubench_mma.txt gpurun_out/r02_variants.txt gpurun_out/r02_tests_tc.txt gpurun_out/r02_tests_new.txt
