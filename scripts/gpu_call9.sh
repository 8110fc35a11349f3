mkdir -p gpurun_out
timeout 300 python scripts/debug_graph.py > gpurun_out/r02_debug_graph.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_adaptive.py -q -x -k "parameter_gradient or trajectory or fused_adjoint" 2>&1 | tail -15 > gpurun_out/r02_tests_pg.txt
timeout 600 python scripts/adjoint_bench.py 65536 64 > gpurun_out/r02_adjoint_bench2.txt 2>&1
cat gpurun_out/r02_debug_graph.txt | tail -8; cat gpurun_out/r02_tests_pg.txt; tail -3 gpurun_out/r02_adjoint_bench2.txt
