#!/bin/bash
# round 2, call 16: param-grad with relaxed waits (timing + full ncu capture), then the whole GPU test suite
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:param_grad_bf16 python scripts/adjoint_bench.py 65536 0 2>/dev/null | grep param_grad | tail -1 | rev | cut -c1-24 | rev > gpurun_out/r02_pg_relaxed.txt
cat gpurun_out/r02_pg_relaxed.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:param_grad_bf16 -c 1 -f -o gpurun_out/r02_pg2 python scripts/adjoint_bench.py 65536 0 > gpurun_out/r02_ncu_pg2.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_full_c16.txt 2>&1
tail -8 gpurun_out/r02_tests_full_c16.txt
