#!/bin/bash
# round 2, call 15: param-grad kernel ring shapes (bytes in flight), then parity of the default
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/r02_pg_modes2.txt
for m in 8 0 4 6 5; do
  echo "TCDE_PG_MODE=$m" >> gpurun_out/r02_pg_modes2.txt
  TCDE_PG_MODE=$m timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:param_grad_bf16 python scripts/adjoint_bench.py 65536 0 2>/dev/null | grep param_grad | tail -1 | rev | cut -c1-24 | rev >> gpurun_out/r02_pg_modes2.txt
done
cat gpurun_out/r02_pg_modes2.txt
for m in 0 4; do
TCDE_PG_MODE=$m timeout 600 python -m pytest tests/test_gpu_adaptive.py -q -k "parameter_gradient or trajectory or fused_adjoint" > gpurun_out/r02_tests_c15_$m.txt 2>&1
tail -2 gpurun_out/r02_tests_c15_$m.txt
done
