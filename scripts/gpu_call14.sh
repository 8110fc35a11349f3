#!/bin/bash
# round 2, call 14: param-grad kernel A/B (8 / 16 producer warps x producer / issuer proxy fence), then parity of the chosen mode
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for m in 0 1 2 3; do
  echo "TCDE_PG_MODE=$m" >> gpurun_out/r02_pg_modes.txt
  TCDE_PG_MODE=$m TCDE_REPS=4 timeout 300 python scripts/adjoint_bench.py 65536 0 >> gpurun_out/r02_pg_modes.txt 2>&1
  TCDE_PG_MODE=$m timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:param_grad_bf16 python scripts/adjoint_bench.py 65536 0 2>/dev/null | grep param_grad | tail -1 | rev | cut -c1-24 | rev >> gpurun_out/r02_pg_modes.txt
done
cat gpurun_out/r02_pg_modes.txt
TCDE_PG_MODE=1 timeout 600 python -m pytest tests/test_gpu_adaptive.py -q -k "parameter_gradient or trajectory or fused_adjoint" > gpurun_out/r02_tests_c14.txt 2>&1
tail -3 gpurun_out/r02_tests_c14.txt
