#!/bin/bash
# round 2, final evidence run (one GPU): full GPU test suite, smoke, both bench arms, launch list of the bench command, full ncu capture of the solve
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_final.txt 2>&1
tail -4 gpurun_out/r02_tests_final.txt
timeout 600 python scripts/soak_fence.py 40 > gpurun_out/r02_soak_fence.txt 2>&1
tail -7 gpurun_out/r02_soak_fence.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke_final.txt 2>&1
tail -2 gpurun_out/r02_smoke_final.txt
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 1200 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -c 600 gpurun_out/r02_bench_final.err
head -c 1500 gpurun_out/r02_bench_final.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cdeint_tc -c 1 -f -o gpurun_out/r02_tc_last python scripts/profile_targets.py 1 0 > gpurun_out/r02_ncu_tc_last.log 2>&1
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_under_ncu.txt 2>&1
wc -l gpurun_out/r02_launches_final.csv
