#!/bin/bash
# round 2, final tree: full GPU test suite, smoke, both bench arms
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_final.txt 2>&1
tail -4 gpurun_out/r02_tests_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke_final.txt 2>&1
tail -2 gpurun_out/r02_smoke_final.txt
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 1200 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -c 400 gpurun_out/r02_bench_final.err
head -c 600 gpurun_out/r02_bench_final.json
