mkdir -p gpurun_out
timeout 300 python scripts/trace_tc.py > gpurun_out/r02_trace_tc5.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_tests_all.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
cat gpurun_out/r02_trace_tc5.txt gpurun_out/r02_tests_all.txt; tail -5 gpurun_out/r02_bench_a.err; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_a.json'));print({k:d[k] for k in ('value','ms_per_step','clocks')});print(d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['kind']);print(d['roofline']['frac'], d['roofline']['mma']);print(d['config4']);print({k:(v.get('ms'),v.get('kernel_ms'),v.get('frac')) for k,v in d['roofline']['secondary'].items()})"
