mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_tests_all2.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_adjoint.csv python scripts/adjoint_bench.py 65536 64 > gpurun_out/r02_adjoint_bench.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cdeint_tc_kernel -s 2 -c 1 -o gpurun_out/r02_tc_final python scripts/time_variants.py 0 > gpurun_out/r02_ncu_log2.txt 2>&1
cat gpurun_out/r02_tests_all2.txt; cat gpurun_out/r02_adjoint_bench.txt | tail -5; python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r02_launches_adjoint.csv', errors='ignore')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID']
agg = collections.defaultdict(lambda: [0, 0.0])
if hdr:
    h = rows[hdr[0]]
    ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    for r in rows[hdr[0] + 1:]:
        if len(r) > vi:
            try:
                v = float(r[vi].replace(',', ''))
            except ValueError:
                continue
            scale = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 'nsecond': 1e-6}.get(r[ui], 1e-6)
            agg[r[ki][:70]][0] += 1
            agg[r[ki][:70]][1] += v * scale
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print('%6d launches %10.3f ms  %s' % (n, ms, k))
PY
