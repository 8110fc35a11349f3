#!/bin/bash
# round 2, call 10: graph fix + fused fill/Hermite tests, per-kernel times of the training step, ncu of the fill and param-grad kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_builders.py -q -x > gpurun_out/r02_tests_c10.txt 2>&1
tail -5 gpurun_out/r02_tests_c10.txt
python - > gpurun_out/r02_builders_time.txt 2>&1 <<'P'
import math, torch, time
import torchcde_b200 as cde
from torchcde_b200 import _lib
B,L,C=65536,256,8
dev="cuda"
x=torch.randn(B,L,C,device=dev).cumsum(1)/16
xn=x.clone(); hole=torch.rand(x.shape,device=dev)<0.3; hole[:,0]=False; hole[:,-1]=False; xn[hole]=float("nan"); del hole
rows=torch.empty(B,L-1,4*C,device=dev); filled=torch.empty_like(x)
code=_lib.dtype_code(x.dtype); st=_lib.stream_of(x)
def tm(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
print("hermite (no nan) old kernel  ms", tm(lambda:_lib.call("tcde_hermite_bdiff_coeffs",_lib.ptr(x),None,_lib.ptr(rows),B,L,C,code,None,st)))
print("series->hermite (no nan)     ms", tm(lambda:_lib.call("tcde_hermite_bdiff_coeffs_series",_lib.ptr(x),None,_lib.ptr(rows),B,L,C,code,None,st)))
print("series->hermite (30% nan)    ms", tm(lambda:_lib.call("tcde_hermite_bdiff_coeffs_series",_lib.ptr(xn),None,_lib.ptr(rows),B,L,C,code,None,st)))
print("fill (30% nan)               ms", tm(lambda:_lib.call("tcde_linear_fill",_lib.ptr(xn),None,_lib.ptr(filled),B,L,C,code,None,st)))
print("fill (no nan)                ms", tm(lambda:_lib.call("tcde_linear_fill",_lib.ptr(x),None,_lib.ptr(filled),B,L,C,code,None,st)))
print("api hermite nan ms", tm(lambda:cde.hermite_cubic_coefficients_with_backward_differences(xn),5))
P
cat gpurun_out/r02_builders_time.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train.csv python scripts/adjoint_bench.py 65536 0 > gpurun_out/r02_train_under_ncu.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:linear_fill_scan -c 3 -f -o gpurun_out/r02_fill python scripts/profile_fill.py > gpurun_out/r02_ncu_fill.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:param_grad_bf16 -c 1 -f -o gpurun_out/r02_pg python scripts/adjoint_bench.py 65536 0 > gpurun_out/r02_ncu_pg.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
