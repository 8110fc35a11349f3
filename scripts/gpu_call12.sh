#!/bin/bash
# round 2, call 12: the two-threads-per-path solve kernel (variant 5): parity, trace, timing; mask-based gap fill: parity + timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TCDE_VERBOSE=1 TCDE_REPS=6 timeout 300 python scripts/adjoint_bench.py 65536 0 > gpurun_out/r02_adjoint_bench4.txt 2>&1
cat gpurun_out/r02_adjoint_bench4.txt
timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_gpu_builders.py tests/test_gpu_round2.py -q -x -k "variant or fill or linear or hermite or fused" > gpurun_out/r02_tests_c12.txt 2>&1
tail -6 gpurun_out/r02_tests_c12.txt
timeout 300 python scripts/trace_tc.py 5,4 0,5,85,1029 > gpurun_out/r02_trace_tc2k.txt 2>&1
cat gpurun_out/r02_trace_tc2k.txt
timeout 200 python scripts/time_variants.py 4,5 > gpurun_out/r02_variants2.txt 2>&1
cat gpurun_out/r02_variants2.txt
python - > gpurun_out/r02_builders_time2.txt 2>&1 <<'P'
import torch
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
print("series->hermite (no nan)     ms", tm(lambda:_lib.call("tcde_hermite_bdiff_coeffs_series",_lib.ptr(x),None,_lib.ptr(rows),B,L,C,code,None,st)))
print("series->hermite (30% nan)    ms", tm(lambda:_lib.call("tcde_hermite_bdiff_coeffs_series",_lib.ptr(xn),None,_lib.ptr(rows),B,L,C,code,None,st)))
print("fill (30% nan)               ms", tm(lambda:_lib.call("tcde_linear_fill",_lib.ptr(xn),None,_lib.ptr(filled),B,L,C,code,None,st)))
print("fill (no nan)                ms", tm(lambda:_lib.call("tcde_linear_fill",_lib.ptr(x),None,_lib.ptr(filled),B,L,C,code,None,st)))
P
cat gpurun_out/r02_builders_time2.txt
