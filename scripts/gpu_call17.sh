#!/bin/bash
# round 2, call 17: gap fill with hole bytes (parity + timing), double-buffered stage dump (parity + training-step launches)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_builders.py tests/test_gpu_round2.py tests/test_gpu_adaptive.py -q -k "fill or linear or hermite or fused or trajectory or parameter_gradient or adjoint" > gpurun_out/r02_tests_c17.txt 2>&1
tail -6 gpurun_out/r02_tests_c17.txt
python - > gpurun_out/r02_builders_time3.txt 2>&1 <<'P'
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
for v in (0,3):
    _lib.call("tcde_set_natural_variant", v)
    print("variant",v,"series->hermite (no nan)     ms", tm(lambda:_lib.call("tcde_hermite_bdiff_coeffs_series",_lib.ptr(x),None,_lib.ptr(rows),B,L,C,code,None,st)))
    print("variant",v,"series->hermite (30% nan)    ms", tm(lambda:_lib.call("tcde_hermite_bdiff_coeffs_series",_lib.ptr(xn),None,_lib.ptr(rows),B,L,C,code,None,st)))
    print("variant",v,"fill (30% nan)               ms", tm(lambda:_lib.call("tcde_linear_fill",_lib.ptr(xn),None,_lib.ptr(filled),B,L,C,code,None,st)))
    print("variant",v,"fill (no nan)                ms", tm(lambda:_lib.call("tcde_linear_fill",_lib.ptr(x),None,_lib.ptr(filled),B,L,C,code,None,st)))
_lib.call("tcde_set_natural_variant", 0)
P
cat gpurun_out/r02_builders_time3.txt
TCDE_VERBOSE=1 TCDE_REPS=4 timeout 300 python scripts/adjoint_bench.py 65536 0 > gpurun_out/r02_adjoint_bench6.txt 2>&1
cat gpurun_out/r02_adjoint_bench6.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train4.csv python scripts/adjoint_bench.py 65536 0 > /dev/null 2>&1
grep -E "cdeint_tc_kernel|param_grad" gpurun_out/r02_launches_train4.csv | tail -4 | rev | cut -c1-20 | rev
