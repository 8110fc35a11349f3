#!/bin/bash
# round 2, call 20: contraction kernel parity + generic MLP-field timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_adaptive.py -q -k "generic or contract or prod" > gpurun_out/r02_tests_c20.txt 2>&1
tail -4 gpurun_out/r02_tests_c20.txt
python - > gpurun_out/r02_generic_mlp.txt 2>&1 <<'P'
import math, torch
import torchcde_b200 as cde
B,L=65536,256
dev="cuda"
torch.manual_seed(0)
x=torch.randn(B,L,3,device=dev).cumsum(1)/16
class F(torch.nn.Module):
    def __init__(s):
        super().__init__(); s.l1=torch.nn.Linear(8,128); s.l2=torch.nn.Linear(128,24)
    def forward(s,t,z):
        z=s.l2(s.l1(z).relu()).tanh(); return z.view(*z.shape[:-1],8,3)
f=F().to(dev); z0=torch.randn(B,8,device=dev)
with torch.no_grad():
    X=cde.CubicSpline(cde.hermite_cubic_coefficients_with_backward_differences(x)); t=X.interval
    for label,opts in (("kernel loop",{"step_size":1.0}),("cuda graph",{"step_size":1.0,"cuda_graph":True})):
        for _ in range(2): out=cde.cdeint(X,f,z0,t,adjoint=False,method="rk4",options=opts)
        torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): out=cde.cdeint(X,f,z0,t,adjoint=False,method="rk4",options=opts)
        b.record(); torch.cuda.synchronize()
        print(label, a.elapsed_time(b)/3, "ms per solve", B/(a.elapsed_time(b)/3*1e-3)/1e6, "M seq/s")
P
cat gpurun_out/r02_generic_mlp.txt
