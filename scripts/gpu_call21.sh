#!/bin/bash
# round 2, call 21: dopri5 adjoint with the controller on the device: parity with the host-driven adjoint, config-4 timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_adaptive.py tests/test_gpu_dopri5_device.py tests/test_gpu_tricks.py -q -x > gpurun_out/r02_tests_c21.txt 2>&1
tail -25 gpurun_out/r02_tests_c21.txt
python - > gpurun_out/r02_config4_adjoint.txt 2>&1 <<'P'
import math, time, torch
import torchcde_b200 as cde
B,L,C,H=65536,256,8,32
dev="cuda"
gen=torch.Generator(device=dev).manual_seed(0)
x=torch.randn(B,L,C,generator=gen,device=dev).cumsum(1)/math.sqrt(L)
z0=torch.randn(B,H,generator=gen,device=dev)
torch.manual_seed(1)
func=cde.LinearVectorField(H,C).to(dev)
with torch.no_grad():
    X=cde.CubicSpline(cde.hermite_cubic_coefficients_with_backward_differences(x))
for rep in range(3):
    zz=z0.clone().requires_grad_(True); func.zero_grad()
    torch.cuda.synchronize(); t0=time.perf_counter()
    out=cde.cdeint(X,func,zz,X.interval,adjoint=True)
    torch.cuda.synchronize(); t1=time.perf_counter()
    out[:,-1].sum().backward()
    torch.cuda.synchronize(); t2=time.perf_counter()
    print("rep",rep,"forward %.3f s  backward %.3f s"%(t1-t0,t2-t1), "fwd stats",cde.cdeint.last_stats, "adjoint stats",getattr(cde.cdeint,"last_adjoint_stats",None), "gw norm", float(func.linear.weight.grad.norm()), flush=True)
from torchcde_b200 import adaptive
adaptive._device_adaptive_backward = lambda *a, **k: None
zz=z0.clone().requires_grad_(True); gw_dev=func.linear.weight.grad.clone(); func.zero_grad()
out=cde.cdeint(X,func,zz,X.interval,adjoint=True)
torch.cuda.synchronize(); t1=time.perf_counter()
out[:,-1].sum().backward()
torch.cuda.synchronize(); t2=time.perf_counter()
gw_host=func.linear.weight.grad
print("host-driven backward %.3f s; max |gw_dev - gw_host| / scale = %.3e"%(t2-t1, float((gw_dev-gw_host).abs().max()/gw_host.abs().max())))
P
cat gpurun_out/r02_config4_adjoint.txt
compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python -m pytest -q -x tests/test_gpu_solve.py -k "matches_cuda_core_and_oracle and 129" > gpurun_out/r02_sanitizer2.txt 2>&1; echo "exit $?" >> gpurun_out/r02_sanitizer2.txt
compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python -m pytest -q -x tests/test_gpu_adaptive.py -k "device_controlled_dopri5_adjoint and cubic" >> gpurun_out/r02_sanitizer2.txt 2>&1; echo "exit $?" >> gpurun_out/r02_sanitizer2.txt
grep -E "exit|ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitizer2.txt
