"""Time the solve kernel variants at the BASELINE shape (CUDA events around the C-ABI call)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchcde_b200 as cde
from torchcde_b200 import _lib
B, L, C, H = 65536, 256, 8, 32
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, L, C, generator=gen, device=dev).cumsum(1) / math.sqrt(L)
z0 = torch.randn(B, H, generator=gen, device=dev)
torch.manual_seed(1)
func = cde.LinearVectorField(H, C).to(dev)
t = torch.tensor([0.0, L - 1.0])
with torch.no_grad():
    X = cde.CubicSpline(cde.hermite_cubic_coefficients_with_backward_differences(x))
    for v in [int(a) for a in sys.argv[1].split(",")]:
        _lib.call("tcde_set_solve_variant", v)
        for _ in range(2):
            cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options={"step_size": 1.0})
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options={"step_size": 1.0})
        e.record()
        torch.cuda.synchronize()
        print("variant", v, "ms", s.elapsed_time(e) / 5)
    _lib.call("tcde_set_solve_variant", 0)
