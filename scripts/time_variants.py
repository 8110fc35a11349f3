"""Device-resident time of the fused RK4 solve at BASELINE config 3 for each kernel variant (CUDA events, 3 warm-ups,
10 timed) and its worst error against the fp64 CUDA-core solve on a 4,096-path slice."""
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
opts = {"step_size": 1.0}
variants = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "2,3,4,260").split(",")]
with torch.no_grad():
    coeffs = cde.hermite_cubic_coefficients_with_backward_differences(x)
    X = cde.CubicSpline(coeffs)
    f64 = cde.LinearVectorField(H, C, dtype=torch.float64).to(dev)
    f64.linear.weight.copy_(func.linear.weight.double())
    f64.linear.bias.copy_(func.linear.bias.double())
    n = 4096
    _lib.call("tcde_set_solve_variant", 1)
    X64 = cde.CubicSpline(coeffs[:n].double())
    want = cde.cdeint(X64, f64, z0[:n].double(), t.double(), adjoint=False, method="rk4", options=opts)
    scale = float(want.abs().max())
    for v in variants:
        _lib.call("tcde_set_solve_variant", v)
        try:
            for _ in range(3):
                out = cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options=opts)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                out = cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options=opts)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            err = float((out[:n].double() - want).abs().max())
            print("variant {}: {:.3f} ms per solve, {:.2f} M seq/s, max err vs fp64 {:.3e} (scale {:.3e}, ratio {:.2e}), finite={}".format(
                v, ms, B / ms / 1e3, err, scale, err / scale, bool(torch.isfinite(out).all())), flush=True)
        except Exception as exc:
            print("variant {}: FAILED {!r}".format(v, exc), flush=True)
    _lib.call("tcde_set_solve_variant", 0)
