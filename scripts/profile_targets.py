"""Launch each hot kernel a few times at the BASELINE shapes -- the target of the ncu captures
whose summaries are committed under profiles/ (see profiles/README.md for the commands)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import torchcde_b200 as cde  # noqa: E402
from torchcde_b200 import _lib  # noqa: E402

B, L, C, H = 65536, 256, 8, 32
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1]
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, L, C, generator=gen, device=dev).cumsum(1) / math.sqrt(L)
z0 = torch.randn(B, H, generator=gen, device=dev)
torch.manual_seed(1)
func = cde.LinearVectorField(H, C).to(dev)
xn = x.clone()
mask = torch.rand(B, L, C, generator=gen, device=dev) < 0.3
mask[:, 0] = False
mask[:, -1] = False
xn[mask] = float("nan")
del mask
t = torch.tensor([0.0, L - 1.0])
with torch.no_grad():
    for _ in range(reps):
        coeffs = cde.hermite_cubic_coefficients_with_backward_differences(x)
        nat = cde.natural_cubic_coeffs(x)
        del nat
        filled = cde.linear_interpolation_coeffs(xn)
        del filled
        X = cde.CubicSpline(coeffs)
        for v in variants:
            _lib.call("tcde_set_solve_variant", v)
            out = cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options={"step_size": 1.0})
        _lib.call("tcde_set_solve_variant", 0)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
