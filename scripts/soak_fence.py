"""Soak test of the issuer-side proxy fence: the fused solve must give the SAME BITS with the fence only in the MMA issuer
(default) as with a fence in every row thread as well (debug bit 1), over many launches and several problems; and repeated
training steps (dumping solves + parameter-gradient GEMM) must reproduce their gradients bit for bit."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchcde_b200 as cde
from torchcde_b200 import _lib

L, C, H = 256, 8, 32
dev = torch.device("cuda")
opts = {"step_size": 1.0}
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad = 0
for seed, B in ((0, 65536), (1, 65536), (2, 37 * 256 + 77), (3, 148 * 256)):
    gen = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(B, L, C, generator=gen, device=dev).cumsum(1) / math.sqrt(L)
    z0 = torch.randn(B, H, generator=gen, device=dev)
    torch.manual_seed(seed + 10)
    func = cde.LinearVectorField(H, C).to(dev)
    with torch.no_grad():
        X = cde.CubicSpline(cde.hermite_cubic_coefficients_with_backward_differences(x))
        t = torch.tensor([0.0, 100.5, L - 1.0])
        _lib.call("tcde_set_solve_variant", 4 + 16 * 2)
        ref = cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options=opts)
        _lib.call("tcde_set_solve_variant", 0)
        for i in range(reps):
            out = cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options=opts)
            if not torch.equal(out, ref):
                bad += 1
                print("MISMATCH seed", seed, "rep", i, float((out - ref).abs().max()), flush=True)
    print("seed", seed, "batch", B, ":", reps, "solves identical to the row-fence result:", bad == 0, flush=True)
    del X, x
# training step reproducibility
gen = torch.Generator(device=dev).manual_seed(5)
B = 65536
x = torch.randn(B, L, C, generator=gen, device=dev).cumsum(1) / math.sqrt(L)
z0 = torch.randn(B, H, generator=gen, device=dev)
torch.manual_seed(15)
func = cde.LinearVectorField(H, C).to(dev)
with torch.no_grad():
    X = cde.CubicSpline(cde.hermite_cubic_coefficients_with_backward_differences(x))
first = None
for i in range(8):
    zz = z0.clone().requires_grad_(True)
    func.zero_grad()
    out = cde.cdeint(X, func, zz, X.interval, adjoint=True, method="rk4", options=opts)
    out[:, -1].sum().backward()
    g = (func.linear.weight.grad.clone(), func.linear.bias.grad.clone(), zz.grad.clone())
    if first is None:
        first = g
    elif not all(torch.equal(a, b) for a, b in zip(g, first)):
        bad += 1
        print("TRAINING MISMATCH rep", i, [float((a - b).abs().max()) for a, b in zip(g, first)], flush=True)
print("8 training steps reproduce their gradients bit for bit:", bad == 0)
print("SOAK", "OK" if bad == 0 else "FAILED")
sys.exit(1 if bad else 0)
