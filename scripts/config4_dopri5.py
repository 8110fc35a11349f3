"""BASELINE config 4: cdeint dopri5 (the reference's default method), forward and adjoint backward,
at the config-3 shapes.  Prints wall-clock seconds (CUDA-synchronised) and step counts."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchcde_b200 as cde

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
L, C, H = 256, 8, 32
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, L, C, generator=gen, device=dev).cumsum(1) / math.sqrt(L)
z0 = torch.randn(B, H, generator=gen, device=dev)
torch.manual_seed(1)
func = cde.LinearVectorField(H, C).to(dev)
with torch.no_grad():
    X = cde.CubicSpline(cde.hermite_cubic_coefficients_with_backward_differences(x))
t = torch.tensor([0.0, L - 1.0])
for trial in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.no_grad():
        out = cde.cdeint(X, func, z0, t, adjoint=False)          # dopri5, rtol 1e-4, atol 1e-6
    torch.cuda.synchronize(); fwd = time.perf_counter() - t0
print("dopri5 forward: B={} {:.3f} s  -> {:.0f} sequences/s  (|z_T|max {:.2f})".format(B, fwd, B / fwd, float(out.abs().max())))
Bb = min(B, 8192)
zb = z0[:Bb].clone().requires_grad_(True)
Xb = cde.CubicSpline(X._rows()[:Bb].contiguous())
torch.cuda.synchronize(); t0 = time.perf_counter()
outb = cde.cdeint(Xb, func, zb, t, adjoint=True)
outb[:, -1].sum().backward()
torch.cuda.synchronize(); both = time.perf_counter() - t0
print("dopri5 forward + adjoint backward: B={} {:.3f} s -> {:.0f} sequences/s".format(Bb, both, Bb / both))
