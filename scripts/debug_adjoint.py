import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchcde_b200 as cde
from torchcde_b200 import adaptive, solver
DEV = "cuda"
torch.manual_seed(11)
B, L, C, H = 512, 14, 8, 32
x = torch.randn(B, L, C, device=DEV).cumsum(1) / 3
X = cde.CubicSpline(cde.hermite_cubic_coefficients_with_backward_differences(x))
func = cde.LinearVectorField(H, C).to(DEV)
with torch.no_grad():
    func.linear.weight.mul_(0.5)
z0 = torch.randn(B, H, device=DEV)
import sys as _s
t = torch.tensor([0.0, 5.5, L - 1.0], device=DEV) if len(_s.argv) > 1 else torch.tensor([0.0, L - 1.0], device=DEV)

def run(**kw):
    zz = z0.clone().requires_grad_(True)
    func.zero_grad()
    out = cde.cdeint(X, func, zz, t, adjoint=True, **kw)
    (out[:, -1].sum() + (out[:, 1] ** 2).sum()).backward() if t.numel() == 3 else out[:, -1].sum().backward()
    return out.detach(), zz.grad.clone(), func.linear.weight.grad.clone(), func.linear.bias.grad.clone()

truth = run(method="rk4", options={"step_size": 0.02})
print("truth (rk4 h=0.02): |gz| max", float(truth[1].abs().max()), "|gw| max", float(truth[2].abs().max()))
for tol in (1e-4, 1e-6, 1e-8):
    kw = {"rtol": tol, "atol": tol * 1e-2}
    real = adaptive._device_adaptive_backward
    dev = run(**kw)
    st = cde.cdeint.last_adjoint_stats
    adaptive._device_adaptive_backward = lambda *a, **k: None
    host = run(**kw)
    adaptive._device_adaptive_backward = real
    for name, i in (("gz", 1), ("gw", 2), ("gb", 3)):
        sc = float(truth[i].abs().max())
        print("tol %g %s: device-truth %.3e  host-truth %.3e  device-host %.3e (rel. to max |truth|)  stats %s" % (
            tol, name, float((dev[i] - truth[i]).abs().max()) / sc, float((host[i] - truth[i]).abs().max()) / sc,
            float((dev[i] - host[i]).abs().max()) / sc, st if i == 1 else ""))
    print("   forward out: device==host", bool(torch.equal(dev[0], host[0])), " |out-truth| max", float((dev[0] - truth[0]).abs().max()))
