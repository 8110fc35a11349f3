"""Device-controlled dopri5 against the host-driven driver: step counts, agreement, timing."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchcde_b200 as cde
from torchcde_b200 import adaptive
dev = "cuda"
def problem(batch, length, seed=0, linear=False):
    gen = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(batch, length, 8, generator=gen, device=dev).cumsum(1) / math.sqrt(length)
    z0 = torch.randn(batch, 32, generator=gen, device=dev)
    torch.manual_seed(1)
    func = cde.LinearVectorField(32, 8).to(dev)
    with torch.no_grad():
        X = cde.LinearInterpolation(x) if linear else cde.CubicSpline(cde.hermite_cubic_coefficients_with_backward_differences(x))
    return X, func, z0
avail = adaptive.device_dopri5_available
for batch, length, linear in ((1000, 24, False), (257, 17, True), (128, 40, False), (65536, 256, False)):
    X, func, z0 = problem(batch, length, linear=linear)
    t = torch.tensor([0.0, 0.37 * (length - 1), length - 1.0])
    with torch.no_grad():
        res = {}
        for name, fn in (("device", avail), ("host", lambda *a: False)):
            if name == "host" and batch > 8192:
                continue
            adaptive.device_dopri5_available = fn
            cde.cdeint(X, func, z0, t, adjoint=False)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = cde.cdeint(X, func, z0, t, adjoint=False)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            res[name] = (out, dict(cde.cdeint.last_stats), dt)
        adaptive.device_dopri5_available = avail
        n = min(batch, 512)
        Xs = type(X)((X._rows() if not linear else X._coeffs)[:n].contiguous())
        tight = cde.cdeint(Xs, func, z0[:n].contiguous(), t, adjoint=False, method="rk4", options={"step_size": 1.0 / 32})
        scale = float(tight.abs().max())
        for name, (out, stats, dt) in res.items():
            print("B={} L={} linear={} {:6s}: {} | {:.4f} s | err vs tight rk4 {:.3e} (scale {:.2e}) finite={}".format(
                batch, length, linear, name, stats, dt, float((out[:n] - tight).abs().max()), scale, bool(torch.isfinite(out).all())), flush=True)
        if "host" in res:
            print("   device vs host: {:.3e}".format(float((res["device"][0] - res["host"][0]).abs().max())))
