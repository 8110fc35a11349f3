"""Time cdeint(adjoint=True) forward + backward at the BASELINE field shape (length 256, channels 8,
hidden 32, rk4, step 1) with the fused adjoint stage kernel and with autograd serving the backward solve.
usage: python scripts/adjoint_bench.py [batch_fused] [batch_autograd]"""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import torchcde_b200 as cde  # noqa: E402
from torchcde_b200 import solver  # noqa: E402

L, C, H = 256, 8, 32
dev = torch.device("cuda")


def run(batch, fused, method="rk4", reps=int(os.environ.get("TCDE_REPS", "2"))):
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(batch, L, C, generator=gen, device=dev).cumsum(1) / math.sqrt(L)
    z0 = torch.randn(batch, H, generator=gen, device=dev)
    torch.manual_seed(1)
    func = cde.LinearVectorField(H, C).to(dev)
    with torch.no_grad():
        X = cde.CubicSpline(cde.hermite_cubic_coefficients_with_backward_differences(x))
    t = X.interval
    real = solver._kernel_vjp
    if not fused:
        solver._kernel_vjp = lambda *a, **k: None
    kw = {"method": method, "options": {"step_size": 1.0}} if method != "dopri5" else {"method": method}
    try:
        best = None
        for _ in range(reps):
            zz = z0.clone().requires_grad_(True)
            func.zero_grad()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = cde.cdeint(X, func, zz, t, adjoint=True, **kw)
            out[:, -1].sum().backward()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if os.environ.get("TCDE_VERBOSE"):
                st = torch.cuda.memory_stats()
                print("rep: {:.4f} s  cudaMalloc calls so far {}  retries {}".format(dt, st.get("num_device_alloc", -1), st.get("num_alloc_retries", -1)), flush=True)
            best = dt if best is None else min(best, dt)
        return {"batch": batch, "fused_stage": fused, "method": method, "seconds_fwd_bwd": best,
                "sequences_per_s": batch / best, "grad_weight_norm": float(func.linear.weight.grad.norm())}
    finally:
        solver._kernel_vjp = real


if __name__ == "__main__":
    b_fused = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    b_auto = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    rows = [run(b_fused, True)] + ([run(b_auto, False), run(b_auto, True)] if b_auto > 0 else [])
    for r in rows:
        print(json.dumps(r))
