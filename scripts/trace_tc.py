"""In-kernel clock stamps of the round-2 tensor-core solve (variants 3 / 4) and timings of its profiling experiments."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchcde_b200 as cde
from torchcde_b200 import _lib
L, C, H = 256, 8, 32
dev = torch.device("cuda")
opts = {"step_size": 1.0}
t = torch.tensor([0.0, L - 1.0])
def problem(B):
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, L, C, generator=gen, device=dev).cumsum(1) / math.sqrt(L)
    z0 = torch.randn(B, H, generator=gen, device=dev)
    torch.manual_seed(1)
    func = cde.LinearVectorField(H, C).to(dev)
    with torch.no_grad():
        X = cde.CubicSpline(cde.hermite_cubic_coefficients_with_backward_differences(x))
    return X, func, z0
with torch.no_grad():
    X, func, z0 = problem(148 * 256)          # exactly one CTA per SM, both tiles live
    trace = torch.zeros(64, 8, dtype=torch.int64, device=dev)
    for variant in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '4,3').split(',')]:
        _lib.call("tcde_set_solve_variant", variant)
        cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options=opts)
        trace.zero_()
        _lib.call("tcde_set_trace_buffer", _lib.ptr(trace))
        cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options=opts)
        torch.cuda.synchronize()
        _lib.call("tcde_set_trace_buffer", None)
        tr = trace.cpu()
        print("variant", variant, "stamps: issue_start commit_issued | before_wait after_wait after_contract after_rk after_split+issue")
        base = int(tr[20, 0])
        for st in range(20, 26):
            print(st, [int(v) - base for v in tr[st, :7]])
        f = lambda a: float(a.float().mean())
        print("period {:.0f}: issue {:.0f} | commit->d_ready seen {:.0f} | contraction {:.0f} | rk {:.0f} | split+store+arrive+issue(next) {:.0f} | "
              "shadow (after issue -> before wait) {:.0f}".format(
                  f(tr[21:41, 0] - tr[20:40, 0]), f(tr[20:40, 1] - tr[20:40, 0]), f(tr[20:40, 3] - tr[20:40, 1]),
                  f(tr[20:40, 4] - tr[20:40, 3]), f(tr[20:40, 5] - tr[20:40, 4]), f(tr[20:40, 6] - tr[20:40, 5]),
                  f(tr[21:41, 2] - tr[20:40, 6])))
    X, func, z0 = problem(65536)
    for variant in [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else '0,1028,260,1284,3,2').split(',')]:
        _lib.call("tcde_set_solve_variant", variant)
        for _ in range(2):
            cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options=opts)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options=opts)
        b.record()
        torch.cuda.synchronize()
        print("variant {} debug flags {:2d}: {:.3f} ms".format(variant & 15, variant >> 4, a.elapsed_time(b) / 5), flush=True)
    _lib.call("tcde_set_solve_variant", 0)
