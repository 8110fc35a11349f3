"""In-kernel clock stamps of the tensor-core solve (variant 2): where does a stage's time go?"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchcde_b200 as cde
from torchcde_b200 import _lib
B, L, C, H = 4096, 256, 8, 32
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, L, C, generator=gen, device=dev).cumsum(1) / math.sqrt(L)
z0 = torch.randn(B, H, generator=gen, device=dev)
torch.manual_seed(1)
func = cde.LinearVectorField(H, C).to(dev)
t = torch.tensor([0.0, L - 1.0])
trace = torch.zeros(64, 8, dtype=torch.int64, device=dev)
with torch.no_grad():
    X = cde.CubicSpline(cde.hermite_cubic_coefficients_with_backward_differences(x))
    for variant in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "2").split(",")]:
        _lib.call("tcde_set_solve_variant", variant)
        cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options={"step_size": 1.0})
        trace.zero_()
        _lib.call("tcde_set_trace_buffer", _lib.ptr(trace))
        cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options={"step_size": 1.0})
        torch.cuda.synchronize()
        _lib.call("tcde_set_trace_buffer", None)
        tr = trace.cpu()
        print("variant", variant, "columns: mma_start commit_issued | row: before_wait after_wait after_contract after_rk after_arrive")
        base = int(tr[20, 0])
        for st in range(20, 30):
            print(st, [int(v) - base for v in tr[st, :7]])
        per = (tr[40, 0] - tr[20, 0]).item() / 20
        print("period per stage (cycles):", per,
              " issue->commit:", (tr[20:40, 1] - tr[20:40, 0]).float().mean().item(),
              " commit->row sees d_ready:", (tr[20:40, 3] - tr[20:40, 1]).float().mean().item(),
              " contraction:", (tr[20:40, 4] - tr[20:40, 3]).float().mean().item(),
              " rk:", (tr[20:40, 5] - tr[20:40, 4]).float().mean().item(),
              " split+store+arrive:", (tr[20:40, 6] - tr[20:40, 5]).float().mean().item(),
              " arrive->next mma start:", (tr[21:41, 0] - tr[20:40, 6]).float().mean().item(),
              " pre-wait work:", (tr[21:41, 2] - tr[20:40, 6]).float().mean().item())
    _lib.call("tcde_set_solve_variant", 0)
