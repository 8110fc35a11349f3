import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchcde_b200 as cde
from torchcde_b200 import solver
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(64, 20, 3, device=dev).cumsum(1) / 4
class F(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.l1 = torch.nn.Linear(8, 128); self.l2 = torch.nn.Linear(128, 24)
    def forward(self, t, z):
        z = self.l2(self.l1(z).relu()).tanh()
        return z.view(*z.shape[:-1], 8, 3)
func = F().to(dev)
z0 = torch.randn(64, 8, device=dev)
t = torch.tensor([0.0, 7.3, 19.0])
X = cde.CubicSpline(cde.natural_cubic_coeffs(x))
with torch.no_grad():
    for scale in (1.0, 0.5, 2.0):
        fast = cde.cdeint(X, func, z0 * scale, t, adjoint=False, method="rk4", options={"step_size": 1.0})
        graph = cde.cdeint(X, func, z0 * scale, t, adjoint=False, method="rk4", options={"step_size": 1.0, "cuda_graph": True})
        slow = solver._generic_solve(X, func, z0 * scale, t, "rk4", 1.0, False, True)
        print(scale, "fast-slow", float((fast - slow).abs().max()), "graph-fast", float((graph - fast).abs().max()),
              "graph t0 row ok", bool(torch.equal(graph[:, 0], z0 * scale)))
