"""ncu target: the gap-fill kernel and the fused series -> Hermite kernel at the BASELINE shapes (30 % NaN)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torchcde_b200 import _lib  # noqa: E402

B, L, C = 65536, 256, 8
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(B, L, C, device=dev).cumsum(1) / 16
hole = torch.rand(x.shape, device=dev) < 0.3
hole[:, 0] = False
hole[:, -1] = False
xn = x.masked_fill(hole, float("nan"))
rows = torch.empty(B, L - 1, 4 * C, device=dev)
filled = torch.empty_like(x)
code = _lib.dtype_code(x.dtype)
st = _lib.stream_of(x)
_lib.call("tcde_linear_fill", _lib.ptr(xn), None, _lib.ptr(filled), B, L, C, code, None, st)
_lib.call("tcde_hermite_bdiff_coeffs_series", _lib.ptr(xn), None, _lib.ptr(rows), B, L, C, code, None, st)
_lib.call("tcde_hermite_bdiff_coeffs_series", _lib.ptr(x), None, _lib.ptr(rows), B, L, C, code, None, st)
torch.cuda.synchronize()
print("ok")
