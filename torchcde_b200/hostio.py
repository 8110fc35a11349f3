"""Solve with HOST-resident inputs: chunked, double-buffered host<->device pipeline.

The reference workflow keeps spline coefficients as the dataset on the host
(interpolation_cubic.py:216-226) and moves each batch to the device before ``cdeint``.  At
the BASELINE shapes that batch is 2.1 GB, so the copy -- not the solve -- is what a user
waits for unless the two overlap.  ``cdeint_from_host`` cuts the batch into chunks and runs
copy-in / fused solve / copy-out of consecutive chunks on two CUDA streams, so PCIe and the
SMs work at the same time.  Paths are independent, so chunking does not change any result.
"""
import torch

from . import _lib
from .controls import CubicSpline, LinearInterpolation
from .solver import cdeint


class HostPipeline:
    """Reusable staging buffers + streams for ``cdeint_from_host`` (allocate once, run many times)."""

    def __init__(self, device, chunk_paths, control_shape, hidden, n_out, dtype):
        self.device = torch.device(device)
        self.chunk_paths = chunk_paths
        self.streams = [torch.cuda.Stream(self.device) for _ in range(2)]
        self.control = [torch.empty(chunk_paths, *control_shape, dtype=dtype, device=self.device) for _ in range(2)]
        self.z0 = [torch.empty(chunk_paths, hidden, dtype=dtype, device=self.device) for _ in range(2)]
        self.n_out = n_out


def cdeint_from_host(control_host, func, z0_host, t, out_host=None, kind="cubic", knots=None, chunk_paths=8192,
                     pipeline=None, device=None, **kwargs):
    """``cdeint`` for pinned host tensors ``control_host`` (P, rows, width) and ``z0_host`` (P, H).

    Returns a pinned host tensor (P, len(t), H).  ``kwargs`` are ``cdeint``'s (``method``,
    ``options`` ...).  ``kind``: 'cubic' (coefficients of ``CubicSpline``) or 'linear' (knots of
    ``LinearInterpolation``)."""
    device = torch.device(device if device is not None else "cuda")
    n_paths, hidden = z0_host.shape
    n_out = t.numel()
    if out_host is None:
        out_host = torch.empty(n_paths, n_out, hidden, dtype=z0_host.dtype, pin_memory=True)
    if pipeline is None:
        pipeline = HostPipeline(device, min(chunk_paths, n_paths), tuple(control_host.shape[1:]), hidden, n_out,
                                z0_host.dtype)
    cp = pipeline.chunk_paths
    kwargs.setdefault("adjoint", False)
    current = torch.cuda.current_stream(device)
    for s in pipeline.streams:
        s.wait_stream(current)
    t_dev = t.detach().cpu()          # a CPU tensor: the schedule is host-side, nothing to sync on
    k_dev = None if knots is None else knots.to(device)
    with torch.no_grad():
        for i, lo in enumerate(range(0, n_paths, cp)):
            hi = min(lo + cp, n_paths)
            slot = i & 1
            with torch.cuda.stream(pipeline.streams[slot]):
                c_dev = pipeline.control[slot][:hi - lo]
                z_dev = pipeline.z0[slot][:hi - lo]
                c_dev.copy_(control_host[lo:hi], non_blocking=True)
                z_dev.copy_(z0_host[lo:hi], non_blocking=True)
                X = CubicSpline(c_dev, k_dev) if kind == "cubic" else LinearInterpolation(c_dev, k_dev)
                out = cdeint(X, func, z_dev, t_dev, **kwargs)
                out_host[lo:hi].copy_(out, non_blocking=True)
    for s in pipeline.streams:
        current.wait_stream(s)
    return out_host


class SeriesPipeline:
    """Persistent streams and staging buffers for ``cdeint_from_host_series`` (reused across calls:
    creating streams / device buffers per call would go through cudaMalloc every time).

    Four slots, not two: a chunk's solve (one CTA-time, ~2.3 ms at the BASELINE shapes, whatever the chunk size)
    takes longer than its copy (1.2 ms for 8,192 series), and a slot's next copy has to wait for that slot's
    previous solve -- with two slots the PCIe link idles half the time, with four the copies run back to back."""

    SLOTS = 4

    def __init__(self, device, chunk_paths, length, channels, hidden, dtype):
        self.device = torch.device(device)
        self.chunk_paths = chunk_paths
        self.key = (chunk_paths, length, channels, hidden, dtype)
        n = self.SLOTS
        self.streams = [torch.cuda.Stream(self.device) for _ in range(n)]
        self.x = [torch.empty(chunk_paths, length, channels, dtype=dtype, device=self.device) for _ in range(n)]
        self.z0 = [torch.empty(chunk_paths, hidden, dtype=dtype, device=self.device) for _ in range(n)]
        self.fused_builder = True              # tcde_hermite_bdiff_coeffs_series until it reports UNSUPPORTED
        self.filled = [None] * n               # staging of the two-launch builder, allocated only if that is needed
        self.coeffs = [torch.empty(chunk_paths, length - 1, 4 * channels, dtype=dtype, device=self.device)
                       for _ in range(n)]


    def filled_slot(self, slot):
        if self.filled[slot] is None:
            self.filled[slot] = torch.empty_like(self.x[slot])
        return self.filled[slot]


_series_pipelines = {}


def cdeint_from_host_series(x_host, func, z0_host, t, out_host=None, chunk_paths=8192, device=None, **kwargs):
    """The whole user pipeline from raw series kept on the host: per chunk, copy ``x`` (pinned,
    (P, L, C)) in, build the Hermite backward-difference coefficients on the device
    (``hermite_cubic_coefficients_with_backward_differences``'s kernel), solve, copy the result out.
    Moving ``x`` instead of its coefficients cuts the PCIe traffic 4x; rebuilding the coefficients
    costs ~0.5 ms per 65,536 paths on the device.  Missing values (NaN) are handled like the public builder does
    (interpolation_hermite_cubic_bdiff.py:33: linear gap filling first) -- inside the same launch
    (``tcde_hermite_bdiff_coeffs_series``: a warp fills its path in shared memory when it has gaps), so no NaN flag has to
    travel back to the host mid-pipeline."""
    device = torch.device(device if device is not None else "cuda")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    n_paths, hidden = z0_host.shape
    length, channels = x_host.shape[1], x_host.shape[2]
    n_out = t.numel()
    if out_host is None:
        out_host = torch.empty(n_paths, n_out, hidden, dtype=z0_host.dtype, pin_memory=True)
    cp = min(chunk_paths, n_paths)
    key = (device.index, cp, length, channels, hidden, x_host.dtype)
    pipe = _series_pipelines.get(key)
    if pipe is None:
        pipe = SeriesPipeline(device, cp, length, channels, hidden, x_host.dtype)
        _series_pipelines.clear()              # keep one configuration's buffers alive, not every one ever seen
        _series_pipelines[key] = pipe
    kwargs.setdefault("adjoint", False)
    code = _lib.dtype_code(x_host.dtype)
    current = torch.cuda.current_stream(device)
    for s in pipe.streams:
        s.wait_stream(current)
    t_cpu = t.detach().cpu()
    with torch.no_grad():
        for i, lo in enumerate(range(0, n_paths, cp)):
            hi = min(lo + cp, n_paths)
            slot = i % pipe.SLOTS
            with torch.cuda.stream(pipe.streams[slot]):
                n = hi - lo
                x_dev, z_dev, coeffs = pipe.x[slot][:n], pipe.z0[slot][:n], pipe.coeffs[slot][:n]
                x_dev.copy_(x_host[lo:hi], non_blocking=True)
                z_dev.copy_(z0_host[lo:hi], non_blocking=True)
                if pipe.fused_builder:
                    try:
                        _lib.call("tcde_hermite_bdiff_coeffs_series", _lib.ptr(x_dev), None, _lib.ptr(coeffs), n, length,
                                  channels, code, None, _lib.stream_of(x_dev))
                    except NotImplementedError:
                        pipe.fused_builder = False            # a path does not fit a warp's tile: two launches
                if not pipe.fused_builder:
                    filled = pipe.filled_slot(slot)[:n]
                    _lib.call("tcde_linear_fill", _lib.ptr(x_dev), None, _lib.ptr(filled), n, length, channels, code, None,
                              _lib.stream_of(x_dev))
                    _lib.call("tcde_hermite_bdiff_coeffs", _lib.ptr(filled), None, _lib.ptr(coeffs), n, length, channels,
                              code, None, _lib.stream_of(x_dev))
                out = cdeint(CubicSpline(coeffs), func, z_dev, t_cpu, **kwargs)
                out_host[lo:hi].copy_(out, non_blocking=True)
    for s in pipe.streams:
        current.wait_stream(s)
    return out_host
