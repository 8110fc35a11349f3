"""Coefficient construction -- the host side of hot path (i).

Same names, arguments, shapes, exception types and messages as the reference builders
(interpolation_linear.py:131-171, interpolation_hermite_cubic_bdiff.py:23-44,
interpolation_cubic.py:173-265, misc.py:70-126); the arithmetic runs in the CUDA kernels of
``csrc/hermite.cu``, ``natural.cu`` and ``fill.cu`` through the C ABI.  CUDA tensors only; float32 / float64.
"""
import math
import warnings

import torch

from . import _diff, _lib


# --------------------------------------------------------------------------- validation
def validate_input_path(x, t):
    """misc.py:70-100.  One vectorised monotonicity test instead of a Python loop over ``t``."""
    if not x.is_floating_point():
        raise ValueError("X must both be floating point.")

    if x.ndimension() < 2:
        raise ValueError("X must have at least two dimensions, corresponding to time and channels. It instead has "
                         "shape {}.".format(tuple(x.shape)))

    if t is None:
        t = torch.linspace(0, x.size(-2) - 1, x.size(-2), dtype=x.dtype, device=x.device)

    if not t.is_floating_point():
        raise ValueError("t must both be floating point.")
    if len(t.shape) != 1:
        raise ValueError("t must be one dimensional. It instead has shape {}.".format(tuple(t.shape)))
    if t.numel() > 0 and (bool((t[1:] <= t[:-1]).any()) or not bool(t[0] > -math.inf)):
        raise ValueError("t must be monotonically increasing.")

    if x.size(-2) != t.size(0):
        raise ValueError("The time dimension of X must equal the length of t. X has shape {} and t has shape {}, "
                         "corresponding to time dimensions of {} and {} respectively."
                         .format(tuple(x.shape), tuple(t.shape), x.size(-2), t.size(0)))

    if t.size(0) < 2:
        raise ValueError("Must have a time dimension of size at least 2. It instead has shape {}, corresponding to a "
                         "time dimension of size {}.".format(tuple(t.shape), t.size(0)))

    return t


def _differentiable(kernel, formula, x, t):
    """The builders are differentiable in the data and in the knots like the reference's (test_tricks.py:21-49):
    ``kernel`` always produces the values; when ``x`` or ``t`` require grad the result carries a backward pass that
    recomputes ``formula`` (torchcde_b200/_diff.py) with torch operators.  ``t`` may be None."""
    if t is None:
        return _diff.with_kernel_forward(lambda xx: kernel(xx, None), lambda xx: formula(xx, None), x)
    return _diff.with_kernel_forward(kernel, formula, x, t)


def _paths(x):
    """(..., L, C) -> contiguous (P, L, C) view / copy plus the batch shape."""
    batch = x.shape[:-2]
    flat = x.reshape(-1, x.size(-2), x.size(-1))
    if not flat.is_contiguous():
        flat = flat.contiguous()
    return flat, batch


def _knots_arg(t_given, x):
    """Device knots in x's dtype for the kernels, or None for the default 0..L-1."""
    if t_given is None:
        return None
    return t_given.to(device=x.device, dtype=x.dtype).contiguous()


def _flags(x):
    return torch.zeros(1, dtype=torch.int32, device=x.device)


def _has_nan(flat):
    if flat.numel() == 0:
        return False
    flags = _flags(flat)
    _lib.call("tcde_nan_flag", _lib.ptr(flat), flat.numel(), _lib.dtype_code(flat.dtype), _lib.ptr(flags),
              _lib.stream_of(flat))
    return bool(flags.item() & _lib.FLAG_NAN_SEEN)


def _fill(flat, knots, flags=None):
    out = torch.empty_like(flat)
    p, length, channels = flat.shape
    _lib.call("tcde_linear_fill", _lib.ptr(flat), _lib.ptr(knots), _lib.ptr(out), p, length, channels,
              _lib.dtype_code(flat.dtype), _lib.ptr(flags), _lib.stream_of(flat))
    return out


# --------------------------------------------------------------------------- linear / fills
def forward_fill(x, fill_index=-2):
    """misc.forward_fill (misc.py:103-126)."""
    assert isinstance(x, torch.Tensor)
    assert x.dim() >= 2
    _lib.require_cuda(x)
    _lib.dtype_code(x.dtype)
    moved = x.movedim(fill_index, -2) if fill_index not in (-2, x.dim() - 2) else x

    def kernel(src):
        with torch.cuda.device(src.device):
            flat, batch = _paths(src)
            out = torch.empty_like(flat)
            flags = _flags(flat)
            p, length, channels = flat.shape
            if flat.numel() > 0:
                _lib.call("tcde_forward_fill", _lib.ptr(flat), _lib.ptr(out), p, length, channels,
                          _lib.dtype_code(flat.dtype), _lib.ptr(flags), _lib.stream_of(flat))
        return out.view(*batch, length, channels)

    out = _diff.with_kernel_forward(kernel, _diff.forward_fill, moved)
    return out.movedim(-2, fill_index) if moved is not x else out


def _prepare_rectilinear_interpolation(data, time_index):
    """interpolation_linear.py:87-128 (same assertions).  Returns (prepared, starts_with_nan)."""
    n_channels = data.size(-1)
    assert isinstance(time_index, int), "Index of the time channel must be an integer in [0, {}]".format(n_channels - 1)
    assert 0 <= time_index < n_channels, "Time index must be in [0, {}], was given {}." \
                                         "".format(n_channels - 1, time_index)
    flat, batch = _paths(data)
    p, length, channels = flat.shape
    out = torch.empty(p, 2 * length - 1, channels, dtype=flat.dtype, device=flat.device)
    flags = _flags(flat)
    _lib.call("tcde_rectilinear_prepare", _lib.ptr(flat), _lib.ptr(out), p, length, channels, time_index,
              _lib.dtype_code(flat.dtype), _lib.ptr(flags), _lib.stream_of(flat))
    seen = int(flags.item())
    assert not (seen & _lib.FLAG_NAN_TIME), \
        "There exist nan values in the time column which is not allowed. If the " \
        "times are padded with nans after final time, a simple solution is to " \
        "forward fill the final time."
    return out.view(*batch, 2 * length - 1, channels), bool(seen & _lib.FLAG_NAN_FIRST_ROW)


def _linear_kernel(x, t, rectilinear):
    """The kernel path of ``linear_interpolation_coeffs``: returns ``x`` itself when nothing had to be done."""
    with torch.cuda.device(x.device):
        if rectilinear is not None:
            if x.ndimension() < 2:
                validate_input_path(x, t)
            x, starts_with_nan = _prepare_rectilinear_interpolation(x, rectilinear)
            if starts_with_nan:
                warnings.warn("The data `x` begins with missing values in some channels. The path will be constructed "
                              "by backward-filling the first observed value, which is not causal. Raising a warning as "
                              "the `rectilinear` argument has also been passed, which is nearly always only used when "
                              "causality is desired. If you need causality then fill in the missing value at the start "
                              "of each channel with whatever you'd like it to be. (The mean over that channel is a "
                              "common choice.)")
            t_full = validate_input_path(x, t)
            if not starts_with_nan:
                return x
            flat, batch = _paths(x)
            knots = None if t is None else _knots_arg(t_full, x)
            out = _fill(flat, knots)
            return out.view(*batch, x.size(-2), x.size(-1))
        # one pass: the fill kernel also reports whether there was anything to fill (linear.py:169-171)
        t_full = validate_input_path(x, t)
        flat, batch = _paths(x)
        if flat.numel() == 0:
            return x
        knots = None if t is None else _knots_arg(t_full, x)
        flags = _flags(flat)
        out = _fill(flat, knots, flags)
        if not flags.item() & _lib.FLAG_NAN_SEEN:
            return x
        return out.view(*batch, x.size(-2), x.size(-1))


def linear_interpolation_coeffs(x, t=None, rectilinear=None):
    """interpolation_linear.py:131-171.  Without NaNs (and without ``rectilinear``) the input
    tensor itself is returned, like the reference (:169-171)."""
    if not x.is_floating_point():
        raise ValueError("X must both be floating point.")
    if rectilinear is None:
        validate_input_path(x, t)
    _lib.require_cuda(x, t)
    _lib.dtype_code(x.dtype)
    with torch.no_grad():
        out = _linear_kernel(x.detach(), None if t is None else t.detach(), rectilinear)
    if rectilinear is None and out.data_ptr() == x.data_ptr() and out.shape == x.shape:
        return x                                   # nothing was missing: the input itself (and its autograd history)

    def formula(xx, tt):
        if rectilinear is not None:
            xx = _diff.rectilinear(xx, rectilinear)
        return _diff.linear_fill(xx, tt)

    return _differentiable(lambda *_: out, formula, x, t)


# --------------------------------------------------------------------------- Hermite
def _hermite_kernel(x, t):
    t_full = validate_input_path(x, t)
    with torch.cuda.device(x.device):
        flat, batch = _paths(x)
        p, length, channels = flat.shape
        code = _lib.dtype_code(flat.dtype)
        knots = None if t is None else _knots_arg(t_full, x)
        out = torch.empty(p, length - 1, 4 * channels, dtype=flat.dtype, device=flat.device)
        if p == 0:
            return out.view(*batch, length - 1, 4 * channels)
        stream = _lib.stream_of(flat)
        try:
            # fill (where a path has gaps) + coefficients in one launch; nothing is read back, so the call is asynchronous
            _lib.call("tcde_hermite_bdiff_coeffs_series", _lib.ptr(flat), _lib.ptr(knots), _lib.ptr(out), p, length,
                      channels, code, None, stream)
        except NotImplementedError:
            # a path too large for a warp's tile: the reference's two passes (bdiff.py:33, :36-43), still without a sync
            filled = _fill(flat, knots)
            _lib.call("tcde_hermite_bdiff_coeffs", _lib.ptr(filled), _lib.ptr(knots), _lib.ptr(out), p, length,
                      channels, code, None, stream)
    return out.view(*batch, length - 1, 4 * channels)


def hermite_cubic_coefficients_with_backward_differences(x, t=None):
    """interpolation_hermite_cubic_bdiff.py:23-44: (..., L, C) -> (..., L-1, 4C)."""
    validate_input_path(x, t)
    _lib.require_cuda(x, t)
    return _differentiable(_hermite_kernel, _diff.hermite, x, t)


# --------------------------------------------------------------------------- natural cubic
def _natural(x, t, version, name):
    validate_input_path(x, t)
    _lib.require_cuda(x, t)
    return _differentiable(lambda xx, tt: _natural_kernel(xx, tt, version), lambda xx, tt: _diff.natural(xx, tt, version),
                           x, t)


def _natural_kernel(x, t, version):
    t_full = validate_input_path(x, t)
    with torch.cuda.device(x.device):
        flat, batch = _paths(x)
        p, length, channels = flat.shape
        code = _lib.dtype_code(flat.dtype)
        knots = None if t is None else _knots_arg(t_full, x)
        out = torch.empty(p, length - 1, 4 * channels, dtype=flat.dtype, device=flat.device)
        flags = _flags(flat)
        stream = _lib.stream_of(flat)
        if p == 0:
            return out.view(*batch, length - 1, 4 * channels)
        workspace = torch.empty(4 * length + 8, dtype=flat.dtype, device=flat.device)
        try:
            _lib.call("tcde_natural_cubic_coeffs", _lib.ptr(flat), _lib.ptr(knots), _lib.ptr(out),
                      _lib.ptr(workspace), p, length, channels, code, _lib.ptr(flags), stream)
            per_series = bool(flags.item() & _lib.FLAG_NAN_SEEN)
        except NotImplementedError:
            # length x channels too large for the shared-memory kernel: the per-series kernel handles
            # any shape (and NaN-free data just as well, in the reference's exact operation order)
            per_series = True
        if per_series:
            nbytes = _lib.load().tcde_natural_cubic_missing_scratch_bytes(p, length, channels, code)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=flat.device)
            _lib.call("tcde_natural_cubic_coeffs_missing", _lib.ptr(flat), _lib.ptr(knots), _lib.ptr(out),
                      _lib.ptr(scratch), p, length, channels, version, code, stream)
    return out.view(*batch, length - 1, 4 * channels)


def natural_cubic_spline_coeffs(x, t=None):
    """Deprecated variant (ends copied, interpolation_cubic.py:193-230)."""
    return _natural(x, t, 0, "natural_cubic_spline_coeffs")


def natural_cubic_coeffs(x, t=None):
    """interpolation_cubic.py:233-265: (..., L, C) -> (..., L-1, 4C)."""
    return _natural(x, t, 1, "natural_cubic_coeffs")
