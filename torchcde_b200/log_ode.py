"""``logsig_windows`` / ``logsignature_windows`` -- the log-ODE preprocessing transform (reference torchcde/log_ode.py:15-133).

Same arguments, window construction, NaN handling and return values as the reference.  The reference computes each
window's logsignature with the optional third-party package ``signatory`` (log_ode.py:1-8, :56-58) in a Python loop over
the windows; here all windows of all paths are ONE launch of ``tcde_logsignature_windows`` (csrc/logsig.cu), which restates
what ``signatory.Logsignature(depth)`` computes for a piecewise-linear path in its default "words" basis: the coefficients
of the Lyndon words (ordered by length, then lexicographically) of the logarithm of the signature in the truncated tensor
algebra.  ``signatory`` itself is not installable offline, so -- like the ODE stepping -- this is anchored on an independent
fp64 restatement (oracle/logsig_oracle.py) and on analytic facts (level 1 = increments, level 2 = Levy areas, a straight
line has no higher terms, Chen's identity across windows), not on the package's binary.
"""
import torch

from . import _lib
from .coeffs import linear_interpolation_coeffs, validate_input_path


def lyndon_words(channels, depth):
    """All Lyndon words over ``range(channels)`` of length <= depth, ordered by length, then lexicographically (Duval)."""
    words = []
    w = [-1]
    while w:
        w[-1] += 1
        if len(w) <= depth:
            words.append(tuple(w))
        m = len(w)
        while len(w) < depth:
            w.append(w[-m])
        while w and w[-1] == channels - 1:
            w.pop()
    return sorted(words, key=lambda u: (len(u), u))


def logsignature_channels(channels, depth):
    """Number of logsignature channels (``signatory.logsignature_channels``; Witt's formula summed over the levels)."""
    return len(lyndon_words(channels, depth))


def _window_knots(t, window_length):
    """The window end points (in ``t``'s dtype, the last one clipped to the final time), where each of them sits in the merged
    knot sequence, and the end points that are not already observation times -- the bookkeeping of log_ode.py:18-40 on the
    host (``t`` is one dimensional and short), as one merge of two sorted lists."""
    span = t[-1] - t[0]
    n_windows = int((span / window_length).ceil().item())
    ends = torch.linspace(t[0], t[0] + n_windows * window_length, n_windows + 1, dtype=t.dtype, device=t.device)
    ends = torch.min(ends, t.max())
    knots = t.tolist()
    rtol, atol = 1e-5, 1e-8                        # torch.allclose's defaults: when an end point counts as an observation time
    positions, extra = [], []
    k = 0
    for end, end_tensor in zip(ends.tolist(), ends):
        while True:
            same = abs(end - knots[k]) <= atol + rtol * abs(knots[k])
            if same or end <= knots[k]:
                break
            k += 1
        positions.append(k + len(extra))
        if not same:
            extra.append(end_tensor.unsqueeze(0))
    return ends, positions, extra


def _logsignature_windows(x, depth, window_length, t, _version):
    t = validate_input_path(x, t)
    _lib.require_cuda(x)
    code = _lib.dtype_code(x.dtype)
    t_host = t.detach().cpu()
    new_t, new_t_indices, new_t_unique = _window_knots(t_host, window_length)
    batch_dimensions = x.shape[:-2]
    channels = x.size(-1)

    if len(new_t_unique) > 0:                       # window ends that are not observations: NaN rows, filled linearly below
        t_merged, indices = torch.cat([t_host, *new_t_unique]).sort()
        missing = torch.full((1,), float('nan'), dtype=x.dtype, device=x.device).expand(*batch_dimensions, 1, channels)
        x = torch.cat([x, missing], dim=-2)[..., indices.clamp(0, x.size(-2)).to(x.device), :]
        t_fill = t_merged.to(device=x.device, dtype=x.dtype)
    else:
        t_fill = t.to(device=x.device, dtype=x.dtype)
    x = linear_interpolation_coeffs(x, t_fill)       # the gap-fill kernel (signatures are linear between observations anyway)

    words = lyndon_words(channels, depth)
    table = []
    for w in words:
        flat = 0
        for letter in w:
            flat = flat * channels + letter
        table += [len(w), flat]
    n_windows = len(new_t_indices) - 1
    flat_x = x.detach().reshape(-1, x.size(-2), channels).contiguous()
    with torch.cuda.device(x.device):
        window_index = torch.tensor(new_t_indices, dtype=torch.int32, device=x.device)
        word_table = torch.tensor(table, dtype=torch.int32, device=x.device)
        sigs = torch.empty(flat_x.size(0), n_windows, len(words), dtype=x.dtype, device=x.device)
        _lib.call("tcde_logsignature_windows", _lib.ptr(flat_x), flat_x.size(0), flat_x.size(1), channels,
                  _lib.ptr(window_index), n_windows, depth, _lib.ptr(word_table), len(words), _lib.ptr(sigs), code,
                  _lib.stream_of(flat_x))
    sigs = sigs.view(*batch_dimensions, n_windows, len(words))
    if _version == 0:
        widths = (new_t[1:] - new_t[:-1]).to(device=x.device, dtype=x.dtype)
        sigs = sigs * widths.unsqueeze(-1)
    elif _version != 1:
        raise RuntimeError
    first_increment = torch.zeros(*batch_dimensions, 1, len(words), dtype=x.dtype, device=x.device)
    first_increment[..., 0, :channels] = x[..., 0, :]
    logsignatures = torch.cat([first_increment, sigs], dim=-2).cumsum(dim=-2)
    if _version == 0:
        return logsignatures, new_t.to(x.device)
    return logsignatures


def logsignature_windows(x, depth, window_length, t=None):
    """DEPRECATED variant kept for backward compatibility (log_ode.py:80-107): returns ``(values, times)`` with each window's
    logsignature scaled by the window's width."""
    return _logsignature_windows(x, depth, window_length, t, _version=0)


def logsig_windows(x, depth, window_length, t=None):
    """Logsignatures over windows, as in the log-ODE method (log_ode.py:110-133): ``(..., length, channels)`` ->
    ``(..., 1 + number of windows, logsignature_channels(channels, depth))``; the corresponding times are 0, 1, 2, ..."""
    return _logsignature_windows(x, depth, window_length, t, _version=1)
