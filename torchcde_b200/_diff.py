"""Differentiable restatements of the coefficient builders and of spline evaluation -- BACKWARD PASSES ONLY.

The forward pass of every public function is a CUDA kernel (``coeffs.py`` / ``controls.py``).  The reference's
builders are differentiable with respect to the data ``x`` *and* the knots ``t`` because they are chains of torch
operators (test/test_tricks.py:21-49 backpropagates into both).  To offer the same without writing -- and
validating -- a hand-made adjoint for every builder and every argument, the ``torch.autograd.Function``s below run
the kernel in ``forward`` and, only when a gradient is actually requested, recompute the same mathematics here with
torch operators under ``enable_grad`` and pull the cotangent back through it.  Nothing in this file runs on a
forward pass; it is vectorised over the batch (the reference recurses per scalar series in Python,
interpolation_linear.py:74-84, interpolation_cubic.py:56-75), so a backward pass costs a few dozen launches
independent of the batch size.

Formulas restated (file:line of the reference): linear gap filling interpolation_linear.py:13-71, Hermite backward
differences interpolation_hermite_cubic_bdiff.py:5-20, natural cubic interpolation_cubic.py:7-53 (+ :78-167 for
missing values), forward fill misc.py:103-126, rectilinear interpolation_linear.py:87-128, spline evaluation
interpolation_cubic.py:324-336 / interpolation_linear.py:205-225.
"""
import torch


# ----------------------------------------------------------------------------------- helpers
def _default_knots(x):
    return torch.linspace(0, x.size(-2) - 1, x.size(-2), dtype=x.dtype, device=x.device)


def _valid_neighbours(valid):
    """valid: (..., L, C) bool.  For every position the index of the nearest valid position at or before it
    (``prev``, -1 if none) and at or after it (``nxt``, L if none)."""
    length = valid.size(-2)
    pos = torch.arange(length, device=valid.device).view(*([1] * (valid.dim() - 2)), length, 1)
    prev = torch.where(valid, pos, torch.full_like(pos, -1)).cummax(dim=-2).values
    nxt = torch.where(valid, pos, torch.full_like(pos, length)).flip(-2).cummin(dim=-2).values.flip(-2)
    return prev.expand_as(valid), nxt.expand_as(valid)


def linear_fill(x, t):
    """interpolation_linear.py:13-71, every scalar series at once: NaNs between two observations are interpolated
    linearly in time, leading / trailing NaNs take the first / last observation, an all-NaN series becomes zeros."""
    if t is None:
        t = _default_knots(x)
    valid = ~torch.isnan(x)
    length = x.size(-2)
    prev, nxt = _valid_neighbours(valid)
    has_prev, has_next = prev >= 0, nxt < length
    clean = torch.where(valid, x, torch.zeros_like(x))
    p = prev.clamp(min=0)
    n = nxt.clamp(max=length - 1)
    xp = clean.gather(-2, p)
    xn = clean.gather(-2, n)
    tp, tn = t[p], t[n]
    tt = t.view(*([1] * (x.dim() - 2)), length, 1)
    gap = torch.where(n > p, tn - tp, torch.ones_like(tn))
    ratio = (tt - tp) / gap
    inside = xp + ratio * (xn - xp)
    out = torch.where(has_prev & has_next, inside, torch.where(has_prev, xp, xn))
    out = torch.where(valid, x, out)
    return torch.where(has_prev | has_next, out, torch.zeros_like(out))


def forward_fill(x):
    """misc.py:103-126: hold the last observation; leading NaNs stay NaN."""
    valid = ~torch.isnan(x)
    prev, _ = _valid_neighbours(valid)
    clean = torch.where(valid, x, torch.zeros_like(x))
    held = clean.gather(-2, prev.clamp(min=0))
    return torch.where(prev >= 0, held, torch.full_like(held, float("nan")))


def rectilinear(x, time_index):
    """interpolation_linear.py:87-128: forward fill, repeat every row twice, lag the time channel by one."""
    filled = forward_fill(x)
    doubled = filled.repeat_interleave(2, dim=-2)
    lagged = torch.cat([doubled[..., 1:, time_index], doubled[..., -1:, time_index]], dim=-1)
    doubled = torch.cat([doubled[..., :time_index], lagged.unsqueeze(-1), doubled[..., time_index + 1:]], dim=-1)
    return doubled[..., :-1, :]


def hermite(x, t):
    """interpolation_hermite_cubic_bdiff.py:5-44 (after the linear fill of :33)."""
    if t is None:
        t = _default_knots(x)
    x = linear_fill(x, t)
    dt = (t[1:] - t[:-1]).unsqueeze(-1)
    derivs = (x[..., 1:, :] - x[..., :-1, :]) / dt
    derivs_prev = torch.cat([derivs[..., :1, :], derivs[..., :-1, :]], dim=-2)
    a = x[..., :-1, :]
    b = derivs_prev
    two_c = 2 * (3 * (derivs - derivs_prev) - derivs + derivs_prev) / dt
    three_d = (derivs - derivs_prev) / dt ** 2 - two_c / dt
    return torch.cat([a, b, two_c, three_d], dim=-1)


def _piece_coeffs(x0, x1, k0, k1, dt):
    """Closed-form cubic on one piece from end values and end derivatives (interpolation_cubic.py:44-51)."""
    rdt = dt.reciprocal()
    six = 6 * (x1 - x0)
    two_c = (six * rdt - 4 * k0 - 2 * k1) * rdt
    three_d = (-six * rdt + 3 * (k0 + k1)) * rdt ** 2
    return x0, k0, two_c, three_d


def natural(x, t, version=1, chunk=2048):
    """interpolation_cubic.py:7-53 and :78-167: the knot derivatives of the natural cubic spline solve a tridiagonal
    system.  Without NaNs the matrix depends on the (shared, 1-D) knots only and is inverted once; with NaNs every
    scalar series has its own knots: the system is assembled densely per series (``chunk`` series at a time) with
    identity rows for the missing knots, and every piece is then re-expanded around each original knot it covers."""
    if t is None:
        t = _default_knots(x)
    length, channels = x.size(-2), x.size(-1)
    if not bool(torch.isnan(x).any()):
        if length == 2:
            a = x[..., :1, :]
            b = (x[..., 1:, :] - x[..., :1, :]) / (t[1:] - t[:1]).unsqueeze(-1)
            zero = torch.zeros_like(a)
            return torch.cat([a, b, zero, zero], dim=-1)
        r = (t[1:] - t[:-1]).reciprocal()
        diag = torch.zeros(length, dtype=x.dtype, device=x.device)
        diag = 2 * (torch.cat([r, r.new_zeros(1)]) + torch.cat([r.new_zeros(1), r]))
        A = torch.diag(diag) + torch.diag(r, 1) + torch.diag(r, -1)
        scaled = 3 * (x[..., 1:, :] - x[..., :-1, :]) * (r ** 2).unsqueeze(-1)
        zero_row = torch.zeros_like(scaled[..., :1, :])
        rhs = torch.cat([scaled, zero_row], dim=-2) + torch.cat([zero_row, scaled], dim=-2)      # (..., L, C)
        k = torch.einsum("ij,...jc->...ic", torch.linalg.inv(A), rhs)
        dt = (t[1:] - t[:-1]).unsqueeze(-1)
        a, b, two_c, three_d = _piece_coeffs(x[..., :-1, :], x[..., 1:, :], k[..., :-1, :], k[..., 1:, :], dt)
        return torch.cat([a, b, two_c, three_d], dim=-1)

    # ---- missing values: one scalar series per row -------------------------------------------------------------
    batch = x.shape[:-2]
    series = x.movedim(-1, -2).reshape(-1, length)                      # (N, L)
    outs = []
    for lo in range(0, series.size(0), chunk):
        outs.append(_natural_missing(series[lo:lo + chunk], t, version))
    rows = torch.cat(outs, dim=0)                                        # (N, L-1, 4)
    rows = rows.view(*batch, channels, length - 1, 4).movedim(-3, -1)     # (..., L-1, 4, C)
    return rows.reshape(*batch, length - 1, 4 * channels)


def _natural_missing(s, t, version):
    n_series, length = s.shape
    valid = ~torch.isnan(s)
    any_valid = valid.any(dim=1, keepdim=True)
    pos = torch.arange(length, device=s.device).unsqueeze(0)
    prev, nxt = _valid_neighbours(valid.unsqueeze(-1))
    prev, nxt = prev.squeeze(-1), nxt.squeeze(-1)
    clean = torch.where(valid, s, torch.zeros_like(s))
    first = nxt[:, :1].clamp(max=length - 1)            # first valid index per series
    last = prev[:, -1:].clamp(min=0)
    if version == 0:
        # the first / last observation is imputed AT the first / last knot only (interpolation_cubic.py:101-118)
        fill_first = (~valid[:, :1]) & any_valid
        fill_last = (~valid[:, -1:]) & any_valid
        v0 = torch.where(fill_first, clean.gather(1, first), clean[:, :1])
        v1 = torch.where(fill_last, clean.gather(1, last), clean[:, -1:])
        clean = torch.cat([v0, clean[:, 1:-1], v1], dim=1)
        valid = torch.cat([valid[:, :1] | fill_first, valid[:, 1:-1], valid[:, -1:] | fill_last], dim=1)
    else:
        # filled backward / forward from the first / last observation (interpolation_cubic.py:119-131)
        before = (pos < first) & any_valid
        after = (pos > last) & any_valid
        clean = torch.where(before, clean.gather(1, first).expand_as(clean), clean)
        clean = torch.where(after, clean.gather(1, last).expand_as(clean), clean)
        valid = valid | before | after
    prev, nxt = _valid_neighbours(valid.unsqueeze(-1))
    prev, nxt = prev.squeeze(-1), nxt.squeeze(-1)

    # neighbours of a VALID knot among the valid knots
    shifted_prev = torch.cat([torch.full_like(prev[:, :1], -1), prev[:, :-1]], dim=1)       # last valid < i
    shifted_next = torch.cat([nxt[:, 1:], torch.full_like(nxt[:, :1], length)], dim=1)      # first valid > i
    has_l = valid & (shifted_prev >= 0)
    has_r = valid & (shifted_next < length)
    l_idx = shifted_prev.clamp(min=0)
    r_idx = shifted_next.clamp(max=length - 1)
    tt = t.unsqueeze(0).expand(n_series, length)
    one = torch.ones_like(tt)
    rl = torch.where(has_l, (tt - t[l_idx]), one).reciprocal() * has_l
    rr = torch.where(has_r, (t[r_idx] - tt), one).reciprocal() * has_r
    diag = torch.where(valid, 2 * (rl + rr), one)
    # a series with a single valid knot cannot happen after the imputation (length >= 2); an all-NaN series is all identity
    A = torch.diag_embed(diag)
    A = A.scatter_add(2, l_idx.unsqueeze(-1), rl.unsqueeze(-1))
    A = A.scatter_add(2, r_idx.unsqueeze(-1), rr.unsqueeze(-1))
    xl = clean.gather(1, l_idx)
    xr = clean.gather(1, r_idx)
    rhs = 3 * (clean - xl) * rl ** 2 + 3 * (xr - clean) * rr ** 2
    k = torch.linalg.solve(A, rhs.unsqueeze(-1)).squeeze(-1)                                  # (N, L); 0 at missing knots

    # the piece covering original interval j runs from P = last valid <= j to Q = first valid >= j + 1
    P = prev[:, :-1].clamp(min=0)
    Q = nxt[:, 1:].clamp(max=length - 1)
    dt = torch.where(Q > P, t[Q] - t[P], torch.ones_like(t[Q]))
    a, b, two_c, three_d = _piece_coeffs(clean.gather(1, P), clean.gather(1, Q), k.gather(1, P), k.gather(1, Q), dt)
    offset = t[P] - tt[:, :-1]
    a_inner = (0.5 * two_c - three_d * offset / 3) * offset
    a_j = a + (a_inner - b) * offset
    b_j = b + (three_d * offset - two_c) * offset
    two_c_j = two_c - 2 * three_d * offset
    out = torch.stack([a_j, b_j, two_c_j, three_d], dim=-1)
    return torch.where(any_valid.unsqueeze(-1), out, torch.zeros_like(out))


# ------------------------------------------------------------------------------ spline evaluation
def cubic_eval(a, b, two_c, three_d, knots, t, index, derivative):
    """interpolation_cubic.py:324-336 with the interval ``index`` (integer tensor, shape of ``t``) already known."""
    frac = (t - knots[index]).unsqueeze(-1)
    pick = lambda buf: buf.index_select(-2, index.reshape(-1)).view(*buf.shape[:-2], *index.shape, buf.size(-1))  # noqa: E731
    if derivative:
        inner = pick(two_c) + pick(three_d) * frac
        return pick(b) + inner * frac
    inner = 0.5 * pick(two_c) + pick(three_d) * frac / 3
    inner = pick(b) + inner * frac
    return pick(a) + inner * frac


def linear_eval(coeffs, knots, t, index, derivative):
    """interpolation_linear.py:205-225."""
    pick = lambda buf, idx: buf.index_select(-2, idx.reshape(-1)).view(*buf.shape[:-2], *idx.shape, buf.size(-1))  # noqa: E731
    prev, nxt = pick(coeffs, index), pick(coeffs, index + 1)
    dt = (knots[index + 1] - knots[index]).unsqueeze(-1)
    if derivative:
        return (nxt - prev) / dt
    frac = (t - knots[index]).unsqueeze(-1)
    return prev + frac * ((nxt - prev) / dt)


# -------------------------------------------------------------------------- forward = kernel, backward = the above
class KernelForward(torch.autograd.Function):
    """``forward``: ``kernel(*tensors)`` (no graph).  ``backward``: recompute ``formula(*tensors)`` with torch
    operators under autograd for the inputs that need a gradient and pull the cotangent back through it."""

    @staticmethod
    def forward(ctx, kernel, formula, *tensors):
        ctx.formula = formula
        ctx.save_for_backward(*[x for x in tensors if isinstance(x, torch.Tensor)])
        ctx.layout = [isinstance(x, torch.Tensor) for x in tensors]
        ctx.others = [None if isinstance(x, torch.Tensor) else x for x in tensors]
        with torch.no_grad():
            return kernel(*[x.detach() if isinstance(x, torch.Tensor) else x for x in tensors])

    @staticmethod
    def backward(ctx, grad_out):
        saved = list(ctx.saved_tensors)
        args, wanted = [], []
        for i, is_tensor in enumerate(ctx.layout):
            if not is_tensor:
                args.append(ctx.others[i])
                continue
            x = saved.pop(0)
            if ctx.needs_input_grad[2 + i] and x.is_floating_point():
                x = x.detach().requires_grad_(True)
                wanted.append((i, x))
            args.append(x)
        grads = [None] * len(ctx.layout)
        if wanted:
            with torch.enable_grad():
                out = ctx.formula(*args)
                got = torch.autograd.grad(out, [x for _, x in wanted], grad_out, allow_unused=True)
            for (i, _), g in zip(wanted, got):
                grads[i] = g
        return (None, None, *grads)


def with_kernel_forward(kernel, formula, *tensors):
    """Route through ``KernelForward`` only when a gradient can be asked for; otherwise just run the kernel."""
    if torch.is_grad_enabled() and any(isinstance(x, torch.Tensor) and x.requires_grad for x in tensors):
        return KernelForward.apply(kernel, formula, *tensors)
    return kernel(*tensors)
