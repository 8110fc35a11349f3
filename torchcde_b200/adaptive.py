"""Adaptive (dopri5) stepping and the adjoint backward pass -- this package's own drivers.

Where the reference calls ``torchdiffeq.odeint`` / ``odeint_adjoint`` with the default method
(solver.py:226-227; ``example/time_series_classification.py:83-86`` uses dopri5 + adjoint), we
run the drivers below instead.  They restate the *published* algorithms of torchdiffeq 0.2.x
(``_impl/rk_common.py`` ``RKAdaptiveStepsizeODESolver`` with the Dormand-Prince tableau,
``_impl/adjoint.py``) -- the package itself is not installable here, so, like the fixed-step
port, the stepping is "parity unpinned" against a torchdiffeq binary; an adaptive solver's
output is only defined up to its tolerances anyway, and that is what the tests check
(against tight-tolerance fp64 solutions).

Design for the GPU: time, step size and the controller live on the HOST as Python floats
(double precision, the dtype torchdiffeq uses for them); the state, the seven stage slopes and
the error estimate live on the device; each attempted step costs exactly one device->host
read (the error ratio), which is also what torchdiffeq pays (``if accept_step``).  The error
norm is the RMS over the *whole* state tensor, so a batch shares one step sequence
(SURVEY.md 8e caveat 1).  The vector field is a callable ``field(t_float, y) -> dy``; for the
linear vector field it is one fused kernel launch (``tcde_vector_field_linear``).

The backward pass of ``adjoint=True`` for the flagship linear field (fp32, hidden 32, channels 8) does not go
through autograd: a fixed-step backward solve is three launches per segment (``_fused_fixed_backward`` ->
``solver._kernel_vjp.segment``), an adaptive one gets the field and its vector-Jacobian products from one launch
per evaluation (``tcde_vector_field_linear_vjp``); the stage combinations and the error ratio of a dopri5
attempt are single launches too (``tcde_linear_combination`` / ``tcde_error_ratio_sumsq``).
"""
import ctypes
import math

import torch

from . import _lib

# Dormand-Prince 5(4), FSAL.  alpha, beta (rows), solution weights, error weights, midpoint weights
_ALPHA = (1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0)
_BETA = (
    (1 / 5,),
    (3 / 40, 9 / 40),
    (44 / 45, -56 / 15, 32 / 9),
    (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
    (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656),
    (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84),
)
_C_ERROR = (
    35 / 384 - 1951 / 21600, 0.0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
    -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0,
)
_C_MID = (
    6025192743 / 30085553152 / 2, 0.0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
    187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2,
)
_ORDER = 5


def _rms(x):
    return float(x.detach().pow(2).mean().sqrt())


def _kernels_apply(*tensors):
    """The single-launch kernels serve plain CUDA float tensors of one shape outside autograd; anything else
    (CPU tensors, states that are being differentiated with adjoint=False) keeps the torch operators."""
    first = tensors[0]
    if not (first.is_cuda and first.dtype in (torch.float32, torch.float64) and first.numel() > 0):
        return False
    grad = torch.is_grad_enabled()
    for x in tensors:
        if x.shape != first.shape or x.dtype != first.dtype or x.device != first.device or not x.is_contiguous():
            return False
        if grad and x.requires_grad:
            return False
    return True


def _term_arrays(ks, coefs):
    n = len(ks)
    return (ctypes.c_void_p * n)(*[k.data_ptr() for k in ks]), (ctypes.c_double * n)(*coefs), n


def _combine(base, ks, weights, dt, out=None):
    """base + sum_j (weights[j] * dt) * ks[j], skipping zero weights -- one launch of ``tcde_linear_combination``
    where that applies (``base`` may be None for a pure combination; ``out`` may alias ``base``)."""
    pairs = [(k, w * dt) for k, w in zip(ks, weights) if w != 0.0]
    terms = [k for k, _ in pairs]
    if terms and len(terms) <= 7 and _kernels_apply(*(terms + ([base] if base is not None else []))):
        out = torch.empty_like(terms[0]) if out is None else out
        ptrs, coefs, n = _term_arrays(terms, [c for _, c in pairs])
        with torch.cuda.device(out.device):
            _lib.call("tcde_linear_combination", _lib.ptr(out), _lib.ptr(base), ptrs, coefs, n, out.numel(),
                      _lib.dtype_code(out.dtype), _lib.stream_of(out))
        return out
    res = base if base is not None else torch.zeros_like(ks[0])
    for k, c in pairs:
        res = res + k * c
    if out is not None:
        out.copy_(res)
        return out
    return res


def _error_ratio(y0, y1, ks, weights, dt, atol, rtol, norm):
    """norm(err / (atol + rtol max(|y0|, |y1|))) with err = dt sum_j weights[j] ks[j]; for the default RMS norm one
    launch of ``tcde_error_ratio_sumsq`` plus the sum of its per-CTA partials (the one host read per attempt)."""
    pairs = [(k, w * dt) for k, w in zip(ks, weights) if w != 0.0]
    terms = [k for k, _ in pairs]
    if norm is _rms and terms and len(terms) <= 7 and _kernels_apply(y0, y1, *terms):
        n_part = _lib.load().tcde_error_ratio_partials(y0.numel())
        partials = torch.empty(n_part, dtype=torch.float64, device=y0.device)
        ptrs, coefs, n = _term_arrays(terms, [c for _, c in pairs])
        with torch.cuda.device(y0.device):
            _lib.call("tcde_error_ratio_sumsq", _lib.ptr(y0), _lib.ptr(y1), ptrs, coefs, n, float(atol), float(rtol),
                      y0.numel(), _lib.dtype_code(y0.dtype), _lib.ptr(partials), _lib.stream_of(y0))
        return float(partials.sum().div_(y0.numel()).sqrt_())
    err = _combine(None, ks, weights, dt)
    tol = atol + rtol * torch.max(y0.abs(), y1.abs())
    return norm(err / tol)


def _initial_step(field, t0, y0, f0, rtol, atol):
    """Hairer's starting step (torchdiffeq ``_select_initial_step`` with order - 1 = 4)."""
    scale = atol + y0.abs() * rtol
    d0 = _rms(y0 / scale)
    d1 = _rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    f1 = field(t0 + h0, y0 + h0 * f0)
    d2 = _rms((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / _ORDER)
    return min(100 * h0, h1)


def _next_step(dt, ratio, safety=0.9, ifactor=10.0, dfactor=0.2):
    if ratio == 0:
        return dt * ifactor
    if ratio < 1:
        dfactor = 1.0
    return dt * min(ifactor, max(safety / ratio ** (1.0 / _ORDER), dfactor))


class Dopri5:
    """One adaptive solve of dy/dt = field(t, y) over increasing output times ``t`` (floats)."""

    def __init__(self, field, y0, rtol, atol, first_step=None, max_num_steps=2 ** 31 - 1, norm=None, jump_t=None):
        self.field, self.rtol, self.atol = field, float(rtol), float(atol)
        self.max_num_steps = max_num_steps
        self.norm = norm or _rms
        self.first_step = first_step
        self.y0 = y0
        self.n_accepted = 0
        self.n_rejected = 0
        # times at which the field is discontinuous (README.md:194-200, `options=dict(jump_t=X.grid_points)`):
        # no step may straddle one, and the slope is re-evaluated just after it
        self.jump_t = sorted(float(v) for v in jump_t) if jump_t is not None else []

    def _attempt(self, t0, dt, y0, f0):
        ks = [f0]
        y1 = y0
        for alpha, beta in zip(_ALPHA, _BETA):
            ti = t0 + dt if alpha == 1.0 else t0 + alpha * dt
            y1 = _combine(y0, ks, beta, dt)
            ks.append(self.field(ti, y1))
        # the last beta row equals the solution weights (FSAL): y1 is the 5th-order solution
        return y1, ks, _error_ratio(y0, y1, ks, _C_ERROR, dt, self.atol, self.rtol, self.norm)

    def integrate(self, times, targets=None):
        """``targets``: optional 1-D tensor holding the same output times; when given, the dense output is evaluated at
        the tensor entries so that autograd sees the dependence on the requested times (adjoint=False)."""
        out = [self.y0]
        t0 = times[0]
        y0 = self.y0
        f0 = self.field(t0, y0)
        dt = self.first_step if self.first_step is not None else _initial_step(self.field, t0, y0, f0, self.rtol,
                                                                               self.atol)
        t_lo, t_hi = t0, t0
        coeff = None
        jumps = [v for v in self.jump_t if v > t0]
        for n_target, target in enumerate(times[1:], 1):
            tries = 0
            while target > t_hi:
                if tries >= self.max_num_steps:
                    raise RuntimeError("max_num_steps exceeded ({}>={})".format(tries, self.max_num_steps))
                step = dt
                on_jump = bool(jumps) and t_hi < jumps[0] < t_hi + dt
                if on_jump:
                    step = jumps[0] - t_hi
                y1, ks, ratio = self._attempt(t_hi, step, y0, f0)
                if ratio <= 1:
                    t_new = jumps[0] if on_jump else t_hi + step
                    if t_new >= target:          # the dense output is only ever evaluated inside the last step
                        y_mid = _combine(y0, ks, _C_MID, step)
                        coeff = self._fit(y0, y1, y_mid, ks[0], ks[-1], step)
                    t_lo, t_hi = t_hi, t_new
                    y0, f0 = y1, ks[-1]
                    if on_jump:
                        jumps.pop(0)
                        f0 = self.field(t_hi, y0, 1)      # the slope just AFTER the discontinuity
                    self.n_accepted += 1
                else:
                    self.n_rejected += 1
                dt = _next_step(step, ratio)
                tries += 1
            out.append(self._evaluate(coeff, t_lo, t_hi, target if targets is None else targets[n_target]))
        return torch.stack(out, dim=0)

    @staticmethod
    def _fit(y0, y1, y_mid, f0, f1, dt):
        a = 2 * dt * (f1 - f0) - 8 * (y1 + y0) + 16 * y_mid
        b = dt * (5 * f0 - 3 * f1) + 18 * y0 + 14 * y1 - 32 * y_mid
        c = dt * (f1 - 4 * f0) - 11 * y0 - 5 * y1 + 16 * y_mid
        return (y0, dt * f0, c, b, a)

    @staticmethod
    def _evaluate(coeff, t_lo, t_hi, t):
        x = (t - t_lo) / (t_hi - t_lo)
        total = coeff[0] + x * coeff[1]
        power = x
        for cf in coeff[2:]:
            power = power * x
            total = total + power * cf
        return total


def odeint_dopri5(field, y0, times, rtol, atol, options=None, targets=None):
    """``times``: increasing Python floats.  Returns (len(times), *y0.shape)."""
    options = dict(options or {})
    jump_t = options.pop("jump_t", None)
    if isinstance(jump_t, torch.Tensor):
        jump_t = jump_t.detach().cpu().tolist()
    solver = Dopri5(field, y0, rtol, atol, first_step=options.pop("first_step", None),
                    max_num_steps=options.pop("max_num_steps", 2 ** 31 - 1), jump_t=jump_t)
    if options:
        raise NotImplementedError("dopri5: unsupported options {}".format(sorted(options)))
    out = solver.integrate(list(times), targets)
    return out, solver


# ----------------------------------------------------- dopri5 with the controller on the device (linear field)
_DEVICE_CHUNK = 48          # launches enqueued between two reads of the done flag


def device_dopri5_available(z0, channels):
    return (z0.is_cuda and z0.dtype == torch.float32 and z0.size(-1) == 32 and channels == 8 and z0.numel() > 0)


def odeint_dopri5_device(control, kind, n_rows, knots, weight, bias, field, y0, times, rtol, atol, sign, first_step=None):
    """The same algorithm as ``Dopri5.integrate`` with accept / reject, the step-size rule and the dense output on the
    device (``tcde_dopri5_linear_attempts``: one launch per attempted step, no host read per attempt).  ``field`` is only
    used for the slope at the start and Hairer's initial step.  ``times``: increasing floats (already negated when
    ``sign`` < 0).  Returns ``(out (len(times), *y0.shape), stats)``."""
    lib = _lib.load()
    shape = y0.shape
    hidden = shape[-1]
    yf = y0.detach().reshape(-1, hidden).contiguous()
    n_paths = yf.size(0)
    dev = yf.device
    t0 = times[0]
    with torch.no_grad():
        f0 = field(t0, yf)
        dt = first_step if first_step is not None else _initial_step(field, t0, yf, f0, rtol, atol)
        state = torch.empty(5, n_paths, hidden, dtype=torch.float32, device=dev)
        state[0].copy_(yf)
        state[2].copy_(f0)
        grid = lib.tcde_dopri5_linear_grid(n_paths)
        partials = torch.zeros(2, grid, dtype=torch.float64, device=dev)
        ctl_host = torch.zeros(2, 24, dtype=torch.float64)
        ctl_host[1, 0], ctl_host[1, 1], ctl_host[1, 2] = t0, dt, times[-1]
        ctl_host[1, 3], ctl_host[1, 4] = rtol, atol
        ctl_host[1, 10] = 1                       # next output index
        ctl = ctl_host.to(dev)
        out = torch.empty(n_paths, len(times), hidden, dtype=torch.float32, device=dev)
        out[:, 0].copy_(yf)
        out_times = torch.tensor(times, dtype=torch.float64, device=dev)
        seq = 0
        with torch.cuda.device(dev):
            stream = _lib.stream_of(yf)
            while True:
                _lib.call("tcde_dopri5_linear_attempts", _lib.ptr(control), kind, n_rows, _lib.ptr(knots), _lib.ptr(weight),
                          _lib.ptr(bias), _lib.ptr(state), _lib.ptr(partials), _lib.ptr(ctl), _lib.ptr(out),
                          _lib.ptr(out_times), len(times), n_paths, 8, hidden, float(sign), seq, _DEVICE_CHUNK,
                          _lib.F32, stream)
                seq += _DEVICE_CHUNK
                last = ctl[(seq - 1) & 1].cpu()              # the one host read per chunk
                if last[6] != 0:
                    break
                if not math.isfinite(float(last[1])) or float(last[1]) == 0.0 or seq > 10_000_000:
                    raise RuntimeError("dopri5: step size underflow / non-finite error estimate at t = {}".format(float(last[0])))
    stats = {"n_accepted": int(last[8]), "n_rejected": int(last[9]), "launches": seq, "device_controlled": True}
    return out.movedim(1, 0).reshape(len(times), *shape), stats


# ------------------------------------------------------------------------- fixed grids (generic)
def odeint_fixed(field, y0, times, method, step_size):
    """Fixed-grid euler / midpoint / rk4 (3/8 rule) on float times -- used by the adjoint pass."""
    t0, t_end = times[0], times[-1]
    if step_size is None:
        grid = list(times)
    else:
        import math
        n = int(math.ceil((t_end - t0) / step_size + 1))
        grid = [t0 + i * step_size for i in range(n)]
        grid[-1] = t_end
    out = [y0]
    y = y0
    j = 1
    for a, b in zip(grid[:-1], grid[1:]):
        dt = b - a
        if method == "rk4":
            k1 = field(a, y)
            k2 = field(a + dt / 3, y + dt * k1 / 3)
            k3 = field(a + dt * 2 / 3, y + dt * (k2 - k1 / 3))
            k4 = field(b, y + dt * (k1 - k2 + k3))
            y1 = y + (k1 + 3 * (k2 + k3) + k4) * (dt * 0.125)
        elif method == "midpoint":
            y1 = y + dt * field(a + dt / 2, y + field(a, y) * (dt / 2))
        else:
            y1 = y + dt * field(a, y)
        while j < len(times) and b >= times[j]:
            if times[j] == b:
                out.append(y1)
            else:
                out.append(y + ((times[j] - a) / (b - a)) * (y1 - y))
            j += 1
        y = y1
    return torch.stack(out, dim=0)


# ------------------------------------------------------------------------------------- adjoint
class _Adjoint(torch.autograd.Function):
    """Continuous adjoint (torchdiffeq ``odeint_adjoint``): the forward pass keeps only the outputs;
    the backward pass integrates ``(y, a_y, a_params)`` backwards between consecutive output times
    with the same solver family, evaluating the vector field under autograd for its VJPs."""

    @staticmethod
    def forward(ctx, forward_solve, vf, times, solve_aug, fused_vjp, fixed_spec, flipped, t_tensor, y0, *params):
        ctx.vf, ctx.times, ctx.solve_aug, ctx.fused_vjp, ctx.fixed_spec = vf, times, solve_aug, fused_vjp, fixed_spec
        ctx.flipped = flipped
        ctx.t_needs_grad = t_tensor is not None and t_tensor.requires_grad
        ctx.t_meta = None if t_tensor is None else (t_tensor.dtype, t_tensor.device)
        with torch.no_grad():
            ys = forward_solve(y0)
        ctx.save_for_backward(ys, *params)
        return ys

    @staticmethod
    def backward(ctx, grad_ys):
        ys, *params = ctx.saved_tensors
        vf, times = ctx.vf, ctx.times
        params = tuple(params)
        if ctx.fused_vjp is not None and ctx.fixed_spec is not None and not ctx.t_needs_grad:
            with torch.no_grad():
                a_y, a_p = _fused_fixed_backward(ctx.fused_vjp, times, ys, grad_ys, *ctx.fixed_spec)
            return (None, None, None, None, None, None, None, None, a_y, *a_p)
        if ctx.t_needs_grad:
            return _Adjoint._backward_with_times(ctx, ys, params, grad_ys)
        if ctx.fused_vjp is not None and getattr(ctx.fused_vjp, "adaptive_spec", None) is not None:
            with torch.no_grad():
                done = _device_adaptive_backward(ctx.fused_vjp, times, ys, grad_ys)
            if done is not None:
                return (None, None, None, None, None, None, None, None, done[0], *done[1])
        shapes = [ys[0].shape, ys[0].shape] + [p.shape for p in params]
        sizes = [s.numel() for s in shapes]

        def pack(parts):
            return torch.cat([p.reshape(-1) for p in parts])

        def unpack(flat):
            outs, off = [], 0
            for shape, n in zip(shapes, sizes):
                outs.append(flat[off:off + n].view(shape))
                off += n
            return outs

        def aug_field(t, flat):
            y, a_y, *_ = unpack(flat)
            if ctx.fused_vjp is not None:
                # one launch: the field and its three vector-Jacobian products with the cotangent -a_y
                f, vjp_y, vjp_p = ctx.fused_vjp(t, y, a_y, -1.0)
                return pack([f, vjp_y] + vjp_p)
            with torch.enable_grad():
                y_ = y.detach().requires_grad_(True)
                f = vf(t, y_)
                # retain_graph: the part of the graph that is shared between calls (views of a coefficient tensor that is
                # itself an adjoint parameter) must survive, like in torchdiffeq's augmented dynamics
                grads = torch.autograd.grad(f, (y_,) + params, -a_y, allow_unused=True, retain_graph=True)
            vjp_y = grads[0] if grads[0] is not None else torch.zeros_like(y)
            vjp_p = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads[1:], params)]
            return pack([f.detach(), vjp_y] + vjp_p)

        with torch.no_grad():
            a_y = grad_ys[-1].clone()
            a_p = [torch.zeros_like(p) for p in params]
            for i in range(len(times) - 1, 0, -1):
                flat0 = pack([ys[i], a_y] + a_p)
                # integrate from times[i] back to times[i-1]: substitute s = -t
                flat1 = ctx.solve_aug(lambda s, v: -aug_field(-s, v), flat0, [-times[i], -times[i - 1]])[-1]
                _, a_y, *a_p = [x.clone() for x in unpack(flat1)]
                a_y = a_y + grad_ys[i - 1]
        return (None, None, None, None, None, None, None, None, a_y, *a_p)

    @staticmethod
    def _backward_with_times(ctx, ys, params, grad_ys):
        """The same backward solve with torchdiffeq's time gradients: the augmented state carries one more scalar,
        the integral of -a^T df/dt, and every output time t_i gets dL/dt_i = <f(t_i, y_i), dL/dy_i> (``odeint_adjoint``
        with ``t.requires_grad``).  Autograd serves the field (its time dependence is the control's dX/dt)."""
        vf, times = ctx.vf, ctx.times
        shapes = [torch.Size([]), ys[0].shape, ys[0].shape] + [p.shape for p in params]
        sizes = [max(1, s.numel()) for s in shapes]

        def pack(parts):
            return torch.cat([p.reshape(-1) for p in parts])

        def unpack(flat):
            outs, off = [], 0
            for shape, n in zip(shapes, sizes):
                outs.append(flat[off:off + n].view(shape))
                off += n
            return outs

        def aug_field(t, flat):
            _, y, a_y, *_ = unpack(flat)
            with torch.enable_grad():
                t_ = torch.tensor(t, dtype=torch.float64, device=ys.device, requires_grad=True)
                y_ = y.detach().requires_grad_(True)
                f = vf(t_, y_)
                grads = torch.autograd.grad(f, (t_, y_) + params, -a_y, allow_unused=True, retain_graph=True)
            vjp_t = grads[0].to(ys.dtype) if grads[0] is not None else torch.zeros((), dtype=ys.dtype, device=ys.device)
            vjp_y = grads[1] if grads[1] is not None else torch.zeros_like(y)
            vjp_p = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads[2:], params)]
            return pack([vjp_t, f.detach(), vjp_y] + vjp_p)

        with torch.no_grad():
            time_vjps = [None] * len(times)
            a_t = torch.zeros((), dtype=ys.dtype, device=ys.device)
            a_y = grad_ys[-1].clone()
            a_p = [torch.zeros_like(p) for p in params]
            for i in range(len(times) - 1, 0, -1):
                f_i = vf(times[i], ys[i])
                d_cur = (f_i.reshape(-1) * grad_ys[i].reshape(-1)).sum()
                a_t = a_t - d_cur
                time_vjps[i] = d_cur
                flat0 = pack([a_t, ys[i], a_y] + a_p)
                flat1 = ctx.solve_aug(lambda s, v: -aug_field(-s, v), flat0, [-times[i], -times[i - 1]])[-1]
                a_t, _, a_y, *a_p = [x.clone() for x in unpack(flat1)]
                a_y = a_y + grad_ys[i - 1]
            time_vjps[0] = a_t
            grad_t = torch.stack(time_vjps)
            if ctx.flipped:                # the solve ran in s = -t
                grad_t = -grad_t
            grad_t = grad_t.to(dtype=ctx.t_meta[0], device=ctx.t_meta[1])
        return (None, None, None, None, None, None, None, grad_t, a_y, *a_p)


def _device_adaptive_backward(stage, times, ys, grad_ys):
    """The dopri5 backward solve of ``_Adjoint`` with the controller on the device (``stage.adaptive_segment``), segment by
    segment between the output times like torchdiffeq's ``odeint_adjoint``.  Returns ``(a_y, grads)`` or None when a segment
    could not run there (the host-driven backward then starts from scratch)."""
    rtol, atol = stage.adaptive_spec
    grads = stage.new_grads()
    gw = next((g for g, r in zip(grads, stage.roles) if r == "w"), None)
    gb = next((g for g, r in zip(grads, stage.roles) if r == "b"), None)
    from .solver import cdeint
    hint = getattr(stage, "slots_hint", None) or 256
    totals = {}
    a_y = grad_ys[-1].clone()
    for i in range(len(times) - 1, 0, -1):
        try:
            a_lo = stage.adaptive_segment(times[i], times[i - 1], ys[i], a_y, rtol, atol, gw, gb, hint)
        except NotImplementedError:           # TCDE_ERR_UNSUPPORTED from a kernel: the host-driven backward takes any problem
            a_lo = None
        if a_lo is None:
            return None
        for key, val in stage.adaptive_stats.items():
            totals[key] = max(totals.get(key, 0), val) if key == "slots" else totals.get(key, 0) + val
        a_y = a_lo + grad_ys[i - 1]
    try:
        cdeint.last_adjoint_stats = dict(totals, device_controlled=True)
    except Exception:
        pass
    return a_y, grads


def _fused_fixed_backward(stage, times, ys, grad_ys, method, step_size):
    """The backward solve of ``_Adjoint`` for a fixed-step method when the fused adjoint stage kernel serves the
    field: same grid, same stage times and the same Runge-Kutta combinations as ``odeint_fixed`` on the packed
    state, but (z, a) stay one (2, ..., H) tensor that the kernel reads and writes directly, and the parameter
    gradients -- whose slopes do not depend on themselves -- are accumulated in place by the kernel with the
    stage's weight, instead of being carried through ``torch.cat`` as 8,448 extra state components."""
    import math
    grads = stage.new_grads()
    gw = next((g for g, r in zip(grads, stage.roles) if r == "w"), None)
    gb = next((g for g, r in zip(grads, stage.roles) if r == "b"), None)
    shape = ys[0].shape
    U = torch.empty((2,) + tuple(shape), dtype=ys.dtype, device=ys.device)
    V = torch.empty_like(U)
    K = [torch.empty_like(U) for _ in range(4)]
    a_y = grad_ys[-1].clone()

    def field(idx, frac, state, out, weight):
        # d(z, a)/ds = (-f, +a^T df/dz), d(grad)/ds = +a^T df/dp  (s = -t), the latter times the stage weight
        if weight == 0.0:
            stage.launch(idx, frac, state[0], state[1], out[0], out[1], None, None, -1.0, 1.0, 0.0)
        else:
            stage.launch(idx, frac, state[0], state[1], out[0], out[1], gw, gb, -1.0, 1.0, weight)

    for i in range(len(times) - 1, 0, -1):
        # whole segment in three launches (two tensor-core solves that keep their stage inputs + one contraction)
        a_lo = stage.segment(times[i], times[i - 1], ys[i], a_y, method, step_size, gw, gb)
        if a_lo is not None:
            a_y = a_lo + grad_ys[i - 1]
            continue
        s0, s_end = -times[i], -times[i - 1]
        if step_size is None:
            grid = [s0, s_end]
        else:
            n = int(math.ceil((s_end - s0) / step_size + 1))
            grid = [s0 + j * step_size for j in range(n)]
            grid[-1] = s_end
        stage_times = []
        for lo, hi in zip(grid[:-1], grid[1:]):
            ds = hi - lo
            if method == "rk4":
                stage_times += [lo, lo + ds / 3, lo + ds * 2 / 3, hi]
            elif method == "midpoint":
                stage_times += [lo, lo + ds / 2]
            else:
                stage_times += [lo]
        index, frac = stage.locate_many([-s for s in stage_times])
        U[0].copy_(ys[i])
        U[1].copy_(a_y)
        e = 0
        for lo, hi in zip(grid[:-1], grid[1:]):
            ds = hi - lo
            if method == "rk4":
                # the 3/8 rule of odeint_fixed, each stage state / the step itself as one combination launch
                field(index[e], frac[e], U, K[0], ds * 0.125)
                _combine(U, K[:1], (1.0 / 3,), ds, out=V)
                field(index[e + 1], frac[e + 1], V, K[1], ds * 0.375)
                _combine(U, K[:2], (-1.0 / 3, 1.0), ds, out=V)
                field(index[e + 2], frac[e + 2], V, K[2], ds * 0.375)
                _combine(U, K[:3], (1.0, -1.0, 1.0), ds, out=V)
                field(index[e + 3], frac[e + 3], V, K[3], ds * 0.125)
                _combine(U, K, (0.125, 0.375, 0.375, 0.125), ds, out=U)
                e += 4
            elif method == "midpoint":
                field(index[e], frac[e], U, K[0], 0.0)
                torch.add(U, K[0], alpha=ds / 2, out=V)
                field(index[e + 1], frac[e + 1], V, K[1], ds)
                U.add_(K[1], alpha=ds)
                e += 2
            else:
                field(index[e], frac[e], U, K[0], ds)
                U.add_(K[0], alpha=ds)
                e += 1
        a_y = U[1] + grad_ys[i - 1]
    return a_y, grads


def solve_with_adjoint(forward_solve, vf, times, solve_aug, y0, params, fused_vjp=None, fixed_spec=None, t_tensor=None,
                       flipped=False):
    """``forward_solve(y0) -> ys`` (time first); ``vf(t_float, y)`` differentiable in y and ``params``.

    ``fused_vjp(params)``, if given, returns ``None`` or a callable ``(t, y, a, scale) -> (f, scale * a^T df/dy,
    [scale * a^T df/dp for p in params])`` that replaces autograd in the backward solve."""
    params = tuple(p for p in params if p.requires_grad and p is not t_tensor)
    stage = fused_vjp(params) if fused_vjp is not None else None
    return _Adjoint.apply(forward_solve, vf, times, solve_aug, stage, fixed_spec, flipped, t_tensor, y0, *params)
