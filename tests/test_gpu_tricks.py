"""Gradient plumbing on the device -- what the reference checks in test/test_tricks.py:21-131, restated: gradients must reach
the knots, the data behind the coefficients, z0, the vector field's parameters and the output times, for rk4 and dopri5, with
and without the adjoint method; stacked CDEs must not traverse an earlier graph twice; parameter gradients must not depend on
whether the output times require grad.  Plus what the reference does not check: the two routes agree in value, the builders'
and the controls' backward passes agree with finite differences.

``TCDE_TRICKS_ON_REFERENCE=1`` runs the plumbing tests of this file against the reference package on the CPU (fixed-step
methods only: the stand-in for torchdiffeq has no dopri5) -- a self-check of the test logic where no GPU exists."""
import os

import pytest
import torch

if os.environ.get("TCDE_TRICKS_ON_REFERENCE"):
    from oracle.reference_loader import load_reference
    torchcde = load_reference()
    DEV = "cpu"
    METHODS = ("rk4",)
else:
    import torchcde_b200 as torchcde
    DEV = "cuda"
    METHODS = ("rk4", "dopri5")

pytestmark = pytest.mark.gpu


class _Func(torch.nn.Module):
    """z -> sigmoid(z) broadcast over the input channels, plus one learnt offset per channel (one path, asserted)."""

    def __init__(self, input_size, hidden_size):
        super(_Func, self).__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.variable = torch.nn.Parameter(torch.rand(1, 1, input_size, device=DEV))

    def forward(self, t, z):
        assert z.shape == (1, self.hidden_size)
        out = z.sigmoid().unsqueeze(-1) + self.variable
        assert out.shape == (1, self.hidden_size, self.input_size)
        return out


@pytest.mark.parametrize("use_adjoint", [True, False])
@pytest.mark.parametrize("method", METHODS)
def test_every_input_receives_a_gradient(method, use_adjoint):
    knots = torch.linspace(0, 9, 10, device=DEV, requires_grad=True)
    series = torch.rand(1, 10, 3, device=DEV, requires_grad=True)
    coeffs = torchcde.natural_cubic_coeffs(series, knots)
    control = torchcde.CubicSpline(coeffs, knots)
    start = torch.rand(1, 3, device=DEV, requires_grad=True)
    field = _Func(input_size=3, hidden_size=3)
    out_times = torch.tensor([0., 9.], device=DEV, requires_grad=True)
    leaves = {"knots": knots, "series": series, "z0": start, "field parameter": field.variable, "output times": out_times}

    extra = {"adjoint_params": tuple(field.parameters()) + (coeffs, knots)} if use_adjoint else {}
    z = torchcde.cdeint(X=control, func=field, z0=start, t=out_times, adjoint=use_adjoint, method=method, rtol=1e-4,
                        atol=1e-6, **extra)
    assert z.shape == (1, 2, 3)
    for name, leaf in leaves.items():
        assert leaf.grad is None, name
    z[:, 1].sum().backward()
    for name, leaf in leaves.items():
        assert isinstance(leaf.grad, torch.Tensor), (name, method, use_adjoint)
        assert bool(torch.isfinite(leaf.grad).all()), (name, method, use_adjoint)


def test_gradients_of_the_two_routes_agree():
    """Beyond existence: adjoint and direct backpropagation give the same numbers (rk4, fp64), for every input."""
    grads = []
    for adjoint in (True, False):
        torch.manual_seed(3)
        t = torch.linspace(0, 9, 10, device=DEV, dtype=torch.float64).requires_grad_(True)
        path = torch.rand(1, 10, 3, device=DEV, dtype=torch.float64).requires_grad_(True)
        # the coefficients are built from a detached copy of the knots: with BOTH the coefficients and the knots they
        # were built from in adjoint_params, the adjoint method counts the knots' influence through the coefficients
        # twice (autograd.grad of the augmented dynamics already follows that path) -- in torchdiffeq just the same
        coeffs = torchcde.natural_cubic_coeffs(path, t.detach())
        X = torchcde.CubicSpline(coeffs, t)
        z0 = torch.rand(1, 3, device=DEV, dtype=torch.float64).requires_grad_(True)
        func = _Func(3, 3).double()
        t_ = torch.tensor([0., 4.5, 9.], device=DEV, dtype=torch.float64, requires_grad=True)
        kwargs = dict(adjoint_params=tuple(func.parameters()) + (coeffs, t)) if adjoint else {}
        z = torchcde.cdeint(X=X, func=func, z0=z0, t=t_, adjoint=adjoint, method='rk4', options=dict(step_size=0.05),
                            **kwargs)
        (z[:, 1].sum() + 2 * z[:, 2].sum()).backward()
        grads.append([g.clone() for g in (path.grad, z0.grad, func.variable.grad, t_.grad, t.grad)])
    for name, a, b in zip(("path", "z0", "variable", "t_", "knots"), *grads):
        if name == "t_":
            # the fixed grid is anchored at t_[0]: backpropagating through the solver attributes the sensitivity of an
            # output time that falls ON the grid to t_[0], the continuous adjoint attributes it to that time itself
            # (torchdiffeq's odeint / odeint_adjoint differ in the same way); total and final-time gradients agree
            a, b = torch.stack([a.sum(), a[-1]]), torch.stack([b.sum(), b[-1]])
        # the continuous adjoint discretises the backward ODE itself: agreement to the step error, not to rounding
        assert torch.allclose(a, b, rtol=2e-3, atol=2e-4), (name, float((a - b).abs().max()))


class _OnlyOnce(torch.autograd.Function):
    """Identity whose backward fails the test when it is entered a second time."""

    @staticmethod
    def forward(ctx, label, x):
        ctx.label = label
        return x

    @staticmethod
    def backward(ctx, grad):
        if getattr(ctx, "visited", False):
            pytest.fail("the graph behind the {} path was traversed twice".format(ctx.label))
        ctx.visited = True
        return None, grad


_KINDS = {"linear": (lambda x, t=None: torchcde.linear_interpolation_coeffs(x, t), lambda c, t=None: torchcde.LinearInterpolation(c, t)),
          "cubic": (lambda x, t=None: torchcde.natural_cubic_coeffs(x, t), lambda c, t=None: torchcde.CubicSpline(c, t))}


@pytest.mark.parametrize("use_adjoint", [False, True])
@pytest.mark.parametrize("lower", ["linear", "cubic"])
@pytest.mark.parametrize("upper", ["linear", "cubic"])
def test_stacked_cdes_backpropagate_once(use_adjoint, lower, upper):
    """A CDE driven by the solution of another CDE: one backward pass, every graph visited once, gradients down to the data."""
    def solve(control, field, times, params):
        extra = {"adjoint_params": params} if use_adjoint else {}
        return torchcde.cdeint(X=control, func=field, z0=torch.rand(1, 2, device=DEV), t=times, adjoint=use_adjoint,
                               method="rk4", options={"step_size": 10}, **extra)

    build_lower, wrap_lower = _KINDS[lower]
    build_upper, wrap_upper = _KINDS[upper]
    data = torch.rand(1, 100, 2, device=DEV, requires_grad=True)
    lower_coeffs = build_lower(data)
    lower_field = _Func(input_size=2, hidden_size=2)
    mid_times = torch.linspace(0, 99, 20, device=DEV)
    mid_path = solve(wrap_lower(lower_coeffs), lower_field, mid_times, tuple(lower_field.parameters()) + (lower_coeffs,))
    mid_path = _OnlyOnce.apply("middle", mid_path)

    upper_coeffs = build_upper(mid_path, mid_times)
    upper_field = _Func(input_size=2, hidden_size=2)
    top_times = torch.linspace(0, 99, 10, device=DEV)
    top_path = solve(wrap_upper(upper_coeffs, mid_times), upper_field, top_times,
                     tuple(upper_field.parameters()) + (upper_coeffs, mid_times))
    top_path = _OnlyOnce.apply("top", top_path)

    for leaf in (lower_field.variable, upper_field.variable, data):
        assert leaf.grad is None
    top_path[:, -1].sum().backward()
    for leaf in (upper_field.variable, lower_field.variable, data):
        assert isinstance(leaf.grad, torch.Tensor)


@pytest.mark.parametrize("use_adjoint", [True, False])
def test_parameter_gradient_does_not_depend_on_time_requiring_grad(use_adjoint):
    control = torchcde.CubicSpline(torchcde.natural_cubic_coeffs(torch.rand(1, 10, 3, device=DEV)))
    field = _Func(input_size=3, hidden_size=3)
    start = torch.rand(1, 3, device=DEV)
    seen = []
    for times_need_grad in (True, False):
        out_times = torch.tensor([0., 9.], device=DEV, requires_grad=times_need_grad)
        z = torchcde.cdeint(X=control, z0=start, func=field, t=out_times, adjoint=use_adjoint, method="rk4",
                            options={"step_size": 0.5})
        z[:, -1].sum().backward()
        seen.append(field.variable.grad.clone())
        field.variable.grad.zero_()
    assert torch.equal(seen[0], seen[1])


@pytest.mark.parametrize("nan", [0.0, 0.3])
def test_builders_backward_against_finite_differences(nan):
    """Kernel forward + _diff.py backward: directional derivative of every builder against central differences (fp64)."""
    torch.manual_seed(0)
    x = torch.randn(3, 9, 2, device=DEV, dtype=torch.float64)
    if nan:
        x = x.masked_fill(torch.rand(x.shape, device=DEV) < nan, float("nan"))
        x[:, 0] = torch.randn(3, 2, device=DEV, dtype=torch.float64)
    t = (torch.rand(9, device=DEV, dtype=torch.float64) + 0.3).cumsum(0)
    dx = torch.where(torch.isnan(x), torch.zeros_like(x), torch.randn_like(x))
    dtv = 0.05 * torch.randn_like(t)
    builders = [torchcde.hermite_cubic_coefficients_with_backward_differences, torchcde.natural_cubic_coeffs,
                torchcde.natural_cubic_spline_coeffs]
    if nan:
        builders.append(torchcde.linear_interpolation_coeffs)
    for build in builders:
        xa, ta = x.clone().requires_grad_(True), t.clone().requires_grad_(True)
        out = build(xa, ta)
        cot = torch.randn_like(out)
        out.backward(cot)
        analytic = (torch.nan_to_num(xa.grad) * dx).sum() + (ta.grad * dtv).sum()
        eps = 1e-6
        with torch.no_grad():
            numeric = ((build(x + eps * dx, t + eps * dtv) - build(x - eps * dx, t - eps * dtv)) * cot).sum() / (2 * eps)
        assert abs(float(analytic - numeric)) <= 1e-5 * max(1.0, abs(float(numeric))), (build.__name__, float(analytic), float(numeric))


def test_evaluate_and_derivative_are_differentiable():
    """X.evaluate / X.derivative carry gradients to the coefficients, the knots and the query times."""
    torch.manual_seed(1)
    x = torch.randn(2, 7, 3, device=DEV, dtype=torch.float64)
    t = (torch.rand(7, device=DEV, dtype=torch.float64) + 0.3).cumsum(0)
    q = torch.tensor([0.7, 1.9, 3.3], device=DEV, dtype=torch.float64)
    for make in (lambda c, k: torchcde.CubicSpline(torchcde.natural_cubic_coeffs(c, k), k),
                 lambda c, k: torchcde.LinearInterpolation(c, k)):
        for deriv in (False, True):
            xa, ta, qa = (v.clone().requires_grad_(True) for v in (x, t, q))
            X = make(xa, ta)
            out = X.derivative(qa) if deriv else X.evaluate(qa)
            cot = torch.randn_like(out)
            out.backward(cot)
            assert xa.grad is not None and ta.grad is not None
            if isinstance(X, torchcde.LinearInterpolation) and deriv:
                assert qa.grad is None          # a piecewise-constant derivative does not depend on the query time
                continue
            assert qa.grad is not None
            eps = 1e-6
            dq = torch.randn_like(q)
            with torch.no_grad():
                Xn = make(x, t)
                f = (lambda qq: Xn.derivative(qq)) if deriv else (lambda qq: Xn.evaluate(qq))
                numeric = ((f(q + eps * dq) - f(q - eps * dq)) * cot).sum() / (2 * eps)
            analytic = (qa.grad * dq).sum()
            assert abs(float(analytic - numeric)) <= 1e-5 * max(1.0, abs(float(numeric)))
