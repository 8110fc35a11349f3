"""Port of the reference's test/test_tricks.py (gradient plumbing) to the device: gradients must reach the knots,
the data behind the coefficients, z0, the vector field's parameters and the output times, for rk4 and dopri5, with and
without the adjoint method; stacked CDEs must not traverse earlier graph twice; parameter gradients must not depend
on whether the output times require grad.  Same structure and assertions as the reference file (test_tricks.py:21-131)."""
import pytest
import torch

import torchcde_b200 as torchcde

pytestmark = pytest.mark.gpu
DEV = "cuda"


class _Func(torch.nn.Module):
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


def test_grad_paths():
    for method in ('rk4', 'dopri5'):
        for adjoint in (True, False):
            t = torch.linspace(0, 9, 10, device=DEV, requires_grad=True)
            path = torch.rand(1, 10, 3, device=DEV, requires_grad=True)
            coeffs = torchcde.natural_cubic_coeffs(path, t)
            cubic_spline = torchcde.CubicSpline(coeffs, t)
            z0 = torch.rand(1, 3, device=DEV, requires_grad=True)
            func = _Func(input_size=3, hidden_size=3)
            t_ = torch.tensor([0., 9.], device=DEV, requires_grad=True)

            if adjoint:
                kwargs = dict(adjoint_params=tuple(func.parameters()) + (coeffs, t))
            else:
                kwargs = {}
            z = torchcde.cdeint(X=cubic_spline, func=func, z0=z0, t=t_, adjoint=adjoint, method=method, rtol=1e-4,
                                atol=1e-6, **kwargs)
            assert z.shape == (1, 2, 3)
            assert t.grad is None
            assert path.grad is None
            assert z0.grad is None
            assert func.variable.grad is None
            assert t_.grad is None
            z[:, 1].sum().backward()
            assert isinstance(t.grad, torch.Tensor), (method, adjoint)
            assert isinstance(path.grad, torch.Tensor), (method, adjoint)
            assert isinstance(z0.grad, torch.Tensor), (method, adjoint)
            assert isinstance(func.variable.grad, torch.Tensor), (method, adjoint)
            assert isinstance(t_.grad, torch.Tensor), (method, adjoint)
            for g in (t.grad, path.grad, z0.grad, func.variable.grad, t_.grad):
                assert bool(torch.isfinite(g).all())


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


def test_stacked_paths():
    class Record(torch.autograd.Function):
        @staticmethod
        def forward(ctx, name, x):
            ctx.name = name
            return x

        @staticmethod
        def backward(ctx, x):
            if hasattr(ctx, 'been_here_before'):
                pytest.fail(ctx.name)
            ctx.been_here_before = True
            return None, x

    coeff_paths = [(torchcde.linear_interpolation_coeffs, torchcde.LinearInterpolation),
                   (torchcde.natural_cubic_coeffs, torchcde.CubicSpline)]
    for adjoint in (False, True):
        for first_coeffs, First in coeff_paths:
            for second_coeffs, Second in coeff_paths:
                first_path = torch.rand(1, 100, 2, device=DEV, requires_grad=True)
                first_coeff = first_coeffs(first_path)
                first_X = First(first_coeff)
                first_func = _Func(input_size=2, hidden_size=2)

                second_t = torch.linspace(0, 99, 20, device=DEV)
                if adjoint:
                    kwargs = dict(adjoint_params=tuple(first_func.parameters()) + (first_coeff,))
                else:
                    kwargs = {}
                second_path = torchcde.cdeint(X=first_X, func=first_func, z0=torch.rand(1, 2, device=DEV),
                                              t=second_t, adjoint=adjoint, method='rk4', options=dict(step_size=10),
                                              **kwargs)
                second_path = Record.apply('second', second_path)
                second_coeff = second_coeffs(second_path, second_t)
                second_X = Second(second_coeff, second_t)
                second_func = _Func(input_size=2, hidden_size=2)

                third_t = torch.linspace(0, 99, 10, device=DEV)
                if adjoint:
                    kwargs = dict(adjoint_params=tuple(second_func.parameters()) + (second_coeff, second_t))
                else:
                    kwargs = {}
                third_path = torchcde.cdeint(X=second_X, func=second_func, z0=torch.rand(1, 2, device=DEV),
                                             t=third_t, adjoint=adjoint, method='rk4', options=dict(step_size=10),
                                             **kwargs)
                third_path = Record.apply('third', third_path)
                assert first_func.variable.grad is None
                assert second_func.variable.grad is None
                assert first_path.grad is None
                third_path[:, -1].sum().backward()
                assert isinstance(second_func.variable.grad, torch.Tensor)
                assert isinstance(first_func.variable.grad, torch.Tensor)
                assert isinstance(first_path.grad, torch.Tensor)


def test_detach_trick():
    path = torch.rand(1, 10, 3, device=DEV)
    interp = torchcde.CubicSpline(torchcde.natural_cubic_coeffs(path))

    func = _Func(input_size=3, hidden_size=3)

    for adjoint in (True, False):
        variable_grads = []
        z0 = torch.rand(1, 3, device=DEV)
        for t_grad in (True, False):
            t_ = torch.tensor([0., 9.], device=DEV, requires_grad=t_grad)
            z = torchcde.cdeint(X=interp, z0=z0, func=func, t=t_, adjoint=adjoint, method='rk4',
                                options=dict(step_size=0.5))
            z[:, -1].sum().backward()
            variable_grads.append(func.variable.grad.clone())
            func.variable.grad.zero_()

        for elem in variable_grads[1:]:
            assert (elem == variable_grads[0]).all()


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
