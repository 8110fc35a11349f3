"""The gradient plumbing of the package (KernelForward, the adjoint's time / knot gradients, the differentiable
schedule) exercised on the CPU: the CUDA kernels are replaced by the torch restatements of torchcde_b200/_diff.py
(pinned to the live reference in test_diff_formulas.py), everything around them is the product code.  The same test
bodies run against the real kernels in tests/test_gpu_tricks.py."""
import pytest
import torch

import test_gpu_tricks as T
from torchcde_b200 import _diff, _lib, coeffs, controls


@pytest.fixture
def kernels_as_formulas(monkeypatch):
    def fake_eval(control, knots, n_rows, channels, index, frac, kind, derivative):
        t = knots[index] + frac
        if kind == _lib.CONTROL_CUBIC:
            c = channels
            return _diff.cubic_eval(control[..., :c], control[..., c:2 * c], control[..., 2 * c:3 * c], control[..., 3 * c:],
                                    knots, t, index, derivative)
        return _diff.linear_eval(control, knots, t, index, derivative)

    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
    monkeypatch.setattr(coeffs, "_natural_kernel", lambda x, t, v: _diff.natural(x, t, v))
    monkeypatch.setattr(coeffs, "_hermite_kernel", lambda x, t: _diff.hermite(x, t))
    monkeypatch.setattr(coeffs, "_linear_kernel",
                        lambda x, t, r: x if not bool(torch.isnan(x).any()) and r is None else _diff.linear_fill(x, t))
    monkeypatch.setattr(controls, "_eval_kernel", fake_eval)
    monkeypatch.setattr(T, "DEV", "cpu")


@pytest.mark.parametrize("use_adjoint", [True, False])
@pytest.mark.parametrize("method", ["rk4", "dopri5"])
def test_grad_paths_cpu(kernels_as_formulas, method, use_adjoint):
    T.test_every_input_receives_a_gradient(method, use_adjoint)


def test_two_routes_agree_cpu(kernels_as_formulas):
    T.test_gradients_of_the_two_routes_agree()


@pytest.mark.parametrize("use_adjoint", [False, True])
@pytest.mark.parametrize("lower", ["linear", "cubic"])
@pytest.mark.parametrize("upper", ["linear", "cubic"])
def test_stacked_paths_cpu(kernels_as_formulas, use_adjoint, lower, upper):
    T.test_stacked_cdes_backpropagate_once(use_adjoint, lower, upper)


@pytest.mark.parametrize("use_adjoint", [True, False])
def test_detach_trick_cpu(kernels_as_formulas, use_adjoint):
    T.test_parameter_gradient_does_not_depend_on_time_requiring_grad(use_adjoint)


def test_evaluate_and_derivative_cpu(kernels_as_formulas):
    T.test_evaluate_and_derivative_are_differentiable()


def test_device_adaptive_backward_chains_segments_and_falls_back():
    """Host logic of the dopri5 backward with the controller on the device (adaptive._device_adaptive_backward): segments run
    from the last output time to the first, every output time injects its incoming gradient, statistics add up, and any
    segment that cannot run there (None, or an unsupported-shape report from a kernel) hands the whole backward to the
    host-driven path."""
    import torch
    import torchcde_b200 as cde
    from torchcde_b200 import adaptive

    calls = []

    def make_stage(mode):
        def stage():
            pass

        stage.adaptive_spec = (1e-4, 1e-6)
        stage.roles = ["w", "b"]
        stage.new_grads = lambda: [torch.zeros(3), torch.zeros(2)]
        stage.slots_hint = 32

        def adaptive_segment(t_hi, t_lo, y_hi, a_hi, rtol, atol, gw, gb, hint):
            calls.append((t_hi, t_lo, hint, rtol, atol))
            if mode == "none" and len(calls) == 2:
                return None
            if mode == "unsupported":
                raise NotImplementedError("shape")
            gw += 1.0
            gb += 2.0
            stage.adaptive_stats = {"n_accepted": 5, "n_rejected": 2, "launches": 48, "slots": hint, "flushes": 1}
            return a_hi * 2.0                      # a(t_lo) = 2 a(t_hi)

        stage.adaptive_segment = adaptive_segment
        return stage

    times = [0.0, 1.0, 3.0]
    ys = torch.zeros(3, 4, 2)
    grad_ys = torch.stack([torch.full((4, 2), 1.0), torch.full((4, 2), 10.0), torch.full((4, 2), 100.0)])
    cde.cdeint.last_adjoint_stats = None
    a_y, grads = adaptive._device_adaptive_backward(make_stage("ok"), times, ys, grad_ys)
    assert [c[:2] for c in calls] == [(3.0, 1.0), (1.0, 0.0)] and calls[0][2:] == (32, 1e-4, 1e-6)
    assert torch.equal(a_y, torch.full((4, 2), (100.0 * 2 + 10.0) * 2 + 1.0))
    assert torch.equal(grads[0], torch.full((3,), 2.0)) and torch.equal(grads[1], torch.full((2,), 4.0))
    assert cde.cdeint.last_adjoint_stats == {"n_accepted": 10, "n_rejected": 4, "launches": 96, "slots": 32, "flushes": 2,
                                             "device_controlled": True}
    calls.clear()
    assert adaptive._device_adaptive_backward(make_stage("none"), times, ys, grad_ys) is None
    calls.clear()
    assert adaptive._device_adaptive_backward(make_stage("unsupported"), times, ys, grad_ys) is None
