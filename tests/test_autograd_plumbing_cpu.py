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


def test_grad_paths_cpu(kernels_as_formulas):
    T.test_grad_paths()


def test_two_routes_agree_cpu(kernels_as_formulas):
    T.test_gradients_of_the_two_routes_agree()


def test_stacked_paths_cpu(kernels_as_formulas):
    T.test_stacked_paths()


def test_detach_trick_cpu(kernels_as_formulas):
    T.test_detach_trick()


def test_evaluate_and_derivative_cpu(kernels_as_formulas):
    T.test_evaluate_and_derivative_are_differentiable()
