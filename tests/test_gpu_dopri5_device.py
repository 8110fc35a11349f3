"""Config 4's forward pass: dopri5 with the controller on the device (solve_dopri5.cu) against the host-driven driver
(adaptive.Dopri5: same algorithm, one host read per attempt) and against a tight fixed-step solution."""
import math

import pytest
import torch

import torchcde_b200 as cde
from torchcde_b200 import adaptive

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _problem(batch, length, seed=0, linear=False):
    gen = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(batch, length, 8, generator=gen, device=DEV).cumsum(1) / math.sqrt(length)
    z0 = torch.randn(batch, 32, generator=gen, device=DEV)
    torch.manual_seed(1)
    func = cde.LinearVectorField(32, 8).to(DEV)
    with torch.no_grad():
        X = cde.LinearInterpolation(x) if linear else cde.CubicSpline(cde.hermite_cubic_coefficients_with_backward_differences(x))
    return X, func, z0


@pytest.mark.parametrize("batch,length,linear", [(1000, 24, False), (257, 17, True), (128, 40, False)])
def test_device_controller_matches_the_host_driver(batch, length, linear, monkeypatch):
    X, func, z0 = _problem(batch, length, linear=linear)
    t = torch.tensor([0.0, 0.37 * (length - 1), length - 1.0])
    with torch.no_grad():
        dev_out = cde.cdeint(X, func, z0, t, adjoint=False)            # default method: dopri5
        dev_stats = dict(cde.cdeint.last_stats)
        assert dev_stats["device_controlled"]
        monkeypatch.setattr(adaptive, "device_dopri5_available", lambda *a: False)
        host_out = cde.cdeint(X, func, z0, t, adjoint=False)
        host_stats = dict(cde.cdeint.last_stats)
        assert not host_stats["device_controlled"]
        tight = cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options={"step_size": 1.0 / 32})
    scale = float(tight.abs().max())
    assert torch.equal(dev_out[:, 0], z0)
    # same algorithm, different rounding of the field (tensor core split vs CUDA cores): the step sequences agree up to the
    # odd borderline accept / reject, the solutions to well inside the tolerance the solver works to (rtol 1e-4)
    assert abs(dev_stats["n_accepted"] - host_stats["n_accepted"]) <= max(3, host_stats["n_accepted"] // 8), (dev_stats, host_stats)
    assert float((dev_out - host_out).abs().max()) <= 1e-2 * scale     # both are ~3e-3 * scale from the tight solution
    assert float((dev_out - tight).abs().max()) <= 2e-2 * scale, (float((dev_out - tight).abs().max()), scale)
    assert float((host_out - tight).abs().max()) <= 2e-2 * scale


def test_device_controller_decreasing_times():
    X, func, z0 = _problem(300, 12, seed=3)
    t = torch.tensor([11.0, 4.0, 0.0])
    with torch.no_grad():
        out = cde.cdeint(X, func, z0, t, adjoint=False)
        assert cde.cdeint.last_stats["device_controlled"]
        tight = cde.cdeint(X, func, z0, t, adjoint=False, method="rk4", options={"step_size": 1.0 / 32})
    scale = float(tight.abs().max())
    assert float((out - tight).abs().max()) <= 2e-2 * scale


def test_config4_full_size_forward():
    """BASELINE config 4 shapes (batch 65,536, length 256, dopri5): a sub-sample against a tight fixed-step solve, step
    counts reported, and the batch-coupled error norm means a sub-batch solved alone takes a (slightly) different path."""
    X, func, z0 = _problem(65536, 256)
    with torch.no_grad():
        out = cde.cdeint(X, func, z0, X.interval, adjoint=False)
        stats = dict(cde.cdeint.last_stats)
        assert stats["device_controlled"] and stats["n_accepted"] >= 255 // 2
        pick = torch.arange(0, 65536, 257, device=DEV)
        Xs = cde.CubicSpline(X._rows()[pick].contiguous())
        tight = cde.cdeint(Xs, func, z0[pick].contiguous(), X.interval, adjoint=False, method="rk4", options={"step_size": 1.0 / 16})
    print("config 4 forward: accepted {} rejected {} launches {}".format(stats["n_accepted"], stats["n_rejected"], stats["launches"]))
    scale = float(tight.abs().max())
    assert bool(torch.isfinite(out).all())
    assert float((out[pick] - tight).abs().max()) <= 2e-2 * scale
