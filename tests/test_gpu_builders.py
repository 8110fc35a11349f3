"""GPU parity, hot path (i): coefficient builders through the C ABI vs the reference fixtures
and the CPU oracle.  Bit-exact unless stated (the NaN-free natural cubic is within a few ulp:
its back substitution multiplies by a reciprocal instead of dividing)."""
import math
import warnings

import pytest
import torch

import torchcde_b200 as cde
from conftest import Golden, same
from oracle import cde_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close_ulps(got, want, ulps=64):
    """|got - want| <= ulps * eps * max(1, per-tensor scale); NaN must match NaN."""
    assert got.shape == want.shape and got.dtype == want.dtype
    nan_ok = torch.isnan(got) == torch.isnan(want)
    eps = torch.finfo(want.dtype).eps
    scale = max(1.0, float(want[~torch.isnan(want)].abs().max())) if want.numel() else 1.0
    diff = (got - want).abs()
    diff[torch.isnan(diff)] = 0
    return bool(nan_ok.all()) and float(diff.max() if diff.numel() else 0.0) <= ulps * eps * scale


def test_builders_against_reference_fixtures():
    g = Golden("builders")
    for i in range(g.count):
        k = "c{:03d}".format(i)
        x = g.t(k + "_in_x").to(DEV)
        t = g.t(k + "_in_t").to(DEV) if g.has(k + "_in_t") else None
        dense = not bool(torch.isnan(x).any())
        assert same(cde.linear_interpolation_coeffs(x, t).cpu(), g.t(k + "_ref_linear")), k
        assert same(cde.hermite_cubic_coefficients_with_backward_differences(x, t).cpu(), g.t(k + "_ref_hermite")), k
        assert same(cde.misc.forward_fill(x).cpu(), g.t(k + "_ref_ffill")), k
        for fn, key in ((cde.natural_cubic_coeffs, "_ref_natural_v1"), (cde.natural_cubic_spline_coeffs, "_ref_natural_v0")):
            got = fn(x, t).cpu()
            if dense:
                assert _close_ulps(got, g.t(k + key)), (k, key)
            else:
                assert same(got, g.t(k + key)), (k, key)       # per-series path: reference order, exact


def test_linear_coeffs_returns_the_input_object_when_dense():
    x = torch.randn(3, 5, 2, device=DEV)
    assert cde.linear_interpolation_coeffs(x) is x              # interpolation_linear.py:169-171


def test_rectilinear_known_answers_and_fixtures():
    g = Golden("rectilinear")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x = g.t("k_in_x").to(DEV)
        assert torch.equal(cde.linear_interpolation_coeffs(x, rectilinear=0).cpu(), g.t("k_known"))
        assert torch.equal(cde.linear_interpolation_coeffs(x[:, :, [1, 0]].contiguous(), rectilinear=1).cpu(),
                           g.t("k_ref_swapped"))
        assert torch.equal(cde.linear_interpolation_coeffs(x[0], rectilinear=0).cpu(), g.t("k_known")[0])
        x4 = torch.stack([x, x])
        assert torch.equal(cde.linear_interpolation_coeffs(x4, rectilinear=0).cpu(), torch.stack([g.t("k_known")] * 2))
        bad = x.clone()
        bad[0, 1, 0] = float("nan")
        with pytest.raises(AssertionError):
            cde.linear_interpolation_coeffs(bad, rectilinear=0)
        for i in range(g.count):
            k = "r{:02d}".format(i)
            got = cde.linear_interpolation_coeffs(g.t(k + "_in_x").to(DEV), rectilinear=int(g.z[k + "_time_index"]))
            assert same(got.cpu(), g.t(k + "_ref")), k
    with pytest.warns(UserWarning, match="begins with missing values"):
        cde.linear_interpolation_coeffs(g.t("k_in_x").to(DEV), rectilinear=0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("shape", [(257, 256, 8), (33, 100, 3), (5, 7, 1), (2, 3, 40, 5), (4, 2, 16), (3, 300, 300)])
def test_builders_against_oracle_seeded(dtype, shape):
    gen = torch.Generator().manual_seed(hash(shape) % 1000)
    x = torch.randn(shape, generator=gen, dtype=torch.float64).to(dtype)
    t = (torch.rand(shape[-2], generator=gen, dtype=torch.float64) + 0.1).cumsum(0).to(dtype)
    for tt in (None, t):
        td = None if tt is None else tt.to(DEV)
        assert same(cde.hermite_cubic_coefficients_with_backward_differences(x.to(DEV), td).cpu(),
                    O.hermite_backward_difference_coeffs(x, tt))
        assert _close_ulps(cde.natural_cubic_coeffs(x.to(DEV), td).cpu(), O.natural_cubic_coeffs(x, tt, 1))
        xn = x.clone()
        xn[torch.rand(shape, generator=gen) < 0.3] = float("nan")
        assert same(cde.linear_interpolation_coeffs(xn.to(DEV), td).cpu(), O.linear_knots(xn, tt))
        assert same(cde.hermite_cubic_coefficients_with_backward_differences(xn.to(DEV), td).cpu(),
                    O.hermite_backward_difference_coeffs(xn, tt))
        assert same(cde.misc.forward_fill(xn.to(DEV)).cpu(), O.carry_forward(xn))
        if xn[..., 0].numel() <= 4096:        # the per-series oracle is a Python loop
            assert same(cde.natural_cubic_coeffs(xn.to(DEV), td).cpu(), O.natural_cubic_coeffs(xn, tt, 1))
            assert same(cde.natural_cubic_spline_coeffs(xn.to(DEV), td).cpu(), O.natural_cubic_coeffs(xn, tt, 0))


def test_edge_cases():
    nan = float("nan")
    # all-NaN series -> zeros; a single observation -> constant; two knots
    x = torch.tensor([[[nan, 1.0], [nan, nan], [nan, nan], [nan, 4.0]],
                      [[nan, nan], [2.0, nan], [nan, nan], [nan, nan]]], device=DEV)
    for fn, ofn in ((cde.linear_interpolation_coeffs, O.linear_knots),
                    (cde.hermite_cubic_coefficients_with_backward_differences, O.hermite_backward_difference_coeffs),
                    (cde.natural_cubic_coeffs, lambda v: O.natural_cubic_coeffs(v, None, 1)),
                    (cde.natural_cubic_spline_coeffs, lambda v: O.natural_cubic_coeffs(v, None, 0))):
        assert same(fn(x).cpu(), ofn(x.cpu())), fn.__name__
    two = torch.randn(6, 2, 3, device=DEV)
    assert same(cde.natural_cubic_coeffs(two).cpu(), O.natural_cubic_coeffs(two.cpu(), None, 1))
    assert same(cde.hermite_cubic_coefficients_with_backward_differences(two).cpu(),
                O.hermite_backward_difference_coeffs(two.cpu()))
    empty = torch.zeros(0, 5, 3, device=DEV)
    assert cde.hermite_cubic_coefficients_with_backward_differences(empty).shape == (0, 4, 12)
    # non-contiguous input
    xt = torch.randn(4, 3, 9, device=DEV).transpose(-1, -2)
    assert same(cde.hermite_cubic_coefficients_with_backward_differences(xt).cpu(),
                O.hermite_backward_difference_coeffs(xt.cpu().contiguous()))


@pytest.mark.parametrize("shape", [(70, 256, 8), (9, 100, 4), (5, 33, 12), (3, 20, 64), (4, 3, 8), (2, 5, 128),
                                   (131, 17, 72), (40, 1000, 4), (1, 64, 8)])
@pytest.mark.parametrize("knots", ["unit", "uneven", "stiff"])
def test_natural_cubic_kernel_variants(shape, knots):
    """fp32 with channels % 4 == 0 takes the warp-per-path kernel (variant 0); it must agree with the CTA-per-path
    kernel (variant 2), the sequential one-thread-per-series kernel (variant 1) and the oracle.  'stiff' knots
    (spacings over three decades) make the warm-up windows long and uneven."""
    from torchcde_b200 import _lib
    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=gen, dtype=torch.float64).to(torch.float32)
    if knots == "unit":
        t = None
    elif knots == "uneven":
        t = (torch.rand(shape[-2], generator=gen, dtype=torch.float64) + 0.1).cumsum(0).to(torch.float32)
    else:
        t = (10.0 ** (3 * torch.rand(shape[-2], generator=gen, dtype=torch.float64) - 2)).cumsum(0).to(torch.float32)
    want = O.natural_cubic_coeffs(x, t, 1)
    td = None if t is None else t.to(DEV)
    try:
        for variant in (0, 2, 1):
            _lib.call("tcde_set_natural_variant", variant)
            got = cde.natural_cubic_coeffs(x.to(DEV), td).cpu()
            assert _close_ulps(got, want, ulps=256 if knots == "stiff" else 64), (variant, shape, knots)
    finally:
        _lib.call("tcde_set_natural_variant", 0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("shape", [(70, 256, 8), (9, 100, 4), (5, 33, 12), (3, 20, 32), (4, 3, 8), (6, 2, 1),
                                   (131, 17, 7), (40, 1000, 4), (3, 64, 1), (2, 2000, 16)])
@pytest.mark.parametrize("rate", [0.3, 0.9])
def test_linear_fill_kernel_variants(dtype, shape, rate):
    """Channels <= 32 take the scan kernel (variant 0); it must give the bits of the ballot kernel (variant 2), of
    the one-thread-per-series kernel (variant 1) and of the oracle -- including all-NaN series, leading / trailing
    gaps, gaps longer than a chunk and -0.0 observations next to an imputed end point."""
    from torchcde_b200 import _lib
    gen = torch.Generator().manual_seed(sum(shape) + int(rate * 10))
    x = torch.randn(shape, generator=gen, dtype=torch.float64).to(dtype)
    x[torch.rand(shape, generator=gen) < 0.05] = -0.0
    x[torch.rand(shape, generator=gen) < rate] = float("nan")
    x[0, :, 0] = float("nan")                              # one series with nothing observed
    if shape[1] > 2:
        x[-1, 0, :] = float("nan")                         # leading / trailing gaps on a whole path
        x[-1, -1, :] = float("nan")
    for t in (None, (torch.rand(shape[-2], generator=gen, dtype=torch.float64) + 0.1).cumsum(0).to(dtype)):
        want = O.linear_knots(x, t)
        td = None if t is None else t.to(DEV)
        try:
            for variant in (0, 2, 1):
                _lib.call("tcde_set_natural_variant", variant)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    got = cde.linear_interpolation_coeffs(x.to(DEV), td).cpu()
                assert same(got, want), (variant, shape, rate, t is None)
        finally:
            _lib.call("tcde_set_natural_variant", 0)


def test_hermite_unit_time_closed_form():
    """The reference's independent restatement for unit knots (test_hermite_cubic.py:5-22):
    2c = 4 (d_next - d_prev), 3d = -3 (d_next - d_prev)."""
    x = torch.randn(2, 3, 10, 6, dtype=torch.float64, device=DEV)
    got = cde.hermite_cubic_coefficients_with_backward_differences(x)
    d_next = x[..., 1:, :] - x[..., :-1, :]
    d_prev = torch.cat([d_next[..., :1, :], d_next[..., :-1, :]], dim=-2)
    want = torch.cat([x[..., :-1, :], d_prev, 4 * (d_next - d_prev), -3 * (d_next - d_prev)], dim=-1)
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_full_size_config2_properties():
    """BASELINE config 2 (batch 65536, length 256, channels 8, fp32): size-independent checks."""
    gen = torch.Generator(device=DEV).manual_seed(0)
    B, L, C = 65536, 256, 8
    x = torch.randn(B, L, C, generator=gen, device=DEV).cumsum(1) / math.sqrt(L)
    co = cde.hermite_cubic_coefficients_with_backward_differences(x)
    assert co.shape == (B, L - 1, 4 * C)
    a, b, two_c, three_d = co[..., :C], co[..., C:2 * C], co[..., 2 * C:3 * C], co[..., 3 * C:]
    assert torch.equal(a, x[:, :-1])                                           # passes through the knots
    assert bool((two_c[:, 0] == 0).all()) and bool((three_d[:, 0] == 0).all())   # first piece is linear
    end = a + (b + (0.5 * two_c + three_d / 3))                                # value at the right knot (dt = 1)
    assert torch.allclose(end, x[:, 1:], rtol=0, atol=2e-5)
    slope_end = b + two_c + three_d                                            # derivative continuity
    assert torch.allclose(slope_end[:, :-1], b[:, 1:], rtol=0, atol=2e-5)
    pick = torch.arange(0, B, 4099, device=DEV)
    assert torch.equal(co[pick].cpu(), O.hermite_backward_difference_coeffs(x[pick].cpu()))
    nat = cde.natural_cubic_coeffs(x)
    assert _close_ulps(nat[pick].cpu(), O.natural_cubic_coeffs(x[pick].cpu(), None, 1))
    n2c = nat[..., 2 * C:3 * C]
    assert float(n2c[:, 0].abs().max()) < 1e-4                                 # natural boundary: c(0) = 0
    del co, nat
    xn = x.clone()
    mask = torch.rand(B, L, C, generator=gen, device=DEV) < 0.3
    mask[:, 0] = False
    mask[:, -1] = False
    xn[mask] = float("nan")
    filled = cde.linear_interpolation_coeffs(xn)
    assert not bool(torch.isnan(filled).any())
    assert torch.equal(filled[~mask], x[~mask])                                # observations untouched
    assert torch.equal(cde.linear_interpolation_coeffs(filled), filled)        # idempotent
    assert torch.equal(filled[pick].cpu(), O.linear_knots(xn[pick].cpu()))
