"""oracle/logsig_oracle.py against mathematics that does not depend on any implementation (the package it stands in for,
signatory, is not installable offline): level 1 = total increment, level 2 = Levy areas, a straight line has nothing above
level 1, Chen's identity, invariance under re-sampling, and the number of Lyndon words (Witt's formula)."""
import numpy as np

from oracle import logsig_oracle as O


def test_level_one_and_levy_area():
    rng = np.random.default_rng(0)
    path = rng.standard_normal((7, 3))
    ls = O.logsignature(path, 2)
    words = O.lyndon_words(3, 2)
    assert words == [(0,), (1,), (2,), (0, 1), (0, 2), (1, 2)]
    assert np.allclose(ls[:3], path[-1] - path[0])
    x = path - path[0]
    dx = np.diff(path, axis=0)
    mid = 0.5 * (x[1:] + x[:-1])
    for q, (i, j) in enumerate(words[3:], 3):
        area = 0.5 * np.sum(mid[:, i] * dx[:, j] - mid[:, j] * dx[:, i])       # 1/2 int (x_i dx_j - x_j dx_i), exact for linear pieces
        assert np.isclose(ls[q], area)


def test_straight_line_and_resampling():
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal(4), rng.standard_normal(4)
    line = np.stack([a + s * (b - a) for s in (0.0, 0.2, 0.7, 1.0)])
    ls = O.logsignature(line, 4)
    assert np.allclose(ls[:4], b - a) and np.allclose(ls[4:], 0.0, atol=1e-13)
    path = rng.standard_normal((5, 2))
    finer = np.concatenate([[path[0]], *[[0.5 * (path[i] + path[i + 1]), path[i + 1]] for i in range(4)]])
    assert np.allclose(O.logsignature(path, 4), O.logsignature(finer, 4))


def test_chen_identity_and_counts():
    rng = np.random.default_rng(2)
    path = rng.standard_normal((9, 2))
    whole = O.signature(path, 4)
    parts = O._mul(O.signature(path[:5], 4), O.signature(path[4:], 4), 4)
    for k in range(5):
        assert np.allclose(whole[k], parts[k])
    # Witt: number of Lyndon words of length n over c letters = (1/n) sum_{d | n} mu(d) c^(n/d)
    assert [len([w for w in O.lyndon_words(3, 4) if len(w) == n]) for n in (1, 2, 3, 4)] == [3, 3, 8, 18]
    from torchcde_b200.log_ode import lyndon_words, logsignature_channels
    for c, d in ((1, 3), (2, 5), (3, 4), (5, 2)):
        assert lyndon_words(c, d) == O.lyndon_words(c, d)
        assert logsignature_channels(c, d) == len(O.lyndon_words(c, d))


def test_window_knots_bookkeeping():
    """log_ode.py:18-40 as restated in torchcde_b200.log_ode._window_knots: the merged knot sequence stays sorted, every
    window end sits at the position it is given, end points that coincide with observation times are not duplicated, and
    the last window is clipped to the final time."""
    import math
    import random
    import torch
    from torchcde_b200.log_ode import _window_knots

    random.seed(1)
    for trial in range(300):
        n = random.randint(2, 30)
        dtype = random.choice([torch.float32, torch.float64])
        t = torch.linspace(0, n - 1, n, dtype=dtype) if trial % 3 == 0 else (torch.rand(n, dtype=torch.float64) + 0.05).cumsum(0).to(dtype)
        wl = random.choice([0.5, 1.0, 2.0, float(t[-1] - t[0]) / random.randint(1, 5), random.uniform(0.3, 4.0)])
        ends, positions, extra = _window_knots(t, wl)
        assert float(ends[0]) == float(t[0]) and float(ends[-1]) == float(t[-1])
        assert ends.numel() == int(math.ceil(float((t[-1] - t[0]) / wl))) + 1 or ends.numel() == int((((t[-1] - t[0]) / wl).ceil()).item()) + 1
        merged = torch.cat([t] + extra).sort().values if extra else t
        assert all(b > a for a, b in zip(positions[:-1], positions[1:]))
        assert merged.numel() == n + len(extra)
        for e, p in zip(ends.tolist(), positions):
            assert abs(float(merged[p]) - e) <= 1e-8 + 1e-5 * abs(e)
