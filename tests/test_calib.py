"""Efficiency criteria from sufficient statistics (hydpy_b200/calib.py) against the reference's own documented values
(`hydpy/auxs/statstools.py`: nse :584-606, kge :956-973, rmse) and the window-merging rule."""
import numpy as np
import pytest

from hydpy_b200 import calib


def crit(fn, sim, obs):
    return float(fn(calib.statistics_host([sim], [obs]))[0])


def test_nse_values_of_the_reference_docstring():
    obs = [1.0, 2.0, 3.0]
    assert crit(calib.nse, [2.0, 2.0, 2.0], obs) == pytest.approx(0.0, abs=1e-15)
    assert crit(calib.nse, [0.0, 2.0, 4.0], obs) == pytest.approx(0.0, abs=1e-15)
    assert crit(calib.nse, [3.0, 2.0, 1.0], obs) == pytest.approx(-3.0)
    assert crit(calib.nse, [1.0, 2.0, 2.0], obs) == pytest.approx(0.5)
    assert crit(calib.nse, [1.0, 2.0, 3.0], obs) == pytest.approx(1.0)


def test_kge_values_of_the_reference_docstring():
    assert crit(calib.kge, [1.0, 2.0, 3.0], [1.0, 2.0, 3.0]) == pytest.approx(1.0)
    assert crit(calib.kge, [3.0, 2.0, 1.0], [1.0, 2.0, 3.0]) == pytest.approx(-1.0)
    assert crit(calib.kge, [3.0, 2.0, 1.0], [6.0, 4.0, 2.0]) == pytest.approx(0.5)
    assert crit(calib.kge, [3.0, 2.0, 1.0], [4.0, 2.0, 0.0]) == pytest.approx(0.5)
    assert crit(calib.kge, [3.0, 2.0, 1.0], [2.0, 2.0, 1.0]) == pytest.approx(0.495489, abs=5e-7)


def test_criteria_equal_the_numpy_formulas_of_the_reference():
    rng = np.random.default_rng(0)
    sim, obs = rng.gamma(2.0, 3.0, size=(6, 500)), rng.gamma(2.0, 3.0, size=(6, 500))
    st = calib.statistics_host(sim, obs)
    for i in range(6):
        s, o = sim[i], obs[i]
        assert calib.nse(st)[i] == pytest.approx(1.0 - np.sum((s - o) ** 2) / np.sum((o - np.mean(o)) ** 2), rel=1e-13)
        assert calib.rmse(st)[i] == pytest.approx(np.sqrt(np.mean((s - o) ** 2)), rel=1e-13)
        r = np.corrcoef(s, o)[0, 1]
        b = np.mean(s) / np.mean(o)
        g = (np.std(s) / np.mean(s)) / (np.std(o) / np.mean(o))
        assert calib.kge(st)[i] == pytest.approx(1.0 - ((r - 1) ** 2 + (b - 1) ** 2 + (g - 1) ** 2) ** 0.5, rel=1e-12)
        assert calib.corr(st)[i] == pytest.approx(r, rel=1e-13)


def test_windows_combine_to_the_whole_period():
    rng = np.random.default_rng(1)
    sim, obs = rng.gamma(2.0, 3.0, size=(4, 700)), rng.gamma(2.0, 3.0, size=(4, 700))
    obs[1, 100:130] = np.nan
    whole = calib.statistics_host(sim, obs, skip_nan=True)
    parts = [calib.statistics_host(sim[:, a:b], obs[:, a:b], skip_nan=True) for a, b in ((0, 240), (240, 480), (480, 700))]
    np.testing.assert_allclose(calib.combine(parts), whole, rtol=1e-12)
