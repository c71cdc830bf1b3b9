"""Efficiency criteria of calibration ensembles from device-reduced statistics.

`Engine.node_statistics` returns, per (node, observation series) pair, the sufficient statistics `layout.STATS`; the
functions below turn them into the criteria of the reference's `hydpy/auxs/statstools.py` — vectorised over pairs, so
that an ensemble of M parameter sets is judged with one call instead of M `nse(node=...)` calls on host series
(the pattern of `CalibrationInterface.apply_values` + target function, ref: hydpy/auxs/calibtools.py:2134-2190).
"""
from __future__ import annotations

import numpy as np

from . import layout

_I = {name: i for i, name in enumerate(layout.STATS)}


def _c(stats: np.ndarray, name: str) -> np.ndarray:
    return np.asarray(stats, dtype=np.float64)[..., _I[name]]


def nse(stats: np.ndarray) -> np.ndarray:
    """1 - Σ(sim-obs)² / Σ(obs-mean(obs))²  (ref: statstools.py:574-616)."""
    return 1.0 - _c(stats, "ss_err") / _c(stats, "ss_obs")


def rmse(stats: np.ndarray) -> np.ndarray:
    """sqrt(mean((sim-obs)²))  (ref: statstools.py:529-553)."""
    return np.sqrt(_c(stats, "ss_err") / _c(stats, "n"))


def corr(stats: np.ndarray) -> np.ndarray:
    """Pearson correlation, `numpy.corrcoef(sim, obs)[0, 1]`."""
    return _c(stats, "sp") / np.sqrt(_c(stats, "ss_sim") * _c(stats, "ss_obs"))


def bias_abs(stats: np.ndarray) -> np.ndarray:
    """mean(sim - obs)."""
    return _c(stats, "mean_sim") - _c(stats, "mean_obs")


def bias_rel(stats: np.ndarray) -> np.ndarray:
    """mean(sim) / mean(obs) - 1."""
    return _c(stats, "mean_sim") / _c(stats, "mean_obs") - 1.0


def std_ratio(stats: np.ndarray) -> np.ndarray:
    """std(sim) / std(obs) - 1."""
    return np.sqrt(_c(stats, "ss_sim") / _c(stats, "ss_obs")) - 1.0


def kge(stats: np.ndarray) -> np.ndarray:
    """Kling-Gupta efficiency after Kling et al. (2012)  (ref: statstools.py:944-987)."""
    n = _c(stats, "n")
    r = corr(stats)
    m_sim, m_obs = _c(stats, "mean_sim"), _c(stats, "mean_obs")
    s_sim, s_obs = np.sqrt(_c(stats, "ss_sim") / n), np.sqrt(_c(stats, "ss_obs") / n)
    b = m_sim / m_obs
    g = (s_sim / m_sim) / (s_obs / m_obs)
    return 1.0 - ((r - 1.0) ** 2 + (b - 1.0) ** 2 + (g - 1.0) ** 2) ** 0.5


def combine(parts: list[np.ndarray]) -> np.ndarray:
    """Statistics of consecutive windows merged into the statistics of the whole period (pairwise update of centred
    sums after Chan et al.), so that criteria can be evaluated over a calibration period simulated in windows."""
    acc = np.array(parts[0], dtype=np.float64, copy=True)
    for part in parts[1:]:
        b = np.asarray(part, dtype=np.float64)
        na, nb = acc[..., _I["n"]], b[..., _I["n"]]
        n = na + nb
        ds = b[..., _I["mean_sim"]] - acc[..., _I["mean_sim"]]
        do = b[..., _I["mean_obs"]] - acc[..., _I["mean_obs"]]
        w = na * nb / n
        out = acc.copy()
        out[..., _I["n"]] = n
        out[..., _I["mean_sim"]] = acc[..., _I["mean_sim"]] + ds * nb / n
        out[..., _I["mean_obs"]] = acc[..., _I["mean_obs"]] + do * nb / n
        out[..., _I["ss_err"]] = acc[..., _I["ss_err"]] + b[..., _I["ss_err"]]
        out[..., _I["ss_sim"]] = acc[..., _I["ss_sim"]] + b[..., _I["ss_sim"]] + ds * ds * w
        out[..., _I["ss_obs"]] = acc[..., _I["ss_obs"]] + b[..., _I["ss_obs"]] + do * do * w
        out[..., _I["sp"]] = acc[..., _I["sp"]] + b[..., _I["sp"]] + ds * do * w
        acc = out
    return acc


def statistics_host(sim: np.ndarray, obs: np.ndarray, skip_nan: bool = False) -> np.ndarray:
    """The same statistics from host arrays ([n_pair, nt] each) — used by the tests as the NumPy side of the comparison."""
    sim, obs = np.atleast_2d(np.asarray(sim, float)), np.atleast_2d(np.asarray(obs, float))
    out = np.zeros((sim.shape[0], len(layout.STATS)))
    for i, (s, o) in enumerate(zip(sim, obs)):
        if skip_nan:
            keep = ~(np.isnan(s) | np.isnan(o))
            s, o = s[keep], o[keep]
        ms, mo = np.mean(s), np.mean(o)
        out[i] = (len(s), ms, mo, np.sum((s - o) ** 2), np.sum((s - ms) ** 2), np.sum((o - mo) ** 2),
                  np.sum((s - ms) * (o - mo)))
    return out
