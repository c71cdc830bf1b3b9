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


# ---------------------------------------------------------------------------------------------------------------------
# batched calibration caller: M parameter sets -> ONE ensemble run -> M efficiencies
# ---------------------------------------------------------------------------------------------------------------------
CRITERIA = {"nse": nse, "kge": kge, "rmse": rmse, "corr": corr, "bias_abs": bias_abs, "bias_rel": bias_rel,
            "std_ratio": std_ratio}


def _entries(mods) -> list[tuple]:
    return [(n, m, a, None) for n, (m, a) in mods.items()] if isinstance(mods, dict) else list(mods)


def apply_mods(land, mods, member: int):
    """NumPy twin of the parameter arithmetic of `hb_set_members`: the land bundle of ensemble member `member` — every
    listed parameter becomes `value * mul[member] + add[member]` (entries in order, masked elements only).  For tests
    and for running single members through the ordinary path."""
    import copy

    out = copy.deepcopy(land)
    out.normalise()
    zone_elem = np.repeat(np.arange(out.n_elem), np.diff(out.zone_ptr))
    for name, mul, add, mask in _entries(mods):
        mul_m = float(np.asarray(mul, float)[member]) if np.ndim(mul) else float(mul)
        add_m = float(np.asarray(add, float)[member]) if np.ndim(add) else float(add)
        if name in layout.ZP:
            row = out.zpar[layout.ZP[name]]
            sel = slice(None) if mask is None else np.asarray(mask, bool)[zone_elem]
        elif name in layout.EP:
            row = out.epar[layout.EP[name]]
            sel = slice(None) if mask is None else np.asarray(mask, bool)
        else:
            raise ValueError(f"unknown parameter {name!r}")
        row[sel] = row[sel] * mul_m + add_m
    return out


def evaluate(problem, members: int, mods, nodes, obs, *, criterion: str = "nse", engine=None, window: int | None = None,
             skip_nan: bool = False, device: int = 0) -> np.ndarray:
    """Efficiency of `members` parameter sets in ONE run: the basin of `problem` becomes an ensemble on the device
    (`Engine.set_members`, constants folded there), every window is simulated for all members at once and reduced to
    sufficient statistics on the device (`hb_node_statistics`), merged over the windows on the host.

    `nodes`: node indices of the basin to judge, `obs`: their observed series [len(nodes), nt].  Returns
    [members, len(nodes)] values of `criterion` (a key of `CRITERIA`) — what M sequential
    `CalibrationInterface.apply_values()` calls (ref: hydpy/auxs/calibtools.py:2134-2150) with the target functions of
    hydpy/auxs/statstools.py would give."""
    from .simulate import get_engine

    eng = engine if engine is not None else get_engine(device)
    nodes = np.asarray(nodes, np.int64)
    obs = np.ascontiguousarray(np.atleast_2d(np.asarray(obs, np.float64)))
    nt, nn = problem.nt, problem.river.n_node
    if obs.shape != (len(nodes), nt):
        raise ValueError(f"obs must have shape ({len(nodes)}, {nt}), not {obs.shape}")
    eng.set_land(problem.land)
    eng.set_land_states(problem.states)
    eng.set_river(problem.river)
    eng.set_river_states(problem.discharge, problem.mct_states)
    eng.select_series([])
    eng.set_members(members, mods)
    pairs = (np.arange(members)[:, None] * nn + nodes[None, :]).reshape(-1)
    rows = np.tile(np.arange(len(nodes)), members)
    window = nt if not window or window <= 0 else min(int(window), nt)
    parts = []
    for t0 in range(0, nt, window):
        t1 = min(t0 + window, nt)
        eng.set_forcing(problem.forcing[:, t0:t1, :], problem.doy[t0:t1])
        eng.set_external(problem.ext[:, t0:t1])
        eng.run(land=True, river=True)
        parts.append(eng.node_statistics(pairs, obs[:, t0:t1], rows, skip_nan=skip_nan))
    stats = combine(parts) if len(parts) > 1 else parts[0]
    eng.set_members(1)
    return CRITERIA[criterion](stats).reshape(members, len(nodes))


class CalibrationAdapter:
    """Batched counterpart of the reference's `CalibrationInterface` (ref: hydpy/auxs/calibtools.py:1181) for projects the
    engine covers: takes the SAME `Rule` objects (`Replace`, `Add`, `Multiply`; ref: `:892, :995, :1087`), but judges M
    value vectors with one ensemble run instead of M times `apply_values()` → `hp.simulate()` → target function.

        adapter = CalibrationAdapter(hp, rules, nodes=[hp.nodes.lahn_kalk], criterion="nse")
        efficiencies = adapter.perform_calibrationsteps(values)      # values [M, len(rules)] -> [M, len(nodes)]

    A rule becomes a modification of `hb_set_members`: Replace → (0, v), Add → (1, v), Multiply → (v, 0), with v scaled
    to the simulation step by the parameter's own `apply_timefactor` under the rule's `parameterstep`, and masked to
    the elements of the rule's selections.  The reference recomputes derived parameters in `update_parameters()`; the
    only one that depends on a calibratable value here is TTM = TT + DTTM, which follows `tt` automatically (possible
    when DTTM is zero or, for Replace, uniform).  Observations are the `obs` series of the nodes."""

    def __init__(self, hp, rules, nodes, criterion: str = "nse", engine=None, window: int | None = None):
        import hydpy

        from .extract import extract

        self.hp, self.rules, self.criterion, self.engine, self.window = hp, list(rules), criterion, engine, window
        self.i0, self.i1 = hydpy.pub.timegrids.simindices
        self.problem = extract(hp, self.i0, self.i1)
        names = [n.name for n in self.problem.nodes]
        self.nodes = [names.index(getattr(n, "name", n)) for n in nodes]
        self.obs = np.array([np.asarray(self.problem.nodes[i].sequences.obs.series[self.i0:self.i1], float)
                             for i in self.nodes])
        self._elem_index = {el.name: i for i, el in enumerate(self.problem.land_elements)}

    def mods(self, values: np.ndarray) -> list[tuple]:
        """The modifications of `hb_set_members` for the value vectors `values` [M, len(rules)]."""
        import hydpy

        from .extract import UnsupportedModelError

        values = np.atleast_2d(np.asarray(values, float))
        m = values.shape[0]
        land = self.problem.land
        out: list[tuple] = []
        for j, rule in enumerate(self.rules):
            name, kind = rule.parametername, type(rule).__name__
            if name not in layout.ZP and name not in layout.EP:
                raise UnsupportedModelError(f"rule `{rule.name}`: parameter `{name}` is not a calibratable constant of the engine")
            if getattr(rule, "keyword", None) is not None or getattr(rule, "adaptor", None):
                raise UnsupportedModelError(f"rule `{rule.name}`: keyword arguments and adaptors are not supported")
            mask = np.zeros(land.n_elem, bool)
            for element in rule.element2parameters:
                if element.name in self._elem_index:
                    mask[self._elem_index[element.name]] = True
            parameter = next(iter(rule))
            with hydpy.pub.options.parameterstep(rule.parameterstep):
                scaled = np.array([float(parameter.apply_timefactor(np.array(v))) for v in values[:, j]])
            one, zero = np.ones(m), np.zeros(m)
            if kind in ("Replace", "LogReplace"):
                mul, add = zero, scaled
            elif kind == "Add":
                mul, add = one, scaled
            elif kind == "Multiply":
                mul, add = values[:, j].copy(), zero
            else:
                raise UnsupportedModelError(f"rule `{rule.name}` of type {kind} is not supported")
            out.append((name, mul, add, mask))
            if name == "fc":   # the aet submodel's MaxSoilWater follows FC in `update_parameters()` (prepare_maxsoilwater)
                out.append(("maxsoilwater", mul, add, mask))
            if name == "tt":   # derived TTM = TT + DTTM (ref: hydpy/models/hland/hland_derived.py:461)
                zone_elem = np.repeat(np.arange(land.n_elem), np.diff(land.zone_ptr))
                dttm = (land.zp("ttm") - land.zp("tt"))[mask[zone_elem]]
                if kind == "Add":
                    out.append(("ttm", one, scaled, mask))
                elif np.all(dttm == 0.0):
                    out.append(("ttm", mul, add, mask))
                elif kind != "Multiply" and np.all(dttm == dttm[0]):
                    out.append(("ttm", zero, scaled + dttm[0], mask))
                else:
                    raise UnsupportedModelError(f"rule `{rule.name}`: TTM = TT + DTTM cannot follow this rule on the device")
        return out

    def perform_calibrationsteps(self, values: np.ndarray) -> np.ndarray:
        """[M, len(rules)] calibration values (not transformed) → [M, len(nodes)] efficiencies."""
        values = np.atleast_2d(np.asarray(values, float))
        return evaluate(self.problem, values.shape[0], self.mods(values), self.nodes, self.obs, criterion=self.criterion,
                        engine=self.engine, window=self.window)
