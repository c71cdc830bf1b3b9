"""The drop-in for `HydPy.simulate()` (ref: hydpy/core/hydpytools.py:2769-2952) on the B200 engine.

    import hydpy_b200
    hydpy_b200.simulate(hp)          # instead of hp.simulate()
    hydpy_b200.install()             # or: make hp.simulate() itself use the engine

Contract after return, for `[i0, i1) = pub.timegrids.simindices` (seam #1):
  * `node.sequences.sim.series[i0:i1]` and `sim.value` of every node the run simulates;
  * `series[i0:i1]` of every hland_96 factor/flux/state sequence and musk_classic inflow/outflow kept in RAM;
  * final `states` (new and old) and `logs.quh`, so that a following window continues seamlessly.
Projects containing anything but hland_96(+evap_aet_hbv96/evap_pet_hbv96, rconc_uh) and musk_classic raise
`UnsupportedModelError` — there is no silent CPU fallback.
"""
from __future__ import annotations

from typing import Any, Iterable

import numpy as np

from .abi import LandStates
from .engine import Engine
from .extract import (Problem, check_reach_series, extract, requested_mct_series, requested_reach_states, requested_series,
                      write_back)

_engines: dict[int, Engine] = {}


def get_engine(device: int = 0) -> Engine:
    """The process-wide engine of a device (created on first use)."""
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]


def run_problem(problem: Problem, engine: Engine | None = None, *, series: Iterable[str] = (),
                window: int | None = None, device: int = 0, reach_states: bool = False,
                mct_series: Iterable[str] = ()) -> dict[str, Any]:
    """Run an extracted problem on the engine, optionally in time windows of `window` steps.

    Returns the keyword arguments of `extract.write_back` (states, discharge, node_rows, land_rows, reach_out,
    reach_in, series and — with `reach_states` — reach_state_series = musk_classic `states.discharge` per step; with
    `mct_series` (names of `layout.MCT_SERIES`) — mct_series = {name: [nt, n_seg]}, the 1-D musk_mct sequences per step).
    """
    eng = engine if engine is not None else get_engine(device)
    series = list(series)
    nt = problem.nt
    eng.set_land(problem.land)
    eng.set_land_states(problem.states)
    eng.set_river(problem.river)
    eng.set_river_states(problem.discharge, problem.mct_states)
    eng.select_series(series)
    eng.record_reach_states(reach_states)
    mct_series = list(mct_series) if problem.river.any_mct else []
    eng.record_mct_series(mct_series)
    mct_parts: dict[str, list[np.ndarray]] = {name: [] for name in mct_series}
    state_parts: list[np.ndarray] = []
    window = nt if not window or window <= 0 else min(int(window), nt)
    node_rows = np.zeros((problem.river.n_node, nt))
    land_rows = np.zeros((problem.land.n_elem, nt))
    reach_out = np.zeros((problem.river.n_reach, nt))
    reach_in = np.zeros((problem.river.n_reach, nt))
    out_series = {name: [] for name in series}
    t0 = 0
    while t0 < nt or (nt == 0 and t0 == 0):
        t1 = min(t0 + max(window, 1), nt)
        eng.set_forcing(problem.forcing[:, t0:t1, :], problem.doy[t0:t1])
        eng.set_external(problem.ext[:, t0:t1])
        eng.run(land=True, river=True)
        if t1 > t0:
            land_rows[:, t0:t1] = eng.get_land_outflow()
            node_rows[:, t0:t1] = eng.get_node_series()
            reach_out[:, t0:t1] = eng.get_reach_outflow()
            reach_in[:, t0:t1] = eng.get_reach_inflow()
            for name in series:
                out_series[name].append(eng.get_series(name))
            if reach_states:
                state_parts.append(eng.get_reach_state_series())
            for name in mct_series:
                mct_parts[name].append(eng.get_mct_series(name))
        t0 = t1
        if nt == 0:
            break
    states: LandStates = eng.get_land_states()
    discharge = eng.get_river_states()
    result = dict(
        states=states, discharge=discharge, node_rows=node_rows, land_rows=land_rows, reach_out=reach_out,
        reach_in=reach_in, series={n: np.concatenate(v, axis=0) for n, v in out_series.items() if v},
    )
    if reach_states:
        result["reach_state_series"] = (np.concatenate(state_parts, axis=0) if state_parts
                                        else np.zeros((0, problem.river.n_seg)))
    if problem.river.any_mct:
        result["mct_states"] = eng.get_river_mct_states()
    if mct_series:
        result["mct_series"] = {n: (np.concatenate(v, axis=0) if v else np.zeros((0, problem.river.n_seg)))
                                for n, v in mct_parts.items()}
    return result


def simulate(hp: Any, *, device: int = 0, window: int | None = None, engine: Engine | None = None) -> None:
    """Perform the simulation run of `hp` over `pub.timegrids.sim` on the B200 engine."""
    import hydpy  # the reference package owns objects, files and time grids

    i0, i1 = hydpy.pub.timegrids.simindices
    problem = extract(hp, i0, i1)
    check_reach_series(problem)
    result = run_problem(problem, engine, series=requested_series(problem), window=window, device=device,
                         reach_states=requested_reach_states(problem), mct_series=requested_mct_series(problem))
    write_back(problem, **result)


_original_simulate = None


def install(device: int = 0, window: int | None = None) -> None:
    """Monkey-patch `hydpy.HydPy.simulate` so that existing scripts use the engine unchanged."""
    global _original_simulate
    from hydpy.core import hydpytools

    if _original_simulate is None:
        _original_simulate = hydpytools.HydPy.simulate

    def _simulate(self: Any) -> None:
        simulate(self, device=device, window=window)

    _simulate.__doc__ = simulate.__doc__
    hydpytools.HydPy.simulate = _simulate  # type: ignore[method-assign]


def uninstall() -> None:
    """Undo `install`."""
    global _original_simulate
    if _original_simulate is not None:
        from hydpy.core import hydpytools

        hydpytools.HydPy.simulate = _original_simulate  # type: ignore[method-assign]
        _original_simulate = None
