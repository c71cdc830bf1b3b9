"""Worker of tests/test_dropin_reference.py (own process: HydPy keeps global state in `hydpy.pub`).

The drop-in `hydpy_b200.simulate(hp)` on REAL objects of the unmodified reference (built by oracle/build_ref.py in
this container): extract → run → write_back must leave `hp` exactly as the reference's own `hp.simulate()` leaves it.
There is no GPU here, so the engine call in the middle is replaced by the CPU oracle (bit-exact with the reference),
which makes the comparison a bit-for-bit check of the host-side layers."""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import build_ref, oracle  # noqa: E402

SRC, STUBS = build_ref.scratch_paths()
sys.path[:0] = [STUBS, SRC]

import hydpy  # noqa: E402
from hydpy import pub  # noqa: E402
from hydpy.core import testtools  # noqa: E402

import hydpy_b200  # noqa: E402

simulate_module = sys.modules["hydpy_b200.simulate"]   # (the package exports the function under the same name)


def oracle_run_problem(problem, engine=None, *, series=(), window=None, device=0, reach_states=False):
    """Same contract as hydpy_b200.run_problem, executed by the CPU oracle."""
    st, dis = problem.states.copy(), problem.discharge.copy()
    qt, ser = oracle.land_run(problem.land, st, problem.forcing, problem.doy, series=list(series), threads=2)
    dser = np.zeros((problem.nt, problem.river.n_seg)) if reach_states else None
    node_rows, rout, rin = oracle.river_run(problem.river, dis, qt, problem.ext, dis_series=dser)
    out = dict(states=st, discharge=dis, node_rows=node_rows, land_rows=qt, reach_out=rout, reach_in=rin, series=ser)
    if reach_states:
        out["reach_state_series"] = dser
    return out


def snapshot(hp):
    out = {}
    for node in hp.nodes:
        out[f"node:{node.name}"] = node.sequences.sim.series.copy()
        out[f"nodevalue:{node.name}"] = np.array(node.sequences.sim.value)
    for element in hp.elements:
        seqs = element.model.sequences
        for group in ("factors", "fluxes", "states", "outlets"):
            for seq in getattr(seqs, group, ()):
                if getattr(seq, "ramflag", False):
                    out[f"{element.name}:{group}.{seq.name}:series"] = np.array(seq.series)
                if group == "states":
                    out[f"{element.name}:states.{seq.name}:new"] = np.array(seq.new)
                    out[f"{element.name}:states.{seq.name}:old"] = np.array(seq.old)
        rc = getattr(element.model, "rconcmodel", None)
        if rc is not None:
            out[f"{element.name}:quh"] = np.array(rc.sequences.logs.quh.value)
    return out


def main() -> None:
    del pub.projectname
    del pub.timegrids
    pub.options.prepare_testing(usecython=True, threads=0)
    with contextlib.redirect_stdout(io.StringIO()):
        hp, _, _ = testtools.prepare_full_example_2(lastdate="1996-03-01")
    for element in hp.elements:
        seqs = element.model.sequences
        for group in ("factors", "fluxes", "states"):
            for seq in getattr(seqs, group):
                seq.prepare_series(allocate_ram=True)
    # two consecutive simulation windows inside the initialisation period, as in hydpy/core/hydpytools.py:2804-2825
    windows = (("1996-01-01", "1996-02-01"), ("1996-02-01", "1996-03-01"))

    import copy

    initial = copy.deepcopy(hp.conditions)

    def run(simulate):
        hp.conditions = copy.deepcopy(initial)
        for node in hp.nodes:
            node.sequences.sim.series[:] = np.nan
        for element in hp.elements:
            for group in ("factors", "fluxes", "states", "outlets"):
                for seq in getattr(element.model.sequences, group, ()):
                    if getattr(seq, "ramflag", False):
                        seq.series[:] = np.nan
        snaps = []
        for first, last in windows:
            pub.timegrids.sim.firstdate = first
            pub.timegrids.sim.lastdate = last
            with contextlib.redirect_stdout(io.StringIO()):
                simulate()
            snaps.append(snapshot(hp))
        return snaps

    ref = run(hp.simulate)
    simulate_module.run_problem = oracle_run_problem            # the engine call, replaced by the oracle
    got = run(lambda: hydpy_b200.simulate(hp))
    hydpy_b200.install()
    got_installed = run(hp.simulate)                            # the monkey-patched HydPy.simulate
    hydpy_b200.uninstall()
    if os.environ.get("DROPIN_DEBUG"):
        a, b = ref[0], got[0]
        for key in a:
            if key.startswith("land_dill_assl") and key.endswith("series") and not np.array_equal(a[key], b[key], equal_nan=True):
                x, y = np.array(a[key]), np.array(b[key])
                bad = np.argwhere(x != y)
                print("DIFF", key, x.shape, "first", bad[0], x[tuple(bad[0])], y[tuple(bad[0])])
    n = 0
    for which, snaps in (("simulate", got), ("install", got_installed)):
        for w, (a, b) in enumerate(zip(ref, snaps)):
            assert a.keys() == b.keys()
            for key in a:
                if not np.array_equal(a[key], b[key], equal_nan=True):
                    idx = np.flatnonzero(~np.isclose(np.ravel(a[key]), np.ravel(b[key]), rtol=0, atol=0, equal_nan=True))
                    raise AssertionError(f"{which}, window {w}: {key} differs at {idx[:8]} of {np.size(a[key])}: "
                                         f"{np.ravel(a[key])[idx[:4]]} vs {np.ravel(b[key])[idx[:4]]}")
                n += 1
    print(f"DROPIN_OK compared {n} arrays over {len(windows)} windows, {len(hp.elements)} elements, {len(hp.nodes)} nodes")


if __name__ == "__main__":
    main()
