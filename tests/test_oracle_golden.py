"""Pin the CPU oracle (oracle/hland_oracle.c) against fixtures produced by the reference itself.

The fixtures (tests/golden/*.npz, generator tests/golden/make_golden.py) hold raw float64 results of the reference's
Cython path for: the 11 `hland_96` integration tests (ref: hydpy/models/hland_96.py:232-1653), the 3 `musk_classic`
integration tests (ref: hydpy/models/musk_classic.py:74-180), HydPy-H-Lahn daily 1996-1997 (ref:
hydpy/core/hydpytools.py:2791-2947) and a miniature of config 2 (Lahn hourly, 3 parameter members).

Same compiler family, same libm, no FMA contraction ⇒ the restatement is expected to be BIT-IDENTICAL; the asserted
bar is nevertheless the north_star tolerance (1e-9) plus an explicit report of exactness.
"""
import numpy as np
import pytest

from hydpy_b200 import layout
from oracle import oracle
from tests.util import assert_close, load_scenarios

HLAND = load_scenarios("hland_96_integration.npz")
HLAND_P = load_scenarios("hland_96p_integration.npz")
HLAND_C = load_scenarios("hland_96c_integration.npz")
HLAND_RETIO = load_scenarios("hland_96_ret_io.npz")
MUSK = load_scenarios("musk_classic_integration.npz")
MUSK_MCT = load_scenarios("musk_mct_integration.npz")
LAHN = load_scenarios("lahn_daily.npz")
LAHN_H = load_scenarios("lahn_hourly_members.npz")


def run_oracle(sc, threads=1):
    p = sc.problem
    states = p.states.copy()
    qt, series = oracle.land_run(p.land, states, p.forcing, p.doy, series=sc.expected_series, threads=threads)
    discharge = p.discharge.copy()
    mct = None if p.mct_states is None else p.mct_states.copy()
    dser = np.full((p.nt, p.river.n_seg), np.nan)
    mser = np.full((len(layout.MCT_SERIES), p.nt, p.river.n_seg), np.nan) if p.river.any_mct else None
    node_rows, rout, rin = oracle.river_run(p.river, discharge, qt, p.ext, dis_series=dser, mct_states=mct,
                                            mct_series=mser)
    return dict(states=states, qt=qt, series=series, discharge=discharge, node_rows=node_rows, rout=rout, rin=rin,
                mct_states=mct, dis_series=dser, mct_series=mser)


def check(sc, res, exact=True):
    worst = 0.0
    for name in sc.expected_series:
        worst = max(worst, assert_close(res["series"][name], sc.expected(name), f"{sc.name}:{name}"))
    wb = sc.problem.node_writeback  # nodes in `oldsim` mode keep their given series (threadingtools.py:236,257)
    assert_close(res["node_rows"][wb], sc.expected("node_rows")[wb], f"{sc.name}:node series")
    for n in res["states"].NAMES:
        if sc.expected("state_" + n) is not None:   # fixtures of ABI 1 do not hold the sibling models' states
            assert_close(getattr(res["states"], n), sc.expected("state_" + n), f"{sc.name}:final state {n}")
    assert_close(res["discharge"], sc.expected("discharge"), f"{sc.name}:final discharge")
    if sc.expected("reach_out") is not None:
        assert_close(res["rout"], sc.expected("reach_out"), f"{sc.name}:reach outflow")
        assert_close(res["rin"], sc.expected("reach_in"), f"{sc.name}:reach inflow")
    if exact:
        # the stronger statement: bit-identical to the reference build
        for name in sc.expected_series:
            assert np.array_equal(res["series"][name], sc.expected(name), equal_nan=True), f"{sc.name}:{name} not exact"
        assert np.array_equal(res["node_rows"][wb], sc.expected("node_rows")[wb]), f"{sc.name}: node series not exact"
    return worst


@pytest.mark.parametrize("name", sorted(HLAND))
def test_hland_96_integration(name):
    sc = HLAND[name]
    assert len(sc.expected_series) == 44  # every factor/flux/state sequence + the 8 of the evap submodels
    check(sc, run_oracle(sc))


@pytest.mark.parametrize("name", sorted(HLAND_RETIO))
def test_hland_96_with_evap_ret_io(name):
    """`evap_ret_io` as the PET submodel of `evap_aet_hbv96` (ref: hydpy/models/evap_ret_io.py:53-64): extract.py maps it
    onto the evap_pet_hbv96 equations with the identity setting (all factors but EvapotranspirationFactor zero), which
    must reproduce the reference bit for bit — every sequence of both elements (without / with capillary flux)."""
    sc = HLAND_RETIO[name]
    zp = sc.problem.land.zp
    assert not zp("airtemperaturefactor").any() and not zp("altitudefactor").any() and not zp("precipitationfactor").any()
    assert "referenceevapotranspiration" in sc.expected_series and "ea" in sc.expected_series
    check(sc, run_oracle(sc))


@pytest.mark.parametrize("name", sorted(HLAND_P))
def test_hland_96p_integration(name):
    """hland_96p + rconc_nash (ref: hydpy/models/hland_96p.py integration tests)."""
    sc = HLAND_P[name]
    assert len(sc.expected_series) in (49, 50)  # every sequence of hland_96p + 8 of the evap submodels + rconc_nash states.sc
    check(sc, run_oracle(sc))


@pytest.mark.parametrize("name", sorted(HLAND_C))
def test_hland_96c_integration(name):
    """hland_96c + rconc_nash (ref: hydpy/models/hland_96c.py integration tests)."""
    sc = HLAND_C[name]
    assert len(sc.expected_series) in (44, 45)  # every sequence of hland_96c + 8 of the evap submodels + rconc_nash states.sc
    check(sc, run_oracle(sc))


@pytest.mark.parametrize("name", sorted(MUSK))
def test_musk_classic_integration(name):
    sc = MUSK[name]
    check(sc, run_oracle(sc))


@pytest.mark.parametrize("name", sorted(MUSK_MCT))
def test_musk_mct_integration(name):
    """musk_mct + wq_trapeze_strickler (ref: hydpy/models/musk_mct.py integration tests): outflow,
    states.discharge of every step and the final conditions (Courant/Reynolds numbers, reference water depths)."""
    sc = MUSK_MCT[name]
    res = run_oracle(sc)
    check(sc, res)
    exp = sc.expected("dis_series")
    assert np.array_equal(res["dis_series"][~np.isnan(exp)], exp[~np.isnan(exp)]), "states.discharge series not exact"
    assert np.array_equal(res["mct_states"], sc.expected("mct_states"), equal_nan=True)
    # every 1-D sequence of every step (hydraulic factors, reference discharge, Courant and Reynolds numbers): bit for bit
    for k, seq in enumerate(layout.MCT_SERIES):
        assert np.array_equal(res["mct_series"][k], sc.expected("mct_" + seq), equal_nan=True), f"series of {seq}"


def test_lahn_daily():
    sc = LAHN["lahn_daily"]
    res = run_oracle(sc)
    check(sc, res)
    names = [str(n) for n in sc.data["node_names"]]
    # the reference's docstring goldens (6 digits), ref: hydpy/core/hydpytools.py:2791-2794, 2846-2847;
    # hydpy/core/modeltools.py:2489-2492
    for node, vals in {
        "lahn_kalk": [54.019332, 37.257552, 31.865302, 28.359538],
        "lahn_leun": [42.34647, 27.157463, 22.880985, 20.156832],
        "dill_assl": [11.757521, 8.865071, 7.10181, 5.994192],
    }.items():
        np.testing.assert_allclose(res["node_rows"][names.index(node)][:4], vals, rtol=0, atol=6e-7)


def test_lahn_daily_threaded_is_identical():
    sc = LAHN["lahn_daily"]
    a, b = run_oracle(sc, threads=1), run_oracle(sc, threads=3)
    assert np.array_equal(a["node_rows"], b["node_rows"])


@pytest.mark.parametrize("name", sorted(LAHN_H))
def test_lahn_hourly_members(name):
    sc = LAHN_H[name]
    check(sc, run_oracle(sc))
