"""Parity of the CUDA engine (through the C-ABI) against the reference-generated golden fixtures.

Bar (BASELINE.json north_star): ≤1e-9 relative on discharge and states.  Everything that does not pass through
libm (`pow`/`exp`) is expected — and asserted — to be bit-identical, because the kernels evaluate the same IEEE
operations in the same order (-fmad=false) and the seasonal sine is taken from the host libm.
"""
import numpy as np
import pytest

from hydpy_b200 import Engine, layout, run_problem
from tests.util import assert_close, load_scenarios

pytestmark = pytest.mark.gpu

HLAND = load_scenarios("hland_96_integration.npz")
HLAND_P = load_scenarios("hland_96p_integration.npz")
HLAND_C = load_scenarios("hland_96c_integration.npz")
HLAND_RETIO = load_scenarios("hland_96_ret_io.npz")
MUSK = load_scenarios("musk_classic_integration.npz")
MUSK_MCT = load_scenarios("musk_mct_integration.npz")
LAHN = load_scenarios("lahn_daily.npz")
LAHN_H = load_scenarios("lahn_hourly_members.npz")

# sequences whose values never depend on pow/exp (not even through the interception storage, which feels exp() via
# the potential evaporation): must be bit-identical
EXACT_SERIES = ("tc", "fracrain", "cfact", "gact", "pc")


@pytest.fixture(scope="module", params=["auto", "general", "grouped"])
def engine(request):
    """Both land paths ('auto' = two-kernel fast path for the eligible elements, 'general' = general kernel only) and
    both routing sweeps ('auto': headwater kernels + persistent wavefront, 'general': one launch per anti-diagonal)."""
    with Engine(0) as eng:
        eng.set_land_path("auto" if request.param == "grouped" else request.param)
        eng.set_river_sweep({"auto": "persistent", "general": "launches", "grouped": "grouped"}[request.param])
        eng.path = request.param
        yield eng


def run_and_check(engine, sc, window=None):
    p = sc.problem
    res = run_problem(p, engine, series=sc.expected_series, window=window)
    worst = {}
    for name in sc.expected_series:
        worst[name] = assert_close(res["series"][name], sc.expected(name), f"{sc.name}:{name}")
    wb = p.node_writeback
    assert_close(res["node_rows"][wb], sc.expected("node_rows")[wb], f"{sc.name}:node series")
    for n in res["states"].NAMES:
        if sc.expected("state_" + n) is not None:   # fixtures of ABI 1 do not hold the sibling models' states
            assert_close(getattr(res["states"], n), sc.expected("state_" + n), f"{sc.name}:final state {n}")
    assert_close(res["discharge"], sc.expected("discharge"), f"{sc.name}:final discharge")
    if sc.expected("reach_out") is not None:
        assert_close(res["reach_out"], sc.expected("reach_out"), f"{sc.name}:reach outflow")
        assert_close(res["reach_in"], sc.expected("reach_in"), f"{sc.name}:reach inflow")
    # (the fast path multiplies by host-folded reciprocals instead of dividing: <= 1 ulp, so only the general kernel is
    # held to bit-exactness here)
    for name in EXACT_SERIES if engine.path == "general" else ():
        if name in res["series"]:
            assert np.array_equal(res["series"][name], sc.expected(name), equal_nan=True), f"{sc.name}:{name} not exact"
    # discharge-only run (no series requested): the eligible elements take the fast path
    res0 = run_problem(p, engine, window=window)
    assert_close(res0["node_rows"][wb], sc.expected("node_rows")[wb], f"{sc.name}:node series (discharge-only run)")
    for n in res0["states"].NAMES:
        if sc.expected("state_" + n) is not None:
            assert_close(getattr(res0["states"], n), sc.expected("state_" + n), f"{sc.name}:final state {n} (discharge-only)")
    assert_close(res0["discharge"], sc.expected("discharge"), f"{sc.name}:final discharge (discharge-only run)")
    return res, worst


@pytest.mark.parametrize("name", sorted(HLAND))
def test_hland_96_integration(engine, name):
    run_and_check(engine, HLAND[name])


@pytest.mark.parametrize("name", sorted(HLAND_RETIO))
def test_hland_96_with_evap_ret_io(engine, name):
    """`evap_ret_io` behind `evap_aet_hbv96` (external reference evapotranspiration; ref: hydpy/models/evap_ret_io.py:53-64)
    on the kernels of the evap_pet_hbv96 equations in their identity setting — exp(-0.0 * pc) must be exactly 1."""
    run_and_check(engine, HLAND_RETIO[name])


@pytest.mark.parametrize("name", sorted(HLAND_P))
def test_hland_96p_integration(engine, name):
    """hland_96p + rconc_nash, all 41 sequences of the reference's six integration tests."""
    run_and_check(engine, HLAND_P[name])


@pytest.mark.parametrize("name", sorted(HLAND_C))
def test_hland_96c_integration(engine, name):
    """hland_96c + rconc_nash, all 36 sequences of the reference's six integration tests."""
    run_and_check(engine, HLAND_C[name])


@pytest.mark.parametrize("name", sorted(MUSK))
def test_musk_classic_integration(engine, name):
    sc = MUSK[name]
    res, _ = run_and_check(engine, sc)
    # routing has no libm call at all: bit-identical
    assert np.array_equal(res["reach_out"], sc.expected("reach_out"))
    assert np.array_equal(res["discharge"], sc.expected("discharge"))


@pytest.mark.parametrize("name", sorted(MUSK_MCT))
def test_musk_mct_integration(engine, name):
    """musk_mct + wq_trapeze_strickler — outflow, states.discharge of every step, final Courant/Reynolds
    numbers and reference water depths of the reference's five integration tests (root search included)."""
    sc = MUSK_MCT[name]
    p = sc.problem
    for window in (None, 37):
        res = run_problem(p, engine, reach_states=True, window=window, mct_series=layout.MCT_SERIES)
        for seq in layout.MCT_SERIES:   # hydraulic factors, reference discharge, Courant/Reynolds numbers of every step
            assert_close(res["mct_series"][seq], sc.expected("mct_" + seq), f"{sc.name}: series of {seq} (window {window})")
        assert_close(res["reach_out"], sc.expected("reach_out"), f"{sc.name}: outflow (window {window})")
        assert_close(res["reach_in"], sc.expected("reach_in"), f"{sc.name}: inflow")
        assert_close(res["node_rows"][p.node_writeback], sc.expected("node_rows")[p.node_writeback], f"{sc.name}: nodes")
        assert_close(res["discharge"], sc.expected("discharge"), f"{sc.name}: final discharge")
        exp = sc.expected("dis_series")
        assert_close(np.where(np.isnan(exp), np.nan, res["reach_state_series"]), exp, f"{sc.name}: states.discharge series")
        assert_close(res["mct_states"], sc.expected("mct_states"), f"{sc.name}: final Courant/Reynolds numbers, water depths")


def test_lahn_daily(engine):
    sc = LAHN["lahn_daily"]
    res, worst = run_and_check(engine, sc)
    names = [str(n) for n in sc.data["node_names"]]
    # the reference's docstring goldens, ref: hydpy/core/hydpytools.py:2791-2794
    np.testing.assert_allclose(res["node_rows"][names.index("lahn_kalk")][:4],
                               [54.019332, 37.257552, 31.865302, 28.359538], rtol=0, atol=6e-7)
    print("lahn_daily max rel deviations:", {k: f"{v:.2e}" for k, v in worst.items()})


def test_lahn_daily_windowed_equals_single_window(engine):
    sc = LAHN["lahn_daily"]
    a = run_problem(sc.problem, engine)
    b = run_problem(sc.problem, engine, window=100)
    assert np.array_equal(a["node_rows"], b["node_rows"])
    for n in a["states"].NAMES:
        assert np.array_equal(getattr(a["states"], n), getattr(b["states"], n))
    assert np.array_equal(a["discharge"], b["discharge"])


@pytest.mark.parametrize("name", sorted(LAHN_H))
def test_lahn_hourly_members(engine, name):
    run_and_check(engine, LAHN_H[name])
