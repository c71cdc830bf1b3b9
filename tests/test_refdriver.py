"""The binaries of the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.py) driven directly: they must
reproduce the golden fixtures (which came from the same reference through its full Python API) and agree with the C
restatement on random ragged problems — this is what licenses `cpu_baseline.kind = "reference"` in bench.py."""
import numpy as np
import pytest

from oracle import oracle, refdriver
from tests.randland import random_forcing, random_land, random_river
from tests.util import load_scenarios

pytestmark = pytest.mark.skipif(not refdriver.available(), reason="oracle/_ref not built (needs /root/reference)")


def test_lahn_daily_through_reference_binaries():
    sc = load_scenarios("lahn_daily.npz")["lahn_daily"]
    p = sc.problem
    st = p.states.copy()
    els = refdriver.prepare_land(p.land, st, p.forcing, p.doy, keep_series=("sm", "uz", "qt"))
    qt = refdriver.run_land(els, threads=4)
    for el in els:
        el.store_states(p.land, st)
    dis = p.discharge.copy()
    node_rows, rout, rin = refdriver.run_river(p.river, dis, qt, p.ext)
    assert np.array_equal(node_rows, sc.expected("node_rows"))
    assert np.array_equal(rout, sc.expected("reach_out"))
    for n in st.NAMES_V1:
        assert np.array_equal(getattr(st, n), sc.expected("state_" + n)), n
    assert np.array_equal(dis, sc.expected("discharge"))
    sm = np.concatenate([el.series["sm"] for el in els], axis=1)
    assert np.array_equal(sm, sc.expected("sm"))


@pytest.mark.parametrize("seed", [3, 4])
def test_restatement_equals_reference_binaries_on_random_problems(seed):
    ne, nt = 40, 48
    land, states = random_land(ne, seed=seed)
    forcing, doy = random_forcing(nt, ne, seed=seed)
    river, discharge, _ = random_river(ne, 20, seed=seed)
    rng = np.random.default_rng(seed)
    ext = rng.uniform(0, 3, size=(river.n_ext, nt))
    ext[1::2][rng.uniform(size=ext[1::2].shape) < 0.3] = np.nan
    st_a, st_b = states.copy(), states.copy()
    qt_a, _ = oracle.land_run(land, st_a, forcing, doy)
    els = refdriver.prepare_land(land, st_b, forcing, doy)
    qt_b = refdriver.run_land(els, threads=2)
    for el in els:
        el.store_states(land, st_b)
    assert np.array_equal(qt_a, qt_b)
    for n in st_a.NAMES:
        assert np.array_equal(getattr(st_a, n), getattr(st_b, n), equal_nan=True), n
    dis_a, dis_b = discharge.copy(), discharge.copy()
    na, oa, ia = oracle.river_run(river, dis_a, qt_a, ext)
    nb, ob, ib = refdriver.run_river(river, dis_b, qt_b, ext)
    assert np.array_equal(na, nb, equal_nan=True) and np.array_equal(oa, ob, equal_nan=True)
    assert np.array_equal(dis_a, dis_b, equal_nan=True)
