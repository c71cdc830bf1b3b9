"""Error conventions at the C-ABI (include/hydpy_b200.h): every failure is a negative HB_E* code plus a message that the
Python layer raises as EngineError — the engine never guesses and never falls back (last row)."""
import numpy as np
import pytest

from hydpy_b200 import Engine, EngineError, layout, synth
from hydpy_b200.abi import RiverBundle

pytestmark = pytest.mark.gpu


@pytest.fixture()
def engine():
    with Engine(0) as eng:
        yield eng


def small(ne=8, nt=6):
    land, states = synth.make_land(ne, "1h")
    river, discharge = synth.make_tree_river(ne, "1h", levels=3)
    forcing, doy = synth.make_forcing(nt, "1h")
    return land, states, river, discharge, forcing, doy


def test_call_order_is_enforced(engine):
    land, states, river, discharge, forcing, doy = small()
    with pytest.raises(EngineError, match="hb_set_land first"):
        engine._check(engine._lib.hb_run(engine._h, 3), "hb_run")
    engine.set_land(land)
    with pytest.raises(EngineError, match="hb_set_forcing first"):
        engine.run()
    engine.set_forcing(forcing, doy)
    with pytest.raises(EngineError, match="hb_set_land_states first"):
        engine.run(land=True, river=False)
    engine.set_land_states(states)
    with pytest.raises(EngineError, match="hb_set_river first"):
        engine.run()
    engine.set_river(river)
    engine.set_river_states(discharge)
    engine.run()                                   # n_ext == 0: no external rows needed
    with pytest.raises(EngineError, match="not selected"):
        engine.get_series("sm")
    with pytest.raises(EngineError, match="not recorded"):
        engine.get_reach_state_series()


def test_cyclic_network_is_rejected(engine):
    land, states, river, *_ = small()
    # reach 0 routes node 1 to node 0; let one more reach route node 0 back into node 1
    bad = RiverBundle(
        entry_ptr=river.entry_ptr.copy(), entry_src=river.entry_src.copy(), node_mode=river.node_mode, node_ext=river.node_ext,
        inlet_ptr=np.append(river.inlet_ptr, river.inlet_ptr[-1] + 1), inlet_node=np.append(river.inlet_node, 0),
        reach_outlet=np.append(river.reach_outlet, 1), seg_ptr=np.append(river.seg_ptr, river.seg_ptr[-1] + 2),
        coef=np.concatenate([river.coef, np.array([[0.2], [0.6], [0.2]])], axis=1), n_ext=0)
    src = bad.entry_src.tolist()
    end1 = int(bad.entry_ptr[2])
    src.insert(end1, -1 - (bad.n_reach - 1))
    bad.entry_src = np.array(src, np.int32)
    bad.entry_ptr = bad.entry_ptr.copy()
    bad.entry_ptr[2:] += 1
    engine.set_land(land)
    with pytest.raises(EngineError, match="cycle"):
        engine.set_river(bad)


def test_invalid_descriptions_are_rejected(engine):
    land, states, river, discharge, forcing, doy = small()
    broken = land.zonetype.copy()
    broken[3] = 9
    land_bad = type(land)(**{**{f: getattr(land, f) for f in land.__dataclass_fields__}, "zonetype": broken})
    with pytest.raises(EngineError, match="invalid type"):
        engine.set_land(land_bad)
    engine.set_land(land)
    river_bad = RiverBundle(**{**{f: getattr(river, f) for f in river.__dataclass_fields__},
                               "entry_src": np.where(river.entry_src >= 0, river.entry_src + 100, river.entry_src)})
    with pytest.raises(EngineError, match="land element"):
        engine.set_river(river_bad)
    engine.set_forcing(forcing, doy)
    with pytest.raises(EngineError, match="outside the staged block"):
        engine.select_window(3, 10)
    with pytest.raises(EngineError, match="unknown option"):
        engine._check(engine._lib.hb_set_option(engine._h, 99, 0), "hb_set_option")
    with pytest.raises(ValueError):
        engine.set_forcing(forcing[:3], doy)       # caught by the Python layer before the call


def test_nan_forcing_is_handled_like_in_the_reference(engine):
    """NaN inputs are legal (ref: hydpy/core/hydpytools.py:2889-2906): no exception.  Where the NaN ends up follows from
    Cython's lowering of min/max (NaN-asymmetric): soil moisture and UZ of the affected element turn NaN,
    its discharge does not (q0 drops to zero, the lower zone keeps draining) — engine and oracle must agree cell by cell."""
    from oracle import oracle

    land, states, river, discharge, forcing, doy = small(ne=40, nt=12)
    forcing = forcing.copy()
    col = int(land.forcing_col[5])
    forcing[0, 4, col] = np.nan                    # precipitation of one column at step 4
    for path in ("auto", "general"):
        engine.set_land_path(path)
        engine.set_land(land); engine.set_land_states(states); engine.set_river(river); engine.set_river_states(discharge)
        engine.set_forcing(forcing, doy); engine.set_external(np.zeros((0, 12)))
        engine.run()
        rows = engine.get_land_outflow()
        got = engine.get_land_states()
        st = states.copy()
        qt, _ = oracle.land_run(land, st, forcing, doy, threads=1)
        assert np.array_equal(np.isnan(rows), np.isnan(qt)), path
        for n in st.NAMES:
            assert np.array_equal(np.isnan(getattr(got, n)), np.isnan(getattr(st, n))), (path, n)
        affected = np.flatnonzero(land.forcing_col == col)
        assert np.isnan(got.uz[affected]).all() and np.isfinite(got.uz[np.setdiff1d(np.arange(land.n_elem), affected)]).all()
        ok = ~np.isnan(qt)
        np.testing.assert_allclose(rows[ok], qt[ok], rtol=1e-9, atol=1e-12)


def test_mct_series_need_a_request_and_musk_mct_reaches(engine):
    """hb_get_mct_series: recorded only for the bits of HB_OPT_RIVER_MCT_SERIES and only when the river holds musk_mct
    reaches — everything else is an error, never a buffer of zeros."""
    land, states, river, discharge, forcing, doy = small()
    engine.set_land(land); engine.set_land_states(states); engine.set_river(river); engine.set_river_states(discharge)
    engine.set_forcing(forcing, doy)
    engine.record_mct_series(("celerity",))
    engine.run()
    with pytest.raises(EngineError, match="was not recorded"):
        engine.get_mct_series("celerity")              # musk_classic reaches only
    with pytest.raises(EngineError, match="unknown sequence"):
        engine._check(engine._lib.hb_get_mct_series(engine._h, 99, None), "hb_get_mct_series")
    with pytest.raises(EngineError, match="invalid musk_mct series mask"):
        engine._check(engine._lib.hb_set_option(engine._h, layout.OPT_RIVER_MCT_SERIES, 1 << len(layout.MCT_SERIES)),
                      "hb_set_option")
    with pytest.raises(ValueError, match="unknown musk_mct sequence"):
        engine.record_mct_series(("discharge",))
    engine.record_mct_series(())
