"""CUDA engine (through the C-ABI) vs the CPU oracle on seeded random problems at sizes the oracle finishes in seconds,
plus size-independent properties at larger sizes (window splitting, element permutation, ensemble replication)."""
import numpy as np
import pytest

from hydpy_b200 import Engine, Problem, layout, run_problem, synth
from oracle import oracle
from tests.randland import add_mct, add_siblings, random_forcing, random_land, random_river
from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["auto", "general", "grouped"])
def engine(request):
    """Both land paths ('auto' = two-kernel fast path for the eligible elements, 'general' = general kernel only) and the
    three routing sweeps ('auto': headwater kernels + persistent warp-per-reach wavefront, 'general': one launch per
    anti-diagonal, 'grouped': fast land path + persistent wavefront with (32 reaches, chunk) tasks)."""
    with Engine(0) as eng:
        eng.set_land_path("auto" if request.param == "grouped" else request.param)
        eng.set_river_sweep({"auto": "persistent", "general": "launches", "grouped": "grouped"}[request.param])
        yield eng


def oracle_run(problem, series=()):
    st = problem.states.copy()
    qt, ser = oracle.land_run(problem.land, st, problem.forcing, problem.doy, series=series, threads=4)
    dis = problem.discharge.copy()
    dser = np.zeros((problem.nt, problem.river.n_seg))
    mct = None if problem.mct_states is None else problem.mct_states.copy()
    mser = None if mct is None else np.full((len(layout.MCT_SERIES), problem.nt, problem.river.n_seg), np.nan)
    node_rows, rout, rin = oracle.river_run(problem.river, dis, qt, problem.ext, dis_series=dser, mct_states=mct,
                                            mct_series=mser)
    return dict(states=st, land_rows=qt, series=ser, discharge=dis, node_rows=node_rows, reach_out=rout, reach_in=rin,
                reach_state_series=dser, mct_states=mct,
                mct_series=None if mser is None else dict(zip(layout.MCT_SERIES, mser)))


def compare(res, ref, what, series=()):
    assert_close(res["land_rows"], ref["land_rows"], f"{what}: land outflow")
    assert_close(res["node_rows"], ref["node_rows"], f"{what}: node series")
    assert_close(res["reach_out"], ref["reach_out"], f"{what}: reach outflow")
    assert_close(res["reach_in"], ref["reach_in"], f"{what}: reach inflow")
    assert_close(res["discharge"], ref["discharge"], f"{what}: final discharge")
    for n in ref["states"].NAMES:
        assert_close(getattr(res["states"], n), getattr(ref["states"], n), f"{what}: final {n}")
    for name in series:
        assert_close(res["series"][name], ref["series"][name], f"{what}: series {name}")


def check_window_splitting(problem, engine, window, keys, states=False):
    """Splitting the period into windows must not change a single bit (states stay on the device) — with one documented
    exception: elements with a finite SMax on the speculative fast path (`HB_OPT_LAND_SPECULATE`) run a window in which
    their snow exceeds the limit in the general kernel (the reference's operations) and every other window on the fast path
    (reciprocal multiplications, <= 1 ulp per operation), so where the windows are cut decides the last bits.  Bit
    equality is therefore demanded with the speculation off, the parity bar of the unsplit run with it on."""
    engine.set_land_speculate(False)
    try:
        a = run_problem(problem, engine)
        b = run_problem(problem, engine, window=window)
    finally:
        engine.set_land_speculate(True)
    for key in keys:
        assert np.array_equal(a[key], b[key], equal_nan=True), f"window splitting changes {key}"
    if states:
        for n in a["states"].NAMES:
            assert np.array_equal(getattr(a["states"], n), getattr(b["states"], n), equal_nan=True), n
    c = run_problem(problem, engine, window=window)
    for key in keys:
        assert_close(c[key], a[key], f"window splitting with speculation: {key}")
    if states:
        for n in a["states"].NAMES:
            assert_close(getattr(c["states"], n), getattr(a["states"], n), f"window splitting with speculation: final {n}")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_ragged_problem_all_series(engine, seed):
    ne, nt = 150, 60
    land, states = random_land(ne, seed=seed)
    forcing, doy = random_forcing(nt, ne, seed=seed)
    river, discharge, picks = random_river(ne, 60, seed=seed)
    rng = np.random.default_rng(seed)
    ext = rng.uniform(0, 3, size=(river.n_ext, nt))
    ext[1::2][rng.uniform(size=ext[1::2].shape) < 0.3] = np.nan  # obs_newsim rows have gaps
    problem = Problem(land=land, states=states, river=river, discharge=discharge, forcing=forcing, doy=doy, ext=ext)
    ref = oracle_run(problem, series=layout.SERIES)
    res = run_problem(problem, engine, series=layout.SERIES)
    compare(res, ref, f"seed {seed}", series=layout.SERIES)
    # without series the eligible elements take the fast path: same bar against the oracle
    res1 = run_problem(problem, engine)
    compare(res1, ref, f"seed {seed} (no series)")
    check_window_splitting(problem, engine, 17, ("land_rows", "node_rows", "reach_out", "discharge"))


@pytest.mark.parametrize("seed", [0, 1])
def test_random_ragged_problem_with_sibling_models(engine, seed):
    """hland_96, hland_96p and hland_96c elements with rconc_uh / rconc_nash / no submodel side by side in one basin
    (SURVEY.md §8f-3): every sequence against the oracle; cells of sequences a model does not own are NaN on both sides."""
    ne, nt = 160, 72
    land, states = random_land(ne, seed=seed + 20)
    add_siblings(land, states, seed=seed + 20)
    assert len(set(land.model.tolist())) == 3 and len(set(land.rconc.tolist())) == 2
    forcing, doy = random_forcing(nt, ne, seed=seed + 20)
    river, discharge, _ = random_river(ne, 50, seed=seed + 20, n_ext=0)
    problem = Problem(land=land, states=states, river=river, discharge=discharge, forcing=forcing, doy=doy,
                      ext=np.zeros((0, nt)))
    ref = oracle_run(problem, series=layout.SERIES)
    res = run_problem(problem, engine, series=layout.SERIES)
    compare(res, ref, f"siblings seed {seed}", series=layout.SERIES)
    assert np.isnan(res["series"]["dp"][:, np.repeat(land.model, np.diff(land.zone_ptr)) != layout.MODEL_HLAND_96P]).all()
    res1 = run_problem(problem, engine)
    compare(res1, ref, f"siblings seed {seed} (no series)")
    check_window_splitting(problem, engine, 13, ("land_rows", "node_rows", "reach_out", "discharge"), states=True)


@pytest.mark.parametrize("seed", [0, 1])
def test_random_network_with_musk_mct_reaches(engine, seed):
    """musk_classic and musk_mct reaches (1-3 trapezes, 0..40 segments, one or two runs per segment, fresh and warm start
    values of the root search) side by side in a random DAG with deploy modes."""
    ne, nt = 40, 80
    land, states = random_land(ne, seed=seed + 40)
    forcing, doy = random_forcing(nt, ne, seed=seed + 40)
    river, discharge, _ = random_river(ne, 30, seed=seed + 40)
    mct = add_mct(river, seed=seed + 40)
    assert river.any_mct and (river.reach_model == layout.REACH_CLASSIC).any()
    rng = np.random.default_rng(seed)
    ext = rng.uniform(0, 3, size=(river.n_ext, nt))
    ext[1::2][rng.uniform(size=ext[1::2].shape) < 0.3] = np.nan
    problem = Problem(land=land, states=states, river=river, discharge=discharge, forcing=forcing, doy=doy, ext=ext,
                      mct_states=mct)
    ref = oracle_run(problem)
    res = run_problem(problem, engine, reach_states=True, mct_series=layout.MCT_SERIES)
    compare(res, ref, f"musk_mct seed {seed}")
    assert_close(res["reach_state_series"], ref["reach_state_series"], "musk_mct: states.discharge of every step")
    assert_close(res["mct_states"], ref["mct_states"], "musk_mct: final Courant/Reynolds numbers and water depths")
    for seq in layout.MCT_SERIES:   # NaN in the cells of musk_classic reaches and at the last point of every reach
        assert_close(res["mct_series"][seq], ref["mct_series"][seq], f"musk_mct: series of {seq}")
    some = ("celerity", "coefficient2", "reynoldsnumber")
    res2 = run_problem(problem, engine, window=23, mct_series=some)
    assert sorted(res2["mct_series"]) == sorted(some)
    for key in ("node_rows", "reach_out", "discharge", "mct_states"):
        assert np.array_equal(res[key], res2[key], equal_nan=True), f"window splitting changes {key}"
    for seq in some:
        assert np.array_equal(res["mct_series"][seq], res2["mct_series"][seq], equal_nan=True), f"window splitting changes {seq}"
    with pytest.raises(Exception, match="was not recorded"):
        engine.get_mct_series("wettedarea")
    engine.record_mct_series(())


def test_nan_forcing_in_a_mixed_basin(engine):
    """NaN inputs are legal (ref: hydpy/core/hydpytools.py:2889-2906) and spread as Cython's min/max lowering dictates
     — through hland_96p / hland_96c, the storage cascade of rconc_nash and the root search of musk_mct
    (ref: rootutils.pyx:35-78: a NaN reference discharge ends the search with NaN).  Engine and oracle must agree on the
    NaN pattern of every array and on every finite value."""
    ne, nt = 120, 48
    land, states = random_land(ne, seed=77)
    add_siblings(land, states, seed=77)
    forcing, doy = random_forcing(nt, ne, seed=77)
    forcing = forcing.copy()
    rng = np.random.default_rng(77)
    for col in rng.choice(ne, size=12, replace=False):
        forcing[int(rng.integers(0, 4)), int(rng.integers(5, nt - 5)), col] = np.nan
    river, discharge, _ = random_river(ne, 40, seed=77, n_ext=0)
    mct = add_mct(river, seed=77)
    problem = Problem(land=land, states=states, river=river, discharge=discharge, forcing=forcing, doy=doy,
                      ext=np.zeros((0, nt)), mct_states=mct)
    ref = oracle_run(problem, series=layout.SERIES)
    assert np.isnan(ref["land_rows"]).any() and np.isfinite(ref["land_rows"]).any()
    mct_reach = np.flatnonzero(river.reach_model == layout.REACH_MCT)
    assert np.isnan(ref["reach_out"][mct_reach]).any(), "the test must push a NaN through a musk_mct reach"
    res = run_problem(problem, engine, series=layout.SERIES, reach_states=True, mct_series=layout.MCT_SERIES)
    compare(res, ref, "NaN forcing", series=layout.SERIES)
    for seq in layout.MCT_SERIES:
        assert_close(res["mct_series"][seq], ref["mct_series"][seq], f"NaN forcing: series of {seq}")
    assert_close(res["reach_state_series"], ref["reach_state_series"], "NaN forcing: states.discharge of every step")
    assert_close(res["mct_states"], ref["mct_states"], "NaN forcing: final Courant/Reynolds numbers and water depths")
    res1 = run_problem(problem, engine)
    compare(res1, ref, "NaN forcing (no series)")
    assert_close(res1["mct_states"], ref["mct_states"], "NaN forcing (no series): musk_mct factors")
    check_window_splitting(problem, engine, 11, ("land_rows", "node_rows", "reach_out", "discharge", "mct_states"))


def test_long_reaches_more_than_32_segments(engine):
    ne, nt = 8, 150
    land, states = random_land(ne, seed=5, snowlimit_fraction=0.0)
    forcing, doy = random_forcing(nt, ne, seed=5)
    river, discharge, _ = random_river(ne, 12, seed=5, smax=100, n_ext=0)
    problem = Problem(land=land, states=states, river=river, discharge=discharge, forcing=forcing, doy=doy,
                      ext=np.zeros((0, nt)))
    ref = oracle_run(problem)
    res = run_problem(problem, engine, reach_states=True)
    compare(res, ref, "long reaches")
    assert int(np.diff(river.seg_ptr).max()) > 33
    # musk_classic states.discharge of every step, given the engine's own inflow: recompute with the oracle from the
    # engine's land rows so that the comparison is bit for bit (routing has no libm call)
    dis = discharge.copy()
    dser = np.zeros((nt, river.n_seg))
    oracle.river_run(river, dis, res["land_rows"], np.zeros((0, nt)), dis_series=dser)
    assert np.array_equal(res["reach_state_series"], dser)


def test_synthetic_config4_slice_vs_oracle(engine):
    """A 2048-element slice of BASELINE config 4 (hourly, river tree) for 3 days against the oracle."""
    ne, nt = 2048, 72
    land, states = synth.make_land(ne, "1h", variants=True)
    forcing, doy = synth.make_forcing(nt, "1h")
    river, discharge = synth.make_tree_river(ne, "1h", levels=40)
    problem = Problem(land=land, states=states, river=river, discharge=discharge, forcing=forcing, doy=doy,
                      ext=np.zeros((0, nt)))
    ref = oracle_run(problem)
    res = run_problem(problem, engine)
    compare(res, ref, "config-4 slice")


@pytest.mark.parametrize("model", ["hland_96p", "hland_96c"])
def test_synthetic_sibling_basin_vs_oracle(engine, model):
    """A 2048-element slice of the config-4 basin with hland_96p / hland_96c + rconc_nash elements: on
    the 'auto' path they run through the sibling instances of the two-kernel fast path."""
    ne, nt = 2048, 72
    land, states = synth.make_land(ne, "1h", variants=True, model=model)
    forcing, doy = synth.make_forcing(nt, "1h")
    river, discharge = synth.make_tree_river(ne, "1h", levels=40)
    problem = Problem(land=land, states=states, river=river, discharge=discharge, forcing=forcing, doy=doy,
                      ext=np.zeros((0, nt)))
    ref = oracle_run(problem)
    res = run_problem(problem, engine)
    compare(res, ref, f"{model} slice")
    assert np.ptp(ref["land_rows"]) > 0.1 and np.isfinite(ref["land_rows"]).all()
    res2 = run_problem(problem, engine, window=25)
    assert np.array_equal(res["node_rows"], res2["node_rows"])


def test_synthetic_capillary_flux_vs_oracle(engine):
    """CFlux > 0 couples the zones to the element's UZ of the previous step: fused kernel on the 'auto' path."""
    ne, nt = 1500, 96
    land, states = synth.make_land(ne, "1h", variants=True, cflux=0.02)
    forcing, doy = synth.make_forcing(nt, "1h")
    river, discharge = synth.make_tree_river(ne, "1h", levels=30)
    problem = Problem(land=land, states=states, river=river, discharge=discharge, forcing=forcing, doy=doy,
                      ext=np.zeros((0, nt)))
    ref = oracle_run(problem)
    assert np.any(ref["states"].sm != oracle_run(Problem(
        land=synth.make_land(ne, "1h", variants=True)[0], states=states, river=river, discharge=discharge,
        forcing=forcing, doy=doy, ext=np.zeros((0, nt))))["states"].sm), "the capillary flux has no effect in this test"
    res = run_problem(problem, engine)
    compare(res, ref, "capillary flux")
    res2 = run_problem(problem, engine, window=29)
    for key in ("land_rows", "node_rows", "discharge"):
        assert np.array_equal(res[key], res2[key]), f"window splitting changes {key}"


def test_synthetic_config3_daily_slice_vs_oracle(engine):
    ne, nt = 512, 40
    land, states = synth.make_land(ne, "1d")
    forcing, doy = synth.make_forcing(nt, "1d")
    from hydpy_b200.abi import RiverBundle

    river = RiverBundle.empty(0)
    problem = Problem(land=land, states=states, river=river, discharge=np.zeros(0), forcing=forcing, doy=doy,
                      ext=np.zeros((0, nt)))
    st = states.copy()
    qt, _ = oracle.land_run(land, st, forcing, doy, threads=4)
    res = run_problem(problem, engine)
    assert_close(res["land_rows"], qt, "config-3 slice: land outflow")
    for n in st.NAMES:
        assert_close(getattr(res["states"], n), getattr(st, n), f"config-3 slice: final {n}")


def test_full_size_properties(engine):
    """At BASELINE size (10^5 elements) the oracle is too slow; check size-independent properties instead:
    (1) the first 1024 elements of the big run equal a stand-alone run of those elements bit for bit (elements are
    independent), (2) elements sharing parameters and forcing column give identical results."""
    ne, nt = 100_000, 24
    land, states = synth.make_land(ne, "1h")
    forcing, doy = synth.make_forcing(nt, "1h")
    from hydpy_b200.abi import RiverBundle

    empty = RiverBundle.empty(0)
    big = run_problem(Problem(land=land, states=states, river=empty, discharge=np.zeros(0), forcing=forcing, doy=doy,
                              ext=np.zeros((0, nt))), engine)
    sub_land, sub_states = synth.make_land(ne, "1h")
    n = 1024
    z = int(sub_land.zone_ptr[n])
    u = int(sub_land.uh_ptr[n])
    from hydpy_b200.abi import LandBundle, LandStates

    small_land = LandBundle(
        zone_ptr=sub_land.zone_ptr[:n + 1], snow_ptr=sub_land.snow_ptr[:n + 1], sfdist_ptr=sub_land.sfdist_ptr[:n + 1],
        uh_ptr=sub_land.uh_ptr[:n + 1], sred_ptr=sub_land.sred_ptr[:n + 1], recstep=sub_land.recstep[:n],
        eflags=sub_land.eflags[:n], forcing_col=sub_land.forcing_col[:n], outlet_node=sub_land.outlet_node[:n],
        epar=sub_land.epar[:, :n], zonetype=sub_land.zonetype[:z], zflags=sub_land.zflags[:z], zorder=sub_land.zorder[:z],
        zpar=sub_land.zpar[:, :z], sfdist=sub_land.sfdist[:n], uh=sub_land.uh[:u], sred_from=sub_land.sred_from,
        sred_to=sub_land.sred_to, sred_adjust=sub_land.sred_adjust)
    small_states = LandStates(ic=sub_states.ic[:z], sm=sub_states.sm[:z], sp=sub_states.sp[:z], wc=sub_states.wc[:z],
                              uz=sub_states.uz[:n], lz=sub_states.lz[:n], quh=sub_states.quh[:u])
    small = run_problem(Problem(land=small_land, states=small_states, river=empty, discharge=np.zeros(0),
                                forcing=forcing, doy=doy, ext=np.zeros((0, nt))), engine)
    assert np.array_equal(big["land_rows"][:n], small["land_rows"])
    assert np.array_equal(big["states"].sm[:z], small["states"].sm)
    assert np.all(np.isfinite(big["land_rows"]))
    # oracle spot check on 64 elements of the big run
    st = small_states.copy()
    qt, _ = oracle.land_run(small_land, st, forcing, doy, threads=4)
    assert_close(small["land_rows"], qt, "full-size spot check")


def test_ensemble_members_share_forcing_and_member0_equals_base(engine):
    ne, nt, m = 64, 48, 5
    land, states = synth.make_land(ne, "1h", bank=ne)
    forcing, doy = synth.make_forcing(nt, "1h", bank=ne)
    river, discharge = synth.make_tree_river(ne, "1h", levels=8)
    base = run_problem(Problem(land=land, states=states, river=river, discharge=discharge, forcing=forcing, doy=doy,
                               ext=np.zeros((0, nt))), engine)
    eland, estates = synth.replicate_members(land, states, m, n_node=river.n_node)
    eriver, edis = synth.replicate_river(river, discharge, m, ne)
    ens = run_problem(Problem(land=eland, states=estates, river=eriver, discharge=edis, forcing=forcing, doy=doy,
                              ext=np.zeros((0, nt))), engine)
    assert np.array_equal(ens["node_rows"][:river.n_node], base["node_rows"])
    assert not np.array_equal(ens["node_rows"][river.n_node:2 * river.n_node], base["node_rows"])
    ref = oracle_run(Problem(land=eland, states=estates, river=eriver, discharge=edis, forcing=forcing, doy=doy,
                             ext=np.zeros((0, nt))))
    assert_close(ens["node_rows"], ref["node_rows"], "ensemble: node series")


def test_pipelined_windows_equal_synchronous_windows(engine):
    """hb_run_async / hb_get_node_series_async (land of window w+1 overlapping routing and copies of window w) must give
    exactly what the synchronous calls give, for host-staged forcing as well as for a resident forcing block."""
    ne, win, nw = 3000, 24, 5
    land, states = synth.make_land(ne, "1h", variants=True)
    river, discharge = synth.make_tree_river(ne, "1h", levels=30)
    forcing, doy = synth.make_forcing(win * nw, "1h")
    no_ext = np.zeros((0, win))

    def setup():
        engine.set_land(land)
        engine.set_land_states(states)
        engine.set_river(river)
        engine.set_river_states(discharge)
        engine.select_series([])

    setup()
    sync_rows = []
    for w in range(nw):
        engine.set_forcing(forcing[:, w * win:(w + 1) * win], doy[w * win:(w + 1) * win])
        engine.set_external(no_ext)
        engine.run()
        sync_rows.append(engine.get_node_series())
    sync_states, sync_dis = engine.get_land_states(), engine.get_river_states()

    staged = engine.pinned((nw, 4, win, forcing.shape[2]))          # asynchronous staging reads pinned memory
    for w in range(nw):
        staged.array[w] = forcing[:, w * win:(w + 1) * win]
    for resident in (False, True, "async"):
        setup()
        bufs = [engine.pinned((river.n_node, win)) for _ in range(nw)]
        if resident is True:
            engine.set_forcing(forcing, doy)
        engine.timing()                      # reset the sums
        for w in range(nw):
            if resident is True:
                engine.select_window(w * win, win)
            elif resident == "async":
                engine.set_forcing(staged.array[w], doy[w * win:(w + 1) * win], asynchronous=True)
            else:
                engine.set_forcing(forcing[:, w * win:(w + 1) * win], doy[w * win:(w + 1) * win])
            engine.set_external(no_ext)
            engine.run_async()
            engine.get_node_series_async(bufs[w].array)
        engine.wait()
        for w in range(nw):
            assert np.array_equal(bufs[w].array, sync_rows[w]), f"window {w} (resident={resident})"
        st, dis = engine.get_land_states(), engine.get_river_states()
        for n in st.NAMES:
            assert np.array_equal(getattr(st, n), getattr(sync_states, n)), n
        assert np.array_equal(dis, sync_dis)
        tm = engine.timing()
        assert tm["n_runs"] == nw and tm["land_ms"] > 0 and tm["river_ms"] > 0
        for b in bufs:
            b.free()
    staged.free()


def test_objective_statistics_of_an_ensemble_on_the_device(engine):
    """hb_node_statistics: nse/kge of every member against the same gauge observations, reduced on the GPU, equal the
    NumPy formulas of hydpy/auxs/statstools.py evaluated on the downloaded series."""
    from hydpy_b200 import calib

    ne, nt, m = 48, 240, 6
    land, states = synth.make_land(ne, "1h", bank=ne)
    forcing, doy = synth.make_forcing(nt, "1h", bank=ne)
    river, discharge = synth.make_tree_river(ne, "1h", levels=6)
    eland, estates = synth.replicate_members(land, states, m, n_node=river.n_node)
    eriver, edis = synth.replicate_river(river, discharge, m, ne)
    res = run_problem(Problem(land=eland, states=estates, river=eriver, discharge=edis, forcing=forcing, doy=doy,
                              ext=np.zeros((0, nt))), engine)
    gauges = np.array([0, 3, 7])                              # nodes of the base network that carry observations
    rng = np.random.default_rng(3)
    obs = res["node_rows"][gauges] * rng.uniform(0.8, 1.2, size=(len(gauges), nt))   # "observed" = perturbed member 0
    obs[1, 50:60] = np.nan
    nodes = np.concatenate([gauges + i * river.n_node for i in range(m)])
    rows = np.tile(np.arange(len(gauges)), m)
    st = engine.node_statistics(nodes, obs, obs_row=rows, skip_nan=True)
    ref = calib.statistics_host(res["node_rows"][nodes], obs[rows], skip_nan=True)
    np.testing.assert_allclose(st, ref, rtol=1e-12, atol=1e-300)
    for fn in (calib.nse, calib.kge, calib.rmse):
        np.testing.assert_allclose(fn(st), fn(ref), rtol=1e-10)
    assert calib.nse(st).reshape(m, -1)[0].min() > 0.5        # member 0 generated the observations
    # without skip_nan the NaN gap propagates like in NumPy
    st2 = engine.node_statistics(nodes[:3], obs, obs_row=rows[:3], skip_nan=False)
    assert np.isnan(calib.nse(st2)[1]) and np.isfinite(calib.nse(st2)[[0, 2]]).all()


def test_full_size_routing_sweeps_agree_bit_for_bit():
    """BASELINE size (10^5 reaches, ≈220 levels): the persistent wavefront (ordered queue, progress flags, ld.cg reads)
    must reproduce the one-launch-per-diagonal sweep exactly — at the size where races would show — and twice in a row
    (pipelined windows).  The land paths must agree within the parity bar on the same basin."""
    ne, win, nw = 100_000, 96, 2
    land, states = synth.make_land(ne, "1h")
    river, discharge = synth.make_tree_river(ne, "1h", levels=200)
    forcing, doy = synth.make_forcing(win * nw, "1h")
    no_ext = np.zeros((0, win))
    results = {}
    with Engine(0) as eng:
        for sweep, path in (("launches", "auto"), ("persistent", "auto"), ("persistent", "general"), ("grouped", "auto")):
            eng.set_land_path(path)
            eng.set_river_sweep(sweep)
            eng.set_land(land); eng.set_land_states(states); eng.set_river(river); eng.set_river_states(discharge)
            bufs = [eng.pinned((river.n_node, win)) for _ in range(nw)]
            for w in range(nw):
                eng.set_forcing(forcing[:, w * win:(w + 1) * win], doy[w * win:(w + 1) * win])
                eng.set_external(no_ext)
                eng.run_async()
                eng.get_node_series_async(bufs[w].array)
            eng.wait()
            results[(sweep, path)] = (np.concatenate([b.array for b in bufs], axis=1), eng.get_river_states())
            for b in bufs:
                b.free()
    a, b, c = results[("launches", "auto")], results[("persistent", "auto")], results[("persistent", "general")]
    g = results[("grouped", "auto")]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[0], g[0]) and np.array_equal(a[1], g[1]), "grouped wavefront differs from the diagonal sweep"
    assert np.isfinite(a[0]).all() and a[0][0].max() > 0.0
    assert_close(b[0], c[0], "full size: fast vs general land path, node series")


@pytest.mark.parametrize("sweep", ["persistent", "grouped"])
def test_plain_routing_levels_give_the_same_bits(sweep):
    """HB_OPT_RIVER_PLAIN_MIN: upstream levels routed by plain per-level launches (one thread per reach) instead of wavefront
    tasks.  Random networks (deploy modes, external rows, reaches with 0 and with more than 32 segments, musk_mct reaches
    that break the run of plain levels) and a river tree: every setting gives the bits of the one-launch-per-diagonal sweep,
    over two windows."""
    cases = []
    ne = 300
    land, states = random_land(ne, seed=7)
    forcing, doy = random_forcing(64, ne, seed=7)
    river, discharge, _ = random_river(ne, 260, seed=7)
    rng = np.random.default_rng(7)
    ext = rng.uniform(0, 3, size=(river.n_ext, 64))
    cases.append(Problem(land=land, states=states, river=river, discharge=discharge, forcing=forcing, doy=doy, ext=ext))
    river2, discharge2, _ = random_river(ne, 260, seed=8)
    mct2 = add_mct(river2, seed=8, fraction=0.1)
    cases.append(Problem(land=land, states=states, river=river2, discharge=discharge2, forcing=forcing, doy=doy,
                         ext=rng.uniform(0, 3, size=(river2.n_ext, 64)), mct_states=mct2))
    ne3 = 3000
    land3, states3 = synth.make_land(ne3, "1h")
    river3, discharge3 = synth.make_tree_river(ne3, "1h", levels=40)
    forcing3, doy3 = synth.make_forcing(64, "1h")
    cases.append(Problem(land=land3, states=states3, river=river3, discharge=discharge3, forcing=forcing3, doy=doy3,
                         ext=np.zeros((0, 64))))
    with Engine(0) as eng:
        for i, problem in enumerate(cases):
            eng.set_river_sweep("launches")
            ref = run_problem(problem, eng, window=32)
            eng.set_river_sweep(sweep)
            for plain_min in (0, 1, 25, 10**9):
                eng.set_river_plain_min(plain_min)
                res = run_problem(problem, eng, window=32)
                for key in ("node_rows", "reach_out", "reach_in", "discharge"):
                    assert np.array_equal(res[key], ref[key], equal_nan=True), (i, plain_min, key)
                if problem.mct_states is not None:
                    assert np.array_equal(res["mct_states"], ref["mct_states"], equal_nan=True), (i, plain_min)


def test_element_slabs_and_time_chunks_do_not_change_a_bit():
    """HB_OPT_LAND_SLABS / HB_OPT_LAND_CHUNK only decide which stream and launch computes what: every combination must
    give the bits of the single-slab, single-chunk run — pipelined windows included (slabs of consecutive windows overlap)."""
    ne, win, nw = 5000, 36, 4
    land, states = synth.make_land(ne, "1h", variants=True)
    river, discharge = synth.make_tree_river(ne, "1h", levels=25)
    forcing, doy = synth.make_forcing(win * nw, "1h")
    no_ext = np.zeros((0, win))
    results = []
    with Engine(0) as eng:
        for slabs, chunk in ((1, 0), (2, 0), (3, 7), (8, 5), (5, 36)):
            eng.set_land(land)
            eng.set_land_states(states)
            eng.set_river(river)
            eng.set_river_states(discharge)
            eng.set_land_slabs(slabs)
            eng.set_land_chunk(chunk)
            eng.set_forcing(forcing, doy)
            bufs = [eng.pinned((river.n_node, win)) for _ in range(nw)]
            for w in range(nw):
                eng.select_window(w * win, win)
                eng.set_external(no_ext)
                eng.run_async()
                eng.get_node_series_async(bufs[w].array)
            eng.wait()
            st = eng.get_land_states()
            results.append(([b.array.copy() for b in bufs], st, eng.get_river_states()))
            for b in bufs:
                b.free()
    rows0, st0, dis0 = results[0]
    assert np.isfinite(rows0[-1]).all() and rows0[-1].max() > 0.0
    for rows, st, dis in results[1:]:
        for a, b in zip(rows0, rows):
            assert np.array_equal(a, b)
        assert np.array_equal(dis0, dis)
        for n in st0.NAMES:
            assert np.array_equal(getattr(st0, n), getattr(st, n)), n


@pytest.mark.parametrize("case", ["lahn", "variants"])
def test_members_built_on_the_device_equal_the_replicated_bundle(case):
    """hb_set_members (constants folded on the device from the base basin + per-member parameter modifications, river and
    conditions replicated on the device) must give what the same ensemble gives when every member is uploaded as
    additional elements through hb_set_land: bit for bit, states included."""
    from tests.util import load_scenarios

    members = 7
    if case == "lahn":
        p = load_scenarios("lahn_hourly_members.npz")["member0"].problem
        land, states, river, discharge, forcing, doy = p.land, p.states, p.river, p.discharge, p.forcing, p.doy
    else:
        land, states = synth.make_land(300, "1h", variants=True, seed=9)
        river, discharge = synth.make_tree_river(300, "1h", levels=9, seed=10)
        discharge = np.random.default_rng(3).uniform(0, 2, river.n_seg)
        forcing, doy = synth.make_forcing(72, "1h")
    nt = forcing.shape[1]
    eland, estates = synth.replicate_members(land, states, members, n_node=river.n_node)
    eriver, edis = synth.replicate_river(river, discharge, members, land.n_elem)
    with Engine(0) as eng:
        ref = run_problem(Problem(land=eland, states=estates, river=eriver, discharge=edis, forcing=forcing, doy=doy,
                                  ext=np.zeros((0, nt))), eng)
        eng.set_land(land)
        eng.set_land_states(states)
        eng.set_river(river)
        eng.set_river_states(discharge)
        eng.set_members(members, synth.member_mods(members))
        eng.set_forcing(forcing, doy)
        eng.set_external(np.zeros((0, nt)))
        eng.run()
        rows = eng.get_node_series()
        st = eng.get_land_states()
        dis = eng.get_river_states()
        assert rows.shape == (members * river.n_node, nt)
        assert np.array_equal(rows, ref["node_rows"])
        assert np.array_equal(dis, ref["discharge"])
        for n in ("ic", "sm", "sp", "wc", "uz", "lz", "quh"):
            assert np.array_equal(getattr(st, n), getattr(ref["states"], n)), n
        assert not np.array_equal(rows[:river.n_node], rows[river.n_node:2 * river.n_node])
        # a second window continues every member from its own conditions; going back to one basin keeps member 0
        eng.set_forcing(forcing, doy)
        eng.set_external(np.zeros((0, nt)))
        eng.run()
        rows2 = eng.get_node_series()
        sm2 = eng.get_land_states().sm
        eng.set_members(1)
        assert np.array_equal(eng.get_land_states().sm, sm2[:land.n_zone])
        eng.set_forcing(forcing, doy)
        eng.set_external(np.zeros((0, nt)))
        eng.run()
        assert eng.get_node_series().shape == (river.n_node, nt)
    ref2_states = ref["states"]
    with Engine(0) as eng:   # the replicated bundle, second window
        eng.set_land(eland); eng.set_land_states(ref2_states); eng.set_river(eriver); eng.set_river_states(ref["discharge"])
        eng.set_forcing(forcing, doy); eng.set_external(np.zeros((0, nt))); eng.run()
        assert np.array_equal(eng.get_node_series(), rows2)


def test_sixty_four_parameter_sets_in_one_run_equal_sequential_oracle_runs():
    """The batched calibration caller: `calib.evaluate` — M = 64 parameter sets as one ensemble on the
    device, objective reduction on the device, windows merged — against 64 sequential oracle runs judged with NumPy:
    NSE and KGE within the parity bar."""
    from hydpy_b200 import calib
    from tests.util import load_scenarios

    p = load_scenarios("lahn_hourly_members.npz")["member0"].problem
    members, nt = 64, p.nt
    mods = synth.member_mods(members, seed=7, jitter=0.3)
    mask = np.array([True, False, True, True])                        # one more rule, bound to three of the four elements
    entries = [(n, m, a, None) for n, (m, a) in mods.items()] + [("icmax", np.linspace(0.5, 1.5, members), np.zeros(members), mask)]
    nodes = [p.river.n_node - 1, 0]
    base = oracle_run(p)
    rng = np.random.default_rng(1)
    obs = base["node_rows"][nodes] * rng.uniform(0.8, 1.25, size=(len(nodes), nt))
    with Engine(0) as eng:
        got_nse = calib.evaluate(p, members, entries, nodes, obs, criterion="nse", engine=eng, window=100)
        got_kge = calib.evaluate(p, members, entries, nodes, obs, criterion="kge", engine=eng)
    ref_nse, ref_kge = np.zeros((members, len(nodes))), np.zeros((members, len(nodes)))
    for m in range(members):
        land_m = calib.apply_mods(p.land, entries, m)
        r = oracle_run(Problem(land=land_m, states=p.states, river=p.river, discharge=p.discharge, forcing=p.forcing,
                               doy=p.doy, ext=p.ext, mct_states=p.mct_states))
        st = calib.statistics_host(r["node_rows"][nodes], obs)
        ref_nse[m], ref_kge[m] = calib.nse(st), calib.kge(st)
    assert_close(got_nse, ref_nse, "NSE of 64 parameter sets")
    assert_close(got_kge, ref_kge, "KGE of 64 parameter sets")
    assert np.ptp(ref_nse[:, 0]) > 1e-3            # the sets really differ


@pytest.mark.parametrize("zones", [1, 2, 3, 5, 7, 10])
def test_fused_zone_kernel_instances_for_small_zone_counts(zones):
    """`zone_kernel<…, ZW>` exists for blocks of 1, 2, 3, 4, 6, 8 and 12 zone warps (the smallest that holds the basin's
    largest zone count): every instance against the term buffers (bit for bit) and the oracle."""
    from hydpy_b200.abi import RiverBundle

    ne, nt = 700, 40
    land, states = synth.make_land(ne, "1h", zones=zones, seed=zones)
    forcing, doy = synth.make_forcing(nt, "1h", t0=4000)
    out = []
    with Engine(0) as eng:
        for fuse in (False, True):
            eng.set_land(land)
            eng.set_land_states(states)
            eng.set_river(RiverBundle.empty(0))
            eng.set_land_fuse(fuse)
            eng.set_forcing(forcing, doy)
            eng.run(land=True, river=False)
            tm = eng.timing()
            assert tm["n_fused"] == (ne if fuse else 0)
            out.append((eng.get_land_outflow(), eng.get_land_states()))
    (q0, s0), (q1, s1) = out
    assert np.array_equal(q0, q1)
    for n in s0.NAMES:
        assert np.array_equal(getattr(s0, n), getattr(s1, n)), n
    st = states.copy()
    qt, _ = oracle.land_run(land, st, forcing, doy, threads=4)
    assert_close(q1, qt, f"{zones} zones: land outflow")


@pytest.mark.parametrize("case", ["synthetic", "ragged"])
def test_zone_terms_added_inside_the_zone_kernel_give_the_same_bits(case):
    """HB_OPT_LAND_FUSE: groups of 32 consecutive elements whose zones all run the branch-free step have their zone terms
    added inside the thread-per-zone kernel (in zone order) instead of the thread-per-element kernel.  Same bits as the
    hand-over through the term buffers, for mixed basins (fused and other groups side by side), ragged zone counts,
    element slabs and time chunks; and the same values as the oracle."""
    if case == "synthetic":
        ne, nt = 4200, 60
        land, states = synth.make_land(ne, "1h", variants=True, seed=5)
        forcing, doy = synth.make_forcing(nt, "1h", t0=2000)
    else:
        ne, nt = 700, 48
        land, states = random_land(ne, seed=11, zmax=12, bank=64, soil_runs=70)
        forcing, doy = random_forcing(nt, 64, seed=3)
    from hydpy_b200.abi import RiverBundle

    out = []
    with Engine(0) as eng:
        for fuse, slabs, chunk in ((False, 1, 0), (True, 1, 0), (True, 3, 7)):
            eng.set_land(land)
            eng.set_land_states(states)
            eng.set_river(RiverBundle.empty(0))
            eng.set_land_fuse(fuse)
            eng.set_land_slabs(slabs)
            eng.set_land_chunk(chunk)
            eng.set_forcing(forcing, doy)
            eng.run(land=True, river=False)
            out.append((eng.get_land_outflow(), eng.get_land_states(), eng.timing()))
    q0, s0, _ = out[0]
    assert np.isfinite(q0).all() and q0.max() > 0.0
    for q, s, tm in out[1:]:
        assert np.array_equal(q0, q)
        for n in s0.NAMES:
            assert np.array_equal(getattr(s0, n), getattr(s, n)), n
    st = states.copy()
    qt, _ = oracle.land_run(land, st, forcing, doy, threads=4)
    assert_close(q0, qt, f"{case}: land outflow")
    for n in st.NAMES:
        assert_close(getattr(s0, n), getattr(st, n), f"{case}: final {n}")


def test_recstep_chain_with_out_of_range_power_arguments():
    """The short form of the RecStep chain evaluates dt*k*(uz/contriarea)^(1+alpha) as 2^(y*log2(uz) + c2) for exponents
    below 800 in magnitude; beyond that (here: tiny storages under exponents up to 64, huge ones under large exponents)
    and for parameters outside its bounds the element repeats the step with the general loop — same values as the oracle."""
    from hydpy_b200.abi import RiverBundle

    ne, nt = 256, 30
    land, states = synth.make_land(ne, "1h", seed=9)
    EP = layout.EP
    rng = np.random.default_rng(5)
    # (steep responses only with tiny rate constants: dt*k*(uz/ca)^41 near uz/ca = 1 makes the explicit sub-steps unstable —
    # the last bits of ANY power function then decide the result, the reference's included)
    alpha = rng.choice([0.0, 0.5, 1.0, 3.0, 20.0, 40.0, 63.0, 70.0], size=ne)
    land.epar[EP["alpha"]] = alpha
    land.epar[EP["k"]] = np.where(alpha > 3.0, rng.choice([1e-40, 1e-12], size=ne), rng.choice([1e-40, 1e-12, 1e-3, 0.05], size=ne))
    states.uz[:] = rng.choice([0.0, 1e-300, 1e-200, 1e-12, 1e-3, 5.0, 1e4, 1e30], size=ne)
    forcing, doy = synth.make_forcing(nt, "1h", t0=3000)
    # ... and no recharge for them: a steep response under strong recharge settles where (1 + alpha) * q0 > 2 * uz per
    # sub-step, i.e. where the explicit scheme oscillates, whatever the rate constant
    forcing[0][:, np.unique(land.forcing_col[alpha > 3.0])] = 0.0
    states.sp[:] = 0.0
    states.wc[:] = 0.0
    with Engine(0) as eng:
        eng.set_land(land)
        eng.set_land_states(states)
        eng.set_river(RiverBundle.empty(0))
        eng.set_forcing(forcing, doy)
        eng.run(land=True, river=False)
        q, st_e = eng.get_land_outflow(), eng.get_land_states()
        tm = eng.timing()
    assert 0 < tm["elem_steps_run"] < tm["elem_steps"]
    st = states.copy()
    qt, _ = oracle.land_run(land, st, forcing, doy, threads=4)
    assert np.isfinite(qt).all()
    assert_close(q, qt, "extreme response parameters: land outflow")
    for n in st.NAMES:
        assert_close(getattr(st_e, n), getattr(st, n), f"extreme response parameters: final {n}")


@pytest.mark.parametrize("model", ["hland_96p", "hland_96c"])
def test_sibling_series_come_from_the_fast_path(model):
    """Requested series of hland_96p / hland_96c elements (their own zone and element sequences included) are recorded by
    the sibling instances of the two fast-path kernels: the general kernel is not launched, every sequence equals the
    oracle's, sequences of the other sibling model read NaN."""
    from hydpy_b200.abi import RiverBundle

    ne, nt = 1200, 48
    land, states = synth.make_land(ne, "1h", variants=True, model=model, seed=6)
    forcing, doy = synth.make_forcing(nt, "1h", t0=1500)
    st = states.copy()
    qt, ser = oracle.land_run(land, st, forcing, doy, threads=4, series=layout.SERIES)
    with Engine(0) as eng:
        eng.set_land(land)
        eng.set_land_states(states)
        eng.set_river(RiverBundle.empty(0))
        eng.select_series(layout.SERIES)
        eng.set_forcing(forcing, doy)
        eng.run(land=True, river=False)
        tm = eng.timing()
        assert tm["n_fast"] == ne and tm["general_ms"] == 0.0
        assert_close(eng.get_land_outflow(), qt, f"{model} with series: land outflow")
        for name in layout.SERIES:
            assert_close(eng.get_series(name), ser[name], f"{model}: series {name}")
        other = layout.SERIES_96C_ONLY if model == "hland_96p" else layout.SERIES_96P_ONLY
        assert all(np.isnan(eng.get_series(name)).all() for name in other)
        st_e = eng.get_land_states()
        for n in st.NAMES:
            assert_close(getattr(st_e, n), getattr(st, n), f"{model} with series: final {n}")
        eng.select_series([])


def test_snow_classes_run_on_the_fast_path():
    """SClass > 1 (ref: hland_model.py:484-499, 1166-1187, 1318-1341, 1434-1448: the `for c in range(con.sclass)` loops):
    the fast path's general zone step loops over the classes, the general kernel is not launched — unless series are
    recorded, which sends these elements (and only them) back to it.  Same values as the oracle either way."""
    import copy

    from hydpy_b200.abi import RiverBundle

    ne, nt = 1500, 72
    land1, states1 = synth.make_land(ne, "1h", variants=True, seed=3)
    nz = np.diff(land1.zone_ptr)
    ncls = 1 + (np.arange(ne) % 3)                      # 1, 2, 3 classes side by side
    land = copy.deepcopy(land1)
    land.sfdist_ptr = np.concatenate([[0], np.cumsum(ncls)]).astype(np.int64)
    rng = np.random.default_rng(9)
    land.sfdist = np.concatenate([(lambda w: w / w.mean())(rng.uniform(0.5, 1.5, size=c)) for c in ncls])
    land.snow_ptr = np.concatenate([[0], np.cumsum(nz * ncls)]).astype(np.int64)
    land.normalise()
    states = states1.copy()
    ns = int(land.snow_ptr[-1])
    states.sp = rng.uniform(0, 20, ns) * (rng.uniform(size=ns) < 0.6)
    states.wc = rng.uniform(0, 1, ns)
    for e in range(ne):                                 # lake zones carry no snow
        lk = np.tile(land.zonetype[land.zone_ptr[e]:land.zone_ptr[e + 1]] == layout.ILAKE, int(ncls[e]))
        states.sp[land.snow_ptr[e]:land.snow_ptr[e + 1]][lk] = 0.0
        states.wc[land.snow_ptr[e]:land.snow_ptr[e + 1]][lk] = 0.0
    forcing, doy = synth.make_forcing(nt, "1h", t0=24 * 20)        # snowfall, melt and refreezing side by side:
    forcing[1] -= 10.0                                  # around the freezing point
    st = states.copy()
    qt, ser = oracle.land_run(land, st, forcing, doy, threads=4, series=("sp", "wc", "in_"))
    with Engine(0) as eng:
        eng.set_land(land)
        eng.set_land_states(states)
        eng.set_river(RiverBundle.empty(0))
        eng.set_forcing(forcing, doy)
        eng.run(land=True, river=False)
        tm = eng.timing()
        q, st_e = eng.get_land_outflow(), eng.get_land_states()
        assert tm["n_fast"] == ne and tm["general_ms"] == 0.0
        assert_close(q, qt, "snow classes: land outflow")
        for n in st.NAMES:
            assert_close(getattr(st_e, n), getattr(st, n), f"snow classes: final {n}")
        eng.set_land_states(states)
        eng.select_series(["sp", "wc", "in_"])
        eng.run(land=True, river=False)
        tm = eng.timing()
        assert tm["general_ms"] > 0.0                   # the multi-class elements, with their series
        assert_close(eng.get_land_outflow(), qt, "snow classes with series: land outflow")
        for name in ("sp", "wc", "in_"):
            assert_close(eng.get_series(name), ser[name], f"snow classes: series {name}")
        eng.select_series([])


def test_results_do_not_depend_on_where_an_element_sits():
    """The fast path groups 32 consecutive elements into warps and lets warp-wide votes pick between equivalent code paths
    (branch-free or general zone step, dry shortcuts, unswitched RecStep loop).  Shuffling the elements — what sharding a
    basin over several GPUs does — must therefore not change a single bit of any element's outflow or state."""
    from hydpy_b200.abi import RiverBundle
    from hydpy_b200.shard import subset_land

    ne, nt = 2100, 96
    land, states = synth.make_land(ne, "1h", variants=True, seed=21)
    forcing, doy = synth.make_forcing(nt, "1h")
    perm = np.random.default_rng(4).permutation(ne)
    pland, pstates = subset_land(land, states, perm, np.full(ne, -1, np.int32))
    out = []
    with Engine(0) as eng:
        for ld, st in ((land, states), (pland, pstates)):
            eng.set_land(ld)
            eng.set_land_states(st)
            eng.set_river(RiverBundle.empty(0))
            eng.set_forcing(forcing, doy)
            eng.run(land=True, river=False)
            out.append((eng.get_land_outflow(), eng.get_land_states()))
    (q0, s0), (q1, s1) = out
    assert np.array_equal(q0[perm], q1)
    assert np.array_equal(s0.uz[perm], s1.uz) and np.array_equal(s0.lz[perm], s1.lz)
    zone_of = np.concatenate([np.arange(land.zone_ptr[e], land.zone_ptr[e + 1]) for e in perm])
    assert np.array_equal(s0.sm[zone_of], s1.sm) and np.array_equal(s0.ic[zone_of], s1.ic)


def test_finite_smax_runs_the_fast_path_speculatively():
    """Finite SMax without capillary flux (ref: hland_model.py:584-605 Calc_SPL_WCL_SP_WC_V1, 876-969
    Calc_SPG_WCG_SP_WC_V1): while no snow class exceeds its limit both methods leave every state alone, so such elements
    take the fast path and its zone kernel watches the limit (`HB_OPT_LAND_SPECULATE`).  Elements whose snow exceeds it
    somewhere in a window repeat that window in the general kernel from the window's initial conditions (and skip the
    speculation of the next window if the general kernel saw a loss).  Two consecutive
    windows (conditions carried on the device) against the oracle, limits of 5-40 mm, three times that and out of reach
    side by side under snowfall: exactly the element-windows in which the oracle sees a loss (or saw one in the window
    before) are repeated; with the
    speculation switched off the general kernel runs as in round 1; while series are recorded the elements stay with the
    general kernel."""
    import copy

    from hydpy_b200.abi import RiverBundle

    ne, nt = 900, 60
    land, states0 = random_land(ne, seed=31, snowlimit_fraction=0.85, fast_fraction=0.0, recstep=10)
    ZP = layout.ZP
    land = copy.deepcopy(land)
    land.zpar[ZP["cflux"]] = 0.0                             # (capillary flux would send the elements to other kernels)
    factor = np.random.default_rng(3).choice([1.0, 3.0, 1000.0], size=ne)
    nz = np.diff(land.zone_ptr)
    land.zpar[ZP["smax"]] = land.zpar[ZP["smax"]] * np.repeat(factor, nz)
    forcing, doy = random_forcing(2 * nt, ne, seed=5)
    forcing[1] -= 6.0                                        # snowfall, melt and refreezing side by side
    live = np.isfinite(land.zp("smax")) & (land.zonetype != layout.ILAKE)        # (lake zones carry no snow)
    limited = np.array([live[land.zone_ptr[e]:land.zone_ptr[e + 1]].any() for e in range(ne)])
    windows = [(np.ascontiguousarray(forcing[:, w * nt:(w + 1) * nt]), doy[w * nt:(w + 1) * nt]) for w in range(2)]
    names = ("sp", "wc", "spl", "wcl", "spg", "wcg")

    st_ref = states0.copy()
    q_ref, ser_ref, expected = [], None, 0
    touchy, lost_before = np.zeros(ne, bool), np.zeros(ne, bool)
    for f, d in windows:
        q, ser_ref = oracle.land_run(land, st_ref, f, d, threads=4, series=names)
        q_ref.append(q)
        loss = (ser_ref["spl"] + ser_ref["wcl"]) > 0.0
        lost = np.array([bool(loss[:, land.zone_ptr[e]:land.zone_ptr[e + 1]].any()) for e in range(ne)])
        # repeated: the elements that lose snow in this window and those that lost some in the previous one (they skip
        # the speculation of this window and go straight to the general kernel)
        expected += int((lost | lost_before).sum())
        lost_before = lost
        # a snow pack of less than 1e-9 mm in the REFERENCE's own run (the rounding residue of an interception storage at
        # its capacity, frozen): evap_aet_hbv96's snow-cover indicator (sp > 0) turns it into a third less soil
        # evapotranspiration of a three-class zone, so the last bit of the precipitation decides about 1e-7 of the soil
        # moisture — the fast path's reciprocal multiplications (<= 1 ulp) may fall on the other side
        residue = (ser_ref["sp"] > 0.0) & (ser_ref["sp"] < 1e-9)
        touchy |= np.array([bool(residue[:, land.snow_ptr[e]:land.snow_ptr[e + 1]].any()) for e in range(ne)])
    q_ref = np.concatenate(q_ref, axis=1)
    assert 600 < limited.sum() < ne and 300 < expected < 2 * limited.sum() - 300     # both outcomes well represented
    assert touchy.sum() < 0.06 * ne
    firm = ~touchy
    zone_firm, snow_firm = np.repeat(firm, nz), np.repeat(firm, np.diff(land.snow_ptr))
    uh_firm = np.repeat(firm, np.diff(land.uh_ptr))

    def engine_two_windows(eng, series=()):
        eng.set_land(land)
        eng.set_land_states(states0)
        eng.set_river(RiverBundle.empty(0))
        out = []
        eng.set_forcing(*windows[0])
        eng.run(land=True, river=False)
        out.append(eng.get_land_outflow())
        eng.select_series(list(series))
        eng.set_forcing(*windows[1])
        eng.run(land=True, river=False)
        out.append(eng.get_land_outflow())
        ser = {name: eng.get_series(name) for name in series}
        eng.select_series([])
        return np.concatenate(out, axis=1), eng.get_land_states(), eng.timing(), ser

    def check(what, q, st):
        assert_close(q[firm], q_ref[firm], f"{what}: land outflow of both windows")
        assert_close(q[touchy], q_ref[touchy], f"{what}: land outflow of both windows (residual snow)", rtol=1e-3)
        for n in st_ref.NAMES:
            a, b = getattr(st, n), getattr(st_ref, n)
            mask = {ne: firm, len(zone_firm): zone_firm, len(snow_firm): snow_firm, len(uh_firm): uh_firm}.get(len(b))
            if mask is None or len(b) == 0:
                continue
            assert_close(a[mask], b[mask], f"{what}: final {n}")
            assert_close(a[~mask], b[~mask], f"{what}: final {n} (residual snow)", rtol=1e-3, atol=1e-6)

    with Engine(0) as eng:
        q, st, tm, _ = engine_two_windows(eng)
        assert tm["n_spec"] == limited.sum() and tm["n_fast"] == ne, tm
        assert tm["spec_reruns"] == expected, (tm["spec_reruns"], expected)
        check("speculative fast path", q, st)
        # series recorded in the second window: the general kernel takes the limited elements there, with their series
        q, st, tm, ser = engine_two_windows(eng, series=names)
        assert tm["n_spec"] == 0
        check("series in window 2", q, st)
        for name in names:
            mask = zone_firm if ser_ref[name].shape[1] == len(zone_firm) else snow_firm
            assert_close(ser[name][:, mask], ser_ref[name][:, mask], f"series in window 2: {name}")
    with Engine(0) as eng:
        eng.set_land_speculate(False)
        q, st, tm, _ = engine_two_windows(eng)
        assert tm["n_spec"] == 0 and tm["spec_reruns"] == 0 and tm["n_fast"] == ne - limited.sum(), tm
        check("speculation off", q, st)
