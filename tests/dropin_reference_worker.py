"""Worker of tests/test_dropin_reference.py (own process: HydPy keeps global state in `hydpy.pub`).

The drop-in `hydpy_b200.simulate(hp)` on REAL objects of the unmodified reference (built by oracle/build_ref.py in
this container): extract → run → write_back must leave `hp` exactly as the reference's own `hp.simulate()` leaves it.
Without `--engine` (no GPU: the build container) the engine call in the middle is replaced by the CPU oracle (bit-exact
with the reference), which makes the comparison a bit-for-bit check of the host-side layers.  With `--engine` (GPU box,
tests/test_gpu_dropin.py) nothing is replaced: extract → CUDA engine → write_back runs as one piece and every array must
agree with the reference's own `hp.simulate()` within the parity bar of tests/util.py (rtol 1e-9, atol 1e-12)."""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import build_ref, oracle  # noqa: E402

_paths = build_ref.import_paths()
if _paths is None:
    raise SystemExit("no importable reference: run oracle/build_ref.py where /root/reference exists")
SRC, STUBS = _paths
sys.path[:0] = [STUBS, SRC]
ENGINE = "--engine" in sys.argv

import hydpy  # noqa: E402
from hydpy import pub  # noqa: E402
from hydpy.core import testtools  # noqa: E402

import hydpy_b200  # noqa: E402

simulate_module = sys.modules["hydpy_b200.simulate"]   # (the package exports the function under the same name)


def oracle_run_problem(problem, engine=None, *, series=(), window=None, device=0, reach_states=False, mct_series=()):
    """Same contract as hydpy_b200.run_problem, executed by the CPU oracle."""
    st, dis = problem.states.copy(), problem.discharge.copy()
    qt, ser = oracle.land_run(problem.land, st, problem.forcing, problem.doy, series=list(series), threads=2)
    dser = np.zeros((problem.nt, problem.river.n_seg)) if reach_states else None
    mct = None if problem.mct_states is None else problem.mct_states.copy()
    mct_series = list(mct_series) if problem.river.any_mct else []
    mser = np.full((len(hydpy_b200.layout.MCT_SERIES), problem.nt, problem.river.n_seg), np.nan) if mct_series else None
    node_rows, rout, rin = oracle.river_run(problem.river, dis, qt, problem.ext, dis_series=dser, mct_states=mct,
                                            mct_series=mser)
    out = dict(states=st, discharge=dis, node_rows=node_rows, land_rows=qt, reach_out=rout, reach_in=rin, series=ser)
    if reach_states:
        out["reach_state_series"] = dser
    if mct is not None:
        out["mct_states"] = mct
    if mct_series:
        out["mct_series"] = {name: mser[hydpy_b200.layout.MS[name]] for name in mct_series}
    return out


def snapshot(hp):
    out = {}
    for node in hp.nodes:
        out[f"node:{node.name}"] = node.sequences.sim.series.copy()
        out[f"nodevalue:{node.name}"] = np.array(node.sequences.sim.value)
    for element in hp.elements:
        seqs = element.model.sequences
        for group in ("inputs", "factors", "fluxes", "states", "outlets"):
            for seq in getattr(seqs, group, ()):
                if getattr(seq, "ramflag", False):
                    out[f"{element.name}:{group}.{seq.name}:series"] = np.array(seq.series)
                if group == "states":
                    out[f"{element.name}:states.{seq.name}:new"] = np.array(seq.new)
                    out[f"{element.name}:states.{seq.name}:old"] = np.array(seq.old)
        # sequences of the submodels (`prepare_allseries()` keeps them in RAM as well; ref: modelutils.py:1127-1167)
        for path in ("aetmodel", "aetmodel.petmodel", "rconcmodel"):
            sub = element.model
            for part in path.split("."):
                sub = getattr(sub, part, None)
            if sub is None:
                continue
            for group in ("factors", "fluxes", "states"):
                for seq in getattr(sub.sequences, group, ()) or ():
                    if getattr(seq, "ramflag", False):
                        out[f"{element.name}:{path}.{group}.{seq.name}:series"] = np.array(seq.series)
        rwd = getattr(getattr(seqs, "factors", None), "referencewaterdepth", None)
        if rwd is not None:
            out[f"{element.name}:referencewaterdepth"] = np.array(rwd.value)
        pet = getattr(getattr(element.model, "aetmodel", None), "petmodel", None)
        if pet is not None:
            for seq in pet.sequences.inputs:
                if getattr(seq, "ramflag", False):
                    out[f"{element.name}:pet.inputs.{seq.name}:series"] = np.array(seq.series)
        rc = getattr(element.model, "rconcmodel", None)
        if rc is not None and rc.name == "rconc_uh":
            out[f"{element.name}:quh"] = np.array(rc.sequences.logs.quh.value)
        elif rc is not None:
            out[f"{element.name}:sc:new"] = np.array(rc.sequences.states.sc.new)
            out[f"{element.name}:sc:old"] = np.array(rc.sequences.states.sc.old)
    return out


def sibling_project():
    """A small basin of the sibling models built through the reference's own API: one hland_96p and one
    hland_96c element (four zones each, every zone type but ILAKE/…, rconc_nash) draining through a musk_classic reach."""
    from hydpy import Element, Elements, HydPy, Node, Nodes

    pub.timegrids = "2000-01-01", "2000-01-09", "1h"
    n_up, n_out, n_end = Node("sib_up"), Node("sib_out"), Node("sib_end")
    e_p = Element("sib_land_p", outlets=n_up)
    e_c = Element("sib_land_c", outlets=n_up)
    e_r = Element("sib_reach", inlets=n_up, outlets=n_out)
    e_m = Element("sib_reach_mct", inlets=n_out, outlets=n_end)
    rng = np.random.default_rng(11)
    common = """
parameterstep("1h")
area(12.0)
nmbzones(5)
sclass(2)
zonetype(FIELD, FOREST, GLACIER, ILAKE, SEALED)
zonearea(3.0, 3.0, 2.0, 2.0, 2.0)
psi(1.0)
zonez(2.0, 3.0, 9.0, 1.0, 2.0)
pcorr(1.1); pcalt(0.1); rfcf(1.1); sfcf(1.3); tcorr(0.6); tcalt(0.6); icmax(2.0); sfdist(0.8, 1.2); smax(inf)
sred(0.0); tt(0.0); ttint(2.0); dttm(1.0); cfmax(0.5); cfvar(0.1); gmelt(1.0); gvar(0.2); cfr(0.1); whc(0.4)
fc(200.0); beta(2.0)
{special}
with model.add_aetmodel_v1("evap_aet_hbv96"):
    temperaturethresholdice(0.0)
    soilmoisturelimit(0.8)
    excessreduction(0.5)
    with model.add_petmodel_v1("evap_pet_hbv96"):
        evapotranspirationfactor(0.7)
        airtemperaturefactor(0.1)
        altitudefactor(-0.1)
        precipitationfactor(0.1)
with model.add_rconcmodel_v1("rconc_nash"):
    retentiontime(3.0)
    nmbstorages(4)
    nmbsteps(20)
"""
    special = {
        "hland_96p": "percmax(2.0); sgr(20.0); sg1max(50.0); k0(2.0); k1(10.0); k2(20.0); k3(100.0)",
        "hland_96c": "h1(10.0); tab1(2.0); tvs1(2.0); h2(10.0); tab2(10.0); tvs2(10.0); k4(0.005); gamma(0.0)",
    }
    for element, name in ((e_p, "hland_96p"), (e_c, "hland_96c")):
        ns: dict = {}
        exec(f"from hydpy.models.{name} import *\n" + common.format(special=special[name]), ns)   # like a control file
        element.model = ns["model"]
    ns = {}
    exec('from hydpy.models.musk_classic import *\nparameterstep("1h")\nnmbsegments(3)\ncoefficients(damp=0.2)\n', ns)
    reach = ns["model"]
    e_r.model = reach
    # a musk_mct reach with a two-trapeze cross section downstream (ref: hydpy/models/musk_mct.py:40-72)
    ns = {}
    exec("""from hydpy.models.musk_mct import *
parameterstep("1h")
nmbsegments(4)
catchmentarea(2500.0)
length(20.0)
bottomslope(0.0005)
with model.add_wqmodel_v1("wq_trapeze_strickler"):
    nmbtrapezes(2)
    bottomlevels(0.0, 1.5)
    bottomwidths(4.0, 10.0)
    sideslopes(2.0, 4.0)
    stricklercoefficients(30.0, 25.0)
    calibrationfactors(1.0, 1.0)
""", ns)
    e_m.model = ns["model"]
    hp = HydPy("sibling_project")
    hp.update_devices(nodes=Nodes(n_up, n_out, n_end), elements=Elements(e_p, e_c, e_r, e_m))
    hp.update_parameters()
    hp.prepare_allseries()                                    # musk_mct included: every factor, flux and state series
    nt = len(pub.timegrids.init)
    for element in (e_p, e_c):
        model = element.model
        t = 2.0 + 8.0 * np.sin(np.arange(nt) / 24.0) + rng.normal(0.0, 2.0, nt)
        model.sequences.inputs.p.series = rng.gamma(0.4, 2.0, nt) * (rng.uniform(size=nt) < 0.5)
        model.sequences.inputs.t.series = t
        pet = model.aetmodel.petmodel.sequences.inputs
        pet.normalairtemperature.series = t - 1.0
        pet.normalevapotranspiration.series = rng.uniform(0.0, 0.4, nt)
        sta = model.sequences.states
        sta.ic(0.5); sta.sp(1.0); sta.wc(0.1); sta.sm(100.0)
        if model.name == "hland_96p":
            sta.suz(1.0); sta.sg1(10.0); sta.sg2(10.0); sta.sg3(10.0)
        else:
            sta.bw1(2.0); sta.bw2(3.0); sta.lz(10.0)
        model.rconcmodel.sequences.states.sc(0.05)
    reach.sequences.states.discharge(0.0)
    e_m.model.sequences.states.discharge(0.5)
    e_m.model.sequences.states.courantnumber(0.5)
    e_m.model.sequences.states.reynoldsnumber(1.0)
    return hp, (("2000-01-01", "2000-01-04"), ("2000-01-04", "2000-01-09"))


def ret_io_project():
    """Two hland_96 elements whose evap_aet_hbv96 submodel takes its potential evapotranspiration from `evap_ret_io`
    (an external time series of reference evapotranspiration, adjusted per zone; ref: hydpy/models/evap_ret_io.py:53-64)
    — one without capillary flux (fast path of the engine), one with (coupled kernel) — on a common outlet node."""
    from hydpy import Element, Elements, HydPy, Node, Nodes

    pub.timegrids = "2000-01-01", "2000-01-09", "1h"
    n_out = Node("retio_out")
    e_a, e_b = Element("retio_land_a", outlets=n_out), Element("retio_land_b", outlets=n_out)
    rng = np.random.default_rng(23)
    control = """
from hydpy.models.hland_96 import *
parameterstep("1h")
area(20.0)
nmbzones(5)
sclass(1)
zonetype(FIELD, FOREST, GLACIER, ILAKE, SEALED)
zonearea(6.0, 5.0, 3.0, 2.0, 4.0)
psi(1.0)
zonez(2.0, 4.0, 11.0, 1.0, 2.0)
pcorr(1.05); pcalt(0.08); rfcf(1.1); sfcf(1.25); tcorr(0.3); tcalt(0.6); icmax(1.5); sfdist(1.0); smax(inf)
sred(0.0); tt(0.2); ttint(2.0); dttm(0.5); cfmax(0.6); cfvar(0.1); gmelt(1.0); gvar(0.2); cfr(0.1); whc(0.3)
fc(180.0); beta(2.3); percmax(0.6); cflux({cflux}); resparea(True); recstep(50); alpha(1.3); k(0.002); k4(0.006)
gamma(0.1)
with model.add_aetmodel_v1("evap_aet_hbv96"):
    temperaturethresholdice(0.0)
    soilmoisturelimit(0.8)
    excessreduction(0.5)
    with model.add_petmodel_v1("evap_ret_io"):
        evapotranspirationfactor(0.8, 1.2, 0.5, 1.3, 1.0)
with model.add_rconcmodel_v1("rconc_uh"):
    uh("triangle", tb=4.0)
"""
    for element, cflux in ((e_a, 0.0), (e_b, 0.1)):
        ns: dict = {}
        exec(control.format(cflux=cflux), ns)   # like a control file
        element.model = ns["model"]
    hp = HydPy("ret_io_project")
    hp.update_devices(nodes=Nodes(n_out), elements=Elements(e_a, e_b))
    hp.update_parameters()
    hp.prepare_allseries()
    nt = len(pub.timegrids.init)
    for element in (e_a, e_b):
        model = element.model
        model.sequences.inputs.p.series = rng.gamma(0.4, 2.5, nt) * (rng.uniform(size=nt) < 0.4)
        model.sequences.inputs.t.series = 1.0 + 7.0 * np.sin(np.arange(nt) / 24.0) + rng.normal(0.0, 2.0, nt)
        ret = rng.uniform(0.0, 0.5, nt)
        ret[rng.uniform(size=nt) < 0.2] = 0.0          # nights without any demand
        model.aetmodel.petmodel.sequences.inputs.referenceevapotranspiration.series = ret
        sta = model.sequences.states
        sta.ic(0.4); sta.sp(2.0); sta.wc(0.2); sta.sm(120.0); sta.uz(3.0); sta.lz(9.0)
        model.rconcmodel.sequences.logs.quh(0.05)
    return hp, (("2000-01-01", "2000-01-05"), ("2000-01-05", "2000-01-09"))


def add_input_nodes(hp):
    """Feed three of the four meteorological inputs of two Lahn sub-basins through input nodes in different deploy
    modes (ref: hydpy/core/devicetools.py `Element.inputs`, hydpy/core/modeltools.py:1533 `_connect_inputs`)."""
    from hydpy import Element, Node, Nodes
    from hydpy.models.evap.evap_inputs import NormalAirTemperature, NormalEvapotranspiration
    from hydpy.models.hland.hland_inputs import P, T

    nt = len(pub.timegrids.init)
    rng = np.random.default_rng(11)
    t_node, p_node = Node("t_dill", variable=T), Node("p_dill", variable=P)
    nat_node, net_node = Node("nat_marb", variable=NormalAirTemperature), Node("net_marb", variable=NormalEvapotranspiration)
    p2_node = Node("p_marb", variable=P)
    Element("land_dill_assl", inputs=(t_node, p_node))
    Element("land_lahn_marb", inputs=(nat_node, net_node, p2_node))
    new = Nodes(t_node, p_node, nat_node, net_node, p2_node)
    hp.update_devices(nodes=hp.nodes + new, elements=hp.elements)
    for name in ("land_dill_assl", "land_lahn_marb"):
        hp.elements[name].model.connect()
    for node in new:
        node.prepare_allseries()
    with pub.options.checkseries(False):
        t_node.deploymode = "oldsim"
        t_node.sequences.sim.series = 4.0 * np.sin(np.arange(nt) / 9.0) + rng.normal(0.0, 1.0, nt)
        p_node.deploymode = "obs"
        p_node.sequences.obs.series = rng.gamma(0.5, 3.0, nt)
        nat_node.deploymode = "obs_oldsim"
        nat_node.sequences.sim.series = 1.0
        gaps = rng.uniform(0.0, 6.0, nt)
        gaps[rng.uniform(size=nt) < 0.4] = np.nan
        nat_node.sequences.obs.series = gaps
        net_node.deploymode = "obs_newsim"        # gaps fall back to the reset value 0.0
        gaps = rng.uniform(0.0, 3.0, nt)
        gaps[rng.uniform(size=nt) < 0.3] = np.nan
        net_node.sequences.obs.series = gaps
        p2_node.deploymode = "newsim"             # nothing feeds it: 0.0 at every step
    # the elements' own series must be ignored from now on
    hp.elements.land_dill_assl.model.sequences.inputs.t.series = 1e6
    hp.elements.land_dill_assl.model.sequences.inputs.p.series = 1e6


def main() -> None:
    del pub.projectname
    del pub.timegrids
    pub.options.prepare_testing(usecython=True, threads=0)
    if "--siblings" in sys.argv:
        with contextlib.redirect_stdout(io.StringIO()):
            hp, windows = sibling_project()
    elif "--retio" in sys.argv:
        with contextlib.redirect_stdout(io.StringIO()):
            hp, windows = ret_io_project()
    else:
        with contextlib.redirect_stdout(io.StringIO()):
            hp, _, _ = testtools.prepare_full_example_2(lastdate="1996-03-01")
        for element in hp.elements:
            seqs = element.model.sequences
            for group in ("factors", "fluxes", "states"):
                for seq in getattr(seqs, group):
                    seq.prepare_series(allocate_ram=True)
            for sub in element.model.find_submodels(include_mainmodel=False).values():   # evap_aet/pet, rconc
                for group in ("factors", "fluxes", "states"):
                    for seq in getattr(sub.sequences, group, ()) or ():
                        seq.prepare_series(allocate_ram=True)
        if "--inputnodes" in sys.argv:
            add_input_nodes(hp)
        # two consecutive simulation windows inside the initialisation period, as in hydpy/core/hydpytools.py:2804-2825
        windows = (("1996-01-01", "1996-02-01"), ("1996-02-01", "1996-03-01"))

    import copy

    initial = copy.deepcopy(hp.conditions)

    # (factors.referencewaterdepth of musk_mct — the start value of its root search — is no condition: reset it by hand)
    initial_rwd = {el: np.array(el.model.sequences.factors.referencewaterdepth.value) for el in hp.elements
                   if el.model.name == "musk_mct"}

    def run(simulate):
        hp.conditions = copy.deepcopy(initial)
        for el, value in initial_rwd.items():
            el.model.sequences.factors.referencewaterdepth.value = value
        for node in hp.nodes:
            node.sequences.sim.series[:] = np.nan
        for element in hp.elements:
            for group in ("factors", "fluxes", "states", "outlets"):
                for seq in getattr(element.model.sequences, group, ()):
                    if getattr(seq, "ramflag", False):
                        seq.series[:] = np.nan
            for sub in element.model.find_submodels(include_mainmodel=False).values():
                for group in ("factors", "fluxes", "states"):
                    for seq in getattr(sub.sequences, group, ()) or ():
                        if getattr(seq, "ramflag", False):
                            seq.series[:] = np.nan
            pet = getattr(getattr(element.model, "aetmodel", None), "petmodel", None)
            for seq in list(getattr(element.model.sequences, "inputs", ())) + list(pet.sequences.inputs if pet else ()):
                if getattr(seq, "node2idx", None) and seq.ramflag:
                    seq.series[:] = 1e6               # node-fed inputs: overwritten by what the node delivers
        snaps = []
        for first, last in windows:
            pub.timegrids.sim.firstdate = first
            pub.timegrids.sim.lastdate = last
            with contextlib.redirect_stdout(io.StringIO()):
                simulate()
            snaps.append(snapshot(hp))
        return snaps

    import time

    t0 = time.perf_counter()
    ref = run(hp.simulate)
    t_ref = time.perf_counter() - t0
    if not ENGINE:
        simulate_module.run_problem = oracle_run_problem        # the engine call, replaced by the oracle
    else:
        hydpy_b200.simulate(hp)                                 # first call: CUDA context, library load (not timed)
    t0 = time.perf_counter()
    got = run(lambda: hydpy_b200.simulate(hp))
    t_got = time.perf_counter() - t0
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
    worst = 0.0
    for which, snaps in (("simulate", got), ("install", got_installed)):
        for w, (a, b) in enumerate(zip(ref, snaps)):
            assert a.keys() == b.keys()
            for key in a:
                if ENGINE:
                    x, y = np.asarray(a[key], float), np.asarray(b[key], float)
                    assert x.shape == y.shape, key
                    if not np.allclose(x, y, rtol=1e-9, atol=1e-12, equal_nan=True):
                        idx = np.flatnonzero(~np.isclose(np.ravel(x), np.ravel(y), rtol=1e-9, atol=1e-12, equal_nan=True))
                        raise AssertionError(f"{which}, window {w}: {key} differs beyond 1e-9 at {idx[:8]} of {x.size}: "
                                             f"{np.ravel(x)[idx[:4]]} vs {np.ravel(y)[idx[:4]]}")
                    with np.errstate(invalid="ignore", divide="ignore"):
                        dev = np.abs(x - y) / np.maximum(np.abs(x), 1e-3)
                    worst = max(worst, float(np.nanmax(dev)) if dev.size and np.isfinite(dev).any() else 0.0)
                elif not np.array_equal(a[key], b[key], equal_nan=True):
                    idx = np.flatnonzero(~np.isclose(np.ravel(a[key]), np.ravel(b[key]), rtol=0, atol=0, equal_nan=True))
                    raise AssertionError(f"{which}, window {w}: {key} differs at {idx[:8]} of {np.size(a[key])}: "
                                         f"{np.ravel(a[key])[idx[:4]]} vs {np.ravel(b[key])[idx[:4]]}")
                n += 1
    if "--retio" in sys.argv:
        # a negative reference evapotranspiration cannot pass through the HBV96 equations unchanged: refused, not clipped
        series = hp.elements.retio_land_a.model.aetmodel.petmodel.sequences.inputs.referenceevapotranspiration.series
        series[-3] = -0.1                                        # (inside the last simulation window)
        try:
            hydpy_b200.simulate(hp)
        except hydpy_b200.UnsupportedModelError as exc:
            assert "negative reference evapotranspiration" in str(exc)
        else:
            raise AssertionError("negative evap_ret_io input was not refused")
    print(f"DROPIN_OK compared {n} arrays over {len(windows)} windows, {len(hp.elements)} elements, {len(hp.nodes)} nodes"
          + (f"; CUDA engine in the middle, worst deviation {worst:.2e} (bar 1e-9); wall: reference hp.simulate() "
             f"{t_ref:.3f} s, hydpy_b200.simulate(hp) {t_got:.3f} s for both windows" if ENGINE else
             "; oracle in the middle, bit for bit"))


if __name__ == "__main__":
    main()
