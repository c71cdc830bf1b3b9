#!/usr/bin/env python
"""Generate the committed golden fixtures by RUNNING THE REFERENCE ITSELF (only possible where /root/reference
exists, i.e. in the build container; the fixtures travel, the reference does not).

What is produced (all under tests/golden/):

* ``hland_96_integration.npz``  — the 11 integration tests of the reference's own ``hydpy/models/hland_96.py``
  docstring (96 hourly steps each): the extracted problem (parameters, initial conditions, forcing) plus EVERY
  factor/flux/state series, the outlet node series and the final conditions as raw float64.  While harvesting, the
  docstring is executed by ``doctest``, so the reference build is at the same time checked against its own printed
  6-digit tables.
* ``hland_96p_integration.npz``, ``hland_96c_integration.npz`` — the same for the sibling models (``--siblings``).
* ``musk_mct_integration.npz`` — the 5 integration tests of ``hydpy/models/musk_mct.py`` (``--musk_mct``).
* ``musk_classic_integration.npz`` — the 3 integration tests of ``hydpy/models/musk_classic.py``.
* ``hland_96_ret_io.npz`` — two `hland_96` elements (five zones of every type; one without, one with capillary flux)
  whose `evap_aet_hbv96` submodel reads its potential evapotranspiration from `evap_ret_io` (``--ret_io``; the project
  is the one of tests/dropin_reference_worker.py), 192 hourly steps, every sequence of every step.
* ``lahn_daily.npz`` — the HydPy-H-Lahn example project, 1996-01-01…1998-01-01 daily (config 1 of BASELINE.json):
  problem + all four gauge series + land outflows + final conditions + a few state series.
* ``lahn_hourly_members.npz`` — config 2 in miniature: Lahn with ``simulationstep 1h`` (synthetic hourly
  disaggregation of the daily forcing) for 14 days and 3 parameter-ensemble members.

Usage:  python tests/golden/make_golden.py            (builds the scratch reference first if necessary)
"""
from __future__ import annotations

import contextlib
import doctest
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402

SRC, STUBS = build_ref.scratch_paths()
if not os.path.exists(os.path.join(os.path.dirname(SRC), ".built")):
    build_ref.build()
sys.path[:0] = [STUBS, SRC]

import hydpy  # noqa: E402
from hydpy import pub  # noqa: E402
from hydpy.core import testtools  # noqa: E402

from hydpy_b200 import layout  # noqa: E402
from hydpy_b200.extract import extract  # noqa: E402


def gather_expected(problem, hp) -> dict[str, np.ndarray]:
    """Collect the reference's results in the ABI layouts."""
    land, i0, i1 = problem.land, problem.i0, problem.i1
    nt = i1 - i0
    out: dict[str, np.ndarray] = {}
    for name in layout.SERIES:
        group = layout.SERIES_GROUP[name]
        width = land.n_zone if name in layout.SERIES_ZONE else land.n_snow if name in layout.SERIES_SNOW else land.n_uh if name in layout.SERIES_UH else land.n_elem
        arr = np.full((nt, width), np.nan)
        have = False
        for e, element in enumerate(problem.land_elements):
            seq = getattr(getattr(element.model.sequences, group), name, None)
            if seq is None or not seq.ramflag:
                continue
            have = True
            data = np.asarray(seq.series[i0:i1], dtype=np.float64)
            if name in layout.SERIES_ZONE:
                arr[:, land.zone_ptr[e]:land.zone_ptr[e + 1]] = data
            elif name in layout.SERIES_SNOW:
                arr[:, land.snow_ptr[e]:land.snow_ptr[e + 1]] = data.reshape(nt, -1)
            else:
                arr[:, e] = data
        if have:
            out["exp_" + name] = arr
    # sequences of the evap submodels (third block of layout.SERIES): read from the submodel that owns them
    for (path, group, seqname), name in layout.SUBMODEL_SERIES.items():
        if name not in layout.SERIES_ZONE3 or seqname != name:
            continue
        arr = np.full((nt, land.n_zone), np.nan)
        have = False
        for e, element in enumerate(problem.land_elements):
            sub = element.model
            for part in path.split("."):
                sub = getattr(sub, part, None)
            seq = getattr(getattr(sub.sequences, group), seqname, None) if sub is not None else None
            if seq is None or not seq.ramflag:
                continue
            have = True
            arr[:, land.zone_ptr[e]:land.zone_ptr[e + 1]] = np.asarray(seq.series[i0:i1], dtype=np.float64)
        if have:
            out["exp_" + name] = arr
    arr = np.full((nt, land.n_uh), np.nan)
    have = False
    for e, element in enumerate(problem.land_elements):
        rc = getattr(element.model, "rconcmodel", None)
        seq = getattr(getattr(rc.sequences, "states", None), "sc", None) if rc is not None else None
        if seq is not None and seq.ramflag:
            have = True
            arr[:, land.uh_ptr[e]:land.uh_ptr[e + 1]] = np.asarray(seq.series[i0:i1], dtype=np.float64)
    if have:
        out["exp_sc"] = arr
    out["exp_node_rows"] = np.stack(
        [np.asarray(n.sequences.sim.series[i0:i1], dtype=np.float64) for n in problem.nodes]
    ) if problem.nodes else np.zeros((0, nt))
    after = extract(hp, i0, i1)
    for n in after.states.NAMES:
        out["exp_state_" + n] = getattr(after.states, n)
    out["exp_discharge"] = after.discharge
    if after.mct_states is not None:
        out["exp_mct_states"] = after.mct_states
    # states.discharge of every step (and, for musk_mct, the other per-segment sequences), stored like the state: [nt, n_seg]
    per_seg = {"dis_series": ("states", "discharge"), "cn_series": ("states", "courantnumber"),
               "rn_series": ("states", "reynoldsnumber"), "rwd_series": ("factors", "referencewaterdepth")}
    for name in layout.MCT_SERIES:   # every 1-D musk_mct sequence: exp_mct_<name>
        per_seg["mct_" + name] = (layout.MCT_SERIES_GROUP.get(name, "factors"), name)
    for key, (group, name) in per_seg.items():
        arr = np.full((nt, problem.river.n_seg), np.nan)
        have = False
        for r, element in enumerate(problem.river_elements):
            seq = getattr(getattr(element.model.sequences, group, None), name, None)
            if seq is None or not seq.ramflag:
                continue
            data = np.asarray(seq.series[i0:i1], dtype=np.float64).reshape(nt, -1)
            g0 = int(problem.river.seg_ptr[r])
            arr[:, g0:g0 + data.shape[1]] = data
            have = True
        if have:
            out["exp_" + key] = arr
    rin, rout = [], []
    for element in problem.river_elements:
        flu = element.model.sequences.fluxes
        if flu.inflow.ramflag and flu.outflow.ramflag:
            rin.append(np.asarray(flu.inflow.series[i0:i1], dtype=np.float64))
            rout.append(np.asarray(flu.outflow.series[i0:i1], dtype=np.float64))
    if rin and len(rin) == len(problem.river_elements):
        out["exp_reach_in"] = np.stack(rin)
        out["exp_reach_out"] = np.stack(rout)
    return out


def harvest_module(modulename: str, capture_all: bool = False) -> dict[str, dict[str, np.ndarray]]:
    """Execute the module docstring with doctest and snapshot every `test("name")` call (`capture_all`: also the calls
    that ask for the conditions of a later date, which the musk_mct tests all do)."""
    captured: dict[str, dict[str, np.ndarray]] = {}
    orig_call = testtools.IntegrationTest.__call__

    def hooked(self, filename=None, *, axis1=None, axis2=None, update_parameters=True, get_conditions=None,
               use_conditions=None):
        # same steps as the reference's IntegrationTest.__call__ (hydpy/core/testtools.py:626-645) minus plotting
        self.prepare_model(update_parameters=update_parameters, use_conditions=use_conditions)
        problem = None
        if (get_conditions is None or capture_all) and filename:
            i0, i1 = pub.timegrids.simindices
            problem = extract(self.hydpy, i0, i1)
        seq2value = self._perform_simulation(get_conditions)
        if problem is not None:
            d = problem.to_npz_dict()
            d.update(gather_expected(problem, self.hydpy))
            captured[str(filename)] = d
        self.print_table()
        return seq2value

    module = __import__(modulename, fromlist=["_"])
    del pub.projectname
    del pub.timegrids
    pub.options.prepare_testing(usecython=True, threads=0)
    testtools.IntegrationTest.plotting_options = testtools.PlottingOptions()
    testtools.IntegrationTest.__call__ = hooked
    try:
        from hydpy.core import devicetools

        tests = [t for t in doctest.DocTestFinder().find(module) if t.name == modulename]
        assert len(tests) == 1, [t.name for t in tests]
        runner = doctest.DocTestRunner(verbose=False)
        buf = io.StringIO()
        with devicetools.clear_registries_temporarily(), contextlib.redirect_stdout(buf):
            runner.run(tests[0], out=buf.write)
        res = runner.summarize(verbose=False)
        if res.failed:
            print(buf.getvalue()[-6000:])
            raise SystemExit(f"{modulename}: {res.failed} of {res.attempted} reference doctest examples failed")
        print(f"{modulename}: {res.attempted} reference doctest examples pass; {len(captured)} scenarios captured")
    finally:
        testtools.IntegrationTest.__call__ = orig_call
    return captured


def save_scenarios(path: str, scenarios: dict[str, dict[str, np.ndarray]]) -> None:
    flat = {}
    for sname, d in scenarios.items():
        for k, v in d.items():
            flat[f"{sname}/{k}"] = np.asarray(v)
    flat["__scenarios__"] = np.array(sorted(scenarios))
    np.savez_compressed(path, **flat)
    print(f"wrote {path}: {os.path.getsize(path)/1024:.0f} KiB, scenarios: {', '.join(sorted(scenarios))}")


def prepare_lahn(lastdate: str):
    del pub.projectname
    del pub.timegrids
    pub.options.prepare_testing(usecython=True, threads=0)
    with contextlib.redirect_stdout(io.StringIO()):
        hp, _, _ = testtools.prepare_full_example_2(lastdate=lastdate)
    return hp


def lahn_daily() -> dict[str, np.ndarray]:
    hp = prepare_lahn("1998-01-01")
    for element in hp.elements:
        if element.model.name == "hland_96":
            seqs = element.model.sequences
            for seq in (seqs.states.sm, seqs.states.sp, seqs.states.uz, seqs.states.lz, seqs.fluxes.qt,
                        seqs.fluxes.q0, seqs.fluxes.perc, seqs.factors.contriarea):
                seq.prepare_series(allocate_ram=True)
        else:
            element.model.sequences.fluxes.inflow.prepare_series(allocate_ram=True)
            element.model.sequences.fluxes.outflow.prepare_series(allocate_ram=True)
    i0, i1 = pub.timegrids.simindices
    problem = extract(hp, i0, i1)
    with contextlib.redirect_stdout(io.StringIO()):
        hp.simulate()
    d = problem.to_npz_dict()
    d.update(gather_expected(problem, hp))
    d["node_names"] = np.array([n.name for n in problem.nodes])
    d["land_names"] = np.array([e.name for e in problem.land_elements])
    d["reach_names"] = np.array([e.name for e in problem.river_elements])
    # the reference's own docstring goldens (hydpy/core/hydpytools.py:2791-2794) as a cross-check
    kalk = d["exp_node_rows"][list(d["node_names"]).index("lahn_kalk")][:4]
    assert np.allclose(kalk, [54.019332, 37.257552, 31.865302, 28.359538], rtol=0, atol=5e-7), kalk
    return d


def lahn_hourly_members(n_members: int = 3, days: int = 14) -> dict[str, dict[str, np.ndarray]]:
    """Config 2 in miniature.  Hourly forcing = deterministic disaggregation of the daily Lahn series
    (P/24 and NormalEvapotranspiration/24 per hour; T and NormalAirTemperature repeated), member 0 = shipped
    parameter values, further members drawn with default_rng(20260101) inside plausible calibration bounds."""
    import datetime

    del pub.projectname
    del pub.timegrids
    pub.options.prepare_testing(usecython=True, threads=0)
    first = datetime.datetime(1996, 1, 1)
    last = first + datetime.timedelta(days=days)
    with contextlib.redirect_stdout(io.StringIO()):
        testtools.prepare_full_example_1()
        with testtools.TestIO():
            hp = hydpy.HydPy("HydPy-H-Lahn")
            pub.timegrids = first.strftime("%Y-%m-%d"), last.strftime("%Y-%m-%d"), "1d"
            hp.prepare_network()
            hp.prepare_models()
            hp.load_conditions()
            hp.prepare_inputseries()
            hp.load_inputseries()
            daily = {}
            for element in hp.elements:
                m = element.model
                if m.name != "hland_96":
                    continue
                pet = m.aetmodel.petmodel
                daily[element.name] = (
                    m.sequences.inputs.p.series.copy(), m.sequences.inputs.t.series.copy(),
                    pet.sequences.inputs.normalairtemperature.series.copy(),
                    pet.sequences.inputs.normalevapotranspiration.series.copy(),
                )
            # switch to hourly: re-create the project with simulationstep 1h
            del pub.timegrids
            hydpy.Node.clear_all()
            hydpy.Element.clear_all()
            hp = hydpy.HydPy("HydPy-H-Lahn")
            pub.timegrids = first.strftime("%Y-%m-%d"), last.strftime("%Y-%m-%d"), "1h"
            hp.prepare_network()
            hp.prepare_models()
            hp.load_conditions()
            hp.prepare_nodeseries()
            hp.prepare_inputseries()
    for element in hp.elements:
        m = element.model
        if m.name != "hland_96":
            continue
        p, t, nat, net = daily[element.name]
        pet = m.aetmodel.petmodel
        m.sequences.inputs.p.series = np.repeat(p / 24.0, 24)
        m.sequences.inputs.t.series = np.repeat(t, 24)
        pet.sequences.inputs.normalairtemperature.series = np.repeat(nat, 24)
        pet.sequences.inputs.normalevapotranspiration.series = np.repeat(net / 24.0, 24)
    rng = np.random.default_rng(20260101)
    conditions = hp.conditions
    scenarios = {}
    for member in range(n_members):
        if member > 0:
            for element in hp.elements:
                m = element.model
                if m.name == "hland_96":
                    con = m.parameters.control
                    con.fc(float(rng.uniform(100.0, 300.0)))
                    con.beta(float(rng.uniform(1.0, 4.0)))
                    con.percmax(float(rng.uniform(0.5, 2.0)))
                    con.k(float(rng.uniform(0.001, 0.01)))
                    con.k4(float(rng.uniform(0.005, 0.05)))
                    con.alpha(float(rng.uniform(0.5, 1.5)))
                    con.cfmax(float(rng.uniform(2.0, 5.0)))
                    con.tt(float(rng.uniform(-1.0, 1.0)))
                    con.whc(float(rng.uniform(0.05, 0.2)))
                    con.cflux(float(rng.uniform(0.0, 1.0)))
                    m.aetmodel.parameters.control.soilmoisturelimit(float(rng.uniform(0.6, 1.0)))
                    m.aetmodel.parameters.control.maxsoilwater(con.fc)
                else:
                    con = m.parameters.control
                    con.nmbsegments(lag=float(rng.uniform(0.0, 0.8)))
                    con.coefficients(damp=float(rng.uniform(0.0, 1.0)))
                m.update_parameters()
            hp.conditions = conditions
            for element in hp.elements:
                m = element.model
                if m.name == "musk_classic":
                    m.sequences.states.discharge(0.0)
        i0, i1 = pub.timegrids.simindices
        problem = extract(hp, i0, i1)
        with contextlib.redirect_stdout(io.StringIO()):
            hp.simulate()
        d = problem.to_npz_dict()
        d.update(gather_expected(problem, hp))
        scenarios[f"member{member}"] = d
    return scenarios


def write_template(daily: dict[str, np.ndarray], hourly: dict[str, np.ndarray]) -> None:
    """`hydpy_b200/data/template.npz`: internal parameter values of the real reference element `land_dill_assl`
    (element 0 of the Lahn fixtures) for both simulation steps — the seed of `hydpy_b200.synth`."""
    out = {}
    for step, d in (("1d", daily), ("1h", hourly)):
        z0, z1 = int(d["land_zone_ptr"][0]), int(d["land_zone_ptr"][1])
        u0, u1 = int(d["land_uh_ptr"][0]), int(d["land_uh_ptr"][1])
        out[f"{step}/zpar"] = d["land_zpar"][:, z0:z1]
        out[f"{step}/epar"] = d["land_epar"][:, :1]
        out[f"{step}/zonetype"] = d["land_zonetype"][z0:z1]
        out[f"{step}/zflags"] = d["land_zflags"][z0:z1]
        out[f"{step}/uh"] = d["land_uh"][u0:u1]
        out[f"{step}/recstep"] = d["land_recstep"][:1]
        out[f"{step}/eflags"] = d["land_eflags"][:1]
    path = os.path.join(ROOT, "hydpy_b200", "data", "template.npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)/1024:.0f} KiB")


def ret_io() -> dict[str, dict[str, np.ndarray]]:
    """`evap_ret_io` behind `evap_aet_hbv96` (ref: hydpy/models/evap_ret_io.py:53-64): the project is built by
    tests/dropin_reference_worker.py::ret_io_project through the reference's own API."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("dropin_reference_worker",
                                                  os.path.join(ROOT, "tests", "dropin_reference_worker.py"))
    worker = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(worker)
    del pub.projectname
    del pub.timegrids
    pub.options.prepare_testing(usecython=True, threads=0)
    with contextlib.redirect_stdout(io.StringIO()):
        hp, _ = worker.ret_io_project()
    i0, i1 = pub.timegrids.simindices
    problem = extract(hp, i0, i1)
    with contextlib.redirect_stdout(io.StringIO()):
        hp.simulate()
    d = problem.to_npz_dict()
    d.update(gather_expected(problem, hp))
    return {"ret_io": d}


def main() -> None:
    if "--ret_io" in sys.argv:
        save_scenarios(os.path.join(HERE, "hland_96_ret_io.npz"), ret_io())
        return
    if "--musk_mct" in sys.argv:
        save_scenarios(os.path.join(HERE, "musk_mct_integration.npz"),
                       harvest_module("hydpy.models.musk_mct", capture_all=True))
        return
    if "--siblings" in sys.argv:
        # the integration tests of the sibling models' own docstrings (hland_96p: 4, hland_96c: 6 distinct
        # scenarios, all with rconc_nash)
        for name in ("hland_96p", "hland_96c"):
            save_scenarios(os.path.join(HERE, f"{name}_integration.npz"), harvest_module(f"hydpy.models.{name}"))
        return
    save_scenarios(os.path.join(HERE, "hland_96_integration.npz"), harvest_module("hydpy.models.hland_96"))
    save_scenarios(os.path.join(HERE, "musk_classic_integration.npz"), harvest_module("hydpy.models.musk_classic"))
    daily = lahn_daily()
    save_scenarios(os.path.join(HERE, "lahn_daily.npz"), {"lahn_daily": daily})
    hourly = lahn_hourly_members()
    save_scenarios(os.path.join(HERE, "lahn_hourly_members.npz"), hourly)
    write_template(daily, hourly["member0"])


if __name__ == "__main__":
    main()
