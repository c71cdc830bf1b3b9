#!/usr/bin/env python
"""(Lives under tests/ because it uses the CPU oracle as its checker.)  Secondary measurements beside bench.py (which holds the contract workload, BASELINE.json configs[3]): the other
configurations of BASELINE.json that fit one GPU, and the sibling models.  One JSON line per case on stdout.

  config2   HydPy-H-Lahn hourly x 4096 parameter-ensemble members (real Lahn structure and parameters from the committed
            fixture tests/golden/lahn_hourly_members.npz, member 0 = shipped values; 336 hourly steps per window,
            members share the four forcing columns)                                     [BASELINE.json configs[1]]
  config3   synthetic 10^5 hland_96 elements x 12 zones, DAILY (RecStep = 1200 per step), runoff generation only,
            365 daily steps per window                                                   [BASELINE.json configs[2]]
  config5   one GPU's share of the continental case in miniature: 83 334 elements x 12 zones (10^6 zones) + river tree,
            hourly, x `--members5` ensemble members (default 16 -> 1.6*10^7 zones, 1.3*10^6 reaches on one GPU; the full
            case is 512 members over 8 GPUs = 64 per GPU)                                [BASELINE.json configs[4]]
  hland_96p / hland_96c   the configs[3] basin with sibling-model elements + rconc_nash

Timing: medians of the CUDA-event times of the engine (hb_get_timing) over `--reps` synchronous windows after one warm-up window, inputs
resident, two warm-up windows; `e2e_ms` = wall clock of the same window through the public API including the H2D copy of the forcing and the
D2H copy of all node series.  Every case is checked against the CPU oracle on an element sample first.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hydpy_b200 import Engine, Problem, synth  # noqa: E402
from hydpy_b200.shard import subset_land  # noqa: E402


def measure(eng: Engine, name: str, land, states, river, discharge, forcings, doys, *, route: bool, reps: int,
            note: str) -> dict:
    eng.set_land(land)
    eng.set_land_states(states)
    eng.set_river(river)
    eng.set_river_states(discharge)
    eng.select_series([])
    nt = forcings[0].shape[1]
    no_ext = np.zeros((0, nt))
    dev, e2e = [], []
    warm = 2                         # both sets of the engine's double-buffered window buffers get allocated
    for rep in range(reps + warm):
        f, d = forcings[rep % len(forcings)], doys[rep % len(doys)]
        t0 = time.perf_counter()
        eng.set_forcing(f, d)
        eng.set_external(no_ext)
        eng.timing()
        eng.run(land=True, river=route)
        tm = eng.timing()
        rows = eng.get_node_series() if route else eng.get_land_outflow()
        t1 = time.perf_counter()
        if rep >= warm:              # the first windows warm up (allocations, module load)
            dev.append(tm)
            e2e.append(1e3 * (t1 - t0))
    zs = land.n_zone * nt
    land_ms = float(np.median([t["land_ms"] for t in dev]))
    river_ms = float(np.median([t["river_ms"] for t in dev]))
    total = land_ms + (river_ms if route else 0.0)
    return {
        "case": name, "note": note, "elements": land.n_elem, "zones": land.n_zone, "reaches": river.n_reach,
        "window_steps": nt, "reps": reps, "land_ms": land_ms, "land_ms_reps": [round(t["land_ms"], 3) for t in dev], "zone_ms": float(np.mean([t["zone_ms"] for t in dev])),
        "element_ms": float(np.mean([t["element_ms"] for t in dev])),
        "general_ms": float(np.mean([t["general_ms"] for t in dev])), "river_ms": river_ms if route else None,
        "n_fast": int(dev[-1]["n_fast"]), "n_spec": int(dev[-1]["n_spec"]),
        "spec_reruns_per_window": float(np.mean([t["spec_reruns"] for t in dev])), "value": zs / (total * 1e-3), "unit": "zone-timesteps/s (device, synchronous windows)",
        "e2e_ms": float(np.median(e2e)), "e2e_value": zs / (float(np.median(e2e)) * 1e-3),
        "finite": bool(np.isfinite(rows).all()), "mean_discharge": float(rows.mean()),
    }


def check_sample(eng: Engine, land, states, forcing, doy, n: int = 64) -> float:
    """Runoff generation of the first `n` elements against the CPU oracle (relative deviation of the outflow)."""
    from oracle import oracle
    from tests.util import assert_close

    idx = np.arange(min(n, land.n_elem))
    sub, st = subset_land(land, states, idx, np.full(len(idx), -1, np.int32))
    from hydpy_b200.abi import RiverBundle

    res_st = st.copy()
    qt, _ = oracle.land_run(sub, res_st, forcing, doy, threads=4)
    eng.set_land(sub)
    eng.set_land_states(st)
    eng.set_river(RiverBundle.empty(0))
    eng.select_series([])
    eng.set_forcing(forcing, doy)
    eng.run(land=True, river=False)
    return assert_close(eng.get_land_outflow(), qt, "sample vs oracle")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="config2,config3,hland_96p,hland_96c")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--members", type=int, default=4096)
    ap.add_argument("--elements", type=int, default=100_000)
    ap.add_argument("--members5", type=int, default=16)
    ap.add_argument("--land-slabs", type=int, default=0, help="element slabs of the fast land path (0 = engine default)")
    ap.add_argument("--no-land-fuse", action="store_true", help="zone terms through HBM (A/B timing)")
    args = ap.parse_args()
    eng = Engine(0)
    if args.land_slabs:
        eng.set_land_slabs(args.land_slabs)
    if args.no_land_fuse:
        eng.set_land_fuse(False)
    info = eng.device_info()
    for case in args.cases.split(","):
        if case == "config2":
            from tests.util import load_scenarios

            p = load_scenarios("lahn_hourly_members.npz")["member0"].problem
            worst = check_sample(eng, p.land, p.states, p.forcing, p.doy)
            land, states = synth.replicate_members(p.land, p.states, args.members, n_node=p.river.n_node)
            river, discharge = synth.replicate_river(p.river, p.discharge, args.members, p.land.n_elem)
            out = measure(eng, case, land, states, river, discharge, [p.forcing], [p.doy], route=True, reps=args.reps,
                          note=f"HydPy-H-Lahn hourly x {args.members} members (4 hland_96 elements, 49 zones, 3 musk_classic "
                               f"reaches per member); CFlux = 0 everywhere: two-kernel fast path")
        elif case == "config3":
            land, states = synth.make_land(args.elements, "1d")
            river, discharge = synth.make_tree_river(args.elements, "1d", levels=4)
            fs, ds = zip(*[synth.make_forcing(365, "1d", t0=365 * i) for i in range(2)])
            worst = check_sample(eng, land, states, fs[0][:, :30], ds[0][:30])
            out = measure(eng, case, land, states, river, discharge, list(fs), list(ds), route=False, reps=args.reps,
                          note="synthetic daily basin, RecStep = 1200 sub-steps per step, runoff generation only")
        elif case == "config5":
            base_land, base_states = synth.make_land(83_334, "1h")
            base_river, base_dis = synth.make_tree_river(83_334, "1h", levels=200)
            fs, ds = zip(*[synth.make_forcing(240, "1h", t0=240 * i) for i in range(2)])
            worst = check_sample(eng, base_land, base_states, fs[0][:, :72], ds[0][:72])
            land, states = synth.replicate_members(base_land, base_states, args.members5, n_node=base_river.n_node)
            river, discharge = synth.replicate_river(base_river, base_dis, args.members5, base_land.n_elem)
            del base_land, base_states, base_river, base_dis
            out = measure(eng, case, land, states, river, discharge, list(fs), list(ds), route=True, reps=args.reps,
                          note=f"10^6-zone basin x {args.members5} ensemble members on one GPU (members = additional elements "
                               f"sharing the forcing bank; zone terms are processed in time chunks bounded by device memory)")
        elif case in ("hland_96p", "hland_96c"):
            land, states = synth.make_land(args.elements, "1h", model=case)
            river, discharge = synth.make_tree_river(args.elements, "1h", levels=200)
            fs, ds = zip(*[synth.make_forcing(240, "1h", t0=240 * i) for i in range(2)])
            worst = check_sample(eng, land, states, fs[0][:, :72], ds[0][:72])
            out = measure(eng, case, land, states, river, discharge, list(fs), list(ds), route=True, reps=args.reps,
                          note=f"configs[3] basin with {case} + rconc_nash elements (sibling instances of the fast path)")
        elif case in ("smax_high", "smax_low", "smax_high_general", "smax_low_general"):
            # finite SMax on every zone (ref: hland_model.py:584-605, 876-969) under a winter forcing (12 K colder: the
            # snow packs grow by about 25 mm per window): speculative fast path with a limit the packs do not reach during
            # the run (400 mm) and one most elements reach after a few windows (60 mm: the worst case — fast path AND a
            # repetition in the general kernel; an element that lost snow goes straight to the general kernel in the next
            # window), each against the general kernel alone (HB_OPT_LAND_SPECULATE = 0, round 1's way)
            from hydpy_b200 import layout

            land, states = synth.make_land(args.elements, "1h")
            land.zpar[layout.ZP["smax"]] = 60.0 if case.startswith("smax_low") else 400.0
            river, discharge = synth.make_tree_river(args.elements, "1h", levels=200)
            fs, ds = zip(*[synth.make_forcing(240, "1h", t0=240 * i) for i in range(2)])
            for f in fs:
                f[1] -= 12.0
            eng.set_land_speculate(not case.endswith("_general"))
            worst = check_sample(eng, land, states, fs[0][:, :72], ds[0][:72])
            out = measure(eng, case, land, states, river, discharge, list(fs), list(ds), route=True, reps=args.reps,
                          note=f"configs[3] basin, winter, SMax = {land.zpar[layout.ZP['smax']][0]:.0f} mm on every zone: " +
                               ("general kernel (speculation off)" if case.endswith("_general") else
                                "speculative fast path (element-windows with snow above the limit repeated by the general kernel)"))
            eng.set_land_speculate(True)
        else:
            raise SystemExit(f"unknown case {case}")
        out["oracle_sample_max_rel_deviation"] = worst
        out["device"] = info["name"]
        print(json.dumps(out), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
