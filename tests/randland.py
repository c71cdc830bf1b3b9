"""Random ragged test problems (test infrastructure): arbitrary zone counts, snow classes, zone types, finite SMax with
snow-redistribution edges, cflux > 0, optional unit hydrographs.  Oracle and engine consume the same bundle, so the
values only have to be plausible, not hydrologically calibrated."""
from __future__ import annotations

import numpy as np

from hydpy_b200 import layout
from hydpy_b200.abi import LandBundle, LandStates, RiverBundle


def random_land(n_elem: int, seed: int = 0, zmax: int = 14, cmax: int = 3, recstep: int = 20, bank: int | None = None,
                snowlimit_fraction: float = 0.3, fast_fraction: float = 0.4, soil_runs: int = 0
                ) -> tuple[LandBundle, LandStates]:
    """`fast_fraction` of the elements get one snow class, CFlux = 0 and SMax = inf, i.e. they qualify for the engine's
    two-kernel fast path (all zone types still occur); the others exercise the general kernel.  `soil_runs` > 0: every
    other run of that many consecutive elements is fast with FIELD/FOREST zones and FC > 0 only (ragged zone counts) —
    the 32-element groups inside such runs qualify for the reduction inside the zone kernel."""
    rng = np.random.default_rng(seed)
    ne = n_elem
    nz = rng.integers(1, zmax + 1, size=ne)
    nc = rng.integers(1, cmax + 1, size=ne)
    fast = np.random.default_rng(seed + 4242).uniform(size=ne) < fast_fraction
    in_run = ((np.arange(ne) // soil_runs) % 2 == 0) if soil_runs > 0 else np.zeros(ne, bool)
    fast |= in_run
    # (the fast path takes several snow classes too: half of the fast elements outside the soil runs keep theirs)
    one_class = fast & (in_run | (np.random.default_rng(seed + 777).uniform(size=ne) < 0.5))
    nc[one_class] = 1
    nu = rng.integers(0, 6, size=ne)
    zone_ptr = np.concatenate([[0], np.cumsum(nz)]).astype(np.int64)
    snow_ptr = np.concatenate([[0], np.cumsum(nz * nc)]).astype(np.int64)
    sfdist_ptr = np.concatenate([[0], np.cumsum(nc)]).astype(np.int64)
    uh_ptr = np.concatenate([[0], np.cumsum(nu)]).astype(np.int64)
    NZ = int(zone_ptr[-1])
    zonetype = rng.choice([1, 2, 3, 4, 5], size=NZ, p=[0.35, 0.35, 0.1, 0.1, 0.1]).astype(np.int32)
    run_zone = np.repeat(in_run, nz)
    zonetype = np.where(run_zone, 1 + (np.arange(NZ) + seed) % 2, zonetype).astype(np.int32)
    ZP, EP = layout.ZP, layout.EP
    zpar = np.zeros((len(layout.ZPAR), NZ))
    U = lambda lo, hi, n=NZ: rng.uniform(lo, hi, size=n)  # noqa: E731
    zpar[ZP["zonez"]] = U(1, 12)
    zpar[ZP["pcorr"]] = U(0.9, 1.2); zpar[ZP["pcalt"]] = U(0.0, 0.15); zpar[ZP["rfcf"]] = U(0.9, 1.2)
    zpar[ZP["sfcf"]] = U(0.9, 1.4); zpar[ZP["tcorr"]] = U(-0.5, 0.5); zpar[ZP["tcalt"]] = U(0.4, 0.8)
    zpar[ZP["icmax"]] = U(0.0, 2.0); zpar[ZP["tt"]] = U(-1, 1); zpar[ZP["ttint"]] = U(0.5, 3.0)
    zpar[ZP["ttm"]] = zpar[ZP["tt"]] + U(-0.5, 0.5); zpar[ZP["cfmax"]] = U(0.05, 0.3); zpar[ZP["cfvar"]] = U(0, 0.05)
    zpar[ZP["gmelt"]] = U(0.05, 0.3); zpar[ZP["gvar"]] = U(0, 0.05); zpar[ZP["cfr"]] = U(0, 0.1)
    zpar[ZP["whc"]] = U(0, 0.3); zpar[ZP["fc"]] = np.where(rng.uniform(size=NZ) < 0.03, 0.0, U(50, 300))
    zpar[ZP["fc"]] = np.where(run_zone & (zpar[ZP["fc"]] == 0.0), 120.0, zpar[ZP["fc"]])
    zpar[ZP["beta"]] = U(1, 4); zpar[ZP["cflux"]] = np.where(rng.uniform(size=NZ) < 0.5, 0.0, U(0, 0.05))
    zpar[ZP["temperaturethresholdice"]] = np.where(rng.uniform(size=NZ) < 0.3, np.nan, U(-1, 1))
    zpar[ZP["maxsoilwater"]] = np.maximum(zpar[ZP["fc"]], 1.0); zpar[ZP["soilmoisturelimit"]] = U(0.5, 1.0)
    zpar[ZP["excessreduction"]] = U(0, 1); zpar[ZP["hrualtitude"]] = 100 * zpar[ZP["zonez"]]
    zpar[ZP["evapotranspirationfactor"]] = U(0.6, 1.2); zpar[ZP["altitudefactor"]] = U(-0.1, 0.1)
    zpar[ZP["precipitationfactor"]] = U(0, 0.3); zpar[ZP["airtemperaturefactor"]] = U(0, 0.2)
    zpar[ZP["smax"]] = np.inf
    zflags = np.zeros(NZ, np.int32)
    soil = (zonetype == 1) | (zonetype == 2)
    zflags |= np.where(zonetype == 4, layout.ZF_WATER, 0).astype(np.int32)
    zflags |= np.where(soil | (zonetype == 5), layout.ZF_INTERCEPTION, 0).astype(np.int32)
    zflags |= np.where(soil & (rng.uniform(size=NZ) < 0.95), layout.ZF_SOIL, 0).astype(np.int32)
    zflags |= np.where(zonetype == 2, layout.ZF_TREE, 0).astype(np.int32)
    epar = np.zeros((len(layout.EPAR), ne))
    zorder = np.zeros(NZ, np.int32)
    sred_ptr = np.zeros(ne + 1, np.int64)
    sf, st_, sa = [], [], []
    eflags = np.zeros(ne, np.int32)
    for e in range(ne):
        z0, z1 = zone_ptr[e], zone_ptr[e + 1]
        n = z1 - z0
        area = rng.uniform(0.2, 1.0, size=n)
        rel = area / area.sum()
        zpar[ZP["relzonearea"], z0:z1] = rel
        zpar[ZP["hruareafraction"], z0:z1] = rel
        zt = zonetype[z0:z1]
        s = (zt == 1) | (zt == 2)
        epar[EP["z"], e] = np.sum(rel * zpar[ZP["zonez"], z0:z1])
        epar[EP["relsoilarea"], e] = rel[s].sum()
        epar[EP["rellandarea"], e] = rel[zt != 4].sum()
        up = rel[s | (zt == 3)].sum()
        lo = rel[zt != 5].sum()
        # keep the weights finite: degenerate area sums would divide by zero in both implementations
        epar[EP["relupperzonearea"], e] = up if up > 0 else 1.0
        epar[EP["rellowerzonearea"], e] = lo
        epar[EP["petaltitude"], e] = np.sum(rel * zpar[ZP["hrualtitude"], z0:z1])
        order = np.argsort(zpar[ZP["zonez"], z0:z1], kind="stable")
        zorder[z0:z1] = order
        eflags[e] = layout.EF_RESPAREA if rng.uniform() < 0.5 else 0
        edges = 0
        if fast[e]:
            zpar[ZP["cflux"], z0:z1] = 0.0
        if rng.uniform() < snowlimit_fraction and not fast[e]:
            zpar[ZP["smax"], z0:z1] = rng.uniform(5.0, 40.0, size=n)
            # redistribution edges from higher to lower non-lake zones (a DAG in descending altitude)
            desc = order[::-1]
            land_desc = [k for k in desc if zt[k] != 4]
            ends = np.ones(n, bool)
            for i, f in enumerate(land_desc[:-1]):
                targets = land_desc[i + 1:i + 3]
                w = rng.uniform(0.2, 1.0, size=len(targets))
                w /= w.sum()
                for t, wt in zip(targets, w):
                    sf.append(f); st_.append(t); sa.append(float(rel[f] / rel[t] * wt)); edges += 1
                ends[f] = False
            ends[zt == 4] = False
            zflags[z0:z1] |= np.where(ends, layout.ZF_SREDEND, 0).astype(np.int32)
        sred_ptr[e + 1] = sred_ptr[e] + edges
    epar[EP["percmax"]] = rng.uniform(0.02, 0.1, ne); epar[EP["dt"]] = 1.0 / recstep
    epar[EP["alpha"]] = np.where(rng.uniform(size=ne) < 0.3, 1.0, rng.uniform(0.3, 1.5, ne))
    epar[EP["k"]] = rng.uniform(1e-4, 1e-2, ne); epar[EP["k4"]] = rng.uniform(1e-3, 1e-2, ne)
    epar[EP["gamma"]] = np.where(rng.uniform(size=ne) < 0.5, 0.0, rng.uniform(0, 0.5, ne))
    epar[EP["qfactor"]] = rng.uniform(1, 30, ne)
    sfdist = np.concatenate([(lambda w: w / w.mean())(rng.uniform(0.5, 1.5, size=c)) for c in nc])
    uh = np.concatenate([(lambda w: w / w.sum())(rng.uniform(0.1, 1.0, size=u)) for u in nu] + [np.zeros(0)])
    land = LandBundle(
        zone_ptr=zone_ptr, snow_ptr=snow_ptr, sfdist_ptr=sfdist_ptr, uh_ptr=uh_ptr, sred_ptr=sred_ptr,
        recstep=np.full(ne, recstep, np.int32), eflags=eflags,
        forcing_col=(np.arange(ne) % (bank or ne)).astype(np.int32), outlet_node=np.arange(ne, dtype=np.int32),
        epar=epar, zonetype=zonetype, zflags=zflags, zorder=zorder, zpar=zpar, sfdist=sfdist, uh=uh,
        sred_from=np.array(sf, np.int32), sred_to=np.array(st_, np.int32), sred_adjust=np.array(sa, np.float64),
    )
    land.normalise()
    NS = int(snow_ptr[-1])
    lake_zone = zonetype == 4
    states = LandStates(
        ic=np.where(lake_zone, 0.0, rng.uniform(0, 1, NZ)), sm=np.where(soil, rng.uniform(20, 150, NZ), 0.0),
        sp=rng.uniform(0, 20, NS) * (rng.uniform(size=NS) < 0.7), wc=rng.uniform(0, 2, NS),
        uz=rng.uniform(0, 10, ne), lz=rng.uniform(0, 20, ne), quh=rng.uniform(0, 0.5, int(uh_ptr[-1])),
    )
    # lake zones carry no snow
    for e in range(ne):
        n, c = int(nz[e]), int(nc[e])
        lk = np.tile(zonetype[zone_ptr[e]:zone_ptr[e + 1]] == 4, c)
        states.sp[snow_ptr[e]:snow_ptr[e + 1]][lk] = 0.0
        states.wc[snow_ptr[e]:snow_ptr[e + 1]][lk] = 0.0
    return land, states


def add_siblings(land: LandBundle, states: LandStates, seed: int = 0, fraction: float = 0.6, nash_fraction: float = 0.6
                 ) -> None:
    """Turn `fraction` of the elements into hland_96p / hland_96c elements and give `nash_fraction` of
    ALL elements a rconc_nash submodel (their uh range becomes the storage cascade) — in place, with its own random
    stream, so the hland_96 part of the bundle stays what `random_land` drew."""
    rng = np.random.default_rng(seed + 99)
    land.normalise()
    states.normalise()
    ne, NZ = land.n_elem, land.n_zone
    ZP, EP = layout.ZP, layout.EP
    pick = rng.uniform(size=ne)
    land.model[:] = np.where(pick < fraction / 2, layout.MODEL_HLAND_96P,
                             np.where(pick < fraction, layout.MODEL_HLAND_96C, layout.MODEL_HLAND_96))
    zmodel = np.repeat(land.model, np.diff(land.zone_ptr))
    U = lambda lo, hi, n=NZ: rng.uniform(lo, hi, size=n)  # noqa: E731
    upper = np.isin(land.zonetype, (1, 2, 3))
    # hland_96p: control SGR, K2, SG1Max; derived W0 = exp(-1/K0), W1 = exp(-1/K1), W2 = exp(-1/K2), …
    p = zmodel == layout.MODEL_HLAND_96P
    k2 = U(10, 200)
    for name, val in (("sgr", U(2, 30)), ("w0", np.exp(-1.0 / U(2, 20))), ("w1", np.exp(-1.0 / U(10, 100))),
                      ("k2", k2), ("w2", np.exp(-1.0 / k2)), ("sg1max", U(5, 60))):
        land.zpar[ZP[name]] = np.where(p, val, 0.0)
    ep_ = land.model == layout.MODEL_HLAND_96P
    k3, k4 = rng.uniform(50, 800, ne), rng.uniform(500, 5000, ne)
    land.epar[EP["k3"]] = np.where(ep_, k3, 0.0)
    land.epar[EP["w3"]] = np.where(ep_, np.exp(-1.0 / k3), 0.0)
    land.epar[EP["w4"]] = np.where(ep_, np.exp(-1.0 / k4), 0.0)
    land.epar[EP["fsg"]] = np.where(ep_, 8.0 / 9.0, 0.0)
    land.epar[EP["k4"]] = np.where(ep_, k4, land.epar[EP["k4"]])
    land.epar[EP["percmax"]] = np.where(ep_, rng.uniform(0.05, 1.5, ne), land.epar[EP["percmax"]])
    # hland_96c: control H1, TAb1, TVs1, H2, TAb2, TVs2 (zero and infinite time constants hit the special branches)
    c = zmodel == layout.MODEL_HLAND_96C

    def tconst(lo, hi):
        v = U(lo, hi)
        r = rng.uniform(size=NZ)
        return np.where(r < 0.08, 0.0, np.where(r < 0.16, np.inf, v))

    for name, val in (("h1", U(0, 15)), ("tab1", tconst(1, 40)), ("tvs1", tconst(1, 80)), ("h2", U(0, 30)),
                      ("tab2", tconst(5, 100)), ("tvs2", tconst(5, 200))):
        land.zpar[ZP[name]] = np.where(c, val, 0.0)
    # (two degenerate time constants at once divide 0/0 or inf/inf in the reference as well: keep one regular)
    for a, b in (("tab1", "tvs1"), ("tab2", "tvs2")):
        both = (~np.isfinite(land.zpar[ZP[a]]) | (land.zpar[ZP[a]] == 0.0)) & \
               (~np.isfinite(land.zpar[ZP[b]]) | (land.zpar[ZP[b]] == 0.0)) & c
        land.zpar[ZP[b]] = np.where(both, 7.0, land.zpar[ZP[b]])
    sib = land.model != layout.MODEL_HLAND_96
    land.epar[EP["dt"]] = np.where(sib, 1.0, land.epar[EP["dt"]])
    land.recstep[:] = np.where(sib, 0, land.recstep)
    land.eflags[:] = np.where(sib, 0, land.eflags)
    states.suz[:] = np.where(p & upper, U(0, 10), 0.0)
    states.sg1[:] = np.where(p & upper, U(0, 20), 0.0)
    states.bw1[:] = np.where(c & upper, U(0, 20), 0.0)
    states.bw2[:] = np.where(c & upper, U(0, 40), 0.0)
    states.sg2[:] = np.where(ep_, rng.uniform(-2, 30, ne), 0.0)
    states.sg3[:] = np.where(ep_, rng.uniform(-2, 60, ne), 0.0)
    states.uz[:] = np.where(sib, 0.0, states.uz)
    states.lz[:] = np.where(ep_, 0.0, states.lz)
    # rconc_nash
    nash = rng.uniform(size=ne) < nash_fraction
    land.rconc[:] = np.where(nash, layout.RCONC_NASH, layout.RCONC_UH)
    steps = rng.integers(1, 9, size=ne)
    land.nash_steps[:] = np.where(nash, steps, 0)
    ksc = np.where(rng.uniform(size=ne) < 0.1, np.inf, rng.uniform(0.05, 2.5, ne))
    land.epar[EP["nash_ksc"]] = np.where(nash, ksc, 0.0)
    land.epar[EP["nash_dt"]] = np.where(nash, 1.0 / steps, 0.0)
    land.normalise()


def add_mct(river: RiverBundle, seed: int = 0, fraction: float = 0.5) -> np.ndarray:
    """Turn `fraction` of the reaches into musk_mct reaches with 1-3 trapeze cross sections, in place;
    returns their conditions [3, n_seg] (Courant numbers, Reynolds numbers, reference water depths — NaN = no start
    value yet, like a freshly prepared reference model)."""
    rng = np.random.default_rng(seed + 555)
    river.normalise()
    nr = river.n_reach
    is_mct = rng.uniform(size=nr) < fraction
    river.reach_model = np.where(is_mct, layout.REACH_MCT, layout.REACH_CLASSIC).astype(np.int32)
    river.nmbruns = np.where(is_mct, rng.integers(1, 3, size=nr), 1).astype(np.int32)
    RP, TP = layout.RP, layout.TP
    rpar = np.zeros((len(layout.RPAR), nr))
    rpar[RP["tolerancewaterdepth"]] = np.where(rng.uniform(size=nr) < 0.5, 0.0, 1e-4)
    rpar[RP["tolerancedischarge"]] = rng.uniform(1e-4, 2e-2, nr)
    rpar[RP["tolerancenegativeinflow"]] = np.where(rng.uniform(size=nr) < 0.5, -np.inf, -0.5)
    slope = rng.uniform(2e-4, 3e-3, nr)
    rpar[RP["bottomslope"]] = slope
    rpar[RP["wq_bottomslope"]] = slope
    rpar[RP["seconds"]] = 3600.0
    rpar[RP["segmentlength"]] = rng.uniform(1.0, 8.0, nr)
    river.rpar = rpar
    ntr = np.where(is_mct, rng.integers(1, 4, size=nr), 0)
    river.trap_ptr = np.concatenate([[0], np.cumsum(ntr)]).astype(np.int64)
    tp = np.zeros((len(layout.TPAR), int(river.trap_ptr[-1])))
    for q in range(nr):
        n = int(ntr[q])
        if not n:
            continue
        t0 = int(river.trap_ptr[q])
        heights = rng.uniform(0.5, 2.5, n)
        levels = np.concatenate([[0.0], np.cumsum(heights[:-1])])
        ss = rng.uniform(0.0, 6.0, n)
        th = np.full(n, np.inf)
        th[:-1] = np.diff(levels)
        ws = np.full(n, np.inf)
        ws[:-1] = 2.0 * ss[:-1] * th[:-1]
        tp[TP["bottomwidth"], t0:t0 + n] = np.concatenate([[rng.uniform(3.0, 20.0)], rng.uniform(0.0, 30.0, n - 1)])
        tp[TP["sideslope"], t0:t0 + n] = ss
        tp[TP["stricklercoefficient"], t0:t0 + n] = rng.uniform(15.0, 40.0, n)
        tp[TP["calibrationfactor"], t0:t0 + n] = rng.uniform(0.8, 1.2, n)
        tp[TP["bottomdepth"], t0:t0 + n] = levels - levels[0]
        tp[TP["trapezeheight"], t0:t0 + n] = th
        tp[TP["slopewidth"], t0:t0 + n] = ws
        tp[TP["perimeterderivative"], t0:t0 + n] = 2.0 * (1.0 + ss**2.0) ** 0.5
    river.tpar = tp
    river.normalise()
    mct = np.zeros((3, river.n_seg))
    for q in range(nr):
        if not is_mct[q]:
            continue
        g0, g1 = int(river.seg_ptr[q]), int(river.seg_ptr[q + 1]) - 1
        fresh = rng.uniform() < 0.4
        mct[0, g0:g1] = 0.0 if fresh else rng.uniform(0.2, 1.2, g1 - g0)
        mct[1, g0:g1] = 0.0 if fresh else rng.uniform(0.3, 3.0, g1 - g0)
        mct[2, g0:g1] = np.nan if fresh else rng.uniform(0.3, 2.0, g1 - g0)
    return mct


def random_forcing(nt: int, ncol: int, seed: int = 0) -> tuple[np.ndarray, np.ndarray]:
    rng = np.random.default_rng(seed + 1000)
    p = rng.gamma(0.4, 2.0, size=(nt, ncol)) * (rng.uniform(size=(nt, ncol)) < 0.5)
    t = rng.normal(1.0, 6.0, size=(nt, ncol))
    nat = t - rng.normal(0.0, 2.0, size=(nt, ncol))
    net = rng.uniform(0.0, 0.3, size=(nt, ncol))
    doy = ((np.arange(nt) // 24 + 40) % 366).astype(np.int64)
    return np.stack([p, t, nat, net]), doy


def random_river(n_land: int, n_reach: int, seed: int = 0, smax: int = 40, n_ext: int = 2
                 ) -> tuple[RiverBundle, np.ndarray, np.ndarray]:
    """Random DAG: nodes 0..n_land-1 are fed by the land elements, further nodes by reaches; some nodes use external
    rows (deploy modes oldsim/obs and obs_newsim with NaN gaps).  Returns (bundle, discharge, ext_mask_nan_rows)."""
    rng = np.random.default_rng(seed + 7)
    n_node = n_land + n_reach
    entries: list[list[int]] = [[e] for e in range(n_land)] + [[] for _ in range(n_reach)]
    inlet_ptr = [0]
    inlet_node: list[int] = []
    reach_outlet = np.zeros(n_reach, np.int32)
    for r in range(n_reach):
        target = n_land + r
        k = int(rng.integers(1, 4))
        src = rng.choice(np.arange(0, target), size=min(k, target), replace=False)
        inlet_node.extend(int(s) for s in src)
        inlet_ptr.append(len(inlet_node))
        reach_outlet[r] = target
        entries[target].append(-1 - r)
    entry_ptr = np.concatenate([[0], np.cumsum([len(x) for x in entries])]).astype(np.int64)
    entry_src = np.array([s for x in entries for s in x], np.int32)
    nseg = rng.integers(0, smax + 1, size=n_reach)
    nseg[rng.uniform(size=n_reach) < 0.2] = 0
    seg_ptr = np.concatenate([[0], np.cumsum(nseg + 1)]).astype(np.int64)
    k = rng.uniform(0.5, 3.0, n_reach); x = rng.uniform(0.0, 0.5, n_reach)
    den = k * (1 - x) + 0.5
    coef = np.stack([(0.5 - k * x) / den, (k * x + 0.5) / den, (k * (1 - x) - 0.5) / den])
    node_mode = np.zeros(n_node, np.int32)
    node_ext = np.full(n_node, -1, np.int32)
    picks = rng.choice(np.arange(n_node), size=min(n_ext, n_node), replace=False)
    for i, n in enumerate(picks):
        node_mode[n] = layout.NODE_EXT if i % 2 == 0 else layout.NODE_OBSNEWSIM
        node_ext[n] = i
    river = RiverBundle(entry_ptr=entry_ptr, entry_src=entry_src, node_mode=node_mode, node_ext=node_ext,
                        inlet_ptr=np.array(inlet_ptr, np.int64), inlet_node=np.array(inlet_node, np.int32),
                        reach_outlet=reach_outlet, seg_ptr=seg_ptr, coef=coef, n_ext=len(picks))
    river.normalise()
    discharge = rng.uniform(0, 5, river.n_seg)
    return river, discharge, picks
