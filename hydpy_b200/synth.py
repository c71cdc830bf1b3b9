"""Synthetic basins of BASELINE.json's configs, generated directly as SoA bundles (bulk API).

Constructing 10^5 parameterised `hland_96` elements through the reference's object model takes >15 min and ≈19 GB
(hard part 7), so the large configurations are fed to the engine as flat arrays.  All values are the
reference's INTERNAL values (after `update_parameters()`, time-scaled to the simulation step): the generators start
from a template that was extracted from a real reference element (HydPy-H-Lahn `land_dill_assl`, 12 zones
FIELD/FOREST, `hydpy_b200/data/template.npz`, written by tests/golden/make_golden.py) and apply the recipe of
SURVEY.md §8(d):

* zones: 12 per element, `zonearea = area/12`, `zonez = linspace(2, 7, 12)`, optional 5 % GLACIER/ILAKE/SEALED variants;
* Lahn-like control values jittered ±10 % per element (`default_rng(1)`), `smax = inf`, `sred = 0`;
* forcing bank of 1024 series (seed 0), element e reads column e % 1024;
* river network (config 4): random tree (seed 2) with ≈200 topological levels, one `musk_classic` reach per
  non-outlet node, `lag ~ U(0, 1) d`, `damp ~ U(0, 1)`.
"""
from __future__ import annotations

import os

import numpy as np

from . import layout
from .abi import LandBundle, LandStates, RiverBundle

HERE = os.path.dirname(os.path.abspath(__file__))
TEMPLATE = os.path.join(HERE, "data", "template.npz")

#: parameters jittered per element (zone-level ones get one factor per element, like a regional calibration factor)
JITTER_ZONE = ("pcorr", "rfcf", "sfcf", "icmax", "cfmax", "cfr", "whc", "fc", "beta")
JITTER_ELEM = ("percmax", "alpha", "k", "k4")


def load_template(step: str) -> dict[str, np.ndarray]:
    """Internal parameter values of one real 12-zone reference element for simulation step '1h' or '1d'."""
    if step not in ("1h", "1d"):
        raise ValueError("step must be '1h' or '1d'")
    with np.load(TEMPLATE) as z:
        return {k[len(step) + 1:]: z[k] for k in z.files if k.startswith(step + "/")}


def make_land(n_elem: int, step: str = "1h", *, zones: int = 12, seed: int = 1, bank: int = 1024,
              variants: bool = False, jitter: float = 0.1, area: float = 100.0,
              own_nodes: bool = True, cflux: float = 0.0, model: str = "hland_96",
              select: np.ndarray | None = None) -> tuple[LandBundle, LandStates]:
    """`n_elem` synthetic hland_96 elements (configs 3-5).  `cflux` > 0 (mm per simulation step, jittered like the other
    parameters) switches the capillary flux on, i.e. couples the zones to the element's upper zone storage.

    `select` (ascending global element indices): build only these elements of the `n_elem`-element basin — the random
    draws are those of the whole basin, so a shard of a large basin gets exactly the elements the whole would have
    (forcing columns follow the global index; the outlet nodes are numbered locally, 0 … len(select)-1)."""
    tpl = load_template(step)
    tz = tpl["zpar"].shape[1]
    rng = np.random.default_rng(seed)
    n_all, nz = int(n_elem), int(zones)
    sel = np.arange(n_all) if select is None else np.asarray(select, np.int64)
    ne = len(sel)

    def draw(size) -> np.ndarray:     # one draw per element of the WHOLE basin, then the selection
        return rng.uniform(-1.0, 1.0, size=size)[sel]

    idx = np.arange(nz) % tz                      # template zone used for every slot (FIELD/FOREST alternating)
    zpar = np.repeat(tpl["zpar"][:, idx][:, None, :], ne, axis=1)        # [P, ne, nz]
    zonetype = np.repeat(tpl["zonetype"][idx][None, :], ne, axis=0).astype(np.int32)
    zflags = np.repeat(tpl["zflags"][idx][None, :], ne, axis=0).astype(np.int32)
    ZP = layout.ZP
    for name in JITTER_ZONE:
        f = 1.0 + jitter * draw((n_all, 1))
        zpar[ZP[name]] *= f
    zpar[ZP["tt"]] += jitter * draw((n_all, 1))
    zpar[ZP["ttm"]] = zpar[ZP["tt"]]                                      # dttm = 0 as in the Lahn project
    zpar[ZP["maxsoilwater"]] = zpar[ZP["fc"]]
    zpar[ZP["smax"]] = np.inf
    zpar[ZP["cflux"]] = 0.0 * zpar[ZP["cflux"]] + cflux * (1.0 + jitter * draw((n_all, 1)))
    zonez = np.linspace(2.0, 7.0, nz)
    zpar[ZP["zonez"]] = zonez[None, :]
    zpar[ZP["hrualtitude"]] = 100.0 * zonez[None, :]
    if variants:
        # 5 % of the elements get one GLACIER, one ILAKE and one SEALED zone (last three slots)
        special = rng.uniform(size=n_all)[sel] < 0.05
        for slot, ztype, flags in ((nz - 1, layout.GLACIER, 0), (nz - 2, layout.ILAKE, layout.ZF_WATER),
                                   (nz - 3, layout.SEALED, layout.ZF_INTERCEPTION)):
            zonetype[special, slot] = ztype
            zflags[special, slot] = flags
        zpar[ZP["gmelt"]] = np.where(zonetype == layout.GLACIER, zpar[ZP["cfmax"]] * 1.5, zpar[ZP["gmelt"]])
        zpar[ZP["temperaturethresholdice"]] = np.where(zonetype == layout.ILAKE, 0.0, zpar[ZP["temperaturethresholdice"]])
    rel = np.full((ne, nz), 1.0 / nz)
    zpar[ZP["relzonearea"]] = rel
    zpar[ZP["hruareafraction"]] = rel

    def rsum(mask: np.ndarray) -> np.ndarray:
        return np.sum(np.where(mask, rel, 0.0), axis=1)

    soil = (zonetype == layout.FIELD) | (zonetype == layout.FOREST)
    epar = np.repeat(tpl["epar"][:, :1], ne, axis=1)
    EP = layout.EP
    for name in JITTER_ELEM:
        epar[EP[name]] *= 1.0 + jitter * draw(n_all)
    epar[EP["z"]] = np.sum(rel * zonez[None, :], axis=1)
    epar[EP["relsoilarea"]] = rsum(soil)
    epar[EP["rellandarea"]] = rsum(zonetype != layout.ILAKE)
    epar[EP["relupperzonearea"]] = rsum(soil | (zonetype == layout.GLACIER))
    epar[EP["rellowerzonearea"]] = rsum(zonetype != layout.SEALED)
    epar[EP["petaltitude"]] = np.sum(rel * 100.0 * zonez[None, :], axis=1)
    seconds = 3600.0 if step == "1h" else 86400.0
    epar[EP["qfactor"]] = area * 1000.0 / seconds
    nu = len(tpl["uh"])
    zorder = np.repeat(np.argsort(zonez, kind="stable")[None, :], ne, axis=0).astype(np.int32)
    land = LandBundle(
        zone_ptr=np.arange(ne + 1, dtype=np.int64) * nz, snow_ptr=np.arange(ne + 1, dtype=np.int64) * nz,
        sfdist_ptr=np.arange(ne + 1, dtype=np.int64), uh_ptr=np.arange(ne + 1, dtype=np.int64) * nu,
        sred_ptr=np.zeros(ne + 1, np.int64), recstep=np.full(ne, int(tpl["recstep"][0]), np.int32),
        eflags=np.full(ne, int(tpl["eflags"][0]), np.int32), forcing_col=(sel % bank).astype(np.int32),
        outlet_node=(np.arange(ne) if own_nodes else np.full(ne, -1)).astype(np.int32), epar=epar,
        zonetype=zonetype.reshape(-1), zflags=zflags.reshape(-1), zorder=zorder.reshape(-1),
        zpar=zpar.reshape(zpar.shape[0], ne * nz), sfdist=np.ones(ne), uh=np.tile(tpl["uh"], ne),
        sred_from=np.zeros(0, np.int32), sred_to=np.zeros(0, np.int32), sred_adjust=np.zeros(0),
    )
    land.normalise()
    fc = land.zp("fc")
    lakeless = land.zonetype != layout.ILAKE
    states = LandStates(
        ic=np.where(soil.reshape(-1), 0.5, 0.0) * np.ones(ne * nz), sm=np.where(soil.reshape(-1), 0.7 * fc, 0.0),
        sp=np.where(lakeless, 5.0, 0.0) * np.ones(ne * nz), wc=np.where(lakeless, 0.2, 0.0) * np.ones(ne * nz),
        uz=np.full(ne, 5.0), lz=np.full(ne, 8.0), quh=np.zeros(ne * nu),
    )
    if model != "hland_96":
        to_sibling(land, states, model, step, jitter=jitter, seed=seed)
    return land, states


def to_sibling(land: LandBundle, states: LandStates, model: str, step: str = "1h", *, jitter: float = 0.1,
               seed: int = 1, nash: bool = True) -> None:
    """Turn a synthetic hland_96 bundle into `hland_96p` / `hland_96c` elements, in place: the snow
    and soil parameters stay, the runoff response gets the values of the reference's integration tests
    (ref: hydpy/models/hland_96p.py:83-95, hland_96c.py:83-96, internal = simulation-step related values), jittered per
    element; `nash` swaps the unit hydrograph for a `rconc_nash` cascade of as many storages."""
    land.normalise()
    states.normalise()
    rng = np.random.default_rng(seed + 77)
    ne, nzt = land.n_elem, land.n_zone
    per_day = 24.0 if step == "1h" else 1.0
    ZP, EP = layout.ZP, layout.EP
    zone_elem = np.repeat(np.arange(ne), np.diff(land.zone_ptr))
    jit = lambda: (1.0 + jitter * rng.uniform(-1.0, 1.0, size=ne))  # noqa: E731
    upper = np.isin(land.zonetype, (layout.FIELD, layout.FOREST, layout.GLACIER))
    land.model[:] = layout.MODELS[model]
    land.eflags[:] = 0
    land.recstep[:] = 0
    land.epar[EP["dt"]] = 1.0
    land.zpar[ZP["cflux"]] = 0.0
    if model == "hland_96p":
        k0, k1, k2 = 2.0 * per_day * jit(), 10.0 * per_day * jit(), 20.0 * per_day * jit()
        k3, k4 = 100.0 * per_day * jit(), 300.0 * per_day * jit()
        land.zpar[ZP["sgr"]] = (20.0 * jit())[zone_elem]
        land.zpar[ZP["sg1max"]] = (50.0 * jit())[zone_elem]
        land.zpar[ZP["w0"]] = np.exp(-1.0 / k0)[zone_elem]
        land.zpar[ZP["w1"]] = np.exp(-1.0 / k1)[zone_elem]
        land.zpar[ZP["k2"]] = k2[zone_elem]
        land.zpar[ZP["w2"]] = np.exp(-1.0 / k2)[zone_elem]
        land.epar[EP["percmax"]] = 2.0 / per_day * jit()
        land.epar[EP["k3"]], land.epar[EP["w3"]] = k3, np.exp(-1.0 / k3)
        land.epar[EP["k4"]], land.epar[EP["w4"]] = k4, np.exp(-1.0 / k4)
        land.epar[EP["fsg"]] = 8.0 / 9.0
        states.suz[:] = np.where(upper, 1.0, 0.0)
        states.sg1[:] = np.where(upper, 10.0, 0.0)
        states.sg2[:] = 10.0
        states.sg3[:] = 10.0
        states.uz[:] = 0.0
        states.lz[:] = 0.0
    elif model == "hland_96c":
        for name, days in (("tab1", 2.0), ("tvs1", 2.0), ("tab2", 10.0), ("tvs2", 10.0)):
            land.zpar[ZP[name]] = (days * per_day * jit())[zone_elem]
        land.zpar[ZP["h1"]] = (10.0 * jit())[zone_elem]
        land.zpar[ZP["h2"]] = (10.0 * jit())[zone_elem]
        states.bw1[:] = np.where(upper, 2.0, 0.0)
        states.bw2[:] = np.where(upper, 3.0, 0.0)
        states.uz[:] = 0.0
    else:
        raise ValueError(f"unknown sibling model {model!r}")
    if nash:
        nst = np.diff(land.uh_ptr)
        land.rconc[:] = np.where(nst > 0, layout.RCONC_NASH, layout.RCONC_UH)
        steps = 10
        land.nash_steps[:] = np.where(nst > 0, steps, 0)
        # retention time 3 h spread over the storages (ref: rconc_derived.KSC: ksc = nmbstorages / retentiontime)
        land.epar[EP["nash_ksc"]] = np.where(nst > 0, nst / (3.0 * jit()), 0.0)
        land.epar[EP["nash_dt"]] = np.where(nst > 0, 1.0 / steps, 0.0)
        land.uh[:] = 0.0
        states.quh[:] = 0.05
    land.normalise()


#: mean daily precipitation depth of the "storms" forcing = 0.3 * STORM_SCALE mm (Gamma(shape 0.3) keeps the survey's
#: wet-day statistics; the scale puts the annual sum at ≈ 880 mm against ≈ 600 mm of potential evapotranspiration, the
#: humid balance of the Lahn basin the parameters come from)
STORM_SCALE = 8.0


def make_forcing(nt: int, step: str = "1h", *, bank: int = 1024, seed: int = 0, t0: int = 0, kind: str = "storms"
                 ) -> tuple[np.ndarray, np.ndarray]:
    """Forcing bank [4, nt, bank] + doy[nt] for steps t0 … t0+nt-1.

    kind = "storms" (default): every DAY draws, per bank column, a precipitation depth ~ Gamma(0.3, STORM_SCALE) mm, an
    event of 1-6 consecutive wet hours inside the day that carries it (hourly step), a temperature anomaly ~ N(0, 3) K on
    top of the seasonal normal 8 + 10 sin(2 pi doy / 365) °C (plus a diurnal cycle of ±3 K in hourly runs), and a seasonal
    normal evapotranspiration of 1.6 + 1.3 sin(2 pi doy / 365) mm/d.  Days are seeded with (seed, day), so any window of
    any length reproduces the same series.
    kind = "dry": the first round's forcing — an independent Gamma(0.3, 3)/24 drizzle in every hour, which the canopy
    evaporates completely (0.9 mm/d against 2 mm/d of potential evapotranspiration: a basin in terminal drought).
    """
    per_day = 24 if step == "1h" else 1
    steps = np.arange(t0, t0 + nt)
    day = steps // per_day
    doy = (day % 365).astype(np.int64)
    if kind == "dry":
        rng = np.random.default_rng([seed, t0])
        p = rng.gamma(0.3, 3.0, size=(nt, bank)) / per_day
        t = 8.0 + 10.0 * np.sin(2.0 * np.pi * doy / 365.0)[:, None] + rng.normal(0.0, 3.0, size=(nt, bank))
        nat = np.full((nt, bank), 8.0)
        net = np.full((nt, bank), 2.0 / per_day)
        return np.stack([p, t, nat, net]), doy
    if kind != "storms":
        raise ValueError(f"unknown forcing kind {kind!r}")
    season = np.sin(2.0 * np.pi * doy / 365.0)
    p = np.zeros((nt, bank))
    t = np.zeros((nt, bank))
    hour = steps % per_day
    for d in range(int(day[0]) if nt else 0, int(day[-1]) + 1 if nt else 0):
        rng = np.random.default_rng([seed, 7919, d])
        depth = rng.gamma(0.3, STORM_SCALE, size=bank)
        dur = rng.integers(1, 7, size=bank)
        start = (rng.uniform(size=bank) * (per_day - dur + 1)).astype(np.int64)
        anomaly = rng.normal(0.0, 3.0, size=bank)
        rows = np.nonzero(day == d)[0]
        if per_day == 1:
            p[rows] = depth
        else:
            h = hour[rows][:, None]
            wet = (h >= start[None, :]) & (h < (start + dur)[None, :])
            p[rows] = np.where(wet, (depth / dur)[None, :], 0.0)
        t[rows] = anomaly[None, :]
    diurnal = 3.0 * np.sin(2.0 * np.pi * (hour - 9.0) / 24.0) if per_day > 1 else np.zeros(nt)
    normal_t = 8.0 + 10.0 * season
    t += (normal_t + diurnal)[:, None]
    nat = np.repeat(normal_t[:, None], bank, axis=1)
    net = np.repeat(((1.6 + 1.3 * season) / per_day)[:, None], bank, axis=1)
    return np.stack([p, t, nat, net]), doy


def make_tree_river(n_elem: int, step: str = "1h", *, seed: int = 2, levels: int = 200
                    ) -> tuple[RiverBundle, np.ndarray]:
    """Random river tree for config 4: node i is fed by land element i and by the reaches of its children; reach i-1
    routes node i to its parent.  The parent window is chosen so that the tree is ≈`levels` reaches deep."""
    ne = int(n_elem)
    rng = np.random.default_rng(seed)
    window = max(1, int(round(2.0 * ne / max(levels, 1))))
    parent = np.zeros(ne, np.int64)
    if ne > 1:
        i = np.arange(1, ne)
        lo = np.maximum(0, i - window)
        parent[1:] = lo + (rng.uniform(size=ne - 1) * (i - lo)).astype(np.int64)
    nr = ne - 1
    per_day = 24.0 if step == "1h" else 1.0
    lag = rng.uniform(0.0, 1.0, size=nr) * per_day          # in simulation steps
    nseg = np.rint(lag).astype(np.int64)                    # ref: musk_control.py:93-95 int(round(lag)), banker's
    damp = rng.uniform(0.0, 1.0, size=nr)
    c0 = damp / (1.0 + damp)                                # ref: musk_control.py:288-291
    coef = np.stack([c0, 1.0 - 2.0 * c0, c0])
    # node entries: land element first, then the child reaches in ascending order
    children: list[list[int]] = [[] for _ in range(ne)]
    for child in range(1, ne):
        children[parent[child]].append(child - 1)           # reach index = child node - 1
    entry_ptr = np.zeros(ne + 1, np.int64)
    entry_src: list[int] = []
    for n in range(ne):
        entry_src.append(n)
        entry_src.extend(-1 - r for r in children[n])
        entry_ptr[n + 1] = len(entry_src)
    seg_ptr = np.zeros(nr + 1, np.int64)
    seg_ptr[1:] = np.cumsum(nseg + 1)
    river = RiverBundle(
        entry_ptr=entry_ptr, entry_src=np.array(entry_src, np.int32), node_mode=np.zeros(ne, np.int32),
        node_ext=np.full(ne, -1, np.int32), inlet_ptr=np.arange(nr + 1, dtype=np.int64),
        inlet_node=np.arange(1, ne, dtype=np.int32), reach_outlet=parent[1:].astype(np.int32), seg_ptr=seg_ptr,
        coef=coef, n_ext=0,
    )
    river.normalise()
    return river, np.zeros(river.n_seg)


def tree_depth(river: RiverBundle) -> int:
    """Number of reaches on the longest path (diagnostics for the generator)."""
    depth = np.zeros(river.n_node, np.int64)
    # reach r: inlet node r+1 → outlet parent; nodes are ordered so that parents have smaller indices
    for r in range(river.n_reach - 1, -1, -1):
        child, par = int(river.inlet_node[r]), int(river.reach_outlet[r])
        depth[par] = max(depth[par], depth[child] + 1)
    return int(depth.max()) if river.n_node else 0


def replicate_members(land: LandBundle, states: LandStates, members: int, *, seed: int = 20260101,
                      jitter: float = 0.25, n_node: int | None = None) -> tuple[LandBundle, LandStates]:
    """Calibration-style ensemble (config 2): every member is a complete copy of the basin with its own parameter
    set (member 0 = the given values), sharing structure and forcing columns.  Members become additional elements."""
    m = int(members)
    rng = np.random.default_rng(seed)
    land.normalise()
    ne, nz = land.n_elem, land.n_zone
    node_stride = int(n_node) if n_node is not None else (int(land.outlet_node.max()) + 1 if ne else 0)

    def tile_ptr(p: np.ndarray) -> np.ndarray:
        total = int(p[-1])
        return np.concatenate([p[:-1] + i * total for i in range(m)] + [np.array([m * total])]).astype(np.int64)

    zpar = np.tile(land.zpar, (1, m)).reshape(len(layout.ZPAR), m, nz)
    epar = np.tile(land.epar, (1, m)).reshape(len(layout.EPAR), m, ne)
    ZP, EP = layout.ZP, layout.EP
    f = lambda: 1.0 + jitter * rng.uniform(-1.0, 1.0, size=(m, 1))  # noqa: E731
    for name in ("fc", "beta", "cfmax", "whc"):
        fac = f()
        fac[0] = 1.0
        zpar[ZP[name]] *= fac
    zpar[ZP["maxsoilwater"]] = zpar[ZP["fc"]]
    shift = jitter * rng.uniform(-1.0, 1.0, size=(m, 1))
    shift[0] = 0.0
    zpar[ZP["tt"]] += shift
    zpar[ZP["ttm"]] += shift
    for name in ("percmax", "k", "k4", "alpha"):
        fac = f()
        fac[0] = 1.0
        epar[EP[name]] *= fac
    out = LandBundle(
        zone_ptr=tile_ptr(land.zone_ptr), snow_ptr=tile_ptr(land.snow_ptr), sfdist_ptr=tile_ptr(land.sfdist_ptr),
        uh_ptr=tile_ptr(land.uh_ptr), sred_ptr=tile_ptr(land.sred_ptr), recstep=np.tile(land.recstep, m),
        eflags=np.tile(land.eflags, m), forcing_col=np.tile(land.forcing_col, m),
        outlet_node=np.concatenate([np.where(land.outlet_node >= 0, land.outlet_node + i * node_stride, -1)
                                    for i in range(m)]).astype(np.int32) if ne else np.zeros(0, np.int32),
        epar=epar.reshape(len(layout.EPAR), m * ne), zonetype=np.tile(land.zonetype, m), zflags=np.tile(land.zflags, m),
        zorder=np.tile(land.zorder, m), zpar=zpar.reshape(len(layout.ZPAR), m * nz), sfdist=np.tile(land.sfdist, m),
        uh=np.tile(land.uh, m), sred_from=np.tile(land.sred_from, m), sred_to=np.tile(land.sred_to, m),
        sred_adjust=np.tile(land.sred_adjust, m), model=np.tile(land.model, m), rconc=np.tile(land.rconc, m),
        nash_steps=np.tile(land.nash_steps, m),
    )
    out.normalise()
    states.normalise()
    st = LandStates(**{n: np.tile(getattr(states, n), m) for n in LandStates.NAMES})
    return out, st


def member_mods(members: int, *, seed: int = 20260101, jitter: float = 0.25) -> dict[str, tuple[np.ndarray, np.ndarray]]:
    """The parameter sets of `replicate_members` as modifications for `Engine.set_members` (`hb_set_members`): parameter
    name → (mul[members], add[members]), member 0 = the given values.  Same random draws, same arithmetic (one
    multiplication or one addition per value), so the ensemble built on the device equals the replicated bundle bit for
    bit."""
    m = int(members)
    rng = np.random.default_rng(seed)
    one, zero = np.ones(m), np.zeros(m)
    mods: dict[str, tuple[np.ndarray, np.ndarray]] = {}

    def fac() -> np.ndarray:
        f = 1.0 + jitter * rng.uniform(-1.0, 1.0, size=(m, 1))
        f[0] = 1.0
        return f[:, 0]

    for name in ("fc", "beta", "cfmax", "whc"):
        mods[name] = (fac(), zero)
    mods["maxsoilwater"] = mods["fc"]
    shift = jitter * rng.uniform(-1.0, 1.0, size=(m, 1))
    shift[0] = 0.0
    mods["tt"] = (one, shift[:, 0])
    mods["ttm"] = (one, shift[:, 0])
    for name in ("percmax", "k", "k4", "alpha"):
        mods[name] = (fac(), zero)
    return mods


def replicate_river(river: RiverBundle, discharge: np.ndarray, members: int, n_land: int
                    ) -> tuple[RiverBundle, np.ndarray]:
    """Copy the river network once per ensemble member (each member routes its own discharges)."""
    m = int(members)
    river.normalise()
    nn, nr = river.n_node, river.n_reach

    def tile_ptr(p: np.ndarray) -> np.ndarray:
        total = int(p[-1])
        return np.concatenate([p[:-1] + i * total for i in range(m)] + [np.array([m * total])]).astype(np.int64)

    entry_src = np.concatenate([
        np.where(river.entry_src >= 0, river.entry_src + i * n_land, river.entry_src - i * nr) for i in range(m)
    ]).astype(np.int32) if len(river.entry_src) else np.zeros(0, np.int32)
    out = RiverBundle(
        entry_ptr=tile_ptr(river.entry_ptr), entry_src=entry_src, node_mode=np.tile(river.node_mode, m),
        node_ext=np.tile(river.node_ext, m), inlet_ptr=tile_ptr(river.inlet_ptr),
        inlet_node=np.concatenate([river.inlet_node + i * nn for i in range(m)]).astype(np.int32),
        reach_outlet=np.concatenate([np.where(river.reach_outlet >= 0, river.reach_outlet + i * nn, -1)
                                     for i in range(m)]).astype(np.int32),
        seg_ptr=tile_ptr(river.seg_ptr), coef=np.tile(river.coef, (1, m)), n_ext=river.n_ext,
        reach_model=np.tile(river.reach_model, m), nmbruns=np.tile(river.nmbruns, m), rpar=np.tile(river.rpar, (1, m)),
        trap_ptr=tile_ptr(river.trap_ptr), tpar=np.tile(river.tpar, (1, m)),
    )
    out.normalise()
    return out, np.tile(discharge, m)
