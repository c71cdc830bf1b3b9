"""HydPy objects  →  SoA bundles, and results  →  HydPy objects.

This is the staging layer that replaces the reference's per-step `load_data`/`save_data`
(ref: hydpy/cythons/modelutils.py:1080-1167) and the pointer exchange between nodes and models
(ref: hydpy/core/modeltools.py:2550-2584, hydpy/cythons/pointerutils.pyx) for the covered models:

* values are read from `model.parameters.<group>.<name>.value` after `update_parameters()` has run, i.e. the
  reference's internal, simulation-step related values;
* conditions come from `model.sequences.states.*` and `rconcmodel.sequences.logs.quh`;
* forcing comes from `inputs.<name>.series` (must be in RAM — NetCDF just-in-time reading is refused exactly like
  the reference's threaded mode does, ref: hydpy/core/hydpytools.py:3007-3015).

Nothing here imports `hydpy`: all objects are duck-typed, so the module also works against look-alike objects.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Iterable

import numpy as np

from . import layout
from .abi import LandBundle, LandStates, RiverBundle

LAND_MODELS = ("hland_96", "hland_96p", "hland_96c")
RIVER_MODELS = ("musk_classic", "musk_mct")

_ZONE_CONTROL = (
    "zonez", "pcorr", "pcalt", "rfcf", "sfcf", "tcorr", "tcalt", "icmax", "smax", "tt", "ttint", "cfmax", "cfvar",
    "gmelt", "gvar", "cfr", "whc", "fc", "beta",
)
# model-specific zone parameters: (control names, derived names)
_ZONE_EXTRA = {
    "hland_96": (("cflux",), ()),
    "hland_96p": (("sgr", "k2", "sg1max"), ("w0", "w1", "w2")),
    "hland_96c": (("h1", "tab1", "tvs1", "h2", "tab2", "tvs2"), ()),
}
_AET_CONTROL = ("temperaturethresholdice", "maxsoilwater", "soilmoisturelimit", "excessreduction")
_PET_CONTROL = (
    "hrualtitude", "evapotranspirationfactor", "altitudefactor", "precipitationfactor", "airtemperaturefactor",
)
#: PET submodels of `evap_aet_hbv96` the engine serves.  `evap_ret_io` (reference evapotranspiration from an external
#: time series, adjusted per zone; ref: hydpy/models/evap_ret_io.py:53-64, evap_model.py:4987-4993, 5420-5424) runs on
#: the same kernels: its input series takes the place of NormalEvapotranspiration, and with AirTemperatureFactor =
#: AltitudeFactor = PrecipitationFactor = 0 the evap_pet_hbv96 equations return  ret * 1.0 * exp(-0.0) = ret  bit for
#: bit (for ret >= 0, which `_pet_inputs` checks: the HBV96 clipping to [0, 2 * NormalEvapotranspiration] is the only
#: step that would treat a negative value differently).
_PET_MODELS = ("evap_pet_hbv96", "evap_ret_io")


class UnsupportedModelError(NotImplementedError):
    """The project contains something the B200 engine does not cover (no silent fallback)."""


def _f(par: Any, n: int | None = None) -> np.ndarray:
    a = np.array(par.value, dtype=np.float64)
    if n is not None:
        a = np.broadcast_to(a, (n,)).copy() if a.ndim == 0 else a
        if a.shape != (n,):
            raise ValueError(f"parameter `{par.name}` has shape {a.shape}, expected ({n},)")
    return a


def _modelname(model: Any) -> str:
    return getattr(model, "name", type(model).__module__.rsplit(".", 1)[-1])


@dataclasses.dataclass
class Problem:
    """Everything one engine run needs, plus the objects the results go back to."""

    land: LandBundle
    states: LandStates
    river: RiverBundle
    discharge: np.ndarray                 # musk_classic / musk_mct states.discharge, concatenated
    forcing: np.ndarray                   # [4, T, n_col]  (T = i1-i0)
    doy: np.ndarray                       # [T] int64
    ext: np.ndarray                       # [n_ext, T]
    land_elements: list[Any] = dataclasses.field(default_factory=list)
    river_elements: list[Any] = dataclasses.field(default_factory=list)
    nodes: list[Any] = dataclasses.field(default_factory=list)
    node_writeback: np.ndarray | None = None  # bool[n_node]: does the run overwrite sim.series of the node?
    i0: int = 0
    i1: int = 0
    mct_states: np.ndarray | None = None  # [3, n_seg] musk_mct courantnumber, reynoldsnumber, referencewaterdepth

    @property
    def nt(self) -> int:
        return int(self.forcing.shape[1])

    def to_npz_dict(self) -> dict[str, np.ndarray]:
        d = {**self.land.to_dict(), **self.river.to_dict()}
        for n in LandStates.NAMES:
            d["state_" + n] = getattr(self.states, n)
        d["discharge"] = self.discharge
        d["forcing"] = self.forcing
        d["doy"] = self.doy
        d["ext"] = self.ext
        if self.node_writeback is not None:
            d["node_writeback"] = self.node_writeback
        if self.mct_states is not None:
            d["mct_states"] = self.mct_states
        return d

    @classmethod
    def from_npz_dict(cls, d: Any) -> "Problem":
        from .abi import states_from_dict

        return cls(
            land=LandBundle.from_dict(d), states=states_from_dict(d), river=RiverBundle.from_dict(d),
            discharge=np.array(d["discharge"], dtype=np.float64), forcing=np.array(d["forcing"], dtype=np.float64),
            doy=np.array(d["doy"], dtype=np.int64), ext=np.array(d["ext"], dtype=np.float64),
            node_writeback=np.array(d["node_writeback"], dtype=bool) if "node_writeback" in d else None,
            mct_states=np.array(d["mct_states"], dtype=np.float64) if "mct_states" in d else None,
        )


def _check_land_model(element: Any) -> None:
    model = element.model
    aet = getattr(model, "aetmodel", None)
    if aet is None or _modelname(aet) != "evap_aet_hbv96":
        raise UnsupportedModelError(
            f"element `{element.name}`: {_modelname(model)} needs an `evap_aet_hbv96` AET submodel, found "
            f"`{None if aet is None else _modelname(aet)}`"
        )
    pet = getattr(aet, "petmodel", None)
    if pet is None or _modelname(pet) not in _PET_MODELS:
        raise UnsupportedModelError(
            f"element `{element.name}`: evap_aet_hbv96 needs an `evap_pet_hbv96` or `evap_ret_io` PET submodel, found "
            f"`{None if pet is None else _modelname(pet)}`"
        )
    if getattr(aet, "snowycanopymodel", None) is not None:
        raise UnsupportedModelError(f"element `{element.name}`: snowy-canopy submodels are not supported")
    rconc = getattr(model, "rconcmodel", None)
    if rconc is not None and _modelname(rconc) not in ("rconc_uh", "rconc_nash"):
        raise UnsupportedModelError(
            f"element `{element.name}`: only `rconc_uh`, `rconc_nash` (or no) runoff concentration submodel is supported, "
            f"found `{_modelname(rconc)}`"
        )
    for seq in (model.sequences.inputs.p, model.sequences.inputs.t, *_pet_input_sequences(pet)):
        if seq is None:
            continue
        if getattr(seq, "node2idx", None):
            continue                                  # fed by an input node: see `_input_series`
        if not seq.ramflag:
            raise RuntimeError(
                f"element `{element.name}`: the time series of input sequence `{seq.name}` is not available in "
                f"RAM (just-in-time NetCDF reading is not possible with the B200 engine)"
            )
    if len(tuple(getattr(element, "outputs", ()))):
        raise UnsupportedModelError(f"element `{element.name}`: output nodes are not supported")


def _pet_input_sequences(pet: Any) -> tuple[Any, Any]:
    """The input sequences behind forcing rows 2 and 3 (NormalAirTemperature, NormalEvapotranspiration) of a PET
    submodel; `evap_ret_io` has no temperature input (row 2 stays zero) and its ReferenceEvapotranspiration is row 3."""
    inp = pet.sequences.inputs
    if _modelname(pet) == "evap_ret_io":
        return None, inp.referenceevapotranspiration
    return inp.normalairtemperature, inp.normalevapotranspiration


def _input_series(element: Any, seq: Any, i0: int, i1: int) -> np.ndarray:
    """The values `load_data` hands to one input sequence during [i0, i1): its own series or, when an input node feeds it
    (`inputflag`, ref: hydpy/cythons/modelutils.py:1080-1125 + hydpy/core/devicetools.py `Node.get_double("inputs")`),
    what that node holds at each step.  Per deploy mode (ref: hydpy/core/hydpytools.py:2704-2722, 2726-2733):
    `oldsim` → sim.series; `obs` → obs.series; `obs_oldsim` → obs.series where available, else sim.series; `newsim` /
    `obs_newsim` → the node is reset to zero at every step (nothing inside the run feeds it), observations on top."""
    nodes = list(getattr(seq, "node2idx", None) or ())
    if not nodes:
        return np.array(seq.series[i0:i1], dtype=np.float64)
    if len(nodes) > 1:
        raise UnsupportedModelError(
            f"element `{element.name}`: input sequence `{seq.name}` is connected to {len(nodes)} nodes")
    node = nodes[0]
    if len(tuple(getattr(node, "entries", ()))):
        raise UnsupportedModelError(
            f"element `{element.name}`: input node `{node.name}` is fed by other elements within the run, which the "
            f"B200 engine does not support")
    mode = str(node.deploymode)
    nt = i1 - i0

    def series(name: str) -> np.ndarray:
        sub = getattr(node.sequences, name)
        if not sub.ramflag:
            raise RuntimeError(f"input node `{node.name}`: the time series of `{name}` is not available in RAM")
        return np.array(sub.series[i0:i1], dtype=np.float64)

    if mode in ("oldsim", "oldsim_bi"):
        return series("sim")
    if mode in ("obs", "obs_bi"):
        return series("obs")
    if mode == "obs_oldsim":
        obs = series("obs")
        return np.where(np.isnan(obs), series("sim"), obs)
    if mode == "newsim":
        return np.zeros(nt)
    if mode == "obs_newsim":
        obs = series("obs")
        return np.where(np.isnan(obs), 0.0, obs)
    raise UnsupportedModelError(f"input node `{node.name}`: deploy mode `{mode}` is not supported")


def extract_land(elements: Iterable[Any], node_index: dict[Any, int], i0: int, i1: int
                 ) -> tuple[LandBundle, LandStates, np.ndarray, np.ndarray]:
    """Gather all hland_96 elements into one `LandBundle` (+ conditions, forcing [4,T,E], doy[T])."""
    elements = list(elements)
    ne = len(elements)
    zone_ptr = np.zeros(ne + 1, np.int64)
    snow_ptr = np.zeros(ne + 1, np.int64)
    sfdist_ptr = np.zeros(ne + 1, np.int64)
    uh_ptr = np.zeros(ne + 1, np.int64)
    sred_ptr = np.zeros(ne + 1, np.int64)
    recstep = np.zeros(ne, np.int32)
    eflags = np.zeros(ne, np.int32)
    outlet_node = np.full(ne, -1, np.int32)
    epar = np.zeros((len(layout.EPAR), ne))
    zt_l, zf_l, zo_l, zp_l, sf_l, uh_l, sr_f, sr_t, sr_a = [], [], [], [], [], [], [], [], []
    ic_l, sm_l, sp_l, wc_l, quh_l = [], [], [], [], []
    suz_l, sg1_l, bw1_l, bw2_l = [], [], [], []
    uz = np.zeros(ne)
    lz = np.zeros(ne)
    sg2 = np.zeros(ne)
    sg3 = np.zeros(ne)
    model_id = np.zeros(ne, np.int32)
    rconc_kind = np.zeros(ne, np.int32)
    nash_steps = np.zeros(ne, np.int32)
    nt = i1 - i0
    forcing = np.zeros((4, nt, ne))
    doy = None
    for e, element in enumerate(elements):
        _check_land_model(element)
        model = element.model
        con, der = model.parameters.control, model.parameters.derived
        aet = model.aetmodel
        pet = aet.petmodel
        nz = int(con.nmbzones.value)
        nc = int(con.sclass.value)
        zone_ptr[e + 1] = zone_ptr[e] + nz
        snow_ptr[e + 1] = snow_ptr[e] + nz * nc
        sfdist_ptr[e + 1] = sfdist_ptr[e] + nc
        mname = _modelname(model)
        model_id[e] = layout.MODELS[mname]
        ep = layout.EP
        epar[ep["z"], e] = der.z.value
        epar[ep["rellandarea"], e] = der.rellandarea.value
        epar[ep["rellowerzonearea"], e] = der.rellowerzonearea.value
        epar[ep["qfactor"], e] = der.qfactor.value
        epar[ep["dt"], e] = 1.0
        if mname == "hland_96":
            # ref: hydpy/models/hland_96.py (control RecStep, RespArea, PercMax, Alpha, K, K4, Gamma)
            recstep[e] = int(con.recstep.value)
            eflags[e] = layout.EF_RESPAREA if bool(con.resparea.value) else 0
            epar[ep["relsoilarea"], e] = der.relsoilarea.value
            epar[ep["relupperzonearea"], e] = der.relupperzonearea.value
            epar[ep["percmax"], e] = con.percmax.value
            epar[ep["dt"], e] = der.dt.value
            epar[ep["alpha"], e] = con.alpha.value
            epar[ep["k"], e] = con.k.value
            epar[ep["k4"], e] = con.k4.value
            epar[ep["gamma"], e] = con.gamma.value
        elif mname == "hland_96p":
            # ref: hydpy/models/hland_96p.py (control PercMax, K3; derived W3, K4, W4; fixed FSG)
            epar[ep["percmax"], e] = con.percmax.value
            epar[ep["k3"], e] = con.k3.value
            epar[ep["w3"], e] = der.w3.value
            epar[ep["k4"], e] = der.k4.value
            epar[ep["w4"], e] = der.w4.value
            epar[ep["fsg"], e] = model.parameters.fixed.fsg.value
        else:
            # ref: hydpy/models/hland_96c.py (control K4, Gamma)
            epar[ep["k4"], e] = con.k4.value
            epar[ep["gamma"], e] = con.gamma.value
        ret_io = _modelname(pet) == "evap_ret_io"
        epar[ep["petaltitude"], e] = 0.0 if ret_io else pet.parameters.derived.altitude.value
        zp = np.zeros((len(layout.ZPAR), nz))
        for name in _ZONE_CONTROL:
            zp[layout.ZP[name]] = _f(getattr(con, name), nz)
        for name in _ZONE_EXTRA[mname][0]:
            zp[layout.ZP[name]] = _f(getattr(con, name), nz)
        for name in _ZONE_EXTRA[mname][1]:
            zp[layout.ZP[name]] = _f(getattr(der, name), nz)
        zp[layout.ZP["ttm"]] = _f(der.ttm, nz)
        zp[layout.ZP["relzonearea"]] = _f(der.relzoneareas, nz)
        for name in _AET_CONTROL:
            zp[layout.ZP[name]] = _f(getattr(aet.parameters.control, name), nz)
        for name in _PET_CONTROL:
            if ret_io and name != "evapotranspirationfactor":
                zp[layout.ZP[name]] = 0.0        # (see _PET_MODELS: the identity setting of the HBV96 equations)
            else:
                zp[layout.ZP[name]] = _f(getattr(pet.parameters.control, name), nz)
        zp[layout.ZP["hruareafraction"]] = _f(pet.parameters.derived.hruareafraction, nz)
        zp_l.append(zp)
        zt_l.append(np.array(con.zonetype.value, dtype=np.int32).reshape(nz))
        acon = aet.parameters.control
        flags = (
            np.array(acon.water.value, dtype=bool).reshape(nz) * layout.ZF_WATER
            + np.array(acon.interception.value, dtype=bool).reshape(nz) * layout.ZF_INTERCEPTION
            + np.array(acon.soil.value, dtype=bool).reshape(nz) * layout.ZF_SOIL
            + np.array(acon.tree.value, dtype=bool).reshape(nz) * layout.ZF_TREE
            + np.array(der.sredend.value, dtype=bool).reshape(nz) * layout.ZF_SREDEND
        )
        zf_l.append(flags.astype(np.int32))
        zo_l.append(np.array(der.indiceszonez.value, dtype=np.int32).reshape(nz))
        sf_l.append(_f(con.sfdist, nc))
        nsr = int(der.srednumber.value)
        sred_ptr[e + 1] = sred_ptr[e] + nsr
        if nsr:
            order = np.array(der.sredorder.value, dtype=np.int64).reshape(-1, 2)[:nsr]
            ratios = np.array(der.zonearearatios.value, dtype=np.float64)
            sred = np.array(con.sred.value, dtype=np.float64)
            sr_f.append(order[:, 0].astype(np.int32))
            sr_t.append(order[:, 1].astype(np.int32))
            sr_a.append(ratios[order[:, 0], order[:, 1]] * sred[order[:, 0], order[:, 1]])
        rconc = getattr(model, "rconcmodel", None)
        if rconc is not None and _modelname(rconc) == "rconc_nash":
            # ref: hydpy/models/rconc_nash.py — the uh range holds the NmbStorages cells of states.sc
            rconc_kind[e] = layout.RCONC_NASH
            rcon, rder = rconc.parameters.control, rconc.parameters.derived
            nst = int(rcon.nmbstorages.value)
            nash_steps[e] = int(rcon.nmbsteps.value)
            epar[ep["nash_ksc"], e] = rder.ksc.value
            epar[ep["nash_dt"], e] = rder.dt.value
            sc = np.array(rconc.sequences.states.sc.value, dtype=np.float64).reshape(-1)
            if len(sc) != nst:
                raise ValueError(f"element `{element.name}`: `sc` must have nmbstorages entries")
            uh_l.append(np.zeros(nst))
            quh_l.append(sc)
            uh_ptr[e + 1] = uh_ptr[e] + nst
        elif rconc is not None:
            uh = np.array(rconc.parameters.control.uh.value, dtype=np.float64).reshape(-1)
            quh = np.array(rconc.sequences.logs.quh.value, dtype=np.float64).reshape(-1)
            if quh.shape != uh.shape:
                raise ValueError(f"element `{element.name}`: shapes of `uh` and `quh` differ")
            uh_l.append(uh)
            quh_l.append(quh)
            uh_ptr[e + 1] = uh_ptr[e] + len(uh)
        else:
            uh_ptr[e + 1] = uh_ptr[e]
        sta = model.sequences.states
        ic_l.append(_f(sta.ic, nz))
        sm_l.append(_f(sta.sm, nz))
        sp_l.append(np.array(sta.sp.value, dtype=np.float64).reshape(nc * nz))
        wc_l.append(np.array(sta.wc.value, dtype=np.float64).reshape(nc * nz))
        zeros = np.zeros(nz)
        if mname == "hland_96":
            uz[e] = sta.uz.value
            lz[e] = sta.lz.value
        elif mname == "hland_96p":
            sg2[e] = sta.sg2.value
            sg3[e] = sta.sg3.value
        else:
            lz[e] = sta.lz.value
        suz_l.append(_f(sta.suz, nz) if mname == "hland_96p" else zeros)
        sg1_l.append(_f(sta.sg1, nz) if mname == "hland_96p" else zeros)
        bw1_l.append(_f(sta.bw1, nz) if mname == "hland_96c" else zeros)
        bw2_l.append(_f(sta.bw2, nz) if mname == "hland_96c" else zeros)
        inp = model.sequences.inputs
        seq_nat, seq_net = _pet_input_sequences(pet)
        forcing[0, :, e] = _input_series(element, inp.p, i0, i1)
        forcing[1, :, e] = _input_series(element, inp.t, i0, i1)
        if seq_nat is not None:
            forcing[2, :, e] = _input_series(element, seq_nat, i0, i1)
        forcing[3, :, e] = _input_series(element, seq_net, i0, i1)
        if ret_io and np.any(forcing[3, :, e] < 0.0):
            raise UnsupportedModelError(
                f"element `{element.name}`: negative reference evapotranspiration in the input series of `evap_ret_io` "
                f"(the B200 engine serves it through the evap_pet_hbv96 equations, which clip at zero)")
        this_doy = np.array(der.doy.value, dtype=np.int64)[i0:i1]
        if doy is None:
            doy = this_doy
        elif not np.array_equal(doy, this_doy):
            raise ValueError("all hland_96 elements must share one `doy` index")
        out = model.sequences.outlets.q
        for node in getattr(out, "node2idx", {}):
            outlet_node[e] = node_index[node]

    def cat(parts: list[np.ndarray], dtype: Any) -> np.ndarray:
        return np.concatenate(parts).astype(dtype) if parts else np.zeros(0, dtype)

    land = LandBundle(
        zone_ptr=zone_ptr, snow_ptr=snow_ptr, sfdist_ptr=sfdist_ptr, uh_ptr=uh_ptr, sred_ptr=sred_ptr,
        recstep=recstep, eflags=eflags, forcing_col=np.arange(ne, dtype=np.int32), outlet_node=outlet_node,
        epar=epar, zonetype=cat(zt_l, np.int32), zflags=cat(zf_l, np.int32), zorder=cat(zo_l, np.int32),
        zpar=np.concatenate(zp_l, axis=1) if zp_l else np.zeros((len(layout.ZPAR), 0)),
        sfdist=cat(sf_l, np.float64), uh=cat(uh_l, np.float64), sred_from=cat(sr_f, np.int32),
        sred_to=cat(sr_t, np.int32), sred_adjust=cat(sr_a, np.float64),
        model=model_id, rconc=rconc_kind, nash_steps=nash_steps,
    )
    land.normalise()
    states = LandStates(
        ic=cat(ic_l, np.float64), sm=cat(sm_l, np.float64), sp=cat(sp_l, np.float64), wc=cat(wc_l, np.float64),
        uz=uz, lz=lz, quh=cat(quh_l, np.float64), suz=cat(suz_l, np.float64), sg1=cat(sg1_l, np.float64),
        bw1=cat(bw1_l, np.float64), bw2=cat(bw2_l, np.float64), sg2=sg2, sg3=sg3,
    )
    if doy is None:
        doy = np.zeros(nt, np.int64)
    return land, states, forcing, doy


def extract_river(elements: Iterable[Any], nodes: list[Any], node_index: dict[Any, int],
                  land_index: dict[Any, int], i0: int, i1: int
                  ) -> tuple[RiverBundle, np.ndarray, np.ndarray, np.ndarray, np.ndarray | None]:
    """Gather musk_classic / musk_mct elements + the node graph.  Returns (bundle, discharge, ext rows, node_writeback,
    musk_mct conditions [3, n_seg] or None)."""
    elements = list(elements)
    reach_index = {el: i for i, el in enumerate(elements)}
    nr, nn, nt = len(elements), len(nodes), i1 - i0
    inlet_ptr = np.zeros(nr + 1, np.int64)
    seg_ptr = np.zeros(nr + 1, np.int64)
    reach_outlet = np.full(nr, -1, np.int32)
    coef = np.zeros((3, nr))
    inlet_node: list[int] = []
    dis: list[np.ndarray] = []
    reach_model = np.zeros(nr, np.int32)
    nmbruns_a = np.ones(nr, np.int32)
    rpar = np.zeros((len(layout.RPAR), nr))
    trap_ptr = np.zeros(nr + 1, np.int64)
    tp_l: list[np.ndarray] = []
    mct_l: list[np.ndarray] = []
    for r, element in enumerate(elements):
        model = element.model
        con = model.parameters.control
        ns = int(con.nmbsegments.value)
        nmbruns = int(model.parameters.solver.nmbruns.value) if hasattr(model.parameters, "solver") else 1
        trap_ptr[r + 1] = trap_ptr[r]
        mct = np.zeros((3, ns + 1))
        if _modelname(model) == "musk_mct":
            # ref: hydpy/models/musk_mct.py; cross section: hydpy/models/wq_trapeze_strickler.py
            wq = getattr(model, "wqmodel", None)
            if wq is None or _modelname(wq) != "wq_trapeze_strickler":
                raise UnsupportedModelError(
                    f"element `{element.name}`: musk_mct needs a `wq_trapeze_strickler` cross-section submodel, found "
                    f"`{None if wq is None else _modelname(wq)}`")
            reach_model[r] = layout.REACH_MCT
            nmbruns_a[r] = nmbruns
            sol, der = model.parameters.solver, model.parameters.derived
            rp = layout.RP
            rpar[rp["tolerancewaterdepth"], r] = sol.tolerancewaterdepth.value
            rpar[rp["tolerancedischarge"], r] = sol.tolerancedischarge.value
            rpar[rp["tolerancenegativeinflow"], r] = sol.tolerancenegativeinflow.value
            rpar[rp["bottomslope"], r] = con.bottomslope.value
            rpar[rp["seconds"], r] = der.seconds.value
            rpar[rp["segmentlength"], r] = der.segmentlength.value
            wcon, wder = wq.parameters.control, wq.parameters.derived
            rpar[rp["wq_bottomslope"], r] = wcon.bottomslope.value
            ntr = int(wcon.nmbtrapezes.value)
            tp = np.zeros((len(layout.TPAR), ntr))
            for name, par in (("bottomwidth", wcon.bottomwidths), ("sideslope", wcon.sideslopes),
                              ("stricklercoefficient", wcon.stricklercoefficients),
                              ("calibrationfactor", wcon.calibrationfactors), ("bottomdepth", wder.bottomdepths),
                              ("trapezeheight", wder.trapezeheights), ("slopewidth", wder.slopewidths),
                              ("perimeterderivative", wder.perimeterderivatives)):
                tp[layout.TP[name]] = _f(par, ntr)
            tp_l.append(tp)
            trap_ptr[r + 1] = trap_ptr[r] + ntr
            sta = model.sequences.states
            mct[0, :ns] = _f(sta.courantnumber, ns) if ns else 0.0
            mct[1, :ns] = _f(sta.reynoldsnumber, ns) if ns else 0.0
            mct[2, :ns] = _f(model.sequences.factors.referencewaterdepth, ns) if ns else 0.0
        else:
            if nmbruns != 1:
                raise UnsupportedModelError(f"element `{element.name}`: musk_classic with nmbruns != 1 is not supported")
            coef[:, r] = np.array(con.coefficients.value, dtype=np.float64).reshape(3)
        mct_l.append(mct)
        d = np.array(model.sequences.states.discharge.value, dtype=np.float64).reshape(-1)
        if len(d) != ns + 1:
            raise ValueError(f"element `{element.name}`: `discharge` must have nmbsegments+1 entries")
        dis.append(d)
        seg_ptr[r + 1] = seg_ptr[r] + ns + 1
        node2idx = model.sequences.inlets.q.node2idx
        ordered = sorted(node2idx.items(), key=lambda kv: -1 if kv[1] is None else kv[1])
        for node, _ in ordered:
            inlet_node.append(node_index[node])
        inlet_ptr[r + 1] = inlet_ptr[r] + len(ordered)
        for node in model.sequences.outlets.q.node2idx:
            reach_outlet[r] = node_index[node]
    entry_ptr = np.zeros(nn + 1, np.int64)
    entry_src: list[int] = []
    node_mode = np.zeros(nn, np.int32)
    node_ext = np.full(nn, -1, np.int32)
    writeback = np.zeros(nn, bool)
    ext_rows: list[np.ndarray] = []
    for n, node in enumerate(nodes):
        for element in node.entries:
            if element in land_index:
                entry_src.append(land_index[element])
            elif element in reach_index:
                entry_src.append(-1 - reach_index[element])
            else:
                raise UnsupportedModelError(
                    f"node `{node.name}` is fed by element `{element.name}`, whose model is not covered by the "
                    f"B200 engine"
                )
        entry_ptr[n + 1] = len(entry_src)
        mode = str(node.deploymode)
        seqs = node.sequences
        if mode == "newsim":
            writeback[n] = True
        elif mode in ("oldsim", "oldsim_bi"):
            node_mode[n] = layout.NODE_EXT
            ext_rows.append(np.array(seqs.sim.series[i0:i1], dtype=np.float64))
        elif mode in ("obs", "obs_bi"):
            node_mode[n] = layout.NODE_EXT
            ext_rows.append(np.array(seqs.obs.series[i0:i1], dtype=np.float64))
            writeback[n] = True
        elif mode == "obs_newsim":
            node_mode[n] = layout.NODE_OBSNEWSIM
            ext_rows.append(np.array(seqs.obs.series[i0:i1], dtype=np.float64))
            writeback[n] = True
        elif mode == "obs_oldsim":
            node_mode[n] = layout.NODE_EXT
            obs = np.array(seqs.obs.series[i0:i1], dtype=np.float64)
            old = np.array(seqs.sim.series[i0:i1], dtype=np.float64)
            ext_rows.append(np.where(np.isnan(obs), old, obs))
        else:
            raise UnsupportedModelError(f"node `{node.name}`: deploy mode `{mode}` is not supported")
        if node_mode[n] != layout.NODE_NEWSIM:
            node_ext[n] = len(ext_rows) - 1
    river = RiverBundle(
        entry_ptr=entry_ptr, entry_src=np.array(entry_src, np.int32), node_mode=node_mode, node_ext=node_ext,
        inlet_ptr=inlet_ptr, inlet_node=np.array(inlet_node, np.int32), reach_outlet=reach_outlet, seg_ptr=seg_ptr,
        coef=coef, n_ext=len(ext_rows), reach_model=reach_model, nmbruns=nmbruns_a, rpar=rpar, trap_ptr=trap_ptr,
        tpar=np.concatenate(tp_l, axis=1) if tp_l else np.zeros((len(layout.TPAR), 0)),
    )
    river.normalise()
    discharge = np.concatenate(dis) if dis else np.zeros(0)
    ext = np.stack(ext_rows) if ext_rows else np.zeros((0, nt))
    mct_states = np.concatenate(mct_l, axis=1) if (mct_l and river.any_mct) else None
    return river, discharge, ext, writeback, mct_states


def extract(hp: Any, i0: int, i1: int) -> Problem:
    """Translate the complete `HydPy` project `hp` for the simulation window `[i0, i1)`.

    Every element must handle `hland_96` or `musk_classic`; anything else raises `UnsupportedModelError`
    (explicit error instead of a silent CPU fallback).
    """
    nodes = list(hp.nodes)
    node_index = {node: i for i, node in enumerate(nodes)}
    land_elements, river_elements = [], []
    for element in hp.elements:
        name = _modelname(element.model)
        if name in LAND_MODELS:
            land_elements.append(element)
        elif name in RIVER_MODELS:
            river_elements.append(element)
        else:
            raise UnsupportedModelError(
                f"element `{element.name}` handles model `{name}`; the B200 engine covers "
                f"{', '.join(LAND_MODELS + RIVER_MODELS)} only"
            )
        if len(tuple(getattr(element, "receivers", ()))) or len(tuple(getattr(element, "senders", ()))) or len(
            tuple(getattr(element, "observers", ()))
        ):
            raise UnsupportedModelError(f"element `{element.name}`: receiver/sender/observer nodes are not supported")
    land, states, forcing, doy = extract_land(land_elements, node_index, i0, i1)
    land_index = {el: i for i, el in enumerate(land_elements)}
    river, discharge, ext, writeback, mct_states = extract_river(river_elements, nodes, node_index, land_index, i0, i1)
    return Problem(
        land=land, states=states, river=river, discharge=discharge, forcing=forcing, doy=doy, ext=ext,
        land_elements=land_elements, river_elements=river_elements, nodes=nodes, node_writeback=writeback,
        i0=i0, i1=i1, mct_states=mct_states,
    )


# ---------------------------------------------------------------------------------------------------------------------
# results → HydPy objects   (contract: SURVEY.md §8b seam #1)
# ---------------------------------------------------------------------------------------------------------------------


def _submodel(model: Any, path: str) -> Any:
    sub = model
    for part in path.split("."):
        sub = getattr(sub, part, None)
        if sub is None:
            return None
    return sub


def _io_sequences(model: Any):
    """(group name, sequence) of every IO sequence of a model (no submodels)."""
    for group in ("inputs", "factors", "fluxes", "states", "inlets", "outlets", "receivers", "senders", "observers"):
        for seq in getattr(model.sequences, group, None) or ():
            yield group, seq


_MEAN_SOURCE = {"meanpet": "potentialevapotranspiration", "meanret": "referenceevapotranspiration"}

#: the models below a land element: path from the main model ("" = the main model itself)
_LAND_PATHS = ("", "aetmodel", "aetmodel.petmodel", "rconcmodel")


def check_deliverable(problem: Problem) -> None:
    """Everything the reference's `save_data` / `load_data` would serve during the run must be something the engine can
    serve (ref: hydpy/cythons/modelutils.py:1080-1167): refuse just-in-time NetCDF reading and writing (`diskflag`) and
    series kept in RAM (`ramflag`, set by `prepare_allseries()` down to every submodel) that the engine does not compute —
    instead of leaving them stale (no silent fallback)."""
    known = {(path, group, name) for (path, group, name) in layout.SUBMODEL_SERIES} | set(layout.SUBMODEL_DERIVED)
    for element in problem.land_elements:
        for path in _LAND_PATHS:
            sub = _submodel(element.model, path) if path else element.model
            if sub is None:
                continue
            for group, seq in _io_sequences(sub):
                if getattr(seq, "diskflag_reading", False) or getattr(seq, "diskflag_writing", False):
                    raise UnsupportedModelError(
                        f"element `{element.name}`: sequence `{seq.name}` of `{_modelname(sub)}` reads or writes its "
                        f"time series just in time (NetCDF); the B200 engine only serves series kept in RAM")
                if not getattr(seq, "ramflag", False) or group in ("inputs", "outlets", "inlets"):
                    continue
                if path == "":
                    if seq.name in layout.SERIES:
                        continue
                elif (path, group, seq.name) in known:
                    continue
                raise UnsupportedModelError(
                    f"element `{element.name}`: the series of `{group}.{seq.name}` of `{_modelname(sub)}` is requested "
                    f"(ramflag), but the B200 engine does not deliver it")
    for element in problem.river_elements:
        for group, seq in _io_sequences(element.model):
            if getattr(seq, "diskflag_reading", False) or getattr(seq, "diskflag_writing", False):
                raise UnsupportedModelError(
                    f"element `{element.name}`: sequence `{seq.name}` reads or writes its time series just in time "
                    f"(NetCDF); the B200 engine only serves series kept in RAM")
    for node in problem.nodes:
        for name in ("sim", "obs"):
            seq = getattr(node.sequences, name, None)
            if seq is not None and (getattr(seq, "diskflag_reading", False) or getattr(seq, "diskflag_writing", False)):
                raise UnsupportedModelError(
                    f"node `{node.name}`: sequence `{name}` reads or writes its time series just in time (NetCDF); the "
                    f"B200 engine only serves series kept in RAM")


def requested_series(problem: Problem) -> list[str]:
    """Names of the land series at least one element keeps in RAM (`ramflag`) — its own sequences and those of its
    evap / rconc submodels (`layout.SUBMODEL_SERIES`: each is a copy of an engine series)."""
    check_deliverable(problem)
    wanted: list[str] = []
    for name in layout.SERIES:
        group = layout.SERIES_GROUP[name]
        for element in problem.land_elements:
            seq = getattr(getattr(element.model.sequences, group), name, None)   # sibling models own different sequences
            if seq is not None and seq.ramflag:
                wanted.append(name)
                break
    extra = dict(layout.SUBMODEL_SERIES)
    extra[("aetmodel.petmodel", "fluxes", "meanpotentialevapotranspiration")] = "potentialevapotranspiration"
    extra[("aetmodel.petmodel", "fluxes", "meanreferenceevapotranspiration")] = "referenceevapotranspiration"
    for (path, group, seqname), name in extra.items():
        if name in wanted:
            continue
        for element in problem.land_elements:
            sub = _submodel(element.model, path)
            seq = getattr(getattr(sub.sequences, group, None), seqname, None) if sub is not None else None
            if seq is not None and getattr(seq, "ramflag", False):
                wanted.append(name)
                break
    return [n for n in layout.SERIES if n in wanted]


def check_reach_series(problem: Problem) -> None:
    """musk_mct: inflow, outflow, states.discharge and the 1-D sequences of `layout.MCT_SERIES` are recorded per step —
    that is every sequence the model owns; a sequence this engine does not know (a future reference version) is an
    error, not a silent gap."""
    for element in problem.river_elements:
        model = element.model
        if _modelname(model) != "musk_mct":
            continue
        for group in ("factors", "fluxes", "states"):
            for seq in getattr(model.sequences, group, ()):
                if getattr(seq, "ramflag", False) and seq.name not in ("inflow", "outflow", "discharge") and (
                        layout.MCT_SERIES_GROUP.get(seq.name, "factors") != group or seq.name not in layout.MS):
                    raise UnsupportedModelError(
                        f"element `{element.name}`: the B200 engine does not record the time series of musk_mct "
                        f"sequence `{seq.name}`")


def requested_mct_series(problem: Problem) -> list[str]:
    """The 1-D musk_mct sequences at least one element keeps in RAM (`seq.ramflag`), in `layout.MCT_SERIES` order."""
    wanted = []
    for name in layout.MCT_SERIES:
        group = layout.MCT_SERIES_GROUP.get(name, "factors")
        for element in problem.river_elements:
            if _modelname(element.model) != "musk_mct":
                continue
            seq = getattr(getattr(element.model.sequences, group), name, None)
            if seq is not None and getattr(seq, "ramflag", False):
                wanted.append(name)
                break
    return wanted


def requested_reach_states(problem: Problem) -> bool:
    """True when at least one musk_classic element keeps `states.discharge` in RAM."""
    return any(getattr(el.model.sequences.states.discharge, "ramflag", False) for el in problem.river_elements)


def _set_state(seq: Any, value: Any) -> None:
    seq.new = value
    seq.old = value


def write_back(problem: Problem, *, states: LandStates, discharge: np.ndarray, node_rows: np.ndarray,
               land_rows: np.ndarray, reach_out: np.ndarray | None, reach_in: np.ndarray | None,
               series: dict[str, np.ndarray], reach_state_series: np.ndarray | None = None,
               mct_states: np.ndarray | None = None, mct_series: dict[str, np.ndarray] | None = None) -> None:
    """Store the results of one run in the HydPy objects exactly where the reference leaves them."""
    land, i0, i1 = problem.land, problem.i0, problem.i1
    last = i1 - i0 - 1
    for e, element in enumerate(problem.land_elements):
        model = element.model
        z0, z1 = int(land.zone_ptr[e]), int(land.zone_ptr[e + 1])
        s0, s1 = int(land.snow_ptr[e]), int(land.snow_ptr[e + 1])
        nz = z1 - z0
        nc = (s1 - s0) // max(nz, 1)
        sta = model.sequences.states
        _set_state(sta.ic, states.ic[z0:z1])
        _set_state(sta.sm, states.sm[z0:z1])
        _set_state(sta.sp, states.sp[s0:s1].reshape(nc, nz))
        _set_state(sta.wc, states.wc[s0:s1].reshape(nc, nz))
        mid = int(land.model[e])
        if mid == layout.MODEL_HLAND_96:
            _set_state(sta.uz, float(states.uz[e]))
            _set_state(sta.lz, float(states.lz[e]))
        elif mid == layout.MODEL_HLAND_96P:
            _set_state(sta.suz, states.suz[z0:z1])
            _set_state(sta.sg1, states.sg1[z0:z1])
            _set_state(sta.sg2, float(states.sg2[e]))
            _set_state(sta.sg3, float(states.sg3[e]))
        else:
            _set_state(sta.bw1, states.bw1[z0:z1])
            _set_state(sta.bw2, states.bw2[z0:z1])
            _set_state(sta.lz, float(states.lz[e]))
        # input sequences fed by nodes keep what they received (save_data, ref: modelutils.py:1127-1167)
        col = int(land.forcing_col[e])
        for k, seq in enumerate((model.sequences.inputs.p, model.sequences.inputs.t,
                                 *_pet_input_sequences(model.aetmodel.petmodel))):
            if seq is not None and getattr(seq, "node2idx", None) and seq.ramflag:
                seq.series[i0:i1] = problem.forcing[k, :, col]
        u0, u1 = int(land.uh_ptr[e]), int(land.uh_ptr[e + 1])
        if u1 > u0:
            if int(land.rconc[e]) == layout.RCONC_NASH:
                _set_state(model.rconcmodel.sequences.states.sc, states.quh[u0:u1])
            else:
                model.rconcmodel.sequences.logs.quh.value = states.quh[u0:u1]
        out = model.sequences.outlets.q
        if getattr(out, "ramflag", False):
            out.series[i0:i1] = land_rows[e]
        if last >= 0:
            out.value = float(land_rows[e, last])
        for name, data in series.items():
            group = layout.SERIES_GROUP[name]
            seq = getattr(getattr(model.sequences, group), name, None)
            if seq is None or not seq.ramflag:
                continue
            if name in layout.SERIES_ZONE:
                seq.series[i0:i1] = data[:, z0:z1]
                if last >= 0 and group != "states":
                    seq.value = data[last, z0:z1]
            elif name in layout.SERIES_SNOW:
                seq.series[i0:i1] = data[:, s0:s1].reshape(-1, nc, nz)
                if last >= 0 and group != "states":
                    seq.value = data[last, s0:s1].reshape(nc, nz)
            else:
                seq.series[i0:i1] = data[:, e]
                if last >= 0 and group != "states":
                    seq.value = float(data[last, e])
        # sequences of the submodels: copies of engine series (`layout.SUBMODEL_SERIES`) and three derived ones
        for (path, group, seqname), name in layout.SUBMODEL_SERIES.items():
            sub = _submodel(model, path)
            seq = getattr(getattr(sub.sequences, group, None), seqname, None) if sub is not None else None
            if seq is None or name not in series:
                continue
            data = series[name]
            if name in layout.SERIES_UH:       # rconc_nash states.sc: the state itself is set from `states.quh` above
                if getattr(seq, "ramflag", False) and int(land.rconc[e]) == layout.RCONC_NASH:
                    seq.series[i0:i1] = data[:, u0:u1]
            elif name in layout.SERIES_ZONE:
                if getattr(seq, "ramflag", False):
                    seq.series[i0:i1] = data[:, z0:z1]
                if last >= 0:
                    seq.value = data[last, z0:z1]
            else:
                if getattr(seq, "ramflag", False):
                    seq.series[i0:i1] = data[:, e]
                if last >= 0:
                    seq.value = float(data[last, e])
        for (path, group, seqname), kind in layout.SUBMODEL_DERIVED.items():
            sub = _submodel(model, path)
            seq = getattr(getattr(sub.sequences, group, None), seqname, None) if sub is not None else None
            if seq is None or not getattr(seq, "ramflag", False):
                continue
            if kind == "zero":
                seq.series[i0:i1] = 0.0
            elif kind == "input_t":
                seq.series[i0:i1] = problem.forcing[1, :, col]
            elif kind in ("meanpet", "meanret") and _MEAN_SOURCE[kind] in series:
                # ref: evap_model.py:5489-5498 — Σ hruareafraction[k] * potentialevapotranspiration[k], k = 0 … Z-1;
                # evap_model.py:5451-5461 (evap_ret_io) — the same sum over referenceevapotranspiration
                frac = land.zp("hruareafraction")[z0:z1]
                pet = series[_MEAN_SOURCE[kind]][:, z0:z1]
                acc = np.zeros(pet.shape[0])
                for k in range(nz):
                    acc = acc + frac[k] * pet[:, k]
                seq.series[i0:i1] = acc
    river = problem.river
    for r, element in enumerate(problem.river_elements):
        model = element.model
        g0, g1 = int(river.seg_ptr[r]), int(river.seg_ptr[r + 1])
        _set_state(model.sequences.states.discharge, discharge[g0:g1])
        if mct_states is not None and int(river.reach_model[r]) == layout.REACH_MCT:
            _set_state(model.sequences.states.courantnumber, mct_states[0, g0:g1 - 1])
            _set_state(model.sequences.states.reynoldsnumber, mct_states[1, g0:g1 - 1])
            model.sequences.factors.referencewaterdepth.value = mct_states[2, g0:g1 - 1]
        if mct_series and int(river.reach_model[r]) == layout.REACH_MCT:
            for name, data in mct_series.items():
                seq = getattr(getattr(model.sequences, layout.MCT_SERIES_GROUP.get(name, "factors")), name)
                if getattr(seq, "ramflag", False):
                    seq.series[i0:i1] = data[:, g0:g1 - 1]
                if last >= 0 and name not in ("courantnumber", "reynoldsnumber", "referencewaterdepth"):
                    seq.value = data[last, g0:g1 - 1]
        if reach_state_series is not None and getattr(model.sequences.states.discharge, "ramflag", False):
            model.sequences.states.discharge.series[i0:i1] = reach_state_series[:, g0:g1]
        flu = model.sequences.fluxes
        if reach_in is not None:
            if flu.inflow.ramflag:
                flu.inflow.series[i0:i1] = reach_in[r]
            if last >= 0:
                flu.inflow.value = float(reach_in[r, last])
        if reach_out is not None:
            if flu.outflow.ramflag:
                flu.outflow.series[i0:i1] = reach_out[r]
            out = model.sequences.outlets.q
            if getattr(out, "ramflag", False):
                out.series[i0:i1] = reach_out[r]
            if last >= 0:
                flu.outflow.value = float(reach_out[r, last])
                out.value = float(reach_out[r, last])
    for n, node in enumerate(problem.nodes):
        if problem.node_writeback is not None and not problem.node_writeback[n]:
            continue
        sim = node.sequences.sim
        if sim.ramflag:
            sim.series[i0:i1] = node_rows[n]
        if last >= 0:
            # ref: hydpy/core/hydpytools.py:3094-3100 — sim.value = sim.series[last index]
            sim.value = float(node_rows[n, last])
