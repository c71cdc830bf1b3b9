"""ctypes mirror of the structs in `include/hydpy_b200.h` and the SoA "bundles" that fill them.

A bundle is the engine-side image of what the reference keeps per element in
`model.parameters.<group>.fastaccess` / `model.sequences.<group>.fastaccess`: ragged
per-element arrays are concatenated and CSR-indexed, so 10^5 elements are a handful of flat NumPy arrays.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Any

import numpy as np

from . import layout

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)


class hb_land_desc(C.Structure):
    _fields_ = [
        ("n_elem", C.c_int64), ("n_zone", C.c_int64), ("n_snow", C.c_int64), ("n_sfdist", C.c_int64),
        ("n_uh", C.c_int64), ("n_sred", C.c_int64),
        ("zone_ptr", c_i64p), ("snow_ptr", c_i64p), ("sfdist_ptr", c_i64p), ("uh_ptr", c_i64p), ("sred_ptr", c_i64p),
        ("recstep", c_i32p), ("eflags", c_i32p), ("forcing_col", c_i32p), ("outlet_node", c_i32p),
        ("epar", c_f64p),
        ("zonetype", c_i32p), ("zflags", c_i32p), ("zorder", c_i32p),
        ("zpar", c_f64p), ("sfdist", c_f64p), ("uh", c_f64p),
        ("sred_from", c_i32p), ("sred_to", c_i32p), ("sred_adjust", c_f64p),
        ("model", c_i32p), ("rconc", c_i32p), ("nash_steps", c_i32p),
    ]


class hb_land_states(C.Structure):
    _fields_ = [(n, c_f64p) for n in ("ic", "sm", "sp", "wc", "uz", "lz", "quh", "suz", "sg1", "bw1", "bw2", "sg2", "sg3")]


class hb_river_desc(C.Structure):
    _fields_ = [
        ("n_node", C.c_int64), ("n_reach", C.c_int64), ("n_entry", C.c_int64), ("n_inlet", C.c_int64),
        ("n_seg", C.c_int64), ("n_ext", C.c_int64),
        ("entry_ptr", c_i64p), ("entry_src", c_i32p), ("node_mode", c_i32p), ("node_ext", c_i32p),
        ("inlet_ptr", c_i64p), ("inlet_node", c_i32p), ("reach_outlet", c_i32p), ("seg_ptr", c_i64p),
        ("coef", c_f64p),
        ("reach_model", c_i32p), ("nmbruns", c_i32p), ("rpar", c_f64p), ("n_trap", C.c_int64), ("trap_ptr", c_i64p),
        ("tpar", c_f64p),
    ]


class hb_river_states(C.Structure):
    _fields_ = [(n, c_f64p) for n in ("discharge", "courantnumber", "reynoldsnumber", "referencewaterdepth")]


class hb_timing(C.Structure):
    _fields_ = [
        ("land_ms", C.c_double), ("river_ms", C.c_double), ("total_ms", C.c_double),
        ("land_launches", C.c_int64), ("river_launches", C.c_int64), ("n_levels", C.c_int64),
        ("zone_ms", C.c_double), ("element_ms", C.c_double), ("general_ms", C.c_double), ("n_fast", C.c_int64),
        ("n_runs", C.c_int64),
        ("zone_warpsteps", C.c_int64), ("zone_warpsteps_dry", C.c_int64), ("zone_steps", C.c_int64),
        ("zone_steps_wet", C.c_int64), ("elem_steps", C.c_int64), ("elem_steps_run", C.c_int64),
        ("substeps_pow", C.c_int64), ("elem_warpsteps", C.c_int64), ("elem_warpsteps_run", C.c_int64),
        ("n_fused", C.c_int64), ("n_spec", C.c_int64), ("spec_reruns", C.c_int64),
    ]


def _arr(a: Any, dtype: Any, shape: tuple[int, ...] | None = None) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None and out.shape != shape:
        raise ValueError(f"array of shape {out.shape} given where {shape} is required")
    return out


def ptr(a: np.ndarray, ctype: Any) -> Any:
    """Pointer to the data of a C-contiguous array (a 1-element dummy keeps empty arrays addressable)."""
    if a.size == 0:
        a = np.zeros(1, dtype=a.dtype)
    return a.ctypes.data_as(C.POINTER(ctype))


@dataclasses.dataclass
class LandStates:
    """hland_96 states + rconc_uh log (`Sequences.conditions`, ref: core/sequencetools.py:807-828); `quh` also holds
    `states.sc` of rconc_nash elements.  The states of the sibling models (hland_96p: suz, sg1 per zone, sg2, sg3 per
    element; hland_96c: bw1, bw2 per zone) default to zeros of the matching size."""

    ic: np.ndarray
    sm: np.ndarray
    sp: np.ndarray
    wc: np.ndarray
    uz: np.ndarray
    lz: np.ndarray
    quh: np.ndarray
    suz: np.ndarray | None = None
    sg1: np.ndarray | None = None
    bw1: np.ndarray | None = None
    bw2: np.ndarray | None = None
    sg2: np.ndarray | None = None
    sg3: np.ndarray | None = None

    NAMES_V1 = ("ic", "sm", "sp", "wc", "uz", "lz", "quh")
    NAMES_ZONE2 = ("suz", "sg1", "bw1", "bw2")
    NAMES_ELEM2 = ("sg2", "sg3")
    NAMES = NAMES_V1 + NAMES_ZONE2 + NAMES_ELEM2

    def copy(self) -> "LandStates":
        self.normalise()
        return LandStates(**{n: getattr(self, n).copy() for n in self.NAMES})

    def normalise(self) -> None:
        for n in self.NAMES_ZONE2:
            if getattr(self, n) is None:
                setattr(self, n, np.zeros(len(self.ic)))
        for n in self.NAMES_ELEM2:
            if getattr(self, n) is None:
                setattr(self, n, np.zeros(len(self.uz)))
        for n in self.NAMES:
            setattr(self, n, _arr(getattr(self, n), np.float64))

    def as_struct(self) -> hb_land_states:
        self.normalise()
        s = hb_land_states()
        for n in self.NAMES:
            setattr(s, n, ptr(getattr(self, n), C.c_double))
        s._keep = self  # type: ignore[attr-defined]
        return s


@dataclasses.dataclass
class LandBundle:
    """All hland_96 elements of one run as structure-of-arrays (fields = `hb_land_desc`)."""

    zone_ptr: np.ndarray
    snow_ptr: np.ndarray
    sfdist_ptr: np.ndarray
    uh_ptr: np.ndarray
    sred_ptr: np.ndarray
    recstep: np.ndarray
    eflags: np.ndarray
    forcing_col: np.ndarray
    outlet_node: np.ndarray
    epar: np.ndarray          # [len(EPAR), n_elem]
    zonetype: np.ndarray
    zflags: np.ndarray
    zorder: np.ndarray
    zpar: np.ndarray          # [len(ZPAR), n_zone]
    sfdist: np.ndarray
    uh: np.ndarray
    sred_from: np.ndarray
    sred_to: np.ndarray
    sred_adjust: np.ndarray
    model: np.ndarray | None = None        # [n_elem] layout.MODEL_* (None: all hland_96)
    rconc: np.ndarray | None = None        # [n_elem] layout.RCONC_* (None: all rconc_uh / no submodel)
    nash_steps: np.ndarray | None = None   # [n_elem] rconc_nash NmbSteps

    I64 = ("zone_ptr", "snow_ptr", "sfdist_ptr", "uh_ptr", "sred_ptr")
    I32 = ("recstep", "eflags", "forcing_col", "outlet_node", "zonetype", "zflags", "zorder", "sred_from", "sred_to",
           "model", "rconc", "nash_steps")
    F64 = ("epar", "zpar", "sfdist", "uh", "sred_adjust")

    @property
    def n_elem(self) -> int:
        return len(self.zone_ptr) - 1

    @property
    def n_zone(self) -> int:
        return int(self.zone_ptr[-1])

    @property
    def n_snow(self) -> int:
        return int(self.snow_ptr[-1])

    @property
    def n_uh(self) -> int:
        return int(self.uh_ptr[-1])

    def normalise(self) -> None:
        for n in ("model", "rconc", "nash_steps"):
            if getattr(self, n) is None:
                setattr(self, n, np.zeros(self.n_elem, np.int32))
        for n in self.I64:
            setattr(self, n, _arr(getattr(self, n), np.int64))
        for n in self.I32:
            setattr(self, n, _arr(getattr(self, n), np.int32))
        for n in self.F64:
            setattr(self, n, _arr(getattr(self, n), np.float64))
        ne, nz = self.n_elem, self.n_zone
        # bundles of ABI 1 (committed fixtures, hland_96 only): the rows of the sibling models are zero
        if self.zpar.shape == (layout.N_ZPAR_V1, nz):
            self.zpar = np.concatenate([self.zpar, np.zeros((len(layout.ZPAR) - layout.N_ZPAR_V1, nz))])
        if self.epar.shape == (layout.N_EPAR_V1, ne):
            self.epar = np.concatenate([self.epar, np.zeros((len(layout.EPAR) - layout.N_EPAR_V1, ne))])
        if self.epar.shape != (len(layout.EPAR), ne):
            raise ValueError(f"epar must have shape {(len(layout.EPAR), ne)}, not {self.epar.shape}")
        if self.zpar.shape != (len(layout.ZPAR), nz):
            raise ValueError(f"zpar must have shape {(len(layout.ZPAR), nz)}, not {self.zpar.shape}")
        for n in ("snow_ptr", "sfdist_ptr", "uh_ptr", "sred_ptr"):
            if len(getattr(self, n)) != ne + 1:
                raise ValueError(f"{n} must have n_elem+1 entries")
        for n in ("recstep", "eflags", "forcing_col", "outlet_node", "model", "rconc", "nash_steps"):
            if len(getattr(self, n)) != ne:
                raise ValueError(f"{n} must have n_elem entries")
        for n in ("zonetype", "zflags", "zorder"):
            if len(getattr(self, n)) != nz:
                raise ValueError(f"{n} must have n_zone entries")
        if len(self.sfdist) != self.sfdist_ptr[-1] or len(self.uh) != self.uh_ptr[-1]:
            raise ValueError("sfdist/uh do not match their CSR pointers")
        if not (len(self.sred_from) == len(self.sred_to) == len(self.sred_adjust) == self.sred_ptr[-1]):
            raise ValueError("snow redistribution arrays do not match sred_ptr")
        sclass = np.diff(self.sfdist_ptr)
        if np.any(np.diff(self.snow_ptr) != sclass * np.diff(self.zone_ptr)):
            raise ValueError("snow_ptr must advance by SClass*NmbZones per element")

    def as_struct(self) -> hb_land_desc:
        self.normalise()
        d = hb_land_desc()
        d.n_elem, d.n_zone, d.n_snow = self.n_elem, self.n_zone, self.n_snow
        d.n_sfdist, d.n_uh, d.n_sred = len(self.sfdist), len(self.uh), len(self.sred_from)
        for n in self.I64:
            setattr(d, n, ptr(getattr(self, n), C.c_int64))
        for n in self.I32:
            setattr(d, n, ptr(getattr(self, n), C.c_int32))
        for n in self.F64:
            setattr(d, n, ptr(getattr(self, n), C.c_double))
        d._keep = self  # type: ignore[attr-defined]
        return d

    def zp(self, name: str) -> np.ndarray:
        return self.zpar[layout.ZP[name]]

    def ep(self, name: str) -> np.ndarray:
        return self.epar[layout.EP[name]]

    def to_dict(self, prefix: str = "land_") -> dict[str, np.ndarray]:
        self.normalise()
        return {prefix + f.name: np.asarray(getattr(self, f.name)) for f in dataclasses.fields(self)}

    @classmethod
    def from_dict(cls, d: Any, prefix: str = "land_") -> "LandBundle":
        return cls(**{f.name: np.asarray(d[prefix + f.name]) for f in dataclasses.fields(cls) if prefix + f.name in d})


@dataclasses.dataclass
class RiverBundle:
    """All musk_classic elements and the node graph (fields = `hb_river_desc`)."""

    entry_ptr: np.ndarray
    entry_src: np.ndarray
    node_mode: np.ndarray
    node_ext: np.ndarray
    inlet_ptr: np.ndarray
    inlet_node: np.ndarray
    reach_outlet: np.ndarray
    seg_ptr: np.ndarray
    coef: np.ndarray          # [3, n_reach]
    n_ext: int = 0
    # musk_mct reaches (None: all musk_classic)
    reach_model: np.ndarray | None = None   # [n_reach] layout.REACH_*
    nmbruns: np.ndarray | None = None       # [n_reach]
    rpar: np.ndarray | None = None          # [len(layout.RPAR), n_reach]
    trap_ptr: np.ndarray | None = None      # [n_reach+1]
    tpar: np.ndarray | None = None          # [len(layout.TPAR), n_trap]

    I64 = ("entry_ptr", "inlet_ptr", "seg_ptr", "trap_ptr")
    I32 = ("entry_src", "node_mode", "node_ext", "inlet_node", "reach_outlet", "reach_model", "nmbruns")

    @property
    def n_node(self) -> int:
        return len(self.entry_ptr) - 1

    @property
    def n_reach(self) -> int:
        return len(self.inlet_ptr) - 1

    @property
    def n_seg(self) -> int:
        return int(self.seg_ptr[-1])

    @property
    def any_mct(self) -> bool:
        return self.reach_model is not None and bool(np.any(np.asarray(self.reach_model) == layout.REACH_MCT))

    def normalise(self) -> None:
        nr = self.n_reach
        if self.reach_model is None:
            self.reach_model = np.zeros(nr, np.int32)
        if self.nmbruns is None:
            self.nmbruns = np.ones(nr, np.int32)
        if self.rpar is None:
            self.rpar = np.zeros((len(layout.RPAR), nr))
        if self.trap_ptr is None:
            self.trap_ptr = np.zeros(nr + 1, np.int64)
        if self.tpar is None:
            self.tpar = np.zeros((len(layout.TPAR), 0))
        for n in self.I64:
            setattr(self, n, _arr(getattr(self, n), np.int64))
        for n in self.I32:
            setattr(self, n, _arr(getattr(self, n), np.int32))
        self.coef = _arr(self.coef, np.float64)
        self.rpar = _arr(self.rpar, np.float64)
        self.tpar = _arr(self.tpar, np.float64)
        if self.rpar.shape != (len(layout.RPAR), nr) or len(self.reach_model) != nr or len(self.nmbruns) != nr:
            raise ValueError("reach_model/nmbruns/rpar do not match n_reach")
        if len(self.trap_ptr) != nr + 1 or self.tpar.shape != (len(layout.TPAR), int(self.trap_ptr[-1])):
            raise ValueError("trap_ptr/tpar do not match")
        if self.coef.shape != (3, self.n_reach):
            raise ValueError(f"coef must have shape (3, {self.n_reach}), not {self.coef.shape}")
        if len(self.node_mode) != self.n_node or len(self.node_ext) != self.n_node:
            raise ValueError("node_mode/node_ext must have n_node entries")
        if len(self.reach_outlet) != self.n_reach or len(self.seg_ptr) != self.n_reach + 1:
            raise ValueError("reach_outlet/seg_ptr do not match n_reach")
        if len(self.entry_src) != self.entry_ptr[-1] or len(self.inlet_node) != self.inlet_ptr[-1]:
            raise ValueError("entry/inlet arrays do not match their CSR pointers")
        self.n_ext = int(self.n_ext)

    def as_struct(self) -> hb_river_desc:
        self.normalise()
        d = hb_river_desc()
        d.n_node, d.n_reach = self.n_node, self.n_reach
        d.n_entry, d.n_inlet, d.n_seg, d.n_ext = len(self.entry_src), len(self.inlet_node), self.n_seg, self.n_ext
        for n in self.I64:
            setattr(d, n, ptr(getattr(self, n), C.c_int64))
        for n in self.I32:
            setattr(d, n, ptr(getattr(self, n), C.c_int32))
        d.coef = ptr(self.coef, C.c_double)
        d.rpar = ptr(self.rpar, C.c_double)
        d.tpar = ptr(self.tpar, C.c_double)
        d.n_trap = int(self.trap_ptr[-1])
        d._keep = self  # type: ignore[attr-defined]
        return d

    def to_dict(self, prefix: str = "river_") -> dict[str, np.ndarray]:
        self.normalise()
        return {prefix + f.name: np.asarray(getattr(self, f.name)) for f in dataclasses.fields(self)}

    @classmethod
    def from_dict(cls, d: Any, prefix: str = "river_") -> "RiverBundle":
        kw = {f.name: np.asarray(d[prefix + f.name]) for f in dataclasses.fields(cls) if prefix + f.name in d}
        kw["n_ext"] = int(kw["n_ext"])
        return cls(**kw)

    @classmethod
    def empty(cls, n_node: int = 0) -> "RiverBundle":
        z64 = np.zeros(1, np.int64)
        return cls(
            entry_ptr=np.zeros(n_node + 1, np.int64), entry_src=np.zeros(0, np.int32),
            node_mode=np.zeros(n_node, np.int32), node_ext=np.full(n_node, -1, np.int32),
            inlet_ptr=z64.copy(), inlet_node=np.zeros(0, np.int32), reach_outlet=np.zeros(0, np.int32),
            seg_ptr=z64.copy(), coef=np.zeros((3, 0)), n_ext=0,
        )


def states_to_dict(st: LandStates, prefix: str = "state_") -> dict[str, np.ndarray]:
    st.normalise()
    return {prefix + n: np.asarray(getattr(st, n)) for n in LandStates.NAMES}


def states_from_dict(d: Any, prefix: str = "state_") -> LandStates:
    return LandStates(**{n: np.array(d[prefix + n], dtype=np.float64) for n in LandStates.NAMES if prefix + n in d})
