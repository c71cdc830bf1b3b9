"""ctypes wrapper of `oracle/liboracle.so` (the C restatement in `hland_oracle.c`).  TEST INFRASTRUCTURE ONLY.

Mirrors the call sequence of the engine (`hydpy_b200.engine.Engine`) closely enough that tests can run the same
problem through both and compare.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

from hydpy_b200 import layout
from hydpy_b200.abi import (LandBundle, LandStates, RiverBundle, hb_land_desc, hb_land_states, hb_river_desc,
                            hb_river_states, ptr)

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "hland_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "hydpy_b200.h")
    stale = (not os.path.exists(LIBPATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(LIBPATH) for p in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", HERE, "liboracle.so"] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return LIBPATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIBPATH)
        f64p, i64p = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        _lib.oracle_land_run_range.restype = C.c_int
        _lib.oracle_land_run_range.argtypes = [
            C.POINTER(hb_land_desc), C.POINTER(hb_land_states), C.c_int64, C.c_int64, f64p, i64p, f64p,
            C.POINTER(f64p), C.c_int64, C.c_int64,
        ]
        _lib.oracle_river_run.restype = C.c_int
        _lib.oracle_river_run.argtypes = [C.POINTER(hb_river_desc), f64p, C.c_int64, f64p, f64p, f64p, f64p, f64p]
        _lib.oracle_river_run2.restype = C.c_int
        _lib.oracle_river_run2.argtypes = [C.POINTER(hb_river_desc), f64p, C.c_int64, f64p, f64p, f64p, f64p, f64p, f64p]
        _lib.oracle_river_run3.restype = C.c_int
        _lib.oracle_river_run3.argtypes = [C.POINTER(hb_river_desc), C.POINTER(hb_river_states), C.c_int64, f64p, f64p,
                                           f64p, f64p, f64p, f64p]
        _lib.oracle_river_run4.restype = C.c_int
        _lib.oracle_river_run4.argtypes = _lib.oracle_river_run3.argtypes + [f64p]
        _lib.oracle_sinfactor.restype = C.c_double
        _lib.oracle_sinfactor.argtypes = [C.c_int64]
    return _lib


def land_run(land: LandBundle, states: LandStates, forcing: np.ndarray, doy: np.ndarray,
             series: tuple[str, ...] | list[str] = (), threads: int = 1
             ) -> tuple[np.ndarray, dict[str, np.ndarray]]:
    """Advance all land elements over the window; `states` is updated in place.

    Returns (qt rows [n_elem, nt], {series name: array in the ABI layout}).  `threads` > 1 splits the elements over
    Python threads (the C call releases the GIL) — the same element-major parallelisation as the reference's
    `simulate_multithreaded` (ref: hydpy/core/threadingtools.py:342-394).
    """
    L = lib()
    forcing = np.ascontiguousarray(forcing, dtype=np.float64)
    doy = np.ascontiguousarray(doy, dtype=np.int64)
    nt, ncol = forcing.shape[1], forcing.shape[2]
    d = land.as_struct()
    s = states.as_struct()
    qt = np.zeros((land.n_elem, nt))
    out: dict[str, np.ndarray] = {}
    arr = (C.POINTER(C.c_double) * len(layout.SERIES))()
    for name in series:
        width = land.n_zone if name in layout.SERIES_ZONE else land.n_snow if name in layout.SERIES_SNOW else land.n_uh if name in layout.SERIES_UH else land.n_elem
        out[name] = np.full((nt, width), np.nan)
        arr[layout.S[name]] = ptr(out[name], C.c_double)
    sp = arr if series else None

    def work(e0: int, e1: int) -> None:
        rc = L.oracle_land_run_range(C.byref(d), C.byref(s), nt, ncol, ptr(forcing, C.c_double),
                                     ptr(doy, C.c_int64), ptr(qt, C.c_double), sp, e0, e1)
        if rc != 0:
            raise RuntimeError(f"oracle_land_run_range failed with code {rc}")

    ne = land.n_elem
    threads = max(1, min(int(threads), ne))
    if threads == 1:
        work(0, ne)
    else:
        bounds = np.linspace(0, ne, threads + 1).astype(int)
        errors: list[BaseException] = []

        def guarded(e0: int, e1: int) -> None:
            try:
                work(e0, e1)
            except BaseException as exc:  # pragma: no cover
                errors.append(exc)

        ts = [threading.Thread(target=guarded, args=(int(bounds[i]), int(bounds[i + 1]))) for i in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errors:
            raise errors[0]
    return qt, out


def river_run(river: RiverBundle, discharge: np.ndarray, land_rows: np.ndarray, ext: np.ndarray | None = None,
              dis_series: np.ndarray | None = None, mct_states: np.ndarray | None = None,
              mct_series: np.ndarray | None = None) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Route the window; `discharge` (and `mct_states` [3, n_seg], the musk_mct conditions) are updated in place.
    Returns (node rows, reach outflow, reach inflow).  `dis_series` (optional, [nt, n_seg]) receives
    `states.discharge` after every step, `mct_series` (optional, [len(layout.MCT_SERIES), nt, n_seg], NaN-filled by the
    caller) the 1-D musk_mct sequences after every step."""
    L = lib()
    land_rows = np.ascontiguousarray(land_rows, dtype=np.float64)
    nt = land_rows.shape[1] if land_rows.ndim == 2 and land_rows.shape[0] else (ext.shape[1] if ext is not None else 0)
    if ext is None:
        ext = np.zeros((0, nt))
    ext = np.ascontiguousarray(ext, dtype=np.float64)
    d = river.as_struct()
    node_rows = np.zeros((river.n_node, nt))
    rout = np.zeros((river.n_reach, nt))
    rin = np.zeros((river.n_reach, nt))
    assert discharge.dtype == np.float64 and discharge.flags.c_contiguous
    if dis_series is not None:
        assert dis_series.shape == (nt, river.n_seg) and dis_series.dtype == np.float64 and dis_series.flags.c_contiguous
    st = hb_river_states()
    st.discharge = ptr(discharge, C.c_double)
    if river.any_mct:
        if mct_states is None:
            raise ValueError("the river bundle holds musk_mct reaches: pass their conditions (mct_states [3, n_seg])")
        assert mct_states.shape == (3, river.n_seg) and mct_states.dtype == np.float64 and mct_states.flags.c_contiguous
        st.courantnumber = ptr(mct_states[0], C.c_double)
        st.reynoldsnumber = ptr(mct_states[1], C.c_double)
        st.referencewaterdepth = ptr(mct_states[2], C.c_double)
    if mct_series is not None:
        assert mct_series.shape == (len(layout.MCT_SERIES), nt, river.n_seg) and mct_series.dtype == np.float64
        assert mct_series.flags.c_contiguous
    rc = L.oracle_river_run4(C.byref(d), C.byref(st), nt, ptr(land_rows, C.c_double),
                             ptr(ext, C.c_double), ptr(node_rows, C.c_double), ptr(rout, C.c_double),
                             ptr(rin, C.c_double), ptr(dis_series, C.c_double) if dis_series is not None else None,
                             ptr(mct_series, C.c_double) if mct_series is not None else None)
    if rc != 0:
        raise RuntimeError(f"oracle_river_run failed with code {rc}")
    return node_rows, rout, rin
