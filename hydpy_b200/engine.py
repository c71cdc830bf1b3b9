"""ctypes binding of `libhydpy_b200.so` — the only compute path of this package.

`Engine` mirrors the C-ABI of `include/hydpy_b200.h` one to one.  There is no CPU implementation behind it: if the
shared library is missing or no sm_100 device is usable, every entry point raises (`EngineUnavailable`), it never
falls back.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Sequence

import numpy as np

from . import layout
from .abi import (LandBundle, LandStates, RiverBundle, hb_land_desc, hb_land_states, hb_river_desc, hb_river_states,
                  hb_timing, ptr)

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("HYDPY_B200_LIB") or os.path.join(HERE, "libhydpy_b200.so")

EXPORTS = (
    "hb_abi_version", "hb_create", "hb_destroy", "hb_last_error", "hb_host_alloc", "hb_host_free", "hb_set_land",
    "hb_set_land_states", "hb_get_land_states", "hb_set_river", "hb_set_river_states", "hb_get_river_states",
    "hb_set_river_states2", "hb_get_river_states2",
    "hb_select_series", "hb_set_option", "hb_set_forcing", "hb_set_forcing_async", "hb_select_window", "hb_set_external", "hb_run",
    "hb_run_async", "hb_wait", "hb_get_land_outflow", "hb_get_node_series", "hb_get_node_series_async",
    "hb_node_statistics", "hb_get_reach_outflow", "hb_get_reach_inflow", "hb_get_reach_state_series", "hb_get_mct_series", "hb_get_series", "hb_comm_unique_id", "hb_comm_init",
    "hb_comm_send_rows", "hb_comm_recv_ext", "hb_comm_barrier", "hb_get_timing", "hb_timer_start", "hb_timer_stop",
    "hb_measure_fp64_peak",
    "hb_device_info", "hb_set_members",
)


class EngineUnavailable(RuntimeError):
    """The CUDA engine cannot be used (library not built, or no B200-class device)."""


class EngineError(RuntimeError):
    """A call into libhydpy_b200.so failed; the message carries `hb_last_error`."""


_lib: C.CDLL | None = None


def load_library() -> C.CDLL:
    """Load `libhydpy_b200.so` (built in-tree by `python -m hydpy_b200.build`) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise EngineUnavailable(
            f"{LIBPATH} does not exist; build it with `python -m hydpy_b200.build` (needs nvcc). "
            "hydpy_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIBPATH)
    H = C.c_void_p
    f64p, i32p, i64p = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    proto = {
        "hb_abi_version": (C.c_int, []),
        "hb_create": (C.c_int, [C.c_int, C.POINTER(H)]),
        "hb_destroy": (None, [H]),
        "hb_last_error": (C.c_char_p, [H]),
        "hb_host_alloc": (C.c_int, [H, C.c_int64, C.POINTER(C.c_void_p)]),
        "hb_host_free": (C.c_int, [H, C.c_void_p]),
        "hb_set_land": (C.c_int, [H, C.POINTER(hb_land_desc)]),
        "hb_set_land_states": (C.c_int, [H, C.POINTER(hb_land_states)]),
        "hb_get_land_states": (C.c_int, [H, C.POINTER(hb_land_states)]),
        "hb_set_river": (C.c_int, [H, C.POINTER(hb_river_desc)]),
        "hb_set_river_states2": (C.c_int, [H, C.POINTER(hb_river_states)]),
        "hb_get_river_states2": (C.c_int, [H, C.POINTER(hb_river_states)]),
        "hb_set_river_states": (C.c_int, [H, f64p]),
        "hb_get_river_states": (C.c_int, [H, f64p]),
        "hb_select_series": (C.c_int, [H, C.c_int, C.c_int]),
        "hb_set_option": (C.c_int, [H, C.c_int, C.c_int64]),
        "hb_set_members": (C.c_int, [H, C.c_int64, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.POINTER(C.c_uint8)]),
        "hb_set_forcing": (C.c_int, [H, C.c_int64, C.c_int64, f64p, i64p]),
        "hb_set_forcing_async": (C.c_int, [H, C.c_int64, C.c_int64, f64p, i64p]),
        "hb_select_window": (C.c_int, [H, C.c_int64, C.c_int64]),
        "hb_set_external": (C.c_int, [H, C.c_int64, f64p]),
        "hb_run": (C.c_int, [H, C.c_int]),
        "hb_run_async": (C.c_int, [H, C.c_int]),
        "hb_wait": (C.c_int, [H]),
        "hb_get_node_series_async": (C.c_int, [H, C.c_int64, i32p, f64p]),
        "hb_get_land_outflow": (C.c_int, [H, f64p]),
        "hb_get_node_series": (C.c_int, [H, C.c_int64, i32p, f64p]),
        "hb_get_reach_outflow": (C.c_int, [H, f64p]),
        "hb_get_reach_inflow": (C.c_int, [H, f64p]),
        "hb_node_statistics": (C.c_int, [H, C.c_int64, i32p, i32p, C.c_int64, f64p, C.c_int, f64p]),
        "hb_get_reach_state_series": (C.c_int, [H, f64p]),
        "hb_get_mct_series": (C.c_int, [H, C.c_int, f64p]),
        "hb_get_series": (C.c_int, [H, C.c_int, f64p]),
        "hb_comm_unique_id": (C.c_int, [H, C.c_void_p]),
        "hb_comm_init": (C.c_int, [H, C.c_int, C.c_int, C.c_void_p]),
        "hb_comm_send_rows": (C.c_int, [H, C.c_int, C.c_int, C.c_int64, i32p]),
        "hb_comm_recv_ext": (C.c_int, [H, C.c_int, C.c_int64, i32p]),
        "hb_comm_barrier": (C.c_int, [H]),
        "hb_get_timing": (C.c_int, [H, C.POINTER(hb_timing)]),
        "hb_timer_start": (C.c_int, [H]),
        "hb_timer_stop": (C.c_int, [H, C.POINTER(C.c_double)]),
        "hb_measure_fp64_peak": (C.c_int, [H, C.POINTER(C.c_double)]),
        "hb_device_info": (C.c_int, [H, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]),
    }
    assert set(proto) == set(EXPORTS)
    for name, (res, args) in proto.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.hb_abi_version() != layout.ABI_VERSION:
        raise EngineUnavailable(
            f"{LIBPATH} has ABI version {lib.hb_abi_version()}, the Python layer expects {layout.ABI_VERSION}; rebuild"
        )
    _lib = lib
    return lib


def device_count() -> int:
    """Number of CUDA devices the driver offers (0 if the driver is absent)."""
    try:
        cudart = C.CDLL("libcudart.so")
    except OSError:
        try:
            import glob

            cands = sorted(glob.glob("/usr/local/cuda/lib64/libcudart.so*"))
            cudart = C.CDLL(cands[-1]) if cands else None
        except OSError:
            cudart = None
    if cudart is None:
        return 0
    n = C.c_int(0)
    rc = cudart.cudaGetDeviceCount(C.byref(n))
    return int(n.value) if rc == 0 else 0


class PinnedArray:
    """A float64/int64 NumPy array backed by page-locked host memory of the engine (`hb_host_alloc`)."""

    def __init__(self, engine: "Engine", shape: Sequence[int], dtype=np.float64):
        self._engine = engine
        self.shape = tuple(int(s) for s in shape)
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape, dtype=np.int64)) * dtype.itemsize
        p = C.c_void_p()
        engine._check(engine._lib.hb_host_alloc(engine._h, nbytes, C.byref(p)), "hb_host_alloc")
        self._p = p
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(self.shape, dtype=np.int64))).reshape(self.shape)

    def free(self) -> None:
        if self._p is not None and self._engine._h:
            self.array = None  # type: ignore[assignment]
            self._engine._lib.hb_host_free(self._engine._h, self._p)
        self._p = None


class _Members:
    """Sizes of an ensemble on the device: `members` copies of a base bundle (hb_set_members)."""

    def __init__(self, base, members: int, names: tuple[str, ...]):
        self.base, self.members = base, int(members)
        for n in names:
            setattr(self, n, getattr(base, n) * self.members)

    def __getattr__(self, name):          # everything that is not a size (any_mct, n_ext, …) is the base's
        return getattr(self.base, name)


class Engine:
    """One engine = one GPU (`hb_create`).  Methods map 1:1 to the C entry points."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        rc = self._lib.hb_create(int(device), C.byref(self._h))
        if rc != 0:
            msg = (self._lib.hb_last_error(None) or b"").decode()
            self._h = C.c_void_p()
            raise EngineUnavailable(f"hb_create(device={device}) failed with code {rc}: {msg}")
        self.device = int(device)
        self.land: LandBundle | None = None
        self.river: RiverBundle | None = None
        self.nt = 0
        self._keep: list[object] = []

    # -- plumbing ---------------------------------------------------------------------------------------------------
    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = (self._lib.hb_last_error(self._h) or b"").decode()
            raise EngineError(f"{what} failed with code {rc}: {msg}")

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.hb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self) -> "Engine":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def pinned(self, shape: Sequence[int], dtype=np.float64) -> PinnedArray:
        return PinnedArray(self, shape, dtype)

    def device_info(self) -> dict[str, object]:
        sm, khz = C.c_int(0), C.c_int(0)
        name = C.create_string_buffer(256)
        self._check(self._lib.hb_device_info(self._h, C.byref(sm), C.byref(khz), name, 256), "hb_device_info")
        return {"name": name.value.decode(), "sm_count": sm.value, "clock_khz": khz.value}

    # -- setup ------------------------------------------------------------------------------------------------------
    def set_land(self, land: LandBundle) -> None:
        d = land.as_struct()
        self._check(self._lib.hb_set_land(self._h, C.byref(d)), "hb_set_land")
        self.land = land

    def set_members(self, members: int, mods=None) -> None:
        """Turn the basin on the device (land, river and their conditions) into `members` copies of itself — the batched
        calibration caller (`hb_set_members`).  `mods`: parameter name (a row of `layout.ZPAR` or `layout.EPAR`) →
        (mul[members], add[members]) — member m computes with `value * mul[m] + add[m]` for every zone / element — or a
        list of (name, mul, add, mask) with `mask` = None or a bool array over the basin's elements (a rule bound to a
        selection); list entries act in the given order.  All results then hold `members` times the rows / cells of the
        basin, member after member."""
        mods = mods or {}
        entries = [(n, m_, a_, None) for n, (m_, a_) in mods.items()] if isinstance(mods, dict) else list(mods)
        base_land = self.land.base if isinstance(self.land, _Members) else self._need_land()
        base_river = getattr(self, "river", None)
        base_river = base_river.base if isinstance(base_river, _Members) else base_river
        ids, mul, add, masks = [], [], [], []
        for name, m_, a_, mask_ in entries:
            masks.append(None if mask_ is None else np.asarray(mask_, bool))
            if name in layout.ZP:
                ids.append(layout.ZP[name])
            elif name in layout.EP:
                ids.append(layout.MOD_EPAR + layout.EP[name])
            else:
                raise ValueError(f"unknown parameter {name!r}")
            mul.append(np.broadcast_to(np.asarray(m_, np.float64), (members,)))
            add.append(np.broadcast_to(np.asarray(a_, np.float64), (members,)))
        par = np.ascontiguousarray(ids, dtype=np.int32)
        mul_a = np.ascontiguousarray(np.stack(mul, axis=1) if mul else np.zeros((members, 0)))
        add_a = np.ascontiguousarray(np.stack(add, axis=1) if add else np.zeros((members, 0)))
        mask_a = None
        if any(k is not None for k in masks):
            mask_a = np.ascontiguousarray(np.stack([np.ones(base_land.n_elem, bool) if k is None else k for k in masks])
                                          .astype(np.uint8))
            if mask_a.shape != (len(ids), base_land.n_elem):
                raise ValueError("a modification mask must have one entry per element of the basin")
        self._check(self._lib.hb_set_members(self._h, int(members), len(ids), ptr(par, C.c_int32) if len(ids) else None,
                                             ptr(mul_a, C.c_double) if len(ids) else None,
                                             ptr(add_a, C.c_double) if len(ids) else None,
                                             ptr(mask_a, C.c_uint8) if mask_a is not None else None), "hb_set_members")
        self.members = int(members)
        self.land = _Members(base_land, members, ("n_elem", "n_zone", "n_snow", "n_uh")) if members > 1 else base_land
        if base_river is not None:
            self.river = (_Members(base_river, members, ("n_node", "n_reach", "n_seg")) if members > 1 else base_river)

    def set_land_states(self, states: LandStates) -> None:
        s = states.as_struct()
        self._check(self._lib.hb_set_land_states(self._h, C.byref(s)), "hb_set_land_states")

    def get_land_states(self) -> LandStates:
        land = self._need_land()
        st = LandStates(
            ic=np.zeros(land.n_zone), sm=np.zeros(land.n_zone), sp=np.zeros(land.n_snow), wc=np.zeros(land.n_snow),
            uz=np.zeros(land.n_elem), lz=np.zeros(land.n_elem), quh=np.zeros(land.n_uh),
        )
        s = st.as_struct()
        self._check(self._lib.hb_get_land_states(self._h, C.byref(s)), "hb_get_land_states")
        return st

    def set_river(self, river: RiverBundle) -> None:
        d = river.as_struct()
        self._check(self._lib.hb_set_river(self._h, C.byref(d)), "hb_set_river")
        self.river = river

    def set_river_states(self, discharge: np.ndarray, mct_states: np.ndarray | None = None) -> None:
        """`states.discharge` of all reaches, concatenated; `mct_states` [3, n_seg] = Courant numbers, Reynolds numbers and
        reference water depths of the musk_mct reaches (required when the network holds such reaches)."""
        river = self._need_river()
        discharge = np.ascontiguousarray(discharge, dtype=np.float64)
        if discharge.shape != (river.n_seg,):
            raise ValueError(f"discharge must have shape ({river.n_seg},), not {discharge.shape}")
        if mct_states is None and not river.any_mct:
            self._check(self._lib.hb_set_river_states(self._h, ptr(discharge, C.c_double)), "hb_set_river_states")
            return
        if mct_states is None:
            raise ValueError("the network holds musk_mct reaches: their conditions (mct_states [3, n_seg]) are required")
        mct_states = np.ascontiguousarray(mct_states, dtype=np.float64)
        if mct_states.shape != (3, river.n_seg):
            raise ValueError(f"mct_states must have shape (3, {river.n_seg}), not {mct_states.shape}")
        st = hb_river_states()
        st.discharge = ptr(discharge, C.c_double)
        st.courantnumber = ptr(mct_states[0], C.c_double)
        st.reynoldsnumber = ptr(mct_states[1], C.c_double)
        st.referencewaterdepth = ptr(mct_states[2], C.c_double)
        self._check(self._lib.hb_set_river_states2(self._h, C.byref(st)), "hb_set_river_states2")

    def get_river_mct_states(self) -> np.ndarray:
        """[3, n_seg]: Courant numbers, Reynolds numbers, reference water depths (zeros for musk_classic reaches)."""
        river = self._need_river()
        out = np.zeros((3, river.n_seg))
        dis = np.zeros(river.n_seg)
        st = hb_river_states()
        st.discharge = ptr(dis, C.c_double)
        st.courantnumber = ptr(out[0], C.c_double)
        st.reynoldsnumber = ptr(out[1], C.c_double)
        st.referencewaterdepth = ptr(out[2], C.c_double)
        self._check(self._lib.hb_get_river_states2(self._h, C.byref(st)), "hb_get_river_states2")
        return out

    def get_river_states(self) -> np.ndarray:
        river = self._need_river()
        out = np.zeros(river.n_seg)
        self._check(self._lib.hb_get_river_states(self._h, ptr(out, C.c_double)), "hb_get_river_states")
        return out

    def select_series(self, names: Iterable[str]) -> None:
        names = set(names)
        unknown = names - set(layout.SERIES)
        if unknown:
            raise ValueError(f"unknown series: {sorted(unknown)}")
        for name in layout.SERIES:
            self._check(self._lib.hb_select_series(self._h, layout.S[name], int(name in names)), "hb_select_series")

    def set_land_path(self, path: str = "auto") -> None:
        """'auto': fast two-kernel path where the parameters allow it; 'general': the general kernel for all elements."""
        value = {"auto": layout.LAND_AUTO, "general": layout.LAND_GENERAL}[path]
        self._check(self._lib.hb_set_option(self._h, layout.OPT_LAND_PATH, value), "hb_set_option")

    def set_river_sweep(self, mode: str = "persistent") -> None:
        """'persistent': one persistent wavefront launch, a warp per (reach, chunk); 'grouped': persistent wavefront with
        (32 reaches, chunk) tasks, lane = reach; 'launches': one launch per anti-diagonal."""
        value = {"persistent": layout.RIVER_PERSISTENT, "launches": layout.RIVER_LAUNCHES,
                 "grouped": layout.RIVER_GROUPED}[mode]
        self._check(self._lib.hb_set_option(self._h, layout.OPT_RIVER_SWEEP, value), "hb_set_option")

    def set_land_slabs(self, slabs: int) -> None:
        """Element slabs (streams) of the fast runoff-generation path, 1..8 (performance only; same results)."""
        self._check(self._lib.hb_set_option(self._h, layout.OPT_LAND_SLABS, int(slabs)), "hb_set_option")

    def set_river_plain_min(self, reaches: int) -> None:
        """Minimum number of simple reaches for which a level is routed by plain per-level launches (takes effect with the
        next `set_river`; performance only, same results)."""
        self._check(self._lib.hb_set_option(self._h, layout.OPT_RIVER_PLAIN_MIN, int(reaches)), "hb_set_option")

    def set_land_fuse(self, on: bool) -> None:
        """Zone terms added inside the thread-per-zone kernel where a 32-element group allows (default on; same results)."""
        self._check(self._lib.hb_set_option(self._h, layout.OPT_LAND_FUSE, 1 if on else 0), "hb_set_option")

    def set_land_speculate(self, on: bool) -> None:
        """Finite `SMax`: fast path while no snow class exceeds its limit, windows in which one does are repeated by the
        general kernel (default on; read by the next `set_land`; same results)."""
        self._check(self._lib.hb_set_option(self._h, layout.OPT_LAND_SPECULATE, 1 if on else 0), "hb_set_option")

    def set_land_chunk(self, steps: int) -> None:
        """Time chunk of the fast runoff-generation path in steps (0 = as long as the scratch memory allows; -1 = the engine's default)."""
        self._check(self._lib.hb_set_option(self._h, layout.OPT_LAND_CHUNK, int(steps)), "hb_set_option")

    def set_wave_blocks(self, blocks_per_sm: int) -> None:
        """Persistent routing wavefront blocks per SM (1..8): how much of every SM the routing of window w takes from the
        runoff generation of window w+1."""
        self._check(self._lib.hb_set_option(self._h, layout.OPT_RIVER_WAVE_BLOCKS, int(blocks_per_sm)), "hb_set_option")

    # -- window -----------------------------------------------------------------------------------------------------
    def set_forcing(self, forcing: np.ndarray, doy: np.ndarray, asynchronous: bool = False) -> None:
        """Stage the forcing block [4, nt, n_col] of the next window.  `asynchronous=True` only enqueues the copy
        (`hb_set_forcing_async`): `forcing` must then be a C-contiguous float64 view of PINNED memory (`Engine.pinned`)
        that stays untouched until `wait()`."""
        if asynchronous and not (forcing.dtype == np.float64 and forcing.flags.c_contiguous):
            raise ValueError("asynchronous staging needs a C-contiguous float64 array in pinned memory")
        forcing = np.ascontiguousarray(forcing, dtype=np.float64)
        if forcing.ndim != 3 or forcing.shape[0] != len(layout.FORCING):
            raise ValueError(f"forcing must have shape (4, nt, n_col), not {forcing.shape}")
        doy = np.ascontiguousarray(doy, dtype=np.int64)
        nt, ncol = forcing.shape[1], forcing.shape[2]
        if doy.shape != (nt,):
            raise ValueError(f"doy must have shape ({nt},), not {doy.shape}")
        land = self._need_land()
        if land.n_elem and int(land.forcing_col.max()) >= ncol:
            raise ValueError("forcing has fewer columns than `forcing_col` refers to")
        fn = self._lib.hb_set_forcing_async if asynchronous else self._lib.hb_set_forcing
        self._check(fn(self._h, nt, ncol, ptr(forcing, C.c_double), ptr(doy, C.c_int64)), "hb_set_forcing")
        self.nt = nt

    def select_window(self, t0: int, nt: int) -> None:
        """Run the next `run()` over steps [t0, t0+nt) of the staged forcing block."""
        self._check(self._lib.hb_select_window(self._h, int(t0), int(nt)), "hb_select_window")
        self.nt = int(nt)

    def set_external(self, rows: np.ndarray) -> None:
        river = self._need_river()
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(river.n_ext, self.nt)
        self._check(self._lib.hb_set_external(self._h, self.nt, ptr(rows, C.c_double)), "hb_set_external")

    def run(self, land: bool = True, river: bool = True) -> None:
        flags = (layout.RUN_LAND if land else 0) | (layout.RUN_RIVER if river else 0)
        self._check(self._lib.hb_run(self._h, flags), "hb_run")

    def run_async(self, land: bool = True, river: bool = True) -> None:
        """Enqueue the window and return: the next window's runoff generation overlaps this window's routing and
        result copies (see `hb_run_async`).  `wait()` — or any synchronous call — collects the work."""
        flags = (layout.RUN_LAND if land else 0) | (layout.RUN_RIVER if river else 0)
        self._check(self._lib.hb_run_async(self._h, flags), "hb_run_async")

    def wait(self) -> None:
        self._check(self._lib.hb_wait(self._h), "hb_wait")

    def get_node_series_async(self, out: np.ndarray, nodes: np.ndarray | None = None) -> None:
        """Copy the node series of the current window — all nodes, or the rows of `nodes` (e.g. the gauged ones) — into
        the PINNED array `out` behind its routing; valid after `wait()`."""
        river = self._need_river()
        n_rows = river.n_node if nodes is None else len(nodes)
        if out.shape != (n_rows, self.nt) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape (number of rows, nt)")
        if nodes is None:
            self._check(self._lib.hb_get_node_series_async(self._h, 0, None, ptr(out, C.c_double)),
                        "hb_get_node_series_async")
        elif n_rows:
            sel = np.ascontiguousarray(nodes, dtype=np.int32)
            self._check(self._lib.hb_get_node_series_async(self._h, n_rows, ptr(sel, C.c_int32), ptr(out, C.c_double)),
                        "hb_get_node_series_async")

    @property
    def window_steps(self) -> int:
        return self.nt

    # -- results ----------------------------------------------------------------------------------------------------
    def get_land_outflow(self, out: np.ndarray | None = None) -> np.ndarray:
        land = self._need_land()
        out = self._out(out, (land.n_elem, self.nt))
        self._check(self._lib.hb_get_land_outflow(self._h, ptr(out, C.c_double)), "hb_get_land_outflow")
        return out

    def get_node_series(self, nodes: Sequence[int] | None = None, out: np.ndarray | None = None) -> np.ndarray:
        river = self._need_river()
        if nodes is None:
            out = self._out(out, (river.n_node, self.nt))
            self._check(self._lib.hb_get_node_series(self._h, 0, None, ptr(out, C.c_double)), "hb_get_node_series")
            return out
        sel = np.ascontiguousarray(nodes, dtype=np.int32)
        out = self._out(out, (len(sel), self.nt))
        if len(sel):
            self._check(self._lib.hb_get_node_series(self._h, len(sel), ptr(sel, C.c_int32), ptr(out, C.c_double)),
                        "hb_get_node_series")
        return out

    def node_statistics(self, nodes: Sequence[int], obs: np.ndarray, obs_row: Sequence[int] | None = None,
                        skip_nan: bool = False) -> np.ndarray:
        """Sufficient statistics (columns = `layout.STATS`) of the simulated series of `nodes` (last window) against the
        rows of `obs` — pair i compares `nodes[i]` with `obs[obs_row[i]]` (default: row i).  See `hydpy_b200.calib`."""
        nodes = np.ascontiguousarray(nodes, dtype=np.int32)
        obs = np.ascontiguousarray(obs, dtype=np.float64).reshape(-1, self.nt)
        rows = np.arange(len(nodes), dtype=np.int32) if obs_row is None else np.ascontiguousarray(obs_row, dtype=np.int32)
        if len(rows) != len(nodes):
            raise ValueError("`obs_row` must have one entry per node")
        out = np.zeros((len(nodes), len(layout.STATS)))
        self._check(self._lib.hb_node_statistics(self._h, len(nodes), ptr(nodes, C.c_int32), ptr(rows, C.c_int32),
                                                 obs.shape[0], ptr(obs, C.c_double), int(bool(skip_nan)),
                                                 ptr(out, C.c_double)), "hb_node_statistics")
        return out

    def get_reach_outflow(self) -> np.ndarray:
        river = self._need_river()
        out = np.zeros((river.n_reach, self.nt))
        self._check(self._lib.hb_get_reach_outflow(self._h, ptr(out, C.c_double)), "hb_get_reach_outflow")
        return out

    def record_reach_states(self, on: bool = True) -> None:
        """Record musk_classic `states.discharge` of every step (read with `get_reach_state_series`)."""
        self._check(self._lib.hb_set_option(self._h, layout.OPT_RIVER_STATE_SERIES, int(bool(on))), "hb_set_option")

    def get_reach_state_series(self) -> np.ndarray:
        river = self._need_river()
        out = np.zeros((self.nt, river.n_seg))
        self._check(self._lib.hb_get_reach_state_series(self._h, ptr(out, C.c_double)), "hb_get_reach_state_series")
        return out

    def record_mct_series(self, names: Iterable[str] = ()) -> None:
        """musk_mct sequences (names of `layout.MCT_SERIES`) to record at every step (read with `get_mct_series`)."""
        mask = 0
        for name in names:
            if name not in layout.MS:
                raise ValueError(f"unknown musk_mct sequence `{name}` (known: {', '.join(layout.MCT_SERIES)})")
            mask |= 1 << layout.MS[name]
        self._check(self._lib.hb_set_option(self._h, layout.OPT_RIVER_MCT_SERIES, mask), "hb_set_option")

    def get_mct_series(self, name: str) -> np.ndarray:
        """[nt, n_seg]: segment i of reach r at seg_ptr[r] + i, NaN elsewhere."""
        river = self._need_river()
        out = np.zeros((self.nt, river.n_seg))
        self._check(self._lib.hb_get_mct_series(self._h, layout.MS[name], ptr(out, C.c_double)), "hb_get_mct_series")
        return out

    def get_reach_inflow(self) -> np.ndarray:
        river = self._need_river()
        out = np.zeros((river.n_reach, self.nt))
        self._check(self._lib.hb_get_reach_inflow(self._h, ptr(out, C.c_double)), "hb_get_reach_inflow")
        return out

    def get_series(self, name: str) -> np.ndarray:
        land = self._need_land()
        width = land.n_zone if name in layout.SERIES_ZONE else land.n_snow if name in layout.SERIES_SNOW else land.n_uh if name in layout.SERIES_UH else land.n_elem
        out = np.zeros((self.nt, width))
        self._check(self._lib.hb_get_series(self._h, layout.S[name], ptr(out, C.c_double)), "hb_get_series")
        return out

    def timer_start(self) -> None:
        self._check(self._lib.hb_timer_start(self._h), "hb_timer_start")

    def timer_stop(self) -> float:
        """Milliseconds of device time on the engine stream since `timer_start` (synchronises)."""
        v = C.c_double(0.0)
        self._check(self._lib.hb_timer_stop(self._h, C.byref(v)), "hb_timer_stop")
        return float(v.value)

    def measure_fp64_peak(self) -> float:
        """Sustained DFMA throughput of the device in TFLOP/s (the FP64 roofline denominator)."""
        v = C.c_double(0.0)
        self._check(self._lib.hb_measure_fp64_peak(self._h, C.byref(v)), "hb_measure_fp64_peak")
        return float(v.value)

    def timing(self) -> dict[str, float]:
        t = hb_timing()
        self._check(self._lib.hb_get_timing(self._h, C.byref(t)), "hb_get_timing")
        return {n: getattr(t, n) for n, _ in hb_timing._fields_}

    # -- multi-GPU --------------------------------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(layout.NCCL_ID_BYTES)
        self._check(self._lib.hb_comm_unique_id(self._h, buf), "hb_comm_unique_id")
        return buf.raw

    def comm_init(self, rank: int, world: int, uid: bytes) -> None:
        if len(uid) != layout.NCCL_ID_BYTES:
            raise ValueError("NCCL unique id must have 128 bytes")
        buf = C.create_string_buffer(uid, layout.NCCL_ID_BYTES)
        self._check(self._lib.hb_comm_init(self._h, rank, world, buf), "hb_comm_init")

    def comm_send_rows(self, peer: int, rows: Sequence[int], kind: str = "node") -> None:
        sel = np.ascontiguousarray(rows, dtype=np.int32)
        k = {"node": layout.ROW_NODE, "reach": layout.ROW_REACH}[kind]
        self._check(self._lib.hb_comm_send_rows(self._h, peer, k, len(sel), ptr(sel, C.c_int32)), "hb_comm_send_rows")

    def comm_recv_ext(self, peer: int, rows: Sequence[int]) -> None:
        sel = np.ascontiguousarray(rows, dtype=np.int32)
        self._check(self._lib.hb_comm_recv_ext(self._h, peer, len(sel), ptr(sel, C.c_int32)), "hb_comm_recv_ext")

    def comm_barrier(self) -> None:
        self._check(self._lib.hb_comm_barrier(self._h), "hb_comm_barrier")

    # -- helpers ----------------------------------------------------------------------------------------------------
    def _need_land(self) -> LandBundle:
        if self.land is None:
            raise EngineError("no land description set (call set_land first)")
        return self.land

    def _need_river(self) -> RiverBundle:
        if self.river is None:
            raise EngineError("no river description set (call set_river first)")
        return self.river

    @staticmethod
    def _out(out: np.ndarray | None, shape: tuple[int, ...]) -> np.ndarray:
        if out is None:
            return np.zeros(shape)
        if out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError(f"output array must be C-contiguous float64 of shape {shape}")
        return out
