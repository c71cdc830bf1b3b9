#!/usr/bin/env python
"""bench.py — zone-timesteps/s of the hland_96 + musk_classic hot path on N B200s (contract: see the task prompt).

Workload (``config.workload``): BASELINE.json configs[3] — the synthetic 10^5-element hland_96 + 10^5-reach
musk_classic river tree (≈200 topological levels), HOURLY, discharge-only — the configuration the
north_star's ≥100x target is quoted on.  One *step* = one simulation window of ``--window`` hours (default 438: the 20
timed steps of the driver cover one year) for ALL elements followed by routing through the whole network; consecutive
steps continue the same simulation (states stay on the device).

* ``value``  — whole-job zone-timesteps/s with the forcing of all K windows already resident in HBM (device time, CUDA
  events on the engine stream, max over ranks).
* ``e2e``    — the same loop through the public API with HOST buffers: every step copies that window's forcing from
  pinned host memory to the device and reads every node's discharge series back to pinned host memory.
* N > 1     — ONE river basin, cut by ``hydpy_b200.shard`` into N sub-networks of equal weight (one process per GPU):
  every boundary reach hands its outflow series to the rank that owns the downstream node with NCCL send/recv over
  NVLink, and the receiving rank's routing WAITS for it (the one real exchange step of the path).
  ``--scaling weak`` (default): the basin grows with N (N x 10^5 elements, the same depth); ``--scaling strong``: the
  10^5-element basin of N = 1 is cut into N parts.  The last rank re-runs the first window of the WHOLE basin on its own
  GPU and compares its nodes bit for bit (``multi_gpu_parity``).
* ``--config 2`` — BASELINE.json configs[2]: the 10^5-element basin with DAILY steps (1200 RecStep sub-steps per step), runoff
  generation only; a step is one simulated year.
* ``--config 1`` / ``--config 4`` — BASELINE.json configs[1] (HydPy-H-Lahn hourly x 4096 parameter-ensemble members) and
  configs[4] (10^6 zones x 512 members), member-sharded over the ranks (no exchange: every member is a complete basin).
* ``--impl reference`` — the reference's own CPU implementation of the same window on a bounded element sample with all
  host threads (oracle/_ref = unmodified reference binaries when present, else the C restatement).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "zone-timesteps/sec (hland_96+musk)"
UNIT = "zone-timesteps/s"


def parse_args() -> argparse.Namespace:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("engine", "reference"), default="engine")
    ap.add_argument("--elements", type=int, default=100_000, help="hland_96 elements per GPU")
    ap.add_argument("--window", type=int, default=438,
                    help="simulation steps (hours) per bench step; 20 timed steps of 438 h span one year")
    ap.add_argument("--forcing", choices=("storms", "dry"), default="storms",
                    help="synthetic forcing (hydpy_b200.synth.make_forcing): 'storms' = event-based, humid water balance "
                         "(default); 'dry' = round 1's hourly drizzle that never reaches the soil")
    ap.add_argument("--levels", type=int, default=200)
    ap.add_argument("--cpu-elements", type=int, default=0, help="element sample of the CPU baseline (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--wave-blocks", type=int, default=0, help="persistent routing blocks per SM (0 = engine default)")
    ap.add_argument("--river-sweep", default="", help="routing sweep: persistent | grouped | launches (default: engine's)")
    ap.add_argument("--scaling", choices=("weak", "strong"), default="weak",
                    help="N > 1: weak = one basin of N x --elements elements, strong = the --elements basin cut into N parts")
    ap.add_argument("--config", type=int, choices=(1, 2, 3, 4), default=3, help="BASELINE.json configs[i] (default 3)")
    ap.add_argument("--members", type=int, default=0, help="ensemble members of --config 1 / 4 in total (0 = 4096 / 512)")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the bit-for-bit check against the unsharded run")
    ap.add_argument("--land-slabs", type=int, default=0, help="element slabs of the fast land path (0 = engine default)")
    ap.add_argument("--river-plain-min", type=int, default=-1,
                    help="minimum width of a routing level for plain per-level launches (-1 = engine default)")
    ap.add_argument("--no-land-fuse", action="store_true",
                    help="hand the zone terms over through HBM instead of adding them inside the zone kernel (A/B timing)")
    ap.add_argument("--land-chunk", type=int, default=0, help="time chunk of the fast land path in steps (0 = engine default)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# distributed plumbing (torch.distributed only for rendezvous / barrier / reductions; the data path is the engine's NCCL)
# ---------------------------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self) -> None:
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.pg = None
        if self.world > 1:
            import torch.distributed as dist

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(backend="gloo", rank=self.rank, world_size=self.world)
            self.pg = dist

    def barrier(self) -> None:
        if self.pg is not None:
            self.pg.barrier()

    def max(self, value: float) -> float:
        if self.pg is None:
            return value
        import torch

        t = torch.tensor([value], dtype=torch.float64)
        self.pg.all_reduce(t, op=self.pg.ReduceOp.MAX)
        return float(t[0])

    def min(self, value: float) -> float:
        return -self.max(-value)

    def bcast_float(self, value: float, src: int) -> float:
        if self.pg is None:
            return value
        import torch

        t = torch.tensor([value], dtype=torch.float64)
        self.pg.broadcast(t, src=src)
        return float(t[0])

    def all_gather(self, values: list[float]) -> list[list[float]]:
        """One row of `values` per rank."""
        if self.pg is None:
            return [list(values)]
        import torch

        mine = torch.tensor(values, dtype=torch.float64)
        out = [torch.zeros_like(mine) for _ in range(self.world)]
        self.pg.all_gather(out, mine)
        return [[float(x) for x in t] for t in out]

    def sum(self, value: float) -> float:
        if self.pg is None:
            return value
        import torch

        t = torch.tensor([value], dtype=torch.float64)
        self.pg.all_reduce(t, op=self.pg.ReduceOp.SUM)
        return float(t[0])

    def bcast_bytes(self, data: bytes | None, n: int) -> bytes:
        if self.pg is None:
            assert data is not None
            return data
        import torch

        t = torch.zeros(n, dtype=torch.uint8)
        if self.rank == 0:
            t[:] = torch.frombuffer(bytearray(data), dtype=torch.uint8)
        self.pg.broadcast(t, src=0)
        return bytes(t.numpy().tobytes())

    def close(self) -> None:
        if self.pg is not None:
            self.pg.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int) -> None:
        self.proc = None
        self.path = None
        exe = shutil.which("nvidia-smi")
        if exe:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                [exe, f"--id={device}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self) -> dict | None:
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------------
class Workload:
    """What one rank simulates: its bundles, the exchange plan of a window and how to describe / verify it."""

    def __init__(self, land, states, river, discharge, recv=(), send=(), *, desc: dict, forcing_seed: int = 0,
                 forcing_fixed=None, members: int = 1, verify=None, step: str = "1h"):
        self.step = step                  # simulation step of the synthetic forcing
        self.land, self.states, self.river, self.discharge = land, states, river, discharge
        self.recv, self.send = list(recv), list(send)
        self.desc, self.forcing_seed, self.forcing_fixed = desc, forcing_seed, forcing_fixed
        self.members = members            # ensemble members on this rank (created on the device: hb_set_members)
        self.verify = verify


def build_shard(args: argparse.Namespace, rank: int, world: int) -> Workload:
    """configs[3]: this rank's part of ONE river basin.  N = 1: the whole basin.  The basin is generated
    as a whole where that is cheap (river tree, random draws of the land parameters); the land bundles — the bulk of the
    memory — only for the rank's own elements."""
    from hydpy_b200 import shard, synth

    n_total = args.elements * (world if args.scaling == "weak" else 1)
    river_g, dis_g = synth.make_tree_river(n_total, "1h", seed=2, levels=args.levels)
    desc = {"workload": "configs[3]: synthetic 10^5-element hland_96 + musk_classic river tree, hourly, discharge-only",
            "basin_elements": n_total}
    if world == 1:
        land, states = synth.make_land(n_total, "1h", seed=1)
        return Workload(land, states, river_g, dis_g, desc=desc)
    outlet = np.arange(n_total)
    node_rank = shard.assign_nodes(None, river_g, world, outlet_node=outlet, zones=np.full(n_total, 12.0))
    s = shard.partition(None, None, river_g, dis_g, world, node_rank=node_rank, ranks=[rank], outlet_node=outlet)[rank]
    # the shard's elements in the order of their forcing columns (a shard is a scattered subset of the basin's elements:
    # in index order the 32 elements of a warp would read 32 distant columns of the forcing bank, in this order they read
    # the same or neighbouring ones, like the consecutive elements of the whole basin do).  Element g drains to node g.
    order = np.argsort(s.elements % 1024, kind="stable")
    land, states = synth.make_land(n_total, "1h", seed=1, select=s.elements[order])
    land.outlet_node = order.astype(np.int32)              # local node of local element j = the old local position
    inv = np.empty(len(order), np.int64)
    inv[order] = np.arange(len(order))
    src = s.river.entry_src.astype(np.int64)
    s.river.entry_src = np.where(src >= 0, inv[np.maximum(src, 0)], src).astype(np.int32)
    land.normalise()
    s.river.normalise()
    desc.update({"shard_elements": int(len(s.elements)), "shard_stage": int(s.stage),
                 "boundary_rows_received": int(sum(len(r) for _, r in s.recv)),
                 "boundary_rows_sent": int(sum(len(r) for _, r in s.send))})

    def verify(eng, forcing, doy, rows_sharded: np.ndarray) -> bool:
        """The whole basin on this rank's GPU, first window from the initial states: this rank's nodes, bit for bit."""
        g_land, g_states = synth.make_land(n_total, "1h", seed=1)
        eng.set_land(g_land)
        eng.set_land_states(g_states)
        eng.set_river(river_g)
        eng.set_river_states(dis_g)
        eng.set_forcing(forcing, doy)
        eng.set_external(np.zeros((0, forcing.shape[1])))
        eng.run(land=True, river=True)
        return bool(np.array_equal(eng.get_node_series(nodes=s.nodes), rows_sharded))

    return Workload(land, states, s.river, s.discharge, s.recv, s.send, desc=desc, verify=verify)


def build_members(args: argparse.Namespace, rank: int, world: int) -> Workload:
    """configs[1] / configs[4]: a parameter ensemble, member-sharded (every member is a complete basin, so nothing crosses
    ranks).  The rank describes ONE basin to the engine; its members are created on the device (`hb_set_members`)."""
    from hydpy_b200 import synth

    total = args.members or (4096 if args.config == 1 else 512)
    per_rank = max(1, total // world)
    lo = rank * per_rank
    mods_all = synth.member_mods(per_rank * world)            # member 0 = the shipped parameter values
    mods = {k: (m[lo:lo + per_rank], a[lo:lo + per_rank]) for k, (m, a) in mods_all.items()}
    if args.config == 1:
        from tests.util import load_scenarios                  # (the committed fixture of the real Lahn basin)

        p = load_scenarios("lahn_hourly_members.npz")["member0"].problem
        desc = {"workload": f"configs[1]: HydPy-H-Lahn hourly x {per_rank * world} parameter-ensemble members "
                            f"(4 hland_96 elements, 49 zones, 3 musk_classic reaches per member), members created on the "
                            f"device (hb_set_members)", "basin_elements": p.land.n_elem}
        wl = Workload(p.land, p.states, p.river, p.discharge, desc=desc, forcing_fixed=(p.forcing, p.doy), members=per_rank)
    else:
        land, states = synth.make_land(83_334, "1h", seed=1)
        river, discharge = synth.make_tree_river(83_334, "1h", seed=2, levels=args.levels)
        desc = {"workload": f"configs[4]: 10^6-zone hland_96 + musk_classic basin (83 334 elements) x {per_rank * world} "
                            f"ensemble members, hourly, members created on the device (hb_set_members)",
                "basin_elements": land.n_elem}
        wl = Workload(land, states, river, discharge, desc=desc, members=per_rank)
    wl.mods = mods
    wl.desc.update({"members_total": per_rank * world, "members_per_gpu": per_rank})
    return wl


def build_daily(args: argparse.Namespace, rank: int, world: int) -> Workload:
    """configs[2]: 10^5 hland_96 elements x 12 zones, DAILY steps (RecStep = 1200 sub-steps per step), runoff generation only:
    every element drains into a node of its own, no reaches.  N > 1: independent replicas of the basin's elements (the
    elements do not interact: nothing to exchange)."""
    from hydpy_b200 import synth
    from hydpy_b200.abi import RiverBundle

    ne = args.elements
    land, states = synth.make_land(ne, "1d", seed=1 + rank)
    river = RiverBundle.empty(ne)
    river.entry_ptr = np.arange(ne + 1, dtype=np.int64)
    river.entry_src = np.arange(ne, dtype=np.int32)
    river.normalise()
    desc = {"workload": "configs[2]: synthetic 10^5-element hland_96 basin, 12 zones per element, daily steps (1200 RecStep "
                        "sub-steps per step), runoff generation only (every element drains into its own node, no reaches)",
            "basin_elements": ne * world}
    return Workload(land, states, river, np.zeros(river.n_seg), desc=desc, step="1d")


def build_workload(args: argparse.Namespace, dist: "Dist") -> Workload:
    if args.config == 3:
        return build_shard(args, dist.rank, dist.world)
    if args.config == 2:
        return build_daily(args, dist.rank, dist.world)
    return build_members(args, dist.rank, dist.world)


def pin_to_gpu_numa_node(device: int) -> None:
    """Run this rank (and allocate its pinned staging buffers) on the CPU cores next to its GPU: with 8 ranks the
    device→host copies of the e2e loop otherwise cross the socket interconnect."""
    try:
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(device))
    except Exception:   # no NVML, no permission, single-socket box: nothing to gain
        pass


def run_engine(args: argparse.Namespace, dist: Dist) -> dict:
    from hydpy_b200 import Engine, synth

    if dist.world > 1:
        pin_to_gpu_numa_node(dist.local_rank)

    K, W, win = args.steps, args.warmup, args.window
    t_setup = time.perf_counter()
    eng = Engine(dist.local_rank)
    info = eng.device_info()
    wl = build_workload(args, dist)
    if wl.forcing_fixed is not None:
        win = wl.forcing_fixed[0].shape[1]
    elif args.config == 2 and win == 438:
        win = 365        # one simulated year per step
    elif args.config == 4 and win == 438:
        win = 120        # 5·10^6 nodes per GPU: the node/reach rows of a 438-h window would not fit beside the constants
    states, discharge = wl.states, wl.discharge
    if args.river_plain_min >= 0:
        eng.set_river_plain_min(args.river_plain_min)
    eng.set_land(wl.land)
    eng.set_land_states(states)
    eng.set_river(wl.river)
    eng.set_river_states(discharge)
    if wl.members > 1:
        eng.set_members(wl.members, wl.mods)
    land, river = eng.land, eng.river       # (sizes of everything on the device: all members)
    ensemble = wl.members > 1
    import resource

    print(f"[bench] rank {dist.rank}: workload on the device after {time.perf_counter() - t_setup:.1f} s "
          f"({land.n_elem} elements, {river.n_reach} reaches; host peak RSS "
          f"{resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6:.1f} GB)", file=sys.stderr, flush=True)

    def reset_conditions() -> None:
        """Back to the initial conditions (an ensemble simply goes on: its conditions live on the device only)."""
        if not ensemble:
            eng.set_land_states(states)
            eng.set_river_states(discharge)

    if args.river_sweep:
        eng.set_river_sweep(args.river_sweep)
    if args.wave_blocks:
        eng.set_wave_blocks(args.wave_blocks)
    if args.land_slabs:
        eng.set_land_slabs(args.land_slabs)
    if args.no_land_fuse:
        eng.set_land_fuse(False)
    if args.land_chunk:
        eng.set_land_chunk(args.land_chunk)
    if dist.world > 1:
        uid = dist.bcast_bytes(eng.comm_unique_id() if dist.rank == 0 else None, 128)
        eng.comm_init(dist.rank, dist.world, uid)
    own_ext = np.zeros((river.n_ext, win))      # (no rows of the global external block in these workloads: all zero)

    def step_exchange_and_route() -> None:
        """land → (NVLink hand-over of the boundary series) → routing, for the currently staged window.  Enqueued on
        the engine's pipeline (hb_run_async): the next window's runoff generation overlaps this window's routing, and
        an upstream rank works on window w+1 while the ranks below it still route window w."""
        if not wl.recv:
            eng.set_external(own_ext)
            eng.run_async(land=True, river=True)
        else:
            eng.run_async(land=True, river=False)
            for peer, rows in wl.recv:                      # straight into the external rows the reaches read
                eng.comm_recv_ext(peer, rows)
            eng.run_async(land=False, river=True)           # (waits on the device for the received rows)
        for peer, reaches in wl.send:                       # straight from the routing result buffer
            eng.comm_send_rows(peer, reaches, kind="reach")

    n_steps = W + K
    # ---- forcing of all windows: generated up front (synthetic, seeded per window) in pinned host memory ----------------
    bank = wl.forcing_fixed[0].shape[2] if wl.forcing_fixed is not None else 1024
    pinned = eng.pinned((n_steps, 4, win, bank))
    doys = np.zeros((n_steps, win), np.int64)
    for s in range(n_steps):
        if wl.forcing_fixed is not None:
            f, doys[s] = wl.forcing_fixed
        else:
            f, doys[s] = synth.make_forcing(win, wl.step, seed=wl.forcing_seed, t0=s * win, kind=args.forcing)
        pinned.array[s] = f
    zone_steps_per_step = land.n_zone * win
    launches = 0

    # ================= device-resident run (value) =====================================================================
    block = np.ascontiguousarray(np.concatenate(list(pinned.array), axis=1))      # [4, n_steps*win, 1024]
    eng.set_forcing(block, doys.reshape(-1))
    del block
    dist.barrier()
    for s in range(W):
        eng.select_window(s * win, win)
        step_exchange_and_route()
    eng.wait()
    eng.timing()                                   # reset the engine's per-kernel sums
    dist.barrier()
    sampler = ClockSampler(dist.local_rank)
    eng.timer_start()
    t_wall = time.perf_counter()
    for s in range(W, W + K):
        eng.select_window(s * win, win)
        step_exchange_and_route()
    device_ms = eng.timer_stop()                   # joins the land, routing and copy streams, then synchronises
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    clocks = sampler.stop()
    tm = eng.timing()                              # CUDA-event sums over the K timed steps, per kernel family
    land_ms, river_ms = [tm["land_ms"] / K], [tm["river_ms"] / K]
    zone_ms, element_ms, general_ms = [tm["zone_ms"] / K], [tm["element_ms"] / K], [tm["general_ms"] / K]
    launches = int(tm["land_launches"] + tm["river_launches"])
    n_fast = int(tm["n_fast"])
    dist.barrier()
    per_rank = dist.all_gather([device_ms / K, tm["land_ms"] / K, tm["river_ms"] / K, wall_ms / K])
    device_ms_max = dist.max(device_ms)
    total_zone_steps = dist.sum(float(zone_steps_per_step * K))
    value = total_zone_steps / (device_ms_max * 1e-3)
    n_levels = int(tm["n_levels"])
    states_after_value = None if ensemble else eng.get_land_states()

    # ================= end-to-end run through the public API with host buffers (e2e) =====================================
    # every step: forcing of the window host -> device, node series device -> host.  A single basin returns ALL its node
    # series; an ensemble the series of every member's outlet (what a calibration run reads)
    reset_conditions()
    if ensemble:
        nn_base = wl.river.n_node
        terminal = int(np.setdiff1d(np.arange(nn_base), wl.river.inlet_node)[0]) if nn_base else 0
        e2e_nodes = (np.arange(wl.members) * nn_base + terminal).astype(np.int32)
    else:
        e2e_nodes = None
    n_out = river.n_node if e2e_nodes is None else len(e2e_nodes)
    outs = [eng.pinned((n_out, win)), eng.pinned((n_out, win))]
    h2d = 4 * win * bank * 8 + win * 8
    d2h = n_out * win * 8
    dist.barrier()
    for s in range(W):
        eng.set_forcing(pinned.array[s], doys[s], asynchronous=True)   # host → device copy of this window's forcing
        step_exchange_and_route()
        eng.get_node_series_async(outs[s & 1].array, e2e_nodes)   # device → host copy of the node series
    eng.wait()
    dist.barrier()
    t0 = time.perf_counter()
    for s in range(W, W + K):
        eng.set_forcing(pinned.array[s], doys[s], asynchronous=True)
        step_exchange_and_route()
        eng.get_node_series_async(outs[s & 1].array, e2e_nodes)
    eng.wait()                                                    # all K windows' results are in host memory
    e2e_s = time.perf_counter() - t0
    dist.barrier()
    e2e_s_max = dist.max(e2e_s)
    e2e_value = total_zone_steps / e2e_s_max
    checksum = float(outs[(W + K - 1) & 1].array[0].sum())
    if ensemble:
        same = None
    else:
        states_after_e2e = eng.get_land_states()
        same = all(np.array_equal(getattr(states_after_value, n), getattr(states_after_e2e, n)) for n in states.NAMES)

    # ---- the same with the gauged nodes only (1 % of the nodes, hb_get_node_series_async with a selection): what a user
    # who keeps the series of the gauges pays; at 8 ranks the copy of ALL node series is bound by the host's memory
    gauges = np.sort(np.random.default_rng(5).choice(river.n_node, size=max(1, river.n_node // 100), replace=False))
    if ensemble:
        gauges = np.asarray(e2e_nodes)
    gouts = [eng.pinned((len(gauges), win)), eng.pinned((len(gauges), win))]
    reset_conditions()
    dist.barrier()
    for s in range(W if not ensemble else 0):
        eng.set_forcing(pinned.array[s], doys[s], asynchronous=True)
        step_exchange_and_route()
        eng.get_node_series_async(gouts[s & 1].array, gauges)
    eng.wait()
    dist.barrier()
    t0 = time.perf_counter()
    for s in range(W, W + K) if not ensemble else ():
        eng.set_forcing(pinned.array[s], doys[s], asynchronous=True)
        step_exchange_and_route()
        eng.get_node_series_async(gouts[s & 1].array, gauges)
    eng.wait()
    e2e_g_s = time.perf_counter() - t0 if not ensemble else e2e_s
    dist.barrier()
    e2e_gauged = {"value": total_zone_steps / dist.max(e2e_g_s), "unit": UNIT, "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": int(len(gauges)) * win * 8, "ms_per_step": 1e3 * dist.max(e2e_g_s) / K,
                  "nodes_returned_per_rank": int(len(gauges)),
                  "same_rows_as_full_copy": None if ensemble else bool(np.array_equal(
                      gouts[(W + K - 1) & 1].array, outs[(W + K - 1) & 1].array[gauges]))}
    for o in gouts:
        o.free()

    # ================= N > 1: the sharded run against the unsharded one ==================================================
    parity = None
    if dist.world > 1 and not args.no_verify and wl.verify is not None:
        reset_conditions()
        eng.set_forcing(np.ascontiguousarray(pinned.array[0]), doys[0])
        step_exchange_and_route()
        eng.wait()
        rows0 = eng.get_node_series()
        dist.barrier()
        ok = 1.0
        checker = dist.world - 1                             # the last rank holds the outlet (and the highest stage)
        if dist.rank == checker:
            ok = 1.0 if wl.verify(eng, np.ascontiguousarray(pinned.array[0]), doys[0], rows0) else 0.0
            # (the engine now holds the whole basin; nothing below uses the shard again)
        ok = dist.min(ok)
        parity = {"bit_identical_to_unsharded_run": bool(ok == 1.0), "checked_on_rank": checker,
                  "nodes_compared": int(dist.bcast_float(float(river.n_node), checker)), "window": 0}
        dist.barrier()

    # ================= the same kernels timed alone (synchronous windows, nothing overlapping) ===========================
    iso = {"zone_ms": 0.0, "element_ms": 0.0, "general_ms": 0.0, "river_ms": 0.0}
    n_iso = 3
    if dist.world == 1:
        eng.set_forcing(np.ascontiguousarray(np.concatenate(list(pinned.array[:n_iso]), axis=1)), doys[:n_iso].reshape(-1))
        eng.set_land_slabs(1)      # whole-basin launches: the per-kernel events then bracket ALL elements
        eng.timing()
        for s in range(n_iso):
            eng.select_window(s * win, win)
            eng.set_external(own_ext)
            eng.run(land=True, river=True)
        t_iso = eng.timing()
        iso = {k: t_iso[k] / n_iso for k in iso}

    # ================= roofline of the dominant kernel ====================================================================
    # Discharge-only mode moves 40 B per element-step, so the bound is the FP64 pipe.
    # Work in SURVEY's convention (add/mul/cmp/div/pow/exp = 1 op each): 130 ops per zone-step in the thread-per-zone
    # kernel, 14 R + 2 U + 32 ops per element-step in the thread-per-element kernel — counted as EXECUTED: the device
    # counters of the timed region (hb_get_timing) say how many element-steps ran the RecStep chain (6 ops per sub-step)
    # and how many sub-steps evaluated the power function (8 more ops, one of them the pow); skipped work does not count.
    fp64_peak = eng.measure_fp64_peak()
    peak_ops = fp64_peak / 2.0                                             # T FP64 instructions/s (an FMA = 1 op)
    R, U, Z = int(land.recstep[0]), int(land.uh_ptr[1] - land.uh_ptr[0]), land.n_zone / land.n_elem
    cnt = {k: float(tm[k]) / K for k in ("zone_warpsteps", "zone_warpsteps_dry", "zone_steps", "zone_steps_wet",
                                        "elem_steps", "elem_steps_run", "substeps_pow")}
    executed = {
        "recstep_executed_frac": cnt["elem_steps_run"] / cnt["elem_steps"] if cnt["elem_steps"] else None,
        "recstep_substeps_with_pow_frac": (cnt["substeps_pow"] / (R * cnt["elem_steps"])) if cnt["elem_steps"] and R else None,
        "soil_input_frac": cnt["zone_steps_wet"] / cnt["zone_steps"] if cnt["zone_steps"] else None,
        "dry_shortcut_warpstep_frac": cnt["zone_warpsteps_dry"] / cnt["zone_warpsteps"] if cnt["zone_warpsteps"] else None,
    }
    zone_ops_alg = 130.0 * land.n_zone * win
    element_ops_alg = (14.0 * R + 2.0 * U + 32.0) * land.n_elem * win
    # executed: a zone-step of a dry warp-step leaves out the exp half of its pow (still one op: not subtracted);
    # the element kernel executes 6 ops per sub-step of a running chain and 8 more where uz > 0
    zone_ops = zone_ops_alg
    element_ops = (2.0 * U + 32.0) * land.n_elem * win + 6.0 * R * cnt["elem_steps_run"] + 8.0 * cnt["substeps_pow"]
    # inside the timed region the land kernels run in element slabs on streams of their own and the events bracket the
    # launches of slab 0: its share of the elements (engine.cu run_land: slabs of 128-element blocks, block s*n/slabs)
    n_slabs = args.land_slabs or 2
    nblk = (land.n_elem + 127) // 128
    slab0_share = min(land.n_elem, (nblk // n_slabs) * 128) / land.n_elem if n_slabs > 1 and nblk >= n_slabs else 1.0
    fused = int(tm.get("n_fused", 0)) == land.n_elem              # every group's terms are added in the zone kernel
    land_ms_avg, zone_ms_avg, element_ms_avg = float(np.mean(land_ms)), float(np.mean(zone_ms)), float(np.mean(element_ms))
    general_ms_avg = float(np.mean(general_ms))

    def tops(ops: float, ms: float) -> float:
        return ops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0

    # hardware counters of the same kernels (ncu --set full of one steady-state window of THIS workload), written by
    # tools/ncu_summary.py json into profiles/; absent or for another workload: null
    ncu = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_kernels.json")) as f:
            rec = json.load(f)
        if rec.get("workload") == {"elements": land.n_elem, "zones": land.n_zone, "window": win, "forcing": args.forcing}:
            ncu = rec
    except (OSError, ValueError):
        pass

    def ncu_of(kernel: str, key: str):
        return ncu["kernels"].get(kernel, {}).get(key) if ncu else None

    # algorithmic HBM bytes per launch: zone kernel = the forcing of its elements (32 B per
    # element-step, shared by the zones) + the hand-over; element kernel = the hand-over + 8 B of outflow per element-step.
    # Hand-over: 16 B per element-step when the zone kernel adds the terms itself (HB_OPT_LAND_FUSE), else 16 B per zone-step
    handover = 16 * (land.n_elem if fused else land.n_zone) * win
    kernels = {
        "zone_kernel": {"ms_slab0": zone_ms_avg, "slab0_share_of_elements": slab0_share,
                        "ops_per_launch": zone_ops, "achieved": tops(zone_ops * slab0_share, zone_ms_avg),
                        "frac": tops(zone_ops * slab0_share, zone_ms_avg) / peak_ops,
                        "ms_alone": iso["zone_ms"], "frac_alone": tops(zone_ops, iso["zone_ms"]) / peak_ops,
                        "algorithmic_bytes_per_launch": handover + 32 * land.n_elem * win,
                        "terms_added_in_zone_kernel": bool(fused),
                        "ncu_fp64_pipe_pct": ncu_of("zone_kernel", "fp64_pipe_pct"),
                        "ncu_issue_active_pct": ncu_of("zone_kernel", "issue_active_pct"),
                        "ncu_dram_bytes": ncu_of("zone_kernel", "dram_bytes")},
        "element_kernel": {"ms_slab0": element_ms_avg, "slab0_share_of_elements": slab0_share,
                           "ops_per_launch": element_ops,
                           "ops_per_launch_algorithmic": element_ops_alg,
                           "achieved": tops(element_ops * slab0_share, element_ms_avg),
                           "frac": tops(element_ops * slab0_share, element_ms_avg) / peak_ops,
                           "ms_alone": iso["element_ms"], "frac_alone": tops(element_ops, iso["element_ms"]) / peak_ops,
                           "algorithmic_bytes_per_launch": handover + 8 * land.n_elem * win,
                           "ncu_fp64_pipe_pct": ncu_of("element_kernel", "fp64_pipe_pct"),
                           "ncu_issue_active_pct": ncu_of("element_kernel", "issue_active_pct"),
                           "ncu_dram_bytes": ncu_of("element_kernel", "dram_bytes")},
        "land_kernel (general)": {"ms": general_ms_avg},
        "routing (reach_head + reach_bulk + reach_wave kernels)": {
            "ms": float(np.mean(river_ms)), "ms_alone": iso["river_ms"],
            "algorithmic_bytes_per_step": 16 * river.n_reach * win + 8 * (river.n_node + land.n_elem) * win},
    }
    # dominant kernel = the one with the larger share of the step
    if not n_fast:
        dominant, dom_ms, dom_ops = "land_kernel (general)", general_ms_avg, zone_ops + element_ops
        dom_bytes = 40 * land.n_elem * win
    elif (iso["element_ms"] > iso["zone_ms"]) if (iso["zone_ms"] > 0.0 and iso["element_ms"] > 0.0) else (element_ms_avg > zone_ms_avg):
        # (by the whole-basin launches timed alone where they exist: the events inside the pipeline bracket one slab's
        # kernel together with whatever shares the SMs with it)
        dominant, dom_ms, dom_ops = "element_kernel", element_ms_avg, element_ops * slab0_share
        dom_bytes = kernels["element_kernel"]["algorithmic_bytes_per_launch"] * slab0_share
    else:
        dominant, dom_ms, dom_ops = "zone_kernel", zone_ms_avg, zone_ops * slab0_share
        dom_bytes = kernels["zone_kernel"]["algorithmic_bytes_per_launch"] * slab0_share
    achieved_tops = tops(dom_ops, dom_ms)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    forcing_note = {
        "storms": "synth.make_forcing(kind='storms'): per day and bank column a Gamma(0.3, 8) mm depth in an event of 1-6 "
                  "wet hours, seasonal temperature and evapotranspiration (P 880 mm/a against 580 mm/a potential "
                  "evapotranspiration: the humid balance of the Lahn basin); the timed steps cover one year",
        "dry": "synth.make_forcing(kind='dry'), round 1's forcing: gamma(0.3, 3)/24 mm of rain in every hour, all of it "
               "evaporated from the canopy - the upper zone stays empty and the kernels take their exact dry shortcuts",
    }[args.forcing]
    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": dist.world, "steps": K, "warmup": W,
        "ms_per_step": device_ms_max / K, "higher_is_better": True,
        "scaling": args.scaling if args.config == 3 else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            **wl.desc,
            "elements_per_gpu": land.n_elem, "zones_per_element": Z, "reaches_per_gpu": river.n_reach,
            "river_levels": n_levels, "window_steps": win,
            "simulated_days_timed": K * win / (24.0 if wl.step == "1h" else 1.0),
            "recstep_per_step": R, "uh_ordinates": U,
            "forcing_bank": bank,
            "parallelism": (f"one basin, element-sharded x{dist.world} by hydpy_b200.shard.partition, NCCL boundary "
                            f"series feed the downstream reaches" if dist.world > 1 and args.config == 3 else
                            f"member-sharded x{dist.world}" if dist.world > 1 else "single GPU"),
            "nvlink_bytes_per_step_this_rank": 8 * win * (sum(len(r) for _, r in wl.recv) + sum(len(r) for _, r in wl.send)),
            "l2": "per step the forcing block, the zone-to-element hand-over (16 B per element-step) and the node/reach rows "
                  "(16 B per reach-step) stream through HBM: about 2.5 GB per 438-h step, far beyond the 126 MB L2; no flush "
                  "needed",
            "forcing": forcing_note, **executed,
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s_max / K,
                "same_final_states_as_resident_run": None if same is None else bool(same),
                "node_series_returned": "all nodes" if e2e_nodes is None else "the outlet node of every member"},
        "e2e_gauged_nodes": e2e_gauged, "multi_gpu_parity": parity,
        "per_rank_ms_per_step": {"columns": ["device (timed region / K)", "runoff generation", "routing", "host wall"],
                                 "rows": [[round(x, 3) for x in row] for row in per_rank]},
        "gpu_launches": launches,
        "kernel_ms_per_step": {"runoff_generation": land_ms_avg, "routing": float(np.mean(river_ms)),
                               "wall": wall_ms / K,
                               "note": "the two run on different streams and overlap across windows (hb_run_async)"},
        "roofline": {
            "bound": "fp64", "achieved": achieved_tops, "peak": peak_ops, "unit": "Top/s (FP64 ops, FMA = 1)",
            "frac": achieved_tops / peak_ops,
            "traffic": kernels[dominant].get("ncu_dram_bytes") if dominant in kernels else None,
            "frac_alone": (kernels[dominant]["frac_alone"] if n_fast else None),
            "kernel": dominant, "ops_per_launch": dom_ops,
            "ops": "EXECUTED operations in SURVEY.md section 8d's convention, from the device counters of the timed "
                   "region (config.recstep_executed_frac …); a pow counts as one op but costs about 25 FP64 "
                   "instructions, so ncu's FP64-pipe utilisation (kernels.*.ncu_fp64_pipe_pct) is the tighter figure",
            "peak_source": "hb_measure_fp64_peak (DFMA micro-benchmark, this run); MEASURED_PEAKS.json has no FP64 figure",
            "note": "ms_slab0/frac: CUDA events around the launches of element slab 0 inside the timed, pipelined region "
                    "(the other slab's kernels and the routing of the previous window share the SMs), ops scaled to the "
                    "slab; ms_alone/frac_alone: whole-basin launches in synchronous windows right after it.  The RecStep "
                    "chain is sequential per element: 10^5 elements are 5.3 warps per scheduler, each a dependent chain, "
                    "which bounds the element kernel near 0.7 instructions per cycle and scheduler (DESIGN.md section 4)",
            "fp64_tflops_measured": fp64_peak,
            "ncu_source": ncu.get("source") if ncu else None,
            "hbm": {"achieved": dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s",
                    "frac": (dom_bytes / (dom_ms * 1e-3) / 1e9 / hbm_peak) if dom_ms > 0 else 0.0,
                    "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
            "kernels": kernels,
        },
        "clocks": clocks, "device": info, "checksum": checksum,
    }
    pinned.free()
    for o in outs:
        o.free()
    eng.close()
    return result


# ---------------------------------------------------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------------------------------------------------
def cpu_run(args: argparse.Namespace, steps: int, warmup: int, elements: int) -> dict:
    """The reference's CPU implementation on a bounded sample of the same workload: `elements` land elements of shard 0
    (+ their river sub-tree), `window` hours per step, all host threads."""
    from hydpy_b200 import synth
    from oracle import oracle, refdriver

    cores = os.cpu_count() or 1
    win = args.window
    daily = args.config == 2                     # configs[2]: daily steps, runoff generation only
    step = "1d" if daily else "1h"
    if daily and win == 438:
        win = 365
    land, states = synth.make_land(elements, step, seed=1)
    if daily:
        from hydpy_b200.abi import RiverBundle

        river = RiverBundle.empty(elements)
        river.entry_ptr = np.arange(elements + 1, dtype=np.int64)
        river.entry_src = np.arange(elements, dtype=np.int32)
        river.normalise()
        discharge = np.zeros(0)
    else:
        river, discharge = synth.make_tree_river(elements, step, seed=2,
                                                 levels=max(2, args.levels * elements // args.elements))
    kind = "reference" if refdriver.available() else "port"
    zone_steps = land.n_zone * win
    times = []
    for s in range(warmup + steps):
        forcing, doy = synth.make_forcing(win, step, seed=0, t0=s * win, kind=args.forcing)
        if kind == "reference":
            if s == 0:
                els = refdriver.prepare_land(land, states, forcing, doy)   # object set-up (once) is not timed
            else:
                for el in els:                                             # next window: the conditions stay in the cymodels
                    el.set_forcing(land, forcing, doy)
            t0 = time.perf_counter()
            qt = refdriver.run_land(els, threads=cores)
            t_land = time.perf_counter() - t0
        else:
            t0 = time.perf_counter()
            qt, _ = oracle.land_run(land, states, forcing, doy, threads=cores)
            t_land = time.perf_counter() - t0
        t0 = time.perf_counter()
        oracle.river_run(river, discharge, qt, None)                       # routing: C restatement (≈1 % of the time)
        t_river = time.perf_counter() - t0
        if s >= warmup:
            times.append(t_land + t_river)
    total = float(np.sum(times))
    value = zone_steps * steps / total
    return {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{elements} of {args.elements} elements ({land.n_zone} zones) x {win} {'daily' if daily else 'hourly'} steps per step, "
                      f"{steps} steps, runoff generation on {cores} threads"
                      f" ({'unmodified reference binaries oracle/_ref' if kind == 'reference' else 'C restatement'})"
                      f" + routing", "ms_per_step": 1e3 * total / steps}


def auto_cpu_elements(args: argparse.Namespace) -> int:
    if args.cpu_elements:
        return args.cpu_elements
    cores = os.cpu_count() or 1
    # ≈3 µs per element-step and thread: aim at ≈2 s of wall time per step
    # (daily steps: 1200 instead of 50 RecStep sub-steps per step, about 30 µs per element-step)
    n = int(2.0 * cores / ((30e-6 if args.config == 2 else 3e-6) * (365 if args.config == 2 and args.window == 438 else args.window)))
    return int(max(256, min(args.elements, n // 64 * 64)))


def main() -> None:
    # stdout carries the ONE JSON line only: whatever native libraries print (NCCL's version banner, NCCL_DEBUG output)
    # goes to stderr — file descriptor 1 points to stderr until the line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict) -> None:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    args = parse_args()
    dist = Dist()
    if args.gpus != dist.world and dist.world > 1:
        args.gpus = dist.world
    try:
        if args.impl == "reference":
            if dist.rank == 0:
                base = cpu_run(args, args.steps, args.warmup, auto_cpu_elements(args))
                line = {
                    "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                    "data": "synthetic",
                    "config": {"workload": ("configs[2]: synthetic 10^5-element hland_96 basin, daily steps, runoff generation "
                                            "only (bounded element sample on the host CPU)" if args.config == 2 else
                                            "configs[3]: synthetic 10^5-element hland_96 + musk_classic river tree, "
                                            "hourly, discharge-only (bounded element sample on the host CPU)"),
                               "window_steps": 365 if args.config == 2 and args.window == 438 else args.window},
                    "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                    "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0,
                }
                emit(line)
            return
        result = run_engine(args, dist)
        if dist.rank == 0:
            if dist.world == 1 and not args.no_cpu_baseline:
                base = cpu_run(args, steps=2, warmup=1, elements=auto_cpu_elements(args))
                result["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
            emit(result)
    finally:
        dist.close()


if __name__ == "__main__":
    main()
