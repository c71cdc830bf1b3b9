"""Element sharding of one basin over several GPUs.

Runoff generation has no cross-element dependency, so land elements simply follow the node they drain to.  The river
DAG is cut into `world` parts along reach→node edges: where the outlet node of a reach lives on another rank, the
producer sends the reach's outflow series and the consumer sees it as an *external row* that takes the reach's place
in the node's entry list — same position, same addition order, hence bit-identical node sums
(ref: hydpy/core/threadingtools.py:396-446 for the order of additions).  Nothing else crosses ranks.

The reference has no counterpart (single process); the correctness oracle for a sharded run is the unsharded run.
`Shard.send`/`Shard.recv` are the exchange plan one window needs; `run_window` executes it on an `Engine` (NCCL over
NVLink, rows travel device to device), tests/test_shard.py executes the same plan with the CPU oracle over gloo.
"""
from __future__ import annotations

import dataclasses
from typing import Sequence

import numpy as np

from . import layout
from .abi import LandBundle, LandStates, RiverBundle


@dataclasses.dataclass
class Shard:
    rank: int
    stage: int                     # 0 = no upstream shard; a shard runs its routing after all shards it receives from
    elements: np.ndarray           # global indices of the land elements (local order)
    nodes: np.ndarray              # global node indices (local order)
    reaches: np.ndarray            # global reach indices (local order)
    land: LandBundle
    states: LandStates
    river: RiverBundle
    discharge: np.ndarray
    ext_global: np.ndarray         # local external rows 0..len-1 are these rows of the global external block …
    recv: list[tuple[int, list[int]]]   # … followed by boundary rows: (peer rank, local external rows), peer-sorted
    send: list[tuple[int, list[int]]]   # (peer rank, local reach indices), peer-sorted; order = peer's recv order
    recv_global: dict[int, list[int]] = dataclasses.field(default_factory=dict)  # peer -> global reaches, row order
    mct_states: np.ndarray | None = None   # [3, n_seg] conditions of the shard's musk_mct reaches

    def local_ext(self, ext: np.ndarray, nt: int) -> np.ndarray:
        """The rows of the global external block this shard reads, padded with zero rows for the boundary series."""
        out = np.zeros((self.river.n_ext, nt))
        if len(self.ext_global):
            out[:len(self.ext_global)] = ext[self.ext_global]
        return out


def subset_land(land: LandBundle, states: LandStates, elements: np.ndarray, outlet_node: np.ndarray
                ) -> tuple[LandBundle, LandStates]:
    """The elements `elements` (global indices) as a bundle of their own; `outlet_node` = their local outlet nodes."""
    elements = np.asarray(elements, np.int64)

    def csr(ptr_: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        lens = (ptr_[1:] - ptr_[:-1])[elements]
        new_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        idx = np.concatenate([np.arange(ptr_[e], ptr_[e + 1]) for e in elements] + [np.zeros(0, np.int64)]).astype(np.int64)
        return new_ptr, idx

    land.normalise()
    states.normalise()
    zp, zi = csr(land.zone_ptr)
    sp, si = csr(land.snow_ptr)
    fp, fi = csr(land.sfdist_ptr)
    up, ui = csr(land.uh_ptr)
    rp, ri = csr(land.sred_ptr)
    sub = LandBundle(
        zone_ptr=zp, snow_ptr=sp, sfdist_ptr=fp, uh_ptr=up, sred_ptr=rp, recstep=land.recstep[elements],
        eflags=land.eflags[elements], forcing_col=land.forcing_col[elements], outlet_node=outlet_node,
        epar=land.epar[:, elements], zonetype=land.zonetype[zi], zflags=land.zflags[zi], zorder=land.zorder[zi],
        zpar=land.zpar[:, zi], sfdist=land.sfdist[fi], uh=land.uh[ui], sred_from=land.sred_from[ri],
        sred_to=land.sred_to[ri], sred_adjust=land.sred_adjust[ri], model=land.model[elements],
        rconc=land.rconc[elements], nash_steps=land.nash_steps[elements])
    sub.normalise()
    st = LandStates(ic=states.ic[zi], sm=states.sm[zi], sp=states.sp[si], wc=states.wc[si], uz=states.uz[elements],
                    lz=states.lz[elements], quh=states.quh[ui], suz=states.suz[zi], sg1=states.sg1[zi],
                    bw1=states.bw1[zi], bw2=states.bw2[zi], sg2=states.sg2[elements], sg3=states.sg3[elements])
    return sub, st


def scatter_states(shard: Shard, local: LandStates, land: LandBundle, out: LandStates) -> None:
    """Write a shard's final states back into the global state arrays."""
    local.normalise()
    out.normalise()
    for which, ptr_ in (("ic", land.zone_ptr), ("sm", land.zone_ptr), ("sp", land.snow_ptr), ("wc", land.snow_ptr),
                        ("quh", land.uh_ptr), ("suz", land.zone_ptr), ("sg1", land.zone_ptr), ("bw1", land.zone_ptr),
                        ("bw2", land.zone_ptr)):
        src = getattr(local, which)
        dst = getattr(out, which)
        pos = 0
        for e in shard.elements:
            n = int(ptr_[e + 1] - ptr_[e])
            dst[ptr_[e]:ptr_[e] + n] = src[pos:pos + n]
            pos += n
    out.uz[shard.elements] = local.uz
    out.lz[shard.elements] = local.lz
    out.sg2[shard.elements] = local.sg2
    out.sg3[shard.elements] = local.sg3


def assign_nodes(land: LandBundle, river: RiverBundle, world: int) -> np.ndarray:
    """Rank of every node: contiguous upstream sub-networks of roughly equal zone count.

    Nodes that feed the same reach are kept together (a reach reads all its inlet nodes locally); a reach lives with
    its inlet nodes.  Greedy bottom-up accumulation: walking the DAG upstream-first, a sub-network is split off as soon
    as it has collected total/world zones."""
    nn, nr = river.n_node, river.n_reach
    parent = np.arange(nn)

    def find(a: int) -> int:
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for q in range(nr):
        ins = river.inlet_node[river.inlet_ptr[q]:river.inlet_ptr[q + 1]]
        for n in ins[1:]:
            ra, rb = find(int(ins[0])), find(int(n))
            if ra != rb:
                parent[rb] = ra
    group = np.array([find(n) for n in range(nn)])
    weight = np.zeros(nn)
    zones = np.diff(land.zone_ptr).astype(float)
    for e, n in enumerate(land.outlet_node):
        if n >= 0:
            weight[group[n]] += zones[e]
    segs = np.diff(river.seg_ptr).astype(float)
    down = np.full(nn, -1)                       # downstream group of a group (first reach with an outlet elsewhere)
    children: list[list[int]] = [[] for _ in range(nn)]
    indeg = np.zeros(nn, int)
    for q in range(nr):
        if river.inlet_ptr[q + 1] == river.inlet_ptr[q]:
            continue
        g = group[river.inlet_node[river.inlet_ptr[q]]]
        weight[g] += 0.05 * segs[q]
        o = river.reach_outlet[q]
        if o >= 0 and group[o] != g and down[g] < 0:
            down[g] = group[o]
            children[group[o]].append(g)
            indeg[group[o]] += 1
    roots = [g for g in np.unique(group)]
    rank = np.full(nn, -1)
    target = weight[np.unique(group)].sum() / max(world, 1)
    acc = weight.copy()
    order = [g for g in roots if indeg[g] == 0]
    next_rank = 0
    i = 0
    while i < len(order):
        g = order[i]
        i += 1
        if next_rank < world - 1 and acc[g] >= target:
            stack = [g]
            while stack:                          # the not yet assigned part of the sub-network above g
                u = stack.pop()
                if rank[u] >= 0:
                    continue
                rank[u] = next_rank
                stack.extend(c for c in children[u] if rank[c] < 0)
            next_rank += 1
            acc[g] = 0.0
        d = down[g]
        if d >= 0:
            acc[d] += acc[g]
            indeg[d] -= 1
            if indeg[d] == 0:
                order.append(d)
    rank[rank < 0] = next_rank                    # the remainder (holds the outlets)
    return rank[group]


def partition(land: LandBundle, states: LandStates, river: RiverBundle, discharge: np.ndarray, world: int,
              node_rank: np.ndarray | None = None, mct_states: np.ndarray | None = None) -> list[Shard]:
    """Cut the basin into `world` shards (see module docstring).  `node_rank` overrides the built-in assignment."""
    land.normalise()
    river.normalise()
    nn, nr = river.n_node, river.n_reach
    node_rank = assign_nodes(land, river, world) if node_rank is None else np.asarray(node_rank)
    reach_rank = np.zeros(nr, int)
    for q in range(nr):
        ins = river.inlet_node[river.inlet_ptr[q]:river.inlet_ptr[q + 1]]
        reach_rank[q] = node_rank[ins[0]] if len(ins) else (node_rank[river.reach_outlet[q]] if river.reach_outlet[q] >= 0 else 0)
        if len(ins) and np.any(node_rank[ins] != reach_rank[q]):
            raise ValueError(f"reach {q} reads inlet nodes of different ranks")
    elem_rank = np.where(land.outlet_node >= 0, node_rank[np.maximum(land.outlet_node, 0)], 0)
    shards: list[Shard] = []
    # boundary edges, in the order the consumer meets them: (consumer, producer) -> [global reach]
    for r in range(world):
        nodes = np.flatnonzero(node_rank == r)
        reaches = np.flatnonzero(reach_rank == r)
        elements = np.flatnonzero(elem_rank == r)
        node_l = {int(n): i for i, n in enumerate(nodes)}
        reach_l = {int(q): i for i, q in enumerate(reaches)}
        elem_l = {int(e): i for i, e in enumerate(elements)}
        ext_global: list[int] = []
        ext_l: dict[int, int] = {}

        def ext_row(g: int) -> int:
            if g not in ext_l:
                ext_l[g] = len(ext_global)
                ext_global.append(g)
            return ext_l[g]

        # first pass: global external rows used here (their local numbers come first)
        for n in nodes:
            if river.node_ext[n] >= 0:
                ext_row(int(river.node_ext[n]))
            for s in river.entry_src[river.entry_ptr[n]:river.entry_ptr[n + 1]]:
                if s <= layout.ENTRY_EXT:
                    ext_row(int(layout.ENTRY_EXT - s))
        boundary: dict[int, list[int]] = {}       # producer rank -> global reaches, in order of first use
        brow: dict[int, int] = {}
        for n in nodes:
            for s in river.entry_src[river.entry_ptr[n]:river.entry_ptr[n + 1]]:
                if layout.ENTRY_EXT < s < 0 and reach_rank[-1 - s] != r:
                    boundary.setdefault(int(reach_rank[-1 - s]), []).append(int(-1 - s))
        n_own_ext = len(ext_global)
        recv: list[tuple[int, list[int]]] = []
        k = n_own_ext
        for peer in sorted(boundary):
            rows = []
            for q in boundary[peer]:
                brow[q] = k
                rows.append(k)
                k += 1
            recv.append((peer, rows))
        entry_ptr = [0]
        entry_src: list[int] = []
        for n in nodes:
            for s in river.entry_src[river.entry_ptr[n]:river.entry_ptr[n + 1]]:
                s = int(s)
                if s >= 0:
                    entry_src.append(elem_l[s])
                elif s <= layout.ENTRY_EXT:
                    entry_src.append(layout.ENTRY_EXT - ext_l[layout.ENTRY_EXT - s])
                elif reach_rank[-1 - s] == r:
                    entry_src.append(-1 - reach_l[-1 - s])
                else:
                    entry_src.append(layout.ENTRY_EXT - brow[-1 - s])
            entry_ptr.append(len(entry_src))
        inlet_ptr = [0]
        inlet_node: list[int] = []
        seg_ptr = [0]
        dis_parts = []
        mct_parts: list[np.ndarray] = []
        trap_ptr = [0]
        trap_idx: list[int] = []
        for q in reaches:
            inlet_node.extend(node_l[int(n)] for n in river.inlet_node[river.inlet_ptr[q]:river.inlet_ptr[q + 1]])
            inlet_ptr.append(len(inlet_node))
            seg_ptr.append(seg_ptr[-1] + int(river.seg_ptr[q + 1] - river.seg_ptr[q]))
            dis_parts.append(discharge[river.seg_ptr[q]:river.seg_ptr[q + 1]])
            if mct_states is not None:
                mct_parts.append(mct_states[:, river.seg_ptr[q]:river.seg_ptr[q + 1]])
            trap_ptr.append(trap_ptr[-1] + int(river.trap_ptr[q + 1] - river.trap_ptr[q]))
            trap_idx.extend(range(int(river.trap_ptr[q]), int(river.trap_ptr[q + 1])))
        outlet = np.array([node_l.get(int(river.reach_outlet[q]), -1) if reach_rank[q] == r and
                           river.reach_outlet[q] >= 0 and node_rank[river.reach_outlet[q]] == r else -1
                           for q in reaches], np.int32)
        sub_river = RiverBundle(
            entry_ptr=np.array(entry_ptr, np.int64), entry_src=np.array(entry_src, np.int32),
            node_mode=river.node_mode[nodes], node_ext=np.array(
                [ext_l[int(x)] if x >= 0 else -1 for x in river.node_ext[nodes]], np.int32),
            inlet_ptr=np.array(inlet_ptr, np.int64), inlet_node=np.array(inlet_node, np.int32), reach_outlet=outlet,
            seg_ptr=np.array(seg_ptr, np.int64), coef=river.coef[:, reaches], n_ext=k,
            reach_model=river.reach_model[reaches], nmbruns=river.nmbruns[reaches], rpar=river.rpar[:, reaches],
            trap_ptr=np.array(trap_ptr, np.int64), tpar=river.tpar[:, np.array(trap_idx, np.int64)])
        sub_river.normalise()
        sub_land, sub_states = subset_land(
            land, states, elements,
            np.array([node_l.get(int(land.outlet_node[e]), -1) for e in elements], np.int32))
        shards.append(Shard(
            rank=r, stage=0, elements=elements, nodes=nodes, reaches=reaches, land=sub_land, states=sub_states,
            river=sub_river, discharge=np.concatenate(dis_parts + [np.zeros(0)]),
            ext_global=np.array(ext_global, np.int64), recv=recv, send=[], recv_global=boundary,
            mct_states=(np.concatenate(mct_parts + [np.zeros((3, 0))], axis=1) if mct_states is not None else None)))
    # sends mirror the receives; stages follow the shard DAG
    for s in shards:
        for peer, _rows in s.recv:
            reach_l = {int(q): i for i, q in enumerate(shards[peer].reaches)}
            shards[peer].send.append((s.rank, [reach_l[q] for q in s.recv_global[peer]]))
    for s in shards:
        s.send.sort(key=lambda x: x[0])
    changed = True
    while changed:
        changed = False
        for s in shards:
            st = max([shards[p].stage + 1 for p, _ in s.recv], default=0)
            if st > s.stage:
                if st > world:
                    raise ValueError("the shard graph is cyclic; choose another node assignment")
                s.stage = st
                changed = True
    return shards


def run_window(engine, shard: Shard, ext: np.ndarray | None = None) -> None:
    """One window on one rank: runoff generation, receive the upstream boundary series (NCCL, device to device),
    routing, send the own boundary series downstream.  The forcing of the window must already be staged
    (`engine.set_forcing` / `select_window`); `ext` = this shard's rows of the global external block (or None)."""
    nt = engine.window_steps
    rows_ext = shard.local_ext(ext, nt) if ext is not None and ext.size else np.zeros((shard.river.n_ext, nt))
    if not shard.recv:
        engine.set_external(rows_ext)
        engine.run(land=True, river=True)
    else:
        engine.run(land=True, river=False)
        engine.set_external(rows_ext)
        for peer, rows in shard.recv:
            engine.comm_recv_ext(peer, rows)
        engine.run(land=False, river=True)
    for peer, reaches in shard.send:
        engine.comm_send_rows(peer, reaches, kind="reach")
