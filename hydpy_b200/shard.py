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


def _gather(ptr_: np.ndarray, items: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """CSR rows `items` of a CSR structure: (new row pointer, flat indices into the old value arrays)."""
    items = np.asarray(items, np.int64)
    lens = (ptr_[1:] - ptr_[:-1])[items].astype(np.int64)
    new_ptr = np.zeros(len(items) + 1, np.int64)
    np.cumsum(lens, out=new_ptr[1:])
    total = int(new_ptr[-1])
    if total == 0:
        return new_ptr, np.zeros(0, np.int64)
    idx = np.arange(total, dtype=np.int64) - np.repeat(new_ptr[:-1], lens) + np.repeat(ptr_[:-1][items].astype(np.int64), lens)
    return new_ptr, idx


def subset_land(land: LandBundle, states: LandStates, elements: np.ndarray, outlet_node: np.ndarray
                ) -> tuple[LandBundle, LandStates]:
    """The elements `elements` (global indices) as a bundle of their own; `outlet_node` = their local outlet nodes."""
    elements = np.asarray(elements, np.int64)
    land.normalise()
    states.normalise()
    zp, zi = _gather(land.zone_ptr, elements)
    sp, si = _gather(land.snow_ptr, elements)
    fp, fi = _gather(land.sfdist_ptr, elements)
    up, ui = _gather(land.uh_ptr, elements)
    rp, ri = _gather(land.sred_ptr, elements)
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
        _, idx = _gather(ptr_, shard.elements)
        getattr(out, which)[idx] = getattr(local, which)
    out.uz[shard.elements] = local.uz
    out.lz[shard.elements] = local.lz
    out.sg2[shard.elements] = local.sg2
    out.sg3[shard.elements] = local.sg3


def assign_nodes(land: LandBundle | None, river: RiverBundle, world: int, *, outlet_node: np.ndarray | None = None,
                 zones: np.ndarray | None = None, method: str = "subtrees", granularity: int = 8) -> np.ndarray:
    """Rank of every node: `world` parts of equal weight (zones of the land elements + a share for the reach segments).

    Nodes that feed the same reach are kept together (a reach reads all its inlet nodes locally; a reach lives with its
    inlet nodes).  Two methods, both on an UPSTREAM-FIRST ordering of the network (the post-order of a depth-first walk
    from the outlets towards the headwaters: sub-networks are contiguous in it):

    * "subtrees" (river trees; the default): the largest sub-basins are split at their outlet node until none is heavier
      than 1/(granularity * world) of the basin; the split nodes form the *trunk*, the remaining complete sub-basins are
      packed into `world` bins of equal weight (largest first), the trunk joins the last one.  Ranks 0 … world-2 then own
      complete sub-basins only — they never wait for anybody — and the last rank routes the trunk: the shard graph is
      one stage deep, so the skew between ranks is ONE routing pass whatever the number of GPUs;
    * "ranges": the order is cut into `world` ranges of equal weight; boundary series only travel from lower to higher
      ranks (acyclic for any DAG), but the stages form a chain.  Fallback for networks with bifurcations."""
    nn, nr = river.n_node, river.n_reach
    if outlet_node is None:
        outlet_node = land.outlet_node if land is not None else np.zeros(0, np.int32)
    if zones is None:
        zones = np.diff(land.zone_ptr).astype(float) if land is not None else np.ones(len(outlet_node))
    # groups of nodes that feed the same reach (union-find over the reaches with several inlet nodes only)
    parent = np.arange(nn)

    def find(a: int) -> int:
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    n_in = np.diff(river.inlet_ptr)
    for q in np.flatnonzero(n_in > 1):
        ins = river.inlet_node[river.inlet_ptr[q]:river.inlet_ptr[q + 1]]
        for n in ins[1:]:
            ra, rb = find(int(ins[0])), find(int(n))
            if ra != rb:
                parent[rb] = ra
    if np.any(n_in > 1):
        group = np.array([find(n) for n in range(nn)])
    else:
        group = parent
    weight = np.zeros(nn)
    ok = np.asarray(outlet_node) >= 0
    np.add.at(weight, group[np.asarray(outlet_node)[ok]], np.asarray(zones, float)[ok])
    has_in = n_in > 0
    first_in = np.zeros(nr, np.int64)
    first_in[has_in] = river.inlet_node[river.inlet_ptr[:-1][has_in]]
    g_of_reach = group[first_in]
    np.add.at(weight, g_of_reach[has_in], 0.05 * np.diff(river.seg_ptr).astype(float)[has_in])
    # upstream adjacency of the groups: edge g -> d for every reach of g that ends in a node of another group d
    out_ok = has_in & (river.reach_outlet >= 0)
    src_g = g_of_reach[out_ok]
    dst_g = group[river.reach_outlet[out_ok]]
    keep = src_g != dst_g
    src_g, dst_g = src_g[keep], dst_g[keep]
    if len(src_g):                                           # (parallel reaches between two groups: one edge)
        pair = np.unique(np.stack([dst_g, src_g], axis=1), axis=0)
        dst_g, src_g = pair[:, 0], pair[:, 1]
    up_of = src_g                                            # upstream groups, sorted by downstream group
    up_ptr = np.zeros(nn + 1, np.int64)
    np.add.at(up_ptr, dst_g + 1, 1)
    np.cumsum(up_ptr, out=up_ptr)
    is_group = np.zeros(nn, bool)
    is_group[group] = True
    n_down = np.zeros(nn, np.int64)
    np.add.at(n_down, src_g, 1)
    roots = np.flatnonzero(is_group & (n_down == 0))
    visited = np.zeros(nn, bool)
    post: list[int] = []
    up_ptr_l, up_of_l = up_ptr.tolist(), up_of.tolist()
    for root in roots.tolist():
        if visited[root]:
            continue
        visited[root] = True
        stack = [(root, up_ptr_l[root])]
        while stack:
            u, i = stack[-1]
            if i < up_ptr_l[u + 1]:
                stack[-1] = (u, i + 1)
                c = up_of_l[i]
                if not visited[c]:
                    visited[c] = True
                    stack.append((c, up_ptr_l[c]))
            else:
                post.append(u)
                stack.pop()
    post_arr = np.array(post, np.int64)
    if len(post_arr) != int(is_group.sum()):                 # groups on a cycle (refused later by the engine): append them
        rest = np.flatnonzero(is_group & ~visited)
        post_arr = np.concatenate([post_arr, rest])
    w = weight[post_arr]
    total = float(w.sum())
    if total <= 0.0:
        w = np.ones(len(post_arr))
        total = float(len(post_arr))
    rank = np.zeros(nn, np.int64)
    is_tree = bool(np.all(n_down[post_arr] <= 1)) and len(post_arr) == len(post)
    if method == "subtrees" and is_tree and world > 1:
        import heapq

        # sub-basin of a group = a contiguous stretch of the post-order that ends at the group
        pos = np.zeros(nn, np.int64)
        pos[post_arr] = np.arange(len(post_arr))
        down = np.full(nn, -1, np.int64)
        down[src_g] = dst_g
        acc = np.zeros(nn)
        acc[post_arr] = w
        size = np.zeros(nn, np.int64)
        size[post_arr] = 1
        acc_l, size_l, down_l = acc.tolist(), size.tolist(), down.tolist()
        for g in post_arr.tolist():                          # upstream first: every group hands its sum downstream
            d = down_l[g]
            if d >= 0:
                acc_l[d] += acc_l[g]
                size_l[d] += size_l[g]
        limit = total / (max(granularity, 1) * world)
        heap = [(-acc_l[r], r) for r in roots.tolist()]
        heapq.heapify(heap)
        trunk: list[int] = []
        parts: list[tuple[float, int]] = []
        while heap:
            neg, g = heapq.heappop(heap)
            if -neg <= limit or up_ptr_l[g] == up_ptr_l[g + 1]:
                parts.append((-neg, g))
                continue
            trunk.append(g)
            for i in range(up_ptr_l[g], up_ptr_l[g + 1]):
                c = up_of_l[i]
                heapq.heappush(heap, (-acc_l[c], c))
        bins = [0.0] * world
        bins[world - 1] = float(sum(weight[g] for g in trunk))
        rank_of_group = np.full(nn, world - 1, np.int64)     # the trunk stays with the last rank
        for wt, g in sorted(parts, reverse=True):
            b = min(range(world), key=lambda i: (bins[i], -i))
            bins[b] += wt
            rank_of_group[post_arr[pos[g] - size_l[g] + 1:pos[g] + 1]] = b
        return rank_of_group[group]
    mid = np.cumsum(w) - 0.5 * w
    rank_of_group = np.minimum((mid / (total / max(world, 1))).astype(np.int64), world - 1)
    rank[post_arr] = rank_of_group
    return rank[group]


def partition(land: LandBundle | None, states: LandStates | None, river: RiverBundle, discharge: np.ndarray, world: int,
              node_rank: np.ndarray | None = None, mct_states: np.ndarray | None = None,
              ranks: Sequence[int] | None = None, outlet_node: np.ndarray | None = None) -> list[Shard]:
    """Cut the basin into `world` shards (see module docstring).  `node_rank` overrides the built-in assignment.

    `ranks`: the shards to build completely (default: all); the others only carry what the exchange plan needs
    (`nodes`, `reaches`, `elements`, `recv`, `send`, `stage`) — on a large basin every rank builds its own bundles only.
    `land`/`states` may be None when the caller brings the land elements of its shard itself (`outlet_node` then says
    which node each global element drains to)."""
    if land is not None:
        land.normalise()
    river.normalise()
    nn, nr = river.n_node, river.n_reach
    if outlet_node is None:
        outlet_node = land.outlet_node if land is not None else np.zeros(0, np.int32)
    outlet_node = np.asarray(outlet_node, np.int64)
    node_rank = (assign_nodes(land, river, world, outlet_node=outlet_node) if node_rank is None
                 else np.asarray(node_rank)).astype(np.int64)
    n_in = np.diff(river.inlet_ptr)
    has_in = n_in > 0
    first_in = np.zeros(nr, np.int64)
    first_in[has_in] = river.inlet_node[river.inlet_ptr[:-1][has_in]]
    reach_rank = np.where(has_in, node_rank[first_in],
                          np.where(river.reach_outlet >= 0, node_rank[np.maximum(river.reach_outlet, 0)], 0)).astype(np.int64)
    if len(river.inlet_node):
        inlet_reach = np.repeat(np.arange(nr), n_in)
        bad = node_rank[river.inlet_node] != reach_rank[inlet_reach]
        if np.any(bad):
            raise ValueError(f"reach {int(inlet_reach[np.argmax(bad)])} reads inlet nodes of different ranks")
    elem_rank = np.where(outlet_node >= 0, node_rank[np.maximum(outlet_node, 0)], 0)
    build = set(range(world)) if ranks is None else {int(r) for r in ranks}
    # every boundary edge of the network at once: entries of reaches that live on another rank than the node
    entry_node = np.repeat(np.arange(nn), np.diff(river.entry_ptr))
    esrc = river.entry_src.astype(np.int64)
    is_reach_entry = (esrc < 0) & (esrc > layout.ENTRY_EXT)
    eq = -1 - esrc[is_reach_entry]
    consumer = node_rank[entry_node[is_reach_entry]]
    producer = reach_rank[eq]
    cross = consumer != producer
    b_cons, b_prod, b_reach = consumer[cross], producer[cross], eq[cross]
    order = np.lexsort((b_reach, b_prod, b_cons))
    b_cons, b_prod, b_reach = b_cons[order], b_prod[order], b_reach[order]
    if len(b_reach) > 1:   # (a reach enters one node: duplicates would mean a malformed network)
        dup = (b_cons[1:] == b_cons[:-1]) & (b_reach[1:] == b_reach[:-1])
        keep = np.concatenate([[True], ~dup])
        b_cons, b_prod, b_reach = b_cons[keep], b_prod[keep], b_reach[keep]
    shards: list[Shard] = []
    for r in range(world):
        nodes = np.flatnonzero(node_rank == r)
        reaches = np.flatnonzero(reach_rank == r)
        elements = np.flatnonzero(elem_rank == r)
        mine = b_cons == r
        boundary: dict[int, list[int]] = {}
        for peer in np.unique(b_prod[mine]).tolist():
            boundary[int(peer)] = b_reach[mine & (b_prod == peer)].tolist()
        # global external rows used here come first (ascending), then the boundary rows, peer by peer
        eptr, eidx = _gather(river.entry_ptr, nodes)
        src = esrc[eidx]
        used_ext = np.concatenate([river.node_ext[nodes][river.node_ext[nodes] >= 0].astype(np.int64),
                                   layout.ENTRY_EXT - src[src <= layout.ENTRY_EXT]])
        ext_global = np.unique(used_ext)
        n_own_ext = len(ext_global)
        recv: list[tuple[int, list[int]]] = []
        k = n_own_ext
        brow_q: list[int] = []
        for peer in sorted(boundary):
            rows = list(range(k, k + len(boundary[peer])))
            brow_q.extend(boundary[peer])
            k += len(rows)
            recv.append((peer, rows))
        sh = Shard(rank=r, stage=0, elements=elements, nodes=nodes, reaches=reaches, land=None, states=None, river=None,
                   discharge=None, ext_global=ext_global, recv=recv, send=[], recv_global=boundary, mct_states=None)
        shards.append(sh)
        if r not in build:
            continue
        node_l = np.full(nn, -1, np.int64)
        node_l[nodes] = np.arange(len(nodes))
        reach_l = np.full(nr, -1, np.int64)
        reach_l[reaches] = np.arange(len(reaches))
        elem_l = np.full(len(outlet_node), -1, np.int64)
        elem_l[elements] = np.arange(len(elements))
        new_src = np.zeros(len(src), np.int64)
        m_land = src >= 0
        new_src[m_land] = elem_l[src[m_land]]
        m_ext = src <= layout.ENTRY_EXT
        new_src[m_ext] = layout.ENTRY_EXT - np.searchsorted(ext_global, layout.ENTRY_EXT - src[m_ext])
        m_reach = ~m_land & ~m_ext
        q = -1 - src[m_reach]
        local = reach_rank[q] == r
        val = np.zeros(len(q), np.int64)
        val[local] = -1 - reach_l[q[local]]
        if np.any(~local):
            brow_arr = np.array(brow_q, np.int64)
            sorter = np.argsort(brow_arr)
            pos = sorter[np.searchsorted(brow_arr, q[~local], sorter=sorter)]
            val[~local] = layout.ENTRY_EXT - (n_own_ext + pos)
        new_src[m_reach] = val
        iptr, iidx = _gather(river.inlet_ptr, reaches)
        sptr, sidx = _gather(river.seg_ptr, reaches)
        tptr, tidx = _gather(river.trap_ptr, reaches)
        ro = river.reach_outlet[reaches].astype(np.int64)
        outlet = np.where((ro >= 0) & (node_rank[np.maximum(ro, 0)] == r), node_l[np.maximum(ro, 0)], -1).astype(np.int32)
        node_ext_g = river.node_ext[nodes].astype(np.int64)
        node_ext_l = np.where(node_ext_g >= 0, np.searchsorted(ext_global, np.maximum(node_ext_g, 0)), -1).astype(np.int32)
        sub_river = RiverBundle(
            entry_ptr=eptr, entry_src=new_src.astype(np.int32), node_mode=river.node_mode[nodes], node_ext=node_ext_l,
            inlet_ptr=iptr, inlet_node=node_l[river.inlet_node[iidx]].astype(np.int32), reach_outlet=outlet,
            seg_ptr=sptr, coef=river.coef[:, reaches], n_ext=k,
            reach_model=river.reach_model[reaches], nmbruns=river.nmbruns[reaches], rpar=river.rpar[:, reaches],
            trap_ptr=tptr, tpar=river.tpar[:, tidx])
        sub_river.normalise()
        sh.river = sub_river
        sh.discharge = np.ascontiguousarray(discharge[sidx])
        sh.mct_states = np.ascontiguousarray(mct_states[:, sidx]) if mct_states is not None else None
        if land is not None:
            local_outlet = np.where(outlet_node[elements] >= 0, node_l[np.maximum(outlet_node[elements], 0)], -1)
            sh.land, sh.states = subset_land(land, states, elements, local_outlet.astype(np.int32))
    # sends mirror the receives; stages follow the shard DAG
    for s in shards:
        for peer, _rows in s.recv:
            local_q = np.searchsorted(shards[peer].reaches, np.array(s.recv_global[peer], np.int64))
            shards[peer].send.append((s.rank, local_q.tolist()))
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
