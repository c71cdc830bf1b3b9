"""Host-side multi-GPU logic on CPU: the sharding plan (`hydpy_b200/shard.py`) executed with the CPU oracle must give
the unsharded result bit for bit — once in a single process (stage order), once as a real world-size-2 job over
`torch.distributed` gloo, which is how the engine ranks exchange boundary series (there: NCCL)."""
import os
import socket

import numpy as np
import pytest

from hydpy_b200 import layout, shard, synth
from oracle import oracle
from tests.randland import add_mct, random_forcing, random_land, random_river


def make_problem(ne=240, nt=40, levels=14, seed=3):
    land, states = synth.make_land(ne, "1h", variants=True, seed=seed, bank=64)
    forcing, doy = synth.make_forcing(nt, "1h", bank=64)
    river, discharge = synth.make_tree_river(ne, "1h", levels=levels, seed=seed + 1)
    rng = np.random.default_rng(seed)
    discharge = rng.uniform(0, 3, river.n_seg)
    return land, states, river, discharge, forcing, doy


def run_unsharded(land, states, river, discharge, forcing, doy, ext=None):
    st, dis = states.copy(), discharge.copy()
    qt, _ = oracle.land_run(land, st, forcing, doy, threads=2)
    node_rows, rout, _ = oracle.river_run(river, dis, qt, ext)
    return node_rows, rout, st, dis


def run_shard_cpu(s: shard.Shard, forcing, doy, ext_rows):
    """One rank's window with the oracle in place of the engine; returns (node rows, reach outflow, states, discharge)."""
    st, dis = s.states.copy(), s.discharge.copy()
    qt, _ = oracle.land_run(s.land, st, forcing, doy, threads=1)
    node_rows, rout, _ = oracle.river_run(s.river, dis, qt, ext_rows)
    return node_rows, rout, st, dis


@pytest.mark.parametrize("world", [2, 3, 5])
def test_partition_executed_in_stage_order_is_bit_identical(world):
    land, states, river, discharge, forcing, doy = make_problem()
    nt = forcing.shape[1]
    ref_nodes, ref_rout, ref_st, ref_dis = run_unsharded(land, states, river, discharge, forcing, doy)
    shards = shard.partition(land, states, river, discharge, world)
    assert sum(len(s.elements) for s in shards) == land.n_elem
    assert sum(len(s.nodes) for s in shards) == river.n_node
    assert sum(len(s.reaches) for s in shards) == river.n_reach
    assert sum(1 for s in shards if len(s.elements)) >= 2, "the partition left everything on one rank"
    out_rows: dict[int, np.ndarray] = {}
    nodes = np.zeros_like(ref_nodes)
    rout_g = np.zeros_like(ref_rout)
    got_st = ref_st.copy()
    for n in got_st.NAMES:
        getattr(got_st, n)[:] = np.nan
    got_dis = np.full_like(ref_dis, np.nan)
    for s in sorted(shards, key=lambda x: x.stage):
        ext = s.local_ext(np.zeros((0, nt)), nt)
        for peer, rows in s.recv:
            assert shards[peer].stage < s.stage
            sent = dict(shards[peer].send)[s.rank]
            assert len(sent) == len(rows)
            for row, q_local in zip(rows, sent):
                ext[row] = out_rows[peer][q_local]
        n_rows, r_rows, st, dis = run_shard_cpu(s, forcing, doy, ext)
        out_rows[s.rank] = r_rows
        nodes[s.nodes] = n_rows
        rout_g[s.reaches] = r_rows
        shard.scatter_states(s, st, land, got_st)
        for i, q in enumerate(s.reaches):
            got_dis[river.seg_ptr[q]:river.seg_ptr[q + 1]] = dis[s.river.seg_ptr[i]:s.river.seg_ptr[i + 1]]
    assert np.array_equal(nodes, ref_nodes)
    assert np.array_equal(rout_g, ref_rout)
    assert np.array_equal(got_dis, ref_dis)
    for n in got_st.NAMES:
        assert np.array_equal(getattr(got_st, n), getattr(ref_st, n)), n


def test_partition_keeps_external_rows_and_deploy_modes():
    ne, nt = 40, 30
    land, states = random_land(ne, seed=11)
    forcing, doy = random_forcing(nt, ne, seed=11)
    river, discharge, _ = random_river(ne, 30, seed=11, smax=12, n_ext=2)
    # one inlet node per reach (a reach reads all its inlet nodes on its own rank)
    first = river.inlet_ptr[:-1]
    river.inlet_node = river.inlet_node[first]
    river.inlet_ptr = np.arange(river.n_reach + 1, dtype=np.int64)
    river.normalise()
    rng = np.random.default_rng(0)
    ext = rng.uniform(0, 3, size=(river.n_ext, nt))
    ext[1][rng.uniform(size=nt) < 0.3] = np.nan
    ref_nodes, ref_rout, _, ref_dis = run_unsharded(land, states, river, discharge, forcing, doy, ext)
    # a hand-made cut: nodes below an index threshold form rank 0.  Reaches only flow towards higher node indices in
    # this generator, so the shard graph is acyclic (rank 0 -> rank 1)
    node_rank = (np.arange(river.n_node) >= ne + 8).astype(int)
    shards = shard.partition(land, states, river, discharge, 2, node_rank=node_rank)
    assert shards[1].recv and shards[0].send
    s0, s1 = shards
    e0 = s0.local_ext(ext, nt)
    n0, r0, _, d0 = run_shard_cpu(s0, forcing, doy, e0)
    e1 = s1.local_ext(ext, nt)
    for (peer, rows), (_, sent) in zip(s1.recv, s0.send):
        for row, q_local in zip(rows, sent):
            e1[row] = r0[q_local]
    n1, r1, _, d1 = run_shard_cpu(s1, forcing, doy, e1)
    nodes = np.zeros_like(ref_nodes)
    nodes[s0.nodes], nodes[s1.nodes] = n0, n1
    assert np.array_equal(nodes, ref_nodes, equal_nan=True)
    rout = np.zeros_like(ref_rout)
    rout[s0.reaches], rout[s1.reaches] = r0, r1
    assert np.array_equal(rout, ref_rout, equal_nan=True)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank: int, world: int, port: int, result_path: str) -> None:
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        land, states, river, discharge, forcing, doy = make_problem(ne=160, nt=24, levels=10)
        nt = forcing.shape[1]
        shards = shard.partition(land, states, river, discharge, world)
        s = shards[rank]
        ext = s.local_ext(np.zeros((0, nt)), nt)
        # the exchange plan of shard.run_window with gloo send/recv in place of NCCL
        st, dis = s.states.copy(), s.discharge.copy()
        qt, _ = oracle.land_run(s.land, st, forcing, doy, threads=1)
        for peer, rows in s.recv:
            buf = torch.zeros((len(rows), nt), dtype=torch.float64)
            dist.recv(buf, src=peer)
            ext[rows] = buf.numpy()
        node_rows, rout, _ = oracle.river_run(s.river, dis, qt, ext)
        for peer, reaches in s.send:
            dist.send(torch.from_numpy(np.ascontiguousarray(rout[reaches])), dst=peer)
        # gather the node series on rank 0 (max-over-ranks style reduction: every rank contributes its own nodes)
        full = torch.zeros((river.n_node, nt), dtype=torch.float64)
        full[torch.from_numpy(s.nodes)] = torch.from_numpy(node_rows)
        dist.all_reduce(full, op=dist.ReduceOp.SUM)
        if rank == 0:
            np.save(result_path, full.numpy())
    finally:
        dist.destroy_process_group()


def test_world_size_2_over_gloo_matches_unsharded(tmp_path):
    import torch.multiprocessing as mp

    world = 2
    path = str(tmp_path / "nodes.npy")
    mp.spawn(_gloo_worker, args=(world, _free_port(), path), nprocs=world, join=True)
    land, states, river, discharge, forcing, doy = make_problem(ne=160, nt=24, levels=10)
    ref_nodes, _, _, _ = run_unsharded(land, states, river, discharge, forcing, doy)
    assert np.array_equal(np.load(path), ref_nodes)


def test_partition_with_musk_mct_reaches():
    """A cut through a network of musk_classic and musk_mct reaches: parameters, trapezes and conditions
    of the musk_mct reaches travel with their shard; executed in stage order the result is the unsharded one bit for bit."""
    ne, nt = 40, 30
    land, states = random_land(ne, seed=13)
    forcing, doy = random_forcing(nt, ne, seed=13)
    river, discharge, _ = random_river(ne, 30, seed=13, smax=12, n_ext=0)
    first = river.inlet_ptr[:-1]
    river.inlet_node = river.inlet_node[first]
    river.inlet_ptr = np.arange(river.n_reach + 1, dtype=np.int64)
    river.normalise()
    mct = add_mct(river, seed=13)
    st, dis, m = states.copy(), discharge.copy(), mct.copy()
    qt, _ = oracle.land_run(land, st, forcing, doy)
    ref_nodes, ref_rout, _ = oracle.river_run(river, dis, qt, None, mct_states=m)
    node_rank = (np.arange(river.n_node) >= ne + 8).astype(int)
    s0, s1 = shard.partition(land, states, river, discharge, 2, node_rank=node_rank, mct_states=mct)
    assert s0.river.any_mct and s1.river.any_mct and s1.recv
    outs = {}
    got_mct = np.full_like(mct, np.nan)
    for s in (s0, s1):
        ext = s.local_ext(np.zeros((0, nt)), nt)
        for (peer, rows), (_, sent) in zip(s.recv, s0.send):
            for row, q_local in zip(rows, sent):
                ext[row] = outs[peer][1][q_local]
        st_s, dis_s, m_s = s.states.copy(), s.discharge.copy(), s.mct_states.copy()
        qt_s, _ = oracle.land_run(s.land, st_s, forcing, doy)
        outs[s.rank] = oracle.river_run(s.river, dis_s, qt_s, ext, mct_states=m_s)[:2]
        for i, q in enumerate(s.reaches):
            got_mct[:, river.seg_ptr[q]:river.seg_ptr[q + 1]] = m_s[:, s.river.seg_ptr[i]:s.river.seg_ptr[i + 1]]
    nodes = np.zeros_like(ref_nodes)
    nodes[s0.nodes], nodes[s1.nodes] = outs[0][0], outs[1][0]
    rout = np.zeros_like(ref_rout)
    rout[s0.reaches], rout[s1.reaches] = outs[0][1], outs[1][1]
    assert np.array_equal(nodes, ref_nodes, equal_nan=True) and np.array_equal(rout, ref_rout, equal_nan=True)
    assert np.array_equal(got_mct, m, equal_nan=True)


@pytest.mark.parametrize("scaling", ["weak", "strong"])
def test_bench_workload_of_several_ranks_is_the_unsharded_basin(scaling):
    """bench.py's per-rank workloads for N > 1 (ONE basin cut by shard.partition, land elements built per rank with
    `make_land(select=…)`) executed with the oracle in stage order must give the unsharded basin bit for bit — boundary
    series feed downstream REACHES here, not only outlet nodes."""
    import argparse

    import bench

    world, nt = 4, 30
    args = argparse.Namespace(elements=90 if scaling == "weak" else 360, levels=12, scaling=scaling, config=3)
    wls = [bench.build_shard(args, rank, world) for rank in range(world)]
    n_total = 360
    assert sum(w.land.n_elem for w in wls) == n_total and all(w.land.n_elem > 40 for w in wls)
    assert sum(w.river.n_reach for w in wls) == n_total - 1
    for w in wls:
        w.land.as_struct(); w.states.as_struct(); w.river.as_struct()
        assert len(w.discharge) == w.river.n_seg
    assert any(w.recv for w in wls) and any(w.send for w in wls)
    forcing, doy = synth.make_forcing(nt, "1h", seed=0)
    land, states = synth.make_land(n_total, "1h", seed=1)
    river, discharge = synth.make_tree_river(n_total, "1h", seed=2, levels=12)
    ref_nodes, _, _, _ = run_unsharded(land, states, river, discharge, forcing, doy)
    node_rank = shard.assign_nodes(None, river, world, outlet_node=np.arange(n_total), zones=np.full(n_total, 12.0))
    stages = [w.desc["shard_stage"] for w in wls]
    out_rows: dict[int, np.ndarray] = {}
    nodes = np.zeros_like(ref_nodes)
    fed_reach = False
    for rank in sorted(range(world), key=lambda r: stages[r]):
        w = wls[rank]
        ext = np.zeros((w.river.n_ext, nt))
        for peer, rows in w.recv:
            sent = dict(wls[peer].send)[rank]
            assert stages[peer] < stages[rank] and len(sent) == len(rows)
            for row, q_local in zip(rows, sent):
                ext[row] = out_rows[peer][q_local]
        st, dis = w.states.copy(), w.discharge.copy()
        qt, _ = oracle.land_run(w.land, st, forcing, doy, threads=1)
        n_rows, r_rows, _ = oracle.river_run(w.river, dis, qt, ext)
        out_rows[rank] = r_rows
        nodes[np.flatnonzero(node_rank == rank)] = n_rows
        # some received row enters a node that a local reach drains: the exchange is on the routing's critical path
        ext_nodes = {n for n in range(w.river.n_node)
                     for s_ in w.river.entry_src[w.river.entry_ptr[n]:w.river.entry_ptr[n + 1]] if s_ <= layout.ENTRY_EXT}
        fed_reach = fed_reach or bool(ext_nodes & set(w.river.inlet_node.tolist()))
    assert fed_reach
    assert np.array_equal(nodes, ref_nodes)
