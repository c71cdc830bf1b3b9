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


@pytest.mark.parametrize("rank", [0, 1, 3])
def test_bench_shards_are_consistent_bundles(rank):
    """bench.py's per-rank sub-basins (N > 1: an extra boundary reach or extra external entries) must stay valid bundles —
    the C structs are built from them on every rank."""
    import argparse

    import bench

    args = argparse.Namespace(elements=300, levels=12)
    land, states, river, discharge = bench.build_shard(args, rank, 4)
    land.as_struct()
    states.as_struct()
    d = river.as_struct()
    assert d.n_reach == river.n_reach and len(discharge) == river.n_seg
    assert len(river.reach_model) == river.n_reach and river.rpar.shape[1] == river.n_reach
    if rank == 0:
        assert river.n_ext == 3
    else:
        assert river.n_reach == 300 and river.reach_outlet[-1] == -1
