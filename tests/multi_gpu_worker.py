"""Worker of tests/test_gpu_multi.py (one process per GPU, launched with torch.distributed.run): a basin cut by
`hydpy_b200.shard.partition`, boundary series handed over with NCCL (hb_comm_send_rows / hb_comm_recv_ext), must give
the single-GPU result bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hydpy_b200 import Engine, Problem, run_problem, shard, synth  # noqa: E402


def main() -> None:
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ne, win, nw = 6000, 48, 3
    land, states = synth.make_land(ne, "1h", variants=True, seed=5)
    river, discharge = synth.make_tree_river(ne, "1h", levels=40, seed=6)
    discharge = np.random.default_rng(7).uniform(0, 2, river.n_seg)
    forcing, doy = synth.make_forcing(win * nw, "1h")
    shards = shard.partition(land, states, river, discharge, world)
    s = shards[rank]
    with Engine(local) as eng:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid[:] = torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8)
        dist.broadcast(uid, src=0)
        eng.comm_init(rank, world, bytes(uid.numpy().tobytes()))
        eng.set_land(s.land)
        eng.set_land_states(s.states)
        eng.set_river(s.river)
        eng.set_river_states(s.discharge)
        full = torch.zeros((river.n_node, win * nw), dtype=torch.float64)
        for w in range(nw):
            eng.set_forcing(forcing[:, w * win:(w + 1) * win], doy[w * win:(w + 1) * win])
            shard.run_window(eng, s)
            rows = eng.get_node_series()
            full[torch.from_numpy(s.nodes), w * win:(w + 1) * win] = torch.from_numpy(rows)
        eng.comm_barrier()
        dist.all_reduce(full, op=dist.ReduceOp.SUM)
        if rank == 0:
            ref = run_problem(Problem(land=land, states=states, river=river, discharge=discharge, forcing=forcing,
                                      doy=doy, ext=np.zeros((0, win * nw))), eng, window=win)
            same = np.array_equal(full.numpy(), ref["node_rows"])
            sizes = [len(x.elements) for x in shards]
            if not same:
                got, want = full.numpy(), ref["node_rows"]
                bad = np.argwhere(got != want)
                node_rank = np.zeros(river.n_node, int)
                for x in shards:
                    node_rank[x.nodes] = x.rank
                nodes_bad = np.unique(bad[:, 0])
                print(f"MULTI_GPU_DIFF {len(bad)} cells, {len(nodes_bad)} nodes on ranks "
                      f"{np.bincount(node_rank[nodes_bad], minlength=world).tolist()}, first step {bad[:, 1].min()}, "
                      f"windows {np.unique(bad[:, 1] // win).tolist()}, max abs diff {np.abs(got - want).max():.3e}, "
                      f"example node {bad[0, 0]} step {bad[0, 1]}: {got[bad[0, 0], bad[0, 1]]!r} vs {want[bad[0, 0], bad[0, 1]]!r}",
                      flush=True)
            print(f"MULTI_GPU_PARITY world={world} shard_elements={sizes} stages={[x.stage for x in shards]} "
                  f"bit_identical={same}", flush=True)
            if not same:
                sys.exit(3)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
