"""Development aid: routing of single shards of the bench's partitioned basin, alone on one GPU (synchronous windows)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from hydpy_b200 import Engine, synth

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--ranks", default="0,7")
ap.add_argument("--scaling", default="weak")
ap.add_argument("--window", type=int, default=438)
a = ap.parse_args()
args = argparse.Namespace(elements=100_000, levels=200, scaling=a.scaling, config=3)
with Engine(0) as eng:
    for rank in [int(r) for r in a.ranks.split(",")]:
        wl = bench.build_shard(args, rank, a.world) if a.world > 1 else bench.build_shard(args, 0, 1)
        r = wl.river
        nseg = np.diff(r.seg_ptr) - 1
        print(f"rank {rank}: {wl.land.n_elem} elements, {r.n_reach} reaches, {r.n_node} nodes, n_ext {r.n_ext}, "
              f"entries/node max {np.diff(r.entry_ptr).max()}, segments mean {nseg.mean():.1f}", flush=True)
        eng.set_land(wl.land); eng.set_land_states(wl.states); eng.set_river(r); eng.set_river_states(wl.discharge)
        for w in range(3):
            f, d = synth.make_forcing(a.window, "1h", t0=w * a.window)
            eng.set_forcing(f, d)
            eng.set_external(np.zeros((r.n_ext, a.window)))
            eng.timing()
            eng.run()
            tm = eng.timing()
            print(f"   window {w}: land {tm['land_ms']:.2f} ms, routing {tm['river_ms']:.2f} ms, {int(tm['river_launches'])} launches, "
                  f"{int(tm['n_levels'])} levels", flush=True)
