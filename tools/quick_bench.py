"""Ad-hoc device timing of the two kernel families (development aid; bench.py is the contract)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from hydpy_b200 import Engine, synth

ap = argparse.ArgumentParser()
ap.add_argument("--elements", type=int, default=100_000)
ap.add_argument("--step", default="1h")
ap.add_argument("--nt", type=int, default=168)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-river", action="store_true")
ap.add_argument("--levels", type=int, default=200)
ap.add_argument("--series", default="", help="comma-separated land series to record ('all' = every hland_96 sequence)")
ap.add_argument("--path", default="auto")
ap.add_argument("--cflux", type=float, default=0.0)
ap.add_argument("--sweep", default="", help="routing sweep: persistent | grouped | launches")
ap.add_argument("--model", default="hland_96", choices=("hland_96", "hland_96p", "hland_96c"))
ap.add_argument("--kind", default="storms", help="forcing kind of synth.make_forcing: storms | dry")
ap.add_argument("--spinup", type=int, default=0, help="windows simulated (untimed output) before the first rep")
ap.add_argument("--wave-blocks", type=int, default=0)
ap.add_argument("--land-slabs", type=int, default=0)
ap.add_argument("--no-fuse", action="store_true")
args = ap.parse_args()

t = time.time()
land, states = synth.make_land(args.elements, args.step, cflux=args.cflux, model=args.model)
river, dis = synth.make_tree_river(args.elements, args.step, levels=args.levels)
print(f"generated in {time.time()-t:.1f}s: zones={land.n_zone} reaches={river.n_reach} segs={river.n_seg}")
eng = Engine(0)
print(eng.device_info())
t = time.time()
eng.set_land(land); eng.set_land_states(states); eng.set_river(river); eng.set_river_states(dis)
from hydpy_b200 import layout
series = list(layout.SERIES_96) if args.series == "all" else [x for x in args.series.split(",") if x]
eng.select_series(series)
eng.set_land_path(args.path)
if args.sweep:
    eng.set_river_sweep(args.sweep)
nbytes = 8 * args.nt * sum(land.n_zone if x in layout.SERIES_ZONE else land.n_snow if x in layout.SERIES_SNOW else land.n_elem for x in series)
print(f"uploaded in {time.time()-t:.2f}s")
if args.wave_blocks:
    eng.set_wave_blocks(args.wave_blocks)
if args.land_slabs:
    eng.set_land_slabs(args.land_slabs)
if args.no_fuse:
    eng.set_land_fuse(False)
for rep in range(-args.spinup, args.reps):
    forcing, doy = synth.make_forcing(args.nt, args.step, t0=(rep + args.spinup) * args.nt, kind=args.kind)
    t = time.time()
    eng.set_forcing(forcing, doy); eng.set_external(np.zeros((0, args.nt)))
    t1 = time.time()
    eng.run(land=True, river=not args.no_river)
    t2 = time.time()
    tm = eng.timing()
    zs = land.n_zone * args.nt
    print(f"rep {rep}: h2d {1e3*(t1-t):.1f} ms  run wall {1e3*(t2-t1):.1f} ms  land {tm['land_ms']:.2f} ms (zone {tm['zone_ms']:.2f} element {tm['element_ms']:.2f})  river {tm['river_ms']:.2f} ms "
          f"({tm['river_launches']} launches, {tm['n_levels']} levels)  -> {zs/tm['total_ms']*1e3:.3e} zone-steps/s (device)"
          + (f"  series {nbytes/1e9:.2f} GB -> {nbytes/1e6/tm['land_ms']:.0f} GB/s of land time" if series else "")
          + (f"  | RecStep chain ran in {tm['elem_steps_run']/tm['elem_steps']:.1%} of the element-steps, pow in "
             f"{tm['substeps_pow']/max(1, tm['elem_steps'])/max(1, int(land.recstep[0])):.1%} of the sub-steps; water at the soil in "
             f"{tm['zone_steps_wet']/max(1, tm['zone_steps']):.1%} of the zone-steps, dry shortcut in "
             f"{tm['zone_warpsteps_dry']/max(1, tm['zone_warpsteps']):.1%} of the warp-steps" if tm['elem_steps'] else ""))
t = time.time()
rows = eng.get_node_series() if not args.no_river else eng.get_land_outflow()
print(f"d2h {rows.nbytes/1e6:.0f} MB in {1e3*(time.time()-t):.1f} ms; finite={np.isfinite(rows).all()} mean={rows.mean():.4f}")
