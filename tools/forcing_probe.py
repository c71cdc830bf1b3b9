#!/usr/bin/env python
"""How hard does a synthetic forcing work the hot loops?  (CPU oracle on an element sample; test infrastructure.)

Per bench window it prints the share of element-steps that run the RecStep chain of Calc_Q0_Perc_UZ_V1 (everything but
`uz == 0 and inuz == 0`), the share of zone-steps in which water reaches the soil (`in_ > 0`, the steps that pay for the
exp half of Calc_R_SM_V1's power function), and the water balance terms.  Usage:

    python tools/forcing_probe.py [--kind storms|dry] [--elements 256] [--window 438] [--windows 25]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from hydpy_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="storms")
    ap.add_argument("--elements", type=int, default=256)
    ap.add_argument("--window", type=int, default=438)
    ap.add_argument("--windows", type=int, default=25)
    ap.add_argument("--step", default="1h")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    land, states = synth.make_land(a.elements, a.step, seed=1)
    per_day = 24 if a.step == "1h" else 1
    print("window  days      recstep_run  soil_input   P mm/d  EI mm/d  EA mm/d  qt mm/d    sm mm   uz mm   T degC")
    tot_run = tot_in = 0.0
    for w in range(a.windows):
        forcing, doy = synth.make_forcing(a.window, a.step, seed=0, t0=w * a.window, kind=a.kind)
        uz_before = states.uz.copy()
        _, ser = oracle.land_run(land, states, forcing, doy, series=("uz", "inuz", "in_", "ei", "ea", "rt", "sm"),
                                 threads=a.threads)
        uz_prev = np.vstack([uz_before[None, :], ser["uz"][:-1]])
        run = ~((uz_prev == 0.0) & (ser["inuz"] == 0.0))
        wet = ser["in_"] > 0.0
        cols = land.forcing_col
        p = forcing[0][:, cols].mean() * per_day
        print(f"{w:5d}  {w * a.window / per_day:5.0f}-{(w + 1) * a.window / per_day:<5.0f}  {run.mean():9.3f}  "
              f"{wet.mean():10.3f}  {p:7.2f}  {ser['ei'].mean() * per_day:7.2f}  {ser['ea'].mean() * per_day:7.2f}  "
              f"{ser['rt'].mean() * per_day:7.2f}  {ser['sm'].mean():7.1f}  {ser['uz'].mean():6.2f}  "
              f"{forcing[1][:, cols].mean():7.1f}")
        if w >= 5:
            tot_run += run.mean()
            tot_in += wet.mean()
    n = max(1, a.windows - 5)
    print(f"windows 5..{a.windows - 1}: recstep_executed_frac {tot_run / n:.3f}, soil_input_frac {tot_in / n:.3f}")


if __name__ == "__main__":
    main()
