"""Development aid: fast path with several snow classes against the oracle, one element, growing window lengths."""
import copy, sys
import numpy as np
sys.path.insert(0, ".")
from hydpy_b200 import Engine, layout
from hydpy_b200.abi import RiverBundle
from oracle import oracle
from tests.randland import random_land, random_forcing

ne, nt = 900, 60
land, states0 = random_land(ne, seed=31, snowlimit_fraction=0.85, fast_fraction=0.0, recstep=10)
ZP = layout.ZP
land = copy.deepcopy(land)
land.zpar[ZP["cflux"]] = 0.0
land.zpar[ZP["smax"]] = np.inf
forcing, doy = random_forcing(2 * nt, ne, seed=5)
forcing[1] -= 6.0
E = int(sys.argv[1]) if len(sys.argv) > 1 else 746
z0, z1 = int(land.zone_ptr[E]), int(land.zone_ptr[E + 1]); s0, s1 = int(land.snow_ptr[E]), int(land.snow_ptr[E + 1])
nc = int(land.sfdist_ptr[E + 1] - land.sfdist_ptr[E]); nz = z1 - z0
print("element", E, "nz", nz, "nc", nc, "types", land.zonetype[z0:z1], "sfdist", land.sfdist[land.sfdist_ptr[E]:land.sfdist_ptr[E + 1]])
for name in ("whc", "cfmax", "cfvar", "cfr", "gmelt", "gvar", "fc", "icmax", "ttm"):
    print("  ", name, land.zp(name)[z0:z1])
print("   sp0", states0.sp[s0:s1].reshape(nc, nz)); print("   wc0", states0.wc[s0:s1].reshape(nc, nz))
with Engine(0) as eng:
    eng.set_land(land); eng.set_river(RiverBundle.empty(0))
    for n in (1, 2, 3, 5, 10, 20, 40, 60):
        f, d = np.ascontiguousarray(forcing[:, :n]), doy[:n]
        st = states0.copy()
        q_ref, _ = oracle.land_run(land, st, f, d, threads=4)
        eng.set_land_states(states0); eng.set_forcing(f, d); eng.run(land=True, river=False)
        q, ste = eng.get_land_outflow(), eng.get_land_states()
        line = [f"nt={n:3d} q rel {np.max(np.abs(q[E]-q_ref[E])/np.maximum(np.abs(q_ref[E]),1e-12)):.2e}"]
        for name, a, b in (("ic", ste.ic[z0:z1], st.ic[z0:z1]), ("sm", ste.sm[z0:z1], st.sm[z0:z1]), ("sp", ste.sp[s0:s1], st.sp[s0:s1]),
                           ("wc", ste.wc[s0:s1], st.wc[s0:s1]), ("uz", ste.uz[E:E+1], st.uz[E:E+1]), ("lz", ste.lz[E:E+1], st.lz[E:E+1])):
            dev = np.abs(a - b) / np.maximum(np.abs(b), 1e-12)
            line.append(f"{name} {dev.max():.1e}@{int(dev.argmax())}")
        print("  ".join(line))
        if n == 60:
            print("   sp eng", ste.sp[s0:s1].reshape(nc, nz)); print("   sp ref", st.sp[s0:s1].reshape(nc, nz))
            print("   wc eng", ste.wc[s0:s1].reshape(nc, nz)); print("   wc ref", st.wc[s0:s1].reshape(nc, nz))
print("T of the element:", np.round(forcing[1, :20, E], 2), "P:", np.round(forcing[0, :20, E], 2))
