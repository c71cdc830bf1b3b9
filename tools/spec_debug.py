"""Development aid: where does the speculative SMax path deviate from the oracle?"""
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
factor = np.random.default_rng(3).choice([1.0, 3.0, 1000.0], size=ne)
nz = np.diff(land.zone_ptr)
land.zpar[ZP["smax"]] = land.zpar[ZP["smax"]] * np.repeat(factor, nz)
forcing, doy = random_forcing(2 * nt, ne, seed=5)
forcing[1] -= 6.0
live = np.isfinite(land.zp("smax")) & (land.zonetype != layout.ILAKE)
limited = np.array([live[land.zone_ptr[e]:land.zone_ptr[e + 1]].any() for e in range(ne)])
windows = [(np.ascontiguousarray(forcing[:, w * nt:(w + 1) * nt]), doy[w * nt:(w + 1) * nt]) for w in range(2)]
st_ref = states0.copy()
q_ref, lossw = [], []
for f, d in windows:
    q, ser = oracle.land_run(land, st_ref, f, d, threads=4, series=("spl", "wcl"))
    q_ref.append(q)
    loss = (ser["spl"] + ser["wcl"]) > 0.0
    lossw.append(np.array([bool(loss[:, land.zone_ptr[e]:land.zone_ptr[e + 1]].any()) for e in range(ne)]))
q_ref = np.concatenate(q_ref, axis=1)

def run(spec, lnd):
    with Engine(0) as eng:
        eng.set_land_speculate(spec)
        eng.set_land(lnd); eng.set_land_states(states0); eng.set_river(RiverBundle.empty(0))
        out = []
        for f, d in windows:
            eng.set_forcing(f, d); eng.run(land=True, river=False); out.append(eng.get_land_outflow())
        return np.concatenate(out, axis=1), eng.timing()

for label, spec in (("spec on", True), ("spec off", False)):
    q, tm = run(spec, land)
    rel = np.abs(q - q_ref) / np.maximum(np.abs(q_ref), 1e-12)
    per_e = rel.max(axis=1)
    order = np.argsort(-per_e)[:8]
    print(label, "n_spec", tm["n_spec"], "reruns", tm["spec_reruns"], "worst rel", per_e.max())
    for e in order:
        print(f"   e={e} rel={per_e[e]:.2e} at t={rel[e].argmax()} limited={limited[e]} factor={factor[e]} loss_w0={lossw[0][e]} loss_w1={lossw[1][e]} "
              f"nc={land.sfdist_ptr[e+1]-land.sfdist_ptr[e]} nz={nz[e]} types={land.zonetype[land.zone_ptr[e]:land.zone_ptr[e+1]]}")
# the same basin without any limit: pure fast path against the oracle
land_inf = copy.deepcopy(land); land_inf.zpar[ZP["smax"]] = np.inf
st = states0.copy(); qi = []
for f, d in windows:
    q, _ = oracle.land_run(land_inf, st, f, d, threads=4); qi.append(q)
qi = np.concatenate(qi, axis=1)
q, tm = run(True, land_inf)
rel = np.abs(q - qi) / np.maximum(np.abs(qi), 1e-12)
print("no limits: worst rel", rel.max(), "element", rel.max(axis=1).argmax())
