O=gpurun_out/r03s; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -12 $O/pytest_gpu.log
python - <<'PY' > $O/classes.log 2>&1
import numpy as np, time
from hydpy_b200 import Engine, synth
from hydpy_b200.abi import RiverBundle
# 10^5 elements with two snow classes: fast path vs general kernel (240 h)
land, states = synth.make_land(100_000, "1h")
ne = land.n_elem
nz = np.diff(land.zone_ptr)
# give every element 2 snow classes
import copy
land2 = copy.deepcopy(land)
land2.sfdist_ptr = np.arange(ne + 1, dtype=np.int64) * 2
land2.sfdist = np.tile(np.array([0.8, 1.2]), ne)
land2.snow_ptr = np.concatenate([[0], np.cumsum(nz * 2)]).astype(np.int64)
land2.normalise()
st2 = states.copy()
st2.sp = np.zeros(int(land2.snow_ptr[-1])); st2.wc = np.zeros(int(land2.snow_ptr[-1]))
forcing, doy = synth.make_forcing(240, "1h", t0=0)
with Engine(0) as eng:
    for path in ("auto", "general"):
        eng.set_land_path(path)
        eng.set_land(land2); eng.set_land_states(st2); eng.set_river(RiverBundle.empty(0))
        eng.set_forcing(forcing, doy)
        for rep in range(2):
            eng.set_land_states(st2)
            eng.timing(); eng.run(land=True, river=False); tm = eng.timing()
        q = eng.get_land_outflow()
        print(path, "land ms", round(tm["land_ms"], 2), "zone", round(tm["zone_ms"], 2), "elem", round(tm["element_ms"], 2), "general", round(tm["general_ms"], 2), "n_fast", tm["n_fast"], "mean q", q.mean())
        if path == "auto": q_auto = q
    print("max rel dev auto vs general", np.max(np.abs(q_auto - q) / np.maximum(np.abs(q), 1e-12)))
PY
cat $O/classes.log | tail -5
