"""Worker of tests/test_dropin_reference.py::test_calibration_rules… (own process: HydPy keeps global state).

`CalibrationAdapter` turns the reference's own `Rule` objects into the parameter modifications of `hb_set_members`.
Here every rule value is ALSO applied the reference's way (`rule.apply_value()` + `update_parameters()`,
ref: hydpy/auxs/calibtools.py:2134-2150) and the project extracted again: the bundle of member m built from the
modifications (`calib.apply_mods`, the NumPy twin of the device arithmetic) must hold the same parameter values (to the
last bit where no time factor is involved, else to one rounding)."""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402

SRC, STUBS = build_ref.import_paths()
sys.path[:0] = [STUBS, SRC]

import hydpy  # noqa: E402
from hydpy import Add, Multiply, Replace, pub  # noqa: E402
from hydpy.core import testtools  # noqa: E402

from hydpy_b200 import calib, layout  # noqa: E402
from hydpy_b200.extract import extract  # noqa: E402


def main() -> None:
    del pub.projectname
    del pub.timegrids
    pub.options.prepare_testing(usecython=True, threads=0)
    with contextlib.redirect_stdout(io.StringIO()):
        hp, _, _ = testtools.prepare_full_example_2(lastdate="1996-02-01")
    for node in hp.nodes:
        node.prepare_obsseries(allocate_ram=True)
        node.sequences.obs.series = 1.0
    rules = [
        Replace(name="fc_head", parameter="fc", value=200.0, model="hland_96", selections=("headwaters",)),
        Replace(name="fc_rest", parameter="fc", value=150.0, model="hland_96", selections=("nonheadwaters",)),
        Multiply(name="percmax", parameter="percmax", value=1.0, model="hland_96"),
        Add(name="tt", parameter="tt", value=0.0, model="hland_96"),
        Replace(name="k4", parameter="k4", value=0.01, parameterstep="1d", model="hland_96"),
        Multiply(name="beta", parameter="beta", value=1.0, model="hland_96"),
        Add(name="cfmax", parameter="cfmax", value=0.0, parameterstep="1d", model="hland_96"),
    ]
    adapter = calib.CalibrationAdapter(hp, rules, nodes=[hp.nodes.lahn_kalk, hp.nodes.dill_assl])
    rng = np.random.default_rng(2)
    values = np.column_stack([rng.uniform(120, 280, 5), rng.uniform(100, 250, 5), rng.uniform(0.5, 2.0, 5),
                              rng.uniform(-1, 1, 5), rng.uniform(0.005, 0.05, 5), rng.uniform(0.8, 1.3, 5),
                              rng.uniform(-0.5, 1.0, 5)])
    mods = adapter.mods(values)
    base = adapter.problem.land
    i0, i1 = pub.timegrids.simindices
    worst = 0.0
    for m in range(values.shape[0]):
        for rule, v in zip(rules, values[m]):
            rule.value = float(v)
            rule.apply_value()
        for element in hp.elements:
            element.model.update_parameters()
        ref = extract(hp, i0, i1).land
        got = calib.apply_mods(base, mods, m)
        for rows, names in ((("zpar", layout.ZPAR)), (("epar", layout.EPAR))):
            a, b = getattr(ref, rows), getattr(got, rows)
            for i, name in enumerate(names):
                x, y = a[i], b[i]
                ok = np.isclose(x, y, rtol=4e-16, atol=0.0, equal_nan=True)
                assert np.all(ok), (m, name, x[~ok][:4], y[~ok][:4])
                with np.errstate(invalid="ignore", divide="ignore"):
                    d = np.abs(x - y) / np.abs(x)
                worst = max(worst, float(np.nanmax(np.where(np.isfinite(d), d, 0.0))) if d.size else 0.0)
        for name in ("zonetype", "zflags", "uh", "recstep"):
            assert np.array_equal(getattr(ref, name), getattr(got, name)), name
        for rule in rules:
            rule.reset_parameters()
    print(f"CALIB_RULES_OK {values.shape[0]} value vectors x {len(rules)} rules; largest relative difference of a "
          f"parameter value {worst:.1e}")


if __name__ == "__main__":
    main()
