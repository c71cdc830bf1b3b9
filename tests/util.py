"""Shared helpers of the test-suite: golden fixtures and comparison with the stated tolerances."""
from __future__ import annotations

import os

import numpy as np

from hydpy_b200 import layout
from hydpy_b200.extract import Problem

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

#: parity bar of BASELINE.json's north_star: ≤1e-9 relative on discharge and states
RTOL = 1e-9
#: absolute floor for values that are (nearly) zero (fluxes of 1e-300 carry no information)
ATOL = 1e-12


class Scenario:
    def __init__(self, name: str, data: dict[str, np.ndarray]):
        self.name = name
        self.data = data
        self.problem = Problem.from_npz_dict(data)

    def expected(self, key: str) -> np.ndarray | None:
        return self.data.get("exp_" + key)

    @property
    def expected_series(self) -> list[str]:
        return [n for n in layout.SERIES if "exp_" + n in self.data]


def load_scenarios(filename: str) -> dict[str, Scenario]:
    with np.load(os.path.join(GOLDEN, filename), allow_pickle=False) as z:
        names = [str(n) for n in z["__scenarios__"]]
        out = {}
        for name in names:
            prefix = name + "/"
            out[name] = Scenario(name, {k[len(prefix):]: z[k] for k in z.files if k.startswith(prefix)})
    return out


def assert_close(actual: np.ndarray, desired: np.ndarray, what: str, rtol: float = RTOL, atol: float = ATOL) -> float:
    actual = np.asarray(actual, dtype=np.float64)
    desired = np.asarray(desired, dtype=np.float64)
    assert actual.shape == desired.shape, f"{what}: shape {actual.shape} != {desired.shape}"
    nan_a, nan_d = np.isnan(actual), np.isnan(desired)
    assert np.array_equal(nan_a, nan_d), f"{what}: NaN pattern differs"
    a, d = actual[~nan_a], desired[~nan_d]
    same = a == d                                    # equal infinities included (inf - inf would be NaN)
    err = np.zeros(a.shape)
    err[~same] = np.abs(a[~same] - d[~same])
    tol = atol + rtol * np.abs(d)
    if np.any(err > tol):
        i = int(np.argmax(err - tol))
        raise AssertionError(
            f"{what}: max violation at flat index {i}: actual {a[i]!r} desired {d[i]!r} "
            f"(abs err {err[i]:.3e}, rel {err[i] / max(abs(d[i]), 1e-300):.3e})"
        )
    denom = np.maximum(np.abs(d), atol / max(rtol, 1e-300))
    return float(np.max(err / denom)) if err.size else 0.0
