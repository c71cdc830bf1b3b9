"""Run the doctests of one model module of the unmodified reference with `HydPy.simulate` replaced by the drop-in
(`hydpy_b200.install()`), the way hydpy/tests/run_doctests.py prepares them (own process: HydPy keeps global state).

    python tests/doctest_reference_worker.py hland_96 [--engine]

The integration tests of the docstrings (ref: hydpy/models/hland_96.py:232-1653, musk_classic.py:74-180, …) print every
sequence of every step with six digits after `IntegrationTest._perform_simulation` (ref: hydpy/core/testtools.py:647-663)
has called `hp.simulate()`.  With `--engine` the CUDA engine computes them (GPU box); without it the CPU oracle stands
in for the engine, which checks the host-side layers where no GPU is available."""
import doctest
import importlib
import io
import os
import sys
import unittest
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import dropin_reference_worker as worker  # noqa: E402  (puts the reference on sys.path)

import hydpy  # noqa: E402
from hydpy import pub  # noqa: E402
from hydpy.core import devicetools, testtools  # noqa: E402

import hydpy_b200  # noqa: E402


def main() -> None:
    name = sys.argv[1]
    engine = "--engine" in sys.argv
    module = importlib.import_module(f"hydpy.models.{name}")
    calls = {"n": 0, "fallback": 0}
    original = hydpy.HydPy.simulate
    if not engine:
        worker.simulate_module.run_problem = worker.oracle_run_problem
    hydpy_b200.install()
    installed = hydpy.HydPy.simulate

    def counting(self):
        calls["n"] += 1
        return installed(self)

    hydpy.HydPy.simulate = counting
    testtools.IntegrationTest.plot = lambda self, *args, **kwargs: None     # plotly is not installed
    del pub.projectname
    del pub.timegrids
    pub.options.prepare_testing(usecython=True, threads=0)
    testtools.IntegrationTest.plotting_options = testtools.PlottingOptions()
    suite = unittest.TestSuite()
    suite.addTest(doctest.DocTestSuite(module))
    stream = io.StringIO()
    with warnings.catch_warnings(), devicetools.clear_registries_temporarily():
        warnings.filterwarnings(action="ignore")
        result = unittest.TextTestRunner(stream=stream).run(suite)
    hydpy.HydPy.simulate = original
    problems = result.errors + result.failures
    for _, text in problems:
        print(text[-6000:])
    assert not problems, f"{len(problems)} doctest(s) of hydpy.models.{name} fail with the drop-in installed"
    assert calls["n"] > 0, "the doctests never called HydPy.simulate"
    print(f"DOCTEST_OK hydpy.models.{name}: {result.testsRun} docstring(s), HydPy.simulate called {calls['n']} times "
          f"through {'the CUDA engine' if engine else 'the oracle'}")


if __name__ == "__main__":
    main()
