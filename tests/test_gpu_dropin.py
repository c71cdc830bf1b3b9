"""The drop-in end to end ON THE GPU with real objects of the unmodified reference (`oracle/_ref` is an importable
installation of it, built by oracle/build_ref.py; it travels to the GPU box): `hydpy_b200.simulate(hp)` and the installed
`HydPy.simulate` — extract → CUDA engine → write_back — against the reference's own `hp.simulate()` on the same objects
(ref: hydpy/core/hydpytools.py:2769-2952), and the reference's integration-test doctests with the engine in the middle
(ref: hydpy/core/testtools.py:647-663)."""
import os
import subprocess
import sys

import pytest

from oracle import build_ref

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
AVAILABLE = build_ref.import_paths() is not None


def run_worker(*args: str) -> str:
    proc = subprocess.run([sys.executable, os.path.join(HERE, "dropin_reference_worker.py"), "--engine", *args],
                          capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "DROPIN_OK" in proc.stdout
    print(proc.stdout[-600:])
    return proc.stdout


@pytest.mark.skipif(not AVAILABLE, reason="no importable reference (oracle/_ref) in this environment")
def test_lahn_through_the_engine_leaves_hydpy_objects_like_the_reference_does():
    out = run_worker()
    assert "CUDA engine in the middle" in out


@pytest.mark.skipif(not AVAILABLE, reason="no importable reference (oracle/_ref) in this environment")
def test_sibling_models_through_the_engine():
    run_worker("--siblings")


@pytest.mark.skipif(not AVAILABLE, reason="no importable reference (oracle/_ref) in this environment")
def test_evap_ret_io_through_the_engine():
    run_worker("--retio")


@pytest.mark.skipif(not AVAILABLE, reason="no importable reference (oracle/_ref) in this environment")
def test_input_nodes_through_the_engine():
    run_worker("--inputnodes")


@pytest.mark.skipif(not AVAILABLE, reason="no importable reference (oracle/_ref) in this environment")
@pytest.mark.parametrize("module", ["hland_96", "musk_classic", "hland_96p", "hland_96c", "musk_mct"])
def test_reference_integration_doctests_with_the_engine_installed(module):
    """`IntegrationTest.__call__` of the reference's own docstrings with `HydPy.simulate` replaced by the engine: every
    printed table (all sequences, 6 digits) must come out as the docstring states it."""
    proc = subprocess.run([sys.executable, os.path.join(HERE, "doctest_reference_worker.py"), module, "--engine"],
                          capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-4000:] + proc.stderr[-3000:]
    assert "DOCTEST_OK" in proc.stdout
    print(proc.stdout[-400:])
