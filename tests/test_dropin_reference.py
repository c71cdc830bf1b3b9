"""The drop-in boundary on real reference objects (CPU, only where the reference build of oracle/build_ref.py exists —
i.e. in the build container; skipped on the GPU box, which has no /root/reference)."""
import os
import subprocess
import sys

import pytest

from oracle import build_ref

HERE = os.path.dirname(os.path.abspath(__file__))
BUILT = build_ref.import_paths() is not None


@pytest.mark.skipif(not BUILT, reason="no reference build in this environment (run oracle/build_ref.py)")
def test_simulate_leaves_hydpy_objects_like_the_reference_does():
    proc = subprocess.run([sys.executable, os.path.join(HERE, "dropin_reference_worker.py")], capture_output=True,
                          text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "DROPIN_OK" in proc.stdout


@pytest.mark.skipif(not BUILT, reason="no reference build in this environment (run oracle/build_ref.py)")
def test_simulate_sibling_models_like_the_reference_does():
    """hland_96p + hland_96c + rconc_nash elements built through the reference API."""
    proc = subprocess.run([sys.executable, os.path.join(HERE, "dropin_reference_worker.py"), "--siblings"],
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "DROPIN_OK" in proc.stdout


@pytest.mark.skipif(not BUILT, reason="no reference build in this environment (run oracle/build_ref.py)")
def test_simulate_with_evap_ret_io_like_the_reference_does():
    """`evap_ret_io` (external reference evapotranspiration) as the PET submodel of `evap_aet_hbv96`, all series of the
    main model and the submodels kept in RAM (ref: hydpy/models/evap_ret_io.py:53-64)."""
    proc = subprocess.run([sys.executable, os.path.join(HERE, "dropin_reference_worker.py"), "--retio"],
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "DROPIN_OK" in proc.stdout


@pytest.mark.skipif(not BUILT, reason="no reference build in this environment (run oracle/build_ref.py)")
def test_simulate_with_input_nodes_like_the_reference_does():
    """Meteorological inputs fed through input nodes (`Element.inputs`) in the deploy modes oldsim, obs, obs_oldsim,
    obs_newsim and newsim: the same values reach the model, the input sequences keep them, the nodes end as in the
    reference (ref: hydpy/core/hydpytools.py:2704-2733, hydpy/cythons/modelutils.py:1080-1167)."""
    proc = subprocess.run([sys.executable, os.path.join(HERE, "dropin_reference_worker.py"), "--inputnodes"],
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "DROPIN_OK" in proc.stdout


@pytest.mark.skipif(not BUILT, reason="no reference build in this environment (run oracle/build_ref.py)")
@pytest.mark.parametrize("module", ["hland_96", "musk_classic", "hland_96p", "hland_96c", "musk_mct"])
def test_reference_integration_doctests_with_the_dropin_installed(module):
    """The reference's own integration-test docstrings (every sequence of every step, six digits) with `HydPy.simulate`
    replaced by `hydpy_b200.install()`; the oracle stands in for the engine here (tests/test_gpu_dropin.py: the engine)."""
    proc = subprocess.run([sys.executable, os.path.join(HERE, "doctest_reference_worker.py"), module],
                          capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-4000:] + proc.stderr[-3000:]
    assert "DOCTEST_OK" in proc.stdout


@pytest.mark.skipif(not BUILT, reason="no reference build in this environment (run oracle/build_ref.py)")
def test_calibration_rules_become_the_modifications_the_reference_would_apply():
    """`CalibrationAdapter.mods` (Replace / Add / Multiply rules with selections and parameter steps → hb_set_members
    modifications) against `rule.apply_value()` + `update_parameters()` of the reference on real objects
    (ref: hydpy/auxs/calibtools.py:892-1178, 2134-2150)."""
    proc = subprocess.run([sys.executable, os.path.join(HERE, "calib_reference_worker.py")], capture_output=True,
                          text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "CALIB_RULES_OK" in proc.stdout
