"""hydpy_b200 — B200-native engine for HydPy's data-parallel core (hland_96 + musk_classic).

Public API:
    simulate(hp)            drop-in for `HydPy.simulate()`
    install()/uninstall()   patch / restore `hydpy.HydPy.simulate`
    Engine                  1:1 Python face of the C-ABI (`include/hydpy_b200.h`)
    extract / Problem       HydPy objects → SoA bundles
    synth                   synthetic basins of BASELINE.json's configs (bulk SoA API)
    shard                   element sharding over several GPUs + the NCCL exchange plan
    calib                   nse/kge/rmse of calibration ensembles from device-reduced statistics
"""
from .abi import LandBundle, LandStates, RiverBundle
from .engine import Engine, EngineError, EngineUnavailable, device_count
from .extract import Problem, UnsupportedModelError, extract, requested_reach_states, requested_series, write_back
from .simulate import get_engine, install, run_problem, simulate, uninstall

__all__ = [
    "Engine", "EngineError", "EngineUnavailable", "LandBundle", "LandStates", "Problem", "RiverBundle",
    "UnsupportedModelError", "device_count", "extract", "get_engine", "install", "requested_series", "run_problem",
    "simulate", "uninstall", "write_back",
]
