"""Host logic of `hydpy_b200.extract` that needs no reference objects: what an input sequence receives from an input node
in each deploy mode (ref: hydpy/core/hydpytools.py:2704-2733; the same cases run against real reference objects in
tests/test_dropin_reference.py where the reference build exists) and which musk_mct series a project asks for."""
import types

import numpy as np
import pytest

from hydpy_b200 import layout
from hydpy_b200.extract import UnsupportedModelError, _input_series, check_reach_series, requested_mct_series


def seq(series=None, ramflag=True, nodes=(), name="t"):
    return types.SimpleNamespace(series=None if series is None else np.asarray(series, dtype=float), ramflag=ramflag,
                                 node2idx={n: None for n in nodes}, name=name)


class Node:
    def __init__(self, mode, sim=None, obs=None, entries=()):
        self.name, self.deploymode, self.entries = "n", mode, entries
        self.sequences = types.SimpleNamespace(sim=seq(sim, sim is not None), obs=seq(obs, obs is not None))


ELEMENT = types.SimpleNamespace(name="land")
SIM = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
OBS = np.array([10.0, np.nan, 30.0, np.nan, 50.0])


def test_own_series_without_a_node():
    assert np.array_equal(_input_series(ELEMENT, seq(SIM), 1, 4), SIM[1:4])


@pytest.mark.parametrize("mode, expected", [
    ("oldsim", SIM), ("oldsim_bi", SIM), ("obs", OBS), ("obs_bi", OBS),
    ("obs_oldsim", np.array([10.0, 2.0, 30.0, 4.0, 50.0])),     # observations where available, else the old simulation
    ("newsim", np.zeros(5)),                                     # reset at every step, nothing feeds the node
    ("obs_newsim", np.array([10.0, 0.0, 30.0, 0.0, 50.0])),
])
def test_node_fed_input_per_deploy_mode(mode, expected):
    node = Node(mode, sim=SIM, obs=OBS)
    got = _input_series(ELEMENT, seq(np.full(5, 1e6), nodes=(node,)), 0, 5)     # the own series is ignored
    assert np.array_equal(got, expected, equal_nan=True)
    assert np.array_equal(_input_series(ELEMENT, seq(None, False, nodes=(node,)), 2, 4), expected[2:4], equal_nan=True)


def test_node_fed_input_errors():
    with pytest.raises(UnsupportedModelError, match="fed by other elements"):
        _input_series(ELEMENT, seq(SIM, nodes=(Node("newsim", entries=("upstream",)),)), 0, 5)
    with pytest.raises(UnsupportedModelError, match="connected to 2 nodes"):
        _input_series(ELEMENT, seq(SIM, nodes=(Node("oldsim", sim=SIM), Node("obs", obs=OBS))), 0, 5)
    with pytest.raises(UnsupportedModelError, match="deploy mode"):
        _input_series(ELEMENT, seq(SIM, nodes=(Node("newsim_update", sim=SIM),)), 0, 5)
    with pytest.raises(RuntimeError, match="not available in RAM"):
        _input_series(ELEMENT, seq(SIM, nodes=(Node("oldsim"),)), 0, 5)


class Group(types.SimpleNamespace):
    """Sequence sub-group look-alike: attribute access by name, iteration over the sequences."""

    def __iter__(self):
        return iter(vars(self).values())


def mct_element(ram):
    groups = {"factors": Group(), "fluxes": Group(), "states": Group()}
    for name in layout.MCT_SERIES + ("inflow", "outflow", "discharge"):
        group = layout.MCT_SERIES_GROUP.get(name, "states" if name == "discharge" else
                                            "fluxes" if name in ("inflow", "outflow") else "factors")
        setattr(groups[group], name, types.SimpleNamespace(name=name, ramflag=name in ram))
    return types.SimpleNamespace(name="reach", model=types.SimpleNamespace(name="musk_mct",
                                                                           sequences=types.SimpleNamespace(**groups)))


def test_requested_mct_series_follow_the_ramflags():
    problem = types.SimpleNamespace(river_elements=[mct_element({"celerity", "courantnumber", "outflow"}),
                                                    mct_element({"referencedischarge"})])
    assert requested_mct_series(problem) == ["referencedischarge", "celerity", "courantnumber"]
    check_reach_series(problem)                                  # every owned sequence is recordable
    odd = mct_element({"celerity"})
    odd.model.sequences.factors.froudenumber = types.SimpleNamespace(name="froudenumber", ramflag=True)
    with pytest.raises(UnsupportedModelError, match="froudenumber"):
        check_reach_series(types.SimpleNamespace(river_elements=[odd]))


def test_elements_with_receiver_sender_or_observer_nodes_are_refused():
    """Elements fed by `newsim`/`obs_newsim` RECEIVER nodes (or sending / observing) exchange values inside a time step
    (ref: hydpy/core/threadingtools.py:117-131, hydpytools.py:3102-3123); the whole-period engine cannot honour that and
    must say so instead of simulating something else."""
    import types

    import pytest

    from hydpy_b200.extract import UnsupportedModelError, extract

    def element(**links):
        model = types.SimpleNamespace(name="hland_96")
        return types.SimpleNamespace(name="land_x", model=model, receivers=links.get("receivers", ()),
                                     senders=links.get("senders", ()), observers=links.get("observers", ()))

    class Node:
        name = "n"

    node = Node()
    for kind in ("receivers", "senders", "observers"):
        hp = types.SimpleNamespace(nodes=[node], elements=[element(**{kind: (node,)})])
        with pytest.raises(UnsupportedModelError, match="receiver/sender/observer"):
            extract(hp, 0, 10)


def test_just_in_time_netcdf_and_undeliverable_series_are_refused():
    """`diskflag_reading` / `diskflag_writing` (just-in-time NetCDF, ref: hydpy/core/sequencetools.py:1876-1992) and series
    in RAM the engine does not compute must raise instead of staying stale (no silent fallback)."""
    import types

    import pytest

    from hydpy_b200.extract import Problem, UnsupportedModelError, check_deliverable

    def seq(name, **flags):
        return types.SimpleNamespace(name=name, ramflag=flags.get("ram", False),
                                     diskflag_reading=flags.get("reading", False),
                                     diskflag_writing=flags.get("writing", False))

    def model(name, **groups):
        return types.SimpleNamespace(name=name, sequences=types.SimpleNamespace(**groups))

    def problem(land_model=None, river_model=None, node=None):
        p = Problem.__new__(Problem)
        p.land_elements = [types.SimpleNamespace(name="land_x", model=land_model)] if land_model else []
        p.river_elements = [types.SimpleNamespace(name="stream_x", model=river_model)] if river_model else []
        p.nodes = [node] if node else []
        return p

    ok = model("hland_96", fluxes=[seq("pc", ram=True)], states=[seq("sm", ram=True)])
    ok.aetmodel = model("evap_aet_hbv96", fluxes=[seq("soilevapotranspiration", ram=True)], factors=[seq("snowycanopy", ram=True)])
    ok.aetmodel.petmodel = model("evap_pet_hbv96", fluxes=[seq("meanpotentialevapotranspiration", ram=True)])
    ok.rconcmodel = model("rconc_nash", states=[seq("sc", ram=True)], fluxes=[seq("inflow", ram=True)])
    check_deliverable(problem(land_model=ok))

    bad = model("hland_96", fluxes=[seq("pc", writing=True)])
    with pytest.raises(UnsupportedModelError, match="just in time"):
        check_deliverable(problem(land_model=bad))
    bad = model("hland_96", fluxes=[])
    bad.aetmodel = model("evap_aet_hbv96", fluxes=[seq("somethingnew", ram=True)])
    with pytest.raises(UnsupportedModelError, match="does not deliver"):
        check_deliverable(problem(land_model=bad))
    with pytest.raises(UnsupportedModelError, match="just in time"):
        check_deliverable(problem(river_model=model("musk_classic", fluxes=[seq("outflow", reading=True)])))
    node = types.SimpleNamespace(name="n", sequences=types.SimpleNamespace(sim=seq("sim", writing=True), obs=seq("obs")))
    with pytest.raises(UnsupportedModelError, match="node `n`"):
        check_deliverable(problem(node=node))
