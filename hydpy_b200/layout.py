"""Constants of the C-ABI (`include/hydpy_b200.h`) mirrored for Python.

`tests/test_abi.py` parses the header and checks that both sides agree, so the header stays the single source of
truth.  Names follow the reference's parameter/sequence names (`hydpy/models/hland/hland_control.py`,
`hland_derived.py`, `hland_fluxes.py`, …; `hydpy/models/evap/evap_control.py`).
"""
from __future__ import annotations

ABI_VERSION = 1

# zone types, `hydpy/models/hland/hland_constants.py`
FIELD, FOREST, GLACIER, ILAKE, SEALED = 1, 2, 3, 4, 5

# per-zone flags
ZF_WATER, ZF_INTERCEPTION, ZF_SOIL, ZF_TREE, ZF_SREDEND = 1, 2, 4, 8, 16
# per-element flags
EF_RESPAREA = 1

ZPAR = (
    "zonez", "pcorr", "pcalt", "rfcf", "sfcf", "tcorr", "tcalt", "icmax", "smax", "tt", "ttint", "ttm", "cfmax",
    "cfvar", "gmelt", "gvar", "cfr", "whc", "fc", "beta", "cflux", "relzonearea", "temperaturethresholdice",
    "maxsoilwater", "soilmoisturelimit", "excessreduction", "hrualtitude", "evapotranspirationfactor",
    "altitudefactor", "precipitationfactor", "airtemperaturefactor", "hruareafraction",
)
ZP = {name: i for i, name in enumerate(ZPAR)}

EPAR = (
    "z", "relsoilarea", "rellandarea", "relupperzonearea", "rellowerzonearea", "percmax", "dt", "alpha", "k", "k4",
    "gamma", "qfactor", "petaltitude",
)
EP = {name: i for i, name in enumerate(EPAR)}

FORCING = ("p", "t", "normalairtemperature", "normalevapotranspiration")

SERIES_ZONE = (
    "tc", "fracrain", "cfact", "gact", "pc", "ei", "tf", "spl", "wcl", "spg", "wcg", "glmelt", "in_", "r", "sr",
    "ea", "cf", "el", "ic", "sm",
)
SERIES_SNOW = ("swe", "melt", "refr", "sp", "wc")
SERIES_ELEM = ("contriarea", "inuz", "perc", "q0", "q1", "inrc", "outrc", "rt", "qt", "uz", "lz")
SERIES = SERIES_ZONE + SERIES_SNOW + SERIES_ELEM
S = {name: i for i, name in enumerate(SERIES)}

#: sequence subgroup of every land series in the reference model (`model.sequences.<group>.<name>`)
SERIES_GROUP = {
    **{n: "factors" for n in ("tc", "fracrain", "cfact", "gact", "swe", "contriarea")},
    **{n: "states" for n in ("ic", "sm", "sp", "wc", "uz", "lz")},
}
for _n in SERIES:
    SERIES_GROUP.setdefault(_n, "fluxes")

NODE_NEWSIM, NODE_EXT, NODE_OBSNEWSIM = 0, 1, 2

RUN_LAND, RUN_RIVER = 1, 2
OPT_LAND_PATH = 0
OPT_RIVER_SWEEP = 1
OPT_RIVER_STATE_SERIES = 2
RIVER_PERSISTENT, RIVER_LAUNCHES = 0, 1
LAND_AUTO, LAND_GENERAL = 0, 1
ROW_NODE, ROW_REACH = 0, 1
ENTRY_EXT = -(1 << 30)

NCCL_ID_BYTES = 128

#: rows of `hb_node_statistics` (enum hb_stat)
STATS = ("n", "mean_sim", "mean_obs", "ss_err", "ss_sim", "ss_obs", "sp")
