"""Constants of the C-ABI (`include/hydpy_b200.h`) mirrored for Python.

`tests/test_abi.py` parses the header and checks that both sides agree, so the header stays the single source of
truth.  Names follow the reference's parameter/sequence names (`hydpy/models/hland/hland_control.py`,
`hland_derived.py`, `hland_fluxes.py`, …; `hydpy/models/evap/evap_control.py`).
"""
from __future__ import annotations

ABI_VERSION = 5

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
    # hland_96p
    "sgr", "w0", "w1", "k2", "w2", "sg1max",
    # hland_96c
    "h1", "tab1", "tvs1", "h2", "tab2", "tvs2",
)
N_ZPAR_V1 = 32
ZP = {name: i for i, name in enumerate(ZPAR)}

EPAR = (
    "z", "relsoilarea", "rellandarea", "relupperzonearea", "rellowerzonearea", "percmax", "dt", "alpha", "k", "k4",
    "gamma", "qfactor", "petaltitude",
    # hland_96p
    "k3", "w3", "w4", "fsg",
    # rconc_nash
    "nash_ksc", "nash_dt",
)
N_EPAR_V1 = 13

# application model of a land element / its runoff concentration submodel
MODEL_HLAND_96, MODEL_HLAND_96P, MODEL_HLAND_96C = 0, 1, 2
MODELS = {"hland_96": MODEL_HLAND_96, "hland_96p": MODEL_HLAND_96P, "hland_96c": MODEL_HLAND_96C}
RCONC_UH, RCONC_NASH = 0, 1
EP = {name: i for i, name in enumerate(EPAR)}

FORCING = ("p", "t", "normalairtemperature", "normalevapotranspiration")

SERIES_ZONE1 = (
    "tc", "fracrain", "cfact", "gact", "pc", "ei", "tf", "spl", "wcl", "spg", "wcg", "glmelt", "in_", "r", "sr",
    "ea", "cf", "el", "ic", "sm",
)
SERIES_SNOW = ("swe", "melt", "refr", "sp", "wc")
SERIES_ELEM1 = ("contriarea", "inuz", "perc", "q0", "q1", "inrc", "outrc", "rt", "qt", "uz", "lz")
#: the sequences of hland_96 (the first block of enum hb_series)
SERIES_96 = SERIES_ZONE1 + SERIES_SNOW + SERIES_ELEM1
# sibling models (second block): hland_96p zone/element sequences, hland_96c zone sequences
SERIES_ZONE2 = ("dp", "rs", "ri", "gr1", "rg1", "suz", "sg1", "qab1", "qvs1", "qab2", "qvs2", "bw1", "bw2")
SERIES_ELEM2 = ("gr2", "gr3", "rg2", "rg3", "sg2", "sg3")
# evap submodels (third block), zone level
SERIES_ZONE3 = ("referenceevapotranspiration", "potentialevapotranspiration", "interceptionevaporation",
                "soilevapotranspiration", "waterevaporation", "interceptedwater", "soilwater", "snowcover")
#: storage level [nt, n_uh]: rconc_nash states.sc
SERIES_UH = ("sc",)
SERIES = SERIES_96 + SERIES_ZONE2 + SERIES_ELEM2 + SERIES_ZONE3 + SERIES_UH
SERIES_ZONE = SERIES_ZONE1 + SERIES_ZONE2 + SERIES_ZONE3
SERIES_ELEM = SERIES_ELEM1 + SERIES_ELEM2
S = {name: i for i, name in enumerate(SERIES)}
#: which application models own a sequence (cells of other elements stay NaN)
SERIES_96P_ONLY = ("dp", "rs", "ri", "gr1", "rg1", "suz", "sg1", "gr2", "gr3", "rg2", "rg3", "sg2", "sg3")
SERIES_96C_ONLY = ("qab1", "qvs1", "qab2", "qvs2", "bw1", "bw2")
SERIES_NOT_96P = ("cf", "contriarea", "inuz", "perc", "q0", "q1", "uz", "lz")
SERIES_NOT_96C = ("cf", "contriarea", "inuz", "perc", "q0", "uz")

#: sequence subgroup of every land series in the reference model (`model.sequences.<group>.<name>`)
SERIES_GROUP = {
    **{n: "factors" for n in ("tc", "fracrain", "cfact", "gact", "swe", "contriarea")},
    **{n: "states" for n in ("ic", "sm", "sp", "wc", "uz", "lz", "suz", "sg1", "sg2", "sg3", "bw1", "bw2")},
}
SERIES_GROUP["sc"] = "states"          # (of the rconc_nash submodel)
for _n in SERIES:
    SERIES_GROUP.setdefault(_n, "fluxes")

#: sequences of the submodels of a land element → the engine series that holds their values
#: (`Element.prepare_allseries()` keeps them in RAM too; ref: hydpy/cythons/modelutils.py:1127-1167 `save_data` of every
#: submodel).  Key: (path from the main model, sequence group, sequence name).
SUBMODEL_SERIES = {
    ("aetmodel.petmodel", "fluxes", "referenceevapotranspiration"): "referenceevapotranspiration",
    ("aetmodel.petmodel", "fluxes", "potentialevapotranspiration"): "potentialevapotranspiration",
    ("aetmodel.petmodel", "fluxes", "precipitation"): "pc",
    ("aetmodel", "fluxes", "potentialinterceptionevaporation"): "potentialevapotranspiration",
    ("aetmodel", "fluxes", "potentialsoilevapotranspiration"): "potentialevapotranspiration",
    ("aetmodel", "fluxes", "potentialwaterevaporation"): "potentialevapotranspiration",
    ("aetmodel", "fluxes", "interceptionevaporation"): "interceptionevaporation",
    ("aetmodel", "fluxes", "soilevapotranspiration"): "soilevapotranspiration",
    ("aetmodel", "fluxes", "waterevaporation"): "waterevaporation",
    ("aetmodel", "factors", "airtemperature"): "tc",
    ("aetmodel", "factors", "interceptedwater"): "interceptedwater",
    ("aetmodel", "factors", "soilwater"): "soilwater",
    ("aetmodel", "factors", "snowcover"): "snowcover",
    ("rconcmodel", "fluxes", "inflow"): "inrc",
    ("rconcmodel", "fluxes", "outflow"): "outrc",
    ("rconcmodel", "states", "sc"): "sc",
}
#: … and the ones the host layer derives: snowycanopy = 0 (no snowy-canopy submodel, ref: evap_model.py:4259-4261),
#: meanairtemperature = the input T (ref: hland_model.py:4267-4270), meanpotentialevapotranspiration = Σ hruareafraction ·
#: potentialevapotranspiration in zone order (ref: evap_model.py:5489-5498)
SUBMODEL_DERIVED = {
    ("aetmodel", "factors", "snowycanopy"): "zero",
    ("aetmodel.petmodel", "factors", "meanairtemperature"): "input_t",
    ("aetmodel.petmodel", "fluxes", "meanpotentialevapotranspiration"): "meanpet",
    ("aetmodel.petmodel", "fluxes", "meanreferenceevapotranspiration"): "meanret",   # (evap_ret_io)
}

NODE_NEWSIM, NODE_EXT, NODE_OBSNEWSIM = 0, 1, 2

# routing model of a reach; parameters of musk_mct reaches (enum hb_rpar) and of their trapezes (enum hb_tpar)
REACH_CLASSIC, REACH_MCT = 0, 1
RPAR = ("tolerancewaterdepth", "tolerancedischarge", "tolerancenegativeinflow", "bottomslope", "seconds",
        "segmentlength", "wq_bottomslope")
RP = {name: i for i, name in enumerate(RPAR)}
TPAR = ("bottomwidth", "sideslope", "stricklercoefficient", "calibrationfactor", "bottomdepth", "trapezeheight",
        "slopewidth", "perimeterderivative")
TP = {name: i for i, name in enumerate(TPAR)}
#: musk_mct conditions beside the discharge state, rows of the `mct_states` array [3, n_seg]
MCT_STATES = ("courantnumber", "reynoldsnumber", "referencewaterdepth")
#: 1-D musk_mct sequences recorded per step on request (enum hb_mct_series; factors, fluxes.referencedischarge, states)
MCT_SERIES = ("referencedischarge", "referencewaterdepth", "wettedarea", "surfacewidth", "celerity", "correctingfactor",
              "coefficient1", "coefficient2", "coefficient3", "courantnumber", "reynoldsnumber")
MS = {name: i for i, name in enumerate(MCT_SERIES)}
#: sequence group of each of them in the reference (`hydpy/models/musk/musk_factors.py`, `musk_fluxes.py`, `musk_states.py`)
MCT_SERIES_GROUP = {"referencedischarge": "fluxes", "courantnumber": "states", "reynoldsnumber": "states"}

RUN_LAND, RUN_RIVER = 1, 2
OPT_LAND_PATH = 0
OPT_RIVER_SWEEP = 1
OPT_RIVER_STATE_SERIES = 2
OPT_RIVER_WAVE_BLOCKS = 3
OPT_RIVER_MCT_SERIES = 4
MOD_EPAR = 1000
OPT_LAND_SLABS = 5
OPT_LAND_CHUNK = 6
OPT_LAND_FUSE = 7
OPT_RIVER_PLAIN_MIN = 8
OPT_LAND_SPECULATE = 9
RIVER_PERSISTENT, RIVER_LAUNCHES, RIVER_GROUPED = 0, 1, 2
LAND_AUTO, LAND_GENERAL = 0, 1
ROW_NODE, ROW_REACH = 0, 1
ENTRY_EXT = -(1 << 30)

NCCL_ID_BYTES = 128

#: rows of `hb_node_statistics` (enum hb_stat)
STATS = ("n", "mean_sim", "mean_obs", "ss_err", "ss_sim", "ss_obs", "sp")
