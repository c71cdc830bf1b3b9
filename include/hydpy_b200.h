/* hydpy_b200.h — C-ABI of the B200-native engine for HydPy's data-parallel core.
 *
 * The reference (hydpy 6.5dev0, /root/reference) has NO C interface for this path: its native layer is one
 * generated Cython extension type per model (`hydpy/cythons/modelutils.py:758-2999` writes
 * `hydpy/cythons/autogen/c_<model>.pyx`), driven from Python by `HydPy.simulate`
 * (`hydpy/core/hydpytools.py:2769-2963`) or `simulate_multithreaded` (`hydpytools.py:3055-3123`).  The entry
 * points below are therefore what a binding of this path has to offer: every function names
 * the reference interface it replaces.  Plain C: pointers + sizes only, caller owns all host buffers, the library
 * owns device memory behind the opaque handle.  All floating point data is IEEE binary64, integers are int32/int64
 * as declared.  Every function returns HB_OK (0) or a negative HB_E* code; `hb_last_error` gives the message.
 *
 * Units and time scaling: all values are the reference's *internal* values, i.e. what
 * `model.parameters.<group>.fastaccess.<name>` holds after `model.update_parameters()` (time-dependent parameters
 * already converted to the simulation step, `hydpy/core/parametertools.py:1464-1620`).
 *
 * Threading: one caller thread per handle (same as `HydPy.simulate`, which is synchronous and not re-entrant).
 */
#ifndef HYDPY_B200_H
#define HYDPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_ABI_VERSION 5

/* ---- return codes ------------------------------------------------------------------------------------------ */
#define HB_OK 0
#define HB_EINVAL (-1)   /* inconsistent or missing argument                                        */
#define HB_ECUDA (-2)    /* CUDA runtime error (message holds cudaGetErrorString)                    */
#define HB_ESTATE (-3)   /* call order violated (e.g. hb_run before hb_set_land/hb_set_forcing)      */
#define HB_ENOMEM (-4)   /* host or device allocation failed                                         */
#define HB_ENCCL (-5)    /* NCCL error / NCCL not loadable                                           */
#define HB_ECYCLE (-6)   /* the node/reach network is not a DAG                                      */

/* ---- zone types: `hydpy/models/hland/hland_constants.py` ------------------------------------------------- */
#define HB_FIELD 1
#define HB_FOREST 2
#define HB_GLACIER 3
#define HB_ILAKE 4
#define HB_SEALED 5

/* ---- per-zone flags (evap_aet_hbv96 control + hland derived), bit set of `zflags[]` ------------------------- */
#define HB_ZF_WATER 1         /* evap_control.Water          (`hydpy/models/evap/evap_control.py`)             */
#define HB_ZF_INTERCEPTION 2  /* evap_control.Interception                                                     */
#define HB_ZF_SOIL 4          /* evap_control.Soil                                                             */
#define HB_ZF_TREE 8          /* evap_control.Tree                                                             */
#define HB_ZF_SREDEND 16      /* hland_derived.SRedEnd       (`hydpy/models/hland/hland_derived.py:403`)       */

/* ---- per-element flags, bit set of `eflags[]` --------------------------------------------------------------- */
#define HB_EF_RESPAREA 1      /* hland_control.RespArea                                                        */

/* ---- application model of a land element (sibling models on the hland_96 skeleton) ------------ *
 * `hydpy/models/hland_96.py`, `hland_96p.py` (PREVAH response: SUZ/SG1 per zone, SG2/SG3 per element),
 * `hland_96c.py` (COSERO response: BW1/BW2 per zone, LZ per element).                                            */
#define HB_MODEL_HLAND_96 0
#define HB_MODEL_HLAND_96P 1
#define HB_MODEL_HLAND_96C 2
/* ---- runoff concentration submodel of a land element (`hydpy/models/rconc_uh.py`, `rconc_nash.py`) ------------- */
#define HB_RCONC_UH 0     /* rconc_uh, or no submodel when the element's uh range is empty                       */
#define HB_RCONC_NASH 1   /* rconc_nash: the uh range holds NmbStorages cells of `states.sc` (uh[] unused)        */

/* ---- per-zone double parameters: rows of `zpar[HB_ZP_COUNT][n_zone]` ---------------------------------------- *
 * hland_96 control (`hydpy/models/hland/hland_control.py`), derived (`hland_derived.py`), evap_aet_hbv96 and
 * evap_pet_hbv96 control/derived (`hydpy/models/evap/evap_control.py`, `evap_derived.py`).                      */
enum hb_zpar {
  HB_ZP_ZONEZ = 0, HB_ZP_PCORR, HB_ZP_PCALT, HB_ZP_RFCF, HB_ZP_SFCF, HB_ZP_TCORR, HB_ZP_TCALT, HB_ZP_ICMAX,
  HB_ZP_SMAX, HB_ZP_TT, HB_ZP_TTINT, HB_ZP_TTM /* derived TTM = TT + DTTM */, HB_ZP_CFMAX, HB_ZP_CFVAR,
  HB_ZP_GMELT, HB_ZP_GVAR, HB_ZP_CFR, HB_ZP_WHC, HB_ZP_FC, HB_ZP_BETA, HB_ZP_CFLUX,
  HB_ZP_RELZONEAREA,                     /* hland_derived.RelZoneAreas                                         */
  HB_ZP_TEMPERATURETHRESHOLDICE,         /* aet control                                                        */
  HB_ZP_MAXSOILWATER, HB_ZP_SOILMOISTURELIMIT, HB_ZP_EXCESSREDUCTION,
  HB_ZP_HRUALTITUDE,                     /* pet control                                                        */
  HB_ZP_EVAPOTRANSPIRATIONFACTOR, HB_ZP_ALTITUDEFACTOR, HB_ZP_PRECIPITATIONFACTOR, HB_ZP_AIRTEMPERATUREFACTOR,
  HB_ZP_HRUAREAFRACTION,                 /* pet derived                                                        */
  /* hland_96p only: control SGR, K2, SG1Max; derived W0, W1, W2 (`hland_derived.py`)                          */
  HB_ZP_SGR, HB_ZP_W0, HB_ZP_W1, HB_ZP_K2, HB_ZP_W2, HB_ZP_SG1MAX,
  /* hland_96c only: control H1, TAb1, TVs1, H2, TAb2, TVs2                                                     */
  HB_ZP_H1, HB_ZP_TAB1, HB_ZP_TVS1, HB_ZP_H2, HB_ZP_TAB2, HB_ZP_TVS2,
  HB_ZP_COUNT
};

/* ---- per-element double parameters: rows of `epar[HB_EP_COUNT][n_elem]` ------------------------------------- */
enum hb_epar {
  HB_EP_Z = 0 /* hland_derived.Z */, HB_EP_RELSOILAREA, HB_EP_RELLANDAREA, HB_EP_RELUPPERZONEAREA,
  HB_EP_RELLOWERZONEAREA, HB_EP_PERCMAX, HB_EP_DT /* hland_derived.DT = 1/recstep */, HB_EP_ALPHA, HB_EP_K,
  HB_EP_K4 /* hland_96, hland_96c: control K4; hland_96p: derived K4 */, HB_EP_GAMMA,
  HB_EP_QFACTOR /* hland_derived.QFactor */, HB_EP_PETALTITUDE /* evap_derived.Altitude */,
  /* hland_96p only: control K3, derived W3, W4, fixed FSG (`hland_fixed.py`)                                    */
  HB_EP_K3, HB_EP_W3, HB_EP_W4, HB_EP_FSG,
  /* rconc_nash only: derived KSC, DT (`hydpy/models/rconc/rconc_derived.py`)                                    */
  HB_EP_NASH_KSC, HB_EP_NASH_DT,
  HB_EP_COUNT
};

/* ---- forcing fields: rows of `forcing[HB_FORCING_COUNT][nt][n_col]` ------------------------------------------ *
 * hland_inputs.P, hland_inputs.T (`hydpy/models/hland/hland_inputs.py`), evap_inputs.NormalAirTemperature,
 * evap_inputs.NormalEvapotranspiration (`hydpy/models/evap/evap_inputs.py`).                                     */
enum hb_forcing { HB_F_P = 0, HB_F_T, HB_F_NORMALAIRTEMPERATURE, HB_F_NORMALEVAPOTRANSPIRATION, HB_FORCING_COUNT };

/* ---- optional per-step output series of the land model (hland_96 factor/flux/state sequences) -------------- *
 * Layouts on the host: zone sequences [nt][n_zone], snow sequences [nt][n_snow] (element-wise [class][zone]),
 * element sequences [nt][n_elem].  (`hydpy/models/hland/hland_{factors,fluxes,states}.py`)                       */
enum hb_series {
  /* zone level */
  HB_S_TC = 0, HB_S_FRACRAIN, HB_S_CFACT, HB_S_GACT, HB_S_PC, HB_S_EI, HB_S_TF, HB_S_SPL, HB_S_WCL, HB_S_SPG,
  HB_S_WCG, HB_S_GLMELT, HB_S_IN, HB_S_R, HB_S_SR, HB_S_EA, HB_S_CF, HB_S_EL, HB_S_IC, HB_S_SM,
  /* snow-class level */
  HB_S_SWE, HB_S_MELT, HB_S_REFR, HB_S_SP, HB_S_WC,
  /* element level */
  HB_S_CONTRIAREA, HB_S_INUZ, HB_S_PERC, HB_S_Q0, HB_S_Q1, HB_S_INRC, HB_S_OUTRC, HB_S_RT, HB_S_QT, HB_S_UZ,
  HB_S_LZ,
  /* sibling models, zone level: hland_96p fluxes DP, RS, RI, GR1, RG1 and states SUZ, SG1; hland_96c fluxes QAb1, QVs1,
   * QAb2, QVs2 and states BW1, BW2 (cells of elements of another model keep NaN) */
  HB_S_DP, HB_S_RS, HB_S_RI, HB_S_GR1, HB_S_RG1, HB_S_SUZ, HB_S_SG1,
  HB_S_QAB1, HB_S_QVS1, HB_S_QAB2, HB_S_QVS2, HB_S_BW1, HB_S_BW2,
  /* sibling models, element level: hland_96p fluxes GR2, GR3, RG2, RG3 and states SG2, SG3 */
  HB_S_GR2, HB_S_GR3, HB_S_RG2, HB_S_RG3, HB_S_SG2, HB_S_SG3,
  /* sequences of the evap submodels (every land model has them), zone level: evap_pet_hbv96 fluxes
   * ReferenceEvapotranspiration, PotentialEvapotranspiration; evap_aet_hbv96 fluxes InterceptionEvaporation,
   * SoilEvapotranspiration, WaterEvaporation and factors InterceptedWater, SoilWater, SnowCover
   * (`hydpy/models/evap/evap_fluxes.py`, `evap_factors.py`).  The other submodel sequences are copies the host layer
   * makes: AirTemperature = TC, Precipitation = PC, the three potential fluxes of evap_aet_hbv96 = PotentialEvapotranspiration,
   * rconc Inflow/Outflow = InRC/OutRC. */
  HB_S_REFERENCEEVAPOTRANSPIRATION, HB_S_POTENTIALEVAPOTRANSPIRATION, HB_S_INTERCEPTIONEVAPORATION,
  HB_S_SOILEVAPOTRANSPIRATION, HB_S_WATEREVAPORATION, HB_S_INTERCEPTEDWATER, HB_S_SOILWATER, HB_S_SNOWCOVER,
  /* storage level [nt][n_uh]: rconc_nash states.sc of every step (cells of rconc_uh elements keep NaN) */
  HB_S_SC,
  HB_S_COUNT
};
#define HB_S_FIRST_SNOW HB_S_SWE
#define HB_S_FIRST_ELEM HB_S_CONTRIAREA
#define HB_S_FIRST_ZONE2 HB_S_DP    /* second block of zone-level series */
#define HB_S_FIRST_ELEM2 HB_S_GR2   /* second block of element-level series */
#define HB_S_FIRST_ZONE3 HB_S_REFERENCEEVAPOTRANSPIRATION   /* third block: zone-level series of the evap submodels */
/* level of a series: 0 zone [nt][n_zone], 1 snow [nt][n_snow], 2 element [nt][n_elem], 3 storage [nt][n_uh] */
#define HB_S_LEVEL(s) ((s) < HB_S_FIRST_SNOW ? 0 : (s) < HB_S_FIRST_ELEM ? 1 : (s) < HB_S_FIRST_ZONE2 ? 2 : (s) < HB_S_FIRST_ELEM2 ? 0 : (s) < HB_S_FIRST_ZONE3 ? 2 : (s) < HB_S_SC ? 0 : 3)

/* ---- land description: all hland_96 / hland_96p / hland_96c elements (+ evap_aet_hbv96/evap_pet_hbv96, optional
 * rconc_uh / rconc_nash) --------------------------------------------------------------------------------------- *
 * Replaces the per-element `Parameters`/`ControlParameters`/`DerivedParameters` extension types of
 * `c_hland_96.pyx`, `c_evap_aet_hbv96.pyx`, `c_evap_pet_hbv96.pyx`, `c_rconc_uh.pyx`
 * (generator: `hydpy/cythons/modelutils.py:1000-1050`).  Ragged dimensions are CSR-encoded.                      */
typedef struct hb_land_desc {
  int64_t n_elem;          /* number of hland_96 elements (ensemble members are additional elements)            */
  int64_t n_zone;          /* sum of NmbZones                                                                   */
  int64_t n_snow;          /* sum of SClass*NmbZones                                                            */
  int64_t n_sfdist;        /* sum of SClass                                                                     */
  int64_t n_uh;            /* sum of len(rconc_uh.UH) (0 for elements without rconc submodel)                   */
  int64_t n_sred;          /* sum of hland_derived.SRedNumber                                                   */
  const int64_t* zone_ptr;    /* [n_elem+1] zones of element e = zone_ptr[e]..zone_ptr[e+1]-1                   */
  const int64_t* snow_ptr;    /* [n_elem+1] snow cells of e, index inside = c*NmbZones + k                      */
  const int64_t* sfdist_ptr;  /* [n_elem+1] → sfdist[]; SClass = sfdist_ptr[e+1]-sfdist_ptr[e]                  */
  const int64_t* uh_ptr;      /* [n_elem+1] → uh[] and quh[]; empty range ⇒ `model.rconcmodel is None`          */
  const int64_t* sred_ptr;    /* [n_elem+1] → sred_from/to/adjust                                               */
  const int32_t* recstep;     /* [n_elem]  hland_control.RecStep (already scaled to the simulation step)        */
  const int32_t* eflags;      /* [n_elem]  HB_EF_*                                                              */
  const int32_t* forcing_col; /* [n_elem]  column of the forcing arrays this element reads                      */
  const int32_t* outlet_node; /* [n_elem]  node receiving `outlets.q` (−1: none)                                */
  const double* epar;         /* [HB_EP_COUNT][n_elem]                                                          */
  const int32_t* zonetype;    /* [n_zone]  HB_FIELD..HB_SEALED                                                  */
  const int32_t* zflags;      /* [n_zone]  HB_ZF_*                                                              */
  const int32_t* zorder;      /* [n_zone]  hland_derived.IndicesZoneZ (zone index local to its element)         */
  const double* zpar;         /* [HB_ZP_COUNT][n_zone]                                                          */
  const double* sfdist;       /* [n_sfdist] hland_control.SFDist                                                */
  const double* uh;           /* [n_uh]    rconc_control.UH ordinates                                           */
  const int32_t* sred_from;   /* [n_sred]  hland_derived.SRedOrder[:,0] (local zone index)                      */
  const int32_t* sred_to;     /* [n_sred]  hland_derived.SRedOrder[:,1]                                         */
  const double* sred_adjust;  /* [n_sred]  ZoneAreaRatios[f,t]*SRed[f,t]  (`hland_model.py:897`)                 */
  /* ABI 2: sibling models; each pointer may be NULL (all elements hland_96 / rconc_uh)                           */
  const int32_t* model;       /* [n_elem]  HB_MODEL_*                                                           */
  const int32_t* rconc;       /* [n_elem]  HB_RCONC_*                                                           */
  const int32_t* nash_steps;  /* [n_elem]  rconc_control.NmbSteps (already scaled to the simulation step)       */
} hb_land_desc;

/* ---- land conditions: hland_96 states + rconc_uh log (`Sequences.conditions`, `sequencetools.py:807-828`) ---- */
typedef struct hb_land_states {
  double* ic;   /* [n_zone] */
  double* sm;   /* [n_zone] */
  double* sp;   /* [n_snow] */
  double* wc;   /* [n_snow] */
  double* uz;   /* [n_elem] */
  double* lz;   /* [n_elem] */
  double* quh;  /* [n_uh]   rconc_uh logs.quh; rconc_nash states.sc                                             */
  /* ABI 2: states of the sibling models; a pointer may be NULL when no element of that model exists             */
  double* suz;  /* [n_zone] hland_96p */
  double* sg1;  /* [n_zone] hland_96p */
  double* bw1;  /* [n_zone] hland_96c */
  double* bw2;  /* [n_zone] hland_96c */
  double* sg2;  /* [n_elem] hland_96p */
  double* sg3;  /* [n_elem] hland_96p */
} hb_land_states;

/* ---- river description: all musk_classic elements and the node graph --------------------------------------- *
 * Replaces `c_musk_classic.pyx` objects plus the node⇄element pointer wiring of
 * `hydpy/cythons/pointerutils.pyx` / `hydpy/core/threadingtools.py:396-446`.
 * A node's simulated series is 0.0 + Σ of its entries in the listed order (`threadingtools.py:401-415`); a reach's
 * inflow is 0.0 + Σ over its inlet nodes in the listed order (`musk_model.py:26-31`).                              */
typedef struct hb_river_desc {
  int64_t n_node;
  int64_t n_reach;
  int64_t n_entry;            /* total number of node entries                                                   */
  int64_t n_inlet;            /* total number of reach inlet links                                              */
  int64_t n_seg;              /* sum over reaches of (NmbSegments+1) = length of the discharge state            */
  int64_t n_ext;              /* number of external (given) series rows                                         */
  const int64_t* entry_ptr;   /* [n_node+1]                                                                     */
  const int32_t* entry_src;   /* [n_entry] ≥0: land element index;  <0: reach index = -1-value;
                                 ≤ HB_ENTRY_EXT: external row = HB_ENTRY_EXT-value (an upstream shard's outflow)  */
  const int32_t* node_mode;   /* [n_node]  HB_NODE_*                                                            */
  const int32_t* node_ext;    /* [n_node]  row of the external series used by HB_NODE_EXT/OBSNEWSIM (−1: none)  */
  const int64_t* inlet_ptr;   /* [n_reach+1]                                                                    */
  const int32_t* inlet_node;  /* [n_inlet]                                                                      */
  const int32_t* reach_outlet;/* [n_reach] node receiving the reach outflow (−1: none)                          */
  const int64_t* seg_ptr;     /* [n_reach+1] → discharge[]; NmbSegments = seg_ptr[r+1]-seg_ptr[r]-1             */
  const double* coef;         /* [3][n_reach] musk_control.Coefficients c0,c1,c2 (`musk_control.py:128-315`)    */
  /* ABI 2: musk_mct reaches (`hydpy/models/musk_mct.py`) with their `wq_trapeze_strickler` cross
   * sections (`hydpy/models/wq_trapeze_strickler.py`).  reach_model may be NULL (all musk_classic); the other arrays are
   * only read for HB_REACH_MCT reaches.                                                                           */
  const int32_t* reach_model; /* [n_reach] HB_REACH_*                                                           */
  const int32_t* nmbruns;     /* [n_reach] musk_solver.NmbRuns                                                  */
  const double* rpar;         /* [HB_RP_COUNT][n_reach]                                                         */
  int64_t n_trap;             /* total number of trapezes                                                       */
  const int64_t* trap_ptr;    /* [n_reach+1] → tpar columns; wq_control.NmbTrapezes = trap_ptr[r+1]-trap_ptr[r]  */
  const double* tpar;         /* [HB_TP_COUNT][n_trap]                                                          */
} hb_river_desc;

#define HB_REACH_CLASSIC 0
#define HB_REACH_MCT 1
/* per-reach double parameters of musk_mct: solver tolerances (`musk_solver.py`), control BottomSlope, derived Seconds and
 * SegmentLength [km] (`musk_derived.py`), and the BottomSlope of the wq submodel (`wq_control.py`) */
enum hb_rpar {
  HB_RP_TOLERANCEWATERDEPTH = 0, HB_RP_TOLERANCEDISCHARGE, HB_RP_TOLERANCENEGATIVEINFLOW, HB_RP_BOTTOMSLOPE,
  HB_RP_SECONDS, HB_RP_SEGMENTLENGTH, HB_RP_WQ_BOTTOMSLOPE, HB_RP_COUNT
};
/* per-trapeze double parameters of wq_trapeze_strickler: control BottomWidths, SideSlopes, StricklerCoefficients,
 * CalibrationFactors; derived BottomDepths, TrapezeHeights, SlopeWidths, PerimeterDerivatives (`wq_derived.py`) */
enum hb_tpar {
  HB_TP_BOTTOMWIDTH = 0, HB_TP_SIDESLOPE, HB_TP_STRICKLERCOEFFICIENT, HB_TP_CALIBRATIONFACTOR, HB_TP_BOTTOMDEPTH,
  HB_TP_TRAPEZEHEIGHT, HB_TP_SLOPEWIDTH, HB_TP_PERIMETERDERIVATIVE, HB_TP_COUNT
};
/* musk_mct conditions beside `states.discharge`: states CourantNumber, ReynoldsNumber and the factor
 * ReferenceWaterDepth (the start value of the next step's root search, `musk_model.py:380-399`), each stored like the
 * discharge state, i.e. [n_seg] with the last cell of every reach unused.  NULL pointers are allowed when no reach is
 * a musk_mct reach. */
typedef struct hb_river_states {
  double* discharge;           /* [n_seg] */
  double* courantnumber;       /* [n_seg] */
  double* reynoldsnumber;      /* [n_seg] */
  double* referencewaterdepth; /* [n_seg] */
} hb_river_states;

#define HB_ENTRY_EXT (-1073741824) /* -(1<<30) */

/* what a node hands to its downstream reaches (`devicetools.py:2205-2335`, `threadingtools.py:433-446`) */
#define HB_NODE_NEWSIM 0     /* freshly simulated series                                                       */
#define HB_NODE_EXT 1        /* the external row only (deploy modes oldsim / obs / oldsim_bi / obs_bi)          */
#define HB_NODE_OBSNEWSIM 2  /* external row where it is not NaN, else the simulated series (obs_newsim)        */

/* ---- life cycle -------------------------------------------------------------------------------------------- */
typedef struct hb_engine hb_engine;

/* ABI version of the loaded library (compare with HB_ABI_VERSION). */
int hb_abi_version(void);
/* Create an engine on CUDA device `device`.  Fails (HB_ECUDA) when no sm_100 device is usable: there is no CPU
 * fallback.  Replaces: instantiating the Cython `Model` objects (`hydpy/core/importtools.py:83-108`). */
int hb_create(int device, hb_engine** out);
void hb_destroy(hb_engine* h);
/* Message of the last failing call on `h` (or of the last failing hb_create when h is NULL). */
const char* hb_last_error(const hb_engine* h);

/* Pinned (page-locked) host memory for the staging buffers of hb_set_forcing / hb_get_* so that copies run at full
 * PCIe speed and asynchronously.  Plain pageable memory is accepted everywhere, too. */
int hb_host_alloc(hb_engine* h, int64_t bytes, void** out);
int hb_host_free(hb_engine* h, void* p);

/* ---- setup --------------------------------------------------------------------------------------------------- */
/* Upload structure + parameters of all land elements.  May be called again with the same structure to change
 * parameter values (calibration loop: `hydpy/auxs/calibtools.py:2134-2150`). */
int hb_set_land(hb_engine* h, const hb_land_desc* desc);
/* Upload / download the conditions.  Replaces `Model.load_conditions`/`Sequences.conditions`
 * (`hydpy/core/modeltools.py:2256`, `sequencetools.py:807-828`). */
int hb_set_land_states(hb_engine* h, const hb_land_states* st);
int hb_get_land_states(hb_engine* h, hb_land_states* st);
/* Upload the river network.  Must follow hb_set_land (entries refer to land elements). */
int hb_set_river(hb_engine* h, const hb_river_desc* desc);
/* musk_classic `states.discharge` (old == new between steps), concatenated, length n_seg. */
int hb_set_river_states(hb_engine* h, const double* discharge);
int hb_get_river_states(hb_engine* h, double* discharge);
/* The same including the musk_mct conditions (ABI 2). */
int hb_set_river_states2(hb_engine* h, const hb_river_states* st);
int hb_get_river_states2(hb_engine* h, hb_river_states* st);
/* Select which optional land series `hb_run` records (on ≠ 0).  Replaces the `ramflag` switches of
 * `IOSequence.prepare_series` (`hydpy/core/sequencetools.py:1912`, save_data `modelutils.py:1127-1167`). */
int hb_select_series(hb_engine* h, int series, int on);
/* Engine options.  HB_OPT_LAND_PATH: HB_LAND_AUTO (default) runs the elements that allow it (one snow class, SMax = inf,
 * every CFlux == 0, no series requested) through the two-kernel fast path and the rest through the general kernel;
 * HB_LAND_GENERAL forces the general kernel for all elements (results agree within the parity bar; used by tests). */
#define HB_OPT_LAND_PATH 0
#define HB_LAND_AUTO 0
#define HB_LAND_GENERAL 1
/* HB_OPT_RIVER_SWEEP: how the (topological level x 32-step chunk) wavefront of the routing runs.  HB_RIVER_PERSISTENT
 * (default): headwater reaches in two plain launches, all other reaches in ONE persistent launch whose warps draw tasks
 * from an ordered queue and wait on per-reach progress counters.  HB_RIVER_LAUNCHES: one launch per anti-diagonal.
 * Both are bit-identical to the reference. */
#define HB_OPT_RIVER_SWEEP 1
#define HB_RIVER_PERSISTENT 0
#define HB_RIVER_LAUNCHES 1
/* HB_RIVER_GROUPED: persistent wavefront whose tasks are (group of 32 reaches of one level, chunk): lanes over time build
 * the inflow rows, then lane = reach runs the recurrences — a tenth of the instructions of the warp-per-reach wavefront
 * and one small block per SM, so that the routing of window w takes little from the runoff generation of window w+1. */
#define HB_RIVER_GROUPED 2
/* HB_OPT_RIVER_STATE_SERIES (0/1): record musk_classic `states.discharge` of every step for hb_get_reach_state_series
 * (the `ramflag` of that sequence, `hydpy/models/musk/musk_states.py`). */
#define HB_OPT_RIVER_STATE_SERIES 2
/* HB_OPT_RIVER_WAVE_BLOCKS (1..8, default 3): persistent wavefront blocks per SM.  The routing of window w shares the GPU
 * with the runoff generation of window w+1 (hb_run_async); fewer blocks leave it more of every SM. */
#define HB_OPT_RIVER_WAVE_BLOCKS 3
/* HB_OPT_RIVER_MCT_SERIES (bit mask over enum hb_mct_series, default 0): musk_mct sequences to record at every step for
 * hb_get_mct_series — the `ramflag` of the 1-D factor, flux and state sequences of `hydpy/models/musk/musk_factors.py`,
 * `musk_fluxes.py:14` and `musk_states.py:10-17` (values of the last of the NmbRuns repetitions, as `save_data` sees
 * them, ref: hydpy/cythons/modelutils.py:1127-1167). */
#define HB_OPT_RIVER_MCT_SERIES 4
/* HB_OPT_LAND_SLABS (1..8, default 2): the fast path of the runoff generation works on this many slabs of elements, each
 * on a stream of its own: while the thread-per-element kernel of one slab follows its sequential RecStep chains (few
 * warps, latency-bound) the thread-per-zone kernel of another slab fills the SMs.  Results do not depend on it. */
#define HB_OPT_LAND_SLABS 5
/* HB_OPT_LAND_CHUNK (steps; 0 = as long as the scratch memory allows; default -1 = 128 once hb_comm_init has been called,
 * else 0): time chunk of the fast path.  Shorter kernels let kernels that need a whole CTA slot (the NCCL hand-over of the
 * boundary rows) onto the SMs sooner. */
#define HB_OPT_LAND_CHUNK 6
/* HB_OPT_LAND_FUSE (0 | 1, default 1): where every zone of a group of 32 consecutive elements runs the branch-free
 * FIELD/FOREST step (and no land series is selected), the thread-per-zone kernel adds the zone terms of Calc_InUZ_V1 /
 * Calc_ContriArea_V1 itself — in zone order, i.e. the same bits — and hands two sums per element-step to the
 * thread-per-element kernel instead of two terms per zone-step.  0 keeps the term buffers (development, A/B timing). */
#define HB_OPT_LAND_FUSE 7
/* HB_OPT_RIVER_PLAIN_MIN (reaches, default 4096; read by the next hb_set_river): the upstream levels of the network are
 * routed level by level with one thread per reach over the whole window while a level offers at least this many simple
 * reaches (musk_classic, at most 32 segments, fed by such reaches only); headwater reaches always are.  The rest of the
 * network — the narrow, deep part of a river tree — runs in the persistent wavefront.  Same bits either way. */
#define HB_OPT_RIVER_PLAIN_MIN 8
/* HB_OPT_LAND_SPECULATE (0 | 1, default 1; read by the next hb_set_land): hland_96 elements with a finite SMax (without
 * capillary flux) take the fast path SPECULATIVELY while no series is recorded.  As long as no snow class of any zone
 * exceeds its limit, Calc_SPL_WCL_SP_WC_V1 / Calc_SPG_WCG_SP_WC_V1 (ref: hydpy/models/hland/hland_model.py:584-605,
 * 876-969) change nothing: all losses, gains and excesses are zero.  The zone kernel watches the limit; an element whose
 * snow exceeds it somewhere in a window gets the conditions of the window's start back and repeats the window in the
 * general kernel, which redistributes (hb_timing.spec_reruns counts these element-windows); if the general kernel saw a
 * loss, the element skips the speculation of the next window and goes straight to the general kernel there.  Which
 * kernel computes a window thus depends on where the windows are cut: the fast path multiplies by host-folded reciprocals
 * where the general kernel divides (<= 1 ulp per operation), so splitting a period differently can change the last bits of
 * such an element (never more than the parity bar); with 0 every split gives the same bits.  0: such elements always run
 * the general kernel (round 1's behaviour). */
#define HB_OPT_LAND_SPECULATE 9
enum hb_mct_series {
  HB_MS_REFERENCEDISCHARGE = 0, HB_MS_REFERENCEWATERDEPTH, HB_MS_WETTEDAREA, HB_MS_SURFACEWIDTH, HB_MS_CELERITY,
  HB_MS_CORRECTINGFACTOR, HB_MS_COEFFICIENT1, HB_MS_COEFFICIENT2, HB_MS_COEFFICIENT3, HB_MS_COURANTNUMBER,
  HB_MS_REYNOLDSNUMBER, HB_MS_COUNT
};
int hb_set_option(hb_engine* h, int option, int64_t value);

/* ---- one simulation window ------------------------------------------------------------------------------------ */
/* Stage the forcing of the next window: `forcing` = [HB_FORCING_COUNT][nt][n_col] doubles, `doy` = [nt] values of
 * hland_derived.DOY at the window's steps (`hland_model.py:1056`; 0-based day of a leap year).  Replaces
 * `load_data` (`modelutils.py:1080-1125`) reading `inputs.<name>._<name>_array[idx]`. */
int hb_set_forcing(hb_engine* h, int64_t nt, int64_t n_col, const double* forcing, const int64_t* doy);
/* Same for window pipelines: the copy is only enqueued (own stream, behind the land run that last read the target
 * buffer) and overlaps the land part of the current window.  `forcing` must be page-locked (hb_host_alloc) and stay
 * untouched until hb_wait; `doy` is consumed before the call returns. */
int hb_set_forcing_async(hb_engine* h, int64_t nt, int64_t n_col, const double* forcing, const int64_t* doy);
/* Select the sub-window [t0, t0+nt) of the staged forcing for the next hb_run (default after hb_set_forcing: the whole
 * block).  Lets a caller stage a long forcing block once and advance in chunks — `pub.timegrids.sim` windows inside
 * one `init` period, ref: hydpy/core/hydpytools.py:2804-2825. */
int hb_select_window(hb_engine* h, int64_t t0, int64_t nt);
/* External node rows of the window: [n_ext][nt].  (`oldsim`/`obs` deploy modes, multi-GPU boundary hand-over.) */
int hb_set_external(hb_engine* h, int64_t nt, const double* rows);
/* Advance all land elements by the staged `nt` steps (HB_RUN_LAND), then route (HB_RUN_RIVER).  States stay on
 * the device, so consecutive windows continue seamlessly (`hydpytools.py:2817-2825`).  Replaces
 * `Model.simulate_period(i0, i1)` for every element plus the node updates of `Worker.run`
 * (`modelutils.py:1625-1634`, `threadingtools.py:358-446`).  Returns after results are complete on the device. */
#define HB_RUN_LAND 1
#define HB_RUN_RIVER 2
int hb_run(hb_engine* h, int flags);
/* Pipelined variant for window loops: enqueues the same work and returns at once.  Runoff generation of the next window
 * (after hb_set_forcing / hb_select_window) then overlaps routing and result copies of this one; the engine keeps two
 * sets of per-window buffers, so results of a window stay readable until the window after the next is staged.
 * hb_wait returns when everything enqueued so far is complete; every synchronous call waits implicitly. */
int hb_run_async(hb_engine* h, int flags);
int hb_wait(hb_engine* h);

/* ---- results of the last window -------------------------------------------------------------------------------- */
/* `outlets.q` (= fluxes.qt) of every land element: host [n_elem][nt] (element-major rows). */
int hb_get_land_outflow(hb_engine* h, double* rows);
/* node.sequences.sim.series of the window: host [n_node][nt].  n_sel = 0 ⇒ all nodes in order, else rows of
 * the listed nodes. */
int hb_get_node_series(hb_engine* h, int64_t n_sel, const int32_t* nodes, double* rows);
/* Same, enqueued behind the routing of the current window on the copy stream; `rows` must be page-locked
 * (hb_host_alloc) and stay untouched until hb_wait. */
int hb_get_node_series_async(hb_engine* h, int64_t n_sel, const int32_t* nodes, double* rows);
/* musk_classic fluxes.outflow / outlets.q of every reach: host [n_reach][nt]. */
int hb_get_reach_outflow(hb_engine* h, double* rows);
/* musk_classic states.discharge per step: host [nt][n_seg] (points of all reaches concatenated like the state itself);
 * needs HB_OPT_RIVER_STATE_SERIES. */
int hb_get_reach_state_series(hb_engine* h, double* out);
/* One recorded musk_mct sequence per step: host [nt][n_seg]; segment i of reach r sits at seg_ptr[r] + i (like the
 * musk_mct conditions of hb_river_states), all other cells are NaN.  Needs the bit in HB_OPT_RIVER_MCT_SERIES. */
int hb_get_mct_series(hb_engine* h, int which, double* out);
/* musk_classic fluxes.inflow of every reach: host [n_reach][nt]. */
int hb_get_reach_inflow(hb_engine* h, double* rows);
/* One selected land series (see enum hb_series for the layout). */
int hb_get_series(hb_engine* h, int series, double* out);

/* ---- objective reduction for calibration ensembles -------------------------------------------------- *
 * Sufficient statistics of simulated node series of the last window against observation rows, reduced on the device:
 * from them follow `nse`, `kge`, `rmse`, `corr`, `bias_abs/rel`, `std_ratio` of `hydpy/auxs/statstools.py` (:529-987)
 * without moving the series to the host.  Pair i compares node `nodes[i]` with row `obs_row[i]` of `obs[n_obs][nt]`;
 * `out` is [n_pair][HB_ST_COUNT].  skip_nan ≠ 0 drops steps where either value is NaN (`prepare_arrays(skip_nan=True)`). */
enum hb_stat {
  HB_ST_N = 0,      /* number of steps used                              */
  HB_ST_MEAN_SIM,   /* mean(sim)                                          */
  HB_ST_MEAN_OBS,   /* mean(obs)                                          */
  HB_ST_SS_ERR,     /* sum((sim - obs)^2)                                 */
  HB_ST_SS_SIM,     /* sum((sim - mean(sim))^2)                           */
  HB_ST_SS_OBS,     /* sum((obs - mean(obs))^2)                           */
  HB_ST_SP,         /* sum((sim - mean(sim)) * (obs - mean(obs)))         */
  HB_ST_COUNT
};
int hb_node_statistics(hb_engine* h, int64_t n_pair, const int32_t* nodes, const int32_t* obs_row, int64_t n_obs,
                       const double* obs, int skip_nan, double* out);

/* ---- ensemble members: the batched calibration caller ------------------------------------------------ *
 * After ONE basin has been described (hb_set_land, hb_set_river) and its conditions set (hb_set_land_states,
 * hb_set_river_states), hb_set_members turns it into `members` copies ON THE DEVICE: member m's elements, zones, nodes,
 * reaches and segments follow member m-1's in every array of the ABI (results are [members * n_node][nt] and so on);
 * forcing columns and external rows are shared; every member starts from the conditions of the basin.  The members differ
 * in control parameters: modification j gives parameter `mod_par[j]` (HB_ZP_* for a zone parameter, HB_MOD_EPAR + HB_EP_*
 * for an element parameter) of every zone / element of member m the value
 *        value * mul[m * n_mod + j] + add[m * n_mod + j]
 * (`mod_mask`, NULL or [n_mod][n_elem] bytes: the elements of the basin a modification applies to — a rule bound to a
 * `Selection`; modifications of the same parameter act in the listed order on the running value)
 * — the Replace / Add / Multiply rules of `hydpy/auxs/calibtools.py:1181` (`CalibrationInterface.apply_values`
 * `:2134-2150`) are (0, v), (1, v) and (v, 0).  The constants the kernels work with are folded from the modified values on
 * the device with the operations hb_set_land uses on the host, so a member equals a basin uploaded with those values bit for
 * bit, and an ensemble of 10^8 zones never exists on the host.  Structural parameters (SMax, areas, altitudes, derived
 * geometry) cannot vary; derived values the reference recomputes (TTM = TT + DTTM, MaxSoilWater = FC, W0…W4 of hland_96p)
 * are the caller's to list.  hb_set_members(h, 1, 0, …) returns to the single basin.  Calling it again replaces the
 * members (the conditions are then those of member 0). */
#define HB_MOD_EPAR 1000
int hb_set_members(hb_engine* h, int64_t members, int64_t n_mod, const int32_t* mod_par, const double* mul,
                   const double* add, const uint8_t* mod_mask);

/* ---- multi-GPU boundary hand-over (NCCL over NVLink) ---------------------------------------------- */
#define HB_NCCL_ID_BYTES 128
/* Fill `id` (HB_NCCL_ID_BYTES) on rank 0; broadcast it out of band; then every rank calls hb_comm_init. */
int hb_comm_unique_id(hb_engine* h, void* id);
int hb_comm_init(hb_engine* h, int rank, int world, const void* id);
/* Send the window's simulated rows of `n` local nodes (kind HB_ROW_NODE) or the outflow rows of `n` local reaches
 * (kind HB_ROW_REACH) to rank `peer` / receive `n` rows into the external rows `ext_rows[]` from rank `peer`.  Calls
 * between matching ranks must pair up with the same `n`: the rows of a call travel as ONE message (packed by a gather
 * kernel behind the routing; received straight into the external-row buffer when `ext_rows[]` is consecutive, through a
 * staging buffer otherwise).  Both calls only enqueue (stream of their own); the routing of the window waits on the device
 * for rows that feed a reach. */
#define HB_ROW_NODE 0
#define HB_ROW_REACH 1
int hb_comm_send_rows(hb_engine* h, int peer, int kind, int64_t n, const int32_t* rows);
int hb_comm_recv_ext(hb_engine* h, int peer, int64_t n, const int32_t* ext_rows);
int hb_comm_barrier(hb_engine* h);

/* ---- measurement ------------------------------------------------------------------------------------------------ */
typedef struct hb_timing {
  double land_ms;        /* CUDA-event time of the runoff-generation kernel(s), on the land stream               */
  double river_ms;       /* CUDA-event time of all routing kernels, on the routing stream                        */
  double total_ms;       /* land_ms + river_ms                                                                   */
  int64_t land_launches;   /* kernels launched for the land part                                                */
  int64_t river_launches;  /* kernels launched for the river part                                               */
  int64_t n_levels;        /* topological levels of the river network                                           */
  double zone_ms;          /* runoff generation, fast path: thread-per-zone kernel (all chunks of the window)       */
  double element_ms;       /* runoff generation, fast path: thread-per-element kernel                               */
  double general_ms;       /* runoff generation, general kernel (elements the fast path does not cover, or series)  */
  int64_t n_fast;          /* land elements on the fast path in the last hb_run                                     */
  int64_t n_runs;          /* hb_run / hb_run_async calls the sums above cover                                      */
  /* what the fast-path kernels of hland_96 EXECUTED (device counters; the kernels take exact shortcuts in dry spells):  */
  int64_t zone_warpsteps;      /* warp-steps (32 zones x 1 step) of the branch-free FIELD/FOREST zone step               */
  int64_t zone_warpsteps_dry;  /* ... in which no zone receives water at the soil (no exp of Calc_R_SM_V1's power)       */
  int64_t zone_steps;          /* zone-steps of those warps                                                             */
  int64_t zone_steps_wet;      /* ... with in_ > 0                                                                      */
  int64_t elem_steps;          /* element-steps of the element kernel                                                   */
  int64_t elem_steps_run;      /* ... that ran the RecStep chain of Calc_Q0_Perc_UZ_V1 (not: uz == 0 and inuz == 0)     */
  int64_t substeps_pow;        /* RecStep sub-steps that evaluated the power function (of recstep * elem_steps_run)      */
  int64_t elem_warpsteps;      /* warp-steps (32 elements x 1 step) of the element kernel                               */
  int64_t elem_warpsteps_run;  /* ... in which at least one lane ran the chain                                          */
  int64_t n_fused;             /* elements whose zone terms the zone kernel added itself (HB_OPT_LAND_FUSE), last run    */
  int64_t n_spec;              /* elements with a finite SMax on the speculative fast path (HB_OPT_LAND_SPECULATE), last run */
  int64_t spec_reruns;         /* ... element-windows the general kernel had to repeat (snow above SMax somewhere)       */
} hb_timing;
/* Times and launch counts are SUMS over all runs since the previous hb_get_timing call (which resets them); after a
 * single hb_run they are that run's figures.  Waits for pending asynchronous work. */
int hb_get_timing(hb_engine* h, hb_timing* out);
/* Generic device timer on the engine's stream (CUDA events): everything enqueued between start and stop — kernels,
 * copies, NCCL sends/receives — is covered.  hb_timer_stop synchronises the stream and returns milliseconds. */
int hb_timer_start(hb_engine* h);
int hb_timer_stop(hb_engine* h, double* ms);
/* FP64 roofline denominator: sustained DFMA throughput of the device in TFLOP/s (2 flops per FMA), measured with an
 * independent-accumulator micro-benchmark on the engine's stream. */
int hb_measure_fp64_peak(hb_engine* h, double* tflops);
/* Device properties the bench reports (SM count, clock in kHz, name into buf). */
int hb_device_info(hb_engine* h, int* sm_count, int* clock_khz, char* name, int name_len);

#ifdef __cplusplus
}
#endif
#endif /* HYDPY_B200_H */
