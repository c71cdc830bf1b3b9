// land_kernel_v2.cuh — runoff generation, fast path: two persistent kernels per time chunk.
//
// Why (profiles/r01_land_v1b_fastmath_details.csv): the one-thread-per-element kernel (land_kernel.cuh) re-reads its
// 26 folded constants per zone and step from HBM (280 MB per simulated step for 10^5 elements, more than the L2), is
// limited to 1.3 waves of 128-register threads and therefore spends 65 % of its stall cycles on the long scoreboard
// with the FP64 pipe 12.8 % busy.  The decomposition below follows BASELINE.json's north_star instead:
//
//   zone_kernel    : ONE THREAD PER ZONE (x member), persistent over the chunk.  Every folded constant and the four
//                    zone states (ic, sp, wc, sm) live in registers for all steps; the only memory traffic per step is
//                    the element's forcing (4 doubles, prefetched one step ahead, shared by the element's zones through
//                    L1/L2) and two result terms per zone.  Thread (k, e) sits at k*E_pad + e, so a warp holds zone k of
//                    32 consecutive elements: all loads/stores are coalesced and ragged zone counts only retire warps.
//   element_kernel : one thread per element, persistent over the chunk: adds the zone terms in ZONE ORDER (the
//                    reference's summation order: Calc_InUZ_V1, Calc_ContriArea_V1, Calc_LZ_V1, Calc_EL_LZ_V1,
//                    Calc_InRC_V1) — unless the zone kernel has done that (below) —, then the sequential RecStep loop of
//                    Calc_Q0_Perc_UZ_V1, the lower zone and the unit hydrograph.  All 10^5 chains are resident at once
//                    (5.3 warps per scheduler), which is what hides the dependent chain of the 50 sub-steps.
//
// Round 2 (profiles/r02n_summary.txt):
//   * zone_kernel<…, ZW>: where all zones of 32 consecutive elements are FIELD/FOREST, the block is the group's zone warps
//     and adds the zone terms itself, in zone order, every four steps (north_star's in-kernel zone→element reduction):
//     16 B per element-step through HBM instead of 32 B per zone-step;
//   * the RecStep chain in a short form (exact rewrite of percolation/outflow, base-2 power function with the step's
//     invariants folded in: 15 instead of 26 dependent FP64 instructions per sub-step), chosen per element by its own
//     data; warp-uniform exits for dry sub-steps;
//   * several snow classes in the general zone step; series of the sibling models.
//
// The split is legal exactly when the zone equations do not read the element state of the previous step, i.e. when
// every CFlux of the element is zero (then Calc_CF_SM_V1 yields min(0*(1-sm/fc), uz+r, fc-sm) = the value without the
// uz term, because uz >= 0 and r >= 0; ref: hydpy/models/hland/hland_model.py:1904-1919) — the Lahn default and the
// synthetic configs.  Elements with CFlux > 0 (land_kernel_coupled.cuh) or a finite SMax (snow redistribution) run
// elsewhere; land_kernel.cuh covers everything.
//
// Numerics: same formulas and operand order as land_kernel.cuh; deliberate deviations (each <= 1 ulp per operation,
// inside the 1e-9 parity bar, measured in tests/test_gpu_*): divisions by the step-invariant constants RfCF, SfCF,
// TTInt, FC and SoilMoistureLimit*MaxSoilWater are multiplications by reciprocals folded on the host.
#pragma once
#include <cstdint>

#include "land_kernel.cuh"

namespace hb {

struct LandV2Dev {
  int64_t n_elem, e_pad;
  int32_t zmax;
  // slab of elements this launch works on (multiples of 128; the whole basin: e_lo = 0, slab_blocks = E_pad / 128).
  // Slabs run on streams of their own, so that the latency-bound element kernel of one slab shares the SMs with the
  // throughput-bound zone kernel of another (engine.cu: run_land)
  int64_t e_lo;
  int32_t slab_blocks;
  const int32_t* nz;       // [E_pad]
  const int32_t* nc;       // [E_pad] snow classes (fast path: class 0 in registers, further classes [c][zmax][E_pad] in place)
  int32_t skip_classes;    // this run records series: elements with several snow classes (and speculative ones) go to
                           // the general kernel
  const int32_t* nu;       // [E_pad]
  const int32_t* recstep;  // [E_pad]
  const int32_t* einfo;    // [E_pad]  HB_EI_* incl. HB_EI_FAST / HB_EI_SOILONLY
  const int32_t* fcol;     // [E_pad]
  const int32_t* zinfo;    // [zmax][E_pad]
  const double* kz;        // [zmax][KZ_COUNT][E_pad]
  const double* ke;        // [KE_COUNT][E_pad]
  const double* sfdist;    // [cmax][E_pad] (class 0 used)
  const double* uh;        // [umax][E_pad]
  double* ic;              // [zmax][E_pad]
  double* sm;
  double* sp;              // class 0
  double* wc;
  double* uz;              // [E_pad]
  double* lz;
  double* quh;             // [umax][E_pad]
  int64_t nt;              // steps of this chunk
  int64_t n_col, nt_total, t_first;  // forcing block [4][nt_total][n_col]; first step of the chunk inside it
  const double* forcing;
  const double* sinfactor;  // [nt_total]
  const MathTables* tables;
  const ChainTables* chain_tables;   // RecStep chain (hb_chain_pow2), staged by the hland_96 element kernel
  double* term_a;          // [nt][zmax][E_pad]
  double* term_b;          // [nt][zmax][E_pad]
  // fused reduction (zone_kernel<…, ZW>): the zone terms of an element are added IN ZONE ORDER inside the zone kernel and
  // leave it as two sums per element-step — 16 B instead of 16 B per ZONE-step through HBM
  double* red_a;           // [nt][E_pad] InUZ (Calc_InUZ_V1)
  double* red_b;           // [nt][E_pad] sum of weight * log(sm/fc) (Calc_ContriArea_V1)
  const int32_t* gwarps;   // [E_pad/32] zone warps of the 32-element group (max zone count) if the group is fused, else 0
  int32_t fused;           // this run reduces the fused groups in the zone kernel (they carry HB_EI_FUSED)
  double* qt;              // [nt_window][E_pad], chunk starts at row t_out
  int64_t t_out;
  // optional series (SERIES kernels).  Element level: ABI layout [nt_window][n_elem], written in place.  Zone and
  // snow level: the zone kernel writes a chunk-local staging buffer in ITS layout [nt][zmax][E_pad] (coalesced);
  // series_scatter_kernel then moves the chunk into the ABI layout [nt_window][n_zone] with coalesced stores.
  int64_t n_zone, n_snow, n_uh;
  const int64_t* uh0;      // [E_pad] first storage of the element in ABI order (series HB_S_SC)
  double* series[HB_S_COUNT];
  // fused kernel for elements with capillary flux (land_kernel_coupled.cuh)
  const int32_t* clist;    // their indices
  int64_t n_clist;
  // sibling models (hland_96p / hland_96c) and rconc_nash
  const int32_t* nash_steps;  // [E_pad] see LandDev
  int32_t nash_stage;      // element kernel: dynamic shared memory for the cascades is present
  int32_t uh_stage;        // … with room for this many rconc_uh ordinates per element (>= HB_NASH_STAGE)
  int32_t uh_weights;      // … and for their weights behind them ([uh_stage][64] memory, [uh_stage][64] weights)
  double* zx1;             // [zmax][E_pad] hland_96p suz | hland_96c bw1
  double* zx2;             // [zmax][E_pad] hland_96p sg1 | hland_96c bw2
  double* ex1;             // [E_pad] hland_96p sg2
  double* ex2;             // [E_pad] hland_96p sg3
  // what the kernels actually executed (hb_get_timing): see enum below; one atomicAdd per warp and counter at kernel end
  unsigned long long* counters;
  int32_t* spec_flag;      // [E_pad] HB_EI_SPEC elements: 1 = set by the zone kernel when a snow class exceeded its SMax in this
                           // window; 2 = left by the general kernel's repetition of the PREVIOUS window when the element lost
                           // snow there (the fast path then skips the element: it would only be repeated again); else 0
};

// slots of LandV2Dev::counters
enum { HB_CNT_ZONE_WARPSTEPS = 0,   // warp-steps of the branch-free (FIELD/FOREST) zone step
       HB_CNT_ZONE_WARPSTEPS_DRY,   // … that took the dry shortcut (no water at the soil of any of the 32 zones)
       HB_CNT_ZONE_STEPS,           // zone-steps of those warps (active lanes)
       HB_CNT_ZONE_STEPS_WET,       // … with in_ > 0 (water reaches the soil: Calc_R_SM_V1 pays for its power function)
       HB_CNT_ELEM_STEPS,           // element-steps of the element kernel (hland_96)
       HB_CNT_ELEM_STEPS_RUN,       // … that ran the RecStep chain of Calc_Q0_Perc_UZ_V1
       HB_CNT_SUBSTEPS_POW,         // RecStep sub-steps that evaluated the power function, per element (short form: every
                                    // sub-step its loop runs; general loop: those with uz > 0; a quadratic response: none)
       HB_CNT_ELEM_WARPSTEPS,       // warp-steps of the element kernel (32 elements x 1 step)
       HB_CNT_ELEM_WARPSTEPS_RUN,   // … in which at least one lane ran the chain (what the FP64 pipe is busy with)
       HB_CNT_SPEC_RERUN,           // element-windows of the speculative fast path (HB_EI_SPEC) the general kernel repeated
       HB_CNT_COUNT };

__device__ __forceinline__ void stage_tables(MathTables* s, const MathTables* g) {
  const double* src = reinterpret_cast<const double*>(g);
  double* dst = reinterpret_cast<double*>(s);
  for (int i = threadIdx.x; i < (int)(sizeof(MathTables) / sizeof(double)); i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// hland_96p, one FIELD/FOREST/GLACIER zone: Calc_SUZ_V1, Calc_DP_SUZ_V1, Calc_RS_RI_SUZ_V1, Calc_GR1_V1, Calc_RG1_SG1_V1
// (ref: hland_model.py:2132-2140, 2525-2535, 3012-3033, 3233-3243, 3290-3304).  fast = rs + ri + rg1 (Calc_InRC_V2 term),
// deep = dp - gr1 (Calc_GR2_GR3_V1 term).
__device__ __forceinline__ void response_96p(double r, double percmax, double sgr, double omw0, double omw1, double k2,
                                             double w2, double omw2k2, double sg1max, double& suz, double& sg1,
                                             double& dp, double& rs, double& ri, double& gr1, double& rg1) {
  suz += r;
  dp = HB_PYMIN(suz, percmax);
  suz -= dp;
  rs = (suz > sgr) ? (suz - sgr) * omw0 : 0.0;
  ri = suz * omw1;
  suz -= rs + ri;
  if (suz < 0.0) {
    const double f = 1.0 - suz / (rs + ri);
    rs *= f;
    ri *= f;
    suz = 0.0;
  }
  const double sg1_old = sg1;
  const double b = (sg1max - sg1_old) / k2;
  gr1 = HB_PYMIN(dp, b);
  const double a = sg1_old + gr1 - sg1max;
  gr1 -= HB_PYMAX(a, 0.0);
  sg1 = w2 * sg1_old + omw2k2 * gr1;
  rg1 = sg1_old + gr1 - sg1;
}

// ---------------------------------------------------------------------------------------------------------------------
// zone kernel: everything of hland_96.run() that works on one zone (ref: hydpy/models/hland_96.py:1675-1705, the methods
// Calc_TC_V1 … Calc_EA_SM_AETModel_V1 and the per-zone factors of Calc_InUZ_V1 / Calc_ContriArea_V1 / Calc_LZ_V1 /
// Calc_EL_LZ_V1 / Calc_InRC_V1)
// ---------------------------------------------------------------------------------------------------------------------
#ifndef HB_ZONE_BLOCKS
#define HB_ZONE_BLOCKS 6
#endif
// SOIL = true : the warps whose 32 zones are all FIELD/FOREST with FC > 0 (branch-free step); SOIL = false: all other warps
// SERIES = true additionally records the requested hland_96 zone sequences (the HBM-bound regime of SURVEY.md §8d-ii)
// MODEL = HB_MODEL_HLAND_96 / _96P / _96C: every instance works on the fast-path elements of its model only (einfo carries
// the model).  The sibling models share everything up to the soil routine; their zone response (hland_96p: SUZ, SG1;
// hland_96c: BW1, BW2) replaces the contributing-area factor, needs no pow() and seven more constants (5 blocks per SM).
// ZW > 0 (SOIL, no SERIES, hland_96 only): FUSED REDUCTION.  A block is the ZW zone warps of ONE group of 32 consecutive
// elements (warp k = zone k, lane = element) instead of one zone of 128 elements.  The warps leave the two terms of G
// consecutive steps in shared memory; then — between two block barriers — 2 G warps add them per element in ZONE ORDER
// (the reference's order of additions in Calc_InUZ_V1 / Calc_ContriArea_V1, same bits as the element kernel's loop over
// the term buffers), one thread per (step, sum, element), and store two sums per element-step.  HBM traffic of the
// hand-over: 16 B per element-step instead of 2 x 16 B per zone-step.  (Measured: a barrier and a reduction every step
// — two warps adding while the others run ahead into the next barrier — cost the zone kernel 30 %; every G = 4 steps 2 %.)
template <int ZW> struct ZoneGeom {
  static constexpr int threads = ZW > 0 ? ZW * 32 : 128;
  static constexpr int blocks = ZW > 0 ? (24 / ZW > 0 ? 24 / ZW : 1) : HB_ZONE_BLOCKS;
  static constexpr int G = ZW == 6 ? 3 : 4;      // steps per reduction (shared memory: ZW*32*(224 + 16 G) bytes per block)
};
template <bool SOIL, bool SERIES, int MODEL, int ZW = 0>
__global__ void __launch_bounds__(ZoneGeom<ZW>::threads, MODEL == HB_MODEL_HLAND_96 ? ZoneGeom<ZW>::blocks : 5) zone_kernel(const LandV2Dev d) {
  static_assert(ZW == 0 || (SOIL && !SERIES && MODEL == HB_MODEL_HLAND_96), "fused reduction: branch-free hland_96 step only");
  constexpr int NTH = ZoneGeom<ZW>::threads;
  __shared__ MathTables s_tables;
  stage_tables(&s_tables, d.tables);
  const TabRef T = hb_tabref(&s_tables);
  const int64_t ep = d.e_pad;
  const int lane = threadIdx.x & 31;
  int k, gw = 0;          // zone index; fused: zone warps of this group
  int64_t e;
  if (ZW > 0) {
    const int64_t g = d.e_lo / 32 + blockIdx.x;
    gw = d.gwarps[g];
    k = threadIdx.x >> 5;
    e = g * 32 + lane;
    if (k >= gw) return;                  // group not fused (gw = 0), or a warp beyond its zone counts: whole warps leave
  } else {
    k = (int)(blockIdx.x / (unsigned)d.slab_blocks);
    e = d.e_lo + (int64_t)(blockIdx.x - (unsigned)k * (unsigned)d.slab_blocks) * 128 + threadIdx.x;
  }
  // fused: the lanes of a live warp stay for the block barriers; `active` lanes own a zone.  A lane without a zone (its
  // element has fewer zones than the group's largest, or lies beyond the basin) shadows zone 0 of its element (of the
  // group's first element): the same instructions on valid data, nothing stored — cheaper than a branch around the step,
  // which costs ~25 register moves per step at its join.  (The warp votes of the step only choose between code paths
  // with identical results, so shadows cannot change a bit.)
  const int nz_e = (ZW > 0) ? ((e < d.n_elem) ? d.nz[e] : 0) : 0;
  const bool active = (ZW > 0) ? (k < nz_e) : true;
  const int64_t e_own = e;
  if (ZW > 0 && !active) { if (e >= d.n_elem) e = e - lane; k = 0; }
  const int64_t slot = (int64_t)k * ep + e;   // k*E_pad + e
  if (ZW == 0) {
    if (e >= d.n_elem) return;
    const int einfo0 = d.einfo[e];
    if (!(einfo0 & HB_EI_FAST) || HB_EI_MODEL(einfo0) != MODEL || k >= d.nz[e]) return;
    if (d.fused && (einfo0 & HB_EI_FUSED)) return;   // reduced by the fused instance of this run
  }
  const int einfo = d.einfo[e];
  // several snow classes (hland_96 only, never in a fused group): the general step below loops over them
  const int ncls = (ZW == 0 && MODEL == HB_MODEL_HLAND_96) ? d.nc[e] : 1;
  // finite SMax (hland_96 only, never in a fused group): the general step below watches the limit (see HB_EI_SPEC)
  const bool spec = ZW == 0 && MODEL == HB_MODEL_HLAND_96 && (einfo & HB_EI_SPEC) != 0;
  if (ZW == 0 && (ncls > 1 || spec) && d.skip_classes) return;
  // (2: the element lost snow at its limit in the previous window — it goes straight to the general kernel in this one)
  if (spec && d.spec_flag[e] == 2) return;

  const double* kzk = d.kz + (int64_t)k * KZ_COUNT * ep + e;
#define KZ(name) kzk[(int64_t)(name) * ep]
  const int zi = d.zinfo[slot];
  const int zt = HB_ZI_TYPE(zi), zf = HB_ZI_FLAGS(zi);
  const bool soil = (zt == HB_FIELD) || (zt == HB_FOREST);
  const bool icept = soil || (zt == HB_SEALED);        // zone types with an interception storage
  const bool resp = (einfo & HB_EI_RESPAREA) != 0;
  // ---- constants (coalesced loads: consecutive lanes = consecutive elements) ----
  const double tcorr = KZ(KZ_TCORR), tcalt_dz = KZ(KZ_TCALT_DZ), tt_hi = KZ(KZ_TT_HI), tt_lo = KZ(KZ_TT_LO),
               inv_ttint = KZ(KZ_INV_TTINT), pcalt_f = KZ(KZ_PCALT_F), pcorr = KZ(KZ_PCORR), inv_rfcf = KZ(KZ_INV_RFCF),
               inv_sfcf = KZ(KZ_INV_SFCF), icmax = KZ(KZ_ICMAX), cfmax = KZ(KZ_CFMAX), cfvar = KZ(KZ_CFVAR),
               ttm = KZ(KZ_TTM), cfr_cfmax = KZ(KZ_CFR_CFMAX), whc = KZ(KZ_WHC), fc = KZ(KZ_FC), inv_fc = KZ(KZ_INV_FC),
               beta = KZ(KZ_BETA), thresh = KZ(KZ_THRESH), inv_thresh = KZ(KZ_INV_THRESH), excessred = KZ(KZ_EXCESSRED),
               atf = KZ(KZ_ATF), etf = KZ(KZ_ETF), alt_f = KZ(KZ_ALT_F), neg_pf = KZ(KZ_NEG_PF);
  // type-dependent constants share registers
  // hland_96:  term_a = w_a * (r - cf) | pc | r,  term_b = weight * log(sm/fc) (contributing area) | w_b * el
  // hland_96p: term_a = relzonearea * (rs + ri + rg1) | relzonearea * r (SEALED) | el (ILAKE),  term_b = w_b * (dp - gr1) | w_b * pc
  // hland_96c: term_a = w_land * (qab1 + qab2) | w_land * r (SEALED) | el (ILAKE),              term_b = w_b * qvs2 | w_b * pc
  const double w_a = MODEL == HB_MODEL_HLAND_96P   ? KZ(KZ_RELZONEAREA)
                     : MODEL == HB_MODEL_HLAND_96C ? KZ(KZ_W_LAND)
                     : soil || zt == HB_GLACIER    ? KZ(KZ_W_UZ)
                     : zt == HB_ILAKE              ? KZ(KZ_W_LZ)
                                                   : KZ(KZ_RELZONEAREA);
  const double w_b = (MODEL == HB_MODEL_HLAND_96 && soil) ? KZ(KZ_W_CONTRI) : KZ(KZ_W_LZ);
  const bool upper = soil || zt == HB_GLACIER;
  const double g_melt = zt == HB_GLACIER ? KZ(KZ_GMELT) : 0.0, g_var = zt == HB_GLACIER ? KZ(KZ_GVAR) : 0.0;
  const double tti = zt == HB_ILAKE ? KZ(KZ_TTI) : 0.0;
#undef KZ
#define KZX(name) kzk[(int64_t)(name) * ep]
  const double sf = d.sfdist[e];
  double ic = d.ic[slot], sm = d.sm[slot], sp = d.sp[slot], wc = d.wc[slot];
  double zx1 = 0.0, zx2 = 0.0;
  if (MODEL != HB_MODEL_HLAND_96) { zx1 = d.zx1[slot]; zx2 = d.zx2[slot]; }
  // The constants move on to shared memory ([constant][thread], conflict-free, 29 KB per block): 72 registers per thread
  // instead of 128, 7 instead of 4 warps per scheduler, and room for the routing blocks of the previous window beside them.
  // (ncu: the step is then issue-bound — ≈410 warp instructions per zone-step at ≈80 % issue-slot utilisation.)
  // (constants that are used together share a 16-byte slot: one LDS.128 fetches both)
  extern __shared__ double2 s_dyn[];      // fused: constants [14][NTH] + term buffers [2][2][ZW][32] doubles
  __shared__ double2 s_static[ZW > 0 ? 1 : (MODEL == HB_MODEL_HLAND_96 ? 28 : 36) / 2][ZW > 0 ? 1 : 128];
  double2 (*const s_c)[NTH] = ZW > 0 ? reinterpret_cast<double2 (*)[NTH]>(s_dyn) : reinterpret_cast<double2 (*)[NTH]>(s_static);
  double* const s_terms = reinterpret_cast<double*>(s_dyn + 14 * NTH);   // (fused only) [G][2][ZW][32]
  constexpr int G = ZoneGeom<ZW>::G;
  if (MODEL != HB_MODEL_HLAND_96) {
    s_c[14][threadIdx.x].x = KZX(KZ_X0); s_c[14][threadIdx.x].y = KZX(KZ_X1); s_c[15][threadIdx.x].x = KZX(KZ_X2);
    s_c[15][threadIdx.x].y = KZX(KZ_X3); s_c[16][threadIdx.x].x = KZX(KZ_X4); s_c[16][threadIdx.x].y = KZX(KZ_X5);
    s_c[17][threadIdx.x].x = KZX(KZ_X6);
    s_c[17][threadIdx.x].y = d.ke[KE_DT_PERCMAX * ep + e];   // hland_96p: PercMax (dt = 1)
  }
#define kx0 s_c[14][threadIdx.x].x
#define kx1 s_c[14][threadIdx.x].y
#define kx2 s_c[15][threadIdx.x].x
#define kx3 s_c[15][threadIdx.x].y
#define kx4 s_c[16][threadIdx.x].x
#define kx5 s_c[16][threadIdx.x].y
#define kx6 s_c[17][threadIdx.x].x
#define kpercmax s_c[17][threadIdx.x].y
  s_c[0][threadIdx.x].x = tcorr;
  s_c[0][threadIdx.x].y = tcalt_dz;
  s_c[1][threadIdx.x].x = tt_hi;
  s_c[1][threadIdx.x].y = tt_lo;
  s_c[2][threadIdx.x].x = inv_ttint;
  s_c[2][threadIdx.x].y = pcalt_f;
  s_c[3][threadIdx.x].x = pcorr;
  s_c[3][threadIdx.x].y = inv_rfcf;
  s_c[4][threadIdx.x].x = inv_sfcf;
  s_c[4][threadIdx.x].y = icmax;
  s_c[5][threadIdx.x].x = cfmax;
  s_c[5][threadIdx.x].y = cfvar;
  s_c[6][threadIdx.x].x = ttm;
  s_c[6][threadIdx.x].y = cfr_cfmax;
  s_c[7][threadIdx.x].x = whc;
  s_c[7][threadIdx.x].y = fc;
  s_c[8][threadIdx.x].x = inv_fc;
  s_c[8][threadIdx.x].y = beta;
  s_c[9][threadIdx.x].x = thresh;
  s_c[9][threadIdx.x].y = inv_thresh;
  s_c[10][threadIdx.x].x = excessred;
  s_c[10][threadIdx.x].y = atf;
  s_c[11][threadIdx.x].x = etf;
  s_c[11][threadIdx.x].y = alt_f;
  s_c[12][threadIdx.x].x = neg_pf;
  s_c[12][threadIdx.x].y = w_a;
  s_c[13][threadIdx.x].x = w_b;
  s_c[13][threadIdx.x].y = sf;
#define tcorr s_c[0][threadIdx.x].x
#define tcalt_dz s_c[0][threadIdx.x].y
#define tt_hi s_c[1][threadIdx.x].x
#define tt_lo s_c[1][threadIdx.x].y
#define inv_ttint s_c[2][threadIdx.x].x
#define pcalt_f s_c[2][threadIdx.x].y
#define pcorr s_c[3][threadIdx.x].x
#define inv_rfcf s_c[3][threadIdx.x].y
#define inv_sfcf s_c[4][threadIdx.x].x
#define icmax s_c[4][threadIdx.x].y
#define cfmax s_c[5][threadIdx.x].x
#define cfvar s_c[5][threadIdx.x].y
#define ttm s_c[6][threadIdx.x].x
#define cfr_cfmax s_c[6][threadIdx.x].y
#define whc s_c[7][threadIdx.x].x
#define fc s_c[7][threadIdx.x].y
#define inv_fc s_c[8][threadIdx.x].x
#define beta s_c[8][threadIdx.x].y
#define thresh s_c[9][threadIdx.x].x
#define inv_thresh s_c[9][threadIdx.x].y
#define excessred s_c[10][threadIdx.x].x
#define atf s_c[10][threadIdx.x].y
#define etf s_c[11][threadIdx.x].x
#define alt_f s_c[11][threadIdx.x].y
#define neg_pf s_c[12][threadIdx.x].x
#define w_a s_c[12][threadIdx.x].y
#define w_b s_c[13][threadIdx.x].x
#define sf s_c[13][threadIdx.x].y

  const int64_t col = d.fcol[e];
  // all per-step addresses advance by pointer bumps (no 64-bit multiplies in the loop)
  const double* fP = d.forcing + (0 * d.nt_total + d.t_first) * d.n_col + col;
  const double* fT = d.forcing + (1 * d.nt_total + d.t_first) * d.n_col + col;
  const double* fA = d.forcing + (2 * d.nt_total + d.t_first) * d.n_col + col;
  const double* fE = d.forcing + (3 * d.nt_total + d.t_first) * d.n_col + col;
  const double* sinp = d.sinfactor + d.t_first;
  const int64_t tstride = (int64_t)d.zmax * ep;
  const int64_t fstride = d.n_col;
  double* out_a = d.term_a + slot;
  double* out_b = d.term_b + slot;

  double n_p = 0.0, n_t = 0.0, n_nat = 0.0, n_net = 0.0, n_sin = 0.0;
  if (d.nt > 0) { n_p = ldg(fP); n_t = ldg(fT); n_nat = ldg(fA); n_net = ldg(fE); n_sin = ldg(sinp); }
#define HB_ZONE_STEP_BEGIN                                                                                    \
  const double in_p = n_p, in_t = n_t, in_nat = n_nat, in_net = n_net, sinf = n_sin;                             \
  if (t + 1 < nt) { /* prefetch the next step's forcing; consumed one iteration later */                         \
    fP += fstride; fT += fstride; fA += fstride; fE += fstride; ++sinp;                                          \
    n_p = ldg(fP); n_t = ldg(fT); n_nat = ldg(fA); n_net = ldg(fE); n_sin = ldg(sinp);                           \
  }
#define HB_ZONE_STEP_END(va, vb)                                                                   \
  if (ZW > 0) { s_terms[tb_off + zw_off] = (va); s_terms[tb_off + ZW * 32 + zw_off] = (vb); }       \
  else { *out_a = (va); *out_b = (vb); out_a += tstride; out_b += tstride; }
  const int nt = (int)d.nt;
#define SERZ(id, v) do { if (SERIES && d.series[id]) d.series[id][(int64_t)t * tstride + slot] = (v); } while (0)
#define SERS(id, v) SERZ(id, v)

  // ---- warps whose 32 zones are all FIELD/FOREST with FC > 0 (the bulk of any basin) run a BRANCH-FREE step: every
  // data-dependent `if` of the reference becomes a select on values that are computed unconditionally.  Results are
  // the same bit for bit (x - 0.0 == x, x * 1 == x …); the point is one large basic block in which the scheduler can
  // interleave the independent chains (snow, exp of PET, pow of R, pow of ContriArea) — with 4 warps per scheduler the
  // kernel is latency-bound otherwise (ncu: 62 % issue utilisation, half of the stall cycles are fixed-latency waits).
  const unsigned amask = __activemask();                                             // (fused: all 32 lanes)
  const unsigned omask = ZW > 0 ? __ballot_sync(0xffffffffu, active) : amask;         // lanes that own a zone
  if (ZW == 0 && __all_sync(amask, soil && fc > 0.0 && ncls == 1 && !spec) != SOIL) return;
  // fused: offsets into the term buffers [G][2][ZW][32]
  const int zw_off = (int)(threadIdx.x >> 5) * 32 + lane;   // (a shadow lane writes its own, never read, cell)
  int tb_off = 0, gs = 0;
  if (SOIL) {
    const bool f_icept = (zf & HB_ZF_INTERCEPTION) != 0, f_soil = (zf & HB_ZF_SOIL) != 0;
    // hland_96: log(sm/fc) of the soil moisture at the end of a step serves the contributing area of that step AND the
    // power function of Calc_R_SM_V1 of the next one (same argument, same bits as two evaluations)
    double lgn = (MODEL == HB_MODEL_HLAND_96) ? hb_log_normal(sm * inv_fc, T) : 0.0;
    unsigned n_dry = 0, n_wet = 0;   // warp-uniform: dry warp-steps, lanes with in_ > 0 (summed over the steps)
    for (int t = 0; t < nt; ++t) {
      HB_ZONE_STEP_BEGIN
      // ref: hland_model.py:71-77 Calc_TC_V1, 125-135 Calc_FracRain_V1, 213-225 Calc_PC_V1
      const double tc = in_t + tcorr - tcalt_dz;
      double fracrain = (tc - tt_lo) * inv_ttint;
      fracrain = (tc <= tt_lo) ? 0.0 : fracrain;
      fracrain = (tc >= tt_hi) ? 1.0 : fracrain;
      const double pc0 = in_p * pcalt_f;
      // pcorr/x as pcorr * rcp(x), rcp correctly rounded (<= 1 ulp from the division, a third of its instructions)
      double pc = pc0 * (pcorr * __drcp_rn(fracrain * inv_rfcf + (1.0 - fracrain) * inv_sfcf));
      pc = (pc0 <= 0.0) ? 0.0 : pc;
      // ref: hland_model.py:299-309 Calc_TF_Ic_V1
      const double a = pc - (icmax - ic);
      const double tf = HB_PYMAX(a, 0.0);
      ic += pc - tf;
      SERZ(HB_S_INTERCEPTEDWATER, ic);   // evap_aet_hbv96 factors.InterceptedWater: what the submodel is shown
      // ref: hland_model.py:484-499 Calc_SP_WC_V1, 1050-1062 Calc_CFAct_V1
      wc += sf * (tf * fracrain);
      sp += sf * (tf * (1.0 - fracrain));
      const double cfa = cfmax + sinf * cfvar;
      const double cfact = HB_PYMAX(cfa, 0.0);
      // ref: hland_model.py:1166-1187 Calc_Melt_SP_WC_V1, 1318-1341 Calc_Refr_SP_WC_V1
      {
        const double potmelt = cfact * (tc - ttm);
        double melt = HB_PYMIN(potmelt, sp);
        melt = (tc > ttm) ? melt : 0.0;
        sp -= melt; wc += melt;
        const double potrefr = cfr_cfmax * (ttm - tc);
        double refr = HB_PYMIN(potrefr, wc);
        refr = (tc < ttm) ? refr : 0.0;
        sp += refr; wc -= refr;
        SERS(HB_S_MELT, melt); SERS(HB_S_REFR, refr);
      }
      // ref: hland_model.py:1434-1448 Calc_In_WC_V1
      const double wc_old = wc, lim = whc * sp;
      wc = HB_PYMIN(wc_old, lim);
      const double in_ = wc_old - wc;
      const double ncover = (sp > 0.0) ? 1.0 : 0.0;
      if (SERIES) {
        SERZ(HB_S_TC, tc); SERZ(HB_S_FRACRAIN, fracrain); SERZ(HB_S_PC, pc); SERZ(HB_S_TF, tf); SERZ(HB_S_CFACT, cfact);
        SERZ(HB_S_GACT, 0.0); SERZ(HB_S_SPL, 0.0); SERZ(HB_S_WCL, 0.0); SERZ(HB_S_SPG, 0.0); SERZ(HB_S_WCG, 0.0);
        SERZ(HB_S_GLMELT, 0.0); SERZ(HB_S_SR, 0.0); SERZ(HB_S_EL, 0.0); SERZ(HB_S_IN, in_);
        SERS(HB_S_SWE, sp + wc); SERS(HB_S_SP, sp); SERS(HB_S_WC, wc);   // ref: hland_model.py:1485-1495 Calc_SWE_V1
      }
      // evap_pet_hbv96.run(): ref: evap_model.py:5111-5125, 5420-5424, 5312-5326
      double pet;
      {
        double ret = in_net * (1.0 + atf * (in_t - in_nat));
        const double lo = HB_PYMAX(ret, 0.0);
        const double hi = 2.0 * in_net;
        ret = HB_PYMIN(lo, hi);
        ret *= etf;
        pet = ret * alt_f;
        const double damp = hb_exp(neg_pf * pc, T);
        pet = (pet <= 0.0) ? 0.0 : pet * damp;
        SERZ(HB_S_REFERENCEEVAPOTRANSPIRATION, ret); SERZ(HB_S_POTENTIALEVAPOTRANSPIRATION, pet);
      }
      // ref: evap_model.py:6068-6086 Calc_InterceptionEvaporation_V1; hland_model.py:383-396 Calc_EI_Ic_AETModel_V1
      double aet_ie = HB_PYMIN(pet, ic);
      aet_ie = f_icept ? aet_ie : 0.0;
      if (SERIES) {
        SERZ(HB_S_INTERCEPTIONEVAPORATION, aet_ie); SERZ(HB_S_SNOWCOVER, ncover);
        // ref: evap_model.py:5673-5681 Calc_WaterEvaporation_V1 (all zones, whatever their type)
        SERZ(HB_S_WATEREVAPORATION, ((zf & HB_ZF_WATER) && tc > KZX(KZ_TTI)) ? pet : 0.0);
      }
      {
        const double ei = HB_PYMIN(aet_ie, ic);
        ic -= ei;
        SERZ(HB_S_EI, ei); SERZ(HB_S_IC, ic);
      }
      // ref: hland_model.py:1776-1790 Calc_R_SM_V1
      double r;
      if (MODEL == HB_MODEL_HLAND_96) {
        // nothing reaches the soil of any zone of this warp (dry hour, or the canopy holds the rain back): in_ * pow(...)
        // is in_ itself — 0.0 times a finite power (0 < sm/fc <= 1) — and the exp half of the power function can stay out
        const double x = sm * inv_fc;
        const unsigned none = __ballot_sync(amask, in_ == 0.0);   // (in_ >= 0 or NaN: every other lane counts as wet)
        n_wet += __popc(omask & ~none);
        if (none == amask && __all_sync(amask, x > 0.0 && x <= 1.0)) { r = in_; ++n_dry; }
        else r = in_ * hb_pow_from_log(x, lgn, beta, T);
      } else r = in_ * hb_pow(sm * inv_fc, beta, T);
      {
        const double a2 = sm + in_ - fc;
        r = HB_PYMAX(r, a2);
      }
      sm += in_ - r;
      // ref: hland_model.py:1904-1919 Calc_CF_SM_V1 with CFlux == 0 (hland_96 only)
      double cf = 0.0;
      if (MODEL == HB_MODEL_HLAND_96) {
        cf = 0.0 * (1.0 - sm * inv_fc);
        const double b = fc - sm;
        cf = HB_PYMIN(cf, b);
        sm += cf;
      }
      SERZ(HB_S_SOILWATER, sm);          // evap_aet_hbv96 factors.SoilWater
      // ref: evap_model.py:7630-7658 Determine_SoilEvapotranspiration_V1 (6657-6672, 7001-7018, 7060-7069)
      double set = pet;
      {
        double s1 = set * (sm * inv_thresh);
        s1 = (sm < thresh) ? s1 : set;
        s1 = (sm < 0.0) ? 0.0 : s1;
        set = (set > 0.0) ? s1 : set;
        const double petc = excessred * pet + (1.0 - excessred) * (pet + pet) / 2.0;
        const double excess = set + aet_ie - petc;
        const bool adjust = ((petc >= 0.0) && (excess > 0.0)) || ((petc < 0.0) && (excess < 0.0));
        const double adjusted = set - excessred * excess;
        set = adjust ? adjusted : set;
        set *= 1.0 - ncover;
        set = f_soil ? set : 0.0;
      }
      SERZ(HB_S_SOILEVAPOTRANSPIRATION, set);
      // ref: hland_model.py:2002-2018 Calc_EA_SM_AETModel_V1
      const double ea = HB_PYMIN(set, sm);
      sm -= ea;
      {
        const bool over = sm > fc;
        const double r_over = r + (sm - fc);
        r = over ? r_over : r;
        sm = over ? fc : sm;
      }
      SERZ(HB_S_R, r); SERZ(HB_S_CF, cf); SERZ(HB_S_EA, ea); SERZ(HB_S_SM, sm);
      if (MODEL == HB_MODEL_HLAND_96) {
        // ref: hland_model.py:2092-2101 Calc_InUZ_V1 (term), 2224-2236 Calc_ContriArea_V1: the product of the zones'
        // (sm/fc)^weight, raised to beta, is evaluated as exp(beta * sum(weight * log(sm/fc))) — one log per zone-step here,
        // one exp per element-step in the element kernel instead of a pow (log + exp) per zone-step (-26 of 386 warp
        // instructions per zone-step); relative deviation from the product form <= 1e-14
        const double x_end = sm * inv_fc;
        lgn = hb_log_normal(x_end, T);
        const double contri = w_b * hb_log_from_normal(x_end, lgn);
        HB_ZONE_STEP_END(w_a * (r - cf), resp ? contri : 0.0)
      } else if (MODEL == HB_MODEL_HLAND_96P) {
        double dp, rs, ri, gr1, rg1;
        response_96p(r, kpercmax, kx0, kx1, kx2, kx3, kx4, kx5, kx6, zx1, zx2, dp, rs, ri, gr1, rg1);
        SERZ(HB_S_DP, dp); SERZ(HB_S_RS, rs); SERZ(HB_S_RI, ri); SERZ(HB_S_GR1, gr1); SERZ(HB_S_RG1, rg1);
        SERZ(HB_S_SUZ, zx1); SERZ(HB_S_SG1, zx2);
        HB_ZONE_STEP_END(w_a * (rs + ri + rg1), w_b * (dp - gr1))
      } else {
        // ref: hland_model.py:2833-2853 Calc_QAb1_QVs1_BW1_V1, 2927-2947 Calc_QAb2_QVs2_BW2_V1
        double qab1 = 0.0, qvs1 = 0.0, qab2 = 0.0, qvs2 = 0.0;
        qab_qvs_bw_t<true>(kx0, kx1, kx2, zx1, r, qab1, qvs1, T);
        qab_qvs_bw_t<true>(kx3, kx4, kx5, zx2, qvs1, qab2, qvs2, T);
        SERZ(HB_S_QAB1, qab1); SERZ(HB_S_QVS1, qvs1); SERZ(HB_S_QAB2, qab2); SERZ(HB_S_QVS2, qvs2);
        SERZ(HB_S_BW1, zx1); SERZ(HB_S_BW2, zx2);
        HB_ZONE_STEP_END(w_a * (qab1 + qab2), w_b * qvs2)
      }
      if (ZW > 0) {
        tb_off += 2 * ZW * 32;
        if (++gs == G || t == nt - 1) {
          // every lane of every live warp: the terms of the last gs steps are complete
          __syncwarp();
          asm volatile("bar.sync 1, %0;" ::"r"(gw * 32) : "memory");
          // warp w2 < 2 gs: step w2 / 2 of the batch, sum w2 & 1 (0: Calc_InUZ_V1, ref: hland_model.py:2092-2101; 1: the
          // weighted logarithms of Calc_ContriArea_V1, :2224-2236), lane = element, zones in their order
          for (int w2 = (int)(threadIdx.x >> 5); w2 < 2 * gs; w2 += gw) {
            const double* src = s_terms + (w2 >> 1) * (2 * ZW * 32) + (w2 & 1) * (ZW * 32) + lane;
            double sum = 0.0;
            if (__all_sync(0xffffffffu, nz_e == ZW)) {   // the usual case: all look-ups in flight at once
              double v[ZW > 0 ? ZW : 1];
#pragma unroll
              for (int kk = 0; kk < ZW; ++kk) v[kk] = src[kk * 32];
#pragma unroll
              for (int kk = 0; kk < ZW; ++kk) sum += v[kk];
            } else
              for (int kk = 0; kk < nz_e; ++kk) sum += src[kk * 32];
            if (e_own < d.n_elem) ((w2 & 1) ? d.red_b : d.red_a)[(int64_t)(t + 1 - gs + (w2 >> 1)) * ep + e_own] = sum;
          }
          asm volatile("bar.sync 1, %0;" ::"r"(gw * 32) : "memory");   // the buffers are free again
          gs = 0;
          tb_off = 0;
        }
      }
    }
    if (ZW > 0 && !active) return;
    d.ic[slot] = ic; d.sm[slot] = sm; d.sp[slot] = sp; d.wc[slot] = wc;
    if (MODEL != HB_MODEL_HLAND_96) { d.zx1[slot] = zx1; d.zx2[slot] = zx2; }
    if (MODEL == HB_MODEL_HLAND_96 && d.counters && (threadIdx.x & 31) == __ffs(omask) - 1) {
      atomicAdd(d.counters + HB_CNT_ZONE_WARPSTEPS, (unsigned long long)nt);
      atomicAdd(d.counters + HB_CNT_ZONE_WARPSTEPS_DRY, (unsigned long long)n_dry);
      atomicAdd(d.counters + HB_CNT_ZONE_STEPS, (unsigned long long)nt * __popc(omask));
      atomicAdd(d.counters + HB_CNT_ZONE_STEPS_WET, (unsigned long long)n_wet);
    }
    return;
  }

  // ---- general step (any zone type) ------------------------------------------------------------------------------------
  // ref: hland_model.py:584-605 Calc_SPL_WCL_SP_WC_V1 — a loss arises where sp + wc of a snow class exceeds SMax right
  // after the snowfall; as long as none does, that method and Calc_SPG_WCG_SP_WC_V1 (:876-969) leave every state as it is
  const double smax_lim = spec ? KZX(KZ_SMAX) : HUGE_VAL;
  bool exceeded = false;
  for (int t = 0; t < nt; ++t) {
    HB_ZONE_STEP_BEGIN
    // ref: hland_model.py:71-77 Calc_TC_V1, 125-135 Calc_FracRain_V1, 213-225 Calc_PC_V1
    const double tc = in_t + tcorr - tcalt_dz;
    double fracrain;
    if (tc >= tt_hi) fracrain = 1.0;
    else if (tc <= tt_lo) fracrain = 0.0;
    else fracrain = (tc - tt_lo) * inv_ttint;
    double pc = in_p * pcalt_f;
    if (pc <= 0.0) pc = 0.0;
    // (the very expression of the branch-free step above: which of the two steps a FIELD/FOREST zone runs depends on the
    // other 31 zones of its warp, i.e. on where its element sits — the results must not)
    else pc = pc * (pcorr * __drcp_rn(fracrain * inv_rfcf + (1.0 - fracrain) * inv_sfcf));
    // ref: hland_model.py:299-309 Calc_TF_Ic_V1
    double tf;
    if (icept) {
      const double a = pc - (icmax - ic);
      tf = HB_PYMAX(a, 0.0);
      ic += pc - tf;
    } else { tf = pc; ic = 0.0; }
    SERZ(HB_S_INTERCEPTEDWATER, ic);
    double in_ = tf, ncover = 0.0;
    bool bare = true;
    double s_cfact = 0.0, s_gact = 0.0, s_melt = 0.0, s_refr = 0.0, s_glmelt = 0.0;   // recorded only
    if (zt != HB_ILAKE) {
      // ref: hland_model.py:484-499 Calc_SP_WC_V1
      wc += sf * (tf * fracrain);
      sp += sf * (tf * (1.0 - fracrain));
      exceeded |= (sp + wc) - smax_lim > 0.0;
      // ref: hland_model.py:1050-1062 Calc_CFAct_V1
      const double cfa = cfmax + sinf * cfvar;
      const double cfact = HB_PYMAX(cfa, 0.0);
      s_cfact = cfact;
      // ref: hland_model.py:1166-1187 Calc_Melt_SP_WC_V1, 1318-1341 Calc_Refr_SP_WC_V1
      if (tc > ttm) {
        const double potmelt = cfact * (tc - ttm);
        const double melt = HB_PYMIN(potmelt, sp);
        sp -= melt; wc += melt;
        s_melt = melt;
      } else if (tc < ttm) {
        const double potrefr = cfr_cfmax * (ttm - tc);
        const double refr = HB_PYMIN(potrefr, wc);
        sp += refr; wc -= refr;
        s_refr = refr;
      }
      // ref: hland_model.py:1434-1448 Calc_In_WC_V1 (one snow class: the division by SClass is exact)
      const double wc_old = wc, lim = whc * sp;
      wc = HB_PYMIN(wc_old, lim);
      in_ = wc_old - wc;
      bare = sp <= 0.0;
      ncover = (sp > 0.0) ? 1.0 : 0.0;
      int nbare = bare ? 1 : 0;
      if (ncls > 1) {
        // further snow classes (ref: the `for c in range(con.sclass)` loops of the same four methods): their states stay
        // in the state arrays [class][zone][element], read and written in place (coalesced, L1/L2-resident)
        const double dnc = (double)ncls;
        in_ = 0.0 + in_ / dnc;
        const double rain = tf * fracrain, snow = tf * (1.0 - fracrain);
        for (int c = 1; c < ncls; ++c) {
          const int64_t ick = (int64_t)c * tstride + slot;
          const double sfc = d.sfdist[(int64_t)c * ep + e];
          double sp_c = d.sp[ick] + sfc * snow, wc_c = d.wc[ick] + sfc * rain;
          exceeded |= (sp_c + wc_c) - smax_lim > 0.0;
          if (tc > ttm) {
            const double potmelt = cfact * (tc - ttm);
            const double melt = HB_PYMIN(potmelt, sp_c);
            sp_c -= melt; wc_c += melt;
          } else if (tc < ttm) {
            const double potrefr = cfr_cfmax * (ttm - tc);
            const double refr = HB_PYMIN(potrefr, wc_c);
            sp_c += refr; wc_c -= refr;
          }
          const double wo = wc_c, lm = whc * sp_c;
          wc_c = HB_PYMIN(wo, lm);
          in_ += (wo - wc_c) / dnc;
          d.sp[ick] = sp_c; d.wc[ick] = wc_c;
          nbare += (sp_c <= 0.0) ? 1 : 0;
          ncover += (sp_c > 0.0) ? 1.0 : 0.0;
        }
        ncover = ncover / dnc;
      }
      // ref: hland_model.py:1607-1619 Calc_GAct_V1, 1688-1701 Calc_GlMelt_In_V1
      if (zt == HB_GLACIER) {
        const double ga = g_melt + sinf * g_var;
        const double gact = HB_PYMAX(ga, 0.0);
        s_gact = gact;
        if (ncls > 1) {
          if (tc > ttm) {
            const double glmeltpot = gact / (double)ncls * (tc - ttm);
            for (int c = 0; c < nbare; ++c) { s_glmelt += glmeltpot; in_ += glmeltpot; }
          }
        } else if (tc > ttm && bare) {
          s_glmelt = gact * (tc - ttm);
          in_ += s_glmelt;
        }
      }
    } else {
      sp = 0.0; wc = 0.0;
      for (int c = 1; c < ncls; ++c) { d.sp[(int64_t)c * tstride + slot] = 0.0; d.wc[(int64_t)c * tstride + slot] = 0.0; }
    }
    if (SERIES) {
      SERZ(HB_S_TC, tc); SERZ(HB_S_FRACRAIN, fracrain); SERZ(HB_S_PC, pc); SERZ(HB_S_TF, tf); SERZ(HB_S_CFACT, s_cfact);
      SERZ(HB_S_GACT, s_gact); SERZ(HB_S_SPL, 0.0); SERZ(HB_S_WCL, 0.0); SERZ(HB_S_SPG, 0.0); SERZ(HB_S_WCG, 0.0);
      SERZ(HB_S_GLMELT, s_glmelt); SERZ(HB_S_SR, zt == HB_SEALED ? in_ : 0.0); SERZ(HB_S_IN, in_);
      SERS(HB_S_MELT, s_melt); SERS(HB_S_REFR, s_refr); SERS(HB_S_SWE, sp + wc); SERS(HB_S_SP, sp); SERS(HB_S_WC, wc);
    }
    // evap_pet_hbv96.run(): ref: evap_model.py:5111-5125, 5420-5424, 5312-5326
    double pet;
    {
      double ret = in_net * (1.0 + atf * (in_t - in_nat));
      const double lo = HB_PYMAX(ret, 0.0);
      const double hi = 2.0 * in_net;
      ret = HB_PYMIN(lo, hi);
      ret *= etf;
      pet = ret * alt_f;
      if (pet <= 0.0) pet = 0.0;
      else pet *= hb_exp(neg_pf * pc, T);
      SERZ(HB_S_REFERENCEEVAPOTRANSPIRATION, ret); SERZ(HB_S_POTENTIALEVAPOTRANSPIRATION, pet);
    }
    // ref: evap_model.py:6068-6086 Calc_InterceptionEvaporation_V1; hland_model.py:383-396 Calc_EI_Ic_AETModel_V1
    double aet_ie = 0.0;
    if (zf & HB_ZF_INTERCEPTION) aet_ie = HB_PYMIN(pet, ic);
    if (SERIES) {
      SERZ(HB_S_INTERCEPTIONEVAPORATION, aet_ie); SERZ(HB_S_SNOWCOVER, ncover);
      SERZ(HB_S_WATEREVAPORATION, ((zf & HB_ZF_WATER) && tc > KZX(KZ_TTI)) ? pet : 0.0);
    }
    double s_ei = 0.0;
    if (icept) { s_ei = HB_PYMIN(aet_ie, ic); ic -= s_ei; }
    else ic = 0.0;
    SERZ(HB_S_EI, s_ei); SERZ(HB_S_IC, ic);
    double ta = 0.0, tb = 0.0;
    double r_z = in_;                 // r of the zone (non-soil zones: r = in_)
    if (soil) {
      // ref: hland_model.py:1776-1790 Calc_R_SM_V1
      double r;
      if (fc > 0.0) {
        r = in_ * hb_pow(sm * inv_fc, beta, T);
        const double a = sm + in_ - fc;
        r = HB_PYMAX(r, a);
      } else r = in_;
      sm += in_ - r;
      // ref: hland_model.py:1904-1919 Calc_CF_SM_V1 with CFlux == 0: cf = min(0*(1 - sm/fc), fc - sm); hland_96 only
      double cf = 0.0;
      if (MODEL == HB_MODEL_HLAND_96 && fc > 0.0) {
        cf = 0.0 * (1.0 - sm * inv_fc);
        const double b = fc - sm;
        cf = HB_PYMIN(cf, b);
      }
      sm += cf;
      SERZ(HB_S_SOILWATER, sm);
      // ref: evap_model.py:7630-7658 Determine_SoilEvapotranspiration_V1 (6657-6672, 7001-7018, 7060-7069)
      double set = 0.0;
      if (zf & HB_ZF_SOIL) {
        set = pet;
        if (set > 0.0) {
          if (sm < 0.0) set = 0.0;
          else if (sm < thresh) set *= sm * inv_thresh;
        }
        const double petc = excessred * pet + (1.0 - excessred) * (pet + pet) / 2.0;
        const double excess = set + aet_ie - petc;
        if ((petc >= 0.0) && (excess > 0.0)) set -= excessred * excess;
        else if ((petc < 0.0) && (excess < 0.0)) set -= excessred * excess;
        set *= 1.0 - ncover;
      }
      SERZ(HB_S_SOILEVAPOTRANSPIRATION, set);
      // ref: hland_model.py:2002-2018 Calc_EA_SM_AETModel_V1
      const double ea = HB_PYMIN(set, sm);
      sm -= ea;
      if (sm > fc) { r += sm - fc; sm = fc; }
      // ref: hland_model.py:2092-2101 Calc_InUZ_V1 (term), 2224-2236 Calc_ContriArea_V1 (factor)
      r_z = r;
      if (MODEL == HB_MODEL_HLAND_96) {
        ta = w_a * (r - cf);
        if (resp && fc > 0.0) tb = w_b * hb_log(sm * inv_fc, T);   // weighted log of the contributing-area factor
      }
      SERZ(HB_S_R, r); SERZ(HB_S_CF, cf); SERZ(HB_S_EA, ea); SERZ(HB_S_SM, sm); SERZ(HB_S_EL, 0.0);
    } else {
      sm = 0.0;
      double el = 0.0;
      if (zt == HB_ILAKE) {
        // ref: hland_model.py:3113-3124 Calc_LZ_V1 (term), 3724-3737 Calc_EL_LZ_AETModel_V1 (term);
        // evap_model.py:5673-5681 Calc_WaterEvaporation_V1
        el = ((zf & HB_ZF_WATER) && tc > tti) ? pet : 0.0;
        if (MODEL == HB_MODEL_HLAND_96) { ta = w_a * pc; tb = w_b * el; }
        else { ta = el; tb = w_b * pc; }   // hland_96p: Calc_EL_SG2_SG3 / Calc_GR2_GR3 terms; hland_96c: Calc_EL_LZ / Calc_LZ_V2
      } else ta = w_a * (in_ - 0.0);  // GLACIER: InUZ term; SEALED: InRC term (r = in_, ref: hland_model.py:1528-1535)
      SERZ(HB_S_R, in_); SERZ(HB_S_CF, 0.0); SERZ(HB_S_EA, 0.0); SERZ(HB_S_SM, 0.0); SERZ(HB_S_EL, el);
      // (a zone without soil: the submodel sees an empty soil; its Soil flag may still be set)
      if (SERIES) {
        SERZ(HB_S_SOILWATER, 0.0);
        double set = 0.0;
        if (zf & HB_ZF_SOIL) {   // ref: evap_model.py:6657-6672, 7001-7018, 7060-7069 with soilwater = 0
          set = pet;
          if (set > 0.0) set = (0.0 < thresh) ? set * (0.0 * inv_thresh) : set;
          const double petc = excessred * pet + (1.0 - excessred) * (pet + pet) / 2.0;
          const double excess = set + aet_ie - petc;
          if ((petc >= 0.0) && (excess > 0.0)) set -= excessred * excess;
          else if ((petc < 0.0) && (excess < 0.0)) set -= excessred * excess;
          set *= 1.0 - ncover;
        }
        SERZ(HB_S_SOILEVAPOTRANSPIRATION, set);
      }
    }
    if (MODEL != HB_MODEL_HLAND_96) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;   // recorded only: dp rs ri gr1 rg1 | qab1 qvs1 qab2 qvs2
      if (upper) {
        if (MODEL == HB_MODEL_HLAND_96P) {
          double dp, rs, ri, gr1, rg1;
          response_96p(r_z, kpercmax, kx0, kx1, kx2, kx3, kx4, kx5, kx6, zx1, zx2, dp, rs, ri, gr1, rg1);
          ta = w_a * (rs + ri + rg1);
          tb = w_b * (dp - gr1);
          s0 = dp; s1 = rs; s2 = ri; s3 = gr1; s4 = rg1;
        } else {
          double qab1 = 0.0, qvs1 = 0.0, qab2 = 0.0, qvs2 = 0.0;
          qab_qvs_bw_t<true>(kx0, kx1, kx2, zx1, r_z, qab1, qvs1, T);
          qab_qvs_bw_t<true>(kx3, kx4, kx5, zx2, qvs1, qab2, qvs2, T);
          ta = w_a * (qab1 + qab2);
          tb = w_b * qvs2;
          s0 = qab1; s1 = qvs1; s2 = qab2; s3 = qvs2;
        }
      } else { zx1 = 0.0; zx2 = 0.0; }   // ILAKE (terms set above), SEALED (ta = w_a * r)
      if (SERIES) {
        if (MODEL == HB_MODEL_HLAND_96P) {
          SERZ(HB_S_DP, s0); SERZ(HB_S_RS, s1); SERZ(HB_S_RI, s2); SERZ(HB_S_GR1, s3); SERZ(HB_S_RG1, s4);
          SERZ(HB_S_SUZ, zx1); SERZ(HB_S_SG1, zx2);
        } else {
          SERZ(HB_S_QAB1, s0); SERZ(HB_S_QVS1, s1); SERZ(HB_S_QAB2, s2); SERZ(HB_S_QVS2, s3);
          SERZ(HB_S_BW1, zx1); SERZ(HB_S_BW2, zx2);
        }
      }
    }
    HB_ZONE_STEP_END(ta, tb)
  }
#undef HB_ZONE_STEP_BEGIN
#undef HB_ZONE_STEP_END
#undef SERZ
#undef SERS
#undef tcorr
#undef tcalt_dz
#undef tt_hi
#undef tt_lo
#undef inv_ttint
#undef pcalt_f
#undef pcorr
#undef inv_rfcf
#undef inv_sfcf
#undef icmax
#undef cfmax
#undef cfvar
#undef ttm
#undef cfr_cfmax
#undef whc
#undef fc
#undef inv_fc
#undef beta
#undef thresh
#undef inv_thresh
#undef excessred
#undef atf
#undef etf
#undef alt_f
#undef neg_pf
#undef w_a
#undef w_b
#undef sf
#undef kx0
#undef kx1
#undef kx2
#undef kx3
#undef kx4
#undef kx5
#undef kx6
#undef kpercmax
#undef KZX
  d.ic[slot] = ic; d.sm[slot] = sm; d.sp[slot] = sp; d.wc[slot] = wc;
  if (MODEL != HB_MODEL_HLAND_96) { d.zx1[slot] = zx1; d.zx2[slot] = zx2; }
  if (exceeded) d.spec_flag[e] = 1;   // (every zone of the element that saw it stores the same value; never beside a 2)
}

// Speculative fast path (HB_EI_SPEC): the elements of `list` whose flag is set get the conditions of the window's start
// back, so that the general kernel can repeat the window for them (one thread per element; rare).
struct SpecSnap {
  double *ic, *sm, *sp, *wc, *uz, *lz, *quh;                            // live states
  const double *s_ic, *s_sm, *s_sp, *s_wc, *s_uz, *s_lz, *s_quh;        // … at the start of the window
};
__global__ void __launch_bounds__(128) spec_restore_kernel(const int32_t* __restrict__ list, int64_t n,
                                                           const int32_t* __restrict__ flag, int64_t ep, int zmax,
                                                           const int32_t* __restrict__ nz, const int32_t* __restrict__ nc,
                                                           const int32_t* __restrict__ nu, const SpecSnap s,
                                                           unsigned long long* counters) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t e = list[i];
  if (!flag[e]) return;
  const int64_t zs = (int64_t)zmax * ep;
  for (int k = 0; k < nz[e]; ++k) {
    const int64_t slot = (int64_t)k * ep + e;
    s.ic[slot] = s.s_ic[slot];
    s.sm[slot] = s.s_sm[slot];
    for (int c = 0; c < nc[e]; ++c) {
      s.sp[c * zs + slot] = s.s_sp[c * zs + slot];
      s.wc[c * zs + slot] = s.s_wc[c * zs + slot];
    }
  }
  s.uz[e] = s.s_uz[e];
  s.lz[e] = s.s_lz[e];
  for (int j = 0; j < nu[e]; ++j) s.quh[(int64_t)j * ep + e] = s.s_quh[(int64_t)j * ep + e];
  if (counters) atomicAdd(counters + HB_CNT_SPEC_RERUN, 1ull);
}

// staging layout [nt][zmax][E_pad]  →  ABI layout [nt_window][width] for the zones (or snow cells) of the fast-path
// elements.  A block takes 32 consecutive elements: their staging rows are read coalesced into a shared-memory tile,
// their ABI cells form ONE contiguous range (zones of consecutive elements follow each other), written coalesced.
// slot_of[cell] = k*E_pad + e of the ABI cell, or -1 when the cell belongs to an element of the general kernel.
__global__ void __launch_bounds__(256) series_scatter_kernel(const double* __restrict__ staged, double* __restrict__ abi,
                                                             const int64_t* __restrict__ cell0,   // [E_pad+1] first cell
                                                             const int32_t* __restrict__ slot_of, int64_t n_elem,
                                                             int64_t e_pad, int zmax, int64_t width, int nt, int64_t t_out) {
  extern __shared__ double tile[];                      // [zmax][33]
  const int64_t e0 = (int64_t)blockIdx.x * 32;
  const int64_t e1 = e0 + 32 < n_elem ? e0 + 32 : n_elem;
  const int64_t z_begin = cell0[e0], z_end = cell0[e1];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int t = blockIdx.y; t < nt; t += gridDim.y) {
    const double* src = staged + (int64_t)t * zmax * e_pad + e0;
    for (int k = wib; k < zmax; k += 8) tile[k * 33 + lane] = src[(int64_t)k * e_pad + lane];
    __syncthreads();
    double* dst = abi + (t_out + t) * width;
    for (int64_t z = z_begin + threadIdx.x; z < z_end; z += 256) {
      const int32_t sl = slot_of[z];
      if (sl >= 0) dst[z] = tile[(sl / e_pad) * 33 + (sl % e_pad - e0)];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// element kernel: ordered zone sums + Calc_Q0_Perc_UZ_V1 … Calc_QT_V1
// ---------------------------------------------------------------------------------------------------------------------
// rconc_uh (ref: hydpy/models/rconc/rconc_model.py:89-95 Determine_Outflow_V1) with the element's `quh` resident in shared
// memory ([ordinate][thread], 64 threads per block) for the whole launch: the per-step loads of the global copy wait for
// the L2 (stores bypass the L1), a quarter of the element kernel's stall cycles once the zone sums arrive ready-made
__device__ __forceinline__ double uh_outflow_staged(const double* __restrict__ uh, int64_t ep, int64_t e, int nu,
                                                    double inrc, double* s_q, const double* s_w) {
  if (s_w != nullptr) {   // weights staged as well
    const double outrc = s_w[0] * inrc + s_q[0];
    for (int j = 1; j < nu; ++j) s_q[(j - 1) * 64] = s_w[j * 64] * inrc + s_q[j * 64];
    return outrc;
  }
  const double outrc = ldg(uh + e) * inrc + s_q[0];
  for (int j = 1; j < nu; ++j) s_q[(j - 1) * 64] = ldg(uh + (int64_t)j * ep + e) * inrc + s_q[j * 64];
  return outrc;
}

#ifndef HB_ELEM_PREFETCH
#define HB_ELEM_PREFETCH 1
#endif
#ifndef HB_ELEM_BLOCKS
#define HB_ELEM_BLOCKS 12
#endif
// FUSED: the elements whose zone terms were added inside the zone kernel (HB_EI_FUSED in a run with d.fused): two sums per
// step arrive instead of two terms per zone
template <bool SERIES, int MODEL, bool FUSED = false>
__global__ void __launch_bounds__(64, HB_ELEM_BLOCKS) element_kernel(const LandV2Dev d) {
  static_assert(!FUSED || (!SERIES && MODEL == HB_MODEL_HLAND_96), "fused reduction: hland_96 without series only");
  __shared__ MathTables s_tables;
  extern __shared__ double s_sc_dyn[];   // rconc_nash cascade of the block's elements, [storage][thread]; only allocated
  double* const s_sc = d.nash_stage ? s_sc_dyn : nullptr;   // (launch parameter) when some element has a cascade
  if (MODEL == HB_MODEL_HLAND_96) hb_chain_stage(d.chain_tables);   // (before the barrier of stage_tables)
  stage_tables(&s_tables, d.tables);
  const TabRef T = hb_tabref(&s_tables);
  const ChainRef CT = hb_chainref(nullptr);
  const int64_t e = d.e_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.n_elem || e >= d.e_lo + (int64_t)d.slab_blocks * 128) return;
  const int einfo = d.einfo[e];
  if (!(einfo & HB_EI_FAST) || HB_EI_MODEL(einfo) != MODEL) return;
  if (((einfo & HB_EI_FUSED) != 0 && d.fused != 0) != FUSED) return;
  if (MODEL == HB_MODEL_HLAND_96 && d.skip_classes && (d.nc[e] > 1 || (einfo & HB_EI_SPEC))) return;
  if (MODEL == HB_MODEL_HLAND_96 && !FUSED && (einfo & HB_EI_SPEC) && d.spec_flag[e] == 2) return;
  const int64_t ep = d.e_pad;
  const int nz = d.nz[e], nu = d.nu[e], recstep = d.recstep[e];
  const int nash_steps = d.nash_steps ? d.nash_steps[e] : 0;
// (development builds for A/B timing, `python -m hydpy_b200.build --variant x -DHB_ELEM_NO_UHSMEM` etc.: the unit hydrograph in
// global memory, the RecStep chain in its general form only, no early exits for dry sub-steps)
#ifdef HB_ELEM_NO_UHSMEM
  const bool uh_smem = false;
#else
  const bool uh_smem = s_sc != nullptr && nash_steps == 0 && nu > 0 && nu <= d.uh_stage;
#endif
  const double* const s_uw = (uh_smem && d.uh_weights) ? s_sc + (size_t)d.uh_stage * 64 + threadIdx.x : nullptr;
  if (uh_smem)
    for (int j = 0; j < nu; ++j) {
      s_sc[threadIdx.x + j * 64] = d.quh[(int64_t)j * ep + e];
      if (d.uh_weights) s_sc[(size_t)d.uh_stage * 64 + threadIdx.x + j * 64] = d.uh[(int64_t)j * ep + e];
    }
  const double dt = d.ke[KE_DT * ep + e], dt_percmax = d.ke[KE_DT_PERCMAX * ep + e], dt_k = d.ke[KE_DT_K * ep + e],
               alpha1 = d.ke[KE_ALPHA1 * ep + e], k4 = d.ke[KE_K4 * ep + e], gamma1 = d.ke[KE_GAMMA1 * ep + e],
               relupper = d.ke[KE_RELUPPER * ep + e], rellower = d.ke[KE_RELLOWER * ep + e],
               up_low = d.ke[KE_UP_LOW * ep + e], qfactor = d.ke[KE_QFACTOR * ep + e],
               beta_last = d.ke[KE_BETA_LAST * ep + e];
  const bool resp = (einfo & HB_EI_RESPAREA) != 0, soilonly = (einfo & HB_EI_SOILONLY) != 0;
  const bool square = (alpha1 == 2.0);
  // RecStep chain in its short form (recstep_chain below): log2(dt*k) once per thread; the form needs a positive normal
  // dt*k and a finite exponent, everything else keeps the general loop
  // (the bounds make an out-of-range argument of the power function visible in its exponent, see hb_chain_pow2)
  const bool chain_ok = MODEL == HB_MODEL_HLAND_96 && dt_k >= 1e-30 && dt_k <= 1e20 && alpha1 >= 1.0 && alpha1 <= 64.0;
  const double log2_dtk = chain_ok ? log2(dt_k) : 0.0;
  double uz = d.uz[e], lz = d.lz[e];
  double sg2 = 0.0, sg3 = 0.0;
  if (MODEL == HB_MODEL_HLAND_96P) { sg2 = d.ex1[e]; sg3 = d.ex2[e]; }
  const int64_t tstride = (int64_t)d.zmax * ep;
  unsigned n_run = 0, n_pow = 0;   // element-steps that ran the RecStep chain, sub-steps that evaluated the power function
  unsigned n_wrun = 0;             // (warp-uniform) steps in which some lane of the warp ran it
  const unsigned emask = __activemask();
  double nx_a = 0.0, nx_b = 0.0;    // FUSED: the sums of the next step, loaded one step ahead
  if (FUSED && d.nt > 0) { nx_a = ldg(d.red_a + e); nx_b = ldg(d.red_b + e); }
  for (int64_t t = 0; t < d.nt; ++t) {
    const double* ta = FUSED ? nullptr : d.term_a + t * tstride + e;
    const double* tb = FUSED ? nullptr : d.term_b + t * tstride + e;
#if HB_ELEM_PREFETCH
    // the terms of the NEXT step are pulled towards the SM while the RecStep loop of this step runs (they were written
    // by the zone kernel gigabytes ago, i.e. they come from HBM); costs no registers
    if (!FUSED && t + 1 < d.nt)
      for (int k = 0; k < nz; ++k) {
#if HB_ELEM_PREFETCH == 1
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ta + tstride + (int64_t)k * ep));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(tb + tstride + (int64_t)k * ep));
#else
        asm volatile("prefetch.global.L1 [%0];" ::"l"(ta + tstride + (int64_t)k * ep));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(tb + tstride + (int64_t)k * ep));
#endif
      }
#endif
    if (MODEL == HB_MODEL_HLAND_96P) {
      // ref: hland_model.py:3356-3372 Calc_GR2_GR3_V1 (tb = weight * (dp - gr1) | weight * pc, in zone order)
      const double fsg = d.ke[KE_FSG * ep + e], omfsg = d.ke[KE_OMFSG * ep + e];
      double gr2 = 0.0, gr3 = 0.0;
      for (int k = 0; k < nz; ++k) {
        if (!soilonly && HB_ZI_TYPE(d.zinfo[(int64_t)k * ep + e]) == HB_SEALED) continue;
        const double total = tb[(int64_t)k * ep];
        gr2 += fsg * total;
        gr3 += omfsg * total;
      }
      // ref: hland_model.py:3556-3576 Calc_RG2_SG2_V1, 3646-3665 Calc_RG3_SG3_V1
      double rg2, rg3;
      {
        double s_ = sg2, g_ = gr2, k3 = d.ke[KE_K3 * ep + e], w3 = d.ke[KE_W3 * ep + e];
        if ((s_ < 0.0) && (0.0 < g_)) {
          const double a = -sg2;
          const double add = HB_PYMIN(a, g_);
          k3 *= g_ / add;
          w3 = hb_exp(-1.0 / k3, T);
          s_ += add;
          g_ -= add;
        }
        if (s_ >= 0.0) { sg2 = w3 * s_ + (1.0 - w3) * k3 * g_; rg2 = s_ + g_ - sg2; }
        else { sg2 = s_; rg2 = 0.0; }
      }
      {
        double s_ = sg3, g_ = gr3, k4_ = k4, w4 = d.ke[KE_W4 * ep + e];
        if ((s_ < 0.0) && (0.0 < g_)) {
          const double a = -sg3;
          const double add = HB_PYMIN(a, g_);
          k4_ *= g_ / add;
          w4 = hb_exp(-1.0 / k4_, T);
          s_ += add;
          g_ -= add;
        }
        if (s_ >= 0.0) { sg3 = w4 * s_ + (1.0 - w4) * k4_ * g_; rg3 = s_ + g_ - sg3; }
        else { sg3 = s_; rg3 = 0.0; }
      }
      // ref: hland_model.py:3444-3460 Calc_EL_SG2_SG3_AETModel_V1 (ta of an ILAKE zone = el)
      if (einfo & HB_EI_LAKE)
        for (int k = 0; k < nz; ++k)
          if (HB_ZI_TYPE(d.zinfo[(int64_t)k * ep + e]) == HB_ILAKE) {
            const double weight = d.kz[((int64_t)k * KZ_COUNT + KZ_W_LZ) * ep + e], el = ta[(int64_t)k * ep];
            sg2 -= fsg * weight * el;
            sg3 -= omfsg * weight * el;
          }
      // ref: hland_model.py:3975-3984 Calc_InRC_V2 (ta = relzonearea * (rs + ri + rg1) | relzonearea * r)
      double inrc = rellower * (rg2 + rg3);
      for (int k = 0; k < nz; ++k) {
        if ((einfo & HB_EI_LAKE) && HB_ZI_TYPE(d.zinfo[(int64_t)k * ep + e]) == HB_ILAKE) continue;
        inrc += ta[(int64_t)k * ep];
      }
      const double outrc = uh_smem ? uh_outflow_staged(d.uh, ep, e, nu, inrc, s_sc + threadIdx.x, s_uw)
                                 : rconc_outflow<true>(d.uh, d.quh, d.ke, ep, e, nu, nash_steps, inrc, s_sc + threadIdx.x, 64);
      d.qt[(d.t_out + t) * ep + e] = qfactor * outrc;
      if (SERIES) {
#define SERE(id, v) do { if (d.series[id]) d.series[id][(d.t_out + t) * d.n_elem + e] = (v); } while (0)
        record_sc(d.series[HB_S_SC], d.t_out + t, d.n_uh, d.uh0[e], d.quh, ep, e, nu, nash_steps);
        SERE(HB_S_GR2, gr2); SERE(HB_S_GR3, gr3); SERE(HB_S_RG2, rg2); SERE(HB_S_RG3, rg3); SERE(HB_S_SG2, sg2);
        SERE(HB_S_SG3, sg3); SERE(HB_S_INRC, inrc); SERE(HB_S_OUTRC, outrc); SERE(HB_S_RT, outrc);
        SERE(HB_S_QT, qfactor * outrc);
#undef SERE
      }
      continue;
    }
    if (MODEL == HB_MODEL_HLAND_96C) {
      // ref: hland_model.py:4039-4051 Calc_InRC_V3 (ta = weight * (qab1 + qab2) | weight * r, in zone order)
      double inrc = 0.0;
      for (int k = 0; k < nz; ++k) {
        if ((einfo & HB_EI_LAKE) && HB_ZI_TYPE(d.zinfo[(int64_t)k * ep + e]) == HB_ILAKE) continue;
        inrc += ta[(int64_t)k * ep];
      }
      const double outrc = uh_smem ? uh_outflow_staged(d.uh, ep, e, nu, inrc, s_sc + threadIdx.x, s_uw)
                                 : rconc_outflow<true>(d.uh, d.quh, d.ke, ep, e, nu, nash_steps, inrc, s_sc + threadIdx.x, 64);
      // ref: hland_model.py:3172-3181 Calc_LZ_V2 (tb = weight * pc | weight * qvs2)
      for (int k = 0; k < nz; ++k) {
        if (!soilonly && HB_ZI_TYPE(d.zinfo[(int64_t)k * ep + e]) == HB_SEALED) continue;
        lz += tb[(int64_t)k * ep];
      }
      // ref: hland_model.py:3724-3737 Calc_EL_LZ_AETModel_V1 (ta of an ILAKE zone = el)
      if (einfo & HB_EI_LAKE)
        for (int k = 0; k < nz; ++k)
          if (HB_ZI_TYPE(d.zinfo[(int64_t)k * ep + e]) == HB_ILAKE)
            lz -= d.kz[((int64_t)k * KZ_COUNT + KZ_W_LZ) * ep + e] * ta[(int64_t)k * ep];
      // ref: hland_model.py:3832-3840 Calc_Q1_LZ_V1, 4168-4171 Calc_RT_V2, 4197-4200 Calc_QT_V1
      double q1 = 0.0;
      if (lz > 0.0) q1 = k4 * hb_pow(lz, gamma1, T);
      lz -= q1;
      const double rt = d.ke[KE_RELLANDAREA * ep + e] * outrc + rellower * q1;
      d.qt[(d.t_out + t) * ep + e] = qfactor * rt;
      if (SERIES) {
#define SERE(id, v) do { if (d.series[id]) d.series[id][(d.t_out + t) * d.n_elem + e] = (v); } while (0)
        record_sc(d.series[HB_S_SC], d.t_out + t, d.n_uh, d.uh0[e], d.quh, ep, e, nu, nash_steps);
        SERE(HB_S_Q1, q1); SERE(HB_S_INRC, inrc); SERE(HB_S_OUTRC, outrc); SERE(HB_S_RT, rt); SERE(HB_S_QT, qfactor * rt);
        SERE(HB_S_LZ, lz);
#undef SERE
      }
      continue;
    }
    double inuz = 0.0, contriarea = 1.0, logsum = 0.0;   // logsum: sum of weight * log(sm/fc) over the soil zones
    if (FUSED) {
      inuz = nx_a; logsum = nx_b;
      if (t + 1 < d.nt) { nx_a = ldg(d.red_a + (t + 1) * ep + e); nx_b = ldg(d.red_b + (t + 1) * ep + e); }
    } else if (soilonly) {
      for (int k = 0; k < nz; ++k) {
        inuz += ta[(int64_t)k * ep];
        logsum += tb[(int64_t)k * ep];
      }
    } else {
      for (int k = 0; k < nz; ++k) {
        const int zt = HB_ZI_TYPE(d.zinfo[(int64_t)k * ep + e]);
        if (zt == HB_FIELD || zt == HB_FOREST) { inuz += ta[(int64_t)k * ep]; logsum += tb[(int64_t)k * ep]; }
        else if (zt == HB_GLACIER) inuz += ta[(int64_t)k * ep];
      }
    }
    // ref: hland_model.py:2224-2236 Calc_ContriArea_V1: (prod (sm/fc)^weight)^beta = exp(beta * logsum)
    if (resp) contriarea = hb_exp(beta_last * logsum, T);
    // ref: hland_model.py:2456-2484 Calc_Q0_Perc_UZ_V1
    double perc_sum = 0.0, q0_sum = 0.0;
    // an empty upper zone without recharge stays empty through all RecStep sub-steps (uz = max(0 + 0, 0), perc = q0 = 0,
    // no error correction): dry spells skip the chain — the same bits, nothing to wait for
    if (MODEL == HB_MODEL_HLAND_96) n_wrun += __any_sync(emask, !(uz == 0.0 && inuz == 0.0)) ? 1u : 0u;
    if (!(uz == 0.0 && inuz == 0.0)) {
      ++n_run;
      const double uz_old = uz;
      const double dt_inuz = dt * inuz;
      const double perc_pot = dt_percmax * contriarea;
      // The loop below is the reference's; every instruction of it sits on one dependent chain.  Three exact or
      // (power function only) 1e-14-equivalent forms take its place where an element's OWN data allow — the choice never
      // depends on the warp neighbours, so an element's bits do not depend on where it is placed:
      //  * with uz >= 0 and recharge >= 0, `max(uz + dt*inuz, 0)` is the sum itself, and percolation + outflow become
      //    d = a - perc_pot; perc = d > 0 ? perc_pot : a;  u = d - q; q0 = u < 0 ? d : q;  uz = u > 0 ? u : 0
      //    (the same values: x < y <=> x - y < 0 in IEEE arithmetic with gradual underflow);
      //  * the power function runs on d whether or not d > 0 (no branch on the chain; lanes without outflow drop the
      //    garbage), as 2^(y*log2(d) + c2) with the step-invariant factors folded into c2 (hb_chain_pow2);
      //  * a quadratic response (alpha = 1, the Lahn default) needs no power function at all: exact.
      const bool lane_plain = chain_ok && uz >= 0.0 && uz <= 1e300 && dt_inuz >= 0.0 && dt_inuz <= 1e300 &&
                              perc_pot >= 0.0 && perc_pot <= 1e300 && contriarea >= 1e-30 && contriarea <= 2.0;
#ifdef HB_ELEM_NO_FAST
      bool redo = true;
      if (false) {
#else
      bool redo = !lane_plain;
      if (lane_plain) {
#endif
        // (whole warp) no sub-step leaves water for the outflow: uz + dt*inuz <= perc_pot now, and dt*inuz <= perc_pot
        // from then on because every sub-step empties the storage — percolation takes everything, 50 additions remain
        if (__all_sync(__activemask(), uz + dt_inuz <= perc_pot)) {
          perc_sum = uz + dt_inuz;
          for (int i = 1; i < recstep; ++i) perc_sum += dt_inuz;
          uz = 0.0;
        } else if (square) {
          const double inv_ca = 1.0 / contriarea;
          double a = uz + dt_inuz, u = 0.0;
          bool pos = false;
          const bool refill = dt_inuz > perc_pot;    // recharge alone can lift the storage above the percolation again
          // sub-steps in runs of 2, 6 and the rest: after the first two runs, if every storage of the warp is empty for good, percolation
          // takes the recharge of each remaining sub-step — additions only, what the reference's loop does in dry hours
          // without its branches.  (The test sits between plain counted loops: inside the loop it costs the unrolling.)
          for (int i = 0, hi = 2; i < recstep; hi = hi == 2 ? 8 : recstep) {
            const int end = hi < recstep ? hi : recstep;
            for (; i < end; ++i) {
              const double dd = a - perc_pot;
              const bool p = dd > 0.0;
              perc_sum += p ? perc_pot : a;
              const double x = dd * inv_ca;
              const double q = dt_k * (x * x);
              u = dd - q;
              pos = p && (u > 0.0);
              if (p) q0_sum += pos ? q : dd;      // min(q, dd): u == 0 means q == dd
              a = pos ? u + dt_inuz : dt_inuz;
            }
#ifndef HB_ELEM_NO_EARLYOUT
            if (i < recstep && __all_sync(__activemask(), !pos && !refill)) {
              for (; i < recstep; ++i) perc_sum += dt_inuz;
              break;
            }
#endif
          }
          uz = pos ? u : 0.0;
        } else {
          // c2 = log2(dt*k) - y*log2(contriarea); ln(contriarea) = beta*logsum is what contriarea was made from
          const double c2 = resp ? fma(-alpha1 * HB_CLIT(HB_C_L1), beta_last * logsum, log2_dtk) : log2_dtk;
          double a = uz + dt_inuz, u = 0.0;
          bool pos = false, bad = false;
          int n_eval = recstep;                     // power functions evaluated: every sub-step this loop runs
          const bool refill = dt_inuz > perc_pot;
          for (int i = 0, hi = 2; i < recstep; hi = hi == 2 ? 8 : recstep) {   // (runs of 2, 6 and the rest: see the quadratic loop)
            const int end = hi < recstep ? hi : recstep;
            for (; i < end; ++i) {
              const double dd = a - perc_pot;
              const bool p = dd > 0.0;
              perc_sum += p ? perc_pot : a;
              double tq;
              const double q = hb_chain_pow2(dd, alpha1, c2, CT, tq);
              bad |= p && !(fabs(tq) < 800.0);
              u = dd - q;
              pos = p && (u > 0.0);
              if (p) q0_sum += pos ? q : dd;      // min(q, dd): u == 0 means q == dd
              a = pos ? u + dt_inuz : dt_inuz;
            }
#ifndef HB_ELEM_NO_EARLYOUT
            if (i < recstep && __all_sync(__activemask(), !pos && !refill)) {
              n_eval = i;
              for (; i < recstep; ++i) perc_sum += dt_inuz;
              break;
            }
#endif
          }
          uz = pos ? u : 0.0;
          if (bad) { redo = true; uz = uz_old; perc_sum = 0.0; q0_sum = 0.0; }   // out-of-range power: general loop
          else n_pow += (unsigned)n_eval;
        }
      }
      if (redo) {
        const double inv_ca = 1.0 / contriarea;
        for (int i = 0; i < recstep; ++i) {
          const double a = uz + dt_inuz;
          uz = HB_PYMAX(a, 0.0);
          const double perc = HB_PYMIN(perc_pot, uz);
          uz -= perc;
          perc_sum += perc;
          if (uz > 0.0) {
            double q0;
            if (contriarea > 0.0) {
              const double x = uz * inv_ca;
              const double q = dt_k * (square ? x * x : hb_pow(x, alpha1, T));
              q0 = HB_PYMIN(q, uz);
            } else q0 = uz;
            uz -= q0;
            q0_sum += q0;
            ++n_pow;
          }
        }
      }
      const double error = uz - (uz_old + inuz - perc_sum - q0_sum);
      if (error > 0.0) {
        const double factor = 1.0 - error / (perc_sum + q0_sum);
        perc_sum *= factor;
        q0_sum *= factor;
      }
    } else uz = 0.0;   // (-0.0 + 0.0 = +0.0 in the loop)
    // ref: hland_model.py:3113-3124 Calc_LZ_V1, 3724-3737 Calc_EL_LZ_AETModel_V1
    if (rellower > 0.0) {
      lz += up_low * perc_sum;
      if (einfo & HB_EI_LAKE)
        for (int k = 0; k < nz; ++k)
          if (HB_ZI_TYPE(d.zinfo[(int64_t)k * ep + e]) == HB_ILAKE) lz += ta[(int64_t)k * ep];
    } else lz = 0.0;
    if (einfo & HB_EI_LAKE)
      for (int k = 0; k < nz; ++k)
        if (HB_ZI_TYPE(d.zinfo[(int64_t)k * ep + e]) == HB_ILAKE) lz -= tb[(int64_t)k * ep];
    // ref: hland_model.py:3832-3840 Calc_Q1_LZ_V1
    double q1 = 0.0;
    if (lz > 0.0) q1 = k4 * hb_pow(lz, gamma1, T);
    lz -= q1;
    // ref: hland_model.py:3905-3912 Calc_InRC_V1
    double inrc = relupper * q0_sum + rellower * q1;
    if (einfo & HB_EI_SEALED)
      for (int k = 0; k < nz; ++k)
        if (HB_ZI_TYPE(d.zinfo[(int64_t)k * ep + e]) == HB_SEALED) inrc += ta[(int64_t)k * ep];
    // ref: hland_model.py:4109-4116 Calc_OutRC_V1; rconc_model.py:89-95 Determine_Outflow_V1
    const double outrc = uh_smem ? uh_outflow_staged(d.uh, ep, e, nu, inrc, s_sc + threadIdx.x, s_uw)
                                 : rconc_outflow<true>(d.uh, d.quh, d.ke, ep, e, nu, nash_steps, inrc, s_sc + threadIdx.x, 64);
    if (SERIES) record_sc(d.series[HB_S_SC], d.t_out + t, d.n_uh, d.uh0[e], d.quh, ep, e, nu, nash_steps);
    // ref: hland_model.py:4139-4141 Calc_RT_V1, 4197-4200 Calc_QT_V1
    const double qt = qfactor * outrc;
    d.qt[(d.t_out + t) * ep + e] = qt;
    if (SERIES) {
#define SERE(id, v) do { if (d.series[id]) d.series[id][(d.t_out + t) * d.n_elem + e] = (v); } while (0)
      SERE(HB_S_INUZ, inuz); SERE(HB_S_CONTRIAREA, contriarea); SERE(HB_S_PERC, perc_sum); SERE(HB_S_Q0, q0_sum);
      SERE(HB_S_Q1, q1); SERE(HB_S_INRC, inrc); SERE(HB_S_OUTRC, outrc); SERE(HB_S_RT, outrc); SERE(HB_S_QT, qt);
      SERE(HB_S_UZ, uz); SERE(HB_S_LZ, lz);
#undef SERE
    }
  }
  d.uz[e] = uz;
  d.lz[e] = lz;
  if (uh_smem)
    for (int j = 0; j < nu; ++j) d.quh[(int64_t)j * ep + e] = s_sc[threadIdx.x + j * 64];
  if (MODEL == HB_MODEL_HLAND_96P) { d.ex1[e] = sg2; d.ex2[e] = sg3; }
  if (MODEL == HB_MODEL_HLAND_96 && d.counters) {
    const unsigned m = __activemask();
    const unsigned run = __reduce_add_sync(m, n_run), pw = __reduce_add_sync(m, n_pow);
    if ((threadIdx.x & 31) == __ffs(m) - 1) {
      atomicAdd(d.counters + HB_CNT_ELEM_STEPS, (unsigned long long)d.nt * __popc(m));
      atomicAdd(d.counters + HB_CNT_ELEM_STEPS_RUN, (unsigned long long)run);
      atomicAdd(d.counters + HB_CNT_SUBSTEPS_POW, (unsigned long long)pw);
      atomicAdd(d.counters + HB_CNT_ELEM_WARPSTEPS, (unsigned long long)d.nt);
      atomicAdd(d.counters + HB_CNT_ELEM_WARPSTEPS_RUN, (unsigned long long)n_wrun);
    }
  }
}

}  // namespace hb
