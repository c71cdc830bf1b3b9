// land_kernel.cuh — runoff generation: hland_96 + evap_aet_hbv96 + evap_pet_hbv96 + rconc_uh as ONE kernel.
//
// Replaces the generated Cython of the reference (c_hland_96.pyx & friends, generator
// hydpy/cythons/modelutils.py:758-2999) for `Model.simulate_period` (modelutils.py:1625-1634).
//
// Decomposition: one thread per element (ensemble members are elements), persistent over the whole time window.
//   * uz, lz and all element constants live in registers for all steps;
//   * zone states (ic, sm, sp[c], wc[c]) and the folded zone constants are element-interleaved SoA
//     ([slot][E_pad], consecutive threads = consecutive elements ⇒ every access is a coalesced 256-byte warp
//     transaction); the states stay L2-resident between steps;
//   * the reference runs method-major (every method over all zones); here a step runs zone-major, which yields
//     bit-identical results because zones only couple through (a) the uz value of the PREVIOUS step, (b) ordered
//     sums/products that are accumulated in zone order, and (c) snow redistribution, which gets its own two-phase
//     path (`sred` elements);
//   * forcing is read through the read-only path, 4 doubles per element-step.
//
// Arithmetic: IEEE binary64, compiled with -fmad=false so that every a*b+c rounds twice exactly like the
// reference's gcc -O2 x86-64 build; min/max follow Cython's lowering ((b<a)?b:a / (b>a)?b:a).  The only deviations
// from the CPU path are pow/exp (fastmath.cuh: relative error < 1e-14 for the model's arguments, glibc < 1 ulp) and
// the reciprocal-multiply inside the UZ sub-step loop.  sin() is evaluated on the host with glibc.
#pragma once
#include <cstdint>

#include "../../include/hydpy_b200.h"
#include "fastmath.cuh"

namespace hb {

#define HB_PYMAX(a, b) (((b) > (a)) ? (b) : (a))
#define HB_PYMIN(a, b) (((b) < (a)) ? (b) : (a))

// folded per-zone constants (host: fold_zone_constants in engine.cu); rows of kz[slot][KZ_COUNT][E_pad]
enum kz {
  KZ_TCORR = 0,   // tcorr
  KZ_TCALT_DZ,    // tcalt*(zonez-z)
  KZ_TT_HI,       // tt + ttint/2
  KZ_TT_LO,       // tt - ttint/2
  KZ_TTINT,
  KZ_PCALT_F,     // 1 + pcalt*(zonez-z)
  KZ_PCORR, KZ_RFCF, KZ_SFCF, KZ_ICMAX,
  KZ_CFMAX, KZ_CFVAR, KZ_TTM,
  KZ_CFR_CFMAX,   // cfr*cfmax
  KZ_WHC, KZ_FC, KZ_BETA, KZ_CFLUX,
  KZ_W_UZ,        // relzonearea/relupperzonearea
  KZ_W_CONTRI,    // relzonearea/relsoilarea
  KZ_THRESH,      // soilmoisturelimit*maxsoilwater
  KZ_EXCESSRED,
  KZ_ATF,         // airtemperaturefactor
  KZ_ETF,         // evapotranspirationfactor
  KZ_ALT_F,       // 1 - altitudefactor/100*(hrualtitude-altitude)
  KZ_NEG_PF,      // -precipitationfactor
  // rarely used (loaded only by the zone types / options that need them)
  KZ_SMAX, KZ_GMELT, KZ_GVAR,
  KZ_W_LZ,        // relzonearea/rellowerzonearea
  KZ_RELZONEAREA, KZ_TTI,
  // reciprocals used by the fast path only (land_kernel_v2.cuh)
  KZ_INV_TTINT, KZ_INV_RFCF, KZ_INV_SFCF, KZ_INV_FC, KZ_INV_THRESH,
  // sibling models (a zone belongs to one model, so the rows are shared):
  //   hland_96p: sgr, 1-w0, 1-w1, k2, w2, (1-w2)*k2, sg1max        hland_96c: h1, tab1, tvs1, h2, tab2, tvs2
  KZ_X0, KZ_X1, KZ_X2, KZ_X3, KZ_X4, KZ_X5, KZ_X6,
  KZ_COUNT
};
// hland_96c weighs its zone outflows with relzonearea/rellandarea (Calc_InRC_V3); the row of the (unused) upper-zone weight holds it
#define KZ_W_LAND KZ_W_UZ

// per-element constants; rows of ke[KE_COUNT][E_pad]
enum ke {
  KE_DT = 0,
  KE_DT_PERCMAX,      // dt*percmax
  KE_DT_K,            // dt*k
  KE_ALPHA1,          // 1+alpha
  KE_K4, KE_GAMMA1,   // 1+gamma
  KE_RELUPPER, KE_RELLOWER,
  KE_UP_LOW,          // relupperzonearea/rellowerzonearea
  KE_QFACTOR, KE_RELSOILAREA, KE_RELLANDAREA,
  KE_BETA_LAST,       // beta of the last zone (Calc_ContriArea_V1 quirk, ref: hland_model.py:2236)
  // hland_96p (KE_DT_PERCMAX = percmax because dt = 1 there, KE_K4 = derived k4)
  KE_K3, KE_W3, KE_W4, KE_FSG,
  KE_OMFSG,           // 1 - fsg
  // rconc_nash
  KE_NASH_DT,
  KE_NASH_DTKSC,      // dt*ksc
  KE_COUNT
};

// element info bits (einfo[e])
#define HB_EI_RESPAREA 1   // resparea && relsoilarea > 0
#define HB_EI_LAKE 2       // has ILAKE zones
#define HB_EI_SEALED 4     // has SEALED zones
#define HB_EI_SNOWLIMIT 8  // some zone has finite SMax ⇒ Calc_SPL/Calc_SPG are live
#define HB_EI_FAST 16      // eligible for the two-kernel fast path: one snow class, no snow limit, every CFlux == 0
#define HB_EI_SOILONLY 32  // all zones are FIELD or FOREST
#define HB_EI_COUPLED 64   // like HB_EI_FAST but with CFlux > 0 somewhere: fused kernel (land_kernel_coupled.cuh)
#define HB_EI_FUSED 128    // HB_EI_FAST hland_96 element of a 32-element group whose zones are all FIELD/FOREST with FC > 0:
                           // the zone kernel can add the group's zone terms itself (land_kernel_v2.cuh, zone_kernel<…, ZW>)
#define HB_EI_SPEC 1024     // HB_EI_FAST hland_96 element with HB_EI_SNOWLIMIT: the fast path runs it SPECULATIVELY — as long as no
                           // snow class of any zone exceeds its SMax, Calc_SPL/Calc_SPG change nothing (all losses, gains and
                           // excesses are zero); a window in which one does is repeated by the general kernel (engine.cu)
#define HB_EI_MODEL_SHIFT 8   // bits 8..9: HB_MODEL_* of the element
#define HB_EI_MODEL(i) (((i) >> HB_EI_MODEL_SHIFT) & 3)

// zone info: bits 0..2 zone type, bits 8.. zflags (HB_ZF_*)
#define HB_ZI_TYPE(i) ((i) & 7)
#define HB_ZI_FLAGS(i) ((i) >> 8)

struct LandDev {
  int64_t n_elem, e_pad, n_zone, n_snow, n_uh;
  const int64_t* uh0;      // [E_pad] first storage / ordinate of the element in ABI order (series HB_S_SC)
  int32_t zmax, cmax, umax;
  int64_t n_list;          // number of elements this launch works on
  const int32_t* elist;    // [n_list] their indices (NULL: elements 0..n_list-1)
  const int32_t* only_flagged;  // [E_pad] or NULL: work on the listed elements whose flag is set only (speculative fast
                                // path: the elements whose snow exceeded SMax in this window, states restored beforehand)
  int32_t* flag_out;            // [E_pad] or NULL: … and leave 2 there if the element lost snow to Calc_SPL_WCL_SP_WC_V1 in
                                // this window (it then skips the speculation of the next window), else 0
  // structure
  const int32_t* nz;       // [E_pad]
  const int32_t* nc;       // [E_pad]
  const int32_t* nu;       // [E_pad]
  const int32_t* recstep;  // [E_pad]
  const int32_t* einfo;    // [E_pad]
  const int32_t* fcol;     // [E_pad]
  const int32_t* zinfo;    // [zmax][E_pad]
  const int64_t* zone0;    // [E_pad] first zone in ABI order (series output)
  const int64_t* snow0;    // [E_pad]
  // constants
  const double* kz;        // [zmax][KZ_COUNT][E_pad]
  const double* ke;        // [KE_COUNT][E_pad]
  const double* sfdist;    // [cmax][E_pad]
  const double* uh;        // [umax][E_pad]
  // snow redistribution (CSR in ABI order, local zone indices)
  const int64_t* sred_ptr; // [E_pad+1]
  const int32_t* sred_from;
  const int32_t* sred_to;
  const double* sred_adjust;
  const int32_t* zorder;   // [zmax][E_pad] hland_derived.IndicesZoneZ
  // states
  double* ic;              // [zmax][E_pad]
  double* sm;              // [zmax][E_pad]
  double* sp;              // [cmax][zmax][E_pad]
  double* wc;              // [cmax][zmax][E_pad]
  double* uz;              // [E_pad]
  double* lz;              // [E_pad]
  double* quh;             // [umax][E_pad]  rconc_uh: logs.quh; rconc_nash: states.sc
  const int32_t* nash_steps;  // [E_pad] 0: rconc_uh or none; > 0: rconc_nash NmbSteps; -1: rconc_nash passing its inflow on
  // states of the sibling models: hland_96p suz, sg1 | hland_96c bw1, bw2 per zone; hland_96p sg2, sg3 per element
  double* zx1;             // [zmax][E_pad]
  double* zx2;             // [zmax][E_pad]
  double* ex1;             // [E_pad]
  double* ex2;             // [E_pad]
  // scratch
  double* scr_a;           // [zmax][E_pad]  r of SEALED zones / pc of ILAKE zones
  double* scr_b;           // [zmax][E_pad]  el of ILAKE zones
  double* scr_sred;        // [6][zmax][E_pad] spl, wcl, spg, wcg, spe, wce (only if any element has HB_EI_SNOWLIMIT)
  // window
  int64_t nt, n_col;       // steps of this window
  int64_t nt_total, t0;    // staged forcing block and offset of the window inside it
  const double* forcing;   // [4][nt_total][n_col]
  const double* sinfactor; // [nt_total]
  const MathTables* tables;  // log/exp tables (fastmath.cuh), staged into shared memory by every block
  double* qt;              // [nt][E_pad]
  double* series[HB_S_COUNT];  // optional, ABI layouts
};

__device__ __forceinline__ double ldg(const double* p) { return __ldg(p); }

// One zone of one step, front part: temperature, rain fraction, corrected precipitation.
// ref: hland_model.py:71-77 Calc_TC_V1, 125-135 Calc_FracRain_V1, 213-225 Calc_PC_V1
__device__ __forceinline__ void zone_front(const double* __restrict__ kzk, int64_t ep, double in_p, double in_t,
                                           double& tc, double& fracrain, double& pc) {
  tc = in_t + ldg(kzk + KZ_TCORR * ep) - ldg(kzk + KZ_TCALT_DZ * ep);
  const double tt_hi = ldg(kzk + KZ_TT_HI * ep), tt_lo = ldg(kzk + KZ_TT_LO * ep);
  if (tc >= tt_hi) fracrain = 1.0;
  else if (tc <= tt_lo) fracrain = 0.0;
  else fracrain = (tc - tt_lo) / ldg(kzk + KZ_TTINT * ep);
  pc = in_p * ldg(kzk + KZ_PCALT_F * ep);
  if (pc <= 0.0) pc = 0.0;
  else {
    const double x = fracrain;
    pc *= ldg(kzk + KZ_PCORR * ep) / (x / ldg(kzk + KZ_RFCF * ep) + (1.0 - x) / ldg(kzk + KZ_SFCF * ep));
  }
}

// ref: hland_model.py:2550-2619 Calc_QAb_QVs_BW_V1 (hland_96c): the reference's tail recursion (t0 := t1) as a loop.
// The iteration cap turns a non-advancing t1 — on which the reference would exhaust its stack — into a NaN result
// instead of a hang.  FAST = false (general kernel): libdevice `log`, IEEE divisions, the reference's operations one by
// one.  FAST = true (thread-per-zone kernel, twice per zone-step): the table logarithm of fastmath.cuh and divisions as
// multiplications by correctly rounded reciprocals (`__drcp_rn`: half the instructions of a division, <= 1 ulp from it) —
// inside the 1e-9 parity bar like the other reciprocals of the fast path.
#define HB_QDIV(a, b) (FAST ? (a) * __drcp_rn(b) : (a) / (b))
template <bool FAST>
__device__ __noinline__ void qab_qvs_bw_t(double h_, double k1_, double k2_, double& s0, double qz_, double& qa1,
                                          double& qa2, TabRef T) {
  double t0 = 0.0;
  for (int it = 0; it < 4096; ++it) {
    double s0_ = s0, t1, dt_;
    if ((k1_ == 0.0) && (s0_ > h_)) {
      qa1 += s0_ - h_;
      s0 = s0_ = h_;
    }
    const double h_k2 = HB_QDIV(h_, k2_);
    if ((k1_ == 0.0) && (s0_ == h_) && (qz_ > h_k2)) {
      const double qa2_ = h_k2;
      dt_ = 1.0 - t0;
      qa2 += dt_ * qa2_;
      qa1 += dt_ * (qz_ - qa2_);
      return;
    } else if (k2_ == 0.0) {
      qa2 += s0_ + qz_;
      s0 = 0.0;
      return;
    } else if ((s0_ < h_) || (s0_ == h_ && qz_ <= h_k2)) {
      if ((s0_ == h_) || (qz_ <= h_k2)) t1 = 1.0;
      else if (isinf(k2_)) t1 = HB_QDIV(h_ - s0_, qz_);
      else {
        const double arg = HB_QDIV(qz_ - HB_QDIV(s0_, k2_), qz_ - h_k2);
        t1 = t0 + k2_ * (FAST ? hb_log(arg, T) : log(arg));
      }
      if (0.0 < t1 && t1 < 1.0) {
        qa2 += (t1 - t0) * qz_ - (h_ - s0_);
        s0 = h_;
        t0 = t1;
        continue;
      } else if (isinf(k2_)) {
        s0 += (1.0 - t0) * qz_;
      } else {
        dt_ = 1.0 - t0;
        const double k2qz = k2_ * qz_;
        s0 = k2qz - (k2qz - s0_) * hb_exp(HB_QDIV(-dt_, k2_), T);
        qa2 += s0_ - s0 + dt_ * qz_;
      }
      return;
    } else {
      const double v1 = HB_QDIV(1.0, k1_) + HB_QDIV(1.0, k2_);
      const double v2 = qz_ + HB_QDIV(h_, k1_);
      const double nom = v2 - h_ * v1;
      const double denom = v2 - s0_ * v1;
      bool one = (s0_ == h_) || (denom == 0.0);
      double q = 1.0;
      if (!one) { q = HB_QDIV(nom, denom); one = !(0.0 < q && q <= 1.0); }
      if (one) t1 = 1.0;
      else {
        t1 = t0 - HB_QDIV(1.0, v1) * (FAST ? hb_log(q, T) : log(HB_QDIV(nom, denom)));
        t1 = HB_PYMIN(t1, 1.0);
      }
      dt_ = t1 - t0;
      const double v3 = HB_QDIV(v2 * dt_, v1);
      const double v4 = HB_QDIV(denom, v1 * v1) * (1.0 - hb_exp(-dt_ * v1, T));
      const double qa1_ = HB_QDIV(v3 - v4 - h_ * dt_, k1_);
      const double qa2_ = HB_QDIV(v3 - v4, k2_);
      qa1 += qa1_;
      qa2 += qa2_;
      if (t1 == 1.0) s0 += dt_ * qz_ - qa1_ - qa2_;
      else s0 = h_;
      if (t1 < 1.0) { t0 = t1; continue; }
      return;
    }
  }
  s0 = qa1 = qa2 = __longlong_as_double(0x7ff8000000000000LL);
}
#undef HB_QDIV
__device__ __forceinline__ void qab_qvs_bw(double h_, double k1_, double k2_, double& s0, double qz_, double& qa1,
                                           double& qa2, TabRef T) {
  qab_qvs_bw_t<false>(h_, k1_, k2_, s0, qz_, qa1, qa2, T);
}

// ref: hland_model.py:4109-4116 Calc_OutRC_V1; rconc_model.py:89-95 Determine_Outflow_V1 (rconc_uh), :179-197
// Determine_Outflow_V2 (rconc_nash: explicit Euler over NmbSteps sub-steps through the storage cascade)
// STAGED: keep up to HB_NASH_STAGE storages of the cascade in the block's shared memory ([storage][thread], conflict-free)
// for the NmbSteps sub-steps of a step.  The thread-per-element kernel of the fast path is bound by the latency of this
// chain; through global memory every one of the NmbSteps x NmbStorages updates is a dependent load-modify-store (and
// registers for the cascade would cost the occupancy that hides the chain: measured 3.8 -> 5.1 ms).
#define HB_NASH_STAGE 12
template <bool STAGED = false>
__device__ __forceinline__ double rconc_outflow(const double* __restrict__ uh, double* __restrict__ quh,
                                                const double* __restrict__ ke, int64_t ep, int64_t e, int nu,
                                                int nash_steps, double inrc, double* s_sc = nullptr, int s_stride = 0) {
  if (nu <= 0 || nash_steps < 0) return inrc;
  if (nash_steps > 0) {
    const double dt = ke[KE_NASH_DT * ep + e], dtksc = ke[KE_NASH_DTKSC * ep + e];
    double outflow = 0.0;
    if (STAGED && s_sc != nullptr && nu <= HB_NASH_STAGE) {
      for (int j = 0; j < nu; ++j) s_sc[j * s_stride] = quh[(int64_t)j * ep + e];
      const double add = dt * inrc;
      for (int i = 0; i < nash_steps; ++i) {
        double carry = add;
#pragma unroll
        for (int j = 0; j < HB_NASH_STAGE; ++j)
          if (j < nu) {
            const double v = s_sc[j * s_stride] + carry;
            const double a = dtksc * v;
            const double q = HB_PYMIN(a, v);
            s_sc[j * s_stride] = v - q;
            carry = q;
          }
        outflow += carry;
      }
      for (int j = 0; j < nu; ++j) quh[(int64_t)j * ep + e] = s_sc[j * s_stride];
      return outflow;
    }
    for (int i = 0; i < nash_steps; ++i) {
      double carry = dt * inrc;                       // what enters storage j in this sub-step
      for (int j = 0; j < nu; ++j) {
        double sc = quh[(int64_t)j * ep + e] + carry;
        const double a = dtksc * sc;
        const double q = HB_PYMIN(a, sc);
        sc -= q;
        quh[(int64_t)j * ep + e] = sc;
        carry = q;
      }
      outflow += carry;
    }
    return outflow;
  }
  const double outrc = ldg(uh + e) * inrc + quh[e];
  for (int j = 1; j < nu; ++j) quh[(int64_t)(j - 1) * ep + e] = ldg(uh + (int64_t)j * ep + e) * inrc + quh[(int64_t)j * ep + e];
  return outrc;
}

// rconc_nash states.sc after the step (ref: hydpy/models/rconc/rconc_model.py:179-197; save_data keeps the new state)
__device__ __forceinline__ void record_sc(double* series, int64_t t, int64_t n_uh, int64_t uh0, const double* quh, int64_t ep,
                                          int64_t e, int nu, int nash_steps) {
  if (!series || nash_steps == 0) return;
  for (int j = 0; j < nu; ++j) series[t * n_uh + uh0 + j] = quh[(int64_t)j * ep + e];
}

template <bool SERIES, int MODEL>
__global__ void __launch_bounds__(128, 4) land_kernel(const LandDev d) {
  __shared__ MathTables s_tables;
  {
    const double* src = reinterpret_cast<const double*>(d.tables);
    double* dst = reinterpret_cast<double*>(&s_tables);
    for (int i = threadIdx.x; i < (int)(sizeof(MathTables) / sizeof(double)); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const TabRef T = hb_tabref(&s_tables);
  const int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= d.n_list) return;
  const int64_t e = d.elist ? (int64_t)d.elist[li] : li;
  if (d.only_flagged && !d.only_flagged[e]) return;
  const int64_t ep = d.e_pad;
  const int nz = d.nz[e], nc = d.nc[e], nu = d.nu[e], recstep = d.recstep[e], einfo = d.einfo[e];
  const int64_t col = d.fcol[e];
  const double dt = d.ke[KE_DT * ep + e], dt_percmax = d.ke[KE_DT_PERCMAX * ep + e], dt_k = d.ke[KE_DT_K * ep + e],
               alpha1 = d.ke[KE_ALPHA1 * ep + e], k4 = d.ke[KE_K4 * ep + e], gamma1 = d.ke[KE_GAMMA1 * ep + e],
               relupper = d.ke[KE_RELUPPER * ep + e], rellower = d.ke[KE_RELLOWER * ep + e],
               up_low = d.ke[KE_UP_LOW * ep + e], qfactor = d.ke[KE_QFACTOR * ep + e],
               beta_last = d.ke[KE_BETA_LAST * ep + e];
  const double dnc = (double)nc;
  const bool snowlimit = (einfo & HB_EI_SNOWLIMIT) != 0;
  bool lost = false;       // some snow class exceeded its SMax in this window
  const int64_t zone0 = SERIES ? d.zone0[e] : 0, snow0 = SERIES ? d.snow0[e] : 0;
  double uz = d.uz[e], lz = d.lz[e];
  const int nash_steps = d.nash_steps ? d.nash_steps[e] : 0;
  // hland_96p: slow-response groundwater storages of the element
  double sg2 = 0.0, sg3 = 0.0;
  if (MODEL == HB_MODEL_HLAND_96P) { sg2 = d.ex1[e]; sg3 = d.ex2[e]; }
  const int64_t zstride = (int64_t)d.zmax * ep;  // class stride of sp/wc
  const double* fP = d.forcing + (0 * d.nt_total + d.t0) * d.n_col + col;
  const double* fT = d.forcing + (1 * d.nt_total + d.t0) * d.n_col + col;
  const double* fA = d.forcing + (2 * d.nt_total + d.t0) * d.n_col + col;
  const double* fE = d.forcing + (3 * d.nt_total + d.t0) * d.n_col + col;

#define SERZ(id, k, v) do { if (SERIES && d.series[id]) d.series[id][t * d.n_zone + zone0 + (k)] = (v); } while (0)
#define SERS(id, c, k, v) \
  do { if (SERIES && d.series[id]) d.series[id][t * d.n_snow + snow0 + (int64_t)(c) * nz + (k)] = (v); } while (0)
#define SERE(id, v) do { if (SERIES && d.series[id]) d.series[id][t * d.n_elem + e] = (v); } while (0)

  for (int64_t t = 0; t < d.nt; ++t) {
    const double in_p = ldg(fP + t * d.n_col), in_t = ldg(fT + t * d.n_col), in_nat = ldg(fA + t * d.n_col),
                 in_net = ldg(fE + t * d.n_col);
    const double sinf = ldg(d.sinfactor + d.t0 + t);
    const double d_t = in_t - in_nat;       // meanairtemperature - normalairtemperature
    const double ret_hi = 2.0 * in_net;

    // ---- phase 1 (only elements with a finite SMax): everything up to Calc_SPL, then Calc_SPG over all zones ----
    if (snowlimit) {
      double* spl = d.scr_sred + 0 * zstride; double* wcl = d.scr_sred + 1 * zstride;
      double* spg = d.scr_sred + 2 * zstride; double* wcg = d.scr_sred + 3 * zstride;
      double* spe = d.scr_sred + 4 * zstride; double* wce = d.scr_sred + 5 * zstride;
      for (int k = 0; k < nz; ++k) {
        const int64_t ik = (int64_t)k * ep + e;
        const double* kzk = d.kz + (int64_t)k * KZ_COUNT * ep + e;
        const int zt = HB_ZI_TYPE(d.zinfo[ik]);
        double tc, fracrain, pc;
        zone_front(kzk, ep, in_p, in_t, tc, fracrain, pc);
        double ic = d.ic[ik], tf;
        // ref: hland_model.py:299-309 Calc_TF_Ic_V1
        if (zt == HB_FIELD || zt == HB_FOREST || zt == HB_SEALED) {
          const double a = pc - (ldg(kzk + KZ_ICMAX * ep) - ic);
          tf = HB_PYMAX(a, 0.0);
          ic += pc - tf;
        } else { tf = pc; ic = 0.0; }
        d.ic[ik] = ic;
        SERZ(HB_S_TF, k, tf);
        const double rain = tf * fracrain, snow = tf * (1.0 - fracrain);
        const double smax = ldg(kzk + KZ_SMAX * ep);
        double spl_k = 0.0, wcl_k = 0.0;
        for (int c = 0; c < nc; ++c) {
          const int64_t ick = (int64_t)c * zstride + ik;
          double sp = d.sp[ick], wc = d.wc[ick];
          if (zt != HB_ILAKE) {
            // ref: hland_model.py:484-499 Calc_SP_WC_V1
            const double sf = ldg(d.sfdist + (int64_t)c * ep + e);
            wc += sf * rain;
            sp += sf * snow;
            // ref: hland_model.py:584-605 Calc_SPL_WCL_SP_WC_V1
            if (!isinf(smax)) {
              const double snw = sp + wc, excess = snw - smax;
              if (excess > 0.0) {
                lost = true;
                const double excess_sp = excess * sp / snw, excess_wc = excess * wc / snw;
                spl_k += excess_sp / dnc;
                wcl_k += excess_wc / dnc;
                sp -= excess_sp;
                wc -= excess_wc;
              }
            }
          } else { sp = 0.0; wc = 0.0; }
          d.sp[ick] = sp;
          d.wc[ick] = wc;
        }
        spl[ik] = spl_k; wcl[ik] = wcl_k; spg[ik] = 0.0; wcg[ik] = 0.0; spe[ik] = 0.0; wce[ik] = 0.0;
      }
      // ref: hland_model.py:876-969 Calc_SPG_WCG_SP_WC_V1 (sequential graph walk of this element)
      do {
        const int64_t r0 = d.sred_ptr[e], nr = d.sred_ptr[e + 1] - r0;
        for (int64_t i = 0; i < nr; ++i) {
          const int64_t f = (int64_t)d.sred_from[r0 + i] * ep + e, tz = (int64_t)d.sred_to[r0 + i] * ep + e;
          const double adjust = d.sred_adjust[r0 + i];
          const double gain_frozen = adjust * (spl[f] + spe[f]);
          const double gain_liquid = adjust * (wcl[f] + wce[f]);
          const double gain_total = gain_frozen + gain_liquid;
          const double smax_t = d.kz[((int64_t)d.sred_to[r0 + i] * KZ_COUNT + KZ_SMAX) * ep + e];
          for (int c = 0; c < nc; ++c) {
            const double sf = d.sfdist[(int64_t)c * ep + e];
            const double gain_pot = sf * gain_total;
            if (gain_pot > 0.0) {
              const int64_t ict = (int64_t)c * zstride + tz;
              const double gain_max = smax_t - d.sp[ict] - d.wc[ict];
              const double q = gain_max / gain_pot;
              const double fraction_gain = HB_PYMIN(q, 1.0);
              const double factor_gain = fraction_gain * sf;
              spg[tz] += factor_gain * gain_frozen / dnc;
              wcg[tz] += factor_gain * gain_liquid / dnc;
              d.sp[ict] += factor_gain * gain_frozen;
              d.wc[ict] += factor_gain * gain_liquid;
              const double factor_excess = (1.0 - fraction_gain) * sf;
              spe[tz] += factor_excess * gain_frozen / dnc;
              wce[tz] += factor_excess * gain_liquid / dnc;
            }
          }
        }
        double excess_frozen_basin = 0.0, excess_liquid_basin = 0.0;
        for (int k = 0; k < nz; ++k) {
          const int64_t ik = (int64_t)k * ep + e;
          if (HB_ZI_FLAGS(d.zinfo[ik]) & HB_ZF_SREDEND) {
            const double rza = d.kz[((int64_t)k * KZ_COUNT + KZ_RELZONEAREA) * ep + e];
            excess_frozen_basin += rza * (spe[ik] + spl[ik]);
            excess_liquid_basin += rza * (wce[ik] + wcl[ik]);
          }
        }
        if ((excess_frozen_basin + excess_liquid_basin) <= 0.0) break;
        bool done = false;
        for (int i = 0; i < nz; ++i) {
          const int tl = d.zorder[(int64_t)i * ep + e];
          const int64_t tz = (int64_t)tl * ep + e;
          if (HB_ZI_TYPE(d.zinfo[tz]) == HB_ILAKE) continue;
          const double rza = d.kz[((int64_t)tl * KZ_COUNT + KZ_RELZONEAREA) * ep + e];
          const double smax_t = d.kz[((int64_t)tl * KZ_COUNT + KZ_SMAX) * ep + e];
          const double excess_frozen_zone = excess_frozen_basin / rza;
          const double excess_liquid_zone = excess_liquid_basin / rza;
          const double excess_total_zone = excess_frozen_zone + excess_liquid_zone;
          double gain_max_cum = 0.0;
          for (int c = 0; c < nc; ++c) {
            const int64_t ict = (int64_t)c * zstride + tz;
            gain_max_cum += smax_t - d.sp[ict] - d.wc[ict];
          }
          if (gain_max_cum <= 0.0) continue;
          const double q = gain_max_cum / dnc / excess_total_zone;
          const double fraction_gain_zone = HB_PYMIN(q, 1.0);
          const double efa = fraction_gain_zone * excess_frozen_zone;
          const double ela = fraction_gain_zone * excess_liquid_zone;
          for (int c = 0; c < nc; ++c) {
            const int64_t ict = (int64_t)c * zstride + tz;
            const double fraction_gain_class = (smax_t - d.sp[ict] - d.wc[ict]) / gain_max_cum;
            const double delta_sp_zone = fraction_gain_class * efa;
            const double delta_wc_zone = fraction_gain_class * ela;
            spg[tz] += delta_sp_zone;
            wcg[tz] += delta_wc_zone;
            d.sp[ict] += delta_sp_zone * dnc;
            d.wc[ict] += delta_wc_zone * dnc;
          }
          excess_frozen_basin -= efa * rza;
          excess_liquid_basin -= ela * rza;
          if ((excess_frozen_basin + excess_liquid_basin) <= 0.0) { done = true; break; }
        }
        if (done) break;
        const double rellandarea = d.ke[KE_RELLANDAREA * ep + e];
        const double excess_frozen_land = excess_frozen_basin / rellandarea;
        const double excess_liquid_land = excess_liquid_basin / rellandarea;
        for (int k = 0; k < nz; ++k) {
          const int64_t ik = (int64_t)k * ep + e;
          if (HB_ZI_TYPE(d.zinfo[ik]) != HB_ILAKE) {
            spg[ik] += excess_frozen_land;
            wcg[ik] += excess_liquid_land;
            for (int c = 0; c < nc; ++c) {
              const int64_t ick = (int64_t)c * zstride + ik;
              d.sp[ick] += excess_frozen_land;
              d.wc[ick] += excess_liquid_land;
            }
          }
        }
      } while (0);
    }

    // ---- zone loop (all remaining zone-level methods) ----------------------------------------------------------
    double inuz = 0.0, contriarea = 1.0;
    double gr2 = 0.0, gr3 = 0.0;      // hland_96p: Calc_GR2_GR3_V1 sums
    double inrc_c = 0.0;              // hland_96c: Calc_InRC_V3 sum
    for (int k = 0; k < nz; ++k) {
      const int64_t ik = (int64_t)k * ep + e;
      const double* kzk = d.kz + (int64_t)k * KZ_COUNT * ep + e;
      const int zi = d.zinfo[ik];
      const int zt = HB_ZI_TYPE(zi), zf = HB_ZI_FLAGS(zi);
      double tc, fracrain, pc;
      zone_front(kzk, ep, in_p, in_t, tc, fracrain, pc);
      double ic = d.ic[ik];
      double rain = 0.0, snow = 0.0, tf = pc;
      if (!snowlimit) {
        // ref: hland_model.py:299-309 Calc_TF_Ic_V1
        if (zt == HB_FIELD || zt == HB_FOREST || zt == HB_SEALED) {
          const double a = pc - (ldg(kzk + KZ_ICMAX * ep) - ic);
          tf = HB_PYMAX(a, 0.0);
          ic += pc - tf;
        } else { tf = pc; ic = 0.0; }
        rain = tf * fracrain;
        snow = tf * (1.0 - fracrain);
        SERZ(HB_S_TF, k, tf);
        SERZ(HB_S_SPL, k, 0.0); SERZ(HB_S_WCL, k, 0.0); SERZ(HB_S_SPG, k, 0.0); SERZ(HB_S_WCG, k, 0.0);
      } else if (SERIES) {
        SERZ(HB_S_SPL, k, d.scr_sred[0 * zstride + ik]); SERZ(HB_S_WCL, k, d.scr_sred[1 * zstride + ik]);
        SERZ(HB_S_SPG, k, d.scr_sred[2 * zstride + ik]); SERZ(HB_S_WCG, k, d.scr_sred[3 * zstride + ik]);
      }
      SERZ(HB_S_TC, k, tc); SERZ(HB_S_FRACRAIN, k, fracrain); SERZ(HB_S_PC, k, pc);
      const double ttm = ldg(kzk + KZ_TTM * ep);
      // ref: hland_model.py:1050-1062 Calc_CFAct_V1
      double cfact = 0.0;
      if (zt != HB_ILAKE) { const double a = ldg(kzk + KZ_CFMAX * ep) + sinf * ldg(kzk + KZ_CFVAR * ep); cfact = HB_PYMAX(a, 0.0); }
      SERZ(HB_S_CFACT, k, cfact);
      // ref: hland_model.py:1607-1619 Calc_GAct_V1
      double gact = 0.0;
      if (zt == HB_GLACIER) { const double a = ldg(kzk + KZ_GMELT * ep) + sinf * ldg(kzk + KZ_GVAR * ep); gact = HB_PYMAX(a, 0.0); }
      SERZ(HB_S_GACT, k, gact);
      double in_ = 0.0, ncover = 0.0;
      int nbare = 0;
      if (zt != HB_ILAKE) {
        const double potmelt = cfact * (tc - ttm);
        const double potrefr = ldg(kzk + KZ_CFR_CFMAX * ep) * (ttm - tc);
        const double whc = ldg(kzk + KZ_WHC * ep);
        for (int c = 0; c < nc; ++c) {
          const int64_t ick = (int64_t)c * zstride + ik;
          double sp = d.sp[ick], wc = d.wc[ick];
          if (!snowlimit) {
            // ref: hland_model.py:484-499 Calc_SP_WC_V1
            const double sf = ldg(d.sfdist + (int64_t)c * ep + e);
            wc += sf * rain;
            sp += sf * snow;
          }
          // ref: hland_model.py:1166-1187 Calc_Melt_SP_WC_V1
          double melt = 0.0;
          if (tc > ttm) { melt = HB_PYMIN(potmelt, sp); sp -= melt; wc += melt; }
          // ref: hland_model.py:1318-1341 Calc_Refr_SP_WC_V1
          double refr = 0.0;
          if (tc < ttm) { refr = HB_PYMIN(potrefr, wc); sp += refr; wc -= refr; }
          // ref: hland_model.py:1434-1448 Calc_In_WC_V1
          const double wc_old = wc, lim = whc * sp;
          wc = HB_PYMIN(wc_old, lim);
          in_ += (wc_old - wc) / dnc;
          d.sp[ick] = sp;
          d.wc[ick] = wc;
          nbare += (sp <= 0.0) ? 1 : 0;
          ncover += (double)(sp > 0.0);
          SERS(HB_S_MELT, c, k, melt); SERS(HB_S_REFR, c, k, refr);
          SERS(HB_S_SWE, c, k, sp + wc);   // ref: hland_model.py:1485-1495 Calc_SWE_V1
          SERS(HB_S_SP, c, k, sp); SERS(HB_S_WC, c, k, wc);
        }
      } else {
        in_ = tf;
        for (int c = 0; c < nc; ++c) {
          const int64_t ick = (int64_t)c * zstride + ik;
          d.sp[ick] = 0.0;
          d.wc[ick] = 0.0;
          nbare += 1;
          SERS(HB_S_MELT, c, k, 0.0); SERS(HB_S_REFR, c, k, 0.0); SERS(HB_S_SWE, c, k, 0.0);
          SERS(HB_S_SP, c, k, 0.0); SERS(HB_S_WC, c, k, 0.0);
        }
      }
      // ref: hland_model.py:1528-1535 Calc_SR_V1
      const double sr = (zt == HB_SEALED) ? in_ : 0.0;
      SERZ(HB_S_SR, k, sr);
      // ref: hland_model.py:1688-1701 Calc_GlMelt_In_V1
      double glmelt = 0.0;
      if (zt == HB_GLACIER && tc > ttm) {
        const double glmeltpot = gact / dnc * (tc - ttm);
        for (int c = 0; c < nbare; ++c) { glmelt += glmeltpot; in_ += glmeltpot; }
      }
      SERZ(HB_S_GLMELT, k, glmelt);
      SERZ(HB_S_IN, k, in_);
      // evap_pet_hbv96.run(): ref: evap_model.py:5111-5125, 5420-5424, 5312-5326
      double pet;
      {
        double ret = in_net * (1.0 + ldg(kzk + KZ_ATF * ep) * d_t);
        const double lo = HB_PYMAX(ret, 0.0);
        ret = HB_PYMIN(lo, ret_hi);
        ret *= ldg(kzk + KZ_ETF * ep);
        pet = ret * ldg(kzk + KZ_ALT_F * ep);
        if (pet <= 0.0) pet = 0.0;
        else pet *= hb_exp(ldg(kzk + KZ_NEG_PF * ep) * pc, T);
        SERZ(HB_S_REFERENCEEVAPOTRANSPIRATION, k, ret); SERZ(HB_S_POTENTIALEVAPOTRANSPIRATION, k, pet);
      }
      // ref: evap_model.py:6068-6086 Calc_InterceptionEvaporation_V1 (f = 1: no snow evaporation, hland_model.py:4425-4430)
      double aet_ie = 0.0;
      if (zf & HB_ZF_INTERCEPTION) aet_ie = HB_PYMIN(pet, ic);
      if (SERIES) {   // sequences of the evap submodels: what their save_data keeps at the end of the step
        SERZ(HB_S_INTERCEPTEDWATER, k, ic); SERZ(HB_S_INTERCEPTIONEVAPORATION, k, aet_ie);
        SERZ(HB_S_SNOWCOVER, k, ncover / dnc);
        SERZ(HB_S_WATEREVAPORATION, k, ((zf & HB_ZF_WATER) && tc > ldg(kzk + KZ_TTI * ep)) ? pet : 0.0);
      }
      double aet_set = 0.0, aet_sw = 0.0;
      // ref: hland_model.py:383-396 Calc_EI_Ic_AETModel_V1
      double ei = 0.0;
      if (zt == HB_FIELD || zt == HB_FOREST || zt == HB_SEALED) { ei = HB_PYMIN(aet_ie, ic); ic -= ei; }
      else ic = 0.0;
      d.ic[ik] = ic;
      SERZ(HB_S_EI, k, ei); SERZ(HB_S_IC, k, ic);
      double sm = 0.0, r, cf = 0.0, ea = 0.0;
      if (zt == HB_FIELD || zt == HB_FOREST) {
        sm = d.sm[ik];
        const double fc = ldg(kzk + KZ_FC * ep);
        // ref: hland_model.py:1776-1790 Calc_R_SM_V1
        if (fc > 0.0) {
          r = in_ * hb_pow(sm / fc, ldg(kzk + KZ_BETA * ep), T);
          const double a = sm + in_ - fc;
          r = HB_PYMAX(r, a);
        } else r = in_;
        sm += in_ - r;
        // ref: hland_model.py:1904-1919 Calc_CF_SM_V1 (uz = value at the end of the previous step); hland_96 only
        if (MODEL == HB_MODEL_HLAND_96 && fc > 0.0) {
          cf = ldg(kzk + KZ_CFLUX * ep) * (1.0 - sm / fc);
          const double a = uz + r;
          cf = HB_PYMIN(cf, a);
          const double b = fc - sm;
          cf = HB_PYMIN(cf, b);
        }
        sm += cf;
        aet_sw = sm;
        // ref: evap_model.py:7630-7658 Determine_SoilEvapotranspiration_V1
        double set = 0.0;
        if (zf & HB_ZF_SOIL) {
          // ref: evap_model.py:6657-6672 Calc_SoilEvapotranspiration_V1
          set = pet;
          if (set > 0.0) {
            if (sm < 0.0) set = 0.0;
            else {
              const double thresh = ldg(kzk + KZ_THRESH * ep);
              if (sm < thresh) set *= sm / thresh;
            }
          }
          // ref: evap_model.py:7001-7018 Update_SoilEvapotranspiration_V1
          const double rr = ldg(kzk + KZ_EXCESSRED * ep);
          const double petc = rr * pet + (1.0 - rr) * (pet + pet) / 2.0;
          const double excess = set + aet_ie - petc;
          if ((petc >= 0.0) && (excess > 0.0)) set -= rr * excess;
          else if ((petc < 0.0) && (excess < 0.0)) set -= rr * excess;
          // ref: evap_model.py:7060-7069 Update_SoilEvapotranspiration_V2
          set *= 1.0 - ncover / dnc;
        }
        aet_set = set;
        // ref: hland_model.py:2002-2018 Calc_EA_SM_AETModel_V1
        ea = HB_PYMIN(set, sm);
        sm -= ea;
        if (sm > fc) { r += sm - fc; sm = fc; }
        d.sm[ik] = sm;
        if (MODEL == HB_MODEL_HLAND_96) {
          // ref: hland_model.py:2092-2101 Calc_InUZ_V1
          inuz += ldg(kzk + KZ_W_UZ * ep) * (r - cf);
          // ref: hland_model.py:2224-2236 Calc_ContriArea_V1
          if ((einfo & HB_EI_RESPAREA) && fc > 0.0) contriarea *= hb_pow(sm / fc, ldg(kzk + KZ_W_CONTRI * ep), T);
        }
      } else {
        r = in_;
        d.sm[ik] = 0.0;
        if (zt == HB_GLACIER) { if (MODEL == HB_MODEL_HLAND_96) inuz += ldg(kzk + KZ_W_UZ * ep) * (r - cf); }
        else if (zt == HB_SEALED) d.scr_a[ik] = r;
        else {  // ILAKE
          d.scr_a[ik] = pc;
          // ref: evap_model.py:5673-5681 Calc_WaterEvaporation_V1 (airtemperature = tc, hland_model.py:4242-4248)
          const double el = ((zf & HB_ZF_WATER) && tc > ldg(kzk + KZ_TTI * ep)) ? pet : 0.0;
          d.scr_b[ik] = el;
        }
      }
      SERZ(HB_S_R, k, r); SERZ(HB_S_CF, k, cf); SERZ(HB_S_EA, k, ea); SERZ(HB_S_SM, k, sm);
      SERZ(HB_S_EL, k, (zt == HB_ILAKE) ? d.scr_b[ik] : 0.0);
      if (SERIES) {
        if (!(zt == HB_FIELD || zt == HB_FOREST) && (zf & HB_ZF_SOIL)) {
          // (a zone without soil whose Soil flag is set: the submodel sees an empty soil; evap_model.py:6657-6672 …)
          double set = pet;
          const double thresh = ldg(kzk + KZ_THRESH * ep);
          if (set > 0.0 && 0.0 < thresh) set *= 0.0 / thresh;
          const double rr = ldg(kzk + KZ_EXCESSRED * ep);
          const double petc = rr * pet + (1.0 - rr) * (pet + pet) / 2.0;
          const double excess = set + aet_ie - petc;
          if ((petc >= 0.0) && (excess > 0.0)) set -= rr * excess;
          else if ((petc < 0.0) && (excess < 0.0)) set -= rr * excess;
          aet_set = set * (1.0 - ncover / dnc);
        }
        SERZ(HB_S_SOILWATER, k, aet_sw); SERZ(HB_S_SOILEVAPOTRANSPIRATION, k, aet_set);
      }
      if (MODEL == HB_MODEL_HLAND_96P) {
        // ---- hland_96p: upper zone layer and first-order groundwater storage of the zone (models/hland_96p.py) ----
        double suz = 0.0, sg1 = 0.0, dp = 0.0, rs = 0.0, ri = 0.0, gr1 = 0.0, rg1 = 0.0;
        if (zt == HB_FIELD || zt == HB_FOREST || zt == HB_GLACIER) {
          // ref: hland_model.py:2132-2140 Calc_SUZ_V1, 2525-2535 Calc_DP_SUZ_V1
          suz = d.zx1[ik] + r;
          dp = HB_PYMIN(suz, dt_percmax);          // dt = 1 for hland_96p elements: dt_percmax = PercMax
          suz -= dp;
          // ref: hland_model.py:3012-3033 Calc_RS_RI_SUZ_V1
          const double sgr = ldg(kzk + KZ_X0 * ep);
          if (suz > sgr) rs = (suz - sgr) * ldg(kzk + KZ_X1 * ep);
          ri = suz * ldg(kzk + KZ_X2 * ep);
          suz -= rs + ri;
          if (suz < 0.0) {
            const double f = 1.0 - suz / (rs + ri);
            rs *= f;
            ri *= f;
            suz = 0.0;
          }
          // ref: hland_model.py:3233-3243 Calc_GR1_V1
          const double sg1_old = d.zx2[ik], sg1max = ldg(kzk + KZ_X6 * ep);
          const double b = (sg1max - sg1_old) / ldg(kzk + KZ_X3 * ep);
          gr1 = HB_PYMIN(dp, b);
          const double a = sg1_old + gr1 - sg1max;
          gr1 -= HB_PYMAX(a, 0.0);
          // ref: hland_model.py:3290-3304 Calc_RG1_SG1_V1
          sg1 = ldg(kzk + KZ_X4 * ep) * sg1_old + ldg(kzk + KZ_X5 * ep) * gr1;
          rg1 = sg1_old + gr1 - sg1;
          // ref: hland_model.py:3975-3984 Calc_InRC_V2 (term; weighted and summed after the zone loop)
          d.scr_a[ik] = rs + ri + rg1;
        }
        d.zx1[ik] = suz;
        d.zx2[ik] = sg1;
        // ref: hland_model.py:3356-3372 Calc_GR2_GR3_V1
        if (zt != HB_SEALED) {
          const double weight = ldg(kzk + KZ_W_LZ * ep);
          const double total = (zt == HB_ILAKE) ? weight * pc : weight * (dp - gr1);
          gr2 += d.ke[KE_FSG * ep + e] * total;
          gr3 += d.ke[KE_OMFSG * ep + e] * total;
        }
        SERZ(HB_S_DP, k, dp); SERZ(HB_S_RS, k, rs); SERZ(HB_S_RI, k, ri); SERZ(HB_S_GR1, k, gr1); SERZ(HB_S_RG1, k, rg1);
        SERZ(HB_S_SUZ, k, suz); SERZ(HB_S_SG1, k, sg1);
      }
      if (MODEL == HB_MODEL_HLAND_96C) {
        // ---- hland_96c: surface-flow and interflow storages of the zone (models/hland_96c.py) ----------------------
        double bw1 = 0.0, bw2 = 0.0, qab1 = 0.0, qvs1 = 0.0, qab2 = 0.0, qvs2 = 0.0;
        if (zt == HB_FIELD || zt == HB_FOREST || zt == HB_GLACIER) {
          // ref: hland_model.py:2833-2853 Calc_QAb1_QVs1_BW1_V1, 2927-2947 Calc_QAb2_QVs2_BW2_V1
          bw1 = d.zx1[ik];
          qab_qvs_bw(ldg(kzk + KZ_X0 * ep), ldg(kzk + KZ_X1 * ep), ldg(kzk + KZ_X2 * ep), bw1, r, qab1, qvs1, T);
          bw2 = d.zx2[ik];
          qab_qvs_bw(ldg(kzk + KZ_X3 * ep), ldg(kzk + KZ_X4 * ep), ldg(kzk + KZ_X5 * ep), bw2, qvs1, qab2, qvs2, T);
          d.scr_a[ik] = qvs2;                       // Calc_LZ_V2 term
        }
        d.zx1[ik] = bw1;
        d.zx2[ik] = bw2;
        // ref: hland_model.py:4039-4051 Calc_InRC_V3
        if (zt != HB_ILAKE) {
          const double weight = ldg(kzk + KZ_W_LAND * ep);
          if (zt == HB_SEALED) inrc_c += weight * r;
          else inrc_c += weight * (qab1 + qab2);
        }
        SERZ(HB_S_QAB1, k, qab1); SERZ(HB_S_QVS1, k, qvs1); SERZ(HB_S_QAB2, k, qab2); SERZ(HB_S_QVS2, k, qvs2);
        SERZ(HB_S_BW1, k, bw1); SERZ(HB_S_BW2, k, bw2);
      }
    }
    if (MODEL == HB_MODEL_HLAND_96P) {
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
      // ref: hland_model.py:3444-3460 Calc_EL_SG2_SG3_AETModel_V1
      if (einfo & HB_EI_LAKE)
        for (int k = 0; k < nz; ++k) {
          const int64_t ik = (int64_t)k * ep + e;
          if (HB_ZI_TYPE(d.zinfo[ik]) == HB_ILAKE) {
            const double weight = d.kz[((int64_t)k * KZ_COUNT + KZ_W_LZ) * ep + e], el = d.scr_b[ik];
            sg2 -= d.ke[KE_FSG * ep + e] * weight * el;
            sg3 -= d.ke[KE_OMFSG * ep + e] * weight * el;
          }
        }
      // ref: hland_model.py:3975-3984 Calc_InRC_V2
      double inrc = rellower * (rg2 + rg3);
      for (int k = 0; k < nz; ++k) {
        const int64_t ik = (int64_t)k * ep + e;
        if (HB_ZI_TYPE(d.zinfo[ik]) != HB_ILAKE) inrc += d.kz[((int64_t)k * KZ_COUNT + KZ_RELZONEAREA) * ep + e] * d.scr_a[ik];
      }
      // ref: hland_model.py:4109-4116 Calc_OutRC_V1, 4139-4141 Calc_RT_V1, 4197-4200 Calc_QT_V1
      const double outrc = rconc_outflow(d.uh, d.quh, d.ke, ep, e, nu, nash_steps, inrc);
      if (SERIES) record_sc(d.series[HB_S_SC], t, d.n_uh, d.uh0[e], d.quh, ep, e, nu, nash_steps);
      const double qt = qfactor * outrc;
      d.qt[t * ep + e] = qt;
      SERE(HB_S_GR2, gr2); SERE(HB_S_GR3, gr3); SERE(HB_S_RG2, rg2); SERE(HB_S_RG3, rg3); SERE(HB_S_SG2, sg2);
      SERE(HB_S_SG3, sg3); SERE(HB_S_INRC, inrc); SERE(HB_S_OUTRC, outrc); SERE(HB_S_RT, outrc); SERE(HB_S_QT, qt);
      continue;
    }
    if (MODEL == HB_MODEL_HLAND_96C) {
      // ref: hland_model.py:4109-4116 Calc_OutRC_V1
      const double outrc = rconc_outflow(d.uh, d.quh, d.ke, ep, e, nu, nash_steps, inrc_c);
      if (SERIES) record_sc(d.series[HB_S_SC], t, d.n_uh, d.uh0[e], d.quh, ep, e, nu, nash_steps);
      // ref: hland_model.py:3172-3181 Calc_LZ_V2 (scr_a: pc of ILAKE zones, qvs2 of FIELD/FOREST/GLACIER zones)
      for (int k = 0; k < nz; ++k) {
        const int64_t ik = (int64_t)k * ep + e;
        if (HB_ZI_TYPE(d.zinfo[ik]) != HB_SEALED) lz += d.kz[((int64_t)k * KZ_COUNT + KZ_W_LZ) * ep + e] * d.scr_a[ik];
      }
      // ref: hland_model.py:3724-3737 Calc_EL_LZ_AETModel_V1
      if (einfo & HB_EI_LAKE)
        for (int k = 0; k < nz; ++k) {
          const int64_t ik = (int64_t)k * ep + e;
          if (HB_ZI_TYPE(d.zinfo[ik]) == HB_ILAKE) lz -= d.kz[((int64_t)k * KZ_COUNT + KZ_W_LZ) * ep + e] * d.scr_b[ik];
        }
      // ref: hland_model.py:3832-3840 Calc_Q1_LZ_V1
      double q1 = 0.0;
      if (lz > 0.0) q1 = k4 * hb_pow(lz, gamma1, T);
      lz -= q1;
      // ref: hland_model.py:4168-4171 Calc_RT_V2, 4197-4200 Calc_QT_V1
      const double rt = d.ke[KE_RELLANDAREA * ep + e] * outrc + rellower * q1;
      const double qt = qfactor * rt;
      d.qt[t * ep + e] = qt;
      SERE(HB_S_Q1, q1); SERE(HB_S_INRC, inrc_c); SERE(HB_S_OUTRC, outrc); SERE(HB_S_RT, rt); SERE(HB_S_QT, qt);
      SERE(HB_S_LZ, lz);
      continue;
    }
    if (einfo & HB_EI_RESPAREA) contriarea = hb_pow(contriarea, beta_last, T);
    SERE(HB_S_INUZ, inuz); SERE(HB_S_CONTRIAREA, contriarea);

    // ---- element level ------------------------------------------------------------------------------------------
    // ref: hland_model.py:2456-2484 Calc_Q0_Perc_UZ_V1
    double perc_sum = 0.0, q0_sum = 0.0;
    {
      const double uz_old = uz;
      const double dt_inuz = dt * inuz;
      const double perc_pot = dt_percmax * contriarea;
      // uz/contriarea is evaluated as uz*(1/contriarea) with the reciprocal taken once per step (≤ 1 ulp from the
      // reference's division, exact for contriarea == 1); alpha == 1 (the Lahn default) squares instead of pow()
      const double inv_ca = 1.0 / contriarea;
      const bool square = (alpha1 == 2.0);
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
        }
      }
      const double error = uz - (uz_old + inuz - perc_sum - q0_sum);
      if (error > 0.0) {
        const double factor = 1.0 - error / (perc_sum + q0_sum);
        perc_sum *= factor;
        q0_sum *= factor;
      }
    }
    // ref: hland_model.py:3113-3124 Calc_LZ_V1
    if (rellower > 0.0) {
      lz += up_low * perc_sum;
      if (einfo & HB_EI_LAKE)
        for (int k = 0; k < nz; ++k) {
          const int64_t ik = (int64_t)k * ep + e;
          if (HB_ZI_TYPE(d.zinfo[ik]) == HB_ILAKE) lz += d.kz[((int64_t)k * KZ_COUNT + KZ_W_LZ) * ep + e] * d.scr_a[ik];
        }
    } else lz = 0.0;
    // ref: hland_model.py:3724-3737 Calc_EL_LZ_AETModel_V1
    if (einfo & HB_EI_LAKE)
      for (int k = 0; k < nz; ++k) {
        const int64_t ik = (int64_t)k * ep + e;
        if (HB_ZI_TYPE(d.zinfo[ik]) == HB_ILAKE) lz -= d.kz[((int64_t)k * KZ_COUNT + KZ_W_LZ) * ep + e] * d.scr_b[ik];
      }
    // ref: hland_model.py:3832-3840 Calc_Q1_LZ_V1
    double q1 = 0.0;
    if (lz > 0.0) q1 = k4 * hb_pow(lz, gamma1, T);
    lz -= q1;
    // ref: hland_model.py:3905-3912 Calc_InRC_V1
    double inrc = relupper * q0_sum + rellower * q1;
    if (einfo & HB_EI_SEALED)
      for (int k = 0; k < nz; ++k) {
        const int64_t ik = (int64_t)k * ep + e;
        if (HB_ZI_TYPE(d.zinfo[ik]) == HB_SEALED) inrc += d.kz[((int64_t)k * KZ_COUNT + KZ_RELZONEAREA) * ep + e] * d.scr_a[ik];
      }
    // ref: hland_model.py:4109-4116 Calc_OutRC_V1; rconc_model.py:89-95 Determine_Outflow_V1
    const double outrc = rconc_outflow(d.uh, d.quh, d.ke, ep, e, nu, nash_steps, inrc);
    if (SERIES) record_sc(d.series[HB_S_SC], t, d.n_uh, d.uh0[e], d.quh, ep, e, nu, nash_steps);
    // ref: hland_model.py:4139-4141 Calc_RT_V1, 4197-4200 Calc_QT_V1
    const double qt = qfactor * outrc;
    d.qt[t * ep + e] = qt;
    SERE(HB_S_PERC, perc_sum); SERE(HB_S_Q0, q0_sum); SERE(HB_S_Q1, q1); SERE(HB_S_INRC, inrc);
    SERE(HB_S_OUTRC, outrc); SERE(HB_S_RT, outrc); SERE(HB_S_QT, qt); SERE(HB_S_UZ, uz); SERE(HB_S_LZ, lz);
  }
  if (d.flag_out) d.flag_out[e] = lost ? 2 : 0;
  d.uz[e] = uz;
  d.lz[e] = lz;
  if (MODEL == HB_MODEL_HLAND_96P) { d.ex1[e] = sg2; d.ex2[e] = sg3; }
#undef SERZ
#undef SERS
#undef SERE
}

}  // namespace hb
