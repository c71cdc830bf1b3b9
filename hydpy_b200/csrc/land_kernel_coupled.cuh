// land_kernel_coupled.cuh — runoff generation for elements with capillary flux (CFlux > 0): ONE fused kernel.
//
// With CFlux > 0 the zones of step t read the element's UZ of step t-1 (Calc_CF_SM_V1, ref:
// hydpy/models/hland/hland_model.py:1904-1919), so the zone and element equations cannot be separated in time like in
// land_kernel_v2.cuh.  They are separated in SPACE instead: a block owns 32 elements (lane = element); warp k runs
// zone k of those elements, one more warp runs the element equations (ordered zone sums, RecStep loop, LZ, unit
// hydrograph).  Per step: zone warps → __syncthreads → element warp (publishes UZ through shared memory) →
// __syncthreads.  Zone constants are plain locals under a register cap: what does not fit is spilled to thread-local
// memory, i.e. it is re-read from L1 every step (≤ 2 blocks per SM keep that working set inside the 256 KB L1) instead
// of from HBM as in the one-thread-per-element general kernel.  Same formulas, operand order and reciprocal
// multiplications as land_kernel_v2.cuh; covers one snow class, SMax = inf, at most HB_COUPLED_ZMAX zones per element.
#pragma once
#include <cstdint>

#include "land_kernel_v2.cuh"

namespace hb {

#define HB_COUPLED_ZMAX 25   // shared memory: zw * 32 * (2*8 + 4 + 32*8) + 256 bytes (+ 2.5 KB of tables) <= 227 KB
#define HB_COUPLED_SMEM(zw) ((size_t)(zw) * 32 * (2 * 8 + 4 + 32 * 8) + 256)

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) coupled_kernel(const LandV2Dev d) {
  __shared__ MathTables s_tables;
  extern __shared__ double c_smem[];   // term_a[zw][32], term_b[zw][32], uz[32], int ztype[zw][32], constants[C_COUNT][zw*32]
  stage_tables(&s_tables, d.tables);
  const TabRef T = hb_tabref(&s_tables);
  const int zw = (int)(blockDim.x >> 5) - 1;            // zone warps
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* s_a = c_smem;
  double* s_b = s_a + zw * 32;
  double* s_uz = s_b + zw * 32;
  int* s_zt = reinterpret_cast<int*>(s_uz + 32);
  double* s_c = s_uz + 32 + zw * 16;                    // constants behind the zone-type table (zw*32 ints)
  const int64_t ep = d.e_pad;
  const int64_t li = (int64_t)blockIdx.x * 32 + lane;
  const bool valid = li < d.n_clist;
  const int64_t e = valid ? (int64_t)d.clist[li] : 0;
  const int einfo = valid ? d.einfo[e] : 0;
  const int nz = valid ? d.nz[e] : 0;
  const bool zone_role = w < zw;
  const int k = zone_role ? w : 0;
  const bool active = zone_role && valid && k < nz;
  const int64_t slot = (int64_t)k * ep + e;
  const int nt = (int)d.nt;

  // ---- zone role: constants and states ------------------------------------------------------------------------------
  const double* kzk = d.kz + (int64_t)k * KZ_COUNT * ep + e;
#define KZ(name) (active ? kzk[(int64_t)(name) * ep] : 0.0)
  const int zi = active ? d.zinfo[slot] : 0;
  const int zt = HB_ZI_TYPE(zi), zf = HB_ZI_FLAGS(zi);
  const bool soil = (zt == HB_FIELD) || (zt == HB_FOREST);
  const bool icept = soil || (zt == HB_SEALED);
  const bool resp = (einfo & HB_EI_RESPAREA) != 0;
  // zone constants live in shared memory ([constant][zone thread], conflict-free): registers stay free for the step and
  // two blocks per SM (the element warps' RecStep chains are what has to overlap) fit without spilling to L1/L2
  enum { C_TCORR, C_TCALT_DZ, C_TT_HI, C_TT_LO, C_INV_TTINT, C_PCALT_F, C_PCORR, C_INV_RFCF, C_INV_SFCF, C_ICMAX, C_CFMAX, C_CFVAR, C_TTM, C_CFR_CFMAX, C_WHC, C_FC, C_INV_FC, C_BETA, C_THRESH, C_INV_THRESH, C_EXCESSRED, C_ATF, C_ETF, C_ALT_F, C_NEG_PF, C_CFLUX, C_W_A, C_W_B, C_G_MELT, C_G_VAR, C_TTI, C_SF, C_COUNT };
  const int nzt = zw * 32;
#define CS(i) s_c[(i) * nzt + k * 32 + lane]
  if (zone_role) {
    CS(C_TCORR) = KZ(KZ_TCORR);
    CS(C_TCALT_DZ) = KZ(KZ_TCALT_DZ);
    CS(C_TT_HI) = KZ(KZ_TT_HI);
    CS(C_TT_LO) = KZ(KZ_TT_LO);
    CS(C_INV_TTINT) = KZ(KZ_INV_TTINT);
    CS(C_PCALT_F) = KZ(KZ_PCALT_F);
    CS(C_PCORR) = KZ(KZ_PCORR);
    CS(C_INV_RFCF) = KZ(KZ_INV_RFCF);
    CS(C_INV_SFCF) = KZ(KZ_INV_SFCF);
    CS(C_ICMAX) = KZ(KZ_ICMAX);
    CS(C_CFMAX) = KZ(KZ_CFMAX);
    CS(C_CFVAR) = KZ(KZ_CFVAR);
    CS(C_TTM) = KZ(KZ_TTM);
    CS(C_CFR_CFMAX) = KZ(KZ_CFR_CFMAX);
    CS(C_WHC) = KZ(KZ_WHC);
    CS(C_FC) = KZ(KZ_FC);
    CS(C_INV_FC) = KZ(KZ_INV_FC);
    CS(C_BETA) = KZ(KZ_BETA);
    CS(C_THRESH) = KZ(KZ_THRESH);
    CS(C_INV_THRESH) = KZ(KZ_INV_THRESH);
    CS(C_EXCESSRED) = KZ(KZ_EXCESSRED);
    CS(C_ATF) = KZ(KZ_ATF);
    CS(C_ETF) = KZ(KZ_ETF);
    CS(C_ALT_F) = KZ(KZ_ALT_F);
    CS(C_NEG_PF) = KZ(KZ_NEG_PF);
    CS(C_CFLUX) = KZ(KZ_CFLUX);
    CS(C_W_A) = soil || zt == HB_GLACIER ? KZ(KZ_W_UZ) : zt == HB_ILAKE ? KZ(KZ_W_LZ) : KZ(KZ_RELZONEAREA);
    CS(C_W_B) = soil ? KZ(KZ_W_CONTRI) : KZ(KZ_W_LZ);
    CS(C_G_MELT) = zt == HB_GLACIER ? KZ(KZ_GMELT) : 0.0;
    CS(C_G_VAR) = zt == HB_GLACIER ? KZ(KZ_GVAR) : 0.0;
    CS(C_TTI) = zt == HB_ILAKE ? KZ(KZ_TTI) : 0.0;
    CS(C_SF) = active ? d.sfdist[e] : 0.0;
  }
#undef KZ
  double ic = 0.0, sm = 0.0, sp = 0.0, wc = 0.0;
  if (active) { ic = d.ic[slot]; sm = d.sm[slot]; sp = d.sp[slot]; wc = d.wc[slot]; }
  if (zone_role) s_zt[k * 32 + lane] = active ? zt : 0;

  // ---- element role: constants and states -----------------------------------------------------------------------------
  const bool elem_role = !zone_role && valid;
#define KE(name) (elem_role ? d.ke[(int64_t)(name) * ep + e] : 0.0)
  const double dt = KE(KE_DT), dt_percmax = KE(KE_DT_PERCMAX), dt_k = KE(KE_DT_K), alpha1 = KE(KE_ALPHA1), k4 = KE(KE_K4),
               gamma1 = KE(KE_GAMMA1), relupper = KE(KE_RELUPPER), rellower = KE(KE_RELLOWER), up_low = KE(KE_UP_LOW),
               qfactor = KE(KE_QFACTOR), beta_last = KE(KE_BETA_LAST);
#undef KE
  const int nu = elem_role ? d.nu[e] : 0, recstep = elem_role ? d.recstep[e] : 0;
  const bool square = (alpha1 == 2.0);
  double uz = elem_role ? d.uz[e] : 0.0, lz = elem_role ? d.lz[e] : 0.0;
  if (!zone_role) s_uz[lane] = uz;

  // ---- forcing (zone role), prefetched one step ahead -----------------------------------------------------------------
  const int64_t col = active ? d.fcol[e] : 0;
  const double* fP = d.forcing + (0 * d.nt_total + d.t_first) * d.n_col + col;
  const double* fT = d.forcing + (1 * d.nt_total + d.t_first) * d.n_col + col;
  const double* fA = d.forcing + (2 * d.nt_total + d.t_first) * d.n_col + col;
  const double* fE = d.forcing + (3 * d.nt_total + d.t_first) * d.n_col + col;
  const double* sinp = d.sinfactor + d.t_first;
  const int64_t fstride = d.n_col;
  double n_p = 0.0, n_t = 0.0, n_nat = 0.0, n_net = 0.0, n_sin = 0.0;
  if (active && nt > 0) { n_p = ldg(fP); n_t = ldg(fT); n_nat = ldg(fA); n_net = ldg(fE); n_sin = ldg(sinp); }
  __syncthreads();

  // Both roles run their own loop with the same sequence of block barriers (bar.sync on a named barrier, every warp takes
  // exactly one of the two paths); separate loops let the compiler allocate registers per role instead of for the union.
#define HB_BLOCK_BAR() asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory")
  if (!zone_role) {
    for (int t = 0; t < nt; ++t) {
      HB_BLOCK_BAR();    // the terms of this step are complete
      if (elem_role) {
        double inuz = 0.0, contriarea = 1.0;
        for (int kk = 0; kk < nz; ++kk) {
          const int ztk = s_zt[kk * 32 + lane];
          if (ztk == HB_FIELD || ztk == HB_FOREST) { inuz += s_a[kk * 32 + lane]; contriarea *= s_b[kk * 32 + lane]; }
          else if (ztk == HB_GLACIER) inuz += s_a[kk * 32 + lane];
        }
        if (resp) contriarea = hb_pow(contriarea, beta_last, T);
        // ref: hland_model.py:2456-2484 Calc_Q0_Perc_UZ_V1
        double perc_sum = 0.0, q0_sum = 0.0;
        {
          const double uz_old = uz;
          const double dt_inuz = dt * inuz;
          const double perc_pot = dt_percmax * contriarea;
          const double inv_ca = 1.0 / contriarea;
          if (__all_sync(__activemask(), contriarea > 0.0 && !square)) {   // unswitched, see element_kernel
            for (int i = 0; i < recstep; ++i) {
              const double a = uz + dt_inuz;
              uz = HB_PYMAX(a, 0.0);
              const double perc = HB_PYMIN(perc_pot, uz);
              uz -= perc;
              perc_sum += perc;
              if (uz > 0.0) {
                const double q = dt_k * hb_pow(uz * inv_ca, alpha1, T);
                const double q0 = HB_PYMIN(q, uz);
                uz -= q0;
                q0_sum += q0;
              }
            }
          } else {
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
          }
          const double error = uz - (uz_old + inuz - perc_sum - q0_sum);
          if (error > 0.0) {
            const double factor = 1.0 - error / (perc_sum + q0_sum);
            perc_sum *= factor;
            q0_sum *= factor;
          }
        }
        s_uz[lane] = uz;
        // ref: hland_model.py:3113-3124 Calc_LZ_V1, 3724-3737 Calc_EL_LZ_AETModel_V1
        if (rellower > 0.0) {
          lz += up_low * perc_sum;
          if (einfo & HB_EI_LAKE)
            for (int kk = 0; kk < nz; ++kk)
              if (s_zt[kk * 32 + lane] == HB_ILAKE) lz += s_a[kk * 32 + lane];
        } else lz = 0.0;
        if (einfo & HB_EI_LAKE)
          for (int kk = 0; kk < nz; ++kk)
            if (s_zt[kk * 32 + lane] == HB_ILAKE) lz -= s_b[kk * 32 + lane];
        // ref: hland_model.py:3832-3840 Calc_Q1_LZ_V1
        double q1 = 0.0;
        if (lz > 0.0) q1 = k4 * hb_pow(lz, gamma1, T);
        lz -= q1;
        // ref: hland_model.py:3905-3912 Calc_InRC_V1
        double inrc = relupper * q0_sum + rellower * q1;
        if (einfo & HB_EI_SEALED)
          for (int kk = 0; kk < nz; ++kk)
            if (s_zt[kk * 32 + lane] == HB_SEALED) inrc += s_a[kk * 32 + lane];
        // ref: hland_model.py:4109-4116 Calc_OutRC_V1; rconc_model.py:89-95 Determine_Outflow_V1
        double outrc = inrc;
        if (nu > 0) {
          outrc = ldg(d.uh + e) * inrc + d.quh[e];
          for (int j = 1; j < nu; ++j)
            d.quh[(int64_t)(j - 1) * ep + e] = ldg(d.uh + (int64_t)j * ep + e) * inrc + d.quh[(int64_t)j * ep + e];
        }
        // ref: hland_model.py:4139-4141 Calc_RT_V1, 4197-4200 Calc_QT_V1
        d.qt[(d.t_out + t) * ep + e] = qfactor * outrc;
      }
      HB_BLOCK_BAR();    // UZ of this step is published
    }
    if (elem_role) { d.uz[e] = uz; d.lz[e] = lz; }
    return;
  }

#define tcorr CS(C_TCORR)
#define tcalt_dz CS(C_TCALT_DZ)
#define tt_hi CS(C_TT_HI)
#define tt_lo CS(C_TT_LO)
#define inv_ttint CS(C_INV_TTINT)
#define pcalt_f CS(C_PCALT_F)
#define pcorr CS(C_PCORR)
#define inv_rfcf CS(C_INV_RFCF)
#define inv_sfcf CS(C_INV_SFCF)
#define icmax CS(C_ICMAX)
#define cfmax CS(C_CFMAX)
#define cfvar CS(C_CFVAR)
#define ttm CS(C_TTM)
#define cfr_cfmax CS(C_CFR_CFMAX)
#define whc CS(C_WHC)
#define fc CS(C_FC)
#define inv_fc CS(C_INV_FC)
#define beta CS(C_BETA)
#define thresh CS(C_THRESH)
#define inv_thresh CS(C_INV_THRESH)
#define excessred CS(C_EXCESSRED)
#define atf CS(C_ATF)
#define etf CS(C_ETF)
#define alt_f CS(C_ALT_F)
#define neg_pf CS(C_NEG_PF)
#define cflux CS(C_CFLUX)
#define w_a CS(C_W_A)
#define w_b CS(C_W_B)
#define g_melt CS(C_G_MELT)
#define g_var CS(C_G_VAR)
#define tti CS(C_TTI)
#define sf CS(C_SF)
  // A zone step is cut where it first reads UZ: `pre` (meteorology, snow, evaporation demand, R) of step t+1 runs while the
  // element warp is busy with the RecStep chain of step t; `post` (capillary flux, soil evaporation, contributing area)
  // follows once that UZ is published.
  double c_tc = 0.0, c_pc = 0.0, c_in = 0.0, c_ncover = 0.0, c_pet = 0.0, c_aet = 0.0, c_r = 0.0;
  auto pre = [&](int s_) {
      const double in_p = n_p, in_t = n_t, in_nat = n_nat, in_net = n_net, sinf = n_sin;
      if (s_ + 1 < nt) {
        fP += fstride; fT += fstride; fA += fstride; fE += fstride; ++sinp;
        n_p = ldg(fP); n_t = ldg(fT); n_nat = ldg(fA); n_net = ldg(fE); n_sin = ldg(sinp);
      }
      // ref: hland_model.py:71-77 Calc_TC_V1, 125-135 Calc_FracRain_V1, 213-225 Calc_PC_V1
      const double tc = in_t + tcorr - tcalt_dz;
      double fracrain;
      if (tc >= tt_hi) fracrain = 1.0;
      else if (tc <= tt_lo) fracrain = 0.0;
      else fracrain = (tc - tt_lo) * inv_ttint;
      double pc = in_p * pcalt_f;
      if (pc <= 0.0) pc = 0.0;
      else pc *= pcorr / (fracrain * inv_rfcf + (1.0 - fracrain) * inv_sfcf);
      // ref: hland_model.py:299-309 Calc_TF_Ic_V1
      double tf;
      if (icept) {
        const double a = pc - (icmax - ic);
        tf = HB_PYMAX(a, 0.0);
        ic += pc - tf;
      } else { tf = pc; ic = 0.0; }
      double in_ = tf, ncover = 0.0;
      bool bare = true;
      double s_cfact = 0.0, s_gact = 0.0, s_melt = 0.0, s_refr = 0.0, s_glmelt = 0.0;   // recorded only
      if (zt != HB_ILAKE) {
        // ref: hland_model.py:484-499 Calc_SP_WC_V1
        wc += sf * (tf * fracrain);
        sp += sf * (tf * (1.0 - fracrain));
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
        // ref: hland_model.py:1607-1619 Calc_GAct_V1, 1688-1701 Calc_GlMelt_In_V1
        if (zt == HB_GLACIER) {
          const double ga = g_melt + sinf * g_var;
          const double gact = HB_PYMAX(ga, 0.0);
          s_gact = gact;
          if (tc > ttm && bare) {
            s_glmelt = gact * (tc - ttm);
            in_ += s_glmelt;
          }
        }
      } else { sp = 0.0; wc = 0.0; }
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
      }
      // ref: evap_model.py:6068-6086 Calc_InterceptionEvaporation_V1; hland_model.py:383-396 Calc_EI_Ic_AETModel_V1
      double aet_ie = 0.0;
      if (zf & HB_ZF_INTERCEPTION) aet_ie = HB_PYMIN(pet, ic);
      double s_ei = 0.0;
      if (icept) { s_ei = HB_PYMIN(aet_ie, ic); ic -= s_ei; }
      else ic = 0.0;
      // ref: hland_model.py:1776-1790 Calc_R_SM_V1 (needs the soil moisture the previous step left behind — and that step's
      // zone part is complete when this runs)
      double r_ = in_;
      if (soil && fc > 0.0) {
        r_ = in_ * hb_pow(sm * inv_fc, beta, T);
        const double a = sm + in_ - fc;
        r_ = HB_PYMAX(r_, a);
      }
      c_tc = tc; c_pc = pc; c_in = in_; c_ncover = ncover; c_pet = pet; c_aet = aet_ie; c_r = r_;
  };
  if (active && nt > 0) pre(0);
  for (int t = 0; t < nt; ++t) {
    if (active) {
      const double tc = c_tc, pc = c_pc, in_ = c_in, ncover = c_ncover, pet = c_pet, aet_ie = c_aet;
      const double uz_prev = s_uz[lane];
      (void)tc; (void)pc;
      double ta, tb = 1.0;
      if (soil) {
        double r = c_r;
        sm += in_ - r;
        // ref: hland_model.py:1904-1919 Calc_CF_SM_V1 (uz_prev = UZ at the end of the previous step)
        double cf = 0.0;
        if (fc > 0.0) {
          cf = cflux * (1.0 - sm * inv_fc);
          const double a1 = uz_prev + r;
          cf = HB_PYMIN(cf, a1);
          const double b = fc - sm;
          cf = HB_PYMIN(cf, b);
        }
        sm += cf;
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
        // ref: hland_model.py:2002-2018 Calc_EA_SM_AETModel_V1
        const double ea = HB_PYMIN(set, sm);
        sm -= ea;
        if (sm > fc) { r += sm - fc; sm = fc; }
        // ref: hland_model.py:2092-2101 Calc_InUZ_V1 (term), 2224-2236 Calc_ContriArea_V1 (factor)
        ta = w_a * (r - cf);
        if (resp && fc > 0.0) tb = hb_pow(sm * inv_fc, w_b, T);
      } else {
        sm = 0.0;
        double el = 0.0;
        if (zt == HB_ILAKE) {
          // ref: hland_model.py:3113-3124 Calc_LZ_V1 (term), 3724-3737 Calc_EL_LZ_AETModel_V1 (term);
          // evap_model.py:5673-5681 Calc_WaterEvaporation_V1
          ta = w_a * pc;
          el = ((zf & HB_ZF_WATER) && tc > tti) ? pet : 0.0;
          tb = w_b * el;
        } else ta = w_a * (in_ - 0.0);  // GLACIER: InUZ term; SEALED: InRC term (r = in_, ref: hland_model.py:1528-1535)
      }
      s_a[k * 32 + lane] = ta;
      s_b[k * 32 + lane] = tb;
    }
    HB_BLOCK_BAR();    // the terms of this step are complete
    if (active && t + 1 < nt) pre(t + 1);
    HB_BLOCK_BAR();    // UZ of this step is published, the term buffers are free again
  }
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
#undef cflux
#undef w_a
#undef w_b
#undef g_melt
#undef g_var
#undef tti
#undef sf
#undef CS
  if (active) { d.ic[slot] = ic; d.sm[slot] = sm; d.sp[slot] = sp; d.wc[slot] = wc; }
#undef HB_BLOCK_BAR
}

}  // namespace hb
