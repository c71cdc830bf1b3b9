// fastmath.cuh — table-driven fp64 log / exp / pow for the runoff-generation kernels.
//
// The reference calls libm `pow` once per UZ sub-step (hydpy/models/hland/hland_model.py:2474), per soil zone
// (:1783) and per contributing-area factor (:2233-2236) and `exp` once per zone (evap_model.py:5324) — ≈76 `pow` per
// element-step for the hourly Lahn parameterisation.  libdevice's correctly-special-cased `pow` costs ≈250 SASS
// instructions (ncu: 53 % of all instructions of the first kernel version, profiles/r01_land_v1_details.csv); the
// routines below need ≈24 fp64 instructions + a few integer ones and two shared-memory look-ups:
//
//   log(x)  = e*ln2 + log(c_k) + log1p(r),   c_k = 1 + k/128 the interval centre NEAREST to the mantissa m (k = 0..128),
//             r = m*(1/c_k) - 1, |r| <= 2^-8; m = 1 hits c_0 = 1 exactly, so log(1) == 0 and pow(1, y) == 1 exactly;
//             degree-7 polynomial in Horner form (one literal per FMA, see below);
//   exp(t)  = 2^q * 2^(j/64) * (1 + p(r)),   n = rint(t*64/ln2) = 64*q + j,  r = t - n*ln2/64,  degree-5 polynomial;
//   pow(x,y)= exp(y*log(x))  for normal positive x and |y*log(x)| < 690; everything else (zero, negative, subnormal,
//             inf, nan, overflow, underflow) goes to libdevice `pow` — out of line, so that all IEEE special cases keep
//             the reference's libm semantics without costing registers or instructions on the hot path.
//
// Instruction economy (the zone kernel is issue-bound, profiles/r01_land_v2_zone_element_details.csv): every
// polynomial step is fma(p, r, LITERAL) with ONE literal, taken from constant memory as a direct DFMA operand; a
// two-literal Estrin step costs two extra register moves on sm_100, and full-width literals cost two UMOVs each.
//
// Accuracy (tests/test_fastmath.py compiles THIS file for the host and compares with x87 extended precision; the GPU
// parity tests bound the end effect): log: 4e-16 + 1 ulp absolute, exp: < 2 ulp, pow: relative error
// < 5e-16 + 2*|y*log x|*2^-52 (the exponent is carried in one double), i.e. < 1e-14 for everything the model feeds
// it — five orders of magnitude inside the 1e-9 parity bar.  All fused multiply-adds are explicit `fma()` calls; the
// translation unit is compiled with -fmad=false.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define HB_HD __host__ __device__ __forceinline__
#else
#define HB_HD inline
#endif

namespace hb {

// bit access that works on the device (intrinsics) and on the host (tests/fastmath_host.cpp checks the very same code)
HB_HD int hb_hi(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  int64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
HB_HD int hb_lo(double x) {
#ifdef __CUDA_ARCH__
  return __double2loint(x);
#else
  int64_t b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffff);
#endif
}
HB_HD double hb_make(int hi, int lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, lo);
#else
  const int64_t b = ((int64_t)hi << 32) | (int64_t)(uint32_t)lo; double x; std::memcpy(&x, &b, 8); return x;
#endif
}

#define HB_LOG_TABLE_BITS 7
#define HB_LOG_N ((1 << HB_LOG_TABLE_BITS) + 1)   // centres 1 + k/128, k = 0..128
#define HB_EXP_TABLE_BITS 6
#define HB_EXP_N (1 << HB_EXP_TABLE_BITS)

struct MathTables {   // (hb_ld_log / hb_ld_exp rely on this layout)
  double log_tab[HB_LOG_N][2];   // {1/c_k rounded to double, -log of that ROUNDED reciprocal}: one 16-byte look-up
  double exp2[HB_EXP_N];         // 2^(j/64)
};

// host side: tables in long double precision (x87 80-bit is plenty for 1e-19)
inline void make_math_tables(MathTables& t) {
  for (int k = 0; k < HB_LOG_N; ++k) {
    const long double c = 1.0L + (long double)k / (HB_LOG_N - 1);
    t.log_tab[k][0] = (double)(1.0L / c);
    t.log_tab[k][1] = (k == 0) ? 0.0 : (double)(-logl((long double)t.log_tab[k][0]));
  }
  for (int j = 0; j < HB_EXP_N; ++j) t.exp2[j] = (double)exp2l((long double)j / HB_EXP_N);
}

// Handle of the tables inside a kernel.  On the device it is the 32-bit shared-memory address of the block's copy: the
// look-ups become `ld.shared` with a register base — through a generic pointer the compiler re-derives the shared window
// (S2UR + ULEA) at every look-up, three instructions on the dependent chain of the RecStep loop.
struct TabRef {
#ifdef __CUDA_ARCH__
  uint32_t base;
#else
  const MathTables* p;
#endif
};
HB_HD TabRef hb_tabref(const MathTables* T) {
  TabRef t;
#ifdef __CUDA_ARCH__
  t.base = (uint32_t)__cvta_generic_to_shared(T);
#else
  t.p = T;
#endif
  return t;
}
HB_HD void hb_ld_log(TabRef t, int k, double& invc, double& logc) {
#ifdef __CUDA_ARCH__
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(invc), "=d"(logc) : "r"(t.base + (uint32_t)k * 16u));
#else
  invc = t.p->log_tab[k][0]; logc = t.p->log_tab[k][1];
#endif
}
HB_HD double hb_ld_exp(TabRef t, int j) {
#ifdef __CUDA_ARCH__
  double s;
  asm("ld.shared.f64 %0, [%1];" : "=d"(s) : "r"(t.base + (uint32_t)(HB_LOG_N * 16) + (uint32_t)j * 8u));
  return s;
#else
  return t.p->exp2[j];
#endif
}

// literals (see "instruction economy" above)
enum {
  HB_K_L7 = 0, HB_K_L6, HB_K_L5, HB_K_L4, HB_K_L3, HB_K_LN2HI, HB_K_LN2LO,
  HB_K_INV, HB_K_L64HI, HB_K_L64LO, HB_K_E5, HB_K_E4, HB_K_E3, HB_K_COUNT
};
#define HB_MATH_LITERALS                                                                                              \
  {1.0 / 7.0, -1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0,                                                          \
   6.93147180369123816490e-01 /* 0x1.62e42fee00000p-1: e*ln2hi is exact */, 1.90821492927058770002e-10,             \
   92.33248261689366 /* 64/ln2 */, 0.01083042469326756 /* 0x1.62e42fee00000p-7: n*l64hi exact for |n| < 2^21 */,      \
   2.9815858269852933e-12 /* ln2/64 - l64hi */, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0}
#ifdef __CUDACC__
__constant__ double hb_math_k_dev[HB_K_COUNT] = HB_MATH_LITERALS;
#endif
static const double hb_math_k_host[HB_K_COUNT] = HB_MATH_LITERALS;
#ifdef __CUDA_ARCH__
#define HB_LIT(i) hb_math_k_dev[i]
#else
#define HB_LIT(i) hb_math_k_host[i]
#endif

HB_HD double hb_log_normal(double x, TabRef T) {
  // x must be a positive normal number (for anything else the result is garbage, but the evaluation is harmless)
  const int hi = hb_hi(x), lo = hb_lo(x);
  const int e = (hi >> 20) - 1023;
  const int k = ((hi & 0x000fffff) + (1 << (19 - HB_LOG_TABLE_BITS))) >> (20 - HB_LOG_TABLE_BITS);   // nearest centre
  const double m = hb_make((hi & 0x000fffff) | 0x3ff00000, lo);  // [1, 2)
  double invc, logc;
  hb_ld_log(T, k, invc, logc);
  const double r = fma(m, invc, -1.0);
  const double r2 = r * r;
  // log1p(r) = r + r^2*(-1/2 + r/3 - r^2/4 + r^3/5 - r^4/6 + r^5/7)
  double q = fma(r, HB_LIT(HB_K_L7), HB_LIT(HB_K_L6));
  q = fma(q, r, HB_LIT(HB_K_L5));
  q = fma(q, r, HB_LIT(HB_K_L4));
  q = fma(q, r, HB_LIT(HB_K_L3));
  q = fma(q, r, -0.5);
  const double tail = fma(r2, q, r);
  const double de = (double)e;
  return fma(de, HB_LIT(HB_K_LN2HI), logc) + fma(de, HB_LIT(HB_K_LN2LO), tail);
}

HB_HD double hb_exp_bounded(double t, TabRef T) {
  // |t| < 690: the result is a normal number (for anything else the result is garbage, but the evaluation is harmless)
  const double shift = 6755399441055744.0;          // 1.5 * 2^52
  const double kd = fma(t, HB_LIT(HB_K_INV), shift);
  const int ki = hb_lo(kd);
  const double n = kd - shift;
  double r = fma(n, -HB_LIT(HB_K_L64HI), t);
  r = fma(n, -HB_LIT(HB_K_L64LO), r);
  const double s = hb_ld_exp(T, ki & (HB_EXP_N - 1));
  const int q = ki >> HB_EXP_TABLE_BITS;
  const double r2 = r * r;
  double p = fma(r, HB_LIT(HB_K_E5), HB_LIT(HB_K_E4));
  p = fma(p, r, HB_LIT(HB_K_E3));
  p = fma(p, r, 0.5);
  p = fma(r2, p, r);                                 // e^r - 1
  const double v = fma(s, p, s);
  return hb_make(hb_hi(v) + (q << 20), hb_lo(v));
}

// rare paths, deliberately out of line on the device
#ifdef __CUDACC__
__device__ __noinline__ double hb_pow_slow(double x, double y) { return pow(x, y); }
__device__ __noinline__ double hb_exp_slow(double t) { return exp(t); }
#endif

HB_HD double hb_exp(double t, TabRef T) {
  const double r = hb_exp_bounded(t, T);
  if (fabs(t) < 690.0) return r;
#ifdef __CUDA_ARCH__
  return hb_exp_slow(t);
#else
  return exp(t);
#endif
}

#ifdef __CUDACC__
__device__ __noinline__ double hb_log_slow(double x) { return log(x); }
#endif

// log(x): table path for positive normal x, libm/libdevice for zero, subnormal, negative, inf and nan
HB_HD double hb_log(double x, TabRef T) {
  const bool ok_x = (unsigned)(hb_hi(x) - 0x00100000) < 0x7fe00000u;
  const double r = hb_log_normal(x, T);
  if (ok_x) return r;
#ifdef __CUDA_ARCH__
  return hb_log_slow(x);
#else
  return log(x);
#endif
}

HB_HD double hb_pow(double x, double y, TabRef T) {
  // positive normal x: 0x00100000 <= hi < 0x7ff00000
  const bool ok_x = (unsigned)(hb_hi(x) - 0x00100000) < 0x7fe00000u;
  const double t = y * hb_log_normal(x, T);
  const double r = hb_exp_bounded(t, T);
  if (ok_x && fabs(t) < 690.0) return r;
#ifdef __CUDA_ARCH__
  return hb_pow_slow(x, y);
#else
  return pow(x, y);
#endif
}

// hb_pow(x, y) for a caller that already holds lgn = hb_log_normal(x): the same bits as hb_pow, without the second log
HB_HD double hb_pow_from_log(double x, double lgn, double y, TabRef T) {
  const bool ok_x = (unsigned)(hb_hi(x) - 0x00100000) < 0x7fe00000u;
  const double t = y * lgn;
  const double r = hb_exp_bounded(t, T);
  if (ok_x && fabs(t) < 690.0) return r;
#ifdef __CUDA_ARCH__
  return hb_pow_slow(x, y);
#else
  return pow(x, y);
#endif
}

// hb_log(x) from lgn = hb_log_normal(x)
HB_HD double hb_log_from_normal(double x, double lgn) {
  const bool ok_x = (unsigned)(hb_hi(x) - 0x00100000) < 0x7fe00000u;
  if (ok_x) return lgn;
#ifdef __CUDA_ARCH__
  return hb_log_slow(x);
#else
  return log(x);
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// RecStep chain (Calc_Q0_Perc_UZ_V1, hydpy/models/hland/hland_model.py:2456-2484): q = dt*k * (uz/contriarea)^(1+alpha)
// once per sub-step, every instruction on ONE dependent chain per element (the sub-steps of an element are sequential
// and all 10^5 chains are resident at once, so the kernel's time is instructions-on-the-chain x their latency).  The
// routine below is built for the shortest chain, not for generality:
//
//   q = 2^(y*log2(d) + c2),   c2 = log2(dt*k) - y*log2(contriarea)  (step-invariant, off the chain)
//   log2(d) = e + log2(c_k) + log1p(r)/ln2 with 512 interval centres (|r| <= 2^-10: a degree-4 polynomial reaches
//             2e-16, three dependent FMAs with 1/ln2 folded into the coefficients);
//   2^t     = 2^q * 2^(j/256) * (1 + p(r)),  n = rint(256 t), r = t - n/256 EXACT (no two-part reduction constant),
//             degree-4 polynomial in r with the powers of ln2 folded in.
//
// 15 dependent fp64 instructions from d to q instead of 26, two shared-memory look-ups (10 KB of tables per block).
// Accuracy: relative error of q < 2.3e-16 * (4 + 1.5 T), T = the magnitudes of the terms of log2 q (one double), i.e.
// < 2e-14 on everything the model feeds it — tests/test_fastmath.py checks THIS code on the host against x87 extended
// precision.  Valid for positive normal d and |t| < 1000; the caller checks t and repeats the step with the general code
// otherwise.
#define HB_CHAIN_LOG_BITS 9
#define HB_CHAIN_LOG_N (1 << HB_CHAIN_LOG_BITS)   // centres 1 + k/512, k = 0..511 (a mantissa that rounds up to 2 is the
                                                  // centre 1 of the next binade: the rounding carries into the exponent)
#define HB_CHAIN_EXP_BITS 8
#define HB_CHAIN_EXP_N (1 << HB_CHAIN_EXP_BITS)

struct ChainTables {
  double log_tab[HB_CHAIN_LOG_N][2];   // {1/c_k rounded to double, -log2 of that ROUNDED reciprocal}
  double exp2[HB_CHAIN_EXP_N];         // 2^(j/256)
};

inline void make_chain_tables(ChainTables& t) {
  for (int k = 0; k < HB_CHAIN_LOG_N; ++k) {
    const long double c = 1.0L + (long double)k / HB_CHAIN_LOG_N;
    t.log_tab[k][0] = (double)(1.0L / c);
    t.log_tab[k][1] = (k == 0) ? 0.0 : (double)(-log2l((long double)t.log_tab[k][0]));
  }
  for (int j = 0; j < HB_CHAIN_EXP_N; ++j) t.exp2[j] = (double)exp2l((long double)j / HB_CHAIN_EXP_N);
}

// On the device the tables are ONE block-wide shared-memory object at namespace scope (only kernels that use it pay for
// it): the look-ups are `LDS [index + window + constant]` — through a pointer the table base costs an add on the chain.
#ifdef __CUDACC__
__shared__ ChainTables hb_chain_smem;
__device__ __forceinline__ void hb_chain_stage(const ChainTables* g) {   // all threads of the block, before a barrier
  const double* src = reinterpret_cast<const double*>(g);
  double* dst = reinterpret_cast<double*>(&hb_chain_smem);
  for (int i = threadIdx.x; i < (int)(sizeof(ChainTables) / sizeof(double)); i += blockDim.x) dst[i] = src[i];
}
#endif
struct ChainRef {
#ifndef __CUDA_ARCH__
  const ChainTables* p;
#endif
};
HB_HD ChainRef hb_chainref(const ChainTables* T) {
  ChainRef t;
#ifndef __CUDA_ARCH__
  t.p = T;
#else
  (void)T;
#endif
  return t;
}
HB_HD void hb_ld_clog(ChainRef t, int byte_off, double& invc, double& l2c) {   // byte_off = 16 k
#ifdef __CUDA_ARCH__
  const double2 v = *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(hb_chain_smem.log_tab) + byte_off);
  invc = v.x; l2c = v.y;
#else
  invc = t.p->log_tab[byte_off >> 4][0]; l2c = t.p->log_tab[byte_off >> 4][1];
#endif
}
HB_HD double hb_ld_cexp(ChainRef t, int byte_off) {                             // byte_off = 8 j
#ifdef __CUDA_ARCH__
  return *reinterpret_cast<const double*>(reinterpret_cast<const char*>(hb_chain_smem.exp2) + byte_off);
#else
  return t.p->exp2[byte_off >> 3];
#endif
}

enum {
  HB_C_L4 = 0, HB_C_L3, HB_C_L2, HB_C_L1, HB_C_E4, HB_C_E3, HB_C_E2, HB_C_E1, HB_C_COUNT
};
#define HB_CHAIN_LITERALS                                                                                             \
  {-0.36067376022224085 /* -1/(4 ln2) */, 0.48089834696298783 /* 1/(3 ln2) */, -0.72134752044448170 /* -1/(2 ln2) */, \
   1.4426950408889634 /* 1/ln2 */, 9.6181291076284772e-03 /* ln2^4/24 */, 5.5504108664821580e-02 /* ln2^3/6 */,       \
   2.4022650695910071e-01 /* ln2^2/2 */, 6.9314718055994531e-01 /* ln2 */}
#ifdef __CUDACC__
__constant__ double hb_chain_k_dev[HB_C_COUNT] = HB_CHAIN_LITERALS;
#endif
static const double hb_chain_k_host[HB_C_COUNT] = HB_CHAIN_LITERALS;
#ifdef __CUDA_ARCH__
#define HB_CLIT(i) hb_chain_k_dev[i]
#else
#define HB_CLIT(i) hb_chain_k_host[i]
#endif

// 2^t with t = y*log2(d) + c2, for positive normal d and |t| < 1000; `t` is returned for the caller's range check.  Anything
// else (zero, negative, subnormal, inf, nan d; |t| >= 1000) yields garbage without side effects.  With y >= 1,
// c2 <= 200 y and a subnormal or zero d the garbage announces itself: t < -800.
HB_HD double hb_chain_pow2(double d, double y, double c2, ChainRef T, double& t) {
  const int hi = hb_hi(d), lo = hb_lo(d);
  const int hr = hi + (1 << (19 - HB_CHAIN_LOG_BITS));      // mantissa rounded to the nearest centre (may carry: see above)
  const int top = hr & (int)0xfff00000;
  const double m = hb_make(hi - top + 0x3ff00000, lo);      // d / 2^e in [1 - 2^-11, 2 - 2^-10)
  double invc, l2c;
  hb_ld_clog(T, (hr >> (16 - HB_CHAIN_LOG_BITS)) & ((HB_CHAIN_LOG_N - 1) << 4), invc, l2c);
  const double r = fma(m, invc, -1.0);
  const double r2 = r * r;
  const double rl = r * HB_CLIT(HB_C_L1);
  // log1p(r)/ln2 = r/ln2 + r^2*(-1/2 + r/3 - r^2/4)/ln2
  double q = fma(r, HB_CLIT(HB_C_L4), HB_CLIT(HB_C_L3));
  q = fma(q, r, HB_CLIT(HB_C_L2));
  const double tail = fma(r2, q, rl);
  const double de = (double)((hr >> 20) - 1023);
  const double base = fma(y, l2c, fma(y, de, c2));          // off the chain: ready before the polynomial
  t = fma(y, tail, base);
  // 2^t
  const double shift = 6755399441055744.0;                  // 1.5 * 2^52
  const double kd = fma(t, (double)HB_CHAIN_EXP_N, shift);
  const int ki = hb_lo(kd);
  const double n = kd - shift;
  const double s = hb_ld_cexp(T, (ki << 3) & ((HB_CHAIN_EXP_N - 1) << 3));
  const double f = fma(n, -1.0 / HB_CHAIN_EXP_N, t);        // exact
  const double f2 = f * f;
  const double fl = f * HB_CLIT(HB_C_E1);
  double p = fma(f, HB_CLIT(HB_C_E4), HB_CLIT(HB_C_E3));
  p = fma(p, f, HB_CLIT(HB_C_E2));
  p = fma(f2, p, fl);                                       // 2^f - 1
  const double v = fma(s, p, s);
  return hb_make(hb_hi(v) + ((ki >> HB_CHAIN_EXP_BITS) << 20), hb_lo(v));
}

}  // namespace hb
