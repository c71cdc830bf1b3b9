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

}  // namespace hb
