// fastmath.cuh — table-driven fp64 log / exp / pow for the runoff-generation kernel.
//
// The reference calls libm `pow` once per UZ sub-step (hydpy/models/hland/hland_model.py:2474), per soil zone
// (:1783) and per contributing-area factor (:2233-2236) and `exp` once per zone (evap_model.py:5324) — ≈76 `pow` per
// element-step for the hourly Lahn parameterisation.  libdevice's correctly-special-cased `pow` costs ≈250 SASS
// instructions (ncu: 53 % of all instructions of the first kernel version, profiles/r01_land_v1_*.txt); the routines
// below need ≈25 fp64 instructions + a few integer ones and two shared-memory look-ups:
//
//   log(x)  = e*ln2 + log(c_k) + log1p(r),   r = m*(1/c_k) - 1,  |r| <= 2^-7 (2^-6 in the interval at 1.0, where
//             c_0 = 1 keeps log(1) == 0 and hence pow(1, y) == 1 exactly), degree-8 polynomial;
//   exp(t)  = 2^q * 2^(j/64) * (1 + p(r)),   n = rint(t*64/ln2) = 64*q + j,  r = t - n*ln2/64,  degree-5 polynomial;
//   pow(x,y)= exp(y*log(x))  for normal positive x and |y*log(x)| < 690; everything else (zero, negative, subnormal,
//             inf, nan, overflow, underflow) goes to libdevice `pow` so that all IEEE special cases keep the
//             reference's libm semantics.
//
// Accuracy (tests/test_fastmath.py compiles THIS file for the host and compares with x87 extended precision; the GPU
// parity tests bound the end effect): log: 4e-16 + 1 ulp absolute, exp: < 2 ulp, pow: relative error
// < 5e-16 + 2*|y*log x|*2^-52 (the exponent is carried in one double), i.e. < 1e-14 for everything the model feeds
// it — five orders of magnitude inside the 1e-9 parity bar.  All fused multiply-adds are explicit `fma()` calls; the translation unit is compiled with -fmad=false.
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

#define HB_LOG_TABLE_BITS 6
#define HB_LOG_N (1 << HB_LOG_TABLE_BITS)
#define HB_EXP_TABLE_BITS 6
#define HB_EXP_N (1 << HB_EXP_TABLE_BITS)

struct MathTables {
  double invc[HB_LOG_N];   // 1/c_k rounded to double (c_0 = 1)
  double logc[HB_LOG_N];   // -log(invc[k]) — consistent with the ROUNDED reciprocal
  double exp2[HB_EXP_N];   // 2^(j/64)
};

// host side: tables in long double precision (x87 80-bit is plenty for 1e-19)
inline void make_math_tables(MathTables& t) {
  for (int k = 0; k < HB_LOG_N; ++k) {
    const long double c = (k == 0) ? 1.0L : 1.0L + ((long double)k + 0.5L) / HB_LOG_N;
    t.invc[k] = (double)(1.0L / c);
    t.logc[k] = (k == 0) ? 0.0 : (double)(-logl((long double)t.invc[k]));
  }
  for (int j = 0; j < HB_EXP_N; ++j) t.exp2[j] = (double)exp2l((long double)j / HB_EXP_N);
}

HB_HD double hb_log_normal(double x, const MathTables* __restrict__ T) {
  // x must be a positive normal number
  const int hi = hb_hi(x), lo = hb_lo(x);
  const int e = (hi >> 20) - 1023;
  const int k = (hi >> (20 - HB_LOG_TABLE_BITS)) & (HB_LOG_N - 1);
  const double m = hb_make((hi & 0x000fffff) | 0x3ff00000, lo);  // [1, 2)
  const double r = fma(m, T->invc[k], -1.0);
  const double r2 = r * r;
  // log1p(r) = r - r^2/2 + r^3/3 - ... - r^8/8; coefficient pairs first to shorten the dependency chain
  const double p78 = fma(r, -1.0 / 8.0, 1.0 / 7.0);
  const double p56 = fma(r, -1.0 / 6.0, 1.0 / 5.0);
  const double p34 = fma(r, -1.0 / 4.0, 1.0 / 3.0);
  double q = fma(r2, p78, p56);
  q = fma(r2, q, p34);             // 1/3 - r/4 + r^2/5 - ... - r^5/8
  q = fma(r, q, -0.5);             // -1/2 + r/3 - ...
  const double tail = fma(r2, q, r);
  const double de = (double)e;
  const double ln2hi = 6.93147180369123816490e-01;   // 0x1.62e42fee00000p-1: e*ln2hi is exact
  const double ln2lo = 1.90821492927058770002e-10;
  return fma(de, ln2hi, T->logc[k]) + fma(de, ln2lo, tail);
}

HB_HD double hb_exp_bounded(double t, const MathTables* __restrict__ T) {
  // |t| < 690: the result is a normal number
  const double inv = 92.33248261689366;             // 64/ln2
  const double l64hi = 0.01083042469326756;         // 0x1.62e42fee00000p-7: n*l64hi is exact for |n| < 2^21
  const double l64lo = 2.9815858269852933e-12;      // ln2/64 - l64hi
  const double shift = 6755399441055744.0;          // 1.5 * 2^52
  const double kd = fma(t, inv, shift);
  const int ki = hb_lo(kd);
  const double n = kd - shift;
  double r = fma(n, -l64hi, t);
  r = fma(n, -l64lo, r);
  const double s = T->exp2[ki & (HB_EXP_N - 1)];
  const int q = ki >> HB_EXP_TABLE_BITS;
  const double r2 = r * r;
  double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(r, p, 1.0 / 6.0);
  p = fma(r, p, 0.5);
  p = fma(r2, p, r);                                 // e^r - 1
  const double v = fma(s, p, s);
  return hb_make(hb_hi(v) + (q << 20), hb_lo(v));
}

HB_HD double hb_exp(double t, const MathTables* __restrict__ T) {
  if (fabs(t) < 690.0) return hb_exp_bounded(t, T);
  return exp(t);
}

HB_HD double hb_pow(double x, double y, const MathTables* __restrict__ T) {
  const int hi = hb_hi(x);
  // positive normal x: 0x00100000 <= hi < 0x7ff00000
  if ((unsigned)(hi - 0x00100000) < 0x7fe00000u) {
    const double t = y * hb_log_normal(x, T);
    if (fabs(t) < 690.0) return hb_exp_bounded(t, T);
  }
  return pow(x, y);
}


}  // namespace hb
