// Host build of hydpy_b200/csrc/fastmath.cuh (the very same source the kernel uses) for accuracy tests on the CPU.
// g++ -O2 -ffp-contract=off -shared -fPIC tests/fastmath_host.cpp -o tests/_fastmath_host.so
#include "../hydpy_b200/csrc/fastmath.cuh"

static hb::MathTables g_tables;
static bool g_init = false;
static void init() { if (!g_init) { hb::make_math_tables(g_tables); g_init = true; } }

extern "C" {
void fm_pow(const double* x, const double* y, double* out, long n) { init(); for (long i = 0; i < n; ++i) out[i] = hb::hb_pow(x[i], y[i], hb::hb_tabref(&g_tables)); }
void fm_exp(const double* x, double* out, long n) { init(); for (long i = 0; i < n; ++i) out[i] = hb::hb_exp(x[i], hb::hb_tabref(&g_tables)); }
void fm_log(const double* x, double* out, long n) { init(); for (long i = 0; i < n; ++i) out[i] = hb::hb_log_normal(x[i], hb::hb_tabref(&g_tables)); }
void fm_logfull(const double* x, double* out, long n) { init(); for (long i = 0; i < n; ++i) out[i] = hb::hb_log(x[i], hb::hb_tabref(&g_tables)); }
// errors against x87 extended precision libm
// RecStep chain: q = 2^(y*log2(d) + c2) with c2 = log2(dtk) - y*log2(ca); reference value dtk * (d/ca)^y in extended precision
void fm_chain(const double* d, const double* y, const double* dtk, const double* ca, double* out, double* rel, int* ok, long n) {
  static hb::ChainTables ct; static bool ci = false;
  if (!ci) { hb::make_chain_tables(ct); ci = true; }
  for (long i = 0; i < n; ++i) {
    const double c2 = (double)(log2l((long double)dtk[i]) - (long double)y[i] * log2l((long double)ca[i]));
    double t;
    out[i] = hb::hb_chain_pow2(d[i], y[i], c2, hb::hb_chainref(&ct), t);
    ok[i] = fabs(t) < 800.0;    // (the caller's range check: land_kernel_v2.cuh, element_kernel)
    const long double r = (long double)dtk[i] * powl((long double)d[i] / (long double)ca[i], (long double)y[i]);
    rel[i] = (double)fabsl(((long double)out[i] - r) / r);
  }
}
void fm_pow_relerr(const double* x, const double* y, const double* got, double* rel, long n) {
  for (long i = 0; i < n; ++i) { long double r = powl((long double)x[i], (long double)y[i]); rel[i] = (double)fabsl(((long double)got[i] - r) / r); }
}
void fm_exp_relerr(const double* x, const double* got, double* rel, long n) {
  for (long i = 0; i < n; ++i) { long double r = expl((long double)x[i]); rel[i] = (double)fabsl(((long double)got[i] - r) / r); }
}
void fm_log_abserr(const double* x, const double* got, double* err, long n) {
  for (long i = 0; i < n; ++i) { long double r = logl((long double)x[i]); err[i] = (double)fabsl((long double)got[i] - r); }
}
}
