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
