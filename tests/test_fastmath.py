"""Accuracy of the engine's table-driven log/exp/pow (hydpy_b200/csrc/fastmath.cuh), host build of the SAME source,
against x87 extended-precision libm.  The bounds asserted here are the ones DESIGN.md quotes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_fastmath_host.so")


@pytest.fixture(scope="module")
def fm():
    src = os.path.join(HERE, "fastmath_host.cpp")
    hdr = os.path.join(HERE, "..", "hydpy_b200", "csrc", "fastmath.cuh")
    if not os.path.exists(SO) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(SO):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", SO], check=True)
    lib = C.CDLL(SO)
    return lib


def p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_pow_relative_error(fm):
    rng = np.random.default_rng(0)
    n = 400_000
    # the model's ranges: ratios in (0, ~1e3], exponents 0.05 … 5
    x = np.concatenate([rng.uniform(1e-8, 2.0, n // 2), np.exp(rng.uniform(-30, 12, n // 2))])
    y = rng.uniform(0.05, 5.0, n)
    out, rel = np.zeros(n), np.zeros(n)
    fm.fm_pow(p(x), p(y), p(out), n)
    fm.fm_pow_relerr(p(x), p(y), p(out), p(rel), n)
    # the exponent t = y*log(x) is held in one double: its two roundings (log, product) bound the relative error of
    # the result by ≈ 2*|t|*2^-53, plus ≈ 2 ulp from the table/polynomial steps
    t = np.abs(y * np.log(x))
    bound = 5e-16 + 2.0 * t * 2.3e-16
    assert np.all(rel <= bound), (rel.max(), x[np.argmax(rel / bound)], y[np.argmax(rel / bound)])
    typical = t < 20.0                     # what the model produces (ratios 1e-4 … 1e3, exponents <= 5)
    assert rel[typical].max() < 1e-14
    print("pow: max rel err %.3e, median %.3e" % (rel.max(), np.median(rel)))


def test_pow_special_cases_follow_libm(fm):
    x = np.array([0.0, 1.0, 1.0, 2.0, 0.5, np.inf, np.nan, -1.0, 1e-310, 4.0, 1e300, 1e-300, 7.3])
    y = np.array([2.0, 3.7, 0.0, 0.0, -1.0, 1.5, 1.0, 2.0, 0.5, 0.5, 2.0, 2.0, 1.0])
    out = np.zeros(len(x))
    fm.fm_pow(p(x), p(y), p(out), len(x))
    with np.errstate(all="ignore"):
        ref = np.power(x, y)
    assert np.array_equal(out[:4], [0.0, 1.0, 1.0, 1.0])       # exact: pow(0,y)=0, pow(1,y)=1, pow(x,0)=1
    np.testing.assert_allclose(out, ref, rtol=1e-15, equal_nan=True)
    assert out[10] == np.inf and out[11] == 0.0


def test_exp_relative_error(fm):
    rng = np.random.default_rng(1)
    n = 400_000
    x = np.concatenate([rng.uniform(-40, 5, n // 2), rng.uniform(-700, 700, n // 2)])
    out, rel = np.zeros(n), np.zeros(n)
    fm.fm_exp(p(x), p(out), n)
    fm.fm_exp_relerr(p(x), p(out), p(rel), n)
    assert rel.max() < 4.5e-16, rel.max()
    z = np.array([0.0, -0.0, 710.0, -750.0, np.nan])
    o = np.zeros(5)
    fm.fm_exp(p(z), p(o), 5)
    assert o[0] == 1.0 and o[1] == 1.0 and o[2] == np.inf and o[3] == 0.0 and np.isnan(o[4])
    print("exp: max rel err %.3e" % rel.max())


def test_log_absolute_error(fm):
    rng = np.random.default_rng(2)
    n = 400_000
    x = np.concatenate([rng.uniform(0.5, 2.0, n // 2), np.exp(rng.uniform(-700, 700, n // 2))])
    out, err = np.zeros(n), np.zeros(n)
    fm.fm_log(p(x), p(out), n)
    fm.fm_log_abserr(p(x), p(out), p(err), n)
    tol = 4e-16 + 2.3e-16 * np.abs(np.log(x))
    assert np.all(err <= tol), (err.max(), x[np.argmax(err / tol)])
    one = np.array([1.0])
    o = np.zeros(1)
    fm.fm_log(p(one), p(o), 1)
    assert o[0] == 0.0


def test_log_special_cases_and_contributing_area_identity(fm):
    """`hb_log` (used for Calc_ContriArea_V1 in the fast path: (prod (sm/fc)^w)^beta = exp(beta * sum(w * log(sm/fc)))):
    libm semantics for zero, subnormal, negative, inf, nan; and the identity holds to 1e-14 for the model's ranges."""
    x = np.array([0.0, 1.0, 1e-310, -1.0, np.inf, np.nan, 0.5])
    out = np.zeros(len(x))
    fm.fm_logfull(p(x), p(out), len(x))
    with np.errstate(all="ignore"):
        ref = np.log(x)
    assert out[0] == -np.inf and out[1] == 0.0 and np.isnan(out[3]) and out[4] == np.inf and np.isnan(out[5])
    np.testing.assert_allclose(out, ref, rtol=1e-15, equal_nan=True)
    rng = np.random.default_rng(3)
    nz, ne = 12, 20_000
    ratio = rng.uniform(1e-4, 1.0, (ne, nz))                  # sm/fc of the soil zones of an element
    w = rng.uniform(0.2, 1.0, (ne, nz)); w /= w.sum(axis=1, keepdims=True)
    beta = rng.uniform(1.0, 4.0, ne)
    lg = np.zeros(ne * nz)
    fm.fm_logfull(p(np.ascontiguousarray(ratio.reshape(-1))), p(lg), ne * nz)
    t = beta * np.sum(w * lg.reshape(ne, nz), axis=1)
    got = np.zeros(ne)
    fm.fm_exp(p(np.ascontiguousarray(t)), p(got), ne)
    ref = np.prod(np.power(ratio.astype(np.longdouble), w.astype(np.longdouble)), axis=1) ** beta.astype(np.longdouble)
    rel = np.abs((got - ref) / ref).astype(np.float64)
    assert rel.max() < 1e-14, rel.max()


def test_recstep_chain_pow(fm):
    """`hb_chain_pow2`, the base-2 power function of the RecStep chain (Calc_Q0_Perc_UZ_V1): dt*k*(uz/contriarea)^(1+alpha)
    over the model's ranges within 2e-14 of extended precision; out-of-range arguments are flagged, never trusted."""
    rng = np.random.default_rng(4)
    n = 400_000
    d = np.concatenate([np.exp(rng.uniform(-30, 8, n // 2)), rng.uniform(1e-6, 200.0, n // 2)])   # uz [mm]
    y = rng.uniform(1.0, 4.0, n)                        # 1 + alpha
    dtk = np.exp(rng.uniform(-14, 0, n))                # k / recstep
    ca = np.exp(rng.uniform(-12, 0, n))                 # contributing area (0, 1]
    out, rel = np.zeros(n), np.zeros(n)
    ok = np.zeros(n, np.int32)
    fm.fm_chain(p(d), p(y), p(dtk), p(ca), p(out), p(rel), ok.ctypes.data_as(C.POINTER(C.c_int)), n)
    assert ok.all()
    # the exponent log2(q) is accumulated in one double (like hb_pow's): c2 and the partial sums round at the magnitude
    # of their terms, the table/polynomial steps add about 4 ulp
    t = np.abs(np.log2(dtk)) + y * np.abs(np.log2(ca)) + y * np.abs(np.log2(d))
    bound = 2.3e-16 * (4.0 + 1.5 * t)
    assert np.all(rel <= bound), (rel.max(), (rel / bound).max())
    typical = (d > 1e-3) & (ca > 0.01)                          # what the model produces
    assert rel[typical].max() < 2e-14, rel[typical].max()
    print("chain pow: max rel err %.3e, median %.3e" % (rel.max(), np.median(rel)))
    # flagged: zero, negative, subnormal, inf, nan arguments and exponents beyond +-1000
    d = np.array([0.0, -1.0, 1e-310, np.inf, np.nan, 1e300, 1e-300, 1.0])
    y = np.array([2.0, 2.0, 2.0, 2.0, 2.0, 4.0, 4.0, 2.0])
    dtk = np.full(8, 0.01); ca = np.full(8, 0.5)
    out, rel = np.zeros(8), np.zeros(8)
    ok = np.ones(8, np.int32)
    fm.fm_chain(p(d), p(y), p(dtk), p(ca), p(out), p(rel), ok.ctypes.data_as(C.POINTER(C.c_int)), 8)
    assert list(ok) == [0, 0, 0, 0, 0, 0, 0, 1]
    assert out[7] == pytest.approx(0.04, rel=1e-15)
