"""The engine's transcendental functions (pyhelp_b200/csrc/help_math_glibc.h).

Contract: they return the SAME BITS as the libm the reference links in this image (glibc 2.39, x86-64,
the FMA ifunc variants) -- checked here against the library itself, millions of arguments per function,
through the two oracle builds ("det" = help_math_glibc.h, "glibc" = direct libm calls).  Accuracy against
exact values (tests/golden/math_golden.npz, made with mpmath by tests/golden/make_math_golden.py) is the
library's: pow/log < 0.53 ulp, the REAL*4 routines < 1 ulp.  GPU: the device functions return the same
bits as the CPU build (this is what makes engine-vs-oracle parity exact)."""
import os.path as osp
import platform

import numpy as np
import pytest

from oracle import oracle

G = np.load(osp.join(osp.dirname(osp.abspath(__file__)), "golden", "math_golden.npz"))


def _libm_is_the_cloned_one() -> bool:
    if platform.machine() != "x86_64" or platform.libc_ver() != ("glibc", "2.39"):
        return False
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return " fma " in flags and " avx2 " in flags          # the ifunc variants help_math_glibc.h follows


needs_glibc_239 = pytest.mark.skipif(not _libm_is_the_cloned_one(),
                                     reason="bit identity is defined against glibc 2.39 / x86-64 / FMA (this image)")


def ulps(got, exact):
    exact = exact.astype(np.float64)
    ok = np.isfinite(exact) & (exact != 0) & (np.abs(exact) > 1e-300) & (np.abs(exact) < 1e300)
    return np.abs(got[ok] - exact[ok]) / np.spacing(np.abs(exact[ok]))


def _args(fn, n, rng):
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    if fn == "pow":            # the routing loop: relative saturations ** Brooks-Corey exponents; then everything
        x = np.concatenate([rng.uniform(0, 1, n // 2), 2.0 ** rng.uniform(-1070, 1023, n // 4), rng.uniform(-4, 4, n // 4)])
        y = np.concatenate([rng.uniform(-40, 40, n // 2), rng.uniform(-3, 3, n // 4), rng.integers(-9, 9, n // 4).astype(float)])
        return x, y
    if fn == "r4_pow":
        x = np.concatenate([rng.uniform(0, 4, n // 2), 2.0 ** rng.uniform(-149, 128, n // 4), rng.uniform(-4, 4, n // 4)])
        y = np.concatenate([rng.uniform(-20, 20, n // 2), rng.uniform(-3, 3, n // 4), rng.integers(-9, 9, n // 4).astype(float)])
        return f32(x), f32(y)
    if fn == "log":
        return np.concatenate([rng.uniform(0, 3, n // 3), rng.uniform(0.9, 1.1, n // 3), 2.0 ** rng.uniform(-1074, 1023, n - 2 * (n // 3))]), None
    if fn in ("r4_log", "r4_log10"):
        return f32(np.concatenate([rng.uniform(0, 50, n // 2), 2.0 ** rng.uniform(-149, 128, n // 2)])), None
    if fn == "r4_exp":
        return f32(np.concatenate([rng.uniform(-110, 95, n // 2), rng.uniform(-1, 1, n // 2) * 2.0 ** rng.uniform(-30, 8, n // 2)])), None
    if fn in ("r4_sin", "r4_cos", "r4_tan"):
        return f32(np.concatenate([rng.uniform(-10, 10, n // 2), rng.uniform(-120, 120, n // 4),
                                   rng.uniform(-1, 1, n // 4) * 2.0 ** rng.uniform(-20, 127, n // 4)])), None
    if fn == "r4_acos":
        return f32(np.concatenate([rng.uniform(-1, 1, n // 2), rng.uniform(-1, 1, n // 2) * 2.0 ** rng.uniform(-30, 0, n // 2)])), None
    if fn in ("r4_sqrt", "r4_pow8", "r4_pow7", "r4_pow4"):
        return f32(np.concatenate([rng.uniform(0, 1.2, n // 2), rng.uniform(0, 400, n // 2)])), None
    if fn in ("sin", "cos"):   # angles of drain-layer slopes and the whole range the library's first three branches cover
        return np.concatenate([rng.uniform(0, 1.5708, n // 2), rng.uniform(-2.426, 2.426, n // 4),
                               rng.uniform(-1, 1, n // 4) * 2.0 ** rng.uniform(-60, 0, n // 4)]), None
    if fn == "atan":
        return np.concatenate([rng.uniform(0, 1, n // 2), rng.uniform(-16, 16, n // 4),
                               rng.uniform(-1, 1, n // 4) * 2.0 ** rng.uniform(-60, 0, n // 4)]), None
    raise KeyError(fn)


LIBM_FUNCTIONS = ("pow", "log", "sin", "cos", "atan", "r4_exp", "r4_log", "r4_log10", "r4_sin", "r4_cos", "r4_tan", "r4_acos",
                  "r4_pow", "r4_sqrt", "r4_pow8", "r4_pow7", "r4_pow4")


@needs_glibc_239
@pytest.mark.parametrize("fn", LIBM_FUNCTIONS)
def test_same_bits_as_the_system_libm(fn):
    """help_math_glibc.h vs libm.so.6 (pow, log, sin, cos, atan, expf, logf, log10f, sinf, cosf, tanf, acosf, powf
    and powf with the exponents 0.5 / 8 / 7 / 4 the Fortran's `**` passes): every result bit-identical"""
    x, y = _args(fn, 2_000_000, np.random.default_rng(11))
    ours = oracle.math_eval(fn, x, y, libm="det")
    libm = oracle.math_eval(fn, x, y, libm="glibc")
    same = (ours.view(np.uint64) == libm.view(np.uint64)) | (np.isnan(ours) & np.isnan(libm))
    assert same.all(), (fn, int((~same).sum()), x[~same][:3], None if y is None else y[~same][:3])


@needs_glibc_239
def test_every_slope_the_text_format_can_express_gives_libm_bits():
    """atan / sin / cos are evaluated on drain-layer slopes only (LATFLO, HELP3O.FOR:4686-4712): a REAL*4 read from a
    6-character field (preprocessing.py:170 writes '{:>6.2f}').  All of them, exhaustively."""
    k = np.arange(100_000, dtype=np.float64)
    slope = np.unique(np.concatenate([(k / 10.0 ** m).astype(np.float32) for m in range(6)])).astype(np.float64) / 100.0
    slope = slope[slope < 16.0]
    a = oracle.math_eval("atan", slope, libm="det")
    assert np.array_equal(a.view(np.uint64), oracle.math_eval("atan", slope, libm="glibc").view(np.uint64))
    for fn in ("sin", "cos"):
        assert np.array_equal(oracle.math_eval(fn, a, libm="det").view(np.uint64), oracle.math_eval(fn, a, libm="glibc").view(np.uint64)), fn
    assert slope.size > 300_000


@pytest.mark.parametrize("fn,xk,yk,ek,tol", [
    ("pow", "pow_x", "pow_y", "pow", 1.1), ("exp", "exp_x", None, "exp", 1.1), ("log", "log_x", None, "log", 1.1),
    ("sin", "sin_x", None, "sin", 2.5), ("cos", "sin_x", None, "cos", 2.5), ("atan", "atan_x", None, "atan", 2.5)])
def test_double_accuracy(fn, xk, yk, ek, tol):
    """error vs the exact value ROUNDED to double: < 0.6 ulp true error shows as <= 1 ulp here"""
    got = oracle.math_eval(fn, G[xk], None if yk is None else G[yk])
    u = ulps(got, G[ek])
    assert u.max() <= tol, (fn, u.max())
    if fn in ("pow", "exp", "log"):
        assert (u == 0).mean() > 0.97           # correctly rounded in all but a few per cent of cases


@pytest.mark.parametrize("fn,xk,yk,ek,cr", [
    ("r4_pow", "r4pow_x", "r4pow_y", "r4pow", 0.995), ("r4_exp", "r4exp_x", None, "r4exp", 0.995), ("r4_log", "r4log_x", None, "r4log", 0.995),
    ("r4_log10", "r4log_x", None, "r4log10", 0.7), ("r4_sin", "r4sin_x", None, "r4sin", 0.97), ("r4_cos", "r4sin_x", None, "r4cos", 0.97),
    ("r4_tan", "r4tan_x", None, "r4tan", 0.7), ("r4_acos", "r4acos_x", None, "r4acos", 0.7), ("r4_pow8", "r4pow8_x", None, "r4pow8", 0.995),
    ("r4_pow7", "r4pow8_x", None, "r4pow7", 0.995), ("r4_pow4", "r4pow8_x", None, "r4pow4", 0.995)])
def test_real4_accuracy(fn, xk, yk, ek, cr):
    """the library's REAL*4 routines: never more than 1 ulp from the exact value, correctly rounded in most cases
    (expf/logf/powf/sinf/cosf almost always; the older tanf/acosf/log10f less often)"""
    got = oracle.math_eval(fn, G[xk], None if yk is None else G[yk]).astype(np.float32)
    want = G[ek].astype(np.float32)       # exact -> double -> float; double rounding is harmless at 53 vs 24 bits
    assert (got == want).mean() >= cr, (fn, float((got == want).mean()))
    ok = np.isfinite(want)
    assert (np.abs(got[ok].astype(np.float64) - want[ok]) <= np.spacing(np.abs(want[ok])).astype(np.float64)).all(), fn


def test_special_values():
    ev = oracle.math_eval
    assert ev("pow", [1.0, 5.0, 0.0, 0.0, 2.0, 4.0], [123.4, 0.0, 2.0, 0.0, 10.0, 0.5]).tolist() == [1.0, 1.0, 0.0, 1.0, 1024.0, 2.0]
    assert np.isnan(ev("pow", [-2.0], [0.5])[0]) and ev("pow", [-2.0], [3.0])[0] == -8.0
    assert np.isinf(ev("pow", [0.0], [-1.0])[0]) and ev("pow", [1e-310], [2.0])[0] == 0.0
    assert ev("exp", [0.0, 800.0, -800.0]).tolist() == [1.0, np.inf, 0.0]
    # NaN operands return early from the special-case routine: the library's results
    nan = float("nan")
    got = ev("pow", [nan, nan, nan, nan, 1.0, 2.0, -1.0, nan, -nan], [3.0, 2.5, 0.0, -0.0, nan, nan, nan, np.inf, 3.0])
    assert np.isnan(got[[0, 1, 5, 6, 7, 8]]).all() and (got[[2, 3, 4]] == 1.0).all()
    assert ev("log", [1.0])[0] == 0.0 and np.isnan(ev("log", [-1.0])[0]) and ev("log", [0.0])[0] == -np.inf
    assert ev("sin", [0.0])[0] == 0.0 and ev("cos", [0.0])[0] == 1.0 and ev("atan", [0.0])[0] == 0.0


def test_dual_pow_returns_the_same_bits():
    """det_pow2 (two chains side by side, used by the kernel's capillary pre-pass) == det_pow"""
    x = np.concatenate([G["pow_x"], [0.0, -1.0, 1.0, 1e-320, np.inf, 2.0, 1e-5]])
    y = np.concatenate([G["pow_y"], [2.0, 2.0, 7.0, 0.5, 0.5, 1e200, 200.0]])
    a = oracle.math_eval("pow", x, y)
    b = oracle.math_eval("pow2", x, y)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_shared_logarithm_pow_returns_the_same_bits():
    """det_pow_samebase (x**(y-1) and x**y from one logarithm: the Newton step of the drainage
    solve, pyhelp/HELP3O.FOR:4470-4473) == det_pow"""
    x = np.concatenate([G["pow_x"], [0.0, -1.0, 1.0, 1e-320, np.inf, 2.0, 1e-5]])
    y = np.concatenate([G["pow_y"], [2.0, 2.0, 7.0, 0.5, 0.5, 1e200, 200.0]])
    a = oracle.math_eval("pow", x, y)
    b = oracle.math_eval("pow_sb", x, y)
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_division_through_the_rounded_reciprocal_is_the_ieee_quotient():
    """div_by_const(a, b, RN(1/b)) == a / b bit for bit: the routing loop divides by the
    segment constant UL-RS this way (help_image.h fdiv_by)"""
    rng = np.random.default_rng(3)
    n = 2_000_000
    b = np.concatenate([rng.uniform(1e-3, 60.0, n), rng.uniform(0.5, 1.0, n).astype(np.float32).astype(np.float64)])
    a = np.concatenate([rng.uniform(-5.0, 60.0, n), rng.uniform(0.0, 1.0, n) * b[n:]])
    # mantissas of all ones / few bits, exact quotients, zero numerators
    b[:1000] = np.nextafter(2.0 ** rng.integers(-8, 6, 1000), 0.0)
    a[1000:2000] = b[1000:2000] * rng.integers(1, 1 << 20, 1000)
    a[2000:2100] = 0.0
    got = oracle.math_eval("fdiv", a, b)
    assert np.array_equal(got.view(np.uint64), (a / b).view(np.uint64))


@pytest.mark.gpu
@pytest.mark.parametrize("fn", list(oracle.MATH_FN))
def test_device_functions_match_cpu_bits(fn):
    import ctypes as C
    from pyhelp_b200 import _native as N
    rng = np.random.default_rng(7)
    n = 200_000
    if fn == "fdiv":
        x = np.concatenate([rng.uniform(-5.0, 60.0, n // 2), rng.uniform(0.0, 1.0, n // 2)])
        y = np.concatenate([rng.uniform(1e-3, 60.0, n // 2), rng.uniform(0.5, 1.0, n // 2).astype(np.float32).astype(np.float64)])
    elif fn in ("pow", "r4_pow", "pow2", "pow_sb"):
        x = np.concatenate([rng.uniform(1e-6, 1.0, n // 2), 10 ** rng.uniform(-12, 4, n // 2)])
        y = np.concatenate([rng.uniform(0.01, 60.0, n // 2), rng.uniform(-40, 40, n // 2)])
    elif fn in ("exp", "r4_exp"):
        x, y = rng.uniform(-720, 720, n), None
    elif fn in ("log", "r4_log", "r4_log10", "r4_sqrt"):
        x, y = 10 ** rng.uniform(-30, 30, n), None
    elif fn in ("r4_acos",):
        x, y = rng.uniform(-1, 1, n), None
    elif fn in ("r4_pow8", "r4_pow7", "r4_pow4"):
        x, y = rng.uniform(0.2, 350.0, n), None
    else:
        x, y = rng.uniform(-50, 50, n), None
    if fn.startswith("r4_"):
        x = x.astype(np.float32).astype(np.float64)
        y = None if y is None else y.astype(np.float32).astype(np.float64)
    cpu = oracle.math_eval(fn, x, y)
    dev = np.empty_like(x)
    N.check(N.lib().help_b200_math_eval(oracle.MATH_FN[fn], x.ctypes.data, None if y is None else y.ctypes.data,
                                        dev.ctypes.data, x.size), "math_eval")
    assert np.array_equal(cpu.view(np.uint64), dev.view(np.uint64)), (fn, int((cpu.view(np.uint64) != dev.view(np.uint64)).sum()))
