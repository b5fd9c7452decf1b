"""The deterministic transcendental functions (pyhelp_b200/csrc/help_math.h).

CPU: accuracy against exact values (tests/golden/math_golden.npz, made with mpmath by
tests/golden/make_math_golden.py) -- pow/exp/log < 0.6 ulp, sin/cos/atan < 2 ulp, the REAL*4
wrappers correctly rounded.  GPU: the device functions return the SAME BITS as the CPU build
(this is what makes engine-vs-oracle parity exact, DESIGN.md "Numerics")."""
import os.path as osp

import numpy as np
import pytest

from oracle import oracle

G = np.load(osp.join(osp.dirname(osp.abspath(__file__)), "golden", "math_golden.npz"))


def ulps(got, exact):
    exact = exact.astype(np.float64)
    ok = np.isfinite(exact) & (exact != 0) & (np.abs(exact) > 1e-300) & (np.abs(exact) < 1e300)
    return np.abs(got[ok] - exact[ok]) / np.spacing(np.abs(exact[ok]))


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


@pytest.mark.parametrize("fn,xk,yk,ek", [
    ("r4_pow", "r4pow_x", "r4pow_y", "r4pow"), ("r4_exp", "r4exp_x", None, "r4exp"), ("r4_log", "r4log_x", None, "r4log"),
    ("r4_log10", "r4log_x", None, "r4log10"), ("r4_sin", "r4sin_x", None, "r4sin"), ("r4_cos", "r4sin_x", None, "r4cos"),
    ("r4_tan", "r4tan_x", None, "r4tan"), ("r4_acos", "r4acos_x", None, "r4acos"), ("r4_pow8", "r4pow8_x", None, "r4pow8"),
    ("r4_pow7", "r4pow8_x", None, "r4pow7"), ("r4_pow4", "r4pow8_x", None, "r4pow4")])
def test_real4_wrappers_correctly_rounded(fn, xk, yk, ek):
    got = oracle.math_eval(fn, G[xk], None if yk is None else G[yk]).astype(np.float32)
    want = G[ek].astype(np.float32)       # exact -> double -> float; double rounding is harmless at 53 vs 24 bits
    assert (got == want).mean() >= 0.9995, fn
    assert np.abs(got.astype(np.float64) - want).max() <= np.spacing(np.abs(want)).max()


def test_special_values():
    ev = oracle.math_eval
    assert ev("pow", [1.0, 5.0, 0.0, 0.0, 2.0, 4.0], [123.4, 0.0, 2.0, 0.0, 10.0, 0.5]).tolist() == [1.0, 1.0, 0.0, 1.0, 1024.0, 2.0]
    assert np.isnan(ev("pow", [-2.0], [0.5])[0]) and ev("pow", [-2.0], [3.0])[0] == -8.0
    assert np.isinf(ev("pow", [0.0], [-1.0])[0]) and ev("pow", [1e-310], [2.0])[0] == 0.0
    assert ev("exp", [0.0, 800.0, -800.0]).tolist() == [1.0, np.inf, 0.0]
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


def test_glibc_build_differs_only_in_last_place():
    """the second oracle build uses the host libm; same functions to within an ulp or two"""
    a = oracle.math_eval("pow", G["pow_x"][:500], G["pow_y"][:500])
    b = oracle.math_eval("pow", G["pow_x"][:500], G["pow_y"][:500], libm="glibc")
    ok = np.isfinite(a) & (a != 0)
    assert (np.abs(a[ok] - b[ok]) / np.spacing(np.abs(a[ok]))).max() <= 1.0


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
