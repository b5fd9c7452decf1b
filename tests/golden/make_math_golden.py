"""Golden vectors for tests/test_help_math.py: inputs and the exact values (mpmath, 200 bits,
rounded to the nearest double) of the functions help_math.h implements.
Run here: python tests/golden/make_math_golden.py  ->  tests/golden/math_golden.npz"""
import os
import numpy as np
import mpmath as mp

mp.mp.prec = 200
rng = np.random.default_rng(20260101)
N = 3000
out = {}


def exact(fn, xs, ys=None):
    r = []
    for i, a in enumerate(xs):
        v = fn(mp.mpf(float(a))) if ys is None else fn(mp.mpf(float(a)), mp.mpf(float(ys[i])))
        r.append(float(v))
    return np.array(r, dtype=np.float64)


# pow on the engine's domain: relative saturations in (0, 1], exponents 3 + 2/lambda and -1/lambda
x = np.concatenate([rng.uniform(1e-4, 1.0, N), 10 ** rng.uniform(-8, 3, N)])
y = np.concatenate([rng.uniform(0.02, 45.0, N), rng.uniform(-30, 30, N)])
out["pow_x"], out["pow_y"], out["pow"] = x, y, exact(lambda a, b: a ** b, x, y)
x = rng.uniform(-700, 700, N); out["exp_x"], out["exp"] = x, exact(mp.exp, x)
x = np.concatenate([10 ** rng.uniform(-300, 300, N), rng.uniform(0.5, 2.0, N)]); out["log_x"], out["log"] = x, exact(mp.log, x)
x = rng.uniform(-20, 20, N); out["sin_x"], out["sin"], out["cos"] = x, exact(mp.sin, x), exact(mp.cos, x)
x = np.concatenate([rng.uniform(-3, 3, N), 10 ** rng.uniform(-6, 6, N)]); out["atan_x"], out["atan"] = x, exact(mp.atan, x)
# REAL*4 wrappers: float inputs, exact value as double (the test rounds it to float)
xf = rng.uniform(0.01, 1.0, N).astype(np.float32); yf = rng.uniform(0.1, 3.0, N).astype(np.float32)
out["r4pow_x"], out["r4pow_y"], out["r4pow"] = xf, yf, exact(lambda a, b: a ** b, xf, yf)
xf = rng.uniform(-30, 8, N).astype(np.float32); out["r4exp_x"], out["r4exp"] = xf, exact(mp.exp, xf)
xf = rng.uniform(1e-3, 1e3, N).astype(np.float32); out["r4log_x"], out["r4log"], out["r4log10"] = xf, exact(mp.log, xf), exact(mp.log10, xf)
xf = rng.uniform(-7, 7, N).astype(np.float32); out["r4sin_x"], out["r4sin"], out["r4cos"] = xf, exact(mp.sin, xf), exact(mp.cos, xf)
xf = rng.uniform(-1.5, 1.5, N).astype(np.float32); out["r4tan_x"], out["r4tan"] = xf, exact(mp.tan, xf)
xf = rng.uniform(-1, 1, N).astype(np.float32); out["r4acos_x"], out["r4acos"] = xf, exact(mp.acos, xf)
xf = rng.uniform(0.5, 1.3, N).astype(np.float32); out["r4pow8_x"], out["r4pow8"], out["r4pow7"], out["r4pow4"] = (
    xf, exact(lambda a: a ** 8, xf), exact(lambda a: a ** 7, xf), exact(lambda a: a ** 4, xf))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "math_golden.npz"), **out)
print("ok", {k: v.shape for k, v in out.items() if not k.endswith(("_x", "_y"))})
