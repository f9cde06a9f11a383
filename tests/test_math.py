"""CPU: accuracy of the product's portable elementary functions (csrc/tofrt_math.h),
evaluated through the oracle's device-math twin, against mpmath."""
import numpy as np
import pytest

mp = pytest.importorskip("mpmath")
from oracle.oracle import _dp  # noqa: E402


def _eval(o, fn, x, y=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(np.zeros_like(x) if y is None else y, dtype=np.float64)
    out = np.empty_like(x)
    assert o.lib.tofo_math_eval(fn, _dp(x), _dp(y), len(x), _dp(out)) == 0
    return out


def _max_ulp(got, exact, args):
    mp.mp.prec = 160
    worst = 0.0
    for i in range(len(got)):
        ex = exact(*[mp.mpf(float(a[i])) for a in args])
        exf = float(ex)
        if exf == 0.0 or not np.isfinite(exf):
            continue
        err = abs((mp.mpf(float(got[i])) - ex) / mp.mpf(float(np.spacing(abs(exf)))))
        worst = max(worst, float(err))
    return worst


def test_trig_accuracy(oracle_dm):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-7, 7, 6000), rng.uniform(-1e-3, 1e-3, 500),
                        np.arange(-8, 9) * np.pi / 2 + rng.uniform(-1e-9, 1e-9, 17), rng.uniform(-1e5, 1e5, 500)])
    assert _max_ulp(_eval(oracle_dm, 0, x), mp.sin, [x]) < 1.5
    assert _max_ulp(_eval(oracle_dm, 1, x), mp.cos, [x]) < 1.5
    assert _max_ulp(_eval(oracle_dm, 2, x), mp.tan, [x]) < 2.5


def test_inverse_trig_accuracy(oracle_dm):
    rng = np.random.default_rng(2)
    xa = np.concatenate([rng.uniform(-3, 3, 5000), rng.standard_cauchy(1500) * 10, 10.0 ** rng.uniform(-10, 18, 500)])
    assert _max_ulp(_eval(oracle_dm, 3, xa), mp.atan, [xa]) < 2.0
    yy, xx = rng.uniform(-1, 1, 6000), rng.uniform(-1, 1, 6000)
    assert _max_ulp(_eval(oracle_dm, 4, yy, xx), mp.atan2, [yy, xx]) < 2.0
    xs = np.concatenate([rng.uniform(-1, 1, 5000), 1 - 10.0 ** rng.uniform(-12, -1, 500)])
    assert _max_ulp(_eval(oracle_dm, 5, xs), mp.asin, [xs]) < 3.0
    # the fused sin/cos of atan2(y, x): (y, x) / hypot -- closer to exact than chaining atan2 with sin/cos
    yy, xx = rng.uniform(-0.1, 0.1, 6000), rng.uniform(-0.1, 0.1, 6000)
    assert _max_ulp(_eval(oracle_dm, 6, yy, xx), lambda y, x: y / mp.sqrt(x * x + y * y), [yy, xx]) < 2.5
    assert _max_ulp(_eval(oracle_dm, 7, yy, xx), lambda y, x: x / mp.sqrt(x * x + y * y), [yy, xx]) < 2.5


def test_special_values(oracle_dm, oracle_libm):
    sp = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 1.0000001])
    y = np.array([1.0, -0.0, np.inf, -np.inf, 1.0, 0.0, -0.0, 1.0])
    for fn in range(6):   # (the fused pair 6, 7 is defined for finite lens coordinates only)
        a, b = _eval(oracle_dm, fn, sp, y), _eval(oracle_libm, fn, sp, y)
        assert np.array_equal(np.isnan(a), np.isnan(b)), fn
        ok = ~np.isnan(a)
        assert np.allclose(a[ok], b[ok], rtol=4e-16, atol=0), fn
        nz = ok & (b != 0)  # sin/tan do not preserve the sign of a zero argument (documented)
        assert np.array_equal(np.signbit(a[nz]), np.signbit(b[nz])), fn
    # beyond the documented domain the trigonometric functions return NaN, never garbage
    assert np.isnan(_eval(oracle_dm, 0, np.array([2e9, -3e300]))).all()
