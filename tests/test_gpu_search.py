"""GPU: the return-path search on ADVERSARIAL lines, against NumPy's literal evaluation of
main_rt.py:78-99 (np.sign / np.diff on the engine's own lens table).  Uses the line-injection test
hook (tofrt_debug_search_host), which runs the production kernel with the reflected lines replaced,
so both search modes are exercised exactly as shipped."""
import numpy as np
import pytest

from ray_tracing_ultrasound_b200 import reference_setup as rs

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("math_mode")]


def numpy_rule(x, z, m, b):
    """main_rt.py:78-99 (+ :84-96), verbatim NumPy; same idx encoding as tofrt_ray_records."""
    out = np.empty(len(m), dtype=np.int32)
    with np.errstate(all="ignore"):
        for i in range(len(m)):
            diffs = z - (m[i] * x + b[i])
            sc = np.where(np.diff(np.sign(diffs)) != 0)[0]
            if len(sc):
                out[i] = sc[0]
                continue
            on_line = np.isclose(diffs, 0)
            out[i] = -2 - int(np.where(on_line)[0][0]) if np.any(on_line) else -1
    return out


def numpy_intersection(x_curve, y_curve, m, b):
    """main_rt.py:78-168 for a non-vertical line given as (m, b): restated in NumPy scalar arithmetic
    (np.isclose defaults, Python min/max), returning (x, y) or (nan, nan) for (None, None)."""
    nan = (np.nan, np.nan)
    with np.errstate(all="ignore"):
        diffs = y_curve - (m * x_curve + b)
        sign_changes = np.where(np.diff(np.sign(diffs)) != 0)[0]
        if len(sign_changes) == 0:
            on_line = np.isclose(diffs, 0)
            if np.any(on_line):
                i = np.where(on_line)[0][0]
                return x_curve[i], y_curve[i]
            return nan
        idx = sign_changes[0]
        x1, y1 = x_curve[idx], y_curve[idx]
        x2, y2 = x_curve[idx + 1], y_curve[idx + 1]
        if np.isclose(x1, x2):
            y_int = m * x1 + b
            if y_int >= min(y1, y2) - 1e-9 and y_int <= max(y1, y2) + 1e-9:
                return x1, y_int
            return nan
        m_seg = (y2 - y1) / (x2 - x1)
        b_seg = y1 - m_seg * x1
        if np.isclose(m, m_seg):
            if np.isclose(b, b_seg):
                x_int = (x1 + x2) / 2.0
                return x_int, m * x_int + b
            return nan
        x_int = (b_seg - b) / (m - m_seg)
        y_int = m * x_int + b
        if (x_int >= min(x1, x2) - 1e-9 and x_int <= max(x1, x2) + 1e-9 and
                y_int >= min(y1, y2) - 1e-9 and y_int <= max(y1, y2) + 1e-9):
            return x_int, y_int
        return nan


def adversarial_lines(x, z, rng):
    n = len(x)
    seg_m = np.diff(z) / np.diff(x)
    seg_b = z[:-1] - seg_m * x[:-1]
    ms, bs = [], []

    def add(m, b):
        ms.append(np.atleast_1d(np.asarray(m, dtype=np.float64)))
        bs.append(np.atleast_1d(np.asarray(b, dtype=np.float64)))

    j = rng.integers(0, n, 600)
    m = rng.standard_cauchy(600) * 3
    add(m, z[j] - m * x[j])                                        # through a sample (diff ~ 0, +-1 ulp)
    for eps in (0.0, 1e-17, -1e-17, 1e-15, -1e-15, 1e-12, -1e-12, 1e-9, -1e-9, 3e-9, 1e-8, -1e-8, 1.1e-8, 1e-6):
        add(seg_m, seg_b + eps)                                    # collinear with / just off every chord
        jj = rng.integers(0, n - 1, 200)
        tm = 0.5 * (seg_m[jj] + seg_m[np.minimum(jj + 1, n - 2)])  # tangent-like: slope between two chords
        add(tm, z[jj + 1] - tm * x[jj + 1] + eps)
    for j0 in (0, 1, n - 2, n - 1):                                # hugging the first / last samples
        for eps in (1e-9, -1e-9, 5e-9, -5e-9, 2e-8, -2e-8, 0.0):
            mm = np.concatenate([rng.uniform(-5, 5, 40), seg_m[[0, -1]] * (1 + rng.uniform(-1e-3, 1e-3, 2))])
            add(mm, z[j0] - mm * x[j0] + eps)
    m = 10.0 ** rng.uniform(-6, 13, 800) * rng.choice([-1, 1], 800)  # steep / shallow, crossing anywhere
    xc = rng.uniform(x[0] * 1.3, x[-1] * 1.3, 800)
    add(m, rng.uniform(z.min() - 0.02, z.max() + 0.02, 800) - m * xc)
    add(rng.uniform(-2, 2, 800), rng.uniform(-0.1, 0.3, 800))       # generic: 0, 1 or 2 crossings
    add([np.nan, 1.0, np.inf, -np.inf, 1.0, 0.0, 1e200, 1e160, -1e155, 0.0],
        [0.0, np.nan, 0.0, 1.0, np.inf, z[5], 0.0, -1e160, 1e150, -np.inf])
    return np.concatenate(ms), np.concatenate(bs)


@pytest.mark.parametrize("n_alpha", [905, 64, 1810])
def test_search_matches_numpy_rule_on_adversarial_lines(engine, n_alpha):
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(n_alpha), rs.focus_line(n_alpha)
    engine.setup(alpha, zf, x_a[32], z_a[32], x_a)
    assert engine.lens_shape == 1
    x, z = engine.lens()
    rng = np.random.default_rng(n_alpha)
    m, b = adversarial_lines(x, z, rng)
    want = numpy_rule(x, z, m, b)
    assert (want >= 0).sum() > 1000 and (want == -1).sum() > 100 and (want < -1).sum() > 20
    pts = np.array([numpy_intersection(x, z, np.float64(mi), np.float64(bi)) for mi, bi in zip(m, b)])
    assert np.isnan(pts[:, 0]).sum() > 100 and (~np.isnan(pts[:, 0])).sum() > 1000
    for mode in ("guided", "scan"):
        engine.set_search(mode)
        got, xi, zi = engine.debug_search(m, b, with_points=True)
        bad = np.nonzero(got != want)[0]
        assert len(bad) == 0, (mode, bad[:10], got[bad[:10]], want[bad[:10]], m[bad[:10]], b[bad[:10]])
        # the intersection point of main_rt.py:106-168 (collinear / parallel / out-of-bounds branches
        # included), bit for bit
        assert np.array_equal(xi, pts[:, 0], equal_nan=True), mode
        assert np.array_equal(zi, pts[:, 1], equal_nan=True), mode
    engine.set_search("guided")
