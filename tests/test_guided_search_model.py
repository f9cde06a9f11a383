"""CPU: an executable model of the guided search of docs/guided_search.md (the bracket invariant, the fused probe,
the post-loop certainty test with the SAME probe, the predicted bracket of section 4b with its warp-wide
all-lanes conditions), checked against the oracle's literal first-sign-change index on the reference's fixture
geometries -- including the degenerate ones whose |f| sit right at `tol`, which is where a probe/diff mix-up in
the post-loop test shows (the model can be switched to that wrong variant and must then disagree).

This models the ALGORITHM the kernel implements (csrc/tofrt_table_kernel.cuh, block "guided search"); the kernel
itself is checked against the oracle on the GPU (tests/test_gpu_parity.py, tests/test_gpu_search.py)."""
import numpy as np
import pytest

from ray_tracing_ultrasound_b200 import reference_setup as rs

W, PRED_STEPS = 4, 3


def _fma(a, b, c):  # one rounding (80-bit intermediate: exact enough for a double fma here)
    return float(np.float64(np.longdouble(a) * np.longdouble(b) + np.longdouble(c)))


def _cls(v):
    return 2 if v > 0 else (0 if v < 0 else (1 if v == 0 else 3))


def guided_indices(lx, lz, m_all, b_all, same_probe_after_loop=True, predict=True):
    """-> idx per ray (-1: no sign change), None where the ray is not eligible (the kernel scans those)."""
    N = len(lx)
    seg = np.append((lz[1:] - lz[:-1]) / (lx[1:] - lx[:-1]), 0.0)
    X, Z, sh = np.abs(lx).max(), np.abs(lz).max(), 1.0
    log2n = int(np.ceil(np.log2(N)))
    out = [None] * len(m_all)
    used_prediction = 0
    h1 = np.full(32, -1)
    h2 = np.full(32, -1)
    for c0 in range(0, len(m_all), 32):
        lanes = range(c0, min(c0 + 32, len(m_all)))
        st = {}
        for k in lanes:
            m, b = m_all[k], b_all[k]
            if not (abs(m) < 1e150 and abs(b) < 1e150):
                continue
            diff = lambda j, m=m, b=b: lz[j] - (m * lx[j] + b)
            tol = (abs(m) * X + abs(b) + Z) * 2.0 ** -48 + 1e-290
            g0, gN = sh * diff(0), sh * diff(N - 1)
            pos0, neg0 = g0 > tol, g0 < -tol
            if pos0 and gN > tol:
                out[k] = -1
                continue
            if not (pos0 or neg0):
                continue
            sgn = sh if pos0 else -sh
            ms_eff = -np.inf if pos0 else sh * m
            cm, cbt = sgn * m, sgn * b + tol
            probe = lambda j, cm=cm, cbt=cbt, sgn=sgn: (sgn * lz[j]) > _fma(cm, lx[j], cbt)
            cond = lambda j, probe=probe, ms_eff=ms_eff: probe(j) and (sh * seg[j] > ms_eff)
            st[k] = dict(diff=diff, tol=tol, neg0=neg0, probe=probe, cond=cond, lo=0, hi=N - 1)
        # predicted bracket: only if EVERY bracketing lane has a history and both of its probes certify it
        nsteps = log2n
        if predict and st and all(h1[k - c0] >= 0 and h2[k - c0] >= 0 for k in st):
            cand = {}
            for k, s in st.items():
                p = 2 * h1[k - c0] - h2[k - c0]
                plo = min(max(p - W, 0), N - 2)
                phi = min(plo + 2 * W, N - 1)
                cand[k] = (plo, phi, s["cond"](plo) and not s["cond"](phi))
            if all(ok for _, _, ok in cand.values()):
                for k, (plo, phi, _) in cand.items():
                    st[k]["lo"], st[k]["hi"] = plo, phi
                nsteps = PRED_STEPS
                used_prediction += 1
        for k, s in st.items():
            lo, hi = s["lo"], s["hi"]
            for _ in range(nsteps):
                mid = (lo + hi) >> 1
                if hi - lo > 1:
                    if s["cond"](mid):
                        lo = mid
                    else:
                        hi = mid
            assert hi - lo == 1
            s["lo"], s["hi"] = lo, hi
        h2 = h1.copy()
        h1 = np.full(32, -1)
        for k, s in st.items():
            h1[k - c0] = s["lo"]
            lo, hi, diff = s["lo"], s["hi"], s["diff"]
            bracket = True
            if s["neg0"]:
                certainly_negative = s["probe"](hi) if same_probe_after_loop else (sh * diff(hi) < -s["tol"])
                if certainly_negative:
                    bracket = False
            if bracket:
                cp, j = _cls(diff(lo)), lo
                while j + 1 < N:
                    cn = _cls(diff(j + 1))
                    if cp == 3 or cn == 3 or cp != cn:
                        break
                    cp, j = cn, j + 1
                out[k] = j if j + 1 < N else -1
            else:
                out[k] = -1
    return out, used_prediction


def _case(oracle_dm, r, off):
    x_a, z_a = rs.transducer()
    alpha, zf = rs.launch_angles(), rs.focus_line()
    res = oracle_dm.shoot_rays(oracle_dm.params(), r, off, x_a[32], z_a[32], zf, alpha)
    ref = np.where(res["idx"] >= 0, res["idx"], -1)    # on-sample encodings (-2 - j) are "no change"
    return res["lens_1_x"], res["lens_1_z"], res["a_l"], res["b_l"], ref


@pytest.mark.parametrize("r,off", [(0.01, 0.0), (0.02, 0.0), (0.01, 0.01), (0.05, 0.003), (0.1, -0.01)])
def test_guided_model_returns_the_literal_index(oracle_dm, r, off):
    lx, lz, m, b, ref = _case(oracle_dm, r, off)
    for predict in (False, True):
        idx, used = guided_indices(lx, lz, m, b, predict=predict)
        elig = [k for k, v in enumerate(idx) if v is not None]
        assert len(elig) >= 850
        assert all(idx[k] == ref[k] for k in elig)
        if predict and (r, off) != (0.01, 0.01):
            assert used >= 24    # the predicted bracket is the common path: 26 of the 27 chunks that have a history


def test_post_loop_test_must_use_the_loops_probe(oracle_dm):
    """The degenerate fixture geometry r = 0.01, offset = 0: certifying the far end with the computed diff instead
    of the loop's fused probe declares 'no crossing' for rays whose |f_hi| is within a few percent of tol."""
    lx, lz, m, b, ref = _case(oracle_dm, 0.01, 0.0)
    wrong, _ = guided_indices(lx, lz, m, b, same_probe_after_loop=False, predict=False)
    bad = [k for k, v in enumerate(wrong) if v is not None and v != ref[k]]
    assert 1 <= len(bad) <= 20 and all(wrong[k] == -1 for k in bad)
