"""NumPy restatement of the reference's hot path ("oracle-A" of SURVEY.md section 7).  TEST INFRASTRUCTURE ONLY.

The reference (`/root/reference/main_rt.py`) is NumPy code that loops in Python over rays (`:384`, one
`find_line_curve_intersection` + `np.polyfit` per ray) and over (receiver, ray) pairs (`:487-501`, one scalar
`np.isclose` each).  This module evaluates the SAME NumPy expressions, in the same operation order, vectorised over
the K rays of a shot: the return-path search becomes one K x N matrix of `diffs` (`:78-79`), the table assembly one
K x E `np.isclose`.  One deliberate deviation, the same as oracle/tofrt_oracle.c: the reflected line is
used as computed, `(a_l, b_l)` of `:375-376`, instead of being re-fitted by `np.polyfit` over N synthetic points
(LAPACK rounding noise).  Rays that miss the pipe propagate NaN where the reference crashes.

It exists to time "the reference's NumPy CPU path" on the GPU box's host cores (bench.py `cpu_baseline.numpy`): the
verbatim reference cannot travel there (and is 400 x slower: 36 ray-solves/s per core against ~1.5e4 here).  Pinned by
tests/test_oracle_numpy.py against the reference's fixtures and the live-reference goldens.  Only tests/ and bench.py's
cpu_baseline leg import it; the product never does.
"""
import numpy as np


class Setup:
    """The module globals of main_rt.py:449-455 plus the per-iteration arrays of :469-479."""

    def __init__(self, x_a, z_a, zf, alpha, shooter, c1=6400.0, c2=1483.0, l0=0.12156646438729327,
                 h0=0.08843353561270673):
        self.c1, self.c2 = np.float64(c1), np.float64(c2)
        self.l0, self.h0 = np.float64(l0), np.float64(h0)
        self.d = self.l0 + self.h0
        self.x_a, self.z_a, self.zf, self.alpha = (np.asarray(v, dtype=np.float64) for v in (x_a, z_a, zf, alpha))
        self.shooter = int(shooter)
        # geometry-independent part of shoot_rays (main_rt.py:338-349), evaluated once
        self.x_p, self.z_p = self.x_z_from_alpha(self.alpha)
        xa0, za0 = self.x_a[self.shooter], self.z_a[self.shooter]
        phi_ap = np.arctan2(za0 - self.z_p, xa0 - self.x_p)
        d_zh, d_xh = self.dz_dx_from_alpha(self.alpha)
        self.phi_pq = self.refraction(phi_ap, np.arctan2(d_zh, d_xh), self.c1, self.c2)
        self.a_pq = np.tan(self.phi_pq)
        self.b_pq = self.z_p - self.a_pq * self.x_p
        self.leg1 = np.sqrt((xa0 - self.x_p) ** 2 + (za0 - self.z_p) ** 2) / self.c1

    # ---- main_rt.py:171-234 ----
    @staticmethod
    def roots_bhaskara(a, b, c):
        sqdelta = np.sqrt(np.square(b) - 4 * a * c)
        return (-b + sqdelta) / (2 * a), (-b - sqdelta) / (2 * a)

    def h_from_alpha(self, alpha):
        T = self.l0 / self.c1 + self.h0 / self.c2
        A = np.square(self.c1) / np.square(self.c2) - 1
        B = 2 * self.d * np.cos(alpha) - 2 * T * np.square(self.c1) / self.c2
        C = np.square(self.c1) * np.square(T) - np.square(self.d)
        return self.roots_bhaskara(A, B, C)[1]

    def dh_from_alpha(self, alpha):
        T = self.l0 / self.c1 + self.h0 / self.c2
        A = np.square(self.c1) / np.square(self.c2) - 1
        C = np.square(self.c1) * np.square(T) - np.square(self.d)
        phi_1 = -1 / (2 * A)
        phi_2 = (-2 * T * np.square(self.c1)) / self.c2
        phi_3 = 2 * self.d
        B = phi_2 + phi_3 * np.cos(alpha)
        dB = -phi_3 * np.sin(alpha)
        dB2 = 2 * B * dB
        sq = np.sqrt(np.square(B) - 4 * A * C)
        return phi_1 * (dB + (1 / (2 * sq)) * dB2)

    def x_z_from_alpha(self, alpha):
        h = self.h_from_alpha(alpha)
        return h * np.sin(alpha), h * np.cos(alpha)

    def dz_dx_from_alpha(self, alpha):
        h, dh = self.h_from_alpha(alpha), self.dh_from_alpha(alpha)
        return dh * np.cos(alpha) - h * np.sin(alpha), dh * np.sin(alpha) + h * np.cos(alpha)

    # ---- main_rt.py:267-292 (phi_slope supplied by the caller) ----
    @staticmethod
    def refraction(phi_in, phi_slope, v1, v2):
        theta_1 = phi_in - (phi_slope + np.pi / 2)
        with np.errstate(invalid="ignore"):
            theta_2 = np.arcsin((v2 / v1) * np.sin(theta_1))
        return phi_slope - np.pi / 2 + theta_2

    @staticmethod
    def reflection(phi_in, phi_slope):
        theta_1 = phi_in - (phi_slope + np.pi / 2)
        return phi_slope - np.pi / 2 + (-theta_1)


def shoot_rays(S, r_outer, pipe_offset):
    """main_rt.py:351-405 for all K rays at once -> dict(pipe_x, pipe_z, lens_2_x, lens_2_z, target_x, idx)."""
    r_outer, pipe_offset = np.float64(r_outer), np.float64(pipe_offset)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        # :351-364 pipe intersection, upper root
        A = np.square(S.a_pq) + 1
        B = 2 * (S.a_pq * S.b_pq - pipe_offset)
        C = np.square(pipe_offset) + np.square(S.b_pq) - np.square(r_outer)
        x_q1, x_q2 = Setup.roots_bhaskara(A, B, C)
        z_q1, z_q2 = S.a_pq * x_q1 + S.b_pq, S.a_pq * x_q2 + S.b_pq
        up = z_q1 > z_q2
        x_q, z_q = np.where(up, x_q1, x_q2), np.where(up, z_q1, z_q2)
        # :367-376 reflection (dzdx_pipe ignores the offset, :237-238)
        slope = -x_q / np.sqrt(np.square(r_outer) - np.square(x_q))
        phi_l = Setup.reflection(S.phi_pq, np.arctan(slope))
        a_l = np.tan(phi_l)
        b_l = z_q - a_l * x_q
        # :78-99 first sign change, all rays at once (K x N)
        diffs = S.z_p[None, :] - (a_l[:, None] * S.x_p[None, :] + b_l[:, None])
        sg = np.sign(diffs)
        chg = np.diff(sg, axis=1) != 0                      # NaN - x is NaN, and NaN != 0 is True (as in the reference)
        has = chg.any(axis=1)
        idx = np.where(has, chg.argmax(axis=1), -1)
        K, N = diffs.shape
        ix, iz = np.full(K, np.nan), np.full(K, np.nan)
        # :84-96 no crossing: first sample within 1e-8 of the line
        none = ~has
        if none.any():
            close = np.isclose(diffs[none], 0.0, atol=1e-8)
            anyc = close.any(axis=1)
            j = close.argmax(axis=1)
            rows = np.nonzero(none)[0]
            ix[rows[anyc]], iz[rows[anyc]] = S.x_p[j[anyc]], S.z_p[j[anyc]]
            idx[rows[anyc]] = -2 - j[anyc]
        # :106-168 chord intersection for the rays with a crossing
        r = np.nonzero(has)[0]
        i0 = idx[r]
        x1, y1, x2, y2 = S.x_p[i0], S.z_p[i0], S.x_p[i0 + 1], S.z_p[i0 + 1]
        m, b = a_l[r], b_l[r]
        m_seg = (y2 - y1) / (x2 - x1)
        b_seg = y1 - m_seg * x1
        x_int = (b_seg - b) / (m - m_seg)
        y_int = m * x_int + b
        ok = ((x_int >= np.minimum(x1, x2) - 1e-9) & (x_int <= np.maximum(x1, x2) + 1e-9) &
              (y_int >= np.minimum(y1, y2) - 1e-9) & (y_int <= np.maximum(y1, y2) + 1e-9))
        vert = np.isclose(x1, x2)                            # :110-123 (never on the reference's lens; kept for completeness)
        par = ~vert & np.isclose(m, m_seg)                   # :131-144
        xi = np.where(ok, x_int, np.nan)
        yi = np.where(ok, y_int, np.nan)
        if par.any():
            same = par & np.isclose(b, b_seg)
            xm = (x1 + x2) / 2.0
            xi = np.where(par, np.where(same, xm, np.nan), xi)
            yi = np.where(par, np.where(same, m * xm + b, np.nan), yi)
        if vert.any():
            yv = m * x1 + b
            okv = (yv >= np.minimum(y1, y2) - 1e-9) & (yv <= np.maximum(y1, y2) + 1e-9)
            xi = np.where(vert, np.where(okv, x1, np.nan), xi)
            yi = np.where(vert, np.where(okv, yv, np.nan), yi)
        ix[r], iz[r] = xi, yi
        # :396-405 refraction back into the lens, landing point
        alpha_i = np.arctan2(ix, iz)
        d_zi, d_xi = S.dz_dx_from_alpha(alpha_i)
        phi_last = Setup.refraction(phi_l, np.arctan2(d_zi, d_xi), S.c2, S.c1)
        a_i = np.tan(phi_last)
        b_i = iz - a_i * ix
        target_x = (S.zf - b_i) / a_i
    return {"pipe_x": x_q, "pipe_z": z_q, "lens_2_x": ix, "lens_2_z": iz, "target_x": target_x, "idx": idx,
            "a_l": a_l, "b_l": b_l}


def table_row(S, res, rtol=1e-5, atol=1e-6):
    """main_rt.py:487-501 for one geometry -> (k_hit[E], tof[E]): first matching ray, legs summed left to right."""
    tx = res["target_x"]
    close = np.isclose(tx[:, None], S.x_a[None, :], rtol=rtol, atol=atol)        # K x E
    hit = close.any(axis=0)
    k = np.where(hit, close.argmax(axis=0), -1)
    kk = np.where(hit, k, 0)
    with np.errstate(invalid="ignore"):
        dist = lambda xa, za, xb, zb: np.sqrt((xa - xb) ** 2 + (za - zb) ** 2)
        t = S.leg1[kk]
        t = t + dist(S.x_p[kk], S.z_p[kk], res["pipe_x"][kk], res["pipe_z"][kk]) / S.c2
        t = t + dist(res["pipe_x"][kk], res["pipe_z"][kk], res["lens_2_x"][kk], res["lens_2_z"][kk]) / S.c2
        t = t + dist(res["lens_2_x"][kk], res["lens_2_z"][kk], tx[kk], S.zf[kk]) / S.c1
    return k.astype(np.int32), np.where(hit, t, 0.0)


def table(S, radii, offsets, rtol=1e-5, atol=1e-6):
    """The sweep main_rt.py:464-501 -> (tof[G][E], k_hit[G][E])."""
    G, E = len(radii), len(S.x_a)
    tof, k_hit = np.zeros((G, E)), np.full((G, E), -1, dtype=np.int32)
    for g in range(G):
        k_hit[g], tof[g] = table_row(S, shoot_rays(S, radii[g], offsets[g]), rtol, atol)
    return tof, k_hit


def _worker(args):
    x_a, z_a, zf, alpha, shooter, radii, offsets = args
    S = Setup(x_a, z_a, zf, alpha, shooter)
    import time
    t0 = time.perf_counter()
    table(S, radii, offsets)
    return time.perf_counter() - t0


def timed_rate(x_a, z_a, zf, alpha, shooter, radii, offsets, processes):
    """ray-solves/s of `processes` worker processes, each tracing its own slice of the geometries (bench.py)."""
    import multiprocessing as mp
    import time
    parts = [(x_a, z_a, zf, alpha, shooter, radii[i::processes], offsets[i::processes]) for i in range(processes)]
    ctx = mp.get_context("spawn")   # the caller may hold a CUDA context: never fork it
    with ctx.Pool(processes) as pool:
        warm = [(x_a, z_a, zf, alpha, shooter, radii[:1], offsets[:1])] * processes
        pool.map_async(_worker, warm).get(timeout=120)                                              # warm the workers
        t0 = time.perf_counter()
        pool.map_async(_worker, parts).get(timeout=300)        # (a timeout, so a stuck worker cannot hang the caller)
        dt = time.perf_counter() - t0
    return len(x_a) * len(radii) / dt, dt
