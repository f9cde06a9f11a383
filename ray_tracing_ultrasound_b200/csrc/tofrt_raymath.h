/*
 * tofrt_raymath.h -- the two interface laws of the return path in VECTOR form
 * (TOFRT_OPT_MATH = TOFRT_MATH_VECTOR).
 *
 * The reference keeps every direction as an angle: reflection (main_rt.py:283-292 with the pipe
 * slope of :237-238) returns phi_l = 2 atan(slope) - phi_pq and the line slope tan(phi_l)
 * (:367-375); the refraction back into the lens (:267-280, :396-401) forms
 * theta_1 = phi_l - (atan2(dz, dx) + pi/2), theta_2 = asin((c1/c2) sin(theta_1)) and
 * tan(atan2(dz, dx) - pi/2 + theta_2).  Every one of these is a rational function of the
 * direction COSINES -- the angles are bookkeeping -- so the same quantities follow from
 * (cos, sin) pairs by the addition theorems, with two square roots and three divisions instead of
 * two atan2, one atan, one asin, one sin and two tan per ray, and with fewer roundings.  The branch
 * conventions are the reference's: atan in (-pi/2, pi/2] for the pipe slope, principal asin
 * (cos(theta_2) >= 0), NaN on total internal reflection and where sqrt(r^2 - x_q^2) is not real.
 *
 * Range: the reflected direction is carried scaled by r^2, so radii beyond ~1e150 m overflow here where the
 * angle form would still return a finite slope (TOFRT_MATH_ANGLE remains available for such inputs).
 *
 * Shared bit for bit between nvcc (the kernel) and gcc (the oracle's device-math twin, which
 * switches to these functions with tofo_set_math(1)).  IEEE operations only; fused operations are
 * explicit fma() calls.
 */
#ifndef TOFRT_RAYMATH_H
#define TOFRT_RAYMATH_H

#include "tofrt_math.h"

/* Direction of the reflected ray phi_l = 2 phi_s - phi_pq, phi_s = atan(-x_q / sqrt(rsq - x_q^2))
 * (dzdx_pipe ignores the pipe offset, main_rt.py:237-238), from (cos, sin) of phi_pq; returned SCALED
 * by r^2: (cl, sl) = r^2 (cos phi_l, sin phi_l).
 * r^2 cos(2 phi_s) = s^2 - x_q^2, r^2 sin(2 phi_s) = -2 x_q s with s = sqrt(r^2 - x_q^2). */
TOFRT_HD void tofrt_vec_reflect(double x_q, double rsq, double cpq, double spq, double *cl, double *sl)
{
    double xx = x_q * x_q;
    double ss = rsq - xx;
    double s = sqrt(ss);                       /* NaN when |x_q| > r, like the reference's slope */
    double c2 = ss - xx;
    double s2 = (-2.0 * x_q) * s;
    *cl = fma(c2, cpq, s2 * spq);
    *sl = fma(s2, cpq, -(c2 * spq));
}

/* Constants of the lens surface h(alpha) (main_rt.py:181-212) as the vector form uses them; built by
 * tofrt_vec_lens_consts from the quantities tofrt_setup / the oracle derive in the reference's operation order. */
typedef struct {
    double twod;      /* 2 d                      (main_rt.py:184 / phi_3 of :201) */
    double K2;        /* 2 T c1^2 / c2            (:184 / -phi_2 of :200)          */
    double fourAC;    /* 4 A C                    (:172)                          */
    double inv_twoA;  /* 1 / (2 A)                                                */
    double pt;        /* -phi_1 * 2 d = d / A     (:199, :201)                    */
    double k12;       /* c1 / c2: v2 / v1 of the refraction back into the lens    */
    double h0;
    int flat;         /* extension: plane z = h0 instead of the lens              */
} tofrt_vec_lens;

TOFRT_HD void tofrt_vec_lens_consts(tofrt_vec_lens *L, double twod, double K2, double fourAC, double twoA, double phi1,
                                    double c1_over_c2, double h0, int flat)
{
    L->twod = twod; L->K2 = K2; L->fourAC = fourAC;
    L->inv_twoA = 1.0 / twoA;
    L->pt = -(phi1 * twod);
    L->k12 = c1_over_c2; L->h0 = h0; L->flat = flat;
}

/* Pipe intersection of main_rt.py:351-364 without the slope/intercept detour: the refracted ray leaves the lens
 * sample P = (xp, zp) along the unit vector u = (cpq, spq) = (cos, sin)(phi_pq); |P + t u - (off, 0)|^2 = r^2 is
 * t^2 + 2 hh t + cc = 0 with hh = u . (P - C), cc = |P - C|^2 - r^2.  The reference's "upper" root (z_q1 > z_q2,
 * :362) is the one with the larger t * spq: t = -hh + copysign(sqrt(hh^2 - cc), spq).  No division (the reference
 * divides twice by 2 (a_pq^2 + 1), with a_pq ~ 1e16 for the centre ray); NaN when the ray misses the pipe. */
TOFRT_HD void tofrt_vec_pipe(double xp, double zp, double cpq, double spq, double off, double rsq, double *x_q,
                             double *z_q)
{
    double wx = xp - off;
    double hh = fma(cpq, wx, spq * zp);
    double cc = fma(wx, wx, zp * zp) - rsq;
    double s = sqrt(fma(hh, hh, -cc));
    double t = copysign(s, spq) - hh;
    *x_q = fma(t, cpq, xp);
    *z_q = fma(t, spq, zp);
}

/* Refraction back into the lens and landing point (main_rt.py:396-405) for a return ray with direction
 * (cl, sl) = r^2 (cos, sin)(phi_l) that enters the lens at I = (ix, iz); rsq2 = (r^2)^2.
 *
 * The reference forms alpha_i = arctan2(ix, iz), h(alpha_i), dh/dalpha, the tangent (dz, dx), theta_1, theta_2 =
 * asin((c1/c2) sin theta_1) and tan(atan2(dz, dx) - pi/2 + theta_2).  Everything after alpha_i depends on the
 * tangent only through its DIRECTION, so it is carried scaled by the positive factor sd * rho^3 (rho = |I|,
 * sd = sqrt(B^2 - 4AC)), which removes every division of h, dh/dalpha and of the normalisations:
 *   rho B  = 2 d iz - K2 rho,      rho sd = sqrt((rho B)^2 - 4AC rho^2) =: sdr,       sum = sdr + rho B
 *   DH = dh sd rho^2 = (pt ix) sum (from dh = phi_1 dB (1 + B / sd)),     HH = h sd rho^2 = (-sum / 2A) sdr
 *   (Tz, Tx) = (DH iz - HH ix, DH ix + HH iz) = sd rho^3 (dz, dx)
 *   N = -(c1/c2) (cl Tx + sl Tz),  M^2 = |T|^2 r^4:   sin theta_2 = N / M,  cos theta_2 = sqrt(M^2 - N^2) / M =: W / M
 *   slope of the last leg = (Tz N - Tx W) / (Tz W + Tx N)      (M cancels; NaN on total internal reflection)
 *   target_x = ix + (z_f - iz) (Tz W + Tx N) / (Tz N - Tx W)   (= (z_f - b_i) / a_i with b_i = iz - a_i ix)
 * Three square roots and one division instead of four square roots and six divisions.
 * Returns M^2 - N^2 through *w2 (negative: total internal reflection, main_rt.py:277 -> NaN). */
TOFRT_HD void tofrt_vec_return_head(const tofrt_vec_lens *L, double ix, double iz, double cl, double sl, double rsq2,
                                    double *tz, double *tx, double *n, double *w2)
{
    double Tz, Tx;
    if (!L->flat) {
        double rho2 = fma(ix, ix, iz * iz);
        double rho = sqrt(rho2);
        double Br = fma(L->twod, iz, -(L->K2 * rho));
        double sdr = sqrt(fma(Br, Br, -(L->fourAC * rho2)));
        double sum = sdr + Br;
        double DH = (L->pt * ix) * sum;
        double HH = (-sum * L->inv_twoA) * sdr;
        Tz = fma(DH, iz, -(HH * ix));
        Tx = fma(DH, ix, HH * iz);
    } else { /* plane z = h0: tangent (dz, dx) = (0, +); NaN entry points must stay NaN */
        Tz = 0.0 * (ix + iz);
        Tx = 1.0 + Tz;
    }
    double N = -L->k12 * fma(cl, Tx, sl * Tz);
    double M2 = fma(Tx, Tx, Tz * Tz) * rsq2;
    *tz = Tz; *tx = Tx; *n = N;
    *w2 = fma(-N, N, M2);
}

TOFRT_HD double tofrt_vec_return_tail(double ix, double iz, double z_f, double tz, double tx, double n, double w2)
{
    double W = sqrt(w2);
    double den = fma(tz, W, tx * n);
    double num = fma(tz, n, -(tx * W));
    return fma(z_f - iz, den / num, ix);
}

#endif /* TOFRT_RAYMATH_H */
