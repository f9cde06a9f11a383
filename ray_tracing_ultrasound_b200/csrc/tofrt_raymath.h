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

/* sin(theta_1) of the refraction back into the lens for an incoming direction (cl, sl)/rsq and the
 * surface tangent (dx, dz): theta_1 = phi_l - (atan2(dz, dx) + pi/2)  =>  sin(theta_1) = -cos(phi_l - A). */
TOFRT_HD double tofrt_vec_sin_theta1(double cl, double sl, double rsq, double dz, double dx)
{
    double hyp = sqrt(fma(dx, dx, dz * dz));
    return -fma(cl, dx, sl * dz) / (hyp * rsq);
}

/* tan(A - pi/2 + theta_2) with sin(theta_2) = sa, cos(theta_2) = sqrt(1 - sa^2) >= 0 (principal asin; NaN for
 * |sa| > 1) and A = atan2(dz, dx):  (dz sa - dx c) / (dz c + dx sa)  -- the tangent's length cancels. */
TOFRT_HD double tofrt_vec_refract_slope(double sa, double dz, double dx)
{
    double c = sqrt((1.0 - sa) * (1.0 + sa));
    return fma(dz, sa, -(dx * c)) / fma(dz, c, dx * sa);
}

#endif /* TOFRT_RAYMATH_H */
