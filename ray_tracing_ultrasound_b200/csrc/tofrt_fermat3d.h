/*
 * tofrt_fermat3d.h -- per-pair arithmetic of the 3-D two-point Fermat solve
 * (SURVEY.md row f4 / BASELINE config 5: matrix array, 3-D volume, curved surface).
 *
 * The reference is strictly 2-D, so nothing in /root/reference corresponds to this file:
 * PARITY UNPINNED.  What is reused from the reference is the lens profile h(alpha)
 * (h_from_alpha, main_rt.py:180-189): the 3-D lens is its surface of revolution about the
 * z axis.  Because h depends on alpha only through cos(alpha) (main_rt.py:184), the surface
 * is P = f(n_z) n for a unit direction n, with f(c) = h(arccos c).  Directions are
 * parametrised by two angles that are regular on the whole lens cap (poles on the y axis):
 *
 *     n(a, b) = (sin a cos b, sin b, cos a cos b)
 *
 * so b = 0 is the reference's x-z plane with a = alpha, where the surface point has the bits of the
 * 2-D lens_point_t of tofrt_common.cuh.  For an element A on the array plane and a focal point F
 * in the water,
 *
 *     t(a, b) = |A - P(a, b)| / c1 + |P(a, b) - F| / c2
 *
 * is minimised by a projected Newton iteration with analytic gradient and Hessian and
 * backtracking, started from the best sample of a coarse (a, b) grid and confined to the
 * aperture [lo, hi]^2.
 *
 * Like tofrt_math.h this header is shared bit for bit between nvcc (the product) and gcc
 * (the oracle's device-math twin): IEEE operations only, no contraction except explicit fma().
 */
#ifndef TOFRT_FERMAT3D_H
#define TOFRT_FERMAT3D_H

#include "tofrt_math.h"

typedef struct {
    double inv_c1, inv_c2;            /* 1/c1, 1/c2 */
    double twod, K2, fourAC, twoA;    /* lens constants of main_rt.py:181-185 (as in DevSetup) */
    double h0;                        /* flat-interface extension: plane z = h0 */
    int surface;                      /* 0 lens, 1 flat */
    double lo, hi;                    /* aperture: a, b in [lo, hi] */
} tofrt_surf3d;

/* radial function f(c) = h(arccos c) and its first two derivatives */
TOFRT_HD void tofrt_f3d_radial(const tofrt_surf3d *S, double c, double *f, double *f1, double *f2)
{
    if (S->surface == 0) {
        double B = S->twod * c - S->K2;               /* main_rt.py:184 */
        double sd = sqrt(B * B - S->fourAC);
        *f = (-B - sd) / S->twoA;                     /* main_rt.py:187-189, root [1] */
        double k = S->twod / S->twoA;
        *f1 = -k * (1.0 + B / sd);
        *f2 = (k * S->twod * S->fourAC) / (sd * sd * sd);
    } else {
        double v = S->h0 / c;
        double v1 = -v / c;
        *f = v; *f1 = v1; *f2 = -2.0 * v1 / c;
    }
}

/* surface point only (grid samples) */
TOFRT_HD void tofrt_f3d_point(const tofrt_surf3d *S, double a, double b, double P[3])
{
    double sa, ca, sb, cb, f, f1, f2;
    tofrt_sincos(a, &sa, &ca);
    tofrt_sincos(b, &sb, &cb);
    double nx = sa * cb, ny = sb, nz = ca * cb;
    tofrt_f3d_radial(S, nz, &f, &f1, &f2);
    P[0] = f * nx; P[1] = f * ny; P[2] = f * nz;
}

/* a*b + c*d and a*b + c*d + e*f with explicit fused steps (no reference operation order to keep
 * here -- the reference has no 3-D code -- so the products are fused: half the FP64 instructions) */
TOFRT_HD double tofrt_f3d_dot2(double a, double b, double c, double d) { return fma(a, b, c * d); }
TOFRT_HD double tofrt_f3d_dot3(double a, double b, double c, double d, double e, double f)
{
    return fma(a, b, fma(c, d, e * f));
}

/* travel time of a given surface point */
TOFRT_HD double tofrt_f3d_time(const tofrt_surf3d *S, const double P[3], const double A[3], const double F[3])
{
    double ux = P[0] - A[0], uy = P[1] - A[1], uz = P[2] - A[2];
    double vx = P[0] - F[0], vy = P[1] - F[1], vz = P[2] - F[2];
    double r1 = sqrt(tofrt_f3d_dot3(ux, ux, uz, uz, uy, uy)), r2 = sqrt(tofrt_f3d_dot3(vx, vx, vz, vz, vy, vy));
    return fma(r1, S->inv_c1, r2 * S->inv_c2);
}

/* t, gradient g = (t_a, t_b) and Hessian H = (t_aa, t_ab, t_bb) at (a, b).
 * With P = f(n_z) n and the derivatives of the unit vector
 *     n_a = (n_z, 0, -n_x)   n_b = (-sa sb, cb, -ca sb)   n_ab = (-ca sb, 0, sa sb)
 *     n_aa = (-n_x, 0, -n_z) n_bb = -n                    (n_z)_a = -n_x, (n_z)_b = -ca sb
 * every product u . P_x (u = P - A or P - F) is a combination of the five scalars u.n, u.n_a, u.n_b,
 * u.n_ab, u.n_aa with the derivatives of f, and n.n_a = n.n_b = n_a.n_b = 0, |n_a| = cb, |n_b| = 1 give
 * |P_a|^2 = f_a^2 + f^2 cb^2, P_a.P_b = f_a f_b, |P_b|^2 = f_b^2 + f^2: no derivative vector is formed. */
TOFRT_HD void tofrt_f3d_eval(const tofrt_surf3d *S, double a, double b, const double A[3], const double F[3],
                             double *t, double g[2], double H[3])
{
    double sa, ca, sb, cb, f, f1, f2;
    tofrt_sincos(a, &sa, &ca);
    tofrt_sincos(b, &sb, &cb);
    const double nx = sa * cb, ny = sb, nz = ca * cb;
    const double sasb = sa * sb, casb = ca * sb;
    tofrt_f3d_radial(S, nz, &f, &f1, &f2);
    const double Px = f * nx, Py = f * ny, Pz = f * nz;
    const double ka = -(f1 * nx), kb = -(f1 * casb);       /* f_a, f_b */
    const double f1nz = f1 * nz, f2nx = f2 * nx;
    const double kaa = fma(f2nx, nx, -f1nz);               /* f_aa */
    const double kab = fma(f2nx, casb, f1 * sasb);         /* f_ab */
    const double kbb = fma(f2 * casb, casb, -f1nz);        /* f_bb */
    const double ff = f * f;
    const double paa = fma(ka, ka, ff * (cb * cb)), pab = ka * kb, pbb = fma(kb, kb, ff);
    double tt = 0.0, ga = 0.0, gb = 0.0, haa = 0.0, hab = 0.0, hbb = 0.0;
    for (int leg = 0; leg < 2; ++leg) {
        const double *X = leg ? F : A;
        const double ic = leg ? S->inv_c2 : S->inv_c1;
        const double ux = Px - X[0], uy = Py - X[1], uz = Pz - X[2];
        const double r = sqrt(tofrt_f3d_dot3(ux, ux, uz, uz, uy, uy));
        const double ir = 1.0 / r;
        const double w = ir * ic, w3 = w * ir * ir;
        const double dxz = tofrt_f3d_dot2(ux, nx, uz, nz);           /* -u.n_aa */
        const double dn = fma(uy, ny, dxz);                          /* u.n */
        const double dna = fma(ux, nz, -(uz * nx));                  /* u.n_a */
        const double dnb = fma(uy, cb, -tofrt_f3d_dot2(ux, sasb, uz, casb));  /* u.n_b */
        const double dnab = fma(uz, sasb, -(ux * casb));             /* u.n_ab */
        const double ua = tofrt_f3d_dot2(ka, dn, f, dna);            /* u.P_a */
        const double ub = tofrt_f3d_dot2(kb, dn, f, dnb);            /* u.P_b */
        const double qaa = fma(kaa, dn, fma(2.0 * ka, dna, -(f * dxz)));
        const double qab = fma(kab, dn, fma(ka, dnb, fma(kb, dna, f * dnab)));
        const double qbb = fma(kbb, dn, fma(2.0 * kb, dnb, -(f * dn)));
        tt = fma(r, ic, tt);
        ga = fma(ua, w, ga);
        gb = fma(ub, w, gb);
        const double uaw3 = ua * w3;
        haa += fma(paa + qaa, w, -(ua * uaw3));
        hab += fma(pab + qab, w, -(ub * uaw3));
        hbb += fma(pbb + qbb, w, -(ub * ub * w3));
    }
    *t = tt; g[0] = ga; g[1] = gb; H[0] = haa; H[1] = hab; H[2] = hbb;
}

TOFRT_HD double tofrt_f3d_clamp(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* One step proposal from the state (a, b, g, H): projected Newton on the free variables
 * (a variable sitting on a bound with the gradient pushing outwards is held), falling back to a
 * bounded steepest-descent step where the reduced Hessian is not positive definite. */
TOFRT_HD void tofrt_f3d_step(const tofrt_surf3d *S, double a, double b, const double g[2], const double H[3],
                             double cell, double *da, double *db)
{
    const int free_a = !((a <= S->lo && g[0] >= 0.0) || (a >= S->hi && g[0] <= 0.0));
    const int free_b = !((b <= S->lo && g[1] >= 0.0) || (b >= S->hi && g[1] <= 0.0));
    double sa_ = 0.0, sb_ = 0.0;
    if (free_a && free_b) {
        const double det = H[0] * H[2] - H[1] * H[1];
        if (H[0] > 0.0 && det > 0.0) {
            sa_ = -(H[2] * g[0] - H[1] * g[1]) / det;
            sb_ = -(H[0] * g[1] - H[1] * g[0]) / det;
        } else {
            const double gn = sqrt(g[0] * g[0] + g[1] * g[1]);
            sa_ = gn > 0.0 ? -g[0] / gn * cell : 0.0;
            sb_ = gn > 0.0 ? -g[1] / gn * cell : 0.0;
        }
    } else if (free_a) {
        sa_ = H[0] > 0.0 ? -g[0] / H[0] : (g[0] > 0.0 ? -cell : (g[0] < 0.0 ? cell : 0.0));
    } else if (free_b) {
        sb_ = H[2] > 0.0 ? -g[1] / H[2] : (g[1] > 0.0 ? -cell : (g[1] < 0.0 ? cell : 0.0));
    }
    *da = sa_; *db = sb_;
}

#define TOFRT_F3D_MAX_EVALS 24
#define TOFRT_F3D_STEP_TOL 1e-6 /* rad: below this the quadratic model is exact to ~1e-22 s */

/* Refinement from (a0, b0): evaluate the trial point, keep it if the time did not increase
 * (the start is always kept), otherwise shrink the step; propose the next trial from the kept
 * state.  One evaluation site, so the GPU code stays small and a warp's lanes stay in step.
 * Returns the number of evaluations used. */
TOFRT_HD int tofrt_f3d_refine(const tofrt_surf3d *S, double a0, double b0, double cell, const double A[3],
                              const double F[3], double *t_out, double *a_out, double *b_out)
{
    double a = a0, b = b0, t = 0.0, g[2] = {0.0, 0.0}, H[3] = {0.0, 0.0, 0.0};
    double na = a0, nb = b0, shrink = 1.0;
    int evals = 0;
    for (;;) {
        double tn, gn[2], Hn[3];
        tofrt_f3d_eval(S, na, nb, A, F, &tn, gn, Hn);
        ++evals;
        if (evals == 1 || tn <= t) {
            a = na; b = nb; t = tn; g[0] = gn[0]; g[1] = gn[1]; H[0] = Hn[0]; H[1] = Hn[1]; H[2] = Hn[2];
            shrink = 1.0;
        } else {
            shrink *= 0.25;
        }
        double da, db;
        tofrt_f3d_step(S, a, b, g, H, cell, &da, &db);
        na = tofrt_f3d_clamp(a + da * shrink, S->lo, S->hi);
        nb = tofrt_f3d_clamp(b + db * shrink, S->lo, S->hi);
        da = na - a; db = nb - b;
        if (!(fabs(da) > TOFRT_F3D_STEP_TOL || fabs(db) > TOFRT_F3D_STEP_TOL)) {
            /* converged (or NaN): the last step is taken through the quadratic model */
            t = t + (g[0] * da + g[1] * db) + 0.5 * (H[0] * da * da + 2.0 * H[1] * da * db + H[2] * db * db);
            a = na; b = nb;
            break;
        }
        if (evals >= TOFRT_F3D_MAX_EVALS) break;
    }
    *t_out = t; *a_out = a; *b_out = b;
    return evals;
}

#endif /* TOFRT_FERMAT3D_H */
