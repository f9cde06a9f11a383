/*
 * tofrt_math.h -- portable FP64 elementary functions of the TOF engine.
 *
 * sin, cos, tan, atan, atan2, asin (and sincos-of-atan2) built only from IEEE-754 operations that
 * are correctly rounded on every target (+, -, *, /, sqrt, fma), so the very
 * same source gives bit-identical results when compiled
 *   - by nvcc for sm_100a (the product: csrc/tofrt_kernels.cu and its .cuh parts), and
 *   - by gcc for the host (the oracle's "device-math" twin, oracle/Makefile),
 * which is what makes GPU-vs-CPU parity of the discrete indices and of the
 * travel times exact instead of "within libm noise".  The reference's own
 * elementary functions are NumPy's SIMD loops (np.sin/np.cos/np.tan/
 * np.arctan/np.arctan2/np.arcsin, call sites main_rt.py:184-405); they differ
 * from these by <= ~2 ulp, i.e. <= 1e-19 s on a travel time.
 *
 * Written for the GPU first: straight-line code (selects, no data-dependent
 * branches, so 32 rays stay converged), one division per function, and the
 * polynomial coefficients live in constant memory on the device so that each
 * Horner step is a single DFMA with a c[bank][offset] operand.
 *
 * Accuracy (tests/test_math.py, against mpmath): sin, cos < 1.5 ulp; sincos-of-atan2 < 2.5 ulp; tan, atan,
 * atan2 < 2.5 ulp; asin < 3 ulp.  Domain of the trigonometric reduction:
 * |x| < 2^30 (angles in this engine never leave [-4 pi, 4 pi]); larger finite
 * arguments return NaN instead of a value with no correct digits.  sin/tan of
 * -0.0 return +0.0 (the sign of a zero is not preserved).
 *
 * Compile without floating-point contraction (nvcc -fmad=false,
 * gcc -ffp-contract=off): every fused operation below is an explicit fma().
 */
#ifndef TOFRT_MATH_H
#define TOFRT_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TOFRT_HD __device__ __forceinline__
#ifdef TOFRT_MATH_NOINLINE /* tuning knob: one shared copy of the two big functions */
#define TOFRT_HD_BIG __device__ __noinline__
#else
#define TOFRT_HD_BIG __device__ __forceinline__
#endif
#define TOFRT_TABLE static __constant__ double
#else
#define TOFRT_HD static inline
#define TOFRT_HD_BIG static inline
#define TOFRT_TABLE static const double
#endif

/* One table for every constant; indices below.  (gcc folds tofrt_k[IDX] into
 * the literal; nvcc turns it into a constant-bank operand.) */
enum {
    TK_PIO2_HI, TK_PIO2_MID, TK_PIO2_LO, TK_TWO_OVER_PI, TK_MAGIC, TK_PI_HI, TK_PI_LO, TK_PIO4_HI, TK_PIO4_LO,
    TK_S1, TK_S2, TK_S3, TK_S4, TK_S5, TK_S6,
    TK_C1, TK_C2, TK_C3, TK_C4, TK_C5, TK_C6,
    TK_A0, TK_A1, TK_A2, TK_A3, TK_A4, TK_A5, TK_A6, TK_A7, TK_A8, TK_A9, TK_A10,
    TK_COUNT
};
TOFRT_TABLE tofrt_k[TK_COUNT] = {
    1.5707963267948966,      /* PIO2_HI  0x1.921fb54442d18p+0  */
    6.123233995736766e-17,   /* PIO2_MID 0x1.1a62633145c07p-54 */
    -1.4973849048591698e-33, /* PIO2_LO                        */
    0.6366197723675814,      /* 2/pi                           */
    6755399441055744.0,      /* 1.5 * 2^52: round to nearest integer */
    3.141592653589793, 1.2246467991473532e-16, /* pi hi, lo    */
    0.7853981633974483, 3.061616997868383e-17, /* pi/4 hi, lo  */
    /* sin(x) = x + x^3 (S1 + x^2 (S2 + ...)) on [-pi/4, pi/4] */
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    /* cos(x) = 1 - x^2/2 + x^4 (C1 + x^2 (C2 + ...)) */
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11,
    /* atan(u) = u - u^3 (A0 + u^2 (A1 + ...)) on [-7/16, 7/16] */
    3.33333333333329318027e-01, -1.99999999998764832476e-01, 1.42857142725034663711e-01,
    -1.11111104054623557880e-01, 9.09088713343650656196e-02, -7.69187620504482999495e-02,
    6.66107313738753120669e-02, -5.83357013379057348645e-02, 4.97687799461593236017e-02,
    -3.65315727442169155270e-02, 1.62858201153657823623e-02,
};
#define TK(name) tofrt_k[TK_##name]

/* low 32 bits of a double's representation */
TOFRT_HD uint32_t tofrt_lo32(double v)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(v);
#else
    uint64_t u;
    memcpy(&u, &v, sizeof u);
    return (uint32_t)u;
#endif
}

/* sin and cos for |x| <= pi/4 (+ a few ulp) */
TOFRT_HD void tofrt_sincos_kernel(double x, double *sn, double *cs)
{
    double z = x * x;
    /* sin */
    double rs = fma(z, fma(z, fma(z, fma(z, fma(z, TK(S6), TK(S5)), TK(S4)), TK(S3)), TK(S2)), TK(S1));
    *sn = fma(z * x, rs, x);
    /* cos */
    double rc = z * fma(z, fma(z, fma(z, fma(z, fma(z, TK(C6), TK(C5)), TK(C4)), TK(C3)), TK(C2)), TK(C1));
    double hz = 0.5 * z;
    double w = 1.0 - hz;
    *cs = w + (((1.0 - w) - hz) + z * rc);
}

/* Quadrant-reduced sine and cosine: x = n*(pi/2) + r by a two-term Cody-Waite reduction whose
 * first step is exact (fma) and whose second step rounds once, relative to r itself, so the
 * result keeps its relative accuracy near multiples of pi/2 (the neglected third term is
 * n * 1.5e-33).  Returns sin(r), cos(r) and n mod 4. */
TOFRT_HD_BIG int tofrt_sincos_reduced(double x, double *s, double *c)
{
    double t = fma(x, TK(TWO_OVER_PI), TK(MAGIC));
    double n = t - TK(MAGIC);
    double r = fma(-n, TK(PIO2_MID), fma(-n, TK(PIO2_HI), x));
    tofrt_sincos_kernel(r, s, c);
    return (int)(tofrt_lo32(t) & 3u); /* the magic constant leaves n in the low mantissa bits */
}

/* sine and cosine; NaN for NaN, infinities and |x| >= 2^30 */
TOFRT_HD void tofrt_sincos(double x, double *sn, double *cs)
{
    double s, c;
    int q = tofrt_sincos_reduced(x, &s, &c);
    double vs = (q & 1) ? c : s, vc = (q & 1) ? s : c;
    vs = (q & 2) ? -vs : vs;
    vc = ((q + 1) & 2) ? -vc : vc;
    int ok = fabs(x) < 1073741824.0;
    *sn = ok ? vs : NAN;
    *cs = ok ? vc : NAN;
}

TOFRT_HD double tofrt_sin(double x)
{
    double s, c;
    tofrt_sincos(x, &s, &c);
    return s;
}

TOFRT_HD double tofrt_cos(double x)
{
    double s, c;
    tofrt_sincos(x, &s, &c);
    return c;
}

/* tan = sin/cos of the reduced argument (one division) */
TOFRT_HD double tofrt_tan(double x)
{
    double s, c;
    int q = tofrt_sincos_reduced(x, &s, &c);
    double num = (q & 1) ? -c : s, den = (q & 1) ? s : c;
    double v = num / den;
    return (fabs(x) < 1073741824.0) ? v : NAN;
}

/* atan2 with a single division.  With a = min(|x|,|y|), b = max(|x|,|y|):
 * atan(a/b) = atan(u) for u = a/b when a < 7/16 b, else pi/4 + atan(u) with
 * u = (a-b)/(a+b) in (-0.392, 0]; then the octant is undone with hi/lo splits
 * of pi/2 and pi. */
TOFRT_HD_BIG double tofrt_atan2(double y, double x)
{
    double ax = fabs(x), ay = fabs(y);
    int swap = ay > ax;
    double a = swap ? ax : ay, b = swap ? ay : ax; /* NaN stays NaN through the arithmetic below */
    int small = a < 0.4375 * b;
    double num = small ? a : a - b;
    double den = small ? b : a + b;
    double u = num / den;
    double z = u * u;
    double w = z * z;
    double s1 = z * fma(w, fma(w, fma(w, fma(w, fma(w, TK(A10), TK(A8)), TK(A6)), TK(A4)), TK(A2)), TK(A0));
    double s2 = w * fma(w, fma(w, fma(w, fma(w, TK(A9), TK(A7)), TK(A5)), TK(A3)), TK(A1));
    double us = u * (s1 + s2);
    double r = small ? u - us : TK(PIO4_HI) - ((us - TK(PIO4_LO)) - u);
    r = (b == 0.0) ? 0.0 : r;             /* atan2(+-0, +-0): 0/0 above */
    r = (a == INFINITY && b == INFINITY) ? TK(PIO4_HI) : r; /* both infinite: inf/inf above */
    r = swap ? TK(PIO2_HI) - (r - TK(PIO2_MID)) : r;
    r = signbit(x) ? TK(PI_HI) - (r - TK(PI_LO)) : r;
    return copysign(r, y);
}

TOFRT_HD double tofrt_atan(double x) { return tofrt_atan2(x, 1.0); }

/* sin and cos of atan2(y, x) as one elementary function: (y, x) / hypot(x, y), no angle, no
 * polynomial.  Used where the reference takes alpha = np.arctan2(x', z') and immediately evaluates
 * sin(alpha), cos(alpha) (main_rt.py:396-397 -> :184, :203-204, :218-233).  < 2.5 ulp (one square
 * root, one reciprocal, one product each), i.e. closer to the exact value than chaining a 2.5-ulp
 * atan2 with 1.5-ulp sin/cos.  Domain: finite arguments of moderate size (lens coordinates);
 * (0, +-0) gives (0, +-1) like the chained form, NaN propagates. */
TOFRT_HD void tofrt_sincos_atan2(double y, double x, double *sn, double *cs)
{
    double r = sqrt(fma(x, x, y * y));
    double inv = 1.0 / r;
    int zero = (r == 0.0);
    *sn = zero ? y : y * inv;
    *cs = zero ? copysign(1.0, x) : x * inv;
}

/* asin(x) = atan2(x, sqrt((1 - x)(1 + x))); NaN for |x| > 1 (total internal
 * reflection in main_rt.py:277) */
TOFRT_HD double tofrt_asin(double x) { return tofrt_atan2(x, sqrt((1.0 - x) * (1.0 + x))); }

#endif /* TOFRT_MATH_H */
