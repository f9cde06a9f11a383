/*
 * tofrt_math.h -- portable FP64 elementary functions of the TOF engine.
 *
 * sin, cos, tan, atan, atan2, asin built only from IEEE-754 operations that
 * are correctly rounded on every target (+, -, *, /, sqrt, fma), so the very
 * same source gives bit-identical results when compiled
 *   - by nvcc for sm_100a (the product: csrc/tofrt_kernels.cu), and
 *   - by gcc for the host (the oracle's "device-math" twin, oracle/Makefile),
 * which is what makes GPU-vs-CPU parity of the discrete indices and of the
 * travel times exact instead of "within libm noise".  The reference's own
 * elementary functions are NumPy's SIMD loops (np.sin/np.cos/np.tan/
 * np.arctan/np.arctan2/np.arcsin, call sites main_rt.py:184-405); they differ
 * from these by <= ~2 ulp, i.e. <= 1e-19 s on a travel time.
 *
 * Accuracy (tests/test_math.py, against mpmath): <= 2 ulp on the engine's
 * argument ranges.  Domain of the trigonometric reduction: |x| < 2^30 (angles
 * in this engine never leave [-4 pi, 4 pi]); larger finite arguments return
 * NaN instead of a value with no correct digits.
 *
 * Compile without floating-point contraction (nvcc -fmad=false,
 * gcc -ffp-contract=off): every fused operation below is an explicit fma().
 */
#ifndef TOFRT_MATH_H
#define TOFRT_MATH_H

#include <math.h>

#if defined(__CUDACC__)
#define TOFRT_HD __host__ __device__ __forceinline__
#else
#define TOFRT_HD static inline
#endif

#define TOFRT_PIO2_HI 1.5707963267948966     /* 0x1.921fb54442d18p+0  */
#define TOFRT_PIO2_MID 6.123233995736766e-17 /* 0x1.1a62633145c07p-54 */
#define TOFRT_PIO2_LO -1.4973849048591698e-33
#define TOFRT_TWO_OVER_PI 0.6366197723675814
#define TOFRT_PI_HI 3.141592653589793
#define TOFRT_PI_LO 1.2246467991473532e-16
#define TOFRT_PIO4_HI 0.7853981633974483
#define TOFRT_PIO4_LO 3.061616997868383e-17
#define TOFRT_ATAN_HALF_HI 0.4636476090008061 /* atan(1/2) */
#define TOFRT_ATAN_HALF_LO 2.2698777452961687e-17

/* x = n*(pi/2) + (r + rl), |r| <= pi/4 (+ a few ulp); returns n mod 4 in *q.
 * Three-term Cody-Waite with an exact first step (fma) and a double-double
 * remainder, so the result keeps full relative accuracy near multiples of pi/2. */
TOFRT_HD void tofrt_reduce_pio2(double x, double *r, double *rl, int *q)
{
    const double magic = 6755399441055744.0; /* 1.5 * 2^52: round to nearest integer */
    double n = fma(x, TOFRT_TWO_OVER_PI, magic) - magic;
    double r1 = fma(-n, TOFRT_PIO2_HI, x); /* exact */
    double ph = n * TOFRT_PIO2_MID;
    double pl = fma(n, TOFRT_PIO2_MID, -ph); /* ph + pl = n * MID exactly */
    double s = r1 - ph;                      /* two-sum of r1 + (-ph) */
    double bb = s - r1;
    double e = (r1 - (s - bb)) + (-ph - bb);
    *r = s;
    *rl = fma(-n, TOFRT_PIO2_LO, e - pl);
    *q = (int)((long long)n & 3);
}

/* sin(x + y) for |x| <= pi/4, y the tail of x */
TOFRT_HD double tofrt_sin_kernel(double x, double y)
{
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = x * x;
    double r = fma(z, fma(z, fma(z, fma(z, S6, S5), S4), S3), S2);
    double v = z * x;
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

/* cos(x + y) for |x| <= pi/4 */
TOFRT_HD double tofrt_cos_kernel(double x, double y)
{
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = x * x;
    double r = z * fma(z, fma(z, fma(z, fma(z, fma(z, C6, C5), C4), C3), C2), C1);
    double hz = 0.5 * z;
    double w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * y));
}

TOFRT_HD double tofrt_sin(double x)
{
    if (!(fabs(x) < 1073741824.0)) return x - x + NAN; /* NaN, inf, out of domain */
    double r, rl;
    int q;
    tofrt_reduce_pio2(x, &r, &rl, &q);
    double v = (q & 1) ? tofrt_cos_kernel(r, rl) : tofrt_sin_kernel(r, rl);
    return (q & 2) ? -v : v;
}

TOFRT_HD double tofrt_cos(double x)
{
    if (!(fabs(x) < 1073741824.0)) return x - x + NAN;
    double r, rl;
    int q;
    tofrt_reduce_pio2(x, &r, &rl, &q);
    double v = (q & 1) ? tofrt_sin_kernel(r, rl) : tofrt_cos_kernel(r, rl);
    return ((q + 1) & 2) ? -v : v;
}

/* both at once; bit-identical to tofrt_sin / tofrt_cos */
TOFRT_HD void tofrt_sincos(double x, double *sn, double *cs)
{
    if (!(fabs(x) < 1073741824.0)) { *sn = *cs = x - x + NAN; return; }
    double r, rl;
    int q;
    tofrt_reduce_pio2(x, &r, &rl, &q);
    double s = tofrt_sin_kernel(r, rl), c = tofrt_cos_kernel(r, rl);
    double vs = (q & 1) ? c : s, vc = (q & 1) ? s : c;
    *sn = (q & 2) ? -vs : vs;
    *cs = ((q + 1) & 2) ? -vc : vc;
}

/* tan = sin/cos on the reduced argument (one division) */
TOFRT_HD double tofrt_tan(double x)
{
    if (!(fabs(x) < 1073741824.0)) return x - x + NAN;
    double r, rl;
    int q;
    tofrt_reduce_pio2(x, &r, &rl, &q);
    double s = tofrt_sin_kernel(r, rl), c = tofrt_cos_kernel(r, rl);
    return (q & 1) ? (-c) / s : s / c;
}

/* atan(a/b) for 0 <= a <= b, result in [0, pi/4].  The argument is brought
 * into the polynomial's interval with the classic breakpoints 7/16, 11/16
 * applied to numerator and denominator, so there is a single division. */
TOFRT_HD double tofrt_atan_core(double a, double b)
{
    const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01,
                 aT2 = 1.42857142725034663711e-01, aT3 = -1.11111104054623557880e-01,
                 aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
                 aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02,
                 aT8 = 4.97687799461593236017e-02, aT9 = -3.65315727442169155270e-02,
                 aT10 = 1.62858201153657823623e-02;
    double num, den, hi, lo;
    if (a < 0.4375 * b) {
        num = a; den = b; hi = 0.0; lo = 0.0;
    } else if (a < 0.6875 * b) {
        num = 2.0 * a - b; den = 2.0 * b + a; hi = TOFRT_ATAN_HALF_HI; lo = TOFRT_ATAN_HALF_LO;
    } else {
        num = a - b; den = a + b; hi = TOFRT_PIO4_HI; lo = TOFRT_PIO4_LO;
    }
    double u = num / den;
    double z = u * u;
    double w = z * z;
    double s1 = z * fma(w, fma(w, fma(w, fma(w, fma(w, aT10, aT8), aT6), aT4), aT2), aT0);
    double s2 = w * fma(w, fma(w, fma(w, fma(w, aT9, aT7), aT5), aT3), aT1);
    return hi - ((u * (s1 + s2) - lo) - u);
}

TOFRT_HD double tofrt_atan2(double y, double x)
{
    if (x != x || y != y) return x + y;
    double ax = fabs(x), ay = fabs(y);
    int swap = ay > ax;
    double a = swap ? ax : ay, b = swap ? ay : ax;
    double r;
    if (b == 0.0) r = 0.0;                      /* atan2(+-0, +-0) */
    else if (a == INFINITY) r = TOFRT_PIO4_HI;  /* both infinite */
    else r = tofrt_atan_core(a, b);
    if (swap) r = TOFRT_PIO2_HI - (r - TOFRT_PIO2_MID);
    if (signbit(x)) r = TOFRT_PI_HI - (r - TOFRT_PI_LO);
    return copysign(r, y);
}

TOFRT_HD double tofrt_atan(double x)
{
    if (x != x) return x;
    double ax = fabs(x);
    int swap = ax > 1.0;
    double r = swap ? tofrt_atan_core(1.0, ax) : tofrt_atan_core(ax, 1.0);
    if (swap) r = TOFRT_PIO2_HI - (r - TOFRT_PIO2_MID);
    return copysign(r, x);
}

/* asin(x) = atan2(x, sqrt((1 - x)(1 + x))); NaN for |x| > 1 (total internal
 * reflection in main_rt.py:277) */
TOFRT_HD double tofrt_asin(double x)
{
    return tofrt_atan2(x, sqrt((1.0 - x) * (1.0 + x)));
}

#endif /* TOFRT_MATH_H */
