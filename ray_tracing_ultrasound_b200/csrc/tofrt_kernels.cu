// tofrt_kernels.cu -- sm_100a kernels + C ABI of the TOF ray-tracing engine.
//
// What is computed is the reference's pulse-echo trace (main_rt.py:337-405) and
// table assembly (main_rt.py:487-501); how it is computed is B200-first:
//
//   tofrt_prepare_kernel, tofrt_chord_kernel
//       geometry-independent work, once per setup: lens samples, tangents, chords, incident leg
//       and first refraction of every shooter (main_rt.py:338-349) -> small tables in HBM/L2.
//   tofrt_table_kernel<HG, GUIDED>
//       persistent grid, one CTA per SM.  Lens samples and chords are staged once per CTA in
//       shared memory; every WARP owns one (shooter, geometry) at a time (the CTA draws kWarps
//       consecutive items from a dynamic work counter and starts them together), walks its K rays
//       32 at a time with all per-ray state in registers and keeps the per-receiver first-hit
//       reduction warp-private: two block barriers per geometry, none inside it.
//       The return-path entry index of find_line_curve_intersection (main_rt.py:78-99) comes from
//         GUIDED = true   a warp-uniform bisection that brackets the first sign change with every
//                         earlier sign proved (docs/guided_search.md), O(log N) diffs per ray;
//         GUIDED = false  the reference's literal O(N) scan: broadcast LDS.128 of one lens sample,
//                         DMUL+DADD+DADD (never fused: NumPy rounds twice) and a 3-input integer
//                         min/max on the high words ("sign change or exact zero in this chunk"),
//                         warp-wide early exit, exact index resolved once afterwards.
//       Both return identical bits.  HG = true keeps the first-hit state in the output rows
//       (many receivers).
//   tofrt_table_kernel_f32      the same pipeline in float (FP32 mode, reported separately).
//   tofrt_count_hits / tofrt_compact_hits (+ CUB scan)   sparse hit list for the host.
//   tofrt_fermat_kernel         two-point minimum-time entry point (extension).
//   tofrt_fp64_peak_kernel      FP64 issue-rate micro-benchmark (the roofline denominator).
//
// Compile with -fmad=false: the only fused operations are the explicit fma()
// calls inside tofrt_math.h (shared bit-for-bit with the oracle's twin build).
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <string>
#include <vector>

#include "../../include/tofrt.h"
#include "tofrt_math.h"
#include "tofrt_fermat3d.h"
#include "tofrt_raymath.h"

#define HALF_PI 1.5707963267948966 /* np.pi / 2 */

namespace {

#ifndef TOFRT_THREADS
#define TOFRT_THREADS 896
#endif
#ifndef TOFRT_MIN_BLOCKS
#define TOFRT_MIN_BLOCKS 1
#endif
constexpr int kThreadsScan = TOFRT_THREADS;
constexpr int kThreads = kThreadsScan;  // warps per CTA = kThreads / 32 (literal scan, FP32 mode)
constexpr int kWarps = kThreads / 32;
// The guided-search instantiation fits 64 registers, so it runs 1024 threads (32 warps, 8 per
// scheduler): measured 1.2 % faster than 896; the literal scan is 2 % faster at 896 x 72 registers.
#ifndef TOFRT_THREADS_GUIDED
#define TOFRT_THREADS_GUIDED 1024
#endif
constexpr int kThreadsGuided = TOFRT_THREADS_GUIDED;
#ifndef TOFRT_SCAN_CHUNK
#define TOFRT_SCAN_CHUNK 32
#endif
constexpr int kChunk = TOFRT_SCAN_CHUNK;  // lens samples per early-exit check of the literal scan (<= 32)
constexpr int kNoHit = 0x7fffffff;
#ifndef TOFRT_CHUNK_MB
#define TOFRT_CHUNK_MB 64  // output bytes per pipelined chunk of tofrt_table_host
#endif

// Everything the kernels need that does not depend on the geometry.
struct DevSetup {
    // derived constants, computed on the host in the reference's operation order
    double c1, c2, d;
    double A, fourAC, twoA, twod, K2, phi1;  // lens surface (main_rt.py:181-212)
    double c2_over_c1, c1_over_c2;           // v2 / v1 of the two refractions
    int n_shooters;                          // S: shots per geometry (full-matrix capture when S = E)
    const double *x_sh, *z_sh;               // [S] shooter positions
    double rtol, atol, tol_max;              // np.isclose tolerances, and an upper bound of atol+rtol|x_e|
    double rx_min, rx_max;                   // finite receiver range
    int K, Npad, E, E_sorted, has_nonfinite_rx, tables_ok;
    int hits_in_global;  // 1: E too large for the per-warp shared-memory first-hit state -> keep it in the output rows
    int shape;      // +1: sampled lens strictly concave in x, -1: strictly convex, 0: neither (no guided search)
    int guided;     // search mode: 1 = guided (bracketed) search, 0 = the reference's literal O(N) scan
    int vector_math;  // 1: return-path interface laws in vector form (tofrt_raymath.h), 0: the reference's angles
    int log2n;      // ceil(log2(K)) bisection steps
    double X, Z;    // max |x_p|, max |z_p|  (rounding-error bound of a diff)
    int surface;    // 0: the reference's constant-TOF lens h(alpha); 1: flat interface z = h0 (extension)
    double h0;
    // tables in global memory
    const double2 *lens;   // [Npad] (x_p, z_p), padded with copies of the last sample
    const double2 *seg;    // [K-1] chord (m_seg, b_seg) of main_rt.py:127-128
    const double *a_pq, *b_pq, *phi_pq, *z_f, *leg1;  // [K]
    const double *cpq, *spq;  // [K] cos / sin of phi_pq (vector form of the interface laws)
    const double *x_rx;    // [E] as given
    const double *xs;      // [E_sorted] finite receivers, ascending
    const int *perm;       // [E_sorted] original index of xs[i]
    // candidate lookup: bucket[c] = first i with xs[i] >= bx0 + (c-1)/binv (one cell of slack), nb cells
    const int *bucket;
    int nb;
    double bx0, binv;
};

struct TableArgs {
    const double *radius, *offset;
    long long G;
    double *tof;
    int *k_hit;
    tofrt_ray_records rec;
    int any_rec;           // any per-ray record requested (besides legs)
    int s_begin, s_count;  // shooters [s_begin, s_begin + s_count) of the setup; outputs are [s_count][G][...]
    const double *inject_m, *inject_b;  // test hook (tofrt_debug_search_host): replace the reflected line of ray [w][k]
    unsigned long long *counter;
    unsigned long long *stats;  // [4]: rays traced, rays that needed the O(N) scan, warp-chunks that ran it, on-sample passes
};

// ---- small device helpers ---------------------------------------------------

__device__ __forceinline__ unsigned hi32(double v) { return (unsigned)__double2hiint(v); }

// np.sign class: 0 negative, 1 zero, 2 positive, 3 NaN
__device__ __forceinline__ int sign_class(double v) { return v > 0 ? 2 : (v < 0 ? 0 : (v == 0 ? 1 : 3)); }
// np.diff(np.sign(d)) != 0 (NaN - x is NaN and NaN != 0 is True)
__device__ __forceinline__ bool class_change(int c0, int c1) { return c0 == 3 || c1 == 3 || c0 != c1; }

// np.isclose(a, b, rtol, atol)
__device__ __forceinline__ bool np_isclose(double a, double b, double rtol, double atol)
{
    return ((fabs(a - b) <= atol + rtol * fabs(b)) && isfinite(b)) || (a == b);
}
// Python min/max on two floats
__device__ __forceinline__ double py_min(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double py_max(double a, double b) { return (b > a) ? b : a; }

__device__ __forceinline__ double dist2d(double x1, double z1, double x2, double z2)
{
    return sqrt((x1 - x2) * (x1 - x2) + (z1 - z2) * (z1 - z2));
}

// vertical difference of main_rt.py:78-79, two roundings
__device__ __forceinline__ double line_diff(double2 s, double m, double b) { return s.y - (m * s.x + b); }

// h(alpha), dz/dalpha, dx/dalpha (main_rt.py:180-234); B of :184 and :203 are the same bits
template <bool FLAT>
__device__ __forceinline__ void lens_point_t(const DevSetup &S, double sn, double cs, double *h, double *dz,
                                             double *dx)
{
    double hh, dh;
    if (!FLAT) {
        double B = S.twod * cs - S.K2;
        double sd = sqrt(B * B - S.fourAC);
        hh = (-B - sd) / S.twoA;
        double dB = -S.twod * sn;
        double dB2 = 2 * B * dB;
        double dS = (1 / (2 * sd)) * dB2;
        dh = S.phi1 * (dB + dS);
    } else {  // flat interface z = h0 in the same polar form: h = h0 / cos(alpha) (extension, no reference analogue)
        hh = S.h0 / cs;
        dh = S.h0 * sn / (cs * cs);
    }
    *h = hh;
    *dz = dh * cs - hh * sn;
    *dx = dh * sn + hh * cs;
}
// run-time surface switch, for the kernels where it does not matter (setup, two-point solve)
__device__ __forceinline__ void lens_point(const DevSetup &S, double sn, double cs, double *h, double *dz, double *dx)
{
    if (S.surface == 0) lens_point_t<false>(S, sn, cs, h, dz, dx);
    else lens_point_t<true>(S, sn, cs, h, dz, dx);
}

// refraction of main_rt.py:267-280 with the tuple form of dzdx
__device__ __forceinline__ double refract(double phi_in, double dz, double dx, double v2_over_v1)
{
    double phi_slope = tofrt_atan2(dz, dx);
    double theta_1 = phi_in - (phi_slope + HALF_PI);
    double theta_2 = tofrt_asin(v2_over_v1 * tofrt_sin(theta_1));
    return phi_slope - HALF_PI + theta_2;
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier (cp.async.bulk, sm_90+) ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}

// ---- prepare kernel: main_rt.py:338-349 for every launch angle ----------------
__global__ void tofrt_prepare_kernel(DevSetup S, const double *alpha, double2 *lens, double *a_pq,
                                     double *b_pq, double *phi_pq, double *leg1, double *cpq, double *spq, int *bad)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= S.Npad) return;
    int kk = k < S.K ? k : S.K - 1;  // padding repeats the last sample
    double al = alpha[kk];
    double sn, cs, h, dz, dx;
    tofrt_sincos(al, &sn, &cs);
    lens_point(S, sn, cs, &h, &dz, &dx);
    double x_p = h * sn, z_p = h * cs;
    lens[k] = make_double2(x_p, z_p);
    if (k >= S.K) return;
    if (!(fabs(x_p) < 1e150 && fabs(z_p) < 1e150)) atomicOr(bad, 1);
    for (int s = 0; s < S.n_shooters; ++s) {  // main_rt.py:341-349 for every shooter
        const double x_sh = S.x_sh[s], z_sh = S.z_sh[s];
        double phi_ap = tofrt_atan2(z_sh - z_p, x_sh - x_p);
        double phi = refract(phi_ap, dz, dx, S.c2_over_c1);
        double a = tofrt_tan(phi);
        size_t o = (size_t)s * S.K + k;
        a_pq[o] = a;
        b_pq[o] = z_p - a * x_p;
        phi_pq[o] = phi;
        tofrt_sincos(phi, &spq[o], &cpq[o]);
        leg1[o] = dist2d(x_sh, z_sh, x_p, z_p) / S.c1;
    }
}

// chords of the sampled lens, main_rt.py:127-128 (geometry-independent, so computed once)
__global__ void tofrt_chord_kernel(int K, const double2 *lens, double2 *seg)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= K - 1) return;
    double2 p1 = lens[j], p2 = lens[j + 1];
    double m_seg = (p2.y - p1.y) / (p2.x - p1.x);
    double b_seg = p1.y - m_seg * p1.x;
    seg[j] = make_double2(m_seg, b_seg);
}

// ---- exact (slow) pieces of find_line_curve_intersection ------------------------

// first j in [j0, N-2] with a sign-class change between samples j and j+1, else -1
__device__ __noinline__ int scan_exact_from(const double2 *lens, int N, double m, double b, int j0)
{
    int cprev = sign_class(line_diff(lens[j0], m, b));
    for (int j = j0; j + 1 < N; ++j) {
        int cnext = sign_class(line_diff(lens[j + 1], m, b));
        if (class_change(cprev, cnext)) return j;
        cprev = cnext;
    }
    return -1;
}

// main_rt.py:106-168: intersect the line with chord (idx, idx+1)
__device__ __forceinline__ void chord_intersection(const double2 *lens, const double2 *seg, int idx, double m,
                                                   double b, double *xi, double *zi)
{
    double2 p1 = lens[idx], p2 = lens[idx + 1], sg = seg[idx];
    double x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
    double rx = NAN, ry = NAN;
    if (np_isclose(x1, x2, 1e-5, 1e-8)) {  // :110-123 vertical chord
        double y_int = m * x1 + b;
        if (y_int >= py_min(y1, y2) - 1e-9 && y_int <= py_max(y1, y2) + 1e-9) { rx = x1; ry = y_int; }
    } else {
        double m_seg = sg.x, b_seg = sg.y;  // :127-128
        if (np_isclose(m, m_seg, 1e-5, 1e-8)) {  // :131-144 parallel
            if (np_isclose(b, b_seg, 1e-5, 1e-8)) { rx = (x1 + x2) / 2.0; ry = m * rx + b; }
        } else {  // :147-168
            double x_int = (b_seg - b) / (m - m_seg);
            double y_int = m * x_int + b;
            if (x_int >= py_min(x1, x2) - 1e-9 && x_int <= py_max(x1, x2) + 1e-9 &&
                y_int >= py_min(y1, y2) - 1e-9 && y_int <= py_max(y1, y2) + 1e-9) {
                rx = x_int; ry = y_int;
            }
        }
    }
    *xi = rx; *zi = ry;
}

// ---- the table kernel -----------------------------------------------------------
// HG: first-hit state in the output rows (global memory) instead of warp-private shared memory
// GUIDED: compile the guided search (its rare fallback is then a plain per-ray exact scan) or the
// warp-wide literal scan -- two lean instruction streams instead of one kernel carrying both.
// VEC: the two return-path interface laws in vector form (tofrt_raymath.h) instead of the reference's angles
template <bool HG, bool GUIDED, bool FLAT, bool VEC>
__global__ void __launch_bounds__(GUIDED ? kThreadsGuided : kThreads, TOFRT_MIN_BLOCKS)
    tofrt_table_kernel(DevSetup S, TableArgs T)
{
    constexpr int kThreads = GUIDED ? kThreadsGuided : kThreadsScan;  // shadows the namespace constants
    constexpr int kWarps = kThreads / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = S.K, N = S.K, Npad = S.Npad, E = S.E;
    // shared-memory carve-up
    double2 *s_lens = reinterpret_cast<double2 *>(smem_raw);
    double2 *s_seg = s_lens + Npad;                  // [K-1] (+1 pad)
    double *s_xs = reinterpret_cast<double *>(s_seg + K);
    double *s_tof = s_xs + S.E_sorted;               // [kWarps][E] (absent when HG)
    int *s_perm = reinterpret_cast<int *>(s_tof + (HG ? (size_t)0 : (size_t)kWarps * E));
    int *s_khit = s_perm + S.E_sorted;               // [kWarps][E] (absent when HG)
    int *s_bucket = s_khit + (HG ? (size_t)0 : (size_t)kWarps * E);  // [nb]

    // Stage the sampled surface once per CTA: two TMA bulk copies (lens samples, chords; 16-byte
    // aligned, multiples of 16 bytes) complete on one mbarrier while the threads load the small
    // receiver table themselves.
    __shared__ __align__(8) unsigned long long s_bar;
    const unsigned lens_bytes = (unsigned)Npad * (unsigned)sizeof(double2), seg_bytes = (unsigned)K * (unsigned)sizeof(double2);
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&s_bar, lens_bytes + seg_bytes);
        tma_bulk_g2s(s_lens, S.lens, lens_bytes, &s_bar);
        tma_bulk_g2s(s_seg, S.seg, seg_bytes, &s_bar);
    }
    for (int i = threadIdx.x; i < S.E_sorted; i += kThreads) { s_xs[i] = S.xs[i]; s_perm[i] = S.perm[i]; }
    for (int i = threadIdx.x; i < S.nb; i += kThreads) s_bucket[i] = S.bucket[i];
    while (!mbar_try_wait(&s_bar, 0)) {}
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // first-hit state of the geometry this warp owns: warp-private shared memory, or (large E) the
    // geometry's own output rows in global memory -- read back with volatile loads, because the
    // atomicMin below happens in L2 and must not be shadowed by a stale L1 line
    typedef typename std::conditional<HG, volatile double, double>::type tof_t;
    typedef typename std::conditional<HG, volatile int, int>::type khit_t;
    tof_t *w_tof = s_tof + (size_t)warp * E;
    khit_t *w_khit = s_khit + (size_t)warp * E;
    const int nchunks = (K + 31) >> 5;
    const unsigned FULL = 0xffffffffu;
    const double cand_lo = S.rx_min - S.tol_max, cand_hi = S.rx_max + S.tol_max;  // aperture +- largest tolerance
    unsigned st_rays = 0, st_scan_rays = 0, st_scan_warps = 0, st_onsample = 0;

    // Work distribution: the CTA takes kWarps consecutive (shooter, geometry) items at a time, one per
    // warp, and its warps start them together (two block barriers per ~40 us of work).  Warps that
    // begin a geometry in step run the same code at the same time, which is worth 3.7 % over
    // fetching per warp (the instruction stream of one chunk is ~40 KB); any tighter coupling
    // (a barrier every chunk, or every few) costs more than it gives -- measured.
    __shared__ unsigned long long s_base;
    const unsigned long long total = (unsigned long long)T.G * (unsigned long long)T.s_count;
    for (;;) {
        __syncthreads();  // everyone has read the previous s_base
        if (threadIdx.x == 0) s_base = atomicAdd(T.counter, (unsigned long long)kWarps);
        __syncthreads();
        if (s_base >= total) break;
        const unsigned long long g64 = s_base + warp;
        if (g64 >= total) continue;  // tail: fewer items than warps
        const long long w = (long long)g64;               // work item = (shooter, geometry)
        long long g = w, s_local = 0;
        if (T.s_count > 1) { s_local = w / T.G; g = w - s_local * T.G; }  // 64-bit division only when several shooters
        const size_t sk = (size_t)(T.s_begin + s_local) * K;  // offset of the shooter's per-ray tables
        const double r_outer = T.radius[g], pipe_offset = T.offset[g];
        const double rsq = r_outer * r_outer, offsq = pipe_offset * pipe_offset;
        if (HG) { w_tof = T.tof + (size_t)w * E; w_khit = T.k_hit + (size_t)w * E; }
        for (int e = lane; e < E; e += 32) { w_khit[e] = kNoHit; w_tof[e] = 0.0; }
        __syncwarp();

        int h1 = -1, h2 = -1;  // guided search: this lane's bracket starts in the previous two chunks
        for (int c = 0; c < nchunks; ++c) {
            const int k = (c << 5) + lane;
            const bool valid = k < K;
            const int kk = valid ? k : K - 1;
            // per-ray constants of the shot: one coalesced, L1-resident load per 32 rays
            const double a_pq = __ldg(S.a_pq + sk + kk), b_pq = __ldg(S.b_pq + sk + kk);
            const double phi_pq = VEC ? 0.0 : __ldg(S.phi_pq + sk + kk);
            const double cpq = VEC ? __ldg(S.cpq + sk + kk) : 0.0, spq = VEC ? __ldg(S.spq + sk + kk) : 0.0;
            const double z_f = __ldg(S.z_f + kk);

            // ---- main_rt.py:351-364 pipe intersection, upper root ----
            double x_q, z_q;
            {
                double A2 = a_pq * a_pq + 1;
                double B2 = 2 * (a_pq * b_pq - pipe_offset);
                double C2 = offsq + b_pq * b_pq - rsq;
                double sd = sqrt(B2 * B2 - 4 * A2 * C2);
                double den = 2 * A2;
                double x_q1 = (-B2 + sd) / den, x_q2 = (-B2 - sd) / den;
                double z_q1 = a_pq * x_q1 + b_pq, z_q2 = a_pq * x_q2 + b_pq;
                bool up = z_q1 > z_q2;
                x_q = up ? x_q1 : x_q2;
                z_q = up ? z_q1 : z_q2;
            }
            // ---- main_rt.py:367-376 reflection on the pipe ----
            double phi_l = 0.0, cl = 0.0, sl = 0.0, m, b;  // reflected direction: angle, or r^2 (cos, sin)
            if (VEC) {
                tofrt_vec_reflect(x_q, rsq, cpq, spq, &cl, &sl);
                m = sl / cl;
                b = z_q - m * x_q;
            } else {
                double slope = -x_q / sqrt(rsq - x_q * x_q);  // dzdx_pipe ignores the offset (:237-238)
                double phi_slope = tofrt_atan(slope);
                double theta_1 = phi_pq - (phi_slope + HALF_PI);
                phi_l = phi_slope - HALF_PI + (-theta_1);
                m = tofrt_tan(phi_l);
                b = z_q - m * x_q;
            }
            if (T.inject_m) {  // test hook: search an arbitrary line instead (uniform branch, off in production)
                m = T.inject_m[(size_t)w * K + kk];
                b = T.inject_b[(size_t)w * K + kk];
            }

            // ---- main_rt.py:78-99 first sign change over the sampled lens ----
            // fast path precondition: every diff is finite, so only {-,0,+} occur
            const bool fast = valid && S.tables_ok && fabs(m) < 1e150 && fabs(b) < 1e150;
            int idx = -1;             // >= 0 segment, -1 none
            bool need_scan = fast;    // lanes whose index the O(N) scan below still has to find
            bool skip_on_sample = false;
            if (GUIDED) {
                // ---- guided search: same idx as the scan, in O(log N) diffs ------------------
                // g_j = shape * diff_j is, in exact arithmetic, a concave sequence over the sampled
                // lens (chord slopes strictly monotone, checked by tofrt_setup), and a computed diff
                // differs from the exact one by < tol = 2^-48 (|m| X + |b| + Z).  So a diff beyond
                // +-tol has a certain sign, a concave sequence is >= the smaller of two of its values
                // in between, and it is monotone on either side of its maximum: that brackets the
                // first sign change to [lo, lo+1] with every earlier sign known; the exact scan then
                // starts at lo.  Anything uncertain (|diff| <= tol where it matters) falls back to the
                // full scan, so the result is identical by construction.
                const double sh = (double)S.shape;
                const double tol = (fabs(m) * S.X + fabs(b) + S.Z) * 0x1p-48 + 1e-290;
                const double g0 = sh * line_diff(s_lens[0], m, b);
                const double gN = sh * line_diff(s_lens[N - 1], m, b);
                const bool pos0 = g0 > tol, neg0 = g0 < -tol;
                bool bracket = false, certain_none = false;
                double edge = 0.0;  // lower bound of min|diff| when there is certainly no change
                if (fast && pos0 && gN > tol) { certain_none = true; edge = fmin(g0, gN) - tol; }
                else if (fast && (pos0 || neg0)) bracket = true;
                // One bisection serves both cases.  lo advances to mid iff g_mid certainly has the
                // sign of g_0 and, when g_0 < 0, mid is still before the maximum of g (chord slope
                // on the rising side): then every sample up to mid certainly has that sign too.
                const double sgn = pos0 ? sh : -sh;
                const double ms_eff = pos0 ? -INFINITY : sh * m;
                // "diff_mid certainly has the sign of diff_0", one fused operation per probe:
                //   sgn (z - m x - b) > tol   <=   sgn z > fma(sgn m, x, sgn b + tol)
                // (sgn = +-1: the products are exact; the two roundings are < tol/16, see docs/guided_search.md)
                const double cm = sgn * m, cbt = sgn * b + tol;
                const int zflip = sgn < 0.0 ? (int)0x80000000 : 0;
                const int sflip = S.shape < 0 ? (int)0x80000000 : 0;  // uniform: sh * slope by a sign flip
                const char *lens_b = reinterpret_cast<const char *>(s_lens);
                const int slope_off = (int)(reinterpret_cast<const char *>(s_seg) - lens_b);  // uniform
                // cond(j): sample j certainly has the sign of g_0 and (g_0 < 0) lies on the rising side.  ANY j
                // with cond(j) is a valid lower end of the bracket, any j without it a valid upper end
                // (the concavity arguments do not depend on how j was chosen, docs/guided_search.md).
                auto cond_at = [&](int j) {
                    const char *pa = lens_b + j * 16;  // lens[j]; the chord slope sits slope_off bytes further
                    const double2 pm = *reinterpret_cast<const double2 *>(pa);
                    const double slope = *reinterpret_cast<const double *>(pa + slope_off);
                    const double zs = __hiloint2double(__double2hiint(pm.y) ^ zflip, __double2loint(pm.y));
                    const double ss = __hiloint2double(__double2hiint(slope) ^ sflip, __double2loint(slope));
                    return (zs > fma(cm, pm.x, cbt)) && (ss > ms_eff);
                };
                int lo = 0, hi = N - 1, nsteps = S.log2n;
                // Predicted bracket: along a shot the entry index is a smooth function of the ray index, so
                // this lane's brackets of the previous two chunks extrapolate to within a few samples
                // (measured on config 3: +-4 covers 96 % of the chunks, +-8 99.6 %).  Two probes certify [p - W, p + W] as a
                // bracket for every lane of the warp, and the bisection needs log2(2 W) steps instead of
                // log2(N); otherwise the full bisection runs as before.
#ifndef TOFRT_PRED_W  /* measured: W = 4 (3 steps) 7.81 ms, W = 8 (4 steps) 7.91 ms, W = 16 (5 steps) 8.11 ms; none: 8.44 ms */
#define TOFRT_PRED_W 4
#define TOFRT_PRED_STEPS 3
#endif
                constexpr int kPredW = TOFRT_PRED_W, kPredSteps = TOFRT_PRED_STEPS;  // 2 * kPredW <= 2^kPredSteps
                if (__all_sync(FULL, !bracket || (h1 >= 0 && h2 >= 0))) {
                    const int p = 2 * h1 - h2;
                    const int plo = min(max(p - kPredW, 0), N - 2), phi = min(plo + 2 * kPredW, N - 1);
                    const bool okp = cond_at(plo) && !cond_at(phi);
                    if (__all_sync(FULL, !bracket || okp)) { lo = plo; hi = phi; nsteps = kPredSteps; }
                }
                for (int it = 0; it < nsteps; ++it) {
                    const int mid = (lo + hi) >> 1;  // <= N-2 whenever hi - lo > 1
                    const bool cond = cond_at(mid);
                    const bool act = bracket && (hi - lo > 1);
                    lo = (act && cond) ? mid : lo;
                    hi = (act && !cond) ? mid : hi;
                }
                h2 = h1;
                h1 = bracket ? lo : -1;
                if (bracket && hi - lo > 1) { bracket = false; }  // cannot happen (log2n steps suffice); scan if it did
                if (bracket && neg0) {
                    // hi = lo + 1 failed the test (or is the untested last sample).  If g_hi is still
                    // certainly negative, hi is at/after the maximum while lo is before it: g_hi IS
                    // the maximum of g, so no diff is >= 0.
                    // (lo == 0 was never tested against the slope: if g already falls from the first
                    // sample on, the maximum is g_0 -- so the bound below uses the larger of the two)
                    // ("certainly negative" must be the loop's own probe: hi may have failed exactly that)
                    const double2 ph = s_lens[hi];
                    const double ghi = sh * line_diff(ph, m, b);
                    const double glo = sh * line_diff(s_lens[lo], m, b);
                    const double zh = __hiloint2double(__double2hiint(ph.y) ^ zflip, __double2loint(ph.y));
                    if (zh > fma(cm, ph.x, cbt)) { bracket = false; certain_none = true; edge = -fmax(ghi, glo) - tol; }
                }
                if (bracket) {
                    // exact classes from lo on (normally one step)
                    int cprev = sign_class(line_diff(s_lens[lo], m, b));
                    int j = lo;
                    for (; j + 1 < N; ++j) {
                        int cnext = sign_class(line_diff(s_lens[j + 1], m, b));
                        if (class_change(cprev, cnext)) break;
                        cprev = cnext;
                    }
                    if (j + 1 < N) idx = j;
                    need_scan = false;  // no change after lo either => none at all (on-sample test decides)
                } else if (certain_none) {
                    need_scan = false;
                    skip_on_sample = edge > 1e-8;  // every |diff| > 1e-8: main_rt.py:86-96 returns None
                }
            }
            int flagged = -1;
            unsigned gmin = 0xffffffffu;  // min over all chunks of the high words (see below)
            st_rays += valid; st_scan_rays += need_scan;
            if (GUIDED) {
                // a ray the guided search could not decide (|diff| <= tol where it matters): plain exact scan
                if (__any_sync(FULL, need_scan)) st_scan_warps += (lane == 0);
                if (need_scan) { idx = scan_exact_from(s_lens, N, m, b, 0); flagged = 0; }
            } else if (__any_sync(FULL, need_scan)) {
                st_scan_warps += (lane == 0);
                unsigned hp = hi32(line_diff(s_lens[0], m, b));
                for (int base = 1; base < Npad; base += kChunk) {
                    unsigned mn = hp, mx = hp;
#pragma unroll
                    for (int u = 0; u < kChunk; u += 2) {
                        unsigned ha = hi32(line_diff(s_lens[base + u], m, b));
                        unsigned hb = hi32(line_diff(s_lens[base + u + 1], m, b));
                        mn = __vimin3_u32(mn, ha, hb);
                        mx = __vimax3_u32(mx, ha, hb);
                        hp = hb;
                    }
                    gmin = min(gmin, mn);
                    // sign differs inside the chunk, or a +-0 (high word 0 / 0x80000000) occurred
                    bool flag = ((int)(mn ^ mx) < 0) || ((mn << 1) == 0u);
                    if (flag && flagged < 0) flagged = base;
                    if (__all_sync(FULL, flagged >= 0 || !need_scan)) break;
                }
            }
            if (!GUIDED && need_scan && flagged >= 0) {
                // resolve the exact index inside the flagged chunk without a data-dependent loop:
                // classes (-1, 0, +1) of samples flagged-1 .. flagged+kChunk-1 (all finite here),
                // bit u of `chg` = class changes between samples j0+u and j0+u+1
                const int j0 = flagged - 1;
                double d = line_diff(s_lens[j0], m, b);
                int cprev = (((unsigned)__double2hiint(d) << 1) | (unsigned)__double2loint(d)) ? (__double2hiint(d) >> 31) | 1 : 0;
                unsigned chg = 0;
#pragma unroll
                for (int u = 0; u < kChunk; ++u) {
                    d = line_diff(s_lens[j0 + u + 1], m, b);
                    int cnext = (((unsigned)__double2hiint(d) << 1) | (unsigned)__double2loint(d)) ? (__double2hiint(d) >> 31) | 1 : 0;
                    chg |= (cnext != cprev ? 1u : 0u) << u;
                    cprev = cnext;
                }
                if (chg) idx = j0 + __ffs(chg) - 1;  // padding repeats the last sample: never a change there
            }
            if (valid && idx < 0) {
                if (!fast) idx = scan_exact_from(s_lens, N, m, b, 0);
                // conservative flag without a real change (a denormal diff): keep going exactly
                else if (!GUIDED && need_scan && flagged >= 0 && flagged + kChunk - 1 < N - 1)
                    idx = scan_exact_from(s_lens, N, m, b, flagged + kChunk - 1);
            }
            // ---- main_rt.py:84-96 no crossing: is a sample on the line (|diff| <= 1e-8)? ----
            // All diffs of such a ray have one sign, so the unsigned minimum of their high words is
            // the high word of min|diff|: only rays that come within 1e-8 of a sample (hi <= hi(1e-8))
            // need the test.  It is warp-cooperative: 32 lanes take 32 consecutive samples of one ray.
            int on_sample = -1;
            {
                bool maybe = valid && idx < 0 && !skip_on_sample &&
                             (!fast || !need_scan || flagged >= 0 ||
                              (gmin & 0x7fffffffu) <= 0x3e45798eu /* hi32(1e-8) */);
                unsigned need = __ballot_sync(FULL, maybe);
                st_onsample += (lane == 0) ? __popc(need) : 0;
                while (need) {
                    int src = __ffs(need) - 1;
                    need &= need - 1;
                    double mm = __shfl_sync(FULL, m, src), bb = __shfl_sync(FULL, b, src);
                    int first = kNoHit;
                    for (int j = lane; j < N; j += 32) {
                        double dd = line_diff(s_lens[j], mm, bb);
                        if (fabs(dd) <= 1e-8) { first = j; break; }
                    }
                    first = __reduce_min_sync(FULL, first);
                    if (lane == src && first != kNoHit) on_sample = first;
                }
            }
            double ix = NAN, iz = NAN;
            if (idx >= 0) chord_intersection(s_lens, s_seg, idx, m, b, &ix, &iz);
            else if (on_sample >= 0) { double2 s = s_lens[on_sample]; ix = s.x; iz = s.y; }

            // ---- main_rt.py:396-405 refraction back into the lens, landing point ----
            // NaN in -> NaN out, so whole 32-ray groups that missed the lens, or are totally
            // reflected (|sin| > 1 in main_rt.py:277), skip the rest of the arithmetic.
            double target_x = NAN;
            if (__any_sync(FULL, valid && ix == ix)) {
                // alpha_i = arctan2(ix, iz) is only ever used through sin/cos (main_rt.py:396-397):
                // one fused elementary function, no angle
                double sn, cs, h, dz, dx;
                tofrt_sincos_atan2(ix, iz, &sn, &cs);
                lens_point_t<FLAT>(S, sn, cs, &h, &dz, &dx);
                // refraction(phi_l, (dz, dx), c2, c1), main_rt.py:272-278
                double phi_slope = 0.0, s_arg;
                if (VEC) s_arg = S.c1_over_c2 * tofrt_vec_sin_theta1(cl, sl, rsq, dz, dx);
                else {
                    phi_slope = tofrt_atan2(dz, dx);
                    double theta_1 = phi_l - (phi_slope + HALF_PI);
                    s_arg = S.c1_over_c2 * tofrt_sin(theta_1);
                }
                if (__any_sync(FULL, valid && fabs(s_arg) <= 1.0)) {
                    double a_i;
                    if (VEC) a_i = tofrt_vec_refract_slope(s_arg, dz, dx);
                    else {
                        double theta_2 = tofrt_asin(s_arg);
                        double phi_last = phi_slope - HALF_PI + theta_2;
                        a_i = tofrt_tan(phi_last);
                    }
                    double b_i = iz - a_i * ix;
                    target_x = (z_f - b_i) / a_i;
                }
            }

            // ---- optional per-ray records ----
            if (T.any_rec && valid) {  // one uniform test in table-only mode
                size_t o = (size_t)w * K + k;
                if (T.rec.pipe_x) T.rec.pipe_x[o] = x_q;
                if (T.rec.pipe_z) T.rec.pipe_z[o] = z_q;
                if (T.rec.lens2_x) T.rec.lens2_x[o] = ix;
                if (T.rec.lens2_z) T.rec.lens2_z[o] = iz;
                if (T.rec.target_x) T.rec.target_x[o] = target_x;
                if (T.rec.idx) T.rec.idx[o] = idx >= 0 ? idx : (on_sample >= 0 ? -2 - on_sample : -1);
            }

            // ---- main_rt.py:487-501 which receivers does this ray land on? ----
            // candidates: finite receivers within tol_max of target_x (binary search in the
            // sorted copy); the exact np.isclose decides.  +-inf landing points can only match
            // an identical infinite receiver.
            bool cand = valid && target_x >= cand_lo && target_x <= cand_hi;
            bool special = valid && S.has_nonfinite_rx && isinf(target_x);
            const bool want_legs = T.rec.legs != nullptr;
            if (__any_sync(FULL, cand || special || want_legs)) {
                int lo = 0, hi = 0;
                bool mine = false;
                if (cand) {
                    double tl = target_x - S.tol_max, th = target_x + S.tol_max;
                    // first sorted receiver >= tl: a bucket lookup (start provably at or before it)
                    // plus a short forward walk; binary search when no bucket table was built
                    int l = 0;
                    if (S.nb > 0) {
                        const double cpos = (tl - S.bx0) * S.binv;
                        const int c = cpos > 0.0 ? (cpos < (double)(S.nb - 1) ? (int)cpos : S.nb - 1) : 0;
                        l = s_bucket[c];
                        while (l < S.E_sorted && s_xs[l] < tl) ++l;
                    } else {
                        int r = S.E_sorted;
                        while (l < r) { int mid = (l + r) >> 1; if (s_xs[mid] < tl) l = mid + 1; else r = mid; }
                    }
                    lo = l;
                    hi = l;
                    while (hi < S.E_sorted && s_xs[hi] <= th) ++hi;
                    for (int i = lo; i < hi; ++i)
                        if (np_isclose(target_x, s_xs[i], S.rtol, S.atol)) { atomicMin(const_cast<int *>(&w_khit[s_perm[i]]), k); mine = true; }
                }
                if (special) {
                    for (int e = 0; e < E; ++e)
                        if (target_x == S.x_rx[e]) { atomicMin(const_cast<int *>(&w_khit[e]), k); mine = true; }
                }
                __syncwarp();
                if (mine || (want_legs && valid)) {
                    // main_rt.py:497-500 / main_compare.py:514-517
                    double2 p = s_lens[kk];
                    double t1 = S.leg1[sk + kk];
                    double t2 = dist2d(p.x, p.y, x_q, z_q) / S.c2;
                    double t3 = dist2d(x_q, z_q, ix, iz) / S.c2;
                    double t4 = dist2d(ix, iz, target_x, z_f) / S.c1;
                    if (want_legs) {
                        size_t o = (size_t)w * 4 * K + k;
                        T.rec.legs[o] = t1; T.rec.legs[o + K] = t2;
                        T.rec.legs[o + 2 * (size_t)K] = t3; T.rec.legs[o + 3 * (size_t)K] = t4;
                    }
                    if (mine) {
                        double tof = t1;
                        tof += t2;
                        tof += t3;
                        tof += t4;
                        for (int i = lo; i < hi; ++i) {
                            int e = s_perm[i];
                            if (w_khit[e] == k) w_tof[e] = tof;
                        }
                        if (special)
                            for (int e = 0; e < E; ++e)
                                if (w_khit[e] == k) w_tof[e] = tof;
                    }
                }
                __syncwarp();
            }
        }
        // ---- one geometry done: coalesced row of the [G][E] table ----
        __syncwarp();
        for (int e = lane; e < E; e += 32) {
            int kh = w_khit[e];
            if (!HG) T.tof[(size_t)w * E + e] = w_tof[e];
            T.k_hit[(size_t)w * E + e] = kh == kNoHit ? -1 : kh;
        }
        __syncwarp();
    }
    if (T.stats) {
        st_rays = __reduce_add_sync(FULL, st_rays); st_scan_rays = __reduce_add_sync(FULL, st_scan_rays);
        if (lane == 0) {
            atomicAdd(T.stats + 0, (unsigned long long)st_rays); atomicAdd(T.stats + 1, (unsigned long long)st_scan_rays);
            atomicAdd(T.stats + 2, (unsigned long long)st_scan_warps); atomicAdd(T.stats + 3, (unsigned long long)st_onsample);
        }
    }
}

// =====================================================================================
// FP32 mode (north_star: "FP32 mode reported separately with its measured max error").
// Same pipeline, per-ray arithmetic in float (interface laws in vector form, MUFU square roots and
// reciprocals); the geometry-independent tables are the FP64 ones rounded to float.  No bit-parity is claimed for
// this mode: tests/bench report max |dt| and the hit-decision differences against FP64.
// =====================================================================================
struct DevSetupF {
    float c1, c2, fourAC, twoA, twod, K2, phi1, c1_over_c2;
    float rtol, atol, tol_max, rx_min, rx_max, X, Z;
    int K, E, E_sorted, has_nonfinite_rx, shape, log2n, n_shooters, surface;
    float h0;
    const float2 *lens, *seg;
    const float *a_pq, *b_pq, *phi_pq, *cpq, *spq, *z_f, *leg1, *x_rx, *xs;
    const int *perm;
};

struct TableArgsF {
    const double *radius, *offset;
    long long G;
    float *tof;
    int *k_hit;
    int s_begin, s_count;
    unsigned long long *counter;
};

__global__ void tofrt_to_float_kernel(int n, const double *src, float *dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (float)src[i];
}

// float helpers of the FP32 mode (MUFU-based square roots; divisions are __fdividef).  The interface
// laws are evaluated in vector form (see tofrt_raymath.h), so no float elementary functions are needed.
__device__ __forceinline__ float fsqrt_fast(float x)
{
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// flush-to-zero variant: no denormal pre-scaling around the MUFU (screening passes only)
__device__ __forceinline__ float fsqrt_fast_ftz(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float line_diff_f(float2 s, float m, float b) { return s.y - (m * s.x + b); }
__device__ __forceinline__ int sign_class_f(float v) { return v > 0 ? 2 : (v < 0 ? 0 : (v == 0 ? 1 : 3)); }
__device__ __forceinline__ bool np_isclose_f(float a, float b, float rtol, float atol)
{
    return ((fabsf(a - b) <= atol + rtol * fabsf(b)) && isfinite(b)) || (a == b);
}
__device__ __forceinline__ float dist2d_f(float x1, float z1, float x2, float z2)
{
    return sqrtf((x1 - x2) * (x1 - x2) + (z1 - z2) * (z1 - z2));
}
__device__ __noinline__ int scan_exact_from_f(const float2 *lens, int N, float m, float b, int j0)
{
    int cprev = sign_class_f(line_diff_f(lens[j0], m, b));
    for (int j = j0; j + 1 < N; ++j) {
        int cnext = sign_class_f(line_diff_f(lens[j + 1], m, b));
        if (class_change(cprev, cnext)) return j;
        cprev = cnext;
    }
    return -1;
}

__global__ void __launch_bounds__(kThreads, 1) tofrt_table_kernel_f32(DevSetupF S, TableArgsF T)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = S.K, N = S.K, E = S.E;
    float2 *s_lens = reinterpret_cast<float2 *>(smem_raw);
    float2 *s_seg = s_lens + N;
    float *s_xs = reinterpret_cast<float *>(s_seg + N);
    float *s_tof = s_xs + S.E_sorted;                 // [kWarps][E]
    int *s_perm = reinterpret_cast<int *>(s_tof + (size_t)kWarps * E);
    int *s_khit = s_perm + S.E_sorted;                // [kWarps][E]
    for (int i = threadIdx.x; i < N; i += kThreads) s_lens[i] = S.lens[i];
    for (int i = threadIdx.x; i < N - 1; i += kThreads) s_seg[i] = S.seg[i];
    for (int i = threadIdx.x; i < S.E_sorted; i += kThreads) { s_xs[i] = S.xs[i]; s_perm[i] = S.perm[i]; }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *w_tof = s_tof + (size_t)warp * E;
    int *w_khit = s_khit + (size_t)warp * E;
    const int nchunks = (K + 31) >> 5;
    const unsigned FULL = 0xffffffffu;
    for (;;) {  // per-warp work fetch (the CTA-level fetch of the FP64 kernel measured 5 % slower here)
        unsigned long long g64 = 0;
        if (lane == 0) g64 = atomicAdd(T.counter, 1ull);
        g64 = __shfl_sync(FULL, g64, 0);
        if (g64 >= (unsigned long long)T.G * (unsigned long long)T.s_count) break;
        const long long w = (long long)g64;
        const long long g = w % T.G;
        const size_t sk = (size_t)(T.s_begin + w / T.G) * K;
        const float r_outer = (float)T.radius[g], pipe_offset = (float)T.offset[g];
        const float rsq = r_outer * r_outer, offsq = pipe_offset * pipe_offset;
        for (int e = lane; e < E; e += 32) { w_khit[e] = kNoHit; w_tof[e] = 0.0f; }
        __syncwarp();
        for (int c = 0; c < nchunks; ++c) {
            const int k = (c << 5) + lane;
            const bool valid = k < K;
            const int kk = valid ? k : K - 1;
            const float a_pq = __ldg(S.a_pq + sk + kk), b_pq = __ldg(S.b_pq + sk + kk);
            const float cpq = __ldg(S.cpq + sk + kk), spq = __ldg(S.spq + sk + kk);
            const float z_f = __ldg(S.z_f + kk);
            float x_q, z_q;
            {
                float A2 = a_pq * a_pq + 1;
                float B2 = 2 * (a_pq * b_pq - pipe_offset);
                float C2 = offsq + b_pq * b_pq - rsq;
                float sd = fsqrt_fast(B2 * B2 - 4 * A2 * C2);
                float den = 2 * A2;
                float x_q1 = __fdividef(-B2 + sd, den), x_q2 = __fdividef(-B2 - sd, den);
                float z_q1 = a_pq * x_q1 + b_pq, z_q2 = a_pq * x_q2 + b_pq;
                bool up = z_q1 > z_q2;
                x_q = up ? x_q1 : x_q2;
                z_q = up ? z_q1 : z_q2;
            }
            float cl, sl, m, b;  // reflected direction r^2 (cos, sin): the vector form of tofrt_raymath.h, in float
            {
                const float xx = x_q * x_q, ss = rsq - xx;
                const float s = fsqrt_fast(ss);
                const float c2 = ss - xx, s2 = (-2.0f * x_q) * s;
                cl = fmaf(c2, cpq, s2 * spq);
                sl = fmaf(s2, cpq, -(c2 * spq));
                m = __fdividef(sl, cl);
                b = z_q - m * x_q;
            }
            const bool fin = valid && fabsf(m) < 1e18f && fabsf(b) < 1e18f;
            int idx = -1;
            bool need_scan = fin, check_on_sample = false;
            if (S.shape != 0) {  // the guided search of the FP64 kernel, in float (tol = 2^-19 (...))
                const float sh = (float)S.shape;
                const float tol = (fabsf(m) * S.X + fabsf(b) + S.Z) * 1.9073486328125e-06f /* 2^-19 */ + 1e-30f;
                const float g0 = sh * line_diff_f(s_lens[0], m, b), gN = sh * line_diff_f(s_lens[N - 1], m, b);
                const bool pos0 = g0 > tol, neg0 = g0 < -tol;
                bool bracket = false, none = false;
                float edge = 0.f;
                if (fin && pos0 && gN > tol) { none = true; edge = fminf(g0, gN) - tol; }
                else if (fin && (pos0 || neg0)) bracket = true;
                const float sgn = pos0 ? sh : -sh;
                const float ms_eff = pos0 ? -INFINITY : sh * m;
                int lo = 0, hi = N - 1;
                for (int it = 0; it < S.log2n; ++it) {
                    const int mid = (lo + hi) >> 1;
                    const bool cond = (sgn * line_diff_f(s_lens[mid], m, b) > tol) && (sh * s_seg[mid].x > ms_eff);
                    const bool act = bracket && (hi - lo > 1);
                    lo = (act && cond) ? mid : lo;
                    hi = (act && !cond) ? mid : hi;
                }
                if (bracket && hi - lo > 1) bracket = false;
                if (bracket && neg0) {
                    const float ghi = sh * line_diff_f(s_lens[hi], m, b), glo = sh * line_diff_f(s_lens[lo], m, b);
                    if (ghi < -tol) { bracket = false; none = true; edge = -fmaxf(ghi, glo) - tol; }
                }
                if (bracket) {
                    int cprev = sign_class_f(line_diff_f(s_lens[lo], m, b));
                    int j = lo;
                    for (; j + 1 < N; ++j) {
                        int cnext = sign_class_f(line_diff_f(s_lens[j + 1], m, b));
                        if (class_change(cprev, cnext)) break;
                        cprev = cnext;
                    }
                    if (j + 1 < N) idx = j;
                    need_scan = false;
                    check_on_sample = idx < 0;
                } else if (none) {
                    need_scan = false;
                    check_on_sample = !(edge > 1e-8f);
                }
            }
            if (valid && need_scan) { idx = scan_exact_from_f(s_lens, N, m, b, 0); check_on_sample = idx < 0; }
            else if (valid && !fin) idx = scan_exact_from_f(s_lens, N, m, b, 0);
            float ix = NAN, iz = NAN;
            if (idx >= 0) {
                float2 p1 = s_lens[idx], p2 = s_lens[idx + 1], sg = s_seg[idx];
                if (np_isclose_f(p1.x, p2.x, 1e-5f, 1e-8f)) {
                    float y_int = m * p1.x + b;
                    if (y_int >= fminf(p1.y, p2.y) - 1e-9f && y_int <= fmaxf(p1.y, p2.y) + 1e-9f) { ix = p1.x; iz = y_int; }
                } else if (np_isclose_f(m, sg.x, 1e-5f, 1e-8f)) {
                    if (np_isclose_f(b, sg.y, 1e-5f, 1e-8f)) { ix = (p1.x + p2.x) / 2.0f; iz = m * ix + b; }
                } else {
                    float x_int = __fdividef(sg.y - b, m - sg.x), y_int = m * x_int + b;
                    if (x_int >= fminf(p1.x, p2.x) - 1e-9f && x_int <= fmaxf(p1.x, p2.x) + 1e-9f &&
                        y_int >= fminf(p1.y, p2.y) - 1e-9f && y_int <= fmaxf(p1.y, p2.y) + 1e-9f) { ix = x_int; iz = y_int; }
                }
            } else if (valid && check_on_sample) {
                for (int j = 0; j < N; ++j)
                    if (fabsf(line_diff_f(s_lens[j], m, b)) <= 1e-8f) { ix = s_lens[j].x; iz = s_lens[j].y; break; }
            }
            float target_x = NAN;
            if (__any_sync(FULL, valid && ix == ix)) {
                // sin/cos of atan2(ix, iz) without the angle, as in the FP64 kernel
                float sn, cs;
                {
                    const float r2 = fmaf(ix, ix, iz * iz);
                    float inv;
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(r2));
                    sn = ix * inv;
                    cs = iz * inv;
                }
                float hh, dh;
                if (S.surface == 0) {
                    float B = S.twod * cs - S.K2;
                    float sd = fsqrt_fast(B * B - S.fourAC);
                    hh = __fdividef(-B - sd, S.twoA);
                    float dB = -S.twod * sn;
                    float dS = __fdividef(1.0f, 2 * sd) * (2 * B * dB);
                    dh = S.phi1 * (dB + dS);
                } else {
                    hh = __fdividef(S.h0, cs);
                    dh = __fdividef(S.h0 * sn, cs * cs);
                }
                float dz = dh * cs - hh * sn, dx = dh * sn + hh * cs;
                float ihyp;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ihyp) : "f"(fmaf(dx, dx, dz * dz)));
                float s_arg = S.c1_over_c2 * (-fmaf(cl, dx, sl * dz) * ihyp * __fdividef(1.0f, rsq));
                if (__any_sync(FULL, valid && fabsf(s_arg) <= 1.0f)) {
                    float ct = fsqrt_fast((1.0f - s_arg) * (1.0f + s_arg));   // NaN on total internal reflection
                    float a_i = __fdividef(fmaf(dz, s_arg, -(dx * ct)), fmaf(dz, ct, dx * s_arg));
                    float b_i = iz - a_i * ix;
                    target_x = __fdividef(z_f - b_i, a_i);
                }
            }
            bool cand = valid && target_x >= S.rx_min - S.tol_max && target_x <= S.rx_max + S.tol_max;
            bool special = valid && S.has_nonfinite_rx && isinf(target_x);
            if (__any_sync(FULL, cand || special)) {
                int lo = 0, hi = 0;
                bool mine = false;
                if (cand) {
                    float tl = target_x - S.tol_max, th = target_x + S.tol_max;
                    int l = 0, r = S.E_sorted;
                    while (l < r) { int mid = (l + r) >> 1; if (s_xs[mid] < tl) l = mid + 1; else r = mid; }
                    lo = hi = l;
                    while (hi < S.E_sorted && s_xs[hi] <= th) ++hi;
                    for (int i = lo; i < hi; ++i)
                        if (np_isclose_f(target_x, s_xs[i], S.rtol, S.atol)) { atomicMin(&w_khit[s_perm[i]], k); mine = true; }
                }
                if (special)
                    for (int e = 0; e < E; ++e)
                        if (target_x == S.x_rx[e]) { atomicMin(&w_khit[e], k); mine = true; }
                __syncwarp();
                if (mine) {
                    float2 p = s_lens[kk];
                    float tof = S.leg1[sk + kk];
                    tof += dist2d_f(p.x, p.y, x_q, z_q) / S.c2;
                    tof += dist2d_f(x_q, z_q, ix, iz) / S.c2;
                    tof += dist2d_f(ix, iz, target_x, z_f) / S.c1;
                    for (int i = lo; i < hi; ++i) { int e = s_perm[i]; if (w_khit[e] == k) w_tof[e] = tof; }
                    if (special) for (int e = 0; e < E; ++e) if (w_khit[e] == k) w_tof[e] = tof;
                }
                __syncwarp();
            }
        }
        __syncwarp();
        for (int e = lane; e < E; e += 32) {
            int kh = w_khit[e];
            T.tof[(size_t)w * E + e] = w_tof[e];
            T.k_hit[(size_t)w * E + e] = kh == kNoHit ? -1 : kh;
        }
        __syncwarp();
    }
}

// =====================================================================================
// Two-point Fermat solve (SURVEY.md row f4 -- BASELINE.json north_star's literal wording).
// For every (array element e, focal point F) pair: the lens entry point P(alpha) that minimises
//     t(alpha) = |A_e - P(alpha)| / c1 + |P(alpha) - F| / c2
// over the lens surface of the current setup -- discrete argmin over the N lens samples, then a
// safeguarded secant (Illinois) iteration on dt/dalpha inside the two neighbouring sample
// intervals, using the reference's own h(alpha), dh/dalpha (main_rt.py:180-234).  The reference
// has no such function: parity is unpinned; the checker is a NumPy/SciPy oracle plus analytic
// known answers (tests/test_gpu_fermat.py).
// One CTA works on (element e, 256 focal points) items; the lens samples and the element's
// |A_e - P_j|/c1 row are staged in shared memory and broadcast to the 256 threads.
// =====================================================================================
constexpr int kFermatThreads = 256;

struct FermatArgs {
    const double *fx, *fz;   // [F] focal points
    long long F;
    const double *d1;        // [E][N]  |A_e - P_j| / c1
    const double *alpha;     // [N]
    double *tof;             // [E][F]
    int *kmin;               // [E][F]
    double *alpha_out;       // [E][F] or null
    double z_el;             // array line
    unsigned long long *counter;
};

__global__ void tofrt_fermat_d1_kernel(DevSetup S, double z_el, double *d1)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int e = blockIdx.y;
    if (j >= S.K) return;
    double2 p = S.lens[j];
    d1[(size_t)e * S.K + j] = dist2d(S.x_rx[e], z_el, p.x, p.y) / S.c1;
}

// t(alpha) and dt/dalpha at a continuous lens angle
__device__ __forceinline__ void fermat_eval(const DevSetup &S, double al, double ax, double az, double fx, double fz,
                                            double *t, double *dt)
{
    double sn, cs, h, dz, dx;
    tofrt_sincos(al, &sn, &cs);
    lens_point(S, sn, cs, &h, &dz, &dx);
    double px = h * sn, pz = h * cs;
    double u1x = px - ax, u1z = pz - az, u2x = px - fx, u2z = pz - fz;
    double r1 = sqrt(u1x * u1x + u1z * u1z), r2 = sqrt(u2x * u2x + u2z * u2z);
    *t = r1 / S.c1 + r2 / S.c2;
    *dt = (u1x * dx + u1z * dz) / (S.c1 * r1) + (u2x * dx + u2z * dz) / (S.c2 * r2);
}

__global__ void __launch_bounds__(kFermatThreads) tofrt_fermat_kernel(DevSetup S, FermatArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = S.K;
    double2 *s_lens = reinterpret_cast<double2 *>(smem_raw);
    double *s_d1 = reinterpret_cast<double *>(s_lens + N);
    float2 *s_lens32 = reinterpret_cast<float2 *>(s_d1 + N);  // float copies for the screening pass
    float *s_d132 = reinterpret_cast<float *>(s_lens32 + N);
    __shared__ unsigned long long s_item;
    __shared__ float s_d1max;
    for (int i = threadIdx.x; i < N; i += kFermatThreads) {
        const double2 p = S.lens[i];
        s_lens[i] = p;
        s_lens32[i] = make_float2((float)p.x, (float)p.y);
    }
    const long long tiles = (A.F + kFermatThreads - 1) / kFermatThreads;
    const unsigned long long items = (unsigned long long)tiles * S.E;
    const double inv_c2 = 1.0 / S.c2;
    int cur_e = -1;
    for (;;) {
        __syncthreads();  // previous item done with s_d1 / s_item
        if (threadIdx.x == 0) s_item = atomicAdd(A.counter, 1ull);
        __syncthreads();
        const unsigned long long item = s_item;
        if (item >= items) break;
        const int e = (int)(item / tiles);
        const long long f = (long long)(item % tiles) * kFermatThreads + threadIdx.x;
        if (e != cur_e) {  // uniform across the CTA
            if (threadIdx.x == 0) s_d1max = 0.f;
            __syncthreads();
            float mx = 0.f;
            for (int i = threadIdx.x; i < N; i += kFermatThreads) {
                const double v = A.d1[(size_t)e * N + i];
                s_d1[i] = v;
                s_d132[i] = (float)v;
                mx = fmaxf(mx, (float)fabs(v));
            }
            atomicMax(reinterpret_cast<int *>(&s_d1max), __float_as_int(mx));  // non-negative floats order as ints
            cur_e = e;
            __syncthreads();
        }
        if (f >= A.F) continue;
        const double fx = A.fx[f], fz = A.fz[f];
        const double ax = S.x_rx[e], az = A.z_el;
        // ---- discrete argmin over the sampled surface (first minimum on ties) ----
        // Exact FP64 argmin, found with an FP32 screening pass: t32_j differs from t_j by less than
        // delta = 2^-19 (D/c2 + max_j d1_j), D = X + Z + |fx| + |fz| (float rounding of the four
        // coordinates, of the squared distance, the approximate square root and the fma; >2.5x margin),
        // so the exact minimiser is among the samples with t32_j <= min_j t32_j + 2 delta -- a handful
        // around the minimum -- and only those are evaluated in FP64, in index order.
        double best = INFINITY;
        int jbest = 0;
        {
            const float fx32 = (float)fx, fz32 = (float)fz, ic2 = (float)inv_c2;
            float min32 = INFINITY;
#pragma unroll 4
            for (int j = 0; j < N; ++j) {
                const float2 p = s_lens32[j];
                const float ux = p.x - fx32, uz = p.y - fz32;
                min32 = fminf(min32, fmaf(fsqrt_fast(fmaf(ux, ux, uz * uz)), ic2, s_d132[j]));
            }
            const float D = (float)(S.X + S.Z + fabs(fx) + fabs(fz));
            const float thr = min32 + 2.0f * 1.9073486328125e-06f * (D * ic2 + s_d1max) * 1.0001f;
            const bool screened = thr == thr && thr < INFINITY;  // non-finite inputs: evaluate everything exactly
#pragma unroll 2
            for (int j = 0; j < N; ++j) {
                const float2 p = s_lens32[j];
                const float ux32 = p.x - fx32, uz32 = p.y - fz32;
                const float t32 = fmaf(fsqrt_fast(fmaf(ux32, ux32, uz32 * uz32)), ic2, s_d132[j]);
                if (!screened || t32 <= thr) {
                    const double2 q = s_lens[j];
                    const double ux = q.x - fx, uz = q.y - fz;
                    const double t = fma(sqrt(fma(ux, ux, uz * uz)), inv_c2, s_d1[j]);
                    if (t < best) { best = t; jbest = j; }
                }
            }
        }
        // ---- continuous refinement between the neighbouring samples ----
        double a = A.alpha[max(jbest - 1, 0)], b = A.alpha[min(jbest + 1, N - 1)];
        double ta, ga, tb, gb, tc, gc;
        fermat_eval(S, a, ax, az, fx, fz, &ta, &ga);
        fermat_eval(S, b, ax, az, fx, fz, &tb, &gb);
        double al = A.alpha[jbest], tof;
        fermat_eval(S, al, ax, az, fx, fz, &tof, &gc);
        if (ga < 0.0 && gb > 0.0) {  // a bracketed stationary point: Illinois false position on dt/dalpha
            int side = 0;
            for (int it = 0; it < 8; ++it) {
                double c = (a * gb - b * ga) / (gb - ga);
                if (!(c > a && c < b)) c = 0.5 * (a + b);
                fermat_eval(S, c, ax, az, fx, fz, &tc, &gc);
                al = c; tof = tc;
                if (gc == 0.0) break;
                if (gc > 0.0) { b = c; gb = gc; if (side == 1) ga *= 0.5; side = 1; }
                else { a = c; ga = gc; if (side == -1) gb *= 0.5; side = -1; }
            }
        }
        const size_t o = (size_t)e * A.F + f;
        A.tof[o] = tof;
        A.kmin[o] = jbest;
        if (A.alpha_out) A.alpha_out[o] = al;
    }
}

// =====================================================================================
// 3-D two-point Fermat solve (SURVEY.md row f4 / BASELINE config 5: matrix array, 3-D volume,
// curved surface).  The reference is 2-D: parity unpinned.  The lens is the surface of revolution
// of the reference's profile h(alpha); the per-pair arithmetic lives in tofrt_fermat3d.h (shared
// with the oracle's twin).  For every (matrix-array element e, voxel F) pair:
//   1. discrete search over a coarse n_side x n_side grid of surface samples: one thread per pair,
//      the element's (P_s, |A_e - P_s|/c1) table broadcast from shared memory, FP32 arithmetic
//      (the argmin only selects the start of step 2);
//   2. projected Newton refinement in FP64, state in registers.
// One CTA works on (element, 256 voxels) items; outputs are element-major tof[e][f] so a CTA
// writes 2 KB runs.
// =====================================================================================
constexpr int kF3dThreads = 256;
#ifndef TOFRT_F3D_MIN_BLOCKS
#define TOFRT_F3D_MIN_BLOCKS 3
#endif

struct F3dArgs {
    const double *fx, *fy, *fz;  // [F] voxels
    long long F;
    const float4 *smp;           // [E][ns] (P_s.x, P_s.y, P_s.z, |A_e - P_s| / c1)
    const double *ang;           // [n_side]
    const double *ex, *ey;       // [E] element positions on the plane z = z_el
    int E, n_side;
    double z_el, cell;
    double *tof;                 // [E][F]
    int *kmin;                   // [E][F] or null
    double *ab;                  // [E][F][2] or null
    unsigned long long *counter;
};

__global__ void tofrt_f3d_prepare_kernel(tofrt_surf3d S, F3dArgs A, float4 *smp)
{
    const int ns = A.n_side * A.n_side;
    const int s = blockIdx.x * blockDim.x + threadIdx.x, e = blockIdx.y;
    if (s >= ns) return;
    double P[3];
    tofrt_f3d_point(&S, A.ang[s / A.n_side], A.ang[s % A.n_side], P);
    const double ux = P[0] - A.ex[e], uy = P[1] - A.ey[e], uz = P[2] - A.z_el;
    const double d1 = sqrt(ux * ux + uz * uz + uy * uy) * S.inv_c1;
    smp[(size_t)e * ns + s] = make_float4((float)P[0], (float)P[1], (float)P[2], (float)d1);
}

__global__ void __launch_bounds__(kF3dThreads, TOFRT_F3D_MIN_BLOCKS) tofrt_fermat3d_kernel(tofrt_surf3d S, F3dArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_smp = reinterpret_cast<float4 *>(smem_raw);
    const int ns = A.n_side * A.n_side;
    const long long tiles = (A.F + kF3dThreads - 1) / kF3dThreads;
    const unsigned long long items = (unsigned long long)tiles * A.E;
    const float ic2 = (float)S.inv_c2;
    int cur_e = -1;
    // Dynamic work counter, one (element, 256 voxels) item at a time.  (Dealing the items round-robin
    // without the two barriers measured 4 % slower: the barrier keeps a CTA's warps in step.)
    __shared__ unsigned long long s_item;
    for (;;) {
        __syncthreads();  // previous item done with s_smp / s_item
        if (threadIdx.x == 0) s_item = atomicAdd(A.counter, 1ull);
        __syncthreads();
        const unsigned long long item = s_item;
        if (item >= items) break;
        const int e = (int)(item / tiles);
        const long long f = (long long)(item % tiles) * kF3dThreads + threadIdx.x;
        if (e != cur_e) {  // uniform across the CTA; items are element-major, so this is rare
            for (int i = threadIdx.x; i < ns; i += kF3dThreads) s_smp[i] = A.smp[(size_t)e * ns + i];
            cur_e = e;
            __syncthreads();
        }
        if (f >= A.F) continue;
        const double Fp[3] = {A.fx[f], A.fy[f], A.fz[f]};
        const double Ap[3] = {A.ex[e], A.ey[e], A.z_el};
        // ---- discrete search over the sampled surface (first minimum on ties) ----
        int kb = 0;
        {
            const float gx = (float)Fp[0], gy = (float)Fp[1], gz = (float)Fp[2];
            float best = INFINITY;
#pragma unroll 4
            for (int s = 0; s < ns; ++s) {
                const float4 p = s_smp[s];
                const float ux = p.x - gx, uy = p.y - gy, uz = p.z - gz;
                const float t = fmaf(fsqrt_fast_ftz(fmaf(ux, ux, fmaf(uy, uy, uz * uz))), ic2, p.w);
                const bool lt = t < best;
                best = lt ? t : best;
                kb = lt ? s : kb;
            }
        }
        // ---- continuous refinement ----
        const int ia = kb / A.n_side, ib = kb - ia * A.n_side;
        double t, a, b;
        tofrt_f3d_refine(&S, A.ang[ia], A.ang[ib], A.cell, Ap, Fp, &t, &a, &b);
        const size_t o = (size_t)e * A.F + f;
        A.tof[o] = t;
        if (A.kmin) A.kmin[o] = kb;
        if (A.ab) reinterpret_cast<double2 *>(A.ab)[o] = make_double2(a, b);
    }
}

// ---- sparse form of the table: only the hits, ordered by (row, receiver) -------------------
// 97-98 % of the table entries are misses; a consumer on the host that only wants the hits gets
// them as a list, which cuts the device->host traffic ~30x (the dense D2H is what limits the
// end-to-end rate once 8 GPUs share the host's PCIe/memory bandwidth).  Deterministic: count per
// row, exclusive scan, ordered compaction (no atomics).
__global__ void tofrt_count_hits_kernel(const int *k_hit, long long rows, int E, long long *cnt)
{
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    int n = 0;
    for (int e = lane; e < E; e += 32) n += k_hit[row * E + e] >= 0;
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) cnt[row] = n;
}

__global__ void tofrt_compact_hits_kernel(const int *k_hit, const double *tof, long long rows, int E,
                                          const long long *off, long long row0, long long base, long long capacity,
                                          int *hit_g, int *hit_e, int *hit_k, double *hit_tof)
{
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    long long o = base + off[row];
    for (int e0 = 0; e0 < E; e0 += 32) {
        const int e = e0 + lane;
        const int kh = e < E ? k_hit[row * E + e] : -1;
        const unsigned m = __ballot_sync(0xffffffffu, kh >= 0);
        if (kh >= 0) {
            const long long i = o + __popc(m & ((1u << lane) - 1u));
            if (i < capacity) {
                hit_g[i] = (int)(row0 + row); hit_e[i] = e; hit_k[i] = kh; hit_tof[i] = tof[row * E + e];
            }
        }
        o += __popc(m);
    }
}

// ---- FP64 issue-rate micro-benchmark (the roofline denominator, measured on the box) ----
// 8 independent dependency chains per thread; MODE 0: DFMA, MODE 1: DMUL+DADD (un-fused, as
// the scan issues them).  Each loop iteration issues 16 FP64 instructions per thread.
template <int MODE>
__global__ void __launch_bounds__(256) tofrt_fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (double)(threadIdx.x + i) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { v[i] = fma(v[i], a, b); v[i] = fma(v[i], a, b); }
            else { v[i] = __dmul_rn(v[i], a); v[i] = __dadd_rn(v[i], b); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 12345.678) out[0] = s;  // keep the chains alive
}

// ---- host side ----------------------------------------------------------------------

thread_local std::string g_err;

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

typedef void (*table_kernel_t)(DevSetup, TableArgs);
// the instantiation for (first-hit state in global memory, guided search, flat surface, vector math)
table_kernel_t pick_table_kernel(bool hg, bool guided, bool flat, bool vec)
{
    if (flat) return vec ? tofrt_table_kernel<false, false, true, true> : tofrt_table_kernel<false, false, true, false>;
    if (hg) {
        if (guided) return vec ? tofrt_table_kernel<true, true, false, true> : tofrt_table_kernel<true, true, false, false>;
        return vec ? tofrt_table_kernel<true, false, false, true> : tofrt_table_kernel<true, false, false, false>;
    }
    if (guided) return vec ? tofrt_table_kernel<false, true, false, true> : tofrt_table_kernel<false, true, false, false>;
    return vec ? tofrt_table_kernel<false, false, false, true> : tofrt_table_kernel<false, false, false, false>;
}

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(TOFRT_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
    } while (0)

}  // namespace

struct tofrt_ctx {
    int device = 0;
    int sm_count = 0;
    bool ready = false;
    DevSetup S{};
    int K = 0, E = 0;
    double params_d = 0.0;   // array line z = d
    size_t smem = 0, smem_guided = 0;   // dynamic shared memory of the scan / guided instantiations
    int ctas_per_sm = 1;
    // device buffers owned by the context
    double *d_alpha = nullptr, *d_zf = nullptr, *d_a = nullptr, *d_b = nullptr, *d_phi = nullptr, *d_leg1 = nullptr;
    double *d_cs = nullptr;   // cos / sin of phi_pq, [2][S*K]
    double2 *d_lens = nullptr, *d_seg = nullptr;
    int opt_guided = 1;
    int opt_surface = 0;
    int opt_vector = TOFRT_MATH_DEFAULT;
    float *d_f32 = nullptr;   // FP32 copies of the tables (FP32 mode)
    DevSetupF SF{};
    size_t smem_f32 = 0;
    double *d_xrx = nullptr, *d_xs = nullptr, *d_sh = nullptr;
    int n_shooters = 1;
    int *d_perm = nullptr, *d_bad = nullptr, *d_bucket = nullptr;
    unsigned long long *d_counter = nullptr;
    // host-buffer pipeline (tofrt_table_host)
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    void *ws[2] = {nullptr, nullptr};
    size_t ws_bytes = 0, in_cap = 0;
    double *d_in = nullptr;
    // tofrt_hits_host workspace (grow-only): dense intermediate, counts/offsets, scan scratch, hit list
    void *hw[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t hw_bytes[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    double *d_d1 = nullptr;   // Fermat solver: |A_e - P_j|/c1, built lazily per setup
    // 3-D two-point solve (tofrt_fermat3d_setup)
    bool f3_ready = false;
    tofrt_surf3d f3_S{};
    int f3_E = 0, f3_n = 0;
    double f3_cell = 0.0;
    double *f3_geo = nullptr;   // ang[n] ex[E] ey[E]
    float4 *f3_smp = nullptr;   // [E][n*n]
    void free_all()
    {
        cudaFree(d_alpha); cudaFree(d_zf); cudaFree(d_a); cudaFree(d_b); cudaFree(d_phi); cudaFree(d_leg1);
        cudaFree(d_lens); cudaFree(d_seg); cudaFree(d_f32); cudaFree(d_sh); cudaFree(d_xrx); cudaFree(d_xs); cudaFree(d_perm); cudaFree(d_bad);
        cudaFree(d_bucket); d_bucket = nullptr;
        cudaFree(d_cs); d_cs = nullptr;
        d_alpha = d_zf = d_a = d_b = d_phi = d_leg1 = d_xrx = d_xs = nullptr;
        d_lens = d_seg = nullptr; d_perm = d_bad = nullptr; d_f32 = nullptr; d_sh = nullptr;
    }
};

extern "C" {

int tofrt_version(void) { return TOFRT_VERSION; }
const char *tofrt_last_error(void) { return g_err.c_str(); }

int tofrt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int tofrt_create(tofrt_ctx **out, int device)
{
    if (!out) return fail(TOFRT_E_BADARG, "tofrt_create: ctx is NULL");
    *out = nullptr;
    CUDA_TRY(cudaSetDevice(device));
    tofrt_ctx *c = new tofrt_ctx();
    c->device = device;
    cudaError_t e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_counter, 5 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_bad, sizeof(int));
    if (e != cudaSuccess) {
        delete c;
        return fail(TOFRT_E_CUDA, std::string("tofrt_create: ") + cudaGetErrorString(e));
    }
    *out = c;
    return TOFRT_OK;
}

int tofrt_destroy(tofrt_ctx *c)
{
    if (!c) return TOFRT_OK;
    cudaSetDevice(c->device);
    c->free_all();
    cudaFree(c->d_counter);
    cudaFree(c->ws[0]); cudaFree(c->ws[1]); cudaFree(c->d_in); cudaFree(c->d_d1);
    cudaFree(c->f3_geo); cudaFree(c->f3_smp);
    for (int i = 0; i < 10; ++i) cudaFree(c->hw[i]);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
        if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]);
    }
    if (c->s_compute) cudaStreamDestroy(c->s_compute);
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
    delete c;
    return TOFRT_OK;
}

int tofrt_setup(tofrt_ctx *c, const tofrt_params *p, const double *alpha, const double *z_f, int K,
                double x_shooter, double z_shooter, const double *x_rx, int E, double rtol, double atol)
{
    return tofrt_setup_shooters(c, p, alpha, z_f, K, &x_shooter, &z_shooter, 1, x_rx, E, rtol, atol);
}

int tofrt_setup_shooters(tofrt_ctx *c, const tofrt_params *p, const double *alpha, const double *z_f, int K,
                         const double *x_shooters, const double *z_shooters, int n_shooters, const double *x_rx,
                         int E, double rtol, double atol)
{
    if (!c || !p || !alpha || !z_f || !x_rx || !x_shooters || !z_shooters)
        return fail(TOFRT_E_BADARG, "tofrt_setup: NULL argument");
    if (K < 2) return fail(TOFRT_E_TOOFEW, "Curve needs at least two points.");
    if (E < 1) return fail(TOFRT_E_BADARG, "tofrt_setup: need at least one receiver");
    if (n_shooters < 1) return fail(TOFRT_E_BADARG, "tofrt_setup: need at least one shooter");
    const int NS = n_shooters;
    CUDA_TRY(cudaSetDevice(c->device));
    c->ready = false;
    c->free_all();
    cudaFree(c->d_d1); c->d_d1 = nullptr;
    c->f3_ready = false;
    CUDA_TRY(cudaMalloc(&c->d_bad, sizeof(int)));

    DevSetup S{};
    S.c1 = p->c1; S.c2 = p->c2; S.d = p->d;
    // main_rt.py:181-185 and :199-201, same operation order as NumPy
    const double T = p->l0 / p->c1 + p->h0 / p->c2;
    const double c1sq = p->c1 * p->c1, c2sq = p->c2 * p->c2;
    S.A = (c1sq / c2sq) - 1;
    const double C = (c1sq * (T * T)) - (p->d * p->d);
    S.fourAC = 4 * S.A * C;
    S.twoA = 2 * S.A;
    S.twod = 2 * p->d;
    S.K2 = 2 * T * c1sq / p->c2;
    S.phi1 = -1 / (2 * S.A);
    S.c2_over_c1 = p->c2 / p->c1;
    S.c1_over_c2 = p->c1 / p->c2;
    S.n_shooters = NS;
    S.surface = c->opt_surface; S.h0 = p->h0;
    S.rtol = rtol; S.atol = atol;
    S.K = K; S.E = E;
    S.Npad = 1 + ((K - 1 + kChunk - 1) / kChunk) * kChunk;

    // receivers: sorted finite copy for the candidate search
    std::vector<int> order;
    double amax = 0.0;
    int nonfinite = 0;
    for (int e = 0; e < E; ++e) {
        if (std::isfinite(x_rx[e])) { order.push_back(e); amax = std::max(amax, std::fabs(x_rx[e])); }
        else nonfinite = 1;
    }
    std::stable_sort(order.begin(), order.end(), [&](int i, int j) { return x_rx[i] < x_rx[j]; });
    std::vector<double> xs(order.size());
    for (size_t i = 0; i < order.size(); ++i) xs[i] = x_rx[order[i]];
    S.E_sorted = (int)order.size();
    S.has_nonfinite_rx = nonfinite;
    // upper bound of atol + rtol*|x_e| with slack for the rounding of target_x -+ tol_max
    S.tol_max = (std::fabs(atol) + std::fabs(rtol) * amax) * 1.0000001 + 1e-300;
    if (!std::isfinite(S.tol_max)) S.tol_max = 1.7e308;
    S.rx_min = xs.empty() ? 1.0 : xs.front();
    S.rx_max = xs.empty() ? -1.0 : xs.back();

    // bucket table for the candidate lookup of the table assembly: nb uniform cells over the finite
    // receiver range (+- tol_max); bucket[c] = first sorted receiver >= start of cell c-1, so a lookup
    // that is off by one cell through rounding still starts at or before the wanted receiver
    std::vector<int> bucket;
    S.nb = 0; S.bx0 = 0.0; S.binv = 0.0;
#ifndef TOFRT_NO_BUCKET  /* tuning knob: binary search instead */
    if (S.E_sorted >= 8) {
        const double x0 = xs.front() - S.tol_max, range = (xs.back() + S.tol_max) - x0;
        const int nb = std::min(4 * S.E_sorted, 8192);
        const double cell = range / nb, binv = 1.0 / cell;
        if (std::isfinite(range) && range > 0.0 && std::isfinite(binv) && cell > 0.0) {
            bucket.resize(nb);
            int i = 0;
            for (int cidx = 0; cidx < nb; ++cidx) {
                const double start = x0 + (cidx - 1) * cell;
                while (i < S.E_sorted && xs[i] < start) ++i;
                bucket[cidx] = i;
            }
            S.nb = nb; S.bx0 = x0; S.binv = binv;
        }
    }
#endif
    const size_t smem_tables = (size_t)(S.Npad + K) * sizeof(double2) +
                               (size_t)S.E_sorted * (sizeof(double) + sizeof(int)) + (size_t)S.nb * sizeof(int) + 16;
    size_t smem = smem_tables + (size_t)kWarps * E * (sizeof(double) + sizeof(int));
    size_t smem_guided = smem_tables + (size_t)(kThreadsGuided / 32) * E * (sizeof(double) + sizeof(int));
    S.hits_in_global = 0;
    if (std::max(smem, smem_guided) > 227 * 1024) {  // many receivers: the per-warp first-hit state moves to the output rows
        S.hits_in_global = 1;
        smem = smem_guided = smem_tables;
    }
    if (smem > 227 * 1024)
        return fail(TOFRT_E_TOOLARGE, "tofrt_setup: K/E need " + std::to_string(smem) + " B of shared memory (> 227 KB)");

    CUDA_TRY(cudaMalloc(&c->d_alpha, sizeof(double) * K));
    CUDA_TRY(cudaMalloc(&c->d_zf, sizeof(double) * K));
    CUDA_TRY(cudaMalloc(&c->d_a, sizeof(double) * K * NS));
    CUDA_TRY(cudaMalloc(&c->d_b, sizeof(double) * K * NS));
    CUDA_TRY(cudaMalloc(&c->d_phi, sizeof(double) * K * NS));
    CUDA_TRY(cudaMalloc(&c->d_leg1, sizeof(double) * K * NS));
    CUDA_TRY(cudaMalloc(&c->d_cs, sizeof(double) * 2 * K * NS));
    CUDA_TRY(cudaMalloc(&c->d_lens, sizeof(double2) * S.Npad));
    CUDA_TRY(cudaMalloc(&c->d_seg, sizeof(double2) * K));
    CUDA_TRY(cudaMemset(c->d_seg, 0, sizeof(double2) * K));  // entry K-1 is padding (TMA copies K entries)
    CUDA_TRY(cudaMalloc(&c->d_xrx, sizeof(double) * E));
    CUDA_TRY(cudaMalloc(&c->d_sh, sizeof(double) * 2 * NS));
    CUDA_TRY(cudaMemcpy(c->d_sh, x_shooters, sizeof(double) * NS, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->d_sh + NS, z_shooters, sizeof(double) * NS, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&c->d_xs, sizeof(double) * std::max(1, S.E_sorted)));
    CUDA_TRY(cudaMalloc(&c->d_perm, sizeof(int) * std::max(1, S.E_sorted)));
    CUDA_TRY(cudaMemcpy(c->d_alpha, alpha, sizeof(double) * K, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->d_zf, z_f, sizeof(double) * K, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->d_xrx, x_rx, sizeof(double) * E, cudaMemcpyHostToDevice));
    if (S.E_sorted) {
        CUDA_TRY(cudaMemcpy(c->d_xs, xs.data(), sizeof(double) * S.E_sorted, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(c->d_perm, order.data(), sizeof(int) * S.E_sorted, cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaMemset(c->d_bad, 0, sizeof(int)));
    S.lens = c->d_lens; S.seg = c->d_seg; S.a_pq = c->d_a; S.b_pq = c->d_b; S.phi_pq = c->d_phi; S.z_f = c->d_zf;
    S.x_sh = c->d_sh; S.z_sh = c->d_sh + NS;
    S.cpq = c->d_cs; S.spq = c->d_cs + (size_t)K * NS;
    S.leg1 = c->d_leg1; S.x_rx = c->d_xrx; S.xs = c->d_xs; S.perm = c->d_perm;
    CUDA_TRY(cudaMalloc(&c->d_bucket, sizeof(int) * std::max(1, S.nb)));
    if (S.nb) CUDA_TRY(cudaMemcpy(c->d_bucket, bucket.data(), sizeof(int) * S.nb, cudaMemcpyHostToDevice));
    S.bucket = c->d_bucket;

    tofrt_prepare_kernel<<<(S.Npad + 127) / 128, 128>>>(S, c->d_alpha, c->d_lens, c->d_a, c->d_b, c->d_phi,
                                                        c->d_leg1, c->d_cs, c->d_cs + (size_t)K * NS, c->d_bad);
    CUDA_TRY(cudaGetLastError());
    tofrt_chord_kernel<<<(K + 127) / 128, 128>>>(K, c->d_lens, c->d_seg);
    CUDA_TRY(cudaGetLastError());
    int bad = 0;
    CUDA_TRY(cudaMemcpy(&bad, c->d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    S.tables_ok = bad ? 0 : 1;
    // Facts the guided search relies on, verified here on the actual table: abscissae strictly
    // increasing and chord slopes strictly monotone by more than their own rounding error
    // (=> the sampled lens is a strictly concave / convex polyline in exact arithmetic).
    {
        std::vector<double2> hl(K), hs(K);
        CUDA_TRY(cudaMemcpy(hl.data(), c->d_lens, sizeof(double2) * K, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(hs.data(), c->d_seg, sizeof(double2) * (K - 1), cudaMemcpyDeviceToHost));
        bool incr = true, dec = true, inc = true;
        double X = 0, Z = 0;
        for (int j = 0; j < K; ++j) {
            X = std::max(X, std::fabs(hl[j].x));
            Z = std::max(Z, std::fabs(hl[j].y));
            if (j + 1 < K && !(hl[j + 1].x > hl[j].x)) incr = false;
        }
        for (int j = 0; j + 2 < K; ++j) {
            double a = hs[j].x, b2 = hs[j + 1].x;
            double guard = 16 * 2.220446049250313e-16 * std::max(std::fabs(a), std::fabs(b2));
            if (!(a - b2 > guard)) dec = false;
            if (!(b2 - a > guard)) inc = false;
        }
        S.X = X; S.Z = Z;
        S.shape = (S.tables_ok && incr && K >= 3 && S.surface == 0) ? (dec ? 1 : (inc ? -1 : 0)) : 0;
        int lg = 1;
        while ((1 << lg) < K) ++lg;
        S.log2n = lg;  // ceil-halving an interval of length K-1 <= 2^lg reaches length 1 in lg steps
    }
    S.guided = c->opt_guided;
    S.vector_math = c->opt_vector;
    // ---- FP32 mode: the same tables rounded to float ----
    {
        const int Es = std::max(1, S.E_sorted);
        const size_t nf = (size_t)4 * K + (size_t)6 * K * NS + K + E + Es;  // lens(2K) seg(2K) [a,b,phi,cos,sin,leg1](S*K) zf x_rx xs
        CUDA_TRY(cudaMalloc(&c->d_f32, sizeof(float) * nf));
        float *f = c->d_f32;
        auto conv = [&](const double *src, int n) {
            float *dst = f;
            if (n > 0) tofrt_to_float_kernel<<<(n + 255) / 256, 256>>>(n, src, dst);
            f += n;
            return dst;
        };
        DevSetupF F{};
        F.lens = (const float2 *)conv((const double *)c->d_lens, 2 * K);
        F.seg = (const float2 *)conv((const double *)c->d_seg, 2 * K - 2); f += 2;
        F.a_pq = conv(c->d_a, K * NS); F.b_pq = conv(c->d_b, K * NS); F.phi_pq = conv(c->d_phi, K * NS);
        F.cpq = conv(c->d_cs, K * NS); F.spq = conv(c->d_cs + (size_t)K * NS, K * NS);
        F.z_f = conv(c->d_zf, K); F.leg1 = conv(c->d_leg1, K * NS);
        F.x_rx = conv(c->d_xrx, E); F.xs = conv(c->d_xs, S.E_sorted);
        CUDA_TRY(cudaGetLastError());
        F.perm = c->d_perm;
        F.c1 = (float)S.c1; F.c2 = (float)S.c2; F.fourAC = (float)S.fourAC; F.twoA = (float)S.twoA;
        F.twod = (float)S.twod; F.K2 = (float)S.K2; F.phi1 = (float)S.phi1; F.c1_over_c2 = (float)S.c1_over_c2;
        F.rtol = (float)S.rtol; F.atol = (float)S.atol;
        F.tol_max = (float)(S.tol_max * 1.001) + 1e-30f;
        F.rx_min = (float)S.rx_min; F.rx_max = (float)S.rx_max; F.X = (float)S.X * 1.0001f; F.Z = (float)S.Z * 1.0001f;
        F.K = K; F.E = E; F.E_sorted = S.E_sorted; F.has_nonfinite_rx = S.has_nonfinite_rx;
        F.shape = S.shape; F.log2n = S.log2n; F.n_shooters = NS; F.surface = S.surface; F.h0 = (float)S.h0;
        c->SF = F;
        c->smem_f32 = (size_t)2 * K * sizeof(float2) + (size_t)S.E_sorted * (sizeof(float) + sizeof(int)) +
                      (size_t)kWarps * E * (sizeof(float) + sizeof(int)) + 16;
        if (c->smem_f32 > 227 * 1024) c->smem_f32 = 0;  // FP32 mode keeps its first-hit state in shared memory only
        else
            CUDA_TRY(cudaFuncSetAttribute(tofrt_table_kernel_f32, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)c->smem_f32));
        CUDA_TRY(cudaDeviceSynchronize());
    }

    if (S.surface != 0 && S.hits_in_global)
        return fail(TOFRT_E_TOOLARGE, "tofrt_setup: the flat-interface extension supports the shared-memory receiver budget only");
    const bool flat = S.surface != 0, hg = S.hits_in_global != 0;
    for (int vec = 0; vec < 2; ++vec)
        for (int gd = 0; gd < (flat ? 1 : 2); ++gd)
            CUDA_TRY(cudaFuncSetAttribute(pick_table_kernel(hg, gd != 0, flat, vec != 0),
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(gd ? smem_guided : smem)));
    int occ = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pick_table_kernel(hg, !flat, flat, c->opt_vector != 0),
                                                             flat ? kThreads : kThreadsGuided, flat ? smem : smem_guided));
    c->ctas_per_sm = std::max(1, occ);
    c->smem_guided = smem_guided;
    c->S = S; c->K = K; c->E = E; c->smem = smem; c->n_shooters = NS; c->params_d = p->d;
    c->ready = true;
    return TOFRT_OK;
}

int tofrt_set_option(tofrt_ctx *c, int option, int value)
{
    if (!c) return fail(TOFRT_E_BADARG, "tofrt_set_option: ctx is NULL");
    if (option == TOFRT_OPT_SEARCH) {
        if (value != TOFRT_SEARCH_SCAN && value != TOFRT_SEARCH_GUIDED)
            return fail(TOFRT_E_BADARG, "tofrt_set_option: bad search mode");
        c->opt_guided = value;
        c->S.guided = value;
        return TOFRT_OK;
    }
    if (option == TOFRT_OPT_MATH) {
        if (value != TOFRT_MATH_ANGLE && value != TOFRT_MATH_VECTOR)
            return fail(TOFRT_E_BADARG, "tofrt_set_option: bad math mode");
        c->opt_vector = value;
        c->S.vector_math = value;
        return TOFRT_OK;
    }
    if (option == TOFRT_OPT_SURFACE) {
        if (value != TOFRT_SURFACE_LENS && value != TOFRT_SURFACE_FLAT)
            return fail(TOFRT_E_BADARG, "tofrt_set_option: bad surface type");
        if (value != c->opt_surface) c->ready = false;  // the tables belong to the other surface: set up again
        c->opt_surface = value;
        return TOFRT_OK;
    }
    return fail(TOFRT_E_BADARG, "tofrt_set_option: unknown option");
}

int tofrt_last_stats(tofrt_ctx *c, unsigned long long *stats4)
{
    if (!c || !stats4) return fail(TOFRT_E_BADARG, "tofrt_last_stats: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpy(stats4, c->d_counter + 1, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return TOFRT_OK;
}

int tofrt_get_info(tofrt_ctx *c, int what)
{
    if (!c) return TOFRT_E_BADARG;
    switch (what) {
    case TOFRT_INFO_SHAPE: return c->ready ? c->S.shape : 0;
    case TOFRT_INFO_SEARCH: return c->opt_guided;
    case TOFRT_INFO_MATH: return c->opt_vector;
    default: return TOFRT_E_BADARG;
    }
}

int tofrt_lens_host(tofrt_ctx *c, double *lens_x, double *lens_z)
{
    if (!c || !lens_x || !lens_z) return fail(TOFRT_E_BADARG, "tofrt_lens_host: NULL argument");
    if (!c->ready) return fail(TOFRT_E_NOSETUP, "tofrt_lens_host: call tofrt_setup first");
    CUDA_TRY(cudaSetDevice(c->device));
    std::vector<double2> h(c->K);
    CUDA_TRY(cudaMemcpy(h.data(), c->d_lens, sizeof(double2) * c->K, cudaMemcpyDeviceToHost));
    for (int k = 0; k < c->K; ++k) { lens_x[k] = h[k].x; lens_z[k] = h[k].y; }
    return TOFRT_OK;
}

static int table_dev_impl(tofrt_ctx *c, const double *d_radius, const double *d_offset, long long G, double *d_tof,
                          int *d_k_hit, const tofrt_ray_records *d_rec, void *stream, int s_begin, int s_count,
                          const double *inject_m = nullptr, const double *inject_b = nullptr)
{
    if (!c || G < 0) return fail(TOFRT_E_BADARG, "tofrt_table_dev: bad argument");
    if (!c->ready) return fail(TOFRT_E_NOSETUP, "tofrt_table_dev: call tofrt_setup first");
    if (G == 0) return TOFRT_OK;
    if (!d_radius || !d_offset || !d_tof || !d_k_hit) return fail(TOFRT_E_BADARG, "tofrt_table_dev: NULL buffer");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    TableArgs T{};
    T.radius = d_radius; T.offset = d_offset; T.G = G; T.tof = d_tof; T.k_hit = d_k_hit;
    if (d_rec) T.rec = *d_rec;
    T.any_rec = T.rec.pipe_x || T.rec.pipe_z || T.rec.lens2_x || T.rec.lens2_z || T.rec.target_x || T.rec.idx;
    T.s_begin = s_begin; T.s_count = s_count;
    T.inject_m = inject_m; T.inject_b = inject_b;
    T.counter = c->d_counter;
    T.stats = c->d_counter + 1;
    CUDA_TRY(cudaMemsetAsync(c->d_counter, 0, 5 * sizeof(unsigned long long), st));
    long long warps_needed = G * s_count;
    const bool guided = c->S.guided && c->S.shape != 0 && c->S.surface == 0;
    const int threads = guided ? kThreadsGuided : kThreads, wpc = threads / 32;
    const size_t smem = guided ? c->smem_guided : c->smem;
    int grid = (int)std::min<long long>((long long)c->sm_count * c->ctas_per_sm, (warps_needed + wpc - 1) / wpc);
    // (flat interface, extension: a plane has no strict curvature, so `guided` is false there: literal scan only)
    pick_table_kernel(c->S.hits_in_global != 0, guided, c->S.surface != 0, c->S.vector_math != 0)
        <<<grid, threads, smem, st>>>(c->S, T);
    CUDA_TRY(cudaGetLastError());
    return TOFRT_OK;
}

int tofrt_table_dev(tofrt_ctx *c, const double *d_radius, const double *d_offset, long long G, double *d_tof,
                    int *d_k_hit, const tofrt_ray_records *d_rec, void *stream)
{
    return table_dev_impl(c, d_radius, d_offset, G, d_tof, d_k_hit, d_rec, stream, 0, c ? c->n_shooters : 1);
}

int tofrt_table_f32_dev(tofrt_ctx *c, const double *d_radius, const double *d_offset, long long G, float *d_tof,
                        int *d_k_hit, void *stream)
{
    if (!c || G < 0) return fail(TOFRT_E_BADARG, "tofrt_table_f32_dev: bad argument");
    if (!c->ready) return fail(TOFRT_E_NOSETUP, "tofrt_table_f32_dev: call tofrt_setup first");
    if (G == 0) return TOFRT_OK;
    if (!d_radius || !d_offset || !d_tof || !d_k_hit) return fail(TOFRT_E_BADARG, "tofrt_table_f32_dev: NULL buffer");
    if (c->smem_f32 == 0) return fail(TOFRT_E_TOOLARGE, "tofrt_table_f32_dev: too many receivers for the FP32 mode");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    TableArgsF T{};
    T.radius = d_radius; T.offset = d_offset; T.G = G; T.tof = d_tof; T.k_hit = d_k_hit;
    T.s_begin = 0; T.s_count = c->n_shooters;
    T.counter = c->d_counter;
    CUDA_TRY(cudaMemsetAsync(c->d_counter, 0, 5 * sizeof(unsigned long long), st));
    int grid = (int)std::min<long long>((long long)c->sm_count, (G * c->n_shooters + kWarps - 1) / kWarps);
    tofrt_table_kernel_f32<<<grid, kThreads, c->smem_f32, st>>>(c->SF, T);
    CUDA_TRY(cudaGetLastError());
    return TOFRT_OK;
}

// Host-buffer path: the geometries are cut into chunks whose outputs are ~64 MB; chunk i is
// traced on one stream into one of two device slots while the results of chunk i-1 travel to the
// host on a second stream (pinned host buffers make that copy truly asynchronous), so the
// end-to-end time is max(compute, PCIe) instead of their sum.  The workspace is owned by the
// context and only ever grows.
int tofrt_table_host(tofrt_ctx *c, const double *radius, const double *offset, long long G, double *tof,
                     int *k_hit, const tofrt_ray_records *rec)
{
    if (!c || G < 0) return fail(TOFRT_E_BADARG, "tofrt_table_host: bad argument");
    if (!c->ready) return fail(TOFRT_E_NOSETUP, "tofrt_table_host: call tofrt_setup first");
    if (G == 0) return TOFRT_OK;
    if (!radius || !offset || !tof || !k_hit) return fail(TOFRT_E_BADARG, "tofrt_table_host: NULL buffer");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t K = c->K, E = c->E, g = (size_t)G;
    tofrt_ray_records none{};
    const tofrt_ray_records &R = rec ? *rec : none;
    // bytes of output per geometry, each array 256-byte aligned inside a slot
    size_t per_geom = E * (sizeof(double) + sizeof(int));
    const int n_rec_d = (R.pipe_x != nullptr) + (R.pipe_z != nullptr) + (R.lens2_x != nullptr) +
                        (R.lens2_z != nullptr) + (R.target_x != nullptr) + 4 * (R.legs != nullptr);
    per_geom += (size_t)n_rec_d * K * sizeof(double) + (R.idx ? K * sizeof(int) : 0);
    size_t chunk = std::max<size_t>(1, ((size_t)TOFRT_CHUNK_MB << 20) / per_geom);
    chunk = std::min(chunk, g);
    if (g > chunk) {  // even chunks, each a multiple of 1024 geometries
        size_t n = (g + chunk - 1) / chunk;
        chunk = (((g + n - 1) / n) + 1023) / 1024 * 1024;
    }
    const size_t slot_bytes = chunk * per_geom + 16 * 256;

    // grow-only workspace
    if (!c->s_compute) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
        }
    }
    if (c->ws_bytes < slot_bytes) {
        for (int i = 0; i < 2; ++i) { cudaFree(c->ws[i]); c->ws[i] = nullptr; }
        c->ws_bytes = 0;
        for (int i = 0; i < 2; ++i) CUDA_TRY(cudaMalloc(&c->ws[i], slot_bytes));
        c->ws_bytes = slot_bytes;
    }
    if (c->in_cap < 2 * g) {
        cudaFree(c->d_in); c->d_in = nullptr; c->in_cap = 0;
        CUDA_TRY(cudaMalloc(&c->d_in, sizeof(double) * 2 * g));
        c->in_cap = 2 * g;
    }
    CUDA_TRY(cudaMemcpyAsync(c->d_in, radius, sizeof(double) * g, cudaMemcpyHostToDevice, c->s_compute));
    CUDA_TRY(cudaMemcpyAsync(c->d_in + g, offset, sizeof(double) * g, cudaMemcpyHostToDevice, c->s_compute));

    int rc = TOFRT_OK;
    const size_t n_chunks = (g + chunk - 1) / chunk;
    // one (shooter, chunk) launch at a time; host layout is [S][G][...] so each lands contiguously
    for (size_t i = 0; i < n_chunks * (size_t)c->n_shooters && rc == TOFRT_OK; ++i) {
        const int slot = (int)(i & 1);
        const int sh = (int)(i / n_chunks);
        const size_t done = (i % n_chunks) * chunk;
        const size_t n = std::min(chunk, g - done);
        const size_t hbase = (size_t)sh * g + done;  // first host row of this launch
        // carve the slot
        char *base = (char *)c->ws[slot];
        auto take = [&](size_t bytes) { char *q = base; base += (bytes + 255) / 256 * 256; return q; };
        double *d_tof = (double *)take(n * E * sizeof(double));
        int *d_kh = (int *)take(n * E * sizeof(int));
        tofrt_ray_records d{};
        if (R.pipe_x) d.pipe_x = (double *)take(n * K * sizeof(double));
        if (R.pipe_z) d.pipe_z = (double *)take(n * K * sizeof(double));
        if (R.lens2_x) d.lens2_x = (double *)take(n * K * sizeof(double));
        if (R.lens2_z) d.lens2_z = (double *)take(n * K * sizeof(double));
        if (R.target_x) d.target_x = (double *)take(n * K * sizeof(double));
        if (R.idx) d.idx = (int *)take(n * K * sizeof(int));
        if (R.legs) d.legs = (double *)take(n * K * 4 * sizeof(double));
        if (i >= 2) CUDA_TRY(cudaStreamWaitEvent(c->s_compute, c->ev_free[slot], 0));
        rc = table_dev_impl(c, c->d_in + done, c->d_in + g + done, (long long)n, d_tof, d_kh, rec ? &d : nullptr,
                            (void *)c->s_compute, sh, 1);
        if (rc != TOFRT_OK) break;
        CUDA_TRY(cudaEventRecord(c->ev_done[slot], c->s_compute));
        CUDA_TRY(cudaStreamWaitEvent(c->s_copy, c->ev_done[slot], 0));
        auto back = [&](void *dst, const void *src, size_t bytes) {
            return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->s_copy);
        };
        CUDA_TRY(back(tof + hbase * E, d_tof, n * E * sizeof(double)));
        CUDA_TRY(back(k_hit + hbase * E, d_kh, n * E * sizeof(int)));
        if (R.pipe_x) CUDA_TRY(back(R.pipe_x + hbase * K, d.pipe_x, n * K * sizeof(double)));
        if (R.pipe_z) CUDA_TRY(back(R.pipe_z + hbase * K, d.pipe_z, n * K * sizeof(double)));
        if (R.lens2_x) CUDA_TRY(back(R.lens2_x + hbase * K, d.lens2_x, n * K * sizeof(double)));
        if (R.lens2_z) CUDA_TRY(back(R.lens2_z + hbase * K, d.lens2_z, n * K * sizeof(double)));
        if (R.target_x) CUDA_TRY(back(R.target_x + hbase * K, d.target_x, n * K * sizeof(double)));
        if (R.idx) CUDA_TRY(back(R.idx + hbase * K, d.idx, n * K * sizeof(int)));
        if (R.legs) CUDA_TRY(back(R.legs + hbase * K * 4, d.legs, n * K * 4 * sizeof(double)));
        CUDA_TRY(cudaEventRecord(c->ev_free[slot], c->s_copy));
    }
    cudaError_t e1 = cudaStreamSynchronize(c->s_compute), e2 = cudaStreamSynchronize(c->s_copy);
    if (rc != TOFRT_OK) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess)
        return fail(TOFRT_E_CUDA, std::string("tofrt_table_host: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return TOFRT_OK;
}

// Test hook: the return-path search of the current setup applied to arbitrary lines (m[i], b[i]).
// It runs the production table kernel with the reflected lines replaced, so it exercises exactly
// the code the engine ships (both search modes); idx[i] uses the encoding of tofrt_ray_records.
int tofrt_debug_search_host(tofrt_ctx *c, const double *m, const double *b, long long n, int *idx, double *xi,
                            double *zi)
{
    if (!c || n < 0) return fail(TOFRT_E_BADARG, "tofrt_debug_search_host: bad argument");
    if (!c->ready) return fail(TOFRT_E_NOSETUP, "tofrt_debug_search_host: call tofrt_setup first");
    if (n == 0) return TOFRT_OK;
    if (!m || !b || !idx) return fail(TOFRT_E_BADARG, "tofrt_debug_search_host: NULL buffer");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t K = c->K, E = c->E, G = ((size_t)n + K - 1) / K, tot = G * K;
    std::vector<double> hm(tot, 0.0), hb(tot, 0.0), geo(2 * G);
    std::copy(m, m + n, hm.begin());
    std::copy(b, b + n, hb.begin());
    for (size_t g = 0; g < G; ++g) { geo[g] = 0.05; geo[G + g] = 0.0; }
    double *d_m = nullptr, *d_geo = nullptr, *d_tof = nullptr;
    int *d_kh = nullptr, *d_idx = nullptr;
    double *d_xz = nullptr;
    cudaError_t e = cudaMalloc(&d_m, sizeof(double) * 2 * tot);
    if (e == cudaSuccess) e = cudaMalloc(&d_xz, sizeof(double) * 2 * tot);
    if (e == cudaSuccess) e = cudaMalloc(&d_geo, sizeof(double) * 2 * G);
    if (e == cudaSuccess) e = cudaMalloc(&d_tof, sizeof(double) * G * E);
    if (e == cudaSuccess) e = cudaMalloc(&d_kh, sizeof(int) * G * E);
    if (e == cudaSuccess) e = cudaMalloc(&d_idx, sizeof(int) * tot);
    if (e == cudaSuccess) e = cudaMemcpy(d_m, hm.data(), sizeof(double) * tot, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_m + tot, hb.data(), sizeof(double) * tot, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_geo, geo.data(), sizeof(double) * 2 * G, cudaMemcpyHostToDevice);
    int rc = TOFRT_OK;
    if (e == cudaSuccess) {
        tofrt_ray_records rec{};
        rec.idx = d_idx;
        rec.lens2_x = d_xz; rec.lens2_z = d_xz + tot;
        rc = table_dev_impl(c, d_geo, d_geo + G, (long long)G, d_tof, d_kh, &rec, nullptr, 0, 1, d_m, d_m + tot);
    }
    std::vector<int> hidx(tot);
    if (e == cudaSuccess && rc == TOFRT_OK) e = cudaMemcpy(hidx.data(), d_idx, sizeof(int) * tot, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == TOFRT_OK && xi) e = cudaMemcpy(xi, d_xz, sizeof(double) * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == TOFRT_OK && zi) e = cudaMemcpy(zi, d_xz + tot, sizeof(double) * n, cudaMemcpyDeviceToHost);
    cudaFree(d_xz);
    cudaFree(d_m); cudaFree(d_geo); cudaFree(d_tof); cudaFree(d_kh); cudaFree(d_idx);
    if (rc != TOFRT_OK) return rc;
    if (e != cudaSuccess) return fail(TOFRT_E_CUDA, std::string("tofrt_debug_search_host: ") + cudaGetErrorString(e));
    std::copy(hidx.begin(), hidx.begin() + n, idx);
    return TOFRT_OK;
}

// Page-locked host memory for the callers of the *_host entry points: with pinned buffers the
// device->host copies of tofrt_table_host overlap the kernels; with pageable memory they are staged.
void *tofrt_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        fail(TOFRT_E_CUDA, "tofrt_host_alloc: cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}

void tofrt_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int tofrt_hits_host(tofrt_ctx *c, const double *radius, const double *offset, long long G, long long capacity,
                    int *hit_g, int *hit_e, int *hit_k, double *hit_tof, long long *n_hits)
{
    if (!c || G < 0 || capacity < 0 || !n_hits) return fail(TOFRT_E_BADARG, "tofrt_hits_host: bad argument");
    if (!c->ready) return fail(TOFRT_E_NOSETUP, "tofrt_hits_host: call tofrt_setup first");
    *n_hits = 0;
    if (G == 0) return TOFRT_OK;
    if (!radius || !offset || (capacity > 0 && (!hit_g || !hit_e || !hit_k || !hit_tof)))
        return fail(TOFRT_E_BADARG, "tofrt_hits_host: NULL buffer");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t E = c->E, g = (size_t)G, S = (size_t)c->n_shooters;
    // geometries per pass: the dense intermediate stays below ~6 GB
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(g, ((size_t)6 << 30) / (E * 12 * S)));
    const size_t rows_max = chunk * S;
    int slot = 0;
    auto cleanup = [&]() {};  // the workspace is cached in the context
    auto dalloc = [&](void **q, size_t bytes) {
        bytes = std::max<size_t>(bytes, 16);
        const int i = slot++;
        cudaError_t e = cudaSuccess;
        if (c->hw_bytes[i] < bytes) {
            cudaFree(c->hw[i]); c->hw[i] = nullptr; c->hw_bytes[i] = 0;
            e = cudaMalloc(&c->hw[i], bytes);
            if (e == cudaSuccess) c->hw_bytes[i] = bytes;
        }
        *q = c->hw[i];
        return e;
    };
#define TRY_H(expr)                                                                                    \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess) { cleanup(); return fail(TOFRT_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } \
    } while (0)
    double *d_in = nullptr, *d_tof = nullptr, *d_ht = nullptr;
    int *d_kh = nullptr, *d_hg = nullptr, *d_he = nullptr, *d_hk = nullptr;
    long long *d_cnt = nullptr, *d_off = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    if (rows_max + 1 > 0x7fffffff) return fail(TOFRT_E_TOOLARGE, "tofrt_hits_host: too many rows per pass");
    TRY_H(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt, d_off, (int)(rows_max + 1)));
    TRY_H(dalloc((void **)&d_in, sizeof(double) * 2 * g));
    TRY_H(dalloc((void **)&d_tof, sizeof(double) * rows_max * E));
    TRY_H(dalloc((void **)&d_kh, sizeof(int) * rows_max * E));
    TRY_H(dalloc((void **)&d_cnt, sizeof(long long) * (rows_max + 1)));
    TRY_H(dalloc((void **)&d_off, sizeof(long long) * (rows_max + 1)));
    TRY_H(dalloc(&d_tmp, tmp_bytes));
    TRY_H(dalloc((void **)&d_hg, sizeof(int) * (size_t)capacity));
    TRY_H(dalloc((void **)&d_he, sizeof(int) * (size_t)capacity));
    TRY_H(dalloc((void **)&d_hk, sizeof(int) * (size_t)capacity));
    TRY_H(dalloc((void **)&d_ht, sizeof(double) * (size_t)capacity));
    TRY_H(cudaMemcpy(d_in, radius, sizeof(double) * g, cudaMemcpyHostToDevice));
    TRY_H(cudaMemcpy(d_in + g, offset, sizeof(double) * g, cudaMemcpyHostToDevice));
    long long total = 0;
    if (S > 1 && chunk < g) { cleanup(); return fail(TOFRT_E_TOOLARGE, "tofrt_hits_host: too many geometries x shooters for one pass"); }
    for (size_t done = 0; done < g; done += chunk) {
        const size_t n = std::min(chunk, g - done), rows = n * S;
        int rc = tofrt_table_dev(c, d_in + done, d_in + g + done, (long long)n, d_tof, d_kh, nullptr, nullptr);
        if (rc != TOFRT_OK) { cleanup(); return rc; }
        const int blocks = (int)((rows * 32 + 255) / 256);
        TRY_H(cudaMemsetAsync(d_cnt + rows, 0, sizeof(long long)));
        tofrt_count_hits_kernel<<<blocks, 256>>>(d_kh, (long long)rows, (int)E, d_cnt);
        TRY_H(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cnt, d_off, (int)rows + 1));
        tofrt_compact_hits_kernel<<<blocks, 256>>>(d_kh, d_tof, (long long)rows, (int)E, d_off, (long long)(done * 1),
                                                   total, capacity, d_hg, d_he, d_hk, d_ht);
        TRY_H(cudaGetLastError());
        long long got = 0;
        TRY_H(cudaMemcpy(&got, d_off + rows, sizeof(long long), cudaMemcpyDeviceToHost));
        total += got;
    }
    *n_hits = total;
    const size_t ncopy = (size_t)std::min<long long>(total, capacity);
    if (ncopy) {
        TRY_H(cudaMemcpy(hit_g, d_hg, sizeof(int) * ncopy, cudaMemcpyDeviceToHost));
        TRY_H(cudaMemcpy(hit_e, d_he, sizeof(int) * ncopy, cudaMemcpyDeviceToHost));
        TRY_H(cudaMemcpy(hit_k, d_hk, sizeof(int) * ncopy, cudaMemcpyDeviceToHost));
        TRY_H(cudaMemcpy(hit_tof, d_ht, sizeof(double) * ncopy, cudaMemcpyDeviceToHost));
    }
    cleanup();
#undef TRY_H
    if (total > capacity) return fail(TOFRT_E_TOOLARGE, "tofrt_hits_host: capacity too small (n_hits holds the required size)");
    return TOFRT_OK;
}

int tofrt_fermat_dev(tofrt_ctx *c, const double *d_fx, const double *d_fz, long long F, double *d_tof, int *d_kmin,
                     double *d_alpha, void *stream)
{
    if (!c || F < 0) return fail(TOFRT_E_BADARG, "tofrt_fermat_dev: bad argument");
    if (!c->ready) return fail(TOFRT_E_NOSETUP, "tofrt_fermat_dev: call tofrt_setup first");
    if (F == 0) return TOFRT_OK;
    if (!d_fx || !d_fz || !d_tof || !d_kmin) return fail(TOFRT_E_BADARG, "tofrt_fermat_dev: NULL buffer");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int K = c->K, E = c->E;
    const size_t smem = (size_t)K * (sizeof(double2) + sizeof(double) + sizeof(float2) + sizeof(float)) + 16;
    if (smem > 227 * 1024) return fail(TOFRT_E_TOOLARGE, "tofrt_fermat_dev: too many lens samples");
    if (!c->d_d1) {
        CUDA_TRY(cudaMalloc(&c->d_d1, sizeof(double) * (size_t)K * E));
        dim3 grid((K + 127) / 128, E);
        tofrt_fermat_d1_kernel<<<grid, 128, 0, st>>>(c->S, c->params_d, c->d_d1);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaFuncSetAttribute(tofrt_fermat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    FermatArgs A{};
    A.fx = d_fx; A.fz = d_fz; A.F = F; A.d1 = c->d_d1; A.alpha = c->d_alpha; A.tof = d_tof; A.kmin = d_kmin;
    A.alpha_out = d_alpha; A.z_el = c->params_d; A.counter = c->d_counter;
    CUDA_TRY(cudaMemsetAsync(c->d_counter, 0, 5 * sizeof(unsigned long long), st));
    int occ = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tofrt_fermat_kernel, kFermatThreads, smem));
    const long long items = ((F + kFermatThreads - 1) / kFermatThreads) * E;
    int grid = (int)std::min<long long>((long long)c->sm_count * std::max(1, occ), items);
    tofrt_fermat_kernel<<<grid, kFermatThreads, smem, st>>>(c->S, A);
    CUDA_TRY(cudaGetLastError());
    return TOFRT_OK;
}

int tofrt_fermat_host(tofrt_ctx *c, const double *fx, const double *fz, long long F, double *tof, int *kmin,
                      double *alpha)
{
    if (!c || F < 0) return fail(TOFRT_E_BADARG, "tofrt_fermat_host: bad argument");
    if (!c->ready) return fail(TOFRT_E_NOSETUP, "tofrt_fermat_host: call tofrt_setup first");
    if (F == 0) return TOFRT_OK;
    if (!fx || !fz || !tof || !kmin) return fail(TOFRT_E_BADARG, "tofrt_fermat_host: NULL buffer");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t n = (size_t)F, EF = n * c->E;
    double *d_f = nullptr, *d_t = nullptr, *d_a = nullptr;
    int *d_k = nullptr;
    cudaError_t e = cudaMalloc(&d_f, sizeof(double) * 2 * n);
    if (e == cudaSuccess) e = cudaMalloc(&d_t, sizeof(double) * EF);
    if (e == cudaSuccess) e = cudaMalloc(&d_k, sizeof(int) * EF);
    if (e == cudaSuccess && alpha) e = cudaMalloc(&d_a, sizeof(double) * EF);
    if (e == cudaSuccess) e = cudaMemcpy(d_f, fx, sizeof(double) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_f + n, fz, sizeof(double) * n, cudaMemcpyHostToDevice);
    int rc = TOFRT_OK;
    if (e == cudaSuccess) rc = tofrt_fermat_dev(c, d_f, d_f + n, F, d_t, d_k, d_a, nullptr);
    if (e == cudaSuccess && rc == TOFRT_OK) e = cudaMemcpy(tof, d_t, sizeof(double) * EF, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == TOFRT_OK) e = cudaMemcpy(kmin, d_k, sizeof(int) * EF, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == TOFRT_OK && alpha) e = cudaMemcpy(alpha, d_a, sizeof(double) * EF, cudaMemcpyDeviceToHost);
    cudaFree(d_f); cudaFree(d_t); cudaFree(d_k); cudaFree(d_a);
    if (rc != TOFRT_OK) return rc;
    if (e != cudaSuccess) return fail(TOFRT_E_CUDA, std::string("tofrt_fermat_host: ") + cudaGetErrorString(e));
    return TOFRT_OK;
}

// ---- 3-D two-point solve ------------------------------------------------------------------
int tofrt_fermat3d_setup(tofrt_ctx *c, const double *ang, int n_side, const double *ex, const double *ey, int E3)
{
    if (!c || !ang || !ex || !ey) return fail(TOFRT_E_BADARG, "tofrt_fermat3d_setup: NULL argument");
    if (!c->ready) return fail(TOFRT_E_NOSETUP, "tofrt_fermat3d_setup: call tofrt_setup first (it defines the lens)");
    if (n_side < 2) return fail(TOFRT_E_TOOFEW, "Surface grid needs at least two samples per axis.");
    if (E3 < 1) return fail(TOFRT_E_BADARG, "tofrt_fermat3d_setup: need at least one element");
    const size_t ns = (size_t)n_side * n_side;
    if (ns * sizeof(float4) > 200 * 1024) return fail(TOFRT_E_TOOLARGE, "tofrt_fermat3d_setup: n_side too large (<= 113)");
    for (int i = 0; i + 1 < n_side; ++i)
        if (!(ang[i + 1] > ang[i])) return fail(TOFRT_E_BADARG, "tofrt_fermat3d_setup: ang must increase strictly");
    CUDA_TRY(cudaSetDevice(c->device));
    c->f3_ready = false;
    cudaFree(c->f3_geo); cudaFree(c->f3_smp);
    c->f3_geo = nullptr; c->f3_smp = nullptr;
    CUDA_TRY(cudaMalloc(&c->f3_geo, sizeof(double) * ((size_t)n_side + 2 * (size_t)E3)));
    CUDA_TRY(cudaMalloc(&c->f3_smp, sizeof(float4) * ns * (size_t)E3));
    CUDA_TRY(cudaMemcpy(c->f3_geo, ang, sizeof(double) * n_side, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->f3_geo + n_side, ex, sizeof(double) * E3, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->f3_geo + n_side + E3, ey, sizeof(double) * E3, cudaMemcpyHostToDevice));
    tofrt_surf3d S{};
    S.inv_c1 = 1.0 / c->S.c1; S.inv_c2 = 1.0 / c->S.c2;
    S.twod = c->S.twod; S.K2 = c->S.K2; S.fourAC = c->S.fourAC; S.twoA = c->S.twoA;
    S.h0 = c->S.h0; S.surface = c->S.surface;
    S.lo = ang[0]; S.hi = ang[n_side - 1];
    c->f3_S = S; c->f3_E = E3; c->f3_n = n_side;
    c->f3_cell = (S.hi - S.lo) / (n_side - 1);
    F3dArgs A{};
    A.ang = c->f3_geo; A.ex = c->f3_geo + n_side; A.ey = c->f3_geo + n_side + E3;
    A.E = E3; A.n_side = n_side; A.z_el = c->params_d;
    dim3 grid((unsigned)((ns + 127) / 128), (unsigned)E3);
    tofrt_f3d_prepare_kernel<<<grid, 128>>>(S, A, c->f3_smp);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFuncSetAttribute(tofrt_fermat3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(ns * sizeof(float4))));
    CUDA_TRY(cudaDeviceSynchronize());
    c->f3_ready = true;
    return TOFRT_OK;
}

int tofrt_fermat3d_dev(tofrt_ctx *c, const double *d_fx, const double *d_fy, const double *d_fz, long long F,
                       double *d_tof, int *d_kmin, double *d_ab, void *stream)
{
    if (!c || F < 0) return fail(TOFRT_E_BADARG, "tofrt_fermat3d_dev: bad argument");
    if (!c->ready || !c->f3_ready) return fail(TOFRT_E_NOSETUP, "tofrt_fermat3d_dev: call tofrt_fermat3d_setup first");
    if (F == 0) return TOFRT_OK;
    if (!d_fx || !d_fy || !d_fz || !d_tof) return fail(TOFRT_E_BADARG, "tofrt_fermat3d_dev: NULL buffer");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = c->f3_n, E3 = c->f3_E;
    const size_t smem = (size_t)n * n * sizeof(float4);
    F3dArgs A{};
    A.fx = d_fx; A.fy = d_fy; A.fz = d_fz; A.F = F; A.smp = c->f3_smp;
    A.ang = c->f3_geo; A.ex = c->f3_geo + n; A.ey = c->f3_geo + n + E3;
    A.E = E3; A.n_side = n; A.z_el = c->params_d; A.cell = c->f3_cell;
    A.tof = d_tof; A.kmin = d_kmin; A.ab = d_ab; A.counter = c->d_counter;
    CUDA_TRY(cudaMemsetAsync(c->d_counter, 0, 5 * sizeof(unsigned long long), st));
    int occ = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tofrt_fermat3d_kernel, kF3dThreads, smem));
    const long long items = ((F + kF3dThreads - 1) / kF3dThreads) * E3;
    int grid = (int)std::min<long long>((long long)c->sm_count * std::max(1, occ), items);
    tofrt_fermat3d_kernel<<<grid, kF3dThreads, smem, st>>>(c->f3_S, A);
    CUDA_TRY(cudaGetLastError());
    return TOFRT_OK;
}

// Host buffers: voxels are cut into chunks; chunk i is solved on one stream while the results of
// chunk i-1 travel to the host on a second one (element-major outputs: strided 2-D copies).
int tofrt_fermat3d_host(tofrt_ctx *c, const double *fx, const double *fy, const double *fz, long long F, double *tof,
                        int *kmin, double *ab)
{
    if (!c || F < 0) return fail(TOFRT_E_BADARG, "tofrt_fermat3d_host: bad argument");
    if (!c->ready || !c->f3_ready) return fail(TOFRT_E_NOSETUP, "tofrt_fermat3d_host: call tofrt_fermat3d_setup first");
    if (F == 0) return TOFRT_OK;
    if (!fx || !fy || !fz || !tof) return fail(TOFRT_E_BADARG, "tofrt_fermat3d_host: NULL buffer");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t E3 = (size_t)c->f3_E, n = (size_t)F;
    const size_t per_voxel = E3 * (sizeof(double) + (kmin ? sizeof(int) : 0) + (ab ? 2 * sizeof(double) : 0));
    size_t chunk = std::max<size_t>(256, (((size_t)TOFRT_CHUNK_MB << 20) / per_voxel) / 256 * 256);
    chunk = std::min(chunk, n);
    if (!c->s_compute) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
        }
    }
    double *d_f = nullptr;
    char *slot[2] = {nullptr, nullptr};
    const size_t slot_bytes = chunk * per_voxel + 3 * 256;
    cudaError_t e = cudaMalloc(&d_f, sizeof(double) * 3 * n);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMalloc(&slot[i], slot_bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_f, fx, sizeof(double) * n, cudaMemcpyHostToDevice, c->s_compute);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_f + n, fy, sizeof(double) * n, cudaMemcpyHostToDevice, c->s_compute);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_f + 2 * n, fz, sizeof(double) * n, cudaMemcpyHostToDevice, c->s_compute);
    int rc = TOFRT_OK;
    size_t i = 0;
    for (size_t done = 0; done < n && e == cudaSuccess && rc == TOFRT_OK; done += chunk, ++i) {
        const size_t m = std::min(chunk, n - done);
        const int sl = (int)(i & 1);
        char *base = slot[sl];
        auto take = [&](size_t bytes) { char *q = base; base += (bytes + 255) / 256 * 256; return q; };
        double *d_t = (double *)take(m * E3 * sizeof(double));
        int *d_k = kmin ? (int *)take(m * E3 * sizeof(int)) : nullptr;
        double *d_ab = ab ? (double *)take(m * E3 * 2 * sizeof(double)) : nullptr;
        if (i >= 2) e = cudaStreamWaitEvent(c->s_compute, c->ev_free[sl], 0);
        if (e != cudaSuccess) break;
        rc = tofrt_fermat3d_dev(c, d_f + done, d_f + n + done, d_f + 2 * n + done, (long long)m, d_t, d_k, d_ab,
                                (void *)c->s_compute);
        if (rc != TOFRT_OK) break;
        e = cudaEventRecord(c->ev_done[sl], c->s_compute);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->s_copy, c->ev_done[sl], 0);
        // device chunk is [E][m]; host table is [E][F]
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync(tof + done, n * sizeof(double), d_t, m * sizeof(double), m * sizeof(double), E3,
                                  cudaMemcpyDeviceToHost, c->s_copy);
        if (e == cudaSuccess && kmin)
            e = cudaMemcpy2DAsync(kmin + done, n * sizeof(int), d_k, m * sizeof(int), m * sizeof(int), E3,
                                  cudaMemcpyDeviceToHost, c->s_copy);
        if (e == cudaSuccess && ab)
            e = cudaMemcpy2DAsync(ab + 2 * done, n * 2 * sizeof(double), d_ab, m * 2 * sizeof(double),
                                  m * 2 * sizeof(double), E3, cudaMemcpyDeviceToHost, c->s_copy);
        if (e == cudaSuccess) e = cudaEventRecord(c->ev_free[sl], c->s_copy);
    }
    cudaError_t e1 = cudaStreamSynchronize(c->s_compute), e2 = cudaStreamSynchronize(c->s_copy);
    cudaFree(d_f); cudaFree(slot[0]); cudaFree(slot[1]);
    if (rc != TOFRT_OK) return rc;
    if (e == cudaSuccess) e = e1 != cudaSuccess ? e1 : e2;
    if (e != cudaSuccess) return fail(TOFRT_E_CUDA, std::string("tofrt_fermat3d_host: ") + cudaGetErrorString(e));
    return TOFRT_OK;
}

int tofrt_measure_fp64_peak(tofrt_ctx *c, double *dfma_per_s, double *dmul_dadd_per_s)
{
    if (!c || !dfma_per_s || !dmul_dadd_per_s) return fail(TOFRT_E_BADARG, "tofrt_measure_fp64_peak: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    double *d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 1 << 14, blocks = c->sm_count * 8, threads = 256;
    const double inst = (double)blocks * threads * iters * 16.0;
    double best[2] = {0, 0};
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) tofrt_fp64_peak_kernel<0><<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);
            else tofrt_fp64_peak_kernel<1><<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);
            cudaEventRecord(e1);
            cudaError_t e = cudaEventSynchronize(e1);
            if (e != cudaSuccess) { cudaFree(d_out); return fail(TOFRT_E_CUDA, cudaGetErrorString(e)); }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0) best[mode] = std::max(best[mode], inst / (ms * 1e-3));
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    *dfma_per_s = best[0];
    *dmul_dadd_per_s = best[1];
    return TOFRT_OK;
}

int tofrt_kernel_info(tofrt_ctx *c, int *grid, int *block, int *smem_bytes, int *regs)
{
    if (!c) return fail(TOFRT_E_BADARG, "tofrt_kernel_info: ctx is NULL");
    const bool guided = c->S.guided && c->S.shape != 0 && c->S.surface == 0;
    cudaFuncAttributes fa{};
    CUDA_TRY(cudaFuncGetAttributes(&fa, pick_table_kernel(false, guided, false, c->S.vector_math != 0)));
    if (grid) *grid = c->sm_count * c->ctas_per_sm;
    if (block) *block = guided ? kThreadsGuided : kThreads;
    if (smem_bytes) *smem_bytes = (int)(guided ? c->smem_guided : c->smem);
    if (regs) *regs = fa.numRegs;
    return TOFRT_OK;
}

}  // extern "C"
