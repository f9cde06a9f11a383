// tofrt_common.cuh -- launch constants, the geometry-independent setup (DevSetup), per-call arguments,
// small device helpers (NumPy/Python semantics, lens model, Snell refraction) and the TMA bulk-copy wrappers.
// Part of the single translation unit tofrt_kernels.cu (included inside its anonymous namespace).
#pragma once


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
constexpr int kSplitQueue = 256;  // small-G mode: rays per geometry the CTA-wide on-sample pass can queue
constexpr size_t kSplitQueueBytes = 16 + (size_t)kSplitQueue * (sizeof(double2) + sizeof(int));
// per-ray tables of the FP64 table kernel: ray_t[shooter][c][plane][lane] holds ray min(32 c + lane, K-1)
enum { kPlaneXp, kPlaneZp, kPlaneCpq, kPlaneSpq, kPlaneZf, kPlaneApq, kPlaneBpq, kPlanePhi, kRayPlanes };
// executed-work counters of the instrumented (STATS) instantiation, per thread-ray (tofrt_last_stats)
enum { kStRays, kStBracket, kStProbes, kStDiffs, kStLens2, kStLanding, kStCand, kStWinners, kStFallback, kStScan,
       kStOnSample, kNumStats };
#ifndef TOFRT_CHUNK_MB
#define TOFRT_CHUNK_MB 64  // output bytes per pipelined chunk of tofrt_table_host
#endif

// Everything the kernels need that does not depend on the geometry.
struct DevSetup {
    // derived constants, computed on the host in the reference's operation order
    double c1, c2, d;
    double A, fourAC, twoA, twod, K2, phi1;  // lens surface (main_rt.py:181-212)
    double c2_over_c1, c1_over_c2;           // v2 / v1 of the two refractions
    tofrt_vec_lens VL;                       // the same constants as the vector form uses them (tofrt_raymath.h)
    int n_shooters;                          // S: shots per geometry (full-matrix capture when S = E)
    const double *x_sh, *z_sh;               // [S] shooter positions
    double rtol, atol, tol_max;              // np.isclose tolerances, and an upper bound of atol+rtol|x_e|
    double rx_min, rx_max;                   // finite receiver range
    int K, Npad, E, E_sorted, has_nonfinite_rx, tables_ok;
    int C;                 // 32-ray chunks per shot: ceil(K / 32)
    const double *ray_t;   // [S][C][kRayPlanes][32] per-ray tables, padded with copies of the last ray
    double2 P0, PN;        // first and last lens sample (the guided search tests them for every ray)
    int chords_plain;    // 1: abscissae strictly increasing and no chord is np.isclose-vertical (main_rt.py:110): checked per setup
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
    // candidate lookup: nb cells of width 1/binv from bx0; bucket[c] = first | count << 16 lists the sorted receivers
    // that a landing point whose computed cell index is c can match (one cell of slack either side + tol_max)
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
    // sparse emission (tofrt_hits_host): when hit_slots is set the warp writes the hits of its row, ordered by
    // receiver, into the row's slot of hit_cap (e, k, tof) entries and their number into hit_count[row]; the dense
    // row is then not stored (a row with more than hit_cap hits only counts: the host falls back to the dense path)
    uint4 *hit_slots;
    int *hit_count;
    int hit_cap;
    unsigned long long *counter;
    unsigned long long *stats;  // [kNumStats] executed-work counters (instrumented instantiation only)
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

