/*
 * tofrt.h -- C ABI of the B200-native time-of-flight ray-tracing engine.
 *
 * The reference (edudscrc/ray-tracing-ultrasound) has no FFI: its seam is the
 * Python function shoot_rays() completed by module globals, plus the two
 * script bodies.  Each entry point below names the reference code it replaces
 * (file:line in /root/reference); INTEGRATION.md shows the ctypes binding a
 * maintainer adds to main_rt.py / main_compare.py.
 *
 * Conventions
 *   - plain C types only; every function returns 0 or a negative TOFRT_E_*;
 *     tofrt_last_error() gives the message (thread-local); nothing throws.
 *   - geometric failures are data, not errors: a ray that misses the pipe or
 *     the lens, or is totally reflected, is NaN in the ray records and simply
 *     never matches a receiver (main_rt.py:392-393 stores None as NaN).
 *   - "_dev" functions take device pointers and are asynchronous on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).
 *     "_host" functions take host pointers, copy in/out and synchronise.
 *   - table layout is geometry-major: tof[g*E + e], k_hit[g*E + e]; this is
 *     the row order of database_2.csv (main_rt.py:464-504) and makes a
 *     G-slice (one GPU's shard) one contiguous block.
 */
#ifndef TOFRT_H
#define TOFRT_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOFRT_VERSION 100

#define TOFRT_OK 0
#define TOFRT_E_BADARG (-1)   /* NULL pointer, negative size, mismatched lengths   */
#define TOFRT_E_TOOFEW (-2)   /* < 2 lens samples: ValueError of main_rt.py:24-29  */
#define TOFRT_E_CUDA (-3)     /* CUDA runtime error (message has the CUDA string)  */
#define TOFRT_E_TOOLARGE (-4) /* K/E do not fit the 227 KB shared-memory staging   */
#define TOFRT_E_NOSETUP (-5)  /* tofrt_setup() has not been called on this context */

/* The module globals the reference's functions read (main_rt.py:449-455).
 * d must be passed as the caller computed it (l0 + h0 in float64). */
typedef struct tofrt_params {
    double c1; /* speed of sound in the lens   (main_rt.py:449) */
    double c2; /* speed of sound in the water  (main_rt.py:450) */
    double l0; /* array-to-lens-apex distance  (main_rt.py:453) */
    double h0; /* lens-apex-to-origin distance (main_rt.py:454) */
    double d;  /* l0 + h0                      (main_rt.py:455) */
} tofrt_params;

/* Optional per-ray records of a trace, each [G][K] (NULL = not wanted).  These
 * are the dict entries of shoot_rays (main_rt.py:432-441) that depend on the
 * geometry, the discrete return-path index of main_rt.py:99, and the four
 * travel-time legs of main_compare.py:514-517. */
typedef struct tofrt_ray_records {
    double *pipe_x, *pipe_z;   /* "pipe_x", "pipe_z"     */
    double *lens2_x, *lens2_z; /* "lens_2_x", "lens_2_z" */
    double *target_x;          /* "target_x" ("target_z" is a copy of z_f) */
    int *idx;    /* >=0 first sign-change segment; -1 miss; <=-2: sample -2-idx on the line */
    double *legs; /* [G][4][K] tof_1..tof_4 */
} tofrt_ray_records;

typedef struct tofrt_ctx tofrt_ctx;

int tofrt_version(void);
const char *tofrt_last_error(void);
int tofrt_device_count(void);

/* One context per (process, GPU).  Holds the lens tables and the staged setup. */
int tofrt_create(tofrt_ctx **ctx, int device);
int tofrt_destroy(tofrt_ctx *ctx);

/* Replaces the per-iteration setup of main_rt.py:469-479 and the geometry-
 * independent part of shoot_rays (main_rt.py:338-349): lens samples
 * x_z_from_alpha(alpha), tangents, incident leg from the shooter
 * (x_shooter, z_shooter) and the first refraction.  alpha, z_f: K host
 * doubles (the reference's N lens samples are the same K angles).  x_rx: E
 * receiver abscissae; rtol/atol: the np.isclose tolerances of
 * main_rt.py:495 (1e-5, 1e-6) or main_compare.py:519 (1e-5, 1e-4). */
int tofrt_setup(tofrt_ctx *ctx, const tofrt_params *params, const double *alpha, const double *z_f,
                int K, double x_shooter, double z_shooter, const double *x_rx, int E, double rtol,
                double atol);

/* Full-matrix capture: the same setup for S shooters at once (the reference
 * shoots only element 32, main_rt.py:481-482; any element can shoot).  After this call the
 * table functions trace every shooter for every geometry: outputs are tof[s][g][e],
 * k_hit[s][g][e] and records [s][g][K] -- S blocks in the single-shooter layout. */
int tofrt_setup_shooters(tofrt_ctx *ctx, const tofrt_params *params, const double *alpha,
                         const double *z_f, int K, const double *x_shooters, const double *z_shooters,
                         int n_shooters, const double *x_rx, int E, double rtol, double atol);

/* Options.  TOFRT_OPT_SEARCH selects how the return-path entry index of
 * find_line_curve_intersection (main_rt.py:78-99) is found:
 *   TOFRT_SEARCH_SCAN    the reference's literal first-sign-change scan over all
 *                        N lens samples (early exit per warp);
 *   TOFRT_SEARCH_GUIDED  (default) bracketed search that returns the identical
 *                        index in O(log N) evaluations whenever the sampled lens
 *                        is a strictly concave/convex polyline, and falls back to
 *                        the scan for any ray whose signs are within rounding
 *                        error of zero.  Results are bit-identical by construction. */
#define TOFRT_OPT_SEARCH 1
#define TOFRT_SEARCH_SCAN 0
#define TOFRT_SEARCH_GUIDED 1
/* TOFRT_OPT_SURFACE (extension, parity unpinned -- BASELINE configs 1-2 name a "flat interface" the
 * reference does not have): TOFRT_SURFACE_LENS = the reference's constant-TOF lens h(alpha)
 * (main_rt.py:180-214, default); TOFRT_SURFACE_FLAT = the plane z = h0, sampled at the same angles
 * (h = h0 / cos(alpha)).  Set it BEFORE tofrt_setup; changing it invalidates the current setup. */
#define TOFRT_OPT_SURFACE 2
#define TOFRT_SURFACE_LENS 0
#define TOFRT_SURFACE_FLAT 1
/* TOFRT_OPT_MATH: how the two interface laws of the return path -- reflection on the pipe
 * (main_rt.py:283-292, 367-375) and refraction back into the lens (main_rt.py:267-280, 396-401) -- are
 * evaluated.  TOFRT_MATH_ANGLE: the reference's own bookkeeping, every direction an angle
 * (atan / tan, atan2 / sin / asin / tan), in the reference's operation order.  TOFRT_MATH_VECTOR: the
 * same quantities from direction cosines by the addition theorems (csrc/tofrt_raymath.h): no angle is
 * formed, two square roots and three divisions replace seven elementary-function calls per ray, fewer
 * roundings; results agree with the angle form to ~1e-15 relative, i.e. identical hit decisions and
 * travel times within 1e-17 s (tests), and the first-sign-change index differs only where it is
 * rounding noise (the reference's 12 degenerate fixture geometries).  Both forms are bit-identical
 * to their CPU twin.  Default: TOFRT_MATH_VECTOR (24 % faster; the whole GPU parity suite runs in both
 * modes).  May be changed at any time after tofrt_create. */
#define TOFRT_OPT_MATH 3
#define TOFRT_MATH_ANGLE 0
#define TOFRT_MATH_VECTOR 1
#ifndef TOFRT_MATH_DEFAULT
#define TOFRT_MATH_DEFAULT TOFRT_MATH_VECTOR
#endif
int tofrt_set_option(tofrt_ctx *ctx, int option, int value);

/* Facts about the current setup: TOFRT_INFO_SHAPE -> +1 concave / -1 convex /
 * 0 neither (guided search disabled); TOFRT_INFO_SEARCH -> current search mode. */
#define TOFRT_INFO_SHAPE 1
#define TOFRT_INFO_SEARCH 2
#define TOFRT_INFO_MATH 3
int tofrt_get_info(tofrt_ctx *ctx, int what);

/* Executed-work statistics.  tofrt_set_option(ctx, TOFRT_OPT_STATS, 1) makes the following table calls run the
 * INSTRUMENTED instantiation of the table kernel (same results; a few percent slower -- leave it off when timing);
 * tofrt_last_stats then returns the counters of the most recent table call on this context (synchronises), counted
 * per thread-ray (a warp walks 32 * ceil(K / 32) rays per geometry: indices beyond K-1 repeat the last ray):
 * [0] rays traced, [1] rays whose return path got a certified bracket (guided search), [2] bracket probes,
 * [3] exact diffs of main_rt.py:78 evaluated (bracket check or literal scan), [4] rays that entered the lens again,
 * [5] rays that reached the array line (no total internal reflection), [6] np.isclose receiver tests, [7] rays that
 * won a receiver (four legs computed), [8] rays that needed the fallback bisection, [9] rays that needed the exact
 * O(N) scan (guided mode: undecidable rays; 0 in every run so far), [10] rays that ran the |diff| <= 1e-8 on-sample rule.
 * Entries beyond TOFRT_NUM_STATS are zero.  With statistics off the counters read zero. */
#define TOFRT_OPT_STATS 4
/* Small-G latency mode (default 1 = automatic): a call with fewer (shooter, geometry) items than half the GPU's
 * resident warps -- main_rt.py's own 210-geometry sweep, main_compare.py's single geometry -- runs one CTA per item
 * with one 32-ray chunk per warp and a CTA-wide first-hit reduction, instead of one warp per item walking its
 * ceil(K / 32) chunks serially (needs ceil(K / 32) <= 32, the guided search and the shared-memory receiver budget;
 * otherwise the call silently takes the regular kernel).  Same results, bit for bit.  0 switches it off. */
#define TOFRT_OPT_SMALL_G 5
#define TOFRT_NUM_STATS 11
int tofrt_last_stats(tofrt_ctx *ctx, unsigned long long *stats, int n);

/* Lens samples of the current setup, "lens_1_x"/"lens_1_z" (main_rt.py:338). */
int tofrt_lens_host(tofrt_ctx *ctx, double *lens_x, double *lens_z);

/* Replaces the sweep body main_rt.py:464-501 for G geometries
 * (radius[g], offset[g]) = (r_outer, pipe_offset): traces the K rays of the
 * shot (shoot_rays, main_rt.py:351-405) and assembles, per receiver e, the
 * first matching ray k_hit (-1: none) and tof_total (0.0 when not hit, as the
 * CSV row keeps 0).  rec may be NULL. */
int tofrt_table_dev(tofrt_ctx *ctx, const double *d_radius, const double *d_offset, long long G,
                    double *d_tof, int *d_k_hit, const tofrt_ray_records *d_rec, void *stream);
int tofrt_table_host(tofrt_ctx *ctx, const double *radius, const double *offset, long long G,
                     double *tof, int *k_hit, const tofrt_ray_records *rec);

/* Page-locked host memory for the buffers handed to the *_host entry points (optional: any host
 * pointer works, pinned ones let the result copies overlap the kernels).  NULL on failure. */
void *tofrt_host_alloc(size_t bytes);
void tofrt_host_free(void *p);

/* Sparse form of tofrt_table_host: only the hits (k_hit >= 0), ordered by (row, receiver), where
 * row = geometry index (shooter * G + geometry with several shooters).  The dense table is
 * 97-98 % misses; the list cuts the device->host traffic ~30x.  If more than `capacity` hits
 * exist, the first `capacity` are returned, *n_hits holds the total and the call returns
 * TOFRT_E_TOOLARGE.  Same rows as main_rt.py:487-504 writes with hitted == True. */
int tofrt_hits_host(tofrt_ctx *ctx, const double *radius, const double *offset, long long G,
                    long long capacity, int *hit_g, int *hit_e, int *hit_k, double *hit_tof,
                    long long *n_hits);

/* FP32 mode of tofrt_table_dev (reported separately, no bit-parity claim): the same
 * pipeline with the per-ray arithmetic in float; tof is float[G][E].  radius/offset stay
 * double arrays so both modes share their inputs. */
int tofrt_table_f32_dev(tofrt_ctx *ctx, const double *d_radius, const double *d_offset, long long G,
                        float *d_tof, int *d_k_hit, void *stream);

/* Two-point Fermat solve (no function of the reference corresponds to it --
 * parity unpinned, checked against a NumPy/SciPy oracle and analytic known answers).  For every
 * (receiver e at (x_rx[e], d), focal point (fx[f], fz[f]) below the lens) pair: the lens entry
 * point minimising |A_e - P(alpha)|/c1 + |P(alpha) - F|/c2 over the lens of the current setup
 * (h_from_alpha / dh_from_alpha, main_rt.py:180-214): kmin = discrete argmin over the lens
 * samples, tof = travel time at the refined stationary point, alpha = refined entry angle
 * (may be NULL).  Element-major outputs: tof[e*F + f].  Bit-identical to the oracle's CPU twin of the same
 * operations (oracle/tofrt_oracle.c: tofo_fermat2d). */
int tofrt_fermat_dev(tofrt_ctx *ctx, const double *d_fx, const double *d_fz, long long F, double *d_tof,
                     int *d_kmin, double *d_alpha, void *stream);
int tofrt_fermat_host(tofrt_ctx *ctx, const double *fx, const double *fz, long long F, double *tof,
                      int *kmin, double *alpha);

/* 3-D two-point solve (BASELINE.json config 5 "32x32 matrix array, 3-D volume
 * through a curved surface").  The reference is strictly 2-D (main_rt.py works in the x-z plane
 * only), so no line of it corresponds to these entry points: PARITY UNPINNED.  What is reused is the
 * lens profile of the current tofrt_setup (h_from_alpha, main_rt.py:180-189; or the flat-interface
 * option): the 3-D lens is its surface of revolution about the z axis, P = h(alpha) n with
 * n(a, b) = (sin a cos b, sin b, cos a cos b), cos(alpha) = cos a cos b -- b = 0 is the reference's
 * plane with a = alpha.  For every (element e at (ex[e], ey[e], d), voxel (fx, fy, fz)) pair the
 * entry point minimising |A_e - P|/c1 + |P - F|/c2 is found by a discrete search over the
 * n_side x n_side grid of surface samples (ang[i], ang[j]) followed by a projected Newton
 * refinement inside [ang[0], ang[n_side-1]]^2.
 *   tofrt_fermat3d_setup   host pointers; builds the per-element sample tables on the device.
 *   tofrt_fermat3d_dev     device pointers, asynchronous on `stream`.  Element-major outputs:
 *                          tof[e*F + f]; kmin[e*F + f] = ia*n_side + ib of the best grid sample
 *                          (FP32 search: on a near-tie of two samples either may be reported;
 *                          the refined result is the same); ab[2*(e*F + f) + {0,1}] = refined
 *                          entry angles.  kmin and ab may be NULL.
 *   tofrt_fermat3d_host    host pointers, chunked and pipelined like tofrt_table_host.
 * In-plane rule: a pair whose element and voxel both have y == 0 IS the 2-D problem of the current setup and is
 * solved by tofrt_fermat_dev's arithmetic (its K lens samples, its refinement), so the plane y = 0 of a
 * matrix-array table equals the linear-array table bit for bit (SURVEY.md 8(d), config 5's consistency pin); such
 * a pair reports kmin = -1 - j (j = its 2-D lens sample) and ab = (alpha, 0).  The rule applies when the 3-D
 * aperture equals the 2-D one (ang[0] == alpha[0], ang[n_side-1] == alpha[K-1]); otherwise every pair takes the
 * 3-D search.
 * Checked against tests/fermat3d_oracle.py (NumPy/SciPy), the oracle's bit-exact twin, the 2-D
 * solver in the plane y = 0 and analytic known answers (tests/test_gpu_fermat3d.py). */
int tofrt_fermat3d_setup(tofrt_ctx *ctx, const double *ang, int n_side, const double *ex, const double *ey,
                         int E3);
int tofrt_fermat3d_dev(tofrt_ctx *ctx, const double *d_fx, const double *d_fy, const double *d_fz, long long F,
                       double *d_tof, int *d_kmin, double *d_ab, void *stream);
int tofrt_fermat3d_host(tofrt_ctx *ctx, const double *fx, const double *fy, const double *fz, long long F,
                        double *tof, int *kmin, double *ab);

/* Test hook: apply the return-path search of find_line_curve_intersection (main_rt.py:78-99,
 * current setup, current TOFRT_OPT_SEARCH mode) to n arbitrary lines y = m[i] x + b[i].  Runs the
 * production kernel with the reflected lines replaced; idx[i] is encoded as in tofrt_ray_records;
 * xi/zi (nullable) receive the intersection point of main_rt.py:106-168 (NaN = None). */
int tofrt_debug_search_host(tofrt_ctx *ctx, const double *m, const double *b, long long n, int *idx,
                            double *xi, double *zi);

/* Roofline denominator, measured on the device the context lives on: FP64
 * instructions per second (thread-level) of a DFMA stream and of the un-fused
 * DMUL+DADD stream the search loop issues (SURVEY.md section 8 d). */
int tofrt_measure_fp64_peak(tofrt_ctx *ctx, double *dfma_per_s, double *dmul_dadd_per_s);
/* The same for the FP32 mode (tofrt_table_f32_dev): FFMA stream and un-fused FMUL+FADD stream. */
int tofrt_measure_fp32_peak(tofrt_ctx *ctx, double *ffma_per_s, double *fmul_fadd_per_s);

/* Kernel-side facts for benchmarks/tests. */
int tofrt_kernel_info(tofrt_ctx *ctx, int *grid, int *block, int *smem_bytes, int *regs);

#ifdef __cplusplus
}
#endif
#endif /* TOFRT_H */
