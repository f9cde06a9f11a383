// tofrt_fermat_kernels.cuh -- two-point Fermat solves (extensions): 2-D on the sampled lens, 3-D through the lens of revolution.
// Part of the single translation unit tofrt_kernels.cu (included inside its anonymous namespace).
#pragma once

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

// Continuous refinement of one pair from its discrete minimiser: safeguarded secant (Illinois) on dt/dalpha inside
// the two sample intervals around lens sample jbest.  Shared by the 2-D kernel and the in-plane pairs of the 3-D one.
__device__ __forceinline__ void fermat2d_refine(const DevSetup &S, const double *alpha, int N, int jbest, double ax,
                                                double az, double fx, double fz, double *tof_out, double *al_out)
{
    double a = alpha[max(jbest - 1, 0)], b = alpha[min(jbest + 1, N - 1)];
    double ta, ga, tb, gb, tc, gc;
    fermat_eval(S, a, ax, az, fx, fz, &ta, &ga);
    fermat_eval(S, b, ax, az, fx, fz, &tb, &gb);
    double al = alpha[jbest], tof;
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
    *tof_out = tof; *al_out = al;
}

// One whole pair of the 2-D solve without the kernel's staging: exact FP64 argmin over the lens samples in global
// memory (|A - P_j| / c1 evaluated on the fly with the expression of tofrt_fermat_d1_kernel), then the refinement.
// Same operations, hence the same bits, as tofrt_fermat_kernel computes for this pair.  Used by the 3-D kernel for
// pairs that lie in the plane y = 0 (out of line: its registers are not the 3-D kernel's).
__device__ __noinline__ double2 fermat2d_pair(const DevSetup &S, const double *alpha, double ax, double az, double fx,
                                              double fz, int *jbest_out)
{
    const int N = S.K;
    const double inv_c2 = 1.0 / S.c2;
    double best = INFINITY;
    int jbest = 0;
    for (int j = 0; j < N; ++j) {
        const double2 q = S.lens[j];
        const double d1 = dist2d(ax, az, q.x, q.y) / S.c1;
        const double ux = q.x - fx, uz = q.y - fz;
        const double t = fma(sqrt(fma(ux, ux, uz * uz)), inv_c2, d1);
        if (t < best) { best = t; jbest = j; }
    }
    double tof, al;
    fermat2d_refine(S, alpha, N, jbest, ax, az, fx, fz, &tof, &al);
    *jbest_out = jbest;
    return make_double2(tof, al);
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
        double al, tof;
        fermat2d_refine(S, A.alpha, N, jbest, ax, az, fx, fz, &tof, &al);
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
    // In-plane pairs (element and voxel both at y = 0) are the 2-D problem of the current 2-D setup and are solved by
    // its arithmetic, so that the plane y = 0 of a matrix-array table IS the linear-array table, bit for bit (SURVEY
    // 8(d) config 5's consistency pin): alpha2d = the 2-D setup's lens angles (null: the apertures differ, no dispatch).
    // Such a pair reports kmin = -1 - j (j = its 2-D lens sample) and ab = (alpha, 0).
    const double *alpha2d;
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

__global__ void __launch_bounds__(kF3dThreads, TOFRT_F3D_MIN_BLOCKS)
    tofrt_fermat3d_kernel(tofrt_surf3d S, F3dArgs A, const __grid_constant__ DevSetup S2)
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
        if (A.alpha2d && Ap[1] == 0.0 && Fp[1] == 0.0) {  // in-plane pair: the 2-D solve (see F3dArgs)
            int j2;
            const double2 r = fermat2d_pair(S2, A.alpha2d, Ap[0], Ap[2], Fp[0], Fp[2], &j2);
            const size_t o2 = (size_t)e * A.F + f;
            A.tof[o2] = r.x;
            if (A.kmin) A.kmin[o2] = -1 - j2;
            if (A.ab) reinterpret_cast<double2 *>(A.ab)[o2] = make_double2(r.y, 0.0);
            continue;
        }
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

