// tofrt_table_f32.cuh -- FP32 mode of the table kernel (reported separately, no parity claim).
// Part of the single translation unit tofrt_kernels.cu (included inside its anonymous namespace).
#pragma once

// =====================================================================================
// FP32 mode (north_star: "FP32 mode reported separately with its measured max error").
// Same pipeline, per-ray arithmetic in float (interface laws in vector form, MUFU square roots and
// reciprocals); the geometry-independent tables are the FP64 ones rounded to float.  No bit-parity is claimed for
// this mode: tests/bench report max |dt| and the hit-decision differences against FP64.
// =====================================================================================
struct DevSetupF {
    float c1, c2, fourAC, inv_twoA, twod, K2, pt, c1_over_c2;   // lens constants as the vector form uses them (tofrt_vec_lens)
    float rtol, atol, tol_max, rx_min, rx_max, X, Z;
    int K, C, E, E_sorted, has_nonfinite_rx, shape, log2n, n_shooters, surface;
    float h0;
    float2 P0, PN;             // first and last lens sample
    const float2 *lens, *seg;
    const float *ray_t;        // [S][C][kRayPlanes][32]: the FP64 kernel's per-ray tables, rounded to float
    const float *leg1, *x_rx, *xs;
    const int *perm;
    const int *bucket;         // the FP64 kernel's receiver cell table (its cells are wide against float rounding)
    int nb;
    float bx0, binv;
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

// The FP64 table kernel's pipeline (vector form, guided search with the predicted bracket, cell table) in float.
constexpr int kThreadsF32 = 1024, kWarpsF32 = kThreadsF32 / 32;   // 64 registers: 32 warps per SM
__global__ void __launch_bounds__(kThreadsF32, 1) tofrt_table_kernel_f32(DevSetupF S, TableArgsF T)
{
    constexpr int kThreads = kThreadsF32, kWarps = kWarpsF32;   // shadow the scan kernel's launch constants
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = S.K, N = S.K, E = S.E, C = S.C;
    float2 *s_lens = reinterpret_cast<float2 *>(smem_raw);
    float2 *s_seg = s_lens + N;
    float *s_xs = reinterpret_cast<float *>(s_seg + N);
    float *s_tof = s_xs + S.E_sorted;                 // [kWarps][E]
    int *s_perm = reinterpret_cast<int *>(s_tof + (size_t)kWarps * E);
    int *s_khit = s_perm + S.E_sorted;                // [kWarps][E]
    int *s_bucket = s_khit + (size_t)kWarps * E;      // [nb]
    for (int i = threadIdx.x; i < N; i += kThreads) s_lens[i] = S.lens[i];
    for (int i = threadIdx.x; i < N; i += kThreads) s_seg[i] = i < N - 1 ? S.seg[i] : make_float2(0.f, 0.f);
    for (int i = threadIdx.x; i < S.E_sorted; i += kThreads) { s_xs[i] = S.xs[i]; s_perm[i] = S.perm[i]; }
    for (int i = threadIdx.x; i < S.nb; i += kThreads) s_bucket[i] = S.bucket[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *w_tof = s_tof + (size_t)warp * E;
    int *w_khit = s_khit + (size_t)warp * E;
    const unsigned FULL = 0xffffffffu;
    const float sh = (float)S.shape;
    const int sflip = S.shape < 0 ? (int)0x80000000 : 0;
    for (;;) {  // per-warp work fetch (the CTA-level fetch of the FP64 kernel measured 5 % slower here)
        unsigned long long g64 = 0;
        if (lane == 0) g64 = atomicAdd(T.counter, 1ull);
        g64 = __shfl_sync(FULL, g64, 0);
        if (g64 >= (unsigned long long)T.G * (unsigned long long)T.s_count) break;
        const long long w = (long long)g64;
        const long long g = w % T.G;
        const int shooter = T.s_begin + (int)(w / T.G);
        const float *ray_t = S.ray_t + (size_t)shooter * C * (kRayPlanes * 32) + lane;
        const float r_outer = (float)T.radius[g], pipe_offset = (float)T.offset[g];
        const float rsq = r_outer * r_outer, rsq2 = rsq * rsq;
        for (int e = lane; e < E; e += 32) { w_khit[e] = kNoHit; w_tof[e] = 0.0f; }
        __syncwarp();
        int h1 = -1, h2 = -1;
        for (int c = 0; c < C; ++c) {
            const int k = min((c << 5) + lane, K - 1);
            const float *rt = ray_t + c * (kRayPlanes * 32);
            const float xp = __ldg(rt + kPlaneXp * 32), zp = __ldg(rt + kPlaneZp * 32);
            const float cpq = __ldg(rt + kPlaneCpq * 32), spq = __ldg(rt + kPlaneSpq * 32);
            const float z_f = __ldg(rt + kPlaneZf * 32);
            // pipe intersection (tofrt_vec_pipe) and reflection (tofrt_vec_reflect), in float
            float x_q, z_q, cl, sl, m, b;
            {
                // (discriminant from the perpendicular distance p of the pipe centre to the ray, r^2 - p^2: in float,
                // hh^2 - (|w|^2 - r^2) would cancel five of its seven digits)
                const float wx = xp - pipe_offset;
                const float hh = fmaf(cpq, wx, spq * zp);
                const float pd = fmaf(cpq, zp, -(spq * wx));
                const float t = copysignf(fsqrt_fast(fmaf(-pd, pd, rsq)), spq) - hh;
                x_q = fmaf(t, cpq, xp);
                z_q = fmaf(t, spq, zp);
                const float xx = x_q * x_q, ss = rsq - xx;
                const float s = fsqrt_fast(ss);
                const float c2 = ss - xx, s2 = (-2.0f * x_q) * s;
                cl = fmaf(c2, cpq, s2 * spq);
                sl = fmaf(s2, cpq, -(c2 * spq));
                m = __fdividef(sl, cl);
                b = fmaf(-m, x_q, z_q);
            }
            const bool fin = fabsf(m) < 1e18f && fabsf(b) < 1e18f;
            int idx = -1;
            bool need_scan = fin, check_on_sample = false;
            if (S.shape != 0) {  // the guided search of the FP64 kernel, in float (tol = 2^-19 (...))
                const float tol = (fabsf(m) * S.X + fabsf(b) + S.Z) * 1.9073486328125e-06f /* 2^-19 */ + 1e-30f;
                const float d0 = line_diff_f(S.P0, m, b);
                const float g0 = sh * d0, gN = sh * line_diff_f(S.PN, m, b);
                const bool pos0 = g0 > tol, neg0 = g0 < -tol;
                bool bracket = false, none = false;
                float edge = 0.f;
                if (fin && pos0 && gN > tol) { none = true; edge = fminf(g0, gN) - tol; }
                else if (fin && (pos0 || neg0)) bracket = true;
                const float sgn = pos0 ? sh : -sh;
                const float ms_eff = pos0 ? -INFINITY : sh * m;
                const float cm = sgn * m, cbt = sgn * b + tol;
                auto cond_at = [&](int j) {
                    const float2 pm = s_lens[j];
                    const float ss = __int_as_float(__float_as_int(s_seg[j].x) ^ sflip);
                    return (sgn * pm.y > fmaf(cm, pm.x, cbt)) && (ss > ms_eff);
                };
                int lo = 0, hi = N - 1;
                bool predicted = false;
                if (__all_sync(FULL, !bracket || (h1 >= 0 && h2 >= 0))) {
                    const int p = 2 * h1 - h2;
                    const int plo = min(max(p - 4, 0), N - 2), phi = min(plo + 8, N - 1);
                    const bool okp = cond_at(plo) && !cond_at(phi);
                    if (__all_sync(FULL, !bracket || okp)) { lo = plo; hi = phi; predicted = true; }
                }
                if (predicted) {
#pragma unroll
                    for (int it = 0; it < 3; ++it) {
                        const int mid = (lo + hi) >> 1;
                        const bool cond = cond_at(mid);
                        lo = cond ? mid : lo;
                        hi = cond ? hi : mid;
                    }
                } else {
                    for (int it = 0; it < S.log2n; ++it) {
                        const int mid = (lo + hi) >> 1;
                        const bool cond = cond_at(mid);
                        const bool act = hi - lo > 1;
                        lo = (act && cond) ? mid : lo;
                        hi = (act && !cond) ? mid : hi;
                    }
                }
                h2 = h1;
                h1 = bracket ? lo : -1;
                if (bracket && hi - lo > 1) bracket = false;
                if (bracket && neg0) {
                    const float2 ph = s_lens[hi];
                    if (sgn * ph.y > fmaf(cm, ph.x, cbt)) {
                        const float ghi = sh * line_diff_f(ph, m, b), glo = sh * line_diff_f(s_lens[lo], m, b);
                        bracket = false; none = true; edge = -fmaxf(ghi, glo) - tol;
                    }
                }
                if (bracket) {
                    const float dn = line_diff_f(s_lens[lo + 1], m, b);
                    const bool same = d0 > 0.f ? dn > 0.f : dn < 0.f;
                    if (!same) idx = lo;
                    else if (lo + 2 < N) idx = scan_exact_from_f(s_lens, N, m, b, lo + 1);
                    need_scan = false;
                    check_on_sample = idx < 0;
                } else if (none) {
                    need_scan = false;
                    check_on_sample = !(edge > 1e-8f);
                }
            }
            if (need_scan) { idx = scan_exact_from_f(s_lens, N, m, b, 0); check_on_sample = idx < 0; }
            else if (!fin) idx = (m != m || b != b) ? 0 : scan_exact_from_f(s_lens, N, m, b, 0);
            // chord intersection (main_rt.py:106-168), every lane
            float ix, iz;
            {
                const int j = max(idx, 0);
                const float2 p1 = s_lens[j], p2 = s_lens[j + 1], sg = s_seg[j];
                const float x_int = __fdividef(sg.y - b, m - sg.x), y_int = m * x_int + b;
                bool ok = x_int >= fminf(p1.x, p2.x) - 1e-9f && x_int <= fmaxf(p1.x, p2.x) + 1e-9f &&
                          y_int >= fminf(p1.y, p2.y) - 1e-9f && y_int <= fmaxf(p1.y, p2.y) + 1e-9f;
                float rx = x_int, ry = y_int;
                if (np_isclose_f(p1.x, p2.x, 1e-5f, 1e-8f)) {
                    ry = m * p1.x + b; rx = p1.x;
                    ok = ry >= fminf(p1.y, p2.y) - 1e-9f && ry <= fmaxf(p1.y, p2.y) + 1e-9f;
                } else if (np_isclose_f(m, sg.x, 1e-5f, 1e-8f)) {
                    ok = np_isclose_f(b, sg.y, 1e-5f, 1e-8f);
                    rx = (p1.x + p2.x) / 2.0f; ry = m * rx + b;
                }
                ix = (ok && idx >= 0) ? rx : NAN;
                iz = (ok && idx >= 0) ? ry : NAN;
            }
            if (idx < 0 && check_on_sample) {
                for (int j = 0; j < N; ++j)
                    if (fabsf(line_diff_f(s_lens[j], m, b)) <= 1e-8f) { ix = s_lens[j].x; iz = s_lens[j].y; break; }
            }
            // refraction back into the lens and landing point (tofrt_vec_return_head / _tail), in float
            float target_x = NAN;
            if (__any_sync(FULL, ix == ix)) {
                float Tz, Tx;
                if (S.surface == 0) {
                    const float rho2 = fmaf(ix, ix, iz * iz);
                    const float rho = fsqrt_fast(rho2);
                    const float Br = fmaf(S.twod, iz, -(S.K2 * rho));
                    const float sdr = fsqrt_fast(fmaf(Br, Br, -(S.fourAC * rho2)));
                    const float sum = sdr + Br;
                    const float DH = (S.pt * ix) * sum;
                    const float HH = (-sum * S.inv_twoA) * sdr;
                    Tz = fmaf(DH, iz, -(HH * ix));
                    Tx = fmaf(DH, ix, HH * iz);
                } else {
                    Tz = 0.0f * (ix + iz);
                    Tx = 1.0f + Tz;
                }
                const float Nn = -S.c1_over_c2 * fmaf(cl, Tx, sl * Tz);
                const float M2 = fmaf(Tx, Tx, Tz * Tz) * rsq2;
                const float w2 = fmaf(-Nn, Nn, M2);
                if (__any_sync(FULL, w2 >= 0.0f)) {
                    const float W = fsqrt_fast(w2);   // NaN on total internal reflection
                    const float den = fmaf(Tz, W, Tx * Nn), num = fmaf(Tz, Nn, -(Tx * W));
                    target_x = fmaf(z_f - iz, __fdividef(den, num), ix);
                }
            }
            // table assembly: cell table of the FP64 kernel (cells ~1/16 receiver spacing: wide against float rounding)
            int lo = 0, hi = 0;
            if (S.nb > 0) {
                const float cpos = (target_x - S.bx0) * S.binv;
                const int cell = cpos > 0.0f ? (cpos < (float)(S.nb - 1) ? (int)cpos : S.nb - 1) : 0;
                const int packed = s_bucket[cell];
                lo = packed & 0xffff;
                hi = lo + (int)((unsigned)packed >> 16);
            } else if (target_x >= S.rx_min - S.tol_max && target_x <= S.rx_max + S.tol_max) {
                const float tl = target_x - S.tol_max, th = target_x + S.tol_max;
                int r = S.E_sorted;
                while (lo < r) { int mid = (lo + r) >> 1; if (s_xs[mid] < tl) lo = mid + 1; else r = mid; }
                hi = lo;
                while (hi < S.E_sorted && s_xs[hi] <= th) ++hi;
            }
            const bool special = S.has_nonfinite_rx && isinf(target_x);
            if (__any_sync(FULL, hi > lo || special)) {
                bool mine = false;
                for (int i = lo; i < hi; ++i)
                    if (np_isclose_f(target_x, s_xs[i], S.rtol, S.atol)) { atomicMin(&w_khit[s_perm[i]], k); mine = true; }
                if (special)
                    for (int e = 0; e < E; ++e)
                        if (target_x == S.x_rx[e]) { atomicMin(&w_khit[e], k); mine = true; }
                __syncwarp();
                if (mine) {
                    float tof = S.leg1[(size_t)shooter * K + k];
                    tof += dist2d_f(xp, zp, x_q, z_q) / S.c2;
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
