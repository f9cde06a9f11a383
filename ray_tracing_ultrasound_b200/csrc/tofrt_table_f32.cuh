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

