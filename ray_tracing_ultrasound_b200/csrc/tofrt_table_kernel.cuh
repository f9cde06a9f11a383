// tofrt_table_kernel.cuh -- the hot kernel: pulse-echo trace (main_rt.py:351-405) + table assembly (:487-501), FP64.
// Part of the single translation unit tofrt_kernels.cu (included inside its anonymous namespace).
#pragma once

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

