// tofrt_table_kernel.cuh -- the hot kernel: pulse-echo trace (main_rt.py:351-405) + table assembly (:487-501), FP64.
// Part of the single translation unit tofrt_kernels.cu (included inside its anonymous namespace).
#pragma once

// ---- exact (slow) pieces of find_line_curve_intersection ------------------------

// first j in [j0, N-2] with a sign-class change between samples j and j+1, else -1 (one thread, one ray)
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

// The reference's scan, exactly (np.sign classes incl. zero and NaN): first j in [j0, N-2] with a class change between
// samples j and j+1, else -1 -- for ONE line, evaluated by the whole warp (32 sample pairs per step).  The rare rays
// no cheaper argument decides come here one after the other in the small-G mode, where a per-thread loop over up to N
// samples (scan_exact_from: 900 dependent iterations, ~45 us) would be the critical path of the whole call; in the
// throughput kernel such a ray is one slow warp among thousands and the per-thread loop keeps the hot stream leaner
// (measured: the cooperative form there costs 1.5 %).
__device__ __noinline__ int warp_scan_exact(const double2 *lens, int N, double m, double b, int j0, int lane)
{
    for (int base = j0; base + 1 < N; base += 32) {
        const int j = base + lane;
        bool chg = false;
        if (j + 1 < N) chg = class_change(sign_class(line_diff(lens[j], m, b)), sign_class(line_diff(lens[j + 1], m, b)));
        const unsigned mask = __ballot_sync(0xffffffffu, chg);
        if (mask) return base + __ffs(mask) - 1;
    }
    return -1;
}

// test hook (tofrt_debug_search_host): the reflected line of ray [w][k] replaced by a given one.  Out of line so
// that the production stream carries one uniform branch for it and nothing else.
__device__ __noinline__ double2 inject_line(const double *inject_m, const double *inject_b, size_t o)
{
    return make_double2(inject_m[o], inject_b[o]);
}

// main_rt.py:106-168: intersect the line with chord (idx, idx+1).
// PLAIN (uniform, decided per setup on the actual table): no chord is "vertical" in the sense of :110
// (np.isclose(x1, x2) is false for every j), the abscissae increase and every sample is finite -- the same
// decisions and the same bits with fewer tests per ray.
template <bool PLAIN>
__device__ __forceinline__ void chord_intersection(const double2 *lens, const double2 *seg, int idx, double m,
                                                   double b, double *xi, double *zi)
{
    const double2 p1 = lens[idx], p2 = lens[idx + 1], sg = seg[idx];
    const double x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
    const double m_seg = sg.x, b_seg = sg.y;  // :127-128
    if (PLAIN) {
        // :147-168, the common case
        const double x_int = (b_seg - b) / (m - m_seg);
        const double y_int = m * x_int + b;
        const bool up = y2 > y1;  // finite table entries: one comparison orders them like py_min / py_max
        const double y_lo = up ? y1 : y2, y_hi = up ? y2 : y1;
        bool ok = x_int >= x1 - 1e-9 && x_int <= x2 + 1e-9 && y_int >= y_lo - 1e-9 && y_int <= y_hi + 1e-9;
        double rx = x_int, ry = y_int;
        if (np_isclose(m, m_seg, 1e-5, 1e-8)) {  // :131-144 parallel (rare: a separate, divergent path)
            ok = np_isclose(b, b_seg, 1e-5, 1e-8);
            rx = (x1 + x2) / 2.0;
            ry = m * rx + b;
        }
        *xi = ok ? rx : NAN;
        *zi = ok ? ry : NAN;
        return;
    }
    double rx = NAN, ry = NAN;
    if (np_isclose(x1, x2, 1e-5, 1e-8)) {  // :110-123 vertical chord
        double y_int = m * x1 + b;
        if (y_int >= py_min(y1, y2) - 1e-9 && y_int <= py_max(y1, y2) + 1e-9) { rx = x1; ry = y_int; }
    } else if (np_isclose(m, m_seg, 1e-5, 1e-8)) {  // :131-144 parallel
        if (np_isclose(b, b_seg, 1e-5, 1e-8)) { rx = (x1 + x2) / 2.0; ry = m * rx + b; }
    } else {  // :147-168
        double x_int = (b_seg - b) / (m - m_seg);
        double y_int = m * x_int + b;
        if (x_int >= py_min(x1, x2) - 1e-9 && x_int <= py_max(x1, x2) + 1e-9 &&
            y_int >= py_min(y1, y2) - 1e-9 && y_int <= py_max(y1, y2) + 1e-9) {
            rx = x_int; ry = y_int;
        }
    }
    *xi = rx; *zi = ry;
}

// ---- the table kernel -----------------------------------------------------------
// HG: first-hit state in the output rows (global memory) instead of warp-private shared memory
// GUIDED: compile the guided search (its rare fallback is then a plain per-ray exact scan) or the
// warp-wide literal scan -- two lean instruction streams instead of one kernel carrying both.
// VEC: the return-path arithmetic in vector form (tofrt_raymath.h) instead of the reference's angles
// STATS: count the executed work (rays per stage, probes, exact diffs) into T.stats -- an instrumented twin of
// the production instantiation, used by bench.py outside the timed region (roofline.executed)
//
// Ray order.  A warp owns one (shooter, geometry) and walks its K rays 32 at a time, lane l of chunk c taking ray
// 32 c + l (indices beyond K-1 repeat ray K-1, which is idempotent: same landing point, same atomicMin), so the
// lanes of a warp hold neighbouring rays: they probe neighbouring lens samples (shared-memory broadcast), die
// together (rays that miss the lens or are totally reflected come in runs) and take the same branches.  The
// per-ray tables are stored chunk by chunk, [c][plane][32] (tofrt_ray_table_kernel): every load is one coalesced
// line at a constant offset from the chunk's base.
// (Measured and rejected: lane l walking the CONSECUTIVE rays l C .. l C + C-1, which predicts the next bracket to
// +-1 sample from the lane's own previous ray -- the search gets cheaper, but scattered shared-memory probes, the
// slow paths of the 4 % dead rays now present in every warp and a 2x slower literal scan cost more: 29.8 vs 25.2 ms.)
//
// SPLIT (small-G latency mode): a whole CTA owns one (shooter, geometry) at a time and its warps take one 32-ray
// chunk each (requires ceil(K / 32) <= warps per CTA; warps beyond the last chunk repeat it, which is idempotent),
// with a CTA-wide first-hit reduction in shared memory.  With fewer geometries than resident warps -- the
// reference's own 210-geometry sweep -- this turns 29 serial chunks per warp into one: 0.23 ms -> tens of us.
template <bool HG, bool GUIDED, bool FLAT, bool VEC, bool STATS, bool SPLIT = false>
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
    int *s_bucket = s_khit + (HG ? (size_t)0 : (size_t)kWarps * E);  // [nb] packed (first, count)
    // small-G mode: queue of rays for the CTA-wide on-sample pass (kSplitQueue lines + results, after the cell table)
    double2 *s_qmb = reinterpret_cast<double2 *>((reinterpret_cast<uintptr_t>(s_bucket + S.nb) + 15) & ~(uintptr_t)15);
    int *s_qres = reinterpret_cast<int *>(s_qmb + kSplitQueue);

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
    tof_t *w_tof = s_tof + (SPLIT ? (size_t)0 : (size_t)warp * E);
    khit_t *w_khit = s_khit + (SPLIT ? (size_t)0 : (size_t)warp * E);
    const int C = S.C;               // 32-ray chunks per shot
    const unsigned FULL = 0xffffffffu;
    const bool rare_path = S.has_nonfinite_rx != 0 || T.rec.legs != nullptr;
    tofrt_vec_lens VL = S.VL;  // lens constants of the vector form (uniform)
    VL.flat = FLAT ? 1 : 0;
    // executed-work counters (STATS only): rays, rays with a bracket, cond probes, exact diffs, rays that entered
    // the lens again, rays that reached the landing point, receiver candidates tested, winners (legs computed),
    // rays in a fallback bisection, rays in the exact O(N) scan, on-sample passes
    unsigned st[kNumStats];
#pragma unroll
    for (int i = 0; i < kNumStats; ++i) st[i] = 0;

    // Work distribution: the CTA takes kWarps consecutive (shooter, geometry) items at a time, one per
    // warp, and its warps start them together (two block barriers per ~40 us of work).  Warps that
    // begin a geometry in step run the same code at the same time, which is worth 3.7 % over
    // fetching per warp (the instruction stream of one chunk is ~40 KB); any tighter coupling
    // (a barrier every chunk, or every few) costs more than it gives -- measured.
    __shared__ unsigned long long s_base;
    const unsigned long long total = (unsigned long long)T.G * (unsigned long long)T.s_count;
    for (;;) {
        __syncthreads();  // everyone has read the previous s_base
        if (threadIdx.x == 0) s_base = atomicAdd(T.counter, (unsigned long long)(SPLIT ? 1 : kWarps));
        __syncthreads();
        if (s_base >= total) break;
        const unsigned long long g64 = s_base + (SPLIT ? 0 : warp);
        if (g64 >= total) continue;  // tail: fewer items than warps
        const long long w = (long long)g64;               // work item = (shooter, geometry)
        long long g = w, s_local = 0;
        if (T.s_count > 1) { s_local = w / T.G; g = w - s_local * T.G; }  // 64-bit division only when several shooters
        const int shooter = T.s_begin + (int)s_local;
        const double *ray_t = S.ray_t + (size_t)shooter * C * (kRayPlanes * 32) + lane;  // [c][plane][lane]
        const double r_outer = T.radius[g], pipe_offset = T.offset[g];
        const double rsq = r_outer * r_outer, offsq = pipe_offset * pipe_offset, rsq2 = rsq * rsq;
        if (HG) { w_tof = T.tof + (size_t)w * E; w_khit = T.k_hit + (size_t)w * E; }
        for (int e = SPLIT ? (int)threadIdx.x : lane; e < E; e += SPLIT ? kThreads : 32) { w_khit[e] = kNoHit; w_tof[e] = 0.0; }
        if (SPLIT) __syncthreads(); else __syncwarp();

        int h1 = -1, h2 = -1;  // guided search: this lane's bracket starts in the previous two chunks
        for (int c = SPLIT ? min(warp, C - 1) : 0; c < C; c += SPLIT ? C : 1) {
            const int k = min((c << 5) + lane, K - 1);
            // per-ray constants of the shot: one coalesced, L1-resident load per 32 rays
            const double *rt = ray_t + c * (kRayPlanes * 32);
            const double z_f = __ldg(rt + kPlaneZf * 32);

            double x_q, z_q;
            double phi_l = 0.0, cl = 0.0, sl = 0.0, m, b;  // reflected direction: angle, or r^2 (cos, sin)
            if (VEC) {
                const double xp = __ldg(rt + kPlaneXp * 32), zp = __ldg(rt + kPlaneZp * 32);
                const double cpq = __ldg(rt + kPlaneCpq * 32), spq = __ldg(rt + kPlaneSpq * 32);
                // ---- main_rt.py:351-364 pipe intersection, upper root; :367-376 reflection on the pipe ----
                tofrt_vec_pipe(xp, zp, cpq, spq, pipe_offset, rsq, &x_q, &z_q);
                tofrt_vec_reflect(x_q, rsq, cpq, spq, &cl, &sl);
                m = sl / cl;
                b = fma(-m, x_q, z_q);
            } else {
                const double a_pq = __ldg(rt + kPlaneApq * 32), b_pq = __ldg(rt + kPlaneBpq * 32);
                const double phi_pq = __ldg(rt + kPlanePhi * 32);
                // ---- main_rt.py:351-364 pipe intersection, upper root ----
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
                // ---- main_rt.py:367-376 reflection on the pipe ----
                double slope = -x_q / sqrt(rsq - x_q * x_q);  // dzdx_pipe ignores the offset (:237-238)
                double phi_slope = tofrt_atan(slope);
                double theta_1 = phi_pq - (phi_slope + HALF_PI);
                phi_l = phi_slope - HALF_PI + (-theta_1);
                m = tofrt_tan(phi_l);
                b = z_q - m * x_q;
            }
            if (T.inject_m) { const double2 mb = inject_line(T.inject_m, T.inject_b, (size_t)w * K + k); m = mb.x; b = mb.y; }  // test hook, off in production

            // ---- main_rt.py:78-99 first sign change over the sampled lens ----
            // fast path precondition: every diff is finite, so only {-,0,+} occur
            const bool fast = S.tables_ok && fabs(m) < 1e150 && fabs(b) < 1e150;
            int idx = -1;             // >= 0 segment, -1 none
            int scan_from = -1;       // >= 0: this ray still needs the reference's exact scan from that sample on
            bool need_scan = fast;    // lanes whose index the O(N) scan below still has to find
            bool skip_on_sample = false;
            if (GUIDED) {
                // ---- guided search: same idx as the scan, in O(1) diffs (docs/guided_search.md) --------
                // g_j = shape * diff_j is, in exact arithmetic, a concave sequence over the sampled
                // lens (chord slopes strictly monotone, checked by tofrt_setup), and a computed diff
                // differs from the exact one by < tol = 2^-48 (|m| X + |b| + Z).  So a diff beyond
                // +-tol has a certain sign, a concave sequence is >= the smaller of two of its values
                // in between, and it is monotone on either side of its maximum: that brackets the
                // first sign change to [lo, lo+1] with every earlier sign known; the exact test then
                // starts at lo.  Anything uncertain (|diff| <= tol where it matters) falls back to the
                // full scan, so the result is identical by construction.
                const double sh = (double)S.shape;
                const double tol = (fabs(m) * S.X + fabs(b) + S.Z) * 0x1p-48 + 1e-290;
                const double d0 = line_diff(S.P0, m, b);   // first and last lens sample: setup constants
                const double g0 = sh * d0;
                const double gN = sh * line_diff(S.PN, m, b);
                const bool pos0 = g0 > tol, neg0 = g0 < -tol;
                bool bracket = false, certain_none = false;
                double edge = 0.0;  // lower bound of min|diff| when there is certainly no change
                if (fast && pos0 && gN > tol) { certain_none = true; edge = fmin(g0, gN) - tol; }
                else if (fast && (pos0 || neg0)) bracket = true;
                // cond(j): sample j certainly has the sign of g_0 and, when g_0 < 0, lies on the rising side
                // (chord slope above m): then every sample up to j certainly has that sign too.  ANY j
                // with cond(j) is a valid lower end of the bracket, any j without it a valid upper end
                // (the concavity arguments do not depend on how j was chosen, docs/guided_search.md).
                const double sgn = pos0 ? sh : -sh;
                const double ms_eff = pos0 ? -INFINITY : sh * m;
                // "diff_j certainly has the sign of diff_0", one fused operation per probe:
                //   sgn (z - m x - b) > tol   <=   sgn z > fma(sgn m, x, sgn b + tol)
                // (sgn = +-1: the products are exact; the two roundings are < tol/16, see docs/guided_search.md)
                const double cm = sgn * m, cbt = sgn * b + tol;
                const int zflip = sgn < 0.0 ? (int)0x80000000 : 0;
                const int sflip = S.shape < 0 ? (int)0x80000000 : 0;  // uniform: sh * slope by a sign flip
                const char *lens_b = reinterpret_cast<const char *>(s_lens);
                const int slope_off = (int)(reinterpret_cast<const char *>(s_seg) - lens_b);  // uniform
                auto cond_at = [&](int j) {
                    const char *pa = lens_b + j * 16;  // lens[j]; the chord slope sits slope_off bytes further
                    const double2 pm = *reinterpret_cast<const double2 *>(pa);
                    const double slope = *reinterpret_cast<const double *>(pa + slope_off);
                    const double zs = __hiloint2double(__double2hiint(pm.y) ^ zflip, __double2loint(pm.y));
                    const double ss = __hiloint2double(__double2hiint(slope) ^ sflip, __double2loint(slope));
                    if (STATS) st[kStProbes] += bracket;
                    return (zs > fma(cm, pm.x, cbt)) && (ss > ms_eff);
                };
                int lo = 0, hi = N - 1;
                // Predicted bracket: along a shot the entry index is a smooth function of the ray index, so
                // this lane's brackets of the previous two chunks extrapolate to within a few samples
                // (measured on config 3: +-4 covers 96 % of the chunks, +-8 99.6 %).  Two probes certify [p - W, p + W] as a
                // bracket for every lane of the warp, and three straight-line bisection steps finish it;
                // otherwise the full bisection runs.
#ifndef TOFRT_PRED_W  /* measured: W = 4 (3 steps) 7.81 ms, W = 8 (4 steps) 7.91 ms, W = 16 (5 steps) 8.11 ms; none: 8.44 ms */
#define TOFRT_PRED_W 4
#define TOFRT_PRED_STEPS 3
#endif
                constexpr int kPredW = TOFRT_PRED_W, kPredSteps = TOFRT_PRED_STEPS;  // 2 * kPredW <= 2^kPredSteps
                bool predicted = false;
                if (__all_sync(FULL, !bracket || (h1 >= 0 && h2 >= 0))) {
                    const int p = 2 * h1 - h2;
                    const int plo = min(max(p - kPredW, 0), N - 2), phi = min(plo + 2 * kPredW, N - 1);
                    const bool okp = cond_at(plo) && !cond_at(phi);
                    if (__all_sync(FULL, !bracket || okp)) { lo = plo; hi = phi; predicted = true; }
                }
                if (predicted) {
                    // hi - lo <= 2^kPredSteps: kPredSteps halvings reach hi - lo = 1 (a step at hi - lo = 1 probes
                    // mid = lo again, cond(lo) holds, nothing moves)
#pragma unroll
                    for (int it = 0; it < kPredSteps; ++it) {
                        const int mid = (lo + hi) >> 1;
                        const bool cond = cond_at(mid);
                        lo = cond ? mid : lo;
                        hi = cond ? hi : mid;
                    }
                } else {
                    if (STATS) st[kStFallback] += bracket;
                    for (int it = 0; it < S.log2n; ++it) {
                        const int mid = (lo + hi) >> 1;  // <= N-2 whenever hi - lo > 1
                        const bool cond = cond_at(mid);
                        const bool act = hi - lo > 1;
                        lo = (act && cond) ? mid : lo;
                        hi = (act && !cond) ? mid : hi;
                    }
                }
                h2 = h1;
                h1 = bracket ? lo : -1;
                if (STATS) st[kStBracket] += bracket;
                if (bracket && hi - lo > 1) { bracket = false; }  // cannot happen (log2n steps suffice); scan if it did
                if (bracket && neg0) {
                    // hi = lo + 1 failed the test (or is the untested last sample).  If g_hi is still
                    // certainly negative, hi is at/after the maximum while lo is before it: g_hi IS
                    // the maximum of g, so no diff is >= 0.
                    // (lo == 0 was never tested against the slope: if g already falls from the first
                    // sample on, the maximum is g_0 -- so the bound below uses the larger of the two)
                    // ("certainly negative" must be the loop's own probe: hi may have failed exactly that)
                    const double2 ph = s_lens[hi];
                    const double zh = __hiloint2double(__double2hiint(ph.y) ^ zflip, __double2loint(ph.y));
                    if (zh > fma(cm, ph.x, cbt)) {
                        const double ghi = sh * line_diff(ph, m, b);
                        const double glo = sh * line_diff(s_lens[lo], m, b);
                        bracket = false; certain_none = true; edge = -fmax(ghi, glo) - tol;
                    }
                }
                if (bracket) {
                    // Every sample up to lo has a non-zero computed diff with the sign of diff_0 (the invariant),
                    // so the first change is at lo iff diff_{lo+1} does not strictly share that sign (zero, the
                    // other sign or NaN: np.sign differs); otherwise keep going exactly (rare).
                    const double dn = line_diff(s_lens[lo + 1], m, b);
                    if (STATS) st[kStDiffs] += 1;
                    const bool same = d0 > 0.0 ? dn > 0.0 : dn < 0.0;
                    if (!same) idx = lo;
                    else if (lo + 2 < N) { if (SPLIT) scan_from = lo + 1; else idx = scan_exact_from(s_lens, N, m, b, lo + 1); }
                    need_scan = false;  // no change after lo either => none at all (on-sample test decides)
                } else if (certain_none) {
                    need_scan = false;
                    skip_on_sample = edge > 1e-8;  // every |diff| > 1e-8: main_rt.py:86-96 returns None
                }
            }
            int flagged = -1;
            unsigned gmin = 0xffffffffu;  // min over all chunks of the high words (see below)
            if (STATS) { st[kStRays] += 1; st[kStScan] += need_scan; }
            if (GUIDED) {
                // a ray the guided search could not decide (|diff| <= tol where it matters): plain exact scan
                if (need_scan) { if (SPLIT) scan_from = 0; else idx = scan_exact_from(s_lens, N, m, b, 0); flagged = 0; }
            } else if (__any_sync(FULL, need_scan)) {
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
                    if (STATS) st[kStDiffs] += kChunk;
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
            if (idx < 0) {
                if (!fast) {
                    // a NaN line makes every diff NaN: np.sign changes "between" samples 0 and 1 (main_rt.py:82)
                    if (m != m || b != b) idx = 0;
                    else if (SPLIT) scan_from = 0;
                    else idx = scan_exact_from(s_lens, N, m, b, 0);
                }
                // conservative flag without a real change (a denormal diff): keep going exactly
                else if (!GUIDED && need_scan && flagged >= 0 && flagged + kChunk - 1 < N - 1)
                    idx = scan_exact_from(s_lens, N, m, b, flagged + kChunk - 1);
            }
            if (SPLIT) {  // the exact scans, one ray at a time, 32 sample pairs per step (see warp_scan_exact)
                unsigned need = __ballot_sync(FULL, scan_from >= 0);
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    const int r = warp_scan_exact(s_lens, N, __shfl_sync(FULL, m, src), __shfl_sync(FULL, b, src),
                                                  __shfl_sync(FULL, scan_from, src), lane);
                    if (lane == src) idx = r;
                }
            }
            // ---- main_rt.py:84-96 no crossing: is a sample on the line (|diff| <= 1e-8)? ----
            // All diffs of such a ray have one sign, so the unsigned minimum of their high words is
            // the high word of min|diff|: only rays that come within 1e-8 of a sample (hi <= hi(1e-8))
            // need the test.  It is warp-cooperative: 32 lanes take 32 consecutive samples of one ray.
            int on_sample = -1;
            {
                bool maybe = idx < 0 && !skip_on_sample &&
                             (!fast || !need_scan || flagged >= 0 ||
                              (gmin & 0x7fffffffu) <= 0x3e45798eu /* hi32(1e-8) */);
                unsigned need = __ballot_sync(FULL, maybe);
                if (STATS) st[kStOnSample] += maybe;
                if (SPLIT) {
                    // Small-G mode: the rays of one geometry that need this pass usually sit in one or two chunks (the
                    // ones grazing the lens apex), i.e. in one or two warps, while the other ~30 warps of the CTA would
                    // wait at the next barrier: queue them and let every warp take one ray at a time.
                    __shared__ int s_qn;
                    if (threadIdx.x == 0) s_qn = 0;
                    __syncthreads();
                    int myq = -1;
                    if (maybe) {
                        myq = atomicAdd(&s_qn, 1);
                        if (myq < kSplitQueue) s_qmb[myq] = make_double2(m, b);
                    }
                    need = __ballot_sync(FULL, maybe && myq >= kSplitQueue);   // overflow: the per-warp path below
                    __syncthreads();
                    const int nq = min(s_qn, kSplitQueue);
                    for (int q = warp; q < nq; q += kWarps) {
                        const double2 mb = s_qmb[q];
                        int first = kNoHit;
                        for (int j = lane; j < N; j += 32) {
                            double dd = line_diff(s_lens[j], mb.x, mb.y);
                            if (fabs(dd) <= 1e-8) { first = j; break; }
                        }
                        first = __reduce_min_sync(FULL, first);
                        if (lane == 0) s_qres[q] = first;
                    }
                    __syncthreads();
                    if (myq >= 0 && myq < kSplitQueue && s_qres[myq] != kNoHit) on_sample = s_qres[myq];
                }
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
            // chord intersection for every lane (a lane without a crossing computes on chord 0 and discards it:
            // cheaper than a divergent region around a division)
            double ix, iz;
            if (S.chords_plain) chord_intersection<true>(s_lens, s_seg, max(idx, 0), m, b, &ix, &iz);
            else chord_intersection<false>(s_lens, s_seg, max(idx, 0), m, b, &ix, &iz);
            if (idx < 0) {
                ix = NAN; iz = NAN;
                if (on_sample >= 0) { double2 s = s_lens[on_sample]; ix = s.x; iz = s.y; }
            }

            // ---- main_rt.py:396-405 refraction back into the lens, landing point ----
            // NaN in -> NaN out, so whole 32-ray groups that missed the lens, or are totally
            // reflected (|sin| > 1 in main_rt.py:277), skip the rest of the arithmetic.
            double target_x = NAN;
            if (STATS) st[kStLens2] += ix == ix;
            if (VEC) {
                if (__any_sync(FULL, ix == ix)) {
                    double tz, tx, nn, w2;
                    tofrt_vec_return_head(&VL, ix, iz, cl, sl, rsq2, &tz, &tx, &nn, &w2);
                    if (STATS) st[kStLanding] += w2 >= 0.0;
                    if (__any_sync(FULL, w2 >= 0.0)) target_x = tofrt_vec_return_tail(ix, iz, z_f, tz, tx, nn, w2);
                }
            } else if (__any_sync(FULL, ix == ix)) {
                // alpha_i = arctan2(ix, iz) is only ever used through sin/cos (main_rt.py:396-397):
                // one fused elementary function, no angle
                double sn, cs, h, dz, dx;
                tofrt_sincos_atan2(ix, iz, &sn, &cs);
                lens_point_t<FLAT>(S, sn, cs, &h, &dz, &dx);
                // refraction(phi_l, (dz, dx), c2, c1), main_rt.py:272-278
                double phi_slope = tofrt_atan2(dz, dx);
                double theta_1 = phi_l - (phi_slope + HALF_PI);
                double s_arg = S.c1_over_c2 * tofrt_sin(theta_1);
                if (STATS) st[kStLanding] += fabs(s_arg) <= 1.0;
                if (__any_sync(FULL, fabs(s_arg) <= 1.0)) {
                    double theta_2 = tofrt_asin(s_arg);
                    double phi_last = phi_slope - HALF_PI + theta_2;
                    double a_i = tofrt_tan(phi_last);
                    double b_i = iz - a_i * ix;
                    target_x = (z_f - b_i) / a_i;
                }
            }

            // ---- optional per-ray records ----
            if (T.any_rec) {  // one uniform test in table-only mode
                size_t o = (size_t)w * K + k;
                if (T.rec.pipe_x) T.rec.pipe_x[o] = x_q;
                if (T.rec.pipe_z) T.rec.pipe_z[o] = z_q;
                if (T.rec.lens2_x) T.rec.lens2_x[o] = ix;
                if (T.rec.lens2_z) T.rec.lens2_z[o] = iz;
                if (T.rec.target_x) T.rec.target_x[o] = target_x;
                if (T.rec.idx) T.rec.idx[o] = idx >= 0 ? idx : (on_sample >= 0 ? -2 - on_sample : -1);
            }

            // ---- main_rt.py:487-501 which receivers does this ray land on? ----
            // Candidates: finite receivers within tol_max of target_x, from a cell table over the aperture
            // (cell c lists every receiver that any landing point whose computed cell index is c could match:
            // most cells list none, so most rays leave after one lookup); the exact np.isclose decides.
            // +-inf landing points can only match an identical infinite receiver.
            int lo = 0, hi = 0;
            if (S.nb > 0) {
                const double cpos = (target_x - S.bx0) * S.binv;
                // NaN -> cell 0; out-of-range landing points -> the (empty) guard cells at either end
                const int cell = cpos > 0.0 ? (cpos < (double)(S.nb - 1) ? (int)cpos : S.nb - 1) : 0;
                const int packed = s_bucket[cell];
                lo = packed & 0xffff;
                hi = lo + (int)((unsigned)packed >> 16);
            } else if (target_x >= S.rx_min - S.tol_max && target_x <= S.rx_max + S.tol_max) {
                // no cell table (few receivers, or a tolerance as wide as the aperture): binary search
                const double tl = target_x - S.tol_max, th = target_x + S.tol_max;
                int r = S.E_sorted;
                while (lo < r) { int mid = (lo + r) >> 1; if (s_xs[mid] < tl) lo = mid + 1; else r = mid; }
                hi = lo;
                while (hi < S.E_sorted && s_xs[hi] <= th) ++hi;
            }
            bool special = false, want_legs = false;
            if (rare_path) {  // uniform: infinite receivers or the per-ray leg records -- neither in a production table call
                special = S.has_nonfinite_rx && isinf(target_x);
                want_legs = T.rec.legs != nullptr;
            }
            if (SPLIT || rare_path || __any_sync(FULL, hi > lo)) {  // (SPLIT: a CTA barrier sits inside)
                bool mine = false;
                // np.isclose(target_x, x_e): |target_x - x_e| <= atol + rtol |x_e| for a finite x_e (the sorted copy
                // holds finite receivers only), or equality.  Usually one candidate: straight-line, the rest in a loop.
                if (hi > lo) {
                    if (STATS) st[kStCand] += hi - lo;
                    const double xe = s_xs[lo];
                    if (fabs(target_x - xe) <= S.atol + S.rtol * fabs(xe) || target_x == xe) {
                        atomicMin(const_cast<int *>(&w_khit[s_perm[lo]]), k); mine = true;
                    }
                    for (int i = lo + 1; i < hi; ++i)
                        if (np_isclose(target_x, s_xs[i], S.rtol, S.atol)) { atomicMin(const_cast<int *>(&w_khit[s_perm[i]]), k); mine = true; }
                }
                if (special) {
                    for (int e = 0; e < E; ++e)
                        if (target_x == S.x_rx[e]) { atomicMin(const_cast<int *>(&w_khit[e]), k); mine = true; }
                }
                if (SPLIT) __syncthreads(); else __syncwarp();  // every atomicMin on this row is in: the winners are known
                if (mine || want_legs) {
                    // main_rt.py:497-500 / main_compare.py:514-517
                    if (STATS) st[kStWinners] += 1;
                    const double2 pk = s_lens[k];
                    double t1 = S.leg1[(size_t)shooter * K + k];
                    double t2 = dist2d(pk.x, pk.y, x_q, z_q) / S.c2;
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
        // ---- one geometry done: coalesced row of the [G][E] table, or (sparse emission) only its hits ----
        if (SPLIT) { __syncthreads(); if (warp != 0) continue; } else __syncwarp();  // (SPLIT: warp 0 stores the CTA's row)
        if (!HG && T.hit_slots) {
            int base = 0;
            uint4 *slot = T.hit_slots + (size_t)w * T.hit_cap;
            for (int e0 = 0; e0 < E; e0 += 32) {
                const int e = e0 + lane;
                int kh = 0x7fffffff;  // kNoHit
                if (e < E) kh = w_khit[e];
                const unsigned hits = __ballot_sync(FULL, kh != 0x7fffffff);
                if (kh != 0x7fffffff) {
                    const int pos = base + __popc(hits & ((1u << lane) - 1u));
                    if (pos < T.hit_cap) {
                        const double t = w_tof[e];
                        slot[pos] = make_uint4((unsigned)e, (unsigned)kh, (unsigned)__double2loint(t), (unsigned)__double2hiint(t));
                    }
                }
                base += __popc(hits);
            }
            if (lane == 0) T.hit_count[w] = base;
        } else {
            for (int e = lane; e < E; e += 32) {
                int kh = w_khit[e];
                if (!HG) T.tof[(size_t)w * E + e] = w_tof[e];
                T.k_hit[(size_t)w * E + e] = kh == kNoHit ? -1 : kh;
            }
        }
        __syncwarp();
    }
    if (STATS && T.stats) {
#pragma unroll
        for (int i = 0; i < kNumStats; ++i) {
            unsigned v = __reduce_add_sync(FULL, st[i]);
            if (lane == 0) atomicAdd(T.stats + i, (unsigned long long)v);
        }
    }
}

