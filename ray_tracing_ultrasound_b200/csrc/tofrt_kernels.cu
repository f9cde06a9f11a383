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
// One translation unit; the device code lives in the .cuh files included below (inside this file's
// anonymous namespace): tofrt_common.cuh (setup struct, helpers, TMA wrappers), tofrt_setup_kernels.cuh,
// tofrt_table_kernel.cuh (the hot kernel), tofrt_table_f32.cuh, tofrt_fermat_kernels.cuh,
// tofrt_misc_kernels.cuh; per-pair/per-ray arithmetic shared with the oracle's twin in tofrt_math.h,
// tofrt_raymath.h, tofrt_fermat3d.h.  This file holds the context, the host logic and the C ABI.
//
// Compile with -fmad=false: the only fused operations are the explicit fma()
// calls inside tofrt_math.h (shared bit-for-bit with the oracle's twin build).
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <type_traits>
#include <string>
#include <vector>

#include "../../include/tofrt.h"
#include "tofrt_math.h"
#include "tofrt_fermat3d.h"
#include "tofrt_raymath.h"

#define HALF_PI 1.5707963267948966 /* np.pi / 2 */

namespace {

#include "tofrt_common.cuh"
#include "tofrt_setup_kernels.cuh"
#include "tofrt_table_kernel.cuh"
#include "tofrt_table_f32.cuh"
#include "tofrt_fermat_kernels.cuh"
#include "tofrt_misc_kernels.cuh"

// ---- host side ----------------------------------------------------------------------

thread_local std::string g_err;

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

typedef void (*table_kernel_t)(DevSetup, TableArgs);
// the instantiation for (first-hit state in global memory, guided search, flat surface, vector math); `stats` selects
// the instrumented twin (shared-memory first-hit state, lens surface only)
table_kernel_t pick_table_kernel(bool hg, bool guided, bool flat, bool vec, bool stats = false)
{
    if (stats && !flat && !hg) {
        if (guided) return vec ? tofrt_table_kernel<false, true, false, true, true> : tofrt_table_kernel<false, true, false, false, true>;
        return vec ? tofrt_table_kernel<false, false, false, true, true> : tofrt_table_kernel<false, false, false, false, true>;
    }
    if (flat) return vec ? tofrt_table_kernel<false, false, true, true, false> : tofrt_table_kernel<false, false, true, false, false>;
    if (hg) {
        if (guided) return vec ? tofrt_table_kernel<true, true, false, true, false> : tofrt_table_kernel<true, true, false, false, false>;
        return vec ? tofrt_table_kernel<true, false, false, true, false> : tofrt_table_kernel<true, false, false, false, false>;
    }
    if (guided) return vec ? tofrt_table_kernel<false, true, false, true, false> : tofrt_table_kernel<false, true, false, false, false>;
    return vec ? tofrt_table_kernel<false, false, false, true, false> : tofrt_table_kernel<false, false, false, false, false>;
}

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(TOFRT_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
    } while (0)

}  // namespace

// Largest dynamic shared-memory request of any kernel here (227 KB per CTA minus 1 KB for the kernels' few static
// variables).  The per-function opt-in attribute is per DEVICE, not per context, so it is always raised to this
// maximum: two live contexts with different K / E on one GPU can then never lower each other's limit.
constexpr size_t kMaxDynSmem = 227 * 1024 - 1024;
constexpr int kHitRing = 3;                  // chunk buffers in flight in tofrt_hits_host
constexpr int kHitSlotCap = 16;              // (e, k, tof) entries per row slot of the fused sparse emission
constexpr size_t kHitChunkRows = 32768;      // rows per pipelined chunk of tofrt_hits_host (8 MB of slots; measured best of 32 Ki .. 1 Mi)
constexpr unsigned kCounterSlots = 64, kCounterStride = 16;  // 128 B per slot: [0] work counter, [1..] statistics

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
    double *d_ray_t = nullptr;  // per-ray tables of the table kernel as coalesced planes
    int opt_small_g = 1;        // 1: few geometries run a CTA per geometry (TOFRT_OPT_SMALL_G)
    int opt_stats = 0;          // 1: table calls run the instrumented instantiation (tofrt_last_stats)
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
    // Work counters: every launch draws its own slot (dynamic work counter + statistics) from a ring, zeroed on the
    // caller's stream, so calls in flight on different streams never share one (up to kCounterSlots at a time).
    unsigned long long *d_counter = nullptr;
    std::atomic<unsigned> next_slot{0};
    unsigned last_table_slot = 0;
    unsigned long long *take_counter(cudaStream_t st, cudaError_t *err)
    {
        const unsigned slot = next_slot.fetch_add(1) % kCounterSlots;
        unsigned long long *p = d_counter + (size_t)slot * kCounterStride;
        *err = cudaMemsetAsync(p, 0, kCounterStride * sizeof(unsigned long long), st);
        return p;
    }
    // host-buffer pipeline (tofrt_table_host)
    // s_compute / s_compute2: one compute stream per pipeline slot, so that chunk i+1's kernel fills the SMs chunk i's
    // last CTAs leave idle instead of waiting for its tail; s_copy: device -> host; s_h2d: the chunks' geometries
    cudaStream_t s_compute = nullptr, s_compute2 = nullptr, s_copy = nullptr, s_h2d = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr};
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    void *ws[2] = {nullptr, nullptr};
    size_t ws_bytes = 0, in_cap = 0;
    double *d_in = nullptr;
    // tofrt_hits_host workspace (grow-only): dense intermediate, counts/offsets, scan scratch, hit list
    void *hw[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t hw_bytes[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    // fused sparse emission (tofrt_hits_host): geometries + a ring of three chunk buffers
    double *he_in = nullptr; size_t he_in_bytes = 0;
    uint4 *he_slots[kHitRing] = {};
    int *he_cnt[kHitRing] = {}, *he_flag[kHitRing] = {};
    long long *he_cnt64[kHitRing] = {}, *he_off[kHitRing] = {};
    void *he_tmp[kHitRing] = {}, *he_list[kHitRing] = {};
    size_t he_bytes[kHitRing][7] = {};
    long long *he_total = nullptr;   // pinned: per ring entry (hits in the chunk, overflow flag)
    cudaStream_t s_post = nullptr;   // high priority: the chunks' scan + gather
    cudaEvent_t ev_k[kHitRing] = {}, ev_post[kHitRing] = {}, ev_cnt[kHitRing] = {}, ev_out[kHitRing] = {};
    double *d_d1 = nullptr;   // Fermat solver: |A_e - P_j|/c1, built lazily per setup
    // 3-D two-point solve (tofrt_fermat3d_setup)
    bool f3_ready = false;
    bool f3_inplane = false;    // pairs in the plane y = 0 are solved by the 2-D solver's arithmetic (same aperture)
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
        cudaFree(d_ray_t); d_ray_t = nullptr;
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
    if (e == cudaSuccess) e = cudaMalloc(&c->d_counter, (size_t)kCounterSlots * kCounterStride * sizeof(unsigned long long));
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
    cudaFree(c->he_in);
    for (int i = 0; i < kHitRing; ++i) {
        cudaFree(c->he_slots[i]); cudaFree(c->he_cnt[i]); cudaFree(c->he_flag[i]); cudaFree(c->he_cnt64[i]);
        cudaFree(c->he_off[i]); cudaFree(c->he_tmp[i]); cudaFree(c->he_list[i]);
        for (cudaEvent_t e : {c->ev_k[i], c->ev_post[i], c->ev_cnt[i], c->ev_out[i]}) if (e) cudaEventDestroy(e);
    }
    if (c->s_post) cudaStreamDestroy(c->s_post);
    if (c->he_total) cudaFreeHost(c->he_total);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
        if (c->ev_free[i]) cudaEventDestroy(c->ev_free[i]);
    }
    if (c->s_compute) cudaStreamDestroy(c->s_compute);
    if (c->s_compute2) cudaStreamDestroy(c->s_compute2);
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    for (int i = 0; i < 2; ++i) if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
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
    // tofrt_vec_lens_consts, evaluated on the host (IEEE division: the same bits as on the device)
    S.VL.twod = S.twod; S.VL.K2 = S.K2; S.VL.fourAC = S.fourAC; S.VL.inv_twoA = 1.0 / S.twoA;
    S.VL.pt = -(S.phi1 * S.twod); S.VL.k12 = S.c1_over_c2; S.VL.h0 = S.h0; S.VL.flat = S.surface != 0;
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

    // Cell table for the candidate lookup of the table assembly: nb uniform cells over the finite receiver range
    // (+- tol_max, two guard cells at either end).  A landing point t is filed under c = int((t - bx0) * binv),
    // clamped; the two roundings of that expression move c by at most one cell, so cell c lists every receiver
    // within tol_max of [start(c-1), end(c+1)]: first | count << 16.  ~16 cells per receiver spacing leave most
    // cells empty -- a ray that lands between two receivers is dismissed by one shared-memory load.
    std::vector<int> bucket;
    S.nb = 0; S.bx0 = 0.0; S.binv = 0.0;
#ifndef TOFRT_NO_BUCKET  /* tuning knob: binary search instead */
    if (S.E_sorted >= 1 && S.E_sorted <= 0xffff) {
        // three guard cells at either end: cell 0 reaches up to start(3) + tol_max, just short of the first receiver
        const int nb = std::min(16 * S.E_sorted + 6, 8192);
        const double lo = xs.front() - S.tol_max, range = (xs.back() + S.tol_max) - lo;
        const double cell = range > 0.0 ? range / (nb - 6) : 1e-300, binv = 1.0 / cell, x0 = lo - 3 * cell;
        if (std::isfinite(range) && std::isfinite(binv) && cell > 0.0 && std::isfinite(x0) && nb > 6) {
            bucket.resize(nb);
            int first = 0, last = 0;
            bool ok = true;
            for (int cidx = 0; cidx < nb; ++cidx) {
                const double from = x0 + (cidx - 1) * cell - S.tol_max, to = x0 + (cidx + 2) * cell + S.tol_max;
                while (first < S.E_sorted && xs[first] < from) ++first;
                if (last < first) last = first;
                while (last < S.E_sorted && xs[last] <= to) ++last;
                if (last - first > 0xffff) ok = false;
                bucket[cidx] = first | ((last - first) << 16);
            }
            // the clamped end cells also receive every landing point beyond them: they must stay empty
            if ((bucket[0] >> 16) != 0 || ((unsigned)bucket[nb - 1] >> 16) != 0) ok = false;
            if (ok) { S.nb = nb; S.bx0 = x0; S.binv = binv; } else bucket.clear();
        }
    }
#endif
    const size_t smem_tables = (size_t)(S.Npad + K) * sizeof(double2) +
                               (size_t)S.E_sorted * (sizeof(double) + sizeof(int)) + (size_t)S.nb * sizeof(int) + 16;
    size_t smem = smem_tables + (size_t)kWarps * E * (sizeof(double) + sizeof(int));
    size_t smem_guided = smem_tables + (size_t)(kThreadsGuided / 32) * E * (sizeof(double) + sizeof(int));
    S.hits_in_global = 0;
    if (std::max(smem, smem_guided) > kMaxDynSmem) {  // many receivers: the per-warp first-hit state moves to the output rows
        S.hits_in_global = 1;
        smem = smem_guided = smem_tables;
    }
    if (smem > kMaxDynSmem)
        return fail(TOFRT_E_TOOLARGE, "tofrt_setup: K/E need " + std::to_string(smem) + " B of shared memory (> 227 KB)");

    CUDA_TRY(cudaMalloc(&c->d_alpha, sizeof(double) * K));
    CUDA_TRY(cudaMalloc(&c->d_zf, sizeof(double) * K));
    CUDA_TRY(cudaMalloc(&c->d_a, sizeof(double) * K * NS));
    CUDA_TRY(cudaMalloc(&c->d_b, sizeof(double) * K * NS));
    CUDA_TRY(cudaMalloc(&c->d_phi, sizeof(double) * K * NS));
    CUDA_TRY(cudaMalloc(&c->d_leg1, sizeof(double) * K * NS));
    CUDA_TRY(cudaMalloc(&c->d_cs, sizeof(double) * 2 * K * NS));
    S.C = (K + 31) / 32;
    CUDA_TRY(cudaMalloc(&c->d_ray_t, sizeof(double) * (size_t)kRayPlanes * S.C * 32 * NS));
    S.ray_t = c->d_ray_t;
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
    tofrt_ray_table_kernel<<<(unsigned)(((size_t)S.C * 32 * NS + 127) / 128), 128>>>(S, c->d_ray_t);
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
        bool incr = true, dec = true, inc = true, vertical = false;
        double X = 0, Z = 0;
        for (int j = 0; j < K; ++j) {
            X = std::max(X, std::fabs(hl[j].x));
            Z = std::max(Z, std::fabs(hl[j].y));
            if (j + 1 < K && !(hl[j + 1].x > hl[j].x)) incr = false;
            if (j + 1 < K) {  // np.isclose(x1, x2) of main_rt.py:110, with b = x2
                const double x1 = hl[j].x, x2 = hl[j + 1].x;
                if (((std::fabs(x1 - x2) <= 1e-8 + 1e-5 * std::fabs(x2)) && std::isfinite(x2)) || x1 == x2) vertical = true;
                if (!(hl[j].y == hl[j].y) || !(hl[j + 1].y == hl[j + 1].y)) vertical = true;  // NaN heights: generic path
            }
        }
        S.chords_plain = (S.tables_ok && incr && !vertical) ? 1 : 0;
        for (int j = 0; j + 2 < K; ++j) {
            double a = hs[j].x, b2 = hs[j + 1].x;
            double guard = 16 * 2.220446049250313e-16 * std::max(std::fabs(a), std::fabs(b2));
            if (!(a - b2 > guard)) dec = false;
            if (!(b2 - a > guard)) inc = false;
        }
        S.X = X; S.Z = Z;
        S.P0 = hl[0]; S.PN = hl[K - 1];
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
        const size_t n_ray_t = (size_t)kRayPlanes * S.C * 32 * NS;
        const size_t nf = (size_t)4 * K + n_ray_t + (size_t)K * NS + E + Es;  // lens(2K) seg(2K) ray_t leg1(S*K) x_rx xs
        CUDA_TRY(cudaMalloc(&c->d_f32, sizeof(float) * nf));
        float *f = c->d_f32;
        auto conv = [&](const double *src, size_t n) {
            float *dst = f;
            if (n > 0) tofrt_to_float_kernel<<<(unsigned)((n + 255) / 256), 256>>>((int)n, src, dst);
            f += n;
            return dst;
        };
        DevSetupF F{};
        F.lens = (const float2 *)conv((const double *)c->d_lens, 2 * (size_t)K);
        F.seg = (const float2 *)conv((const double *)c->d_seg, 2 * (size_t)K - 2); f += 2;
        F.ray_t = conv(c->d_ray_t, n_ray_t);
        F.leg1 = conv(c->d_leg1, (size_t)K * NS);
        F.x_rx = conv(c->d_xrx, E); F.xs = conv(c->d_xs, S.E_sorted);
        CUDA_TRY(cudaGetLastError());
        F.perm = c->d_perm;
        F.bucket = c->d_bucket; F.nb = S.nb; F.bx0 = (float)S.bx0; F.binv = (float)S.binv;
        F.c1 = (float)S.c1; F.c2 = (float)S.c2; F.fourAC = (float)S.fourAC; F.inv_twoA = (float)S.VL.inv_twoA;
        F.twod = (float)S.twod; F.K2 = (float)S.K2; F.pt = (float)S.VL.pt; F.c1_over_c2 = (float)S.c1_over_c2;
        F.rtol = (float)S.rtol; F.atol = (float)S.atol;
        F.tol_max = (float)(S.tol_max * 1.001) + 1e-30f;
        F.rx_min = (float)S.rx_min; F.rx_max = (float)S.rx_max; F.X = (float)S.X * 1.0001f; F.Z = (float)S.Z * 1.0001f;
        F.K = K; F.C = S.C; F.E = E; F.E_sorted = S.E_sorted; F.has_nonfinite_rx = S.has_nonfinite_rx;
        F.shape = S.shape; F.log2n = S.log2n; F.n_shooters = NS; F.surface = S.surface; F.h0 = (float)S.h0;
        F.P0 = make_float2((float)S.P0.x, (float)S.P0.y); F.PN = make_float2((float)S.PN.x, (float)S.PN.y);
        // the cell table is only safe in float while a cell is wide against the rounding of a landing point
        if (!(std::fabs(S.bx0) < 1e30 && S.binv < 1e30 && 1.0 / std::max(S.binv, 1e-300) > 64 * 1.2e-7 * std::max(std::fabs(S.rx_min), std::fabs(S.rx_max))))
            F.nb = 0;
        c->SF = F;
        c->smem_f32 = (size_t)2 * K * sizeof(float2) + (size_t)S.E_sorted * (sizeof(float) + sizeof(int)) +
                      (size_t)kWarpsF32 * E * (sizeof(float) + sizeof(int)) + (size_t)F.nb * sizeof(int) + 16;
        if (c->smem_f32 > kMaxDynSmem) c->smem_f32 = 0;  // FP32 mode keeps its first-hit state in shared memory only
        else
            CUDA_TRY(cudaFuncSetAttribute(tofrt_table_kernel_f32, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kMaxDynSmem));
        CUDA_TRY(cudaDeviceSynchronize());
    }

    if (S.surface != 0 && S.hits_in_global)
        return fail(TOFRT_E_TOOLARGE, "tofrt_setup: the flat-interface extension supports the shared-memory receiver budget only");
    const bool flat = S.surface != 0, hg = S.hits_in_global != 0;
    for (int vec = 0; vec < 2; ++vec)
        for (int gd = 0; gd < (flat ? 1 : 2); ++gd)
            for (int stt = 0; stt < 2; ++stt)
                CUDA_TRY(cudaFuncSetAttribute(pick_table_kernel(hg, gd != 0, flat, vec != 0, stt != 0),
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
    CUDA_TRY(cudaFuncSetAttribute(tofrt_table_kernel<false, true, false, true, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
    CUDA_TRY(cudaFuncSetAttribute(tofrt_table_kernel<false, true, false, false, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
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
    if (option == TOFRT_OPT_SMALL_G) {
        if (value != 0 && value != 1) return fail(TOFRT_E_BADARG, "tofrt_set_option: bad small-G switch");
        c->opt_small_g = value;
        return TOFRT_OK;
    }
    if (option == TOFRT_OPT_STATS) {
        if (value != 0 && value != 1) return fail(TOFRT_E_BADARG, "tofrt_set_option: bad statistics switch");
        c->opt_stats = value;
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

int tofrt_last_stats(tofrt_ctx *c, unsigned long long *stats, int n)
{
    if (!c || !stats || n < 0) return fail(TOFRT_E_BADARG, "tofrt_last_stats: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    unsigned long long h[kNumStats];
    CUDA_TRY(cudaMemcpy(h, c->d_counter + (size_t)c->last_table_slot * kCounterStride + 1, sizeof h, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) stats[i] = i < kNumStats ? h[i] : 0ull;
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

struct HitEmit {  // sparse emission of table_dev_impl: per-row slots instead of the dense rows
    uint4 *slots = nullptr;
    int *count = nullptr;
    int cap = 0;
};

static int table_dev_impl(tofrt_ctx *c, const double *d_radius, const double *d_offset, long long G, double *d_tof,
                          int *d_k_hit, const tofrt_ray_records *d_rec, void *stream, int s_begin, int s_count,
                          const double *inject_m = nullptr, const double *inject_b = nullptr,
                          const HitEmit *emit = nullptr)
{
    if (!c || G < 0) return fail(TOFRT_E_BADARG, "tofrt_table_dev: bad argument");
    if (!c->ready) return fail(TOFRT_E_NOSETUP, "tofrt_table_dev: call tofrt_setup first");
    if (G == 0) return TOFRT_OK;
    if (!d_radius || !d_offset || ((!d_tof || !d_k_hit) && !emit)) return fail(TOFRT_E_BADARG, "tofrt_table_dev: NULL buffer");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    TableArgs T{};
    T.radius = d_radius; T.offset = d_offset; T.G = G; T.tof = d_tof; T.k_hit = d_k_hit;
    if (d_rec) T.rec = *d_rec;
    T.any_rec = T.rec.pipe_x || T.rec.pipe_z || T.rec.lens2_x || T.rec.lens2_z || T.rec.target_x || T.rec.idx;
    T.s_begin = s_begin; T.s_count = s_count;
    T.inject_m = inject_m; T.inject_b = inject_b;
    if (emit) { T.hit_slots = emit->slots; T.hit_count = emit->count; T.hit_cap = emit->cap; }
    cudaError_t ce;
    T.counter = c->take_counter(st, &ce);
    CUDA_TRY(ce);
    T.stats = T.counter + 1;
    c->last_table_slot = (unsigned)((T.counter - c->d_counter) / kCounterStride);
    long long warps_needed = G * s_count;
    const bool guided = c->S.guided && c->S.shape != 0 && c->S.surface == 0;
    const int threads = guided ? kThreadsGuided : kThreads, wpc = threads / 32;
    const size_t smem = guided ? c->smem_guided : c->smem;
    int grid = (int)std::min<long long>((long long)c->sm_count * c->ctas_per_sm, (warps_needed + wpc - 1) / wpc);
    // Small-G latency mode: with fewer (shooter, geometry) items than the GPU has resident warps, a warp per item
    // leaves most of the machine idle while each warp walks its 29 chunks one after the other; a CTA per item with a
    // chunk per warp finishes an item in the time of ONE chunk.  (Test hook and instrumented runs keep the main kernel.)
    const bool split = guided && !c->S.hits_in_global && !c->opt_stats && !inject_m && c->opt_small_g &&
                       c->S.C <= kThreadsGuided / 32 && warps_needed < (long long)c->sm_count * c->ctas_per_sm * wpc / 2 &&
                       smem + kSplitQueueBytes <= kMaxDynSmem;
    if (split) {
        grid = (int)std::min<long long>((long long)c->sm_count * c->ctas_per_sm, warps_needed);
        const size_t smem_split = smem + kSplitQueueBytes;
        if (c->S.vector_math) tofrt_table_kernel<false, true, false, true, false, true><<<grid, threads, smem_split, st>>>(c->S, T);
        else tofrt_table_kernel<false, true, false, false, false, true><<<grid, threads, smem_split, st>>>(c->S, T);
        CUDA_TRY(cudaGetLastError());
        return TOFRT_OK;
    }
    // (flat interface, extension: a plane has no strict curvature, so `guided` is false there: literal scan only)
    pick_table_kernel(c->S.hits_in_global != 0, guided, c->S.surface != 0, c->S.vector_math != 0, c->opt_stats != 0)
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
    cudaError_t ce;
    T.counter = c->take_counter(st, &ce);
    CUDA_TRY(ce);
    int grid = (int)std::min<long long>((long long)c->sm_count, (G * c->n_shooters + kWarpsF32 - 1) / kWarpsF32);
    tofrt_table_kernel_f32<<<grid, kThreadsF32, c->smem_f32, st>>>(c->SF, T);
    CUDA_TRY(cudaGetLastError());
    return TOFRT_OK;
}

static int ensure_streams(tofrt_ctx *c)
{
    if (!c->s_compute) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
        }
    }
    if (!c->s_compute2) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_compute2, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) CUDA_TRY(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
    }
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
    { int rcs = ensure_streams(c); if (rcs != TOFRT_OK) return rcs; }
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
        cudaStream_t cs = slot ? c->s_compute2 : c->s_compute;
        if (sh == 0) {  // this chunk's geometries (every shooter reuses them)
            CUDA_TRY(cudaMemcpyAsync(c->d_in + done, radius + done, sizeof(double) * n, cudaMemcpyHostToDevice, c->s_h2d));
            CUDA_TRY(cudaMemcpyAsync(c->d_in + g + done, offset + done, sizeof(double) * n, cudaMemcpyHostToDevice, c->s_h2d));
            CUDA_TRY(cudaEventRecord(c->ev_in[slot], c->s_h2d));
            CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_in[slot], 0));
        } else if (done == 0) {  // next shooter: both compute streams must have seen every geometry copy
            CUDA_TRY(cudaEventRecord(c->ev_in[0], c->s_h2d));
            CUDA_TRY(cudaStreamWaitEvent(c->s_compute, c->ev_in[0], 0));
            CUDA_TRY(cudaStreamWaitEvent(c->s_compute2, c->ev_in[0], 0));
        }
        if (i >= 2) CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_free[slot], 0));
        rc = table_dev_impl(c, c->d_in + done, c->d_in + g + done, (long long)n, d_tof, d_kh, rec ? &d : nullptr,
                            (void *)cs, sh, 1);
        if (rc != TOFRT_OK) break;
        CUDA_TRY(cudaEventRecord(c->ev_done[slot], cs));
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
    cudaError_t e3 = cudaStreamSynchronize(c->s_compute2), e4 = cudaStreamSynchronize(c->s_h2d);
    if (e1 == cudaSuccess) e1 = e3 != cudaSuccess ? e3 : e4;
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

// Dense route to the hit list (many receivers per row, or a row that overflowed its slot): full table on the device,
// count per row -> exclusive scan -> ordered compaction.
static int hits_host_dense(tofrt_ctx *c, const double *radius, const double *offset, long long G, long long capacity,
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

// Hit list through the fused emission: the table kernel writes each row's hits into a small per-row slot (no dense
// table is stored or re-read); per chunk of rows a scan of the counts and an ordered gather ("post") build the list,
// which travels to the host while later chunks are traced.
//
// Streams.  The table kernel is persistent and fills every SM (1024 threads x 64 registers), so nothing else runs
// beside it: chunk i's post kernels only get SMs when the NEXT table kernel runs out of work.  If they sat in front of
// chunk i+2's kernel on the same stream, the GPU would drain at every chunk boundary (measured: +0.9 ms per 1 Mi
// rows).  So the table kernels alternate between two streams and never wait for a post; the posts go to a third,
// HIGH-PRIORITY stream (their blocks win the SMs a finishing table kernel frees, ahead of the next kernel's CTAs), and
// everything they touch is a ring of three buffers: chunk i+3 reuses what chunk i's copy to the host has released.
int tofrt_hits_host(tofrt_ctx *c, const double *radius, const double *offset, long long G, long long capacity,
                    int *hit_g, int *hit_e, int *hit_k, double *hit_tof, long long *n_hits)
{
    if (!c || G < 0 || capacity < 0 || !n_hits) return fail(TOFRT_E_BADARG, "tofrt_hits_host: bad argument");
    if (!c->ready) return fail(TOFRT_E_NOSETUP, "tofrt_hits_host: call tofrt_setup first");
    *n_hits = 0;
    if (G == 0) return TOFRT_OK;
    if (!radius || !offset || (capacity > 0 && (!hit_g || !hit_e || !hit_k || !hit_tof)))
        return fail(TOFRT_E_BADARG, "tofrt_hits_host: NULL buffer");
    if (c->S.hits_in_global)  // the first-hit state lives in the dense rows: nothing to fuse
        return hits_host_dense(c, radius, offset, G, capacity, hit_g, hit_e, hit_k, hit_tof, n_hits);
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t g = (size_t)G, NSH = (size_t)c->n_shooters;
    const int cap = kHitSlotCap;
    size_t chunk_rows = kHitChunkRows;
    if (const char *env = getenv("TOFRT_HIT_CHUNK_ROWS")) {  // tuning knob
        const long long v = atoll(env);
        if (v >= 1024) chunk_rows = (size_t)v;
    }
    const size_t chunk = std::min<size_t>(g, chunk_rows);
    const size_t n_chunks = (g + chunk - 1) / chunk;
    { int rcs = ensure_streams(c); if (rcs != TOFRT_OK) return rcs; }
    if (!c->s_post) {
        int lo_p = 0, hi_p = 0;  // (numerically lowest = highest priority)
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
        CUDA_TRY(cudaStreamCreateWithPriority(&c->s_post, cudaStreamNonBlocking, hi_p));
        for (int i = 0; i < kHitRing; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_post[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_cnt[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
        }
        CUDA_TRY(cudaHostAlloc((void **)&c->he_total, 2 * kHitRing * sizeof(long long), cudaHostAllocDefault));
    }
    // grow-only workspace: geometries, and per ring entry the row slots, counts, offsets, scan scratch, list, flag
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (long long *)nullptr, (long long *)nullptr, (int)(chunk + 1)));
    auto grow = [&](void **q, size_t *have, size_t bytes) {
        bytes = std::max<size_t>(bytes, 16);
        if (*have >= bytes) return cudaSuccess;
        cudaFree(*q); *q = nullptr; *have = 0;
        cudaError_t e = cudaMalloc(q, bytes);
        if (e == cudaSuccess) *have = bytes;
        return e;
    };
    const size_t list_cap = chunk * (size_t)cap;  // most hits one chunk can deliver through the slots
    CUDA_TRY(grow((void **)&c->he_in, &c->he_in_bytes, sizeof(double) * 2 * g));
    for (int s = 0; s < kHitRing; ++s) {
        CUDA_TRY(grow((void **)&c->he_slots[s], &c->he_bytes[s][0], sizeof(uint4) * list_cap));
        CUDA_TRY(grow((void **)&c->he_cnt[s], &c->he_bytes[s][1], sizeof(int) * chunk));
        CUDA_TRY(grow((void **)&c->he_cnt64[s], &c->he_bytes[s][2], sizeof(long long) * (chunk + 1)));
        CUDA_TRY(grow((void **)&c->he_off[s], &c->he_bytes[s][3], sizeof(long long) * (chunk + 1)));
        CUDA_TRY(grow((void **)&c->he_tmp[s], &c->he_bytes[s][4], tmp_bytes));
        CUDA_TRY(grow((void **)&c->he_list[s], &c->he_bytes[s][5], list_cap * 20));
        CUDA_TRY(grow((void **)&c->he_flag[s], &c->he_bytes[s][6], sizeof(int)));
    }

    long long total = 0;
    bool overflow = false;
    int rc = TOFRT_OK;
    const size_t n_items = n_chunks * NSH;
    // drain(i): the host learns chunk i's hit count, then queues its copies at the running offset
    auto drain = [&](size_t i) -> int {
        const int b = (int)(i % kHitRing);
        CUDA_TRY(cudaEventSynchronize(c->ev_cnt[b]));
        const long long got = c->he_total[2 * b];
        if (c->he_total[2 * b + 1]) overflow = true;
        const long long room = std::max<long long>(0, capacity - total);
        const size_t n = (size_t)std::min<long long>(got, room);
        if (n && !overflow) {
            const char *L = (const char *)c->he_list[b];
            CUDA_TRY(cudaStreamWaitEvent(c->s_copy, c->ev_post[b], 0));
            CUDA_TRY(cudaMemcpyAsync(hit_g + total, L, sizeof(int) * n, cudaMemcpyDeviceToHost, c->s_copy));
            CUDA_TRY(cudaMemcpyAsync(hit_e + total, L + 4 * list_cap, sizeof(int) * n, cudaMemcpyDeviceToHost, c->s_copy));
            CUDA_TRY(cudaMemcpyAsync(hit_k + total, L + 8 * list_cap, sizeof(int) * n, cudaMemcpyDeviceToHost, c->s_copy));
            CUDA_TRY(cudaMemcpyAsync(hit_tof + total, L + 12 * list_cap, sizeof(double) * n, cudaMemcpyDeviceToHost, c->s_copy));
        }
        CUDA_TRY(cudaEventRecord(c->ev_out[b], c->s_copy));
        total += got;
        return TOFRT_OK;
    };
    size_t drained = 0;
    for (size_t i = 0; i < n_items && rc == TOFRT_OK && !overflow; ++i) {
        const int b = (int)(i % kHitRing);
        const int sh = (int)(i / n_chunks);
        const size_t done = (i % n_chunks) * chunk, n = std::min(chunk, g - done);
        cudaStream_t cs = (i & 1) ? c->s_compute2 : c->s_compute;
        if (sh == 0) {  // this chunk's geometries (every shooter reuses them)
            const int slot = (int)(i & 1);
            CUDA_TRY(cudaMemcpyAsync(c->he_in + done, radius + done, sizeof(double) * n, cudaMemcpyHostToDevice, c->s_h2d));
            CUDA_TRY(cudaMemcpyAsync(c->he_in + g + done, offset + done, sizeof(double) * n, cudaMemcpyHostToDevice, c->s_h2d));
            CUDA_TRY(cudaEventRecord(c->ev_in[slot], c->s_h2d));
            CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_in[slot], 0));
        } else if (done == 0) {  // next shooter: both compute streams must have seen every geometry copy
            CUDA_TRY(cudaEventRecord(c->ev_in[0], c->s_h2d));
            CUDA_TRY(cudaStreamWaitEvent(c->s_compute, c->ev_in[0], 0));
            CUDA_TRY(cudaStreamWaitEvent(c->s_compute2, c->ev_in[0], 0));
        }
        if (i >= (size_t)kHitRing) CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_out[b], 0));  // ring entry released by chunk i-3's copy
        HitEmit em;
        em.slots = c->he_slots[b]; em.count = c->he_cnt[b]; em.cap = cap;
        rc = table_dev_impl(c, c->he_in + done, c->he_in + g + done, (long long)n, nullptr, nullptr, nullptr,
                            (void *)cs, sh, 1, nullptr, nullptr, &em);
        if (rc != TOFRT_OK) break;
        CUDA_TRY(cudaEventRecord(c->ev_k[b], cs));
        // post (high priority): counts -> offsets -> ordered list of this chunk
        cudaStream_t ps = c->s_post;
        char *L = (char *)c->he_list[b];
        const int blocks = (int)((n * 32 + 255) / 256);
        CUDA_TRY(cudaStreamWaitEvent(ps, c->ev_k[b], 0));
        CUDA_TRY(cudaMemsetAsync(c->he_flag[b], 0, sizeof(int), ps));
        tofrt_widen_counts_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, ps>>>(c->he_cnt[b], (long long)n, c->he_cnt64[b]);
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(c->he_tmp[b], tmp_bytes, c->he_cnt64[b], c->he_off[b], (int)n + 1, ps));
        tofrt_gather_slots_kernel<<<blocks, 256, 0, ps>>>(
            c->he_slots[b], c->he_cnt[b], c->he_off[b], (long long)n, cap, (long long)((size_t)sh * g + done), 0,
            (long long)list_cap, (int *)L, (int *)(L + 4 * list_cap), (int *)(L + 8 * list_cap),
            (double *)(L + 12 * list_cap), c->he_flag[b]);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(c->ev_post[b], ps));
        CUDA_TRY(cudaMemcpyAsync(&c->he_total[2 * b], c->he_off[b] + n, sizeof(long long), cudaMemcpyDeviceToHost, ps));
        CUDA_TRY(cudaMemcpyAsync(&c->he_total[2 * b + 1], c->he_flag[b], sizeof(int), cudaMemcpyDeviceToHost, ps));
        CUDA_TRY(cudaEventRecord(c->ev_cnt[b], ps));
        // chunk i's post runs when chunk i+1's kernel ends: the host stays two chunks ahead of the counts it waits for
        if (i >= 2) { rc = drain(i - 2); drained = i - 1; }
    }
    for (; rc == TOFRT_OK && !overflow && drained < n_items; ++drained) rc = drain(drained);
    cudaError_t e1 = cudaStreamSynchronize(c->s_compute), e2 = cudaStreamSynchronize(c->s_copy);
    cudaError_t e3 = cudaStreamSynchronize(c->s_compute2), e4 = cudaStreamSynchronize(c->s_h2d);
    cudaError_t e5 = cudaStreamSynchronize(c->s_post);
    if (e1 == cudaSuccess) e1 = e3 != cudaSuccess ? e3 : (e4 != cudaSuccess ? e4 : e5);
    if (rc != TOFRT_OK) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess)
        return fail(TOFRT_E_CUDA, std::string("tofrt_hits_host: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    if (overflow)  // some row holds more hits than a slot: the dense route handles any number
        return hits_host_dense(c, radius, offset, G, capacity, hit_g, hit_e, hit_k, hit_tof, n_hits);
    *n_hits = total;
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
    if (smem > kMaxDynSmem) return fail(TOFRT_E_TOOLARGE, "tofrt_fermat_dev: too many lens samples");
    if (!c->d_d1) {
        CUDA_TRY(cudaMalloc(&c->d_d1, sizeof(double) * (size_t)K * E));
        dim3 grid((K + 127) / 128, E);
        tofrt_fermat_d1_kernel<<<grid, 128, 0, st>>>(c->S, c->params_d, c->d_d1);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaFuncSetAttribute(tofrt_fermat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
    }
    FermatArgs A{};
    A.fx = d_fx; A.fz = d_fz; A.F = F; A.d1 = c->d_d1; A.alpha = c->d_alpha; A.tof = d_tof; A.kmin = d_kmin;
    A.alpha_out = d_alpha; A.z_el = c->params_d;
    cudaError_t ce;
    A.counter = c->take_counter(st, &ce);
    CUDA_TRY(ce);
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
    {   // in-plane pairs go to the 2-D solve only when both solvers search the same aperture
        double a_ends[2];
        CUDA_TRY(cudaMemcpy(&a_ends[0], c->d_alpha, sizeof(double), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(&a_ends[1], c->d_alpha + (c->K - 1), sizeof(double), cudaMemcpyDeviceToHost));
        c->f3_inplane = a_ends[0] == S.lo && a_ends[1] == S.hi && c->n_shooters >= 1;
    }
    c->f3_cell = (S.hi - S.lo) / (n_side - 1);
    F3dArgs A{};
    A.ang = c->f3_geo; A.ex = c->f3_geo + n_side; A.ey = c->f3_geo + n_side + E3;
    A.E = E3; A.n_side = n_side; A.z_el = c->params_d;
    dim3 grid((unsigned)((ns + 127) / 128), (unsigned)E3);
    tofrt_f3d_prepare_kernel<<<grid, 128>>>(S, A, c->f3_smp);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFuncSetAttribute(tofrt_fermat3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kMaxDynSmem));
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
    A.tof = d_tof; A.kmin = d_kmin; A.ab = d_ab;
    A.alpha2d = c->f3_inplane ? c->d_alpha : nullptr;
    cudaError_t ce;
    A.counter = c->take_counter(st, &ce);
    CUDA_TRY(ce);
    int occ = 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tofrt_fermat3d_kernel, kF3dThreads, smem));
    const long long items = ((F + kF3dThreads - 1) / kF3dThreads) * E3;
    int grid = (int)std::min<long long>((long long)c->sm_count * std::max(1, occ), items);
    tofrt_fermat3d_kernel<<<grid, kF3dThreads, smem, st>>>(c->f3_S, A, c->S);
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
    { int rcs = ensure_streams(c); if (rcs != TOFRT_OK) return rcs; }
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

int tofrt_measure_fp32_peak(tofrt_ctx *c, double *ffma_per_s, double *fmul_fadd_per_s)
{
    if (!c || !ffma_per_s || !fmul_fadd_per_s) return fail(TOFRT_E_BADARG, "tofrt_measure_fp32_peak: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    float *d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, sizeof(float)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int iters = 1 << 15, blocks = c->sm_count * 8, threads = 256;
    const double inst = (double)blocks * threads * iters * 16.0;
    double best[2] = {0, 0};
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) tofrt_fp32_peak_kernel<0><<<blocks, threads>>>(d_out, iters, 0.999999f, 1e-9f);
            else tofrt_fp32_peak_kernel<1><<<blocks, threads>>>(d_out, iters, 0.999999f, 1e-9f);
            cudaEventRecord(e1);
            cudaError_t e = cudaEventSynchronize(e1);
            if (e != cudaSuccess) { cudaFree(d_out); return fail(TOFRT_E_CUDA, cudaGetErrorString(e)); }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0) best[mode] = std::max(best[mode], inst / (ms * 1e-3));
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    *ffma_per_s = best[0];
    *fmul_fadd_per_s = best[1];
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
