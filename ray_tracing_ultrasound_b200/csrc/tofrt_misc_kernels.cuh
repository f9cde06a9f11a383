// tofrt_misc_kernels.cuh -- sparse hit list (count / compact) and the FP64 issue-rate micro-benchmark.
// Part of the single translation unit tofrt_kernels.cu (included inside its anonymous namespace).
#pragma once

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

// ---- sparse emission: per-row slots written by the table kernel -> ordered hit list ---------------------
// cnt[rows] must be followed by one zero entry; off = exclusive scan of cnt (rows + 1 entries, off[rows] = total).
// flags[0] is raised when a row holds more hits than its slot (the caller then takes the dense path).
__global__ void tofrt_gather_slots_kernel(const uint4 *slots, const int *cnt, const long long *off, long long rows,
                                          int cap, long long row0, long long base, long long capacity, int *hit_g,
                                          int *hit_e, int *hit_k, double *hit_tof, int *flags)
{
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int n = cnt[row];
    if (n > cap) { if (lane == 0) atomicOr(flags, 1); return; }
    const long long o = base + off[row];
    for (int i = lane; i < n; i += 32) {
        const uint4 v = slots[(size_t)row * cap + i];
        const long long j = o + i;
        if (j < capacity) {
            hit_g[j] = (int)(row0 + row); hit_e[j] = (int)v.x; hit_k[j] = (int)v.y;
            hit_tof[j] = __hiloint2double((int)v.w, (int)v.z);
        }
    }
}

__global__ void tofrt_widen_counts_kernel(const int *cnt, long long rows, long long *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= rows) out[i] = i < rows ? (long long)cnt[i] : 0ll;
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


// FP32 twin of the micro-benchmark (roofline denominator of the FP32 mode): MODE 0 FFMA, MODE 1 FMUL+FADD
template <int MODE>
__global__ void __launch_bounds__(256) tofrt_fp32_peak_kernel(float *out, int iters, float a, float b)
{
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { v[i] = fmaf(v[i], a, b); v[i] = fmaf(v[i], a, b); }
            else { v[i] = __fmul_rn(v[i], a); v[i] = __fadd_rn(v[i], b); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;  // keep the chains alive
}
