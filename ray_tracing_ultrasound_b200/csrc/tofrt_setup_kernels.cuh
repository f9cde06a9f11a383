// tofrt_setup_kernels.cuh -- once-per-setup kernels: lens samples, first refraction of every shooter, chords.
// Part of the single translation unit tofrt_kernels.cu (included inside its anonymous namespace).
#pragma once

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


// per-ray tables of the table kernel, chunk by chunk: ray_t[shooter][c][plane][lane] holds ray min(32 c + lane, K-1)
// (the padding repeats the last ray, which the kernel treats as one more, idempotent, copy of it), so a warp's
// loads are coalesced lines at constant offsets from the chunk's base and need no bounds logic
__global__ void tofrt_ray_table_kernel(DevSetup S, double *ray_t)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_shot = (long long)S.C * 32;
    if (i >= per_shot * S.n_shooters) return;
    const int s = (int)(i / per_shot), r = (int)(i - (long long)s * per_shot), c = r >> 5, lane = r & 31;
    const int k = min(r, S.K - 1);
    const size_t o = (size_t)s * S.K + k;
    double *dst = ray_t + ((size_t)s * S.C + c) * (kRayPlanes * 32) + lane;
    const double2 pk = S.lens[k];
    dst[kPlaneXp * 32] = pk.x;
    dst[kPlaneZp * 32] = pk.y;
    dst[kPlaneCpq * 32] = S.cpq[o];
    dst[kPlaneSpq * 32] = S.spq[o];
    dst[kPlaneZf * 32] = S.z_f[k];
    dst[kPlaneApq * 32] = S.a_pq[o];
    dst[kPlaneBpq * 32] = S.b_pq[o];
    dst[kPlanePhi * 32] = S.phi_pq[o];
}
