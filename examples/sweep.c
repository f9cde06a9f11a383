/*
 * sweep.c -- the reference's geometry sweep (main_rt.py:449-504) from plain C through the C ABI.
 *
 * What a host program (or a cgo / JNI / ctypes stub) does with libtofrt.so: build the inputs of
 * main_rt.py:449-479, call tofrt_setup once, trace the 10 x 21 (r_outer, pipe_offset) grid with ONE
 * tofrt_table_host call, and print the rows of database_2.csv (main_rt.py:487-504).
 *
 *   gcc -std=c99 -O2 -Iinclude examples/sweep.c -Lray_tracing_ultrasound_b200 -ltofrt -lm -o sweep
 *   LD_LIBRARY_PATH=ray_tracing_ultrasound_b200 ./sweep [out.csv]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "tofrt.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != TOFRT_OK) {                                                       \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, tofrt_last_error());       \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char **argv)
{
    /* main_rt.py:449-462 */
    tofrt_params p;
    p.c1 = 6400.0; p.c2 = 1483.0;
    p.l0 = 0.12156646438729327; p.h0 = 0.08843353561270673;
    p.d = p.l0 + p.h0;
    const double alpha_max = 50.62033040986099 * (acos(-1.0) / 180.0), pitch = 0.0006;
    enum { NEL = 64, E = NEL + 1, K = 181 * 5, NR = 10, NO = 21, G = NR * NO };

    /* main_rt.py:469-479: centred array with the virtual centre element inserted, launch angles, focus line */
    static double x_a[E], alpha[K], z_f[K], radius[G], offset[G], tof[G * E];
    static int k_hit[G * E];
    for (int i = 0, j = 0; i < NEL; ++i) {
        if (i == NEL / 2) x_a[j++] = 0.0;
        x_a[j++] = i * pitch - 0.5 * (NEL - 1) * pitch;
    }
    for (int k = 0; k < K; ++k) {
        alpha[k] = k == K - 1 ? alpha_max : -alpha_max + k * (2.0 * alpha_max / (K - 1));
        z_f[k] = p.d;
    }
    /* main_rt.py:464-467: radius-major grid */
    for (int r = 1, g = 0; r <= NR; ++r)
        for (int o = -10; o <= 10; ++o, ++g) { radius[g] = r * 1e-2; offset[g] = o * 1e-3; }

    tofrt_ctx *ctx = NULL;
    CHECK(tofrt_create(&ctx, 0));
    CHECK(tofrt_setup(ctx, &p, alpha, z_f, K, x_a[NEL / 2], p.d, x_a, E, 1e-5, 1e-6));   /* shooter = element 32 */
    CHECK(tofrt_table_host(ctx, radius, offset, G, tof, k_hit, NULL));

    FILE *out = argc > 1 ? fopen(argv[1], "wb") : NULL;
    int hits = 0;
    for (int g = 0; g < G; ++g)
        for (int e = 0; e < E; ++e) {
            const int hit = k_hit[g * E + e] >= 0;
            hits += hit;
            if (out)  /* main_rt.py:484-504: elem_idx, offset, radius, hitted, tof_total; CRLF rows */
                fprintf(out, "%d,%.17g,%.17g,%s,%.17g\r\n", e, offset[g], radius[g], hit ? "True" : "False",
                        hit ? tof[g * E + e] : 0.0);
        }
    if (out) fclose(out);
    printf("geometries %d receivers %d rows %d hits %d\n", G, E, G * E, hits);
    CHECK(tofrt_destroy(ctx));
    return 0;
}
