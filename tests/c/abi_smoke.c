/* abi_smoke.c — include/ludwig_b200.h used from plain C (C99): what a cgo / ccall / JNI binding sees.
 * Builds a 2 x 2 x 2-block box (periodic y / z, inlet / outlet x) with the reference's table conventions (1-based,
 * column-major, 0 = none), runs create -> level_create -> init_equilibrium -> step_batch -> flow_stats -> download ->
 * destroy and checks the obvious invariants.  Exit code 0 = ok.  Compiled by tests/c/Makefile, run by
 * tests/test_c_abi.py (the run needs a GPU; compiling and linking do not). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ludwig_b200.h"

#define NB 2
#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc__ = (call);                                                            \
        if (rc__ != LUDWIG_OK) {                                                      \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, ludwig_last_error(ctx));   \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(void) {
    enum { N = NB * NB * NB, CELLS = N * 512 };
    static int32_t block_pointer[N], neighbor_table[N * 27], map_x[N], map_y[N], map_z[N];
    static uint8_t obstacle[CELLS];
    static float sponge[CELLS], wall_dist[CELLS], rho[CELLS];
    ludwig_ctx* ctx = NULL;
    ludwig_level_desc d;
    ludwig_params p;
    double stats[6];
    int32_t idx = -1;
    int b = 0, bx, by, bz, dir, i;

    /* blocks sorted lexicographically with bx major (blocks.jl:89-103); block_pointer[bx,by,bz] column-major */
    for (bx = 1; bx <= NB; ++bx)
        for (by = 1; by <= NB; ++by)
            for (bz = 1; bz <= NB; ++bz) {
                map_x[b] = bx; map_y[b] = by; map_z[b] = bz;
                block_pointer[(bx - 1) + NB * ((by - 1) + NB * (bz - 1))] = ++b;
            }
    for (b = 0; b < N; ++b)
        for (dir = 0; dir < 27; ++dir) {
            int nx = map_x[b] + dir % 3 - 1, ny = map_y[b] + (dir / 3) % 3 - 1, nz = map_z[b] + dir / 9 - 1;
            ny = (ny - 1 + NB) % NB + 1; nz = (nz - 1 + NB) % NB + 1;                     /* periodic y, z */
            neighbor_table[b + N * dir] = (nx < 1 || nx > NB) ? 0 : block_pointer[(nx - 1) + NB * ((ny - 1) + NB * (nz - 1))];
        }
    for (i = 0; i < CELLS; ++i) { obstacle[i] = 0; sponge[i] = 0.0f; wall_dist[i] = 100.0f; }

    memset(&d, 0, sizeof d);
    d.level_id = 1; d.n_blocks = N; d.dim_x = d.dim_y = d.dim_z = NB; d.tau = 0.5006f; d.dx = 1.0;
    d.block_pointer = block_pointer; d.neighbor_table = neighbor_table; d.map_x = map_x; d.map_y = map_y; d.map_z = map_z;
    d.obstacle = obstacle; d.sponge = sponge; d.wall_dist = wall_dist; d.temporal_storage = 0;
    memset(&p, 0, sizeof p);
    p.c_wale = 0.5f; p.nu_sgs_bg = 0.0005f; p.inlet_turbulence = 0.01f; p.q_min_threshold = 0.001f; p.sponge_blend = 1;
    p.domain_nx = p.domain_ny = p.domain_nz = NB * 8; p.strict_fp = 1;

    if (ludwig_ctx_create(&ctx, 0) != LUDWIG_OK) { fprintf(stderr, "ludwig_ctx_create failed (no CUDA device?)\n"); return 2; }
    printf("backend %s\n", ludwig_backend_name());
    CHECK(ludwig_ctx_set_option(ctx, "verbose", "0"));
    if (ludwig_ctx_set_option(ctx, "no_such_option", "1") == LUDWIG_OK) { fprintf(stderr, "unknown option accepted\n"); return 1; }
    CHECK(ludwig_level_create(ctx, &d, &idx));
    CHECK(ludwig_init_equilibrium(ctx));
    CHECK(ludwig_step_batch(ctx, 1, 6, 0.03f, &p));
    CHECK(ludwig_sync(ctx));
    CHECK(ludwig_flow_stats(ctx, 0, stats));
    CHECK(ludwig_level_download(ctx, 0, LUDWIG_RHO, rho));
    printf("n_fluid %.0f rho_mean %.9f rho_min %.7f rho_max %.7f v_max %.6f launches %lld bytes %lld\n", stats[0], stats[1], stats[2], stats[3],
           stats[4], (long long)ludwig_launch_count(ctx), (long long)ludwig_device_bytes(ctx));
    if (idx != 0 || ludwig_num_levels(ctx) != 1 || stats[0] != (double)CELLS) return 1;
    if (!(fabs(stats[1] - 1.0) < 1e-3) || !(stats[4] > 1e-4 && stats[4] < 0.1)) return 1;   /* the inlet pushed a wave into the box */
    for (i = 0; i < CELLS; ++i) if (!(rho[i] >= stats[2] && rho[i] <= stats[3])) return 1;
    if (ludwig_level_step(ctx, 3, 1, 0, 0.0f, 0.03f, &p) == LUDWIG_OK) return 1;               /* bad level index is an error, not a crash */
    CHECK(ludwig_ctx_destroy(ctx));
    printf("C ABI smoke ok\n");
    return 0;
}
