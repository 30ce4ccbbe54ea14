// k1_ghost.cuh — the refinement-interface halo pre-pass (2:1 coarse->fine interpolation with temporal blend and f_neq
// rescaling, physics_interpolation.jl:16-138) as its own kernels.  Textually included inside a per-TU namespace AFTER
// k1_boundary.cuh by k1_fast.cu (FMA contraction on) and k1_strict.cu (-fmad=false: the reference's operation order, bit-exact
// against the in-kernel interpolation of the CPU oracle).
// ---------------------------------------------------------------------------------------------------------
// Interface halo pre-pass.  In the reference every missing-neighbour population of a fine block is interpolated
// inside the stream-collide thread that needs it (interpolate_with_rescaling, physics_interpolation.jl:16-138):
// 8 parent corners x (f_k, rho, u) x (new, old) scattered loads per population, serialised in a handful of
// divergent lanes.  Here the missing in-domain neighbour blocks of a level exist as GHOST blocks, filled before K1,
// which then treats them as ordinary neighbours (interface blocks run the plain kernel).
// The 2x2x2 fine cells inside one parent cell share their 8 parent corners (p0 = (g-1) >> 1) and differ only in the
// weights (0.25 / 0.75 per axis), so one thread handles such a GROUP: the 8-corner rho/u/f_k loads are done once and
// reused for up to 8 ghost cells (4 on a face layer) — ~4x fewer scattered loads than one thread per cell.
// Same arithmetic per population as the reference (its per-direction rho/u interpolation is direction-independent).
__global__ void __launch_bounds__(128) ghost_interp_kernel(const GhostArgs a) {
    // (dealing a group's populations to 4 lanes was measured: 2.7x slower — the kernel is bound by DRAM sectors of the
    // scattered parent values, not by its load chains)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int gq = a.gcell[i];               // ghost block * 64 + group (qz*16 + qy*4 + qx)
    const uint32_t word = a.gmask[i];        // bits 0..26: populations needed by some cell of the group; 27..31 unused
    const uint32_t cells = a.gcells8[i];     // bit (dz*4 + dy*2 + dx): that cell of the group is pulled from
    const int g = gq >> 6, q = gq & 63;
    const int x0 = (q & 3) * 2, y0 = ((q >> 2) & 3) * 2, z0 = (q >> 4) * 2;
    const int4 gb = *reinterpret_cast<const int4*>(a.gcoord + (size_t)g * 4);
    // 1-based fine coordinates of the group's first cell (odd) and its shared parent corner p0
    const int fgx = gb.x * BS + x0 + 1, fgy = gb.y * BS + y0 + 1, fgz = gb.z * BS + z0 + 1;
    const int p0x = (fgx - 1) >> 1, p0y = (fgy - 1) >> 1, p0z = (fgz - 1) >> 1;   // floor((g - 0.5) * 0.5) for both cells of a pair
    const int cx0 = max(1, p0x), cy0 = max(1, p0y), cz0 = max(1, p0z);             // :44-46 (clamped AFTER p1 = p0 + 1)
    const int cx1 = p0x + 1, cy1 = p0y + 1, cz1 = p0z + 1;

    int pb[8], loc[8];   // corner order 000,100,010,110,001,101,011,111 (x fastest)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int pgx = (c & 1) ? cx1 : cx0, pgy = (c & 2) ? cy1 : cy0, pgz = (c & 4) ? cz1 : cz0;
        const int bx = (pgx - 1) >> 3, by = (pgy - 1) >> 3, bz = (pgz - 1) >> 3;
        pb[c] = -1;   // (owner rank << 24) | owner-local block index
        if (bx >= 0 && bx < a.pdimx && by >= 0 && by < a.pdimy && bz >= 0 && bz < a.pdimz) pb[c] = a.pptr[bx + a.pdimx * (by + a.pdimy * bz)];
        loc[c] = ((pgx - 1) & 7) + 8 * ((pgy - 1) & 7) + 64 * ((pgz - 1) & 7);
    }
    // temporal blend old (1 - tw) + new tw (:60-80).  At tw = 0 (the first of a parent step's two child sub-steps) that is
    // the OLD state exactly, so the new state is not loaded at all: half the scattered parent loads of that sub-step.
    const bool only_old = a.use_temporal == 1 && a.tw == 0.0f;
    const bool blend = a.use_temporal == 1 && a.tw < 0.99f && !only_old;
    const float tw = a.tw;
    // rho, u at the corners (invalid corner: (1,0,0,0); corners 1..7 then fall back to corner 0, valid or not)
    float cr[8], cux[8], cuy[8], cuz[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (pb[c] >= 0) {
            const int pr = pb[c] >> PTR_RANK_SHIFT, pl = pb[c] & PTR_LOCAL_MASK;
            const size_t ri = (size_t)pl * BS3 + loc[c], vi = (size_t)pl * 3 * BS3 + loc[c];
            const float* __restrict__ vn = only_old ? a.pvel_old.p[pr] : a.pvel_new.p[pr];
            float r = (only_old ? a.prho_old.p[pr] : a.prho_new.p[pr])[ri], x = vn[vi], y = vn[vi + BS3], z = vn[vi + 2 * BS3];
            if (blend) {
                const float* __restrict__ vo = a.pvel_old.p[pr];
                const float ro = a.prho_old.p[pr][ri], xo = vo[vi], yo = vo[vi + BS3], zo = vo[vi + 2 * BS3];
                r = ro * (1.0f - tw) + r * tw; x = xo * (1.0f - tw) + x * tw; y = yo * (1.0f - tw) + y * tw; z = zo * (1.0f - tw) + z * tw;
            }
            cr[c] = r; cux[c] = x; cuy[c] = y; cuz[c] = z;
        } else if (c == 0) {
            cr[c] = 1.0f; cux[c] = 0.0f; cuy[c] = 0.0f; cuz[c] = 0.0f;
        } else {
            cr[c] = cr[0]; cux[c] = cux[0]; cuy[c] = cuy[0]; cuz[c] = cuz[0];
        }
    }
    // trilinear interpolation in the reference's order x, y, z (:110-118); w = 0.25 for the first cell of a pair, 0.75 for the second
    auto trilin = [](const float* v, float wx, float wy, float wz) {
        float c00 = v[0] * (1.0f - wx) + v[1] * wx;
        float c01 = v[4] * (1.0f - wx) + v[5] * wx;
        float c10 = v[2] * (1.0f - wx) + v[3] * wx;
        float c11 = v[6] * (1.0f - wx) + v[7] * wx;
        float c0 = c00 * (1.0f - wy) + c10 * wy;
        float c1 = c01 * (1.0f - wy) + c11 * wy;
        return c0 * (1.0f - wz) + c1 * wz;
    };
    float rho_i[8], ux_i[8], uy_i[8], uz_i[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        if (cells & (1u << m)) {
            const float wx = (m & 1) ? 0.75f : 0.25f, wy = (m & 2) ? 0.75f : 0.25f, wz = (m & 4) ? 0.75f : 0.25f;
            rho_i[m] = trilin(cr, wx, wy, wz); ux_i[m] = trilin(cux, wx, wy, wz); uy_i[m] = trilin(cuy, wx, wy, wz); uz_i[m] = trilin(cuz, wx, wy, wz);
        }
    }
    const float tau_c = a.tau_parent - 0.5f, tau_f = a.tau - 0.5f;
    const float scale = tau_c > 1.0e-6f ? fminf(fmaxf(tau_f / tau_c, 0.01f), 100.0f) : 1.0f;

    float* __restrict__ dst = a.f_ghost + (size_t)g * (Q * BS3) + (z0 * 64 + y0 * 8 + x0);
    for (uint32_t km = word & 0x7FFFFFFu; km; km &= km - 1) {
        const int k = __ffs(km) - 1;
        const int kx = k % 3 - 1, ky = (k / 3) % 3 - 1, kz = k / 9 - 1;
        const int d2 = kx * kx + ky * ky + kz * kz;
        const float w_k = d2 == 0 ? 8.0f / 27.0f : d2 == 1 ? 2.0f / 27.0f : d2 == 2 ? 1.0f / 54.0f : 1.0f / 216.0f;
        float cf[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (pb[c] >= 0) {
                const int pr = pb[c] >> PTR_RANK_SHIFT, pl = pb[c] & PTR_LOCAL_MASK;
                const size_t fi = ((size_t)pl * Q + k) * BS3 + loc[c];
                float v = (only_old ? a.pf_old.p[pr] : a.pf_new.p[pr])[fi];
                if (blend) v = a.pf_old.p[pr][fi] * (1.0f - tw) + v * tw;
                cf[c] = v;
            } else cf[c] = c == 0 ? w_k : cf[0];
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            if (cells & (1u << m)) {
                const float wx = (m & 1) ? 0.75f : 0.25f, wy = (m & 2) ? 0.75f : 0.25f, wz = (m & 4) ? 0.75f : 0.25f;
                const float f_int = trilin(cf, wx, wy, wz);
                const float feq_int = calc_eq(rho_i[m], ux_i[m], uy_i[m], uz_i[m], w_k, (float)kx, (float)ky, (float)kz);
                const float f_neq = f_int - feq_int;
                dst[k * BS3 + ((m >> 2) & 1) * 64 + ((m >> 1) & 1) * 8 + (m & 1)] = feq_int + f_neq * scale;
            }
        }
    }
}

// Block-cooperative variant of the interface pre-pass (LUDWIG_PREPASS=block): one CTA per ghost block.  Every ghost cell
// of a block interpolates from the same 5 x 5 x 5 parent cells (1-based parent coords 4 gb .. 4 gb + 4 per axis), so the
// CTA first stages those cells — rho, u and every population some cell of the block is pulled in, already blended in time —
// in shared memory with independent, row-coalesced loads (each parent value is fetched ONCE per ghost block instead of once
// per group that touches it), then interpolates out of shared memory, parallel over (group, cell) and (group, direction).
// Same arithmetic in the same order as ghost_interp_kernel: the two produce identical bits.
__global__ void __launch_bounds__(128) ghost_interp_block_kernel(const GhostArgs a) {
    __shared__ float s_val[4 + Q][125];      // 0 rho, 1..3 u, 4 + k populations; index pz * 25 + py * 5 + px
    __shared__ int s_pb[125], s_loc[125];    // rank-encoded parent block (-1 invalid) and in-block cell of every staged cell
    __shared__ float s_cell[56][8][4];       // interpolated rho, u of every ghost cell of the listed groups
    __shared__ uint32_t s_union;
    __shared__ int s_k[Q], s_nk;
    const int g = blockIdx.x, t = threadIdx.x;
    const int i0 = a.gstart[g], ng = a.gstart[g + 1] - i0;
    if (ng == 0) return;
    if (t == 0) s_union = 0;
    __syncthreads();
    if (t < ng) atomicOr(&s_union, a.gmask[i0 + t] & 0x7FFFFFFu);
    const int4 gb = *reinterpret_cast<const int4*>(a.gcoord + (size_t)g * 4);
    if (t < 125) {
        const int px = t % 5, py = (t / 5) % 5, pz = t / 25;
        const int pgx = max(1, 4 * gb.x + px), pgy = max(1, 4 * gb.y + py), pgz = max(1, 4 * gb.z + pz);   // :44-46 clamp (see ghost_interp_kernel)
        const int bx = (pgx - 1) >> 3, by = (pgy - 1) >> 3, bz = (pgz - 1) >> 3;
        int pb = -1;
        if (bx >= 0 && bx < a.pdimx && by >= 0 && by < a.pdimy && bz >= 0 && bz < a.pdimz) pb = a.pptr[bx + a.pdimx * (by + a.pdimy * bz)];
        s_pb[t] = pb;
        s_loc[t] = ((pgx - 1) & 7) + 8 * ((pgy - 1) & 7) + 64 * ((pgz - 1) & 7);
    }
    __syncthreads();
    if (t == 0) {
        int n = 0;
        for (uint32_t km = s_union; km; km &= km - 1) s_k[n++] = __ffs(km) - 1;
        s_nk = n;
    }
    __syncthreads();
    const int nk = s_nk;
    const bool only_old = a.use_temporal == 1 && a.tw == 0.0f;
    const bool blend = a.use_temporal == 1 && a.tw < 0.99f && !only_old;
    const float tw = a.tw;
    // ---- stage: (4 + nk) quantities x 125 cells
    for (int i = t; i < (4 + nk) * 125; i += 128) {
        const int qn = i / 125, c = i - qn * 125;
        const int pb = s_pb[c];
        if (pb < 0) continue;
        const int pr = pb >> PTR_RANK_SHIFT, pl = pb & PTR_LOCAL_MASK;
        const float *pn, *po;
        size_t idx;
        if (qn == 0) { pn = a.prho_new.p[pr]; po = a.prho_old.p[pr]; idx = (size_t)pl * BS3 + s_loc[c]; }
        else if (qn < 4) { pn = a.pvel_new.p[pr]; po = a.pvel_old.p[pr]; idx = ((size_t)pl * 3 + (qn - 1)) * BS3 + s_loc[c]; }
        else { pn = a.pf_new.p[pr]; po = a.pf_old.p[pr]; idx = ((size_t)pl * Q + s_k[qn - 4]) * BS3 + s_loc[c]; }
        float v = (only_old ? po : pn)[idx];
        if (blend) v = po[idx] * (1.0f - tw) + v * tw;
        s_val[qn < 4 ? qn : 4 + s_k[qn - 4]][c] = v;
    }
    __syncthreads();
    auto trilin = [](const float* v, float wx, float wy, float wz) {
        float c00 = v[0] * (1.0f - wx) + v[1] * wx;
        float c01 = v[4] * (1.0f - wx) + v[5] * wx;
        float c10 = v[2] * (1.0f - wx) + v[3] * wx;
        float c11 = v[6] * (1.0f - wx) + v[7] * wx;
        float c0 = c00 * (1.0f - wy) + c10 * wy;
        float c1 = c01 * (1.0f - wy) + c11 * wy;
        return c0 * (1.0f - wz) + c1 * wz;
    };
    // the 8 corners of group q (corner order 000,100,010,110,001,101,011,111, x fastest) as staged-cell indices
    auto corner_cell = [](int q, int c) {
        return ((q >> 4) + ((c >> 2) & 1)) * 25 + (((q >> 2) & 3) + ((c >> 1) & 1)) * 5 + (q & 3) + (c & 1);
    };
    // ---- rho, u of every listed ghost cell
    for (int i = t; i < ng * 8; i += 128) {
        const int gi = i >> 3, m = i & 7;
        if (!(a.gcells8[i0 + gi] & (1u << m))) continue;
        const int q = a.gcell[i0 + gi] & 63;
        const float wx = (m & 1) ? 0.75f : 0.25f, wy = (m & 2) ? 0.75f : 0.25f, wz = (m & 4) ? 0.75f : 0.25f;
#pragma unroll
        for (int qn = 0; qn < 4; ++qn) {
            float cv[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int cc = corner_cell(q, c);
                cv[c] = s_pb[cc] >= 0 ? s_val[qn][cc] : (c == 0 ? (qn == 0 ? 1.0f : 0.0f) : cv[0]);
            }
            s_cell[gi][m][qn] = trilin(cv, wx, wy, wz);
        }
    }
    __syncthreads();
    const float tau_c = a.tau_parent - 0.5f, tau_f = a.tau - 0.5f;
    const float scale = tau_c > 1.0e-6f ? fminf(fmaxf(tau_f / tau_c, 0.01f), 100.0f) : 1.0f;
    // ---- populations: one item per (group, direction of the block's union mask)
    for (int i = t; i < ng * nk; i += 128) {
        const int gi = i / nk, k = s_k[i - gi * nk];
        if (!(a.gmask[i0 + gi] & (1u << k))) continue;
        const int q = a.gcell[i0 + gi] & 63;
        const uint32_t cells = a.gcells8[i0 + gi];
        const int kx = k % 3 - 1, ky = (k / 3) % 3 - 1, kz = k / 9 - 1;
        const int d2 = kx * kx + ky * ky + kz * kz;
        const float w_k = d2 == 0 ? 8.0f / 27.0f : d2 == 1 ? 2.0f / 27.0f : d2 == 2 ? 1.0f / 54.0f : 1.0f / 216.0f;
        float cf[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int cc = corner_cell(q, c);
            cf[c] = s_pb[cc] >= 0 ? s_val[4 + k][cc] : (c == 0 ? w_k : cf[0]);
        }
        const int x0 = (q & 3) * 2, y0 = ((q >> 2) & 3) * 2, z0 = (q >> 4) * 2;
        float* __restrict__ dst = a.f_ghost + (size_t)g * (Q * BS3) + k * BS3 + (z0 * 64 + y0 * 8 + x0);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            if (cells & (1u << m)) {
                const float wx = (m & 1) ? 0.75f : 0.25f, wy = (m & 2) ? 0.75f : 0.25f, wz = (m & 4) ? 0.75f : 0.25f;
                const float f_int = trilin(cf, wx, wy, wz);
                const float feq_int = calc_eq(s_cell[gi][m][0], s_cell[gi][m][1], s_cell[gi][m][2], s_cell[gi][m][3], w_k, (float)kx, (float)ky, (float)kz);
                const float f_neq = f_int - feq_int;
                dst[((m >> 2) & 1) * 64 + ((m >> 1) & 1) * 8 + (m & 1)] = feq_int + f_neq * scale;
            }
        }
    }
}
