// k1_generic.cuh — K1, the fused pull-stream + boundary conditions + interface interpolation +
// bounce-back + sponge + WMLES wall force + WALE + regularized-BGK collision kernel, in the form that
// handles EVERY kind of block (domain faces, refinement interfaces, obstacles).
//
// Replaces stream_collide_kernel_v2! (physics_kernels.jl:9-358) with its device functions
// gradient_noise (physics_utils.jl:17-28), compute_velocity_gradients (physics_utils.jl:45-83) and
// interpolate_with_rescaling (physics_interpolation.jl:16-138).
//
// The floating-point expressions below are written in the reference's operation order.  This header
// is compiled twice: k1_generic_strict.cu with -fmad=false (no FMA contraction: bit-comparable with
// the CPU oracle) and k1_generic_fast.cu with contraction on.  In fast mode this
// kernel only runs the blocks that are not BF_INTERIOR; those go to k1_interior.cu.
//
// One CTA = half a block (256 threads = 4 z-planes); a warp = 4 x-rows of one z-plane, so the in-block
// part of every direction is a set of fully used 32 B sectors.
#pragma once
#include "ludwig_internal.h"

namespace ludwig {
namespace K1_NS {

#include "k1_boundary.cuh"

__device__ __forceinline__ void vel_neighbor(const K1Args& a, const int* s_nbr, int b, int x, int y, int z, int dx, int dy, int dz,
                                             float& ux, float& uy, float& uz) {
    int nx = x + dx, ny = y + dy, nz = z + dz;
    int nbi = b;
    if (((nx | ny | nz) & ~7) != 0) {
        int ox = nx < 0 ? -1 : (nx > 7 ? 1 : 0), oy = ny < 0 ? -1 : (ny > 7 ? 1 : 0), oz = nz < 0 ? -1 : (nz > 7 ? 1 : 0);
        nbi = s_nbr[(ox + 1) + (oy + 1) * 3 + (oz + 1) * 9];
        if (nbi < 0) { nbi = b; nx = x; ny = y; nz = z; }   // fall back to the cell's own value
    }
    size_t vi = (size_t)nbi * 3 * BS3 + ((nz & 7) * 64 + (ny & 7) * 8 + (nx & 7));
    ux = a.vel_in[vi]; uy = a.vel_in[vi + BS3]; uz = a.vel_in[vi + 2 * BS3];
}

__global__ void __launch_bounds__(256, 2) K1_KERNEL_NAME(const K1Args a) {
    __shared__ int s_nbr[27];
    const int slot = blockIdx.x >> 1;
    if (slot >= a.n_list) return;
    const int b = a.list ? a.list[slot] : slot;
    if (threadIdx.x < 27) s_nbr[threadIdx.x] = a.nbr[(size_t)b * 27 + threadIdx.x];
    __syncthreads();

    const int c = (blockIdx.x & 1) * 256 + threadIdx.x;
    const int x = c & 7, y = (c >> 3) & 7, z = c >> 6;
    const int4 bc = *reinterpret_cast<const int4*>(a.bcoord + (size_t)b * 4);
    const int gx = bc.x * BS + x + 1, gy = bc.y * BS + y + 1, gz = bc.z * BS + z + 1;   // 1-based like the reference
    const uint32_t bflags = (uint32_t)bc.w;

    const size_t cell = (size_t)b * BS3 + c;
    const float* __restrict__ fin_cell = a.f_in + (size_t)b * Q * BS3 + c;
    const bool is_obs = (bflags & BF_OBSTACLE) ? (a.obstacle[cell] != 0) : false;

    float rho = 0.0f, jx = 0.0f, jy = 0.0f, jz = 0.0f;
    float f_stored[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        const int cx = lat_cx(k), cy = lat_cy(k), cz = lat_cz(k);
        const int sx = x - cx, sy = y - cy, sz = z - cz;
        float val;
        if (((sx | sy | sz) & ~7) == 0) {
            val = fin_cell[k * BS3 - (cx + 8 * cy + 64 * cz)];
        } else {
            int ox = sx < 0 ? -1 : (sx > 7 ? 1 : 0), oy = sy < 0 ? -1 : (sy > 7 ? 1 : 0), oz = sz < 0 ? -1 : (sz > 7 ? 1 : 0);
            int nbi = s_nbr[(ox + 1) + (oy + 1) * 3 + (oz + 1) * 9];
            if (nbi >= 0) val = a.f_in[((size_t)nbi * Q + k) * BS3 + ((sz & 7) * 64 + (sy & 7) * 8 + (sx & 7))];
            else val = pull_missing(a, fin_cell, k, gx, gy, gz);
        }
        f_stored[k] = val;
        rho += val;
        jx += val * (float)cx;
        jy += val * (float)cy;
        jz += val * (float)cz;
    }

    float* __restrict__ fout_cell = a.f_out + (size_t)b * Q * BS3 + c;
    float* __restrict__ vout_cell = a.vel_out + (size_t)b * 3 * BS3 + c;

    if (is_obs) {
        vout_cell[0] = 0.0f; vout_cell[BS3] = 0.0f; vout_cell[2 * BS3] = 0.0f;
        a.rho_out[cell] = 1.0f;
#pragma unroll
        for (int k = 0; k < 27; ++k) fout_cell[k * BS3] = f_stored[26 - k];
        return;
    }

    rho = fmaxf(rho, 0.01f);
    float inv_rho = 1.0f / rho;
    float ux = jx * inv_rho, uy = jy * inv_rho, uz = jz * inv_rho;

    if (bflags & BF_SPONGE) {
        float sp = a.sponge[cell];
        if (sp > 0.0f) {
            const float rho_target = 1.0f, ux_target = k1_u_inlet(a);
            rho = rho * (1.0f - sp) + rho_target * sp;
            ux = ux * (1.0f - sp) + ux_target * sp;
            uy = uy * (1.0f - sp);
            uz = uz * (1.0f - sp);
            if (a.sponge_blend == 1) {
#pragma unroll
                for (int k = 0; k < 27; ++k) {
                    float feq_target = calc_eq(rho_target, ux_target, 0.0f, 0.0f, lat_w(k), (float)lat_cx(k), (float)lat_cy(k), (float)lat_cz(k));
                    f_stored[k] = f_stored[k] * (1.0f - sp) + feq_target * sp;
                }
            }
        }
    }

    float Fx_wall = 0.0f, Fy_wall = 0.0f, Fz_wall = 0.0f;
    if (a.wm == 1 && (bflags & BF_WALLDIST)) {
        float dist_wall = a.wall_dist[cell];
        if (dist_wall > 0.0f && dist_wall < 10.0f) {
            float u_mag = sqrtf(ux * ux + uy * uy + uz * uz);
            float nu_visc = (a.tau - 0.5f) / 3.0f;
            if (u_mag > 1.0e-6f && nu_visc > 1.0e-10f) {
                float u_tau = u_mag * pow32(nu_visc / (dist_wall * u_mag + 1.0e-10f), 1.0f / 7.0f) * pow32(2.0f * 8.3f, -1.0f / 7.0f);
                u_tau = fmaxf(u_tau, 1.0e-6f);
                float y_p = u_tau * dist_wall / nu_visc;
                if (y_p > 11.81f) {
                    float u_plus_law = (1.0f / KAPPA) * log32(y_p) + 5.2f;
                    if (u_plus_law > 0.1f) {
                        u_tau = u_tau * ((u_mag / u_tau) / u_plus_law);
                        u_tau = fmaxf(u_tau, 1.0e-6f);
                    }
                }
                float tau_wall = rho * u_tau * u_tau;
                float tau_res = rho * nu_visc * (u_mag / dist_wall);
                if (tau_wall > tau_res) {
                    float force_mag = (tau_wall - tau_res) / dist_wall;
                    Fx_wall = -force_mag * ux / u_mag;
                    Fy_wall = -force_mag * uy / u_mag;
                    Fz_wall = -force_mag * uz / u_mag;
                }
            }
        }
    }

    float ux_eq = ux + 0.5f * Fx_wall * inv_rho;
    float uy_eq = uy + 0.5f * Fy_wall * inv_rho;
    float uz_eq = uz + 0.5f * Fz_wall * inv_rho;
    float usq_eq = ux_eq * ux_eq + uy_eq * uy_eq + uz_eq * uz_eq;

    vout_cell[0] = ux; vout_cell[BS3] = uy; vout_cell[2 * BS3] = uz;
    a.rho_out[cell] = rho;

    float uxE, uyE, uzE, uxW, uyW, uzW, uxN, uyN, uzN, uxS, uyS, uzS, uxT, uyT, uzT, uxB, uyB, uzB;
    vel_neighbor(a, s_nbr, b, x, y, z, 1, 0, 0, uxE, uyE, uzE);
    vel_neighbor(a, s_nbr, b, x, y, z, -1, 0, 0, uxW, uyW, uzW);
    vel_neighbor(a, s_nbr, b, x, y, z, 0, 1, 0, uxN, uyN, uzN);
    vel_neighbor(a, s_nbr, b, x, y, z, 0, -1, 0, uxS, uyS, uzS);
    vel_neighbor(a, s_nbr, b, x, y, z, 0, 0, 1, uxT, uyT, uzT);
    vel_neighbor(a, s_nbr, b, x, y, z, 0, 0, -1, uxB, uyB, uzB);
    float g11 = 0.5f * (uxE - uxW), g12 = 0.5f * (uxN - uxS), g13 = 0.5f * (uxT - uxB);
    float g21 = 0.5f * (uyE - uyW), g22 = 0.5f * (uyN - uyS), g23 = 0.5f * (uyT - uyB);
    float g31 = 0.5f * (uzE - uzW), g32 = 0.5f * (uzN - uzS), g33 = 0.5f * (uzT - uzB);

    float gsq11 = g11 * g11 + g12 * g21 + g13 * g31;
    float gsq12 = g11 * g12 + g12 * g22 + g13 * g32;
    float gsq13 = g11 * g13 + g12 * g23 + g13 * g33;
    float gsq21 = g21 * g11 + g22 * g21 + g23 * g31;
    float gsq22 = g21 * g12 + g22 * g22 + g23 * g32;
    float gsq23 = g21 * g13 + g22 * g23 + g23 * g33;
    float gsq31 = g31 * g11 + g32 * g21 + g33 * g31;
    float gsq32 = g31 * g12 + g32 * g22 + g33 * g32;
    float gsq33 = g31 * g13 + g32 * g23 + g33 * g33;
    float tr_gsq = gsq11 + gsq22 + gsq33;
    float tr_term = tr_gsq / 3.0f;
    float Sd11 = gsq11 - tr_term, Sd22 = gsq22 - tr_term, Sd33 = gsq33 - tr_term;
    float Sd12 = 0.5f * (gsq12 + gsq21), Sd13 = 0.5f * (gsq13 + gsq31), Sd23 = 0.5f * (gsq23 + gsq32);
    float S12 = 0.5f * (g12 + g21), S13 = 0.5f * (g13 + g31), S23 = 0.5f * (g23 + g32);
    float OP1 = Sd11 * Sd11 + Sd22 * Sd22 + Sd33 * Sd33 + 2.0f * (Sd12 * Sd12 + Sd13 * Sd13 + Sd23 * Sd23);
    float OP2 = g11 * g11 + g22 * g22 + g33 * g33 + 2.0f * (S12 * S12 + S13 * S13 + S23 * S23);
    float nu_eddy = 0.0f;
    if (OP1 > 1.0e-12f) {
        float OP1_32 = OP1 * sqrtf(OP1);
        float OP2_52 = OP2 * OP2 * sqrtf(fmaxf(OP2, 1.0e-12f));
        float denom = OP2_52 + OP1 * sqrtf(sqrtf(fmaxf(OP1, 1.0e-12f)));
        if (denom > 1.0e-12f) nu_eddy = (a.c_wale * a.c_wale) * OP1_32 / denom;
    }
    nu_eddy = fmaxf(nu_eddy, a.nu_bg);
    float tau_turb = a.tau + nu_eddy * 3.0f;
    float omega = 1.0f / fmaxf(tau_turb, 0.500001f);

    float Pi_xx = 0.0f, Pi_yy = 0.0f, Pi_zz = 0.0f, Pi_xy = 0.0f, Pi_yz = 0.0f, Pi_zx = 0.0f;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        const float cx_f = (float)lat_cx(k), cy_f = (float)lat_cy(k), cz_f = (float)lat_cz(k);
        float cu = cx_f * ux_eq + cy_f * uy_eq + cz_f * uz_eq;
        float feq = rho * lat_w(k) * (1.0f + 3.0f * cu + 4.5f * cu * cu - 1.5f * usq_eq);
        float f_neq = f_stored[k] - feq;
        Pi_xx += f_neq * cx_f * cx_f;
        Pi_yy += f_neq * cy_f * cy_f;
        Pi_zz += f_neq * cz_f * cz_f;
        Pi_xy += f_neq * cx_f * cy_f;
        Pi_yz += f_neq * cy_f * cz_f;
        Pi_zx += f_neq * cz_f * cx_f;
    }
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        const float cx_f = (float)lat_cx(k), cy_f = (float)lat_cy(k), cz_f = (float)lat_cz(k);
        const float w_k = lat_w(k);
        float cu = cx_f * ux_eq + cy_f * uy_eq + cz_f * uz_eq;
        float feq = rho * w_k * (1.0f + 3.0f * cu + 4.5f * cu * cu - 1.5f * usq_eq);
        float force_term = w_k * 3.0f *
                           ((cx_f - ux + 3.0f * cu * cx_f) * Fx_wall + (cy_f - uy + 3.0f * cu * cy_f) * Fy_wall +
                            (cz_f - uz + 3.0f * cu * cz_f) * Fz_wall);
        float Q_xx = cx_f * cx_f - CS2_PHYSICS, Q_yy = cy_f * cy_f - CS2_PHYSICS, Q_zz = cz_f * cz_f - CS2_PHYSICS;
        float f_neq_reg = w_k * 4.5f *
                          (Pi_xx * Q_xx + Pi_yy * Q_yy + Pi_zz * Q_zz +
                           2.0f * (Pi_xy * cx_f * cy_f + Pi_yz * cy_f * cz_f + Pi_zx * cz_f * cx_f));
        fout_cell[k * BS3] = feq + (1.0f - omega) * f_neq_reg + (1.0f - 0.5f * omega) * force_term;
    }
}

}  // namespace K1_NS

void K1_LAUNCH_NAME(const K1Args& a, cudaStream_t s) {
    if (a.n_list <= 0) return;
    K1_NS::K1_KERNEL_NAME<<<2 * a.n_list, 256, 0, s>>>(a);
}

}  // namespace ludwig
