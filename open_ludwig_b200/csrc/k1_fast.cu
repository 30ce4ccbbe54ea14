// k1_fast.cu — K1 (fused pull-stream + BCs + interface interpolation + bounce-back + sponge + WMLES wall
// force + WALE + regularized-BGK collision) for fast mode.  Replaces stream_collide_kernel_v2!
// (physics_kernels.jl:9-358).  Two instantiations of one template:
//
//   PLAIN  blocks whose 26 neighbours exist and that hold no obstacle / sponge / near-wall cell (the bulk of
//          every large level).  The regularized collision needs only the 10 moments rho, j, sum f c c — not the
//          27 populations — so loads are streamed straight into the moment accumulators and no f[27] array
//          lives in registers.
//   FULL   every other block (two instantiations, with / without missing-neighbour handling): missing neighbours (domain faces, refinement interfaces -> k1_boundary.cuh),
//          obstacle cells (full-way bounce-back), sponge blending, wall-model force:
//          bounce-back returns the pulled populations (re-pulled at the end, solid cells skip the collision).  Per-block flag bits gate each feature uniformly.
//
// Arithmetic (fast mode = FMA on, sums regrouped; parity is carried by the strict build):
//   * opposite directions are paired: s_k = f_k + f_(26-k), d_k = f_k - f_(26-k), k = 0..12;
//   * Pi_ab = sum_k (f_k - feq_k) c_a c_b = sum_k f_k c_a c_b - rho (delta_ab/3 + u_a u_b)   (exact identity
//     for the second-order equilibrium on D3Q27);
//   * feq_k, feq_(26-k) share their even part; f_neq_reg and the even part of the Guo force term are even;
//   * sponge blending of the populations (physics_kernels.jl:192-198) is linear, so it is applied to the
//     raw second moments instead of to 27 populations.
//
// Blackwell specifics: one thread owns TWO x-adjacent cells and all arithmetic is packed FP32x2
// (FADD2 / FMUL2 / FFMA2, sm_100+), halving the issue slots of the collision; cx = 0 populations and all
// stores are 64-bit accesses; a warp is one z-plane (64 cells), a CTA (256 threads) one 8^3 block; CTAs walk
// the blocks in Morton order so that halo sectors are L2 hits.
#include <climits>
#include <cstdlib>

#include "ludwig_internal.h"

namespace ludwig {
namespace k1f {

#include "k1_boundary.cuh"

typedef float2 v2;
__device__ __forceinline__ v2 V(float s) { return make_float2(s, s); }
__device__ __forceinline__ v2 vadd(v2 a, v2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ v2 vsub(v2 a, v2 b) { return __ffma2_rn(b, V(-1.0f), a); }
__device__ __forceinline__ v2 vneg(v2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ v2 vmul(v2 a, v2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ v2 vfma(v2 a, v2 b, v2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ v2 vmax(v2 a, float s) { return make_float2(fmaxf(a.x, s), fmaxf(a.y, s)); }
// MUFU approximations (max rel. error 2^-23 / 2^-22): one instruction each, no slow-path call
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ v2 vrcp(v2 a) { return make_float2(rcp_approx(a.x), rcp_approx(a.y)); }
__device__ __forceinline__ v2 vsqrt(v2 a) { return make_float2(sqrt_approx(a.x), sqrt_approx(a.y)); }
// streaming store (st.global.cs): outputs are not re-read by this kernel, keep L2 for the halo sectors of f_in
__device__ __forceinline__ void st2(float* p, v2 v) { __stcs(reinterpret_cast<float2*>(p), v); }
__device__ __forceinline__ v2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }

constexpr float W0 = 8.0f / 27.0f, W1 = 2.0f / 27.0f, W2 = 1.0f / 54.0f, W3 = 1.0f / 216.0f;
__host__ __device__ constexpr float wk(int k) {
    return (lat_cx(k) != 0) + (lat_cy(k) != 0) + (lat_cz(k) != 0) == 0   ? W0
           : (lat_cx(k) != 0) + (lat_cy(k) != 0) + (lat_cz(k) != 0) == 1 ? W1
           : (lat_cx(k) != 0) + (lat_cy(k) != 0) + (lat_cz(k) != 0) == 2 ? W2
                                                                         : W3;
}

struct Moments {
    v2 rho, jx, jy, jz, Pxx, Pyy, Pzz, Pxy, Pyz, Pzx;
};

// add the pair (f_k, f_(26-k)), k in 0..12, to the raw moments.  Populations are shifted by their rest weight
// first (f - w_k, exact by Sterbenz), so every sum below adds numbers of size |f - w| ~ 1e-3 w instead of
// O(0.1): the moments rho - 1 and sum f c c - delta/3 carry ~1e-10 of round-off instead of ~3e-8.
template <int K>
__device__ __forceinline__ void acc_pair(Moments& m, v2 fk, v2 fo) {
    constexpr int cx = lat_cx(K), cy = lat_cy(K), cz = lat_cz(K);
    const v2 s = vadd(vadd(fk, V(-wk(K))), vadd(fo, V(-wk(K)))), d = vsub(fk, fo);
    m.rho = vadd(m.rho, s);
    if (cx == 1) m.jx = vadd(m.jx, d); else if (cx == -1) m.jx = vsub(m.jx, d);
    if (cy == 1) m.jy = vadd(m.jy, d); else if (cy == -1) m.jy = vsub(m.jy, d);
    if (cz == 1) m.jz = vadd(m.jz, d); else if (cz == -1) m.jz = vsub(m.jz, d);
    if (cx != 0) m.Pxx = vadd(m.Pxx, s);
    if (cy != 0) m.Pyy = vadd(m.Pyy, s);
    if (cz != 0) m.Pzz = vadd(m.Pzz, s);
    if (cx * cy == 1) m.Pxy = vadd(m.Pxy, s); else if (cx * cy == -1) m.Pxy = vsub(m.Pxy, s);
    if (cy * cz == 1) m.Pyz = vadd(m.Pyz, s); else if (cy * cz == -1) m.Pyz = vsub(m.Pyz, s);
    if (cz * cx == 1) m.Pzx = vadd(m.Pzx, s); else if (cz * cx == -1) m.Pzx = vsub(m.Pzx, s);
}

// Wall-model force of one cell (physics_kernels.jl:206-236), scalar: only near-wall cells get here.  Fast mode: the
// 1/7 power and the logarithm are MUFU lg2/ex2 sequences (relative error ~1e-7, far below the model's own accuracy and
// the fast-mode round-off budget), divisions are reciprocal multiplies, (2 * 8.3)^(-1/7) is a constant; inlined, so a
// warp with near-wall lanes spends ~60 instructions per cell here instead of two libdevice powf calls and five IEEE divides.
__device__ __forceinline__ float3 wall_force(float dist_wall, float rho, float ux, float uy, float uz, float tau) {
    float3 F = make_float3(0.f, 0.f, 0.f);
    if (dist_wall > 0.0f && dist_wall < 10.0f) {
        const float u_mag = sqrt_approx(ux * ux + uy * uy + uz * uz);
        const float nu_visc = (tau - 0.5f) * (1.0f / 3.0f);
        if (u_mag > 1.0e-6f && nu_visc > 1.0e-10f) {
            const float inv_d = rcp_approx(dist_wall), inv_nu = rcp_approx(nu_visc);
            const float ratio = nu_visc * rcp_approx(dist_wall * u_mag + 1.0e-10f);
            float u_tau = u_mag * exp2f((1.0f / 7.0f) * __log2f(ratio)) * 0.66942024f;   // (2 * 8.3)^(-1/7)
            u_tau = fmaxf(u_tau, 1.0e-6f);
            const float y_p = u_tau * dist_wall * inv_nu;
            if (y_p > 11.81f) {
                const float u_plus_law = (1.0f / KAPPA) * (0.69314718f * __log2f(y_p)) + 5.2f;
                if (u_plus_law > 0.1f) u_tau = fmaxf(u_mag * rcp_approx(u_plus_law), 1.0e-6f);   // u_tau (u_mag / u_tau) / u_plus
            }
            const float tau_wall = rho * u_tau * u_tau;
            const float tau_res = rho * nu_visc * (u_mag * inv_d);
            if (tau_wall > tau_res) {
                const float s = -(tau_wall - tau_res) * inv_d * rcp_approx(u_mag);
                F.x = s * ux; F.y = s * uy; F.z = s * uz;
            }
        }
    }
    return F;
}


#include "k1_ghost.cuh"
#include "k1_common.cuh"



// FULL : see file header.   VELFB : some axis neighbour may lack a velocity field (ghost block or domain face) ->
// fall back to the cell's own value (physics_utils.jl:69).
//        MISS : some neighbour block may be absent (domain face) -> k1_boundary.cuh; blocks with features but all 26
//        neighbours present (the near-body bulk) use FULL without MISS and never carry that code.
template <bool FULL, bool VELFB, bool MISS>
__device__ __forceinline__ void fast_block(const K1Args& a, const int b, const int t, const float* __restrict__ fbase, const long long* s_fo, const long long* s_vo) {
    const int p = t & 3, y = (t >> 2) & 7, z = t >> 5;
    const int x0 = 2 * p;
    const int c0 = 2 * t;   // z*64 + y*8 + x0
    uint32_t bflags = BF_INTERIOR;
    int gx = 0, gy = 0, gz = 0;   // 1-based global coords of cell A (FULL only)
    if (FULL) {
        const int4 bc = *reinterpret_cast<const int4*>(a.bcoord + (size_t)b * 4);
        bflags = (uint32_t)bc.w;
        if (MISS) { gx = bc.x * BS + x0 + 1; gy = bc.y * BS + y + 1; gz = bc.z * BS + z + 1; }
    }
    // domain x faces in closed form (cell A on the inlet plane / cell B on the outlet plane; an inlet source wins over y / z faces
    // and over the outlet test exactly as in pull_missing: physics_kernels.jl:99-113)
    const bool at_inlet = MISS && gx == 1, at_outlet = MISS && gx + 1 == a.nxg && a.nxg > 1;
    XFace xf{0.f, 0.f};
    if (MISS && (at_inlet || at_outlet)) xf = x_face_equilibria(a, gy, gz);
    const float* __restrict__ fin_own = fbase + s_fo[13] + c0;   // own cell A, direction 0

    // source-row bookkeeping per axis: index j = c + 1 for lattice component c in {-1,0,1}; source = coord - c
    int yoff[3], ydir[3], zoff[3], zdir[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        int ys = y - (j - 1), zs = z - (j - 1);
        yoff[j] = (ys & 7) * 8;
        ydir[j] = (ys < 0 ? 0 : (ys > 7 ? 2 : 1)) * 3;
        zoff[j] = (zs & 7) * 64;
        zdir[j] = (zs < 0 ? 0 : (zs > 7 ? 2 : 1)) * 9;
    }
    const int dM = p > 0 ? 1 : 0, xM = p > 0 ? x0 - 1 : 7;   // where cell A's x-1 lives
    const int dP = p < 3 ? 1 : 2, xP = p < 3 ? x0 + 2 : 0;   // where cell B's x+1 lives

    // ---- pull-stream (physics_kernels.jl:62-149).  combo (jy,jz) serves the three directions km,k0,kp.
    auto pull3 = [&](int jy, int jz, v2& fm, v2& f0, v2& fp) {
        const int loc = zoff[jz] + yoff[jy];
        const int dir = zdir[jz] + ydir[jy];
        const int k0 = 1 + 3 * jy + 9 * jz, kp = k0 + 1, km = k0 - 1;
        const long long o0 = s_fo[dir + 1], oM = s_fo[dir + dM], oP = s_fo[dir + dP];
        if (!MISS || (o0 != MISSING && oM != MISSING && oP != MISSING)) {
            const float* __restrict__ P0 = fbase + o0 + (loc + x0);
            const float* __restrict__ PM = fbase + oM + (loc + xM);
            const float* __restrict__ PP = fbase + oP + (loc + xP);
            f0 = ld2(P0 + k0 * BS3);
            fp = make_float2(PM[kp * BS3], P0[kp * BS3]);       // cx=+1: sources x0-1, x0
            fm = make_float2(P0[km * BS3 + 1], PP[km * BS3]);   // cx=-1: sources x0+1, x0+2
        } else {
            // some source block is missing: domain face or refinement interface (rare path)
            if (o0 != MISSING) {
                const float* __restrict__ P0 = fbase + o0 + (loc + x0);
                f0 = ld2(P0 + k0 * BS3); fp.y = P0[kp * BS3]; fm.x = P0[km * BS3 + 1];
            } else {
                f0.x = pull_missing(a, fin_own, k0, gx, gy, gz); f0.y = pull_missing(a, fin_own + 1, k0, gx + 1, gy, gz);
                fp.y = pull_missing(a, fin_own + 1, kp, gx + 1, gy, gz);
                fm.x = pull_missing(a, fin_own, km, gx, gy, gz);
            }
            fp.x = oM != MISSING ? fbase[oM + (loc + xM) + kp * BS3] : at_inlet ? lat_w_of(kp) * xf.p_in : pull_missing(a, fin_own, kp, gx, gy, gz);
            fm.y = oP != MISSING ? fbase[oP + (loc + xP) + km * BS3] : at_outlet ? lat_w_of(km) * xf.p_out : pull_missing(a, fin_own + 1, km, gx + 1, gy, gz);
        }
    };

    bool obsA = false, obsB = false;
    if (FULL && (bflags & BF_OBSTACLE)) {
        const uchar2 o = *reinterpret_cast<const uchar2*>(a.obstacle + (size_t)b * BS3 + c0);
        obsA = o.x != 0; obsB = o.y != 0;
    }
    float* __restrict__ fout = a.f_out + (size_t)b * (Q * BS3) + c0;
    float* __restrict__ vout = a.vel_out + (size_t)b * (3 * BS3) + c0;
    float* __restrict__ rout = a.rho_out + (size_t)b * BS3 + c0;

    // Full-way bounce-back (physics_kernels.jl:154-166): an obstacle cell returns every pulled population in the
    // opposite direction, f_out[26-k] = pulled f_k.  No arithmetic: the values are stored straight from the pull loop below
    // (one pass over f_in for every thread, solid or fluid); a thread with ONE obstacle cell stores its fluid cell with
    // 4-byte stores at the end, a thread whose two cells are both solid skips the collision.
    const bool anyobs = FULL && (obsA || obsB);
    auto bounce3 = [&](int jy, int jz, const v2& fm, const v2& f0, const v2& fp) {
        if (!anyobs) return;
        const int k0 = 1 + 3 * jy + 9 * jz;
        float* __restrict__ o = fout + (26 - (k0 + 1)) * BS3;   // slots 26-(k0+1), 26-k0, 26-(k0-1) are consecutive directions
        if (obsA && obsB) { st2(o, fp); st2(o + BS3, f0); st2(o + 2 * BS3, fm); }
        else if (obsA) { o[0] = fp.x; o[BS3] = f0.x; o[2 * BS3] = fm.x; }
        else { o[1] = fp.y; o[BS3 + 1] = f0.y; o[2 * BS3 + 1] = fm.y; }
    };

    Moments m;
    m.jx = m.jy = m.jz = m.Pxx = m.Pyy = m.Pzz = m.Pxy = m.Pyz = m.Pzx = V(0.f);
    {
        // combos c = jy + 3 jz; combo c and 8-c hold opposite directions: (km,k0,kp)(c) <-> (kp,k0,km)(8-c)
        v2 am, a0, ap, bm, b0, bp;
        pull3(1, 1, am, a0, ap);          // centre combo: k = 12,13,14
        bounce3(1, 1, am, a0, ap);
        // moments are accumulated for f - w_k.  The FP32 weights all carry the same relative error, sum_k w_k =
        // 1 + 7.45e-9 (the reference's equilibrium therefore creates that much mass per step); adding it back here
        // keeps rho = sum_k f_k exactly as the reference computes it, and makes the shifted second moments
        // consistent with its feq (sum_k w_k c_a c_b = (1 + eps) delta/3).
        m.rho = vadd(vadd(a0, V(-W0)), V(7.4505806e-9f));
        acc_pair<12>(m, am, ap);
#define LUDWIG_COMBO(JY, JZ)                                                   \
        pull3(JY, JZ, am, a0, ap);                                             \
        pull3(2 - (JY), 2 - (JZ), bm, b0, bp);                                 \
        bounce3(JY, JZ, am, a0, ap);                                           \
        bounce3(2 - (JY), 2 - (JZ), bm, b0, bp);                               \
        acc_pair<3 * (JY) + 9 * (JZ)>(m, am, bp);                              \
        acc_pair<3 * (JY) + 9 * (JZ) + 1>(m, a0, b0);                          \
        acc_pair<3 * (JY) + 9 * (JZ) + 2>(m, ap, bm);
        LUDWIG_COMBO(0, 0)
        LUDWIG_COMBO(1, 0)
        LUDWIG_COMBO(2, 0)
        LUDWIG_COMBO(0, 1)
#undef LUDWIG_COMBO
    }

    if (FULL && obsA && obsB) {       // both cells solid: bounce-back done, vel = 0, rho = 1 (:155-158)
        st2(vout, V(0.f)); st2(vout + BS3, V(0.f)); st2(vout + 2 * BS3, V(0.f)); st2(rout, V(1.0f));
        return;
    }

    // ---- previous-step velocities of the six axis neighbours (physics_utils.jl:45-83)
    v2 uE[3], uW[3], uN[3], uS[3], uT[3], uB[3];
    {
        const int row = z * 64 + y * 8;
        const float* __restrict__ vo = a.vel_in + s_vo[13] + c0;
        long long oM = s_vo[12 + dM], oP = s_vo[13 + (dP - 1)];
        long long oN = s_vo[y < 7 ? 13 : 16], oS = s_vo[y > 0 ? 13 : 10], oT = s_vo[z < 7 ? 13 : 22], oB = s_vo[z > 0 ? 13 : 4];
        const int lN = z * 64 + ((y + 1) & 7) * 8 + x0, lS = z * 64 + ((y - 1) & 7) * 8 + x0;
        const int lT = ((z + 1) & 7) * 64 + y * 8 + x0, lB = ((z - 1) & 7) * 64 + y * 8 + x0;
#pragma unroll
        for (int cpt = 0; cpt < 3; ++cpt) {
            const v2 own = ld2(vo + cpt * BS3);
            // a missing neighbour block falls back to the cell's own value (physics_utils.jl:69)
            uW[cpt] = make_float2((!VELFB || oM != MISSING) ? a.vel_in[oM + (row + xM) + cpt * BS3] : own.x, own.x);
            uE[cpt] = make_float2(own.y, (!VELFB || oP != MISSING) ? a.vel_in[oP + (row + xP) + cpt * BS3] : own.y);
            uN[cpt] = (!VELFB || oN != MISSING) ? ld2(a.vel_in + oN + lN + cpt * BS3) : own;
            uS[cpt] = (!VELFB || oS != MISSING) ? ld2(a.vel_in + oS + lS + cpt * BS3) : own;
            uT[cpt] = (!VELFB || oT != MISSING) ? ld2(a.vel_in + oT + lT + cpt * BS3) : own;
            uB[cpt] = (!VELFB || oB != MISSING) ? ld2(a.vel_in + oB + lB + cpt * BS3) : own;
        }
    }

    v2 rho = vmax(vadd(m.rho, V(1.0f)), 0.01f);     // :172
    // rho - 1 at full precision (m.rho), except in the (never observed) clamped case
    v2 drho = make_float2(rho.x > 0.01f ? m.rho.x : rho.x - 1.0f, rho.y > 0.01f ? m.rho.y : rho.y - 1.0f);
    const v2 inv_rho = vrcp(rho);
    v2 ux = vmul(m.jx, inv_rho), uy = vmul(m.jy, inv_rho), uz = vmul(m.jz, inv_rho);

    // ---- sponge (:181-199): rho, u relax towards (1, u_inlet, 0, 0); the population blend acts on the moments
    if (FULL && (bflags & BF_SPONGE)) {
        const v2 sp = ld2(a.sponge + (size_t)b * BS3 + c0);   // sp == 0 leaves everything unchanged
        const v2 om = vsub(V(1.0f), sp);
        rho = vfma(rho, om, sp);
        drho = vmul(drho, om);                       // rho' - 1 = (rho - 1)(1 - sp)
        const float u_in = k1_u_inlet(a);
        ux = vfma(ux, om, vmul(V(u_in), sp));
        uy = vmul(uy, om);
        uz = vmul(uz, om);
        if (a.sponge_blend == 1) {
            const float ui2 = u_in * u_in;
            // raw second moments of feq(1, u_inlet, 0, 0): delta/3 + u u  (cross terms vanish)
            // (shifted moments: the delta/3 parts cancel, only u_inlet^2 remains on xx)
            m.Pxx = vfma(m.Pxx, om, vmul(V(ui2), sp));
            m.Pyy = vmul(m.Pyy, om);
            m.Pzz = vmul(m.Pzz, om);
            m.Pxy = vmul(m.Pxy, om); m.Pyz = vmul(m.Pyz, om); m.Pzx = vmul(m.Pzx, om);
        }
    }

    // ---- wall-model force (:202-236)
    v2 Fx = V(0.f), Fy = V(0.f), Fz = V(0.f);
    bool has_force = false;
    if (FULL && a.wm == 1 && (bflags & BF_WALLDIST)) {
        const v2 dw = ld2(a.wall_dist + (size_t)b * BS3 + c0);
        if (dw.x > 0.0f && dw.x < 10.0f && !obsA) { float3 F = wall_force(dw.x, rho.x, ux.x, uy.x, uz.x, a.tau); Fx.x = F.x; Fy.x = F.y; Fz.x = F.z; }
        if (dw.y > 0.0f && dw.y < 10.0f && !obsB) { float3 F = wall_force(dw.y, rho.y, ux.y, uy.y, uz.y, a.tau); Fx.y = F.x; Fy.y = F.y; Fz.y = F.z; }
        has_force = true;
    }
    v2 uxe = ux, uye = uy, uze = uz;
    if (FULL && has_force) {
        const v2 hi = vmul(V(0.5f), inv_rho);        // inv_rho is the PRE-sponge 1/rho (:238, reference quirk)
        uxe = vfma(Fx, hi, ux); uye = vfma(Fy, hi, uy); uze = vfma(Fz, hi, uz);
    }

    // vel_out / rho_out (:155-158, :243-246)
    if (FULL && (obsA || obsB)) {
        st2(vout, make_float2(obsA ? 0.f : ux.x, obsB ? 0.f : ux.y));
        st2(vout + BS3, make_float2(obsA ? 0.f : uy.x, obsB ? 0.f : uy.y));
        st2(vout + 2 * BS3, make_float2(obsA ? 0.f : uz.x, obsB ? 0.f : uz.y));
        st2(rout, make_float2(obsA ? 1.f : rho.x, obsB ? 1.f : rho.y));
    } else {
        st2(vout, ux); st2(vout + BS3, uy); st2(vout + 2 * BS3, uz); st2(rout, rho);
    }

    // ---- WALE eddy viscosity (:251-300)
    v2 omega;
    {
        const v2 h = V(0.5f);
        v2 g11 = vmul(h, vsub(uE[0], uW[0])), g12 = vmul(h, vsub(uN[0], uS[0])), g13 = vmul(h, vsub(uT[0], uB[0]));
        v2 g21 = vmul(h, vsub(uE[1], uW[1])), g22 = vmul(h, vsub(uN[1], uS[1])), g23 = vmul(h, vsub(uT[1], uB[1]));
        v2 g31 = vmul(h, vsub(uE[2], uW[2])), g32 = vmul(h, vsub(uN[2], uS[2])), g33 = vmul(h, vsub(uT[2], uB[2]));
        v2 gsq11 = vfma(g13, g31, vfma(g12, g21, vmul(g11, g11)));
        v2 gsq12 = vfma(g13, g32, vfma(g12, g22, vmul(g11, g12)));
        v2 gsq13 = vfma(g13, g33, vfma(g12, g23, vmul(g11, g13)));
        v2 gsq21 = vfma(g23, g31, vfma(g22, g21, vmul(g21, g11)));
        v2 gsq22 = vfma(g23, g32, vfma(g22, g22, vmul(g21, g12)));
        v2 gsq23 = vfma(g23, g33, vfma(g22, g23, vmul(g21, g13)));
        v2 gsq31 = vfma(g33, g31, vfma(g32, g21, vmul(g31, g11)));
        v2 gsq32 = vfma(g33, g32, vfma(g32, g22, vmul(g31, g12)));
        v2 gsq33 = vfma(g33, g33, vfma(g32, g23, vmul(g31, g13)));
        v2 tr_term = vmul(vadd(vadd(gsq11, gsq22), gsq33), V(1.0f / 3.0f));
        v2 Sd11 = vsub(gsq11, tr_term), Sd22 = vsub(gsq22, tr_term), Sd33 = vsub(gsq33, tr_term);
        v2 Sd12 = vmul(h, vadd(gsq12, gsq21)), Sd13 = vmul(h, vadd(gsq13, gsq31)), Sd23 = vmul(h, vadd(gsq23, gsq32));
        v2 S12 = vmul(h, vadd(g12, g21)), S13 = vmul(h, vadd(g13, g31)), S23 = vmul(h, vadd(g23, g32));
        v2 offd = vfma(Sd23, Sd23, vfma(Sd13, Sd13, vmul(Sd12, Sd12)));
        v2 OP1 = vfma(V(2.0f), offd, vfma(Sd33, Sd33, vfma(Sd22, Sd22, vmul(Sd11, Sd11))));
        v2 offs = vfma(S23, S23, vfma(S13, S13, vmul(S12, S12)));
        v2 OP2 = vfma(V(2.0f), offs, vfma(g33, g33, vfma(g22, g22, vmul(g11, g11))));
        v2 OP1_32 = vmul(OP1, vsqrt(OP1));
        v2 OP2_52 = vmul(vmul(OP2, OP2), vsqrt(vmax(OP2, 1.0e-12f)));
        v2 denom = vfma(OP1, vsqrt(vsqrt(vmax(OP1, 1.0e-12f))), OP2_52);
        v2 q = vmul(vmul(V(a.c_wale * a.c_wale), OP1_32), vrcp(vmax(denom, 1.0e-30f)));
        float ne0 = (OP1.x > 1.0e-12f && denom.x > 1.0e-12f) ? q.x : 0.0f;
        float ne1 = (OP1.y > 1.0e-12f && denom.y > 1.0e-12f) ? q.y : 0.0f;
        v2 nu_eddy = vmax(make_float2(ne0, ne1), a.nu_bg);
        omega = vrcp(vmax(vfma(nu_eddy, V(3.0f), V(a.tau)), 0.500001f));
    }

    // ---- regularized collision (:305-354)
    const v2 usq = vfma(uze, uze, vfma(uye, uye, vmul(uxe, uxe)));
    const v2 third_rho = vmul(drho, V(1.0f / 3.0f));   // (rho - 1)/3: the shifted moments already lack delta/3
    const v2 rux = vmul(rho, uxe), ruy = vmul(rho, uye), ruz = vmul(rho, uze);
    const v2 Pi_xx = vsub(vsub(m.Pxx, third_rho), vmul(rux, uxe));
    const v2 Pi_yy = vsub(vsub(m.Pyy, third_rho), vmul(ruy, uye));
    const v2 Pi_zz = vsub(vsub(m.Pzz, third_rho), vmul(ruz, uze));
    const v2 Pi_xy2 = vmul(V(2.0f), vsub(m.Pxy, vmul(rux, uye)));
    const v2 Pi_yz2 = vmul(V(2.0f), vsub(m.Pyz, vmul(ruy, uze)));
    const v2 Pi_zx2 = vmul(V(2.0f), vsub(m.Pzx, vmul(ruz, uxe)));
    const v2 T = vmul(vadd(vadd(Pi_xx, Pi_yy), Pi_zz), V(1.0f / 3.0f));
    const v2 A = vmul(rho, vfma(V(-1.5f), usq, V(1.0f)));   // rho (1 - 1.5 u^2)
    const v2 r45 = vmul(rho, V(4.5f));
    const v2 r3 = vmul(rho, V(3.0f));
    const v2 g = vmul(vsub(V(1.0f), omega), V(4.5f));       // (1 - omega) 4.5
    // Guo force (:333-337):  w 3 [(c - u + 3 cu c) . F]  =  w 3 [ cF - uF + 3 cu cF ],  cu from u_eq, u from u
    v2 hw = V(0.f), uF = V(0.f);
    if (FULL && has_force) {
        hw = vmul(vfma(V(-0.5f), omega, V(1.0f)), V(3.0f));   // (1 - omega/2) 3
        uF = vfma(uz, Fz, vfma(uy, Fy, vmul(ux, Fx)));
    }

    // the obstacle cell of a half-solid thread already holds its bounced-back populations: store the fluid cell only
    auto store_f = [&](int k, v2 v) {
        if (!anyobs) st2(fout + k * BS3, v);
        else if (obsA) fout[k * BS3 + 1] = v.y;
        else fout[k * BS3] = v.x;
    };
#pragma unroll
    for (int k = 0; k < 13; ++k) {
        const int cx = lat_cx(k), cy = lat_cy(k), cz = lat_cz(k);
        v2 cu = V(0.f), cF = V(0.f);
        bool first = true;
        if (cx != 0) { cu = cx > 0 ? uxe : vneg(uxe); cF = cx > 0 ? Fx : vneg(Fx); first = false; }
        if (cy != 0) {
            cu = first ? (cy > 0 ? uye : vneg(uye)) : (cy > 0 ? vadd(cu, uye) : vsub(cu, uye));
            cF = first ? (cy > 0 ? Fy : vneg(Fy)) : (cy > 0 ? vadd(cF, Fy) : vsub(cF, Fy));
            first = false;
        }
        if (cz != 0) {
            cu = first ? (cz > 0 ? uze : vneg(uze)) : (cz > 0 ? vadd(cu, uze) : vsub(cu, uze));
            cF = first ? (cz > 0 ? Fz : vneg(Fz)) : (cz > 0 ? vadd(cF, Fz) : vsub(cF, Fz));
        }
        v2 R = vneg(T);                                   // Pi : Q_k = sum Pi_ab c_a c_b - tr(Pi)/3
        if (cx != 0) R = vadd(R, Pi_xx);
        if (cy != 0) R = vadd(R, Pi_yy);
        if (cz != 0) R = vadd(R, Pi_zz);
        if (cx * cy == 1) R = vadd(R, Pi_xy2); else if (cx * cy == -1) R = vsub(R, Pi_xy2);
        if (cy * cz == 1) R = vadd(R, Pi_yz2); else if (cy * cz == -1) R = vsub(R, Pi_yz2);
        if (cz * cx == 1) R = vadd(R, Pi_zx2); else if (cz * cx == -1) R = vsub(R, Pi_zx2);
        const float w = wk(k);
        v2 even = vfma(g, R, vfma(r45, vmul(cu, cu), A));   // A + 4.5 rho cu^2 + (1-omega) 4.5 Pi:Q
        v2 odd = vmul(r3, cu);
        if (FULL && has_force) {
            even = vfma(hw, vfma(vmul(V(3.0f), cu), cF, vneg(uF)), even);
            odd = vfma(hw, cF, odd);
        }
        even = vmul(even, V(w)); odd = vmul(odd, V(w));
        store_f(k, vadd(even, odd));
        store_f(26 - k, vsub(even, odd));
    }
    {
        v2 even = vfma(g, vneg(T), A);
        if (FULL && has_force) even = vfma(hw, vneg(uF), even);
        even = vmul(even, V(W0));
        store_f(13, even);
    }
}


// FULL : see file header.   VELFB : some axis neighbour may lack a velocity field (ghost block or domain face) ->
// fall back to the cell's own value (physics_utils.jl:69).
//        MISS : some neighbour block may be absent (domain face) -> k1_boundary.cuh; blocks with features but all 26
//        neighbours present (the near-body bulk) use FULL without MISS and never carry that code.
template <bool FULL, bool VELFB, bool MISS, int MINB, int NT>
__global__ void __launch_bounds__(NT, MINB * (256 / NT)) k1_fast_kernel(const __grid_constant__ K1Args a) {
    __shared__ long long s_fo[27];   // element offset of each neighbour block relative to f_in (MISSING: no block)
    __shared__ long long s_vo[27];   // ... relative to vel_in (MISSING for ghost blocks: they carry populations only)
    constexpr int PARTS = 256 / NT;  // NT = 128 / 64: a CTA takes 4 / 2 of the block's z-planes (option cta_threads)
    const int b = a.list[blockIdx.x / PARTS];
    if (threadIdx.x < 27) neighbour_offsets(a, b, threadIdx.x, (long long)b * (Q * BS3), s_fo, s_vo);
    __syncthreads();
    prefetch_block_part<NT>(a, (int)(blockIdx.x / PARTS), (int)(blockIdx.x % PARTS));
    fast_block<FULL, VELFB, MISS>(a, b, (int)threadIdx.x + (int)(blockIdx.x % PARTS) * NT, a.f_in, s_fo, s_vo);
}

// Merged launch of the plain and the domain-face class (see k1_strict.cu / abi.cu, option merge_face); register budget of the plain kernel.
__global__ void __launch_bounds__(128, 6) k1_fast_mixed_kernel(const __grid_constant__ K1Args a) {
    __shared__ long long s_fo[27], s_vo[27];
    const int b = a.list[blockIdx.x >> 1];
    bool miss = false;
    if (threadIdx.x < 27) { neighbour_offsets(a, b, threadIdx.x, (long long)b * (Q * BS3), s_fo, s_vo); miss = s_fo[threadIdx.x] == MISSING; }
    const bool face = __syncthreads_or(miss) != 0;
    const int t = (int)threadIdx.x + (int)(blockIdx.x & 1) * 128;
    if (face) fast_block<true, true, true>(a, b, t, a.f_in, s_fo, s_vo);
    else fast_block<false, false, false>(a, b, t, a.f_in, s_fo, s_vo);
}

// Persistent form for a small latency-bound class beside the plain launch (see k1_strict.cu / abi.cu, option face_persist).
template <bool FULL, bool VELFB, bool MISS>
__global__ void __launch_bounds__(64, 8) k1_fast_persist_kernel(const __grid_constant__ K1Args a) {
    __shared__ long long s_fo[27], s_vo[27];
    const int total = a.n_list * 4;
    for (int e = blockIdx.x; e < total; e += gridDim.x) {
        const int b = a.list[e >> 2];
        __syncthreads();
        if (threadIdx.x < 27) neighbour_offsets(a, b, threadIdx.x, (long long)b * (Q * BS3), s_fo, s_vo);
        __syncthreads();
        fast_block<FULL, VELFB, MISS>(a, b, (int)threadIdx.x + (e & 3) * 64, a.f_in, s_fo, s_vo);
    }
}

// TMA variant (option fast_kernel = tma): persistent CTAs, the block's own populations staged into shared memory by one
// cp.async.bulk per block, double-buffered (see k1_strict.cu for the full description).  north_star asks for this form
// ("block tiles plus halo staged into shared memory with TMA, or cp.async where measured faster"); the measured A/B against the
// direct-load kernel is in profiles/README.md.  Same fast_block body, same bits.
template <bool FULL, bool VELFB, bool MISS>
__global__ void __launch_bounds__(256, 2) k1_fast_tma_kernel(const __grid_constant__ K1Args a) {
    extern __shared__ __align__(128) float s_tile[];              // [2][TILE_FLOATS]
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ long long s_fo[2][27], s_vo[2][27];
    __shared__ int s_blk[2];          // block of this / the next iteration (-1: the list is exhausted)
    const int t = threadIdx.x;
    if (t == 0) {
        mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned long long i0 = atomicAdd(a.ticket, 1ull) - a.ticket_base;
        s_blk[0] = i0 < (unsigned long long)a.n_list ? a.list[i0] : -1;
    }
    __syncthreads();
    if (t == 0 && s_blk[0] >= 0) {
        mbar_expect_tx(&s_bar[0], TILE_BYTES);
        bulk_load(s_tile, a.f_in + (size_t)s_blk[0] * TILE_FLOATS, TILE_BYTES, &s_bar[0]);
    }
    const float* fbase;   // a.f_in with its global provenance hidden: own-block reads hit shared memory, the loads must be generic
    asm volatile("mov.u64 %0, %1;" : "=l"(fbase) : "l"(a.f_in));
    for (int it = 0;; ++it) {
        const int cur = it & 1;
        const int b = s_blk[cur];
        if (b < 0) break;
        const float* tile = s_tile + cur * TILE_FLOATS;
        if (t < 27) neighbour_offsets(a, b, t, (long long)(((long long)(uintptr_t)tile - (long long)(uintptr_t)a.f_in) >> 2), s_fo[cur], s_vo[cur]);
        if (t == 32) {   // the next ticket: blocks are handed out in list (Morton) order, whichever CTA asks first
            const unsigned long long i1 = atomicAdd(a.ticket, 1ull) - a.ticket_base;
            s_blk[cur ^ 1] = i1 < (unsigned long long)a.n_list ? a.list[i1] : -1;
        }
        // one barrier per iteration: publishes this block's tables and the next block's index, and every warp is done with the
        // OTHER stage (previous iteration) before thread 0 lets the copy engine overwrite it
        __syncthreads();
        const int bn = s_blk[cur ^ 1];
        if (bn >= 0) {
            if (t == 0) {
                mbar_expect_tx(&s_bar[cur ^ 1], TILE_BYTES);
                bulk_load(s_tile + (cur ^ 1) * TILE_FLOATS, a.f_in + (size_t)bn * TILE_FLOATS, TILE_BYTES, &s_bar[cur ^ 1]);
            } else if (t >= 32 && t < 32 + 48) {   // the next block's own velocities (6 KiB = 48 lines) into L2
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.vel_in + (size_t)bn * (3 * BS3) + (t - 32) * 32));
            }
        }
        mbar_wait(&s_bar[cur], (uint32_t)((it >> 1) & 1));
        fast_block<FULL, VELFB, MISS>(a, b, t, fbase, s_fo[cur], s_vo[cur]);
    }
}

}  // namespace k1f

template <bool FULL, bool VELFB, bool MISS, int MINB>
void launch_fast(const K1Args& a, cudaStream_t s) {
    if (a.n_list <= 0) return;
    if (a.persist_grid > 0) {
        if constexpr (FULL && MISS) { k1f::k1_fast_persist_kernel<FULL, VELFB, MISS><<<a.persist_grid, 64, 0, s>>>(a); return; }
    }
    if (a.fast_variant == 2) {
        static const cudaError_t once = cudaFuncSetAttribute(k1f::k1_fast_tma_kernel<FULL, VELFB, MISS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)k1f::TILE_BYTES);
        static const cudaError_t once2 = cudaFuncSetAttribute(k1f::k1_fast_tma_kernel<FULL, VELFB, MISS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        (void)once; (void)once2;
        const int grid = a.n_list < 2 * a.num_sms ? a.n_list : 2 * a.num_sms;
        k1f::k1_fast_tma_kernel<FULL, VELFB, MISS><<<grid, 256, 2 * k1f::TILE_BYTES, s>>>(a);
    } else if (a.cta_threads == 128) k1f::k1_fast_kernel<FULL, VELFB, MISS, MINB, 128><<<2 * a.n_list, 128, 0, s>>>(a);
    else if (a.cta_threads == 64) k1f::k1_fast_kernel<FULL, VELFB, MISS, MINB, 64><<<4 * a.n_list, 64, 0, s>>>(a);
    else k1f::k1_fast_kernel<FULL, VELFB, MISS, MINB, 256><<<a.n_list, 256, 0, s>>>(a);
}
void launch_k1_plain(const K1Args& a, cudaStream_t s) { launch_fast<false, false, false, 3>(a, s); }
void launch_k1_mixed(const K1Args& a, cudaStream_t s) { if (a.n_list > 0) k1f::k1_fast_mixed_kernel<<<2 * a.n_list, 128, 0, s>>>(a); }
void launch_k1_plain_ghost(const K1Args& a, cudaStream_t s) { launch_fast<false, true, false, 3>(a, s); }
// feature blocks: 80 registers / 3 CTAs per SM measured +12 % over 127 registers / 2 CTAs on Wing_5_deg (A/B on one box)
void launch_k1_feat(const K1Args& a, cudaStream_t s) { launch_fast<true, true, false, 3>(a, s); }
void launch_k1_full(const K1Args& a, cudaStream_t s) { launch_fast<true, true, true, 2>(a, s); }
void launch_ghost_interp(const GhostArgs& g, bool block_variant, cudaStream_t s) {
    if (g.n <= 0) return;
    if (block_variant && g.gstart) k1f::ghost_interp_block_kernel<<<g.n_ghost, 128, 0, s>>>(g);
    else k1f::ghost_interp_kernel<<<(g.n + 127) / 128, 128, 0, s>>>(g);
}

}  // namespace ludwig
