// k1_interior.cu — K1 for "plain interior" blocks: all 26 neighbour blocks exist and the block has no
// obstacle, sponge or near-wall cell.  On a production case that is the bulk of every level; on the
// synthetic 512^3 box it is 97 % of the blocks.  Everything else goes to k1_generic.cuh.
//
// Same physics as stream_collide_kernel_v2! (physics_kernels.jl:9-358) restricted to the fluid branch
// without sponge / wall force (:172-176, :238-354), with the sums regrouped (fast mode, FMA on):
//   * opposite directions are paired:  s_k = f_k + f_(26-k),  d_k = f_k - f_(26-k)  (k = 0..12), so that
//     rho, j and the raw second moments need ~110 packed adds instead of 27 x 10 multiply-adds;
//   * Pi_ab = sum_k (f_k - feq_k) c_a c_b  is evaluated as  sum_k f_k c_a c_b - rho (delta_ab/3 + u_a u_b)
//     (exact identity for the second-order equilibrium on D3Q27);
//   * feq_k and feq_(26-k) share their even part, f_neq_reg is even.
//
// Blackwell specifics: one thread owns TWO x-adjacent cells and all arithmetic is packed FP32x2
// (FADD2 / FMUL2 / FFMA2, sm_100+), which halves the issue slots of the ~500-flop collision; populations of
// the cx = 0 directions and all stores are 64-bit accesses; a warp is one z-plane of the block (64 cells),
// one CTA (256 threads) is one 8^3 block, CTAs walk the blocks in Morton order so halo sectors hit L2.
#include "ludwig_internal.h"

namespace ludwig {
namespace {

typedef float2 v2;
__device__ __forceinline__ v2 V(float s) { return make_float2(s, s); }
__device__ __forceinline__ v2 vadd(v2 a, v2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ v2 vsub(v2 a, v2 b) { return __ffma2_rn(b, V(-1.0f), a); }
__device__ __forceinline__ v2 vmul(v2 a, v2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ v2 vfma(v2 a, v2 b, v2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ v2 vmax(v2 a, float s) { return make_float2(fmaxf(a.x, s), fmaxf(a.y, s)); }
// MUFU approximations (max rel. error 2^-23 / 2^-22): one instruction each, no slow-path call
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ v2 vrcp(v2 a) { return make_float2(rcp_approx(a.x), rcp_approx(a.y)); }
__device__ __forceinline__ v2 vsqrt(v2 a) { return make_float2(sqrt_approx(a.x), sqrt_approx(a.y)); }

constexpr float W0 = 8.0f / 27.0f, W1 = 2.0f / 27.0f, W2 = 1.0f / 54.0f, W3 = 1.0f / 216.0f;
__host__ __device__ constexpr float wk(int k) {
    return (lat_cx(k) != 0) + (lat_cy(k) != 0) + (lat_cz(k) != 0) == 0   ? W0
           : (lat_cx(k) != 0) + (lat_cy(k) != 0) + (lat_cz(k) != 0) == 1 ? W1
           : (lat_cx(k) != 0) + (lat_cy(k) != 0) + (lat_cz(k) != 0) == 2 ? W2
                                                                         : W3;
}

__global__ void __launch_bounds__(256, 2) k1_plain_kernel(const K1Args a) {
    __shared__ long long s_fo[27];   // element offset of each neighbour block in f_in
    __shared__ long long s_vo[27];   // ... in vel_in
    const int b = a.list[blockIdx.x];
    const int t = threadIdx.x;
    if (t < 27) {
        int nbi = a.nbr[(size_t)b * 27 + t];
        s_fo[t] = (long long)nbi * (Q * BS3);
        s_vo[t] = (long long)nbi * (3 * BS3);
    }
    __syncthreads();

    const int p = t & 3, y = (t >> 2) & 7, z = t >> 5;
    const int x0 = 2 * p;
    const int c0 = 2 * t;   // z*64 + y*8 + x0

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

    // ---- pull-stream (physics_kernels.jl:62-149, in-block and neighbour-block branches only)
    v2 f[27];
#pragma unroll
    for (int jz = 0; jz < 3; ++jz) {
#pragma unroll
        for (int jy = 0; jy < 3; ++jy) {
            const int loc = zoff[jz] + yoff[jy];
            const int dir = zdir[jz] + ydir[jy];
            const float* __restrict__ P0 = a.f_in + s_fo[dir + 1] + (loc + x0);
            const float* __restrict__ PM = a.f_in + s_fo[dir + dM] + (loc + xM);
            const float* __restrict__ PP = a.f_in + s_fo[dir + dP] + (loc + xP);
            const int k0 = 1 + 3 * jy + 9 * jz, kp = k0 + 1, km = k0 - 1;
            f[k0] = *reinterpret_cast<const float2*>(P0 + k0 * BS3);
            f[kp] = make_float2(PM[kp * BS3], P0[kp * BS3]);       // cx=+1: sources x0-1, x0
            f[km] = make_float2(P0[km * BS3 + 1], PP[km * BS3]);   // cx=-1: sources x0+1, x0+2
        }
    }

    // ---- previous-step velocities of the six axis neighbours (physics_utils.jl:45-83; all blocks exist)
    v2 uE[3], uW[3], uN[3], uS[3], uT[3], uB[3];
    {
        const int row = z * 64 + y * 8;
        const float* __restrict__ vo = a.vel_in + s_vo[13] + c0;
        const float* __restrict__ vM = a.vel_in + s_vo[12 + dM] + (row + xM);
        const float* __restrict__ vP = a.vel_in + s_vo[13 + (dP - 1)] + (row + xP);
        const float* __restrict__ vN = a.vel_in + s_vo[y < 7 ? 13 : 16] + (z * 64 + ((y + 1) & 7) * 8 + x0);
        const float* __restrict__ vS = a.vel_in + s_vo[y > 0 ? 13 : 10] + (z * 64 + ((y - 1) & 7) * 8 + x0);
        const float* __restrict__ vT = a.vel_in + s_vo[z < 7 ? 13 : 22] + (((z + 1) & 7) * 64 + y * 8 + x0);
        const float* __restrict__ vB = a.vel_in + s_vo[z > 0 ? 13 : 4] + (((z - 1) & 7) * 64 + y * 8 + x0);
#pragma unroll
        for (int cpt = 0; cpt < 3; ++cpt) {
            v2 own = *reinterpret_cast<const float2*>(vo + cpt * BS3);
            uW[cpt] = make_float2(vM[cpt * BS3], own.x);
            uE[cpt] = make_float2(own.y, vP[cpt * BS3]);
            uN[cpt] = *reinterpret_cast<const float2*>(vN + cpt * BS3);
            uS[cpt] = *reinterpret_cast<const float2*>(vS + cpt * BS3);
            uT[cpt] = *reinterpret_cast<const float2*>(vT + cpt * BS3);
            uB[cpt] = *reinterpret_cast<const float2*>(vB + cpt * BS3);
        }
    }

    // ---- moments from direction pairs
    v2 s[13], d[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) { s[k] = vadd(f[k], f[26 - k]); d[k] = vsub(f[k], f[26 - k]); }
    v2 rho = f[13];
    v2 jx = V(0.f), jy = V(0.f), jz = V(0.f);
    v2 Pxx = V(0.f), Pyy = V(0.f), Pzz = V(0.f), Pxy = V(0.f), Pyz = V(0.f), Pzx = V(0.f);
#pragma unroll
    for (int k = 0; k < 13; ++k) {
        const int cx = lat_cx(k), cy = lat_cy(k), cz = lat_cz(k);
        rho = vadd(rho, s[k]);
        if (cx == 1) jx = vadd(jx, d[k]); else if (cx == -1) jx = vsub(jx, d[k]);
        if (cy == 1) jy = vadd(jy, d[k]); else if (cy == -1) jy = vsub(jy, d[k]);
        if (cz == 1) jz = vadd(jz, d[k]); else if (cz == -1) jz = vsub(jz, d[k]);
        if (cx != 0) Pxx = vadd(Pxx, s[k]);
        if (cy != 0) Pyy = vadd(Pyy, s[k]);
        if (cz != 0) Pzz = vadd(Pzz, s[k]);
        if (cx * cy == 1) Pxy = vadd(Pxy, s[k]); else if (cx * cy == -1) Pxy = vsub(Pxy, s[k]);
        if (cy * cz == 1) Pyz = vadd(Pyz, s[k]); else if (cy * cz == -1) Pyz = vsub(Pyz, s[k]);
        if (cz * cx == 1) Pzx = vadd(Pzx, s[k]); else if (cz * cx == -1) Pzx = vsub(Pzx, s[k]);
    }

    rho = vmax(rho, 0.01f);                       // :172
    const v2 inv_rho = vrcp(rho);
    const v2 ux = vmul(jx, inv_rho), uy = vmul(jy, inv_rho), uz = vmul(jz, inv_rho);

    // vel_out / rho_out (:243-246)
    {
        float* __restrict__ vout = a.vel_out + (size_t)b * (3 * BS3) + c0;
        *reinterpret_cast<float2*>(vout) = ux;
        *reinterpret_cast<float2*>(vout + BS3) = uy;
        *reinterpret_cast<float2*>(vout + 2 * BS3) = uz;
        *reinterpret_cast<float2*>(a.rho_out + (size_t)b * BS3 + c0) = rho;
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
        v2 sq1 = vsqrt(OP1);
        v2 OP1_32 = vmul(OP1, sq1);
        v2 OP2_52 = vmul(vmul(OP2, OP2), vsqrt(vmax(OP2, 1.0e-12f)));
        v2 denom = vfma(OP1, vsqrt(vsqrt(vmax(OP1, 1.0e-12f))), OP2_52);
        v2 num = vmul(V(a.c_wale * a.c_wale), OP1_32);
        v2 q = vmul(num, vrcp(vmax(denom, 1.0e-30f)));
        float ne0 = (OP1.x > 1.0e-12f && denom.x > 1.0e-12f) ? q.x : 0.0f;
        float ne1 = (OP1.y > 1.0e-12f && denom.y > 1.0e-12f) ? q.y : 0.0f;
        v2 nu_eddy = vmax(make_float2(ne0, ne1), a.nu_bg);
        v2 tau_turb = vfma(nu_eddy, V(3.0f), V(a.tau));
        omega = vrcp(vmax(tau_turb, 0.500001f));
    }

    // ---- regularized collision (:305-354 with F_wall = 0, u_eq = u)
    const v2 usq = vfma(uz, uz, vfma(uy, uy, vmul(ux, ux)));
    const v2 third_rho = vmul(rho, V(1.0f / 3.0f));
    // Pi = raw second moment - rho (delta/3 + u u)
    const v2 rux = vmul(rho, ux), ruy = vmul(rho, uy), ruz = vmul(rho, uz);
    const v2 Pi_xx = vsub(vsub(Pxx, third_rho), vmul(rux, ux));
    const v2 Pi_yy = vsub(vsub(Pyy, third_rho), vmul(ruy, uy));
    const v2 Pi_zz = vsub(vsub(Pzz, third_rho), vmul(ruz, uz));
    const v2 Pi_xy2 = vmul(V(2.0f), vsub(Pxy, vmul(rux, uy)));
    const v2 Pi_yz2 = vmul(V(2.0f), vsub(Pyz, vmul(ruy, uz)));
    const v2 Pi_zx2 = vmul(V(2.0f), vsub(Pzx, vmul(ruz, ux)));
    const v2 T = vmul(vadd(vadd(Pi_xx, Pi_yy), Pi_zz), V(1.0f / 3.0f));

    const v2 A = vmul(rho, vfma(V(-1.5f), usq, V(1.0f)));   // rho (1 - 1.5 u^2)
    const v2 r45 = vmul(rho, V(4.5f));
    const v2 r3 = vmul(rho, V(3.0f));
    const v2 g = vmul(vsub(V(1.0f), omega), V(4.5f));       // (1 - omega) * 4.5

    float* __restrict__ fout = a.f_out + (size_t)b * (Q * BS3) + c0;
#pragma unroll
    for (int k = 0; k < 13; ++k) {
        const int cx = lat_cx(k), cy = lat_cy(k), cz = lat_cz(k);
        v2 cu = V(0.f);
        bool first = true;
        if (cx != 0) { cu = cx > 0 ? ux : vsub(V(0.f), ux); first = false; }
        if (cy != 0) { cu = first ? (cy > 0 ? uy : vsub(V(0.f), uy)) : (cy > 0 ? vadd(cu, uy) : vsub(cu, uy)); first = false; }
        if (cz != 0) { cu = first ? (cz > 0 ? uz : vsub(V(0.f), uz)) : (cz > 0 ? vadd(cu, uz) : vsub(cu, uz)); }
        // Pi : Q_k  =  sum Pi_ab c_a c_b - tr(Pi)/3
        v2 R = vsub(V(0.f), T);
        if (cx != 0) R = vadd(R, Pi_xx);
        if (cy != 0) R = vadd(R, Pi_yy);
        if (cz != 0) R = vadd(R, Pi_zz);
        if (cx * cy == 1) R = vadd(R, Pi_xy2); else if (cx * cy == -1) R = vsub(R, Pi_xy2);
        if (cy * cz == 1) R = vadd(R, Pi_yz2); else if (cy * cz == -1) R = vsub(R, Pi_yz2);
        if (cz * cx == 1) R = vadd(R, Pi_zx2); else if (cz * cx == -1) R = vsub(R, Pi_zx2);
        const float w = wk(k);
        v2 even = vfma(g, R, vfma(r45, vmul(cu, cu), A));   // A + 4.5 rho cu^2 + (1-omega) 4.5 Pi:Q
        even = vmul(even, V(w));
        v2 odd = vmul(vmul(r3, V(w)), cu);
        *reinterpret_cast<float2*>(fout + k * BS3) = vadd(even, odd);
        *reinterpret_cast<float2*>(fout + (26 - k) * BS3) = vsub(even, odd);
    }
    {
        v2 R = vsub(V(0.f), T);
        v2 even = vmul(vfma(g, R, A), V(W0));
        *reinterpret_cast<float2*>(fout + 13 * BS3) = even;
    }
}

}  // namespace

void launch_k1_interior(const K1Args& a, cudaStream_t s) {
    if (a.n_list <= 0) return;
    k1_plain_kernel<<<a.n_list, 256, 0, s>>>(a);
}

}  // namespace ludwig
