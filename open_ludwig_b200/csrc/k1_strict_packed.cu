// k1_strict_packed.cu — K1 for plain interior blocks in STRICT mode: the reference's exact FP32 operation order
// (physics_kernels.jl:62-354, fluid branch without sponge / wall force) evaluated with packed FP32x2 instructions.
//
// FADD2 rounds each half exactly like FADD, so doing two x-adjacent cells per thread changes no bit.  Additions are
// packed (__fadd2_rn; a - b as __ffma2_rn(b, -1, a): the product by -1 is exact, so the single rounding is that of
// a - b).  Multiplications are two scalar __fmul_rn: ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even
// under -fmad=false — and even folds fma(fma(a,b,-0),1,c) into fma(a,b,c) — which changes the rounding; it leaves the
// scalar .rn forms alone.  This file is compiled with -fmad=false.
// Terms the reference multiplies by a zero lattice component are skipped: x + (+-0) = x, and a sum that stays +-0
// only feeds 1 + 3cu etc. where the sign of zero cannot matter.  Division and square root are the IEEE ones.
//
// Result: bit-identical to k1_generic_strict.cu / the CPU oracle (tests/test_k1_single_level_gpu.py,
// tests/test_large_sizes_gpu.py) at ~3x its speed.  Blocks with missing neighbours, obstacles, sponge or near-wall
// cells, and refinement-interface blocks, stay on the generic strict kernel.
#include <climits>

#include "ludwig_internal.h"

namespace ludwig {
namespace k1sp {

typedef float2 v2;
__device__ __forceinline__ v2 V(float s) { return make_float2(s, s); }
__device__ __forceinline__ v2 vadd(v2 a, v2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ v2 vsub(v2 a, v2 b) { return __ffma2_rn(b, V(-1.0f), a); }   // a - b, one rounding
__device__ __forceinline__ v2 vmul(v2 a, v2 b) { return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }   // see header
__device__ __forceinline__ v2 vneg(v2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ v2 vmaxs(v2 a, float s) { return make_float2(fmaxf(a.x, s), fmaxf(a.y, s)); }
__device__ __forceinline__ v2 vsqrt(v2 a) { return make_float2(__fsqrt_rn(a.x), __fsqrt_rn(a.y)); }
__device__ __forceinline__ v2 vdiv(v2 a, v2 b) { return make_float2(__fdiv_rn(a.x, b.x), __fdiv_rn(a.y, b.y)); }
__device__ __forceinline__ void st2(float* p, v2 v) { __stcs(reinterpret_cast<float2*>(p), v); }
__device__ __forceinline__ v2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }

__host__ __device__ constexpr int d2of(int k) { return (lat_cx(k) != 0) + (lat_cy(k) != 0) + (lat_cz(k) != 0); }

// c . u in the reference's order ((cx*ux + cy*uy) + cz*uz), zero components skipped
template <int K>
__device__ __forceinline__ v2 cdot(v2 ux, v2 uy, v2 uz) {
    constexpr int cx = lat_cx(K), cy = lat_cy(K), cz = lat_cz(K);
    v2 r = V(0.f);
    bool have = false;
    if (cx != 0) { r = cx > 0 ? ux : vneg(ux); have = true; }
    if (cy != 0) { r = have ? (cy > 0 ? vadd(r, uy) : vsub(r, uy)) : (cy > 0 ? uy : vneg(uy)); have = true; }
    if (cz != 0) { r = have ? (cz > 0 ? vadd(r, uz) : vsub(r, uz)) : (cz > 0 ? uz : vneg(uz)); }
    return r;
}

template <int K>
__device__ __forceinline__ void moment_step(v2 val, v2& rho, v2& jx, v2& jy, v2& jz) {   // :144-148
    constexpr int cx = lat_cx(K), cy = lat_cy(K), cz = lat_cz(K);
    rho = K == 0 ? val : vadd(rho, val);
    if (cx == 1) jx = vadd(jx, val); else if (cx == -1) jx = vsub(jx, val);
    if (cy == 1) jy = vadd(jy, val); else if (cy == -1) jy = vsub(jy, val);
    if (cz == 1) jz = vadd(jz, val); else if (cz == -1) jz = vsub(jz, val);
}

// feq_k (:313), f_neq and the Pi accumulation (:314-321); returns feq_k (kept for the collision loop)
template <int K>
__device__ __forceinline__ v2 pi_step(v2 fk, v2 rw, v2 ux, v2 uy, v2 uz, v2 usq15, v2& Pxx, v2& Pyy, v2& Pzz, v2& Pxy, v2& Pyz, v2& Pzx) {
    constexpr int cx = lat_cx(K), cy = lat_cy(K), cz = lat_cz(K);
    v2 poly;
    if (K == 13) poly = vsub(V(1.0f), usq15);                                  // cu = 0: ((1 + 0) + 0) - 1.5 usq
    else {
        const v2 cu = cdot<K>(ux, uy, uz);
        poly = vsub(vadd(vadd(V(1.0f), vmul(V(3.0f), cu)), vmul(vmul(V(4.5f), cu), cu)), usq15);
    }
    const v2 feq = vmul(rw, poly);
    const v2 fneq = vsub(fk, feq);
    if (cx != 0) Pxx = vadd(Pxx, fneq);
    if (cy != 0) Pyy = vadd(Pyy, fneq);
    if (cz != 0) Pzz = vadd(Pzz, fneq);
    if (cx * cy == 1) Pxy = vadd(Pxy, fneq); else if (cx * cy == -1) Pxy = vsub(Pxy, fneq);
    if (cy * cz == 1) Pyz = vadd(Pyz, fneq); else if (cy * cz == -1) Pyz = vsub(Pyz, fneq);
    if (cz * cx == 1) Pzx = vadd(Pzx, fneq); else if (cz * cx == -1) Pzx = vsub(Pzx, fneq);
    return feq;
}

// f_coll_k = feq + (1 - omega) f_neq_reg  (:339-348; the force term is exactly zero in plain blocks)
template <int K>
__device__ __forceinline__ v2 collide_step(v2 feq, v2 om1, const v2 (&PQ)[3][2], v2 Pxy, v2 Pyz, v2 Pzx) {
    constexpr int cx = lat_cx(K), cy = lat_cy(K), cz = lat_cz(K);
    // Pi_xx Q_xx + Pi_yy Q_yy + Pi_zz Q_zz with Q = c^2 - 1/3: PQ[a][0] = Pi_aa * (1 - 1/3), PQ[a][1] = Pi_aa * (0 - 1/3)
    const v2 diag = vadd(vadd(PQ[0][cx != 0 ? 0 : 1], PQ[1][cy != 0 ? 0 : 1]), PQ[2][cz != 0 ? 0 : 1]);
    // Pi_xy cx cy + Pi_yz cy cz + Pi_zx cz cx
    v2 off = V(0.f);
    bool have = false;
    if (cx * cy != 0) { off = cx * cy > 0 ? Pxy : vneg(Pxy); have = true; }
    if (cy * cz != 0) { off = have ? (cy * cz > 0 ? vadd(off, Pyz) : vsub(off, Pyz)) : (cy * cz > 0 ? Pyz : vneg(Pyz)); have = true; }
    if (cz * cx != 0) { off = have ? (cz * cx > 0 ? vadd(off, Pzx) : vsub(off, Pzx)) : (cz * cx > 0 ? Pzx : vneg(Pzx)); have = true; }
    const v2 inner = have ? vadd(diag, vmul(V(2.0f), off)) : diag;     // + 2 * 0 changes nothing
    const v2 fnr = vmul(V(lat_w(K) * 4.5f), inner);
    return vadd(feq, vmul(om1, fnr));
}

template <int K, int KEND>
struct Unroll {
    template <typename F> __device__ __forceinline__ static void run(F&& f) { f.template operator()<K>(); Unroll<K + 1, KEND>::run(f); }
};
template <int KEND>
struct Unroll<KEND, KEND> {
    template <typename F> __device__ __forceinline__ static void run(F&&) {}
};

__global__ void __launch_bounds__(256, 2) k1_strict_packed_kernel(const K1Args a) {
    __shared__ long long s_fo[27];
    __shared__ long long s_vo[27];
    const int b = a.list[blockIdx.x];
    const int t = threadIdx.x;
    if (t < 27) {
        const int nbi = a.nbr[(size_t)b * 27 + t];     // plain blocks: all 26 neighbours are local real blocks
        s_fo[t] = (long long)nbi * (Q * BS3);
        s_vo[t] = (long long)nbi * (3 * BS3);
    }
    __syncthreads();
    const int p = t & 3, y = (t >> 2) & 7, z = t >> 5;
    const int x0 = 2 * p, c0 = 2 * t;
    int yoff[3], ydir[3], zoff[3], zdir[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        int ys = y - (j - 1), zs = z - (j - 1);
        yoff[j] = (ys & 7) * 8; ydir[j] = (ys < 0 ? 0 : (ys > 7 ? 2 : 1)) * 3;
        zoff[j] = (zs & 7) * 64; zdir[j] = (zs < 0 ? 0 : (zs > 7 ? 2 : 1)) * 9;
    }
    const int dM = p > 0 ? 1 : 0, xM = p > 0 ? x0 - 1 : 7;
    const int dP = p < 3 ? 1 : 2, xP = p < 3 ? x0 + 2 : 0;

    // pull-stream in k order (combo (jy,jz) yields k0-1, k0, k0+1 = three consecutive directions)
    v2 f[27];
    v2 rho = V(0.f), jx = V(0.f), jy = V(0.f), jz = V(0.f);
#pragma unroll
    for (int jzc = 0; jzc < 3; ++jzc) {
#pragma unroll
        for (int jyc = 0; jyc < 3; ++jyc) {
            const int loc = zoff[jzc] + yoff[jyc], dir = zdir[jzc] + ydir[jyc];
            const float* __restrict__ P0 = a.f_in + s_fo[dir + 1] + (loc + x0);
            const float* __restrict__ PM = a.f_in + s_fo[dir + dM] + (loc + xM);
            const float* __restrict__ PP = a.f_in + s_fo[dir + dP] + (loc + xP);
            const int k0 = 1 + 3 * jyc + 9 * jzc;
            f[k0 - 1] = make_float2(P0[(k0 - 1) * BS3 + 1], PP[(k0 - 1) * BS3]);   // cx=-1: sources x0+1, x0+2
            f[k0] = ld2(P0 + k0 * BS3);
            f[k0 + 1] = make_float2(PM[(k0 + 1) * BS3], P0[(k0 + 1) * BS3]);       // cx=+1: sources x0-1, x0
        }
    }
    Unroll<0, 27>::run([&]<int K>() { moment_step<K>(f[K], rho, jx, jy, jz); });

    // six axis neighbours' previous-step velocities (physics_utils.jl:72-78)
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
        for (int c = 0; c < 3; ++c) {
            const v2 own = ld2(vo + c * BS3);
            uW[c] = make_float2(vM[c * BS3], own.x);
            uE[c] = make_float2(own.y, vP[c * BS3]);
            uN[c] = ld2(vN + c * BS3); uS[c] = ld2(vS + c * BS3); uT[c] = ld2(vT + c * BS3); uB[c] = ld2(vB + c * BS3);
        }
    }

    rho = vmaxs(rho, 0.01f);                                          // :172
    const v2 inv_rho = vdiv(V(1.0f), rho);
    const v2 ux = vmul(jx, inv_rho), uy = vmul(jy, inv_rho), uz = vmul(jz, inv_rho);
    const v2 usq = vadd(vadd(vmul(ux, ux), vmul(uy, uy)), vmul(uz, uz));   // u_eq = u + 0.5*0*inv_rho = u

    float* __restrict__ fout = a.f_out + (size_t)b * (Q * BS3) + c0;
    {
        float* __restrict__ vout = a.vel_out + (size_t)b * (3 * BS3) + c0;
        st2(vout, ux); st2(vout + BS3, uy); st2(vout + 2 * BS3, uz);
        st2(a.rho_out + (size_t)b * BS3 + c0, rho);
    }

    // WALE (:251-300) in the reference's expression order
    v2 omega;
    {
        const v2 h = V(0.5f);
        const v2 g11 = vmul(h, vsub(uE[0], uW[0])), g12 = vmul(h, vsub(uN[0], uS[0])), g13 = vmul(h, vsub(uT[0], uB[0]));
        const v2 g21 = vmul(h, vsub(uE[1], uW[1])), g22 = vmul(h, vsub(uN[1], uS[1])), g23 = vmul(h, vsub(uT[1], uB[1]));
        const v2 g31 = vmul(h, vsub(uE[2], uW[2])), g32 = vmul(h, vsub(uN[2], uS[2])), g33 = vmul(h, vsub(uT[2], uB[2]));
#define DOT3(a1, b1, a2, b2, a3, b3) vadd(vadd(vmul(a1, b1), vmul(a2, b2)), vmul(a3, b3))
        const v2 gsq11 = DOT3(g11, g11, g12, g21, g13, g31), gsq12 = DOT3(g11, g12, g12, g22, g13, g32), gsq13 = DOT3(g11, g13, g12, g23, g13, g33);
        const v2 gsq21 = DOT3(g21, g11, g22, g21, g23, g31), gsq22 = DOT3(g21, g12, g22, g22, g23, g32), gsq23 = DOT3(g21, g13, g22, g23, g23, g33);
        const v2 gsq31 = DOT3(g31, g11, g32, g21, g33, g31), gsq32 = DOT3(g31, g12, g32, g22, g33, g32), gsq33 = DOT3(g31, g13, g32, g23, g33, g33);
        const v2 tr_gsq = vadd(vadd(gsq11, gsq22), gsq33);
        const v2 tr_term = vdiv(tr_gsq, V(3.0f));
        const v2 Sd11 = vsub(gsq11, tr_term), Sd22 = vsub(gsq22, tr_term), Sd33 = vsub(gsq33, tr_term);
        const v2 Sd12 = vmul(h, vadd(gsq12, gsq21)), Sd13 = vmul(h, vadd(gsq13, gsq31)), Sd23 = vmul(h, vadd(gsq23, gsq32));
        const v2 S12 = vmul(h, vadd(g12, g21)), S13 = vmul(h, vadd(g13, g31)), S23 = vmul(h, vadd(g23, g32));
        const v2 OP1 = vadd(DOT3(Sd11, Sd11, Sd22, Sd22, Sd33, Sd33), vmul(V(2.0f), DOT3(Sd12, Sd12, Sd13, Sd13, Sd23, Sd23)));
        const v2 OP2 = vadd(DOT3(g11, g11, g22, g22, g33, g33), vmul(V(2.0f), DOT3(S12, S12, S13, S13, S23, S23)));
#undef DOT3
        const v2 OP1_32 = vmul(OP1, vsqrt(OP1));
        const v2 OP2_52 = vmul(vmul(OP2, OP2), vsqrt(vmaxs(OP2, 1.0e-12f)));
        const v2 denom = vadd(OP2_52, vmul(OP1, vsqrt(vsqrt(vmaxs(OP1, 1.0e-12f)))));
        const float cw2 = __fmul_rn(a.c_wale, a.c_wale);
        float ne0 = 0.0f, ne1 = 0.0f;
        if (OP1.x > 1.0e-12f && denom.x > 1.0e-12f) ne0 = __fdiv_rn(__fmul_rn(cw2, OP1_32.x), denom.x);
        if (OP1.y > 1.0e-12f && denom.y > 1.0e-12f) ne1 = __fdiv_rn(__fmul_rn(cw2, OP1_32.y), denom.y);
        const v2 nu_eddy = vmaxs(make_float2(ne0, ne1), a.nu_bg);
        const v2 tau_turb = vadd(V(a.tau), vmul(nu_eddy, V(3.0f)));
        omega = vdiv(V(1.0f), vmaxs(tau_turb, 0.500001f));
    }

    // Pi loop (:308-322): f[k] is replaced by feq_k
    const v2 usq15 = vmul(V(1.5f), usq);
    const v2 rw0 = vmul(rho, V(lat_w(13))), rw1 = vmul(rho, V(lat_w(12))), rw2 = vmul(rho, V(lat_w(9))), rw3 = vmul(rho, V(lat_w(0)));
    v2 Pxx = V(0.f), Pyy = V(0.f), Pzz = V(0.f), Pxy = V(0.f), Pyz = V(0.f), Pzx = V(0.f);
    Unroll<0, 27>::run([&]<int K>() {
        const v2 rw = d2of(K) == 0 ? rw0 : d2of(K) == 1 ? rw1 : d2of(K) == 2 ? rw2 : rw3;
        f[K] = pi_step<K>(f[K], rw, ux, uy, uz, usq15, Pxx, Pyy, Pzz, Pxy, Pyz, Pzx);
    });

    // collision loop (:324-354)
    const float cs2 = 1.0f / 3.0f;
    const float qa = 1.0f - cs2, qb = 0.0f - cs2;                     // Q = c*c - CS2 for |c| = 1 and c = 0
    const v2 PQ[3][2] = {{vmul(Pxx, V(qa)), vmul(Pxx, V(qb))}, {vmul(Pyy, V(qa)), vmul(Pyy, V(qb))}, {vmul(Pzz, V(qa)), vmul(Pzz, V(qb))}};
    const v2 om1 = vsub(V(1.0f), omega);
    Unroll<0, 27>::run([&]<int K>() { st2(fout + K * BS3, collide_step<K>(f[K], om1, PQ, Pxy, Pyz, Pzx)); });
}

}  // namespace k1sp

void launch_k1_strict_packed(const K1Args& a, cudaStream_t s) {
    if (a.n_list <= 0) return;
    k1sp::k1_strict_packed_kernel<<<a.n_list, 256, 0, s>>>(a);
}

}  // namespace ludwig
