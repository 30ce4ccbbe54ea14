// k1_strict.cu — K1 in STRICT mode: the reference's exact FP32 operation order (physics_kernels.jl:62-354) on the
// B200-native kernel structure of k1_fast.cu (one CTA per 8^3 block, a warp per z-plane, two x-adjacent cells per
// thread, packed FP32x2 arithmetic, ghost blocks + pre-pass for refinement interfaces, remote neighbour blocks through
// peer offsets).  Every class of block runs here: plain, plain + ghost neighbours, feature (obstacle / sponge / wall
// model) and domain-face blocks.  Results are BIT-IDENTICAL to the CPU oracle (tests/test_k1_*_gpu.py,
// tests/test_large_sizes_gpu.py); wall model included (Float32 power / logarithm evaluated in Float64 as Julia does, k1_boundary.cuh).
//
// How the packed arithmetic keeps the reference's roundings (this file is compiled with -fmad=false):
//   * FADD2 rounds each half exactly like FADD;  a - b  is  FFMA2(b, -1, a): the product by -1 is exact, one rounding.
//   * a * b  is  FFMA2(a, b, nz)  with nz = -0.0f passed as a KERNEL ARGUMENT: round(a b + (-0)) = round(a b), and because
//     ptxas cannot see the value of nz it cannot contract the product into the next addition.  (ptxas 12.9 contracts
//     mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under -fmad=false, and folds a CONSTANT -0 addend away.)
//   * terms the reference multiplies by a zero lattice component are skipped: x + (+-0) = x, and a sum that stays +-0 only
//     feeds 1 + 3 cu etc., where the sign of zero cannot matter.  Division and square root are the IEEE ones.
//   * the 27 pulled populations must be kept between the pull, the Pi loop (which needs f_k - feq_k in k order) and the
//     collision loop — the price of the reference's operation order (the fast build streams them into 10 moments): in
//     registers or in a shared-memory stash (FStore below, option strict_stash).
#include <climits>

#include "ludwig_internal.h"

namespace ludwig {
namespace k1s {

#include "k1_boundary.cuh"
#include "k1_ghost.cuh"
#include "k1_common.cuh"

typedef float2 v2;
__device__ __forceinline__ v2 V(float s) { return make_float2(s, s); }
__device__ __forceinline__ v2 vadd(v2 a, v2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ v2 vsub(v2 a, v2 b) { return __ffma2_rn(b, V(-1.0f), a); }   // a - b, one rounding
__device__ __forceinline__ v2 vneg(v2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ v2 vmaxs(v2 a, float s) { return make_float2(fmaxf(a.x, s), fmaxf(a.y, s)); }
__device__ __forceinline__ v2 vsqrt(v2 a) { return make_float2(__fsqrt_rn(a.x), __fsqrt_rn(a.y)); }
__device__ __forceinline__ v2 vdiv(v2 a, v2 b) { return make_float2(__fdiv_rn(a.x, b.x), __fdiv_rn(a.y, b.y)); }
__device__ __forceinline__ void st2(float* p, v2 v) { __stcs(reinterpret_cast<float2*>(p), v); }
__device__ __forceinline__ v2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
// packed product with the rounding of a * b (see the file header); NZ = (-0, -0) from the kernel arguments
#define VMUL(a, b) __ffma2_rn((a), (b), NZ)

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

template <int K, int KEND>
struct Unroll {
    template <typename F> __device__ __forceinline__ static void run(F&& f) { f.template operator()<K>(); Unroll<K + 1, KEND>::run(f); }
};
template <int KEND>
struct Unroll<KEND, KEND> {
    template <typename F> __device__ __forceinline__ static void run(F&&) {}
};

// Wall-model force of one cell in the reference's operation order (physics_kernels.jl:206-236); scalar, near-wall cells only.
// c166 = (2 * 8.3)^(-1/7), evaluated ONCE per context by wall_model_constant_kernel with the same pow32 (a third of this function's cost).
__device__ __noinline__ float3 wall_force(float dist_wall, float rho, float ux, float uy, float uz, float tau, float c166) {
    float3 F = make_float3(0.f, 0.f, 0.f);
    if (dist_wall > 0.0f && dist_wall < 10.0f) {
        float u_mag = sqrtf(ux * ux + uy * uy + uz * uz);
        float nu_visc = (tau - 0.5f) / 3.0f;
        if (u_mag > 1.0e-6f && nu_visc > 1.0e-10f) {
            float u_tau = u_mag * pow32(nu_visc / (dist_wall * u_mag + 1.0e-10f), 1.0f / 7.0f) * c166;
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
                F.x = -force_mag * ux / u_mag;
                F.y = -force_mag * uy / u_mag;
                F.z = -force_mag * uz / u_mag;
            }
        }
    }
    return F;
}


// Where the 27 pulled populations (later: the 27 equilibria) of a thread's two cells live between the pull, the Pi loop and the
// collision loop.  REG: 54 registers (128 registers per thread, 2 CTAs per SM).  STASH: 54 KiB of dynamic shared memory per CTA,
// slot [k][thread] (consecutive threads -> consecutive 8-byte words: conflict-free 64-bit accesses), ~80 registers, 3 CTAs per
// SM — the loads of one more block are in flight while two others compute.  Same values, same operation order, same bits.
template <bool STASH> struct FStore;
template <> struct FStore<false> {
    v2 r[27];
    __device__ __forceinline__ FStore(float2*) {}
    __device__ __forceinline__ v2 get(int k) const { return r[k]; }
    __device__ __forceinline__ void set(int k, v2 v) { r[k] = v; }
};
template <> struct FStore<true> {
    float2* s;   // already offset by the thread index
    __device__ __forceinline__ FStore(float2* base) : s(base + threadIdx.x) {}   // (the stash variant runs whole-block CTAs)
    __device__ __forceinline__ v2 get(int k) const { return s[k * 256]; }
    __device__ __forceinline__ void set(int k, v2 v) { s[k * 256] = v; }
};

// FULL : obstacle / sponge / wall-model handling (per-block flag bits gate each feature uniformly)
// VELFB: some axis neighbour may lack a velocity field (ghost block or domain face) -> the cell's own value (physics_utils.jl:69)
// MISS : some neighbour block may be absent (domain face) -> k1_boundary.cuh
// One 8^3 block: 256 threads, two x-adjacent cells per thread.  fbase + s_fo[d] is neighbour block d's populations (fbase is
// a.f_in, or — TMA variant — the same address with its global provenance hidden, because s_fo[13] then points into shared memory
// and the loads must be generic).
// XONLY (with MISS): the only neighbours the block lacks lie beyond the domain's inlet / outlet plane and it has no feature (the host
// checks both when it builds the work lists): the closed-form x-face populations are all that is needed - no out-of-line boundary code,
// the register budget of the plain kernel.
template <bool FULL, bool VELFB, bool MISS, bool STASH, bool WALE_FIRST, bool MISS_UNIFORM = true, bool XONLY = false>
__device__ __forceinline__ void strict_block(const K1Args& a, const int b, const int t, const float* __restrict__ fbase, const long long* s_fo,
                                             const long long* s_vo, float2* s_stash) {
    const v2 NZ = V(a.negzero);

    const int p = t & 3, y = (t >> 2) & 7, z = t >> 5;
    const int x0 = 2 * p, c0 = 2 * t;
    uint32_t bflags = BF_INTERIOR;
    int gx = 0, gy = 0, gz = 0;   // 1-based global coords of cell A (MISS only)
    if (FULL || MISS) {
        const int4 bc = *reinterpret_cast<const int4*>(a.bcoord + (size_t)b * 4);
        if (FULL) bflags = (uint32_t)bc.w;
        if (MISS) { gx = bc.x * BS + x0 + 1; gy = bc.y * BS + y + 1; gz = bc.z * BS + z + 1; }
    }
    // domain x faces in closed form (cell A on the inlet plane / cell B on the outlet plane; an inlet source wins over y / z faces
    // and over the outlet test exactly as in pull_missing: physics_kernels.jl:99-113)
    const bool at_inlet = MISS && gx == 1, at_outlet = MISS && gx + 1 == a.nxg && a.nxg > 1;
    XFace xf{0.f, 0.f};
    if (MISS && (at_inlet || at_outlet)) xf = x_face_equilibria(a, gy, gz);
    const float* __restrict__ fin_own = fbase + s_fo[13] + c0;

    int yoff[3], ydir[3], zoff[3], zdir[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        int ys = y - (j - 1), zs = z - (j - 1);
        yoff[j] = (ys & 7) * 8; ydir[j] = (ys < 0 ? 0 : (ys > 7 ? 2 : 1)) * 3;
        zoff[j] = (zs & 7) * 64; zdir[j] = (zs < 0 ? 0 : (zs > 7 ? 2 : 1)) * 9;
    }
    const int dM = p > 0 ? 1 : 0, xM = p > 0 ? x0 - 1 : 7;
    const int dP = p < 3 ? 1 : 2, xP = p < 3 ? x0 + 2 : 0;

    // The relaxation rate depends only on the PREVIOUS step's velocities of the six axis neighbours (physics_utils.jl:45-83), not on
    // the populations.  WALE_FIRST: it is computed before the pull, so that the 18 neighbour velocities (36 registers) are dead before
    // the 27 pulled populations (54 registers) become live — the 96- / 80-register forms.  Otherwise (128 registers) the velocity
    // loads follow the pull and the rate is computed just before the Pi loop, all loads of the thread in flight together.
    v2 uE[3], uW[3], uN[3], uS[3], uT[3], uB[3];
    v2 omega;
    auto load_neighbour_velocities = [&]() {
    {
            const int row = z * 64 + y * 8;
            const float* __restrict__ vo = a.vel_in + s_vo[13] + c0;
            const long long oM = s_vo[12 + dM], oP = s_vo[13 + (dP - 1)];
            const long long oN = s_vo[y < 7 ? 13 : 16], oS = s_vo[y > 0 ? 13 : 10], oT = s_vo[z < 7 ? 13 : 22], oB = s_vo[z > 0 ? 13 : 4];
            const int lN = z * 64 + ((y + 1) & 7) * 8 + x0, lS = z * 64 + ((y - 1) & 7) * 8 + x0;
            const int lT = ((z + 1) & 7) * 64 + y * 8 + x0, lB = ((z - 1) & 7) * 64 + y * 8 + x0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const v2 own = ld2(vo + c * BS3);
                uW[c] = make_float2((!VELFB || oM != MISSING) ? a.vel_in[oM + (row + xM) + c * BS3] : own.x, own.x);
                uE[c] = make_float2(own.y, (!VELFB || oP != MISSING) ? a.vel_in[oP + (row + xP) + c * BS3] : own.y);
                uN[c] = (!VELFB || oN != MISSING) ? ld2(a.vel_in + oN + lN + c * BS3) : own;
                uS[c] = (!VELFB || oS != MISSING) ? ld2(a.vel_in + oS + lS + c * BS3) : own;
                uT[c] = (!VELFB || oT != MISSING) ? ld2(a.vel_in + oT + lT + c * BS3) : own;
                uB[c] = (!VELFB || oB != MISSING) ? ld2(a.vel_in + oB + lB + c * BS3) : own;
            }
        }

    };
    auto relaxation_rate = [&]() {      // WALE (:251-300) in the reference's expression order
    {
            const v2 h = V(0.5f);
            const v2 g11 = VMUL(h, vsub(uE[0], uW[0])), g12 = VMUL(h, vsub(uN[0], uS[0])), g13 = VMUL(h, vsub(uT[0], uB[0]));
            const v2 g21 = VMUL(h, vsub(uE[1], uW[1])), g22 = VMUL(h, vsub(uN[1], uS[1])), g23 = VMUL(h, vsub(uT[1], uB[1]));
            const v2 g31 = VMUL(h, vsub(uE[2], uW[2])), g32 = VMUL(h, vsub(uN[2], uS[2])), g33 = VMUL(h, vsub(uT[2], uB[2]));
#define DOT3(a1, b1, a2, b2, a3, b3) vadd(vadd(VMUL(a1, b1), VMUL(a2, b2)), VMUL(a3, b3))
            const v2 gsq11 = DOT3(g11, g11, g12, g21, g13, g31), gsq12 = DOT3(g11, g12, g12, g22, g13, g32), gsq13 = DOT3(g11, g13, g12, g23, g13, g33);
            const v2 gsq21 = DOT3(g21, g11, g22, g21, g23, g31), gsq22 = DOT3(g21, g12, g22, g22, g23, g32), gsq23 = DOT3(g21, g13, g22, g23, g23, g33);
            const v2 gsq31 = DOT3(g31, g11, g32, g21, g33, g31), gsq32 = DOT3(g31, g12, g32, g22, g33, g32), gsq33 = DOT3(g31, g13, g32, g23, g33, g33);
            const v2 tr_gsq = vadd(vadd(gsq11, gsq22), gsq33);
            const v2 tr_term = vdiv(tr_gsq, V(3.0f));
            const v2 Sd11 = vsub(gsq11, tr_term), Sd22 = vsub(gsq22, tr_term), Sd33 = vsub(gsq33, tr_term);
            const v2 Sd12 = VMUL(h, vadd(gsq12, gsq21)), Sd13 = VMUL(h, vadd(gsq13, gsq31)), Sd23 = VMUL(h, vadd(gsq23, gsq32));
            const v2 S12 = VMUL(h, vadd(g12, g21)), S13 = VMUL(h, vadd(g13, g31)), S23 = VMUL(h, vadd(g23, g32));
            const v2 OP1 = vadd(DOT3(Sd11, Sd11, Sd22, Sd22, Sd33, Sd33), VMUL(V(2.0f), DOT3(Sd12, Sd12, Sd13, Sd13, Sd23, Sd23)));
            const v2 OP2 = vadd(DOT3(g11, g11, g22, g22, g33, g33), VMUL(V(2.0f), DOT3(S12, S12, S13, S13, S23, S23)));
#undef DOT3
            const v2 OP1_32 = VMUL(OP1, vsqrt(OP1));
            const v2 OP2_52 = VMUL(VMUL(OP2, OP2), vsqrt(vmaxs(OP2, 1.0e-12f)));
            const v2 denom = vadd(OP2_52, VMUL(OP1, vsqrt(vsqrt(vmaxs(OP1, 1.0e-12f)))));
            const float cw2 = __fmul_rn(a.c_wale, a.c_wale);
            float ne0 = 0.0f, ne1 = 0.0f;
            if (OP1.x > 1.0e-12f && denom.x > 1.0e-12f) ne0 = __fdiv_rn(__fmul_rn(cw2, OP1_32.x), denom.x);
            if (OP1.y > 1.0e-12f && denom.y > 1.0e-12f) ne1 = __fdiv_rn(__fmul_rn(cw2, OP1_32.y), denom.y);
            const v2 nu_eddy = vmaxs(make_float2(ne0, ne1), a.nu_bg);
            const v2 tau_turb = vadd(V(a.tau), VMUL(nu_eddy, V(3.0f)));
            omega = vdiv(V(1.0f), vmaxs(tau_turb, 0.500001f));
        }

    };
    // obstacle flags first: a thread whose two cells are solid only bounces its populations back (:154-166) and needs no relaxation rate
    bool obsA = false, obsB = false;
    if (FULL && (bflags & BF_OBSTACLE)) {
        const uchar2 o = *reinterpret_cast<const uchar2*>(a.obstacle + (size_t)b * BS3 + c0);
        obsA = o.x != 0; obsB = o.y != 0;
    }
    if (WALE_FIRST && !(FULL && obsA && obsB)) { load_neighbour_velocities(); relaxation_rate(); }

    // ---- pull-stream (:62-149) with the moment sums of :144-148 taken in k order as the values arrive; combo (jy,jz) yields the
    // three consecutive directions k0-1, k0, k0+1 and the combos are visited in ascending k0
    FStore<STASH> f(s_stash);
    v2 rho = V(0.f), jx = V(0.f), jy = V(0.f), jz = V(0.f);
#pragma unroll
    for (int jzc = 0; jzc < 3; ++jzc) {
#pragma unroll
        for (int jyc = 0; jyc < 3; ++jyc) {
            const int loc = zoff[jzc] + yoff[jyc], dir = zdir[jzc] + ydir[jyc];
            const int k0 = 1 + 3 * jyc + 9 * jzc, kp = k0 + 1, km = k0 - 1;
            const long long o0 = s_fo[dir + 1], oM = s_fo[dir + dM], oP = s_fo[dir + dP];
            v2 fm, f0, fp;
            if (!MISS || (!MISS_UNIFORM && o0 != MISSING && oM != MISSING && oP != MISSING)) {
                const float* __restrict__ P0 = fbase + o0 + (loc + x0);
                const float* __restrict__ PM = fbase + oM + (loc + xM);
                const float* __restrict__ PP = fbase + oP + (loc + xP);
                fm = make_float2(P0[km * BS3 + 1], PP[km * BS3]);   // cx=-1: sources x0+1, x0+2
                f0 = ld2(P0 + k0 * BS3);
                fp = make_float2(PM[kp * BS3], P0[kp * BS3]);       // cx=+1: sources x0-1, x0
            } else {
                // some source block may be missing (domain face).  MISS_UNIFORM: every thread of a domain-face block takes this path, so
                // that the lanes on the face (a quarter of every warp at an x face) do not make the warp execute both paths
                if (XONLY || o0 != MISSING) {
                    const float* __restrict__ P0 = fbase + o0 + (loc + x0);
                    f0 = ld2(P0 + k0 * BS3); fp.y = P0[kp * BS3]; fm.x = P0[km * BS3 + 1];
                } else {
                    f0.x = pull_missing(a, fin_own, k0, gx, gy, gz); f0.y = pull_missing(a, fin_own + 1, k0, gx + 1, gy, gz);
                    fp.y = pull_missing(a, fin_own + 1, kp, gx + 1, gy, gz);
                    fm.x = pull_missing(a, fin_own, km, gx, gy, gz);
                }
                fp.x = oM != MISSING ? fbase[oM + (loc + xM) + kp * BS3] : (XONLY || at_inlet) ? lat_w_of(kp) * xf.p_in : pull_missing(a, fin_own, kp, gx, gy, gz);
                fm.y = oP != MISSING ? fbase[oP + (loc + xP) + km * BS3] : (XONLY || at_outlet) ? lat_w_of(km) * xf.p_out : pull_missing(a, fin_own + 1, km, gx + 1, gy, gz);
            }
            f.set(km, fm); f.set(k0, f0); f.set(kp, fp);
            // rho += f_k; j += f_k c_k  (k = km: cx = -1, k0: cx = 0, kp: cx = +1; cy = jyc - 1, cz = jzc - 1)
            rho = (jzc == 0 && jyc == 0) ? fm : vadd(rho, fm);
            jx = vsub(jx, fm);
            if (jyc == 2) jy = vadd(jy, fm); else if (jyc == 0) jy = vsub(jy, fm);
            if (jzc == 2) jz = vadd(jz, fm); else if (jzc == 0) jz = vsub(jz, fm);
            rho = vadd(rho, f0);
            if (jyc == 2) jy = vadd(jy, f0); else if (jyc == 0) jy = vsub(jy, f0);
            if (jzc == 2) jz = vadd(jz, f0); else if (jzc == 0) jz = vsub(jz, f0);
            rho = vadd(rho, fp);
            jx = vadd(jx, fp);
            if (jyc == 2) jy = vadd(jy, fp); else if (jyc == 0) jy = vsub(jy, fp);
            if (jzc == 2) jz = vadd(jz, fp); else if (jzc == 0) jz = vsub(jz, fp);
        }
    }

    float* __restrict__ fout = a.f_out + (size_t)b * (Q * BS3) + c0;
    float* __restrict__ vout = a.vel_out + (size_t)b * (3 * BS3) + c0;
    float* __restrict__ rout = a.rho_out + (size_t)b * BS3 + c0;

    // Full-way bounce-back (:154-166): f_out[26-k] = pulled f_k, vel = 0, rho = 1.  No arithmetic.
    const bool anyobs = FULL && (obsA || obsB);
    if (anyobs) {
        if (obsA && obsB) {
#pragma unroll
            for (int k = 0; k < 27; ++k) st2(fout + (26 - k) * BS3, f.get(k));
            st2(vout, V(0.f)); st2(vout + BS3, V(0.f)); st2(vout + 2 * BS3, V(0.f)); st2(rout, V(1.0f));
            return;
        }
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            if (obsA) fout[(26 - k) * BS3] = f.get(k).x; else fout[(26 - k) * BS3 + 1] = f.get(k).y;
        }
    }

    if (!WALE_FIRST) load_neighbour_velocities();

    rho = vmaxs(rho, 0.01f);                                          // :172
    const v2 inv_rho = vdiv(V(1.0f), rho);
    v2 ux = VMUL(jx, inv_rho), uy = VMUL(jy, inv_rho), uz = VMUL(jz, inv_rho);

    // ---- sponge (:181-199).  Lanes with sp = 0 keep their bits: x * (1 - 0) + y * 0 = x.
    if (FULL && (bflags & BF_SPONGE)) {
        const v2 sp = ld2(a.sponge + (size_t)b * BS3 + c0);
        const v2 om = vsub(V(1.0f), sp);
        rho = vadd(VMUL(rho, om), VMUL(V(1.0f), sp));
        const float u_in = k1_u_inlet(a);
        ux = vadd(VMUL(ux, om), VMUL(V(u_in), sp));
        uy = VMUL(uy, om);
        uz = VMUL(uz, om);
        if (a.sponge_blend == 1) {
            Unroll<0, 27>::run([&]<int K>() {
                const float feq_t = calc_eq(1.0f, u_in, 0.0f, 0.0f, lat_w(K), (float)lat_cx(K), (float)lat_cy(K), (float)lat_cz(K));
                f.set(K, vadd(VMUL(f.get(K), om), VMUL(V(feq_t), sp)));
            });
        }
    }

    // ---- wall-model force (:202-236)
    v2 Fx = V(0.f), Fy = V(0.f), Fz = V(0.f);
    bool has_force = false;
    if (FULL && a.wm == 1 && (bflags & BF_WALLDIST)) {
        const v2 dw = ld2(a.wall_dist + (size_t)b * BS3 + c0);
        if (dw.x > 0.0f && dw.x < 10.0f && !obsA) { float3 F = wall_force(dw.x, rho.x, ux.x, uy.x, uz.x, a.tau, a.wm_c166); Fx.x = F.x; Fy.x = F.y; Fz.x = F.z; }
        if (dw.y > 0.0f && dw.y < 10.0f && !obsB) { float3 F = wall_force(dw.y, rho.y, ux.y, uy.y, uz.y, a.tau, a.wm_c166); Fx.y = F.x; Fy.y = F.y; Fz.y = F.z; }
        // The force enters u_eq and the force term of the collision loop only through sums and products in which a zero force
        // contributes +-0 (u + 0.5 * 0 / rho, out + hw * w * (a . 0)): where no cell of the warp's z-plane carries a force — the outer
        // part of the 10-cell near-wall shell, cells the law of the wall leaves alone — both are skipped, as in blocks without wall distance.
        const bool nz = Fx.x != 0.0f || Fx.y != 0.0f || Fy.x != 0.0f || Fy.y != 0.0f || Fz.x != 0.0f || Fz.y != 0.0f;
        has_force = __any_sync(__activemask(), nz) != 0;
    }
    // u_eq = u + 0.5 F inv_rho with the PRE-sponge 1/rho (:238); without a force it is u (+0 changes no value)
    v2 uxe = ux, uye = uy, uze = uz;
    if (FULL && has_force) {
        uxe = vadd(ux, VMUL(VMUL(V(0.5f), Fx), inv_rho));
        uye = vadd(uy, VMUL(VMUL(V(0.5f), Fy), inv_rho));
        uze = vadd(uz, VMUL(VMUL(V(0.5f), Fz), inv_rho));
    }
    const v2 usq = vadd(vadd(VMUL(uxe, uxe), VMUL(uye, uye)), VMUL(uze, uze));

    // vel_out / rho_out (:155-158, :243-246)
    if (anyobs) {
        if (obsA) { vout[0] = 0.f; vout[BS3] = 0.f; vout[2 * BS3] = 0.f; rout[0] = 1.f; vout[1] = ux.y; vout[BS3 + 1] = uy.y; vout[2 * BS3 + 1] = uz.y; rout[1] = rho.y; }
        else { vout[1] = 0.f; vout[BS3 + 1] = 0.f; vout[2 * BS3 + 1] = 0.f; rout[1] = 1.f; vout[0] = ux.x; vout[BS3] = uy.x; vout[2 * BS3] = uz.x; rout[0] = rho.x; }
    } else {
        st2(vout, ux); st2(vout + BS3, uy); st2(vout + 2 * BS3, uz); st2(rout, rho);
    }

    if (!WALE_FIRST) relaxation_rate();

    // ---- Pi loop (:308-322): the stored f_k is replaced by feq_k
    const v2 usq15 = VMUL(V(1.5f), usq);
    const v2 rw0 = VMUL(rho, V(lat_w(13))), rw1 = VMUL(rho, V(lat_w(12))), rw2 = VMUL(rho, V(lat_w(9))), rw3 = VMUL(rho, V(lat_w(0)));
    v2 Pxx = V(0.f), Pyy = V(0.f), Pzz = V(0.f), Pxy = V(0.f), Pyz = V(0.f), Pzx = V(0.f);
    Unroll<0, 27>::run([&]<int K>() {
        constexpr int cx = lat_cx(K), cy = lat_cy(K), cz = lat_cz(K);
        const v2 rw = d2of(K) == 0 ? rw0 : d2of(K) == 1 ? rw1 : d2of(K) == 2 ? rw2 : rw3;
        v2 poly;
        if (K == 13) poly = vsub(V(1.0f), usq15);                                  // cu = 0: ((1 + 0) + 0) - 1.5 usq
        else {
            const v2 cu = cdot<K>(uxe, uye, uze);
            poly = vsub(vadd(vadd(V(1.0f), VMUL(V(3.0f), cu)), VMUL(VMUL(V(4.5f), cu), cu)), usq15);
        }
        const v2 feq = VMUL(rw, poly);
        const v2 fneq = vsub(f.get(K), feq);
        f.set(K, feq);
        if (cx != 0) Pxx = vadd(Pxx, fneq);
        if (cy != 0) Pyy = vadd(Pyy, fneq);
        if (cz != 0) Pzz = vadd(Pzz, fneq);
        if (cx * cy == 1) Pxy = vadd(Pxy, fneq); else if (cx * cy == -1) Pxy = vsub(Pxy, fneq);
        if (cy * cz == 1) Pyz = vadd(Pyz, fneq); else if (cy * cz == -1) Pyz = vsub(Pyz, fneq);
        if (cz * cx == 1) Pzx = vadd(Pzx, fneq); else if (cz * cx == -1) Pzx = vsub(Pzx, fneq);
    });

    // ---- collision loop (:324-354):  f_out = (feq + (1 - omega) f_neq_reg) + (1 - omega/2) force_term
    const float cs2 = 1.0f / 3.0f;
    const float qa = 1.0f - cs2, qb = 0.0f - cs2;                     // Q = c*c - CS2 for |c| = 1 and c = 0
    const v2 PQ[3][2] = {{VMUL(Pxx, V(qa)), VMUL(Pxx, V(qb))}, {VMUL(Pyy, V(qa)), VMUL(Pyy, V(qb))}, {VMUL(Pzz, V(qa)), VMUL(Pzz, V(qb))}};
    const v2 om1 = vsub(V(1.0f), omega);
    v2 hw = V(0.f);
    if (FULL && has_force) hw = vsub(V(1.0f), VMUL(V(0.5f), omega));
    Unroll<0, 27>::run([&]<int K>() {
        constexpr int cx = lat_cx(K), cy = lat_cy(K), cz = lat_cz(K);
        // Pi_xx Q_xx + Pi_yy Q_yy + Pi_zz Q_zz
        const v2 diag = vadd(vadd(PQ[0][cx != 0 ? 0 : 1], PQ[1][cy != 0 ? 0 : 1]), PQ[2][cz != 0 ? 0 : 1]);
        // Pi_xy cx cy + Pi_yz cy cz + Pi_zx cz cx
        v2 off = V(0.f);
        bool have = false;
        if (cx * cy != 0) { off = cx * cy > 0 ? Pxy : vneg(Pxy); have = true; }
        if (cy * cz != 0) { off = have ? (cy * cz > 0 ? vadd(off, Pyz) : vsub(off, Pyz)) : (cy * cz > 0 ? Pyz : vneg(Pyz)); have = true; }
        if (cz * cx != 0) { off = have ? (cz * cx > 0 ? vadd(off, Pzx) : vsub(off, Pzx)) : (cz * cx > 0 ? Pzx : vneg(Pzx)); have = true; }
        const v2 inner = have ? vadd(diag, VMUL(V(2.0f), off)) : diag;     // + 2 * 0 changes nothing
        const v2 fnr = VMUL(V(lat_w(K) * 4.5f), inner);
        v2 out = vadd(f.get(K), VMUL(om1, fnr));
        if (FULL && has_force) {
            // force_term = (w 3) * (((cx - ux + 3 cu cx) Fx + (cy - uy + 3 cu cy) Fy) + (cz - uz + 3 cu cz) Fz), cu from u_eq, u from u
            const v2 cu = K == 13 ? V(0.f) : cdot<K>(uxe, uye, uze);
            const v2 cu3 = VMUL(V(3.0f), cu);
            // (c_a - u_a) + (3 cu) c_a : a zero component contributes (0 - u_a) + (+-0) = -u_a
            const v2 ax = cx != 0 ? vadd(vsub(V((float)cx), ux), cx > 0 ? cu3 : vneg(cu3)) : vsub(V(0.f), ux);
            const v2 ay = cy != 0 ? vadd(vsub(V((float)cy), uy), cy > 0 ? cu3 : vneg(cu3)) : vsub(V(0.f), uy);
            const v2 az = cz != 0 ? vadd(vsub(V((float)cz), uz), cz > 0 ? cu3 : vneg(cu3)) : vsub(V(0.f), uz);
            const v2 dotF = vadd(vadd(VMUL(ax, Fx), VMUL(ay, Fy)), VMUL(az, Fz));
            const v2 force_term = VMUL(V(lat_w(K) * 3.0f), dotF);
            out = vadd(out, VMUL(hw, force_term));
        }
        if (!anyobs) st2(fout + K * BS3, out);
        else if (obsA) fout[K * BS3 + 1] = out.y;
        else fout[K * BS3] = out.x;
    });
}


// NT threads per CTA: 256 = one CTA per block; 128 / 64 = a CTA takes 4 / 2 of the block's z-planes, so that more, smaller CTAs are
// resident per SM and their load and compute phases interleave (the register-limited occupancy is the same number of warps).
// OCC: resident warps per SM in units of four (16 / 20 / 24 warps -> 128 / 96 / 80 registers per thread).  With the relaxation rate
// computed before the pull (see strict_block) 96 registers cost 32 bytes of spills and 80 registers 112 bytes.
// LOOP: consecutive parts of ONE block a CTA works through (64-thread CTAs, LOOP = 4: the CTA walks up the block's four z-plane
// pairs): list entry -> neighbour table -> offsets (two dependent memory round trips and a barrier before the first useful load)
// are paid once per block instead of once per part, and the z-halo planes of one part are the own planes of the next (L1).
template <bool FULL, bool VELFB, bool MISS, bool STASH, int NT, int OCC = 4, int LOOP = 1>
__global__ void __launch_bounds__(NT, STASH ? 3 * (256 / NT) : (OCC * 128) / NT) k1_strict_kernel(const __grid_constant__ K1Args a) {
    extern __shared__ float2 s_stash[];
    __shared__ long long s_fo[27];   // element offset of each neighbour block relative to f_in (MISSING: no block)
    __shared__ long long s_vo[27];   // ... relative to vel_in (MISSING for ghost blocks: they carry populations only)
    constexpr int CTAS = 256 / (NT * LOOP);   // CTAs per block
    const int b = a.list[blockIdx.x / CTAS];
    if (threadIdx.x < 27) neighbour_offsets(a, b, threadIdx.x, (long long)b * (Q * BS3), s_fo, s_vo);
    __syncthreads();
    if (LOOP == 1) prefetch_block_part<NT>(a, (int)(blockIdx.x / CTAS), (int)(blockIdx.x % CTAS));
#pragma unroll 1
    for (int l = 0; l < LOOP; ++l)
        strict_block<FULL, VELFB, MISS, STASH, (OCC > 4 || FULL)>(a, b, (int)threadIdx.x + ((int)(blockIdx.x % CTAS) * LOOP + l) * NT, a.f_in, s_fo, s_vo, s_stash);
}

// Merged launch of the plain and the domain-face class (abi.cu, option merge_face): one grid over both lists in spatial order, a
// CTA-uniform branch picks the body.  The face blocks (3 % of the bench box) then run BETWEEN plain CTAs on the same SMs instead of as
// 28 half-empty waves of latency-bound CTAs after the plain launch; the register budget is the plain kernel's (96), the face body
// spills into it, which only the face CTAs pay.  Same bodies, same bits.
// strict mode: the face blocks of the merged list are the X-ONLY ones (see strict_block), whose body is as lean as the plain one; the
// general domain-face class (y / z faces, features, anything that needs pull_missing) keeps its own launch.
__global__ void __launch_bounds__(64, 10) k1_strict_mixed_kernel(const __grid_constant__ K1Args a) {
    __shared__ long long s_fo[27], s_vo[27];
    const int b = a.list[blockIdx.x >> 2];
    bool miss = false;
    if (threadIdx.x < 27) { neighbour_offsets(a, b, threadIdx.x, (long long)b * (Q * BS3), s_fo, s_vo); miss = s_fo[threadIdx.x] == MISSING; }
    const bool face = __syncthreads_or(miss) != 0;
    const int t = (int)threadIdx.x + (int)(blockIdx.x & 3) * 64;
    if (face) strict_block<false, true, true, false, true, true, true>(a, b, t, a.f_in, s_fo, s_vo, nullptr);
    else strict_block<false, false, false, false, true>(a, b, t, a.f_in, s_fo, s_vo, nullptr);
}

// Persistent form for a SMALL latency-bound class (the domain-face blocks of a large level): a few 64-thread CTAs per SM stride over
// the (block, z-plane pair) parts of the list.  Launched BEFORE the plain kernel on a side stream they keep a fixed, small share of
// every SM (2 CTAs = 16 K registers) for their whole run, the HBM-bound plain launch fills the rest, and the class costs no tail of
// half-empty waves after it (abi.cu, option face_persist).  Same strict_block body, same bits.
template <bool FULL, bool VELFB, bool MISS>
__global__ void __launch_bounds__(64, 8) k1_strict_persist_kernel(const __grid_constant__ K1Args a) {
    __shared__ long long s_fo[27], s_vo[27];
    const int total = a.n_list * 4;
    for (int e = blockIdx.x; e < total; e += gridDim.x) {
        const int b = a.list[e >> 2];
        __syncthreads();                       // every warp is done with the previous part's tables
        if (threadIdx.x < 27) neighbour_offsets(a, b, threadIdx.x, (long long)b * (Q * BS3), s_fo, s_vo);
        __syncthreads();
        strict_block<FULL, VELFB, MISS, false, true>(a, b, (int)threadIdx.x + (e & 3) * 64, a.f_in, s_fo, s_vo, nullptr);
    }
}

// ---- TMA variant: persistent CTAs, the block's own 27 x 2 KiB population planes (one contiguous 54 KiB run in the block-major
// layout) staged into shared memory by ONE cp.async.bulk per block, double-buffered: while a CTA computes block i the copy
// engine already fills the other stage with block i + gridDim.x, whatever the register-limited occupancy (2 CTAs x 128
// registers per SM) would allow the LSU to keep in flight.  What comes from the 26 neighbour blocks (23 % of the pulled values:
// faces, edges, corners — strided 4-byte elements, not expressible as a bulk copy) and the velocities stay direct loads; the next
// block's own velocities are prefetched into L2.  Same strict_block body, same bits.
template <bool FULL, bool VELFB, bool MISS>
__global__ void __launch_bounds__(256, 2) k1_strict_tma_kernel(const __grid_constant__ K1Args a) {
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
        strict_block<FULL, VELFB, MISS, false, false>(a, b, t, fbase, s_fo[cur], s_vo[cur], nullptr);
    }
}

}  // namespace k1s

constexpr int STASH_BYTES = 27 * 256 * (int)sizeof(float2);   // 55 296
// variant: 0 registers (2 CTAs / SM), 1 shared-memory stash (3 CTAs / SM), 2 persistent TMA-staged (2 CTAs / SM, double-buffered tiles)
template <bool FULL, bool VELFB, bool MISS>
void launch_strict(const K1Args& a, int variant, cudaStream_t s) {
    if (a.n_list <= 0) return;
    if (a.persist_grid > 0) {
        if constexpr (FULL && MISS) { k1s::k1_strict_persist_kernel<FULL, VELFB, MISS><<<a.persist_grid, 64, 0, s>>>(a); return; }
    }
    if (variant == 2) {
        static const cudaError_t once = cudaFuncSetAttribute(k1s::k1_strict_tma_kernel<FULL, VELFB, MISS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)k1s::TILE_BYTES);
        static const cudaError_t once2 = cudaFuncSetAttribute(k1s::k1_strict_tma_kernel<FULL, VELFB, MISS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        (void)once; (void)once2;
        const int grid = a.n_list < 2 * a.num_sms ? a.n_list : 2 * a.num_sms;
        k1s::k1_strict_tma_kernel<FULL, VELFB, MISS><<<grid, 256, 2 * k1s::TILE_BYTES, s>>>(a);
    } else if (variant == 1) {
        static const cudaError_t once = cudaFuncSetAttribute(k1s::k1_strict_kernel<FULL, VELFB, MISS, true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, STASH_BYTES);
        static const cudaError_t once2 = cudaFuncSetAttribute(k1s::k1_strict_kernel<FULL, VELFB, MISS, true, 256>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        (void)once; (void)once2;
        k1s::k1_strict_kernel<FULL, VELFB, MISS, true, 256><<<a.n_list, 256, STASH_BYTES, s>>>(a);
    } else if (a.cta_threads == 128) k1s::k1_strict_kernel<FULL, VELFB, MISS, false, 128><<<2 * a.n_list, 128, 0, s>>>(a);
    else if (a.cta_threads == 64) {
        // (the feature / domain-face instantiations would spill ~0.5-1 KB per thread at 96 / 80 registers: plain classes only)
        if constexpr (!FULL) {
            if (a.strict_loop == 4 && a.strict_occ == 5) { k1s::k1_strict_kernel<FULL, VELFB, MISS, false, 64, 5, 4><<<a.n_list, 64, 0, s>>>(a); return; }
            if (a.strict_loop == 2 && a.strict_occ == 5) { k1s::k1_strict_kernel<FULL, VELFB, MISS, false, 64, 5, 2><<<2 * a.n_list, 64, 0, s>>>(a); return; }
            if (a.strict_occ == 6) { k1s::k1_strict_kernel<FULL, VELFB, MISS, false, 64, 6><<<4 * a.n_list, 64, 0, s>>>(a); return; }
            if (a.strict_occ == 5) { k1s::k1_strict_kernel<FULL, VELFB, MISS, false, 64, 5><<<4 * a.n_list, 64, 0, s>>>(a); return; }
        }
        if constexpr (FULL) {
            if (a.strict_feat_occ == 5) { k1s::k1_strict_kernel<FULL, VELFB, MISS, false, 64, 5><<<4 * a.n_list, 64, 0, s>>>(a); return; }
            if (a.strict_feat_occ == 3) { k1s::k1_strict_kernel<FULL, VELFB, MISS, false, 64, 3><<<4 * a.n_list, 64, 0, s>>>(a); return; }
        }
        k1s::k1_strict_kernel<FULL, VELFB, MISS, false, 64, 4><<<4 * a.n_list, 64, 0, s>>>(a);
    } else k1s::k1_strict_kernel<FULL, VELFB, MISS, false, 256><<<a.n_list, 256, 0, s>>>(a);
}
void launch_k1s_plain(const K1Args& a, cudaStream_t s) { launch_strict<false, false, false>(a, a.strict_stash, s); }
void launch_k1s_mixed(const K1Args& a, cudaStream_t s) { if (a.n_list > 0) k1s::k1_strict_mixed_kernel<<<4 * a.n_list, 64, 0, s>>>(a); }
void launch_k1s_plain_ghost(const K1Args& a, cudaStream_t s) { launch_strict<false, true, false>(a, a.strict_stash, s); }
void launch_k1s_feat(const K1Args& a, cudaStream_t s) { launch_strict<true, true, false>(a, a.strict_stash, s); }
void launch_k1s_full(const K1Args& a, cudaStream_t s) { launch_strict<true, true, true>(a, a.strict_stash, s); }
__global__ void wall_model_constant_kernel(float* out) { *out = k1s::pow32(2.0f * 8.3f, -1.0f / 7.0f); }
void launch_wall_model_constant(float* d_out, cudaStream_t s) { wall_model_constant_kernel<<<1, 1, 0, s>>>(d_out); }

void launch_ghost_interp_strict(const GhostArgs& g, cudaStream_t s) {
    if (g.n > 0) k1s::ghost_interp_kernel<<<(g.n + 127) / 128, 128, 0, s>>>(g);
}

}  // namespace ludwig
