// k1_boundary.cuh — device functions for populations whose source cell lies in a block that does not exist:
// domain-face boundary conditions (physics_kernels.jl:88-120,138-140), the inlet noise hash
// (physics_utils.jl:17-28) and the 2:1 coarse->fine interface interpolation with temporal blend and f_neq
// rescaling (physics_interpolation.jl:16-138).  Textually included inside a per-TU namespace by
// k1_generic.cuh (strict and fast builds) and k1_fast.cu, so each build gets its own FP-contraction mode.
// Expressions follow the reference's operation order.

constexpr float KAPPA = 0.41f;
constexpr float CS2_PHYSICS = 1.0f / 3.0f;

__device__ __forceinline__ uint32_t gpu_hash(int32_t x) {
    uint32_t h = (uint32_t)x;
    h = (h ^ (h >> 16)) * 0x85ebca6bu;
    h = (h ^ (h >> 13)) * 0xc2b2ae35u;
    return h ^ (h >> 16);
}
__device__ __forceinline__ float gradient_noise(int32_t gx, int32_t gy, int32_t gz, int32_t seed) {
    uint32_t combined = (uint32_t)gx * 374761393u + (uint32_t)gy * 668265263u + (uint32_t)gz * 1274126177u + (uint32_t)seed;
    uint32_t h = gpu_hash((int32_t)combined);
    return ((float)(h & 0xFFFFu) / 32768.0f) - 1.0f;
}
__device__ __forceinline__ float calc_eq(float rho, float ux, float uy, float uz, float w_k, float cx, float cy, float cz) {
    float cu = cx * ux + cy * uy + cz * uz;
    float usq = ux * ux + uy * uy + uz * uz;
    return rho * w_k * (1.0f + 3.0f * cu + 4.5f * cu * cu - 1.5f * usq);
}

// Julia's Float32 power and logarithm, which the wall model calls (physics_kernels.jl:209,213): Base evaluates both in
// Float64 and rounds once — x^y = Float32(exp2(log2(Float64(x)) * y)) (base/math.jl, pow_body for Float16 / Float32) and
// log(::Float32) through the Float64 tables of base/special/log.jl.  Restated the same way here and in the CPU oracle, the
// strict build matches the oracle bit for bit with the wall model on (libdevice powf / logf are 2-4 ulp functions).
__device__ __forceinline__ float pow32(float x, float y) { return (float)exp2(log2((double)x) * (double)y); }
__device__ __forceinline__ float log32(float x) { return (float)log((double)x); }

// inlet velocity and noise seed of this launch: immediates, or (CUDA-graph replay of a coarse step) derived from device memory
__device__ __forceinline__ float k1_u_inlet(const K1Args& a) { return a.dyn ? a.dyn->u_inlet : a.u_inlet; }
__device__ __forceinline__ int k1_seed(const K1Args& a) {
    return a.dyn ? (int)(((a.dyn->t_coarse << a.dyn_shift) + (long long)a.dyn_add) % 1000000LL) : a.seed;
}

struct Corner { float v[5]; bool ok; };

__device__ __forceinline__ Corner get_blended(const K1Args& a, int pgx, int pgy, int pgz, int k, float w_k) {
    Corner c;
    int pbx = (pgx - 1) / BS, pby = (pgy - 1) / BS, pbz = (pgz - 1) / BS;   // 0-based block coords
    if (pbx >= 0 && pbx < a.pdimx && pby >= 0 && pby < a.pdimy && pbz >= 0 && pbz < a.pdimz) {
        int pb = a.pptr[pbx + a.pdimx * (pby + a.pdimy * pbz)];
        if (pb >= 0) {
            pb &= PTR_LOCAL_MASK;   // single-rank paths only: the owner bits are 0
            int loc = ((pgx - 1) & 7) + 8 * ((pgy - 1) & 7) + 64 * ((pgz - 1) & 7);
            size_t fi = ((size_t)pb * Q + k) * BS3 + loc;
            size_t ri = (size_t)pb * BS3 + loc;
            size_t vi = (size_t)pb * 3 * BS3 + loc;
            float f_new = a.pf_new[fi], rho_new = a.prho_new[ri];
            float ux_new = a.pvel_new[vi], uy_new = a.pvel_new[vi + BS3], uz_new = a.pvel_new[vi + 2 * BS3];
            if (a.use_temporal == 1 && a.tw < 0.99f) {
                float f_old = a.pf_old[fi], rho_old = a.prho_old[ri];
                float ux_old = a.pvel_old[vi], uy_old = a.pvel_old[vi + BS3], uz_old = a.pvel_old[vi + 2 * BS3];
                float tw = a.tw;
                c.v[0] = f_old * (1.0f - tw) + f_new * tw;
                c.v[1] = rho_old * (1.0f - tw) + rho_new * tw;
                c.v[2] = ux_old * (1.0f - tw) + ux_new * tw;
                c.v[3] = uy_old * (1.0f - tw) + uy_new * tw;
                c.v[4] = uz_old * (1.0f - tw) + uz_new * tw;
            } else {
                c.v[0] = f_new; c.v[1] = rho_new; c.v[2] = ux_new; c.v[3] = uy_new; c.v[4] = uz_new;
            }
            c.ok = true;
            return c;
        }
    }
    c.v[0] = w_k; c.v[1] = 1.0f; c.v[2] = 0.0f; c.v[3] = 0.0f; c.v[4] = 0.0f; c.ok = false;
    return c;
}

// physics_interpolation.jl:16-138
__device__ __noinline__ float interpolate_with_rescaling(const K1Args& a, int fine_gx, int fine_gy, int fine_gz, int k) {
    const int d2 = lat_cx(k) * lat_cx(k) + lat_cy(k) * lat_cy(k) + lat_cz(k) * lat_cz(k);
    const float w_k = d2 == 0 ? 8.0f / 27.0f : d2 == 1 ? 2.0f / 27.0f : d2 == 2 ? 1.0f / 54.0f : 1.0f / 216.0f;
    const float cx = (float)lat_cx(k), cy = (float)lat_cy(k), cz = (float)lat_cz(k);
    float px_cont = ((float)fine_gx - 0.5f) * 0.5f;
    float py_cont = ((float)fine_gy - 0.5f) * 0.5f;
    float pz_cont = ((float)fine_gz - 0.5f) * 0.5f;
    int px0 = (int)floorf(px_cont), py0 = (int)floorf(py_cont), pz0 = (int)floorf(pz_cont);
    int px1 = px0 + 1, py1 = py0 + 1, pz1 = pz0 + 1;
    float wx = px_cont - (float)px0, wy = py_cont - (float)py0, wz = pz_cont - (float)pz0;
    px0 = max(1, px0); py0 = max(1, py0); pz0 = max(1, pz0);

    Corner v000 = get_blended(a, px0, py0, pz0, k, w_k);
    Corner v100 = get_blended(a, px1, py0, pz0, k, w_k);
    Corner v010 = get_blended(a, px0, py1, pz0, k, w_k);
    Corner v110 = get_blended(a, px1, py1, pz0, k, w_k);
    Corner v001 = get_blended(a, px0, py0, pz1, k, w_k);
    Corner v101 = get_blended(a, px1, py0, pz1, k, w_k);
    Corner v011 = get_blended(a, px0, py1, pz1, k, w_k);
    Corner v111 = get_blended(a, px1, py1, pz1, k, w_k);
    if (!v100.ok) v100 = v000;
    if (!v010.ok) v010 = v000;
    if (!v110.ok) v110 = v000;
    if (!v001.ok) v001 = v000;
    if (!v101.ok) v101 = v000;
    if (!v011.ok) v011 = v000;
    if (!v111.ok) v111 = v000;
    float r[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        float c00 = v000.v[i] * (1.0f - wx) + v100.v[i] * wx;
        float c01 = v001.v[i] * (1.0f - wx) + v101.v[i] * wx;
        float c10 = v010.v[i] * (1.0f - wx) + v110.v[i] * wx;
        float c11 = v011.v[i] * (1.0f - wx) + v111.v[i] * wx;
        float c0 = c00 * (1.0f - wy) + c10 * wy;
        float c1 = c01 * (1.0f - wy) + c11 * wy;
        r[i] = c0 * (1.0f - wz) + c1 * wz;
    }
    float feq_int = calc_eq(r[1], r[2], r[3], r[4], w_k, cx, cy, cz);
    float f_neq = r[0] - feq_int;
    float tau_c = a.tau_parent - 0.5f, tau_f = a.tau - 0.5f;
    float scale = tau_c > 1.0e-6f ? fminf(fmaxf(tau_f / tau_c, 0.01f), 100.0f) : 1.0f;
    return feq_int + f_neq * scale;
}

// The two x-face cases of pull_missing (below) in closed form, once per thread instead of once per pulled population: every
// population a cell at global x = 1 pulls from x = 0 has c_x = +1 and is the inlet equilibrium w_k * P_in with the same
// P_in = 1 + 3 u + 4.5 u u - 1.5 u u, u = u_inlet + noise(gy, gz, seed) (physics_kernels.jl:99-107: the noise is a function of the
// PULLING cell's gy, gz); every population a cell at x = nx pulls from x = nx + 1 has c_x = -1 and is w_k * P_out with cu = -u_inlet
// (:108-113).  Same expressions in the same order as pull_missing, so the products w_k * P are the same bits.
struct XFace { float p_in, p_out; };
__device__ __forceinline__ XFace x_face_equilibria(const K1Args& a, int gy, int gz) {
    const float u_inlet = k1_u_inlet(a);
    const float noise = a.inlet_turb > 0.0f ? gradient_noise(gy, gz, k1_seed(a), 1234) * a.inlet_turb * u_inlet : 0.0f;
    const float u_inst = u_inlet + noise;
    const float cu_in = 1.0f * u_inst, cu_out = -1.0f * u_inlet;
    XFace f;
    f.p_in = 1.0f + 3.0f * cu_in + 4.5f * cu_in * cu_in - 1.5f * u_inst * u_inst;
    f.p_out = 1.0f + 3.0f * cu_out + 4.5f * cu_out * cu_out - 1.5f * u_inlet * u_inlet;
    return f;
}
__host__ __device__ constexpr float lat_w_of(int k) {
    return ((k % 3 != 1) + ((k / 3) % 3 != 1) + (k / 9 != 1)) == 0   ? 8.0f / 27.0f
           : ((k % 3 != 1) + ((k / 3) % 3 != 1) + (k / 9 != 1)) == 1 ? 2.0f / 27.0f
           : ((k % 3 != 1) + ((k / 3) % 3 != 1) + (k / 9 != 1)) == 2 ? 1.0f / 54.0f
                                                                     : 1.0f / 216.0f;
}

// Everything that can happen to a population whose source cell is in a block that does not exist
// (physics_kernels.jl:88-140).  Kept out of line: it is the rare path.
__device__ __noinline__ float pull_missing(const K1Args& a, const float* __restrict__ fin_cell, int k, int gx, int gy, int gz) {
    const int cx = lat_cx(k), cy = lat_cy(k), cz = lat_cz(k);
    const float w_k = (cx * cx + cy * cy + cz * cz) == 0 ? 8.0f / 27.0f
                      : (cx * cx + cy * cy + cz * cz) == 1 ? 2.0f / 27.0f
                      : (cx * cx + cy * cy + cz * cz) == 2 ? 1.0f / 54.0f
                                                           : 1.0f / 216.0f;
    int src_gx = gx - cx, src_gy = gy - cy, src_gz = gz - cz;
    bool is_inlet = src_gx < 1, is_outlet = src_gx > a.nxg;
    bool is_y_min = src_gy < 1, is_y_max = src_gy > a.nyg;
    bool is_z_min = src_gz < 1, is_z_max = src_gz > a.nzg;
    const float u_inlet = k1_u_inlet(a);
    if (is_inlet) {
        float noise = a.inlet_turb > 0.0f ? gradient_noise(gy, gz, k1_seed(a), 1234) * a.inlet_turb * u_inlet : 0.0f;
        float u_inst = u_inlet + noise;
        float cu_in = (float)cx * u_inst;
        return w_k * (1.0f + 3.0f * cu_in + 4.5f * cu_in * cu_in - 1.5f * u_inst * u_inst);
    } else if (is_outlet) {
        float cu_out = (float)cx * u_inlet;
        return w_k * (1.0f + 3.0f * cu_out + 4.5f * cu_out * cu_out - 1.5f * u_inlet * u_inlet);
    } else if (is_y_min || is_y_max) {   // (the symmetric flag makes no difference, physics_kernels.jl:115-118)
        return fin_cell[(k - 6 * cy) * BS3];
    } else if (is_z_min || is_z_max) {
        return fin_cell[(k - 18 * cz) * BS3];
    } else if (a.is_l1 == 0) {
        return interpolate_with_rescaling(a, src_gx, src_gy, src_gz, k);
    }
    return w_k;
}

