// ludwig_oracle.cpp — CPU restatement of OPEN_Ludwig's per-timestep D3Q27 hot path.
//
// *** TEST INFRASTRUCTURE — NOT PRODUCT CODE. ***
// This file is the parity oracle: a line-by-line C++ restatement of the reference's Julia
// KernelAbstractions kernels and of the host recursion that launches them.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
// The product path (open_ludwig_b200/csrc, libludwig_b200.so) never links or calls it.
//
// It exports the SAME C ABI as include/ludwig_b200.h so that one Python driver can run either
// library on identical inputs.  Arithmetic is FP32, evaluated in the reference's operation
// order; build with -ffp-contract=off (the KernelAbstractions CPU backend does not contract
// a*b+c into FMA).  Arrays are kept in the reference's own layout (Julia column-major):
//   f[x,y,z,b,k]  -> x + 8y + 64z + 512b + 512*nb*k   (0-based here, 1-based in Julia).
//
// Pinning: see oracle/README.md — topology/voxel/Bouzidi integer counts and the Cd / rho_min
// rows of RESULTS_SPHERE_RE1M.txt:165-169 are reproduced through this oracle
// (tests/test_golden_logs.py, tests/golden/).
//
// Reference sections restated (all under /root/reference/src):
//   K0 init_eq!                         main.jl:109-124
//   K1 stream_collide_kernel_v2!        physics_kernels.jl:9-358
//      gradient_noise / gpu_hash        physics_utils.jl:17-28
//      calculate_equilibrium            physics_utils.jl:34-39
//      compute_velocity_gradients       physics_utils.jl:45-83
//      interpolate_with_rescaling       physics_interpolation.jl:16-138
//   K2 bouzidi_correction_kernel_fixed! bouzidi_kernel.jl:13-92
//   K3 map_stresses_kernel!             forces/surface.jl:138-266
//   K4 integrate_forces_kernel!         forces/surface.jl:282-366  (+ host math :467-572)
//   M1 copy_to_old!                     blocks.jl:199-205
//   R1 compute_flow_stats               diagnostics.jl:56-94
//   schedule                            solver_control.jl:21-165, physics_v2.jl:26-117
#include "../include/ludwig_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

constexpr int BS = 8;        // blocks.jl:14
constexpr int BS3 = 512;
constexpr float KAPPA = 0.41f;             // physics_v2.jl:15
constexpr float CS2_PHYSICS = 1.0f / 3.0f; // physics_v2.jl:16

// physics_v2.jl:99-117 build_lattice_arrays_gpu: k ordered dz,dy,dx with dx fastest.
struct Lattice {
    int cx[27], cy[27], cz[27], opp[27], mirror_y[27], mirror_z[27];
    float w[27];
    Lattice() {
        int k = 0;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    cx[k] = dx; cy[k] = dy; cz[k] = dz;
                    int d2 = dx * dx + dy * dy + dz * dz;
                    w[k] = d2 == 0 ? 8.0f / 27.0f : d2 == 1 ? 2.0f / 27.0f : d2 == 2 ? 1.0f / 54.0f : 1.0f / 216.0f;
                    ++k;
                }
        for (int i = 0; i < 27; ++i)
            for (int j = 0; j < 27; ++j) {
                if (cx[j] == -cx[i] && cy[j] == -cy[i] && cz[j] == -cz[i]) opp[i] = j;
                if (cx[j] == cx[i] && cy[j] == -cy[i] && cz[j] == cz[i]) mirror_y[i] = j;
                if (cx[j] == cx[i] && cy[j] == cy[i] && cz[j] == -cz[i]) mirror_z[i] = j;
            }
    }
};
const Lattice LAT;

// physics_utils.jl:17-22
// Julia's x^y and log(x) for Float32 (the wall model, physics_kernels.jl:211,216): Base evaluates both in Float64 and rounds
// once — base/math.jl `pow_body(x::T, y::T) where T<:Union{Float16,Float32} = T(exp2(log2(abs(widen(x))) * y))`, and
// base/special/log.jl computes log(::Float32) with Float64 tables.  glibc's powf / logf are 0.8-ulp functions of their own and
// would differ from that in ~30 % of the calls.
inline float pow32(float x, float y) { return (float)std::exp2(std::log2((double)x) * (double)y); }
inline float log32(float x) { return (float)std::log((double)x); }

inline uint32_t gpu_hash(int32_t x) {
    uint32_t h = static_cast<uint32_t>(x);
    h = (h ^ (h >> 16)) * 0x85ebca6bu;
    h = (h ^ (h >> 13)) * 0xc2b2ae35u;
    return h ^ (h >> 16);
}
// physics_utils.jl:24-28 (wrapping Int32 arithmetic)
inline float gradient_noise(int32_t gx, int32_t gy, int32_t gz, int32_t seed) {
    uint32_t combined = static_cast<uint32_t>(gx) * 374761393u + static_cast<uint32_t>(gy) * 668265263u +
                        static_cast<uint32_t>(gz) * 1274126177u + static_cast<uint32_t>(seed);
    uint32_t h = gpu_hash(static_cast<int32_t>(combined));
    return (static_cast<float>(h & 0xFFFFu) / 32768.0f) - 1.0f;
}
// physics_utils.jl:34-39
inline float calculate_equilibrium(float rho, float ux, float uy, float uz, float w_k, float cx, float cy, float cz) {
    float cu = cx * ux + cy * uy + cz * uz;
    float usq = ux * ux + uy * uy + uz * uz;
    return rho * w_k * (1.0f + 3.0f * cu + 4.5f * cu * cu - 1.5f * usq);
}

inline float half_to_float(uint16_t h) {
    uint32_t sign = (h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            int e = -1;
            do { ++e; man <<= 1; } while ((man & 0x400u) == 0);
            bits = sign | ((127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp - 15 + 127) << 23) | (man << 13);
    float f; std::memcpy(&f, &bits, 4); return f;
}

struct Level {
    int level_id = 0, nb = 0, dimx = 0, dimy = 0, dimz = 0;
    float tau = 0; double dx = 0;
    std::vector<int32_t> block_pointer, neighbor_table, map_x, map_y, map_z;
    std::vector<uint8_t> obstacle;
    std::vector<float> sponge, wall_dist;
    bool temporal = false, bouzidi = false;
    int n_bc = 0;
    std::vector<uint16_t> q_map;
    std::vector<int32_t> cell_block;
    std::vector<int8_t> cell_x, cell_y, cell_z;
    // state (blocks.jl:118-147)
    std::vector<float> f, f_temp, f_post, f_old, rho, rho_old, vel, vel_temp, vel_old;
    size_t idx4(int x, int y, int z, int b) const { return x + 8 * y + 64 * z + 512 * (size_t)b; }
    size_t idx5(int x, int y, int z, int b, int k) const { return idx4(x, y, z, b) + 512 * (size_t)nb * k; }
    int nbr(int b, int dir) const { return neighbor_table[b + (size_t)nb * dir]; }  // 1-based result, 0 none
};

struct Parent {  // what recursive_step_temporal! passes down (solver_control.jl:65-72)
    const float *f_new, *rho_new, *vel_new, *f_old, *rho_old, *vel_old;
    const int32_t* ptr; int dimx, dimy, dimz; int nb; float tau;
};

}  // namespace

struct ludwig_mesh { int n; std::vector<float> cx, cy, cz, nx, ny, nz, area; };
struct ludwig_forces {
    const ludwig_mesh* mesh; double rho_ref, u_ref, area_ref, chord_ref, mc[3]; int symmetric;
    std::vector<float> p, sx, sy, sz;
};
struct ludwig_ctx { std::vector<std::unique_ptr<Level>> levels; std::string err; };

namespace {

// physics_interpolation.jl:16-138
float interpolate_with_rescaling(const Parent& P, int fine_gx, int fine_gy, int fine_gz, int k, float w_k, float cx,
                                 float cy, float cz, float tau_coarse, float tau_fine, float temporal_weight,
                                 int use_temporal_interp) {
    float px_cont = ((float)fine_gx - 0.5f) * 0.5f;
    float py_cont = ((float)fine_gy - 0.5f) * 0.5f;
    float pz_cont = ((float)fine_gz - 0.5f) * 0.5f;
    int px0 = (int)std::floor(px_cont), py0 = (int)std::floor(py_cont), pz0 = (int)std::floor(pz_cont);
    int px1 = px0 + 1, py1 = py0 + 1, pz1 = pz0 + 1;
    float wx = px_cont - (float)px0, wy = py_cont - (float)py0, wz = pz_cont - (float)pz0;
    px0 = std::max(1, px0); py0 = std::max(1, py0); pz0 = std::max(1, pz0);

    struct V { float v[5]; bool ok; };
    auto get_blended = [&](int pgx, int pgy, int pgz) -> V {
        int pbx = (pgx - 1) / BS + 1, pby = (pgy - 1) / BS + 1, pbz = (pgz - 1) / BS + 1;
        if (pbx >= 1 && pbx <= P.dimx && pby >= 1 && pby <= P.dimy && pbz >= 1 && pbz <= P.dimz) {
            int pb = P.ptr[(pbx - 1) + P.dimx * ((pby - 1) + (size_t)P.dimy * (pbz - 1))];
            if (pb > 0) {
                int plx = (pgx - 1) % BS, ply = (pgy - 1) % BS, plz = (pgz - 1) % BS;  // 0-based local
                size_t c = plx + 8 * ply + 64 * plz + 512 * (size_t)(pb - 1);
                size_t s = 512 * (size_t)P.nb;
                float f_new = P.f_new[c + s * k], rho_new = P.rho_new[c];
                float ux_new = P.vel_new[c], uy_new = P.vel_new[c + s], uz_new = P.vel_new[c + 2 * s];
                if (use_temporal_interp == 1 && temporal_weight < 0.99f) {
                    float f_old = P.f_old[c + s * k], rho_old = P.rho_old[c];
                    float ux_old = P.vel_old[c], uy_old = P.vel_old[c + s], uz_old = P.vel_old[c + 2 * s];
                    float tw = temporal_weight;
                    return V{{f_old * (1.0f - tw) + f_new * tw, rho_old * (1.0f - tw) + rho_new * tw,
                              ux_old * (1.0f - tw) + ux_new * tw, uy_old * (1.0f - tw) + uy_new * tw,
                              uz_old * (1.0f - tw) + uz_new * tw}, true};
                }
                return V{{f_new, rho_new, ux_new, uy_new, uz_new}, true};
            }
        }
        return V{{w_k, 1.0f, 0.0f, 0.0f, 0.0f}, false};
    };
    V d000 = get_blended(px0, py0, pz0), d100 = get_blended(px1, py0, pz0), d010 = get_blended(px0, py1, pz0),
      d110 = get_blended(px1, py1, pz0), d001 = get_blended(px0, py0, pz1), d101 = get_blended(px1, py0, pz1),
      d011 = get_blended(px0, py1, pz1), d111 = get_blended(px1, py1, pz1);
    const V& v000 = d000;
    const V& v100 = d100.ok ? d100 : v000; const V& v010 = d010.ok ? d010 : v000;
    const V& v110 = d110.ok ? d110 : v000; const V& v001 = d001.ok ? d001 : v000;
    const V& v101 = d101.ok ? d101 : v000; const V& v011 = d011.ok ? d011 : v000;
    const V& v111 = d111.ok ? d111 : v000;
    auto trilin = [&](int i) {
        float c00 = v000.v[i] * (1.0f - wx) + v100.v[i] * wx;
        float c01 = v001.v[i] * (1.0f - wx) + v101.v[i] * wx;
        float c10 = v010.v[i] * (1.0f - wx) + v110.v[i] * wx;
        float c11 = v011.v[i] * (1.0f - wx) + v111.v[i] * wx;
        float c0 = c00 * (1.0f - wy) + c10 * wy;
        float c1 = c01 * (1.0f - wy) + c11 * wy;
        return c0 * (1.0f - wz) + c1 * wz;
    };
    float f_int = trilin(0), rho_int = trilin(1), ux_int = trilin(2), uy_int = trilin(3), uz_int = trilin(4);
    float feq_int = calculate_equilibrium(rho_int, ux_int, uy_int, uz_int, w_k, cx, cy, cz);
    float f_neq = f_int - feq_int;
    float tau_c = tau_coarse - 0.5f, tau_f = tau_fine - 0.5f;
    float scale = tau_c > 1.0e-6f ? std::min(std::max(tau_f / tau_c, 0.01f), 100.0f) : 1.0f;
    return feq_int + f_neq * scale;
}

// physics_utils.jl:45-70 (0-based local coords here)
inline void get_velocity_neighbor(const Level& L, const float* vel_in, int x, int y, int z, int b, int dx, int dy,
                                  int dz, float& ux, float& uy, float& uz) {
    int nx = x + dx, ny = y + dy, nz = z + dz;
    size_t s = 512 * (size_t)L.nb;
    if (nx >= 0 && nx < BS && ny >= 0 && ny < BS && nz >= 0 && nz < BS) {
        size_t c = L.idx4(nx, ny, nz, b);
        ux = vel_in[c]; uy = vel_in[c + s]; uz = vel_in[c + 2 * s];
        return;
    }
    int ox = nx < 0 ? -1 : (nx >= BS ? 1 : 0), oy = ny < 0 ? -1 : (ny >= BS ? 1 : 0), oz = nz < 0 ? -1 : (nz >= BS ? 1 : 0);
    int dir = (ox + 1) + (oy + 1) * 3 + (oz + 1) * 9;
    int nbi = L.nbr(b, dir);
    if (nbi > 0) {
        int nnx = nx < 0 ? nx + BS : (nx >= BS ? nx - BS : nx);
        int nny = ny < 0 ? ny + BS : (ny >= BS ? ny - BS : ny);
        int nnz = nz < 0 ? nz + BS : (nz >= BS ? nz - BS : nz);
        size_t c = L.idx4(nnx, nny, nnz, nbi - 1);
        ux = vel_in[c]; uy = vel_in[c + s]; uz = vel_in[c + 2 * s];
        return;
    }
    size_t c = L.idx4(x, y, z, b);
    ux = vel_in[c]; uy = vel_in[c + s]; uz = vel_in[c + 2 * s];
}

struct StepArgs {
    float tau_parent, c_wale, nu_bg, u_inlet, inlet_turbulence, temporal_weight;
    int is_level_1, is_symmetric, nx_g, ny_g, nz_g, wall_model_active, time_step_seed, store_post,
        use_temporal, sponge_blend;
};

// physics_kernels.jl:39-357, one (x,y,z,b) work-item
void stream_collide_cell(Level& L, const Parent& P, const StepArgs& A, float* f_out, const float* f_in, float* vel_out,
                         const float* vel_in, int x, int y, int z, int b) {
    const size_t s = 512 * (size_t)L.nb;
    const float tau_molecular = L.tau;
    int gx = (L.map_x[b] - 1) * BS + (x + 1);  // 1-based global cell coords, as the reference
    int gy = (L.map_y[b] - 1) * BS + (y + 1);
    int gz = (L.map_z[b] - 1) * BS + (z + 1);
    const size_t cell = L.idx4(x, y, z, b);
    bool is_obs = L.obstacle[cell] != 0;
    float rho = 0.0f, jx = 0.0f, jy = 0.0f, jz = 0.0f;
    float f_stored[27];

    for (int k = 0; k < 27; ++k) {
        int cx = LAT.cx[k], cy = LAT.cy[k], cz = LAT.cz[k];
        int sx = x - cx, sy = y - cy, sz = z - cz;
        float val = 0.0f;
        if (sx >= 0 && sx < BS && sy >= 0 && sy < BS && sz >= 0 && sz < BS) {
            val = f_in[L.idx5(sx, sy, sz, b, k)];
        } else {
            int ox = sx < 0 ? -1 : (sx >= BS ? 1 : 0), oy = sy < 0 ? -1 : (sy >= BS ? 1 : 0), oz = sz < 0 ? -1 : (sz >= BS ? 1 : 0);
            int dir = (ox + 1) + (oy + 1) * 3 + (oz + 1) * 9;
            int nbi = L.nbr(b, dir);
            if (nbi > 0) {
                int nsx = sx < 0 ? sx + BS : (sx >= BS ? sx - BS : sx);
                int nsy = sy < 0 ? sy + BS : (sy >= BS ? sy - BS : sy);
                int nsz = sz < 0 ? sz + BS : (sz >= BS ? sz - BS : sz);
                val = f_in[L.idx5(nsx, nsy, nsz, nbi - 1, k)];
            } else {
                int src_gx = gx - cx, src_gy = gy - cy, src_gz = gz - cz;
                bool is_inlet = src_gx < 1, is_outlet = src_gx > A.nx_g;
                bool is_y_min = src_gy < 1, is_y_max = src_gy > A.ny_g;
                bool is_z_min = src_gz < 1, is_z_max = src_gz > A.nz_g;
                if (is_inlet) {
                    float noise = A.inlet_turbulence > 0.0f
                                      ? gradient_noise(gy, gz, A.time_step_seed, 1234) * A.inlet_turbulence * A.u_inlet
                                      : 0.0f;
                    float u_inst = A.u_inlet + noise;
                    float cu_in = (float)cx * u_inst;
                    val = LAT.w[k] * (1.0f + 3.0f * cu_in + 4.5f * cu_in * cu_in - 1.5f * u_inst * u_inst);
                } else if (is_outlet) {
                    float cu_out = (float)cx * A.u_inlet;
                    val = LAT.w[k] * (1.0f + 3.0f * cu_out + 4.5f * cu_out * cu_out - 1.5f * A.u_inlet * A.u_inlet);
                } else if (is_y_min && A.is_symmetric == 1) {
                    val = f_in[L.idx5(x, y, z, b, LAT.mirror_y[k])];
                } else if (is_y_min || is_y_max) {
                    val = f_in[L.idx5(x, y, z, b, LAT.mirror_y[k])];
                } else if (is_z_min || is_z_max) {
                    val = f_in[L.idx5(x, y, z, b, LAT.mirror_z[k])];
                } else if (A.is_level_1 == 0) {
                    val = interpolate_with_rescaling(P, src_gx, src_gy, src_gz, k, LAT.w[k], (float)cx, (float)cy, (float)cz,
                                                     A.tau_parent, tau_molecular, A.temporal_weight, A.use_temporal);
                } else {
                    val = LAT.w[k];
                }
            }
        }
        f_stored[k] = val;
        rho += val;
        jx += val * (float)cx;
        jy += val * (float)cy;
        jz += val * (float)cz;
    }

    if (is_obs) {
        vel_out[cell] = 0.0f; vel_out[cell + s] = 0.0f; vel_out[cell + 2 * s] = 0.0f;
        L.rho[cell] = 1.0f;
        for (int k = 0; k < 27; ++k) {
            float f_coll = f_stored[LAT.opp[k]];
            f_out[cell + s * k] = f_coll;
            if (A.store_post == 1) L.f_post[cell + s * k] = f_coll;
        }
        return;
    }

    rho = std::max(rho, 0.01f);
    float inv_rho = 1.0f / rho;
    float ux = jx * inv_rho, uy = jy * inv_rho, uz = jz * inv_rho;

    float sp = L.sponge[cell];
    if (sp > 0.0f) {
        float rho_target = 1.0f, ux_target = A.u_inlet;
        rho = rho * (1.0f - sp) + rho_target * sp;
        ux = ux * (1.0f - sp) + ux_target * sp;
        uy = uy * (1.0f - sp);
        uz = uz * (1.0f - sp);
        if (A.sponge_blend == 1) {
            for (int k = 0; k < 27; ++k) {
                float feq_target = calculate_equilibrium(rho_target, ux_target, 0.0f, 0.0f, LAT.w[k], (float)LAT.cx[k],
                                                         (float)LAT.cy[k], (float)LAT.cz[k]);
                f_stored[k] = f_stored[k] * (1.0f - sp) + feq_target * sp;
            }
        }
    }

    float Fx_wall = 0.0f, Fy_wall = 0.0f, Fz_wall = 0.0f;
    if (A.wall_model_active == 1) {
        float dist_wall = L.wall_dist[cell];
        if (dist_wall > 0.0f && dist_wall < 10.0f) {
            float u_mag = std::sqrt(ux * ux + uy * uy + uz * uz);
            float nu_visc = (tau_molecular - 0.5f) / 3.0f;
            if (u_mag > 1.0e-6f && nu_visc > 1.0e-10f) {
                float u_tau = u_mag * pow32(nu_visc / (dist_wall * u_mag + 1.0e-10f), 1.0f / 7.0f) *
                              pow32(2.0f * 8.3f, -1.0f / 7.0f);
                u_tau = std::max(u_tau, 1.0e-6f);
                float y_p = u_tau * dist_wall / nu_visc;
                if (y_p > 11.81f) {
                    float u_plus_law = (1.0f / KAPPA) * log32(y_p) + 5.2f;
                    if (u_plus_law > 0.1f) {
                        u_tau = u_tau * ((u_mag / u_tau) / u_plus_law);
                        u_tau = std::max(u_tau, 1.0e-6f);
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

    vel_out[cell] = ux; vel_out[cell + s] = uy; vel_out[cell + 2 * s] = uz;
    L.rho[cell] = rho;

    float uE[3], uW[3], uN[3], uS[3], uT[3], uB[3];
    get_velocity_neighbor(L, vel_in, x, y, z, b, 1, 0, 0, uE[0], uE[1], uE[2]);
    get_velocity_neighbor(L, vel_in, x, y, z, b, -1, 0, 0, uW[0], uW[1], uW[2]);
    get_velocity_neighbor(L, vel_in, x, y, z, b, 0, 1, 0, uN[0], uN[1], uN[2]);
    get_velocity_neighbor(L, vel_in, x, y, z, b, 0, -1, 0, uS[0], uS[1], uS[2]);
    get_velocity_neighbor(L, vel_in, x, y, z, b, 0, 0, 1, uT[0], uT[1], uT[2]);
    get_velocity_neighbor(L, vel_in, x, y, z, b, 0, 0, -1, uB[0], uB[1], uB[2]);
    float g11 = 0.5f * (uE[0] - uW[0]), g12 = 0.5f * (uN[0] - uS[0]), g13 = 0.5f * (uT[0] - uB[0]);
    float g21 = 0.5f * (uE[1] - uW[1]), g22 = 0.5f * (uN[1] - uS[1]), g23 = 0.5f * (uT[1] - uB[1]);
    float g31 = 0.5f * (uE[2] - uW[2]), g32 = 0.5f * (uN[2] - uS[2]), g33 = 0.5f * (uT[2] - uB[2]);

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
        float OP1_32 = OP1 * std::sqrt(OP1);
        float OP2_52 = OP2 * OP2 * std::sqrt(std::max(OP2, 1.0e-12f));
        float denom = OP2_52 + OP1 * std::sqrt(std::sqrt(std::max(OP1, 1.0e-12f)));
        if (denom > 1.0e-12f) nu_eddy = (A.c_wale * A.c_wale) * OP1_32 / denom;
    }
    nu_eddy = std::max(nu_eddy, A.nu_bg);
    float tau_turb = tau_molecular + nu_eddy * 3.0f;
    float omega = 1.0f / std::max(tau_turb, 0.500001f);

    float Pi_xx = 0, Pi_yy = 0, Pi_zz = 0, Pi_xy = 0, Pi_yz = 0, Pi_zx = 0;
    for (int k = 0; k < 27; ++k) {
        float cx_f = (float)LAT.cx[k], cy_f = (float)LAT.cy[k], cz_f = (float)LAT.cz[k];
        float cu = cx_f * ux_eq + cy_f * uy_eq + cz_f * uz_eq;
        float feq = rho * LAT.w[k] * (1.0f + 3.0f * cu + 4.5f * cu * cu - 1.5f * usq_eq);
        float f_neq = f_stored[k] - feq;
        Pi_xx += f_neq * cx_f * cx_f;
        Pi_yy += f_neq * cy_f * cy_f;
        Pi_zz += f_neq * cz_f * cz_f;
        Pi_xy += f_neq * cx_f * cy_f;
        Pi_yz += f_neq * cy_f * cz_f;
        Pi_zx += f_neq * cz_f * cx_f;
    }
    for (int k = 0; k < 27; ++k) {
        float cx_f = (float)LAT.cx[k], cy_f = (float)LAT.cy[k], cz_f = (float)LAT.cz[k];
        float w_k = LAT.w[k];
        float cu = cx_f * ux_eq + cy_f * uy_eq + cz_f * uz_eq;
        float feq = rho * w_k * (1.0f + 3.0f * cu + 4.5f * cu * cu - 1.5f * usq_eq);
        float force_term = w_k * 3.0f *
                           ((cx_f - ux + 3.0f * cu * cx_f) * Fx_wall + (cy_f - uy + 3.0f * cu * cy_f) * Fy_wall +
                            (cz_f - uz + 3.0f * cu * cz_f) * Fz_wall);
        float Q_xx = cx_f * cx_f - CS2_PHYSICS, Q_yy = cy_f * cy_f - CS2_PHYSICS, Q_zz = cz_f * cz_f - CS2_PHYSICS;
        float f_neq_reg = w_k * 4.5f *
                          (Pi_xx * Q_xx + Pi_yy * Q_yy + Pi_zz * Q_zz +
                           2.0f * (Pi_xy * cx_f * cy_f + Pi_yz * cy_f * cz_f + Pi_zx * cz_f * cx_f));
        float f_coll = feq + (1.0f - omega) * f_neq_reg + (1.0f - 0.5f * omega) * force_term;
        if (A.store_post == 1) L.f_post[cell + s * k] = f_coll;
        f_out[cell + s * k] = f_coll;
    }
}

// bouzidi_kernel.jl:27-91
void bouzidi_correction(Level& L, float* f_out, float q_min_threshold) {
    const size_t s = 512 * (size_t)L.nb;
    const float* f_post = L.f_post.data();
#pragma omp parallel for schedule(static)
    for (int ci = 0; ci < L.n_bc; ++ci) {
        int b = L.cell_block[ci] - 1, x = L.cell_x[ci] - 1, y = L.cell_y[ci] - 1, z = L.cell_z[ci] - 1;
        size_t cell = L.idx4(x, y, z, b);
        for (int k = 0; k < 27; ++k) {
            float q = half_to_float(L.q_map[cell + s * k]);
            if (q > q_min_threshold && q <= 1.0f) {
                int opp_k = LAT.opp[k];
                float f_k = f_post[cell + s * k];
                if (q < 0.5f) {
                    int nx = x + LAT.cx[opp_k], ny = y + LAT.cy[opp_k], nz = z + LAT.cz[opp_k];
                    float f_ff = f_k;
                    if (nx >= 0 && nx < BS && ny >= 0 && ny < BS && nz >= 0 && nz < BS) {
                        f_ff = f_post[L.idx5(nx, ny, nz, b, k)];
                    } else {
                        int ox = nx < 0 ? -1 : (nx >= BS ? 1 : 0), oy = ny < 0 ? -1 : (ny >= BS ? 1 : 0),
                            oz = nz < 0 ? -1 : (nz >= BS ? 1 : 0);
                        int dir = (ox + 1) + (oy + 1) * 3 + (oz + 1) * 9;
                        int nbi = L.nbr(b, dir);
                        if (nbi > 0) {
                            int nnx = nx < 0 ? nx + BS : (nx >= BS ? nx - BS : nx);
                            int nny = ny < 0 ? ny + BS : (ny >= BS ? ny - BS : ny);
                            int nnz = nz < 0 ? nz + BS : (nz >= BS ? nz - BS : nz);
                            f_ff = f_post[L.idx5(nnx, nny, nnz, nbi - 1, k)];
                        }
                    }
                    float coeff1 = 2.0f * q;
                    f_out[cell + s * opp_k] = coeff1 * f_k + (1.0f - coeff1) * f_ff;
                } else {
                    float f_opp_post = f_post[cell + s * opp_k];
                    float inv_2q = 1.0f / (2.0f * q);
                    float coeff2 = (2.0f * q - 1.0f) * inv_2q;
                    f_out[cell + s * opp_k] = inv_2q * f_k + coeff2 * f_opp_post;
                }
            }
        }
    }
}

// physics_v2.jl:26-97
void perform_timestep(Level& L, const Parent* parent, float parent_tau, float* f_out, const float* f_in, float* vel_out,
                      const float* vel_in, float u_curr, const ludwig_params& p, int64_t timestep, float temporal_weight) {
    if (L.nb == 0) return;
    Parent P{};
    StepArgs A{};
    A.is_level_1 = parent == nullptr ? 1 : 0;
    if (parent) P = *parent;
    int scale = 1 << (L.level_id - 1);
    A.nx_g = p.domain_nx * scale; A.ny_g = p.domain_ny * scale; A.nz_g = p.domain_nz * scale;
    A.tau_parent = parent_tau; A.c_wale = p.c_wale; A.nu_bg = p.nu_sgs_bg; A.u_inlet = u_curr;
    A.inlet_turbulence = p.inlet_turbulence; A.temporal_weight = temporal_weight;
    A.is_symmetric = p.symmetric; A.wall_model_active = p.wall_model_active;
    A.time_step_seed = (int32_t)(timestep % 1000000);
    A.store_post = (L.bouzidi && L.n_bc > 0) ? 1 : 0;
    A.use_temporal = p.use_temporal; A.sponge_blend = p.sponge_blend;
#pragma omp parallel for schedule(dynamic, 8)
    for (int b = 0; b < L.nb; ++b)
        for (int z = 0; z < BS; ++z)
            for (int y = 0; y < BS; ++y)
                for (int x = 0; x < BS; ++x) stream_collide_cell(L, P, A, f_out, f_in, vel_out, vel_in, x, y, z, b);
    if (L.bouzidi && L.n_bc > 0) bouzidi_correction(L, f_out, p.q_min_threshold);
}

// blocks.jl:199-205
void copy_to_old(Level& L, const float* f_current, const float* vel_current) {
    if (!L.temporal) return;
    std::memcpy(L.f_old.data(), f_current, L.f_old.size() * 4);
    std::memcpy(L.rho_old.data(), L.rho.data(), L.rho_old.size() * 4);
    std::memcpy(L.vel_old.data(), vel_current, L.vel_old.size() * 4);
}

// solver_control.jl:21-143 (recursive_step! and recursive_step_temporal! share one body)
void recursive_step(ludwig_ctx* ctx, size_t lvl, int64_t t_sub, const Parent* parent, float parent_tau,
                    float temporal_weight, float u_vel, const ludwig_params& p) {
    if (lvl >= ctx->levels.size()) return;
    Level& L = *ctx->levels[lvl];
    float *f_in, *f_out, *vel_in, *vel_out;
    if (t_sub % 2 == 0) { f_in = L.f.data(); f_out = L.f_temp.data(); vel_in = L.vel.data(); vel_out = L.vel_temp.data(); }
    else { f_in = L.f_temp.data(); f_out = L.f.data(); vel_in = L.vel_temp.data(); vel_out = L.vel.data(); }
    bool has_children = lvl + 1 < ctx->levels.size();
    if (has_children && p.use_temporal && L.temporal) copy_to_old(L, f_in, vel_in);
    perform_timestep(L, parent, parent_tau, f_out, f_in, vel_out, vel_in, u_vel, p, t_sub, temporal_weight);
    if (has_children) {
        Parent me{f_out, L.rho.data(), vel_out, L.f_old.data(), L.rho_old.data(), L.vel_old.data(),
                  L.block_pointer.data(), L.dimx, L.dimy, L.dimz, L.nb, L.tau};
        recursive_step(ctx, lvl + 1, 2 * t_sub, &me, L.tau, 0.0f, u_vel, p);
        recursive_step(ctx, lvl + 1, 2 * t_sub + 1, &me, L.tau, 0.5f, u_vel, p);
    }
}

int fail(ludwig_ctx* ctx, int code, const std::string& msg) { if (ctx) ctx->err = msg; return code; }

}  // namespace

extern "C" {

const char* ludwig_backend_name(void) { return "cpu-oracle"; }

int ludwig_ctx_create(ludwig_ctx** out, int /*device*/) {
    if (!out) return LUDWIG_EINVAL;
    *out = new ludwig_ctx();
    return LUDWIG_OK;
}
int ludwig_ctx_destroy(ludwig_ctx* ctx) { delete ctx; return LUDWIG_OK; }
const char* ludwig_last_error(const ludwig_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int ludwig_sync(ludwig_ctx*) { return LUDWIG_OK; }
int ludwig_num_levels(const ludwig_ctx* ctx) { return ctx ? (int)ctx->levels.size() : LUDWIG_EINVAL; }

// blocks.jl:89-188 constructor allocation rules + domain.jl:238-240 field copies
int ludwig_level_create(ludwig_ctx* ctx, const ludwig_level_desc* d, int32_t* out_index) {
    if (!ctx || !d) return LUDWIG_EINVAL;
    if (d->level_id != (int)ctx->levels.size() + 1) return fail(ctx, LUDWIG_ESTATE, "levels must be created in order 1..L");
    if (d->n_blocks <= 0) return fail(ctx, LUDWIG_EINVAL, "n_blocks must be > 0");
    auto L = std::make_unique<Level>();
    L->level_id = d->level_id; L->nb = d->n_blocks; L->dimx = d->dim_x; L->dimy = d->dim_y; L->dimz = d->dim_z;
    L->tau = d->tau; L->dx = d->dx;
    size_t nb = d->n_blocks, nc = nb * 512;
    L->block_pointer.assign(d->block_pointer, d->block_pointer + (size_t)d->dim_x * d->dim_y * d->dim_z);
    L->neighbor_table.assign(d->neighbor_table, d->neighbor_table + nb * 27);
    L->map_x.assign(d->map_x, d->map_x + nb); L->map_y.assign(d->map_y, d->map_y + nb); L->map_z.assign(d->map_z, d->map_z + nb);
    L->obstacle.assign(d->obstacle, d->obstacle + nc);
    L->sponge.assign(d->sponge, d->sponge + nc);
    L->wall_dist.assign(d->wall_dist, d->wall_dist + nc);
    L->temporal = d->temporal_storage != 0;
    L->bouzidi = d->bouzidi_enabled != 0 && d->n_boundary_cells > 0 && d->q_map_f16 != nullptr;
    L->n_bc = L->bouzidi ? d->n_boundary_cells : 0;
    L->rho.assign(nc, 1.0f); L->vel.assign(nc * 3, 0.0f); L->vel_temp.assign(nc * 3, 0.0f);
    L->f.assign(nc * 27, 0.0f); L->f_temp.assign(nc * 27, 0.0f);
    if (L->temporal) { L->rho_old.assign(nc, 1.0f); L->vel_old.assign(nc * 3, 0.0f); L->f_old.assign(nc * 27, 0.0f); }
    if (L->n_bc > 0) {
        L->f_post.assign(nc * 27, 0.0f);
        L->q_map.assign(d->q_map_f16, d->q_map_f16 + nc * 27);
        L->cell_block.assign(d->cell_block, d->cell_block + L->n_bc);
        L->cell_x.assign(d->cell_x, d->cell_x + L->n_bc);
        L->cell_y.assign(d->cell_y, d->cell_y + L->n_bc);
        L->cell_z.assign(d->cell_z, d->cell_z + L->n_bc);
    }
    ctx->levels.push_back(std::move(L));
    if (out_index) *out_index = (int32_t)ctx->levels.size() - 1;
    return LUDWIG_OK;
}

static std::vector<float>* field_f32(Level& L, int which) {
    switch (which) {
        case LUDWIG_F: return &L.f; case LUDWIG_F_TEMP: return &L.f_temp; case LUDWIG_F_POST: return &L.f_post;
        case LUDWIG_F_OLD: return &L.f_old; case LUDWIG_RHO: return &L.rho; case LUDWIG_RHO_OLD: return &L.rho_old;
        case LUDWIG_VEL: return &L.vel; case LUDWIG_VEL_TEMP: return &L.vel_temp; case LUDWIG_VEL_OLD: return &L.vel_old;
        default: return nullptr;
    }
}
int ludwig_level_upload(ludwig_ctx* ctx, int32_t level, int32_t which, const void* src) {
    if (!ctx || level < 0 || level >= (int)ctx->levels.size() || !src) return fail(ctx, LUDWIG_EINVAL, "bad level/src");
    Level& L = *ctx->levels[level];
    if (which == LUDWIG_OBSTACLE) { std::memcpy(L.obstacle.data(), src, L.obstacle.size()); return LUDWIG_OK; }
    auto* v = field_f32(L, which);
    if (!v || v->empty()) return fail(ctx, LUDWIG_EINVAL, "field not allocated on this level");
    std::memcpy(v->data(), src, v->size() * 4);
    return LUDWIG_OK;
}
int ludwig_level_download(ludwig_ctx* ctx, int32_t level, int32_t which, void* dst) {
    if (!ctx || level < 0 || level >= (int)ctx->levels.size() || !dst) return fail(ctx, LUDWIG_EINVAL, "bad level/dst");
    Level& L = *ctx->levels[level];
    if (which == LUDWIG_OBSTACLE) { std::memcpy(dst, L.obstacle.data(), L.obstacle.size()); return LUDWIG_OK; }
    auto* v = field_f32(L, which);
    if (!v || v->empty()) return fail(ctx, LUDWIG_EINVAL, "field not allocated on this level");
    std::memcpy(dst, v->data(), v->size() * 4);
    return LUDWIG_OK;
}

// io_vtk.jl:52-58,100-111 restated: whole-array semantics, then the listed blocks.
int ludwig_output_gather(ludwig_ctx* ctx, int32_t level, int64_t t_step, const int32_t* blocks, int32_t n_blocks,
                         float* rho_arr, float* vel_mat, uint8_t* obst_arr) {
    if (!ctx || level < 0 || level >= (int)ctx->levels.size() || !blocks || n_blocks < 0 || !rho_arr || !vel_mat || !obst_arr)
        return fail(ctx, LUDWIG_EINVAL, "bad gather args");
    Level& L = *ctx->levels[level];
    const std::vector<float>& vel = (t_step % 2 == 0) ? L.vel_temp : L.vel;   // :56
    auto clean = [](float v) { return std::isfinite(v) ? v : 0.0f; };         // :110-111
    for (int i = 0; i < n_blocks; ++i) {
        const int b = blocks[i] - 1;
        if (b < 0 || b >= L.nb) return fail(ctx, LUDWIG_EINVAL, "gather: block index out of range");
        for (int c = 0; c < 512; ++c) {
            const size_t o = (size_t)i * 512 + c, s = (size_t)b * 512 + c;
            rho_arr[o] = clean(L.rho[s]);
            for (int k = 0; k < 3; ++k) vel_mat[o * 3 + k] = clean(vel[s + 512 * (size_t)L.nb * k]);
            obst_arr[o] = L.obstacle[s] ? 1 : 0;
        }
    }
    return LUDWIG_OK;
}

// io_vtk.jl:17-46 restated: a block is exported unless all 8 of its children (2 bx - 1 + dbx, ... in 1-based coordinates) are
// active blocks of the next finer level; level-major, b_idx ascending (the order of `valid_blocks`).
static std::vector<std::vector<int32_t>> valid_blocks_of(const ludwig_ctx* ctx) {
    std::vector<std::vector<int32_t>> out(ctx->levels.size());
    for (size_t l = 0; l < ctx->levels.size(); ++l) {
        const Level& L = *ctx->levels[l];
        const Level* N = l + 1 < ctx->levels.size() ? ctx->levels[l + 1].get() : nullptr;
        for (int b = 0; b < L.nb; ++b) {
            bool should_export = true;
            if (N) {
                int children = 0;
                for (int dbz = 0; dbz < 2; ++dbz) for (int dby = 0; dby < 2; ++dby) for (int dbx = 0; dbx < 2; ++dbx) {
                    const int x = 2 * L.map_x[b] - 1 + dbx, y = 2 * L.map_y[b] - 1 + dby, z = 2 * L.map_z[b] - 1 + dbz;   // 1-based
                    if (x <= N->dimx && y <= N->dimy && z <= N->dimz && N->block_pointer[(x - 1) + (size_t)N->dimx * ((y - 1) + (size_t)N->dimy * (z - 1))] > 0) ++children;
                }
                if (children == 8) should_export = false;
            }
            if (should_export) out[l].push_back(b + 1);
        }
    }
    return out;
}
int ludwig_output_valid_blocks(ludwig_ctx* ctx, int32_t* n_valid, int32_t* blocks) {
    if (!ctx || !n_valid) return fail(ctx, LUDWIG_EINVAL, "bad valid-block args");
    auto v = valid_blocks_of(ctx);
    size_t o = 0;
    for (size_t l = 0; l < v.size(); ++l) {
        n_valid[l] = (int32_t)v[l].size();
        if (blocks) { std::memcpy(blocks + o, v[l].data(), v[l].size() * sizeof(int32_t)); o += v[l].size(); }
    }
    return LUDWIG_OK;
}
// io_vtk.jl:52-58,100-111 for all of them
int ludwig_output_export(ludwig_ctx* ctx, int64_t t_step, float* rho_arr, float* vel_mat, uint8_t* obst_arr, int32_t* level_arr) {
    if (!ctx || !rho_arr || !vel_mat || !obst_arr) return fail(ctx, LUDWIG_EINVAL, "bad export args");
    auto v = valid_blocks_of(ctx);
    size_t base = 0;
    for (size_t l = 0; l < v.size(); ++l) {
        const size_t o = base * 512;
        int rc = v[l].empty() ? LUDWIG_OK : ludwig_output_gather(ctx, (int32_t)l, t_step, v[l].data(), (int32_t)v[l].size(), rho_arr + o, vel_mat + 3 * o, obst_arr + o);
        if (rc) return rc;
        if (level_arr) std::fill(level_arr + o, level_arr + o + v[l].size() * 512, (int32_t)ctx->levels[l]->level_id);
        base += v[l].size();
    }
    return LUDWIG_OK;
}

int ludwig_mesh_create(ludwig_ctx* ctx, int32_t n, const float* cx, const float* cy, const float* cz, const float* nx,
                       const float* ny, const float* nz, const float* area, ludwig_mesh** out) {
    if (!ctx || n <= 0 || !out) return fail(ctx, LUDWIG_EINVAL, "bad mesh");
    auto* m = new ludwig_mesh();
    m->n = n;
    m->cx.assign(cx, cx + n); m->cy.assign(cy, cy + n); m->cz.assign(cz, cz + n);
    m->nx.assign(nx, nx + n); m->ny.assign(ny, ny + n); m->nz.assign(nz, nz + n);
    m->area.assign(area, area + n);
    *out = m;
    return LUDWIG_OK;
}
int ludwig_mesh_destroy(ludwig_mesh* m) { delete m; return LUDWIG_OK; }

int ludwig_forces_create(ludwig_ctx* ctx, const ludwig_mesh* mesh, double rho_ref, double u_ref, double area_ref,
                         double chord_ref, const double mc[3], int32_t symmetric, ludwig_forces** out) {
    if (!ctx || !mesh || !out) return fail(ctx, LUDWIG_EINVAL, "bad forces args");
    auto* f = new ludwig_forces();
    f->mesh = mesh; f->rho_ref = rho_ref; f->u_ref = u_ref; f->area_ref = area_ref; f->chord_ref = chord_ref;
    f->mc[0] = mc[0]; f->mc[1] = mc[1]; f->mc[2] = mc[2]; f->symmetric = symmetric;
    f->p.assign(mesh->n, 0.0f); f->sx.assign(mesh->n, 0.0f); f->sy.assign(mesh->n, 0.0f); f->sz.assign(mesh->n, 0.0f);
    *out = f;
    return LUDWIG_OK;
}
int ludwig_forces_destroy(ludwig_forces* f) { delete f; return LUDWIG_OK; }

// main.jl:109-135
int ludwig_init_equilibrium(ludwig_ctx* ctx) {
    if (!ctx) return LUDWIG_EINVAL;
    for (auto& Lp : ctx->levels) {
        Level& L = *Lp;
        size_t s = 512 * (size_t)L.nb;
        for (int k = 0; k < 27; ++k)
            for (size_t c = 0; c < s; ++c) {
                L.f[c + s * k] = LAT.w[k];
                L.f_temp[c + s * k] = LAT.w[k];
                if (L.temporal) L.f_old[c + s * k] = LAT.w[k];
            }
        if (L.temporal) { std::fill(L.rho_old.begin(), L.rho_old.end(), 1.0f); std::fill(L.vel_old.begin(), L.vel_old.end(), 0.0f); }
    }
    return LUDWIG_OK;
}

// solver_control.jl:145-165
int ludwig_step_batch(ludwig_ctx* ctx, int64_t t_start, int32_t batch_size, float u_curr, const ludwig_params* params) {
    if (!ctx || !params || ctx->levels.empty()) return fail(ctx, LUDWIG_EINVAL, "bad step args");
    for (int t_offset = 0; t_offset < batch_size; ++t_offset)
        recursive_step(ctx, 0, t_start + t_offset, nullptr, 0.5f, 0.0f, u_curr, *params);
    return LUDWIG_OK;
}

int ludwig_level_snapshot_old(ludwig_ctx* ctx, int32_t level, int64_t t_sub) {
    if (!ctx || level < 0 || level >= (int)ctx->levels.size()) return fail(ctx, LUDWIG_EINVAL, "bad level");
    Level& L = *ctx->levels[level];
    if (t_sub % 2 == 0) copy_to_old(L, L.f.data(), L.vel.data());
    else copy_to_old(L, L.f_temp.data(), L.vel_temp.data());
    return LUDWIG_OK;
}

int ludwig_level_step(ludwig_ctx* ctx, int32_t level, int64_t t_sub, int64_t parent_t_sub, float temporal_weight,
                      float u_curr, const ludwig_params* params) {
    if (!ctx || !params || level < 0 || level >= (int)ctx->levels.size()) return fail(ctx, LUDWIG_EINVAL, "bad level");
    Level& L = *ctx->levels[level];
    float *f_in, *f_out, *vel_in, *vel_out;
    if (t_sub % 2 == 0) { f_in = L.f.data(); f_out = L.f_temp.data(); vel_in = L.vel.data(); vel_out = L.vel_temp.data(); }
    else { f_in = L.f_temp.data(); f_out = L.f.data(); vel_in = L.vel_temp.data(); vel_out = L.vel.data(); }
    if (level == 0) {
        perform_timestep(L, nullptr, 0.5f, f_out, f_in, vel_out, vel_in, u_curr, *params, t_sub, temporal_weight);
    } else {
        Level& Pl = *ctx->levels[level - 1];
        bool even = parent_t_sub % 2 == 0;
        Parent me{even ? Pl.f_temp.data() : Pl.f.data(), Pl.rho.data(), even ? Pl.vel_temp.data() : Pl.vel.data(),
                  Pl.f_old.data(), Pl.rho_old.data(), Pl.vel_old.data(), Pl.block_pointer.data(),
                  Pl.dimx, Pl.dimy, Pl.dimz, Pl.nb, Pl.tau};
        if (params->use_temporal && !Pl.temporal) return fail(ctx, LUDWIG_ESTATE, "parent has no temporal storage");
        perform_timestep(L, &me, Pl.tau, f_out, f_in, vel_out, vel_in, u_curr, *params, t_sub, temporal_weight);
    }
    return LUDWIG_OK;
}

// forces/surface.jl:138-266 (K3), :282-366 (K4), :467-572 (host math)
int ludwig_compute_aerodynamics(ludwig_ctx* ctx, ludwig_forces* F, int32_t level, const double mesh_offset[3],
                                double velocity_scale, double rho_phys, int32_t search_radius, double out[18]) {
    if (!ctx || !F || level < 0 || level >= (int)ctx->levels.size() || !out) return fail(ctx, LUDWIG_EINVAL, "bad aero args");
    Level& L = *ctx->levels[level];
    const ludwig_mesh& M = *F->mesh;
    const float pressure_scale = (float)(rho_phys * velocity_scale * velocity_scale);
    const float stress_scale = pressure_scale;
    const float dx = (float)L.dx;
    const float offx = (float)mesh_offset[0], offy = (float)mesh_offset[1], offz = (float)mesh_offset[2];
    const float tau_molecular = L.tau;
    const size_t s = 512 * (size_t)L.nb;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < M.n; ++i) {
        float tx = M.cx[i] + offx, ty = M.cy[i] + offy, tz = M.cz[i] + offz;
        float n_x = M.nx[i], n_y = M.ny[i], n_z = M.nz[i];
        float gx_f = tx / dx, gy_f = ty / dx, gz_f = tz / dx;
        int g_x = (int)std::floor(gx_f) + 1, g_y = (int)std::floor(gy_f) + 1, g_z = (int)std::floor(gz_f) + 1;
        float best_dist_sq = (float)1e10, best_rho = 1.0f, best_ux = 0, best_uy = 0, best_uz = 0, best_wall_dist = 0.5f;
        bool found = false;
        for (int radius = 0; radius <= search_radius; ++radius) {
            if (found && radius > 1) break;
            for (int dz = -radius; dz <= radius; ++dz)
                for (int dy = -radius; dy <= radius; ++dy)
                    for (int ddx = -radius; ddx <= radius; ++ddx) {
                        if (radius > 0) {
                            bool at_shell = std::abs(ddx) == radius || std::abs(dy) == radius || std::abs(dz) == radius;
                            if (!at_shell) continue;
                        }
                        int cgx = g_x + ddx, cgy = g_y + dy, cgz = g_z + dz;
                        // get_cell_at_position, forces/surface.jl:95-124
                        if (cgx < 1 || cgy < 1 || cgz < 1) continue;
                        int bx = (cgx - 1) / BS + 1, by = (cgy - 1) / BS + 1, bz = (cgz - 1) / BS + 1;
                        if (bx < 1 || bx > L.dimx || by < 1 || by > L.dimy || bz < 1 || bz > L.dimz) continue;
                        int bi = L.block_pointer[(bx - 1) + L.dimx * ((by - 1) + (size_t)L.dimy * (bz - 1))];
                        if (bi <= 0) continue;
                        int lx = (cgx - 1) % BS, ly = (cgy - 1) % BS, lz = (cgz - 1) % BS;
                        size_t c = L.idx4(lx, ly, lz, bi - 1);
                        if (L.obstacle[c]) continue;
                        float ccx = ((float)cgx - 0.5f) * dx, ccy = ((float)cgy - 0.5f) * dx, ccz = ((float)cgz - 0.5f) * dx;
                        float ddx_ = tx - ccx, ddy_ = ty - ccy, ddz_ = tz - ccz;
                        float dist_sq = ddx_ * ddx_ + ddy_ * ddy_ + ddz_ * ddz_;
                        if (dist_sq < best_dist_sq) {
                            best_dist_sq = dist_sq;
                            best_rho = L.rho[c];
                            best_ux = L.vel[c]; best_uy = L.vel[c + s]; best_uz = L.vel[c + 2 * s];
                            best_wall_dist = std::sqrt(dist_sq) / dx;
                            found = true;
                        }
                    }
        }
        float p_val = 0, tau_x = 0, tau_y = 0, tau_z = 0;
        if (found) {
            float wall_dist = std::max(best_wall_dist, 0.5f);
            // compute_stress_from_cell, forces/surface.jl:32-89
            float p_gauge_lat = (best_rho - 1.0f) / 3.0f;
            p_val = p_gauge_lat * pressure_scale;
            float u_dot_n = best_ux * n_x + best_uy * n_y + best_uz * n_z;
            float ut_x = best_ux - u_dot_n * n_x, ut_y = best_uy - u_dot_n * n_y, ut_z = best_uz - u_dot_n * n_z;
            float u_tan_mag = std::sqrt(ut_x * ut_x + ut_y * ut_y + ut_z * ut_z);
            float nu_lat = (tau_molecular - 0.5f) / 3.0f;
            if (u_tan_mag > 1.0e-10f && wall_dist > 0.01f) {
                float tau_lat_mag = best_rho * nu_lat * u_tan_mag / wall_dist;
                float tau_phys_mag = tau_lat_mag * stress_scale;
                tau_x = (ut_x / u_tan_mag) * tau_phys_mag;
                tau_y = (ut_y / u_tan_mag) * tau_phys_mag;
                tau_z = (ut_z / u_tan_mag) * tau_phys_mag;
            }
        }
        F->p[i] = p_val; F->sx[i] = tau_x; F->sy[i] = tau_y; F->sz[i] = tau_z;
    }
    // K4: the reference sums 9 scalars with FP32 atomics (order undefined).  The oracle sums in triangle
    // order in FP32 (what a single-threaded KA CPU backend does).
    float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const float refx = (float)F->mc[0], refy = (float)F->mc[1], refz = (float)F->mc[2];
    for (int i = 0; i < M.n; ++i) {
        float p = F->p[i], tau_x = F->sx[i], tau_y = F->sy[i], tau_z = F->sz[i];
        float nx = M.nx[i], ny = M.ny[i], nz = M.nz[i], A = M.area[i];
        float cx = M.cx[i] + offx, cy = M.cy[i] + offy, cz = M.cz[i] + offz;
        float dFp_x = -p * nx * A, dFp_y = -p * ny * A, dFp_z = -p * nz * A;
        float dFv_x = tau_x * A, dFv_y = tau_y * A, dFv_z = tau_z * A;
        float dFx = dFp_x + dFv_x, dFy = dFp_y + dFv_y, dFz = dFp_z + dFv_z;
        float rx = cx - refx, ry = cy - refy, rz = cz - refz;
        float dMx = ry * dFz - rz * dFy, dMy = rz * dFx - rx * dFz, dMz = rx * dFy - ry * dFx;
        acc[0] += dFp_x; acc[1] += dFp_y; acc[2] += dFp_z;
        acc[3] += dFv_x; acc[4] += dFv_y; acc[5] += dFv_z;
        acc[6] += dMx; acc[7] += dMy; acc[8] += dMz;
    }
    double Fx_p = acc[0], Fy_p = acc[1], Fz_p = acc[2], Fx_v = acc[3], Fy_v = acc[4], Fz_v = acc[5];
    double Mx = acc[6], My = acc[7], Mz = acc[8];
    if (F->symmetric) {
        Fx_p *= 2.0; Fz_p *= 2.0; Fx_v *= 2.0; Fz_v *= 2.0; My *= 2.0;
        Fy_p = 0.0; Fy_v = 0.0; Mx = 0.0; Mz = 0.0;
    }
    double Fx = Fx_p + Fx_v, Fy = Fy_p + Fy_v, Fz = Fz_p + Fz_v;
    double q_inf = 0.5 * F->rho_ref * F->u_ref * F->u_ref;
    double F_ref = q_inf * F->area_ref, M_ref = F_ref * F->chord_ref;
    double Cd = 0, Cl = 0, Cs = 0, Cmx = 0, Cmy = 0, Cmz = 0;
    if (F_ref > 1e-10) { Cd = Fx / F_ref; Cl = Fz / F_ref; Cs = Fy / F_ref; }
    if (M_ref > 1e-10) { Cmx = Mx / M_ref; Cmy = My / M_ref; Cmz = Mz / M_ref; }
    double r[18] = {Fx, Fy, Fz, Mx, My, Mz, Fx_p, Fy_p, Fz_p, Fx_v, Fy_v, Fz_v, Cd, Cl, Cs, Cmx, Cmy, Cmz};
    std::memcpy(out, r, sizeof(r));
    return LUDWIG_OK;
}

int ludwig_forces_download_maps(ludwig_ctx*, const ludwig_forces* F, float* p, float* sx, float* sy, float* sz) {
    if (!F) return LUDWIG_EINVAL;
    size_t n = F->p.size() * 4;
    if (p) std::memcpy(p, F->p.data(), n);
    if (sx) std::memcpy(sx, F->sx.data(), n);
    if (sy) std::memcpy(sy, F->sy.data(), n);
    if (sz) std::memcpy(sz, F->sz.data(), n);
    return LUDWIG_OK;
}

// diagnostics.jl:56-94 (the CUDA branch; sums accumulated in double because the reference's
// reduction order is unspecified)
int ludwig_flow_stats(ludwig_ctx* ctx, int32_t level, double out[6]) {
    if (!ctx || level < 0 || level >= (int)ctx->levels.size() || !out) return fail(ctx, LUDWIG_EINVAL, "bad stats args");
    Level& L = *ctx->levels[level];
    size_t s = 512 * (size_t)L.nb;
    double n_fluid = 0, rho_sum = 0, ke = 0;
    float rho_min = INFINITY, rho_max = -INFINITY, v_max = 0.0f;
    bool nan_rho = false, nan_v = false;   // Julia's minimum / maximum propagate NaN (diagnostics.jl:70-77): a blown-up run must not look healthy
    for (size_t c = 0; c < s; ++c) {
        if (L.obstacle[c]) continue;
        n_fluid += 1;
        float r = L.rho[c];
        float v2 = L.vel[c] * L.vel[c] + L.vel[c + s] * L.vel[c + s] + L.vel[c + 2 * s] * L.vel[c + 2 * s];
        rho_sum += r;
        nan_rho |= r != r; nan_v |= v2 != v2;
        rho_min = std::min(rho_min, r); rho_max = std::max(rho_max, r);
        v_max = std::max(v_max, std::sqrt(v2));
        ke += (double)(r * v2);
    }
    if (n_fluid > 0) {
        out[0] = n_fluid; out[1] = rho_sum / n_fluid; out[2] = nan_rho ? NAN : rho_min; out[3] = nan_rho ? NAN : rho_max;
        out[4] = nan_v ? NAN : v_max; out[5] = 0.5 * ke;
    } else {
        out[0] = 0; out[1] = 1; out[2] = 1; out[3] = 1; out[4] = 0; out[5] = 0;
    }
    return LUDWIG_OK;
}

int64_t ludwig_device_bytes(const ludwig_ctx*) { return 0; }
// bench / test initial condition (no reference counterpart): uniform flow (1, (ux, 0, 0)), see include/ludwig_b200.h
int ludwig_init_uniform_flow(ludwig_ctx* ctx, float ux) {
    if (!ctx) return LUDWIG_EINVAL;
    for (auto& Lp : ctx->levels) {
        Level& L = *Lp;
        const size_t s = 512 * (size_t)L.nb;
        for (size_t c = 0; c < s; ++c) {
            const float u = L.obstacle[c] ? 0.0f : ux;
            for (int k = 0; k < 27; ++k) {
                const float v = calculate_equilibrium(1.0f, u, 0.0f, 0.0f, LAT.w[k], (float)LAT.cx[k], (float)LAT.cy[k], (float)LAT.cz[k]);
                L.f[c + s * k] = v; L.f_temp[c + s * k] = v;
            }
            L.vel[c] = u; L.vel_temp[c] = u; L.vel[c + s] = 0; L.vel_temp[c + s] = 0; L.vel[c + 2 * s] = 0; L.vel_temp[c + 2 * s] = 0;
            L.rho[c] = 1.0f;
            if (!L.rho_old.empty()) L.rho_old[c] = 1.0f;
        }
    }
    return LUDWIG_OK;
}
// multi-GPU entry points: the oracle is single-process; it only shares the (host-side) partition rule
int ludwig_partition_starts(int32_t n_blocks, int32_t world, int32_t* starts) {
    if (n_blocks < 0 || world < 1 || world > 8) return LUDWIG_EINVAL;
    if (starts) for (int r = 0; r <= world; ++r) starts[r] = (int32_t)(((int64_t)n_blocks * r) / world);
    return LUDWIG_OK;
}
int ludwig_block_costs(const ludwig_level_desc* d, float* cost) {
    if (!d || !cost) return LUDWIG_EINVAL;
    for (int b = 0; b < d->n_blocks; ++b) cost[b] = 1.0f;     // the oracle never partitions
    return LUDWIG_OK;
}
int ludwig_partition_plan(const ludwig_level_desc* const*, int32_t, int32_t world, uint64_t* keys) {
    if (!keys || world != 1) return LUDWIG_EINVAL;
    keys[0] = 0; keys[1] = ~0ull;
    return LUDWIG_OK;
}
int ludwig_ctx_set_partition_keys(ludwig_ctx*, const uint64_t*, int32_t) { return LUDWIG_OK; }
int ludwig_ctx_set_partition(ludwig_ctx* ctx, int32_t rank, int32_t world) {
    return (rank == 0 && world == 1) ? LUDWIG_OK : fail(ctx, LUDWIG_ESTATE, "the CPU oracle is single-rank");
}
int ludwig_set_barrier_callback(ludwig_ctx*, int (*)(void*), void*) { return LUDWIG_OK; }
int ludwig_level_local_blocks(ludwig_ctx* ctx, int32_t level, int32_t* n_local, int32_t* ref_indices) {
    if (!ctx || level < 0 || level >= (int)ctx->levels.size()) return LUDWIG_EINVAL;
    int nb = ctx->levels[level]->nb;
    if (n_local) *n_local = nb;
    if (ref_indices) for (int i = 0; i < nb; ++i) ref_indices[i] = i + 1;
    return LUDWIG_OK;
}
int ludwig_level_upload_local(ludwig_ctx* ctx, int32_t level, int32_t which, const void* src) { return ludwig_level_upload(ctx, level, which, src); }
int ludwig_level_download_local(ludwig_ctx* ctx, int32_t level, int32_t which, void* dst) { return ludwig_level_download(ctx, level, which, dst); }
int ludwig_ipc_export(ludwig_ctx*, void*, int64_t, int64_t* needed) { if (needed) *needed = 0; return LUDWIG_OK; }
int ludwig_ipc_attach(ludwig_ctx* ctx, const void*, int64_t) { return fail(ctx, LUDWIG_ESTATE, "the CPU oracle is single-rank"); }
void* ludwig_ctx_stream(ludwig_ctx*) { return nullptr; }
int64_t ludwig_launch_count(const ludwig_ctx*) { return 0; }
int64_t ludwig_graph_replays(const ludwig_ctx*) { return 0; }
int64_t ludwig_ctx_self_check(ludwig_ctx*) { return 0; }
int64_t ludwig_multi_self_check(ludwig_multi*) { return LUDWIG_ESTATE; }
// N2 device entry points: the host restatement (open_ludwig_b200/host/domain_build.cpp) is their CPU counterpart, not this oracle
const char* ludwig_domain_last_error(void) { return "not part of the CPU oracle"; }
int ludwig_domain_voxelize(int, const double*, int64_t, const double*, double, const int32_t*, int32_t, const int32_t*, int32_t, int32_t, int32_t, uint8_t*) { return LUDWIG_ESTATE; }
int64_t ludwig_domain_flood_fill(int, const int32_t*, int32_t, const int32_t*, int32_t, int32_t, int32_t, uint8_t*) { return LUDWIG_ESTATE; }
int64_t ludwig_domain_wall_distance(int, const int32_t*, int32_t, const uint8_t*, double, float*) { return LUDWIG_ESTATE; }
int64_t ludwig_domain_qmap(int, const double*, int64_t, const double*, double, const int32_t*, int32_t, const int32_t*, int32_t, int32_t, int32_t, int64_t, int32_t*, double*, int32_t*) { return LUDWIG_ESTATE; }
int ludwig_multi_init_uniform_flow(ludwig_multi*, float) { return LUDWIG_ESTATE; }
int ludwig_profile_enable(ludwig_ctx*, int32_t) { return LUDWIG_OK; }
int ludwig_profile_classes(ludwig_ctx*, double out[8]) { for (int i = 0; i < 8; ++i) out[i] = 0; return LUDWIG_OK; }
int ludwig_profile_levels(ludwig_ctx*, double* out, int32_t capacity) { for (int i = 0; i < capacity; ++i) out[i] = 0; return LUDWIG_OK; }
int ludwig_partition_rcb(const ludwig_level_desc*, int32_t, int32_t*) { return LUDWIG_EINVAL; }   // multi-GPU only: not part of the oracle
int ludwig_partition_rcb_axes(const ludwig_level_desc*, int32_t, int32_t, int32_t*) { return LUDWIG_EINVAL; }
// The options select between implementations of the CUDA library; the oracle has one code path and accepts them all.
int ludwig_ctx_set_option(ludwig_ctx* ctx, const char* key, const char* value) { return (ctx && key && value) ? LUDWIG_OK : LUDWIG_EINVAL; }
// ludwig_multi (several ranks in one process) is multi-GPU plumbing: not part of the oracle.
struct ludwig_multi;
int ludwig_multi_create(ludwig_multi** out, int32_t, const int32_t*) { if (out) *out = nullptr; return LUDWIG_ESTATE; }
int ludwig_multi_destroy(ludwig_multi*) { return LUDWIG_OK; }
const char* ludwig_multi_last_error(const ludwig_multi*) { return "the CPU oracle is single-rank"; }
int32_t ludwig_multi_num_ranks(const ludwig_multi*) { return LUDWIG_ESTATE; }
ludwig_ctx* ludwig_multi_ctx(ludwig_multi*, int32_t) { return nullptr; }
int ludwig_multi_set_option(ludwig_multi*, const char*, const char*) { return LUDWIG_ESTATE; }
int ludwig_multi_set_partition_plan(ludwig_multi*, const ludwig_level_desc* const*, int32_t) { return LUDWIG_ESTATE; }
int ludwig_multi_level_create(ludwig_multi*, const ludwig_level_desc*, int32_t*) { return LUDWIG_ESTATE; }
int ludwig_multi_level_upload(ludwig_multi*, int32_t, int32_t, const void*) { return LUDWIG_ESTATE; }
int ludwig_multi_level_download(ludwig_multi*, int32_t, int32_t, void*) { return LUDWIG_ESTATE; }
int ludwig_multi_init_equilibrium(ludwig_multi*) { return LUDWIG_ESTATE; }
int ludwig_multi_step_batch(ludwig_multi*, int64_t, int32_t, float, const ludwig_params*) { return LUDWIG_ESTATE; }
int ludwig_multi_sync(ludwig_multi*) { return LUDWIG_ESTATE; }
int ludwig_multi_flow_stats(ludwig_multi*, int32_t, double*) { return LUDWIG_ESTATE; }
int ludwig_multi_forces_create(ludwig_multi*, int32_t, const float*, const float*, const float*, const float*, const float*, const float*, const float*,
                               double, double, double, double, const double*, int32_t, int32_t*) { return LUDWIG_ESTATE; }
int ludwig_multi_compute_aerodynamics(ludwig_multi*, int32_t, int32_t, const double*, double, double, int32_t, double*) { return LUDWIG_ESTATE; }
int ludwig_multi_forces_download_maps(ludwig_multi*, int32_t, float*, float*, float*, float*) { return LUDWIG_ESTATE; }
int64_t ludwig_multi_device_bytes(const ludwig_multi*) { return 0; }
int ludwig_multi_output_valid_blocks(ludwig_multi*, int32_t*, int32_t*) { return LUDWIG_ESTATE; }
int ludwig_multi_output_export(ludwig_multi*, int64_t, float*, float*, uint8_t*, int32_t*) { return LUDWIG_ESTATE; }
int ludwig_attach_inprocess(ludwig_ctx* ctx, ludwig_ctx* const*, int32_t) { return fail(ctx, LUDWIG_ESTATE, "the CPU oracle is single-rank"); }
int ludwig_profile_read(ludwig_ctx*, double* ms, int64_t* n, int64_t* c) { if (ms) *ms = 0; if (n) *n = 0; if (c) *c = 0; return LUDWIG_OK; }

}  // extern "C"
