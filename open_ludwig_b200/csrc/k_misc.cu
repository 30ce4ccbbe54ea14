// k_misc.cu — the small kernels around K1:
//   K0  init_eq!                          main.jl:109-124
//   K2  bouzidi_correction_kernel_fixed!  bouzidi_kernel.jl:13-92   (two-phase, no dense f_post_collision)
//   K3  map_stresses_kernel!              forces/surface.jl:138-266
//   K4  integrate_forces_kernel!          forces/surface.jl:282-366 (deterministic FP64 tree, no atomics)
//   R1  compute_flow_stats                diagnostics.jl:56-94      (one fused masked min/max/sum pass)
//   layout conversion reference <-> internal (the device side of adapt(), main.jl:98)
#include "ludwig_internal.h"

#include <cuda_fp16.h>

namespace ludwig {

// ---------------------------------------------------------------------------------------------
// K0
__global__ void init_eq_kernel(float* __restrict__ f0, float* __restrict__ f1, float* __restrict__ f_old, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per float4 of a (b,k) row
    if (i >= n) return;
    int k = (int)((i / (BS3 / 4)) % Q);
    int cx = k % 3 - 1, cy = (k / 3) % 3 - 1, cz = k / 9 - 1;
    int d2 = cx * cx + cy * cy + cz * cz;
    float w = d2 == 0 ? 8.0f / 27.0f : d2 == 1 ? 2.0f / 27.0f : d2 == 2 ? 1.0f / 54.0f : 1.0f / 216.0f;
    float4 v = make_float4(w, w, w, w);
    reinterpret_cast<float4*>(f0)[i] = v;
    reinterpret_cast<float4*>(f1)[i] = v;
    if (f_old) reinterpret_cast<float4*>(f_old)[i] = v;
}
void launch_init_eq(float* f0, float* f1, float* f_old, int nb, cudaStream_t s) {
    size_t n = (size_t)nb * Q * BS3 / 4;
    init_eq_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(f0, f1, f_old, n);
}

// uniform-flow variant of K0 (ludwig_init_uniform_flow): equilibrium of (1, (ux, 0, 0)) in the reference's expression order
__global__ void init_uniform_kernel(float* __restrict__ f0, float* __restrict__ f1, float* __restrict__ v0, float* __restrict__ v1,
                                    float* __restrict__ r0, float* __restrict__ r1, const uint8_t* __restrict__ obstacle, size_t ncell, float ux) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const size_t b = c >> 9, loc = c & 511;
    const float u = obstacle[c] ? 0.0f : ux;
    for (int k = 0; k < Q; ++k) {
        const int cx = k % 3 - 1, cy = (k / 3) % 3 - 1, cz = k / 9 - 1;
        const int d2 = cx * cx + cy * cy + cz * cz;
        const float w = d2 == 0 ? 8.0f / 27.0f : d2 == 1 ? 2.0f / 27.0f : d2 == 2 ? 1.0f / 54.0f : 1.0f / 216.0f;
        const float cu = __fmul_rn((float)cx, u);
        const float feq = __fmul_rn(w, __fsub_rn(__fadd_rn(__fadd_rn(1.0f, __fmul_rn(3.0f, cu)), __fmul_rn(__fmul_rn(4.5f, cu), cu)), __fmul_rn(1.5f, __fmul_rn(u, u))));
        f0[(b * Q + k) * BS3 + loc] = feq; f1[(b * Q + k) * BS3 + loc] = feq;
    }
    v0[b * 3 * BS3 + loc] = u; v1[b * 3 * BS3 + loc] = u;
    v0[(b * 3 + 1) * BS3 + loc] = 0.f; v1[(b * 3 + 1) * BS3 + loc] = 0.f; v0[(b * 3 + 2) * BS3 + loc] = 0.f; v1[(b * 3 + 2) * BS3 + loc] = 0.f;
    r0[c] = 1.0f;
    if (r1) r1[c] = 1.0f;
}
void launch_init_uniform(float* f0, float* f1, float* v0, float* v1, float* r0, float* r1, const uint8_t* obstacle, int nb, float ux, cudaStream_t s) {
    const size_t n = (size_t)nb * BS3;
    init_uniform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(f0, f1, v0, v1, r0, r1, obstacle, n, ux);
}

__global__ void fill_kernel(float* __restrict__ p, float v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}
void launch_fill(float* p, float v, size_t n, cudaStream_t s) {
    if (n == 0) return;
    unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
    fill_kernel<<<grid, 256, 0, s>>>(p, v, n);
}

// ---------------------------------------------------------------------------------------------
// layout conversion.  Reference: field[x,y,z,b_ref,k] (direction-major); internal: field[b_int][k][512].
// One (b,k) row = 512 floats = 128 float4; one thread per float4.
__global__ void ref_to_int_kernel(const float4* __restrict__ src_k, float4* __restrict__ dst, const int32_t* __restrict__ int2ref,
                                  int nb, int ncomp, int k) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nb * 128) return;
    int b = (int)(i >> 7), q = (int)(i & 127);
    dst[((size_t)b * ncomp + k) * 128 + q] = src_k[(size_t)(int2ref ? int2ref[b] : b) * 128 + q];
}
__global__ void int_to_ref_kernel(const float4* __restrict__ src, float4* __restrict__ dst_k, const int32_t* __restrict__ int2ref,
                                  int nb, int ncomp, int k) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nb * 128) return;
    int b = (int)(i >> 7), q = (int)(i & 127);
    dst_k[(size_t)(int2ref ? int2ref[b] : b) * 128 + q] = src[((size_t)b * ncomp + k) * 128 + q];
}
void launch_ref_to_int(const float* src_ref_k, float* dst, const int32_t* int2ref, int nb, int ncomp, int k, cudaStream_t s) {
    size_t n = (size_t)nb * 128;
    ref_to_int_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const float4*)src_ref_k, (float4*)dst, int2ref, nb, ncomp, k);
}
void launch_int_to_ref(const float* src, float* dst_ref_k, const int32_t* int2ref, int nb, int ncomp, int k, cudaStream_t s) {
    size_t n = (size_t)nb * 128;
    int_to_ref_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const float4*)src, (float4*)dst_ref_k, int2ref, nb, ncomp, k);
}
__global__ void ref_to_int_u8_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, const int32_t* __restrict__ int2ref, int nb, int fwd) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // 512 B per block = 32 uint4
    if (i >= (size_t)nb * 32) return;
    int b = (int)(i >> 5), q = (int)(i & 31);
    if (fwd) dst[(size_t)b * 32 + q] = src[(size_t)int2ref[b] * 32 + q];
    else dst[(size_t)int2ref[b] * 32 + q] = src[(size_t)b * 32 + q];
}
void launch_ref_to_int_u8(const uint8_t* src_ref, uint8_t* dst, const int32_t* int2ref, int nb, cudaStream_t s) {
    size_t n = (size_t)nb * 32;
    ref_to_int_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const uint4*)src_ref, (uint4*)dst, int2ref, nb, 1);
}
void launch_int_to_ref_u8(const uint8_t* src, uint8_t* dst_ref, const int32_t* int2ref, int nb, cudaStream_t s) {
    size_t n = (size_t)nb * 32;
    ref_to_int_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const uint4*)src, (uint4*)dst_ref, int2ref, nb, 0);
}

// per-block flag word: lets K1 skip the obstacle / sponge / wall_dist loads for blocks that have none
__global__ void block_flags_kernel(const uint8_t* __restrict__ obstacle, const float* __restrict__ sponge,
                                   const float* __restrict__ wall_dist, const int32_t* __restrict__ nbr, int32_t* __restrict__ bcoord, int nb) {
    int b = blockIdx.x;
    __shared__ unsigned s_flags;
    if (threadIdx.x == 0) s_flags = 0;
    __syncthreads();
    unsigned fl = 0;
    for (int c = threadIdx.x; c < BS3; c += blockDim.x) {
        size_t i = (size_t)b * BS3 + c;
        if (obstacle[i]) fl |= BF_OBSTACLE;
        if (sponge[i] > 0.0f) fl |= BF_SPONGE;
        float d = wall_dist[i];
        if (d > 0.0f && d < 10.0f) fl |= BF_WALLDIST;
    }
    if (threadIdx.x < 27 && nbr[(size_t)b * 27 + threadIdx.x] < 0) fl |= 0x80000000u;   // some neighbour missing
    if (fl) atomicOr(&s_flags, fl);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned f = s_flags;
        unsigned out = f & (BF_OBSTACLE | BF_SPONGE | BF_WALLDIST);
        if (!(f & 0x80000000u)) out |= BF_INTERIOR;
        bcoord[(size_t)b * 4 + 3] = (int32_t)out;
    }
}
void launch_block_flags(Level& L, cudaStream_t s) {
    block_flags_kernel<<<L.nb, 128, 0, s>>>(L.d_obstacle, L.d_sponge, L.d_wall_dist, L.d_nbr, L.d_bcoord, L.nb);
}

// ---------------------------------------------------------------------------------------------
// K2  Bouzidi.  The reference reads a dense post-collision copy f_post_collision (+108 B/cell written by
// K1) and overwrites f_out in place.  Before K2 runs, f_post_collision == f_out bit for bit
// (physics_kernels.jl:350-353), so here phase A gathers every correction from f_out into a compact
// [n_bc][27] buffer and phase B scatters them — same values, no dense array, no read/write hazard.
// One thread per ACTIVE link (q in (q_min, 1], compacted on the host when q_min is first seen): the reference's
// thread-per-cell loop over 27 mostly inactive directions (bouzidi_kernel.jl:35-38) becomes a dense list.
template <bool STRICT>
__global__ void bouzidi_gather_kernel(const float* __restrict__ f_out, const int32_t* __restrict__ link_cell,
                                      const uint8_t* __restrict__ link_k, const float* __restrict__ link_q,
                                      const int32_t* __restrict__ nbr, const long long* __restrict__ roff,
                                      float* __restrict__ tmp, int n_links) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_links) return;
    const float q = link_q[i];
    const int k = link_k[i];
    const int cell = link_cell[i];
    const int b = cell >> 9, c = cell & 511;
    const float* fb = f_out + (size_t)b * Q * BS3;
    const float f_k = fb[k * BS3 + c];
    float res;
    if (q < 0.5f) {
        const int x = c & 7, y = (c >> 3) & 7, z = c >> 6;
        const int cx = k % 3 - 1, cy = (k / 3) % 3 - 1, cz = k / 9 - 1;
        const int nx = x - cx, ny = y - cy, nz = z - cz;   // x_ff = x + c_opp(k)
        float f_ff = f_k;
        if (((nx | ny | nz) & ~7) == 0) {
            f_ff = fb[k * BS3 + nz * 64 + ny * 8 + nx];
        } else {
            int ox = nx < 0 ? -1 : (nx > 7 ? 1 : 0), oy = ny < 0 ? -1 : (ny > 7 ? 1 : 0), oz = nz < 0 ? -1 : (nz > 7 ? 1 : 0);
            int nbi = nbr[(size_t)b * 27 + (ox + 1) + (oy + 1) * 3 + (oz + 1) * 9];
            const int lc = (nz & 7) * 64 + (ny & 7) * 8 + (nx & 7);
            if (nbi >= REMOTE_BASE) f_ff = f_out[roff[nbi - REMOTE_BASE] + k * BS3 + lc];   // block owned by another GPU
            else if (nbi >= 0) f_ff = f_out[((size_t)nbi * Q + k) * BS3 + lc];
        }
        const float coeff1 = 2.0f * q;
        if (STRICT) res = __fadd_rn(__fmul_rn(coeff1, f_k), __fmul_rn(__fsub_rn(1.0f, coeff1), f_ff));
        else res = coeff1 * f_k + (1.0f - coeff1) * f_ff;
    } else {
        const float f_opp_post = fb[(26 - k) * BS3 + c];
        const float inv_2q = 1.0f / (2.0f * q);
        const float coeff2 = __fmul_rn(__fsub_rn(__fmul_rn(2.0f, q), 1.0f), inv_2q);
        if (STRICT) res = __fadd_rn(__fmul_rn(inv_2q, f_k), __fmul_rn(coeff2, f_opp_post));
        else res = inv_2q * f_k + coeff2 * f_opp_post;
    }
    tmp[i] = res;
}
__global__ void bouzidi_scatter_kernel(float* __restrict__ f_out, const int32_t* __restrict__ link_cell,
                                       const uint8_t* __restrict__ link_k, const float* __restrict__ tmp, int n_links) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_links) return;
    const int cell = link_cell[i];
    f_out[((size_t)(cell >> 9) * Q + (26 - link_k[i])) * BS3 + (cell & 511)] = tmp[i];
}
// phase 1 = gather, 2 = scatter, 0 = both (multi-GPU runs them separately with a cross-rank barrier in between)
void launch_bouzidi(const Level& L, float* f_out, const long long* roff, bool strict, int phase, cudaStream_t s) {
    if (L.n_links == 0) return;
    unsigned grid = (L.n_links + 255) / 256;
    if (phase != 2) {
        if (strict) bouzidi_gather_kernel<true><<<grid, 256, 0, s>>>(f_out, L.d_link_cell, L.d_link_k, L.d_link_q, L.d_nbr, roff, L.d_link_tmp, L.n_links);
        else bouzidi_gather_kernel<false><<<grid, 256, 0, s>>>(f_out, L.d_link_cell, L.d_link_k, L.d_link_q, L.d_nbr, roff, L.d_link_tmp, L.n_links);
    }
    if (phase != 1) bouzidi_scatter_kernel<<<grid, 256, 0, s>>>(f_out, L.d_link_cell, L.d_link_k, L.d_link_tmp, L.n_links);
}

// ---------------------------------------------------------------------------------------------
// N3 output gather (io_vtk.jl:52-58,100-111): the listed blocks only, written in the VTK writer's own array layout
// (rho[N], vel[3][N] component-fastest, obstacle[N]; non-finite values -> 0).  One CTA per listed block.
__global__ void __launch_bounds__(256) output_gather_kernel(const int32_t* __restrict__ sel, const float* __restrict__ rho,
                                                            const float* __restrict__ vel, const uint8_t* __restrict__ obs,
                                                            float* __restrict__ o_rho, float* __restrict__ o_vel, uint8_t* __restrict__ o_obs) {
    const int i = blockIdx.x, b = sel[i];
    for (int c = threadIdx.x; c < BS3; c += blockDim.x) {
        const size_t s = (size_t)b * BS3 + c, o = (size_t)i * BS3 + c;
        auto clean = [](float v) { return isfinite(v) ? v : 0.0f; };
        o_rho[o] = clean(rho[s]);
        const float* vb = vel + (size_t)b * 3 * BS3 + c;
        o_vel[o * 3 + 0] = clean(vb[0]); o_vel[o * 3 + 1] = clean(vb[BS3]); o_vel[o * 3 + 2] = clean(vb[2 * BS3]);
        o_obs[o] = obs[s] ? 1 : 0;
    }
}
void launch_output_gather(const int32_t* sel, int n, const float* rho, const float* vel, const uint8_t* obs, float* o_rho, float* o_vel,
                          uint8_t* o_obs, cudaStream_t s) {
    if (n > 0) output_gather_kernel<<<n, 256, 0, s>>>(sel, rho, vel, obs, o_rho, o_vel, o_obs);
}

// ---------------------------------------------------------------------------------------------
// Packed halo exchange (multi-GPU).  What K1 pulls from a neighbour block owned by another GPU is one LAYER of it: for
// the neighbour at offset d = (dx,dy,dz) the cells on its side facing the local block (x = 0 if dx = +1, x = 7 if dx = -1,
// all 8 if dx = 0; same for y, z) and the populations that can cross that side (c_a = -d_a on every axis with d_a != 0:
// 9 for a face, 3 for an edge, 1 for a corner), plus the three velocity components of a face layer (WALE reads axis
// neighbours only): 768 / 24 / 1 floats.  Read in place over NVLink, an x-face costs a 32-byte sector per float.  So:
//   * halo_pack_kernel (exporter, end of its level step): gathers every layer some peer needs from f_out / vel_out into a
//     contiguous export buffer (local traffic), double-buffered by step parity;
//   * halo_unpack_kernel (importer, start of the next level step, after the cross-rank barrier): pulls its segments of the
//     peers' export buffers with fully coalesced loads over NVLink and scatters them into a LOCAL mirror of each remote
//     block at the same in-block positions.  It runs on a high-priority side stream concurrently with the K1 launch over
//     the blocks that have no remote neighbour; K1 itself only ever reads local memory.
// Both sides derive the same entry order from the global tables (abi.cu, ludwig_level_create).  One warp per entry.
__device__ __forceinline__ int halo_entry_size(int dir) {
    const int dx = dir % 3 - 1, dy = (dir / 3) % 3 - 1, dz = dir / 9 - 1;
    const int nz = (dx != 0) + (dy != 0) + (dz != 0);
    return nz == 1 ? 768 : nz == 2 ? 24 : 1;
}
// element i of the entry for direction dir -> offset in the block's f array (i < npop * ncell) or, for the rest, in its
// velocity array (returned negative minus one)
__device__ __forceinline__ int halo_elem_offset(int dir, int i) {
    const int dx = dir % 3 - 1, dy = (dir / 3) % 3 - 1, dz = dir / 9 - 1;
    const int nx = dx ? 1 : 8, ny = dy ? 1 : 8, nz = dz ? 1 : 8;
    const int x0 = dx < 0 ? 7 : 0, y0 = dy < 0 ? 7 : 0, z0 = dz < 0 ? 7 : 0;
    const int ncell = nx * ny * nz;
    const int px = dx ? 1 : 3, py = dy ? 1 : 3, pz = dz ? 1 : 3;   // free lattice components per axis
    const int nf = px * py * pz * ncell;
    const int ii = i < nf ? i : i - nf;
    const int j = ii / ncell, c = ii - j * ncell;
    const int cell = (z0 + c / (nx * ny)) * 64 + (y0 + (c / nx) % ny) * 8 + x0 + c % nx;
    if (i < nf) {
        const int cx = dx ? -dx : j % px - 1, cy = dy ? -dy : (j / px) % py - 1, cz = dz ? -dz : j / (px * py) - 1;
        return ((cx + 1) + 3 * (cy + 1) + 9 * (cz + 1)) * BS3 + cell;
    }
    return -(j * BS3 + cell) - 1;
}
__global__ void __launch_bounds__(256) halo_pack_kernel(const int4* __restrict__ ent, int n, const float* __restrict__ f,
                                                        const float* __restrict__ vel, float* __restrict__ exp_buf) {
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= n) return;
    const int4 q = ent[e];   // x: local block, y: direction, z: offset in the export buffer
    const float* __restrict__ fb = f + (size_t)q.x * (Q * BS3);
    const float* __restrict__ vb = vel + (size_t)q.x * (3 * BS3);
    float* __restrict__ dst = exp_buf + q.z;
    const int n_el = halo_entry_size(q.y);
#pragma unroll 4
    for (int i = threadIdx.x & 31; i < n_el; i += 32) {
        const int off = halo_elem_offset(q.y, i);
        dst[i] = off >= 0 ? fb[off] : vb[-off - 1];
    }
}
struct HaloSrc { const float* p[MAX_RANKS]; };
__global__ void __launch_bounds__(256) halo_unpack_kernel(const int4* __restrict__ ent, int n, HaloSrc src, float* __restrict__ fmirror,
                                                          float* __restrict__ vmirror) {
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= n) return;
    const int4 q = ent[e];   // x: mirror slot (remote id), y: direction, z: offset in the exporter's buffer, w: exporter rank
    const float* __restrict__ sp = src.p[q.w] + q.z;
    float* __restrict__ fb = fmirror + (size_t)q.x * (Q * BS3);
    float* __restrict__ vb = vmirror + (size_t)q.x * (3 * BS3);
    const int n_el = halo_entry_size(q.y);
#pragma unroll 4
    for (int i = threadIdx.x & 31; i < n_el; i += 32) {
        const float v = sp[i];
        const int off = halo_elem_offset(q.y, i);
        if (off >= 0) fb[off] = v; else vb[-off - 1] = v;
    }
}
void launch_halo_pack(const Level& L, int buf, cudaStream_t s) {
    if (L.n_pack <= 0) return;
    halo_pack_kernel<<<(L.n_pack + 7) / 8, 256, 0, s>>>(L.d_pack, L.n_pack, L.d_f[buf], L.d_vel[buf], L.d_export + (size_t)buf * L.export_floats);
}
void launch_halo_unpack(const Level& L, int buf, cudaStream_t s) {
    if (L.n_unpack <= 0) return;
    HaloSrc src;
    for (int r = 0; r < MAX_RANKS; ++r) src.p[r] = L.peer_export[r] ? L.peer_export[r] + (size_t)buf * L.peer_export_floats[r] : nullptr;
    halo_unpack_kernel<<<(L.n_unpack + 7) / 8, 256, 0, s>>>(L.d_unpack, L.n_unpack, src, L.d_fmirror, L.d_vmirror);
}

// ---------------------------------------------------------------------------------------------
// Cross-GPU barrier (multi-GPU, one process per GPU): every rank owns MAX_RANKS epoch slots in device memory that its
// peers have mapped through CUDA IPC.  Lane r writes this rank's epoch into peer r's slot [my rank] (a store over
// NVLink), then spins on the local slot [r] until peer r has done the same.  Stream-ordered, no host round trip, ~5 us.
// Each rank runs on its OWN GPU, so the spinning kernels never wait for a kernel queued behind them; a time-out (option
// barrier_timeout_s, default 20 s of %globaltimer: a peer died or the ranks issued different numbers of barriers) sets a STICKY
// error flag in mapped host memory instead of hanging the GPU; the host turns it into LUDWIG_ESTATE from every later call.
struct BarrierPeers { unsigned int* slot[MAX_RANKS]; };
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void peer_barrier_kernel(BarrierPeers peers, unsigned int* own, int rank, int world, unsigned int epoch, volatile int* err, volatile int* err_host, long long timeout_ns) {
    const int r = threadIdx.x;
    if (r < world && r != rank) {
        __threadfence_system();
        *((volatile unsigned int*)(peers.slot[r] + rank)) = epoch;   // always publish: the peers must not wait for a failed rank's epoch
        __threadfence_system();
        if (*err != 0) return;                                       // sticky: a context whose barrier timed out never spins again
        const unsigned long long t0 = global_ns();
        while ((int)(*((volatile unsigned int*)(own + r)) - epoch) < 0) {
            if ((long long)(global_ns() - t0) > timeout_ns) { *err = 1; *err_host = 1; __threadfence_system(); break; }
            __nanosleep(100);
        }
        __threadfence_system();
    }
}
void launch_peer_barrier(unsigned int* const* peer_slots, unsigned int* own, int rank, int world, unsigned int epoch, int* err, int* err_host, long long timeout_ns, cudaStream_t s) {
    BarrierPeers p;
    for (int r = 0; r < MAX_RANKS; ++r) p.slot[r] = peer_slots[r];
    peer_barrier_kernel<<<1, 32, 0, s>>>(p, own, rank, world, epoch, err, err_host, timeout_ns);
}

// ---------------------------------------------------------------------------------------------
// K3  nearest-fluid-cell search per triangle + pressure / shear (forces/surface.jl:32-124,138-266).
// The arithmetic is written with explicit _rn intrinsics so that it matches the oracle without FMA.
__global__ void map_stresses_kernel(const PeerPtrs rho, const PeerPtrs vel, const PeerBytes obstacle,
                                    const int32_t* __restrict__ ptr, int dimx, int dimy, int dimz,
                                    const float* __restrict__ tcx, const float* __restrict__ tcy, const float* __restrict__ tcz,
                                    const float* __restrict__ tnx, const float* __restrict__ tny, const float* __restrict__ tnz,
                                    float* __restrict__ p_map, float* __restrict__ sx_map, float* __restrict__ sy_map, float* __restrict__ sz_map,
                                    int n_tri, float dx, float offx, float offy, float offz, float pscale, float sscale,
                                    float tau_molecular, int search_radius, int tri_first, int tri_stride) {
    int i = tri_first + (blockIdx.x * blockDim.x + threadIdx.x) * tri_stride;   // multi-GPU: triangles are dealt round-robin
    if (i >= n_tri) return;
    float tx = __fadd_rn(tcx[i], offx), ty = __fadd_rn(tcy[i], offy), tz = __fadd_rn(tcz[i], offz);
    float n_x = tnx[i], n_y = tny[i], n_z = tnz[i];
    int g_x = (int)floorf(__fdiv_rn(tx, dx)) + 1, g_y = (int)floorf(__fdiv_rn(ty, dx)) + 1, g_z = (int)floorf(__fdiv_rn(tz, dx)) + 1;
    float best_dist_sq = 1e10f, best_rho = 1.0f, best_ux = 0.f, best_uy = 0.f, best_uz = 0.f, best_wall_dist = 0.5f;
    bool found = false;
    for (int radius = 0; radius <= search_radius; ++radius) {
        if (found && radius > 1) break;
        for (int dz = -radius; dz <= radius; ++dz)
            for (int dy = -radius; dy <= radius; ++dy)
                for (int ddx = -radius; ddx <= radius; ++ddx) {
                    if (radius > 0 && !(abs(ddx) == radius || abs(dy) == radius || abs(dz) == radius)) continue;
                    int cgx = g_x + ddx, cgy = g_y + dy, cgz = g_z + dz;
                    if (cgx < 1 || cgy < 1 || cgz < 1) continue;
                    int bx = (cgx - 1) >> 3, by = (cgy - 1) >> 3, bz = (cgz - 1) >> 3;
                    if (bx >= dimx || by >= dimy || bz >= dimz) continue;
                    int enc = ptr[bx + dimx * (by + dimy * bz)];
                    if (enc < 0) continue;
                    const int pr = enc >> PTR_RANK_SHIFT, bi = enc & PTR_LOCAL_MASK;   // owning rank, its local block
                    int loc = ((cgx - 1) & 7) + 8 * ((cgy - 1) & 7) + 64 * ((cgz - 1) & 7);
                    size_t c = (size_t)bi * BS3 + loc;
                    if (obstacle.p[pr][c]) continue;
                    float ccx = __fmul_rn(__fsub_rn((float)cgx, 0.5f), dx), ccy = __fmul_rn(__fsub_rn((float)cgy, 0.5f), dx),
                          ccz = __fmul_rn(__fsub_rn((float)cgz, 0.5f), dx);
                    float ex = __fsub_rn(tx, ccx), ey = __fsub_rn(ty, ccy), ez = __fsub_rn(tz, ccz);
                    float dist_sq = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
                    if (dist_sq < best_dist_sq) {
                        best_dist_sq = dist_sq;
                        best_rho = rho.p[pr][c];
                        size_t vi = (size_t)bi * 3 * BS3 + loc;
                        const float* __restrict__ vp = vel.p[pr];
                        best_ux = vp[vi]; best_uy = vp[vi + BS3]; best_uz = vp[vi + 2 * BS3];
                        best_wall_dist = __fdiv_rn(__fsqrt_rn(dist_sq), dx);
                        found = true;
                    }
                }
    }
    float p_val = 0.f, tau_x = 0.f, tau_y = 0.f, tau_z = 0.f;
    if (found) {
        float wall_dist = fmaxf(best_wall_dist, 0.5f);
        float p_gauge_lat = __fdiv_rn(__fsub_rn(best_rho, 1.0f), 3.0f);
        p_val = __fmul_rn(p_gauge_lat, pscale);
        float u_dot_n = __fadd_rn(__fadd_rn(__fmul_rn(best_ux, n_x), __fmul_rn(best_uy, n_y)), __fmul_rn(best_uz, n_z));
        float ut_x = __fsub_rn(best_ux, __fmul_rn(u_dot_n, n_x)), ut_y = __fsub_rn(best_uy, __fmul_rn(u_dot_n, n_y)),
              ut_z = __fsub_rn(best_uz, __fmul_rn(u_dot_n, n_z));
        float u_tan_mag = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(ut_x, ut_x), __fmul_rn(ut_y, ut_y)), __fmul_rn(ut_z, ut_z)));
        float nu_lat = __fdiv_rn(__fsub_rn(tau_molecular, 0.5f), 3.0f);
        if (u_tan_mag > 1.0e-10f && wall_dist > 0.01f) {
            float tau_lat_mag = __fdiv_rn(__fmul_rn(__fmul_rn(best_rho, nu_lat), u_tan_mag), wall_dist);
            float tau_phys_mag = __fmul_rn(tau_lat_mag, sscale);
            tau_x = __fmul_rn(__fdiv_rn(ut_x, u_tan_mag), tau_phys_mag);
            tau_y = __fmul_rn(__fdiv_rn(ut_y, u_tan_mag), tau_phys_mag);
            tau_z = __fmul_rn(__fdiv_rn(ut_z, u_tan_mag), tau_phys_mag);
        }
    }
    p_map[i] = p_val; sx_map[i] = tau_x; sy_map[i] = tau_y; sz_map[i] = tau_z;
}
void launch_map_stresses(const Level& L, const PeerPtrs& rho, const PeerPtrs& vel, const PeerBytes& obstacle, const ludwig_mesh& M,
                         ludwig_forces& F, float dx, float offx, float offy, float offz, float pscale, float sscale, int radius,
                         int tri_first, int tri_stride, cudaStream_t s) {
    const int n_mine = (M.n - tri_first + tri_stride - 1) / tri_stride;
    if (n_mine <= 0) return;
    map_stresses_kernel<<<(n_mine + 127) / 128, 128, 0, s>>>(rho, vel, obstacle, L.d_ptr, L.dimx, L.dimy, L.dimz, M.cx, M.cy, M.cz,
                                                             M.nx, M.ny, M.nz, F.p, F.sx, F.sy, F.sz, M.n, dx, offx, offy, offz, pscale,
                                                             sscale, L.tau, radius, tri_first, tri_stride);
}

// K4.  Per-triangle contributions are formed in FP32 exactly as the reference does (forces/surface.jl:298-352)
// and then summed in FP64 by a fixed-shape tree (warp shuffles + one CTA): deterministic, unlike the
// reference's 9 same-address FP32 atomics per triangle.
__global__ void __launch_bounds__(1024) integrate_forces_kernel(const float* __restrict__ p_map, const float* __restrict__ sx_map,
                                                                const float* __restrict__ sy_map, const float* __restrict__ sz_map,
                                                                const float* __restrict__ tcx, const float* __restrict__ tcy,
                                                                const float* __restrict__ tcz, const float* __restrict__ tnx,
                                                                const float* __restrict__ tny, const float* __restrict__ tnz,
                                                                const float* __restrict__ areas, int n_tri, float offx, float offy,
                                                                float offz, float refx, float refy, float refz, int tri_first,
                                                                int tri_stride, double* __restrict__ out) {
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tri_first + threadIdx.x * tri_stride; i < n_tri; i += blockDim.x * tri_stride) {
        float p = p_map[i], tau_x = sx_map[i], tau_y = sy_map[i], tau_z = sz_map[i];
        float nx = tnx[i], ny = tny[i], nz = tnz[i], A = areas[i];
        float cx = __fadd_rn(tcx[i], offx), cy = __fadd_rn(tcy[i], offy), cz = __fadd_rn(tcz[i], offz);
        float dFp_x = __fmul_rn(__fmul_rn(-p, nx), A), dFp_y = __fmul_rn(__fmul_rn(-p, ny), A), dFp_z = __fmul_rn(__fmul_rn(-p, nz), A);
        float dFv_x = __fmul_rn(tau_x, A), dFv_y = __fmul_rn(tau_y, A), dFv_z = __fmul_rn(tau_z, A);
        float dFx = __fadd_rn(dFp_x, dFv_x), dFy = __fadd_rn(dFp_y, dFv_y), dFz = __fadd_rn(dFp_z, dFv_z);
        float rx = __fsub_rn(cx, refx), ry = __fsub_rn(cy, refy), rz = __fsub_rn(cz, refz);
        float dMx = __fsub_rn(__fmul_rn(ry, dFz), __fmul_rn(rz, dFy));
        float dMy = __fsub_rn(__fmul_rn(rz, dFx), __fmul_rn(rx, dFz));
        float dMz = __fsub_rn(__fmul_rn(rx, dFy), __fmul_rn(ry, dFx));
        acc[0] += dFp_x; acc[1] += dFp_y; acc[2] += dFp_z;
        acc[3] += dFv_x; acc[4] += dFv_y; acc[5] += dFv_z;
        acc[6] += dMx; acc[7] += dMy; acc[8] += dMz;
    }
    __shared__ double s_part[32][9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        double v = acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][j] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            double v = lane < (int)(blockDim.x >> 5) ? s_part[lane][j] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) out[j] = v;
        }
    }
}
void launch_integrate_forces(const ludwig_mesh& M, ludwig_forces& F, float offx, float offy, float offz, int tri_first, int tri_stride,
                             cudaStream_t s) {
    integrate_forces_kernel<<<1, 1024, 0, s>>>(F.p, F.sx, F.sy, F.sz, M.cx, M.cy, M.cz, M.nx, M.ny, M.nz, M.area, M.n, offx, offy, offz,
                                               (float)F.mc[0], (float)F.mc[1], (float)F.mc[2], tri_first, tri_stride, F.d_acc);
}

// ---------------------------------------------------------------------------------------------
// R1  flow statistics over the non-obstacle cells of one level: n, sum(rho), min, max, max|u|, sum(rho*u^2).
// partial layout per CTA: [n, rho_sum, rho_min, rho_max, v_max, ke].  Julia's minimum / maximum (diagnostics.jl:70-77) propagate
// NaN and fminf / fmaxf drop it, so NaN densities / velocities are tracked separately and turn the CTA's min / max / v_max
// partials into NaN: the rho_min column of the console table is the reference's only divergence indicator.
__global__ void __launch_bounds__(256) flow_stats_kernel(const float* __restrict__ rho, const float* __restrict__ vel,
                                                         const uint8_t* __restrict__ obstacle, size_t ncell, double* __restrict__ partials) {
    double n = 0, rs = 0, ke = 0;
    float rmin = INFINITY, rmax = -INFINITY, vmax = 0.f;
    int bad = 0;   // bit 0: NaN density, bit 1: NaN velocity
    // four x-adjacent cells per thread and iteration (one 16-byte load per array: the pass is a pure stream over 17 bytes per
    // cell, and the bytes in flight per thread are what bounds it); ncell is a multiple of 512
    for (size_t c = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; c < ncell; c += (size_t)gridDim.x * blockDim.x * 4) {
        const uchar4 ob = *reinterpret_cast<const uchar4*>(obstacle + c);
        const size_t b = c >> 9, loc = c & 511;
        const size_t vi = b * 3 * BS3 + loc;
        const float4 r4 = *reinterpret_cast<const float4*>(rho + c);
        const float4 x4 = *reinterpret_cast<const float4*>(vel + vi), y4 = *reinterpret_cast<const float4*>(vel + vi + BS3),
                     z4 = *reinterpret_cast<const float4*>(vel + vi + 2 * BS3);
        const float rr[4] = {r4.x, r4.y, r4.z, r4.w}, xx[4] = {x4.x, x4.y, x4.z, x4.w}, yy[4] = {y4.x, y4.y, y4.z, y4.w}, zz[4] = {z4.x, z4.y, z4.z, z4.w};
        const unsigned char oo[4] = {ob.x, ob.y, ob.z, ob.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (oo[i]) continue;
            const float r = rr[i], ux = xx[i], uy = yy[i], uz = zz[i];
            float v2 = __fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), __fmul_rn(uz, uz));
            n += 1; rs += r; ke += (double)__fmul_rn(r, v2);
            bad |= (r != r ? 1 : 0) | (v2 != v2 ? 2 : 0);
            rmin = fminf(rmin, r); rmax = fmaxf(rmax, r); vmax = fmaxf(vmax, __fsqrt_rn(v2));
        }
    }
    __shared__ double s_d[8][3];
    __shared__ float s_f[8][3];
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n += __shfl_down_sync(0xffffffffu, n, o); rs += __shfl_down_sync(0xffffffffu, rs, o); ke += __shfl_down_sync(0xffffffffu, ke, o);
        rmin = fminf(rmin, __shfl_down_sync(0xffffffffu, rmin, o)); rmax = fmaxf(rmax, __shfl_down_sync(0xffffffffu, rmax, o));
        vmax = fmaxf(vmax, __shfl_down_sync(0xffffffffu, vmax, o));
    }
    if (bad) atomicOr(&s_bad, bad);
    if (lane == 0) { s_d[warp][0] = n; s_d[warp][1] = rs; s_d[warp][2] = ke; s_f[warp][0] = rmin; s_f[warp][1] = rmax; s_f[warp][2] = vmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            n += s_d[w][0]; rs += s_d[w][1]; ke += s_d[w][2];
            rmin = fminf(rmin, s_f[w][0]); rmax = fmaxf(rmax, s_f[w][1]); vmax = fmaxf(vmax, s_f[w][2]);
        }
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        double* o = partials + (size_t)blockIdx.x * 6;
        o[0] = n; o[1] = rs; o[2] = (s_bad & 1) ? qnan : (double)rmin; o[3] = (s_bad & 1) ? qnan : (double)rmax;
        o[4] = (s_bad & 2) ? qnan : (double)vmax; o[5] = ke;
    }
}
void launch_flow_stats(const Level& L, const float* rho, const float* vel, double* d_partials, int nparts, cudaStream_t s) {
    flow_stats_kernel<<<nparts, 256, 0, s>>>(rho, vel, L.d_obstacle, (size_t)L.nb * BS3, d_partials);
}

}  // namespace ludwig
