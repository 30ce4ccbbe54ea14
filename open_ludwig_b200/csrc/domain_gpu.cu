// domain_gpu.cu — N2, the step BEFORE the hot path, on the device: the three brute-force phases of the reference's CPU domain
// build that dominate its run time ("hours for the fine bunny"; the restated host build spends 11.6 s in
// the q-map rays and 3.8 s in the wall distance of the 339 M-cell bunny on 8 cores):
//
//   voxelize_blocks!          domain_generation.jl:34-112   triangle / cell-box SAT test per cell (per-block triangle lists, margin 2 dx)
//   compute_wall_distances!   domain_generation.jl:371-431  nearest obstacle cell among the 26 neighbours
//   compute_q_map! + rays     bouzidi_setup.jl:12-54,64-166, bouzidi_math.jl:9-102   Moeller-Trumbore rays along the 26 lattice
//                                                            directions against the block's triangles (margin 2.5 dx), nearest hit
//
// The resulting tables are integers / Float16 and must be BIT-EXACT, so the geometry is evaluated in
// Float64 in the reference's operation order; this file is compiled with -fmad=false (no FMA contraction: IEEE add / mul / div /
// sqrt give the same bits as the host build, tests/test_domain_gpu.py compares the arrays byte for byte).
//
// Mapping: per-block triangle lists are built as a CSR on the device (count with atomics, exclusive scan, fill, per-block sort so
// that the lists are in ascending triangle order like the reference's enumerate loop); then one CTA per block that has triangles,
// one thread per cell.  The q-map runs twice over those blocks: pass A finds the boundary cells (a cell with at least one
// direction hit within one link), a scan over the per-block counts gives every block its first output row, pass B writes the rows
// in the reference's order (block, z, y, x).
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/ludwig_b200.h"

namespace {

constexpr int BS = 8;
thread_local std::string g_err;

struct V3 { double x, y, z; };
__host__ __device__ inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__host__ __device__ inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   // left-to-right, as StaticArrays
__device__ inline double min3(double a, double b, double c) { return fmin(fmin(a, b), c); }
__device__ inline double max3(double a, double b, double c) { return fmax(fmax(a, b), c); }
__device__ inline V3 vtx(const double* __restrict__ tris, int64_t t, int v, const double* off) {
    const double* p = tris + (t * 3 + v) * 3;
    return {p[0] + off[0], p[1] + off[1], p[2] + off[2]};
}

struct Geo {
    const double* tris; int64_t n_tri; double off[3]; double dx;
    const int32_t* coords; int nb;
    const int32_t* grid; int dimx, dimy, dimz;   // [dimx][dimy][dimz] (C order), 1-based block index, 0 = none
};

// block range a triangle's (offset, margin-expanded) bounding box covers (domain_generation.jl:49-60 / bouzidi_setup.jl:31-42)
__device__ inline void tri_block_range(const Geo& g, int64_t t, double margin, bool bouzidi_variant, int lo[3], int hi[3]) {
    const double* p = g.tris + t * 9;
    const double bsdx = BS * g.dx;
    for (int a = 0; a < 3; ++a) {
        double mn, mx;
        if (bouzidi_variant) {   // min / max of the raw vertices, then + offset
            mn = fmin(p[a], fmin(p[3 + a], p[6 + a])) + g.off[a];
            mx = fmax(p[a], fmax(p[3 + a], p[6 + a])) + g.off[a];
        } else {                 // min / max of the offset vertices
            mn = fmin(p[a] + g.off[a], fmin(p[3 + a] + g.off[a], p[6 + a] + g.off[a]));
            mx = fmax(p[a] + g.off[a], fmax(p[3 + a] + g.off[a], p[6 + a] + g.off[a]));
        }
        lo[a] = max(1, (int)floor((mn - margin) / bsdx) + 1);
        hi[a] = (int)floor((mx + margin) / bsdx) + 1;
    }
    hi[0] = min(hi[0], g.dimx); hi[1] = min(hi[1], g.dimy); hi[2] = min(hi[2], g.dimz);
}

// pass 0 / 1 of the CSR build: count (fill == nullptr) or fill
__global__ void tri_map_kernel(Geo g, double margin, int bouzidi_variant, int32_t* __restrict__ count, const int32_t* __restrict__ start,
                               int32_t* __restrict__ fill) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.n_tri) return;
    int lo[3], hi[3];
    tri_block_range(g, t, margin, bouzidi_variant != 0, lo, hi);
    for (int bz = lo[2]; bz <= hi[2]; ++bz)
        for (int by = lo[1]; by <= hi[1]; ++by)
            for (int bx = lo[0]; bx <= hi[0]; ++bx) {
                const int b = g.grid[((size_t)(bx - 1) * g.dimy + (by - 1)) * g.dimz + (bz - 1)];
                if (b <= 0) continue;
                const int pos = atomicAdd(&count[b - 1], 1);
                if (fill) fill[start[b - 1] + pos] = (int32_t)t;
            }
}
// exclusive scan of n ints by ONE CTA (n <= a few 10^5: tens of microseconds; this is set-up code)
__global__ void scan_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int n) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < n ? in[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = threadIdx.x < (blockDim.x >> 5) ? s_warp[threadIdx.x] : 0;
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const int before = s_carry + ((threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0);
        if (i < n) out[i] = before + x - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = s_carry;
}
// ascending triangle order inside every block's list (the atomics filled them in arbitrary order)
__global__ void sort_lists_kernel(const int32_t* __restrict__ start, int32_t* __restrict__ list, int nb) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    int32_t* l = list + start[b];
    const int n = start[b + 1] - start[b];
    for (int i = 1; i < n; ++i) {
        const int32_t v = l[i];
        int j = i - 1;
        while (j >= 0 && l[j] > v) { l[j + 1] = l[j]; --j; }
        l[j + 1] = v;
    }
}
// blocks that have at least one triangle, ascending
__global__ void flag_nonempty_kernel(const int32_t* __restrict__ start, int32_t* __restrict__ flag, int nb) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) flag[b] = start[b + 1] > start[b] ? 1 : 0;
}
__global__ void compact_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, int32_t* __restrict__ out, int nb) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb && flag[b]) out[pos[b]] = b;
}

// domain_generation.jl:10-32 (9-axis SAT: 3 box axes + 9 edge cross products, no triangle-normal axis)
__device__ bool triangle_intersects_aabb(V3 center, V3 box_half, V3 v1, V3 v2, V3 v3) {
    const double tol = 1.001;
    const V3 h{box_half.x * tol, box_half.y * tol, box_half.z * tol};
    const V3 t1 = sub(v1, center), t2 = sub(v2, center), t3 = sub(v3, center);
    if (min3(t1.x, t2.x, t3.x) > h.x || max3(t1.x, t2.x, t3.x) < -h.x) return false;
    if (min3(t1.y, t2.y, t3.y) > h.y || max3(t1.y, t2.y, t3.y) < -h.y) return false;
    if (min3(t1.z, t2.z, t3.z) > h.z || max3(t1.z, t2.z, t3.z) < -h.z) return false;
    const V3 f[3] = {sub(t2, t1), sub(t3, t2), sub(t1, t3)};
    const V3 u[3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const V3 axis = cross(u[i], f[j]);
            if (dot(axis, axis) < 1e-10) continue;
            const double p1 = dot(t1, axis), p2 = dot(t2, axis), p3 = dot(t3, axis);
            const double r = h.x * fabs(axis.x) + h.y * fabs(axis.y) + h.z * fabs(axis.z);
            if (fmin(p1, fmin(p2, p3)) > r || fmax(p1, fmax(p2, p3)) < -r) return false;
        }
    return true;
}

// domain_generation.jl:74-112: one CTA per block with triangles, one thread per cell
__global__ void __launch_bounds__(512) voxelize_kernel(Geo g, const int32_t* __restrict__ blocks, const int32_t* __restrict__ start,
                                                       const int32_t* __restrict__ list, uint8_t* __restrict__ obstacle) {
    const int b = blocks[blockIdx.x];
    const int c = threadIdx.x, lx = (c & 7) + 1, ly = ((c >> 3) & 7) + 1, lz = (c >> 6) + 1;
    const int bx = g.coords[3 * b], by = g.coords[3 * b + 1], bz = g.coords[3 * b + 2];
    const V3 center{((bx - 1) * BS + lx - 0.5) * g.dx, ((by - 1) * BS + ly - 0.5) * g.dx, ((bz - 1) * BS + lz - 0.5) * g.dx};
    const V3 box_half{0.75 * g.dx, 0.75 * g.dx, 0.75 * g.dx};
    bool is_shell = false;
    for (int i = start[b]; i < start[b + 1] && !is_shell; ++i) {
        const int64_t tid = list[i];
        is_shell = triangle_intersects_aabb(center, box_half, vtx(g.tris, tid, 0, g.off), vtx(g.tris, tid, 1, g.off), vtx(g.tris, tid, 2, g.off));
    }
    if (is_shell) obstacle[(size_t)b * 512 + c] = 1;
}

// domain_generation.jl:371-431: one thread per cell; neighbor_table is the reference's [27][nb] (1-based, 0 = none)
__global__ void wall_distance_kernel(const int32_t* __restrict__ nbr, int nb, const uint8_t* __restrict__ obstacle, float dxf,
                                     float* __restrict__ wall_dist, unsigned long long* __restrict__ n_near) {
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (size_t)nb * 512) return;
    if (obstacle[cell]) return;
    const int b = (int)(cell >> 9), c = (int)(cell & 511);
    const int lx = c & 7, ly = (c >> 3) & 7, lz = c >> 6;
    bool near = false;
    float min_dist = 100.0f;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dxo = -1; dxo <= 1; ++dxo) {
                if (dxo == 0 && dy == 0 && dz == 0) continue;
                const int nx = lx + dxo, ny = ly + dy, nz = lz + dz;
                const int ox = nx < 0 ? -1 : (nx >= BS ? 1 : 0), oy = ny < 0 ? -1 : (ny >= BS ? 1 : 0), oz = nz < 0 ? -1 : (nz >= BS ? 1 : 0);
                const int dir = (ox + 1) + (oy + 1) * 3 + (oz + 1) * 9;
                const int nb_i = dir == 13 ? b : nbr[(size_t)dir * nb + b] - 1;
                if (nb_i < 0) continue;
                if (obstacle[(size_t)nb_i * 512 + ((nz + BS) % BS) * 64 + ((ny + BS) % BS) * 8 + ((nx + BS) % BS)]) {
                    near = true;
                    const float dist = __fmul_rn(__fsqrt_rn((float)(dxo * dxo + dy * dy + dz * dz)), dxf);
                    min_dist = fminf(min_dist, dist);
                }
            }
    if (near) { wall_dist[cell] = min_dist; atomicAdd(n_near, 1ull); }
}

// bouzidi_math.jl:9-47 (Moeller-Trumbore, EPSILON 1e-9)
__device__ inline bool ray_triangle(V3 origin, V3 dir, V3 v1, V3 v2, V3 v3, double& t_out) {
    const double EPSILON = 1e-9;
    const V3 edge1 = sub(v2, v1), edge2 = sub(v3, v1);
    const V3 h = cross(dir, edge2);
    const double a = dot(edge1, h);
    if (fabs(a) < EPSILON) return false;
    const double f = 1.0 / a;
    const V3 s = sub(origin, v1);
    const double u = f * dot(s, h);
    if (u < 0.0 || u > 1.0) return false;
    const V3 q = cross(s, edge1);
    const double v = f * dot(dir, q);
    if (v < 0.0 || u + v > 1.0) return false;
    const double t = f * dot(edge2, q);
    if (t > EPSILON) { t_out = t; return true; }
    return false;
}

// bouzidi_setup.jl:64-166 + bouzidi_math.jl:53-102 for one cell: q[27] (0 = no link) and the 1-based triangle of each link.
// Pruning as in host/domain_build.cpp: a hit with q <= 1 lies within one dx of the cell centre on every axis, so a triangle whose
// offset bounding box misses [centre - 1.001 dx, centre + 1.001 dx] cannot supply the nearest hit of any link.
__device__ bool cell_links(const Geo& g, int b, int c, const int32_t* __restrict__ start, const int32_t* __restrict__ list, double* qv, int32_t* tv) {
    const int lx = (c & 7) + 1, ly = ((c >> 3) & 7) + 1, lz = (c >> 6) + 1;
    const int bx = g.coords[3 * b], by = g.coords[3 * b + 1], bz = g.coords[3 * b + 2];
    const V3 cc{((bx - 1) * BS + lx - 0.5) * g.dx, ((by - 1) * BS + ly - 0.5) * g.dx, ((bz - 1) * BS + lz - 0.5) * g.dx};
    const double r = g.dx * 1.001;
    double min_t[27];
    int32_t best[27];
    for (int k = 0; k < 27; ++k) { min_t[k] = INFINITY; best[k] = -1; }
    for (int i = start[b]; i < start[b + 1]; ++i) {
        const int64_t tid = list[i];
        const V3 v1 = vtx(g.tris, tid, 0, g.off), v2 = vtx(g.tris, tid, 1, g.off), v3 = vtx(g.tris, tid, 2, g.off);
        if (!(min3(v1.x, v2.x, v3.x) <= cc.x + r && max3(v1.x, v2.x, v3.x) >= cc.x - r && min3(v1.y, v2.y, v3.y) <= cc.y + r &&
              max3(v1.y, v2.y, v3.y) >= cc.y - r && min3(v1.z, v2.z, v3.z) <= cc.z + r && max3(v1.z, v2.z, v3.z) >= cc.z - r))
            continue;
        for (int k = 0; k < 27; ++k) {
            if (k == 13) continue;
            const double cx = (double)(k % 3 - 1), cy = (double)((k / 3) % 3 - 1), cz = (double)(k / 9 - 1);
            const double nrm = sqrt(cx * cx + cy * cy + cz * cz);
            const V3 dir{cx / nrm, cy / nrm, cz / nrm};
            double t;
            if (ray_triangle(cc, dir, v1, v2, v3, t) && t < min_t[k]) { min_t[k] = t; best[k] = (int32_t)tid; }   // first minimum in list order
        }
    }
    bool any = false;
    for (int k = 0; k < 27; ++k) {
        qv[k] = 0.0; tv[k] = 0;
        if (k == 13 || !(min_t[k] < INFINITY)) continue;
        const double cx = (double)(k % 3 - 1), cy = (double)((k / 3) % 3 - 1), cz = (double)(k / 9 - 1);
        const double q = min_t[k] / (g.dx * sqrt(cx * cx + cy * cy + cz * cz));
        if (q > 0.0 && q <= 1.0) { qv[k] = q; tv[k] = best[k] + 1; any = true; }
    }
    return any;
}
// pass A: boundary-cell flags + per-block counts.  pass B (rows != nullptr): the rows, in (block, z, y, x) order
__global__ void __launch_bounds__(512) qmap_kernel(Geo g, const int32_t* __restrict__ blocks, const int32_t* __restrict__ start,
                                                   const int32_t* __restrict__ list, uint8_t* __restrict__ flags, int32_t* __restrict__ blk_count,
                                                   const int32_t* __restrict__ blk_first, int32_t* __restrict__ out_cells, double* __restrict__ out_q,
                                                   int32_t* __restrict__ out_tri) {
    const int slot = blockIdx.x, b = blocks[slot], c = threadIdx.x;
    double qv[27];
    int32_t tv[27];
    if (!out_cells) {
        const bool any = cell_links(g, b, c, start, list, qv, tv);
        flags[(size_t)slot * 512 + c] = any ? 1 : 0;
        const int n = __syncthreads_count(any);
        if (c == 0) blk_count[slot] = n;
        return;
    }
    const bool any = flags[(size_t)slot * 512 + c] != 0;
    // rank of this cell among the block's boundary cells (cell index order = z, y, x)
    __shared__ int s_warp[16];
    const unsigned bal = __ballot_sync(0xffffffffu, any);
    const int lane = c & 31, warp = c >> 5;
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = __popc(bal & ((1u << lane) - 1));
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (!any) return;
    cell_links(g, b, c, start, list, qv, tv);
    const size_t row = (size_t)blk_first[slot] + before;
    out_cells[row * 4] = b + 1; out_cells[row * 4 + 1] = (c & 7) + 1; out_cells[row * 4 + 2] = ((c >> 3) & 7) + 1; out_cells[row * 4 + 3] = (c >> 6) + 1;
    for (int k = 0; k < 27; ++k) { out_q[row * 27 + k] = qv[k]; out_tri[row * 27 + k] = tv[k]; }
}

// ---- host side --------------------------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <typename T> T* as() { return (T*)p; }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
};
#define DCU(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) { g_err = std::string(#call) + ": " + cudaGetErrorString(e__); return LUDWIG_ECUDA; } \
    } while (0)

struct TriMap {   // device CSR of the per-block triangle lists + the ascending list of blocks that have any
    DevBuf tris, coords, grid, start, list, blocks;
    int n_blocks_with = 0;
    Geo g{};
};

int build_tri_map(TriMap& M, const double* tris, int64_t n_tri, const double off[3], double dx, const int32_t* coords, int nb, const int32_t* grid,
                  int dimx, int dimy, int dimz, double margin, bool bouzidi_variant) {
    const size_t ngrid = (size_t)dimx * dimy * dimz;
    DCU(M.tris.alloc((size_t)n_tri * 9 * sizeof(double))); DCU(M.coords.alloc((size_t)nb * 3 * sizeof(int32_t))); DCU(M.grid.alloc(ngrid * sizeof(int32_t)));
    DCU(cudaMemcpy(M.tris.p, tris, (size_t)n_tri * 9 * sizeof(double), cudaMemcpyHostToDevice));
    DCU(cudaMemcpy(M.coords.p, coords, (size_t)nb * 3 * sizeof(int32_t), cudaMemcpyHostToDevice));
    DCU(cudaMemcpy(M.grid.p, grid, ngrid * sizeof(int32_t), cudaMemcpyHostToDevice));
    Geo g{M.tris.as<double>(), n_tri, {off[0], off[1], off[2]}, dx, M.coords.as<int32_t>(), nb, M.grid.as<int32_t>(), dimx, dimy, dimz};
    M.g = g;
    DevBuf count, flag, pos;
    DCU(count.alloc((size_t)nb * sizeof(int32_t))); DCU(M.start.alloc(((size_t)nb + 1) * sizeof(int32_t)));
    DCU(cudaMemset(count.p, 0, (size_t)nb * sizeof(int32_t)));
    const unsigned tg = (unsigned)((n_tri + 127) / 128), bg = (unsigned)((nb + 255) / 256);
    tri_map_kernel<<<tg, 128>>>(g, margin, bouzidi_variant ? 1 : 0, count.as<int32_t>(), nullptr, nullptr);
    scan_kernel<<<1, 1024>>>(count.as<int32_t>(), M.start.as<int32_t>(), nb);
    int32_t total = 0;
    DCU(cudaMemcpy(&total, M.start.as<int32_t>() + nb, sizeof(int32_t), cudaMemcpyDeviceToHost));
    DCU(M.list.alloc((size_t)total * sizeof(int32_t)));
    DCU(cudaMemset(count.p, 0, (size_t)nb * sizeof(int32_t)));
    tri_map_kernel<<<tg, 128>>>(g, margin, bouzidi_variant ? 1 : 0, count.as<int32_t>(), M.start.as<int32_t>(), M.list.as<int32_t>());
    sort_lists_kernel<<<bg, 256>>>(M.start.as<int32_t>(), M.list.as<int32_t>(), nb);
    DCU(flag.alloc((size_t)nb * sizeof(int32_t))); DCU(pos.alloc(((size_t)nb + 1) * sizeof(int32_t)));
    flag_nonempty_kernel<<<bg, 256>>>(M.start.as<int32_t>(), flag.as<int32_t>(), nb);
    scan_kernel<<<1, 1024>>>(flag.as<int32_t>(), pos.as<int32_t>(), nb);
    DCU(cudaMemcpy(&M.n_blocks_with, pos.as<int32_t>() + nb, sizeof(int32_t), cudaMemcpyDeviceToHost));
    DCU(M.blocks.alloc((size_t)M.n_blocks_with * sizeof(int32_t)));
    compact_kernel<<<bg, 256>>>(flag.as<int32_t>(), pos.as<int32_t>(), M.blocks.as<int32_t>(), nb);
    DCU(cudaGetLastError());
    return LUDWIG_OK;
}

bool args_ok(const double* tris, int64_t n_tri, const double* off, double dx, const int32_t* coords, int nb, const int32_t* grid, int dx_, int dy_, int dz_) {
    if (!tris || n_tri <= 0 || !off || !(dx > 0) || !coords || nb <= 0 || !grid || dx_ <= 0 || dy_ <= 0 || dz_ <= 0) { g_err = "bad domain-build arguments"; return false; }
    return true;
}

// ---- flood fill (domain_generation.jl:114-203 perform_flood_fill!) ------------------------------------------------------------
// The reference walks a cell-by-cell FIFO from the non-obstacle cells of the min-bx blocks through 6-connected non-obstacle cells of
// existing blocks and then marks every cell it never reached as solid.  The reachable set does not depend on the visiting order, so
// the device computes the same set as a frontier of BLOCKS: one CTA per frontier block (a thread per cell) imports the visited
// cells of the six facing neighbour layers, propagates inside the block in shared memory until nothing changes, and queues the
// neighbours behind every face on which one of its cells became visited.  Visited flags only ever go 0 -> 1, a block is re-queued
// whenever a neighbour face changes after (or while) it was processed, so concurrent CTAs need no ordering; the host launches
// one kernel per frontier (its size is the only thing read back) until the frontier is empty.
__device__ inline int ff_neighbour(const int32_t* __restrict__ grid, int dimx, int dimy, int dimz, int bx, int by, int bz) {
    if (bx < 1 || bx > dimx || by < 1 || by > dimy || bz < 1 || bz > dimz) return -1;
    return grid[((size_t)(bx - 1) * dimy + (by - 1)) * dimz + (bz - 1)] - 1;   // [bx][by][bz] C order, 1-based, 0 = none
}
__global__ void __launch_bounds__(512) flood_fill_kernel(const int32_t* __restrict__ coords, const int32_t* __restrict__ grid, int dimx, int dimy, int dimz,
                                                         const uint8_t* __restrict__ obstacle, uint8_t* visited, const int32_t* __restrict__ frontier,
                                                         int32_t* next, int32_t* next_count, int32_t* queued, int seed_min_x) {
    __shared__ uint8_t s_v[512];
    __shared__ int s_face[6];
    const int b = frontier[blockIdx.x], c = threadIdx.x;
    const int lx = c & 7, ly = (c >> 3) & 7, lz = c >> 6;
    const int bx = coords[3 * b], by = coords[3 * b + 1], bz = coords[3 * b + 2];
    if (c < 6) s_face[c] = 0;
    // From now on a neighbour may queue this block again.  The reset comes BEFORE the reads of the neighbours' layers (fence +
    // barrier): a neighbour that finds the flag still set has published its cells before this CTA looks at them, one that
    // finds it cleared appends the block to the next frontier.  (A block can therefore enter a list twice per launch: the
    // lists hold 2 nb entries.)  Other CTAs' flags are read past the non-coherent L1 (__ldcg).
    if (c == 0) { atomicExch(&queued[b], 0); __threadfence(); }
    __syncthreads();
    const bool solid = obstacle[(size_t)b * 512 + c] != 0;
    const bool v0 = __ldcg(visited + (size_t)b * 512 + c) != 0;
    bool v = v0;
    if (seed_min_x >= 0 && bx == seed_min_x && !solid) v = true;            // the reference's seeds (first launch only)
    if (!solid && !v) {                                                    // import the facing layers of the six neighbours
        const int ddx[6] = {1, -1, 0, 0, 0, 0}, ddy[6] = {0, 0, 1, -1, 0, 0}, ddz[6] = {0, 0, 0, 0, 1, -1};
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int nx = lx + ddx[i], ny = ly + ddy[i], nz = lz + ddz[i];
            if (nx >= 0 && nx < BS && ny >= 0 && ny < BS && nz >= 0 && nz < BS) continue;
            const int t = ff_neighbour(grid, dimx, dimy, dimz, bx + ddx[i], by + ddy[i], bz + ddz[i]);
            if (t >= 0 && __ldcg(visited + (size_t)t * 512 + ((nz + BS) % BS) * 64 + ((ny + BS) % BS) * 8 + ((nx + BS) % BS))) v = true;
        }
    }
    // a block without obstacle cells is internally connected: reached anywhere, it is reached everywhere
    if (!__syncthreads_or(solid) && __syncthreads_or(v)) v = true;
    s_v[c] = v;
    __syncthreads();
    for (;;) {                                                              // in-block propagation to the fixed point
        bool grow = false;
        if (!solid && !v)
            grow = (lx > 0 && s_v[c - 1]) || (lx < 7 && s_v[c + 1]) || (ly > 0 && s_v[c - 8]) || (ly < 7 && s_v[c + 8]) || (lz > 0 && s_v[c - 64]) || (lz < 7 && s_v[c + 64]);
        __syncthreads();
        if (grow) { v = true; s_v[c] = 1; }
        if (!__syncthreads_or(grow)) break;
    }
    if (v && !v0) {
        visited[(size_t)b * 512 + c] = 1;
        if (lx == 7) s_face[0] = 1;
        if (lx == 0) s_face[1] = 1;
        if (ly == 7) s_face[2] = 1;
        if (ly == 0) s_face[3] = 1;
        if (lz == 7) s_face[4] = 1;
        if (lz == 0) s_face[5] = 1;
    }
    __threadfence();
    __syncthreads();
    if (c < 6 && s_face[c]) {
        const int ddx[6] = {1, -1, 0, 0, 0, 0}, ddy[6] = {0, 0, 1, -1, 0, 0}, ddz[6] = {0, 0, 0, 0, 1, -1};
        const int t = ff_neighbour(grid, dimx, dimy, dimz, bx + ddx[c], by + ddy[c], bz + ddz[c]);
        if (t >= 0 && atomicExch(&queued[t], 1) == 0) next[atomicAdd(next_count, 1)] = t;
    }
}
__global__ void ff_seed_kernel(const int32_t* __restrict__ coords, int nb, int min_x, int32_t* frontier, int32_t* count, int32_t* queued) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb && coords[3 * b] == min_x) { queued[b] = 1; frontier[atomicAdd(count, 1)] = b; }
}
__global__ void __launch_bounds__(512) ff_finish_kernel(uint8_t* obstacle, const uint8_t* __restrict__ visited, unsigned long long* filled) {
    const size_t i = (size_t)blockIdx.x * 512 + threadIdx.x;
    const bool fill = !obstacle[i] && !visited[i];
    if (fill) obstacle[i] = 1;
    const int n = __syncthreads_count(fill);
    if (threadIdx.x == 0 && n) atomicAdd(filled, (unsigned long long)n);
}

// the q-map is asked for twice (count, then fill): the device state of the counting call is kept for the fill call
struct QmapState {
    TriMap M;
    DevBuf flags, blk_count, blk_first;
    int64_t n_rows = 0;
    const void *k_tris = nullptr, *k_coords = nullptr; int64_t k_ntri = 0; int k_nb = 0; double k_dx = 0; int k_dev = -1;
};
QmapState* g_qmap = nullptr;

}  // namespace

extern "C" {

const char* ludwig_domain_last_error(void) { return g_err.c_str(); }

int ludwig_domain_voxelize(int device, const double* tris, int64_t n_tri, const double offset[3], double dx, const int32_t* coords, int32_t nb,
                           const int32_t* grid_ptr, int32_t dimx, int32_t dimy, int32_t dimz, uint8_t* obstacle) {
    if (!args_ok(tris, n_tri, offset, dx, coords, nb, grid_ptr, dimx, dimy, dimz) || !obstacle) return LUDWIG_EINVAL;
    DCU(cudaSetDevice(device));
    TriMap M;
    int rc = build_tri_map(M, tris, n_tri, offset, dx, coords, nb, grid_ptr, dimx, dimy, dimz, dx * 2, false);
    if (rc) return rc;
    DevBuf obs;
    DCU(obs.alloc((size_t)nb * 512));
    DCU(cudaMemcpy(obs.p, obstacle, (size_t)nb * 512, cudaMemcpyHostToDevice));
    if (M.n_blocks_with > 0)
        voxelize_kernel<<<M.n_blocks_with, 512>>>(M.g, M.blocks.as<int32_t>(), M.start.as<int32_t>(), M.list.as<int32_t>(), obs.as<uint8_t>());
    DCU(cudaGetLastError());
    DCU(cudaMemcpy(obstacle, obs.p, (size_t)nb * 512, cudaMemcpyDeviceToHost));
    return LUDWIG_OK;
}

int64_t ludwig_domain_flood_fill(int device, const int32_t* coords, int32_t nb, const int32_t* grid_ptr, int32_t dimx, int32_t dimy, int32_t dimz,
                                 uint8_t* obstacle) {
    if (!coords || nb <= 0 || !grid_ptr || dimx <= 0 || dimy <= 0 || dimz <= 0 || !obstacle) { g_err = "bad flood-fill arguments"; return LUDWIG_EINVAL; }
    DCU(cudaSetDevice(device));
    const size_t nc = (size_t)nb * 512, ngrid = (size_t)dimx * dimy * dimz;
    DevBuf dco, dgrid, obs, vis, lists, counts, queued, filled;
    DCU(dco.alloc((size_t)nb * 3 * sizeof(int32_t))); DCU(dgrid.alloc(ngrid * sizeof(int32_t))); DCU(obs.alloc(nc)); DCU(vis.alloc(nc));
    DCU(lists.alloc((size_t)nb * 4 * sizeof(int32_t))); DCU(counts.alloc(2 * sizeof(int32_t))); DCU(queued.alloc((size_t)nb * sizeof(int32_t)));
    DCU(filled.alloc(sizeof(unsigned long long)));
    DCU(cudaMemcpy(dco.p, coords, (size_t)nb * 3 * sizeof(int32_t), cudaMemcpyHostToDevice));
    DCU(cudaMemcpy(dgrid.p, grid_ptr, ngrid * sizeof(int32_t), cudaMemcpyHostToDevice));
    DCU(cudaMemcpy(obs.p, obstacle, nc, cudaMemcpyHostToDevice));
    DCU(cudaMemset(vis.p, 0, nc)); DCU(cudaMemset(counts.p, 0, 2 * sizeof(int32_t))); DCU(cudaMemset(queued.p, 0, (size_t)nb * sizeof(int32_t)));
    DCU(cudaMemset(filled.p, 0, sizeof(unsigned long long)));
    int min_x = coords[0];
    for (int i = 1; i < nb; ++i) min_x = std::min(min_x, coords[3 * i]);
    int32_t* list[2] = {lists.as<int32_t>(), lists.as<int32_t>() + 2 * (size_t)nb};
    int32_t* cnt = counts.as<int32_t>();
    ff_seed_kernel<<<(nb + 255) / 256, 256>>>(dco.as<int32_t>(), nb, min_x, list[0], cnt, queued.as<int32_t>());
    int32_t n = 0;
    DCU(cudaMemcpy(&n, cnt, sizeof(int32_t), cudaMemcpyDeviceToHost));
    int cur = 0, seed = min_x;
    while (n > 0) {
        DCU(cudaMemset(cnt + (cur ^ 1), 0, sizeof(int32_t)));
        flood_fill_kernel<<<n, 512>>>(dco.as<int32_t>(), dgrid.as<int32_t>(), dimx, dimy, dimz, obs.as<uint8_t>(), vis.as<uint8_t>(), list[cur], list[cur ^ 1],
                                      cnt + (cur ^ 1), queued.as<int32_t>(), seed);
        seed = -1;
        cur ^= 1;
        DCU(cudaMemcpy(&n, cnt + cur, sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    ff_finish_kernel<<<nb, 512>>>(obs.as<uint8_t>(), vis.as<uint8_t>(), filled.as<unsigned long long>());
    DCU(cudaGetLastError());
    unsigned long long f = 0;
    DCU(cudaMemcpy(&f, filled.p, sizeof(f), cudaMemcpyDeviceToHost));
    DCU(cudaMemcpy(obstacle, obs.p, nc, cudaMemcpyDeviceToHost));
    return (int64_t)f;
}

int64_t ludwig_domain_wall_distance(int device, const int32_t* neighbor_table, int32_t nb, const uint8_t* obstacle, double dx, float* wall_dist) {
    if (!neighbor_table || nb <= 0 || !obstacle || !(dx > 0) || !wall_dist) { g_err = "bad wall-distance arguments"; return LUDWIG_EINVAL; }
    DCU(cudaSetDevice(device));
    DevBuf nbr, obs, wd, cnt;
    const size_t nc = (size_t)nb * 512;
    DCU(nbr.alloc((size_t)nb * 27 * sizeof(int32_t))); DCU(obs.alloc(nc)); DCU(wd.alloc(nc * sizeof(float))); DCU(cnt.alloc(sizeof(unsigned long long)));
    DCU(cudaMemcpy(nbr.p, neighbor_table, (size_t)nb * 27 * sizeof(int32_t), cudaMemcpyHostToDevice));
    DCU(cudaMemcpy(obs.p, obstacle, nc, cudaMemcpyHostToDevice));
    DCU(cudaMemcpy(wd.p, wall_dist, nc * sizeof(float), cudaMemcpyHostToDevice));
    DCU(cudaMemset(cnt.p, 0, sizeof(unsigned long long)));
    wall_distance_kernel<<<(unsigned)((nc + 255) / 256), 256>>>(nbr.as<int32_t>(), nb, obs.as<uint8_t>(), (float)dx, wd.as<float>(), cnt.as<unsigned long long>());
    DCU(cudaGetLastError());
    unsigned long long n = 0;
    DCU(cudaMemcpy(&n, cnt.p, sizeof(n), cudaMemcpyDeviceToHost));
    DCU(cudaMemcpy(wall_dist, wd.p, nc * sizeof(float), cudaMemcpyDeviceToHost));
    return (int64_t)n;
}

int64_t ludwig_domain_qmap(int device, const double* tris, int64_t n_tri, const double offset[3], double dx, const int32_t* coords, int32_t nb,
                           const int32_t* grid_ptr, int32_t dimx, int32_t dimy, int32_t dimz, int64_t capacity, int32_t* out_cells, double* out_q,
                           int32_t* out_tri) {
    if (!args_ok(tris, n_tri, offset, dx, coords, nb, grid_ptr, dimx, dimy, dimz)) return LUDWIG_EINVAL;
    DCU(cudaSetDevice(device));
    const bool cached = g_qmap && g_qmap->k_tris == tris && g_qmap->k_coords == coords && g_qmap->k_ntri == n_tri && g_qmap->k_nb == nb &&
                        g_qmap->k_dx == dx && g_qmap->k_dev == device;
    if (!cached) {
        delete g_qmap;
        g_qmap = new QmapState();
        QmapState& S = *g_qmap;
        int rc = build_tri_map(S.M, tris, n_tri, offset, dx, coords, nb, grid_ptr, dimx, dimy, dimz, dx * 2.5, true);
        if (rc) { delete g_qmap; g_qmap = nullptr; return rc; }
        const int nw = S.M.n_blocks_with;
        if (nw > 0) {
            DCU(S.flags.alloc((size_t)nw * 512)); DCU(S.blk_count.alloc((size_t)nw * sizeof(int32_t))); DCU(S.blk_first.alloc(((size_t)nw + 1) * sizeof(int32_t)));
            qmap_kernel<<<nw, 512>>>(S.M.g, S.M.blocks.as<int32_t>(), S.M.start.as<int32_t>(), S.M.list.as<int32_t>(), S.flags.as<uint8_t>(),
                                     S.blk_count.as<int32_t>(), nullptr, nullptr, nullptr, nullptr);
            scan_kernel<<<1, 1024>>>(S.blk_count.as<int32_t>(), S.blk_first.as<int32_t>(), nw);
            int32_t n = 0;
            DCU(cudaMemcpy(&n, S.blk_first.as<int32_t>() + nw, sizeof(int32_t), cudaMemcpyDeviceToHost));
            S.n_rows = n;
        }
        S.k_tris = tris; S.k_coords = coords; S.k_ntri = n_tri; S.k_nb = nb; S.k_dx = dx; S.k_dev = device;
    }
    QmapState& S = *g_qmap;
    const int64_t n = S.n_rows;
    if (!out_cells || !out_q || !out_tri || capacity < n) return n;   // counting call: the device state stays for the fill call
    if (n > 0) {
        DevBuf cells, q, tri;
        DCU(cells.alloc((size_t)n * 4 * sizeof(int32_t))); DCU(q.alloc((size_t)n * 27 * sizeof(double))); DCU(tri.alloc((size_t)n * 27 * sizeof(int32_t)));
        qmap_kernel<<<S.M.n_blocks_with, 512>>>(S.M.g, S.M.blocks.as<int32_t>(), S.M.start.as<int32_t>(), S.M.list.as<int32_t>(), S.flags.as<uint8_t>(),
                                                S.blk_count.as<int32_t>(), S.blk_first.as<int32_t>(), cells.as<int32_t>(), q.as<double>(), tri.as<int32_t>());
        DCU(cudaGetLastError());
        DCU(cudaMemcpy(out_cells, cells.p, (size_t)n * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost));
        DCU(cudaMemcpy(out_q, q.p, (size_t)n * 27 * sizeof(double), cudaMemcpyDeviceToHost));
        DCU(cudaMemcpy(out_tri, tri.p, (size_t)n * 27 * sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    delete g_qmap;
    g_qmap = nullptr;
    return n;
}

}  // extern "C"
