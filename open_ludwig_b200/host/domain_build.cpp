// domain_build.cpp — host-side restatement of the reference's one-off CPU domain build, the step BEFORE the
// hot path: it produces every table the kernels read.  It mirrors the kept Julia driver files
//   domain_topology.jl:9-52      get_active_blocks_for_level   (triangle-AABB block marking)
//   domain_generation.jl:10-112  triangle_intersects_aabb, build_block_triangle_map, voxelize_blocks!
//   domain_generation.jl:114-203 perform_flood_fill!
//   domain_generation.jl:205-289 smooth_sponge_profile, apply_sponge!
//   domain_generation.jl:371-431 compute_wall_distances!
//   bouzidi_setup.jl:12-54,64-166 + bouzidi_math.jl:9-102   q-map / triangle map ray casting
// in Float64 with the reference's operation order (build with -ffp-contract=off), because the resulting integer
// tables must be bit-exact (block / flood-fill / boundary-cell counts of the golden logs).
// Not part of the GPU hot path; called from open_ludwig_b200/host/domain.py through ctypes.
//
// Array conventions: obstacle/sponge/wall_dist are [nb][8][8][8] (= Julia [x,y,z,b]); coords are 1-based
// (bx,by,bz) triples sorted lexicographically; triangles are [n][3][3] Float64 in STL coordinates.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace {
constexpr int BS = 8;

struct V3 { double x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   // StaticArrays: left-to-right sum
inline double min3(double a, double b, double c) { return std::min(std::min(a, b), c); }
inline double max3(double a, double b, double c) { return std::max(std::max(a, b), c); }
inline V3 vtx(const double* tris, int64_t t, int v, const double* off) {
    const double* p = tris + (t * 3 + v) * 3;
    return {p[0] + off[0], p[1] + off[1], p[2] + off[2]};
}
inline int64_t key3(int bx, int by, int bz) { return ((int64_t)bx << 42) | ((int64_t)by << 21) | (int64_t)bz; }

// domain_generation.jl:10-32 (9-axis SAT: 3 box axes + 9 edge cross products, NO triangle-normal axis)
bool triangle_intersects_aabb(V3 center, V3 box_half, V3 v1, V3 v2, V3 v3) {
    const double tol = 1.001;
    V3 h{box_half.x * tol, box_half.y * tol, box_half.z * tol};
    V3 t1 = sub(v1, center), t2 = sub(v2, center), t3 = sub(v3, center);
    if (min3(t1.x, t2.x, t3.x) > h.x || max3(t1.x, t2.x, t3.x) < -h.x) return false;
    if (min3(t1.y, t2.y, t3.y) > h.y || max3(t1.y, t2.y, t3.y) < -h.y) return false;
    if (min3(t1.z, t2.z, t3.z) > h.z || max3(t1.z, t2.z, t3.z) < -h.z) return false;
    V3 f[3] = {sub(t2, t1), sub(t3, t2), sub(t1, t3)};
    V3 u[3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            V3 axis = cross(u[i], f[j]);
            if (dot(axis, axis) < 1e-10) continue;
            double p1 = dot(t1, axis), p2 = dot(t2, axis), p3 = dot(t3, axis);
            double r = h.x * std::fabs(axis.x) + h.y * std::fabs(axis.y) + h.z * std::fabs(axis.z);
            if (std::min(p1, std::min(p2, p3)) > r || std::max(p1, std::max(p2, p3)) < -r) return false;
        }
    return true;
}

// Per-block triangle lists (domain_generation.jl:34-72 with margin 2 dx; bouzidi_setup.jl:12-54 with 2.5 dx).
// Triangle indices are appended in ascending order, exactly like the reference's enumerate loop.
void build_block_triangle_map(const double* tris, int64_t n_tri, const int32_t* coords, int nb, double dx, const double* off,
                              double margin, bool bouzidi_variant, std::vector<std::vector<int32_t>>& out) {
    out.assign(nb, {});
    std::unordered_map<int64_t, int> lookup;
    lookup.reserve((size_t)nb * 2);
    for (int i = 0; i < nb; ++i) lookup[key3(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2])] = i;
    const double bsdx = BS * dx;
    for (int64_t t = 0; t < n_tri; ++t) {
        double mn[3], mx[3];
        for (int a = 0; a < 3; ++a) {
            const double* p = tris + t * 9;
            if (bouzidi_variant) {   // min/max of the raw vertices, then + offset (bouzidi_setup.jl:31-35)
                mn[a] = std::min(p[a], std::min(p[3 + a], p[6 + a])) + off[a];
                mx[a] = std::max(p[a], std::max(p[3 + a], p[6 + a])) + off[a];
            } else {                 // min/max of the offset vertices (domain_generation.jl:49-55)
                mn[a] = std::min(p[a] + off[a], std::min(p[3 + a] + off[a], p[6 + a] + off[a]));
                mx[a] = std::max(p[a] + off[a], std::max(p[3 + a] + off[a], p[6 + a] + off[a]));
            }
        }
        int lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            lo[a] = (int)std::floor((mn[a] - margin) / bsdx) + 1;
            hi[a] = (int)std::floor((mx[a] + margin) / bsdx) + 1;
        }
        for (int bz = std::max(1, lo[2]); bz <= hi[2]; ++bz)
            for (int by = std::max(1, lo[1]); by <= hi[1]; ++by)
                for (int bx = std::max(1, lo[0]); bx <= hi[0]; ++bx) {
                    auto it = lookup.find(key3(bx, by, bz));
                    if (it != lookup.end()) out[it->second].push_back((int32_t)t);
                }
    }
}

// bouzidi_math.jl:9-47 (Moeller-Trumbore, EPSILON 1e-9)
inline bool ray_triangle(V3 origin, V3 dir, V3 v1, V3 v2, V3 v3, double& t_out) {
    const double EPSILON = 1e-9;
    V3 edge1 = sub(v2, v1), edge2 = sub(v3, v1);
    V3 h = cross(dir, edge2);
    double a = dot(edge1, h);
    if (std::fabs(a) < EPSILON) return false;
    double f = 1.0 / a;
    V3 s = sub(origin, v1);
    double u = f * dot(s, h);
    if (u < 0.0 || u > 1.0) return false;
    V3 q = cross(s, edge1);
    double v = f * dot(dir, q);
    if (v < 0.0 || u + v > 1.0) return false;
    double t = f * dot(edge2, q);
    if (t > EPSILON) { t_out = t; return true; }
    return false;
}
}  // namespace

extern "C" {

// domain_topology.jl:9-52.  grid is uint8 [bx_max][by_max][bz_max] (C order, bx major), set to 1 where marked.
void ludwig_host_mark_surface_blocks(const double* tris, int64_t n_tri, const double* off, double dx, int bx_max, int by_max,
                                     int bz_max, uint8_t* grid) {
    const double margin = dx * 0.01;
    const double inv_bs_dx = 1.0 / (BS * dx);
    for (int64_t t = 0; t < n_tri; ++t) {
        double mn[3], mx[3];
        const double* p = tris + t * 9;
        for (int a = 0; a < 3; ++a) {
            mn[a] = std::min(p[a] + off[a], std::min(p[3 + a] + off[a], p[6 + a] + off[a]));
            mx[a] = std::max(p[a] + off[a], std::max(p[3 + a] + off[a], p[6 + a] + off[a]));
        }
        int lo[3], hi[3];
        const int lim[3] = {bx_max, by_max, bz_max};
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::max(1, (int)std::floor((mn[a] - margin) * inv_bs_dx) + 1);
            hi[a] = std::min((int)std::floor((mx[a] + margin) * inv_bs_dx) + 1, lim[a]);
        }
        for (int bz = lo[2]; bz <= hi[2]; ++bz)
            for (int by = lo[1]; by <= hi[1]; ++by)
                for (int bx = lo[0]; bx <= hi[0]; ++bx) grid[((size_t)(bx - 1) * by_max + (by - 1)) * bz_max + (bz - 1)] = 1;
    }
}

// domain_generation.jl:74-112
void ludwig_host_voxelize(const double* tris, int64_t n_tri, const int32_t* coords, int nb, double dx, const double* off,
                          uint8_t* obstacle) {
    std::vector<std::vector<int32_t>> map;
    build_block_triangle_map(tris, n_tri, coords, nb, dx, off, dx * 2, false, map);
    const V3 box_half{0.75 * dx, 0.75 * dx, 0.75 * dx};
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < nb; ++i) {
        const auto& rel = map[i];
        if (rel.empty()) continue;
        const int bx = coords[3 * i], by = coords[3 * i + 1], bz = coords[3 * i + 2];
        for (int lz = 1; lz <= BS; ++lz)
            for (int ly = 1; ly <= BS; ++ly)
                for (int lx = 1; lx <= BS; ++lx) {
                    V3 center{((bx - 1) * BS + lx - 0.5) * dx, ((by - 1) * BS + ly - 0.5) * dx, ((bz - 1) * BS + lz - 0.5) * dx};
                    bool is_shell = false;
                    for (int32_t tid : rel) {
                        if (triangle_intersects_aabb(center, box_half, vtx(tris, tid, 0, off), vtx(tris, tid, 1, off), vtx(tris, tid, 2, off))) {
                            is_shell = true;
                            break;
                        }
                    }
                    if (is_shell) obstacle[(size_t)i * 512 + (lz - 1) * 64 + (ly - 1) * 8 + (lx - 1)] = 1;
                }
    }
}

// domain_generation.jl:114-203.  block_ptr is the reference's [bx,by,bz] column-major pointer (1-based, 0 none).
// Returns the number of filled voxels.
// Same reachable set as the reference's cell-by-cell BFS from the non-obstacle cells of the min-bx blocks (6-connected),
// computed at two granularities: a block without any obstacle cell is internally connected, so it is visited as ONE node
// (reached as soon as any of its cells is, and it then reaches every non-obstacle cell on the facing layer of its six
// neighbours); only the blocks that hold obstacle cells (the surface shell, a few % of a level) are walked cell by cell.
int64_t ludwig_host_flood_fill(uint8_t* obstacle, const int32_t* coords, int nb, const int32_t* block_ptr, int dimx, int dimy, int dimz) {
    std::vector<uint8_t> has_obs(nb, 0), bvis(nb, 0);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; ++b) {
        uint8_t any = 0;
        for (int c = 0; c < 512; ++c) any |= obstacle[(size_t)b * 512 + c];
        has_obs[b] = any != 0;
    }
    std::vector<uint32_t> cell_slot(nb, 0xFFFFFFFFu);   // index of the block's 512 visited flags, shell blocks only
    size_t n_shell = 0;
    for (int b = 0; b < nb; ++b) if (has_obs[b]) cell_slot[b] = (uint32_t)n_shell++;
    std::vector<uint8_t> visited(n_shell * 512, 0);
    constexpr uint32_t BLOCK_NODE = 0x80000000u;
    std::vector<uint32_t> queue;
    queue.reserve((size_t)nb + n_shell * 64);
    auto visit_cell = [&](int b, int c) {
        uint8_t& v = visited[(size_t)cell_slot[b] * 512 + c];
        if (!v && !obstacle[(size_t)b * 512 + c]) { v = 1; queue.push_back(((uint32_t)b << 9) | (uint32_t)c); }
    };
    auto visit_block = [&](int b) {
        if (!bvis[b]) { bvis[b] = 1; queue.push_back(BLOCK_NODE | (uint32_t)b); }
    };
    int min_x = coords[0];
    for (int i = 0; i < nb; ++i) min_x = std::min(min_x, coords[3 * i]);
    for (int b = 0; b < nb; ++b)
        if (coords[3 * b] == min_x) {
            if (!has_obs[b]) visit_block(b);
            else for (int c = 0; c < 512; ++c) visit_cell(b, c);
        }
    const int ddx[6] = {1, -1, 0, 0, 0, 0}, ddy[6] = {0, 0, 1, -1, 0, 0}, ddz[6] = {0, 0, 0, 0, 1, -1};
    auto neighbour_block = [&](int b, int i) -> int {
        const int nbx = coords[3 * b] + ddx[i], nby = coords[3 * b + 1] + ddy[i], nbz = coords[3 * b + 2] + ddz[i];
        if (nbx < 1 || nbx > dimx || nby < 1 || nby > dimy || nbz < 1 || nbz > dimz) return -1;
        return block_ptr[(nbx - 1) + (size_t)dimx * ((nby - 1) + (size_t)dimy * (nbz - 1))] - 1;   // -1 = none
    };
    size_t head = 0;
    while (head < queue.size()) {
        const uint32_t cur = queue[head++];
        if (cur & BLOCK_NODE) {
            const int b = (int)(cur & ~BLOCK_NODE);
            for (int i = 0; i < 6; ++i) {
                const int t = neighbour_block(b, i);
                if (t < 0) continue;
                if (!has_obs[t]) { visit_block(t); continue; }
                // the 64 cells of the neighbour's layer that faces this block
                for (int u = 0; u < BS; ++u)
                    for (int v = 0; v < BS; ++v) {
                        const int x = ddx[i] ? (ddx[i] > 0 ? 0 : BS - 1) : u;
                        const int y = ddy[i] ? (ddy[i] > 0 ? 0 : BS - 1) : (ddx[i] ? u : v);
                        const int z = ddz[i] ? (ddz[i] > 0 ? 0 : BS - 1) : v;
                        visit_cell(t, z * 64 + y * 8 + x);
                    }
            }
            continue;
        }
        const int b = (int)(cur >> 9), c = (int)(cur & 511);
        const int lx = c & 7, ly = (c >> 3) & 7, lz = c >> 6;
        for (int i = 0; i < 6; ++i) {
            const int nx = lx + ddx[i], ny = ly + ddy[i], nz = lz + ddz[i];
            if (nx >= 0 && nx < BS && ny >= 0 && ny < BS && nz >= 0 && nz < BS) { visit_cell(b, nz * 64 + ny * 8 + nx); continue; }
            const int t = neighbour_block(b, i);
            if (t < 0) continue;
            if (!has_obs[t]) visit_block(t);
            else visit_cell(t, ((nz + BS) % BS) * 64 + ((ny + BS) % BS) * 8 + ((nx + BS) % BS));
        }
    }
    int64_t filled = 0;
#pragma omp parallel for schedule(static) reduction(+ : filled)
    for (int b = 0; b < nb; ++b) {
        if (!has_obs[b]) {
            if (!bvis[b]) { std::memset(obstacle + (size_t)b * 512, 1, 512); filled += 512; }
            continue;
        }
        const uint8_t* v = visited.data() + (size_t)cell_slot[b] * 512;
        for (int c = 0; c < 512; ++c)
            if (!obstacle[(size_t)b * 512 + c] && !v[c]) { obstacle[(size_t)b * 512 + c] = 1; ++filled; }
    }
    return filled;
}

// domain_generation.jl:205-289
static inline double smooth_sponge_profile(double x, double thickness) {
    if (x <= 0.0) return 1.0;
    if (x >= thickness) return 0.0;
    return 0.5 * (1.0 + std::cos(M_PI * x / thickness));
}
void ludwig_host_sponge(const int32_t* coords, int nb, double dx, double Lx, double Ly, double Lz, double sponge_thickness,
                        int symmetric, float* sponge) {
    const double outlet_thickness = Lx * std::max(sponge_thickness, 0.15);
    const double inlet_thickness = Lx * 0.02;
    const double y_t = Ly * sponge_thickness * 0.5, z_t = Lz * sponge_thickness * 0.5;
    const double outlet_start = Lx - outlet_thickness, y_top_start = Ly - y_t, z_back_start = Lz - z_t;
    const double outlet_strength = 1.0, inlet_strength = 0.05, wall_strength = 0.1;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nb; ++i) {
        const int bx = coords[3 * i], by = coords[3 * i + 1], bz = coords[3 * i + 2];
        for (int lz = 1; lz <= BS; ++lz)
            for (int ly = 1; ly <= BS; ++ly)
                for (int lx = 1; lx <= BS; ++lx) {
                    double px = ((bx - 1) * BS + lx - 0.5) * dx, py = ((by - 1) * BS + ly - 0.5) * dx, pz = ((bz - 1) * BS + lz - 0.5) * dx;
                    double v = 0.0;
                    if (px > outlet_start) v = std::max(v, smooth_sponge_profile(outlet_thickness - (px - outlet_start), outlet_thickness) * outlet_strength);
                    if (px < inlet_thickness) v = std::max(v, smooth_sponge_profile(px, inlet_thickness) * inlet_strength);
                    if (!symmetric && py < y_t) v = std::max(v, smooth_sponge_profile(py, y_t) * wall_strength);
                    if (py > y_top_start) v = std::max(v, smooth_sponge_profile(y_t - (py - y_top_start), y_t) * wall_strength);
                    if (pz < z_t) v = std::max(v, smooth_sponge_profile(pz, z_t) * wall_strength);
                    if (pz > z_back_start) v = std::max(v, smooth_sponge_profile(z_t - (pz - z_back_start), z_t) * wall_strength);
                    sponge[(size_t)i * 512 + (lz - 1) * 64 + (ly - 1) * 8 + (lx - 1)] = (float)v;
                }
    }
}

// domain_generation.jl:371-431.  Returns the number of near-wall cells (the reference's own counter is racy).
int64_t ludwig_host_wall_distance(const int32_t* coords, int nb, const uint8_t* obstacle, double dx, float* wall_dist) {
    std::unordered_map<int64_t, int> lookup;
    lookup.reserve((size_t)nb * 2);
    for (int i = 0; i < nb; ++i) lookup[key3(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2])] = i;
    int64_t count = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : count)
    for (int b = 0; b < nb; ++b) {
        const int bx = coords[3 * b], by = coords[3 * b + 1], bz = coords[3 * b + 2];
        int nbidx[27];
        for (int d = 0; d < 27; ++d) {
            auto it = lookup.find(key3(bx + d % 3 - 1, by + (d / 3) % 3 - 1, bz + d / 9 - 1));
            nbidx[d] = it == lookup.end() ? -1 : it->second;
        }
        for (int lz = 0; lz < BS; ++lz)
            for (int ly = 0; ly < BS; ++ly)
                for (int lx = 0; lx < BS; ++lx) {
                    if (obstacle[(size_t)b * 512 + lz * 64 + ly * 8 + lx]) continue;
                    bool near = false;
                    float min_dist = 100.0f;
                    for (int dz = -1; dz <= 1; ++dz)
                        for (int dy = -1; dy <= 1; ++dy)
                            for (int dxo = -1; dxo <= 1; ++dxo) {
                                if (dxo == 0 && dy == 0 && dz == 0) continue;
                                int nx = lx + dxo, ny = ly + dy, nz = lz + dz;
                                int ox = nx < 0 ? -1 : (nx >= BS ? 1 : 0), oy = ny < 0 ? -1 : (ny >= BS ? 1 : 0), oz = nz < 0 ? -1 : (nz >= BS ? 1 : 0);
                                int nb_i = nbidx[(ox + 1) + (oy + 1) * 3 + (oz + 1) * 9];
                                if (nb_i < 0) continue;
                                if (obstacle[(size_t)nb_i * 512 + ((nz + BS) % BS) * 64 + ((ny + BS) % BS) * 8 + ((nx + BS) % BS)]) {
                                    near = true;
                                    float dist = std::sqrt((float)(dxo * dxo + dy * dy + dz * dz)) * (float)dx;
                                    min_dist = std::min(min_dist, dist);
                                }
                            }
                    if (near) { wall_dist[(size_t)b * 512 + lz * 64 + ly * 8 + lx] = min_dist; ++count; }
                }
    }
    return count;
}

// bouzidi_setup.jl:64-166 + bouzidi_math.jl:53-102.  Two-call protocol: first with out_* = NULL to count, then to fill.
// Output is sparse: for every boundary cell (ordered by block, z, y, x = the reference's single-thread order) its
// 1-based (block,x,y,z), 27 Float64 q values and 27 1-based triangle indices (0 = none).
// Pruning: a hit with q <= 1 lies within one dx of the cell centre on every axis, so triangles whose offset AABB
// misses [centre - 1.001 dx, centre + 1.001 dx] cannot change the result (min_t of the reference is taken over
// all hits, but if its nearest hit has q > 1 every hit has).
// The caller asks twice (count, then fill): the counting call keeps its result for the fill call with the same arguments.
namespace {
struct QmapCache {
    const void *tris = nullptr, *coords = nullptr;
    int64_t n_tri = 0; int nb = 0; double dx = 0;
    std::vector<int32_t> cells, tri;
    std::vector<double> q;
    bool matches(const double* t, int64_t nt, const int32_t* c, int n, double d) const {
        return tris == t && coords == c && n_tri == nt && nb == n && dx == d && !cells.empty();
    }
    void clear() { tris = coords = nullptr; cells.clear(); cells.shrink_to_fit(); tri.clear(); tri.shrink_to_fit(); q.clear(); q.shrink_to_fit(); }
} g_qmap_cache;
}  // namespace

int64_t ludwig_host_qmap(const double* tris, int64_t n_tri, const int32_t* coords, int nb, double dx, const double* off,
                         int64_t capacity, int32_t* out_cells /*[n][4]*/, double* out_q /*[n][27]*/, int32_t* out_tri /*[n][27]*/) {
    if (out_cells && out_q && out_tri && g_qmap_cache.matches(tris, n_tri, coords, nb, dx)) {
        const int64_t n = (int64_t)g_qmap_cache.cells.size() / 4;
        if (capacity >= n) {
            std::memcpy(out_cells, g_qmap_cache.cells.data(), (size_t)n * 4 * sizeof(int32_t));
            std::memcpy(out_q, g_qmap_cache.q.data(), (size_t)n * 27 * sizeof(double));
            std::memcpy(out_tri, g_qmap_cache.tri.data(), (size_t)n * 27 * sizeof(int32_t));
            g_qmap_cache.clear();
            return n;
        }
    }
    g_qmap_cache.clear();
    std::vector<std::vector<int32_t>> map;
    build_block_triangle_map(tris, n_tri, coords, nb, dx, off, dx * 2.5, true, map);
    std::vector<std::vector<int32_t>> cells(nb);
    std::vector<std::vector<double>> qs(nb);
    std::vector<std::vector<int32_t>> tr(nb);
    double dirn[27][3], cmag[27];
    for (int k = 0; k < 27; ++k) {
        double c[3] = {(double)(k % 3 - 1), (double)((k / 3) % 3 - 1), (double)(k / 9 - 1)};
        double nrm = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        cmag[k] = nrm;
        for (int a = 0; a < 3; ++a) dirn[k][a] = nrm > 0 ? c[a] / nrm : 0.0;
    }
#pragma omp parallel for schedule(dynamic, 2)
    for (int b = 0; b < nb; ++b) {
        const auto& rel = map[b];
        if (rel.empty()) continue;
        const int bx = coords[3 * b], by = coords[3 * b + 1], bz = coords[3 * b + 2];
        // offset vertices + AABBs of the block's triangles
        std::vector<V3> v(rel.size() * 3);
        std::vector<double> bb(rel.size() * 6);
        for (size_t i = 0; i < rel.size(); ++i) {
            for (int j = 0; j < 3; ++j) v[i * 3 + j] = vtx(tris, rel[i], j, off);
            bb[i * 6 + 0] = min3(v[i * 3].x, v[i * 3 + 1].x, v[i * 3 + 2].x); bb[i * 6 + 1] = max3(v[i * 3].x, v[i * 3 + 1].x, v[i * 3 + 2].x);
            bb[i * 6 + 2] = min3(v[i * 3].y, v[i * 3 + 1].y, v[i * 3 + 2].y); bb[i * 6 + 3] = max3(v[i * 3].y, v[i * 3 + 1].y, v[i * 3 + 2].y);
            bb[i * 6 + 4] = min3(v[i * 3].z, v[i * 3 + 1].z, v[i * 3 + 2].z); bb[i * 6 + 5] = max3(v[i * 3].z, v[i * 3 + 1].z, v[i * 3 + 2].z);
        }
        std::vector<int> cand;
        for (int lz = 1; lz <= BS; ++lz)
            for (int ly = 1; ly <= BS; ++ly)
                for (int lx = 1; lx <= BS; ++lx) {
                    V3 cc{((bx - 1) * BS + lx - 0.5) * dx, ((by - 1) * BS + ly - 0.5) * dx, ((bz - 1) * BS + lz - 0.5) * dx};
                    const double r = dx * 1.001;
                    cand.clear();
                    for (size_t i = 0; i < rel.size(); ++i)
                        if (bb[i * 6] <= cc.x + r && bb[i * 6 + 1] >= cc.x - r && bb[i * 6 + 2] <= cc.y + r && bb[i * 6 + 3] >= cc.y - r &&
                            bb[i * 6 + 4] <= cc.z + r && bb[i * 6 + 5] >= cc.z - r)
                            cand.push_back((int)i);
                    if (cand.empty()) continue;
                    double qv[27];
                    int32_t tv[27];
                    bool any = false;
                    for (int k = 0; k < 27; ++k) {
                        qv[k] = 0.0; tv[k] = 0;
                        if (k == 13) continue;
                        V3 dir{dirn[k][0], dirn[k][1], dirn[k][2]};
                        double min_t = INFINITY;
                        int best = -1;
                        for (int i : cand) {
                            double t;
                            if (ray_triangle(cc, dir, v[i * 3], v[i * 3 + 1], v[i * 3 + 2], t) && t < min_t) { min_t = t; best = rel[i]; }
                        }
                        if (min_t < INFINITY) {
                            double q = min_t / (dx * cmag[k]);
                            if (q > 0.0 && q <= 1.0) { qv[k] = q; tv[k] = best + 1; any = true; }
                        }
                    }
                    if (any) {
                        cells[b].insert(cells[b].end(), {b + 1, lx, ly, lz});
                        qs[b].insert(qs[b].end(), qv, qv + 27);
                        tr[b].insert(tr[b].end(), tv, tv + 27);
                    }
                }
    }
    int64_t n = 0;
    for (int b = 0; b < nb; ++b) n += (int64_t)cells[b].size() / 4;
    if (!out_cells || !out_q || !out_tri || capacity < n) {
        if (n > 0) {   // counting call: keep the result for the fill call
            g_qmap_cache.tris = tris; g_qmap_cache.coords = coords; g_qmap_cache.n_tri = n_tri; g_qmap_cache.nb = nb; g_qmap_cache.dx = dx;
            g_qmap_cache.cells.reserve((size_t)n * 4); g_qmap_cache.q.reserve((size_t)n * 27); g_qmap_cache.tri.reserve((size_t)n * 27);
            for (int b = 0; b < nb; ++b) {
                g_qmap_cache.cells.insert(g_qmap_cache.cells.end(), cells[b].begin(), cells[b].end());
                g_qmap_cache.q.insert(g_qmap_cache.q.end(), qs[b].begin(), qs[b].end());
                g_qmap_cache.tri.insert(g_qmap_cache.tri.end(), tr[b].begin(), tr[b].end());
            }
        }
        return n;
    }
    int64_t o = 0;
    for (int b = 0; b < nb; ++b) {
        int64_t m = (int64_t)cells[b].size() / 4;
        if (!m) continue;
        std::memcpy(out_cells + o * 4, cells[b].data(), m * 4 * sizeof(int32_t));
        std::memcpy(out_q + o * 27, qs[b].data(), m * 27 * sizeof(double));
        std::memcpy(out_tri + o * 27, tr[b].data(), m * 27 * sizeof(int32_t));
        o += m;
    }
    return n;
}

// Dense q_map / tri_map of the reference layout ([k][b][z][y][x], bouzidi_setup.jl:82-85,128-129) from the sparse boundary-cell rows.
// q16 holds the Float16 bit patterns of the rows (converted by the caller), q > 0 selects the written entries.
void ludwig_host_scatter_qmap(const int32_t* cells /*[n][4] 1-based*/, const double* q /*[n][27]*/, const uint16_t* q16 /*[n][27]*/,
                              const int32_t* tri /*[n][27] or NULL*/, int64_t n, int64_t nb, uint16_t* q_map, int32_t* tri_map /* or NULL */) {
    const size_t plane = (size_t)nb * 512;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const size_t cell = (size_t)(cells[i * 4] - 1) * 512 + (size_t)(cells[i * 4 + 3] - 1) * 64 + (size_t)(cells[i * 4 + 2] - 1) * 8 + (size_t)(cells[i * 4 + 1] - 1);
        for (int k = 0; k < 27; ++k)
            if (q[i * 27 + k] > 0.0) {
                q_map[plane * k + cell] = q16[i * 27 + k];
                if (tri_map && tri) tri_map[plane * k + cell] = tri[i * 27 + k];
            }
    }
}

}  // extern "C"
