// abi.cu — the extern "C" entry points of libludwig_b200.so (include/ludwig_b200.h) and the host-side
// schedule that replaces solver_control.jl:21-165 / physics_v2.jl:26-97.
//
// Differences from the reference's schedule that do NOT change results:
//   * no host synchronisation between kernels (the reference blocks after every launch, physics_v2.jl:85,95);
//   * no copy_to_old! (blocks.jl:199-205): after an A-B step the input buffers ARE the old state, so only the
//     density is double-buffered on levels that have children (248 B/cell-update less traffic on parents);
//   * no dense f_post_collision (K2 is two-phase, see k_misc.cu).
#include <algorithm>
#include <parallel/algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <unordered_map>

#include <cuda_fp16.h>

#include "ludwig_internal.h"

using namespace ludwig;

namespace {

int fail(ludwig_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? LUDWIG_ENOMEM : LUDWIG_ECUDA,            \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                            \
    } while (0)

// Every copy goes through the context's (non-blocking) stream: a plain cudaMemcpy on the legacy stream is
// not ordered against kernels launched on ctx->stream.
inline cudaError_t memcpy_sync(cudaStream_t s, void* dst, const void* src, size_t n, cudaMemcpyKind kind) {
    cudaError_t e = cudaMemcpyAsync(dst, src, n, kind, s);
    return e == cudaSuccess ? cudaStreamSynchronize(s) : e;
}

template <typename T>
cudaError_t dalloc(ludwig_ctx* ctx, T** p, size_t n) {
    *p = nullptr;
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e == cudaSuccess) ctx->bytes += (int64_t)(n * sizeof(T));
    return e;
}

inline uint64_t spread3(uint32_t v) {   // 21 bits -> every third bit
    uint64_t x = v & 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}
inline uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) { return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2); }
inline uint64_t morton2(uint32_t x, uint32_t y) {   // 16 bits each
    auto s2 = [](uint64_t v) {
        v &= 0xffff; v = (v | v << 8) & 0x00ff00ffULL; v = (v | v << 4) & 0x0f0f0f0fULL; v = (v | v << 2) & 0x33333333ULL; v = (v | v << 1) & 0x55555555ULL;
        return v;
    };
    return s2(x) | (s2(y) << 1);
}

void free_level(Level* L) {
    if (!L) return;
    void* ptrs[] = {L->d_pack, L->d_unpack, L->d_export, L->d_fmirror, L->d_vmirror, L->d_moff_f[0], L->d_moff_f[1], L->d_moff_v[0], L->d_moff_v[1], L->d_roff_f[0], L->d_roff_f[1], L->d_roff_v[0], L->d_roff_v[1], L->d_link_cell, L->d_link_k, L->d_link_q, L->d_link_tmp, L->d_gstart, L->d_int2ref, L->d_nbr, L->d_bcoord, L->d_ptr, L->d_nbr_fast, L->d_gcoord, L->d_fghost, L->d_gcell, L->d_gmask, L->d_gcells8, L->d_list_plain, L->d_list_plain_g, L->d_list_feat, L->d_list_full, L->d_list_nonplain, L->d_list_plain_full, L->d_list_plain_xface, L->d_list_full_rest,
                    L->d_obstacle, L->d_sponge, L->d_wall_dist, L->d_f[0], L->d_f[1], L->d_vel[0], L->d_vel[1], L->d_rho[0],
                    L->d_rho[1], L->d_f_old, L->d_vel_old, L->d_rho_old};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete L;
}

// The calling rank's own base pointers in the per-rank tables (the other ranks' entries come from ludwig_ipc_attach).
void set_own_peers(ludwig_ctx* ctx, Level& L) {
    for (int i = 0; i < 2; ++i) { L.peer_f[i][ctx->rank] = L.d_f[i]; L.peer_vel[i][ctx->rank] = L.d_vel[i]; L.peer_rho[i][ctx->rank] = L.d_rho[i]; }
    L.peer_obstacle[ctx->rank] = L.d_obstacle;
}

PeerPtrs peers_of(const float* const (&tab)[MAX_RANKS]) {
    PeerPtrs p;
    for (int r = 0; r < MAX_RANKS; ++r) p.p[r] = tab[r];
    return p;
}

// Upload one reference-layout field [ncomp][nb_ref][512] into the internal layout [nb_int][ncomp][512],
// one component at a time through a 2 KiB-per-block staging buffer.
// Staging buffer of the level uploads / downloads: one component of a level in the caller's order.  Kept between calls, grown on
// demand, released by release_stage() when stepping starts (a 339 M-cell case would otherwise hold 1 GB for nothing).
int ensure_stage(ludwig_ctx* ctx, size_t n) {
    if (ctx->stage_floats >= n) return LUDWIG_OK;
    if (ctx->d_stage) { cudaFree(ctx->d_stage); ctx->d_stage = nullptr; ctx->stage_floats = 0; }
    CU(cudaMalloc((void**)&ctx->d_stage, n * sizeof(float)));
    ctx->stage_floats = n;
    return LUDWIG_OK;
}
void release_stage(ludwig_ctx* ctx) {
    if (ctx->d_stage) { cudaFree(ctx->d_stage); ctx->d_stage = nullptr; ctx->stage_floats = 0; }
}
int upload_field(ludwig_ctx* ctx, Level& L, const float* h_src, float* d_dst, int ncomp, bool local_order = false) {
    // global: the host array covers the whole level in reference order, d_int2ref picks the local blocks;
    // local_order: the host array holds only this rank's blocks, already in the library's internal order
    const size_t n = (size_t)(local_order ? L.nb : L.nb_global) * BS3;
    int rc = ensure_stage(ctx, n);
    if (rc) return rc;
    // stream order keeps component k's scatter ahead of component k + 1's copy into the same buffer: no host synchronisation per
    // component (a copy from pageable memory returns once the source has been read), one at the end to report errors
    for (int k = 0; k < ncomp; ++k) {
        CU(cudaMemcpyAsync(ctx->d_stage, h_src + n * k, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        launch_ref_to_int(ctx->d_stage, d_dst, local_order ? nullptr : L.d_int2ref, L.nb, ncomp, k, ctx->stream);
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, LUDWIG_ECUDA, std::string("upload_field: ") + cudaGetErrorString(e));
    return LUDWIG_OK;
}
int download_field(ludwig_ctx* ctx, Level& L, const float* d_src, float* h_dst, int ncomp, bool local_order = false) {
    const size_t n = (size_t)(local_order ? L.nb : L.nb_global) * BS3;   // global: blocks owned by other ranks are returned as zeros
    int rc = ensure_stage(ctx, n);
    if (rc) return rc;
    for (int k = 0; k < ncomp; ++k) {   // (a copy to pageable memory returns when it is complete: the buffer is free for the next component)
        if (!local_order && L.nb != L.nb_global) CU(cudaMemsetAsync(ctx->d_stage, 0, n * 4, ctx->stream));
        launch_int_to_ref(d_src, ctx->d_stage, local_order ? nullptr : L.d_int2ref, L.nb, ncomp, k, ctx->stream);
        CU(cudaMemcpyAsync(h_dst + n * k, ctx->d_stage, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, LUDWIG_ECUDA, std::string("download_field: ") + cudaGetErrorString(e));
    return LUDWIG_OK;
}

int ensure_explicit_old(ludwig_ctx* ctx, Level& L) {
    if (L.explicit_old) return LUDWIG_OK;
    size_t nc = (size_t)L.nb * BS3;
    CU(dalloc(ctx, &L.d_f_old, nc * Q));
    CU(dalloc(ctx, &L.d_vel_old, nc * 3));
    CU(dalloc(ctx, &L.d_rho_old, nc));
    CU(cudaMemsetAsync(L.d_f_old, 0, nc * Q * 4, ctx->stream));
    CU(cudaMemsetAsync(L.d_vel_old, 0, nc * 3 * 4, ctx->stream));
    launch_fill(L.d_rho_old, 1.0f, nc, ctx->stream);
    L.explicit_old = true;
    return LUDWIG_OK;
}

struct ParentView {
    const Level* P = nullptr;
    const float *f_new = nullptr, *f_old = nullptr, *rho_new = nullptr, *rho_old = nullptr, *vel_new = nullptr, *vel_old = nullptr;
    int in = 0, out = 1, rho_new_i = 0, rho_old_i = 0;   // buffer indices (identical on every rank: lock-step schedule)
    bool explicit_old = false;
};

// The parent's buffers as recursive_step_temporal! receives them (solver_control.jl:65-72):
// new = (f_out, level.rho, vel_out) of the parent's step `parent_t_sub`, old = its pre-step state.
ParentView make_parent_view(const Level& P, int64_t parent_t_sub, bool explicit_old) {
    ParentView v;
    v.P = &P;
    int in = (parent_t_sub % 2 == 0) ? 0 : 1, out = 1 - in;
    v.in = in; v.out = out; v.rho_new_i = P.rho_cur; v.rho_old_i = P.d_rho[1] ? 1 - P.rho_cur : P.rho_cur;
    v.f_new = P.d_f[out]; v.vel_new = P.d_vel[out]; v.rho_new = P.d_rho[P.rho_cur];
    if (explicit_old && P.explicit_old) { v.f_old = P.d_f_old; v.vel_old = P.d_vel_old; v.rho_old = P.d_rho_old; v.explicit_old = true; }
    else { v.f_old = P.d_f[in]; v.vel_old = P.d_vel[in]; v.rho_old = P.d_rho[v.rho_old_i]; }
    return v;
}

// Fast-mode tables of one level: ghost blocks for the missing IN-DOMAIN neighbours (refinement interface), the
// neighbour table that addresses them, the interface pre-pass work list and the three K1 work lists.  Built at the
// first fast step because "in-domain" depends on ludwig_params.domain_n{x,y,z} (physics_v2.jl:55-56).
int ensure_fast_tables(ludwig_ctx* ctx, Level& L, const ludwig_params& p) {
    if (L.fast_ready && L.fast_dom[0] == p.domain_nx && L.fast_dom[1] == p.domain_ny && L.fast_dom[2] == p.domain_nz) return LUDWIG_OK;
    CU(cudaStreamSynchronize(ctx->stream));
    for (void* q : {(void*)L.d_gstart, (void*)L.d_nbr_fast, (void*)L.d_gcoord, (void*)L.d_fghost, (void*)L.d_gcell, (void*)L.d_gmask, (void*)L.d_gcells8, (void*)L.d_list_plain,
                    (void*)L.d_list_plain_g, (void*)L.d_list_feat, (void*)L.d_list_full, (void*)L.d_list_nonplain, (void*)L.d_list_plain_full, (void*)L.d_list_plain_xface, (void*)L.d_list_full_rest})
        if (q) cudaFree(q);
    L.d_gstart = nullptr; L.d_nbr_fast = nullptr; L.d_gcoord = nullptr; L.d_fghost = nullptr; L.d_gcell = nullptr; L.d_gmask = nullptr; L.d_gcells8 = nullptr;
    L.d_list_plain = L.d_list_plain_g = L.d_list_feat = L.d_list_full = L.d_list_nonplain = L.d_list_plain_full = L.d_list_plain_xface = L.d_list_full_rest = nullptr;
    const int nb = L.nb;
    const int scale = 1 << (L.level_id - 1);
    const int ext[3] = {p.domain_nx * scale / BS, p.domain_ny * scale / BS, p.domain_nz * scale / BS};   // blocks per axis
    std::vector<int32_t> nbrf(L.h_nbr);
    std::vector<int32_t> gcoord;
    std::unordered_map<uint64_t, int> gid;
    auto key = [](int x, int y, int z) { return ((uint64_t)(uint32_t)x << 42) | ((uint64_t)(uint32_t)y << 21) | (uint64_t)(uint32_t)z; };
    if (L.level_id > 1) {
        for (int b = 0; b < nb; ++b)
            for (int d = 0; d < 27; ++d) {
                if (L.h_nbr[(size_t)b * 27 + d] >= 0) continue;
                int x = L.h_bcoord[(size_t)b * 4] + d % 3 - 1, y = L.h_bcoord[(size_t)b * 4 + 1] + (d / 3) % 3 - 1, z = L.h_bcoord[(size_t)b * 4 + 2] + d / 9 - 1;
                if (x < 0 || x >= ext[0] || y < 0 || y >= ext[1] || z < 0 || z >= ext[2]) continue;   // outside the domain: a face BC, not a ghost
                auto it = gid.find(key(x, y, z));
                int g;
                if (it == gid.end()) { g = (int)gid.size(); gid.emplace(key(x, y, z), g); gcoord.insert(gcoord.end(), {x, y, z, 0}); }
                else g = it->second;
                nbrf[(size_t)b * 27 + d] = nb + g;
            }
    }
    const int ng = (int)gid.size();
    // work list: ghost cell (g,c) must provide population k iff the cell that pulls it, c + c_k, lies in a real block
    std::vector<int32_t> gcell;
    std::vector<uint32_t> gmask;
    std::vector<uint8_t> gcells8;
    std::vector<int32_t> gstart(1, 0);
    auto real_at = [&](int x, int y, int z) {
        if (x < 0 || x >= L.dimx || y < 0 || y >= L.dimy || z < 0 || z >= L.dimz) return false;
        return L.h_ptr[x + (size_t)L.dimx * (y + (size_t)L.dimy * z)] >= 0;
    };
    for (int g = 0; g < ng; ++g) {
        const int gx = gcoord[(size_t)g * 4], gy = gcoord[(size_t)g * 4 + 1], gz = gcoord[(size_t)g * 4 + 2];
        bool realn[27];
        for (int d = 0; d < 27; ++d) realn[d] = real_at(gx + d % 3 - 1, gy + (d / 3) % 3 - 1, gz + d / 9 - 1);
        for (int q = 0; q < 64; ++q) {                  // 2x2x2 groups sharing one parent cell
            const int x0 = (q & 3) * 2, y0 = ((q >> 2) & 3) * 2, z0 = (q >> 4) * 2;
            if (x0 > 0 && x0 < 6 && y0 > 0 && y0 < 6 && z0 > 0 && z0 < 6) continue;
            uint32_t km = 0; uint8_t cm = 0;
            for (int m = 0; m < 8; ++m) {
                const int x = x0 + (m & 1), y = y0 + ((m >> 1) & 1), z = z0 + ((m >> 2) & 1);
                for (int k = 0; k < 27; ++k) {
                    if (k == 13) continue;
                    int dx = x + (k % 3 - 1), dy = y + ((k / 3) % 3 - 1), dz = z + (k / 9 - 1);
                    int ox = dx < 0 ? -1 : (dx > 7 ? 1 : 0), oy = dy < 0 ? -1 : (dy > 7 ? 1 : 0), oz = dz < 0 ? -1 : (dz > 7 ? 1 : 0);
                    if (ox == 0 && oy == 0 && oz == 0) continue;
                    if (realn[(ox + 1) + (oy + 1) * 3 + (oz + 1) * 9]) { km |= 1u << k; cm |= (uint8_t)(1u << m); }
                }
            }
            if (cm) { gcell.push_back(g * 64 + q); gmask.push_back(km); gcells8.push_back(cm); }
        }
        gstart.push_back((int32_t)gcell.size());
    }
    // K1 work lists
    std::vector<int32_t> lp, lg, le, le_g, lf, lx, lfr;   // lx: x-only face blocks (subset of lf), lfr = lf without them; le_g: feature blocks with a ghost neighbour
    const int nxg_level = p.domain_nx << (L.level_id - 1);
    for (int b = 0; b < nb; ++b) {
        bool all = true, ghost = false, xonly = true;
        const int bx = L.h_bcoord[(size_t)b * 4];
        for (int d = 0; d < 27; ++d) {
            int v = nbrf[(size_t)b * 27 + d];
            if (v < 0) {
                all = false;
                const int dx = d % 3 - 1;   // missing only beyond the inlet plane (block at x = 0) or the outlet plane (block ends at nx)
                if (dx == 0 || (dx < 0 && bx != 0) || (dx > 0 && (bx + 1) * BS != nxg_level)) xonly = false;
            } else if (v >= nb && v < REMOTE_BASE) ghost = true;
        }
        const uint32_t feat = (uint32_t)L.h_bcoord[(size_t)b * 4 + 3] & (BF_OBSTACLE | BF_SPONGE | BF_WALLDIST);
        if (all && !feat) (ghost ? lg : lp).push_back(b);
        else if (all) (ghost ? le_g : le).push_back(b);
        else { lf.push_back(b); (xonly && !feat && !ghost ? lx : lfr).push_back(b); }
    }
    // plain blocks without a remote neighbour first: they can run while the halo import is still in flight
    auto has_remote = [&](int32_t b) {
        for (int d = 0; d < 27; ++d) if (nbrf[(size_t)b * 27 + d] >= REMOTE_BASE) return true;
        return false;
    };
    // Order of the plain list in a multi-GPU run.  Packed-mirror mode needs "interior first" (n_plain_int).  With direct NVLink
    // pulls the default stays the Morton order (the configuration measured on 8 GPUs); LUDWIG_REMOTE_ORDER = first | last |
    // interleave moves / spreads the blocks that pull from a peer (unmeasured experiments: where do their longer load
    // latencies hide best?).
    const std::string order = ctx->use_mirror ? "last" : ctx->remote_order;
    L.n_plain_int = 0;
    if (ctx->world > 1 && order != "morton") {
        auto mid = std::stable_partition(lp.begin(), lp.end(), [&](int32_t b) { return !has_remote(b); });
        const size_t ni = (size_t)(mid - lp.begin()), nbd = lp.size() - ni;
        L.n_plain_int = (int)ni;
        if (order == "first") std::rotate(lp.begin(), mid, lp.end());
        else if (order == "interleave" && nbd > 0 && ni > 0) {
            std::vector<int32_t> in(lp.begin(), mid), bd(mid, lp.end()), out;
            out.reserve(lp.size());
            size_t a = 0, c = 0;
            for (size_t j = 0; j < lp.size(); ++j) {   // boundary block whenever its share of the list falls behind
                if (c < nbd && (a >= ni || c * lp.size() < (j + 1) * nbd)) out.push_back(bd[c++]);
                else out.push_back(in[a++]);
            }
            lp.swap(out);
        }
        if (order != "last") L.n_plain_int = 0;   // only the mirror mode splits the launch
    }
    L.n_ghost = ng; L.n_gcell = (int)gcell.size();
    L.n_feat_nog = (int)le.size();
    le.insert(le.end(), le_g.begin(), le_g.end());   // feature list = [no ghost neighbour ..., with ghost neighbour ...]
    L.n_plain = (int)lp.size(); L.n_plain_g = (int)lg.size(); L.n_feat = (int)le.size(); L.n_full = (int)lf.size();
    CU(dalloc(ctx, &L.d_nbr_fast, nbrf.size()));
    CU(memcpy_sync(ctx->stream, L.d_nbr_fast, nbrf.data(), nbrf.size() * 4, cudaMemcpyHostToDevice));
    if (ng > 0) {
        CU(dalloc(ctx, &L.d_gcoord, gcoord.size())); CU(dalloc(ctx, &L.d_fghost, (size_t)ng * Q * BS3));
        CU(dalloc(ctx, &L.d_gcell, gcell.size())); CU(dalloc(ctx, &L.d_gmask, gmask.size())); CU(dalloc(ctx, &L.d_gcells8, gcells8.size()));
        CU(memcpy_sync(ctx->stream, L.d_gcoord, gcoord.data(), gcoord.size() * 4, cudaMemcpyHostToDevice));
        CU(dalloc(ctx, &L.d_gstart, gstart.size()));
        CU(memcpy_sync(ctx->stream, L.d_gstart, gstart.data(), gstart.size() * 4, cudaMemcpyHostToDevice));
        CU(cudaMemsetAsync(L.d_fghost, 0, (size_t)ng * Q * BS3 * 4, ctx->stream));
        if (!gcell.empty()) {
            CU(memcpy_sync(ctx->stream, L.d_gcell, gcell.data(), gcell.size() * 4, cudaMemcpyHostToDevice));
            CU(memcpy_sync(ctx->stream, L.d_gmask, gmask.data(), gmask.size() * 4, cudaMemcpyHostToDevice));
            CU(memcpy_sync(ctx->stream, L.d_gcells8, gcells8.data(), gcells8.size(), cudaMemcpyHostToDevice));
        }
    }
    std::vector<int32_t> ln(lg);
    ln.insert(ln.end(), le.begin(), le.end()); ln.insert(ln.end(), lf.begin(), lf.end());
    std::sort(ln.begin(), ln.end());   // Morton order
    std::vector<int32_t> lpf(lp);
    lpf.insert(lpf.end(), lf.begin(), lf.end());
    std::sort(lpf.begin(), lpf.end());   // internal = spatial order
    std::vector<int32_t> lpx(lp);
    lpx.insert(lpx.end(), lx.begin(), lx.end());
    std::sort(lpx.begin(), lpx.end());
    L.n_xface = (int)lx.size();
    struct { std::vector<int32_t>* v; int32_t** d; } lists[8] = {{&lp, &L.d_list_plain}, {&lg, &L.d_list_plain_g}, {&le, &L.d_list_feat}, {&lf, &L.d_list_full}, {&ln, &L.d_list_nonplain}, {&lpf, &L.d_list_plain_full},
                                                                 {&lpx, &L.d_list_plain_xface}, {&lfr, &L.d_list_full_rest}};
    for (auto& l : lists) {
        CU(dalloc(ctx, l.d, l.v->size()));
        if (!l.v->empty()) CU(memcpy_sync(ctx->stream, *l.d, l.v->data(), l.v->size() * 4, cudaMemcpyHostToDevice));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    L.fast_dom[0] = p.domain_nx; L.fast_dom[1] = p.domain_ny; L.fast_dom[2] = p.domain_nz;
    L.fast_ready = true;
    if (ctx->verbose)
        fprintf(stderr, "[ludwig rank %d] level %d: %d local blocks (plain %d, plain+ghost %d, feature %d, full %d), %d remote, %d ghost blocks, %d ghost groups\n",
                ctx->rank, L.level_id, nb, L.n_plain, L.n_plain_g, L.n_feat, L.n_full, L.n_remote, ng, L.n_gcell);
    return LUDWIG_OK;
}

// Profiling brackets (ludwig_profile_enable): CUDA events on the main stream around one launch, tagged with a class:
// 0 K1 plain, 1 K1 plain+ghost, 2 K1 feature, 3 K1 full (missing neighbours), 4 interface pre-pass, 5 Bouzidi, 6 barrier,
// 7 whole level step.
constexpr int NCLS = 12;   // 8 halo unpack (on the halo stream), 9 halo pack
int prof_begin(ludwig_ctx* ctx, int cls, bool active, cudaStream_t st = nullptr) {
    if (!ctx->profiling || !active) return LUDWIG_OK;
    if (ctx->ev_used + 2 > ctx->ev_pool.size()) {
        cudaEvent_t e0, e1;
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        ctx->ev_pool.push_back(e0); ctx->ev_pool.push_back(e1);
    }
    if (ctx->ev_class.size() < ctx->ev_pool.size() / 2) ctx->ev_class.resize(ctx->ev_pool.size() / 2);
    ctx->ev_class[ctx->ev_used / 2] = cls | (ctx->prof_level << 4);
    CU(cudaEventRecord(ctx->ev_pool[ctx->ev_used], st ? st : ctx->stream));
    return LUDWIG_OK;
}
int prof_end(ludwig_ctx* ctx, bool active, int64_t cells, cudaStream_t st = nullptr) {
    if (!ctx->profiling || !active) return LUDWIG_OK;
    CU(cudaEventRecord(ctx->ev_pool[ctx->ev_used + 1], st ? st : ctx->stream));
    ctx->ev_used += 2;
    ctx->prof_cells += cells;
    return LUDWIG_OK;
}

// Cross-rank barrier after a level step: the registered callback if any, else the native peer-flag kernel.
// Profiling class 6 = device time spent in barriers (mostly waiting for the slowest rank of that level step).
// A failed barrier is STICKY and fatal for the context: the peer-flag kernel reports a time-out through a flag in
// mapped pinned host memory (read here without any synchronisation), a callback through its return value; from then
// on every stepping / result call returns LUDWIG_ESTATE instead of running kernels on unsynchronised peer memory.
int barrier_state(ludwig_ctx* ctx) {
    if (ctx->bar_failed || (ctx->h_bar_err && *(volatile int*)ctx->h_bar_err != 0)) {
        ctx->bar_failed = true;
        return fail(ctx, LUDWIG_ESTATE, "cross-GPU barrier failed (time-out: a peer stopped or the ranks issued different call sequences; or the "
                                        "barrier callback reported an error) - the context's state is no longer consistent");
    }
    return LUDWIG_OK;
}
int rank_barrier(ludwig_ctx* ctx) {
    if (ctx->world <= 1 || ctx->group_managed) return LUDWIG_OK;   // a ludwig_multi places its own (event) barriers
    int rc;
    if ((rc = barrier_state(ctx))) return rc;
    if ((rc = prof_begin(ctx, 6, true))) return rc;
    if (ctx->barrier_cb) {
        if (ctx->barrier_cb(ctx->barrier_user) != 0) { ctx->bar_failed = true; return barrier_state(ctx); }
    } else if (ctx->peers_attached && ctx->d_bar)
        launch_peer_barrier(ctx->peer_bar, ctx->d_bar, ctx->rank, ctx->world, ++ctx->bar_epoch, ctx->d_bar_err, ctx->d_bar_err_dev,
                            (long long)(ctx->barrier_timeout_s * 1e9), ctx->stream);
    return prof_end(ctx, true, 0);
}

// Active Bouzidi links of the local boundary cells for a given q_min (bouzidi_kernel.jl:36-38: q > q_min && q <= 1).
int ensure_bouzidi_links(ludwig_ctx* ctx, Level& L, float q_min) {
    if (L.links_qmin == q_min) return LUDWIG_OK;
    CU(cudaStreamSynchronize(ctx->stream));
    for (void* q : {(void*)L.d_link_cell, (void*)L.d_link_k, (void*)L.d_link_q, (void*)L.d_link_tmp})
        if (q) cudaFree(q);
    L.d_link_cell = nullptr; L.d_link_k = nullptr; L.d_link_q = nullptr; L.d_link_tmp = nullptr;
    std::vector<int32_t> lc; std::vector<uint8_t> lk; std::vector<float> lq;
    for (int i = 0; i < L.n_bc; ++i)
        for (int k = 0; k < 27; ++k) {
            __half_raw hr; hr.x = L.h_bc_q[(size_t)i * 27 + k];
            const float q = __half2float(__half(hr));
            if (q > q_min && q <= 1.0f) { lc.push_back(L.h_bc_cell[i]); lk.push_back((uint8_t)k); lq.push_back(q); }
        }
    L.n_links = (int)lc.size();
    {
        // Order the links by (block, direction, cell): the populations are stored direction-major inside a block, so
        // consecutive threads then touch the same 2 KiB direction plane and x-adjacent boundary cells share 32-byte sectors
        // (the per-cell order of the reference, 27 directions of one cell = 27 different sectors, is the worst case).
        // K2 has no write conflicts and the two-phase kernel no read-after-write either: any order gives the same bits.
        std::vector<int32_t> idx(lc.size());
        std::iota(idx.begin(), idx.end(), 0);
        __gnu_parallel::sort(idx.begin(), idx.end(), [&](int32_t x, int32_t y) {
            const int64_t kx = ((int64_t)(lc[x] >> 9) << 14) | ((int64_t)lk[x] << 9) | (lc[x] & 511);
            const int64_t ky = ((int64_t)(lc[y] >> 9) << 14) | ((int64_t)lk[y] << 9) | (lc[y] & 511);
            return kx < ky;
        });
        std::vector<int32_t> lc2(lc.size()); std::vector<uint8_t> lk2(lk.size()); std::vector<float> lq2(lq.size());
        for (size_t i = 0; i < idx.size(); ++i) { lc2[i] = lc[idx[i]]; lk2[i] = lk[idx[i]]; lq2[i] = lq[idx[i]]; }
        lc.swap(lc2); lk.swap(lk2); lq.swap(lq2);
    }
    if (L.n_links > 0) {
        CU(dalloc(ctx, &L.d_link_cell, lc.size())); CU(dalloc(ctx, &L.d_link_k, lk.size())); CU(dalloc(ctx, &L.d_link_q, lq.size()));
        CU(dalloc(ctx, &L.d_link_tmp, lq.size()));
        CU(memcpy_sync(ctx->stream, L.d_link_cell, lc.data(), lc.size() * 4, cudaMemcpyHostToDevice));
        CU(memcpy_sync(ctx->stream, L.d_link_k, lk.data(), lk.size(), cudaMemcpyHostToDevice));
        CU(memcpy_sync(ctx->stream, L.d_link_q, lq.data(), lq.size() * 4, cudaMemcpyHostToDevice));
    }
    L.links_qmin = q_min;
    return LUDWIG_OK;
}

// perform_timestep_v2! (physics_v2.jl:26-97): K1 then K2, in three phases so that the cross-rank barriers between them can be
// placed by the caller (one context: rank_barrier; several contexts driven by one thread: an event barrier over the group).
enum : int { PH_K1 = 1, PH_GATHER = 2, PH_FINISH = 4, PH_ALL = 7 };
// Set while a coarse step is being captured into a CUDA graph: the kernels then take the step counter and the inlet velocity from
// device memory (DynScalars) instead of immediates, so that the captured graph can be replayed for any later coarse step of the
// same buffer parity.
struct GraphCtl { const DynScalars* dyn; int64_t t_coarse; };
int step_level_phase(ludwig_ctx* ctx, Level& L, const ParentView* pv, int64_t t_sub, float tw, float u_curr, const ludwig_params& p, int phase,
                     const GraphCtl* gc = nullptr) {
    const int in = (t_sub % 2 == 0) ? 0 : 1, out = 1 - in;   // solver_control.jl:35-41
    ctx->prof_level = L.level_id - 1;
    int rho_out = L.rho_cur;
    if (L.d_rho[1]) rho_out = 1 - L.rho_cur;   // keep the pre-step density for the children
  if (phase & PH_K1) {
    // class 7: the whole level step on the main stream (its kernels, side-stream joins and barriers)
    ctx->p7 = (size_t)-1;
    if (ctx->profiling) {
        int rc7 = prof_begin(ctx, 7, true);
        if (rc7) return rc7;
        ctx->p7 = ctx->ev_used; ctx->ev_used += 2;
    }
    K1Args a{};
    a.f_in = L.d_f[in]; a.f_out = L.d_f[out];
    a.vel_in = L.d_vel[in]; a.vel_out = L.d_vel[out];
    a.rho_out = L.d_rho[rho_out];
    a.obstacle = L.d_obstacle; a.sponge = L.d_sponge; a.wall_dist = L.d_wall_dist;
    a.nbr = L.d_nbr; a.bcoord = L.d_bcoord; a.nb = L.nb; a.ghost_delta = 0;
    if (pv) {
        a.pf_new = pv->f_new; a.pf_old = pv->f_old; a.prho_new = pv->rho_new; a.prho_old = pv->rho_old;
        a.pvel_new = pv->vel_new; a.pvel_old = pv->vel_old;
        a.pptr = pv->P->d_ptr; a.pdimx = pv->P->dimx; a.pdimy = pv->P->dimy; a.pdimz = pv->P->dimz;
        a.tau_parent = pv->P->tau; a.is_l1 = 0;
    } else {
        a.pdimx = a.pdimy = a.pdimz = 1; a.tau_parent = 0.5f; a.is_l1 = 1;
    }
    const int scale = 1 << (L.level_id - 1);   // physics_v2.jl:55-56
    a.nxg = p.domain_nx * scale; a.nyg = p.domain_ny * scale; a.nzg = p.domain_nz * scale;
    a.tau = L.tau; a.c_wale = p.c_wale; a.nu_bg = p.nu_sgs_bg; a.u_inlet = u_curr; a.inlet_turb = p.inlet_turbulence;
    a.tw = tw; a.is_symmetric = p.symmetric; a.wm = p.wall_model_active;
    a.seed = (int)(t_sub % 1000000);           // physics_v2.jl:76
    a.use_temporal = p.use_temporal; a.sponge_blend = p.sponge_blend;
    if (gc) { a.dyn = gc->dyn; a.dyn_shift = L.level_id - 1; a.dyn_add = (int)(t_sub - (gc->t_coarse << (L.level_id - 1))); }

    a.roff_f = L.d_roff_f[in]; a.roff_v = L.d_roff_v[in];
    if (ctx->world > 1 && !ctx->peers_attached) return fail(ctx, LUDWIG_ESTATE, "multi-GPU context: call ludwig_ipc_attach before stepping");
    a.negzero = -0.0f; a.wm_c166 = ctx->wm_c166; a.strict_stash = ctx->opt_strict_variant; a.fast_variant = ctx->opt_fast_variant; a.num_sms = ctx->num_sms; a.prefetch_distance = ctx->opt_prefetch_distance; a.strict_occ = ctx->opt_strict_occ; a.strict_loop = ctx->opt_strict_loop; a.strict_feat_occ = ctx->opt_strict_feat_occ; a.cta_threads = ctx->opt_cta_threads ? ctx->opt_cta_threads : (p.strict_fp ? 64 : 128);   // measured best (profiles/README.md)
    const bool strict = p.strict_fp != 0;
    if (strict && ctx->opt_strict_generic) {
        // cross-check path (option "strict_generic"): the one-thread-per-cell kernel with every branch of the reference
        // inside it, in-kernel interface interpolation; single rank only
        if (ctx->world > 1) return fail(ctx, LUDWIG_ESTATE, "option strict_generic is single-GPU only");
        a.list = nullptr; a.n_list = L.nb;
        launch_k1_generic_strict(a, ctx->stream);
        ctx->launches += 1;
    } else {
        // Both FP modes share one schedule: ghost blocks + interface pre-pass, four K1 launch classes, remote neighbours through
        // peer offsets.  strict_fp selects the kernels compiled in the reference's operation order (k1_strict.cu).
        void (*const k_plain)(const K1Args&, cudaStream_t) = strict ? launch_k1s_plain : launch_k1_plain;
        void (*const k_plain_g)(const K1Args&, cudaStream_t) = strict ? launch_k1s_plain_ghost : launch_k1_plain_ghost;
        void (*const k_feat)(const K1Args&, cudaStream_t) = strict ? launch_k1s_feat : launch_k1_feat;
        void (*const k_full)(const K1Args&, cudaStream_t) = strict ? launch_k1s_full : launch_k1_full;
        int rc = ensure_fast_tables(ctx, L, p);
        if (rc) return rc;
        a.nbr = L.d_nbr_fast;
        a.ghost_delta = L.d_fghost ? (long long)(L.d_fghost - L.d_f[in]) : 0;
        bool overlap_pre = false;
        if (L.n_gcell > 0 && pv) {   // interface halo pre-pass: fills the ghost blocks K1 is about to pull from
            GhostArgs g{};
            g.gcell = L.d_gcell; g.gmask = L.d_gmask; g.gcells8 = L.d_gcells8; g.n = L.n_gcell; g.gcoord = L.d_gcoord; g.f_ghost = L.d_fghost;
            g.gstart = L.d_gstart; g.n_ghost = L.n_ghost;
            const Level& P = *pv->P;
            g.pf_new = peers_of(P.peer_f[pv->out]); g.pvel_new = peers_of(P.peer_vel[pv->out]); g.prho_new = peers_of(P.peer_rho[pv->rho_new_i]);
            g.pf_old = peers_of(P.peer_f[pv->in]); g.pvel_old = peers_of(P.peer_vel[pv->in]); g.prho_old = peers_of(P.peer_rho[pv->rho_old_i]);
            if (pv->explicit_old) {   // fine-grained API with real copy_to_old! buffers (single rank only)
                g.pf_old.p[ctx->rank] = pv->f_old; g.pvel_old.p[ctx->rank] = pv->vel_old; g.prho_old.p[ctx->rank] = pv->rho_old;
            }
            g.pptr = pv->P->d_ptr; g.pdimx = pv->P->dimx; g.pdimy = pv->P->dimy; g.pdimz = pv->P->dimz;
            g.tau = L.tau; g.tau_parent = pv->P->tau; g.tw = tw; g.use_temporal = p.use_temporal;
            // The plain K1 launch never touches a ghost block: with side streams the pre-pass runs on its own stream
            // CONCURRENTLY with it, and only the launches that may pull from ghost blocks wait for it.  (The ghost buffer
            // was last read by the previous step's K1 launches, all ordered before ev_pre_fork on the main stream.)
            overlap_pre = ctx->use_side_streams && !ctx->serial_prepass && L.n_plain > 0;
            if (overlap_pre) {
                CU(cudaEventRecord(ctx->ev_pre_fork, ctx->stream));
                CU(cudaStreamWaitEvent(ctx->pre_stream, ctx->ev_pre_fork, 0));
                if ((rc = prof_begin(ctx, 4, true, ctx->pre_stream))) return rc;
                if (strict) launch_ghost_interp_strict(g, ctx->pre_stream); else launch_ghost_interp(g, ctx->opt_block_prepass, ctx->pre_stream);
                if ((rc = prof_end(ctx, true, 0, ctx->pre_stream))) return rc;
                CU(cudaEventRecord(ctx->ev_pre, ctx->pre_stream));
            } else {
                if ((rc = prof_begin(ctx, 4, true))) return rc;
                if (strict) launch_ghost_interp_strict(g, ctx->stream); else launch_ghost_interp(g, ctx->opt_block_prepass, ctx->stream);
                if ((rc = prof_end(ctx, true, 0))) return rc;
            }
            ctx->launches += 1;
        }
        // Multi-GPU halo import: refresh the local mirrors of the remote layers K1 pulls from, on its own stream, while the
        // plain blocks without a remote neighbour (the first n_plain_int of the list) are processed.
        const bool mirror = ctx->world > 1 && ctx->use_mirror && L.n_unpack > 0;
        bool wait_halo = false;
        if (mirror) {
            a.roff_f = L.d_moff_f[in]; a.roff_v = L.d_moff_v[in];
            if (ctx->use_side_streams && L.n_plain_int > 0) {
                CU(cudaEventRecord(ctx->ev_halo_fork, ctx->stream));
                CU(cudaStreamWaitEvent(ctx->halo_stream, ctx->ev_halo_fork, 0));
                if ((rc = prof_begin(ctx, 8, true, ctx->halo_stream))) return rc;
                launch_halo_unpack(L, in, ctx->halo_stream);
                if ((rc = prof_end(ctx, true, 0, ctx->halo_stream))) return rc;
                CU(cudaEventRecord(ctx->ev_halo, ctx->halo_stream));
                wait_halo = true;
            } else {
                if ((rc = prof_begin(ctx, 8, true))) return rc;
                launch_halo_unpack(L, in, ctx->stream);
                if ((rc = prof_end(ctx, true, 0))) return rc;
            }
            ctx->launches += 1;
        }
        // The (up to four) K1 launches of a level step read f_in / vel_in and write disjoint blocks of f_out: on small
        // levels, where each of them is a few waves of latency-bound CTAs, they run concurrently on side streams.
        const bool fork = ctx->use_side_streams && L.nb <= ctx->fork_max_blocks;
        // LUDWIG_FORK_FULL=1 (experiment for large levels, unmeasured): the domain-face blocks (128 registers, latency-bound,
        // 8 GLUPS on the bench box's inlet / outlet faces) run on a side stream UNDER the HBM-bound plain launch
        // face_persist: on a level whose plain launch dwarfs its domain-face class (the bench box: 3 % of the cells on the inlet / outlet
        // planes) that class runs as a few persistent CTAs per SM launched BEFORE the plain kernel, so that it shares every SM with
        // the HBM-bound plain launch for its whole run instead of trailing it as waves of latency-bound CTAs on an otherwise idle GPU
        const bool persist_full = ctx->opt_face_persist > 0 && ctx->use_side_streams && L.n_full > 0 && (long long)L.n_plain > 15LL * L.n_full &&
                                  (long long)L.n_full * 4 > 2LL * ctx->opt_face_persist * ctx->num_sms;
        const bool fork_full = persist_full || (!fork && ctx->fork_full && ctx->use_side_streams && L.n_full > 0 && L.n_plain > 0);
        int used = 0;
        if (fork || fork_full) CU(cudaEventRecord(ctx->ev_fork, ctx->stream));
        auto launch_on = [&](void (*fn)(const K1Args&, cudaStream_t), const int32_t* list, int n, bool main_stream, bool ghosts) -> int {
            if (n <= 0) return LUDWIG_OK;
            a.list = list; a.n_list = n;
            a.persist_grid = (persist_full && fn == k_full) ? ctx->opt_face_persist * ctx->num_sms : 0;
            if ((strict ? a.strict_stash : a.fast_variant) == 2 && !a.persist_grid) {   // persistent variant: its ticket counter (one per launch class)
                const int cls = fn == k_plain ? 0 : fn == k_plain_g ? 1 : fn == k_feat ? 2 : 3;
                const int grid = n < 2 * ctx->num_sms ? n : 2 * ctx->num_sms;
                a.ticket = ctx->d_ticket + cls; a.ticket_base = ctx->ticket_base[cls];
                ctx->ticket_base[cls] += (unsigned long long)n + (unsigned long long)grid;   // every CTA draws one ticket past the end
            }
            if ((!fork && !(fork_full && fn == k_full)) || main_stream) {
                if (overlap_pre && ghosts) { CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_pre, 0)); overlap_pre = false; }
                fn(a, ctx->stream);
            } else {
                cudaStream_t st = ctx->side[used];
                CU(cudaStreamWaitEvent(st, ctx->ev_fork, 0));
                if (overlap_pre) CU(cudaStreamWaitEvent(st, ctx->ev_pre, 0));
                if (wait_halo) CU(cudaStreamWaitEvent(st, ctx->ev_halo, 0));
                fn(a, st);
                CU(cudaEventRecord(ctx->ev_join[used], st));
                ++used;
            }
            ctx->launches += 1;
            return LUDWIG_OK;
        };
        // merge_face: on a level without an interface pre-pass (level 1, the bench box) domain-face blocks ride in the plain launch
        // (a CTA-uniform branch picks the body) instead of trailing it as waves of latency-bound CTAs on an otherwise idle GPU.
        // Fast mode merges the whole class.  Strict mode merges only the X-ONLY face blocks (nothing missing but beyond the inlet /
        // outlet plane, no feature: k1_strict.cu), whose body is as lean as the plain one, and keeps a separate launch for the rest:
        // with the general strict face body in the mix the step got 2.5 % SLOWER (its CTAs live ~4x longer under the plain CTAs'
        // memory traffic and hold a third of the CTA slots meanwhile).  Numbers: profiles/README.md.
        const int n_face_merged = strict ? L.n_xface : L.n_full;
        const bool merged = ctx->opt_merge_face != 0 && !wait_halo && !fork_full && L.n_gcell == 0 && n_face_merged > 0 && L.n_plain > 0 &&
                            (strict ? a.strict_stash == 0 && a.cta_threads == 64 && a.strict_occ == 5 && a.strict_loop == 1 : a.fast_variant == 0 && a.cta_threads == 128);
        const int32_t* list_full = merged && strict ? L.d_list_full_rest : L.d_list_full;
        const int n_full_left = merged ? L.n_full - n_face_merged : L.n_full;
        if (fork_full && (rc = launch_on(k_full, L.d_list_full, L.n_full, false, true))) return rc;
        // feature_first (option, off): while the interface pre-pass runs on its own stream, the feature blocks that pull from no ghost
        // block go first — they are compute-bound, so the latency-bound pre-pass would get the memory system it crawls without beside
        // the plain launch.  Measured on the shipped bunny and wing (both FP modes): no effect (65.4 / 65.5 vs 65.3 / 65.8 ms): the GPU
        // is busy either way, the pre-pass is work, not a wait; reordering launches cannot remove it.
        const bool pc = ctx->profiling && !fork && !fork_full;   // (with profiling on and no forking, classes 1..3 are bracketed too)
        // (not with the persistent ticket-scheduled variants: two launches of one class would share a ticket counter)
        const int n_feat_first = (ctx->opt_feature_first && overlap_pre && L.n_plain > 0 && (strict ? a.strict_stash : a.fast_variant) != 2) ? L.n_feat_nog : 0;
        if (n_feat_first > 0) {   // main stream, no wait for the pre-pass: no entry of this part of the list has a ghost neighbour
            if ((rc = prof_begin(ctx, 2, pc))) return rc;
            if ((rc = launch_on(k_feat, L.d_list_feat, n_feat_first, true, false))) return rc;
            if ((rc = prof_end(ctx, pc, 0))) return rc;
        }
        if ((rc = prof_begin(ctx, 0, L.n_plain > 0))) return rc;
        if (merged) {
            if ((rc = launch_on(strict ? launch_k1s_mixed : launch_k1_mixed, strict ? L.d_list_plain_xface : L.d_list_plain_full, L.n_plain + n_face_merged, true, false))) return rc;
        } else if (wait_halo) {
            if ((rc = launch_on(k_plain, L.d_list_plain, L.n_plain_int, true, false))) return rc;
            CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_halo, 0));
            if ((rc = launch_on(k_plain, L.d_list_plain + L.n_plain_int, L.n_plain - L.n_plain_int, true, false))) return rc;
        } else if ((rc = launch_on(k_plain, L.d_list_plain, L.n_plain, true, false))) return rc;
        if ((rc = prof_end(ctx, L.n_plain > 0, (int64_t)(L.n_plain + (merged ? n_face_merged : 0)) * BS3))) return rc;
        if ((rc = prof_begin(ctx, 1, pc && L.n_plain_g > 0))) return rc;
        if ((rc = launch_on(k_plain_g, L.d_list_plain_g, L.n_plain_g, L.n_plain == 0, true))) return rc;
        if ((rc = prof_end(ctx, pc && L.n_plain_g > 0, 0))) return rc;
        if ((rc = prof_begin(ctx, 2, pc && L.n_feat > n_feat_first))) return rc;
        if ((rc = launch_on(k_feat, L.d_list_feat + n_feat_first, L.n_feat - n_feat_first, L.n_plain == 0 && L.n_plain_g == 0, true))) return rc;
        if ((rc = prof_end(ctx, pc && L.n_feat > n_feat_first, 0))) return rc;
        if ((rc = prof_begin(ctx, 3, pc && n_full_left > 0))) return rc;
        if (!fork_full && (rc = launch_on(k_full, list_full, n_full_left, L.n_plain == 0 && L.n_plain_g == 0 && L.n_feat == 0, true))) return rc;
        if ((rc = prof_end(ctx, pc && n_full_left > 0, 0))) return rc;
        for (int i = 0; i < used; ++i) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_join[i], 0));
        if (overlap_pre) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_pre, 0));   // nothing on the main stream consumed it yet
    }
  }   // PH_K1
    const bool mg = ctx->world > 1;
    // K2 reads f_out of x_ff cells that may belong to another GPU: K1 must be complete everywhere before the gather, and
    // every gather before any scatter (the same two-phase argument as on one GPU, across ranks): the caller puts a
    // cross-rank barrier between the phases.
    if ((phase & PH_GATHER) && L.bouzidi) {
        int rcp;
        int rcb = ensure_bouzidi_links(ctx, L, p.q_min_threshold);
        if (rcb) return rcb;
        if ((rcp = prof_begin(ctx, 5, L.n_links > 0))) return rcp;
        launch_bouzidi(L, L.d_f[out], L.d_roff_f[out], p.strict_fp != 0, 1, ctx->stream);
        if ((rcp = prof_end(ctx, L.n_links > 0, 0))) return rcp;
    }
    if (phase & PH_FINISH) {
        if (L.bouzidi) {
            int rcp;
            if ((rcp = prof_begin(ctx, 5, L.n_links > 0))) return rcp;
            launch_bouzidi(L, L.d_f[out], L.d_roff_f[out], p.strict_fp != 0, 2, ctx->stream);
            if ((rcp = prof_end(ctx, L.n_links > 0, 0))) return rcp;
            if (L.n_links > 0) ctx->launches += 2;
        }
        if (mg && ctx->use_mirror && L.n_pack > 0) {   // the layers the peers pull at the start of the next step of this level
            int rcq;
            if ((rcq = prof_begin(ctx, 9, true))) return rcq;
            launch_halo_pack(L, out, ctx->stream);
            if ((rcq = prof_end(ctx, true, 0))) return rcq;
            ctx->launches += 1;
        }
        L.rho_cur = rho_out;
        L.last_t_sub = t_sub;
    }
    CU(cudaGetLastError());
    return LUDWIG_OK;
}

// A group = the contexts one host thread steps in lock-step: a single context (one GPU, or one process per GPU with the
// peer-flag / callback barrier), or the N contexts of a ludwig_multi (one process, N GPUs or N virtual ranks on one GPU),
// whose barrier is a set of stream-ordered event waits: rank r's stream waits for an event recorded on every other rank's
// stream.  Nothing blocks the host and no kernel spins, so any number of virtual ranks can share one device.
struct Group {
    std::vector<ludwig_ctx*> c;
    std::vector<cudaEvent_t>* ev = nullptr;   // one event per rank (ludwig_multi); null for a single context
};
int group_barrier(Group& g) {
    if (g.c.size() == 1) return rank_barrier(g.c[0]);
    const size_t n = g.c.size();
    for (size_t r = 0; r < n; ++r) {
        ludwig_ctx* ctx = g.c[r];
        CU(cudaSetDevice(ctx->device));
        int rc = prof_begin(ctx, 6, true); if (rc) return rc;
        CU(cudaEventRecord((*g.ev)[r], ctx->stream));
    }
    for (size_t r = 0; r < n; ++r) {
        ludwig_ctx* ctx = g.c[r];
        CU(cudaSetDevice(ctx->device));
        for (size_t q = 0; q < n; ++q) if (q != r) CU(cudaStreamWaitEvent(ctx->stream, (*g.ev)[q], 0));
        int rc = prof_end(ctx, true, 0); if (rc) return rc;
    }
    return LUDWIG_OK;
}

// One level step of every context of the group, phase by phase, with the cross-rank barriers in between.
int group_step_level(Group& g, size_t lvl, const std::vector<ParentView>* pvs, int64_t t_sub, float tw, float u, const ludwig_params& p,
                     const GraphCtl* gc = nullptr) {
    const bool mg = g.c[0]->world > 1;
    const bool bz = g.c[0]->levels[lvl]->bouzidi;
    auto run = [&](int phase) -> int {
        for (size_t r = 0; r < g.c.size(); ++r) {
            ludwig_ctx* ctx = g.c[r];
            if (g.c.size() > 1) CU(cudaSetDevice(ctx->device));
            int rc = step_level_phase(ctx, *ctx->levels[lvl], pvs ? &(*pvs)[r] : nullptr, t_sub, tw, u, p, phase, gc);
            if (rc) { if (ctx != g.c[0]) g.c[0]->err = ctx->err; return rc; }
        }
        return LUDWIG_OK;
    };
    int rc;
    if (!mg || !bz) { if ((rc = run(PH_ALL))) return rc; }
    else {
        if ((rc = run(PH_K1))) return rc;
        if ((rc = group_barrier(g))) return rc;
        if ((rc = run(PH_GATHER))) return rc;
        if ((rc = group_barrier(g))) return rc;
        if ((rc = run(PH_FINISH))) return rc;
    }
    if (mg && (rc = group_barrier(g))) return rc;   // every rank finished this level step
    for (ludwig_ctx* ctx : g.c)
        if (ctx->p7 != (size_t)-1) { CU(cudaEventRecord(ctx->ev_pool[ctx->p7 + 1], ctx->stream)); ctx->p7 = (size_t)-1; }
    return LUDWIG_OK;
}

// Multi-GPU: build every lazily built host table (fast-mode lists, ghost blocks, compacted Bouzidi links) BEFORE the first
// cross-rank barrier of a call, so that no rank sits in a barrier kernel while a peer is still doing seconds of host work.
int prepare_tables(ludwig_ctx* ctx, const ludwig_params& p) {
    if (ctx->world <= 1) return LUDWIG_OK;
    for (Level* L : ctx->levels) {
        int rc = ensure_fast_tables(ctx, *L, p);
        if (rc) return rc;
        if (L->bouzidi && (rc = ensure_bouzidi_links(ctx, *L, p.q_min_threshold))) return rc;
    }
    return LUDWIG_OK;
}

// recursive_step! / recursive_step_temporal! (solver_control.jl:21-143)
int recursive_step(Group& g, size_t lvl, int64_t t_sub, const std::vector<ParentView>* pvs, float tw, float u, const ludwig_params& p,
                   const GraphCtl* gc = nullptr) {
    if (lvl >= g.c[0]->levels.size()) return LUDWIG_OK;
    int rc = group_step_level(g, lvl, pvs, t_sub, tw, u, p, gc);
    if (rc) return rc;
    if (lvl + 1 < g.c[0]->levels.size()) {
        std::vector<ParentView> me;
        for (ludwig_ctx* ctx : g.c) me.push_back(make_parent_view(*ctx->levels[lvl], t_sub, /*explicit_old=*/false));
        if ((rc = recursive_step(g, lvl + 1, 2 * t_sub, &me, 0.0f, u, p, gc))) return rc;
        if ((rc = recursive_step(g, lvl + 1, 2 * t_sub + 1, &me, 0.5f, u, p, gc))) return rc;
    }
    return LUDWIG_OK;
}

// ---- CUDA-graph replay of a coarse step (one context).  The reference synchronises the host after EVERY kernel
// (physics_v2.jl:85,95, solver_control.jl:164); this library launches asynchronously, but a multi-level coarse step is still 7-63
// level steps of up to seven small launches each plus the event fork / join traffic of the concurrent launch classes, and on small
// levels the launch path, not the GPU, sets the pace.  All kernel arguments of a coarse step repeat with period 2 (A-B buffer
// parity of level 1; finer levels take an even number of sub-steps) except the noise seed and the ramped inlet velocity, which the
// kernels read from device memory during replay (DynScalars).  So: capture the whole recursion once per (parity, parameter set)
// with stream capture — the side-stream forks become graph branches — and afterwards one tiny kernel (the two scalars) plus one
// cudaGraphLaunch per coarse step.  Bit-identical to the eager path (tests/test_graph_replay_gpu.py).
__global__ void set_dyn_kernel(DynScalars* d, long long t, float u) { d->t_coarse = t; d->u_inlet = u; }

void drop_graphs(ludwig_ctx* ctx) {
    for (auto& kv : ctx->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    ctx->graphs.clear();
}

bool graphs_enabled(ludwig_ctx* ctx, const ludwig_params& p) {
    if (ctx->world != 1 || ctx->profiling || ctx->graph_failed) return false;
    if ((p.strict_fp ? ctx->opt_strict_variant : ctx->opt_fast_variant) == 2) return false;   // persistent variants draw tickets: per-launch arguments
    if (p.strict_fp && ctx->opt_strict_generic) return false;
    if (ctx->opt_graphs == 0) return false;
    if (ctx->opt_graphs == 1) return true;
    return ctx->levels.size() >= 2;   // auto: multi-level cases (many small launches per coarse step); a single big level gains nothing
}

int graph_coarse_step(Group& g, int64_t t, float u_curr, const ludwig_params& p) {
    ludwig_ctx* ctx = g.c[0];
    if (!ctx->d_dyn) CU(cudaMalloc((void**)&ctx->d_dyn, sizeof(DynScalars)));
    if (std::memcmp(&ctx->graph_params, &p, sizeof(p)) != 0) { drop_graphs(ctx); ctx->graph_params = p; }
    // every lazily built table must exist before the capture starts (building them synchronises and allocates)
    for (Level* L : ctx->levels) {
        int rc = ensure_fast_tables(ctx, *L, p);
        if (rc) return rc;
        if (L->bouzidi && (rc = ensure_bouzidi_links(ctx, *L, p.q_min_threshold))) return rc;
    }
    uint64_t key = (uint64_t)(t & 1);
    for (size_t l = 0; l < ctx->levels.size(); ++l) key |= (uint64_t)(ctx->levels[l]->rho_cur & 1) << (1 + l);
    set_dyn_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_dyn, (long long)t, u_curr);
    auto it = ctx->graphs.find(key);
    if (it == ctx->graphs.end()) {
        const int64_t launches0 = ctx->launches;
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
        int rc = LUDWIG_OK;
        if (e == cudaSuccess) {
            GraphCtl gc{ctx->d_dyn, t};
            rc = recursive_step(g, 0, t, nullptr, 0.0f, u_curr, p, &gc);      // advances the host-side bookkeeping like a real step
            e = cudaStreamEndCapture(ctx->stream, &graph);
        }
        ludwig_ctx::GraphEntry ent;
        if (rc == LUDWIG_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&ent.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc != LUDWIG_OK || e != cudaSuccess || !ent.exec) {
            // capture is an optimisation: fall back to eager stepping for good.  The bookkeeping has advanced but nothing ran: undo it.
            cudaGetLastError();
            ctx->graph_failed = true;
            for (size_t l = 0; l < ctx->levels.size(); ++l) {
                Level& L = *ctx->levels[l];
                if (L.d_rho[1] && ((1ll << l) & 1)) L.rho_cur ^= 1;
                L.last_t_sub = (t << l) - 1;
            }
            ctx->launches = launches0;
            if (rc != LUDWIG_OK) return rc;
            return recursive_step(g, 0, t, nullptr, 0.0f, u_curr, p);
        }
        ent.launches = ctx->launches - launches0;
        it = ctx->graphs.emplace(key, ent).first;
        CU(cudaGraphLaunch(it->second.exec, ctx->stream));
        ctx->launches += 1;
        return LUDWIG_OK;
    }
    CU(cudaGraphLaunch(it->second.exec, ctx->stream));
    ctx->launches += it->second.launches + 1;
    ctx->graph_replays += 1;
    for (size_t l = 0; l < ctx->levels.size(); ++l) {     // the host-side bookkeeping of 2^l steps of level l
        Level& L = *ctx->levels[l];
        if (L.d_rho[1] && l == 0) L.rho_cur ^= 1;          // levels >= 2 take an even number of sub-steps: unchanged
        L.last_t_sub = ((t + 1) << l) - 1;
    }
    return LUDWIG_OK;
}

// execute_timestep_batch! (solver_control.jl:145-165) for a group
int group_step_batch(Group& g, int64_t t_start, int32_t batch_size, float u_curr, const ludwig_params& p) {
    for (ludwig_ctx* ctx : g.c) {
        if (g.c.size() > 1) CU(cudaSetDevice(ctx->device));
        if (ctx->world > 1 && !ctx->peers_attached) return fail(ctx, LUDWIG_ESTATE, "multi-GPU context: attach the peers first");
        release_stage(ctx);   // the upload / download staging buffer is not held while the case steps
        int rcp = prepare_tables(ctx, p);
        if (rcp) return rcp;
        // export the halo layers of the CURRENT state (it may have been uploaded or initialised since the last batch): level l's
        // first sub-step is t_start * 2^l and reads the buffer of that parity
        if (ctx->world > 1 && ctx->use_mirror)
            for (size_t l = 0; l < ctx->levels.size(); ++l) {
                const int64_t t0 = t_start << l;
                launch_halo_pack(*ctx->levels[l], (t0 % 2 == 0) ? 0 : 1, ctx->stream);
            }
    }
    // align the ranks first: a peer may still be uploading / initialising the state this rank is about to pull from
    int rc;
    if (g.c[0]->world > 1 && (rc = group_barrier(g))) return rc;
    for (int t_offset = 0; t_offset < batch_size; ++t_offset) {
        const int64_t t = t_start + t_offset;
        if (g.c.size() == 1 && graphs_enabled(g.c[0], p)) {
            if ((rc = graph_coarse_step(g, t, u_curr, p))) return rc;
        } else if ((rc = recursive_step(g, 0, t, nullptr, 0.0f, u_curr, p))) return rc;
    }
    return LUDWIG_OK;
}

// Remote-block offset tables, relative to the local buffers (K1 adds them to f_in / vel_in); last step of an attach.
int finish_attach(ludwig_ctx* ctx) {
    for (Level* Lp : ctx->levels) {
        Level& L = *Lp;
        set_own_peers(ctx, L);
        if (L.n_remote == 0) continue;
        for (int par = 0; par < 2; ++par) {
            std::vector<long long> of(L.n_remote), ov(L.n_remote);
            for (int i = 0; i < L.n_remote; ++i) {
                const int ow = L.remote_owner[i];
                of[i] = (long long)((L.peer_f[par][ow] + (size_t)L.remote_local[i] * Q * BS3) - L.d_f[par]);
                ov[i] = (long long)((L.peer_vel[par][ow] + (size_t)L.remote_local[i] * 3 * BS3) - L.d_vel[par]);
            }
            CU(dalloc(ctx, &L.d_roff_f[par], (size_t)L.n_remote)); CU(dalloc(ctx, &L.d_roff_v[par], (size_t)L.n_remote));
            CU(memcpy_sync(ctx->stream, L.d_roff_f[par], of.data(), of.size() * 8, cudaMemcpyHostToDevice));
            CU(memcpy_sync(ctx->stream, L.d_roff_v[par], ov.data(), ov.size() * 8, cudaMemcpyHostToDevice));
        }
        if (!ctx->use_mirror) continue;
        // halo mirrors: K1 addresses a local copy of the remote layers (halo_unpack_kernel refreshes it every level step)
        CU(dalloc(ctx, &L.d_fmirror, (size_t)L.n_remote * Q * BS3)); CU(dalloc(ctx, &L.d_vmirror, (size_t)L.n_remote * 3 * BS3));
        CU(cudaMemsetAsync(L.d_fmirror, 0, (size_t)L.n_remote * Q * BS3 * 4, ctx->stream));
        CU(cudaMemsetAsync(L.d_vmirror, 0, (size_t)L.n_remote * 3 * BS3 * 4, ctx->stream));
        for (int par = 0; par < 2; ++par) {
            std::vector<long long> mf(L.n_remote), mv(L.n_remote);
            for (int i = 0; i < L.n_remote; ++i) {
                mf[i] = (long long)((L.d_fmirror + (size_t)i * Q * BS3) - L.d_f[par]);
                mv[i] = (long long)((L.d_vmirror + (size_t)i * 3 * BS3) - L.d_vel[par]);
            }
            CU(dalloc(ctx, &L.d_moff_f[par], (size_t)L.n_remote)); CU(dalloc(ctx, &L.d_moff_v[par], (size_t)L.n_remote));
            CU(memcpy_sync(ctx->stream, L.d_moff_f[par], mf.data(), mf.size() * 8, cudaMemcpyHostToDevice));
            CU(memcpy_sync(ctx->stream, L.d_moff_v[par], mv.data(), mv.size() * 8, cudaMemcpyHostToDevice));
        }
    }
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->peers_attached = true;
    return LUDWIG_OK;
}

bool level_ok(const ludwig_ctx* ctx, int32_t level) { return ctx && level >= 0 && level < (int)ctx->levels.size(); }

}  // namespace

extern "C" {

const char* ludwig_backend_name(void) { return "cuda-sm100a"; }

int ludwig_ctx_create(ludwig_ctx** out, int device) {
    if (!out) return LUDWIG_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return LUDWIG_ECUDA;
    if (cudaSetDevice(device) != cudaSuccess) return LUDWIG_ECUDA;
    auto* ctx = new ludwig_ctx();
    ctx->device = device;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return LUDWIG_ECUDA; }
    {
        bool ok = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < 3 && ok; ++i)
            ok = cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&ctx->pre_stream, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_pre_fork, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_pre, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { delete ctx; return LUDWIG_ECUDA; }
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);   // the halo import must not queue behind the interior K1 CTAs
        ok = cudaStreamCreateWithPriority(&ctx->halo_stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_halo_fork, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { delete ctx; return LUDWIG_ECUDA; }
    }
    // Every behaviour switch is an explicit option (ludwig_ctx_set_option); the library reads no environment variable.
    if (cudaMalloc((void**)&ctx->d_ticket, 4 * sizeof(unsigned long long)) != cudaSuccess || cudaMemset(ctx->d_ticket, 0, 4 * sizeof(unsigned long long)) != cudaSuccess) {
        delete ctx;
        return LUDWIG_ENOMEM;
    }
    {   // the wall model's constant factor, with the device's own Float64 log2 / exp2 (what the strict kernels would compute per cell)
        launch_wall_model_constant((float*)ctx->d_ticket, ctx->stream);
        if (cudaMemcpyAsync(&ctx->wm_c166, ctx->d_ticket, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaMemset(ctx->d_ticket, 0, 4 * sizeof(unsigned long long)) != cudaSuccess) {
            delete ctx;
            return LUDWIG_ECUDA;
        }
    }
    if (cudaMalloc((void**)&ctx->d_stats, 4096 * 6 * sizeof(double)) != cudaSuccess ||
        cudaMallocHost((void**)&ctx->h_stats, 4096 * 6 * sizeof(double)) != cudaSuccess) {
        delete ctx;
        return LUDWIG_ENOMEM;
    }
    *out = ctx;
    return LUDWIG_OK;
}

int ludwig_ctx_destroy(ludwig_ctx* ctx) {
    if (!ctx) return LUDWIG_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ludwig_output_state_free(ctx);
    drop_graphs(ctx);
    if (ctx->d_dyn) cudaFree(ctx->d_dyn);
    for (void* q : ctx->ipc_opened) cudaIpcCloseMemHandle(q);
    for (Level* L : ctx->levels) free_level(L);
    if (ctx->d_bar) cudaFree(ctx->d_bar);
    if (ctx->h_bar_err) cudaFreeHost(ctx->h_bar_err);
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    release_stage(ctx);
    if (ctx->d_ticket) cudaFree(ctx->d_ticket);
    if (ctx->h_stats) cudaFreeHost(ctx->h_stats);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (int i = 0; i < 3; ++i) { if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]); if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]); }
    for (cudaStream_t st : {ctx->pre_stream, ctx->halo_stream}) if (st) cudaStreamDestroy(st);
    for (cudaEvent_t e : {ctx->ev_pre_fork, ctx->ev_pre, ctx->ev_halo, ctx->ev_halo_fork}) if (e) cudaEventDestroy(e);
    delete ctx;
    return LUDWIG_OK;
}

// Behaviour switches that have no counterpart in the reference (it has one device and one kernel per step).  Every one
// defaults to the configuration measured fastest; the alternatives are kept for A/B runs and cross-checks.
int ludwig_ctx_set_option(ludwig_ctx* ctx, const char* key, const char* value) {
    if (!ctx || !key || !value) return fail(ctx, LUDWIG_EINVAL, "set_option: null argument");
    const std::string k(key), v(value);
    const bool on = v == "1" || v == "true" || v == "on";
    const bool before_levels = ctx->levels.empty();
    drop_graphs(ctx);   // captured graphs bake the options in
    auto need_early = [&]() { return fail(ctx, LUDWIG_ESTATE, "option '" + k + "' must be set before the first ludwig_level_create"); };
    if (k == "prepass") {                       // "thread" (default) | "block": block-cooperative interface pre-pass
        if (v != "thread" && v != "block") return fail(ctx, LUDWIG_EINVAL, "prepass: thread | block");
        ctx->opt_block_prepass = v == "block";
    } else if (k == "serial_prepass") ctx->serial_prepass = on;          // pre-pass on the main stream instead of beside the plain K1 launch
    else if (k == "single_stream") ctx->use_side_streams = !on;          // no concurrent launches at all
    else if (k == "fork_full") ctx->fork_full = on;                      // domain-face K1 launch beside the plain launch on large levels
    else if (k == "fork_max_blocks") ctx->fork_max_blocks = atoi(value); // levels up to this size run their K1 classes concurrently
    else if (k == "strict_kernel") {            // where the strict K1 keeps the 27 pulled populations / how it loads them
        if (v == "reg") ctx->opt_strict_variant = 0; else if (v == "stash") ctx->opt_strict_variant = 1; else if (v == "tma") ctx->opt_strict_variant = 2;
        else return fail(ctx, LUDWIG_EINVAL, "strict_kernel: reg | stash | tma");
    } else if (k == "prefetch_distance") {      // K1: every CTA prefetches the lines of the list entry this many blocks ahead into L2 (0 = off)
        ctx->opt_prefetch_distance = atoi(value);
        if (ctx->opt_prefetch_distance < 0) return fail(ctx, LUDWIG_EINVAL, "prefetch_distance >= 0");
    } else if (k == "strict_occupancy") {       // strict K1 (64-thread CTAs): resident warps per SM / 4 -> register budget 128 / 96 / 80
        const int n = atoi(value);
        if (n != 4 && n != 5 && n != 6) return fail(ctx, LUDWIG_EINVAL, "strict_occupancy: 4 | 5 | 6");
        ctx->opt_strict_occ = n;
    } else if (k == "l2_fetch") {               // cudaLimitMaxL2FetchGranularity of the device: 32 | 64 | 128 bytes (a hint to the L2)
        const int n = atoi(value);
        if (n != 32 && n != 64 && n != 128) return fail(ctx, LUDWIG_EINVAL, "l2_fetch: 32 | 64 | 128");
        CU(cudaSetDevice(ctx->device));
        CU(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)n));
    } else if (k == "strict_feature_occupancy") {   // strict feature / domain-face K1: resident warps per SM / 4 -> 128 / 96 registers
        const int n = atoi(value);
        if (n != 3 && n != 4 && n != 5) return fail(ctx, LUDWIG_EINVAL, "strict_feature_occupancy: 3 | 4 | 5");
        ctx->opt_strict_feat_occ = n;
    } else if (k == "feature_first") ctx->opt_feature_first = on;         // feature blocks without a ghost neighbour before the plain launch, beside the pre-pass
    else if (k == "merge_face") ctx->opt_merge_face = on ? 1 : 0;   // domain-face blocks ride in the plain K1 launch (levels without a pre-pass)
    else if (k == "face_persist") {           // persistent CTAs per SM of the domain-face K1 class beside a much larger plain launch (0 = off)
        const int n = atoi(value);
        if (n < 0 || n > 8) return fail(ctx, LUDWIG_EINVAL, "face_persist: 0..8");
        ctx->opt_face_persist = n;
    } else if (k == "strict_loop") {            // strict plain K1 (64-thread CTAs, strict_occupancy 5): z-plane pairs of a block one CTA works through
        const int n = atoi(value);
        if (n != 1 && n != 2 && n != 4) return fail(ctx, LUDWIG_EINVAL, "strict_loop: 1 | 2 | 4");
        ctx->opt_strict_loop = n;
    } else if (k == "cta_threads") {            // threads per CTA of the non-persistent K1 kernels: a CTA takes 8 / 4 / 2 z-planes of a block
        const int n = v == "auto" ? 0 : atoi(value);
        if (n != 0 && n != 256 && n != 128 && n != 64) return fail(ctx, LUDWIG_EINVAL, "cta_threads: auto | 256 | 128 | 64");
        ctx->opt_cta_threads = n;
    } else if (k == "fast_kernel") {
        if (v == "direct") ctx->opt_fast_variant = 0; else if (v == "tma") ctx->opt_fast_variant = 2;
        else return fail(ctx, LUDWIG_EINVAL, "fast_kernel: direct | tma");
    }
    else if (k == "strict_generic") ctx->opt_strict_generic = on;        // strict mode through the one-thread-per-cell cross-check kernel
    else if (k == "graphs") { ctx->opt_graphs = v == "auto" ? -1 : (on ? 1 : 0); }   // CUDA-graph replay of coarse steps: auto (multi-level cases) | 0 | 1
    else if (k == "verbose") ctx->verbose = on;
    else if (k == "barrier_timeout_s") { ctx->barrier_timeout_s = atof(value); if (!(ctx->barrier_timeout_s > 0)) return fail(ctx, LUDWIG_EINVAL, "barrier_timeout_s > 0"); }
    else if (k == "halo_mirror") { if (!before_levels) return need_early(); ctx->use_mirror = on; }   // packed halo exchange into local mirrors
    else if (k == "partition") {                // "morton" (default: cost-weighted Morton ranges / aligned plan) | "rcb" | "rcb_yz"
        if (!before_levels) return need_early();
        if (v == "morton") ctx->partition_mode = 0; else if (v == "rcb") ctx->partition_mode = 1; else if (v == "rcb_yz") ctx->partition_mode = 2;
        else return fail(ctx, LUDWIG_EINVAL, "partition: morton | rcb | rcb_yz");
    } else if (k == "block_order") {            // internal block order within a rank: "morton" | "xslab<T>" (T = 2..64, default xslab8)
        if (!before_levels) return need_early();
        if (v == "morton") ctx->opt_block_order = 0;
        else if (v.rfind("xslab", 0) == 0 && atoi(value + 5) >= 2 && atoi(value + 5) <= 64) ctx->opt_block_order = atoi(value + 5);
        else return fail(ctx, LUDWIG_EINVAL, "block_order: morton | xslab<T> with T = 2..64");
    } else if (k == "remote_order") {           // where the blocks that pull from a peer sit in the plain launch
        if (v != "morton" && v != "first" && v != "last" && v != "interleave") return fail(ctx, LUDWIG_EINVAL, "remote_order: morton | first | last | interleave");
        for (Level* L : ctx->levels) if (L->fast_ready) return fail(ctx, LUDWIG_ESTATE, "remote_order must be set before the first step");
        ctx->remote_order = v;
    } else return fail(ctx, LUDWIG_EINVAL, "unknown option '" + k + "'");
    return LUDWIG_OK;
}

const char* ludwig_last_error(const ludwig_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int ludwig_num_levels(const ludwig_ctx* ctx) { return ctx ? (int)ctx->levels.size() : LUDWIG_EINVAL; }
int64_t ludwig_device_bytes(const ludwig_ctx* ctx) { return ctx ? ctx->bytes : 0; }

void* ludwig_ctx_stream(ludwig_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
// Bounds checks of our own (compute-sanitizer is not available on the target pool): every index table the kernels turn into
// addresses is re-derived / range-checked on the host.  All in-block offsets are bounded by construction (direction * 512 + cell
// < 27 * 512), so a kernel can only leave its buffers through a wrong BLOCK index or a wrong peer offset — exactly what is
// checked here: the neighbour tables (local / ghost / remote ranges), the remote tables against the owners' block counts, the
// peer offsets against offsets recomputed from the mapped base pointers, the rank-encoded block pointer, the work lists and the
// Bouzidi cells.  Returns the number of violations (0 = clean) or a negative error code; the first violation is in last_error.
int64_t ludwig_ctx_self_check(ludwig_ctx* ctx) {
    if (!ctx) return LUDWIG_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    int64_t bad = 0;
    auto flag = [&](const std::string& what) { if (bad++ == 0) ctx->err = "self-check: " + what; };
    auto fetch_i32 = [&](const int32_t* d, size_t n, std::vector<int32_t>& h) -> bool {
        h.resize(n);
        return n == 0 || memcpy_sync(ctx->stream, h.data(), d, n * sizeof(int32_t), cudaMemcpyDeviceToHost) == cudaSuccess;
    };
    for (size_t l = 0; l < ctx->levels.size(); ++l) {
        Level& L = *ctx->levels[l];
        const std::string at = " (level " + std::to_string(l + 1) + ")";
        if ((int)L.part_starts.size() != ctx->world + 1 || L.part_starts[ctx->rank + 1] - L.part_starts[ctx->rank] != L.nb) flag("partition ranges" + at);
        for (int i = 0; i < L.n_remote; ++i) {
            const int ow = L.remote_owner[i];
            if (ow < 0 || ow >= ctx->world || ow == ctx->rank) { flag("remote owner out of range" + at); continue; }
            if (L.remote_local[i] < 0 || L.remote_local[i] >= L.part_starts[ow + 1] - L.part_starts[ow]) flag("remote block index beyond the owner's blocks" + at);
        }
        std::vector<int32_t> h;
        const int32_t* tabs[2] = {L.d_nbr, L.d_nbr_fast};
        for (int ti = 0; ti < 2; ++ti) {
            if (!tabs[ti]) continue;
            if (!fetch_i32(tabs[ti], (size_t)L.nb * 27, h)) return fail(ctx, LUDWIG_ECUDA, "self-check: table download failed");
            for (size_t i = 0; i < h.size(); ++i) {
                const int32_t v = h[i];
                const bool ok = v == -1 || (v >= 0 && v < L.nb) || (ti == 1 && v >= L.nb && v < L.nb + L.n_ghost) || (v >= REMOTE_BASE && v - REMOTE_BASE < L.n_remote);
                if (!ok) { flag("neighbour table entry out of range" + at); break; }
                if (i % 27 == 13 && v != (int32_t)(i / 27)) { flag("neighbour table: direction 13 is not the block itself" + at); break; }
            }
        }
        if (ctx->peers_attached && L.n_remote > 0)
            for (int par = 0; par < 2; ++par) {
                std::vector<long long> of(L.n_remote), ov(L.n_remote);
                if (memcpy_sync(ctx->stream, of.data(), L.d_roff_f[par], of.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
                    memcpy_sync(ctx->stream, ov.data(), L.d_roff_v[par], ov.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
                    return fail(ctx, LUDWIG_ECUDA, "self-check: offset download failed");
                for (int i = 0; i < L.n_remote; ++i) {
                    const int ow = L.remote_owner[i];
                    if (ow < 0 || ow >= ctx->world || !L.peer_f[par][ow] || !L.peer_vel[par][ow]) { flag("peer buffer not mapped" + at); break; }
                    if (of[i] != (long long)((L.peer_f[par][ow] + (size_t)L.remote_local[i] * Q * BS3) - L.d_f[par]) ||
                        ov[i] != (long long)((L.peer_vel[par][ow] + (size_t)L.remote_local[i] * 3 * BS3) - L.d_vel[par])) { flag("peer offset does not address the remote block" + at); break; }
                }
            }
        if (!fetch_i32(L.d_ptr, (size_t)L.dimx * L.dimy * L.dimz, h)) return fail(ctx, LUDWIG_ECUDA, "self-check: pointer download failed");
        size_t n_ptr = 0;
        for (int32_t v : h) {
            if (v < 0) continue;
            ++n_ptr;
            const int ow = v >> PTR_RANK_SHIFT, loc = v & PTR_LOCAL_MASK;
            if (ow >= ctx->world || loc >= L.part_starts[ow + 1] - L.part_starts[ow]) { flag("block pointer entry out of range" + at); break; }
        }
        if (n_ptr != (size_t)L.nb_global) flag("block pointer does not hold every block once" + at);
        if (L.fast_ready) {
            const int32_t* lists[4] = {L.d_list_plain, L.d_list_plain_g, L.d_list_feat, L.d_list_full};
            const int counts[4] = {L.n_plain, L.n_plain_g, L.n_feat, L.n_full};
            std::vector<uint8_t> seen(L.nb, 0);
            for (int li = 0; li < 4; ++li) {
                if (!fetch_i32(lists[li], (size_t)counts[li], h)) return fail(ctx, LUDWIG_ECUDA, "self-check: list download failed");
                for (int32_t b : h) { if (b < 0 || b >= L.nb || seen[b]++) { flag("work lists are not a partition of the blocks" + at); break; } }
            }
            if (L.n_plain + L.n_plain_g + L.n_feat + L.n_full != L.nb) flag("work lists do not cover every block" + at);
        }
        for (int32_t c : L.h_bc_cell) if (c < 0 || c >= L.nb * BS3) { flag("Bouzidi cell out of range" + at); break; }
        if (L.n_links > 0) {
            if (!fetch_i32(L.d_link_cell, (size_t)L.n_links, h)) return fail(ctx, LUDWIG_ECUDA, "self-check: link download failed");
            for (int32_t c : h) if (c < 0 || c >= L.nb * BS3) { flag("Bouzidi link cell out of range" + at); break; }
        }
    }
    return bad;
}

int64_t ludwig_launch_count(const ludwig_ctx* ctx) { return ctx ? ctx->launches : 0; }
int64_t ludwig_graph_replays(const ludwig_ctx* ctx) { return ctx ? ctx->graph_replays : 0; }
int ludwig_profile_enable(ludwig_ctx* ctx, int32_t on) {
    if (!ctx) return LUDWIG_EINVAL;
    ctx->profiling = on != 0;
    return LUDWIG_OK;
}
int ludwig_profile_read(ludwig_ctx* ctx, double* ms_total, int64_t* launches, int64_t* cells) {
    if (!ctx) return LUDWIG_EINVAL;
    CU(cudaStreamSynchronize(ctx->stream));
    double tot = 0;
    int64_t n0 = 0;
    for (int c = 0; c < 8; ++c) ctx->prof_class_ms[c] = 0;
    ctx->prof_level_ms.assign(ctx->levels.size() * NCLS, 0.0);
    for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ctx->ev_pool[i], ctx->ev_pool[i + 1]));
        const int c = ctx->ev_class[i / 2] & 15, lv = ctx->ev_class[i / 2] >> 4;
        if (c < 8) ctx->prof_class_ms[c] += ms;
        if ((size_t)lv < ctx->levels.size() && c < NCLS) ctx->prof_level_ms[(size_t)lv * NCLS + c] += ms;
        if (c == 0) { tot += ms; ++n0; }
    }
    if (ms_total) *ms_total = tot;            // class 0 only: the dominant plain K1 kernel
    if (launches) *launches = n0;
    if (cells) *cells = ctx->prof_cells;
    ctx->ev_used = 0; ctx->prof_cells = 0;
    return LUDWIG_OK;
}

int ludwig_profile_classes(ludwig_ctx* ctx, double out[8]) {
    if (!ctx || !out) return LUDWIG_EINVAL;
    for (int c = 0; c < 8; ++c) out[c] = ctx->prof_class_ms[c];   // filled by the last ludwig_profile_read
    return LUDWIG_OK;
}

int ludwig_profile_levels(ludwig_ctx* ctx, double* out, int32_t capacity) {
    if (!ctx || !out || capacity < (int32_t)ctx->prof_level_ms.size()) return fail(ctx, LUDWIG_EINVAL, "profile_levels: need 12 doubles per level");
    for (size_t i = 0; i < ctx->prof_level_ms.size(); ++i) out[i] = ctx->prof_level_ms[i];   // [level][class], last ludwig_profile_read
    return LUDWIG_OK;
}

int ludwig_sync(ludwig_ctx* ctx) {
    if (!ctx) return LUDWIG_EINVAL;
    CU(cudaStreamSynchronize(ctx->stream));
    return barrier_state(ctx);
}

int ludwig_level_create(ludwig_ctx* ctx, const ludwig_level_desc* d, int32_t* out_index) {
    if (!ctx || !d) return LUDWIG_EINVAL;
    CU(cudaSetDevice(ctx->device));
    if (d->level_id != (int)ctx->levels.size() + 1) return fail(ctx, LUDWIG_ESTATE, "levels must be created in order 1..L");
    if (d->n_blocks <= 0) return fail(ctx, LUDWIG_EINVAL, "n_blocks must be > 0");
    if (d->dim_x <= 0 || d->dim_y <= 0 || d->dim_z <= 0) return fail(ctx, LUDWIG_EINVAL, "block_pointer extents must be > 0");
    if (!d->block_pointer || !d->neighbor_table || !d->map_x || !d->map_y || !d->map_z || !d->obstacle || !d->sponge || !d->wall_dist)
        return fail(ctx, LUDWIG_EINVAL, "null table pointer");
    if (ctx->peers_attached) return fail(ctx, LUDWIG_ESTATE, "levels cannot be added after ludwig_ipc_attach");
    const int nbg = d->n_blocks;                 // blocks of the whole level (all ranks)
    const size_t ncg = (size_t)nbg * BS3;

    Level* Lp = new Level();
    Level& L = *Lp;
    struct Guard { Level* p; ~Guard() { if (p) free_level(p); } } guard{Lp};
    L.level_id = d->level_id; L.nb_global = nbg; L.dimx = d->dim_x; L.dimy = d->dim_y; L.dimz = d->dim_z;
    L.tau = d->tau; L.dx = d->dx; L.temporal = d->temporal_storage != 0;

    // --- permutation: Morton order of the block coordinates (identical on every rank)
    std::vector<uint64_t> key(nbg);
    for (int i = 0; i < nbg; ++i) {
        int bx = d->map_x[i], by = d->map_y[i], bz = d->map_z[i];
        if (bx < 1 || by < 1 || bz < 1 || bx > d->dim_x || by > d->dim_y || bz > d->dim_z)
            return fail(ctx, LUDWIG_EINVAL, "block coordinate outside block_pointer extents");
        key[i] = morton3((uint32_t)(bx - 1), (uint32_t)(by - 1), (uint32_t)(bz - 1));
    }
    // LUDWIG_PARTITION=rcb (experiment): recursive coordinate bisection of THIS level's blocks into `world` compact boxes of
    // equal cost (ludwig_partition_rcb).  The internal order becomes (owner, Morton key), so every rank still owns one
    // contiguous range and walks its blocks along the Morton curve; nothing else in the library depends on how the ranges
    // were chosen.  On the 339 M-cell bunny the Morton ranges have 2-3x the halo surface of RCB boxes.
    std::vector<int32_t> rcb_owner;
    if (ctx->world > 1 && ctx->partition_mode != 0) {
        rcb_owner.resize(nbg);
        if (ludwig_partition_rcb_axes(d, ctx->world, ctx->partition_mode == 2 ? 6 : 7, rcb_owner.data()) != LUDWIG_OK) return fail(ctx, LUDWIG_EINVAL, "rcb partition failed (fewer blocks than ranks?)");
    }
    L.int2ref.resize(nbg);
    std::iota(L.int2ref.begin(), L.int2ref.end(), 0);
    if (rcb_owner.empty()) std::sort(L.int2ref.begin(), L.int2ref.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
    else std::sort(L.int2ref.begin(), L.int2ref.end(), [&](int32_t a, int32_t b) {
            return rcb_owner[a] != rcb_owner[b] ? rcb_owner[a] < rcb_owner[b] : key[a] < key[b]; });
    L.ref2int.resize(nbg);
    for (int i = 0; i < nbg; ++i) L.ref2int[L.int2ref[i]] = i;

    // --- partition: `world` contiguous ranges of the Morton curve with (approximately) equal COST.  A block's cost
    // follows the kernel class it will run in (measured on Wing_5_deg / bunny: feature blocks ~2x a plain block).
    L.part_starts.assign(ctx->world + 1, 0);
    if (ctx->world == 1) { L.part_starts[1] = nbg; }
    else if (!rcb_owner.empty()) {
        for (int i = 0; i < nbg; ++i) L.part_starts[rcb_owner[i] + 1] += 1;
        for (int r = 0; r < ctx->world; ++r) L.part_starts[r + 1] += L.part_starts[r];
    } else if (ctx->has_plan) {
        // spatially aligned cut (ludwig_partition_plan): a block belongs to the rank whose key interval holds its Morton
        // key scaled to the finest level, so parents, children and neighbours of one region live on one GPU
        if (d->level_id > ctx->plan_levels) return fail(ctx, LUDWIG_EINVAL, "level beyond the partition plan");
        const int sh = 3 * (ctx->plan_levels - d->level_id);
        std::vector<uint64_t> sk(nbg);
        for (int gi = 0; gi < nbg; ++gi) sk[gi] = key[L.int2ref[gi]] << sh;
        for (int r = 1; r < ctx->world; ++r) {
            int cut = (int)(std::lower_bound(sk.begin(), sk.end(), ctx->plan_keys[r]) - sk.begin());
            cut = std::max(cut, L.part_starts[r - 1] + 1);       // every rank keeps at least one block of every level
            cut = std::min(cut, nbg - (ctx->world - r));
            L.part_starts[r] = cut;
        }
        L.part_starts[ctx->world] = nbg;
    } else {
        std::vector<float> cost(nbg);
        ludwig_block_costs(d, cost.data());                    // reference order
        std::vector<double> pre(nbg + 1, 0.0);
        for (int gi = 0; gi < nbg; ++gi) pre[gi + 1] = pre[gi] + cost[L.int2ref[gi]];
        for (int r = 1; r < ctx->world; ++r) {
            const double target = pre[nbg] * r / ctx->world;
            int cut = (int)(std::lower_bound(pre.begin(), pre.end(), target) - pre.begin());
            cut = std::max(cut, L.part_starts[r - 1] + 1);
            cut = std::min(cut, nbg - (ctx->world - r));
            L.part_starts[r] = cut;
        }
        L.part_starts[ctx->world] = nbg;
    }
    // --- order WITHIN each rank's range ("block_order").  In the block-major layout an x-face halo layer costs a full 32-byte
    // sector per 4 useful bytes (8x), a y- or z-face layer 1x, so the x neighbours of a block must still be in L2 when it runs:
    // x-slab order = tiles of T x T blocks in (y, z), Morton order over the tiles, and inside a tile x-slices one after the other
    // (x-neighbour distance T^2 blocks, always an L2 hit; only the tile faces in y / z can miss).  Measured on the 512^3 box:
    // profiles/README.md.  The owner of every block is unchanged (the ranges were cut on the Morton / RCB order above).
    if (ctx->opt_block_order > 0) {
        const uint32_t T = (uint32_t)ctx->opt_block_order;
        std::vector<uint64_t> k2(nbg);
        for (int i = 0; i < nbg; ++i) {
            const uint32_t bx = (uint32_t)(d->map_x[i] - 1), by = (uint32_t)(d->map_y[i] - 1), bz = (uint32_t)(d->map_z[i] - 1);
            const uint64_t tile = morton2(by / T, bz / T);
            k2[i] = ((tile * (uint64_t)(d->dim_x + 1) + bx) * T + (by % T)) * T + (bz % T);
        }
        for (int r = 0; r < ctx->world; ++r)
            std::sort(L.int2ref.begin() + L.part_starts[r], L.int2ref.begin() + L.part_starts[r + 1], [&](int32_t a, int32_t b) { return k2[a] < k2[b]; });
        for (int i = 0; i < nbg; ++i) L.ref2int[L.int2ref[i]] = i;
    }
    L.part_start = L.part_starts[ctx->rank];
    const int nb = L.part_starts[ctx->rank + 1] - L.part_start;
    if (nb <= 0) return fail(ctx, LUDWIG_EINVAL, "level has fewer blocks than ranks");
    if (nb > PTR_LOCAL_MASK) return fail(ctx, LUDWIG_EINVAL, "more than 2^24 blocks per rank on one level");
    L.nb = nb;
    const size_t nc = (size_t)nb * BS3;
    auto owner_of = [&](int gi) { return (int)(std::upper_bound(L.part_starts.begin(), L.part_starts.end(), gi) - L.part_starts.begin()) - 1; };

    // --- topology tables of the local blocks
    std::vector<int32_t> nbr((size_t)nb * 27), bcoord((size_t)nb * 4);
    std::unordered_map<int, int> remote_id;
    for (int bi = 0; bi < nb; ++bi) {
        int br = L.int2ref[L.part_start + bi];
        for (int dir = 0; dir < 27; ++dir) {
            int32_t v = d->neighbor_table[br + (size_t)nbg * dir];
            if (v < 0 || v > nbg) return fail(ctx, LUDWIG_EINVAL, "neighbor_table entry out of range");
            int32_t e = -1;
            if (v > 0) {
                const int gj = L.ref2int[v - 1];
                const int ow = owner_of(gj);
                if (ow == ctx->rank) e = gj - L.part_start;
                else {
                    auto it = remote_id.find(gj);
                    int id;
                    if (it == remote_id.end()) {
                        id = (int)remote_id.size(); remote_id.emplace(gj, id);
                        L.remote_owner.push_back(ow); L.remote_local.push_back(gj - L.part_starts[ow]);
                    } else id = it->second;
                    e = REMOTE_BASE + id;
                }
            }
            nbr[(size_t)bi * 27 + dir] = e;
        }
        bcoord[(size_t)bi * 4 + 0] = d->map_x[br] - 1;
        bcoord[(size_t)bi * 4 + 1] = d->map_y[br] - 1;
        bcoord[(size_t)bi * 4 + 2] = d->map_z[br] - 1;
        bcoord[(size_t)bi * 4 + 3] = 0;
    }
    L.n_remote = (int)remote_id.size();
    // --- packed halo exchange plan.  S(i,e) = the (block R owned by e, direction d) pairs for which some block of rank i has
    // R as its neighbour in direction d, sorted by (R, d).  Every rank derives ALL the sets from the global tables, so the
    // exporter's pack order and the importer's unpack order agree without any communication.  Rank e's export buffer is the
    // concatenation of S(0,e), S(1,e), ... ; an entry takes 768 / 24 / 1 floats (face / edge / corner layer).
    if (ctx->world > 1 && ctx->use_mirror) {
        const int W = ctx->world;
        std::vector<uint8_t> own(nbg);
        for (int r = 0; r < W; ++r) for (int gi = L.part_starts[r]; gi < L.part_starts[r + 1]; ++gi) own[gi] = (uint8_t)r;
        std::vector<std::vector<int64_t>> S((size_t)W * W);
        for (int gi = 0; gi < nbg; ++gi) {
            const int i = own[gi], br = L.int2ref[gi];
            for (int dir = 0; dir < 27; ++dir) {
                const int32_t v = d->neighbor_table[br + (size_t)nbg * dir];
                if (v <= 0 || v > nbg) continue;
                const int gj = L.ref2int[v - 1], e = own[gj];
                if (e != i) S[(size_t)i * W + e].push_back((int64_t)gj * 32 + dir);
            }
        }
        auto esz = [](int dir) { const int nzc = (dir % 3 != 1) + ((dir / 3) % 3 != 1) + (dir / 9 != 1); return nzc == 1 ? 768 : nzc == 2 ? 24 : 1; };
        std::vector<size_t> seg((size_t)W * W, 0);
        for (size_t k = 0; k < S.size(); ++k) {
            std::sort(S[k].begin(), S[k].end());
            S[k].erase(std::unique(S[k].begin(), S[k].end()), S[k].end());
            for (int64_t key : S[k]) seg[k] += (size_t)esz((int)(key & 31));
        }
        for (int e = 0; e < W; ++e) {
            size_t tot = 0;
            for (int i = 0; i < W; ++i) tot += seg[(size_t)i * W + e];
            L.peer_export_floats[e] = tot;
        }
        if (L.peer_export_floats[ctx->rank] >= (1ull << 31)) return fail(ctx, LUDWIG_EINVAL, "halo export buffer exceeds 2^31 floats");
        // what this rank exports: importer-major
        size_t off = 0;
        for (int i = 0; i < W; ++i)
            for (int64_t key : S[(size_t)i * W + ctx->rank]) {
                const int gj = (int)(key >> 5), dir = (int)(key & 31);
                L.h_pack.push_back(make_int4(gj - L.part_start, dir, (int)off, 0));
                off += (size_t)esz(dir);
            }
        L.export_floats = off;
        // what this rank imports: its segment of every exporter's buffer
        for (int e = 0; e < W; ++e) {
            if (e == ctx->rank) continue;
            size_t base = 0;
            for (int i = 0; i < ctx->rank; ++i) base += seg[(size_t)i * W + e];
            for (int64_t key : S[(size_t)ctx->rank * W + e]) {
                const int gj = (int)(key >> 5), dir = (int)(key & 31);
                auto it = remote_id.find(gj);
                if (it == remote_id.end()) return fail(ctx, LUDWIG_EINVAL, "halo plan: remote block not in the neighbour table");
                L.h_unpack.push_back(make_int4(it->second, dir, (int)base, e));
                base += (size_t)esz(dir);
            }
        }
        L.n_pack = (int)L.h_pack.size(); L.n_unpack = (int)L.h_unpack.size();
        if (L.n_pack > 0) {
            CU(dalloc(ctx, &L.d_pack, L.h_pack.size()));
            CU(memcpy_sync(ctx->stream, L.d_pack, L.h_pack.data(), L.h_pack.size() * sizeof(int4), cudaMemcpyHostToDevice));
        }
        if (L.n_unpack > 0) {
            CU(dalloc(ctx, &L.d_unpack, L.h_unpack.size()));
            CU(memcpy_sync(ctx->stream, L.d_unpack, L.h_unpack.data(), L.h_unpack.size() * sizeof(int4), cudaMemcpyHostToDevice));
        }
        // allocated even when empty-ish so that every rank exports the same number of IPC handles
        CU(dalloc(ctx, &L.d_export, std::max<size_t>(2 * L.export_floats, 64)));
        CU(cudaMemsetAsync(L.d_export, 0, std::max<size_t>(2 * L.export_floats, 64) * 4, ctx->stream));
        L.peer_export[ctx->rank] = L.d_export;
    }
    // block pointer of the WHOLE level, rank-encoded: (owner << 24) | owner-local index
    size_t nptr = (size_t)d->dim_x * d->dim_y * d->dim_z;
    std::vector<int32_t> ptr(nptr);
    for (size_t i = 0; i < nptr; ++i) {
        int32_t v = d->block_pointer[i];
        if (v < 0 || v > nbg) return fail(ctx, LUDWIG_EINVAL, "block_pointer entry out of range");
        if (v > 0) {
            const int gj = L.ref2int[v - 1], ow = owner_of(gj);
            ptr[i] = (ow << PTR_RANK_SHIFT) | (gj - L.part_starts[ow]);
        } else ptr[i] = -1;
    }
    // device permutation tables: reference index of every LOCAL block
    std::vector<int32_t> loc2ref(L.int2ref.begin() + L.part_start, L.int2ref.begin() + L.part_start + nb);
    CU(dalloc(ctx, &L.d_int2ref, (size_t)nb));
    CU(dalloc(ctx, &L.d_nbr, (size_t)nb * 27)); CU(dalloc(ctx, &L.d_bcoord, (size_t)nb * 4)); CU(dalloc(ctx, &L.d_ptr, nptr));
    CU(memcpy_sync(ctx->stream, L.d_int2ref, loc2ref.data(), (size_t)nb * 4, cudaMemcpyHostToDevice));
    CU(memcpy_sync(ctx->stream, L.d_nbr, nbr.data(), nbr.size() * 4, cudaMemcpyHostToDevice));
    CU(memcpy_sync(ctx->stream, L.d_bcoord, bcoord.data(), bcoord.size() * 4, cudaMemcpyHostToDevice));
    CU(memcpy_sync(ctx->stream, L.d_ptr, ptr.data(), nptr * 4, cudaMemcpyHostToDevice));

    // --- static fields (the host arrays cover the whole level; each rank keeps its own blocks)
    CU(dalloc(ctx, &L.d_obstacle, nc)); CU(dalloc(ctx, &L.d_sponge, nc)); CU(dalloc(ctx, &L.d_wall_dist, nc));
    {
        uint8_t* stage = nullptr;
        CU(cudaMalloc((void**)&stage, ncg));
        cudaError_t e = memcpy_sync(ctx->stream, stage, d->obstacle, ncg, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) { launch_ref_to_int_u8(stage, L.d_obstacle, L.d_int2ref, nb, ctx->stream); e = cudaStreamSynchronize(ctx->stream); }
        cudaFree(stage);
        if (e != cudaSuccess) return fail(ctx, LUDWIG_ECUDA, std::string("obstacle upload: ") + cudaGetErrorString(e));
    }
    int rc;
    if ((rc = upload_field(ctx, L, d->sponge, L.d_sponge, 1))) return rc;
    if ((rc = upload_field(ctx, L, d->wall_dist, L.d_wall_dist, 1))) return rc;

    // --- state, with the constructor's initial values (blocks.jl:118-147): rho = 1, vel = 0, f = 0
    for (int i = 0; i < 2; ++i) {
        CU(dalloc(ctx, &L.d_f[i], nc * Q)); CU(dalloc(ctx, &L.d_vel[i], nc * 3));
        CU(cudaMemsetAsync(L.d_f[i], 0, nc * Q * 4, ctx->stream));
        CU(cudaMemsetAsync(L.d_vel[i], 0, nc * 3 * 4, ctx->stream));
    }
    CU(dalloc(ctx, &L.d_rho[0], nc));
    launch_fill(L.d_rho[0], 1.0f, nc, ctx->stream);

    // --- Bouzidi: compact the dense FP16 q_map rows of the (local) boundary cells
    L.bouzidi = d->bouzidi_enabled != 0 && d->n_boundary_cells > 0 && d->q_map_f16 != nullptr;
    if (L.bouzidi) {
        if (!d->cell_block || !d->cell_x || !d->cell_y || !d->cell_z) return fail(ctx, LUDWIG_EINVAL, "null boundary-cell list");
        std::vector<int32_t> cells;
        std::vector<uint16_t> q;
        for (int i = 0; i < d->n_boundary_cells; ++i) {
            int br = d->cell_block[i] - 1, x = d->cell_x[i] - 1, y = d->cell_y[i] - 1, z = d->cell_z[i] - 1;
            if (br < 0 || br >= nbg || x < 0 || x > 7 || y < 0 || y > 7 || z < 0 || z > 7)
                return fail(ctx, LUDWIG_EINVAL, "boundary cell out of range");
            const int gi = L.ref2int[br];
            if (gi < L.part_start || gi >= L.part_start + nb) continue;   // another rank's cell
            int loc = x + 8 * y + 64 * z;
            cells.push_back((gi - L.part_start) * BS3 + loc);
            for (int k = 0; k < 27; ++k) q.push_back(d->q_map_f16[(size_t)br * BS3 + loc + ncg * k]);
        }
        L.n_bc = (int)cells.size();
        L.h_bc_cell = std::move(cells); L.h_bc_q = std::move(q);
    }

    // --- per-block feature flags (obstacle / sponge / near-wall); the fast-mode work lists are built lazily
    launch_block_flags(L, ctx->stream);
    CU(cudaStreamSynchronize(ctx->stream));
    CU(memcpy_sync(ctx->stream, bcoord.data(), L.d_bcoord, bcoord.size() * 4, cudaMemcpyDeviceToHost));
    L.h_nbr = std::move(nbr); L.h_bcoord = std::move(bcoord); L.h_ptr = std::move(ptr);

    // --- the previous level now has children: double-buffer its density (implicit rho_old)
    if (!ctx->levels.empty()) {
        Level& P = *ctx->levels.back();
        P.has_children = true;
        if (!P.d_rho[1]) {
            size_t pnc = (size_t)P.nb * BS3;
            CU(dalloc(ctx, &P.d_rho[1], pnc));
            launch_fill(P.d_rho[1], 1.0f, pnc, ctx->stream);
        }
    }
    CU(cudaStreamSynchronize(ctx->stream));
    set_own_peers(ctx, L);
    guard.p = nullptr;
    ctx->levels.push_back(Lp);
    if (!ctx->levels.empty() && ctx->levels.size() >= 2) set_own_peers(ctx, *ctx->levels[ctx->levels.size() - 2]);
    if (out_index) *out_index = (int32_t)ctx->levels.size() - 1;
    return LUDWIG_OK;
}

static int resolve_field(ludwig_ctx* ctx, Level& L, int which, bool for_upload, float** p, int* ncomp) {
    const int in = L.last_t_sub >= 0 ? ((L.last_t_sub % 2 == 0) ? 0 : 1) : 0;
    *p = nullptr;
    switch (which) {
        case LUDWIG_F: *p = L.d_f[0]; *ncomp = Q; break;
        case LUDWIG_F_TEMP: *p = L.d_f[1]; *ncomp = Q; break;
        case LUDWIG_VEL: *p = L.d_vel[0]; *ncomp = 3; break;
        case LUDWIG_VEL_TEMP: *p = L.d_vel[1]; *ncomp = 3; break;
        case LUDWIG_RHO: *p = L.d_rho[L.rho_cur]; *ncomp = 1; break;
        case LUDWIG_F_OLD: case LUDWIG_VEL_OLD: case LUDWIG_RHO_OLD: {
            if (!L.temporal) return fail(ctx, LUDWIG_EINVAL, "level has no temporal storage");
            if (for_upload) { int rc = ensure_explicit_old(ctx, L); if (rc) return rc; }
            *ncomp = which == LUDWIG_F_OLD ? Q : which == LUDWIG_VEL_OLD ? 3 : 1;
            if (L.explicit_old) *p = which == LUDWIG_F_OLD ? L.d_f_old : which == LUDWIG_VEL_OLD ? L.d_vel_old : L.d_rho_old;
            else if (L.last_t_sub < 0) return fail(ctx, LUDWIG_ESTATE, "old state not materialised before the first step");
            else *p = which == LUDWIG_F_OLD ? L.d_f[in] : which == LUDWIG_VEL_OLD ? L.d_vel[in] : (L.d_rho[1] ? L.d_rho[1 - L.rho_cur] : nullptr);
            break;
        }
        case LUDWIG_F_POST:
            return fail(ctx, LUDWIG_EINVAL, "f_post_collision is not materialised by this library (two-phase Bouzidi kernel)");
        default: return fail(ctx, LUDWIG_EINVAL, "unknown field code");
    }
    if (!*p) return fail(ctx, LUDWIG_EINVAL, "field not allocated on this level");
    return LUDWIG_OK;
}

int ludwig_level_upload(ludwig_ctx* ctx, int32_t level, int32_t which, const void* src) {
    if (!level_ok(ctx, level) || !src) return fail(ctx, LUDWIG_EINVAL, "bad level/src");
    CU(cudaSetDevice(ctx->device));
    Level& L = *ctx->levels[level];
    if (which == LUDWIG_OBSTACLE) {
        uint8_t* stage = nullptr;
        size_t nc = (size_t)L.nb_global * BS3;
        CU(cudaMalloc((void**)&stage, nc));
        cudaError_t e = memcpy_sync(ctx->stream, stage, src, nc, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) { launch_ref_to_int_u8(stage, L.d_obstacle, L.d_int2ref, L.nb, ctx->stream); e = cudaStreamSynchronize(ctx->stream); }
        cudaFree(stage);
        if (e != cudaSuccess) return fail(ctx, LUDWIG_ECUDA, cudaGetErrorString(e));
        return LUDWIG_OK;
    }
    float* p; int ncomp;
    int rc = resolve_field(ctx, L, which, true, &p, &ncomp);
    if (rc) return rc;
    return upload_field(ctx, L, (const float*)src, p, ncomp);
}

int ludwig_level_download(ludwig_ctx* ctx, int32_t level, int32_t which, void* dst) {
    if (!level_ok(ctx, level) || !dst) return fail(ctx, LUDWIG_EINVAL, "bad level/dst");
    CU(cudaSetDevice(ctx->device));
    Level& L = *ctx->levels[level];
    if (which == LUDWIG_OBSTACLE) {
        uint8_t* stage = nullptr;
        size_t nc = (size_t)L.nb_global * BS3;
        CU(cudaMalloc((void**)&stage, nc));
        CU(cudaMemsetAsync(stage, 0, nc, ctx->stream));
        launch_int_to_ref_u8(L.d_obstacle, stage, L.d_int2ref, L.nb, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = memcpy_sync(ctx->stream, dst, stage, nc, cudaMemcpyDeviceToHost);
        cudaFree(stage);
        if (e != cudaSuccess) return fail(ctx, LUDWIG_ECUDA, cudaGetErrorString(e));
        return LUDWIG_OK;
    }
    float* p; int ncomp;
    int rc = resolve_field(ctx, L, which, false, &p, &ncomp);
    if (rc) return rc;
    return download_field(ctx, L, p, (float*)dst, ncomp);
}

int ludwig_mesh_create(ludwig_ctx* ctx, int32_t n, const float* cx, const float* cy, const float* cz, const float* nx, const float* ny,
                       const float* nz, const float* area, ludwig_mesh** out) {
    if (!ctx || n <= 0 || !out || !cx || !cy || !cz || !nx || !ny || !nz || !area) return fail(ctx, LUDWIG_EINVAL, "bad mesh args");
    CU(cudaSetDevice(ctx->device));
    auto* m = new ludwig_mesh();
    m->n = n;
    float** dst[7] = {&m->cx, &m->cy, &m->cz, &m->nx, &m->ny, &m->nz, &m->area};
    const float* src[7] = {cx, cy, cz, nx, ny, nz, area};
    for (int i = 0; i < 7; ++i) {
        cudaError_t e = dalloc(ctx, dst[i], (size_t)n);
        if (e == cudaSuccess) e = memcpy_sync(ctx->stream, *dst[i], src[i], (size_t)n * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { ludwig_mesh_destroy(m); return fail(ctx, LUDWIG_ECUDA, cudaGetErrorString(e)); }
    }
    *out = m;
    return LUDWIG_OK;
}
int ludwig_mesh_destroy(ludwig_mesh* m) {
    if (!m) return LUDWIG_OK;
    float* p[7] = {m->cx, m->cy, m->cz, m->nx, m->ny, m->nz, m->area};
    for (float* q : p) if (q) cudaFree(q);
    delete m;
    return LUDWIG_OK;
}

int ludwig_forces_create(ludwig_ctx* ctx, const ludwig_mesh* mesh, double rho_ref, double u_ref, double area_ref, double chord_ref,
                         const double mc[3], int32_t symmetric, ludwig_forces** out) {
    if (!ctx || !mesh || !out || !mc) return fail(ctx, LUDWIG_EINVAL, "bad forces args");
    CU(cudaSetDevice(ctx->device));
    auto* f = new ludwig_forces();
    f->mesh = mesh; f->rho_ref = rho_ref; f->u_ref = u_ref; f->area_ref = area_ref; f->chord_ref = chord_ref;
    f->mc[0] = mc[0]; f->mc[1] = mc[1]; f->mc[2] = mc[2]; f->symmetric = symmetric;
    float** dst[4] = {&f->p, &f->sx, &f->sy, &f->sz};
    for (auto d : dst) {
        cudaError_t e = dalloc(ctx, d, (size_t)mesh->n);
        if (e == cudaSuccess) e = cudaMemsetAsync(*d, 0, (size_t)mesh->n * 4, ctx->stream);
        if (e != cudaSuccess) { ludwig_forces_destroy(f); return fail(ctx, LUDWIG_ECUDA, cudaGetErrorString(e)); }
    }
    if (cudaMalloc((void**)&f->d_acc, 9 * sizeof(double)) != cudaSuccess || cudaMallocHost((void**)&f->h_acc, 9 * sizeof(double)) != cudaSuccess) {
        ludwig_forces_destroy(f);
        return fail(ctx, LUDWIG_ENOMEM, "forces accumulators");
    }
    *out = f;
    return LUDWIG_OK;
}
int ludwig_forces_destroy(ludwig_forces* f) {
    if (!f) return LUDWIG_OK;
    float* p[4] = {f->p, f->sx, f->sy, f->sz};
    for (float* q : p) if (q) cudaFree(q);
    if (f->d_acc) cudaFree(f->d_acc);
    if (f->h_acc) cudaFreeHost(f->h_acc);
    delete f;
    return LUDWIG_OK;
}

// main.jl:109-135
int ludwig_init_equilibrium(ludwig_ctx* ctx) {
    if (!ctx) return LUDWIG_EINVAL;
    CU(cudaSetDevice(ctx->device));
    for (Level* Lp : ctx->levels) {
        Level& L = *Lp;
        launch_init_eq(L.d_f[0], L.d_f[1], L.explicit_old ? L.d_f_old : nullptr, L.nb, ctx->stream);
        if (L.explicit_old) {
            launch_fill(L.d_rho_old, 1.0f, (size_t)L.nb * BS3, ctx->stream);
            CU(cudaMemsetAsync(L.d_vel_old, 0, (size_t)L.nb * BS3 * 3 * 4, ctx->stream));
        }
    }
    CU(cudaGetLastError());
    return LUDWIG_OK;
}

// K0 with a prescribed uniform state instead of rest: f = f_temp = feq(rho = 1, u = (ux, 0, 0)), vel = vel_temp = u (0 in obstacle
// cells), rho = 1 on every level.  Not a reference call site (its init_eq! is the rest state): the initial condition of the
// strong-scaling record of bench.py, where an impulsively started flow gives O(1) surface forces after a few coarse steps.
int ludwig_init_uniform_flow(ludwig_ctx* ctx, float ux) {
    if (!ctx) return LUDWIG_EINVAL;
    CU(cudaSetDevice(ctx->device));
    for (Level* Lp : ctx->levels) {
        Level& L = *Lp;
        launch_init_uniform(L.d_f[0], L.d_f[1], L.d_vel[0], L.d_vel[1], L.d_rho[0], L.d_rho[1], L.d_obstacle, L.nb, ux, ctx->stream);
        ctx->launches += 1;
    }
    CU(cudaGetLastError());
    return LUDWIG_OK;
}

// solver_control.jl:145-165
int ludwig_step_batch(ludwig_ctx* ctx, int64_t t_start, int32_t batch_size, float u_curr, const ludwig_params* params) {
    if (!ctx || !params || ctx->levels.empty() || batch_size < 0) return fail(ctx, LUDWIG_EINVAL, "bad step args");
    if (ctx->group_managed) return fail(ctx, LUDWIG_ESTATE, "this context belongs to a ludwig_multi: step it with ludwig_multi_step_batch");
    CU(cudaSetDevice(ctx->device));
    Group g; g.c.push_back(ctx);
    return group_step_batch(g, t_start, batch_size, u_curr, *params);
}

int ludwig_level_step(ludwig_ctx* ctx, int32_t level, int64_t t_sub, int64_t parent_t_sub, float temporal_weight, float u_curr,
                      const ludwig_params* params) {
    if (!level_ok(ctx, level) || !params) return fail(ctx, LUDWIG_EINVAL, "bad level");
    if (ctx->group_managed) return fail(ctx, LUDWIG_ESTATE, "this context belongs to a ludwig_multi");
    CU(cudaSetDevice(ctx->device));
    Level& L = *ctx->levels[level];
    int rcp = prepare_tables(ctx, *params);
    if (rcp) return rcp;
    if (ctx->world > 1 && ctx->use_mirror) launch_halo_pack(L, (t_sub % 2 == 0) ? 0 : 1, ctx->stream);
    if ((rcp = rank_barrier(ctx))) return rcp;
    Group g; g.c.push_back(ctx);
    if (level == 0) return group_step_level(g, 0, nullptr, t_sub, temporal_weight, u_curr, *params);
    Level& P = *ctx->levels[level - 1];
    if (params->use_temporal && !P.temporal) return fail(ctx, LUDWIG_ESTATE, "parent has no temporal storage");
    std::vector<ParentView> pv{make_parent_view(P, parent_t_sub, /*explicit_old=*/true)};
    return group_step_level(g, (size_t)level, &pv, t_sub, temporal_weight, u_curr, *params);
}

// blocks.jl:199-205.  Only the fine-grained API needs real copies; ludwig_step_batch never calls this.
int ludwig_level_snapshot_old(ludwig_ctx* ctx, int32_t level, int64_t t_sub) {
    if (!level_ok(ctx, level)) return fail(ctx, LUDWIG_EINVAL, "bad level");
    CU(cudaSetDevice(ctx->device));
    Level& L = *ctx->levels[level];
    if (!L.temporal) return LUDWIG_OK;   // length(level.f_old) <= 27: the reference's copy_to_old! is a no-op
    int rc = ensure_explicit_old(ctx, L);
    if (rc) return rc;
    const int in = (t_sub % 2 == 0) ? 0 : 1;
    size_t nc = (size_t)L.nb * BS3;
    CU(cudaMemcpyAsync(L.d_f_old, L.d_f[in], nc * Q * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaMemcpyAsync(L.d_rho_old, L.d_rho[L.rho_cur], nc * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaMemcpyAsync(L.d_vel_old, L.d_vel[in], nc * 3 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    return LUDWIG_OK;
}

// forces/surface.jl:389-601
int ludwig_compute_aerodynamics(ludwig_ctx* ctx, ludwig_forces* F, int32_t level, const double mesh_offset[3], double velocity_scale,
                                double rho_phys, int32_t search_radius, double out[18]) {
    if (!level_ok(ctx, level) || !F || !mesh_offset || !out) return fail(ctx, LUDWIG_EINVAL, "bad aero args");
    CU(cudaSetDevice(ctx->device));
    Level& L = *ctx->levels[level];
    const ludwig_mesh& M = *F->mesh;
    const float pscale = (float)(rho_phys * velocity_scale * velocity_scale);   // forces/surface.jl:402-403
    const float offx = (float)mesh_offset[0], offy = (float)mesh_offset[1], offz = (float)mesh_offset[2];
    // K3 reads level.rho and level.vel (NOT vel_temp) whatever the parity — forces/surface.jl:412
    { int rcb = rank_barrier(ctx); if (rcb) return rcb; }   // K3 reads cells owned by other GPUs
    // multi-GPU: triangles are dealt round-robin to the ranks; each rank returns PARTIAL sums (every output of this
    // call is linear in them), the caller adds the 18 numbers over the ranks.
    PeerBytes obs;
    for (int r = 0; r < MAX_RANKS; ++r) obs.p[r] = L.peer_obstacle[r];
    launch_map_stresses(L, peers_of(L.peer_rho[L.rho_cur]), peers_of(L.peer_vel[0]), obs, M, *F, (float)L.dx, offx, offy, offz, pscale, pscale,
                        search_radius, ctx->rank, ctx->world, ctx->stream);
    launch_integrate_forces(M, *F, offx, offy, offz, ctx->rank, ctx->world, ctx->stream);
    CU(cudaMemcpyAsync(F->h_acc, F->d_acc, 9 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    { int rcb = barrier_state(ctx); if (rcb) return rcb; }
    // FP64 sums of the FP32 per-triangle contributions (the reference's FP32 atomics lose ~1e-4 here)
    double Fx_p = F->h_acc[0], Fy_p = F->h_acc[1], Fz_p = F->h_acc[2];
    double Fx_v = F->h_acc[3], Fy_v = F->h_acc[4], Fz_v = F->h_acc[5];
    double Mx = F->h_acc[6], My = F->h_acc[7], Mz = F->h_acc[8];
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

int ludwig_forces_download_maps(ludwig_ctx* ctx, const ludwig_forces* F, float* p, float* sx, float* sy, float* sz) {
    if (!ctx || !F) return fail(ctx, LUDWIG_EINVAL, "bad forces handle");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    size_t n = (size_t)F->mesh->n * 4;
    if (p) CU(memcpy_sync(ctx->stream, p, F->p, n, cudaMemcpyDeviceToHost));
    if (sx) CU(memcpy_sync(ctx->stream, sx, F->sx, n, cudaMemcpyDeviceToHost));
    if (sy) CU(memcpy_sync(ctx->stream, sy, F->sy, n, cudaMemcpyDeviceToHost));
    if (sz) CU(memcpy_sync(ctx->stream, sz, F->sz, n, cudaMemcpyDeviceToHost));
    return LUDWIG_OK;
}

// diagnostics.jl:56-94
int ludwig_flow_stats(ludwig_ctx* ctx, int32_t level, double out[6]) {
    if (!level_ok(ctx, level) || !out) return fail(ctx, LUDWIG_EINVAL, "bad stats args");
    CU(cudaSetDevice(ctx->device));
    Level& L = *ctx->levels[level];
    int nparts = std::min(4096, std::max(1, std::min(ctx->num_sms * 8, (L.nb * BS3 + 255) / 256)));
    launch_flow_stats(L, L.d_rho[L.rho_cur], L.d_vel[0], ctx->d_stats, nparts, ctx->stream);
    CU(cudaMemcpyAsync(ctx->h_stats, ctx->d_stats, (size_t)nparts * 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    { int rcb = barrier_state(ctx); if (rcb) return rcb; }
    double n = 0, rs = 0, ke = 0, rmin = INFINITY, rmax = -INFINITY, vmax = 0;
    bool nan_rho = false, nan_v = false;   // NaN partials (see flow_stats_kernel): std::min / std::max would drop them
    for (int i = 0; i < nparts; ++i) {
        const double* q = ctx->h_stats + (size_t)i * 6;
        n += q[0]; rs += q[1]; ke += q[5];
        nan_rho |= q[2] != q[2]; nan_v |= q[4] != q[4];
        rmin = std::min(rmin, q[2]); rmax = std::max(rmax, q[3]); vmax = std::max(vmax, q[4]);
    }
    if (nan_rho) rmin = rmax = NAN;
    if (nan_v) vmax = NAN;
    if (n > 0) { out[0] = n; out[1] = rs / n; out[2] = rmin; out[3] = rmax; out[4] = vmax; out[5] = 0.5 * ke; }
    else { out[0] = 0; out[1] = 1; out[2] = 1; out[3] = 1; out[4] = 0; out[5] = 0; }
    return LUDWIG_OK;
}

// ---- multi-GPU (one process per GPU) -------------------------------------------------------------------------

int ludwig_level_upload_local(ludwig_ctx* ctx, int32_t level, int32_t which, const void* src) {
    if (!level_ok(ctx, level) || !src || which == LUDWIG_OBSTACLE) return fail(ctx, LUDWIG_EINVAL, "bad level/src/field");
    CU(cudaSetDevice(ctx->device));
    Level& L = *ctx->levels[level];
    float* p; int ncomp;
    int rc = resolve_field(ctx, L, which, true, &p, &ncomp);
    if (rc) return rc;
    return upload_field(ctx, L, (const float*)src, p, ncomp, true);
}
int ludwig_level_download_local(ludwig_ctx* ctx, int32_t level, int32_t which, void* dst) {
    if (!level_ok(ctx, level) || !dst || which == LUDWIG_OBSTACLE) return fail(ctx, LUDWIG_EINVAL, "bad level/dst/field");
    CU(cudaSetDevice(ctx->device));
    Level& L = *ctx->levels[level];
    float* p; int ncomp;
    int rc = resolve_field(ctx, L, which, false, &p, &ncomp);
    if (rc) return rc;
    return download_field(ctx, L, p, (float*)dst, ncomp, true);
}

// Relative cost of every block of a level (reference order), from the kernel class it will run in:
//   1.0 plain; +1.0 obstacle / sponge / near-wall cells (feature kernel); +1.0 a neighbour block missing (domain face or
//   refinement interface: boundary code or ghost pre-pass); +0.004 per Bouzidi boundary cell; 0.6 if every cell is solid.
int ludwig_block_costs(const ludwig_level_desc* d, float* cost) {
    if (!d || !cost || d->n_blocks <= 0) return LUDWIG_EINVAL;
    const int nb = d->n_blocks;
    #pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; ++b) {
        int n_obs = 0; bool sp = false, wd = false;
        for (int c = 0; c < BS3; ++c) {
            const size_t i = (size_t)b * BS3 + c;
            n_obs += d->obstacle[i] != 0;
            sp |= d->sponge[i] > 0.0f;
            const float w = d->wall_dist[i];
            wd |= (w > 0.0f && w < 10.0f);
        }
        bool miss = false, miss_x_only = true;   // x-only: nothing missing but beyond the inlet / outlet plane (level 1)
        for (int dir = 0; dir < 27; ++dir)
            if (d->neighbor_table[b + (size_t)nb * dir] == 0) { miss = true; if (dir % 3 == 1) miss_x_only = false; }
        const bool feat = n_obs > 0 || sp || wd;
        float c = 1.0f;
        if (n_obs == BS3) c = 0.6f;
        else if (feat) c += 1.0f;
        // (a feature-less inlet / outlet block of level 1 rides in the plain launch with a body as lean as the plain one: merge_face)
        if (miss && !(miss_x_only && !feat && d->level_id == 1)) c += 1.0f;
        cost[b] = c;
    }
    if (d->bouzidi_enabled && d->cell_block)
        for (int i = 0; i < d->n_boundary_cells; ++i) {
            const int b = d->cell_block[i] - 1;
            if (b >= 0 && b < nb) cost[b] += 0.004f;
        }
    return LUDWIG_OK;
}

// Cut keys for a spatially aligned, cost-balanced partition of ALL levels: every block contributes its cost x 2^(l-1)
// sub-steps at its Morton key scaled to the finest level; the key axis is cut into `world` intervals of equal weight.
int ludwig_partition_plan(const ludwig_level_desc* const* descs, int32_t n_levels, int32_t world, uint64_t* keys) {
    if (!descs || !keys || n_levels < 1 || world < 1 || world > MAX_RANKS) return LUDWIG_EINVAL;
    std::vector<std::pair<uint64_t, double>> items;
    for (int l = 0; l < n_levels; ++l) {
        const ludwig_level_desc* d = descs[l];
        if (!d || d->n_blocks <= 0) return LUDWIG_EINVAL;
        std::vector<float> cost(d->n_blocks);
        int rc = ludwig_block_costs(d, cost.data());
        if (rc) return rc;
        const int sh = 3 * (n_levels - 1 - l);
        const double sub = (double)(1u << l);
        for (int b = 0; b < d->n_blocks; ++b)
            items.emplace_back(morton3((uint32_t)(d->map_x[b] - 1), (uint32_t)(d->map_y[b] - 1), (uint32_t)(d->map_z[b] - 1)) << sh, cost[b] * sub);
    }
    std::stable_sort(items.begin(), items.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    double total = 0;
    for (auto& it : items) total += it.second;
    keys[0] = 0; keys[world] = ~0ull;
    double acc = 0; size_t i = 0;
    for (int r = 1; r < world; ++r) {
        const double target = total * r / world;
        while (i < items.size() && acc + items[i].second <= target) { acc += items[i].second; ++i; }
        keys[r] = i < items.size() ? items[i].first : ~0ull;
        if (keys[r] < keys[r - 1]) keys[r] = keys[r - 1];
    }
    return LUDWIG_OK;
}

// Recursive coordinate bisection of one level (host code, no GPU): splits the blocks into `world` boxes of equal cost
// (ludwig_block_costs), always across the longest axis of the current box, at the cost-weighted median; ties along the axis
// are ordered by Morton key, so the result is deterministic and identical on every rank.  owner[b] in reference order.
int ludwig_partition_rcb(const ludwig_level_desc* d, int32_t world, int32_t* owner) { return ludwig_partition_rcb_axes(d, world, 7, owner); }

// axes: bit a set = cuts across axis a allowed (7 = x, y and z; 6 = y and z only).  A cut across x puts x-faces on the cut
// surface: in the block layout an x-face layer is 64 separate 32-byte sectors per direction for 4 useful bytes each, so
// every byte K1 pulls through it over NVLink costs 8; y-faces (8 full sectors) and z-faces (256 contiguous bytes) cost 1.
int ludwig_partition_rcb_axes(const ludwig_level_desc* d, int32_t world, int32_t axes, int32_t* owner) {
    if (!d || !owner || world < 1 || world > MAX_RANKS || d->n_blocks < world || (axes & 7) == 0) return LUDWIG_EINVAL;
    const int nb = d->n_blocks;
    std::vector<float> cost(nb);
    int rc = ludwig_block_costs(d, cost.data());
    if (rc) return rc;
    std::vector<uint64_t> key(nb);
    for (int b = 0; b < nb; ++b) key[b] = morton3((uint32_t)(d->map_x[b] - 1), (uint32_t)(d->map_y[b] - 1), (uint32_t)(d->map_z[b] - 1));
    const int32_t* coord[3] = {d->map_x, d->map_y, d->map_z};
    std::vector<int32_t> idx(nb);
    std::iota(idx.begin(), idx.end(), 0);
    struct Node { int lo, hi, first, n; };   // blocks idx[lo, hi) go to ranks [first, first + n)
    std::vector<Node> stack{{0, nb, 0, world}};
    while (!stack.empty()) {
        const Node nd = stack.back();
        stack.pop_back();
        if (nd.n == 1) { for (int i = nd.lo; i < nd.hi; ++i) owner[idx[i]] = nd.first; continue; }
        int axis = 0, best = -1;
        for (int a = 0; a < 3; ++a) {
            if (!((axes >> a) & 1)) continue;
            int mn = INT32_MAX, mx = INT32_MIN;
            for (int i = nd.lo; i < nd.hi; ++i) { mn = std::min(mn, coord[a][idx[i]]); mx = std::max(mx, coord[a][idx[i]]); }
            if (mx - mn > best) { best = mx - mn; axis = a; }
        }
        const int32_t* ca = coord[axis];
        std::sort(idx.begin() + nd.lo, idx.begin() + nd.hi, [&](int32_t x, int32_t y) { return ca[x] != ca[y] ? ca[x] < ca[y] : key[x] < key[y]; });
        const int n_left = nd.n / 2, n_right = nd.n - n_left;
        double total = 0;
        for (int i = nd.lo; i < nd.hi; ++i) total += cost[idx[i]];
        const double target = total * n_left / nd.n;
        double acc = 0;
        int cut = nd.lo;
        while (cut < nd.hi && acc + 0.5 * cost[idx[cut]] < target) acc += cost[idx[cut++]];   // nearest prefix to the target
        cut = std::max(cut, nd.lo + n_left);            // every rank gets at least one block
        cut = std::min(cut, nd.hi - n_right);
        stack.push_back({nd.lo, cut, nd.first, n_left});
        stack.push_back({cut, nd.hi, nd.first + n_left, n_right});
    }
    return LUDWIG_OK;
}

int ludwig_ctx_set_partition_keys(ludwig_ctx* ctx, const uint64_t* keys, int32_t n_levels) {
    if (!ctx || !keys || n_levels < 1) return fail(ctx, LUDWIG_EINVAL, "bad partition keys");
    if (!ctx->levels.empty()) return fail(ctx, LUDWIG_ESTATE, "set the partition plan before creating levels");
    ctx->plan_keys.assign(keys, keys + ctx->world + 1);
    ctx->plan_levels = n_levels;
    ctx->has_plan = true;
    return LUDWIG_OK;
}

int ludwig_partition_starts(int32_t n_blocks, int32_t world, int32_t* starts) {
    if (n_blocks < 0 || world < 1 || world > MAX_RANKS) return LUDWIG_EINVAL;
    if (starts) for (int r = 0; r <= world; ++r) starts[r] = (int32_t)(((int64_t)n_blocks * r) / world);
    return LUDWIG_OK;
}

int ludwig_ctx_set_partition(ludwig_ctx* ctx, int32_t rank, int32_t world) {
    if (!ctx || world < 1 || world > MAX_RANKS || rank < 0 || rank >= world) return fail(ctx, LUDWIG_EINVAL, "bad rank/world (max 8 ranks)");
    if (!ctx->levels.empty()) return fail(ctx, LUDWIG_ESTATE, "set the partition before creating levels");
    ctx->rank = rank; ctx->world = world;
    if (world > 1 && !ctx->d_bar) {
        CU(cudaSetDevice(ctx->device));
        CU(cudaMalloc((void**)&ctx->d_bar, 2 * 1024 * 1024));        // own allocation granule: IPC exports whole allocations
        CU(cudaMemset(ctx->d_bar, 0, 2 * 1024 * 1024));
        ctx->d_bar_err = (int*)(ctx->d_bar + 64);
        CU(cudaHostAlloc((void**)&ctx->h_bar_err, sizeof(int), cudaHostAllocMapped));
        *ctx->h_bar_err = 0;
        CU(cudaHostGetDevicePointer((void**)&ctx->d_bar_err_dev, ctx->h_bar_err, 0));
    }
    return LUDWIG_OK;
}

int ludwig_set_barrier_callback(ludwig_ctx* ctx, int (*fn)(void*), void* user) {
    if (!ctx) return LUDWIG_EINVAL;
    ctx->barrier_cb = fn; ctx->barrier_user = user;
    return LUDWIG_OK;
}

int ludwig_level_local_blocks(ludwig_ctx* ctx, int32_t level, int32_t* n_local, int32_t* ref_indices) {
    if (!level_ok(ctx, level)) return fail(ctx, LUDWIG_EINVAL, "bad level");
    Level& L = *ctx->levels[level];
    if (n_local) *n_local = L.nb;
    if (ref_indices) for (int i = 0; i < L.nb; ++i) ref_indices[i] = L.int2ref[L.part_start + i] + 1;   // 1-based like every table
    return LUDWIG_OK;
}

// 8 handles per level: f, f_temp, vel, vel_temp, rho[0], rho[1] (zeros if absent), obstacle, halo export buffer.
int ludwig_ipc_export(ludwig_ctx* ctx, void* out, int64_t capacity_bytes, int64_t* needed_bytes) {
    if (!ctx) return LUDWIG_EINVAL;
    CU(cudaSetDevice(ctx->device));
    const int64_t need = ((int64_t)ctx->levels.size() * 8 + 1) * (int64_t)sizeof(cudaIpcMemHandle_t);   // + the barrier slots
    if (needed_bytes) *needed_bytes = need;
    if (!out) return LUDWIG_OK;
    if (capacity_bytes < need) return fail(ctx, LUDWIG_EINVAL, "ipc export buffer too small");
    CU(cudaStreamSynchronize(ctx->stream));
    auto* h = (cudaIpcMemHandle_t*)out;
    std::memset(out, 0, (size_t)need);
    for (size_t l = 0; l < ctx->levels.size(); ++l) {
        Level& L = *ctx->levels[l];
        void* ptrs[8] = {L.d_f[0], L.d_f[1], L.d_vel[0], L.d_vel[1], L.d_rho[0], L.d_rho[1], L.d_obstacle, L.d_export};
        for (int i = 0; i < 8; ++i)
            if (ptrs[i]) CU(cudaIpcGetMemHandle(&h[l * 8 + i], ptrs[i]));
    }
    if (ctx->d_bar) CU(cudaIpcGetMemHandle(&h[ctx->levels.size() * 8], ctx->d_bar));
    return LUDWIG_OK;
}

// all_handles: the export buffers of ranks 0..world-1 concatenated (what an all_gather returns).
int ludwig_ipc_attach(ludwig_ctx* ctx, const void* all_handles, int64_t bytes_per_rank) {
    if (!ctx || !all_handles) return fail(ctx, LUDWIG_EINVAL, "bad attach args");
    CU(cudaSetDevice(ctx->device));
    const int64_t need = ((int64_t)ctx->levels.size() * 8 + 1) * (int64_t)sizeof(cudaIpcMemHandle_t);
    if (bytes_per_rank != need) return fail(ctx, LUDWIG_EINVAL, "ipc attach: handle buffer size mismatch (same levels on every rank?)");
    for (int r = 0; r < ctx->world; ++r) {
        if (r == ctx->rank) continue;
        const auto* h = (const cudaIpcMemHandle_t*)((const char*)all_handles + (size_t)r * need);
        if (ctx->d_bar) {
            void* m = nullptr;
            CU(cudaIpcOpenMemHandle(&m, h[ctx->levels.size() * 8], cudaIpcMemLazyEnablePeerAccess));
            ctx->ipc_opened.push_back(m);
            ctx->peer_bar[r] = (unsigned int*)m;
        }
        for (size_t l = 0; l < ctx->levels.size(); ++l) {
            Level& L = *ctx->levels[l];
            void* own[8] = {L.d_f[0], L.d_f[1], L.d_vel[0], L.d_vel[1], L.d_rho[0], L.d_rho[1], L.d_obstacle, L.d_export};
            void* mapped[8] = {};
            for (int i = 0; i < 8; ++i) {
                if (!own[i]) continue;   // same structure on every rank: absent here = absent there
                CU(cudaIpcOpenMemHandle(&mapped[i], h[l * 8 + i], cudaIpcMemLazyEnablePeerAccess));
                ctx->ipc_opened.push_back(mapped[i]);
            }
            L.peer_f[0][r] = (const float*)mapped[0]; L.peer_f[1][r] = (const float*)mapped[1];
            L.peer_vel[0][r] = (const float*)mapped[2]; L.peer_vel[1][r] = (const float*)mapped[3];
            L.peer_rho[0][r] = (const float*)mapped[4]; L.peer_rho[1][r] = (const float*)mapped[5];
            L.peer_obstacle[r] = (const uint8_t*)mapped[6];
            L.peer_export[r] = (const float*)mapped[7];
        }
    }
    return finish_attach(ctx);
}

// Peers that live in THIS process (one host thread driving several contexts, on one device or on several): the peer
// tables are filled from the other contexts directly, no IPC handles.  peers[r] is the context of rank r (peers[rank] ==
// ctx).  With contexts on different devices, peer access is enabled here.  The native peer-flag barrier works unchanged as
// long as the caller enqueues the same level step on every context before any barrier's ~20 s time-out (tools/emulate_ranks.py
// registers a no-op barrier instead and runs one virtual rank at a time to measure its share of a partitioned case).
int ludwig_attach_inprocess(ludwig_ctx* ctx, ludwig_ctx* const* peers, int32_t n_peers) {
    if (!ctx || !peers || n_peers != ctx->world) return fail(ctx, LUDWIG_EINVAL, "attach_inprocess: one context per rank");
    CU(cudaSetDevice(ctx->device));
    for (int r = 0; r < ctx->world; ++r) {
        ludwig_ctx* o = peers[r];
        if (r == ctx->rank) { if (o != ctx) return fail(ctx, LUDWIG_EINVAL, "attach_inprocess: peers[rank] must be this context"); continue; }
        if (!o || o->world != ctx->world || o->rank != r || o->levels.size() != ctx->levels.size())
            return fail(ctx, LUDWIG_EINVAL, "attach_inprocess: peer context has a different partition / level count");
        if (o->device != ctx->device) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, ctx->device, o->device));
            if (!can) return fail(ctx, LUDWIG_ECUDA, "attach_inprocess: no peer access between the two devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
            cudaGetLastError();
        }
        ctx->peer_bar[r] = o->d_bar;
        for (size_t l = 0; l < ctx->levels.size(); ++l) {
            Level& L = *ctx->levels[l];
            const Level& O = *o->levels[l];
            for (int i = 0; i < 2; ++i) { L.peer_f[i][r] = O.d_f[i]; L.peer_vel[i][r] = O.d_vel[i]; L.peer_rho[i][r] = O.d_rho[i]; }
            L.peer_obstacle[r] = O.d_obstacle;
            L.peer_export[r] = O.d_export;
        }
    }
    return finish_attach(ctx);
}


// ---- several GPUs (or several virtual ranks on one GPU) driven by ONE host thread -------------------------------------------------
//
// "Multi-GPU is hidden behind ludwig_ctx_create(n_gpus, device_ids); level creation takes the global
// tables and the library partitions."  A ludwig_multi owns one context per rank, partitions every level as the per-process
// path does, attaches the ranks in-process (plain device pointers, peer access enabled between devices) and steps them in
// lock-step from the calling thread: every phase of a level step is enqueued on every rank's stream, and the cross-rank
// barriers are stream-ordered event waits (group_barrier) — no spinning kernel, no host blocking, no NCCL.  The kept
// single-process Julia driver (main.jl:54-249) can therefore use N GPUs with the same sequence of calls it makes for one.
// Results are bit-identical to the single-context run in both FP modes.
struct ludwig_multi {
    std::vector<ludwig_ctx*> ctx;
    std::vector<cudaEvent_t> ev;
    bool attached = false;
    struct FH { std::vector<ludwig_mesh*> mesh; std::vector<ludwig_forces*> forces; };
    std::vector<FH> fh;
    std::string err;
};

namespace {
int mfail(ludwig_multi* m, int code, const std::string& msg) { if (m) m->err = msg; return code; }
int mpass(ludwig_multi* m, ludwig_ctx* c, int rc) { if (rc && m) m->err = c ? c->err : "error"; return rc; }
int multi_attach(ludwig_multi* m) {
    if (m->attached) return LUDWIG_OK;
    if (m->ctx.size() > 1)
        for (ludwig_ctx* c : m->ctx) {
            int rc = ludwig_attach_inprocess(c, m->ctx.data(), (int32_t)m->ctx.size());
            if (rc) return mpass(m, c, rc);
        }
    m->attached = true;
    return LUDWIG_OK;
}
}  // namespace

int ludwig_multi_create(ludwig_multi** out, int32_t n_ranks, const int32_t* devices) {
    if (!out || n_ranks < 1 || n_ranks > MAX_RANKS) return LUDWIG_EINVAL;
    *out = nullptr;
    auto* m = new ludwig_multi();
    for (int r = 0; r < n_ranks; ++r) {
        ludwig_ctx* c = nullptr;
        int rc = ludwig_ctx_create(&c, devices ? devices[r] : r);
        if (rc == LUDWIG_OK && n_ranks > 1) rc = ludwig_ctx_set_partition(c, r, n_ranks);
        if (rc != LUDWIG_OK) { if (c) ludwig_ctx_destroy(c); ludwig_multi_destroy(m); return rc; }
        c->group_managed = n_ranks > 1;
        m->ctx.push_back(c);
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { ludwig_multi_destroy(m); return LUDWIG_ECUDA; }
        m->ev.push_back(e);
    }
    *out = m;
    return LUDWIG_OK;
}

int ludwig_multi_destroy(ludwig_multi* m) {
    if (!m) return LUDWIG_OK;
    for (ludwig_ctx* c : m->ctx) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }   // nobody frees memory a peer still reads
    for (auto& f : m->fh) {
        for (ludwig_forces* x : f.forces) ludwig_forces_destroy(x);
        for (ludwig_mesh* x : f.mesh) ludwig_mesh_destroy(x);
    }
    for (size_t r = 0; r < m->ctx.size(); ++r) {
        cudaSetDevice(m->ctx[r]->device);
        if (r < m->ev.size() && m->ev[r]) cudaEventDestroy(m->ev[r]);
        ludwig_ctx_destroy(m->ctx[r]);
    }
    delete m;
    return LUDWIG_OK;
}

const char* ludwig_multi_last_error(const ludwig_multi* m) { return m ? m->err.c_str() : "null handle"; }
int32_t ludwig_multi_num_ranks(const ludwig_multi* m) { return m ? (int32_t)m->ctx.size() : LUDWIG_EINVAL; }
ludwig_ctx* ludwig_multi_ctx(ludwig_multi* m, int32_t rank) { return (m && rank >= 0 && rank < (int)m->ctx.size()) ? m->ctx[rank] : nullptr; }

int ludwig_multi_set_option(ludwig_multi* m, const char* key, const char* value) {
    if (!m) return LUDWIG_EINVAL;
    for (ludwig_ctx* c : m->ctx) { int rc = ludwig_ctx_set_option(c, key, value); if (rc) return mpass(m, c, rc); }
    return LUDWIG_OK;
}

int ludwig_multi_set_partition_plan(ludwig_multi* m, const ludwig_level_desc* const* descs, int32_t n_levels) {
    if (!m || !descs) return LUDWIG_EINVAL;
    if (m->ctx.size() == 1) return LUDWIG_OK;
    std::vector<uint64_t> keys(m->ctx.size() + 1);
    int rc = ludwig_partition_plan(descs, n_levels, (int32_t)m->ctx.size(), keys.data());
    if (rc) return mfail(m, rc, "ludwig_partition_plan failed");
    for (ludwig_ctx* c : m->ctx) if ((rc = ludwig_ctx_set_partition_keys(c, keys.data(), n_levels))) return mpass(m, c, rc);
    return LUDWIG_OK;
}

int ludwig_multi_level_create(ludwig_multi* m, const ludwig_level_desc* desc, int32_t* out_index) {
    if (!m || !desc) return LUDWIG_EINVAL;
    if (m->attached) return mfail(m, LUDWIG_ESTATE, "levels cannot be added after the first step");
    for (ludwig_ctx* c : m->ctx) { int rc = ludwig_level_create(c, desc, out_index); if (rc) return mpass(m, c, rc); }
    return LUDWIG_OK;
}

int ludwig_multi_init_equilibrium(ludwig_multi* m) {
    if (!m) return LUDWIG_EINVAL;
    for (ludwig_ctx* c : m->ctx) { int rc = ludwig_init_equilibrium(c); if (rc) return mpass(m, c, rc); }
    return LUDWIG_OK;
}

int ludwig_multi_init_uniform_flow(ludwig_multi* m, float ux) {
    if (!m) return LUDWIG_EINVAL;
    for (ludwig_ctx* c : m->ctx) { int rc = ludwig_init_uniform_flow(c, ux); if (rc) return mpass(m, c, rc); }
    return LUDWIG_OK;
}

int ludwig_multi_sync(ludwig_multi* m) {
    if (!m) return LUDWIG_EINVAL;
    for (ludwig_ctx* c : m->ctx) { int rc = ludwig_sync(c); if (rc) return mpass(m, c, rc); }
    return LUDWIG_OK;
}

int ludwig_multi_step_batch(ludwig_multi* m, int64_t t_start, int32_t batch_size, float u_curr, const ludwig_params* params) {
    if (!m || !params || m->ctx.empty() || m->ctx[0]->levels.empty() || batch_size < 0) return mfail(m, LUDWIG_EINVAL, "bad step args");
    int rc = multi_attach(m);
    if (rc) return rc;
    Group g; g.c = m->ctx; g.ev = &m->ev;
    rc = group_step_batch(g, t_start, batch_size, u_curr, *params);
    if (rc) for (ludwig_ctx* c : m->ctx) if (!c->err.empty()) { m->err = c->err; break; }
    return rc;
}

// Whole-level upload / download in the reference layout: every rank picks / contributes its own blocks.
int ludwig_multi_level_upload(ludwig_multi* m, int32_t level, int32_t which, const void* src) {
    if (!m) return LUDWIG_EINVAL;
    for (ludwig_ctx* c : m->ctx) { int rc = ludwig_level_upload(c, level, which, src); if (rc) return mpass(m, c, rc); }
    return LUDWIG_OK;
}
int ludwig_multi_level_download(ludwig_multi* m, int32_t level, int32_t which, void* dst) {
    if (!m || !dst) return LUDWIG_EINVAL;
    if (m->ctx.size() == 1) return mpass(m, m->ctx[0], ludwig_level_download(m->ctx[0], level, which, dst));
    if (!level_ok(m->ctx[0], level)) return mfail(m, LUDWIG_EINVAL, "bad level");
    const size_t nbg = (size_t)m->ctx[0]->levels[level]->nb_global;
    const int ncomp = which == LUDWIG_OBSTACLE ? 1 : (which == LUDWIG_RHO || which == LUDWIG_RHO_OLD) ? 1 : (which == LUDWIG_VEL || which == LUDWIG_VEL_TEMP || which == LUDWIG_VEL_OLD) ? 3 : Q;
    if (which == LUDWIG_OBSTACLE) {   // bytes: every rank returns zeros outside its blocks -> OR them together
        std::vector<uint8_t> tmp(nbg * BS3);
        std::memset(dst, 0, nbg * BS3);
        for (ludwig_ctx* c : m->ctx) {
            int rc = ludwig_level_download(c, level, which, tmp.data());
            if (rc) return mpass(m, c, rc);
            uint8_t* d = (uint8_t*)dst;
            for (size_t i = 0; i < tmp.size(); ++i) d[i] |= tmp[i];
        }
        return LUDWIG_OK;
    }
    std::vector<float> tmp;
    for (ludwig_ctx* c : m->ctx) {
        Level& L = *c->levels[level];
        tmp.resize((size_t)L.nb * BS3 * ncomp);
        int rc = ludwig_level_download_local(c, level, which, tmp.data());
        if (rc) return mpass(m, c, rc);
        for (int k = 0; k < ncomp; ++k)
            for (int b = 0; b < L.nb; ++b)
                std::memcpy((float*)dst + ((size_t)k * nbg + (size_t)L.int2ref[L.part_start + b]) * BS3, tmp.data() + ((size_t)k * L.nb + b) * BS3, BS3 * sizeof(float));
    }
    return LUDWIG_OK;
}

int ludwig_multi_flow_stats(ludwig_multi* m, int32_t level, double out[6]) {
    if (!m || !out) return LUDWIG_EINVAL;
    double n = 0, rs = 0, ke = 0, rmin = INFINITY, rmax = -INFINITY, vmax = 0;
    bool nan_seen = false, nan_v = false;
    for (ludwig_ctx* c : m->ctx) {
        double q[6];
        int rc = ludwig_flow_stats(c, level, q);
        if (rc) return mpass(m, c, rc);
        if (q[0] <= 0) continue;
        n += q[0]; rs += q[1] * q[0]; ke += q[5];
        if (q[2] != q[2]) nan_seen = true;
        if (q[4] != q[4]) nan_v = true;
        rmin = std::min(rmin, q[2]); rmax = std::max(rmax, q[3]); vmax = std::max(vmax, q[4]);
    }
    if (n > 0) { out[0] = n; out[1] = rs / n; out[2] = rmin; out[3] = rmax; out[4] = vmax; out[5] = ke; }
    else { out[0] = 0; out[1] = 1; out[2] = 1; out[3] = 1; out[4] = 0; out[5] = 0; }
    if (nan_seen) out[2] = out[3] = NAN;
    if (nan_v) out[4] = NAN;
    return LUDWIG_OK;
}

// Mesh + ForceData on every rank (main.jl:101,145); returns a handle index for ludwig_multi_compute_aerodynamics.
int ludwig_multi_forces_create(ludwig_multi* m, int32_t n_triangles, const float* cx, const float* cy, const float* cz, const float* nx,
                               const float* ny, const float* nz, const float* area, double rho_ref, double u_ref, double area_ref,
                               double chord_ref, const double moment_center[3], int32_t symmetric, int32_t* out_handle) {
    if (!m || !out_handle) return LUDWIG_EINVAL;
    ludwig_multi::FH f;
    for (ludwig_ctx* c : m->ctx) {
        ludwig_mesh* mesh = nullptr; ludwig_forces* forces = nullptr;
        int rc = ludwig_mesh_create(c, n_triangles, cx, cy, cz, nx, ny, nz, area, &mesh);
        if (rc == LUDWIG_OK) { f.mesh.push_back(mesh); rc = ludwig_forces_create(c, mesh, rho_ref, u_ref, area_ref, chord_ref, moment_center, symmetric, &forces); }
        if (rc == LUDWIG_OK) f.forces.push_back(forces);
        if (rc) { for (auto* x : f.forces) ludwig_forces_destroy(x); for (auto* x : f.mesh) ludwig_mesh_destroy(x); return mpass(m, c, rc); }
    }
    m->fh.push_back(f);
    *out_handle = (int32_t)m->fh.size() - 1;
    return LUDWIG_OK;
}

// compute_aerodynamics! over all ranks: every rank integrates the triangles dealt to it; all 18 outputs are linear in the partial sums.
int ludwig_multi_compute_aerodynamics(ludwig_multi* m, int32_t handle, int32_t level, const double mesh_offset[3], double velocity_scale,
                                      double rho_phys, int32_t search_radius, double out[18]) {
    if (!m || !out || handle < 0 || handle >= (int)m->fh.size()) return mfail(m, LUDWIG_EINVAL, "bad forces handle");
    int rc = multi_attach(m);
    if (rc) return rc;
    Group g; g.c = m->ctx; g.ev = &m->ev;
    if (m->ctx.size() > 1 && (rc = group_barrier(g))) return rc;   // K3 reads cells owned by other ranks
    for (int i = 0; i < 18; ++i) out[i] = 0.0;
    for (size_t r = 0; r < m->ctx.size(); ++r) {
        double q[18];
        rc = ludwig_compute_aerodynamics(m->ctx[r], m->fh[handle].forces[r], level, mesh_offset, velocity_scale, rho_phys, search_radius, q);
        if (rc) return mpass(m, m->ctx[r], rc);
        for (int i = 0; i < 18; ++i) out[i] += q[i];
    }
    return LUDWIG_OK;
}

// forces/io.jl:28-31: the per-triangle maps; triangle i was computed by rank i % n_ranks.
int ludwig_multi_forces_download_maps(ludwig_multi* m, int32_t handle, float* p, float* sx, float* sy, float* sz) {
    if (!m || handle < 0 || handle >= (int)m->fh.size()) return mfail(m, LUDWIG_EINVAL, "bad forces handle");
    const int n = m->fh[handle].mesh[0]->n, W = (int)m->ctx.size();
    std::vector<float> t[4];
    float* dst[4] = {p, sx, sy, sz};
    for (auto& v : t) v.resize(n);
    for (int r = 0; r < W; ++r) {
        int rc = ludwig_forces_download_maps(m->ctx[r], m->fh[handle].forces[r], t[0].data(), t[1].data(), t[2].data(), t[3].data());
        if (rc) return mpass(m, m->ctx[r], rc);
        for (int j = 0; j < 4; ++j) if (dst[j]) for (int i = r; i < n; i += W) dst[j][i] = t[j][i];
    }
    return LUDWIG_OK;
}

// io_vtk.jl:17-59 over all ranks: the valid-block list is the same on every rank; every rank writes its own blocks' cells
// into the caller's arrays (disjoint positions)
int ludwig_multi_output_valid_blocks(ludwig_multi* m, int32_t* n_valid, int32_t* blocks) {
    if (!m) return LUDWIG_EINVAL;
    return mpass(m, m->ctx[0], ludwig_output_valid_blocks(m->ctx[0], n_valid, blocks));
}
int ludwig_multi_output_export(ludwig_multi* m, int64_t t_step, float* rho_arr, float* vel_mat, uint8_t* obst_arr, int32_t* level_arr) {
    if (!m) return LUDWIG_EINVAL;
    for (ludwig_ctx* c : m->ctx) { int rc = ludwig_output_export(c, t_step, rho_arr, vel_mat, obst_arr, level_arr); if (rc) return mpass(m, c, rc); }
    return LUDWIG_OK;
}

int64_t ludwig_multi_self_check(ludwig_multi* m) {
    if (!m) return LUDWIG_EINVAL;
    int64_t bad = 0;
    for (ludwig_ctx* c : m->ctx) { const int64_t r = ludwig_ctx_self_check(c); if (r < 0) return mpass(m, c, (int)r); if (r > 0 && bad == 0) m->err = c->err; bad += r; }
    return bad;
}

int64_t ludwig_multi_device_bytes(const ludwig_multi* m) {
    int64_t b = 0;
    if (m) for (ludwig_ctx* c : m->ctx) b += c->bytes;
    return b;
}

}  // extern "C"
