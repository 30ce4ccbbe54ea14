// ludwig_internal.h — internal data model of libludwig_b200.so (not part of the ABI).
//
// HBM layout (B200-first; differs from the reference's direction-major SoA, blocks.jl:118-150):
//   * blocks are renumbered along a Morton curve of (bx,by,bz) so that the 26 neighbours of a block
//     are close in launch order and their halo sectors are still resident in the 126 MB L2;
//   * populations are BLOCK-major:  f[b][k][z][y][x]  (one block = 27 x 2 KiB = 54 KiB contiguous),
//     velocities vel[b][c][512], density rho[b][512], flags obstacle/sponge/wall_dist [b][512];
//   * neighbour table nbr[b][27] holds 0-based internal indices, -1 = none.
// ref2int / int2ref keep the permutation so every table can be diffed against the reference order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/ludwig_b200.h"

namespace ludwig {

constexpr int BS = 8;
constexpr int BS3 = 512;
constexpr int Q = 27;
constexpr int MAX_RANKS = 8;
// Neighbour-table encoding: -1 none | [0, nb) local block | [nb, REMOTE_BASE) ghost block nb+g | >= REMOTE_BASE remote
// block (owned by another rank; REMOTE_BASE + index into the level's remote-offset tables).
constexpr int32_t REMOTE_BASE = 1 << 30;
// Block-pointer table encoding (coords -> block): -1 none | (owner_rank << 24) | owner_local_index.
constexpr int PTR_RANK_SHIFT = 24;
constexpr int32_t PTR_LOCAL_MASK = (1 << PTR_RANK_SHIFT) - 1;

// Base pointers of one level's state on every rank (own memory for the calling rank, CUDA-IPC peer mappings for the
// others), passed by value to the kernels that may touch blocks owned by another GPU.
struct PeerPtrs {
    const float* p[MAX_RANKS];
};
struct PeerBytes {
    const uint8_t* p[MAX_RANKS];
};

// The two scalars of a K1 launch that change from one coarse step to the next (physics_v2.jl:76 seed = t_sub % 1000000, main.jl:173
// the ramped inlet velocity), kept in device memory when a coarse step is replayed as a CUDA graph: every other kernel argument of
// a coarse step repeats with period 2 (buffer parity).
struct DynScalars {
    long long t_coarse;
    float u_inlet;
};


// physics_v2.jl:99-117: k = (dx+1) + 3(dy+1) + 9(dz+1), dx fastest.
__host__ __device__ constexpr int lat_cx(int k) { return k % 3 - 1; }
__host__ __device__ constexpr int lat_cy(int k) { return (k / 3) % 3 - 1; }
__host__ __device__ constexpr int lat_cz(int k) { return k / 9 - 1; }
__host__ __device__ constexpr int lat_opp(int k) { return 26 - k; }
__host__ __device__ constexpr int lat_mirror_y(int k) { return k + 3 * (-2 * lat_cy(k)); }
__host__ __device__ constexpr int lat_mirror_z(int k) { return k + 9 * (-2 * lat_cz(k)); }
__host__ __device__ constexpr float lat_w(int k) {
    return (lat_cx(k) * lat_cx(k) + lat_cy(k) * lat_cy(k) + lat_cz(k) * lat_cz(k)) == 0   ? 8.0f / 27.0f
           : (lat_cx(k) * lat_cx(k) + lat_cy(k) * lat_cy(k) + lat_cz(k) * lat_cz(k)) == 1 ? 2.0f / 27.0f
           : (lat_cx(k) * lat_cx(k) + lat_cy(k) * lat_cy(k) + lat_cz(k) * lat_cz(k)) == 2 ? 1.0f / 54.0f
                                                                                         : 1.0f / 216.0f;
}

// per-block flag word (bflags[b])
enum : uint32_t {
    BF_INTERIOR = 1u,    // all 26 neighbour blocks exist -> no boundary / interface handling needed
    BF_OBSTACLE = 2u,    // at least one obstacle cell
    BF_SPONGE = 4u,      // at least one cell with sponge > 0
    BF_WALLDIST = 8u,    // at least one cell with 0 < wall_dist < 10
};

struct Level {
    int level_id = 0;   // 1-based
    int nb = 0;
    int dimx = 0, dimy = 0, dimz = 0;   // extents of the reference's block_pointer
    float tau = 0.f;
    double dx = 0.0;
    bool temporal = false;
    bool has_children = false;

    // partition (one process per GPU): blocks are cut into `world` contiguous ranges of the Morton curve
    int nb_global = 0;
    int part_start = 0;                    // global-internal index of local block 0
    std::vector<int32_t> part_starts;      // [world+1]
    std::vector<int32_t> remote_owner, remote_local;   // remote blocks referenced by local neighbour tables
    int n_remote = 0;
    long long* d_roff_f[2] = {nullptr, nullptr};   // per parity (index = parity of the INPUT buffer)
    long long* d_roff_v[2] = {nullptr, nullptr};
    // packed halo exchange (multi-GPU, opt-in with LUDWIG_HALO_MIRROR=1; see k_misc.cu): export buffer + pack list on the owner,
    // unpack list + local mirrors of the remote blocks on the importer; K1 addresses the mirrors
    std::vector<int4> h_pack, h_unpack;
    int4* d_pack = nullptr; int n_pack = 0;       // {local block, direction, offset in the export buffer, 0}
    int4* d_unpack = nullptr; int n_unpack = 0;   // {mirror slot, direction, offset in the exporter's buffer, exporter rank}
    float* d_export = nullptr;                    // [2][export_floats], index = buffer (parity) the layers were taken from
    size_t export_floats = 0;
    size_t peer_export_floats[MAX_RANKS] = {};    // export_floats of every rank (both sides derive the same plan)
    const float* peer_export[MAX_RANKS] = {};     // the peers' export buffers (IPC mappings)
    float* d_fmirror = nullptr;            // [n_remote][27][512]
    float* d_vmirror = nullptr;            // [n_remote][3][512]
    long long* d_moff_f[2] = {nullptr, nullptr};   // mirror block - local f_in (per input parity), K1's remote offsets
    long long* d_moff_v[2] = {nullptr, nullptr};
    int n_plain_int = 0;                   // the first n_plain_int entries of d_list_plain have no remote neighbour
    // peer base pointers (index = rank); own pointers for this rank
    const float* peer_f[2][MAX_RANKS] = {};
    const float* peer_vel[2][MAX_RANKS] = {};
    const float* peer_rho[2][MAX_RANKS] = {};
    const uint8_t* peer_obstacle[MAX_RANKS] = {};

    // permutation (host + device): GLOBAL internal (Morton) index <-> reference index
    std::vector<int32_t> ref2int, int2ref;
    int32_t* d_ref2int = nullptr;
    int32_t* d_int2ref = nullptr;

    // topology (device)
    int32_t* d_nbr = nullptr;      // [nb][27] internal, -1 none
    int32_t* d_bcoord = nullptr;   // [nb][4]  bx,by,bz (0-based), flags
    int32_t* d_ptr = nullptr;      // [dimx*dimy*dimz] col-major like the reference, internal 0-based, -1 none
    // host copies used to build the fast-mode tables lazily (they depend on the domain extents in ludwig_params)
    std::vector<int32_t> h_nbr, h_bcoord, h_ptr;

    // fast-mode tables (built by ensure_fast_tables at the first fast step)
    bool fast_ready = false;
    int fast_dom[3] = {0, 0, 0};
    int32_t* d_nbr_fast = nullptr;       // [nb][27]: real index, nb + ghost id, or -1 (outside the domain)
    int n_ghost = 0;
    int32_t* d_gcoord = nullptr;         // [n_ghost][4]
    float* d_fghost = nullptr;           // [n_ghost][27][512]
    int32_t* d_gstart = nullptr;         // [n_ghost + 1] offsets of every ghost block in the work list
    int32_t* d_gcell = nullptr;          // interface pre-pass work list
    uint32_t* d_gmask = nullptr;
    uint8_t* d_gcells8 = nullptr;
    int n_gcell = 0;
    int32_t* d_list_plain = nullptr;     // all 26 neighbours real, no feature flag
    int32_t* d_list_plain_g = nullptr;   // all 26 neighbours real or ghost, no feature flag
    int32_t* d_list_feat = nullptr;      // all 26 neighbours present, some feature flag
    int32_t* d_list_full = nullptr;      // some neighbour missing (domain face)
    int32_t* d_list_plain_full = nullptr; // plain + full in internal (spatial) order: the merged launch (option merge_face), fast mode
    int32_t* d_list_plain_xface = nullptr; // plain + x-only face blocks (no feature, nothing missing but beyond the inlet / outlet plane): merged launch, strict mode
    int32_t* d_list_full_rest = nullptr;  // full minus the x-only face blocks
    int n_xface = 0;
    int32_t* d_list_nonplain = nullptr;  // plain_g + feat + full (strict mode: the generic strict kernel)
    int n_plain = 0, n_plain_g = 0, n_feat = 0, n_full = 0;
    int n_feat_nog = 0;                    // the first n_feat_nog entries of d_list_feat have no ghost-block neighbour (they need no pre-pass)

    // static fields (device)
    uint8_t* d_obstacle = nullptr;  // [nb][512]
    float* d_sponge = nullptr;
    float* d_wall_dist = nullptr;

    // state (device).  f[0] = reference `f`, f[1] = `f_temp`; same for vel.
    float* d_f[2] = {nullptr, nullptr};
    float* d_vel[2] = {nullptr, nullptr};
    float* d_rho[2] = {nullptr, nullptr};   // rho[rho_cur] is the reference's level.rho; the other one is
    int rho_cur = 0;                        // the pre-step density (implicit rho_old), only if has_children
    // explicit old-state copies, only materialised by ludwig_level_snapshot_old (fine-grained API)
    float* d_f_old = nullptr;
    float* d_vel_old = nullptr;
    float* d_rho_old = nullptr;
    bool explicit_old = false;
    int64_t last_t_sub = -1;   // parity of the most recent step (for implicit-old bookkeeping)

    // Bouzidi (compact)
    bool bouzidi = false;      // the LEVEL has Bouzidi cells (on some rank); n_bc counts the local ones
    int n_bc = 0;
    std::vector<int32_t> h_bc_cell;  // [n_bc] internal cell index  b*512 + z*64 + y*8 + x
    std::vector<uint16_t> h_bc_q;    // [n_bc][27] fp16 q values (the boundary cells' rows of the dense q_map)
    // active links (q in (q_min, 1]), compacted when ludwig_params.q_min_threshold is first seen
    float links_qmin = -1.0f;
    int n_links = 0;
    int32_t* d_link_cell = nullptr;  // [n_links]
    uint8_t* d_link_k = nullptr;     // [n_links] direction k
    float* d_link_q = nullptr;       // [n_links] q as FP32
    float* d_link_tmp = nullptr;     // [n_links] gathered corrections (two-phase K2)
};

}  // namespace ludwig

struct ludwig_mesh {
    int n = 0;
    float *cx = nullptr, *cy = nullptr, *cz = nullptr, *nx = nullptr, *ny = nullptr, *nz = nullptr, *area = nullptr;
};

struct ludwig_forces {
    const ludwig_mesh* mesh = nullptr;
    double rho_ref = 0, u_ref = 0, area_ref = 0, chord_ref = 0, mc[3] = {0, 0, 0};
    int symmetric = 0;
    float *p = nullptr, *sx = nullptr, *sy = nullptr, *sz = nullptr;   // [n_tri]
    double* d_acc = nullptr;    // [9] reduction result
    double* h_acc = nullptr;    // pinned
};

struct ludwig_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};   // concurrent K1 launches of one level step on small levels
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    cudaStream_t pre_stream = nullptr;       // interface pre-pass, concurrent with the plain K1 launch
    cudaEvent_t ev_pre_fork = nullptr, ev_pre = nullptr;
    cudaStream_t halo_stream = nullptr;      // halo import (multi-GPU), concurrent with the K1 launch over interior blocks
    cudaEvent_t ev_halo = nullptr, ev_halo_fork = nullptr;
    // options (ludwig_ctx_set_option)
    bool use_mirror = false;                 // "halo_mirror": packed halo exchange into local mirrors instead of in-kernel NVLink pulls
    int partition_mode = 0;                  // "partition": 0 Morton ranges / aligned plan, 1 per-level RCB, 2 RCB cutting y and z only
    bool fork_full = false;                  // "fork_full": domain-face K1 launch concurrent with the plain launch on large levels
    int fork_max_blocks = 1 << 30;           // "fork_max_blocks": levels up to this size run their K1 launch classes concurrently (measured: no loss on
                                             // one GPU at any size, +4.5 % on two GPUs where the launch tails wait for NVLink pulls)
    bool use_side_streams = true;            // "single_stream" = 1 turns every concurrent launch off
    bool serial_prepass = false;             // "serial_prepass"
    bool opt_block_prepass = false;          // "prepass" = block
    bool opt_strict_generic = false;         // "strict_generic"
    int opt_strict_variant = 0;              // "strict_kernel" = reg | stash | tma
    int opt_fast_variant = 0;                // "fast_kernel" = direct | tma
    int opt_strict_occ = 5;                  // "strict_occupancy" = 4 | 5 | 6 (5: measured best, profiles/README.md)
    bool opt_feature_first = false;          // "feature_first": feature blocks without a ghost neighbour run BEFORE the plain launch, beside the interface pre-pass (measured: no effect, off)
    int opt_merge_face = 1;                  // "merge_face": domain-face blocks ride in the plain K1 launch on levels without an interface pre-pass (strict: the x-only ones)
    int opt_face_persist = 0;                // "face_persist": persistent CTAs per SM of the domain-face K1 class beside the plain launch (0 = off: measured slower, profiles/README.md)
    int opt_strict_feat_occ = 4;             // "strict_feature_occupancy" = 4 | 5 (128 / 96 registers for the feature and domain-face classes)
    int opt_strict_loop = 1;                 // "strict_loop" = 1 | 2 | 4: z-plane pairs of a block one 64-thread CTA works through
    int opt_block_order = 12;                // "block_order": 0 Morton, T > 0: x-slab order with T x T tiles in (y, z) (12: measured, profiles/README.md)
    int opt_prefetch_distance = 0;           // "prefetch_distance" = blocks
    int opt_cta_threads = 0;                 // "cta_threads" = auto (64 strict / 128 fast) | 256 | 128 | 64
    bool verbose = false;                    // "verbose"
    std::string remote_order = "morton";     // "remote_order"
    double barrier_timeout_s = 20.0;         // "barrier_timeout_s"
    std::vector<ludwig::Level*> levels;
    std::string err;
    int64_t bytes = 0;
    unsigned long long* d_ticket = nullptr;   // [4] ticket counters of the persistent K1 variants, one per launch class
    unsigned long long ticket_base[4] = {0, 0, 0, 0};
    double* d_stats = nullptr;   // flow-stats partials
    float* d_stage = nullptr;    // one component of a whole level in reference order: staging of ludwig_level_upload / download, kept
    size_t stage_floats = 0;     // between calls (a case uploads ~30 components per level) and released when stepping starts
    double* h_stats = nullptr;   // pinned
    int num_sms = 148;
    int rank = 0, world = 1;
    bool has_plan = false;                  // spatially aligned partition (ludwig_ctx_set_partition_keys)
    std::vector<uint64_t> plan_keys;        // [world+1] cut keys at finest-level resolution
    int plan_levels = 0;
    bool peers_attached = false;
    bool group_managed = false;             // one of the contexts of a ludwig_multi: barriers are placed by the group driver
    size_t p7 = (size_t)-1;                 // open "whole level step" profiling bracket
    int (*barrier_cb)(void*) = nullptr;    // cross-rank barrier, stream-ordered or blocking (multi-GPU only); non-zero = failed
    void* barrier_user = nullptr;
    std::vector<void*> ipc_opened;
    // native peer-flag barrier (used when no callback is registered)
    unsigned int* d_bar = nullptr;            // [MAX_RANKS] epoch slots, written by the peers
    unsigned int* peer_bar[ludwig::MAX_RANKS] = {};   // the peers' slot arrays (IPC mappings)
    int* d_bar_err = nullptr;                 // the same flag in device memory (checked by the kernel before it spins)
    int* h_bar_err = nullptr;                 // time-out flag in mapped pinned host memory (the barrier kernel writes, the host reads without sync)
    int* d_bar_err_dev = nullptr;             // its device address
    bool bar_failed = false;                  // sticky: once a barrier failed every stepping / result call returns LUDWIG_ESTATE
    unsigned int bar_epoch = 0;
    int64_t launches = 0;
    float wm_c166 = 0.f;            // see K1Args.wm_c166
    // CUDA-graph replay of coarse steps (abi.cu graph_coarse_step): one instantiated graph per buffer-parity pattern
    struct GraphEntry { cudaGraphExec_t exec = nullptr; int64_t launches = 0; };
    std::map<uint64_t, GraphEntry> graphs;
    ludwig::DynScalars* d_dyn = nullptr;
    ludwig_params graph_params{};
    int opt_graphs = -1;            // "graphs": -1 auto, 0 off, 1 on
    bool graph_failed = false;
    int64_t graph_replays = 0;
    void* output_state = nullptr;   // output.cu: cached valid-block lists + pinned double-buffered staging (N3)
    // K1 profiling (ludwig_profile_enable)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;   // pairs (start, stop)
    std::vector<int> ev_class;          // launch class of each pair
    double prof_class_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<double> prof_level_ms;  // [level][8]
    int prof_level = 0;                 // level index the next bracket is tagged with
    size_t ev_used = 0;
    int64_t prof_cells = 0;
};

void ludwig_output_state_free(ludwig_ctx* ctx);   // output.cu

namespace ludwig {

// Arguments of one K1 launch (physics_kernels.jl:9-38 minus the lattice arrays, which are constexpr here).
struct K1Args {
    const float* f_in; float* f_out;
    const float* vel_in; float* vel_out;
    float* rho_out;
    const uint8_t* obstacle; const float* sponge; const float* wall_dist;
    const int32_t* nbr; const int32_t* bcoord;
    const int32_t* list;   // internal block indices to process (nullptr = identity)
    int n_list;
    int nb;                 // number of local real blocks: see the neighbour-table encoding above
    long long ghost_delta;  // (f_ghost - f_in) in elements
    const long long* roff_f;   // [n_remote] (peer f_in block base - local f_in) in elements, for this step's parity
    const long long* roff_v;   // [n_remote] same for vel_in
    // parent (physics_v2.jl:43-53)
    const float *pf_new, *pf_old, *prho_new, *prho_old, *pvel_new, *pvel_old;
    const int32_t* pptr; int pdimx, pdimy, pdimz;
    float tau, tau_parent, c_wale, nu_bg, u_inlet, inlet_turb, tw;
    int is_l1, is_symmetric, nxg, nyg, nzg, wm, seed, use_temporal, sponge_blend;
    int strict_stash;       // strict build variant: 0 populations in registers (2 CTAs / SM), 1 shared-memory stash (3 CTAs / SM), 2 persistent TMA-staged
    int fast_variant;       // fast build variant: 0 direct loads, 2 persistent TMA-staged
    int num_sms;
    int strict_occ;         // strict K1 at 64 threads per CTA: resident warps per SM / 4 (4, 5 or 6)
    int persist_grid;       // > 0: this launch runs as that many persistent CTAs striding over the list's parts (see abi.cu, "face_persist")
    int strict_feat_occ;    // strict feature / domain-face K1 at 64 threads per CTA: resident warps per SM / 4 (4 or 5)
    int strict_loop;        // strict K1 at 64 threads per CTA: z-plane pairs of a block per CTA (1, 2 or 4)
    int prefetch_distance;  // blocks ahead whose lines a CTA prefetches into L2 (0 = off)
    int cta_threads;        // 256 (one CTA per block), 128 or 64 (a CTA takes 4 / 2 z-planes of a block)
    // persistent (TMA) variants: blocks are handed out in list order through an atomic ticket counter, so that the CTAs in flight always
    // work on a compact window of the Morton curve (halo sectors stay L2 hits); the counter is never reset: block = ticket - ticket_base
    unsigned long long* ticket; unsigned long long ticket_base;
    // graph replay: seed = ((dyn->t_coarse << dyn_shift) + dyn_add) % 1000000 and u_inlet = dyn->u_inlet replace the immediates above
    const DynScalars* dyn; int dyn_shift, dyn_add;
    float wm_c166;          // (2 * 8.3)^(-1/7) of the wall model, evaluated once per context on the device (k1_strict.cu)
    float negzero;          // -0.0f, opaque to ptxas: the strict build's packed multiply is FFMA2(a, b, negzero) (k1_strict.cu)
};

// k1_generic_strict.cu (compiled with -fmad=false): one thread per cell, every branch of the reference inside the kernel
// (in-kernel interface interpolation).  Cross-check only (option "strict_generic"); not on the default path.
void launch_k1_generic_strict(const K1Args& a, cudaStream_t s);
// Arguments of the interface-halo pre-pass (k1_fast.cu ghost_interp_kernel)
struct GhostArgs {
    const int32_t* gcell;    // [n] ghost group id  g*64 + qz*16 + qy*4 + qx  (a group = the 2x2x2 fine cells of one parent cell)
    const uint32_t* gmask;   // [n] bit k set: some real cell pulls population k from a cell of this group
    const uint8_t* gcells8;  // [n] bit (dz*4+dy*2+dx) set: that cell of the group is pulled from
    int n;
    const int32_t* gstart;   // [n_ghost + 1] first work-list entry of every ghost block (block-cooperative variant)
    int n_ghost;
    const int32_t* gcoord;   // [n_ghost][4] ghost block coords (0-based)
    float* f_ghost;          // [n_ghost][27][512]
    PeerPtrs pf_new, pf_old, prho_new, prho_old, pvel_new, pvel_old;   // parent state per owning rank
    const int32_t* pptr; int pdimx, pdimy, pdimz;                      // parent block pointer (rank-encoded)
    float tau, tau_parent, tw;
    int use_temporal;
};

// k1_fast.cu: fast mode.  plain = all 26 neighbours real, no obstacle/sponge/near-wall cell; plain_ghost = same but
// some neighbours are ghost blocks; full = every other block
void launch_k1_plain(const K1Args& a, cudaStream_t s);
void launch_k1_plain_ghost(const K1Args& a, cudaStream_t s);
void launch_k1_feat(const K1Args& a, cudaStream_t s);   // features (obstacle/sponge/wall model), all 26 neighbours present
void launch_k1_full(const K1Args& a, cudaStream_t s);   // features + missing neighbours (domain faces)
void launch_ghost_interp(const GhostArgs& g, bool block_variant, cudaStream_t s);
// k1_strict.cu (-fmad=false): the same four classes and the pre-pass in the reference's operation order (bit-exact
// against the CPU oracle), packed FP32x2
void launch_k1s_plain(const K1Args& a, cudaStream_t s);
void launch_k1s_mixed(const K1Args& a, cudaStream_t s);   // plain + domain-face blocks in one grid (option merge_face)
void launch_k1_mixed(const K1Args& a, cudaStream_t s);
void launch_k1s_plain_ghost(const K1Args& a, cudaStream_t s);
void launch_k1s_feat(const K1Args& a, cudaStream_t s);
void launch_k1s_full(const K1Args& a, cudaStream_t s);
void launch_ghost_interp_strict(const GhostArgs& g, cudaStream_t s);
void launch_wall_model_constant(float* d_out, cudaStream_t s);

// k_misc.cu
void launch_init_eq(float* f0, float* f1, float* f_old, int nb, cudaStream_t s);
void launch_fill(float* p, float v, size_t n, cudaStream_t s);
void launch_init_uniform(float* f0, float* f1, float* v0, float* v1, float* r0, float* r1, const uint8_t* obstacle, int nb, float ux, cudaStream_t s);
void launch_bouzidi(const Level& L, float* f_out, const long long* roff_f_out, bool strict, int phase, cudaStream_t s);
void launch_ref_to_int(const float* src_ref_k, float* dst, const int32_t* int2ref, int nb, int ncomp, int k, cudaStream_t s);
void launch_int_to_ref(const float* src, float* dst_ref_k, const int32_t* int2ref, int nb, int ncomp, int k, cudaStream_t s);
void launch_ref_to_int_u8(const uint8_t* src_ref, uint8_t* dst, const int32_t* int2ref, int nb, cudaStream_t s);
void launch_int_to_ref_u8(const uint8_t* src, uint8_t* dst_ref, const int32_t* int2ref, int nb, cudaStream_t s);
void launch_block_flags(Level& L, cudaStream_t s);
void launch_output_gather(const int32_t* sel, int n, const float* rho, const float* vel, const uint8_t* obs, float* o_rho, float* o_vel,
                          uint8_t* o_obs, cudaStream_t s);
void launch_halo_pack(const Level& L, int buf, cudaStream_t s);
void launch_halo_unpack(const Level& L, int buf, cudaStream_t s);
void launch_peer_barrier(unsigned int* const* peer_slots, unsigned int* own, int rank, int world, unsigned int epoch, int* err, int* err_host, long long timeout_ns, cudaStream_t s);
void launch_map_stresses(const Level& L, const PeerPtrs& rho, const PeerPtrs& vel, const PeerBytes& obstacle, const ludwig_mesh& M,
                         ludwig_forces& F, float dx, float offx, float offy, float offz, float pscale, float sscale, int radius,
                         int tri_first, int tri_stride, cudaStream_t s);
void launch_integrate_forces(const ludwig_mesh& M, ludwig_forces& F, float offx, float offy, float offz, int tri_first, int tri_stride, cudaStream_t s);
void launch_flow_stats(const Level& L, const float* rho, const float* vel, double* d_partials, int nparts, cudaStream_t s);

}  // namespace ludwig
