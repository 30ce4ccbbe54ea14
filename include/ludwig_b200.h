/*
 * ludwig_b200.h — C ABI of libludwig_b200.so
 *
 * Drop-in replacement for the KernelAbstractions launch sites of OPEN_Ludwig's
 * per-timestep D3Q27 hot path.  Every entry point cites the reference call site
 * (file:line under /root/reference/src) it replaces.  The reference has no FFI of
 * its own (it is 100 % Julia); the binding a maintainer adds is the `ccall` glue
 * in open_ludwig_b200/julia/LudwigB200.jl.
 *
 * Conventions
 *  - every function returns 0 on success, a negative LUDWIG_E* code on error and
 *    never throws across the ABI; ludwig_last_error() returns the message.
 *  - all index tables are passed EXACTLY as the reference builds them: 1-based,
 *    0 = "none", Julia column-major.  Field arrays use the reference layout
 *    f[x,y,z,b,k]  <=>  linear = x + 8*y + 64*z + 512*b + 512*n_blocks*k  (0-based).
 *    The library converts to its own HBM layout internally.
 *  - host pointers are borrowed for the duration of the call only.
 *  - handles are opaque and owned by the library.
 *  - step / force calls are asynchronous on the context's stream until
 *    ludwig_sync() or a call that returns host scalars.
 *  - one calling thread per context.
 */
#ifndef LUDWIG_B200_H
#define LUDWIG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LUDWIG_OK 0
#define LUDWIG_EINVAL (-1)   /* bad argument */
#define LUDWIG_ECUDA (-2)    /* CUDA runtime error */
#define LUDWIG_ENOMEM (-3)   /* allocation failure */
#define LUDWIG_ESTATE (-4)   /* call order violated */

typedef struct ludwig_ctx ludwig_ctx;
typedef struct ludwig_mesh ludwig_mesh;
typedef struct ludwig_forces ludwig_forces;

/* One refinement level as the reference's BlockLevel (blocks.jl:16-65) holds it on
 * the host after setup_multilevel_domain (domain.jl:224-236). */
typedef struct ludwig_level_desc {
    int32_t level_id;              /* 1-based (blocks.jl:17) */
    int32_t n_blocks;              /* length(active_block_coords) */
    int32_t dim_x, dim_y, dim_z;   /* size(block_pointer) = max active block coord per axis (blocks.jl:104-115) */
    float tau;                     /* blocks.jl:20 */
    double dx;                     /* blocks.jl:18 (a Float64 holding a Float32 value) */
    const int32_t* block_pointer;  /* [dim_x,dim_y,dim_z] col-major, 1-based block index, 0 = none */
    const int32_t* neighbor_table; /* [n_blocks,27] col-major, dir = (dx+1)+(dy+1)*3+(dz+1)*9 (+1 in Julia) */
    const int32_t* map_x;          /* [n_blocks] 1-based block coordinates */
    const int32_t* map_y;
    const int32_t* map_z;
    const uint8_t* obstacle;       /* Bool[8,8,8,n_blocks] */
    const float* sponge;           /* Float32[8,8,8,n_blocks] */
    const float* wall_dist;        /* Float32[8,8,8,n_blocks] (metres; 100 = far) */
    int32_t temporal_storage;      /* has_temporal_storage(level) (blocks.jl:208) */
    int32_t bouzidi_enabled;       /* blocks.jl:156 */
    int32_t n_boundary_cells;      /* blocks.jl:64 */
    const uint16_t* q_map_f16;     /* Float16[8,8,8,n_blocks,27] bit patterns, or NULL if !bouzidi_enabled */
    const int32_t* cell_block;     /* [n_boundary_cells] 1-based */
    const int8_t* cell_x;          /* [n_boundary_cells] 1-based local coords */
    const int8_t* cell_y;
    const int8_t* cell_z;
} ludwig_level_desc;

/* Scalar arguments of perform_timestep_v2! (physics_v2.jl:26-41) that are constant
 * across a batch (main.jl:60-66,176-180). */
typedef struct ludwig_params {
    float c_wale;
    float nu_sgs_bg;
    float inlet_turbulence;
    float q_min_threshold;      /* Q_MIN_THRESHOLD (config_loader.jl:184) */
    int32_t wall_model_active;
    int32_t use_temporal;       /* TEMPORAL_INTERPOLATION */
    int32_t sponge_blend;       /* SPONGE_BLEND_DISTRIBUTIONS */
    int32_t symmetric;          /* SYMMETRIC_ANALYSIS */
    int32_t domain_nx;          /* level-1 cell counts (main.jl:93) */
    int32_t domain_ny;
    int32_t domain_nz;
    int32_t strict_fp;          /* 1: mirror the reference's FP32 operation order without FMA
                                   contraction (parity build); 0: fast path (FMA, regrouped sums) */
} ludwig_params;

/* which-codes for ludwig_level_upload / ludwig_level_download: the BlockLevel field
 * (blocks.jl:28-47) to move, always in the reference layout. */
enum {
    LUDWIG_F = 0,         /* Float32[8,8,8,nb,27] */
    LUDWIG_F_TEMP = 1,
    LUDWIG_F_POST = 2,    /* only if n_boundary_cells > 0 */
    LUDWIG_F_OLD = 3,     /* only if temporal_storage */
    LUDWIG_RHO = 4,       /* Float32[8,8,8,nb] */
    LUDWIG_RHO_OLD = 5,
    LUDWIG_VEL = 6,       /* Float32[8,8,8,nb,3] */
    LUDWIG_VEL_TEMP = 7,
    LUDWIG_VEL_OLD = 8,
    LUDWIG_OBSTACLE = 9   /* uint8[8,8,8,nb] */
};

/* -- context ------------------------------------------------------------------- */

/* main.jl:75 backend selection.  `device` is the CUDA ordinal. */
int ludwig_ctx_create(ludwig_ctx** out, int device);
/* main.jl:247-248 (grids = nothing; force_cleanup()) */
int ludwig_ctx_destroy(ludwig_ctx* ctx);
/* Behaviour switches without a reference counterpart (the reference has one device and one kernel per step).  The library
 * reads NO environment variable; everything that changes what runs is set here, key / value as strings:
 *   prepass = thread | block          interface pre-pass variant (identical bits)
 *   serial_prepass = 0 | 1            pre-pass on the main stream instead of beside the plain K1 launch
 *   single_stream = 0 | 1             no concurrent launches at all
 *   fork_full = 0 | 1                 domain-face K1 launch beside the plain launch on large levels
 *   fork_max_blocks = N               levels up to N blocks run their K1 launch classes concurrently (default: every level)
 *   strict_kernel = reg | stash | tma strict K1 variant: pulled populations in registers (2 CTAs / SM), in a shared-memory stash
 *                                     (3 CTAs / SM), or persistent CTAs with cp.async.bulk (TMA) staged, double-buffered block tiles
 *   cta_threads = auto | 256 | 128 | 64   threads per CTA of the non-persistent K1 kernels: a CTA takes 8 / 4 / 2 z-planes of a block
 *                                     (auto: 64 in strict mode, 128 in fast mode - the measured optimum)
 *   strict_occupancy = 4 | 5 | 6       strict K1 register budget: 16 / 20 / 24 resident warps per SM (128 / 96 / 80 registers per thread)
 *   prefetch_distance = N             K1 CTAs prefetch the lines of the block N list entries ahead into L2 (0 = off)
 *   fast_kernel = direct | tma        fast K1 variant: direct loads, or the persistent TMA-staged form (identical bits either way)
 *   strict_generic = 0 | 1            strict_fp through the one-thread-per-cell cross-check kernel (single GPU)
 *   block_order = morton | xslab<T>   internal block order within a rank (before the first level): Morton curve, or T x T tiles in (y, z)
 *                                     with the x-slices of a tile one after the other (default xslab12: x-face halo sectors stay in L2)
 *   merge_face = 0 | 1                domain-face blocks ride in the plain K1 launch on levels without an interface pre-pass (default 1;
 *                                     strict mode: only the blocks that lack nothing but what lies beyond the inlet / outlet plane)
 *   feature_first = 0 | 1             feature blocks without a ghost neighbour before the plain launch, beside the pre-pass (measured: no effect, off)
 *   face_persist = N                  domain-face K1 class as N persistent CTAs per SM beside the plain launch (0 = off: measured slower)
 *   strict_loop = 1 | 2 | 4           z-plane pairs of a block one 64-thread strict K1 CTA works through (1: measured best)
 *   strict_feature_occupancy = 3 | 4 | 5   register budget of the strict feature / domain-face classes (166 / 128 / 96; 4: measured best)
 *   l2_fetch = 32 | 64 | 128          cudaLimitMaxL2FetchGranularity of the device (a hint; measured: no effect on DRAM bytes)
 *   partition = morton | rcb | rcb_yz multi-GPU block partition (before the first level)
 *   halo_mirror = 0 | 1               packed halo exchange into local mirrors instead of in-kernel NVLink pulls (before the first level)
 *   remote_order = morton | first | last | interleave   place of the blocks that pull from a peer in the plain launch
 *   graphs = auto | 0 | 1             replay coarse steps as CUDA graphs (auto: multi-level cases on one GPU)
 *   barrier_timeout_s = seconds       time-out of the native cross-GPU barrier (default 20)
 *   verbose = 0 | 1                   per-level table sizes on stderr */
int ludwig_ctx_set_option(ludwig_ctx* ctx, const char* key, const char* value);
const char* ludwig_last_error(const ludwig_ctx* ctx);
/* "cuda-sm100a" for the product library, "cpu-oracle" for the test oracle. */
const char* ludwig_backend_name(void);
/* main.jl:164, solver_control.jl:164, physics_v2.jl:85,95  KernelAbstractions.synchronize */
int ludwig_sync(ludwig_ctx* ctx);

/* -- data upload ---------------------------------------------------------------- */

/* main.jl:98  grids = [adapt(backend, g) for g in cpu_grids]   (blocks.jl:67-87).
 * Levels must be added in order 1..L; returns the 0-based level index in *out_index.
 * Allocates f/f_temp/rho/vel/vel_temp (+ old/post storage as the reference's
 * constructor does, blocks.jl:118-147) with the constructor's initial values. */
int ludwig_level_create(ludwig_ctx* ctx, const ludwig_level_desc* desc, int32_t* out_index);
int ludwig_num_levels(const ludwig_ctx* ctx);

/* adapt() also moves the state arrays; these two move a single field of a level.
 * io_vtk.jl:55-57 / forces/io.jl:28-31 `Array(level.rho)` etc. are the download side. */
int ludwig_level_upload(ludwig_ctx* ctx, int32_t level, int32_t which, const void* src);
int ludwig_level_download(ludwig_ctx* ctx, int32_t level, int32_t which, void* dst);

/* io_vtk.jl:52-58,100-107 (export_merged_mesh_sync, steps "2. Gather data" and the field part of "3."): the reference copies
 * the WHOLE rho / vel (or vel_temp) / obstacle arrays of every level to the host and then picks the blocks that are not
 * covered by a finer level.  This call gathers only the listed blocks on the device, in the order given, straight into the
 * arrays the VTK writer fills: rho_arr Float32[512 n], vel_mat Float32[3, 512 n] (component fastest, as Julia's
 * Matrix{Float32}(3, N)), obst_arr UInt8[512 n] (0 / 1); NaN and +-Inf become 0 as in io_vtk.jl:110-111.  Cell order
 * inside a block is x fastest (cidx of :95).  Velocity buffer by the parity of t_step as :56 (even -> vel_temp).
 * blocks = 1-based reference block indices (b_idx of :27); in a multi-GPU context they must be blocks of this rank. */
int ludwig_output_gather(ludwig_ctx* ctx, int32_t level, int64_t t_step, const int32_t* blocks, int32_t n_blocks,
                         float* rho_arr, float* vel_mat, uint8_t* obst_arr);

/* io_vtk.jl:17-46: the blocks export_merged_mesh_sync writes — every block that is NOT fully covered by the next finer level (all 8
 * children active) — computed on the device from the block-pointer tables and cached.  n_valid[level] = their number per level;
 * blocks (or NULL) receives the 1-based b_idx lists of all levels one after the other (level-major, b_idx ascending: the order of
 * the reference's `valid_blocks`). */
int ludwig_output_valid_blocks(ludwig_ctx* ctx, int32_t* n_valid /* [levels] */, int32_t* blocks /* [sum n_valid] or NULL */);
/* io_vtk.jl:52-58 + 100-111 in one call: the fields of ALL valid blocks in the writer's order, N = 512 * sum(n_valid) cells:
 * rho_arr Float32[N], vel_mat Float32[3, N], obst_arr UInt8[N], level_arr Int32[N] (or NULL).  Only the valid blocks leave the device,
 * through pinned, double-buffered staging (the device gathers chunk i + 1 while chunk i is copied out).  In a multi-GPU context
 * every rank fills the cells of its own blocks and leaves the rest of the arrays untouched. */
int ludwig_output_export(ludwig_ctx* ctx, int64_t t_step, float* rho_arr, float* vel_mat, uint8_t* obst_arr, int32_t* level_arr);

/* main.jl:101  Geometry.upload_mesh_to_gpu (geometry.jl:60-84): Float32 SoA. */
int ludwig_mesh_create(ludwig_ctx* ctx, int32_t n_triangles,
                       const float* cx, const float* cy, const float* cz,
                       const float* nx, const float* ny, const float* nz,
                       const float* area, ludwig_mesh** out);
int ludwig_mesh_destroy(ludwig_mesh* mesh);

/* main.jl:145  ForceData(n_triangles, backend; ...)  (forces/structs.jl:94-119) */
int ludwig_forces_create(ludwig_ctx* ctx, const ludwig_mesh* mesh,
                         double rho_ref, double u_ref, double area_ref, double chord_ref,
                         const double moment_center[3], int32_t symmetric, ludwig_forces** out);
int ludwig_forces_destroy(ludwig_forces* forces);

/* -- time stepping ---------------------------------------------------------------- */

/* main.jl:109-135  init_eq! on every level. */
int ludwig_init_equilibrium(ludwig_ctx* ctx);

/* K0 with a prescribed uniform state instead of rest: f = f_temp = feq(1, (ux, 0, 0)), vel = (ux, 0, 0) (0 in obstacle cells),
 * rho = 1 on every level.  No reference counterpart (init_eq! is the rest state); it is the initial condition of bench.py's
 * strong-scaling record: an impulsively started flow gives O(1) surface forces within a few coarse steps. */
int ludwig_init_uniform_flow(ludwig_ctx* ctx, float ux);

/* solver_control.jl:145-165  execute_timestep_batch!: `batch_size` coarse steps
 * t_start .. t_start+batch_size-1 with the whole recursive 2:1 sub-cycling
 * (recursive_step!, solver_control.jl:21-143) scheduled by the library.  Returns
 * without a host sync. */
int ludwig_step_batch(ludwig_ctx* ctx, int64_t t_start, int32_t batch_size, float u_curr,
                      const ludwig_params* params);

/* Fine-grained entry points for a driver that keeps solver_control.jl verbatim.
 * ludwig_level_step = one perform_timestep_v2! (physics_v2.jl:26-97): K1 (+K2) on
 * `level` with buffer parity derived from t_sub exactly as solver_control.jl:35-41.
 * The parent is level-1 (ignored for level 0); its "new" buffers are the ones its
 * own last step (parity parent_t_sub) wrote, its "old" state the pre-step state
 * (solver_control.jl:65-72). */
int ludwig_level_step(ludwig_ctx* ctx, int32_t level, int64_t t_sub, int64_t parent_t_sub,
                      float temporal_weight, float u_curr, const ludwig_params* params);
/* blocks.jl:199-205 copy_to_old!(level, f_in, vel_in), f_in chosen by parity of t_sub. */
int ludwig_level_snapshot_old(ludwig_ctx* ctx, int32_t level, int64_t t_sub);

/* -- diagnostics / forces ---------------------------------------------------------- */

/* main.jl:197,223  compute_aerodynamics! (forces/surface.jl:592-601): K3 + K4 + the
 * host math of integrate_surface_forces! (:467-572).
 * out[18] = Fx,Fy,Fz, Mx,My,Mz, Fx_p,Fy_p,Fz_p, Fx_v,Fy_v,Fz_v, Cd,Cl,Cs, Cmx,Cmy,Cmz */
int ludwig_compute_aerodynamics(ludwig_ctx* ctx, ludwig_forces* forces, int32_t level,
                                const double mesh_offset[3], double velocity_scale,
                                double rho_phys, int32_t search_radius, double out[18]);
/* forces/io.jl:28-31  Array(force_data.pressure_map) etc.  Each dst is Float32[n_triangles]. */
int ludwig_forces_download_maps(ludwig_ctx* ctx, const ludwig_forces* forces,
                                float* p, float* sx, float* sy, float* sz);

/* main.jl:186  compute_flow_stats(grids[1]) (diagnostics.jl:56-94).
 * out[6] = n_fluid, rho_mean, rho_min, rho_max, v_max, kinetic_energy */
int ludwig_flow_stats(ludwig_ctx* ctx, int32_t level, double out[6]);

/* Device-memory footprint of everything the context owns, in bytes
 * (diagnostics_vram.jl:17 prints the reference's). */
int64_t ludwig_device_bytes(const ludwig_ctx* ctx);

/* -- multi-GPU: one process per GPU (no reference counterpart: the reference is single-device, main.jl:75) -------
 *
 * Every rank creates a context on its own GPU, calls ludwig_ctx_set_partition(rank, world) and then creates the SAME
 * levels from the SAME global tables.  The library orders the blocks of each level along a Morton curve and gives
 * rank r the r-th of `world` contiguous ranges of equal estimated cost (ludwig_block_costs); it allocates state only for its own blocks.
 * After the last level: every rank exports CUDA-IPC handles of its state (ludwig_ipc_export), the host all-gathers
 * the buffers (torch.distributed / MPI) and hands the concatenation to ludwig_ipc_attach.  From then on K1 pulls the
 * populations and velocities of neighbour blocks owned by another GPU directly through the peer mapping (NVLink
 * loads inside the stream-collide kernel: no halo packing, no exchange phase), the interface pre-pass and K3 read
 * remote parents / cells the same way.  The only collective the data path needs is a cross-rank barrier after every
 * level step.  By default the library runs it itself: a one-warp kernel on the context's stream stores this rank's epoch
 * into every peer's flag slots over NVLink and spins on its own slots (stream-ordered, no host round trip, no NCCL; a
 * time-out is sticky and fatal, see ludwig_set_barrier_callback).  ludwig_set_barrier_callback replaces it with a caller-supplied barrier
 * (e.g. a stream-ordered NCCL all-reduce).  ludwig_flow_stats and
 * ludwig_compute_aerodynamics return the calling rank's PARTIAL result (triangles dealt round-robin); all 18
 * aerodynamic outputs are linear in the partial sums, so the caller adds them over the ranks.  Both FP modes. */
int ludwig_partition_starts(int32_t n_blocks, int32_t world, int32_t* starts /* [world+1] or NULL */);  /* equal-count rule */
/* Relative cost of every block (reference order) by the kernel class it will run in; ludwig_level_create cuts the
 * Morton curve into `world` ranges of equal cost (host code, callable without a GPU). */
int ludwig_block_costs(const ludwig_level_desc* desc, float* cost /* [n_blocks] */);
/* Recursive coordinate bisection of one level into `world` compact boxes of equal cost (host code, callable without a GPU):
 * owner[b] for every block in reference order.  ludwig_level_create uses it instead of the Morton ranges when the
 * option partition = rcb is set (2-3x less halo surface on the 339 M-cell bunny). */
int ludwig_partition_rcb(const ludwig_level_desc* desc, int32_t world, int32_t* owner /* [n_blocks] */);
/* The same with a mask of the axes a cut may cross (bit 0 x, 1 y, 2 z).  6 = never cut across x: an x-face halo layer is 64
 * separate 32-byte sectors per direction in the block layout, so pulling it over NVLink moves 8 bytes per useful byte. */
int ludwig_partition_rcb_axes(const ludwig_level_desc* desc, int32_t world, int32_t axes, int32_t* owner /* [n_blocks] */);
int ludwig_ctx_set_partition(ludwig_ctx* ctx, int32_t rank, int32_t world);
/* Optional, multi-level cases: a spatially aligned plan.  ludwig_partition_plan (host code) takes the descriptors of ALL
 * levels and returns world+1 cut keys on the Morton axis of the finest level such that every interval carries the same
 * estimated cost (block cost x 2^(level-1) sub-steps); with ludwig_ctx_set_partition_keys every level is cut at the same
 * places in space, so a fine block, its parent cells and its neighbours live on one GPU except at the cut surfaces. */
int ludwig_partition_plan(const ludwig_level_desc* const* descs, int32_t n_levels, int32_t world, uint64_t* keys /* [world+1] */);
int ludwig_ctx_set_partition_keys(ludwig_ctx* ctx, const uint64_t* keys /* [world+1] */, int32_t n_levels);
/* fn returns 0 on success; a non-zero return marks the context failed (every later stepping / result call returns
 * LUDWIG_ESTATE).  A time-out of the native barrier (option barrier_timeout_s) has the same effect. */
int ludwig_set_barrier_callback(ludwig_ctx* ctx, int (*fn)(void*), void* user);
/* Reference (1-based) indices of the blocks this rank owns on `level`, in the library's internal order. */
int ludwig_level_local_blocks(ludwig_ctx* ctx, int32_t level, int32_t* n_local, int32_t* ref_indices /* or NULL */);
/* Upload / download of ONLY this rank's blocks of a state field: Float32[8,8,8,n_local,ncomp] in the order
 * ludwig_level_local_blocks returns (a 512^3-per-GPU box on 8 GPUs has 116 GB of populations: no host holds it all). */
int ludwig_level_upload_local(ludwig_ctx* ctx, int32_t level, int32_t which, const void* src);
int ludwig_level_download_local(ludwig_ctx* ctx, int32_t level, int32_t which, void* dst);
int ludwig_ipc_export(ludwig_ctx* ctx, void* out, int64_t capacity_bytes, int64_t* needed_bytes);
int ludwig_ipc_attach(ludwig_ctx* ctx, const void* all_handles, int64_t bytes_per_rank);
/* The same attach for contexts that live in ONE process (a single host thread driving several GPUs, or several virtual
 * ranks on one GPU for profiling): peers[r] = the context of rank r, peers[rank] == ctx.  No IPC handles involved. */
int ludwig_attach_inprocess(ludwig_ctx* ctx, ludwig_ctx* const* peers, int32_t n_peers);

/* -- multi-GPU: ONE process, one host thread ("Multi-GPU is hidden behind ludwig_ctx_create(n_gpus,
 * device_ids); level creation takes the global tables and the library partitions") ------------------------------------
 *
 * A ludwig_multi owns one context per rank (devices[r] = CUDA ordinal of rank r; ordinals may repeat: several virtual ranks
 * on one GPU, which is how the partitioned path is tested on a one-GPU box), partitions every level like the per-process
 * path, attaches the ranks in-process and steps them in lock-step from the calling thread.  Cross-rank barriers are
 * stream-ordered event waits.  The kept single-process Julia driver (main.jl:54-249) makes the same sequence of calls as
 * for one GPU: create, level_create x L, init_equilibrium, step_batch, flow_stats / compute_aerodynamics, destroy.
 * Results are bit-identical to the single-context run in both FP modes. */
typedef struct ludwig_multi ludwig_multi;
int ludwig_multi_create(ludwig_multi** out, int32_t n_ranks, const int32_t* devices /* [n_ranks] or NULL = 0..n-1 */);
int ludwig_multi_destroy(ludwig_multi* m);
const char* ludwig_multi_last_error(const ludwig_multi* m);
int32_t ludwig_multi_num_ranks(const ludwig_multi* m);
/* The context of one rank, for the per-rank calls (ludwig_level_local_blocks, ludwig_level_upload_local, profiling ...). */
ludwig_ctx* ludwig_multi_ctx(ludwig_multi* m, int32_t rank);
int ludwig_multi_set_option(ludwig_multi* m, const char* key, const char* value);                 /* every rank */
int ludwig_multi_set_partition_plan(ludwig_multi* m, const ludwig_level_desc* const* descs, int32_t n_levels);  /* optional, before the levels */
int ludwig_multi_level_create(ludwig_multi* m, const ludwig_level_desc* desc, int32_t* out_index);               /* main.jl:98 */
int ludwig_multi_level_upload(ludwig_multi* m, int32_t level, int32_t which, const void* src);     /* whole level, reference layout */
int ludwig_multi_level_download(ludwig_multi* m, int32_t level, int32_t which, void* dst);         /* io_vtk.jl:55-57 */
int ludwig_multi_init_equilibrium(ludwig_multi* m);                                                /* main.jl:109-135 */
int ludwig_multi_step_batch(ludwig_multi* m, int64_t t_start, int32_t batch_size, float u_curr, const ludwig_params* params);  /* solver_control.jl:145-165 */
int ludwig_multi_init_uniform_flow(ludwig_multi* m, float ux);
int ludwig_multi_sync(ludwig_multi* m);
int ludwig_multi_flow_stats(ludwig_multi* m, int32_t level, double out[6]);                        /* main.jl:186, reduced over the ranks */
/* main.jl:101,145: mesh + ForceData on every rank; *out_handle identifies them in the two calls below. */
int ludwig_multi_forces_create(ludwig_multi* m, int32_t n_triangles, const float* cx, const float* cy, const float* cz,
                               const float* nx, const float* ny, const float* nz, const float* area,
                               double rho_ref, double u_ref, double area_ref, double chord_ref,
                               const double moment_center[3], int32_t symmetric, int32_t* out_handle);
int ludwig_multi_compute_aerodynamics(ludwig_multi* m, int32_t handle, int32_t level, const double mesh_offset[3],
                                      double velocity_scale, double rho_phys, int32_t search_radius, double out[18]);   /* main.jl:197,223 */
int ludwig_multi_forces_download_maps(ludwig_multi* m, int32_t handle, float* p, float* sx, float* sy, float* sz);
int ludwig_multi_output_valid_blocks(ludwig_multi* m, int32_t* n_valid, int32_t* blocks);                               /* io_vtk.jl:17-46 */
int ludwig_multi_output_export(ludwig_multi* m, int64_t t_step, float* rho_arr, float* vel_mat, uint8_t* obst_arr, int32_t* level_arr);   /* io_vtk.jl:52-111 */
int64_t ludwig_multi_self_check(ludwig_multi* m);                                                   /* ludwig_ctx_self_check on every rank */
int64_t ludwig_multi_device_bytes(const ludwig_multi* m);

/* -- domain build on the device (N2: the step BEFORE the hot path; the kept Julia driver's setup code calls these instead of its
 * threaded CPU loops).  Host pointers in and out, Float64 geometry in the reference's operation order: the tables are bit-identical
 * to the CPU build.  coords = Int32[nb][3] active block coordinates (1-based, the order of active_block_coords); grid_ptr =
 * Int32[dimx][dimy][dimz] (C order) block index of every block coordinate, 1-based, 0 = none; tris = Float64[n_tri][3][3] in STL
 * coordinates, offset = mesh_offset.  Errors: negative LUDWIG_E* return, message from ludwig_domain_last_error(). ------------------- */
const char* ludwig_domain_last_error(void);
/* domain_generation.jl:34-112  build_block_triangle_map + voxelize_blocks!: sets obstacle[b][z][y][x] = 1 for the shell cells */
int ludwig_domain_voxelize(int device, const double* tris, int64_t n_tri, const double offset[3], double dx, const int32_t* coords, int32_t nb,
                           const int32_t* grid_ptr, int32_t dimx, int32_t dimy, int32_t dimz, uint8_t* obstacle /* in/out [nb][512] */);
/* domain_generation.jl:114-203  perform_flood_fill!: every non-obstacle cell that cannot be reached from the non-obstacle cells of
 * the min-bx blocks through 6-connected non-obstacle cells of existing blocks becomes solid; returns the number of filled cells */
int64_t ludwig_domain_flood_fill(int device, const int32_t* coords, int32_t nb, const int32_t* grid_ptr, int32_t dimx, int32_t dimy, int32_t dimz,
                                 uint8_t* obstacle /* in/out [nb][512] */);
/* domain_generation.jl:371-431  compute_wall_distances!: neighbor_table as in ludwig_level_desc; returns the number of near-wall cells */
int64_t ludwig_domain_wall_distance(int device, const int32_t* neighbor_table, int32_t nb, const uint8_t* obstacle, double dx,
                                    float* wall_dist /* in/out [nb][512], 100 = far */);
/* bouzidi_setup.jl:12-54,64-166 + bouzidi_math.jl:9-102  compute_q_map!: sparse result, one row per boundary cell in the order
 * (block, z, y, x): out_cells Int32[n][4] = 1-based (block, x, y, z), out_q Float64[n][27], out_tri Int32[n][27] (1-based triangle of
 * every link, 0 = none).  Returns n; call with NULL outputs (or a too small capacity) to get n first. */
int64_t ludwig_domain_qmap(int device, const double* tris, int64_t n_tri, const double offset[3], double dx, const int32_t* coords, int32_t nb,
                           const int32_t* grid_ptr, int32_t dimx, int32_t dimy, int32_t dimz, int64_t capacity, int32_t* out_cells, double* out_q,
                           int32_t* out_tri);

/* -- instrumentation (no reference counterpart: the reference only has wall-clock prints, main.jl:37-42,189) -- */

/* The CUDA stream (cudaStream_t) every kernel of this context is launched on, so that a host framework can
 * record its own events on it or order other work against it.  NULL for a CPU backend. */
void* ludwig_ctx_stream(ludwig_ctx* ctx);
/* Bounds checks of the library's own (compute-sanitizer is not available on the target pool): every index table the kernels turn into
 * addresses — neighbour tables, remote-block tables against the owners' block counts, peer offsets against offsets recomputed from
 * the mapped base pointers, the rank-encoded block pointer, the K1 work lists, the Bouzidi cells — is range-checked / re-derived on
 * the host.  Returns the number of violations (0 = clean; the first one is described by ludwig_last_error) or a negative error code. */
int64_t ludwig_ctx_self_check(ludwig_ctx* ctx);
/* Number of kernels this library has launched on the context since creation (kernels inside a replayed CUDA graph included). */
int64_t ludwig_launch_count(const ludwig_ctx* ctx);
/* Number of coarse steps executed by replaying a captured CUDA graph (option graphs). */
int64_t ludwig_graph_replays(const ludwig_ctx* ctx);
/* Per-kernel device timing of K1's dominant kernel: when enabled, every launch of the plain-interior K1 kernel
 * is bracketed by CUDA events on the context's stream.  ludwig_profile_read synchronises, returns the summed
 * duration [ms], the number of bracketed launches and the lattice cells they updated, and resets the counters. */
int ludwig_profile_enable(ludwig_ctx* ctx, int32_t on);
int ludwig_profile_read(ludwig_ctx* ctx, double* ms_total, int64_t* launches, int64_t* cells);
/* Device time [ms] per launch class accumulated by the last ludwig_profile_read: 0 K1 plain, 1 K1 plain+ghost,
 * 2 K1 feature, 3 K1 full, 4 interface pre-pass, 5 Bouzidi, 6 cross-rank barriers, 7 whole level steps (classes 1-3 only when
 * launched on the main stream).  ludwig_profile_levels: per level, out[level * 12 + class], with two more
 * classes: 8 halo unpack (multi-GPU import of the peers' layers), 9 halo pack. */
int ludwig_profile_classes(ludwig_ctx* ctx, double out[8]);
int ludwig_profile_levels(ludwig_ctx* ctx, double* out, int32_t capacity /* >= 12 * levels */);

#ifdef __cplusplus
}
#endif
#endif /* LUDWIG_B200_H */
