# LudwigB200.jl — ccall glue for libludwig_b200.so (include/ludwig_b200.h).
#
# Drop this file next to the reference's src/*.jl, `include("LudwigB200.jl")` in main.jl after blocks.jl, and
# replace the five kernel-boundary call sites as shown in INTEGRATION.md.  Mirrors open_ludwig_b200/cabi.py 1:1
# (that Python twin is what the test-suite exercises: Julia is not installed in the build image).
#
# Julia arrays are column-major and 1-based — exactly the layout the C ABI expects, so BlockLevel fields are
# passed as they are (no copies, no index shifts).  Host arrays are only borrowed during the call: GC.@preserve.
module LudwigB200

const LIB = get(ENV, "LUDWIG_B200_LIB", "libludwig_b200")

struct LevelDesc
    level_id::Int32; n_blocks::Int32
    dim_x::Int32; dim_y::Int32; dim_z::Int32
    tau::Float32; dx::Float64
    block_pointer::Ptr{Int32}; neighbor_table::Ptr{Int32}
    map_x::Ptr{Int32}; map_y::Ptr{Int32}; map_z::Ptr{Int32}
    obstacle::Ptr{UInt8}; sponge::Ptr{Float32}; wall_dist::Ptr{Float32}
    temporal_storage::Int32; bouzidi_enabled::Int32; n_boundary_cells::Int32
    q_map_f16::Ptr{UInt16}; cell_block::Ptr{Int32}
    cell_x::Ptr{Int8}; cell_y::Ptr{Int8}; cell_z::Ptr{Int8}
end

struct Params
    c_wale::Float32; nu_sgs_bg::Float32; inlet_turbulence::Float32; q_min_threshold::Float32
    wall_model_active::Int32; use_temporal::Int32; sponge_blend::Int32; symmetric::Int32
    domain_nx::Int32; domain_ny::Int32; domain_nz::Int32
    strict_fp::Int32
end

mutable struct Context
    h::Ptr{Cvoid}
end

function check(ctx::Context, rc::Cint, what)
    rc == 0 && return
    msg = unsafe_string(ccall((:ludwig_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.h))
    error("$what failed ($rc): $msg")          # caught by the per-case try/catch of main.jl:261-267
end

function Context(device::Integer=0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:ludwig_ctx_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), h, device)
    rc == 0 || error("ludwig_ctx_create failed ($rc)")
    ctx = Context(h[])
    finalizer(c -> ccall((:ludwig_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), ctx)
    return ctx
end

# main.jl:98  grids = [adapt(backend, g) for g in cpu_grids]
function add_level!(ctx::Context, g)   # g::BlockLevel (blocks.jl:16-65), still on the host
    obstacle = convert(Array{UInt8,4}, g.obstacle)
    q16 = g.bouzidi_enabled ? reinterpret(UInt16, g.bouzidi_q_map) : UInt16[]
    idx = Ref{Int32}(-1)
    GC.@preserve g obstacle q16 begin
        d = LevelDesc(g.level_id, length(g.active_block_coords),
                      size(g.block_pointer, 1), size(g.block_pointer, 2), size(g.block_pointer, 3),
                      g.tau, g.dx,
                      pointer(g.block_pointer), pointer(g.neighbor_table),
                      pointer(g.map_x), pointer(g.map_y), pointer(g.map_z),
                      pointer(obstacle), pointer(g.sponge), pointer(g.wall_dist),
                      length(g.f_old) > 27 ? 1 : 0, g.bouzidi_enabled ? 1 : 0, g.n_boundary_cells,
                      g.bouzidi_enabled ? pointer(q16) : Ptr{UInt16}(C_NULL),
                      g.bouzidi_enabled ? pointer(g.bouzidi_cell_block) : Ptr{Int32}(C_NULL),
                      g.bouzidi_enabled ? pointer(g.bouzidi_cell_x) : Ptr{Int8}(C_NULL),
                      g.bouzidi_enabled ? pointer(g.bouzidi_cell_y) : Ptr{Int8}(C_NULL),
                      g.bouzidi_enabled ? pointer(g.bouzidi_cell_z) : Ptr{Int8}(C_NULL))
        check(ctx, ccall((:ludwig_level_create, LIB), Cint, (Ptr{Cvoid}, Ref{LevelDesc}, Ref{Int32}), ctx.h, d, idx), "ludwig_level_create")
    end
    return idx[]
end

# main.jl:101 + main.jl:145
function create_mesh(ctx::Context, mesh)   # Geometry.SolverMesh
    f32(v, i) = Float32[c[i] for c in v]
    cx, cy, cz = f32(mesh.centers, 1), f32(mesh.centers, 2), f32(mesh.centers, 3)
    nx, ny, nz = f32(mesh.normals, 1), f32(mesh.normals, 2), f32(mesh.normals, 3)
    ar = Float32.(mesh.areas)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:ludwig_mesh_create, LIB), Cint,
                     (Ptr{Cvoid}, Int32, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ref{Ptr{Cvoid}}),
                     ctx.h, length(ar), cx, cy, cz, nx, ny, nz, ar, h), "ludwig_mesh_create")
    return h[]
end

function create_forces(ctx::Context, mesh_h, params, symmetric::Bool)
    mc = Float64[params.moment_center...]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:ludwig_forces_create, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Float64, Float64, Ptr{Float64}, Int32, Ref{Ptr{Cvoid}}),
                     ctx.h, mesh_h, params.rho_physical, params.u_physical, params.reference_area, params.reference_chord, mc, symmetric ? 1 : 0, h),
          "ludwig_forces_create")
    return h[]
end

# main.jl:126-135
init_equilibrium!(ctx::Context) = check(ctx, ccall((:ludwig_init_equilibrium, LIB), Cint, (Ptr{Cvoid},), ctx.h), "ludwig_init_equilibrium")

# solver_control.jl:145  execute_timestep_batch!
step_batch!(ctx::Context, t_start::Integer, batch::Integer, u_curr::Float32, p::Params) =
    check(ctx, ccall((:ludwig_step_batch, LIB), Cint, (Ptr{Cvoid}, Int64, Int32, Float32, Ref{Params}), ctx.h, t_start, batch, u_curr, p), "ludwig_step_batch")

# physics_v2.jl:26  perform_timestep_v2!  /  blocks.jl:199  copy_to_old!   (fine-grained: keeps solver_control.jl verbatim)
level_step!(ctx::Context, level::Integer, t_sub::Integer, parent_t_sub::Integer, tw::Float32, u_curr::Float32, p::Params) =
    check(ctx, ccall((:ludwig_level_step, LIB), Cint, (Ptr{Cvoid}, Int32, Int64, Int64, Float32, Float32, Ref{Params}),
                     ctx.h, level, t_sub, parent_t_sub, tw, u_curr, p), "ludwig_level_step")
snapshot_old!(ctx::Context, level::Integer, t_sub::Integer) =
    check(ctx, ccall((:ludwig_level_snapshot_old, LIB), Cint, (Ptr{Cvoid}, Int32, Int64), ctx.h, level, t_sub), "ludwig_level_snapshot_old")

# main.jl:197  compute_aerodynamics!  -> fills the reference's ForceData fields
function compute_aerodynamics!(ctx::Context, forces_h, force_data, finest_level::Integer, params; search_radius::Int=5)
    off = Float64[params.mesh_offset...]
    out = zeros(Float64, 18)
    check(ctx, ccall((:ludwig_compute_aerodynamics, LIB), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Float64}, Float64, Float64, Int32, Ptr{Float64}),
                     ctx.h, forces_h, finest_level, off, params.velocity_scale, params.rho_physical, search_radius, out), "ludwig_compute_aerodynamics")
    fd = force_data
    fd.Fx, fd.Fy, fd.Fz, fd.Mx, fd.My, fd.Mz = out[1:6]
    fd.Fx_pressure, fd.Fy_pressure, fd.Fz_pressure, fd.Fx_viscous, fd.Fy_viscous, fd.Fz_viscous = out[7:12]
    fd.Cd, fd.Cl, fd.Cs, fd.Cmx, fd.Cmy, fd.Cmz = out[13:18]
    return fd
end

# main.jl:186  compute_flow_stats(grids[1])
function flow_stats(ctx::Context, level::Integer=0)
    out = zeros(Float64, 6)
    check(ctx, ccall((:ludwig_flow_stats, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), ctx.h, level, out), "ludwig_flow_stats")
    return (n_fluid=Int(out[1]), rho_mean=out[2], rho_min=out[3], rho_max=out[4], v_max=out[5], kinetic_energy=out[6])
end

# io_vtk.jl:55-57  Array(level.rho) etc.   which: 4 = rho, 6 = vel, 7 = vel_temp, 9 = obstacle (include/ludwig_b200.h)
function download!(ctx::Context, level::Integer, which::Integer, dst::Array)
    GC.@preserve dst check(ctx, ccall((:ludwig_level_download, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}), ctx.h, level, which, pointer(dst)), "ludwig_level_download")
    return dst
end

sync(ctx::Context) = check(ctx, ccall((:ludwig_sync, LIB), Cint, (Ptr{Cvoid},), ctx.h), "ludwig_sync")

# every behaviour switch of the library is an explicit option (it reads no environment variable); see include/ludwig_b200.h
set_option!(ctx::Context, key::AbstractString, value) =
    check(ctx, ccall((:ludwig_ctx_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Cstring), ctx.h, key, string(value)), "ludwig_ctx_set_option($key)")

# io_vtk.jl:17-46  valid_blocks: per level the 1-based b_idx of the blocks that are not fully covered by the next finer level
function output_valid_blocks(ctx::Context, n_levels::Integer)
    n = zeros(Int32, n_levels)
    check(ctx, ccall((:ludwig_output_valid_blocks, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), ctx.h, n, C_NULL), "ludwig_output_valid_blocks")
    flat = Vector{Int32}(undef, sum(n))
    check(ctx, ccall((:ludwig_output_valid_blocks, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}), ctx.h, n, flat), "ludwig_output_valid_blocks")
    return n, flat
end

# io_vtk.jl:52-58 + 100-111 in ONE call: rho_arr, vel_mat (3 x N), obst_arr, level_arr of every valid block in the writer's own order
# (level-major, b_idx ascending).  Only those blocks leave the device, through pinned double-buffered staging.
function output_export!(ctx::Context, t_step::Integer, rho_arr::Vector{Float32}, vel_mat::Matrix{Float32}, obst_arr::Vector{UInt8}, level_arr::Vector{Int32})
    GC.@preserve rho_arr vel_mat obst_arr level_arr check(ctx, ccall((:ludwig_output_export, LIB), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Float32}, Ptr{Float32}, Ptr{UInt8}, Ptr{Int32}),
        ctx.h, t_step, pointer(rho_arr), pointer(vel_mat), pointer(obst_arr), pointer(level_arr)), "ludwig_output_export")
end

# ---- domain build on the device (N2): drop-in for the threaded CPU loops of domain_generation.jl / bouzidi_setup.jl ------------------
# tris: 3 x 3 x n_tri Float64 (vertex component fastest), coords: 3 x nb Int32 (active_block_coords), grid_ptr: the level's FULL
# block-pointer grid as Int32[bz, by, bx]-major memory = permutedims(block_pointer_full, (3, 2, 1)), 1-based, 0 = none
function domain_check(rc, what)
    rc >= 0 && return rc
    error("$what failed ($rc): " * unsafe_string(ccall((:ludwig_domain_last_error, LIB), Cstring, ())))
end
# domain_generation.jl:74-112  voxelize_blocks!
voxelize!(device, tris, offset, dx, coords, grid_ptr, dims, obstacle::Array{UInt8,4}) =
    domain_check(ccall((:ludwig_domain_voxelize, LIB), Cint, (Cint, Ptr{Float64}, Int64, Ptr{Float64}, Float64, Ptr{Int32}, Int32, Ptr{Int32}, Int32, Int32, Int32, Ptr{UInt8}),
                       device, tris, size(tris, 3), Float64[offset...], dx, coords, size(coords, 2), grid_ptr, dims[1], dims[2], dims[3], obstacle), "ludwig_domain_voxelize")
# domain_generation.jl:114-203  perform_flood_fill!  (returns the number of filled voxels)
flood_fill!(device, coords, grid_ptr, dims, obstacle::Array{UInt8,4}) =
    domain_check(ccall((:ludwig_domain_flood_fill, LIB), Int64, (Cint, Ptr{Int32}, Int32, Ptr{Int32}, Int32, Int32, Int32, Ptr{UInt8}),
                       device, coords, size(coords, 2), grid_ptr, dims[1], dims[2], dims[3], obstacle), "ludwig_domain_flood_fill")
# domain_generation.jl:371-431  compute_wall_distances!  (returns the number of near-wall cells)
wall_distance!(device, neighbor_table::Matrix{Int32}, obstacle::Array{UInt8,4}, dx, wall_dist::Array{Float32,4}) =
    domain_check(ccall((:ludwig_domain_wall_distance, LIB), Int64, (Cint, Ptr{Int32}, Int32, Ptr{UInt8}, Float64, Ptr{Float32}),
                       device, neighbor_table, size(neighbor_table, 1), obstacle, dx, wall_dist), "ludwig_domain_wall_distance")
# bouzidi_setup.jl:64-166  compute_q_map!: sparse rows (cells 4 x n, q 27 x n Float64, tri 27 x n); call once with capacity 0 for n
qmap!(device, tris, offset, dx, coords, grid_ptr, dims, capacity, cells, q, tri) =
    domain_check(ccall((:ludwig_domain_qmap, LIB), Int64,
                       (Cint, Ptr{Float64}, Int64, Ptr{Float64}, Float64, Ptr{Int32}, Int32, Ptr{Int32}, Int32, Int32, Int32, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}),
                       device, tris, size(tris, 3), Float64[offset...], dx, coords, size(coords, 2), grid_ptr, dims[1], dims[2], dims[3], capacity,
                       cells === nothing ? C_NULL : cells, q === nothing ? C_NULL : q, tri === nothing ? C_NULL : tri), "ludwig_domain_qmap")

# io_vtk.jl:52-58,100-111 for one level: fills the slices of rho_arr / vel_mat / obst_arr that belong to the level's valid blocks
# (b_idx list of io_vtk.jl:27-45, in the order they appear in valid_blocks) without downloading whole arrays.
function output_gather!(ctx::Context, level::Integer, t_step::Integer, b_idx::Vector{Int32},
                        rho_arr::AbstractVector{Float32}, vel_mat::AbstractMatrix{Float32}, obst_arr::AbstractVector{UInt8})
    GC.@preserve b_idx rho_arr vel_mat obst_arr check(ctx, ccall((:ludwig_output_gather, LIB), Cint,
        (Ptr{Cvoid}, Int32, Int64, Ptr{Int32}, Int32, Ptr{Float32}, Ptr{Float32}, Ptr{UInt8}),
        ctx.h, level, t_step, pointer(b_idx), length(b_idx), pointer(rho_arr), pointer(vel_mat), pointer(obst_arr)), "ludwig_output_gather")
end

# ---- more than one GPU: one process per GPU (INTEGRATION.md "More than one GPU") -----------------------------------
# call right after Context(local_rank), before the first add_level!
set_partition!(ctx::Context, rank::Integer, world::Integer) =
    check(ctx, ccall((:ludwig_ctx_set_partition, LIB), Cint, (Ptr{Cvoid}, Int32, Int32), ctx.h, rank, world), "ludwig_ctx_set_partition")

# after the last add_level!: `allgather(bytes) -> bytes of every rank concatenated` is the host's collective
# (MPI.Allgather(buf, comm) with MPI.jl).  From then on the kernels read the peers' blocks over NVLink.
function attach_peers!(ctx::Context, allgather::Function)
    need = Ref{Int64}(0)
    check(ctx, ccall((:ludwig_ipc_export, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Int64}), ctx.h, C_NULL, 0, need), "ludwig_ipc_export")
    buf = Vector{UInt8}(undef, need[])
    GC.@preserve buf check(ctx, ccall((:ludwig_ipc_export, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Int64}), ctx.h, pointer(buf), need[], need), "ludwig_ipc_export")
    all = allgather(buf)
    GC.@preserve all check(ctx, ccall((:ludwig_ipc_attach, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64), ctx.h, pointer(all), need[]), "ludwig_ipc_attach")
end

# ---- more than one GPU, ONE process (the kept single-process driver, main.jl:54-249, unchanged in structure) ------------------------
# Multi(n) owns one context per GPU; the same sequence of calls as for one GPU: add_level!, init_equilibrium!, step_batch!, flow_stats,
# compute_aerodynamics!.  The library partitions every level, attaches the ranks in-process and steps them in lock-step from this
# thread; cross-rank barriers are stream-ordered event waits.
mutable struct Multi
    h::Ptr{Cvoid}
end
function mcheck(m::Multi, rc::Cint, what)
    rc == 0 && return
    error("$what failed ($rc): " * unsafe_string(ccall((:ludwig_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), m.h)))
end
function Multi(n_ranks::Integer, devices::Vector{Int32}=Int32.(0:n_ranks-1))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:ludwig_multi_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Int32, Ptr{Int32}), h, n_ranks, devices)
    rc == 0 || error("ludwig_multi_create failed ($rc)")
    m = Multi(h[])
    finalizer(x -> ccall((:ludwig_multi_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), m)
    return m
end
set_option!(m::Multi, key::AbstractString, value) =
    mcheck(m, ccall((:ludwig_multi_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Cstring), m.h, key, string(value)), "ludwig_multi_set_option($key)")
# add_level!(m, g): build the LevelDesc exactly as add_level!(ctx, g) does and pass it to ludwig_multi_level_create
level_create!(m::Multi, d::LevelDesc) = (idx = Ref{Int32}(-1);
    mcheck(m, ccall((:ludwig_multi_level_create, LIB), Cint, (Ptr{Cvoid}, Ref{LevelDesc}, Ref{Int32}), m.h, d, idx), "ludwig_multi_level_create"); idx[])
init_equilibrium!(m::Multi) = mcheck(m, ccall((:ludwig_multi_init_equilibrium, LIB), Cint, (Ptr{Cvoid},), m.h), "ludwig_multi_init_equilibrium")
step_batch!(m::Multi, t_start::Integer, batch::Integer, u_curr::Float32, p::Params) =
    mcheck(m, ccall((:ludwig_multi_step_batch, LIB), Cint, (Ptr{Cvoid}, Int64, Int32, Float32, Ref{Params}), m.h, t_start, batch, u_curr, p), "ludwig_multi_step_batch")
sync(m::Multi) = mcheck(m, ccall((:ludwig_multi_sync, LIB), Cint, (Ptr{Cvoid},), m.h), "ludwig_multi_sync")
function flow_stats(m::Multi, level::Integer=0)
    out = zeros(Float64, 6)
    mcheck(m, ccall((:ludwig_multi_flow_stats, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), m.h, level, out), "ludwig_multi_flow_stats")
    return (n_fluid=Int(out[1]), rho_mean=out[2], rho_min=out[3], rho_max=out[4], v_max=out[5], kinetic_energy=out[6])
end
# ludwig_multi_forces_create / ludwig_multi_compute_aerodynamics / ludwig_multi_output_export follow the single-context signatures
# with the handle index instead of the mesh / forces pointers (include/ludwig_b200.h).

end # module
