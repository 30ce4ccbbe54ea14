// output.cu — N3, the step AFTER the hot path: what export_merged_mesh_sync (io_vtk.jl:12-124) needs from the device.
//
// The reference (a) decides on the host which blocks to write — those NOT fully covered by the next finer level (all 8
// children present, io_vtk.jl:17-46) — and (b) copies the WHOLE rho / vel (or vel_temp) / obstacle arrays of every level to
// the host (`Array(level.rho)` ..., :52-58) to pick those blocks' cells one by one (:100-107); its logs show 11-28 s per dump.
// Here:
//   * the valid-block lists are computed ON THE DEVICE from the block-pointer tables (one thread per block probes its 8 child
//     coordinates in the finer level's table; a one-CTA scan compacts the flags in the reference's b_idx order), once, cached;
//   * a gather kernel writes only those blocks, already in the VTK writer's array layout (rho[N], vel[3][N] component-fastest,
//     obstacle[N], level[N]; non-finite -> 0 as :110-111), into a device staging buffer;
//   * staging buffers are persistent, PINNED and double-buffered: while chunk i is copied device -> pinned host -> caller's
//     (pageable) array, the kernel already gathers chunk i + 1.  No allocation per call.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "ludwig_internal.h"

using namespace ludwig;

namespace {

int ofail(ludwig_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}
#define OCU(call)                                                                                                    \
    do {                                                                                                             \
        cudaError_t e__ = (call);                                                                                    \
        if (e__ != cudaSuccess) return ofail(ctx, LUDWIG_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// covered[b] = 1 iff all 8 children (2 bx + {0,1}, ...) of LOCAL block b exist in the finer level's block-pointer table (0-based coords)
__global__ void covered_flags_kernel(const int32_t* __restrict__ bcoord, int nb, const int32_t* __restrict__ cptr, int cdx, int cdy, int cdz,
                                     int32_t* __restrict__ keep) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const int bx = bcoord[4 * b], by = bcoord[4 * b + 1], bz = bcoord[4 * b + 2];
    int children = 0;
    for (int d = 0; d < 8; ++d) {
        const int x = 2 * bx + (d & 1), y = 2 * by + ((d >> 1) & 1), z = 2 * bz + (d >> 2);
        if (x < cdx && y < cdy && z < cdz && cptr[x + (size_t)cdx * (y + (size_t)cdy * z)] >= 0) ++children;
    }
    keep[b] = children == 8 ? 0 : 1;
}

constexpr int CHUNK_BLOCKS = 2048;                                   // 2048 x 512 cells x 21 B = 22 MB per staging buffer
constexpr size_t CHUNK_CELLS = (size_t)CHUNK_BLOCKS * BS3;
constexpr size_t STAGE_BYTES = CHUNK_CELLS * (4 * sizeof(float) + 1);   // rho + vel[3] + obstacle

struct OutputState {
    bool planned = false;
    std::vector<std::vector<int32_t>> valid_local;   // per level: LOCAL block indices of this rank's valid blocks, ascending reference index
    std::vector<std::vector<int32_t>> valid_pos;     // per level: position of each of them in the level's global valid list
    std::vector<int32_t> n_valid_global;             // per level: valid blocks of the whole level (all ranks)
    std::vector<std::vector<int32_t>> valid_ref;     // per level: 1-based reference indices of the global valid list
    uint8_t* d_stage[2] = {nullptr, nullptr};
    uint8_t* h_stage[2] = {nullptr, nullptr};        // pinned
    int32_t* d_sel = nullptr;                        // [2][CHUNK_BLOCKS]
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
};

OutputState* state_of(ludwig_ctx* ctx) {
    if (!ctx->output_state) ctx->output_state = new OutputState();
    return (OutputState*)ctx->output_state;
}

int ensure_staging(ludwig_ctx* ctx, OutputState& S) {
    if (S.d_stage[0]) return LUDWIG_OK;
    for (int i = 0; i < 2; ++i) {
        OCU(cudaMalloc((void**)&S.d_stage[i], STAGE_BYTES));
        OCU(cudaHostAlloc((void**)&S.h_stage[i], STAGE_BYTES, cudaHostAllocDefault));
        OCU(cudaEventCreateWithFlags(&S.ev[i], cudaEventDisableTiming));
    }
    OCU(cudaMalloc((void**)&S.d_sel, 2 * CHUNK_BLOCKS * sizeof(int32_t)));
    OCU(cudaStreamCreateWithFlags(&S.copy_stream, cudaStreamNonBlocking));
    ctx->bytes += (int64_t)(2 * STAGE_BYTES + 2 * CHUNK_BLOCKS * sizeof(int32_t));
    return LUDWIG_OK;
}

// The valid-block lists (io_vtk.jl:17-46).  A block's validity depends only on the finer level's table, which every rank holds
// whole: each rank flags its own blocks on the device; the GLOBAL list (needed for output positions) is derived on the host from
// the same rule applied to the host copy of the tables, so that ranks agree without communicating.
int ensure_plan(ludwig_ctx* ctx, OutputState& S) {
    if (S.planned) return LUDWIG_OK;
    const size_t nl = ctx->levels.size();
    S.valid_local.assign(nl, {}); S.valid_pos.assign(nl, {}); S.valid_ref.assign(nl, {}); S.n_valid_global.assign(nl, 0);
    for (size_t l = 0; l < nl; ++l) {
        Level& L = *ctx->levels[l];
        std::vector<int32_t> keep(L.nb, 1);
        if (l + 1 < nl) {
            Level& Cn = *ctx->levels[l + 1];
            int32_t* d_keep = nullptr;
            OCU(cudaMalloc((void**)&d_keep, (size_t)L.nb * sizeof(int32_t)));
            covered_flags_kernel<<<(L.nb + 255) / 256, 256, 0, ctx->stream>>>(L.d_bcoord, L.nb, Cn.d_ptr, Cn.dimx, Cn.dimy, Cn.dimz, d_keep);
            cudaError_t e = cudaMemcpyAsync(keep.data(), d_keep, (size_t)L.nb * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            cudaFree(d_keep);
            if (e != cudaSuccess) return ofail(ctx, LUDWIG_ECUDA, std::string("valid-block flags: ") + cudaGetErrorString(e));
            ctx->launches += 1;
        }
        // global list in reference order; this rank's share of it.  Other ranks' flags come from the host tables (same rule).
        std::vector<uint8_t> keep_ref(L.nb_global, 1);
        if (l + 1 < nl) {
            Level& Cn = *ctx->levels[l + 1];
            // block coordinates of every block of the level: recover them from the (whole-level) pointer table
            std::vector<int32_t> cx(L.nb_global), cy(L.nb_global), cz(L.nb_global);
            for (int z = 0; z < L.dimz; ++z)
                for (int y = 0; y < L.dimy; ++y)
                    for (int x = 0; x < L.dimx; ++x) {
                        const int32_t enc = L.h_ptr[x + (size_t)L.dimx * (y + (size_t)L.dimy * z)];
                        if (enc < 0) continue;
                        const int ow = enc >> PTR_RANK_SHIFT, loc = enc & PTR_LOCAL_MASK;
                        const int gi = L.part_starts[ow] + loc;
                        const int br = L.int2ref[gi];
                        cx[br] = x; cy[br] = y; cz[br] = z;
                    }
            for (int br = 0; br < L.nb_global; ++br) {
                int children = 0;
                for (int d = 0; d < 8; ++d) {
                    const int x = 2 * cx[br] + (d & 1), y = 2 * cy[br] + ((d >> 1) & 1), z = 2 * cz[br] + (d >> 2);
                    if (x < Cn.dimx && y < Cn.dimy && z < Cn.dimz && Cn.h_ptr[x + (size_t)Cn.dimx * (y + (size_t)Cn.dimy * z)] >= 0) ++children;
                }
                keep_ref[br] = children == 8 ? 0 : 1;
            }
        }
        int pos = 0;
        for (int br = 0; br < L.nb_global; ++br) {
            if (!keep_ref[br]) continue;
            S.valid_ref[l].push_back(br + 1);
            const int loc = L.ref2int[br] - L.part_start;
            if (loc >= 0 && loc < L.nb) {
                if (!keep[loc]) return ofail(ctx, LUDWIG_ESTATE, "valid-block plan: device flags disagree with the host tables");
                S.valid_local[l].push_back(loc); S.valid_pos[l].push_back(pos);
            }
            ++pos;
        }
        S.n_valid_global[l] = pos;
        size_t mine = 0;
        for (int b = 0; b < L.nb; ++b) mine += keep[b] != 0;
        if (mine != S.valid_local[l].size()) return ofail(ctx, LUDWIG_ESTATE, "valid-block plan: device / host counts differ");
    }
    S.planned = true;
    return LUDWIG_OK;
}

// Gathers `n` local blocks (sel, host) of level L into the caller's arrays at cell offsets pos[i] * 512, through the pinned
// double-buffered staging pipeline.  level_arr (optional) is filled on the host.
int gather_blocks(ludwig_ctx* ctx, OutputState& S, Level& L, int64_t t_step, const int32_t* sel, const int32_t* pos, int n, float* rho_arr,
                  float* vel_mat, uint8_t* obst_arr, int32_t* level_arr) {
    int rc = ensure_staging(ctx, S);
    if (rc) return rc;
    const float* vel = (t_step % 2 == 0) ? L.d_vel[1] : L.d_vel[0];   // io_vtk.jl:56
    const int n_chunks = (n + CHUNK_BLOCKS - 1) / CHUNK_BLOCKS;
    auto drain = [&](int c) -> int {   // chunk c: pinned -> caller's arrays
        const int buf = c & 1, first = c * CHUNK_BLOCKS, cnt = std::min(CHUNK_BLOCKS, n - first);
        OCU(cudaEventSynchronize(S.ev[buf]));
        const size_t nc = (size_t)cnt * BS3;
        const float* h_rho = (const float*)S.h_stage[buf];
        const float* h_vel = h_rho + nc;
        const uint8_t* h_obs = (const uint8_t*)(h_vel + 3 * nc);
        for (int i = 0; i < cnt; ++i) {   // consecutive positions are the common case: copy block by block (2 KiB / 6 KiB / 512 B runs)
            const size_t o = (size_t)pos[first + i] * BS3;
            std::memcpy(rho_arr + o, h_rho + (size_t)i * BS3, BS3 * sizeof(float));
            std::memcpy(vel_mat + 3 * o, h_vel + 3 * (size_t)i * BS3, 3 * BS3 * sizeof(float));
            std::memcpy(obst_arr + o, h_obs + (size_t)i * BS3, BS3);
            if (level_arr) std::fill(level_arr + o, level_arr + o + BS3, (int32_t)L.level_id);
        }
        return LUDWIG_OK;
    };
    for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1, first = c * CHUNK_BLOCKS, cnt = std::min(CHUNK_BLOCKS, n - first);
        if (c >= 2 && (rc = drain(c - 2))) return rc;          // the staging buffer about to be reused has been consumed
        const size_t nc = (size_t)cnt * BS3;
        float* d_rho = (float*)S.d_stage[buf];
        float* d_vel = d_rho + nc;
        uint8_t* d_obs = (uint8_t*)(d_vel + 3 * nc);
        OCU(cudaMemcpyAsync(S.d_sel + buf * CHUNK_BLOCKS, sel + first, (size_t)cnt * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        launch_output_gather(S.d_sel + buf * CHUNK_BLOCKS, cnt, L.d_rho[L.rho_cur], vel, L.d_obstacle, d_rho, d_vel, d_obs, ctx->stream);
        OCU(cudaGetLastError());
        ctx->launches += 1;
        OCU(cudaMemcpyAsync(S.h_stage[buf], S.d_stage[buf], nc * (4 * sizeof(float) + 1), cudaMemcpyDeviceToHost, ctx->stream));
        OCU(cudaEventRecord(S.ev[buf], ctx->stream));
        if (c >= 1 && (rc = drain(c - 1))) return rc;          // host copy of chunk c - 1 overlaps the device work of chunk c
    }
    if (n_chunks >= 1 && (rc = drain(n_chunks - 1))) return rc;
    return LUDWIG_OK;
}

}  // namespace

void ludwig_output_state_free(ludwig_ctx* ctx) {
    if (!ctx || !ctx->output_state) return;
    OutputState* S = (OutputState*)ctx->output_state;
    for (int i = 0; i < 2; ++i) {
        if (S->d_stage[i]) cudaFree(S->d_stage[i]);
        if (S->h_stage[i]) cudaFreeHost(S->h_stage[i]);
        if (S->ev[i]) cudaEventDestroy(S->ev[i]);
    }
    if (S->d_sel) cudaFree(S->d_sel);
    if (S->copy_stream) cudaStreamDestroy(S->copy_stream);
    delete S;
    ctx->output_state = nullptr;
}

extern "C" {

// io_vtk.jl:52-58,100-111 for a caller-supplied list of blocks
int ludwig_output_gather(ludwig_ctx* ctx, int32_t level, int64_t t_step, const int32_t* blocks, int32_t n_blocks, float* rho_arr, float* vel_mat,
                         uint8_t* obst_arr) {
    if (!ctx || level < 0 || level >= (int)ctx->levels.size() || !blocks || n_blocks < 0 || !rho_arr || !vel_mat || !obst_arr)
        return ofail(ctx, LUDWIG_EINVAL, "bad gather args");
    if (n_blocks == 0) return LUDWIG_OK;
    OCU(cudaSetDevice(ctx->device));
    Level& L = *ctx->levels[level];
    std::vector<int32_t> sel(n_blocks), pos(n_blocks);
    for (int i = 0; i < n_blocks; ++i) {
        const int br = blocks[i] - 1;
        if (br < 0 || br >= L.nb_global) return ofail(ctx, LUDWIG_EINVAL, "gather: block index out of range");
        const int loc = L.ref2int[br] - L.part_start;
        if (loc < 0 || loc >= L.nb) return ofail(ctx, LUDWIG_EINVAL, "gather: block belongs to another rank");
        sel[i] = loc; pos[i] = i;
    }
    return gather_blocks(ctx, *state_of(ctx), L, t_step, sel.data(), pos.data(), n_blocks, rho_arr, vel_mat, obst_arr, nullptr);
}

// io_vtk.jl:17-46
int ludwig_output_valid_blocks(ludwig_ctx* ctx, int32_t* n_valid, int32_t* blocks) {
    if (!ctx || !n_valid) return ofail(ctx, LUDWIG_EINVAL, "bad valid-block args");
    OCU(cudaSetDevice(ctx->device));
    OutputState& S = *state_of(ctx);
    int rc = ensure_plan(ctx, S);
    if (rc) return rc;
    size_t o = 0;
    for (size_t l = 0; l < ctx->levels.size(); ++l) {
        n_valid[l] = S.n_valid_global[l];
        if (blocks) { std::memcpy(blocks + o, S.valid_ref[l].data(), S.valid_ref[l].size() * sizeof(int32_t)); o += S.valid_ref[l].size(); }
    }
    return LUDWIG_OK;
}

// io_vtk.jl:52-58 + 100-111 for every valid block of every level, in the writer's order (level-major, b_idx ascending)
int ludwig_output_export(ludwig_ctx* ctx, int64_t t_step, float* rho_arr, float* vel_mat, uint8_t* obst_arr, int32_t* level_arr) {
    if (!ctx || !rho_arr || !vel_mat || !obst_arr) return ofail(ctx, LUDWIG_EINVAL, "bad export args");
    OCU(cudaSetDevice(ctx->device));
    OutputState& S = *state_of(ctx);
    int rc = ensure_plan(ctx, S);
    if (rc) return rc;
    size_t base = 0;   // blocks of the coarser levels before this one
    for (size_t l = 0; l < ctx->levels.size(); ++l) {
        const size_t o = base * BS3;
        if (!S.valid_local[l].empty() &&
            (rc = gather_blocks(ctx, S, *ctx->levels[l], t_step, S.valid_local[l].data(), S.valid_pos[l].data(), (int)S.valid_local[l].size(), rho_arr + o,
                                vel_mat + 3 * o, obst_arr + o, level_arr ? level_arr + o : nullptr)))
            return rc;
        base += (size_t)S.n_valid_global[l];
    }
    return LUDWIG_OK;
}

}  // extern "C"
