// k1_common.cuh — pieces shared by the fast (k1_fast.cu) and strict (k1_strict.cu) K1 kernels; textually included inside each
// TU's namespace: the neighbour-offset table of a block and the cp.async.bulk (TMA) / mbarrier primitives of the persistent,
// tile-staged kernel variants.
constexpr long long MISSING = LLONG_MIN;

// neighbour-block offsets of block b relative to f_in / vel_in (threads 0..26); own_f_off: where the block's own populations are
// read from (its place in f_in, or — TMA variant — the shared-memory tile, expressed as an offset from f_in)
__device__ __forceinline__ void neighbour_offsets(const K1Args& a, int b, int t, long long own_f_off, long long* s_fo, long long* s_vo) {
    const int nbi = a.nbr[(size_t)b * 27 + t];
    // indices >= nb address the level's ghost blocks (interface halo, filled by ghost_interp_kernel); >= REMOTE_BASE: another GPU's
    // block, offset of the peer-mapped block relative to the local buffer (pulled over NVLink)
    s_fo[t] = t == 13 ? own_f_off
              : nbi < 0 ? MISSING
              : nbi < a.nb ? (long long)nbi * (Q * BS3)
              : nbi < REMOTE_BASE ? a.ghost_delta + (long long)(nbi - a.nb) * (Q * BS3)
                                  : a.roff_f[nbi - REMOTE_BASE];
    s_vo[t] = nbi < 0 ? MISSING : nbi < a.nb ? (long long)nbi * (3 * BS3) : nbi < REMOTE_BASE ? MISSING : a.roff_v[nbi - REMOTE_BASE];
}

// Software prefetch into L2 (option prefetch_distance = D blocks): the hardware hands out CTAs in list order, so the CTA working on
// list entry i asks the memory system for the part of entry i + D it corresponds to (its z-planes of the 27 population planes and of
// the 3 velocity planes: 128-byte lines) — the loads that arrive D blocks later find their lines in the 126 MB L2 instead of
// paying the full DRAM latency, with no shared-memory staging, no barrier and no extra DRAM traffic (every line is still fetched once).
template <int NT>
__device__ __forceinline__ void prefetch_block_part(const K1Args& a, int entry, int part) {
    if (a.prefetch_distance <= 0) return;
    const int e = entry + a.prefetch_distance;
    if (e >= a.n_list) return;
    const int pb = a.list[e];
    constexpr int LINES_PER_PLANE_PART = (NT * 2 * 4) / 128;            // this CTA's cells of one direction plane, in 128-byte lines (64 threads: 4)
    constexpr int N_LINES = (Q + 3) * LINES_PER_PLANE_PART;
    for (int i = threadIdx.x; i < N_LINES; i += NT) {
        const int plane = i / LINES_PER_PLANE_PART, line = i % LINES_PER_PLANE_PART;
        const float* ptr = plane < Q ? a.f_in + ((size_t)pb * Q + plane) * BS3 + part * (NT * 2) + line * 32
                                     : a.vel_in + ((size_t)pb * 3 + (plane - Q)) * BS3 + part * (NT * 2) + line * 32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
    }
}

// ---- TMA variant primitives: one cp.async.bulk moves a block's own 27 x 2 KiB population planes (one contiguous 54 KiB run in the
// block-major layout) into shared memory and signals an mbarrier with the byte count
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
constexpr int TILE_FLOATS = Q * BS3;                              // 13 824 floats = 55 296 bytes
constexpr uint32_t TILE_BYTES = TILE_FLOATS * (uint32_t)sizeof(float);

