// gem_b200/csrc/gram_tc.cu -- the tall-skinny Gram contraction G = P^T Q on the 5th-gen tensor cores.
//
// This is the ONE dense contraction of the HOPE solver (CholeskyQR Gram and the Rayleigh-Ritz
// projection; replaces numpy.linalg.qr / svd inside scipy svds, _svds.py:508-533).  It is memory bound
// (read n*b fp32 once), so the kernel is a persistent streaming design, one CTA per SM:
//
//   TMA producer (1 thread): a stage of 64 rows is one contiguous 20 KB piece of the row-major block:
//                           cp.async.bulk global -> shared, three raw stages in flight, mbarrier complete_tx
//   transform (8 warps)   : each lane reads 4 consecutive rows (K) of one column (MN) from the raw stage
//                           (conflict-free LDS.32), splits x = hi + lo with hi = rna_tf32(x),
//                           lo = rna_tf32(x - hi), and writes two STS.128 into the UMMA canonical K-major /
//                           no-swizzle layout (element (mn, k) -> (mn/8)*SBO + (mn%8)*16 + (k/4)*LBO + (k%4)*4)
//   MMA issuer (1 thread) : per 8-row k-block three tcgen05.mma.kind::tf32 (hi*hi + hi*lo + lo*hi =
//                           "3xTF32", ~fp32 accuracy), M = 128, N = 32..128, a fresh fp32 TMEM accumulator per stage
//   drain (warps 0-3)     : tcgen05.ld of the finished stage accumulator into fp32 registers (round-to-nearest
//                           adds; two TMEM buffers so the next stage's MMAs overlap), fp64 atomicAdd into G at the end
// tcgen05.commit -> mbarrier hands a stage back to the loaders.  The contraction index is the ROW index of the
// row-major blocks, i.e. A = P^T and B = Q^T arrive "MN-major"; kind::tf32 with MN-major descriptors returned zeros
// on the B200 (scripts/tc_probe.cu), so the loader transposes to K-major on the way into shared memory.
#include "tc_common.cuh"

namespace gemb {



struct GramTcParams {
    int64_t n;
    const float *P, *Q;
    int b1, b2;
    double *G;
    int stage_rows;      // multiple of 8
    int n_pad;           // pad16(b2)
    int m_tiles;         // ceil(b1 / 128)
    uint32_t tile_bytes_p, tile_bytes_q;   // bytes of one (hi or lo) tile
    uint32_t raw_bytes_p, raw_bytes_q;     // bytes of one raw (row-major fp32) stage
    uint32_t tmem_cols;  // power of two >= m_tiles * n_pad
};

// one unit = 4 rows x 32 columns of the RAW stage (row-major fp32, as it lies in global memory): lane -> column
// mn = j32*32 + lane reads rows k4*4 .. k4*4+3 (four conflict-free LDS.32), i.e. exactly the 16-byte K chunk the
// K-major tile wants -- no register transpose.  Rows >= valid_rows read as zero.
__device__ __forceinline__ void transform_unit(const char *raw, int b, int valid_rows, int k4, int j32, int lane,
                                               char *tile_hi, char *tile_lo, uint32_t lbo, uint32_t sbo) {
    const int mn = j32 * 32 + lane;
    if (mn >= b) return;
    const int k0 = k4 * 4;
    const float *src = (const float *)raw + (size_t)k0 * b + mn;
    float w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = (k0 + i < valid_rows) ? src[(size_t)i * b] : 0.f;
    uint4 hi, lo;
    hi.x = tc::to_tf32(w[0]); hi.y = tc::to_tf32(w[1]); hi.z = tc::to_tf32(w[2]); hi.w = tc::to_tf32(w[3]);
    lo.x = tc::to_tf32(w[0] - __uint_as_float(hi.x));
    lo.y = tc::to_tf32(w[1] - __uint_as_float(hi.y));
    lo.z = tc::to_tf32(w[2] - __uint_as_float(hi.z));
    lo.w = tc::to_tf32(w[3] - __uint_as_float(hi.w));
    const uint32_t off = (uint32_t)(mn >> 3) * sbo + (uint32_t)(mn & 7) * 16u + (uint32_t)k4 * lbo;
    *(uint4 *)(tile_hi + off) = hi;
    *(uint4 *)(tile_lo + off) = lo;
}

// NC16 = accumulator width in units of 16 columns (N of the MMA = 16*NC16 >= b2).
// Warp roles: warps 0-7 transform (raw -> UMMA tiles; warps 0-3 also drain TMEM), warp 8 lane 0 is the TMA
// producer and the MMA issuer.  All hand-offs are mbarriers:
//   s_full[3]  TMA complete_tx -> transform      raw_free[3]  transform (8 warps) -> producer
//   tile_full[2] transform (8 warps) -> MMA      s_bar[2]     tcgen05.commit -> transform (tile free, TMEM complete)
// Accumulation is two-level: every stage (<= 64 rows) accumulates into a FRESH TMEM buffer (the tensor core
// adds in fp32 with truncation, which biases long sums of positive terms: measured -7e-5 relative on the
// diagonal at 6.7k rows per CTA), and warps 0-3 fold the finished buffer into fp32 registers with
// round-to-nearest while the next stage is already in flight (two TMEM buffers).
template <bool CROSS, int NC16>
__global__ void __launch_bounds__(288, 1) gram_tc_kernel(GramTcParams p) {
    extern __shared__ __align__(128) char smem[];
    __shared__ __align__(8) uint64_t s_bar[2], s_full[3], s_raw_free[3], s_tile_full[2];
    __shared__ uint32_t s_tmem;
    constexpr int NPAD = 16 * NC16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kblocks = p.stage_rows / 8;              // MMA k-steps (K = 8 rows) per stage
    const int kquads = p.stage_rows / 4;               // 16-byte K chunks per stage
    const uint32_t lbo = 128u;                         // consecutive K chunks of one 8-column group are adjacent
    const uint32_t sbo = (uint32_t)kquads * 128u;      // stride between 8-column (MN) groups
    // shared memory: 3 raw stages [P | (Q)] filled by TMA bulk copies, then 2 tile stages [P_hi | P_lo | (Q_hi | Q_lo)]
    const uint32_t raw_stage = p.raw_bytes_p + (CROSS ? p.raw_bytes_q : 0);
    const uint32_t stage_bytes = 2 * p.tile_bytes_p + (CROSS ? 2 * p.tile_bytes_q : 0);
    char *raw_base = smem;
    char *tile_base = smem + 3 * (size_t)raw_stage;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) { tc::mbar_init(tc::smem_u32(&s_bar[i]), 1); tc::mbar_init(tc::smem_u32(&s_tile_full[i]), 8); }
        for (int i = 0; i < 3; i++) { tc::mbar_init(tc::smem_u32(&s_full[i]), 1); tc::mbar_init(tc::smem_u32(&s_raw_free[i]), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&s_tmem)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    // contiguous row range of this CTA, in whole stages
    const int64_t stages_total = (p.n + p.stage_rows - 1) / p.stage_rows;
    const int64_t per_cta = (stages_total + gridDim.x - 1) / gridDim.x;
    const int64_t s_begin = (int64_t)blockIdx.x * per_cta;
    const int64_t s_end = s_begin + per_cta < stages_total ? s_begin + per_cta : stages_total;
    const int nst = (int)(s_end > s_begin ? s_end - s_begin : 0);

    if (warp == 8) {
        // ================= producer + MMA issuer (one thread) =================
        if (lane == 0 && nst > 0) {
            // instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = NPAD
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) |
                                   ((uint32_t)(NPAD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            auto issue = [&](int it) {
                const int slot = it % 3;
                const int64_t row0 = (s_begin + it) * p.stage_rows;
                const int64_t vr = p.n - row0 < p.stage_rows ? p.n - row0 : p.stage_rows;
                const uint32_t bytes_p = (uint32_t)(vr * p.b1 * 4), bytes_q = CROSS ? (uint32_t)(vr * p.b2 * 4) : 0u;
                const uint32_t bar = tc::smem_u32(&s_full[slot]);
                char *dst = raw_base + (size_t)slot * raw_stage;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes_p + bytes_q) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(tc::smem_u32(dst)), "l"(p.P + row0 * p.b1), "r"(bytes_p), "r"(bar) : "memory");
                if (CROSS)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(tc::smem_u32(dst + p.raw_bytes_p)), "l"(p.Q + row0 * p.b2), "r"(bytes_q), "r"(bar) : "memory");
            };
            for (int it = 0; it < 3 && it < nst; it++) issue(it);
            for (int it = 0; it < nst; it++) {
                const int st = it & 1;
                tc::mbar_wait(tc::smem_u32(&s_tile_full[st]), (uint32_t)(it >> 1) & 1);      // tiles of stage `it` are written
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = tc::smem_u32(tile_base + (size_t)st * stage_bytes), a_lo = a_hi + p.tile_bytes_p;
                const uint32_t b_hi = CROSS ? a_hi + 2 * p.tile_bytes_p : a_hi;
                const uint32_t b_lo = CROSS ? b_hi + p.tile_bytes_q : a_lo;
                const uint32_t d_tmem = tmem + (uint32_t)(st * NPAD);
                uint64_t dah = tc::make_desc(a_hi, lbo, sbo), dal = tc::make_desc(a_lo, lbo, sbo);
                uint64_t dbh = tc::make_desc(b_hi, lbo, sbo), dbl = tc::make_desc(b_lo, lbo, sbo);
                const uint64_t step = (uint64_t)((2u * lbo) >> 4);                           // one MMA consumes two 16-byte K chunks
                for (int kb = 0; kb < kblocks; kb++) {
                    tc::mma_tf32(d_tmem, dah, dbh, idesc, kb > 0 ? 1u : 0u);
                    tc::mma_tf32(d_tmem, dah, dbl, idesc, 1u);
                    tc::mma_tf32(d_tmem, dal, dbh, idesc, 1u);
                    dah += step; dal += step; dbh += step; dbl += step;
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(&s_bar[st])) : "memory");
                if (it + 3 < nst) {                                                            // refill the raw slot stage `it` used
                    tc::mbar_wait(tc::smem_u32(&s_raw_free[it % 3]), (uint32_t)(it / 3) & 1);
                    issue(it + 3);
                }
            }
        }
    } else {
        // ================= transform warps (0-7); warps 0-3 also drain TMEM =================
        float racc[NPAD];
#pragma unroll
        for (int i = 0; i < NPAD; i++) racc[i] = 0.f;
        auto drain = [&](int st) {   // fold TMEM buffer `st` (this warp's 32 lanes) into the register accumulators
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c = 0; c < NC16; c++) {
                uint32_t r[16];
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(st * NPAD + c * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; i++) racc[c * 16 + i] += __uint_as_float(r[i]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        };
        for (int it = 0; it < nst; it++) {
            const int st = it & 1, slot = it % 3;
            char *base = tile_base + (size_t)st * stage_bytes;
            const char *raw = raw_base + (size_t)slot * raw_stage;
            tc::mbar_wait(tc::smem_u32(&s_full[slot]), (uint32_t)(it / 3) & 1);             // stage `it` has landed
            if (it >= 2) {
                tc::mbar_wait(tc::smem_u32(&s_bar[st]), (uint32_t)((it >> 1) - 1) & 1);       // MMAs of stage it-2 are done
                if (warp < 4) drain(st);
            }
            const int64_t row0 = (s_begin + it) * p.stage_rows;
            const int vr = (int)(p.n - row0 < p.stage_rows ? p.n - row0 : p.stage_rows);
            const int units_p = kquads * ((p.b1 + 31) / 32);
#pragma unroll 3
            for (int u = warp; u < units_p; u += 8)
                transform_unit(raw, p.b1, vr, u % kquads, u / kquads, lane, base, base + p.tile_bytes_p, lbo, sbo);
            if (CROSS) {
                const int units_q = kquads * ((p.b2 + 31) / 32);
                char *qb = base + 2 * p.tile_bytes_p;
#pragma unroll 3
                for (int u = warp; u < units_q; u += 8)
                    transform_unit(raw + p.raw_bytes_p, p.b2, vr, u % kquads, u / kquads, lane, qb, qb + p.tile_bytes_q, lbo, sbo);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&s_raw_free[slot])) : "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&s_tile_full[st])) : "memory");
            }
        }
        // ---- drain what is still in flight: stages nst-2 and nst-1 (each use of a buffer is waited for exactly once)
        for (int it = nst >= 2 ? nst - 2 : 0; it < nst; it++) {
            const int st = it & 1;
            tc::mbar_wait(tc::smem_u32(&s_bar[st]), (uint32_t)(it >> 1) & 1);
            if (warp < 4) drain(st);
        }
        if (warp < 4 && nst > 0) {
            const int row = warp * 32 + lane;                                 // TMEM lane = accumulator row
            if (row < p.b1) {
#pragma unroll
                for (int i = 0; i < NPAD; i++)
                    if (i < p.b2) atomicAdd(p.G + (size_t)row * p.b2 + i, (double)racc[i]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

template <bool CROSS, int NC16>
static int gram_tc_launch_t(gemb_ctx *ctx, const GramTcParams &p, int grid, size_t smem_bytes) {
    static size_t attr_bytes = 0;
    if (attr_bytes < smem_bytes) {
        GEMB_CUDA(cudaFuncSetAttribute(gram_tc_kernel<CROSS, NC16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        attr_bytes = smem_bytes;
    }
    gram_tc_kernel<CROSS, NC16><<<grid, 288, smem_bytes, ctx->stream>>>(p);
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    return GEMB_OK;
}

// returns GEMB_ERR_UNSUPPORTED (without setting an error) when the shape does not fit this kernel
int gram_tc_launch(gemb_ctx *ctx, int64_t n, const float *P, int b1, const float *Q, int b2, double *G) {
    const bool cross = (P != Q);
    if (b1 % 4 || b2 % 4 || b1 > 128 || b2 > 128 || n <= 0) return GEMB_ERR_UNSUPPORTED;
    GramTcParams p;
    p.n = n; p.P = P; p.Q = Q; p.b1 = b1; p.b2 = b2; p.G = G;
    // accumulator width: one of 32 / 64 / 80 / 96 / 128 columns
    const int widths[5] = {32, 64, 80, 96, 128};
    int npad = 128;
    for (int w : widths) if (w >= b2) { npad = w; break; }
    p.n_pad = npad;
    p.m_tiles = 1;
    p.tmem_cols = 32;
    while (p.tmem_cols < (uint32_t)(2 * npad)) p.tmem_cols <<= 1;
    // stage rows: largest multiple of 8 (<= 64) such that 3 raw stages + 2 tile stages (hi + lo) fit in ~176 KB
    const size_t cols_raw = (size_t)b1 + (cross ? b2 : 0);
    const size_t cols_tile = (size_t)(b1 + 7) / 8 * 8 + (cross ? (size_t)(b2 + 7) / 8 * 8 : 0);
    const size_t bytes_per_row = 3 * 4 * cols_raw + 2 * 2 * 4 * cols_tile;
    int rows = (int)((176 * 1024) / bytes_per_row) / 8 * 8;
    if (rows > 64) rows = 64;
    if (rows < 8) return GEMB_ERR_UNSUPPORTED;
    p.stage_rows = rows;
    p.raw_bytes_p = (uint32_t)(rows * b1 * 4);
    p.raw_bytes_q = (uint32_t)(rows * b2 * 4);
    p.tile_bytes_p = (uint32_t)((b1 + 7) / 8) * (uint32_t)(rows / 4) * 128u;
    p.tile_bytes_q = (uint32_t)((b2 + 7) / 8) * (uint32_t)(rows / 4) * 128u;
    const size_t stage_bytes = 2 * (size_t)p.tile_bytes_p + (cross ? 2 * (size_t)p.tile_bytes_q : 0);
    // the MMA reads 16 column groups of A and npad/8 groups of B even where the block is narrower:
    // keep those (ignored) reads inside the allocation
    const size_t sbo = (size_t)(rows / 4) * 128;
    const size_t over = std::max<size_t>((size_t)16 * sbo, (size_t)(npad / 8) * sbo);
    const size_t raw_stage = (size_t)p.raw_bytes_p + (cross ? p.raw_bytes_q : 0);
    const size_t smem_bytes = 3 * raw_stage + 2 * stage_bytes + over + 1024;
    if (smem_bytes > 226 * 1024) return GEMB_ERR_UNSUPPORTED;   // + ~64 B of static shared memory <= 227 KB
    GEMB_CUDA(cudaMemsetAsync(G, 0, sizeof(double) * (size_t)b1 * b2, ctx->stream));
    const int64_t stages_total = (n + rows - 1) / rows;
    int grid = ctx->sm_count;
    if (grid > stages_total) grid = (int)stages_total;
#define GEMB_TC_CASE(W)                                                                                \
    case W: return cross ? gram_tc_launch_t<true, W / 16>(ctx, p, grid, smem_bytes)                   \
                         : gram_tc_launch_t<false, W / 16>(ctx, p, grid, smem_bytes);
    switch (npad) {
        GEMB_TC_CASE(32)
        GEMB_TC_CASE(64)
        GEMB_TC_CASE(80)
        GEMB_TC_CASE(96)
        GEMB_TC_CASE(128)
    }
#undef GEMB_TC_CASE
    return GEMB_ERR_UNSUPPORTED;
}

}  // namespace gemb
