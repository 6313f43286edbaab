// gem_b200/csrc/apply_tc.cu -- the tall-skinny product Out = Q * M (n x b1 times b1 x b2) on the 5th-gen tensor
// cores: the second half of CholeskyQR (Q R^-1) and the Ritz rotation (replaces the GEMMs inside numpy.linalg.qr /
// svd of scipy svds, _svds.py:508-533).  Memory bound (read n*b1, write n*b2 fp32), persistent, one CTA per SM:
//
//   producer / MMA (warp 8, one thread): TMA bulk copy of 128 consecutive rows of Q (one contiguous piece of the
//                       row-major block) into a 2-slot raw ring; per tile 3 * b1/8 tcgen05.mma.kind::tf32
//                       (hi*hi + hi*lo + lo*hi, "3xTF32"), M = 128, N = 32..128, accumulator in TMEM (2 buffers)
//   transform (warps 0-7): raw rows -> K-major / no-swizzle UMMA tile, split x = hi + lo (rna_tf32) on the way;
//                       for A = Q the 16-byte K chunk is 4 consecutive floats of a row, i.e. a straight copy
//   epilogue (warps 0-3): tcgen05.ld of the finished accumulator -> row-major staging in the tile's raw slot ->
//                       ONE cp.async.bulk shared -> global per tile (the 128 x b2 output tile is contiguous)
//   B = M^T (K-major) is split once per CTA and stays resident in shared memory.
#include "tc_common.cuh"

namespace gemb {

struct ApplyTcParams {
    int64_t n;
    const float *Q;      // n x b1
    const float *M;      // b1 x b2, leading dimension ldm
    float *Out;          // n x b2, leading dimension ldo (>= b2, multiple of 4)
    int b1, b2, ldm, ldo;
    uint32_t a_lbo;      // byte stride between 16-byte K chunks of the A tile (2048 + 16: bank-conflict free stores)
    uint32_t a_tile;     // bytes of one (hi or lo) A tile
    uint32_t b_tile;     // bytes of one (hi or lo) B tile
    uint32_t raw_slot;   // bytes of one raw / staging slot
    uint32_t tmem_cols;
};

template <int NC16>
__global__ void __launch_bounds__(288, 1) apply_tc_kernel(ApplyTcParams p) {
    extern __shared__ __align__(128) char smem[];
    __shared__ __align__(8) uint64_t s_full[2], s_slot_free[2], s_tile_full, s_mma_done[2];
    __shared__ uint32_t s_tmem;
    constexpr int NPAD = 16 * NC16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kchunks = p.b1 / 4;                 // 16-byte K chunks per row
    const int ksteps = p.b1 / 8;                  // MMAs (K = 8) per split product
    const uint32_t a_sbo = 128u;                  // 8-row groups of the A tile are adjacent
    const uint32_t b_lbo = 128u;                  // B tile: consecutive K chunks of one 8-column group are adjacent
    const uint32_t b_sbo = (uint32_t)kchunks * 128u;
    // shared memory: [raw 0 | raw 1 | A_hi | A_lo | B_hi | B_lo]
    char *raw_base = smem;
    char *a_hi = smem + 2 * (size_t)p.raw_slot, *a_lo = a_hi + p.a_tile;
    char *b_hi = a_lo + p.a_tile, *b_lo = b_hi + p.b_tile;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            tc::mbar_init(tc::smem_u32(&s_full[i]), 1);
            tc::mbar_init(tc::smem_u32(&s_slot_free[i]), 1);
            tc::mbar_init(tc::smem_u32(&s_mma_done[i]), 1);
        }
        tc::mbar_init(tc::smem_u32(&s_tile_full), 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&s_tmem)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- B = M^T, K-major: element (n, k) = M[k][n]  ->  (n/8)*b_sbo + (n%8)*16 + (k/4)*b_lbo + (k%4)*4
    for (int idx = tid; idx < NPAD * kchunks; idx += blockDim.x) {
        const int nn = idx % NPAD, kc = idx / NPAD;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nn < p.b2) {
            const float *m = p.M + (size_t)(4 * kc) * p.ldm + nn;
            v = make_float4(__ldg(m), __ldg(m + p.ldm), __ldg(m + 2 * (size_t)p.ldm), __ldg(m + 3 * (size_t)p.ldm));
        }
        uint4 hi, lo;
        tc::split_tf32(v, hi, lo);
        const uint32_t off = (uint32_t)(nn >> 3) * b_sbo + (uint32_t)(nn & 7) * 16u + (uint32_t)kc * b_lbo;
        *(uint4 *)(b_hi + off) = hi;
        *(uint4 *)(b_lo + off) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    const int64_t tiles_total = (p.n + 127) / 128;
    const int nt = (int)((tiles_total - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles blockIdx.x, + grid, ...
    auto tile_row0 = [&](int t) { return ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * 128; };
    auto tile_rows = [&](int t) { const int64_t r0 = tile_row0(t); return (int)(p.n - r0 < 128 ? p.n - r0 : 128); };

    if (warp == 8) {
        // ================= producer + MMA issuer (one thread) =================
        if (lane == 0 && nt > 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) |
                                   ((uint32_t)(NPAD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            auto issue = [&](int t) {
                const int slot = t & 1;
                const uint32_t bytes = (uint32_t)tile_rows(t) * (uint32_t)p.b1 * 4u;
                const uint32_t bar = tc::smem_u32(&s_full[slot]);
                tc::mbar_expect_tx(bar, bytes);
                tc::bulk_g2s(tc::smem_u32(raw_base + (size_t)slot * p.raw_slot), p.Q + tile_row0(t) * p.b1, bytes, bar);
            };
            issue(0);
            if (nt > 1) issue(1);
            for (int t = 0; t < nt; t++) {
                tc::mbar_wait(tc::smem_u32(&s_tile_full), (uint32_t)t & 1);             // A tile of `t` is written
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem + (uint32_t)((t & 1) * NPAD);
                uint64_t dah = tc::make_desc(tc::smem_u32(a_hi), p.a_lbo, a_sbo), dal = tc::make_desc(tc::smem_u32(a_lo), p.a_lbo, a_sbo);
                uint64_t dbh = tc::make_desc(tc::smem_u32(b_hi), b_lbo, b_sbo), dbl = tc::make_desc(tc::smem_u32(b_lo), b_lbo, b_sbo);
                const uint64_t a_step = (uint64_t)((2u * p.a_lbo) >> 4), b_step = (uint64_t)((2u * b_lbo) >> 4);
                for (int ks = 0; ks < ksteps; ks++) {
                    tc::mma_tf32(d_tmem, dah, dbh, idesc, ks > 0 ? 1u : 0u);
                    tc::mma_tf32(d_tmem, dah, dbl, idesc, 1u);
                    tc::mma_tf32(d_tmem, dal, dbh, idesc, 1u);
                    dah += a_step; dal += a_step; dbh += b_step; dbl += b_step;
                }
                tc::commit(tc::smem_u32(&s_mma_done[t & 1]));
                if (t + 2 < nt) {                                                         // slot of `t`: output staged and stored
                    tc::mbar_wait(tc::smem_u32(&s_slot_free[t & 1]), (uint32_t)(t >> 1) & 1);
                    issue(t + 2);
                }
            }
        }
    } else {
        // ================= transform (warps 0-7) + epilogue (warps 0-3) =================
        auto epilogue = [&](int t) {   // accumulator of tile t -> staging (the tile's raw slot) -> one bulk store
            const int slot = t & 1;
            char *stage = raw_base + (size_t)slot * p.raw_slot;
            tc::mbar_wait(tc::smem_u32(&s_mma_done[slot]), (uint32_t)(t >> 1) & 1);
            if (warp < 4) {
                __syncwarp();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int row = warp * 32 + lane;
                float *dst = (float *)stage + (size_t)row * p.b2;
#pragma unroll
                for (int c = 0; c < NC16; c++) {
                    uint32_t r[16];
                    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(slot * NPAD + c * 16), r);
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        if (c * 16 + q * 4 < p.b2)
                            *(uint4 *)(dst + c * 16 + q * 4) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 128;" ::: "memory");                            // the 4 epilogue warps
                if (p.ldo == p.b2) {
                    if (tid == 0) {
                        tc::bulk_s2g(p.Out + tile_row0(t) * p.b2, tc::smem_u32(stage), (uint32_t)tile_rows(t) * (uint32_t)p.b2 * 4u);
                        tc::bulk_wait_read_all();                                         // staging slot may be overwritten
                        tc::mbar_arrive(tc::smem_u32(&s_slot_free[slot]));
                    }
                } else {
                    // strided output (the two halves of X = [U sqrt(S) | V sqrt(S)], ldo = d): the four epilogue warps copy
                    // the staged tile with coalesced 16-byte stores (128 per-row bulk stores measured 0.43 ms per launch
                    // against 0.22 ms for the contiguous form)
                    const int vr = tile_rows(t), q4 = p.b2 >> 2;
                    float *orow = p.Out + tile_row0(t) * p.ldo;
                    for (int idx = tid; idx < vr * q4; idx += 128) {
                        const int r = idx / q4, c4 = idx - r * q4;
                        *(float4 *)(orow + (size_t)r * p.ldo + 4 * c4) = *(const float4 *)(stage + ((size_t)r * p.b2 + 4 * c4) * 4);
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (tid == 0) tc::mbar_arrive(tc::smem_u32(&s_slot_free[slot]));
                }
            }
        };
        for (int t = 0; t < nt; t++) {
            const int slot = t & 1;
            const char *raw = raw_base + (size_t)slot * p.raw_slot;
            tc::mbar_wait(tc::smem_u32(&s_full[slot]), (uint32_t)(t >> 1) & 1);           // rows of tile t have landed
            if (t >= 1) tc::mbar_wait(tc::smem_u32(&s_mma_done[(t - 1) & 1]), (uint32_t)((t - 1) >> 1) & 1);  // A tile free
            // ---- raw rows -> K-major hi / lo tiles: lane -> chunk (conflict-free LDS), store with the padded LBO
            const int vr = tile_rows(t);
            for (int idx = tid; idx < 128 * kchunks; idx += 256) {
                const int m = idx / kchunks, kc = idx - m * kchunks;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < vr) v = *(const float4 *)(raw + ((size_t)m * p.b1 + 4 * kc) * 4);
                uint4 hi, lo;
                tc::split_tf32(v, hi, lo);
                const uint32_t off = (uint32_t)(m >> 3) * a_sbo + (uint32_t)(m & 7) * 16u + (uint32_t)kc * p.a_lbo;
                *(uint4 *)(a_hi + off) = hi;
                *(uint4 *)(a_lo + off) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tc::smem_u32(&s_tile_full));
            if (t >= 1) epilogue(t - 1);                                                  // overlaps the MMAs of tile t
        }
        if (nt > 0) epilogue(nt - 1);
        if (warp == 0) tc::bulk_wait_read_all();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
    // the bulk stores must be complete (not only read) before the kernel's results are consumed: kernel
    // completion guarantees it (bulk async-groups are flushed at exit of the issuing thread)
    if (warp == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int NC16>
static int apply_tc_launch_t(gemb_ctx *ctx, const ApplyTcParams &p, int grid, size_t smem_bytes) {
    static size_t attr_bytes = 0;
    if (attr_bytes < smem_bytes) {
        GEMB_CUDA(cudaFuncSetAttribute(apply_tc_kernel<NC16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        attr_bytes = smem_bytes;
    }
    apply_tc_kernel<NC16><<<grid, 288, smem_bytes, ctx->stream>>>(p);
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    return GEMB_OK;
}

// returns GEMB_ERR_UNSUPPORTED (without setting an error) when the shape does not fit this kernel
int apply_tc_launch(gemb_ctx *ctx, int64_t n, const float *Q, int b1, const float *M, int ldm, int b2, float *Out, int ldo) {
    if (b1 % 8 || b2 % 4 || b1 > 128 || b2 > 128 || ldo < b2 || ldo % 4 || ((uintptr_t)Out & 15) || n <= 0) return GEMB_ERR_UNSUPPORTED;
    ApplyTcParams p;
    p.n = n; p.Q = Q; p.M = M; p.Out = Out; p.b1 = b1; p.b2 = b2; p.ldm = ldm; p.ldo = ldo;
    const int widths[5] = {32, 64, 80, 96, 128};
    int npad = 128;
    for (int w : widths) if (w >= b2) { npad = w; break; }
    p.tmem_cols = 32;
    while (p.tmem_cols < (uint32_t)(2 * npad)) p.tmem_cols <<= 1;
    p.a_lbo = 2048u + 16u;
    p.a_tile = (uint32_t)(b1 / 4) * p.a_lbo;
    p.b_tile = (uint32_t)(npad / 8) * (uint32_t)(b1 / 4) * 128u;
    const size_t raw = (size_t)128 * (size_t)std::max(b1, b2) * 4;
    p.raw_slot = (uint32_t)((raw + 127) / 128 * 128);
    const size_t smem_bytes = 2 * (size_t)p.raw_slot + 2 * (size_t)p.a_tile + 2 * (size_t)p.b_tile + 256;
    if (smem_bytes > 226 * 1024) return GEMB_ERR_UNSUPPORTED;
    const int64_t tiles = (n + 127) / 128;
    int grid = ctx->sm_count;
    if (grid > tiles) grid = (int)tiles;
    switch (npad) {
        case 32: return apply_tc_launch_t<2>(ctx, p, grid, smem_bytes);
        case 64: return apply_tc_launch_t<4>(ctx, p, grid, smem_bytes);
        case 80: return apply_tc_launch_t<5>(ctx, p, grid, smem_bytes);
        case 96: return apply_tc_launch_t<6>(ctx, p, grid, smem_bytes);
        default: return apply_tc_launch_t<8>(ctx, p, grid, smem_bytes);
    }
}

}  // namespace gemb
