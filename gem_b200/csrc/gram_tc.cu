// gem_b200/csrc/gram_tc.cu -- the tall-skinny Gram contraction G = P^T Q on the 5th-gen tensor cores.
//
// This is the ONE dense contraction of the HOPE solver (CholeskyQR Gram and the Rayleigh-Ritz
// projection; replaces numpy.linalg.qr / svd inside scipy svds, _svds.py:508-533).  It is memory bound
// (read n*b fp32 once), so the kernel is a persistent streaming design, one CTA per SM:
//
//   loader (all 8 warps)  : coalesced LDG.128 of 8-row x 64-byte pieces of the fp32 row-major block,
//                           split x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi) in registers,
//                           STS.128 into the UMMA canonical MN-major / no-swizzle layout
//                           (16-byte chunk (k, j) -> j*SBO + (k/8)*LBO + (k%8)*16); two stages
//   MMA issuer (1 thread) : per 8-row k-block three tcgen05.mma.kind::tf32 (hi*hi + hi*lo + lo*hi =
//                           "3xTF32", ~fp32 accuracy), M = 128, N = pad16(b2), fp32 accumulators in TMEM
//   epilogue (warps 0-3)  : tcgen05.ld of the accumulator, fp64 atomicAdd into G across CTAs
// tcgen05.commit -> mbarrier hands a stage back to the loaders.  A = P^T and B = Q are both "MN-major"
// (the contraction index is the row index of the row-major blocks), which kind::tf32 supports.
#include "common.cuh"

namespace gemb {

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded spin: a descriptor / protocol bug must surface as an error, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 28); it++)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// UMMA shared-memory descriptor, SWIZZLE_NONE, Blackwell version field = 1
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}

}  // namespace tc

struct GramTcParams {
    int64_t n;
    const float *P, *Q;
    int b1, b2;
    double *G;
    int stage_rows;      // multiple of 8
    int n_pad;           // pad16(b2)
    int m_tiles;         // ceil(b1 / 128)
    uint32_t tile_bytes_p, tile_bytes_q;   // bytes of one (hi or lo) tile
    uint32_t tmem_cols;  // power of two >= m_tiles * n_pad
};

// one 8-row x 4-chunk unit: lane -> row k = kb*8 + lane%8, chunk j = jq*4 + lane/8
__device__ __forceinline__ void load_split_store(const float *__restrict__ X, int64_t n, int b, int chunks,
                                                 int64_t row0, int kb, int jq, int lane, char *tile_hi,
                                                 char *tile_lo, uint32_t sbo) {
    const int k = kb * 8 + (lane & 7);
    const int j = jq * 4 + (lane >> 3);
    if (j >= chunks) return;
    const int64_t r = row0 + k;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < n) v = __ldg((const float4 *)(X + r * b) + j);
    uint4 hi, lo;
    hi.x = tc::to_tf32(v.x); hi.y = tc::to_tf32(v.y); hi.z = tc::to_tf32(v.z); hi.w = tc::to_tf32(v.w);
    lo.x = tc::to_tf32(v.x - __uint_as_float(hi.x));
    lo.y = tc::to_tf32(v.y - __uint_as_float(hi.y));
    lo.z = tc::to_tf32(v.z - __uint_as_float(hi.z));
    lo.w = tc::to_tf32(v.w - __uint_as_float(hi.w));
    const uint32_t off = (uint32_t)j * sbo + (uint32_t)kb * 128u + (uint32_t)(lane & 7) * 16u;
    *(uint4 *)(tile_hi + off) = hi;
    *(uint4 *)(tile_lo + off) = lo;
}

template <bool CROSS>
__global__ void __launch_bounds__(256, 1) gram_tc_kernel(GramTcParams p) {
    extern __shared__ __align__(128) char smem[];
    __shared__ __align__(8) uint64_t s_bar[3];   // [0],[1]: stage consumed ; [2]: all MMAs done
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kblocks = p.stage_rows / 8;
    const uint32_t lbo = 128u;
    const uint32_t sbo = (uint32_t)kblocks * 128u;
    const int chunks_p = p.b1 / 4, chunks_q = p.b2 / 4;
    // stage layout: [P_hi | P_lo | (Q_hi | Q_lo)]
    const uint32_t stage_bytes = 2 * p.tile_bytes_p + (CROSS ? 2 * p.tile_bytes_q : 0);

    if (tid == 0) {
        tc::mbar_init(tc::smem_u32(&s_bar[0]), 1);
        tc::mbar_init(tc::smem_u32(&s_bar[1]), 1);
        tc::mbar_init(tc::smem_u32(&s_bar[2]), 1);
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

    // instruction descriptor: D = F32, A = B = TF32, both MN-major, M = 128, N = n_pad
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    // contiguous row range of this CTA, in whole stages
    const int64_t stages_total = (p.n + p.stage_rows - 1) / p.stage_rows;
    const int64_t per_cta = (stages_total + gridDim.x - 1) / gridDim.x;
    const int64_t s_begin = (int64_t)blockIdx.x * per_cta;
    const int64_t s_end = s_begin + per_cta < stages_total ? s_begin + per_cta : stages_total;

    uint32_t fills[2] = {0, 0};
    bool any = false;
    for (int64_t s = s_begin; s < s_end; s++) {
        const int st = (int)((s - s_begin) & 1);
        char *base = smem + (size_t)st * stage_bytes;
        if (fills[st] > 0) tc::mbar_wait(tc::smem_u32(&s_bar[st]), (fills[st] - 1) & 1);   // MMAs of the previous use are done
        const int64_t row0 = s * p.stage_rows;
        // ---- load + split
        const int units_p = kblocks * ((chunks_p + 3) / 4);
        for (int u = warp; u < units_p; u += 8)
            load_split_store(p.P, p.n, p.b1, chunks_p, row0, u % kblocks, u / kblocks, lane, base, base + p.tile_bytes_p, sbo);
        if (CROSS) {
            const int units_q = kblocks * ((chunks_q + 3) / 4);
            char *qb = base + 2 * p.tile_bytes_p;
            for (int u = warp; u < units_q; u += 8)
                load_split_store(p.Q, p.n, p.b2, chunks_q, row0, u % kblocks, u / kblocks, lane, qb, qb + p.tile_bytes_q, sbo);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        fills[st]++;
        // ---- MMA issue (one thread)
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = tc::smem_u32(base), a_lo = a_hi + p.tile_bytes_p;
            const uint32_t b_hi = CROSS ? a_hi + 2 * p.tile_bytes_p : a_hi;
            const uint32_t b_lo = CROSS ? b_hi + p.tile_bytes_q : a_lo;
            for (int mt = 0; mt < p.m_tiles; mt++) {
                const uint32_t d_tmem = tmem + (uint32_t)(mt * p.n_pad);
                const uint32_t a_off = (uint32_t)mt * 32u * sbo;          // 32 chunks = 128 rows of A^T per M tile
                for (int kb = 0; kb < kblocks; kb++) {
                    const uint32_t ko = (uint32_t)kb * lbo;
                    const uint64_t dah = tc::make_desc(a_hi + a_off + ko, lbo, sbo);
                    const uint64_t dal = tc::make_desc(a_lo + a_off + ko, lbo, sbo);
                    const uint64_t dbh = tc::make_desc(b_hi + ko, lbo, sbo);
                    const uint64_t dbl = tc::make_desc(b_lo + ko, lbo, sbo);
                    const uint32_t acc = (any || kb > 0) ? 1u : 0u;
                    tc::mma_tf32(d_tmem, dah, dbh, idesc, acc);
                    tc::mma_tf32(d_tmem, dah, dbl, idesc, 1u);
                    tc::mma_tf32(d_tmem, dal, dbh, idesc, 1u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(&s_bar[st])) : "memory");
        }
        any = true;
    }
    // ---- epilogue
    if (any) {
        if (tid == 0)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(&s_bar[2])) : "memory");
        if (warp < 4) {
            tc::mbar_wait(tc::smem_u32(&s_bar[2]), 0);
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int mt = 0; mt < p.m_tiles; mt++) {
                const int row = mt * 128 + warp * 32 + lane;              // TMEM lane = accumulator row
                for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
                    uint32_t r[16];
                    __syncwarp();
                    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * p.n_pad + c0);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (row < p.b1) {
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const int col = c0 + i;
                            if (col < p.b2) atomicAdd(p.G + (size_t)row * p.b2 + col, (double)__uint_as_float(r[i]));
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
    }
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

// returns GEMB_ERR_UNSUPPORTED (without setting an error) when the shape does not fit this kernel
int gram_tc_launch(gemb_ctx *ctx, int64_t n, const float *P, int b1, const float *Q, int b2, double *G) {
    const bool cross = (P != Q);
    if (b1 % 4 || b2 % 4 || b1 > 256 || b2 > 256 || n <= 0) return GEMB_ERR_UNSUPPORTED;
    GramTcParams p;
    p.n = n; p.P = P; p.Q = Q; p.b1 = b1; p.b2 = b2; p.G = G;
    p.n_pad = (b2 + 15) / 16 * 16;
    p.m_tiles = (b1 + 127) / 128;
    const uint32_t cols = (uint32_t)(p.m_tiles * p.n_pad);
    if (cols > 512) return GEMB_ERR_UNSUPPORTED;
    p.tmem_cols = 32;
    while (p.tmem_cols < cols) p.tmem_cols <<= 1;
    // stage rows: largest multiple of 8 (<= 64) with two stages in <= 192 KB
    const size_t bytes_per_row = (size_t)4 * 2 * (b1 + (cross ? b2 : 0));    // hi + lo
    int rows = (int)((192 * 1024) / (2 * bytes_per_row)) / 8 * 8;
    if (rows > 64) rows = 64;
    if (rows < 8) return GEMB_ERR_UNSUPPORTED;
    p.stage_rows = rows;
    p.tile_bytes_p = (uint32_t)(b1 / 4) * (uint32_t)(rows / 8) * 128u;
    p.tile_bytes_q = (uint32_t)(b2 / 4) * (uint32_t)(rows / 8) * 128u;
    const size_t stage_bytes = 2 * (size_t)p.tile_bytes_p + (cross ? 2 * (size_t)p.tile_bytes_q : 0);
    // the MMA reads 32 chunks per M tile and n_pad/4 chunks of B even where the block is narrower:
    // keep those (ignored) reads inside the allocation
    const size_t sbo = (size_t)(rows / 8) * 128;
    const size_t over = std::max<size_t>((size_t)p.m_tiles * 32 * sbo, (size_t)(p.n_pad / 4) * sbo);
    const size_t smem_bytes = 2 * stage_bytes + over + 1024;
    if (smem_bytes > 227 * 1024) return GEMB_ERR_UNSUPPORTED;
    static bool attr_set = false;
    if (!attr_set) {
        GEMB_CUDA(cudaFuncSetAttribute(gram_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        GEMB_CUDA(cudaFuncSetAttribute(gram_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    GEMB_CUDA(cudaMemsetAsync(G, 0, sizeof(double) * (size_t)b1 * b2, ctx->stream));
    const int64_t stages_total = (n + rows - 1) / rows;
    int grid = ctx->sm_count;
    if (grid > stages_total) grid = (int)stages_total;
    if (cross) gram_tc_kernel<true><<<grid, 256, smem_bytes, ctx->stream>>>(p);
    else gram_tc_kernel<false><<<grid, 256, smem_bytes, ctx->stream>>>(p);
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    return GEMB_OK;
}

}  // namespace gemb
