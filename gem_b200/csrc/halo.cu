// gem_b200/csrc/halo.cu -- multi-GPU HOPE without a per-sweep collective (replaces the ncclAllGather of the whole
// n x b block in front of every SpMM that round 1 used; the reference itself is single-process).
//
// Row-sharded symmetric A, one process per GPU.  At set-up every rank works out
//   * H_p: the sorted distinct REMOTE columns its CSR shard references (the "halo"; an SBM shard at P = 8 needs
//     ~2.75 M of the 7 M remote rows, an all-gather delivers all 7 M), and a second copy of the column ids in which a
//     local column c becomes c - row0 and a remote one n_shard + (its position in H_p);
//   * from the all-gathered H_q: for each of its own rows the list of (peer q, slot in H_q) that reference it.
// Work blocks are laid out [n_shard local rows | halo rows] and every rank maps every peer's blocks and barrier flags
// with CUDA IPC.  The kernel that produces a block writes each local row into the peers' halo slots with plain 16-byte
// stores over NVLink (posted writes: fire-and-forget, they overlap the producer's own gathers and FMAs tile by tile),
// a flag barrier closes the sweep, and the next SpMM gathers from local HBM only.
#include "common.cuh"
#include "nccl_api.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <vector>

namespace gemb {

#define NCCL_TRY(call, what)                                                                   \
    do {                                                                                       \
        ncclResult_t _r = (call);                                                              \
        if (_r != ncclSuccess) { set_error("%s: %s", what, api->GetErrorString(_r)); return GEMB_ERR_NCCL; } \
    } while (0)

struct IsRemote {
    int32_t lo, hi;   // local columns are [lo, hi)
    __host__ __device__ bool operator()(const int32_t &c) const { return c < lo || c >= hi; }
};

__device__ __forceinline__ int64_t lower_bound_i32(const int32_t *a, int64_t n, int32_t x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t m = (lo + hi) >> 1;
        if (a[m] < x) lo = m + 1; else hi = m;
    }
    return lo;
}

__global__ void halo_remap_kernel(int64_t nnz, const int32_t *__restrict__ idx, int32_t lo, int32_t hi,
                                  const int32_t *__restrict__ H, int64_t nH, int32_t n_shard,
                                  int32_t *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = idx[i];
        out[i] = (c >= lo && c < hi) ? c - lo : n_shard + (int32_t)lower_bound_i32(H, nH, c);
    }
}

// seg[2q], seg[2q+1]: the slots of H_q that hold rows of [lo, hi)
__global__ void halo_segments_kernel(int nranks, const int32_t *__restrict__ Hall, int64_t maxH,
                                     const long long *__restrict__ counts, int32_t lo, int32_t hi,
                                     long long *__restrict__ seg) {
    const int q = threadIdx.x;
    if (q >= nranks) return;
    const int32_t *H = Hall + (int64_t)q * maxH;
    seg[2 * q] = lower_bound_i32(H, counts[q], lo);
    seg[2 * q + 1] = lower_bound_i32(H, counts[q], hi);
}

// pass 0: cnt[row]++ ; pass 1: push_dst[push_ptr[row] + cursor[row]++] = (q << 29) | slot
template <int PASS>
__global__ void halo_pushlist_kernel(int q, const int32_t *__restrict__ H, long long s0, long long s1, int32_t row0,
                                     int32_t *__restrict__ cnt, const int32_t *__restrict__ push_ptr,
                                     uint32_t *__restrict__ push_dst) {
    for (long long t = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < s1; t += (long long)gridDim.x * blockDim.x) {
        const int32_t row = H[t] - row0;
        const int pos = atomicAdd(cnt + row, 1);
        if (PASS == 1) push_dst[push_ptr[row] + pos] = ((uint32_t)q << 29) | (uint32_t)t;
    }
}

static int ipc_exchange(gemb_ctx *c, void *const *mine, int count, void **peers /* [count][GEMB_MAX_RANKS] */) {
    NcclApi *api = nccl_api();
    if (!api) return GEMB_ERR_NCCL;
    const int P = c->nranks;
    std::vector<cudaIpcMemHandle_t> h(count), all((size_t)count * P);
    for (int i = 0; i < count; i++) GEMB_CUDA(cudaIpcGetMemHandle(&h[i], mine[i]));
    char *dsend = nullptr, *drecv = nullptr;
    const size_t bytes = sizeof(cudaIpcMemHandle_t) * count;
    GEMB_CUDA(dmalloc(&dsend, bytes));
    GEMB_CUDA(dmalloc(&drecv, bytes * P));
    GEMB_CUDA(cudaMemcpyAsync(dsend, h.data(), bytes, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(api->AllGather(dsend, drecv, bytes, ncclChar, (ncclComm_t)c->comm, c->stream), "ncclAllGather(ipc handles)");
    GEMB_CUDA(cudaMemcpyAsync(all.data(), drecv, bytes * P, cudaMemcpyDeviceToHost, c->stream));
    GEMB_CUDA(cudaStreamSynchronize(c->stream));
    dfree(dsend); dfree(drecv);
    for (int i = 0; i < count; i++)
        for (int q = 0; q < P; q++) {
            void **slot = peers + (size_t)i * GEMB_MAX_RANKS + q;
            if (q == c->rank) { *slot = mine[i]; continue; }
            cudaError_t e = cudaIpcOpenMemHandle(slot, all[(size_t)q * count + i], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                (void)cudaGetLastError();
                set_error("cudaIpcOpenMemHandle (rank %d -> rank %d): %s", c->rank, q, cudaGetErrorString(e));
                return GEMB_ERR_CUDA;
            }
        }
    return GEMB_OK;
}

static int nccl_barrier(gemb_ctx *c) {
    NcclApi *api = nccl_api();
    if (!api) return GEMB_ERR_NCCL;
    int *d = nullptr;
    GEMB_CUDA(dmalloc(&d, sizeof(int)));
    GEMB_CUDA(cudaMemsetAsync(d, 0, sizeof(int), c->stream));
    NCCL_TRY(api->AllReduce(d, d, 1, ncclInt, ncclSum, (ncclComm_t)c->comm, c->stream), "ncclAllReduce(barrier)");
    GEMB_CUDA(cudaStreamSynchronize(c->stream));
    dfree(d);
    return GEMB_OK;
}

int halo_build(gemb_graph *g) {
    gemb_halo &H = g->halo;
    if (H.ready) return GEMB_OK;
    gemb_ctx *c = g->ctx;
    NcclApi *api = nccl_api();
    if (!api) return GEMB_ERR_NCCL;
    GEMB_ARG(c->nranks > 1 && c->nranks <= GEMB_MAX_RANKS, "halo exchange supports 2..8 ranks");
    GEMB_ARG(g->symmetric && !g->replicated, "halo exchange needs a symmetric row shard");
    GEMB_ARG(g->n_shard + 1 < (int64_t)1 << 29, "shard too large for 29-bit halo slots");
    const int P = c->nranks;
    const int64_t nnz = g->A.nnz;
    const int32_t lo = (int32_t)g->row0, hi = (int32_t)std::min<int64_t>(g->row0 + g->n_shard, g->n);
    cudaStream_t st = c->stream;

    // ---- distinct remote columns, sorted
    int32_t *rem = nullptr, *rem_sorted = nullptr, *Hd = nullptr;
    long long *d_num = nullptr;
    GEMB_CUDA(dmalloc(&rem, sizeof(int32_t) * std::max<int64_t>(nnz, 1)));
    GEMB_CUDA(dmalloc(&rem_sorted, sizeof(int32_t) * std::max<int64_t>(nnz, 1)));
    GEMB_CUDA(dmalloc(&Hd, sizeof(int32_t) * std::max<int64_t>(nnz, 1)));
    GEMB_CUDA(dmalloc(&d_num, sizeof(long long) * (2 * GEMB_MAX_RANKS + 2)));
    size_t tb = 0, need = 0;
    void *tmp = nullptr;
    IsRemote pred{lo, hi};
    cub::DeviceSelect::If(nullptr, need, g->A.indices, rem, d_num, nnz, pred, st); tb = need;
    cub::DeviceRadixSort::SortKeys(nullptr, need, rem, rem_sorted, nnz, 0, 32, st); tb = std::max(tb, need);
    cub::DeviceSelect::Unique(nullptr, need, rem_sorted, Hd, d_num, nnz, st); tb = std::max(tb, need);
    GEMB_CUDA(dmalloc(&tmp, tb ? tb : 4));
    long long n_rem = 0, n_H = 0;
    if (nnz > 0) {
        GEMB_CUDA(cub::DeviceSelect::If(tmp, tb, g->A.indices, rem, d_num, nnz, pred, st));
        GEMB_CUDA(cudaMemcpyAsync(&n_rem, d_num, sizeof n_rem, cudaMemcpyDeviceToHost, st));
        GEMB_CUDA(cudaStreamSynchronize(st));
        if (n_rem > 0) {
            GEMB_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tb, rem, rem_sorted, n_rem, 0, 32, st));
            GEMB_CUDA(cub::DeviceSelect::Unique(tmp, tb, rem_sorted, Hd, d_num, n_rem, st));
            GEMB_CUDA(cudaMemcpyAsync(&n_H, d_num, sizeof n_H, cudaMemcpyDeviceToHost, st));
            GEMB_CUDA(cudaStreamSynchronize(st));
        }
    }
    count_launch(3);
    H.halo_rows = n_H;
    GEMB_ARG(n_H < ((int64_t)1 << 29), "halo too large for 29-bit slots");

    // ---- remapped column ids
    GEMB_CUDA(dmalloc(&H.indices_ext, sizeof(int32_t) * (std::max<int64_t>(nnz, 1) + 4)));   // + the x4 padding the bulk copies of spmm.cu read
    if (nnz > 0) {
        halo_remap_kernel<<<c->sm_count * 8, 256, 0, st>>>(nnz, g->A.indices, lo, hi, Hd, n_H, (int32_t)g->n_shard, H.indices_ext);
        GEMB_CUDA(cudaGetLastError());
        count_launch();
    }

    // ---- everyone's halo lists -> who needs my rows
    long long *d_cnt_all = d_num + 2;                       // [P]
    long long h_cnt_all[GEMB_MAX_RANKS];
    GEMB_CUDA(cudaMemcpyAsync(d_num, &n_H, sizeof n_H, cudaMemcpyHostToDevice, st));
    NCCL_TRY(api->AllGather(d_num, d_cnt_all, 1, ncclInt64, (ncclComm_t)c->comm, st), "ncclAllGather(halo counts)");
    GEMB_CUDA(cudaMemcpyAsync(h_cnt_all, d_cnt_all, sizeof(long long) * P, cudaMemcpyDeviceToHost, st));
    GEMB_CUDA(cudaStreamSynchronize(st));
    long long maxH = 1;
    for (int q = 0; q < P; q++) maxH = std::max(maxH, h_cnt_all[q]);
    int32_t *Hpad = nullptr, *Hall = nullptr;
    GEMB_CUDA(dmalloc(&Hpad, sizeof(int32_t) * maxH));
    GEMB_CUDA(dmalloc(&Hall, sizeof(int32_t) * maxH * P));
    GEMB_CUDA(cudaMemsetAsync(Hpad, 0x7f, sizeof(int32_t) * maxH, st));
    if (n_H) GEMB_CUDA(cudaMemcpyAsync(Hpad, Hd, sizeof(int32_t) * n_H, cudaMemcpyDeviceToDevice, st));
    NCCL_TRY(api->AllGather(Hpad, Hall, (size_t)maxH, ncclInt32, (ncclComm_t)c->comm, st), "ncclAllGather(halo lists)");
    long long *d_seg = d_cnt_all + GEMB_MAX_RANKS;          // needs 2P entries: allocate separately
    long long *seg_dev = nullptr;
    GEMB_CUDA(dmalloc(&seg_dev, sizeof(long long) * 2 * GEMB_MAX_RANKS));
    (void)d_seg;
    halo_segments_kernel<<<1, 32, 0, st>>>(P, Hall, maxH, d_cnt_all, lo, hi, seg_dev);
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    long long seg[2 * GEMB_MAX_RANKS];
    GEMB_CUDA(cudaMemcpyAsync(seg, seg_dev, sizeof(long long) * 2 * P, cudaMemcpyDeviceToHost, st));
    GEMB_CUDA(cudaStreamSynchronize(st));

    int32_t *cnt = nullptr;
    const int64_t nl = g->n_local;
    GEMB_CUDA(dmalloc(&cnt, sizeof(int32_t) * (nl + 1)));
    GEMB_CUDA(dmalloc(&H.push_ptr, sizeof(int32_t) * (nl + 1)));
    GEMB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (nl + 1), st));
    long long total = 0;
    for (int q = 0; q < P; q++) {
        if (q == c->rank) continue;
        const long long m = seg[2 * q + 1] - seg[2 * q];
        total += m;
        if (m <= 0) continue;
        const int grid = (int)std::min<long long>((m + 255) / 256, c->sm_count * 8);
        halo_pushlist_kernel<0><<<grid, 256, 0, st>>>(q, Hall + (int64_t)q * maxH, seg[2 * q], seg[2 * q + 1], lo, cnt, nullptr, nullptr);
        GEMB_CUDA(cudaGetLastError());
        count_launch();
    }
    GEMB_ARG(total < ((long long)1 << 31), "push list too long");
    H.push_total = total;
    size_t sb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, sb, cnt, H.push_ptr, (int)(nl + 1), st);
    void *stmp = nullptr;
    GEMB_CUDA(dmalloc(&stmp, sb ? sb : 4));
    GEMB_CUDA(cub::DeviceScan::ExclusiveSum(stmp, sb, cnt, H.push_ptr, (int)(nl + 1), st));
    GEMB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (nl + 1), st));
    GEMB_CUDA(dmalloc(&H.push_dst, sizeof(uint32_t) * std::max<long long>(total, 1)));
    for (int q = 0; q < P; q++) {
        if (q == c->rank) continue;
        const long long m = seg[2 * q + 1] - seg[2 * q];
        if (m <= 0) continue;
        const int grid = (int)std::min<long long>((m + 255) / 256, c->sm_count * 8);
        halo_pushlist_kernel<1><<<grid, 256, 0, st>>>(q, Hall + (int64_t)q * maxH, seg[2 * q], seg[2 * q + 1], lo, cnt, H.push_ptr, H.push_dst);
        GEMB_CUDA(cudaGetLastError());
        count_launch();
    }
    GEMB_CUDA(cudaStreamSynchronize(st));
    dfree(rem); dfree(rem_sorted); dfree(Hd); dfree(d_num); dfree(tmp); dfree(Hpad); dfree(Hall); dfree(seg_dev);
    dfree(cnt); dfree(stmp);

    // ---- barrier flags: owned by the context (plain cudaMalloc: exported through CUDA IPC, never recycled by the block
    //      cache), set up once and shared by every graph of this context
    gemb_halo_pool &PL = c->halo_pool;
    if (!PL.flags) {
        GEMB_CUDA(cudaMalloc(&PL.flags, sizeof(unsigned long long) * GEMB_MAX_RANKS));
        GEMB_CUDA(cudaMemset(PL.flags, 0, sizeof(unsigned long long) * GEMB_MAX_RANKS));
        GEMB_CUDA(cudaMalloc(&PL.timeout_flag, sizeof(int)));
        GEMB_CUDA(cudaMemset(PL.timeout_flag, 0, sizeof(int)));
        void *mine[1] = {PL.flags};
        void *peers[GEMB_MAX_RANKS] = {};
        GEMB_TRY(ipc_exchange(c, mine, 1, peers));
        for (int q = 0; q < P; q++) PL.peer_flags[q] = (unsigned long long *)peers[q];
        PL.epoch = 0;
    }
    H.flags = PL.flags;
    H.timeout_flag = PL.timeout_flag;
    for (int q = 0; q < P; q++) H.peer_flags[q] = PL.peer_flags[q];
    H.ready = true;
    return GEMB_OK;
}

// collective: every rank drops its mappings of the peers' blocks, then frees its own
static int halo_pool_drop_blocks(gemb_ctx *c) {
    gemb_halo_pool &PL = c->halo_pool;
    if (PL.nbuf == 0) return GEMB_OK;
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < PL.nbuf; i++)
        for (int q = 0; q < c->nranks; q++)
            if (q != c->rank && PL.peer_buf[i][q]) { cudaIpcCloseMemHandle(PL.peer_buf[i][q]); PL.peer_buf[i][q] = nullptr; }
    GEMB_TRY(nccl_barrier(c));     // nobody still maps what is freed next
    for (int i = 0; i < PL.nbuf; i++) { cudaFree(PL.buf[i]); PL.buf[i] = nullptr; }
    PL.nbuf = 0; PL.cap_floats = 0;
    return GEMB_OK;
}

int halo_buffers(gemb_graph *g, int nbuf, int width) {
    gemb_halo &H = g->halo;
    gemb_ctx *c = g->ctx;
    gemb_halo_pool &PL = c->halo_pool;
    NcclApi *api = nccl_api();
    if (!api) return GEMB_ERR_NCCL;
    GEMB_ARG(H.ready && nbuf >= 1 && nbuf <= GEMB_HALO_BUFS, "halo_buffers");
    const size_t rows = (size_t)(g->n_shard + H.halo_rows);
    // the pool is (re)built only when some rank needs more than it holds: the decision comes from an all-reduce (max)
    // of the need, so every rank takes the same branch
    long long need = (long long)(rows * (size_t)width), *d_need = nullptr;
    GEMB_CUDA(dmalloc(&d_need, sizeof(long long)));
    GEMB_CUDA(cudaMemcpyAsync(d_need, &need, sizeof need, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(api->AllReduce(d_need, d_need, 1, ncclInt64, ncclMax, (ncclComm_t)c->comm, c->stream), "ncclAllReduce(halo block size)");
    GEMB_CUDA(cudaMemcpyAsync(&need, d_need, sizeof need, cudaMemcpyDeviceToHost, c->stream));
    GEMB_CUDA(cudaStreamSynchronize(c->stream));
    dfree(d_need);
    if (PL.nbuf < nbuf || PL.cap_floats < (size_t)need) {
        GEMB_TRY(halo_pool_drop_blocks(c));
        const size_t cap = (size_t)std::max<long long>(need, 1);
        void *mine[GEMB_HALO_BUFS];
        for (int i = 0; i < nbuf; i++) {
            cudaError_t e = cudaMalloc(&PL.buf[i], sizeof(float) * cap);
            if (e != cudaSuccess) {
                (void)cudaGetLastError();
                gemb_mem_trim();
                e = cudaMalloc(&PL.buf[i], sizeof(float) * cap);
            }
            GEMB_CUDA(e);
            mine[i] = PL.buf[i];
        }
        void *peers[GEMB_HALO_BUFS * GEMB_MAX_RANKS] = {};
        PL.nbuf = nbuf; PL.cap_floats = cap;
        GEMB_TRY(ipc_exchange(c, mine, nbuf, peers));
        for (int i = 0; i < nbuf; i++)
            for (int q = 0; q < c->nranks; q++) PL.peer_buf[i][q] = (float *)peers[(size_t)i * GEMB_MAX_RANKS + q];
    } else {
        // blocks of the previous call are being reused: no rank may still be pushing into them
        GEMB_TRY(nccl_barrier(c));
    }
    for (int i = 0; i < nbuf; i++) {
        GEMB_CUDA(cudaMemsetAsync(PL.buf[i], 0, sizeof(float) * rows * width, c->stream));
        H.buf[i] = PL.buf[i];
        for (int q = 0; q < c->nranks; q++) H.peer_buf[i][q] = PL.peer_buf[i][q];
    }
    H.nbuf = nbuf; H.width = width;
    // the zero fill must be complete everywhere before any peer's first push can land
    GEMB_CUDA(cudaStreamSynchronize(c->stream));
    GEMB_TRY(nccl_barrier(c));
    return GEMB_OK;
}

void halo_push_args(const gemb_graph *g, int bi, HaloPushArgs *out, bool half) {
    const gemb_halo &H = g->halo;
    out->push_ptr = H.push_ptr;
    out->push_dst = H.push_dst;
    for (int q = 0; q < GEMB_MAX_RANKS; q++) out->peer[q] = (float4 *)H.peer_buf[bi][q];
    out->halo_row0 = g->n_shard;
    out->half = half ? 1 : 0;
}

// group of G threads per local row: copy the row into every peer slot that references it
__global__ void __launch_bounds__(256)
halo_push_kernel(int64_t n_rows, int G, int rows_per_cta, const float4 *__restrict__ Y, HaloPushArgs P) {
    const int lr = threadIdx.x / G, c = threadIdx.x - lr * G;
    if (lr >= rows_per_cta) return;
    const int64_t row = (int64_t)blockIdx.x * rows_per_cta + lr;
    if (row >= n_rows) return;
    if (P.push_ptr[row] == P.push_ptr[row + 1]) return;
    halo_push_row(P, row, G, c, Y[row * G + c]);
}

int halo_push_launch(gemb_graph *g, int bi, int width, bool half) {
    gemb_ctx *c = g->ctx;
    if (g->n_local == 0 || g->halo.push_total == 0) return GEMB_OK;
    const int G = width / 4;
    GEMB_ARG(G >= 1 && G <= 256, "width");
    const int rpc = 256 / G;
    HaloPushArgs P;
    halo_push_args(g, bi, &P, half);
    halo_push_kernel<<<(unsigned)((g->n_local + rpc - 1) / rpc), 256, 0, c->stream>>>(g->n_local, G, rpc, (const float4 *)g->halo.buf[bi], P);
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    return GEMB_OK;
}

struct BarrierArgs {
    unsigned long long *peer[GEMB_MAX_RANKS];
};

// One warp: lane q posts this rank's epoch into peer q's flag word, then waits for peer q's.  Stream order puts this
// kernel after the producer kernel, whose (peer) stores are performed before the kernel completes; the fences keep the
// flag behind them.  The wait is bounded (~2 s): a lost rank must surface as an error, never as a hung GPU.
__global__ void halo_barrier_kernel(BarrierArgs A, volatile unsigned long long *mine, unsigned long long epoch, int rank,
                                    int nranks, int *timeout_flag) {
    const int q = threadIdx.x;
    if (q >= nranks || q == rank) return;
    __threadfence_system();
    *(volatile unsigned long long *)(A.peer[q] + rank) = epoch;
    __threadfence_system();
    const long long t0 = clock64();
    while (mine[q] < epoch) {
        if (clock64() - t0 > 4000000000LL) { atomicExch(timeout_flag, 1); break; }
        __nanosleep(100);
    }
    __threadfence_system();
}

int halo_barrier(gemb_graph *g) {
    gemb_halo &H = g->halo;
    gemb_ctx *c = g->ctx;
    BarrierArgs A;
    for (int q = 0; q < GEMB_MAX_RANKS; q++) A.peer[q] = H.peer_flags[q];
    c->halo_pool.epoch++;
    halo_barrier_kernel<<<1, 32, 0, c->stream>>>(A, H.flags, c->halo_pool.epoch, c->rank, c->nranks, H.timeout_flag);
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    return GEMB_OK;
}

int halo_check_timeout(gemb_graph *g) {
    gemb_halo &H = g->halo;
    if (!H.ready) return GEMB_OK;
    int h = 0;
    GEMB_CUDA(cudaMemcpyAsync(&h, H.timeout_flag, sizeof h, cudaMemcpyDeviceToHost, g->ctx->stream));
    GEMB_CUDA(cudaStreamSynchronize(g->ctx->stream));
    if (h) {
        set_error("multi-GPU HOPE: a peer did not reach the sweep barrier within 2 s (rank %d of %d)", g->ctx->rank, g->ctx->nranks);
        return GEMB_ERR_NCCL;
    }
    return GEMB_OK;
}

int halo_free(gemb_graph *g) {
    gemb_halo &H = g->halo;
    if (!H.ready) return GEMB_OK;
    cudaStreamSynchronize(g->ctx->stream);
    dfree(H.indices_ext); dfree(H.push_ptr); dfree(H.push_dst);     // the plan; blocks and flags belong to the context
    H = gemb_halo();
    return GEMB_OK;
}

// context destruction: not collective (the peers may already be gone); a peer that still maps these blocks keeps its
// mapping valid until it closes it or exits
void halo_pool_release(gemb_ctx *c) {
    gemb_halo_pool &PL = c->halo_pool;
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < PL.nbuf; i++)
        for (int q = 0; q < c->nranks; q++)
            if (q != c->rank && PL.peer_buf[i][q]) cudaIpcCloseMemHandle(PL.peer_buf[i][q]);
    for (int q = 0; q < c->nranks; q++)
        if (q != c->rank && PL.peer_flags[q]) cudaIpcCloseMemHandle(PL.peer_flags[q]);
    for (int i = 0; i < PL.nbuf; i++) cudaFree(PL.buf[i]);
    cudaFree(PL.flags); cudaFree(PL.timeout_flag);
    (void)cudaGetLastError();
    PL = gemb_halo_pool();
}

}  // namespace gemb
