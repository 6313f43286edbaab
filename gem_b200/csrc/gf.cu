// gem_b200/csrc/gf.cu -- Graph Factorization: the edge SGD of gem/embedding/gf.py:94-104 (and of its C++ twin
// gem/c_src/gf.cpp:143-164):  for every epoch, for every edge (i, j, w) with j > i:
//        X[i] <- X[i] - eta * ( regu * X[i] - (w - <X[i], X[j]>) * X[j] )
// mode 0 (reference order): ONE warp walks the edge list in the order given, epoch after epoch -- exactly the reference's
//        sequential Gauss-Seidel sweep (fp32 where the Python loop is fp64); for the sizes the reference is used at.
// mode 1 (rows in parallel): one warp per source row; a row's own edges are applied in order with its running x_i (Gauss-Seidel
//        inside the row), the partner rows X[j] are read from the PREVIOUS epoch's table (Jacobi across rows, two tables): deterministic
//        and bit-reproducible whatever the launch shape, which is what lets tests/ compare it with the oracle's restatement.
// A lane owns the dimensions lane, lane + 32, ...; the dot product is a warp shuffle reduction; d <= 1024.
#include "common.cuh"
#include <chrono>

namespace gemb {

constexpr int GF_MAXV = 32;   // dimensions per lane kept in registers (d <= 1024)

template <int NV>
__device__ __forceinline__ void gf_edge(float (&xi)[NV], const float *__restrict__ xj_row, int d, int lane, float w, float eta, float regu) {
    float xj[NV];
    float dot = 0.f;
#pragma unroll
    for (int v = 0; v < NV; v++) {
        const int c = lane + 32 * v;
        xj[v] = c < d ? xj_row[c] : 0.f;
        dot = fmaf(xi[v], xj[v], dot);
    }
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float g = w - dot;
#pragma unroll
    for (int v = 0; v < NV; v++) xi[v] -= eta * (regu * xi[v] - g * xj[v]);
}

// mode 0: one warp, the reference's order
template <int NV>
__global__ void gf_sequential_kernel(int64_t m, const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                                     const float *__restrict__ w, int d, float eta, float regu, int epochs, float *X) {
    const int lane = threadIdx.x;
    for (int ep = 0; ep < epochs; ep++) {
        for (int64_t e = 0; e < m; e++) {
            const int i = src[e], j = dst[e];
            if (j <= i) continue;
            float xi[NV];
            float *xr = X + (int64_t)i * d;
#pragma unroll
            for (int v = 0; v < NV; v++) { const int c = lane + 32 * v; xi[v] = c < d ? xr[c] : 0.f; }
            gf_edge<NV>(xi, X + (int64_t)j * d, d, lane, w ? w[e] : 1.f, eta, regu);
#pragma unroll
            for (int v = 0; v < NV; v++) { const int c = lane + 32 * v; if (c < d) xr[c] = xi[v]; }
            __syncwarp();
        }
    }
}

// mode 1: warp per row; Xold read-only this epoch, Xnew written
template <int NV>
__global__ void __launch_bounds__(256)
gf_rows_kernel(int64_t n, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ dst, const float *__restrict__ w, int d,
               float eta, float regu, const float *__restrict__ Xold, float *__restrict__ Xnew) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        float xi[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) { const int c = lane + 32 * v; xi[v] = c < d ? Xold[i * d + c] : 0.f; }
        for (int64_t e = rowptr[i]; e < rowptr[i + 1]; e++) {
            const int j = dst[e];
            if (j <= i) continue;
            gf_edge<NV>(xi, Xold + (int64_t)j * d, d, lane, w ? w[e] : 1.f, eta, regu);
        }
#pragma unroll
        for (int v = 0; v < NV; v++) { const int c = lane + 32 * v; if (c < d) Xnew[i * d + c] = xi[v]; }
    }
}

}  // namespace gemb

using namespace gemb;

extern "C" int gemb_gf(gemb_ctx *ctx, int64_t n, int64_t m, const int32_t *src, const int32_t *dst, const float *w, int d, float eta,
                       float regu, int max_iter, int mode, const float *X0, float *X_out, double *device_ms_out) {
    GEMB_ARG(ctx && n > 0 && m >= 0 && X0 && X_out, "ctx / n / X0 / X_out");
    GEMB_ARG(m == 0 || (src && dst), "edge arrays");
    GEMB_ARG(d >= 1 && d <= 32 * GF_MAXV, "d must be in 1..1024");
    GEMB_ARG(max_iter >= 0 && (mode == 0 || mode == 1), "max_iter / mode");
    GEMB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    std::vector<int64_t> rowptr;
    for (int64_t e = 0; e < m; e++) {
        GEMB_ARG(src[e] >= 0 && src[e] < n && dst[e] >= 0 && dst[e] < n, "edge endpoint outside [0, n)");
        if (mode == 1 && e > 0) GEMB_ARG(src[e] >= src[e - 1], "mode 1 needs the edges grouped by source row (non-decreasing src)");
    }
    int32_t *d_src = nullptr, *d_dst = nullptr;
    float *d_w = nullptr, *Xa = nullptr, *Xb = nullptr;
    int64_t *d_rp = nullptr;
    const size_t xb = sizeof(float) * (size_t)n * d;
    int status = GEMB_OK;
    auto body = [&]() -> int {
        GEMB_CUDA(dmalloc(&d_dst, sizeof(int32_t) * std::max<int64_t>(m, 1)));
        GEMB_CUDA(cudaMemcpyAsync(d_dst, dst, sizeof(int32_t) * m, cudaMemcpyHostToDevice, st));
        if (w) { GEMB_CUDA(dmalloc(&d_w, sizeof(float) * std::max<int64_t>(m, 1))); GEMB_CUDA(cudaMemcpyAsync(d_w, w, sizeof(float) * m, cudaMemcpyHostToDevice, st)); }
        GEMB_CUDA(dmalloc(&Xa, xb));
        GEMB_CUDA(cudaMemcpyAsync(Xa, X0, xb, cudaMemcpyHostToDevice, st));
        if (mode == 0) {
            GEMB_CUDA(dmalloc(&d_src, sizeof(int32_t) * std::max<int64_t>(m, 1)));
            GEMB_CUDA(cudaMemcpyAsync(d_src, src, sizeof(int32_t) * m, cudaMemcpyHostToDevice, st));
        } else {
            rowptr.assign(n + 1, 0);
            for (int64_t e = 0; e < m; e++) rowptr[src[e] + 1]++;
            for (int64_t i = 0; i < n; i++) rowptr[i + 1] += rowptr[i];
            GEMB_CUDA(dmalloc(&d_rp, sizeof(int64_t) * (n + 1)));
            GEMB_CUDA(cudaMemcpyAsync(d_rp, rowptr.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st));
            GEMB_CUDA(dmalloc(&Xb, xb));
        }
        cudaEvent_t e0, e1;
        GEMB_CUDA(cudaEventCreate(&e0)); GEMB_CUDA(cudaEventCreate(&e1));
        GEMB_CUDA(cudaEventRecord(e0, st));
        const int nv = (d + 31) / 32;
        float *cur = Xa;
#define GF_DISPATCH(CALL)                                                                     \
        do {                                                                                  \
            if (nv <= 1) { CALL(1); } else if (nv <= 2) { CALL(2); } else if (nv <= 4) { CALL(4); } \
            else if (nv <= 8) { CALL(8); } else if (nv <= 16) { CALL(16); } else { CALL(32); } \
        } while (0)
        if (mode == 0) {
            if (m > 0 && max_iter > 0) {
#define SEQ(NV) gf_sequential_kernel<NV><<<1, 32, 0, st>>>(m, d_src, d_dst, d_w, d, eta, regu, max_iter, Xa)
                GF_DISPATCH(SEQ);
#undef SEQ
                GEMB_CUDA(cudaGetLastError());
                count_launch();
            }
        } else {
            const int grid = (int)std::min<int64_t>((n * 32 + 255) / 256, (int64_t)ctx->sm_count * 16);
            float *nxt = Xb;
            for (int ep = 0; ep < max_iter; ep++) {
#define ROWS(NV) gf_rows_kernel<NV><<<grid, 256, 0, st>>>(n, d_rp, d_dst, d_w, d, eta, regu, cur, nxt)
                GF_DISPATCH(ROWS);
#undef ROWS
                std::swap(cur, nxt);
            }
            GEMB_CUDA(cudaGetLastError());
            count_launch(max_iter);
        }
#undef GF_DISPATCH
        GEMB_CUDA(cudaEventRecord(e1, st));
        GEMB_CUDA(cudaMemcpyAsync(X_out, cur, xb, cudaMemcpyDeviceToHost, st));
        GEMB_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (device_ms_out) *device_ms_out = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return GEMB_OK;
    };
    status = body();
    cudaStreamSynchronize(st);
    dfree(d_src); dfree(d_dst); dfree(d_w); dfree(Xa); dfree(Xb); dfree(d_rp);
    return status;
}
