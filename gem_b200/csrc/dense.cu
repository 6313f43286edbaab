// gem_b200/csrc/dense.cu -- tall-skinny dense kernels of the HOPE solver (round-1 CUDA-core version).
//
//   gram   : G = P^T Q      (n x b1, n x b2 -> b1 x b2), the CholeskyQR / Rayleigh-Ritz contraction
//   apply  : Out = Q * M    (n x b1 times b1 x b2)
//   chol_inverse, eigh : single-CTA fp64 factorizations of the b x b matrices
// These replace numpy.linalg.qr / svd inside scipy's svds (hope.py:33 -> _svds.py:508-533).
// fp32 data, fp32 FMA inside a CTA's partial sums, fp64 across CTAs and in the b x b algebra.
#include "common.cuh"
#include <stdlib.h>
#include <algorithm>

namespace gemb {

// ------------------------------------------------------------------------------------ gram
// Output tile (16*TM) x (16*TM) per blockIdx.y, rows strided over blockIdx.x in chunks of KC.
template <int TM>
__global__ void __launch_bounds__(256)
gram_kernel(int64_t n, const float *__restrict__ P, int b1, const float *__restrict__ Q, int b2,
            double *__restrict__ G, int tiles_n) {
    constexpr int BT = 16 * TM;
    constexpr int KC = 32;
    __shared__ float sP[KC][BT];
    __shared__ float sQ[KC][BT];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int tile_m = blockIdx.y / tiles_n, tile_n = blockIdx.y % tiles_n;
    const int m0 = tile_m * BT, n0 = tile_n * BT;
    float acc[TM][TM];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TM; j++) acc[i][j] = 0.f;

    for (int64_t r0 = (int64_t)blockIdx.x * KC; r0 < n; r0 += (int64_t)gridDim.x * KC) {
        for (int idx = tid; idx < KC * BT; idx += 256) {
            const int kk = idx / BT, col = idx - kk * BT;
            const int64_t r = r0 + kk;
            float vp = 0.f, vq = 0.f;
            if (r < n) {
                if (m0 + col < b1) vp = __ldg(P + r * b1 + m0 + col);
                if (n0 + col < b2) vq = __ldg(Q + r * b2 + n0 + col);
            }
            sP[kk][col] = vp;
            sQ[kk][col] = vq;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < KC; kk++) {
            float a[TM], bb[TM];
#pragma unroll
            for (int i = 0; i < TM; i++) a[i] = sP[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TM; j++) bb[j] = sQ[kk][tx * TM + j];
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TM; j++) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int gm = m0 + ty * TM + i;
        if (gm >= b1) continue;
#pragma unroll
        for (int j = 0; j < TM; j++) {
            const int gn = n0 + tx * TM + j;
            if (gn < b2) atomicAdd(G + (size_t)gm * b2 + gn, (double)acc[i][j]);
        }
    }
}

static int pick_tm(int b) {
    // tile edge 16*TM for TM in {4,5,6,8}: minimise padded area, prefer fewer tiles on ties
    int best = 4;
    long best_cost = -1;
    const int cand[4] = {4, 5, 6, 8};
    for (int t = 0; t < 4; t++) {
        int bt = 16 * cand[t];
        long tiles = (b + bt - 1) / bt;
        long cost = tiles * bt;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = cand[t]; }
    }
    return best;
}

int gram_tc_launch(gemb_ctx *ctx, int64_t n, const float *P, int b1, const float *Q, int b2, double *G);

static int gram_mode() {   // GEMB_GRAM=fp32 forces the CUDA-core kernel (A/B testing); default = tcgen05
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("GEMB_GRAM");
        mode = (e && (e[0] == 'f' || e[0] == '0')) ? 0 : 1;
    }
    return mode;
}

int gram_fp32_launch(gemb_ctx *ctx, int64_t n, const float *P, int b1, const float *Q, int b2, double *G);

int gram_launch(gemb_ctx *ctx, int64_t n, const float *P, int b1, const float *Q, int b2, double *G) {
    if (gram_mode() == 1 && n >= 4096) {
        const int s = gram_tc_launch(ctx, n, P, b1, Q, b2, G);
        if (s != GEMB_ERR_UNSUPPORTED) return s;
    }
    return gram_fp32_launch(ctx, n, P, b1, Q, b2, G);
}

int gram_fp32_launch(gemb_ctx *ctx, int64_t n, const float *P, int b1, const float *Q, int b2, double *G) {
    GEMB_CUDA(cudaMemsetAsync(G, 0, sizeof(double) * (size_t)b1 * b2, ctx->stream));
    if (n == 0) return GEMB_OK;
    const int bmax = b1 > b2 ? b1 : b2;
    const int TM = pick_tm(bmax);
    const int BT = 16 * TM;
    const int tiles_m = (b1 + BT - 1) / BT, tiles_n = (b2 + BT - 1) / BT;
    int64_t chunks = (n + 31) / 32;
    int gx = ctx->sm_count * 4 / (tiles_m * tiles_n);
    if (gx < 1) gx = 1;
    if (gx > chunks) gx = (int)chunks;
    dim3 grid(gx, tiles_m * tiles_n), block(256);
    switch (TM) {
        case 4: gram_kernel<4><<<grid, block, 0, ctx->stream>>>(n, P, b1, Q, b2, G, tiles_n); break;
        case 5: gram_kernel<5><<<grid, block, 0, ctx->stream>>>(n, P, b1, Q, b2, G, tiles_n); break;
        case 6: gram_kernel<6><<<grid, block, 0, ctx->stream>>>(n, P, b1, Q, b2, G, tiles_n); break;
        default: gram_kernel<8><<<grid, block, 0, ctx->stream>>>(n, P, b1, Q, b2, G, tiles_n); break;
    }
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    return GEMB_OK;
}

// ------------------------------------------------------------------------------------ apply
// Out[r, n0 + ..] = sum_k Q[r, k] * M[k, ..];  CTA tile: 64 rows x (16*TN) columns, K chunks of 16.
template <int TN>
__global__ void __launch_bounds__(256)
apply_kernel(int64_t n, const float *__restrict__ Q, int b1, const float *__restrict__ M, int ldm,
             int b2, float *__restrict__ Out, int ldo) {
    constexpr int BN = 16 * TN;
    constexpr int BM = 64, BK = 16;
    __shared__ float sA[BK][BM + 1];
    __shared__ float sB[BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // ty -> 4 rows, tx -> TN columns
    const int64_t r0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < b1; k0 += BK) {
        for (int idx = tid; idx < BM * BK; idx += 256) {
            const int rr = idx / BK, kk = idx - rr * BK;
            const int64_t r = r0 + rr;
            sA[kk][rr] = (r < n && k0 + kk < b1) ? __ldg(Q + r * b1 + k0 + kk) : 0.f;
        }
        for (int idx = tid; idx < BK * BN; idx += 256) {
            const int kk = idx / BN, col = idx - kk * BN;
            sB[kk][col] = (k0 + kk < b1 && n0 + col < b2) ? __ldg(M + (size_t)(k0 + kk) * ldm + n0 + col) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[4], bb[TN];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < TN; j++) bb[j] = sB[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t r = r0 + ty * 4 + i;
        if (r >= n) continue;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int col = n0 + tx * TN + j;
            if (col < b2) Out[r * ldo + col] = acc[i][j];
        }
    }
}

int apply_tc_launch(gemb_ctx *ctx, int64_t n, const float *Q, int b1, const float *M, int ldm, int b2, float *Out, int ldo);

static int apply_mode() {   // GEMB_APPLY=fp32 forces the CUDA-core kernel; default = tcgen05 where the shape fits
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("GEMB_APPLY");
        mode = (e && (e[0] == 'f' || e[0] == '0')) ? 0 : 1;
    }
    return mode;
}

int apply_launch(gemb_ctx *ctx, int64_t n, const float *Q, int b1, const float *M, int ldm, int b2,
                 float *Out, int ldo) {
    if (n == 0 || b2 == 0) return GEMB_OK;
    if (apply_mode() == 1 && n >= 4096) {
        const int s = apply_tc_launch(ctx, n, Q, b1, M, ldm, b2, Out, ldo);
        if (s != GEMB_ERR_UNSUPPORTED) return s;
    }
    return apply_fp32_launch(ctx, n, Q, b1, M, ldm, b2, Out, ldo);
}

int apply_fp32_launch(gemb_ctx *ctx, int64_t n, const float *Q, int b1, const float *M, int ldm, int b2,
                      float *Out, int ldo) {
    const int TN = pick_tm(b2);
    const int BN = 16 * TN;
    dim3 grid((unsigned)((n + 63) / 64), (b2 + BN - 1) / BN), block(256);
    switch (TN) {
        case 4: apply_kernel<4><<<grid, block, 0, ctx->stream>>>(n, Q, b1, M, ldm, b2, Out, ldo); break;
        case 5: apply_kernel<5><<<grid, block, 0, ctx->stream>>>(n, Q, b1, M, ldm, b2, Out, ldo); break;
        case 6: apply_kernel<6><<<grid, block, 0, ctx->stream>>>(n, Q, b1, M, ldm, b2, Out, ldo); break;
        default: apply_kernel<8><<<grid, block, 0, ctx->stream>>>(n, Q, b1, M, ldm, b2, Out, ldo); break;
    }
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    return GEMB_OK;
}

// ------------------------------------------------------------------------------------ chol_inverse
// G (b x b fp64, symmetric) -> Minv = R^-1 (fp32, upper triangular) with G = R^T R.
// Work on the diagonally scaled matrix D^-1/2 G D^-1/2 (unit diagonal) so that the rank test is
// scale free.  A pivot below PIV_EPS marks the column numerically dependent: its column of Minv is
// zero (the orthonormalised block then carries a zero column, which stays zero under S).
#define GEMB_PIV_EPS 1e-5
// SMEM: the b x b matrix lives in shared memory for the whole factorization (b <= 160).
template <bool SMEM>
__global__ void __launch_bounds__(1024)
chol_inverse_kernel(int b, double *__restrict__ Gg, float *__restrict__ Minv, double *__restrict__ Minv64,
                    int *__restrict__ rank_out) {
    extern __shared__ double sh[];
    double *dscale = sh;           // b : 1/sqrt(G_jj) (0 if G_jj <= 0)
    double *keep = sh + b;         // b : 1.0 if column j kept, 0.0 if numerically dependent
    double *xdiag = sh + 2 * b;    // b
    double *G = SMEM ? sh + 3 * b : Gg;
    __shared__ int s_rank;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (SMEM) {
        for (int idx = tid; idx < b * b; idx += nt) G[idx] = Gg[idx];
        __syncthreads();
    }
    for (int j = tid; j < b; j += nt) {
        const double d = G[(size_t)j * b + j];
        dscale[j] = d > 0.0 ? rsqrt(d) : 0.0;
    }
    if (tid == 0) s_rank = 0;
    __syncthreads();
    for (int idx = tid; idx < b * b; idx += nt) {
        const int i = idx / b, j = idx - i * b;
        G[idx] = G[idx] * dscale[i] * dscale[j];
    }
    __syncthreads();
    // right-looking Cholesky on the lower triangle: L overwrites G (lower)
    for (int j = 0; j < b; j++) {
        if (tid == 0) {
            const double d = G[(size_t)j * b + j];
            if (d > GEMB_PIV_EPS) {
                G[(size_t)j * b + j] = sqrt(d);
                keep[j] = 1.0;
                s_rank++;
            } else {
                G[(size_t)j * b + j] = 0.0;
                keep[j] = 0.0;
            }
        }
        __syncthreads();
        const double ljj = G[(size_t)j * b + j];
        const double inv = ljj > 0.0 ? 1.0 / ljj : 0.0;
        for (int i = j + 1 + tid; i < b; i += nt) G[(size_t)i * b + j] *= inv;
        __syncthreads();
        // trailing update of the lower triangle: G[i][k] -= L[i][j] * L[k][j], j < k <= i
        const int m = b - j - 1;
        for (int idx = tid; idx < m * m; idx += nt) {
            const int ii = idx / m, kk = idx - ii * m;
            if (kk <= ii) {
                const int i = j + 1 + ii, k = j + 1 + kk;
                G[(size_t)i * b + k] -= G[(size_t)i * b + j] * G[(size_t)k * b + j];
            }
        }
        __syncthreads();
    }
    // R^-1 = D^-1/2 * L^-T.  Column c of L^-1 by forward substitution (one warp per column, the dot
    // products split across lanes); x_i (i > c) is kept in the free strict upper triangle G[c][i].
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    for (int c = warp; c < b; c += nwarps) {
        const bool okc = keep[c] != 0.0;
        const double xc = okc ? 1.0 / G[(size_t)c * b + c] : 0.0;
        if (lane == 0) xdiag[c] = xc;
        for (int i = c + 1; i < b; i++) {
            double s = 0.0;
            if (okc && keep[i] != 0.0) {
                for (int k = c + 1 + lane; k < i; k += 32) s += G[(size_t)i * b + k] * G[(size_t)c * b + k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                s += G[(size_t)i * b + c] * xc;
                s = -s / G[(size_t)i * b + i];
            }
            __syncwarp();
            if (lane == 0) G[(size_t)c * b + i] = s;  // (L^-1)[i][c]
            __syncwarp();
        }
    }
    __syncthreads();
    for (int idx = tid; idx < b * b; idx += nt) {
        const int r = idx / b, c = idx - r * b;  // Minv[r][c] = dscale[r] * (L^-1)[c][r], r <= c
        double v = 0.0;
        if (r == c) v = xdiag[r] * dscale[r];
        else if (r < c) v = G[(size_t)r * b + c] * dscale[r];
        Minv[idx] = (float)v;
        if (Minv64) Minv64[idx] = v;
    }
    if (tid == 0 && rank_out) *rank_out = s_rank;
}

// Fast variant (b*b fp64 fits in shared memory): two block barriers per Cholesky column (the <= 3 warps that own
// the column compute the pivot themselves), and the triangular inverse without block barriers -- each column
// of L^-1 belongs to 8 lanes of one warp that split the dot products.  ~4x faster than the generic kernel
// (0.21 ms -> measured below) for b = 80; same arithmetic (fp64), same pivot rule.
__global__ void __launch_bounds__(1024)
chol_inverse_fast_kernel(int b, const double *__restrict__ Gg, float *__restrict__ Minv, double *__restrict__ Minv64,
                         int *__restrict__ rank_out) {
    extern __shared__ double sh[];
    const int ld = b | 1;          // odd leading dimension: column walks are bank-conflict free
    double *dscale = sh;           // b
    double *ldiag = sh + b;        // b : L_jj (0 if the column was dropped)
    double *colj = sh + 2 * b;     // b : scaled column j of L
    double *G = sh + 3 * b;        // b x ld
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int j = tid; j < b; j += nt) {
        const double d = Gg[(size_t)j * b + j];
        dscale[j] = d > 0.0 ? rsqrt(d) : 0.0;
    }
    __syncthreads();
    for (int idx = tid; idx < b * b; idx += nt) {
        const int i = idx / b, j = idx - i * b;
        G[i * ld + j] = Gg[idx] * dscale[i] * dscale[j];
    }
    __syncthreads();
    const int gi0 = tid / b, gk0 = tid - gi0 * b, gdi = nt / b, gdk = nt - gdi * b;
    for (int j = 0; j < b; j++) {
        const int m = b - j - 1;
        if (tid < m || tid == 0) {   // the column's owners each derive the pivot (no broadcast barrier)
            const double d = G[j * ld + j];
            const bool ok = d > GEMB_PIV_EPS;
            const double ljj = ok ? sqrt(d) : 0.0;
            const double inv = ok ? 1.0 / ljj : 0.0;
            if (tid == 0) ldiag[j] = ljj;
            if (tid < m) {
                const int i = j + 1 + tid;
                const double v = G[i * ld + j] * inv;
                colj[i] = v;
                G[i * ld + j] = v;
            }
        }
        __syncthreads();
        // every thread owns the same (i, k) positions of the b x b grid in all steps (no index division)
        for (int i = gi0, k = gk0; i < b;) {
            if (k > j && k <= i) G[i * ld + k] -= colj[i] * colj[k];
            k += gdk; i += gdi;
            if (k >= b) { k -= b; i++; }
        }
        __syncthreads();
    }
    // X = L^-1 by rows; column c of X is kept in the free strict upper triangle G[c][i] = X[i][c]
    const int lane8 = tid & 7, grp = tid >> 3, ngrp = nt >> 3;
    const unsigned gmask = 0xffu << ((tid & 31) & ~7);
    for (int c = grp; c < b; c += ngrp) {
        const double lcc = ldiag[c];
        const bool okc = lcc > 0.0;
        const double xc = okc ? 1.0 / lcc : 0.0;
        for (int i = c + 1; i < b; i++) {
            const double lii = ldiag[i];
            double sum = 0.0;
            if (okc && lii > 0.0) {
                for (int k = c + 1 + lane8; k < i; k += 8) sum += G[i * ld + k] * G[c * ld + k];
                sum += __shfl_xor_sync(gmask, sum, 4, 8);
                sum += __shfl_xor_sync(gmask, sum, 2, 8);
                sum += __shfl_xor_sync(gmask, sum, 1, 8);
                sum = -(sum + G[i * ld + c] * xc) / lii;
            }
            __syncwarp(gmask);
            if (lane8 == 0) G[c * ld + i] = sum;
            __syncwarp(gmask);
        }
    }
    __syncthreads();
    int rank = 0;
    for (int idx = tid; idx < b * b; idx += nt) {
        const int r = idx / b, c = idx - r * b;  // Minv[r][c] = dscale[r] * X[c][r], r <= c
        double v = 0.0;
        if (r == c) { v = ldiag[r] > 0.0 ? dscale[r] / ldiag[r] : 0.0; }
        else if (r < c) v = G[r * ld + c] * dscale[r];
        Minv[idx] = (float)v;
        if (Minv64) Minv64[idx] = v;
    }
    if (tid == 0 && rank_out) {
        for (int j = 0; j < b; j++) rank += ldiag[j] > 0.0;
        *rank_out = rank;
    }
}

int chol_inverse_launch(gemb_ctx *ctx, int b, double *G, float *Minv, int *rank_out_dev, double *Minv64) {
    const size_t small = sizeof(double) * 3 * (size_t)b;
    const size_t big = small + sizeof(double) * (size_t)b * b;
    const size_t fast = small + sizeof(double) * (size_t)b * (b | 1);
    static const bool generic_only = getenv("GEMB_DENSE_GENERIC") != nullptr;   // A/B switch for the tests
    if (fast <= 200 * 1024 && b <= 128 && !generic_only) {
        static bool attr_fast = false;
        if (!attr_fast) {
            GEMB_CUDA(cudaFuncSetAttribute(chol_inverse_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_fast = true;
        }
        chol_inverse_fast_kernel<<<1, 1024, fast, ctx->stream>>>(b, G, Minv, Minv64, rank_out_dev);
    } else if (big <= 220 * 1024) {
        static bool attr_set = false;
        if (!attr_set) {
            GEMB_CUDA(cudaFuncSetAttribute(chol_inverse_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));  // + 4 B static
            attr_set = true;
        }
        chol_inverse_kernel<true><<<1, 1024, big, ctx->stream>>>(b, G, Minv, Minv64, rank_out_dev);
    } else {
        chol_inverse_kernel<false><<<1, 1024, small, ctx->stream>>>(b, G, Minv, Minv64, rank_out_dev);
    }
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    return GEMB_OK;
}

// ------------------------------------------------------------------------------------ eigh
// Two-sided cyclic Jacobi with round-robin (circle-method) pair ordering, one CTA, fp64.
// A is destroyed; w ascending; Z column j <-> w[j].  Zt is b x b scratch.
// MODE 2: A and Zt in shared memory (b <= 116); MODE 1: A in shared memory (b <= 165); MODE 0: global.
template <int MODE>
__global__ void __launch_bounds__(1024)
eigh_jacobi_kernel(int b, double *__restrict__ Ag, double *__restrict__ w, double *__restrict__ Z,
                   double *__restrict__ Ztg, int max_sweeps, double rel_tol) {
    extern __shared__ double sh[];
    const int m = (b + 1) & ~1;     // even number of players; index >= b is a dummy
    const int half = m / 2;
    double *cs = sh;                // half
    double *sn = sh + half;         // half
    int *pp = (int *)(sh + 2 * half);
    int *qq = pp + half;
    double *mat = sh + 3 * half + 2;
    double *A = MODE >= 1 ? mat : Ag;
    double *Zt = MODE >= 2 ? mat + (size_t)b * b : Ztg;
    __shared__ double s_off, s_diag;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int idx = tid; idx < b * b; idx += nt) {
        if (MODE >= 1) A[idx] = Ag[idx];
        Zt[idx] = (idx / b == idx % b) ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int sweep = 0; sweep < max_sweeps; sweep++) {
        if (tid == 0) { s_off = 0.0; s_diag = 0.0; }
        __syncthreads();
        double off = 0.0, dg = 0.0;
        for (int idx = tid; idx < b * b; idx += nt) {
            const int i = idx / b, j = idx - i * b;
            const double v = A[idx];
            if (i == j) dg += v * v; else off += v * v;
        }
        for (int o = 16; o > 0; o >>= 1) {
            off += __shfl_xor_sync(0xffffffffu, off, o);
            dg += __shfl_xor_sync(0xffffffffu, dg, o);
        }
        if ((tid & 31) == 0) { atomicAdd(&s_off, off); atomicAdd(&s_diag, dg); }
        __syncthreads();
        if (s_off <= rel_tol * rel_tol * (s_diag + s_off) || s_diag + s_off == 0.0) break;
        for (int r = 0; r < m - 1; r++) {
            if (tid < half) {
                int p, q;
                if (tid == 0) { p = m - 1; q = r % (m - 1); }
                else { p = (r + tid) % (m - 1); q = (r + m - 1 - tid) % (m - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                double c = 1.0, s = 0.0;
                if (q < b) {
                    const double apq = A[(size_t)p * b + q];
                    if (apq != 0.0) {
                        const double app = A[(size_t)p * b + p], aqq = A[(size_t)q * b + q];
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = rsqrt(t * t + 1.0);
                        s = t * c;
                    }
                } else { q = -1; }
                pp[tid] = p; qq[tid] = q; cs[tid] = c; sn[tid] = s;
            }
            __syncthreads();
            // columns: A <- A J, Zt <- Zt J   (consecutive threads -> different pairs, same row)
            for (int idx = tid; idx < half * b; idx += nt) {
                const int k = idx / half, pi = idx - k * half;
                const int p = pp[pi], q = qq[pi];
                if (q < 0) continue;
                const double c = cs[pi], s = sn[pi];
                if (s == 0.0) continue;
                double x = A[(size_t)k * b + p], y = A[(size_t)k * b + q];
                A[(size_t)k * b + p] = c * x - s * y;
                A[(size_t)k * b + q] = s * x + c * y;
                x = Zt[(size_t)k * b + p]; y = Zt[(size_t)k * b + q];
                Zt[(size_t)k * b + p] = c * x - s * y;
                Zt[(size_t)k * b + q] = s * x + c * y;
            }
            __syncthreads();
            // rows: A <- J^T A   (consecutive threads -> consecutive columns)
            for (int idx = tid; idx < half * b; idx += nt) {
                const int pi = idx / b, k = idx - pi * b;
                const int p = pp[pi], q = qq[pi];
                if (q < 0) continue;
                const double c = cs[pi], s = sn[pi];
                if (s == 0.0) continue;
                const double x = A[(size_t)p * b + k], y = A[(size_t)q * b + k];
                A[(size_t)p * b + k] = c * x - s * y;
                A[(size_t)q * b + k] = s * x + c * y;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    // sort ascending by rank counting, permute eigenvector columns (one warp per eigenvalue)
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    for (int j = warp; j < b; j += nwarps) {
        const double wj = A[(size_t)j * b + j];
        int rank = 0;
        for (int i = lane; i < b; i += 32) {
            const double wi = A[(size_t)i * b + i];
            rank += (wi < wj) || (wi == wj && i < j);
        }
        for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
        if (lane == 0) w[rank] = wj;
        for (int k = lane; k < b; k += 32) Z[(size_t)k * b + rank] = Zt[(size_t)k * b + j];
    }
}

// Fast variant for b*(b|1)*16 bytes <= shared memory (b <= 112).  The generic kernel is ISSUE bound, not
// bandwidth bound (ncu: 558 warp instructions per warp per round, 63 % issue-active, fp64 pipe 11 %): runtime
// integer divisions and four parameter loads per element.  Here
//   * the two-sided update A <- J^T A J is done per 2x2 BLOCK {p,q} x {r,s} of two rotation pairs by one thread
//     (4 loads, both rotations in registers, 4 stores): half the shared-memory traffic and one block barrier per
//     round less than column phase + row phase;
//   * the eigenvector accumulator is kept transposed so that its update is a row walk;
//   * (pair, column) indices advance incrementally (no division in the loops), rotation parameters are one
//     16-byte and one 8-byte load, the leading dimension is odd (conflict-free row and column walks).
// ZT_GLOBAL (112 < b <= 164, the Rayleigh-Ritz matrix of the thick-restart Lanczos solver): A alone fills the shared
// memory, the eigenvector accumulator lives in global memory (L2 resident, 200 KB) and is updated by coalesced row
// walks -- the generic MODE 1 kernel needed 15.2 ms for b = 160 (10.9 us per Jacobi round, 43 % of an R-MAT solve).
template <bool ZT_GLOBAL>
__global__ void __launch_bounds__(1024)
eigh_jacobi_fast_kernel(int b, const double *__restrict__ Ag, double *__restrict__ w, double *__restrict__ Z,
                        double *__restrict__ Ztg, int max_sweeps, double rel_tol) {
    extern __shared__ __align__(16) unsigned char sh_fast[];
    double *sh = (double *)sh_fast;
    const int m = (b + 1) & ~1;
    const int half = m / 2;
    const int ld = b | 1;
    double2 *csn = (double2 *)sh;                 // half : (c, s)
    int2 *pq = (int2 *)(sh + 2 * half);           // half : (p, q), q = -1 for the dummy partner
    double *A = sh + 3 * half + 2;                // b x ld
    double *ZT = ZT_GLOBAL ? Ztg : A + (size_t)b * ld;   // ZT[j][k] = component k of eigenvector j
    const int ldz = ZT_GLOBAL ? b : ld;
    __shared__ double s_off, s_diag;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int idx = tid; idx < b * b; idx += nt) {
        const int i = idx / b, j = idx - i * b;
        A[i * ld + j] = Ag[idx];
        ZT[i * ldz + j] = (i == j) ? 1.0 : 0.0;
    }
    // incremental (pair, column) walks: idx = tid + t * nt  ->  (idx / div, idx % div)
    const int zb_i0 = tid / b, zb_k0 = tid - zb_i0 * b, zb_di = nt / b, zb_dk = nt - zb_di * b;
    const int bl_i0 = tid / half, bl_j0 = tid - bl_i0 * half, bl_di = nt / half, bl_dj = nt - bl_di * half;
    __syncthreads();
    for (int sweep = 0; sweep < max_sweeps; sweep++) {
        if (tid == 0) { s_off = 0.0; s_diag = 0.0; }
        __syncthreads();
        double off = 0.0, dg = 0.0;
        for (int i = zb_i0, j = zb_k0; i < b;) {
            const double v = A[i * ld + j];
            if (i == j) dg += v * v; else off += v * v;
            j += zb_dk; i += zb_di;
            if (j >= b) { j -= b; i++; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            off += __shfl_xor_sync(0xffffffffu, off, o);
            dg += __shfl_xor_sync(0xffffffffu, dg, o);
        }
        if ((tid & 31) == 0) { atomicAdd(&s_off, off); atomicAdd(&s_diag, dg); }
        __syncthreads();
        if (s_off <= rel_tol * rel_tol * (s_diag + s_off) || s_diag + s_off == 0.0) break;
        for (int r = 0; r < m - 1; r++) {
            if (tid < half) {
                int p, q;
                if (tid == 0) { p = m - 1; q = r % (m - 1); }
                else { p = (r + tid) % (m - 1); q = (r + m - 1 - tid) % (m - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                double c = 1.0, s = 0.0;
                if (q < b) {
                    const double apq = A[p * ld + q];
                    if (apq != 0.0) {
                        const double app = A[p * ld + p], aqq = A[q * ld + q];
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = rsqrt(t * t + 1.0);
                        s = t * c;
                    }
                } else { q = -1; }
                pq[tid] = make_int2(p, q);
                csn[tid] = make_double2(c, s);
            }
            __syncthreads();
            // A <- J^T A J, one 2x2 block {p,q} x {r2,s2} per thread
            for (int pi = bl_i0, rj = bl_j0; pi < half;) {
                const double2 r1 = csn[pi], r2 = csn[rj];
                if (r1.y != 0.0 || r2.y != 0.0) {
                    const int2 a = pq[pi], c2 = pq[rj];
                    const bool hq = a.y >= 0, hs = c2.y >= 0;
                    double *row_p = A + a.x * ld, *row_q = A + (hq ? a.y : a.x) * ld;
                    const int cr = c2.x, cs2 = hs ? c2.y : c2.x;
                    const double apr = row_p[cr], aps = hs ? row_p[cs2] : 0.0;
                    const double aqr = hq ? row_q[cr] : 0.0, aqs = (hq && hs) ? row_q[cs2] : 0.0;
                    const double tpr = r1.x * apr - r1.y * aqr, tqr = r1.y * apr + r1.x * aqr;
                    const double tps = r1.x * aps - r1.y * aqs, tqs = r1.y * aps + r1.x * aqs;
                    row_p[cr] = r2.x * tpr - r2.y * tps;
                    if (hs) row_p[cs2] = r2.y * tpr + r2.x * tps;
                    if (hq) {
                        row_q[cr] = r2.x * tqr - r2.y * tqs;
                        if (hs) row_q[cs2] = r2.y * tqr + r2.x * tqs;
                    }
                }
                rj += bl_dj; pi += bl_di;
                if (rj >= half) { rj -= half; pi++; }
            }
            // Z^T <- J^T Z^T (rows p, q; consecutive threads = consecutive columns)
            if (ZT_GLOBAL) {
                // global (L2) accumulator: batches of 4 independent element pairs -- all 8 loads are issued before the
                // first store, otherwise every element pays a full L2 round trip in sequence
                int pi = zb_i0, k = zb_k0;
                while (pi < half) {
                    double *xp[4], *yp[4];
                    double2 rt[4];
                    double x[4], y[4];
                    bool on[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        on[u] = false;
                        if (pi < half) {
                            rt[u] = csn[pi];
                            const int2 a = pq[pi];
                            if (a.y >= 0 && rt[u].y != 0.0) {
                                on[u] = true;
                                xp[u] = ZT + a.x * ldz + k; yp[u] = ZT + a.y * ldz + k;
                            }
                            k += zb_dk; pi += zb_di;
                            if (k >= b) { k -= b; pi++; }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) if (on[u]) { x[u] = __ldcg(xp[u]); y[u] = __ldcg(yp[u]); }
#pragma unroll
                    for (int u = 0; u < 4; u++) if (on[u]) {
                        __stcg(xp[u], rt[u].x * x[u] - rt[u].y * y[u]);
                        __stcg(yp[u], rt[u].y * x[u] + rt[u].x * y[u]);
                    }
                }
            } else {
                for (int pi = zb_i0, k = zb_k0; pi < half;) {
                    const double2 rt = csn[pi];
                    const int2 a = pq[pi];
                    if (a.y >= 0 && rt.y != 0.0) {
                        double *xp = ZT + a.x * ldz + k, *yp = ZT + a.y * ldz + k;
                        const double x = *xp, y = *yp;
                        *xp = rt.x * x - rt.y * y;
                        *yp = rt.y * x + rt.x * y;
                    }
                    k += zb_dk; pi += zb_di;
                    if (k >= b) { k -= b; pi++; }
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    for (int j = warp; j < b; j += nwarps) {
        const double wj = A[j * ld + j];
        int rank = 0;
        for (int i = lane; i < b; i += 32) {
            const double wi = A[i * ld + i];
            rank += (wi < wj) || (wi == wj && i < j);
        }
        for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
        if (lane == 0) w[rank] = wj;
        for (int k = lane; k < b; k += 32) Z[(size_t)k * b + rank] = ZT_GLOBAL ? __ldcg(ZT + j * ldz + k) : ZT[j * ldz + k];
    }
}

int eigh_launch(gemb_ctx *ctx, int b, double *G, double *w, double *Z, double *Zscratch, double rel_tol) {
    const int half = ((b + 1) & ~1) / 2;
    const size_t base = sizeof(double) * (3 * half + 2);
    const size_t one = sizeof(double) * (size_t)b * b;
    const size_t cap = 220 * 1024;
    static const bool generic_only = getenv("GEMB_DENSE_GENERIC") != nullptr;
    const size_t fast = base + 2 * sizeof(double) * (size_t)b * (b | 1);
    const size_t fast_a = base + sizeof(double) * (size_t)b * (b | 1);       // A only; eigenvectors in global memory
    if ((fast <= cap || fast_a <= cap) && !generic_only) {
        static bool attr_fast = false;
        if (!attr_fast) {
            GEMB_CUDA(cudaFuncSetAttribute(eigh_jacobi_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
            GEMB_CUDA(cudaFuncSetAttribute(eigh_jacobi_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
            attr_fast = true;
        }
        if (fast <= cap) eigh_jacobi_fast_kernel<false><<<1, 1024, fast, ctx->stream>>>(b, G, w, Z, Zscratch, 30, rel_tol);
        else eigh_jacobi_fast_kernel<true><<<1, 1024, fast_a, ctx->stream>>>(b, G, w, Z, Zscratch, 30, rel_tol);
        GEMB_CUDA(cudaGetLastError());
        count_launch();
        return GEMB_OK;
    }
    static bool attr_set = false;
    if (!attr_set) {
        GEMB_CUDA(cudaFuncSetAttribute(eigh_jacobi_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
        GEMB_CUDA(cudaFuncSetAttribute(eigh_jacobi_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
        attr_set = true;
    }
    if (base + 2 * one <= cap)
        eigh_jacobi_kernel<2><<<1, 1024, base + 2 * one, ctx->stream>>>(b, G, w, Z, Zscratch, 30, rel_tol);
    else if (base + one <= cap)
        eigh_jacobi_kernel<1><<<1, 1024, base + one, ctx->stream>>>(b, G, w, Z, Zscratch, 30, rel_tol);
    else
        eigh_jacobi_kernel<0><<<1, 1024, base, ctx->stream>>>(b, G, w, Z, Zscratch, 30, rel_tol);
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    return GEMB_OK;
}

// ------------------------------------------------------------------------------------ small b x b products
// C = op(A) * B, all b x b fp64 row-major (one CTA; used for the Ritz rotation of Gram matrices)
__global__ void __launch_bounds__(1024)
small_gemm_kernel(int b, const double *__restrict__ A, int transA, const double *__restrict__ B,
                  double *__restrict__ C, float *__restrict__ C32) {
    for (int idx = threadIdx.x; idx < b * b; idx += blockDim.x) {
        const int i = idx / b, j = idx - i * b;
        double acc = 0.0;
        if (transA) for (int k = 0; k < b; k++) acc += A[(size_t)k * b + i] * B[(size_t)k * b + j];
        else for (int k = 0; k < b; k++) acc += A[(size_t)i * b + k] * B[(size_t)k * b + j];
        if (C) C[idx] = acc;
        if (C32) C32[idx] = (float)acc;
    }
}

int small_gemm_launch(gemb_ctx *ctx, int b, const double *A, int transA, const double *B, double *C, float *C32) {
    small_gemm_kernel<<<1, 1024, 0, ctx->stream>>>(b, A, transA, B, C, C32);
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    return GEMB_OK;
}

// ------------------------------------------------------------------------------------ misc
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// standard normal per (global row, column): independent of the sharding
__global__ void randn_kernel(int64_t n, int b, uint64_t seed, uint64_t row_offset, float *__restrict__ X) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * b) return;
    const int64_t r = idx / b;
    const int c = (int)(idx - r * b);
    const uint64_t h = splitmix64(seed ^ splitmix64(((uint64_t)(r + row_offset) << 12) ^ (uint64_t)c));
    const uint32_t u1 = (uint32_t)(h >> 32), u2 = (uint32_t)h;
    const float f1 = ((float)u1 + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
    const float f2 = (float)u2 * 2.3283064365386963e-10f;
    X[idx] = sqrtf(-2.0f * logf(f1)) * cospif(2.0f * f2);
}

int randn_launch(gemb_ctx *ctx, int64_t n, int b, uint64_t seed, uint64_t row_offset, float *X) {
    const int64_t tot = n * b;
    if (tot == 0) return GEMB_OK;
    randn_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(n, b, seed, row_offset, X);
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    return GEMB_OK;
}

__global__ void sumsq_kernel(int64_t count, const float *__restrict__ X, double *__restrict__ out) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double v = X[i];
        acc += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

int sumsq_launch(gemb_ctx *ctx, int64_t count, const float *X, double *out_dev) {
    GEMB_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), ctx->stream));
    if (count == 0) return GEMB_OK;
    int grid = ctx->sm_count * 8;
    if ((int64_t)grid * 256 > count) grid = (int)((count + 255) / 256);
    sumsq_kernel<<<grid, 256, 0, ctx->stream>>>(count, X, out_dev);
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    return GEMB_OK;
}

__global__ void scale_kernel(int64_t count, float s, float *__restrict__ X) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x)
        X[i] *= s;
}

int scale_launch(gemb_ctx *ctx, int64_t count, float s, float *X) {
    if (count == 0) return GEMB_OK;
    int grid = ctx->sm_count * 8;
    if ((int64_t)grid * 256 > count) grid = (int)((count + 255) / 256);
    scale_kernel<<<grid, 256, 0, ctx->stream>>>(count, s, X);
    GEMB_CUDA(cudaGetLastError());
    count_launch();
    return GEMB_OK;
}

}  // namespace gemb

extern "C" int gemb_gram(gemb_ctx *c, int64_t n, const float *P, int b1, const float *Q, int b2,
                         int use_tensor_cores, double *G_out) {
    using namespace gemb;
    GEMB_ARG(c && P && G_out && n >= 0 && b1 > 0 && b2 > 0, "ctx/P/G/n/b");
    GEMB_CUDA(cudaSetDevice(c->device));
    float *dP = nullptr, *dQ = nullptr;
    double *dG = nullptr;
    GEMB_CUDA(dmalloc(&dP, sizeof(float) * (size_t)std::max<int64_t>(n, 1) * b1));
    GEMB_CUDA(cudaMemcpyAsync(dP, P, sizeof(float) * (size_t)n * b1, cudaMemcpyHostToDevice, c->stream));
    if (Q) {
        GEMB_CUDA(dmalloc(&dQ, sizeof(float) * (size_t)std::max<int64_t>(n, 1) * b2));
        GEMB_CUDA(cudaMemcpyAsync(dQ, Q, sizeof(float) * (size_t)n * b2, cudaMemcpyHostToDevice, c->stream));
    }
    GEMB_CUDA(dmalloc(&dG, sizeof(double) * (size_t)b1 * b2));
    int s = use_tensor_cores ? gram_tc_launch(c, n, dP, b1, Q ? dQ : dP, b2, dG)
                             : gram_fp32_launch(c, n, dP, b1, Q ? dQ : dP, b2, dG);
    if (s == GEMB_ERR_UNSUPPORTED) set_error("gemb_gram: shape (n=%lld, b1=%d, b2=%d) not supported by the tcgen05 kernel", (long long)n, b1, b2);
    if (s == GEMB_OK) {
        cudaError_t e = cudaMemcpyAsync(G_out, dG, sizeof(double) * (size_t)b1 * b2, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { set_error("gemb_gram: %s", cudaGetErrorString(e)); s = GEMB_ERR_CUDA; }
    }
    dfree(dP); dfree(dQ); dfree(dG);
    return s;
}

extern "C" int gemb_apply(gemb_ctx *c, int64_t n, const float *Q, int b1, const float *M, int b2,
                          int use_tensor_cores, float *Out) {
    using namespace gemb;
    GEMB_ARG(c && Q && M && Out && n >= 0 && b1 > 0 && b2 > 0, "ctx/Q/M/Out/n/b");
    GEMB_CUDA(cudaSetDevice(c->device));
    float *dQ = nullptr, *dM = nullptr, *dO = nullptr;
    GEMB_CUDA(dmalloc(&dQ, sizeof(float) * (size_t)std::max<int64_t>(n, 1) * b1));
    GEMB_CUDA(dmalloc(&dM, sizeof(float) * (size_t)b1 * b2));
    GEMB_CUDA(dmalloc(&dO, sizeof(float) * (size_t)std::max<int64_t>(n, 1) * b2));
    GEMB_CUDA(cudaMemcpyAsync(dQ, Q, sizeof(float) * (size_t)n * b1, cudaMemcpyHostToDevice, c->stream));
    GEMB_CUDA(cudaMemcpyAsync(dM, M, sizeof(float) * (size_t)b1 * b2, cudaMemcpyHostToDevice, c->stream));
    int s = use_tensor_cores ? apply_tc_launch(c, n, dQ, b1, dM, b2, b2, dO, b2)
                             : apply_fp32_launch(c, n, dQ, b1, dM, b2, b2, dO, b2);
    if (s == GEMB_ERR_UNSUPPORTED) set_error("gemb_apply: shape (n=%lld, b1=%d, b2=%d) not supported by the tcgen05 kernel", (long long)n, b1, b2);
    if (s == GEMB_OK) {
        cudaError_t e = cudaMemcpyAsync(Out, dO, sizeof(float) * (size_t)n * b2, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { set_error("gemb_apply: %s", cudaGetErrorString(e)); s = GEMB_ERR_CUDA; }
    }
    dfree(dQ); dfree(dM); dfree(dO);
    return s;
}
