/*
 * include/gemb200.h -- C ABI of libgemb200.so, the B200 (sm_100a) core behind GEM's
 * StaticGraphEmbedding plugin API for HOPE and node2vec.
 *
 * The reference has no FFI for this path: HOPE is four NumPy/SciPy lines
 * (gem/embedding/hope.py:28-36) and node2vec is an argv + text-file hand-off to a prebuilt
 * SNAP executable (gem/embedding/node2vec.py:31-53).  Each entry point below names the reference
 * interface it replaces; INTEGRATION.md shows the ctypes stub a GEM maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 (GEMB_OK) or a negative
 *     gemb_status; gemb_last_error() gives the message of the last failure on this thread.
 *   - the caller owns every host buffer; the library owns device memory behind opaque handles.
 *   - calls are blocking; one gemb_ctx is bound to one CUDA device and is not thread-safe.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     GEMB_ERR_CUDA.
 *   - matrices are row-major; node ids / column ids are int32; CSR offsets are int64 on the
 *     host ABI (node2vec) or int32 (HOPE shards, nnz < 2^31 per shard).
 */
#ifndef GEMB200_H
#define GEMB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GEMB_VERSION 100 /* 0.1.0 */

typedef enum {
    GEMB_OK = 0,
    GEMB_ERR_CUDA = -1,      /* no device / CUDA runtime or kernel failure */
    GEMB_ERR_ARG = -2,       /* bad argument */
    GEMB_ERR_NOMEM = -3,     /* host or device allocation failed */
    GEMB_ERR_DIVERGE = -4,   /* beta * ||A||_2 >= 1: Katz series does not converge */
    GEMB_ERR_NCCL = -5,      /* NCCL missing or collective failed */
    GEMB_ERR_UNSUPPORTED = -6
} gemb_status;

typedef struct gemb_ctx gemb_ctx;     /* one CUDA device + stream (+ optional NCCL communicator) */
typedef struct gemb_graph gemb_graph; /* a CSR row shard (and its transpose) resident in HBM   */

int gemb_version(void);
const char *gemb_last_error(void);
int gemb_device_count(void); /* number of CUDA devices, 0 if none / no driver */
int64_t gemb_launch_count(void); /* kernels this library has launched in this process so far */

int gemb_ctx_create(int device, gemb_ctx **out);
int gemb_ctx_destroy(gemb_ctx *ctx);

/* Pinned (page-locked) host buffers for the host<->device copies of the e2e path. */
int gemb_host_alloc(size_t bytes, void **out);
int gemb_host_free(void *p);

/* Device work buffers of one call (the CSR arrays, the n x (d/2+p) blocks) are kept in a per-device free
 * list when the call returns, so that the next learn_embedding on the same problem shape makes no driver
 * allocation (the reference re-allocates everything per call, hope.py:26-34; at 2.5 GB per call the
 * driver's page mapping costs more than the solve).  GEMB_CACHE_MB caps the list (0 = off).
 * gemb_mem_trim returns every cached block to the driver; gemb_mem_cached_bytes reports the list. */
int gemb_mem_trim(void);
size_t gemb_mem_cached_bytes(void);

/* ---- Graph Factorization.  Replaces the edge SGD of gem/embedding/gf.py:94-104 (its C++ twin:
 * gem/c_src/gf.cpp:143-164; the reference shells out to gem/c_exe/gf when it exists and then runs the Python loop anyway):
 *     for epoch in range(max_iter): for (i, j, w) in edges, j > i:  X[i] -= eta * (regu * X[i] - (w - <X[i], X[j]>) * X[j])
 * src / dst / w: the m directed edges (host; w NULL = 1).  X0: the n x d start (the reference draws 0.01 * randn), X_out: n x d.
 * mode 0: one warp applies the edges in the order given, epoch after epoch (the reference's sequential sweep, fp32).
 * mode 1: one warp per source row (edges grouped by src, in order), partner rows read from the previous epoch's table
 *         (Jacobi across rows, Gauss-Seidel inside a row): deterministic, for graphs beyond the reference's reach. */
int gemb_gf(gemb_ctx *ctx, int64_t n, int64_t m, const int32_t *src, const int32_t *dst, const float *w, int d, float eta,
            float regu, int max_iter, int mode, const float *X0, float *X_out, double *device_ms_out);

/* ---- bench infrastructure: Graph500 R-MAT generator on the device (BASELINE.json configs[3], configs[4]: scale 24).
 * No reference counterpart (GEM ships no generator; gem/tests load fixed fixtures); the host generator
 * gem_b200/synth.py::rmat makes the same kind of graph with NumPy for the small parity cases.
 * Rows [row0, row0 + n_rows) (n_rows < 0: all) of the symmetrised, loop-free, duplicate-free graph with sorted column
 * ids.  indices_out = NULL: only *nnz_out (shard) and *nnz_total_out (graph) are set; otherwise indptr_out (n_rows + 1
 * int64, shard-local offsets) and indices_out (cap >= nnz) are filled.  Counter-based RNG: identical on every rank. */
int gemb_synth_rmat(gemb_ctx *ctx, int scale, int edge_factor, double a, double b, double c, uint64_t seed, int permute,
                    int64_t row0, int64_t n_rows, int64_t *nnz_out, int64_t *nnz_total_out, int64_t *indptr_out,
                    int32_t *indices_out, int64_t cap);

/* ---- multi-GPU: one process per GPU; rank 0 makes the id, every rank calls init.
 * (No reference counterpart: GEM is single-process) */
#define GEMB_UNIQUE_ID_BYTES 128
int gemb_comm_unique_id(void *id_out /* GEMB_UNIQUE_ID_BYTES */);
int gemb_comm_init(gemb_ctx *ctx, int rank, int nranks, const void *id);

/* ---- graph upload.
 * Replaces: nx.to_numpy_matrix(graph) (hope.py:28) and the text edge list written by
 * graph_util.saveGraphToEdgeListTxtn2v (graph_util.py:137-140) + SNAP ReadGraph (bin@0x406550).
 *
 * The shard holds rows [row0, row0+n_local) of the n x n adjacency A in CSR form with GLOBAL
 * column ids, and the same row range of A^T (pass indptr_t == NULL when A is symmetric: A^T = A).
 * data / data_t may be NULL (all weights 1.0).  Single GPU: row0 = 0, n_local = n.
 * Multi GPU (after gemb_comm_init): for HOPE every rank uploads rows [rank*ceil(n/P), ...) (the last rank
 * may own fewer real rows; the library pads); for node2vec every rank uploads the whole graph
 * (row0 = 0, n_local = n: CSR and alias tables are replicated, the walk index space is sharded).  */
int gemb_graph_upload(gemb_ctx *ctx, int64_t n, int64_t row0, int64_t n_local,
                      const int32_t *indptr, const int32_t *indices, const float *data,
                      const int32_t *indptr_t, const int32_t *indices_t, const float *data_t,
                      gemb_graph **out);
int gemb_graph_free(gemb_graph *g);

/* Test hook for the dominant kernel:  Y = X0 + alpha * op(A) * X   (X0 may be NULL).
 * X is the full n x b block, X0 and Y are the n_local x b row shards; all HOST, row-major fp32.
 * b must be a multiple of 4. */
int gemb_spmm(gemb_graph *g, int transpose, int b, float alpha, const float *X, const float *X0,
              float *Y);

/* Test hook for the tensor-core contraction: G (b1 x b2, fp64, row-major) = P^T Q over n rows; P, Q host
 * fp32 row-major (Q == NULL means Q = P).  use_tensor_cores: 1 = tcgen05 kernel (GEMB_ERR_UNSUPPORTED if the
 * shape does not fit it), 0 = CUDA-core fp32 kernel. */
int gemb_gram(gemb_ctx *ctx, int64_t n, const float *P, int b1, const float *Q, int b2, int use_tensor_cores,
              double *G_out);

/* Test hook for the tall-skinny product Out (n x b2) = Q (n x b1) * M (b1 x b2), host fp32 row-major buffers.
 * use_tensor_cores: 1 = tcgen05 kernel, 0 = CUDA-core kernel. */
int gemb_apply(gemb_ctx *ctx, int64_t n, const float *Q, int b1, const float *M, int b2, int use_tensor_cores,
               float *Out);

/* ---- HOPE.  Replaces hope.py:29-36: S = (I - beta A)^-1 beta A is never formed; its top
 * k = d/2 singular triplets come from a block subspace iteration with Rayleigh-Ritz whose
 * operator applications are CSR SpMM sweeps (Katz/Horner sweeps of S and S^T in general; a Chebyshev
 * filter in A itself when A is symmetric, since then S = f(A)).  Output convention = the reference:
 * X = [U sqrt(Sigma) | V sqrt(Sigma)], sigma ASCENDING (scipy svds order). */
typedef struct {
    uint32_t struct_size;  /* = sizeof(gemb_hope_opts) */
    int32_t oversample;    /* extra block columns, default 16 (block b = min(n, d/2 + oversample)) */
    int32_t max_iters;     /* max subspace iterations, default 30 */
    int32_t min_iters;     /* default 2 */
    float tol;             /* stop when every one of the top k singular values moved by less than tol relative to
                              itself between two Rayleigh-Ritz rounds: max_j |sigma_j - sigma_j_prev| /
                              max(sigma_j, 1e-3 sigma_max) <= tol; default 1e-6 */
    int32_t katz_terms;    /* Horner terms J; 0 = choose so that (beta*||A||_2)^J <= katz_tol */
    float katz_tol;        /* default 1e-7 */
    uint64_t seed;         /* start block, default 1234 */
    int32_t compute_residual; /* 1: one extra Katz application to report ||S^T u - sigma v|| */
    int32_t verbose;
    int32_t algorithm;     /* 0 = auto (symmetric shard: 2, or 3 when the first Rayleigh-Ritz round shows a skewed
                              spectrum; else 1); 1 = subspace iteration on S^T S through Katz sweeps (any A);
                              2 = Chebyshev-filtered subspace iteration on A itself, S = f(A) (symmetric A only);
                              3 = thick-restart block Lanczos on A (symmetric A only; power-law spectra) */
    int32_t cheb_degree;   /* filter degree per outer iteration of algorithm 2, default 8 */
    float cheb_range_log2; /* algorithm 2: the filter degree is lowered until the filtered block's column dynamic range
                              T_m(x_L) stays below 2^cheb_range_log2 (fp32 Gram-based orthonormalisation squares it);
                              0 = default (8) */
    int32_t stop_rule;     /* 0 = relative change of every singular value <= tol (default);
                              1 = residual: max_j ||S v_j - sigma_j u_j|| / sigma_max <= tol over the top k, estimated
                                  from the Rayleigh-Ritz products (algorithm 2: |f'(l_j)| ||A v_j - l_j v_j||) */
    int32_t algorithm3_basis; /* algorithm 3 (thick-restart block Lanczos on A, symmetric A): maximum basis width,
                              0 = default (max(2k + 64, 192) rounded to the block) */
    int32_t spectral_mode; /* 0 = HOPE (top d/2 singular triplets of the Katz operator).
                              1 = the d largest ALGEBRAIC eigenpairs of the uploaded symmetric matrix itself, on the same
                              Chebyshev-filtered subspace iteration (beta ignored): X_out = n x d eigenvectors in DESCENDING
                              eigenvalue order, sigma_out = the d eigenvalues.  Laplacian Eigenmaps (lap.py:26-32:
                              eigs(normalized_laplacian, k = d+1, which='SM')) = this on D^-1/2 A D^-1/2 with d+1 pairs */
} gemb_hope_opts;

typedef struct {
    uint32_t struct_size;
    int32_t iters;         /* subspace iterations performed */
    int32_t katz_terms;    /* J actually used */
    int32_t block;         /* b */
    int32_t converged;
    int32_t algorithm;     /* 1 or 2: the solver that ran */
    int64_t spmm_count;    /* SpMM sweeps executed (all of width `block` except norm estimation) */
    double spmm_ms;        /* sum of CUDA-event durations of the block-width SpMM launches */
    double spmm_bytes;     /* algorithmic bytes of ONE block-width sweep: 8*nnz + 4*(n+1) + 8*n*b
                              (4*nnz less when the shard is unweighted) */
    double dense_ms;       /* Gram + apply + small factorizations */
    double comm_ms;        /* NCCL time (multi-GPU) */
    double total_ms;       /* device time of the whole call (events), excluding H2D/D2H */
    double h2d_ms, d2h_ms;
    float norm2_A;         /* estimated ||A||_2 */
    float ritz_change;     /* last max relative Ritz-value change */
    float resid_max;       /* max_j ||S^T u_j - sigma_j v_j|| / sigma_max (compute_residual=1) */
    float resid_est;       /* stop_rule = 1: the residual estimate the last round stopped on (else -1) */
    int32_t mg_mode;       /* 0 = single GPU; 1 = all-gather of the block per sweep; 2 = needed-rows-only exchange over
                              NVLink peer memory (CUDA IPC): rows are stored into the peers' halo slots by the kernel
                              that produces them; 3 = the same with fp16 halo copies: an experiment that did not pay
                              (GEMB_WIRE=fp16-experimental only; see hope.cu) */
    int64_t halo_rows;     /* mg_mode 2: distinct remote rows this shard references */
    int64_t push_rows;     /* mg_mode 2: (row, peer) pairs this rank stores per exchanged block */
    int64_t pushes;        /* mg_mode 2: blocks exchanged in this call (NVLink bytes out = pushes*push_rows*4*block) */
    float beta_used;       /* the beta the solve ran with (differs from the argument when that was negative) */
    double push_bytes;     /* mg_mode 2/3: bytes this rank stored into its peers over NVLink in this call */
} gemb_hope_stats;

/* X_out: n_local x d host buffer, or NULL to leave the result on the device (bench `value`).
 * sigma_out: d/2 floats (ascending) or NULL.
 * beta > 0: hope.py's beta.  beta < 0: beta = |beta| / ||A||_2, ||A||_2 (= rho(A) for symmetric A) estimated by power
 * iteration inside the call -- BASELINE.json configs[3] prescribes beta = 0.5 / rho_hat(A); stats->beta_used reports it. */
int gemb_hope(gemb_graph *g, int d, float beta, const gemb_hope_opts *opts, float *X_out,
              float *sigma_out, gemb_hope_stats *stats);

/* The diagnostic hope.py:38-40 prints: || U diag(s) V^T - S ||_F = || X1 X2^T - S ||_F with S = (I - beta A)^-1 beta A.
 * X: host, n x d row-major fp32 (the embedding gemb_hope returned).  S is never stored whole:
 *   n_probe <= 0 : exact -- S is applied to the identity in column panels (n sweeps' worth of work: meant for
 *                  n <= ~10^4, the sizes at which the reference can run at all);
 *   n_probe  > 0 : Hutchinson estimate sqrt(mean_j ||(X1 X2^T - S) z_j||^2) over n_probe Rademacher vectors
 *                  (SURVEY H8), relative standard error ~ sqrt(2 / n_probe).
 * Single GPU (the graph uploaded with row0 = 0, n_local = n). */
int gemb_hope_svd_error(gemb_graph *g, int d, float beta, const float *X, int n_probe, uint64_t seed,
                        double *err_out);

/* ---- node2vec.  Replaces the SNAP executable GEM shells out to (node2vec.py:31-48):
 * PreprocessTransitionProbs (bin@0x4127f0), node2vec() walks (bin@0x40c420),
 * LearnEmbeddings/TrainModel (bin@0x40ea30 / 0x40d6a0), WriteOutput + loadEmbedding
 * (graph_util.py:161-169).  p = q = 1 uses first-order tables (one per node); any other p, q > 0 builds the reference's
 * second-order tables -- one alias table per directed edge (t -> v) over v's out-neighbours, sum_(t->v) outdeg(v) entries
 * of 12 bytes -- on the device (GEMB_ERR_NOMEM with the size when they do not fit in HBM; the reference keeps the same
 * tables in host memory).
 *
 * The graph must be uploaded single-GPU style (row0 = 0, n_local = n) with every row's column
 * ids sorted ascending (SNAP adjacency order).  weights64: fp64 edge weights in CSR order or
 * NULL for 1.0 (the reference parses the "%f" text into doubles; alias tables are built in fp64
 * so that walks are bit-exact against oracle/n2v_oracle.c). */
int gemb_n2v_alias(gemb_graph *g, const double *weights64, int32_t *K_out /* nnz */,
                   double *U_out /* nnz */);

typedef struct {
    uint32_t struct_size;
    double alias_ms, shuffle_ms, walk_ms, vocab_ms, sgns_ms, total_ms, h2d_ms, d2h_ms, comm_ms;
    int64_t n_tokens;      /* vocabulary size V (includes the phantom token 0 if walks were padded) */
    int64_t n_walks;       /* walks generated by this rank */
    int64_t pairs;         /* (centre, context) pairs trained by this rank */
    double sgns_bytes;     /* algorithmic bytes: pairs * 14 rows * 4*d */
    double walk_bytes;     /* 24 B per transition */
} gemb_n2v_stats;

/* Walks only (parity hook).  nids: the N start nodes in SNAP node-table order (first appearance
 * in the edge list).  seed: the TRnd seed (the binary uses time(NULL)).  Walk w = i*N + j
 * (round i, shuffled position j) reads the Park-Miller stream at offset
 * (i+1)*(N-1) + w*(2*walk_len-3)  -- identical to the single-threaded binary whenever no walk
 * hits a dead end (oracle mode 1).  Walks [w_begin, w_end) are generated;
 * walks_out: (w_end-w_begin) x walk_len int32 host buffer (zero padded after a dead end). */
int gemb_n2v_walks(gemb_graph *g, const double *weights64, const int32_t *nids, int64_t N,
                   int walk_len, int num_walks, double p, double q, int32_t seed, int64_t w_begin,
                   int64_t w_end, int32_t *walks_out, gemb_n2v_stats *stats);

/* Full pipeline.  X_out: n_rows x d host fp32 (row = node id, like loadEmbedding; rows of ids
 * that never appear stay 0) or NULL.  sequential != 0 trains with ONE warp consuming the single
 * TRnd(seed) stream in program order (parity mode: follows oracle/n2v_oracle.c up to fp32
 * rounding); sequential == 0 is the Hogwild production mode. */
int gemb_node2vec(gemb_graph *g, const double *weights64, const int32_t *nids, int64_t N, int d,
                  int walk_len, int num_walks, int con_size, int max_iter, double p, double q,
                  int32_t seed, int sequential, int64_t n_rows, float *X_out,
                  gemb_n2v_stats *stats);

/* ---- reconstruction and its evaluation (the step after learn_embedding in tests/fit_model.py:10).
 * gemb_recon_create replaces the n^2 get_edge_weight calls of static_graph_embedding.py:48-65: A_hat = L R^T with a
 * zero diagonal, kept ON THE DEVICE (n x n fp32; GEMB_ERR_NOMEM with a message when it does not fit).
 *   split = 1: L = X[:, :d/2], R = X[:, d/2:]   (HOPE.get_edge_weight, hope.py:43-44)
 *   split = 0: L = R = X                        (node2vec.get_edge_weight, node2vec.py:56-57)
 * X: host, n x d row-major fp32. */
typedef struct gemb_recon gemb_recon;
int gemb_recon_create(gemb_ctx *ctx, const float *X, int64_t n, int d, int split, gemb_recon **out);
int gemb_recon_free(gemb_recon *r);

/* The dense matrix get_reconstructed_adj returns (static_graph_embedding.py:48-65): adj_out host, n x n row-major. */
int gemb_recon_dense(gemb_recon *r, float *adj_out);

/* A_hat[i[t]][j[t]] for m pairs (0 on the diagonal): the sampled-pairs branch of get_edge_list_from_adj_mtrx
 * (evaluation_util.py:25-28) and the weighted reconstruction error (evaluate_graph_reconstruction.py:37-40). */
int gemb_recon_pairs(gemb_recon *r, const int32_t *i, const int32_t *j, int64_t m, float *out);

/* computeMAP (metrics.py:28-46) without sorting.  The true graph is a CSR by node id (host, int32).  For every
 * true edge e = (i -> j): rank_out[e] = 1-based position of (i, j) in node i's predicted edges sorted by weight
 * (descending, ties in ascending j -- Python's stable sort), or 0 when (i, j) is not a predicted edge (j == i,
 * A_hat[i][j] <= 0, or j < i with is_undirected -- evaluation_util.py:29-35).  n_pred_row[i] = number of
 * predicted edges with source i.  AP and MAP follow from the ranks on the host. */
int gemb_recon_ranks(gemb_recon *r, const int32_t *indptr, const int32_t *indices, int is_undirected,
                     int32_t *rank_out, int32_t *n_pred_row);

/* computePrecisionCurve (metrics.py:6-25): the predicted edges that can be among the max_k heaviest (max_k < 0: all
 * of them), i.e. every valid entry >= the max_k-th largest value; UNORDERED.  *m_out = their number.  Call with
 * cap = 0 to get the count, then with cap >= count and three arrays of that length; the caller orders them
 * (weight descending, then i, then j ascending = the reference's stable sort of the row-major list). */
int gemb_recon_top(gemb_recon *r, int is_undirected, int64_t max_k, int64_t cap, int32_t *i_out, int32_t *j_out,
                   float *w_out, int64_t *m_out);

/* ---- wire formats: the reference's text files, read and written natively and in parallel.
 * HOST code only -- these entry points need no GPU.
 * Edge list: every non-blank line "src dst [weight]" (loadGraphFromEdgeListTxt, graph_util.py:143-158: exactly three
 * tokens -> float(weight), otherwise 1.0).  `skip` leading lines are ignored (the two header lines that
 * saveGraphToEdgeListTxt writes, graph_util.py:131-132).  scan counts the edge lines; parse fills caller arrays of
 * that length (w may be NULL); *all_unit = 1 when no weight differs from 1.0. */
int gemb_edge_list_scan(const char *path, int64_t skip, int64_t *n_edges);
int gemb_edge_list_parse(const char *path, int64_t skip, int64_t n_edges, int64_t *src, int64_t *dst, double *w,
                         int32_t *all_unit);
/* One "%d %d %f\n" per edge (saveGraphToEdgeListTxtn2v, graph_util.py:137-140); header_nodes >= 0 first writes
 * "<nodes>\n<edges>\n" (saveGraphToEdgeListTxt, :129-134).  w == NULL writes 1.000000. */
int gemb_edge_list_write(const char *path, int64_t n_edges, const int64_t *src, const int64_t *dst, const double *w,
                         int64_t header_nodes);
/* ".emb": "<rows> <d>" then "<id> v1 ... vd" (loadEmbedding, graph_util.py:161-169; written by SNAP's WriteOutput,
 * bin@0x406ef0, with 6 significant digits).  read: X == NULL returns only rows and d; otherwise X is rows x d fp64,
 * zeroed by the caller, row = id.  write: ids == NULL means rows 0..n_ids-1; X is indexed by id. */
int gemb_emb_read(const char *path, int64_t *rows, int32_t *d, double *X);
int gemb_emb_write(const char *path, int64_t n_ids, const int64_t *ids, int32_t d, const double *X, int64_t header_rows);

#ifdef __cplusplus
}
#endif
#endif /* GEMB200_H */
