/*
 * oracle/katz_omp.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The Katz operator of the reference's HOPE (gem/embedding/hope.py:29-31: S = (I - beta A)^-1 beta A) applied
 * matrix-free, y = sum_{j=1..J} (beta A)^j x by Horner, over a CSR in fp64 -- the same arithmetic as
 * oracle/hope_oracle.py::katz_apply (scipy.sparse products), with the row loop spread over the host cores by
 * OpenMP so that bench.py's reference arm can run scipy's svds (hope.py:33) on the full 1M-node configuration
 * with "all the host threads it can use".  tests/test_oracle_hope.py checks it against the scipy form.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 */
#include <stdint.h>
#include <string.h>
#include <omp.h>

/* out[r, :] = alpha * sum_c A[r, c] * W[c, :] + (x0 ? x0[r, :] : 0), m columns, row-major */
static void spmm_rows(int64_t n, const int64_t *indptr, const int32_t *idx, const double *w, double alpha, int m,
                      const double *W, const double *x0, double *out) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t r = 0; r < n; r++) {
        double *o = out + r * m;
        if (m == 1) {
            double acc = 0.0;
            for (int64_t i = indptr[r]; i < indptr[r + 1]; i++) acc += (w ? w[i] : 1.0) * W[idx[i]];
            o[0] = alpha * acc + (x0 ? x0[r] : 0.0);
        } else {
            for (int j = 0; j < m; j++) o[j] = 0.0;
            for (int64_t i = indptr[r]; i < indptr[r + 1]; i++) {
                const double v = w ? w[i] : 1.0;
                const double *src = W + (int64_t)idx[i] * m;
                for (int j = 0; j < m; j++) o[j] += v * src[j];
            }
            for (int j = 0; j < m; j++) o[j] = alpha * o[j] + (x0 ? x0[r * m + j] : 0.0);
        }
    }
}

/* y = sum_{j=1..J} (beta A)^j x ;  t1, t2: scratch n*m each.  W <- x; (J-1) times W <- x + beta A W; y = beta A W */
void katz_apply_omp(int64_t n, const int64_t *indptr, const int32_t *idx, const double *w, double beta, int J, int m,
                    const double *x, double *y, double *t1, double *t2) {
    const double *cur = x;
    for (int j = 1; j < J; j++) {
        double *dst = (j & 1) ? t1 : t2;
        spmm_rows(n, indptr, idx, w, beta, m, cur, x, dst);
        cur = dst;
    }
    spmm_rows(n, indptr, idx, w, beta, m, cur, (const double *)0, y);
}

int katz_omp_threads(void) { return omp_get_max_threads(); }
void katz_omp_set_threads(int t) { if (t > 0) omp_set_num_threads(t); }
