"""oracle/lap_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (NumPy / SciPy, fp64) of the reference's Laplacian Eigenmaps, gem/embedding/lap.py:21-37:
    graph = graph.to_undirected()                                  :25
    l_sym = nx.normalized_laplacian_matrix(graph)                  :26   = D^-1/2 (D - A) D^-1/2, 1/sqrt(0) -> 0
    w, v  = scipy.sparse.linalg.eigs(l_sym, k=d+1, which='SM')    :28
    sort ascending; X = v[:, 1:]                                    :29-32
    'Laplacian matrix recon. error (low rank)' = ||V diag(w) V^T - L_sym||_F    :34-36
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.  Pinned (tests/test_oracle_lap.py) against the
reference's own goldens tests/karate_res/LaplacianEigenmaps.txt and tests/smb_res/LaplacianEigenmaps.txt and against outputs of the
unmodified reference class (tests/golden/ref_lap_*.npz, made by tests/golden/make_golden_lap.py).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def undirected_weights(A):
    """nx.DiGraph.to_undirected() on the adjacency matrix (rows in list(graph.nodes) order): the pair {u, v} exists when either
    direction does; when both do, networkx keeps the attributes of the direction it copies LAST -- it walks the nodes in
    order and, for each, its successors, so the edge out of the later node wins: W[u, v] = A[max(u,v), min(u,v)] if that
    entry exists, else A[min(u,v), max(u,v)]."""
    A = sp.csr_matrix(A, dtype=np.float64)
    n = A.shape[0]
    lower = sp.tril(A, -1, format='csr')          # entries (r, c) with r > c: out of the later node
    upper = sp.triu(A, 1, format='csr')
    # pairs present in `lower` take its weight; the others take the upper entry
    has_lower = lower.copy(); has_lower.data[:] = 1.0
    up_only = upper - upper.multiply(has_lower.T)          # upper entries whose mirror does not exist
    up_only.eliminate_zeros()
    # explicit zeros cannot be told from absent entries after the subtraction; weights of 0 are not edges in the reference's fixtures
    half = lower + up_only.T                               # strictly lower triangle of W
    W = half + half.T + sp.diags(A.diagonal())
    return sp.csr_matrix(W)


def normalized_laplacian(W):
    """nx.normalized_laplacian_matrix: D^-1/2 (D - W) D^-1/2 with D = row sums and 1/sqrt(0) = 0."""
    W = sp.csr_matrix(W, dtype=np.float64)
    deg = np.asarray(W.sum(axis=1)).ravel()
    with np.errstate(divide='ignore'):
        dh = 1.0 / np.sqrt(deg)
    dh[np.isinf(dh)] = 0.0
    DH = sp.diags(dh)
    return sp.csr_matrix(DH @ (sp.diags(deg) - W) @ DH)


def lap_dense(A, d):
    """lap.py:25-37 with LAPACK instead of ARPACK (deterministic): (X n x d, w ascending (d+1), V n x (d+1), recon error)."""
    L = normalized_laplacian(undirected_weights(A)).toarray()
    w, v = np.linalg.eigh(L)
    w, v = w[:d + 1], v[:, :d + 1]
    err = float(np.linalg.norm(v @ np.diag(w) @ v.T - L))
    return v[:, 1:], w, v, err


def lap_sparse(A, d, tol=0):
    """The sparse route for sizes where dense eigh does not fit: the d+1 LARGEST eigenpairs of 2I - L_sym (same vectors,
    w = 2 - theta), Lanczos without shift-invert."""
    L = normalized_laplacian(undirected_weights(A))
    n = L.shape[0]
    M = 2.0 * sp.identity(n, format='csr') - L
    th, v = spla.eigsh(M, k=d + 1, which='LA', tol=tol)
    w = 2.0 - th
    idx = np.argsort(w)
    w, v = w[idx], v[:, idx]
    return v[:, 1:], w, v


def align_signs(X, ref):
    """Eigenvectors are defined up to sign: flip each column of X towards ref."""
    s = np.sign(np.sum(X * ref, axis=0))
    s[s == 0] = 1.0
    return X * s
