"""oracle/lle_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (NumPy / SciPy, fp64) of the reference's Locally Linear Embedding, gem/embedding/lle.py:21-35:
    graph = graph.to_undirected()                                    :25
    A = nx.to_scipy_sparse_matrix(graph); normalize(A, 'l1', axis=1) :26-27   P = D^-1 W (rows of zero stay zero)
    u, s, vt = scipy.sparse.linalg.svds(I - P, k=d+1, which='SM')    :28-30
    X = vt.T[:, 1:]                                                   :31-32   right singular vectors, ascending sigma, first dropped
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.  Pinned (tests/test_oracle_lle.py) against the
reference's goldens tests/karate_res/LocallyLinearEmbedding.txt, tests/smb_res/LocallyLinearEmbedding.txt and against outputs of the
unmodified reference class (tests/golden/ref_lle_*.npz, made by tests/golden/make_golden_lle.py with a harness-side shim for the
removed nx.to_scipy_sparse_matrix).
"""
import numpy as np
import scipy.sparse as sp

from lap_oracle import undirected_weights


def row_stochastic(W):
    """sklearn.preprocessing.normalize(W, norm='l1', axis=1): rows divided by the sum of their absolute values, zero rows kept."""
    W = sp.csr_matrix(W, dtype=np.float64)
    s = np.asarray(abs(W).sum(axis=1)).ravel()
    inv = np.where(s > 0, 1.0 / np.where(s > 0, s, 1.0), 0.0)
    return sp.csr_matrix(sp.diags(inv) @ W)


def lle_matrix(A):
    P = row_stochastic(undirected_weights(A))
    return sp.identity(P.shape[0], format='csr') - P


def lle_dense(A, d):
    """lle.py:25-32 with LAPACK's SVD (deterministic): (X n x d, sigma ascending (d+1), V n x (d+1))."""
    M = lle_matrix(A).toarray()
    u, s, vt = np.linalg.svd(M)
    idx = np.argsort(s)[:d + 1]
    s, v = s[idx], vt[idx].T
    return v[:, 1:], s, v
