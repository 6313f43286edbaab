"""oracle/gf_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (NumPy, fp64) of the reference's Graph Factorization, gem/embedding/gf.py:92-105:
    X = 0.01 * np.random.randn(n, d)                                                  :94
    for _ in range(max_iter):                                                         :95
        for i, j, w in graph.edges(data='weight', default=1):  (skip j <= i)          :96-98
            X[i] -= eta * ( -(w - X[i].X[j]) * X[j] + regu * X[i] )                   :99-102
(the C++ twin gem/c_src/gf.cpp:143-164 is the same loop in float).  The start X0 is an argument here: the reference draws it from
the unseeded global NumPy RNG.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.  Pinned
(tests/test_oracle_gf.py) bit-for-bit against the unmodified reference class run with a seeded global RNG (tests/golden/ref_gf_*.npz).
"""
import numpy as np


def gf_sequential(n, src, dst, w, d, eta, regu, max_iter, X0):
    """The reference's sweep, statement by statement (edges in the order given)."""
    X = np.array(X0, dtype=np.float64, copy=True)
    src = np.asarray(src); dst = np.asarray(dst)
    w = np.ones(len(src)) if w is None else np.asarray(w, dtype=np.float64)
    for _ in range(max_iter):
        for i, j, ww in zip(src.tolist(), dst.tolist(), w.tolist()):
            if j <= i:
                continue
            term1 = -(ww - np.dot(X[i, :], X[j, :])) * X[j, :]
            term2 = regu * X[i, :]
            del_phi = term1 + term2
            X[i, :] -= eta * del_phi
    return X


def gf_rows_jacobi(n, src, dst, w, d, eta, regu, max_iter, X0):
    """gemb_gf mode 1: per epoch every row applies ITS edges in order to its running x_i, reading the partners from the previous
    epoch's table."""
    X = np.array(X0, dtype=np.float64, copy=True)
    src = np.asarray(src); dst = np.asarray(dst)
    w = np.ones(len(src)) if w is None else np.asarray(w, dtype=np.float64)
    for _ in range(max_iter):
        Xn = X.copy()
        for e in range(len(src)):
            i, j = int(src[e]), int(dst[e])
            if j <= i:
                continue
            xi = Xn[i]
            Xn[i] = xi - eta * (regu * xi - (w[e] - xi @ X[j]) * X[j])
        X = Xn
    return X
