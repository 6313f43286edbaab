"""oracle/lap_oracle.py pinned against the reference: its own goldens (tests/karate_res/LaplacianEigenmaps.txt,
tests/smb_res/LaplacianEigenmaps.txt) and outputs of the unmodified class gem.embedding.lap.LaplacianEigenmaps
(tests/golden/ref_lap_*.npz, made by tests/golden/make_golden_lap.py in the build container).  CPU only."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import REPO, golden_path, load_karate_nx, load_sbm1024_nx

sys.path.insert(0, os.path.join(REPO, 'oracle'))
import lap_oracle as lo
import hope_oracle as ho


def _adj_from_npz(z):
    e, n = z['edges'], int(z['n'])
    return sp.csr_matrix((e[:, 2], (e[:, 0].astype(int), e[:, 1].astype(int))), shape=(n, n))


@pytest.mark.parametrize('name,d', [('karate', 2), ('karate', 4), ('sbm1024', 16), ('randw120', 8)])
def test_reference_class_outputs(name, d):
    z = np.load(golden_path('ref_lap_%s_d%d.npz' % (name, d)))
    X, w, V, err = lo.lap_dense(_adj_from_npz(z), d)
    ref = np.real(z['X'])
    assert ref.shape == X.shape
    # eigenvectors up to sign; ARPACK converged to machine precision on these sizes.  Degenerate eigenvalues (if any) are
    # compared as subspaces.
    gaps = np.diff(w)
    if gaps.min() > 1e-6:
        assert np.allclose(lo.align_signs(X, ref), ref, atol=1e-7), np.abs(lo.align_signs(X, ref) - ref).max()
    assert ho.principal_angles_deg(X, ref)[0] < 1e-4


def test_reference_goldens_karate_and_sbm():
    import networkx as nx
    G = load_karate_nx()
    A = nx.to_scipy_sparse_array(G, nodelist=list(G.nodes), weight='weight', format='csr')
    X, w, V, err = lo.lap_dense(A, 2)
    gold = np.loadtxt(golden_path('karate_LaplacianEigenmaps.txt'))
    assert np.allclose(lo.align_signs(X, gold), gold, atol=1e-8)
    assert abs(w[0]) < 1e-12                                   # connected graph: the dropped eigenvector is the trivial one
    # SBM fixture, d = 128: the reference's own bar is |mean(target - X)| < 1e-3 (tests/test_sbm.py:94); the subspace is compared too
    S, _ = load_sbm1024_nx()
    As = nx.to_scipy_sparse_array(S, nodelist=list(S.nodes), weight='weight', format='csr')
    Xs, ws, Vs, errs = lo.lap_dense(As, 128)
    golds = np.load(golden_path('sbm1024_LaplacianEigenmaps.npy')).astype(np.float64)
    assert abs(np.mean(golds - Xs)) < 1e-3
    assert ho.principal_angles_deg(Xs[:, :8], golds[:, :8])[0] < 1e-2      # the well separated leading part
    Xsp, wsp, _ = lo.lap_sparse(As, 16)
    assert np.allclose(wsp, ws[:17], atol=1e-9)


def test_to_undirected_rule_matches_networkx():
    import networkx as nx
    rng = np.random.default_rng(3)
    G = nx.DiGraph()
    G.add_nodes_from(range(40))
    for _ in range(300):
        u, v = int(rng.integers(0, 40)), int(rng.integers(0, 40))
        if u != v:
            G.add_edge(u, v, weight=float(rng.uniform(0.5, 2.0)))
    A = nx.to_scipy_sparse_array(G, nodelist=list(G.nodes), weight='weight', format='csr')
    W = lo.undirected_weights(A).toarray()
    Wn = nx.to_numpy_array(G.to_undirected(), nodelist=list(G.nodes), weight='weight')
    assert np.array_equal(W, Wn)
    L = lo.normalized_laplacian(W).toarray()
    Ln = nx.normalized_laplacian_matrix(G.to_undirected(), nodelist=list(G.nodes)).toarray()
    assert np.allclose(L, Ln, atol=1e-14)
