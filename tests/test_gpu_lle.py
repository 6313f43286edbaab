"""Locally Linear Embedding parity on the GPU (gem_b200.embedding.lle.LocallyLinearEmbedding -> gemb_hope, spectral_mode 1 on
c I - M^T M) against the reference's golden tests/karate_res/LocallyLinearEmbedding.txt (allclose up to the sign of each vector),
tests/smb_res/LocallyLinearEmbedding.txt (the reference's bar |mean(target - X)| < 1e-3, tests/test_sbm.py:71-74,94), outputs of
the unmodified reference class (tests/golden/ref_lle_*.npz) and the pinned fp64 oracle (oracle/lle_oracle.py).
Tolerances (fp32 vectors vs fp64): singular values 2e-5 absolute, vectors 1e-4 after sign alignment where the value is simple."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import REPO, golden_path, load_karate_nx, load_sbm1024_nx

pytestmark = [pytest.mark.gpu, pytest.mark.filterwarnings('ignore::RuntimeWarning')]   # tol = 1e-9 runs to max_iters on purpose
sys.path.insert(0, os.path.join(REPO, 'oracle'))


def _fresh(**kw):
    from gem_b200.embedding.lle import LocallyLinearEmbedding
    LocallyLinearEmbedding.hyper_params.clear()
    LocallyLinearEmbedding.hyper_params.update({'method_name': 'lle_svd'})
    return LocallyLinearEmbedding(**kw)


def test_karate_golden(native_lib):
    import lap_oracle as lo
    G = load_karate_nx()
    m = _fresh(d=2, tol=1e-9, oversample=16, max_iters=120)
    X = m.learn_embedding(graph=G, is_weighted=True, no_python=True)
    gold = np.loadtxt(golden_path('karate_LocallyLinearEmbedding.txt'))
    assert X.shape == gold.shape
    assert np.allclose(lo.align_signs(X.astype(np.float64), gold), gold, atol=1e-4), np.abs(lo.align_signs(X.astype(np.float64), gold) - gold).max()
    assert m._s[0] < 2e-3            # sigma_0 = 0 exactly; sqrt of an fp32 eigenvalue difference resolves ~sqrt(c 1e-7)


@pytest.mark.parametrize('name,d', [('karate', 4), ('sbm1024', 16), ('randw120', 8)])
def test_reference_class_outputs(native_lib, name, d):
    import lle_oracle as le
    from gem_b200 import graph as hg
    z = np.load(golden_path('ref_lle_%s_d%d.npz' % (name, d)))
    e, n = z['edges'], int(z['n'])
    A = sp.csr_matrix((e[:, 2], (e[:, 0].astype(int), e[:, 1].astype(int))), shape=(n, n))
    m = _fresh(d=d, tol=1e-9, oversample=24, max_iters=150)
    X = m.learn_embedding(graph=hg.from_scipy(A)).astype(np.float64)
    ref = np.real(z['X'])
    Xo, s, V = le.lle_dense(A, d)
    assert np.allclose(m._s[1:] ** 2, s[1:] ** 2, atol=2e-5), np.abs(m._s ** 2 - s ** 2).max()     # eigenvalues of M^T M
    for j in range(d):
        lo_gap = s[j + 1] ** 2 - s[j] ** 2
        hi_gap = (s[j + 2] ** 2 - s[j + 1] ** 2) if j + 2 < len(s) else 1.0
        if min(lo_gap, hi_gap) > 1e-3:
            sg = np.sign(X[:, j] @ ref[:, j]) or 1.0
            assert np.abs(sg * X[:, j] - ref[:, j]).max() < 2e-4, (j, np.abs(sg * X[:, j] - ref[:, j]).max())
    Qr = np.linalg.qr(ref)[0]
    Xh = X[:, :max(1, d // 2)]
    assert np.linalg.norm(Xh - Qr @ (Qr.T @ Xh), 2) < 2e-3


def test_sbm1024_golden_d128(native_lib):
    import lle_oracle as le
    import networkx as nx
    S, _ = load_sbm1024_nx()
    m = _fresh(d=128, tol=1e-8, oversample=32, max_iters=150)
    X = m.learn_embedding(graph=S).astype(np.float64)
    gold = np.load(golden_path('sbm1024_LocallyLinearEmbedding.npy')).astype(np.float64)
    assert X.shape == gold.shape
    assert abs(np.mean(gold - X)) < 1e-3                                    # tests/test_sbm.py:94
    A = nx.to_scipy_sparse_array(S, nodelist=list(S.nodes), weight='weight', format='csr')
    Xo, s, V = le.lle_dense(A, 128)
    assert np.allclose(m._s[1:] ** 2, s[1:] ** 2, atol=5e-5), np.abs(m._s ** 2 - s ** 2).max()
    # residuals of the singular pairs against the fp64 operator: || M^T M x - s^2 x ||
    M = le.lle_matrix(A)
    R = M.T @ (M @ X) - X * (m._s[1:] ** 2)
    assert np.linalg.norm(R, axis=0).max() < 2e-4
