"""Device R-MAT generator (gemb_synth_rmat; bench infrastructure for BASELINE.json configs[3]/[4]) against the
structural properties SURVEY 8(d) config 4 states and against the host generator gem_b200/synth.py::rmat of the same
family: symmetric, loop-free, duplicate-free, sorted rows, edge count and degree skew of a Graph500 R-MAT; shards are
slices of the whole; every call regenerates the same graph."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def gpu_ctx(native_lib):
    from gem_b200 import _native
    return _native.Context(0)


def test_rmat_structure_and_family(gpu_ctx):
    import scipy.sparse as sp
    from gem_b200 import _native, synth
    scale = 13
    n = 1 << scale
    ip, ix, tot = _native.synth_rmat(gpu_ctx, scale, seed=42)
    assert ip.shape == (n + 1,) and ip[0] == 0 and ip[-1] == tot == ix.shape[0]
    assert np.all(np.diff(ip) >= 0) and ix.min() >= 0 and ix.max() < n
    A = sp.csr_matrix((np.ones(tot, dtype=np.int8), ix, ip), shape=(n, n))
    assert (A != A.T).nnz == 0                                   # symmetric
    assert A.diagonal().sum() == 0                               # no self loops
    rows = np.repeat(np.arange(n), np.diff(ip))
    key = rows.astype(np.int64) * n + ix
    assert np.all(np.diff(key) > 0)                              # sorted, no duplicates
    assert tot <= 2 * 8 * n
    # same family as the host generator: edge count within 2 %, hub degree within a factor 1.5, isolated share within 3 points
    h = synth.rmat(scale=scale, seed=7)
    dh, dd = np.diff(h.indptr), np.diff(ip)
    assert abs(tot / h.nnz - 1) < 0.02, (tot, h.nnz)
    assert 1 / 1.5 < dd.max() / dh.max() < 1.5, (dd.max(), dh.max())
    assert abs((dd == 0).mean() - (dh == 0).mean()) < 0.03
    # deterministic, and a different seed gives a different graph
    ip2, ix2, _ = _native.synth_rmat(gpu_ctx, scale, seed=42)
    assert np.array_equal(ip, ip2) and np.array_equal(ix, ix2)
    ip3, ix3, _ = _native.synth_rmat(gpu_ctx, scale, seed=43)
    assert not (ip3.shape == ip.shape and np.array_equal(ix3[:1000], ix[:1000]))


def test_rmat_shards_are_slices(gpu_ctx):
    from gem_b200 import _native
    scale, world = 12, 3
    n = 1 << scale
    ip, ix, tot = _native.synth_rmat(gpu_ctx, scale, seed=5)
    per = (n + world - 1) // world
    got = 0
    for r in range(world):
        r0 = min(n, r * per)
        rows = min(n, r0 + per) - r0
        sip, six, stot = _native.synth_rmat(gpu_ctx, scale, seed=5, row0=r0, n_rows=rows)
        assert stot == tot
        assert np.array_equal(sip, ip[r0:r0 + rows + 1] - ip[r0])
        assert np.array_equal(six, ix[ip[r0]:ip[r0 + rows]])
        got += six.shape[0]
    assert got == tot


def test_hope_beta_over_rho(gpu_ctx):
    """beta_over_rho = c runs the solve with beta = c / rho_hat(A) (BASELINE configs[3]); same result as passing that beta."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    from gem_b200 import _native
    from gem_b200.graph import HostCSR
    from gem_b200.embedding.hope import HOPE
    ip, ix, tot = _native.synth_rmat(gpu_ctx, 12, seed=9)
    csr = HostCSR(1 << 12, ip, ix, None, symmetric=True)
    A = sp.csr_matrix((np.ones(tot), ix, ip), shape=(1 << 12, 1 << 12))
    rho = float(sla.eigsh(A, k=1, which='LA', return_eigenvectors=False)[0])
    HOPE.hyper_params.clear(); HOPE.hyper_params.update({'method_name': 'hope_gsvd'})
    m = HOPE(d=16, beta=0.0, beta_over_rho=0.5, tol=1e-6, max_iters=200, oversample=32, svd_error_probes=False)
    X = m.learn_embedding(graph=csr)
    assert abs(m._beta * rho / 0.5 - 1) < 5e-3, (m._beta, 0.5 / rho)     # power-iteration estimate of rho
    HOPE.hyper_params.clear(); HOPE.hyper_params.update({'method_name': 'hope_gsvd'})
    m2 = HOPE(d=16, beta=m._beta, tol=1e-6, max_iters=200, oversample=32, svd_error_probes=False)
    X2 = m2.learn_embedding(graph=csr)
    assert np.allclose(m._sigma, m2._sigma, rtol=1e-4)
