"""SpMM parity on the GPU: gemb_spmm (through the C ABI, host buffers) vs scipy.sparse on the same
seeded inputs.  fp32 accumulate in row order -> tolerance 2e-6 relative to the row's |a|.|x| bound."""
import numpy as np
import pytest

from conftest import load_karate_nx, load_sbm1024_nx

pytestmark = pytest.mark.gpu


def _check(ctx, csr, b, alpha, use_x0, transpose, seed=0):
    from gem_b200 import _native
    rng = np.random.default_rng(seed)
    t = csr.transpose()
    g = _native.DeviceGraph(ctx, csr.n, csr.indptr, csr.indices, csr.data_f32(), t.indptr, t.indices, t.data_f32())
    X = rng.standard_normal((csr.n, b)).astype(np.float32)
    X0 = rng.standard_normal((csr.n, b)).astype(np.float32) if use_x0 else None
    Y = g.spmm(X, alpha=alpha, X0=X0, transpose=transpose)
    g.free()
    A = csr.to_scipy().astype(np.float64)
    if transpose:
        A = A.T.tocsr()
    ref = alpha * (A @ X.astype(np.float64))
    if use_x0:
        ref = ref + X0
    bound = abs(alpha) * (abs(A) @ np.abs(X).astype(np.float64)) + (np.abs(X0) if use_x0 else 0) + 1e-30
    err = np.abs(Y - ref) / bound
    assert err.max() < 2e-6, err.max()


@pytest.mark.parametrize('b', [4, 8, 20, 80, 96, 144, 260])
def test_spmm_karate(gpu_ctx, b):
    from gem_b200 import graph as hg
    csr = hg.from_networkx(load_karate_nx())
    _check(gpu_ctx, csr, b, 0.01, True, False)
    _check(gpu_ctx, csr, b, 1.0, False, True)


def test_spmm_sbm1024_weighted(gpu_ctx):
    from gem_b200 import graph as hg
    G, _ = load_sbm1024_nx()
    csr = hg.from_networkx(G)
    rng = np.random.default_rng(1)
    csr.data = rng.uniform(0.1, 2.0, csr.nnz)                   # weighted variant
    for tr in (False, True):
        _check(gpu_ctx, csr, 80, 0.37, True, tr)
        _check(gpu_ctx, csr, 80, -1.5, False, tr)


def test_spmm_rmat_skewed_and_empty_rows(gpu_ctx):
    from gem_b200 import synth
    csr = synth.rmat(scale=12, edge_factor=8, seed=3)           # heavy skew + isolated nodes
    deg = np.diff(csr.indptr)
    assert deg.max() > 50 * max(1, np.median(deg)) and (deg == 0).any()
    _check(gpu_ctx, csr, 80, 0.5, True, False)
    _check(gpu_ctx, csr, 16, 1.0, False, False)
    # hubs of several thousand neighbours: rows above 128 nonzeros go through the chunked heavy-row kernels
    # (several 512-nonzero chunks per hub), weighted and unweighted, wide and narrow blocks, A and A^T
    big = synth.rmat(scale=15, edge_factor=8, seed=4)
    assert np.diff(big.indptr).max() > 4 * 512
    _check(gpu_ctx, big, 80, 0.5, True, False)
    _check(gpu_ctx, big, 8, -1.0, False, True)
    big.data = np.random.default_rng(9).uniform(0.1, 2.0, big.nnz)
    _check(gpu_ctx, big, 80, 0.25, True, True)
    _check(gpu_ctx, big, 144, 1.0, False, False)


def test_spmm_linearity_full_size_property(gpu_ctx):
    """Size-independent property at a size the CPU check would still finish: A(x+y) = Ax + Ay, and the
    Horner epilogue X0 + alpha*A*X is consistent with the plain product."""
    from gem_b200 import _native, synth
    csr = synth.sbm(n=200_000, block=1000, seed=5)
    g = _native.DeviceGraph(gpu_ctx, csr.n, csr.indptr, csr.indices, None)
    rng = np.random.default_rng(2)
    X = rng.standard_normal((csr.n, 80)).astype(np.float32)
    Z = rng.standard_normal((csr.n, 80)).astype(np.float32)
    a, b2, c = g.spmm(X), g.spmm(Z), g.spmm(X + Z)
    assert np.abs(c - (a + b2)).max() <= 2e-5 * np.abs(c).max()
    d = g.spmm(X, alpha=0.25, X0=Z)
    assert np.abs(d - (Z + 0.25 * a)).max() <= 2e-6 * np.abs(d).max()
    ref = csr.to_scipy() @ X.astype(np.float64)
    assert np.abs(a - ref).max() <= 1e-5 * np.abs(ref).max()
    g.free()
