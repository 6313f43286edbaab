"""node2vec parity on the GPU (through the C ABI / plugin class) against oracle/n2v_oracle.c, which
tests/test_oracle_n2v.py pins to the unmodified reference binary.
  * alias tables (K int32, U fp64) and walk matrices: BIT-EXACT.
  * SGNS, sequential parity mode (one warp, the binary's single RNG stream): follows the fp64 oracle up
    to fp32 rounding (tolerance stated per test).
  * SGNS, Hogwild production mode: statistical parity (the reference's own bar, tests/test_karate.py:78 /
    tests/test_sbm.py:94, plus community purity of nearest neighbours vs the oracle run)."""
import numpy as np
import pytest

from conftest import golden_path, load_karate_nx, load_sbm1024_nx, nx_from_npz

pytestmark = pytest.mark.gpu


def _inputs(G):
    from gem_b200 import graph as hg
    return hg.n2v_inputs_from_networkx(G)


def _dev(ctx, csr):
    from gem_b200 import _native
    return _native.DeviceGraph(ctx, csr.n, csr.indptr, csr.indices, None)


def _graphs():
    from gem_b200 import graph as hg, synth
    out = {}
    out['karate'] = _inputs(load_karate_nx())
    out['symw60'] = _inputs(nx_from_npz(np.load(golden_path('n2v_bin_symw60.npz'))))
    out['dirw50'] = _inputs(nx_from_npz(np.load(golden_path('n2v_bin_dirw50_pq.npz'))))
    G, _ = load_sbm1024_nx()
    out['sbm1024'] = _inputs(G)
    r = synth.rmat(scale=11, edge_factor=8, seed=9)               # skewed degrees, isolated ids
    rows = np.repeat(np.arange(r.n), np.diff(r.indptr))
    rng = np.random.default_rng(4)
    w = np.round(rng.uniform(0.05, 4.0, r.nnz), 6)
    out['rmat11w'] = hg.n2v_inputs_from_edges(rows, r.indices, w)
    return out


@pytest.fixture(scope='module')
def graphs():
    return _graphs()


@pytest.mark.parametrize('name', ['karate', 'symw60', 'dirw50', 'sbm1024', 'rmat11w'])
def test_alias_tables_bit_exact(gpu_ctx, n2v_oracle, graphs, name):
    csr, nids = graphs[name]
    g = _dev(gpu_ctx, csr)
    K, U = g.n2v_alias(csr.data)
    g.free()
    Ko, Uo = n2v_oracle.alias_first_order(csr.indptr, csr.data)
    assert np.array_equal(K, Ko)
    assert np.array_equal(U.view(np.int64), Uo.view(np.int64))    # bit pattern of every fp64 threshold


@pytest.mark.parametrize('name,walk_len,num_walks,seed', [
    ('karate', 10, 2, 7), ('karate', 80, 10, 1234), ('symw60', 15, 3, 4242), ('dirw50', 12, 4, 99),
    ('sbm1024', 80, 10, 1), ('sbm1024', 2, 1, 5), ('sbm1024', 1, 2, 5), ('rmat11w', 40, 3, 2147483646)])
def test_walks_bit_exact(gpu_ctx, n2v_oracle, graphs, name, walk_len, num_walks, seed):
    csr, nids = graphs[name]
    g = _dev(gpu_ctx, csr)
    W, st = g.n2v_walks(nids, walk_len, num_walks, seed=seed, weights64=csr.data)
    Wo = n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, walk_len, num_walks, seed=seed, mode=1)
    assert W.shape == Wo.shape and np.array_equal(W, Wo)
    # a shard of the walk index space is the same slice (multi-GPU walk sharding)
    tot = len(nids) * num_walks
    a, b = tot // 3, (2 * tot) // 3 + 1
    Ws, _ = g.n2v_walks(nids, walk_len, num_walks, seed=seed, weights64=csr.data, w_begin=a, w_end=b)
    assert np.array_equal(Ws, Wo[a:b])
    g.free()
    # every walk starts at each node exactly once per round, follows edges, pads with 0 after a dead end
    starts = W[:, 0].reshape(num_walks, len(nids))
    assert all(sorted(r.tolist()) == sorted(nids.tolist()) for r in starts)


def test_walks_equal_the_reference_binary_stream_when_no_dead_ends(gpu_ctx, n2v_oracle, graphs):
    csr, nids = graphs['sbm1024']                                   # symmetric: no dead ends
    g = _dev(gpu_ctx, csr)
    W, _ = g.n2v_walks(nids, 30, 3, seed=77)
    g.free()
    Wseq = n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, 30, 3, seed=77, mode=0)
    assert np.array_equal(W, Wseq)                                  # = single-threaded binary's walks


@pytest.mark.parametrize('case', ['karate_b', 'symw60', 'sbm128', 'offset40'])
def test_sgns_sequential_follows_oracle_and_binary(gpu_ctx, n2v_oracle, case):
    z = np.load(golden_path('n2v_bin_%s.npz' % case))
    G = nx_from_npz(z)
    csr, nids = _inputs(G)
    hp = {k: z[k].item() for k in ('d', 'walk_len', 'num_walks', 'con_size', 'max_iter', 'seed')}
    g = _dev(gpu_ctx, csr)
    X, st = g.node2vec(nids, hp['d'], hp['walk_len'], hp['num_walks'], hp['con_size'], hp['max_iter'],
                       seed=hp['seed'], sequential=True, weights64=csr.data)
    g.free()
    # oracle in the GPU's stream convention (mode 1 walks), fp64
    Wo = n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, hp['walk_len'], hp['num_walks'], seed=hp['seed'], mode=1)
    Xo, tok = n2v_oracle.learn(Wo, csr.n, hp['d'], hp['con_size'], hp['max_iter'], hp['seed'])
    assert st['n_tokens'] == len(tok)
    scale = np.abs(Xo).max()
    assert np.abs(X - Xo).max() < 2e-3 * scale, np.abs(X - Xo).max() / scale     # fp32 vs fp64 trajectory
    assert np.corrcoef(X.ravel(), Xo.ravel())[0, 1] > 0.99999
    if case != 'karate_b':        # no dead ends (or mode-independent) -> also equals the binary's output
        ids, emb = z['ids'], z['emb']
        if case in ('symw60', 'sbm128'):
            assert np.abs(X[ids] - emb).max() < 2e-3 * np.abs(emb).max()


def test_hogwild_statistical_parity_karate(gpu_ctx):
    from gem_b200.embedding.node2vec import node2vec
    node2vec.hyper_params.clear(); node2vec.hyper_params.update({'method_name': 'node2vec_rw'})
    G = load_karate_nx()
    m = node2vec(d=2, max_iter=1, walk_len=80, num_walks=10, con_size=10, ret_p=1, inout_p=1)
    X = m.learn_embedding(graph=G, is_weighted=True, no_python=True)
    target = np.loadtxt(golden_path('karate_node2vec.txt'))
    assert X.shape == target.shape == (34, 2)
    assert abs(np.mean(target - X)) < .3                            # tests/test_karate.py:78
    assert np.isfinite(X).all() and np.abs(X).max() < 50


def _purity(X, labels, k=10):
    Xn = X / (np.linalg.norm(X, axis=1, keepdims=True) + 1e-12)
    S = Xn @ Xn.T
    np.fill_diagonal(S, -np.inf)
    nn = np.argsort(-S, axis=1)[:, :k]
    return float(np.mean(labels[nn] == labels[:, None]))


def test_hogwild_statistical_parity_sbm1024(gpu_ctx, n2v_oracle):
    from gem_b200.embedding.node2vec import node2vec
    node2vec.hyper_params.clear(); node2vec.hyper_params.update({'method_name': 'node2vec_rw'})
    G, z = load_sbm1024_nx()
    labels = z['labels'].astype(np.int64)
    hp = dict(d=32, max_iter=1, walk_len=40, num_walks=5, con_size=5, ret_p=1, inout_p=1)
    m = node2vec(seed=11, **hp)
    X = m.learn_embedding(graph=G)
    csr, nids = _inputs(G)
    Wo = n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, 40, 5, seed=11, mode=1)
    Xo, _ = n2v_oracle.learn(Wo, csr.n, 32, 5, 1, 11)
    assert m.stats['n_tokens'] == 1024 and m.stats['pairs'] > 0
    # the reference's bar against its own golden (different d there, so compare with the oracle run)
    assert abs(np.mean(Xo - X)) < 0.1                               # tests/test_sbm.py:76-79,94
    pg, po = _purity(np.asarray(X, np.float64), labels), _purity(Xo, labels)
    assert po > 0.8 and pg > po - 0.05, (pg, po)
    # same second-moment scale of the learned vectors
    assert 0.7 < np.linalg.norm(X) / np.linalg.norm(Xo) < 1.4


@pytest.mark.parametrize('name,walk_len,num_walks,p,q,seed', [
    ('karate', 20, 3, 0.5, 2.0, 7), ('symw60', 15, 3, 4.0, 0.25, 4242), ('dirw50', 12, 4, 0.7, 1.0, 99),
    ('sbm1024', 40, 2, 1.0, 0.5, 3), ('rmat11w', 30, 2, 2.0, 3.0, 2147483646)])
def test_second_order_walks_bit_exact(gpu_ctx, n2v_oracle, graphs, name, walk_len, num_walks, p, q, seed):
    """p, q != 1 (node2vec.py:22-23,42-43; PreprocessNode bin@0x411f40): one alias table per directed edge (t -> v),
    built on the device in the oracle's fp64 operation order -> walks equal the oracle's, which is pinned to the
    unmodified binary for second-order settings too (tests/test_oracle_n2v.py, n2v_bin_*_pq.npz)."""
    csr, nids = graphs[name]
    g = _dev(gpu_ctx, csr)
    W, st = g.n2v_walks(nids, walk_len, num_walks, p=p, q=q, seed=seed, weights64=csr.data)
    Wo = n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, walk_len, num_walks, p=p, q=q, seed=seed, mode=1)
    assert W.shape == Wo.shape and np.array_equal(W, Wo)
    W1 = n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, walk_len, num_walks, seed=seed, mode=1)
    assert not np.array_equal(Wo, W1)                               # the bias does change the walks
    tot = len(nids) * num_walks
    a, b = tot // 4, tot // 2 + 1
    Ws, _ = g.n2v_walks(nids, walk_len, num_walks, p=p, q=q, seed=seed, weights64=csr.data, w_begin=a, w_end=b)
    g.free()
    assert np.array_equal(Ws, Wo[a:b])


@pytest.mark.parametrize('case', ['dirw50_pq', 'symw60_pq'])
def test_second_order_pipeline_follows_oracle_and_binary(gpu_ctx, n2v_oracle, case):
    """The reference binary's own output for p, q != 1 (goldens made with the time() shim, OMP_NUM_THREADS=1):
    the GPU's sequential-mode embedding follows the fp64 oracle run on the same walks, and -- on the graph without dead
    ends, where the GPU's stream offsets are the binary's -- the binary's printed embedding itself."""
    z = np.load(golden_path('n2v_bin_%s.npz' % case))
    G = nx_from_npz(z)
    csr, nids = _inputs(G)
    hp = {k: z[k].item() for k in ('d', 'walk_len', 'num_walks', 'con_size', 'max_iter', 'seed', 'p', 'q')}
    g = _dev(gpu_ctx, csr)
    X, st = g.node2vec(nids, hp['d'], hp['walk_len'], hp['num_walks'], hp['con_size'], hp['max_iter'], p=hp['p'], q=hp['q'],
                       seed=hp['seed'], sequential=True, weights64=csr.data)
    g.free()
    Wo = n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, hp['walk_len'], hp['num_walks'], p=hp['p'], q=hp['q'],
                          seed=hp['seed'], mode=1)
    Xo, tok = n2v_oracle.learn(Wo, csr.n, hp['d'], hp['con_size'], hp['max_iter'], hp['seed'])
    assert st['n_tokens'] == len(tok)
    assert np.abs(X - Xo).max() < 2e-3 * np.abs(Xo).max()
    if case == 'symw60_pq':
        ids, emb = z['ids'], z['emb']
        assert np.abs(X[ids] - emb).max() < 2e-3 * np.abs(emb).max()


def test_second_order_through_the_plugin_class(gpu_ctx):
    from gem_b200.embedding.node2vec import node2vec
    node2vec.hyper_params.clear(); node2vec.hyper_params.update({'method_name': 'node2vec_rw'})
    m = node2vec(d=4, max_iter=1, walk_len=10, num_walks=2, con_size=3, ret_p=0.5, inout_p=2.0)
    X = m.learn_embedding(graph=load_karate_nx())
    assert X.shape == (34, 4) and np.isfinite(X).all() and np.abs(X).sum() > 0


def test_phantom_zero_and_dead_ends(gpu_ctx, n2v_oracle):
    import networkx as nx
    G = nx.DiGraph([(5, 6), (6, 7), (7, 8)])
    csr, nids = _inputs(G)
    g = _dev(gpu_ctx, csr)
    W, _ = g.n2v_walks(nids, 6, 2, seed=3)
    assert np.array_equal(W, n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, 6, 2, seed=3, mode=1))
    X, st = g.node2vec(nids, 4, 6, 2, 2, 1, seed=3, sequential=True)
    g.free()
    assert st['n_tokens'] == 5 and X.shape == (9, 4)               # phantom token 0
    assert np.all(X[[1, 2, 3, 4]] == 0) and np.abs(X[[0, 5, 6, 7, 8]]).sum(axis=1).min() > 0
