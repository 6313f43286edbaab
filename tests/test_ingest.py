"""Native wire-format readers / writers (gem_b200/csrc/ingest.cu, host code in libgemb200.so)
against the per-line Python of the reference's gem/utils/graph_util.py:129-169 (restated in
gem_b200/utils/graph_util.py and used here as the checker): same bytes out, same values in."""
import os

import numpy as np
import pytest


def _py_load_embedding(file_name):            # graph_util.py:161-169, verbatim loop
    with open(file_name, 'r') as f:
        n, d = f.readline().strip().split()
        X = np.zeros((int(n), int(d)))
        for line in f:
            emb = line.strip().split()
            X[int(emb[0]), :] = [float(e) for e in emb[1:]]
    return X


def _random_digraph(rng, n, m, weighted):
    import networkx as nx
    G = nx.DiGraph()
    for _ in range(m):
        u, v = (int(x) for x in rng.integers(0, n, 2))
        if weighted:
            G.add_edge(u, v, weight=float(rng.choice([rng.uniform(-3, 3), rng.uniform(0, 1e-4), rng.integers(0, 1000), 1.0])))
        else:
            G.add_edge(u, v)
    return G


@pytest.mark.parametrize('weighted', [False, True])
def test_edge_list_round_trip_matches_reference_functions(native_lib, tmp_path, weighted):
    from gem_b200 import graph as hg
    from gem_b200.utils import graph_util as gu
    rng = np.random.default_rng(11)
    G = _random_digraph(rng, 300, 4000, weighted)
    a, b = str(tmp_path / 'ref.txt'), str(tmp_path / 'ref_n2v.txt')
    gu.saveGraphToEdgeListTxt(G, a)                       # per-line Python, byte-for-byte the reference's writer
    gu.saveGraphToEdgeListTxtn2v(G, b)
    # reader: same graph as the reference's loader builds (node order, weights after the %f round trip)
    G2 = gu.loadGraphFromEdgeListTxt(b, directed=True)
    want = hg.from_networkx(G2)                           # rows = list(G2.nodes)
    got = gu.loadEdgeListCSR(b, node_order='appearance')
    assert got.nodes == want.nodes and np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert (got.data is None and want.data is None) or np.array_equal(got.data, want.data)
    by_id = gu.loadEdgeListCSR(a, skip_header=2)
    want_id = hg.from_networkx(G2, by_label=True)
    assert by_id.n == want_id.n and np.array_equal(by_id.indptr, want_id.indptr) and np.array_equal(by_id.indices, want_id.indices)
    assert (by_id.data is None and want_id.data is None) or np.array_equal(by_id.data, want_id.data)
    # writer: the same bytes as the reference's writers for a graph whose edges are in row-major order
    H = hg.from_networkx(G, by_label=True)
    import networkx as nx
    R = nx.DiGraph()
    R.add_nodes_from(range(H.n))
    rows = np.repeat(np.arange(H.n), np.diff(H.indptr))
    for t, (u, v) in enumerate(zip(rows.tolist(), H.indices.tolist())):
        R.add_edge(u, v, weight=1.0 if H.data is None else float(H.data[t]))
    c, d = str(tmp_path / 'py.txt'), str(tmp_path / 'nat.txt')
    gu.saveGraphToEdgeListTxt(R, c); gu.saveEdgeListCSR(H, d)
    assert open(c, 'rb').read() == open(d, 'rb').read()
    gu.saveGraphToEdgeListTxtn2v(R, c); gu.saveEdgeListCSR(H, d, n2v=True)
    assert open(c, 'rb').read() == open(d, 'rb').read()


def test_edge_list_token_rules_and_errors(native_lib, tmp_path):
    from gem_b200.utils import graph_util as gu
    f = str(tmp_path / 'odd.txt')
    with open(f, 'w') as fh:
        fh.write('0 1\n\n  2\t3   0.5  \r\n4 5 1e-3\n6 7 2.5 extra tokens\n8 9 -0.000001\n10 11 .5\n12 13 123456789012345678\n14 15 inf')
    src, dst, w = gu.readEdgeList(f)
    assert src.tolist() == [0, 2, 4, 6, 8, 10, 12, 14] and dst.tolist() == [1, 3, 5, 7, 9, 11, 13, 15]
    # 2 tokens -> 1.0; 3 tokens -> float(); more than 3 -> 1.0 (graph_util.py:151-154)
    assert w.tolist() == [1.0, 0.5, 1e-3, 1.0, -0.000001, 0.5, float('123456789012345678'), float('inf')]
    G = gu.loadGraphFromEdgeListTxt(f)
    assert [G[u][v]['weight'] for u, v in zip(src.tolist(), dst.tolist())] == w.tolist()
    und = gu.loadEdgeListCSR(f, directed=False)
    assert und.is_symmetric() and und.n == 16
    for bad in ('0\n', '0 x\n', '1 2 abc\n'):
        with open(f, 'w') as fh:
            fh.write('0 1\n' + bad)
        with pytest.raises(RuntimeError, match='malformed line'):
            gu.readEdgeList(f)
    open(f, 'w').close()
    src, dst, w = gu.readEdgeList(f)
    assert src.size == 0 and w is None
    with pytest.raises(RuntimeError, match='cannot open'):
        gu.readEdgeList(str(tmp_path / 'missing.txt'))


def test_fast_float_path_is_exact(native_lib, tmp_path):
    """Every '%f' / '%g' / repr rendering of random doubles parses to exactly float(token)."""
    from gem_b200.utils import graph_util as gu
    rng = np.random.default_rng(5)
    vals = np.concatenate((rng.uniform(-1e6, 1e6, 20000), rng.uniform(0, 1, 20000) ** 8, 10.0 ** rng.uniform(-12, 12, 20000)))
    toks = ['%f' % v for v in vals] + ['%g' % v for v in vals] + [repr(float(v)) for v in vals] + ['%.15f' % v for v in vals[:5000]]
    f = str(tmp_path / 'floats.txt')
    with open(f, 'w') as fh:
        fh.write(''.join('%d %d %s\n' % (i, i + 1, t) for i, t in enumerate(toks)))
    _, _, w = gu.readEdgeList(f)
    assert w.tolist() == [float(t) for t in toks]


def test_embedding_files(native_lib, tmp_path):
    from gem_b200.utils import graph_util as gu
    rng = np.random.default_rng(2)
    X = rng.standard_normal((257, 12)) * 10.0 ** rng.integers(-6, 6, (257, 12))
    e = str(tmp_path / 'x.emb')
    gu.saveEmbedding(X, e)                                   # native writer
    lines = open(e).read().split('\n')
    assert lines[0] == '257 12' and lines[1] == '0 ' + ' '.join('%g' % v for v in X[0])      # SNAP's 6 digits
    assert lines[257] == '256 ' + ' '.join('%g' % v for v in X[256]) and lines[258] == ''
    Xn = gu.loadEmbedding(e)                                 # native reader
    assert np.array_equal(Xn, _py_load_embedding(e))         # == the reference's loop on the same file
    assert np.allclose(Xn, X, rtol=1e-5)
    ids = rng.permutation(257)[:100]
    gu.saveEmbedding(X, e, ids=ids.tolist())                 # SNAP writes only the nodes it saw, in its own order
    Xp = gu.loadEmbedding(e)
    assert np.array_equal(Xp, _py_load_embedding(e))
    mask = np.zeros(257, dtype=bool); mask[ids] = True
    assert np.all(Xp[~mask] == 0) and np.allclose(Xp[mask], X[mask], rtol=1e-5)
    with open(e, 'w') as fh:
        fh.write('3 2\n0 1 2\n5 1 2\n')
    with pytest.raises(RuntimeError, match='malformed line'):
        gu.loadEmbedding(e)


def test_binary_csr_round_trip(tmp_path):
    from gem_b200 import synth
    from gem_b200.utils import graph_util as gu
    csr = synth.sbm(n=5000, block=500, seed=1)
    f = str(tmp_path / 'g.npz')
    gu.saveCSR(csr, f)
    back = gu.loadCSR(f)
    assert back.n == csr.n and np.array_equal(back.indptr, csr.indptr) and np.array_equal(back.indices, csr.indices)
    assert back.data is None and back.symmetric is True
