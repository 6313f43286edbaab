"""Pins oracle/n2v_oracle.c against the UNMODIFIED reference binary gem/c_exe/node2vec: the goldens in
tests/golden/n2v_bin_*.npz were produced by that binary with time() interposed (seed = N2V_FAKE_TIME)
and OMP_NUM_THREADS=1 (tests/golden/make_golden.py).  The oracle must reproduce every embedding to the
6 significant digits the binary prints -- which it can only do if its alias tables, shuffles, walks,
vocabulary, negative sampler and SGD updates are all identical.  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import golden_path, load_karate_nx, nx_from_npz

CASES = sorted(os.path.basename(p)[8:-4] for p in glob.glob(golden_path('n2v_bin_*.npz')))


def oracle_inputs(G):
    from gem_b200 import graph as hg
    csr, nids = hg.n2v_inputs_from_networkx(G)
    return csr, nids


@pytest.mark.parametrize('name', CASES)
def test_oracle_reproduces_reference_binary(n2v_oracle, name):
    z = np.load(golden_path('n2v_bin_%s.npz' % name))
    G = nx_from_npz(z)
    csr, nids = oracle_inputs(G)
    hp = {k: z[k].item() for k in ('d', 'walk_len', 'num_walks', 'con_size', 'max_iter', 'p', 'q', 'seed')}
    X, tok, wm = n2v_oracle.node2vec(csr.indptr, csr.indices, csr.data, nids, hp['d'], hp['walk_len'],
                                     hp['num_walks'], hp['con_size'], hp['max_iter'], hp['p'], hp['q'],
                                     hp['seed'], mode=0)
    ids, emb = z['ids'], z['emb']
    assert sorted(ids.tolist()) == sorted(tok.tolist())        # same vocabulary (incl. phantom 0)
    rel = np.abs(X[ids] - emb) / (np.abs(emb) + 1e-12)
    assert rel.max() < 6e-6, rel.max()                          # printing precision of the binary


def test_strided_mode_equals_sequential_without_dead_ends(n2v_oracle):
    z = np.load(golden_path('n2v_bin_symw60.npz'))              # symmetric: no dead ends
    csr, nids = oracle_inputs(nx_from_npz(z))
    a = n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, 15, 3, seed=4242, mode=0)
    b = n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, 15, 3, seed=4242, mode=1)
    assert np.array_equal(a, b)
    # with dead ends (directed Karate) the two modes must differ only AFTER the first early stop
    csr, nids = oracle_inputs(load_karate_nx())
    a = n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, 10, 2, seed=7, mode=0)
    b = n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, 10, 2, seed=7, mode=1)
    assert a.shape == b.shape and not np.array_equal(a, b)
    assert np.array_equal(a[0], b[0])


def test_rng_skip_matches_stepping(n2v_oracle):
    s = 12345
    for k in (0, 1, 2, 1000, 123457):
        x = s
        for _ in range(min(k, 2000)):
            x = (16807 * x) % 2147483647
        if k <= 2000:
            assert n2v_oracle.rng_skip(s, k) == x
    assert n2v_oracle.rng_skip(s, 123457) == pow(16807, 123457, 2147483647) * s % 2147483647


def test_walk_edge_cases(n2v_oracle):
    # chain 5->6->7->8: dead end pads with 0 and creates the phantom token 0
    import networkx as nx
    G = nx.DiGraph([(5, 6), (6, 7), (7, 8)])
    csr, nids = oracle_inputs(G)
    assert nids.tolist() == [5, 6, 7, 8]
    wm = n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, 6, 1, seed=3, mode=1)
    for row in wm:
        nz = row[row != 0]
        assert np.all(np.diff(nz) == 1) and nz[-1] == 8 and np.all(row[len(nz):] == 0)
    X, tok = n2v_oracle.learn(wm, 9, 4, 2, 1, 3)
    assert sorted(tok.tolist()) == [0, 5, 6, 7, 8]
    # walk_len = 1
    wm1 = n2v_oracle.walks(csr.indptr, csr.indices, csr.data, nids, 1, 2, seed=3, mode=1)
    assert wm1.shape == (8, 1) and sorted(wm1[:4, 0].tolist()) == [5, 6, 7, 8]
