"""oracle/eval_oracle.py pinned against the reference's own evaluation functions (goldens made by
tests/golden/make_golden_eval.py from gem.evaluation.* of /root/reference), and its vectorised forms against
the statement-by-statement loops."""
import numpy as np
import pytest

from conftest import eval_golden

CASES = [('eval_karate_hope', ('und', 'dir', 'dirw')), ('eval_karate_n2v', ('und', 'dir', 'dirw')),
         ('eval_randw200_split', ('und', 'dir', 'dirw')), ('eval_randw200_dot', ('und', 'dir', 'dirw')),
         ('eval_sbm1024_hope', ('und',))]


@pytest.mark.parametrize('name,variants', CASES)
def test_oracle_matches_reference_evaluation(eval_oracle, name, variants):
    eo = eval_oracle
    z, n, (indptr, indices, w) = eval_golden(name)
    edges = eo.EdgeSet(n, indptr, indices)
    adj = eo.reconstruct(z['X'], bool(z['split']))
    for tag in variants:
        und = tag == 'und'
        r = eo.evaluate(adj, edges, weights=w, is_undirected=und, is_weighted=(tag == 'dirw'), node_order=z['nodes'])
        assert r['n_pred'] == int(z[tag + '_n_pred'])
        assert abs(r['MAP'] - float(z[tag + '_MAP'])) < 1e-13
        assert np.array_equal(r['prec_curve'][:4096], z[tag + '_prec_head'])
        assert np.array_equal(r['prec_curve'][::997], z[tag + '_prec_stride'])
        if tag == 'dirw':
            assert abs(r['err'] - float(z[tag + '_err'])) < 1e-10
            assert abs(r['err_baseline'] - float(z[tag + '_err_baseline'])) < 1e-12


def test_vectorised_forms_equal_the_loops(eval_oracle):
    import networkx as nx
    eo = eval_oracle
    rng = np.random.default_rng(3)
    n = 60
    G = nx.DiGraph()
    G.add_nodes_from(range(n))
    for _ in range(300):
        u, v = rng.integers(0, n, 2)
        G.add_edge(int(u), int(v))
    X = np.round(rng.standard_normal((n, 6)), 1)          # coarse values -> exact ties
    edges = eo.EdgeSet.from_networkx(G)
    for split in (True, False):
        adj = eo.reconstruct(X, split)
        for und in (True, False):
            pl = eo.edge_list_from_adj_loops(adj, is_undirected=und)
            i, j, w = eo.edge_list_from_adj(adj, is_undirected=und)
            assert [(a, b) for a, b, _ in pl] == list(zip(i.tolist(), j.tolist()))
            ps, df = eo.precision_curve_loops(pl, G)
            pv, dv = eo.precision_curve(i, j, w, edges)
            assert np.array_equal(np.array(ps), pv) and np.array_equal(np.array(df), dv)
            assert abs(eo.compute_map_loops(pl, G, is_undirected=und) - eo.compute_map(i, j, w, edges, is_undirected=und)[0]) < 1e-15
            for mk in (1, 7, 50):
                ps, _ = eo.precision_curve_loops(pl, G, mk)
                assert np.array_equal(np.array(ps), eo.precision_curve(i, j, w, edges, mk)[0])
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, n, (200, 2))]
    adj = eo.reconstruct(X, True)
    pl = eo.edge_list_from_adj_loops(adj, edge_pairs=pairs)
    i, j, w = eo.edge_list_from_adj(adj, edge_pairs=pairs)
    assert [(a, b) for a, b, _ in pl] == list(zip(i.tolist(), j.tolist()))
