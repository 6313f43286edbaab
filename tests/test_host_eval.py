"""Host-side finish of the evaluation (gem_b200/evaluation/metrics.py) on CPU: given the per-edge ranks and the
unordered top candidates that the GPU kernels return (emulated here with NumPy from the fp64 oracle matrix), the
MAP / precision curve must equal the reference goldens (tests/golden/eval_*.npz)."""
import numpy as np
import pytest

from conftest import eval_golden

CASES = [('eval_karate_hope', ('und', 'dir')), ('eval_karate_n2v', ('und', 'dir')),
         ('eval_randw200_split', ('und', 'dir')), ('eval_randw200_dot', ('und', 'dir')), ('eval_sbm1024_hope', ('und',))]


def _emulated_kernel_outputs(adj, indptr, indices, und):
    """What gemb_recon_ranks / gemb_recon_top (max_k = -1) return, computed the slow way."""
    n = adj.shape[0]
    ranks = np.zeros(len(indices), dtype=np.int32)
    cand_i, cand_j, cand_w = [], [], []
    for i in range(n):
        cols = np.arange(i + 1 if und else 0, n)
        cols = cols[cols != i]
        v = adj[i, cols]
        keep = v > 0
        cols, v = cols[keep], v[keep]
        order = np.argsort(-v, kind='stable')
        pos = np.empty(len(cols), dtype=np.int64)
        pos[order] = np.arange(1, len(cols) + 1)
        where = {int(c): int(p) for c, p in zip(cols, pos)}
        for t in range(indptr[i], indptr[i + 1]):
            ranks[t] = where.get(int(indices[t]), 0)
        cand_i.append(np.full(len(cols), i)); cand_j.append(cols); cand_w.append(v)
    return ranks, np.concatenate(cand_i), np.concatenate(cand_j), np.concatenate(cand_w)


@pytest.mark.parametrize('name,variants', CASES)
def test_map_and_precision_curve_from_kernel_outputs(eval_oracle, name, variants):
    from gem_b200.evaluation import metrics
    eo = eval_oracle
    z, n, (indptr, indices, w) = eval_golden(name)
    adj = eo.reconstruct(z['X'], bool(z['split']))
    edges = eo.EdgeSet(n, indptr, indices)
    rng = np.random.default_rng(0)
    for tag in variants:
        und = tag == 'und'
        ranks, ci, cj, cw = _emulated_kernel_outputs(adj, indptr, indices, und)
        MAP, node_ap, count = metrics.map_from_ranks(n, indptr, ranks, und)
        assert abs(MAP - float(z[tag + '_MAP'])) < 1e-13
        perm = rng.permutation(len(cw))                     # the GPU returns the candidates unordered
        prec, delta = metrics.precision_curve_from_top(ci[perm], cj[perm], cw[perm], edges.has_edge, -1)
        assert len(prec) == int(z[tag + '_n_pred'])
        assert np.array_equal(np.array(prec[:4096]), z[tag + '_prec_head'])
        assert np.array_equal(np.array(prec[::997]), z[tag + '_prec_stride'])
        p100, _ = metrics.precision_curve_from_top(ci[perm], cj[perm], cw[perm], edges.has_edge, 100)
        assert p100 == prec[:100]


def test_random_edge_pairs_contract():
    from gem_b200.utils.evaluation_util import get_random_edge_pairs
    p = get_random_edge_pairs(50, 0.1, True, seed=3)
    assert len(p) == 123        # int(0.1 * 50 * 49) = 245 pairs, / 2 = 122.5 -> the reference's loop stops at 123
    s = set(p)
    assert len(s) == len(p) and not any((b, a) in s for a, b in p if a != b)
    assert get_random_edge_pairs(50, 0.1, True, seed=3) == p
    assert len(get_random_edge_pairs(50, 0.1, False, seed=3)) == int(0.1 * 50 * 49)


@pytest.mark.parametrize('name', ['eval_karate_hope', 'eval_karate_n2v', 'eval_randw200_split', 'eval_randw200_dot'])
def test_reference_named_entry_points(eval_oracle, name):
    """computeMAP / computePrecisionCurve / get_edge_list_from_adj_mtrx under the reference's names
    (gem/evaluation/metrics.py:6-46, gem/utils/evaluation_util.py:20-36) reproduce the reference goldens."""
    import networkx as nx
    from gem_b200.evaluation import metrics
    from gem_b200.utils import evaluation_util
    z, n, _ = eval_golden(name)
    adj = eval_oracle.reconstruct(z['X'], bool(z['split']))
    G = nx.DiGraph()
    G.add_nodes_from(int(x) for x in z['nodes'])
    G.add_weighted_edges_from((int(a), int(b), float(c)) for a, b, c in z['edges'])
    for tag, und in (('und', True), ('dir', False)):
        el = evaluation_util.get_edge_list_from_adj_mtrx(adj, is_undirected=und)
        assert len(el) == int(z[tag + '_n_pred'])
        assert abs(metrics.computeMAP(el, G, is_undirected=und) - float(z[tag + '_MAP'])) < 1e-13
        prec, delta = metrics.computePrecisionCurve(el, G)
        assert np.array_equal(np.array(prec[:4096]), z[tag + '_prec_head'])
        p10, _ = metrics.computePrecisionCurve(el, G, max_k=10)
        assert p10 == prec[:10]
    pairs = [(0, 1), (3, 2), (5, 5)]
    got = evaluation_util.get_edge_list_from_adj_mtrx(adj, threshold=-1e9, edge_pairs=pairs)
    assert [(a, b) for a, b, _ in got] == pairs and got[1][2] == adj[3, 2]


def test_true_csr_of_an_undirected_graph_has_both_directions():
    """an nx.Graph stores (5, 2) once, but has_edge(2, 5) is True in the reference (metrics.py:17)."""
    import networkx as nx
    from gem_b200.evaluation.evaluate_graph_reconstruction import _true_csr
    G = nx.Graph()
    G.add_nodes_from(range(6))
    G.add_edges_from([(5, 2), (0, 1), (3, 3)])
    indptr, indices = _true_csr(G, 6)
    rows = np.repeat(np.arange(6), np.diff(indptr))
    assert sorted(zip(rows.tolist(), indices.tolist())) == [(0, 1), (1, 0), (2, 5), (3, 3), (5, 2)]
