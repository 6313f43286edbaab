"""tests/golden/make_golden_eval.py -- goldens of the reconstruction evaluation, produced by the UNMODIFIED reference
functions (gem.evaluation.evaluate_graph_reconstruction.evaluateStaticGraphReconstruction with
gem.evaluation.metrics and gem.utils.evaluation_util, imported from /root/reference) on the reference's own
golden embeddings.  Runs only in the build container; writes tests/golden/eval_*.npz.

Cases
  eval_karate_hope        karate (directed, 78 edges), X = tests/karate_res/HOPE.txt, HOPE.get_edge_weight,
                          is_undirected True and False, is_weighted True
  eval_karate_n2v         karate, X = tests/karate_res/node2vec.txt, node2vec.get_edge_weight
  eval_sbm1024_hope       tests/data/sbm.gpickle (1024 nodes), X = tests/smb_res/HOPE.txt (d = 256)
  eval_randw200           200-node weighted directed random graph, random X (d = 16), split and non-split,
                          undirected False, is_weighted True
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (loaders + REF path)

sys.path.insert(0, mg.REF)
import networkx as nx  # noqa: E402

if not hasattr(nx, 'to_numpy_matrix'):   # harness-side shim: networkx 3 dropped it
    nx.to_numpy_matrix = lambda g, **k: np.asmatrix(nx.to_numpy_array(g, **k))

from gem.evaluation import evaluate_graph_reconstruction as gr  # noqa: E402
from gem.embedding.static_graph_embedding import StaticGraphEmbedding  # noqa: E402


class SplitModel(StaticGraphEmbedding):          # get_edge_weight of hope.py:43-44
    def __init__(self, d):
        self._d = d
        self._X = None

    def get_method_name(self): return 'split'
    def get_method_summary(self): return 'split'
    def learn_embedding(self, graph): raise NotImplementedError

    def get_embedding(self): return self._X

    def get_edge_weight(self, i, j):
        return np.dot(self._X[i, :self._d // 2], self._X[j, self._d // 2:])


class DotModel(SplitModel):                      # get_edge_weight of node2vec.py:56-57
    def get_edge_weight(self, i, j):
        return np.dot(self._X[i, :], self._X[j, :])


def run(name, G, X, split, variants):
    model = (SplitModel if split else DotModel)(X.shape[1])
    out = {'X': X, 'split': np.int32(split)}
    e = np.array([(u, v, w) for u, v, w in G.edges(data='weight', default=1)], dtype=np.float64)
    out['edges'] = e
    out['n'] = np.int64(len(G.nodes))
    out['nodes'] = np.array(list(G.nodes), dtype=np.int64)   # nx.to_numpy_matrix row order (is_weighted error)
    for tag, kw in variants.items():
        MAP, prec, err, err_b = gr.evaluateStaticGraphReconstruction(G, model, X, None, **kw)
        prec = np.asarray(prec, dtype=np.float64)
        out[tag + '_MAP'] = np.float64(MAP)
        out[tag + '_n_pred'] = np.int64(len(prec))
        out[tag + '_prec_head'] = prec[:4096]
        out[tag + '_prec_stride'] = prec[::997]
        out[tag + '_err'] = np.float64(-1.0 if err is None else err)
        out[tag + '_err_baseline'] = np.float64(-1.0 if err_b is None else err_b)
        print(name, tag, 'MAP', MAP, 'n_pred', len(prec), 'err', err, err_b, flush=True)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)


def main():
    V = {'und': dict(is_undirected=True), 'dir': dict(is_undirected=False),
         'dirw': dict(is_undirected=False, is_weighted=True)}
    K = mg.load_karate_nx()
    K.add_nodes_from(range(34))
    run('eval_karate_hope', K, np.loadtxt(os.path.join(mg.REF, 'tests/karate_res/HOPE.txt')), True, V)
    run('eval_karate_n2v', K, np.loadtxt(os.path.join(mg.REF, 'tests/karate_res/node2vec.txt')), False, V)
    rng = np.random.default_rng(7)
    n = 200
    R = nx.DiGraph()
    R.add_nodes_from(range(n))
    for _ in range(1500):
        u, v = rng.integers(0, n, 2)
        R.add_edge(int(u), int(v), weight=float(np.round(rng.uniform(0.1, 2.0), 3)))   # self loops allowed
    X = rng.standard_normal((n, 16)) * 0.4
    X[:, 3] = np.round(X[:, 3], 1)        # coarse column -> exact ties are likely after rounding the others
    Xt = np.round(X, 1)                   # many exactly equal scores: exercises the stable tie order
    run('eval_randw200_split', R, Xt, True, V)
    run('eval_randw200_dot', R, Xt, False, V)
    S = mg.load_sbm_nx()
    run('eval_sbm1024_hope', S, np.loadtxt(os.path.join(mg.REF, 'tests/smb_res/HOPE.txt')), True,
        {'und': dict(is_undirected=True)})


if __name__ == '__main__':
    main()
