"""tests/golden/make_golden_lap.py -- fixtures of the UNMODIFIED reference class gem.embedding.lap.LaplacianEigenmaps
(imported from /root/reference; runs only in the build container).  Writes
  karate_LaplacianEigenmaps.txt, sbm1024_LaplacianEigenmaps.npy   the reference's own goldens (tests/karate_res, tests/smb_res)
  ref_lap_<name>_d<d>.npz                                          X, and the graph as a directed edge list with weights
for karate (d = 2, 4), the 1024-node SBM fixture (d = 16) and a weighted directed random graph whose two directions carry
different weights (d = 8): the to_undirected() rule of oracle/lap_oracle.py::undirected_weights is pinned by it."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
import make_golden as mg          # loaders of the reference fixtures

sys.path.insert(0, mg.REF)
import networkx as nx
from gem.embedding.lap import LaplacianEigenmaps


def run(G, d):
    LaplacianEigenmaps.hyper_params = {'method_name': 'lap_eigmap_svd'}
    m = LaplacianEigenmaps(d=d)
    X = np.asarray(m.learn_embedding(graph=G, is_weighted=True, no_python=True))
    return X


def save(name, G, d):
    X = run(G, d)
    nodes = list(G.nodes)
    idx = {u: i for i, u in enumerate(nodes)}
    e = np.array([[idx[u], idx[v], w] for u, v, w in G.edges(data='weight', default=1.0)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'ref_lap_%s_d%d.npz' % (name, d)), X=X, edges=e, n=len(nodes))
    print(name, d, X.shape)


if __name__ == '__main__':
    import shutil
    shutil.copy(os.path.join(mg.REF, 'tests/karate_res/LaplacianEigenmaps.txt'), os.path.join(HERE, 'karate_LaplacianEigenmaps.txt'))
    np.save(os.path.join(HERE, 'sbm1024_LaplacianEigenmaps.npy'),
            np.loadtxt(os.path.join(mg.REF, 'tests/smb_res/LaplacianEigenmaps.txt')).astype(np.float32))
    K = mg.load_karate_nx()
    save('karate', K, 2)
    save('karate', K, 4)
    save('sbm1024', mg.load_sbm_nx(), 16)
    rng = np.random.default_rng(11)
    R = nx.DiGraph()
    R.add_nodes_from(range(120))
    for _ in range(900):
        u, v = int(rng.integers(0, 120)), int(rng.integers(0, 120))
        if u != v:
            R.add_edge(u, v, weight=float(np.round(rng.uniform(0.2, 3.0), 3)))
    save('randw120', R, 8)
