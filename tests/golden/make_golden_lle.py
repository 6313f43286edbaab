"""tests/golden/make_golden_lle.py -- fixtures of the UNMODIFIED reference class gem.embedding.lle.LocallyLinearEmbedding
(imported from /root/reference; runs only in the build container; harness-side shim: nx.to_scipy_sparse_matrix was removed
from networkx 3 -- the class gets nx.to_scipy_sparse_array wrapped into a float csr_matrix, nothing in the class is edited).
Writes karate_LocallyLinearEmbedding.txt, sbm1024_LocallyLinearEmbedding.npy (the reference's own goldens) and
ref_lle_<name>_d<d>.npz (X + the graph) for karate (d = 2, 4), the SBM fixture (d = 16) and a digraph with asymmetric weights (d = 8)."""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True
import make_golden as mg

sys.path.insert(0, mg.REF)
import networkx as nx
import scipy.sparse as sp
nx.to_scipy_sparse_matrix = lambda g, **k: sp.csr_matrix(nx.to_scipy_sparse_array(g, **k), dtype=float)
from gem.embedding.lle import LocallyLinearEmbedding


def save(name, G, d):
    LocallyLinearEmbedding.hyper_params = {'method_name': 'lle_svd'}
    m = LocallyLinearEmbedding(d=d)
    X = np.asarray(m.learn_embedding(graph=G, is_weighted=True, no_python=True))
    nodes = list(G.nodes)
    idx = {u: i for i, u in enumerate(nodes)}
    e = np.array([[idx[u], idx[v], w] for u, v, w in G.edges(data='weight', default=1.0)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'ref_lle_%s_d%d.npz' % (name, d)), X=X, edges=e, n=len(nodes))
    print(name, d, X.shape)


if __name__ == '__main__':
    shutil.copy(os.path.join(mg.REF, 'tests/karate_res/LocallyLinearEmbedding.txt'), os.path.join(HERE, 'karate_LocallyLinearEmbedding.txt'))
    os.chmod(os.path.join(HERE, 'karate_LocallyLinearEmbedding.txt'), 0o644)
    np.save(os.path.join(HERE, 'sbm1024_LocallyLinearEmbedding.npy'),
            np.loadtxt(os.path.join(mg.REF, 'tests/smb_res/LocallyLinearEmbedding.txt')).astype(np.float32))
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
