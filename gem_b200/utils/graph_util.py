"""Wire formats on either side of the hot path, byte-compatible with the reference
gem/utils/graph_util.py: saveGraphToEdgeListTxt (:129-134), saveGraphToEdgeListTxtn2v (:137-140),
loadGraphFromEdgeListTxt (:143-158), loadEmbedding (:161-169).  saveEmbedding writes the `.emb`
text that SNAP's WriteOutput (bin@0x406ef0) produces and loadEmbedding reads."""
import numpy as np


def saveGraphToEdgeListTxt(graph, file_name):
    with open(file_name, 'w') as f:
        f.write('%d\n' % len(graph.nodes))
        f.write('%d\n' % len(graph.edges))
        for i, j, w in graph.edges(data='weight', default=1):
            f.write('%d %d %f\n' % (i, j, w))


def saveGraphToEdgeListTxtn2v(graph, file_name):
    with open(file_name, 'w') as f:
        for i, j, w in graph.edges(data='weight', default=1):
            f.write('%d %d %f\n' % (i, j, w))


def loadGraphFromEdgeListTxt(file_name, directed=True):
    import networkx as nx
    with open(file_name, 'r') as f:
        G = nx.DiGraph() if directed else nx.Graph()
        for line in f:
            edge = line.strip().split()
            if not edge:
                continue
            w = float(edge[2]) if len(edge) == 3 else 1.0
            G.add_edge(int(edge[0]), int(edge[1]), weight=w)
    return G


def loadEmbedding(file_name):
    with open(file_name, 'r') as f:
        n, d = f.readline().strip().split()
        X = np.zeros((int(n), int(d)))
        for line in f:
            emb = line.strip().split()
            X[int(emb[0]), :] = [float(e) for e in emb[1:]]
    return X


def saveEmbedding(X, file_name, ids=None):
    """'<V> <d>' then '<id> v1 ... vd' with ~6 significant digits (C++ ostream default)."""
    X = np.asarray(X)
    ids = range(X.shape[0]) if ids is None else ids
    with open(file_name, 'w') as f:
        f.write('%d %d\n' % X.shape)
        for i in ids:
            f.write('%d %s\n' % (i, ' '.join('%g' % v for v in X[i])))
