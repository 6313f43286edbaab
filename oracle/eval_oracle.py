"""oracle/eval_oracle.py -- TEST INFRASTRUCTURE ONLY (not shipped, never imported by gem_b200/).

CPU restatement of GEM's graph-reconstruction evaluation, the step that follows learn_embedding in every
reference test (tests/fit_model.py:10):

    reconstruct()            static_graph_embedding.py:48-65  (A_hat[i, j] = get_edge_weight(i, j), diagonal 0)
                             with get_edge_weight of hope.py:43-44 (split=True) or node2vec.py:56-57 (split=False)
    edge_list_from_adj()     evaluation_util.py:20-36   (threshold 0: strict '>' on the full scan, '>=' on sampled pairs)
    precision_curve()        metrics.py:6-25            (stable descending sort: ties keep list order)
    compute_map()            metrics.py:28-46
    evaluate()               evaluate_graph_reconstruction.py:8-46

Two forms of each metric: *_loops follow the reference statement by statement (small inputs only); the
vectorised forms are what the tests use at n = 1024 and above and are checked against the loops and against
goldens produced by the reference's own functions (tests/golden/make_golden_eval.py -> eval_*.npz).
Pinned: yes -- tests/test_oracle_eval.py compares with those goldens.
"""
import numpy as np


# ------------------------------------------------------------------------------------------- reconstruction
def reconstruct(X, split, exact=True):
    """exact=True: one np.dot per (i, j) as the reference does (entries that are 0 in exact arithmetic keep the
    reference's rounding, which decides whether they pass the '> 0' filter); exact=False: one GEMM (large n)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    k = X.shape[1] // 2
    L, R = (X[:, :k], X[:, k:]) if split else (X, X)
    if not exact:
        A = L @ R.T
    else:
        A = np.zeros((n, n))
        for i in range(n):
            li = L[i]
            for j in range(n):
                if i != j:
                    A[i, j] = np.dot(li, R[j])
    np.fill_diagonal(A, 0.0)
    return A


# ------------------------------------------------------------------------------------------- edge list
def edge_list_from_adj(adj, threshold=0.0, is_undirected=True, edge_pairs=None):
    """-> (i, j, w) arrays in the order the reference appends them (row-major scan / pair-list order)."""
    adj = np.asarray(adj)
    n = adj.shape[0]
    if edge_pairs is not None and len(edge_pairs):
        ep = np.asarray(edge_pairs, dtype=np.int64).reshape(-1, 2)
        w = adj[ep[:, 0], ep[:, 1]]
        keep = w >= threshold
        return ep[keep, 0], ep[keep, 1], w[keep]
    mask = adj > threshold
    mask[np.arange(n), np.arange(n)] = False
    if is_undirected:
        mask &= np.triu(np.ones((n, n), dtype=bool), 1)
    i, j = np.nonzero(mask)            # row-major order
    return i, j, adj[i, j]


def edge_list_from_adj_loops(adj, threshold=0.0, is_undirected=True, edge_pairs=None):
    result = []
    n = adj.shape[0]
    if edge_pairs:
        for (st, ed) in edge_pairs:
            if adj[st, ed] >= threshold:
                result.append((st, ed, adj[st, ed]))
    else:
        for i in range(n):
            for j in range(n):
                if j == i:
                    continue
                if is_undirected and i >= j:
                    continue
                if adj[i, j] > threshold:
                    result.append((i, j, adj[i, j]))
    return result


# ------------------------------------------------------------------------------------------- graph access
class EdgeSet:
    """has_edge / out_degree of the true graph from a CSR (sorted column ids per row)."""

    def __init__(self, n, indptr, indices):
        self.n = int(n)
        self.indptr = np.asarray(indptr, dtype=np.int64)
        self.indices = np.asarray(indices, dtype=np.int64)
        rows = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(self.indptr))
        self.keys = rows * self.n + self.indices          # sorted (rows ascending, columns ascending)

    @classmethod
    def from_networkx(cls, G):
        n = len(G.nodes)
        e = np.array([(u, v) for u, v in G.edges()], dtype=np.int64).reshape(-1, 2)
        order = np.lexsort((e[:, 1], e[:, 0]))
        e = e[order]
        indptr = np.zeros(n + 1, dtype=np.int64)
        np.add.at(indptr, e[:, 0] + 1, 1)
        return cls(n, np.cumsum(indptr), e[:, 1])

    def has_edge(self, i, j):
        q = np.asarray(i, dtype=np.int64) * self.n + np.asarray(j, dtype=np.int64)
        pos = np.searchsorted(self.keys, q)
        pos = np.minimum(pos, len(self.keys) - 1) if len(self.keys) else pos
        return (self.keys[pos] == q) if len(self.keys) else np.zeros(q.shape, dtype=bool)

    def out_degree(self):
        return np.diff(self.indptr)


# ------------------------------------------------------------------------------------------- metrics
def precision_curve(i, j, w, edges, max_k=-1):
    """metrics.py:6-25 -> (precision_scores, delta_factors) as fp64 arrays."""
    m = len(w)
    max_k = m if max_k == -1 else min(max_k, m)
    order = np.argsort(-np.asarray(w, dtype=np.float64), kind='stable')[:max_k]   # ties keep list order
    delta = edges.has_edge(np.asarray(i)[order], np.asarray(j)[order]).astype(np.float64)
    correct = np.cumsum(delta)
    prec = correct / np.arange(1, max_k + 1, dtype=np.float64)
    return prec, delta


def precision_curve_loops(pred, G, max_k=-1):
    if max_k == -1:
        max_k = len(pred)
    else:
        max_k = min(max_k, len(pred))
    sorted_edges = sorted(pred, key=lambda x: x[2], reverse=True)
    precision_scores, delta_factors, correct_edge = [], [], 0
    for r in range(max_k):
        if G.has_edge(sorted_edges[r][0], sorted_edges[r][1]):
            correct_edge += 1
            delta_factors.append(1.0)
        else:
            delta_factors.append(0.0)
        precision_scores.append(1.0 * correct_edge / (r + 1))
    return precision_scores, delta_factors


def compute_map(i, j, w, edges, max_k=-1, is_undirected=False):
    """metrics.py:28-46.  -> (MAP, node_ap[n], count)"""
    n = edges.n
    i = np.asarray(i, dtype=np.int64); j = np.asarray(j, dtype=np.int64); w = np.asarray(w, dtype=np.float64)
    node_ap = np.zeros(n, dtype=np.float64)
    outdeg = edges.out_degree()
    order = np.argsort(i, kind='stable')                      # node_edges[st] keeps list order
    i, j, w = i[order], j[order], w[order]
    starts = np.searchsorted(i, np.arange(n + 1))
    count = 0
    for v in range(n):
        if not is_undirected and outdeg[v] == 0:
            continue
        count += 1
        s, e = starts[v], starts[v + 1]
        if e == s:
            continue
        prec, delta = precision_curve(i[s:e], j[s:e], w[s:e], edges, max_k)
        sd = 0.0
        sp = 0.0
        for p, dl in zip(prec, delta):                        # python sum(): sequential fp64 adds
            sp += p * dl
            sd += dl
        node_ap[v] = 0.0 if sd == 0 else float(sp / sd)
    total = 0.0
    for v in range(n):
        total += node_ap[v]
    return (total / count if count else float('nan')), node_ap, count


def compute_map_loops(pred, G, max_k=-1, is_undirected=False):
    node_num = len(G.nodes)
    node_edges = [[] for _ in range(node_num)]
    for (st, ed, w) in pred:
        node_edges[st].append((st, ed, w))
    node_ap = [0.0] * node_num
    count = 0
    for v in range(node_num):
        if not is_undirected and G.out_degree(v) == 0:
            continue
        count += 1
        ps, df = precision_curve_loops(node_edges[v], G, max_k)
        pr = [p * d for p, d in zip(ps, df)]
        node_ap[v] = 0 if sum(df) == 0 else float(sum(pr) / sum(df))
    return sum(node_ap) / count


def evaluate(adj, edges, weights=None, is_undirected=True, is_weighted=False, edge_pairs=None, max_k=-1,
             node_order=None):
    """evaluate_graph_reconstruction.py:8-46 on a given reconstruction `adj` (n x n).
    weights: CSR data of the true graph (None = 1.0), only for is_weighted.
    node_order: list(digraph.nodes).  The weighted error compares nx.to_numpy_matrix(digraph) -- rows and columns in
    list(digraph.nodes) order -- with the reconstruction indexed by node ID (:37-40), so edge (u -> v) is compared
    with adj[pos(u), pos(v)]; the restatement keeps that.  -> dict"""
    i, j, w = edge_list_from_adj(adj, is_undirected=is_undirected, edge_pairs=edge_pairs)
    MAP, node_ap, count = compute_map(i, j, w, edges, is_undirected=is_undirected)
    prec, delta = precision_curve(i, j, w, edges, max_k)
    out = {'MAP': MAP, 'node_ap': node_ap, 'count': count, 'prec_curve': prec, 'delta': delta, 'n_pred': len(w),
           'err': None, 'err_baseline': None}
    if is_weighted:
        rows = np.repeat(np.arange(edges.n, dtype=np.int64), np.diff(edges.indptr))
        a = np.ones(len(rows)) if weights is None else np.asarray(weights, dtype=np.float64)
        pos = np.arange(edges.n, dtype=np.int64)
        if node_order is not None:
            pos = np.empty(edges.n, dtype=np.int64)
            pos[np.asarray(node_order, dtype=np.int64)] = np.arange(edges.n)
        est = np.asarray(adj, dtype=np.float64)[pos[rows], pos[edges.indices]]
        nz = a != 0
        out['err'] = float(np.sqrt(np.sum((a[nz] - est[nz]) ** 2)))
        out['err_baseline'] = float(np.sqrt(np.sum(a ** 2)))
    return out
