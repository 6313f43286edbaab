"""Host-side finish of GEM's metrics (reference gem/evaluation/metrics.py:1-46).

The n^2 part -- ranking every candidate edge of every node -- runs on the GPU (gemb_recon_ranks / gemb_recon_top,
gem_b200/csrc/recon.cu); what is left here is O(nnz) bookkeeping in the reference's own summation order.
"""
import numpy as np

precision_pos = [2, 10, 100, 200, 300, 500, 1000]          # metrics.py:3


def map_from_ranks(n, indptr, ranks, is_undirected, out_degree=None):
    """metrics.py:28-46 given, for every true edge (CSR order), its 1-based rank among the predicted edges of its
    source node (0 = not predicted).  AP_i = (sum over hits, in rank order, of (#hits so far) / rank) / #hits;
    nodes without out-edges are skipped unless is_undirected (:38-39); MAP = sum(AP) / count, summed in node order.
    -> (MAP, node_ap, count)"""
    indptr = np.asarray(indptr, dtype=np.int64)
    ranks = np.asarray(ranks, dtype=np.int64)
    node_ap = [0.0] * n
    outdeg = np.diff(indptr) if out_degree is None else np.asarray(out_degree)
    count = 0
    for v in range(n):
        if not is_undirected and outdeg[v] == 0:
            continue
        count += 1
        r = ranks[indptr[v]:indptr[v + 1]]
        r = np.sort(r[r > 0])
        if r.size == 0:
            continue
        acc = 0.0
        for t, rk in enumerate(r.tolist(), 1):               # python sum(): sequential fp64
            acc += (1.0 * t / rk) * 1.0
        node_ap[v] = float(acc / float(r.size))
    total = 0.0
    for a in node_ap:
        total += a
    return (total / count if count else float('nan')), np.array(node_ap), count


def precision_curve_from_top(i, j, w, has_edge, max_k=-1):
    """metrics.py:6-25 given the (unordered) candidates that reach the max_k-th weight: order them as the reference's
    stable descending sort of the row-major list does (weight desc, then i, then j), cut at max_k.
    -> (precision_scores, delta_factors) as python lists"""
    i = np.asarray(i, dtype=np.int64); j = np.asarray(j, dtype=np.int64); w = np.asarray(w)
    order = np.lexsort((j, i, -w.astype(np.float64)))
    if max_k >= 0:
        order = order[:max_k]
    delta = has_edge(i[order], j[order]).astype(np.float64)
    correct = np.cumsum(delta)
    prec = correct / np.arange(1, order.size + 1, dtype=np.float64)
    return prec.tolist(), delta.tolist()
