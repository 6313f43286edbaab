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
    outdeg = np.diff(indptr) if out_degree is None else np.asarray(out_degree)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    hit = ranks > 0
    hr, rk = rows[hit], ranks[hit]
    order = np.lexsort((rk, hr))                              # by node, then by rank
    hr, rk = hr[order], rk[order]
    hits = np.bincount(hr, minlength=n)                       # sum(delta_factors) per node
    first = np.cumsum(hits) - hits
    t = np.arange(1, hr.size + 1, dtype=np.int64) - first[hr]  # number of hits up to and including this one
    # np.bincount adds its weights one by one in input order = rank order: the reference's sequential sum()
    sums = np.bincount(hr, weights=t.astype(np.float64) / rk.astype(np.float64), minlength=n)
    node_ap = np.where(hits > 0, sums / np.maximum(hits, 1), 0.0)
    counted = np.ones(n, dtype=bool) if is_undirected else (outdeg > 0)
    count = int(counted.sum())
    total = 0.0
    for a in node_ap[counted].tolist():                       # sum(node_ap): sequential, in node order
        total += a
    return (total / count if count else float('nan')), node_ap, count


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
