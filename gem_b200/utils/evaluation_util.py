"""Sampling helper of the evaluation (reference gem/utils/evaluation_util.py:5-18)."""
import secrets

import numpy as np


def get_random_edge_pairs(node_num, sample_ratio=0.01, is_undirected=True, seed=None):
    """Distinct random (st, ed) pairs: int(sample_ratio * n * (n - 1)) of them, half that for undirected graphs, where
    (a, b) and (b, a) count as the same pair; self pairs are possible, as in the reference.  The reference draws from
    `secrets` (not reproducible); `seed` (extra) makes the sample reproducible."""
    num_pairs = int(sample_ratio * node_num * (node_num - 1))
    if is_undirected:
        num_pairs = num_pairs / 2
    rng = np.random.default_rng(secrets.randbits(64) if seed is None else seed)
    chosen = set()
    out = []
    while len(out) < num_pairs:
        a, b = (int(x) for x in rng.integers(0, node_num, 2))
        if (a, b) in chosen or (is_undirected and (b, a) in chosen):
            continue
        chosen.add((a, b))
        out.append((a, b))
    return out


def get_edge_list_from_adj_mtrx(adj, threshold=0.0, is_undirected=True, edge_pairs=None):
    """gem/utils/evaluation_util.py:20-36: [(i, j, adj[i, j]), ...] in row-major order -- entries > threshold off the
    diagonal (i < j only when is_undirected), or the given pairs with adj >= threshold.  Vectorised; same list."""
    adj = np.asarray(adj)
    node_num = adj.shape[0]
    if edge_pairs:
        ep = np.asarray(edge_pairs, dtype=np.int64).reshape(-1, 2)
        w = adj[ep[:, 0], ep[:, 1]]
        keep = w >= threshold
        return [(int(a), int(b), c) for a, b, c in zip(ep[keep, 0], ep[keep, 1], w[keep])]
    mask = adj > threshold
    mask[np.arange(node_num), np.arange(node_num)] = False
    if is_undirected:
        mask &= np.triu(np.ones((node_num, node_num), dtype=bool), 1)
    ii, jj = np.nonzero(mask)                                  # row-major order, like the reference's double loop
    return [(int(a), int(b), c) for a, b, c in zip(ii, jj, adj[ii, jj])]
