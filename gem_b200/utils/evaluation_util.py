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
