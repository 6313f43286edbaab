"""gem_b200/synth.py -- seeded synthetic graphs of BASELINE.json's configs (host, NumPy; not timed).

sbm(n, ...)   : SURVEY 8(d) config 2/3 -- n nodes in equal blocks of `block` consecutive ids, expected
                degree deg_in inside the block + deg_out outside, symmetrised, unit weights, dedup'd.
rmat(scale,.) : Graph500 R-MAT (a,b,c,d) = (.57,.19,.19,.05), edge factor 8 undirected pairs per node,
                symmetrised, self loops and duplicates removed, unit weights.
Both return a gem_b200.graph.HostCSR with sorted column ids (data = None: unit weights)."""
import numpy as np

from gem_b200.graph import HostCSR


def _csr_from_pairs(n, u, v):
    """undirected pairs -> symmetric, dedup'd, loop-free CSR (unit weights)."""
    keep = u != v
    u, v = u[keep], v[keep]
    key = np.concatenate((u * n + v, v * n + u))
    key.sort()                                   # sort + adjacent-difference: 6x faster than np.unique's hash path
    if key.size:
        keep = np.empty(key.size, dtype=bool)
        keep[0] = True
        np.not_equal(key[1:], key[:-1], out=keep[1:])
        key = key[keep]
    src = key // n
    dst = (key - src * n).astype(np.int32)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=indptr[1:])
    return HostCSR(n, indptr, dst, None, symmetric=True)


def sbm(n=1_000_000, block=1000, deg_in=16.0, deg_out=4.0, seed=42):
    rng = np.random.default_rng(seed)
    n = int(n)
    block = min(block, n)
    nb = n // block
    assert nb * block == n, 'n must be a multiple of the block size'
    m_in = int(round(n * deg_in / 2))
    m_out = int(round(n * deg_out / 2))
    # intra-block pairs: pick a block, then two members
    blk = rng.integers(0, nb, m_in, dtype=np.int64)
    u = blk * block + rng.integers(0, block, m_in, dtype=np.int64)
    v = blk * block + rng.integers(0, block, m_in, dtype=np.int64)
    # inter-block pairs: uniform over all nodes (a 1/nb fraction lands inside a block: negligible)
    uo = rng.integers(0, n, m_out, dtype=np.int64)
    vo = rng.integers(0, n, m_out, dtype=np.int64)
    return _csr_from_pairs(n, np.concatenate((u, uo)), np.concatenate((v, vo)))


def rmat(scale=24, edge_factor=8, a=0.57, b=0.19, c=0.19, seed=42, chunk=1 << 24, permute=True):
    """Graph500 R-MAT: edge_factor * 2^scale undirected pairs, symmetrised, self-loops and
    duplicates removed, unit weights.  permute=True relabels the vertices with a random permutation as the Graph500
    generator does: without it the hubs are the lowest ids, and contiguous equal-row shards (multi-GPU) would put most
    of the edges on rank 0."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    m = n * edge_factor
    us, vs = [], []
    ab, abc = a + b, a + b + c
    for start in range(0, m, chunk):
        cnt = min(chunk, m - start)
        u = np.zeros(cnt, dtype=np.int64)
        v = np.zeros(cnt, dtype=np.int64)
        for _ in range(scale):
            r = rng.random(cnt)
            ubit = r >= ab
            vbit = ((r >= a) & (r < ab)) | (r >= abc)
            u = (u << 1) | ubit
            v = (v << 1) | vbit
        us.append(u)
        vs.append(v)
    u, v = np.concatenate(us), np.concatenate(vs)
    if permute:
        perm = rng.permutation(n)
        u, v = perm[u], perm[v]
    return _csr_from_pairs(n, u, v)
