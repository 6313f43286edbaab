"""gem_b200/graph.py -- host-side graph ingestion: networkx / scipy / edge arrays -> CSR (NumPy only).

Replaces `nx.to_numpy_matrix(graph)` (reference gem/embedding/hope.py:28: rows and columns follow
list(graph.nodes), weight attribute 'weight', missing -> 1) and the text edge list of
graph_util.saveGraphToEdgeListTxtn2v (graph_util.py:137-140) + SNAP's ReadGraph (node table in
first-appearance order, adjacency vectors sorted by id).
"""
import numpy as np


class HostCSR:
    """n x n CSR with int64 indptr, int32 column ids sorted within each row, optional fp64 weights
    (None = all 1.0).  `nodes` = the node label of each row (HOPE: list(graph.nodes) order)."""

    def __init__(self, n, indptr, indices, data=None, nodes=None, symmetric=None):
        self.n = int(n)
        indptr = np.asarray(indptr)
        # int32 offsets are kept as they are (what gemb_graph_upload takes; may live in pinned memory)
        self.indptr = indptr if (indptr.dtype == np.int32 and indptr.flags.c_contiguous) \
            else np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = None if data is None else np.ascontiguousarray(data, dtype=np.float64)
        self.nodes = nodes
        self.symmetric = symmetric      # None = unknown (is_symmetric() computes it)
        assert self.indptr.shape[0] == self.n + 1

    @property
    def nnz(self):
        return int(self.indptr[-1])

    def data_f32(self):
        return None if self.data is None else self.data.astype(np.float32)

    def transpose(self):
        rows = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(self.indptr).astype(np.int64))
        return from_edges(self.n, self.indices, rows, self.data, nodes=self.nodes)

    def is_symmetric(self):
        if self.symmetric is not None:
            return bool(self.symmetric)
        t = self.transpose()
        sym = bool(np.array_equal(t.indptr, self.indptr) and np.array_equal(t.indices, self.indices))
        if sym and self.data is not None:
            sym = bool(np.array_equal(t.data, self.data))
        self.symmetric = sym
        return sym

    def row_shard(self, rank, nranks):
        """Rows [rank*ceil(n/P), ...) with shard-local offsets (the layout gemb_graph_upload wants)."""
        per = (self.n + nranks - 1) // nranks
        r0 = min(self.n, rank * per)
        r1 = min(self.n, r0 + per)
        lo, hi = self.indptr[r0], self.indptr[r1]
        return r0, (self.indptr[r0:r1 + 1] - lo), self.indices[lo:hi], (None if self.data is None else self.data[lo:hi])

    def to_scipy(self):
        import scipy.sparse as sp
        d = np.ones(self.nnz) if self.data is None else self.data
        return sp.csr_matrix((d, self.indices, self.indptr), shape=(self.n, self.n))


def from_edges(n, src, dst, w=None, nodes=None, unit_if_all_ones=True, dup='last'):
    """COO (row ids src, col ids dst in 0..n-1) -> CSR with sorted columns.  Duplicate (src, dst)
    pairs keep the LAST weight (SNAP's AddEdge semantics; a DiGraph has none), or, with dup='sum', add up
    (what nx.to_numpy_matrix does with the parallel edges of a MultiGraph)."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    if src.size and (src.min() < 0 or dst.min() < 0 or src.max() >= n or dst.max() >= n):
        raise ValueError('edge endpoint outside [0, n)')
    ww = None if w is None else np.asarray(w, dtype=np.float64)
    key = src * np.int64(n) + dst if n < (1 << 31) else None
    if key is not None and (key.size < 2 or bool(np.all(key[1:] >= key[:-1]))):
        s, t = src, dst                                        # already in row-major order (files we wrote ourselves)
    else:
        # stable: among equal (src, dst) the input order survives, so "last wins" below is the LAST line of a file
        order = np.argsort(key, kind='stable') if key is not None else np.lexsort((dst, src))
        s, t = src[order], dst[order]
        ww = None if ww is None else ww[order]
    if s.size > 1:
        dupm = (s[1:] == s[:-1]) & (t[1:] == t[:-1])
        if dupm.any():
            keep = np.concatenate((~dupm, [True]))  # last of each run
            if dup == 'sum':
                wsum = np.ones(s.size) if ww is None else ww
                run = np.concatenate(([0], np.cumsum(~dupm)))          # run id of every entry
                ww = np.bincount(run, weights=wsum)
                s, t = s[keep], t[keep]
            else:
                s, t = s[keep], t[keep]
                ww = None if ww is None else ww[keep]
    if ww is not None and unit_if_all_ones and np.all(ww == 1.0):
        ww = None
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(s, minlength=n), out=indptr[1:])
    return HostCSR(n, indptr, t.astype(np.int32), ww, nodes=nodes)


def from_scipy(A):
    import scipy.sparse as sp
    A = sp.csr_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    if A.shape[0] != A.shape[1]:
        raise ValueError('adjacency must be square')
    data = np.asarray(A.data, dtype=np.float64)
    if np.all(data == 1.0):
        data = None
    return HostCSR(A.shape[0], A.indptr, A.indices, data, nodes=None)


def from_networkx(graph, by_label=False):
    """by_label=False (HOPE): row r <-> list(graph.nodes)[r]  (hope.py:28 / SURVEY F6).
    by_label=True (node2vec): row = integer node label, n = max label + 1 (graph_util.py:168).
    An undirected nx.Graph gives both directions, as nx.to_numpy_matrix does.  Integer-labelled graphs take a
    vectorised path (one np.fromiter over graph.edges, 2x the speed of the per-edge loop); anything else -- string
    labels, non-numeric weights -- falls back to the loop."""
    import itertools
    import operator
    nodes = list(graph.nodes)
    m = graph.number_of_edges()
    src = dst = w = None
    undirected_done = False
    lab = np.asarray(nodes) if nodes else np.zeros(0, np.int64)
    adj = getattr(graph, '_adj', None)
    if m and lab.ndim == 1 and lab.dtype.kind in 'iu' and isinstance(adj, dict) and not graph.is_multigraph():
        try:
            # dict-of-dict-of-dict walked at C speed: keys of the inner dicts = neighbours, in graph.edges order
            deg = np.fromiter(map(len, adj.values()), dtype=np.int64, count=len(adj))
            tot = int(deg.sum())
            v = np.fromiter(itertools.chain.from_iterable(adj.values()), dtype=np.int64, count=tot)
            w = np.fromiter(map(operator.methodcaller('get', 'weight', 1),
                                itertools.chain.from_iterable(map(dict.values, adj.values()))),
                            dtype=np.float64, count=tot)
            u = np.repeat(np.fromiter(adj.keys(), dtype=np.int64, count=len(adj)), deg)
            if by_label:
                src, dst = u, v
            elif np.array_equal(lab, np.arange(lab.size)):
                src, dst = u, v                                # labels are already the row numbers
            else:
                order = np.argsort(lab, kind='stable')
                sl = lab[order].astype(np.int64)
                src, dst = order[np.searchsorted(sl, u)], order[np.searchsorted(sl, v)]
            undirected_done = True                             # an nx.Graph stores both directions in _adj
        except (TypeError, ValueError, AttributeError):
            src = None
    if src is None:
        index = None if by_label else {u: i for i, u in enumerate(nodes)}
        s_l, d_l, w_l = [], [], []
        for u, v, ww in graph.edges(data='weight', default=1):
            if by_label:
                s_l.append(int(u)); d_l.append(int(v))
            else:
                s_l.append(index[u]); d_l.append(index[v])
            w_l.append(float(ww))
        src, dst, w = np.array(s_l, dtype=np.int64), np.array(d_l, dtype=np.int64), np.array(w_l, dtype=np.float64)
    if by_label:
        n = int(np.asarray([int(x) for x in nodes], dtype=np.int64).max()) + 1 if nodes else 0
    else:
        n = len(nodes)
    if not graph.is_directed() and not undirected_done:
        off = src != dst
        src, dst, w = np.concatenate((src, dst[off])), np.concatenate((dst, src[off])), np.concatenate((w, w[off]))
    return from_edges(n, src, dst, w, nodes=nodes, dup='sum' if graph.is_multigraph() else 'last')


def n2v_inputs_from_networkx(graph):
    """What the SNAP binary would see after GEM wrote the edge list (node2vec.py:34, graph_util.py:137-140):
    CSR by integer label with weights rounded through '%f', and the node table in first-appearance
    order of the edge list (SNAP ReadGraph)."""
    import itertools
    import operator
    adj = getattr(graph, '_adj', None)
    if graph.is_directed() and isinstance(adj, dict) and not graph.is_multigraph() and graph.number_of_edges():
        try:   # the adjacency dicts walked at C speed, in graph.edges order (see from_networkx)
            deg = np.fromiter(map(len, adj.values()), dtype=np.int64, count=len(adj))
            tot = int(deg.sum())
            dst = np.fromiter(itertools.chain.from_iterable(adj.values()), dtype=np.int64, count=tot)
            w = np.fromiter(map(operator.methodcaller('get', 'weight', 1),
                                itertools.chain.from_iterable(map(dict.values, adj.values()))), dtype=np.float64, count=tot)
            src = np.repeat(np.fromiter(adj.keys(), dtype=np.int64, count=len(adj)), deg)
            return n2v_inputs_from_edges(src, dst, w)
        except (TypeError, ValueError, AttributeError):
            pass
    src, dst, w = [], [], []
    for u, v, ww in graph.edges(data='weight', default=1):
        src.append(int(u)); dst.append(int(v)); w.append(float(ww))
    return n2v_inputs_from_edges(np.array(src, np.int64), np.array(dst, np.int64), np.array(w, np.float64))


def round_weights_like_printf_f(w):
    """float('%f' % w): 6 decimals.  round(w*1e6)/1e6 is the correctly rounded double of the decimal
    (both operands exact) except when w*1e6 itself rounds across a .5 boundary (measure-zero)."""
    w = np.asarray(w, dtype=np.float64)
    return np.rint(w * 1e6) / 1e6


def n2v_inputs_from_edges(src, dst, w=None):
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    n = int(max(src.max(), dst.max())) + 1 if src.size else 0
    # first appearance order over the interleaved (src0, dst0, src1, dst1, ...) sequence
    inter = np.empty(2 * src.size, dtype=np.int64)
    inter[0::2], inter[1::2] = src, dst
    _, first = np.unique(inter, return_index=True)
    nids = inter[np.sort(first)].astype(np.int32)
    ww = None if w is None else round_weights_like_printf_f(w)
    csr = from_edges(n, src, dst, ww)
    return csr, nids
