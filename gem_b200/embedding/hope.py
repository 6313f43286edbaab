"""HOPE on a B200 -- drop-in for reference gem/embedding/hope.py:8-44.

Same class name, hyper-parameters (d, beta), method name ('hope_gsvd'), call signature
(learn_embedding(graph=None, is_weighted=False, no_python=False)), error behaviour
(ValueError('graph needed')), row order (list(graph.nodes)), column layout
([U sqrt(S) | V sqrt(S)], sigma ascending) and get_edge_weight.  The arithmetic runs in
libgemb200.so (CUDA, sm_100a); there is no CPU path -- without a GPU learn_embedding raises
RuntimeError.

Extra, optional hyper-parameters (defaults keep reference call sites working unchanged):
    tol, max_iters, min_iters, oversample, katz_terms, katz_tol, seed, compute_residual, verbose,
    algorithm (0 auto / 1 general / 2 symmetric-Chebyshev), cheb_degree
        -> gemb_hope_opts (include/gemb200.h)
    device (int), dtype (np.float32 default | np.float64)
`graph` may also be a scipy.sparse matrix or a gem_b200.graph.HostCSR (rows = 0..n-1) so that
million-node inputs need not go through networkx.
"""
import numpy as np

from gem_b200 import _native
from gem_b200 import graph as _graph
from gem_b200.embedding.static_graph_embedding import StaticGraphEmbedding

_OPT_KEYS = ('tol', 'max_iters', 'min_iters', 'oversample', 'katz_terms', 'katz_tol', 'seed',
             'compute_residual', 'verbose', 'algorithm', 'cheb_degree')


class HOPE(StaticGraphEmbedding):

    _recon_split = True      # get_edge_weight form, for the GPU reconstruction (gemb_recon_create)

    hyper_params = {
        'method_name': 'hope_gsvd'
    }

    def __init__(self, *args, **kwargs):
        """ Initialize the HOPE class

        Args:
            d: dimension of the embedding
            beta: higher order coefficient
        """
        super(HOPE, self).__init__(*args, **kwargs)
        self.stats = None
        self._sigma = None

    def _to_csr(self, graph):
        if isinstance(graph, _graph.HostCSR):
            return graph
        if hasattr(graph, 'nodes') and hasattr(graph, 'edges'):
            return _graph.from_networkx(graph)
        return _graph.from_scipy(graph)

    def learn_embedding(self, graph=None, is_weighted=False, no_python=False, out=None, **ignored):
        if graph is None or (hasattr(graph, '__len__') and len(graph) == 0) or \
                (hasattr(graph, 'shape') and graph.shape[0] == 0):
            raise ValueError('graph needed')
        csr = self._to_csr(graph)
        opts = {k: getattr(self, '_' + k) for k in _OPT_KEYS if hasattr(self, '_' + k)}
        ctx = _native.Context(getattr(self, '_device', 0))
        try:
            if csr.is_symmetric():
                g = _native.DeviceGraph(ctx, csr.n, csr.indptr, csr.indices, csr.data_f32())
            else:
                t = csr.transpose()
                g = _native.DeviceGraph(ctx, csr.n, csr.indptr, csr.indices, csr.data_f32(),
                                        t.indptr, t.indices, t.data_f32())
            try:
                X, sigma, st = g.hope(int(self._d), float(self._beta), out=out, **opts)
            finally:
                g.free()
        finally:
            ctx.close()
        self.stats = st
        self._sigma = sigma
        self._node_num = csr.n
        dt = getattr(self, '_dtype', np.float32)
        self._X = X if np.dtype(dt) == np.float32 else X.astype(dt)
        # hope.py:38-40 prints ||U S V^T - S||_F, which needs the dense n x n S; the part that is
        # computable without S is reported instead.
        if getattr(self, '_verbose', 0):
            print('HOPE: algorithm %d, %d iterations, J=%d Katz terms, block %d, ritz change %.3g' %
                  (st['algorithm'], st['iters'], st['katz_terms'], st['block'], st['ritz_change']))
        return self._X

    def get_edge_weight(self, i, j):
        return np.dot(self._X[i, :self._d // 2], self._X[j, self._d // 2:])
