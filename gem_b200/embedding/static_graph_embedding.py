"""Plugin API of GEM, kept verbatim in behaviour (reference gem/embedding/static_graph_embedding.py:5-83):
same constructor plumbing (class-level hyper_params dict updated by kwargs, then mirrored into
self._<key>: SURVEY F13), same getters, same error strings.  get_reconstructed_adj is the one method
whose n^2 Python loop (reference :59-64) is replaced: a subclass that declares `_recon_split` (True: split
halves, hope.py:43-44; False: dot product, node2vec.py:56-57) gets the product from the GPU
(gemb_recon_create / gemb_recon_dense, fp32 arithmetic, returned as fp64 with a zero diagonal); there is no
CPU path for it -- without a GPU it raises RuntimeError."""
from abc import ABC, abstractmethod

import numpy as np


class StaticGraphEmbedding(ABC):

    def __init__(self, *args, **kwargs):
        """Initialize the Embedding class"""
        self._method_name = None
        self._d = None
        self._X = None
        self.hyper_params.update(kwargs)
        for key in self.hyper_params.keys():
            self.__setattr__('_%s' % key, self.hyper_params[key])
        for dictionary in args:
            for key in dictionary:
                self.__setattr__('_%s' % key, dictionary[key])

    def get_method_name(self):
        return self._method_name

    def get_method_summary(self):
        return '%s_%d' % (self._method_name, self._d)

    def get_embedding(self):
        if self._X is None:
            raise ValueError("Embedding not learned yet")
        return self._X

    def get_reconstructed_adj(self, X=None, node_l=None):
        """Reference :48-65: sets self._X when X is given; A_hat[i, j] = get_edge_weight(i, j), i != j."""
        if X is not None:
            node_num = X.shape[0]
            self._X = X
        else:
            node_num = self._node_num
        split = getattr(self, '_recon_split', None)
        if split is not None:
            from gem_b200 import _native
            ctx = _native.Context(getattr(self, '_device', 0))
            try:
                rec = _native.Reconstruction(ctx, np.asarray(self._X)[:node_num], split)
                try:
                    return rec.dense().astype(np.float64)
                finally:
                    rec.free()
            finally:
                ctx.close()
        adj_mtx_r = np.zeros((node_num, node_num))
        for v_i in range(node_num):
            for v_j in range(node_num):
                if v_i == v_j:
                    continue
                adj_mtx_r[v_i, v_j] = self.get_edge_weight(v_i, v_j)
        return adj_mtx_r

    @abstractmethod
    def learn_embedding(self, graph):
        """Learn the graph embedding from a networkx DiGraph."""

    @abstractmethod
    def get_edge_weight(self, i, j):
        """Weight of the edge between rows i and j of the embedding."""
