"""Semidefinite vertex covers (once per instance, host side).

``pattern_E`` builds P^E_rho exactly as the nested loops of ``CutSolver._get_sdp_vertex_cover``
(cut_select_qp.py:401-522, ch_ext = 0) would: every rho-clique of the off-diagonal sparsity graph plus every
smaller clique (size >= 2) that is contained in no clique one size larger, in lexicographic tuple order
(SURVEY.md A.3).  Implemented on adjacency bitmasks (one Python int per vertex) with an explicit DFS; the
candidate list is then shipped to the GPU once (``sdpcs_set_cover_list``).  The all-subsets cover
(cut_select_qp.py:451-455) is never materialised: the kernels unrank on the fly.

``AggList`` is the lazy stand-in for the reference's ``agg_list`` (cut_select_qp.py:524-540): it holds only the
index tuples; ``agg_list[i]`` builds the reference's 4-tuple ``(setInds, Xarr_inds, Q_slice, max_elem)`` on demand.
"""
import itertools

import numpy as np


def _adj_masks(Q_adj):
    A = np.asarray(Q_adj) != 0
    n = A.shape[0]
    A = A | A.T
    masks = []
    for i in range(n):
        m = 0
        for j in np.nonzero(A[i])[0]:
            if j != i:
                m |= 1 << int(j)
        masks.append(m)
    return masks


def pattern_E(Q_adj, dim):
    """Returns int16 array (N, dim), rows padded with -1, in the reference's order."""
    masks = _adj_masks(Q_adj)
    n = len(masks)
    rows = []
    pad = [-1] * dim

    def grow(clique, common):
        """clique: ascending tuple, common: bitmask of vertices adjacent to all of it."""
        s = len(clique)
        if s == dim:
            rows.append(list(clique))
            return
        if common == 0:                      # no clique one size larger contains it -> emit the smaller clique
            rows.append(list(clique) + pad[s:])
            return
        ext = common >> (clique[-1] + 1)     # extend forward only; backward supersets were emitted earlier
        v = clique[-1] + 1
        while ext:
            if ext & 1:
                grow(clique + (v,), common & masks[v])
            ext >>= 1
            v += 1

    for i1 in range(n):
        ext = masks[i1] >> (i1 + 1)
        i2 = i1 + 1
        while ext:
            if ext & 1:
                grow((i1, i2), masks[i1] & masks[i2])
            ext >>= 1
            i2 += 1
    if not rows:
        return np.zeros((0, dim), dtype=np.int16)
    return np.array(rows, dtype=np.int16)


def xarr_inds(n, set_inds):
    """Flat upper-triangular positions of a subset (cut_select_qp.py:530-531)."""
    return [n * a - a * (a + 1) // 2 + b for a, b in itertools.combinations_with_replacement(set_inds, 2)]


class AggList(object):
    """Sequence view over a vertex cover; elements are built lazily in the reference's tuple format."""

    def __init__(self, n, dim, Q_arr, idx=None, n_all=None, offset=0):
        self.n, self.dim, self.Q_arr = n, dim, Q_arr
        self.idx = idx                 # int16 (N, dim) for list covers, None for the all-subsets cover
        self.n_all = n_all             # C(n, dim) for the all-subsets cover
        self.offset = offset
        self._engine = None            # device context owning this cover (set by CutSolver)

    @property
    def is_all(self):
        return self.idx is None

    def __len__(self):
        return int(self.n_all if self.is_all else self.idx.shape[0])

    def set_inds(self, i):
        if self.is_all:
            from . import _capi
            return [int(v) for v in _capi.unrank(self.n, self.dim, [i])[0]]
        return [int(v) for v in self.idx[i] if v >= 0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            if i == slice(None, None, None):
                return self
            if self.is_all:
                raise IndexError("the all-subsets cover is not materialised; index single elements")
            return AggList(self.n, self.dim, self.Q_arr, idx=self.idx[i])
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        s = self.set_inds(i)
        xi = xarr_inds(self.n, s)
        q = self.Q_arr[xi]
        max_elem = len(s) * abs(max(q, key=abs))          # cut_select_qp.py:536-538
        max_elem += 1 if not max_elem else 0
        return (s, xi, tuple(np.divide(q, max_elem)), max_elem)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def keys(self):
        """Hashable identity of every element (the index tuple) -- used for cover intersection / difference."""
        if self.is_all:
            raise ValueError("all-subsets cover")
        return [tuple(int(v) for v in r if v >= 0) for r in self.idx]
