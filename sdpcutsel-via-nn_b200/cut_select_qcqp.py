"""Drop-in for the selection side of the reference's ``cut_select_qcqp.CutSolverQCQP``.

Keeps the caller pattern of cut_select_qcqp.py:50-98: two vertex covers (objective pattern / constraint-only
pattern, ``__get_vertex_cover`` :314-334), per round a ranking of the objective cover with the chosen strategy,
a feasibility ranking of the constraint-only cover obtained by re-pointing ``self._agg_list``, concatenation and
slicing (not re-sorted, :79), and cut generation that tells the two entry formats apart with
``isinstance(elem[0], int)`` (:91-92).  The OSiL parser and the CPLEX loop stay the reference's.
"""
import numpy as np

from . import _capi, cover
from .cut_select_qp import B200CutSelection, CutSolver, _as_dense_adj as _dense_adj


def _row_keys(idx):
    """(N, dim <= 5) int16 index rows, -1 padded, entries < 511 -> one uint64 key per row."""
    k = np.zeros(idx.shape[0], dtype=np.uint64)
    for t in range(idx.shape[1]):
        k = (k << np.uint64(9)) | (idx[:, t].astype(np.int64) + 1).astype(np.uint64)
    return k


def two_pattern_covers(solver, dim, device_algebra=True):
    """cut_select_qcqp.py:314-334 for any solver object carrying the B200CutSelection surface: sets
    solver._agg_list := P(E_m) intersected with P(E_0) (in P(E_m) order) and returns P(E_m) minus that intersection,
    where P(E_0) is the cover of the objective pattern (_Q_adj) and P(E_m) that of objective + constraints (_Q_adj_cons).
    Both covers are built on the device; intersection and difference are taken there too (sdpcs_cover_filter: binary
    search of every P(E_m) row in P(E_0) + scan compaction) and the index rows come back once for the lazy agg_list views.
    device_algebra=False keeps the set algebra on the host (numpy, on 45-bit row keys) -- the cross-check in the tests."""
    B200CutSelection._get_sdp_vertex_cover(solver, dim)
    agg_obj = solver._agg_list
    Q_adj = solver._Q_adj
    solver._Q_adj = solver._Q_adj_cons
    try:
        B200CutSelection._get_sdp_vertex_cover(solver, dim)
    finally:
        solver._Q_adj = Q_adj
    agg_cons = solver._agg_list
    n, Q_arr = solver._nb_vars, np.asarray(solver._Q_arr, dtype=np.float64)
    empty = cover.AggList(n, dim, Q_arr, idx=np.zeros((0, dim), dtype=np.int16))
    if agg_obj.is_all:                                   # every element of P(E_m) is in P(E_0)
        solver._agg_list = agg_cons
        return empty
    if device_algebra and agg_obj._engine is not None and agg_cons._engine is not None and not agg_cons.is_all:
        eng_obj, eng_int = agg_obj._engine, agg_cons._engine
        eng_diff = solver._new_engine()
        eng_diff.set_cover_pattern(dim, _dense_adj(solver._Q_adj_cons, n))
        eng_int.cover_filter(eng_obj, keep_members=True)
        eng_diff.cover_filter(eng_obj, keep_members=False)
        inter = cover.AggList(n, dim, Q_arr, idx=eng_int.cover_rows())
        diff = cover.AggList(n, dim, Q_arr, idx=eng_diff.cover_rows())
        inter._engine, diff._engine = eng_int, eng_diff
        solver._agg_list = inter
        return diff
    cons_idx = agg_cons.idx if not agg_cons.is_all else \
        _capi.unrank(n, dim, np.arange(len(agg_cons))).astype(np.int16)
    # membership of every P(E_m) row in P(E_0): one 45-bit key per (-1 padded) index row instead of the reference's
    # O(N^2) `el in agg_list` scans (cut_select_qcqp.py:322-331)
    inter = np.isin(_row_keys(cons_idx), _row_keys(agg_obj.idx))
    solver._agg_list = cover.AggList(n, dim, Q_arr, idx=np.ascontiguousarray(cons_idx[inter]))
    return cover.AggList(n, dim, Q_arr, idx=np.ascontiguousarray(cons_idx[~inter]))


class CutSolverQCQP(CutSolver):
    def __init__(self):
        super(CutSolverQCQP, self).__init__()
        self._Q_adj_cons = None

    def set_instance(self, Q_arr, Q_adj, nb_vars, dim=None, my_prob=None, Q_adj_cons=None):
        super(CutSolverQCQP, self).set_instance(Q_arr, Q_adj, nb_vars, dim=dim, my_prob=my_prob)
        self._Q_adj_cons = Q_adj if Q_adj_cons is None else Q_adj_cons

    def __get_vertex_cover(self, dim):                       # -> _CutSolverQCQP__get_vertex_cover
        return two_pattern_covers(self, dim)

    def get_vertex_cover(self, dim):
        return self.__get_vertex_cover(dim)

    def select_and_cut_round(self, strat, vars_values, sel_size, agg_list, agg_list_cons, cut_round=1):
        """One round of cut_select_qcqp.py:63-98 (selection + cut generation, no LP solve).
        Returns (strat_for_next_round, nb_sdp_cuts, nb_opt_cuts, rank_list)."""
        feas_sel, comb_sel = (strat == 1), (strat == 4)
        strat_old = strat
        if comb_sel:
            strat, rank_list_comb_obj = self._sel_eigcut_by_ordering_on_measure(strat, vars_values, cut_round, sel_size=sel_size)
        else:
            rank_list_comb_obj = self._sel_eigcut_by_ordering_on_measure(strat, vars_values, cut_round)
        self._agg_list = agg_list_cons                                   # swap in the constraint-only cover (:75)
        rank_list_feas_cons = self._sel_eigcut_by_ordering_on_measure(1, vars_values, cut_round)
        self._agg_list = agg_list                                        # swap back (:78)
        rank_list = (list(rank_list_comb_obj) + list(rank_list_feas_cons))[0:sel_size]
        nb_opt_cuts = 0
        if feas_sel:
            nb_sdp_cuts = self._gen_eigcuts_selected(strat, sel_size, rank_list, vars_values=vars_values)
        else:
            for elem in rank_list_comb_obj:
                nb_opt_cuts += elem[1] > self._BIG_M
            nb_cuts_combined = 0
            for elem in rank_list:
                nb_cuts_combined += isinstance(elem[0], int)
            nb_cuts_comb = self._gen_eigcuts_selected(1, sel_size - nb_cuts_combined,
                                                      rank_list_feas_cons[0:(sel_size - nb_cuts_combined)], vars_values=vars_values)
            nb_cuts_feas = self._gen_eigcuts_selected(strat_old, nb_cuts_combined, rank_list_comb_obj[0:nb_cuts_combined],
                                                      vars_values=vars_values)
            nb_sdp_cuts = nb_cuts_comb + nb_cuts_feas
        return strat, nb_sdp_cuts, nb_opt_cuts, rank_list
