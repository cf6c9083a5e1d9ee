"""Drop-in for the selection hot path of the reference's ``cut_select_qp.CutSolver``.

Same method names, argument meaning, return formats and error behaviour as the reference for
    _load_neural_nets              (cut_select_qp.py:284-303)
    _get_sdp_vertex_cover          (cut_select_qp.py:377-541)
    _sel_eigcut_by_ordering_on_measure (cut_select_qp.py:543-703, strat 1, 2, 3, 4, 5, -1)
    _gen_eigcuts_selected          (cut_select_qp.py:705-755)
    _get_eigendecomp               (cut_select_qp.py:788-797)
    __preprocess_triangle_ineq / __separate_and_add_triangle (cut_select_qp.py:799-863)
    __gen_dense_eigcuts            (cut_select_qp.py:757-786, strat 0)
but every score, selection and cut is computed by libsdpcutsel on the GPU.  The CPLEX LP loop
(cut_select_algo), the instance readers and McCormick rows stay the reference's: mix this class in front of it
 -- ``class CutSolver(B200CutSelection, reference.CutSolver)`` -- or use it standalone through
``set_instance`` when no LP solver is around (tests, benchmarks).

Narrowing: the ranked list a selection returns is only ever consumed as a prefix of length
<= sel_size <= _SDP_CUTS_PER_ROUND_MAX, so strat 1 / 2 return the first ``_RANK_PREFIX`` (= 5000, or sel_size if
larger) entries of the reference's list instead of all N; ``RankList.n_total`` / ``n_violated`` keep the counts.
Near ties: the device returns the winners plus every candidate within a guard band of the k-th score;
runs of entries closer than the guard are re-scored with the reference's own arithmetic (numpy eigvalsh / the generated
NN function, ``neartie.py``) and re-sorted, so that the prefix equals the reference's also on LP vertices where
thousands of candidates tie; ``RankList.degenerate`` / ``n_near_ties`` report what happened.
strat 3 (exact optimality: the rho-dimensional SDP the reference hands to Mosek) and strat -1 (figure 8: estimate vs
exact) run on the batched SDP solver of sdp_kernels.cuh.  strat 5 (random) shuffles the cover with numpy's legacy generator
exactly like the reference (same permutation for the same seed) and cuts its first sel_size entries on the device.
Out of scope and raising NotImplementedError: ch_ext 1 / 2 (chompack chordal extension).
"""
import operator

import numpy as np

from . import _capi, cover, neartie, nn_weights
from .distributed import ShardedSelector

try:  # the cut sink type of the reference (cut_select_qp.py:747); a plain container when CPLEX is absent
    from cplex import SparsePair
except ImportError:  # pragma: no cover - depends on the environment
    class SparsePair(object):
        def __init__(self, ind=None, val=None):
            self.ind, self.val = list(ind), list(val)


class RankList(list):
    """Prefix of the reference's sorted rank_list plus the totals the full list would have had.
    degenerate: 0 = no near tie touched the prefix, 1 = near ties re-scored with the reference's arithmetic and resolved,
    2 = more near ties than the guard band holds (order inside them is the device's own); n_near_ties = entries re-scored."""
    n_total = 0
    n_violated = 0
    degenerate = 0
    n_near_ties = 0
    _rows = None        # (len, dim) int16 index tuples (-1 padded) and scores the entries were built from: lets
    _scores = None      # _gen_eigcuts_selected skip the walk over the entries while the list is still as returned
    _ends = ()

    def _seal(self, rows, scores):
        self._rows, self._scores, self._ends = rows, scores, tuple(self)

    def _sealed_rows(self):
        """The arrays behind the entries if the list still holds exactly the entries it was returned with, in that order."""
        if self._rows is None or len(self) != len(self._ends) or not all(map(operator.is_, self, self._ends)):
            return None
        return self._rows, self._scores


class _RowSink(object):
    """Stand-in for ``my_prob.linear_constraints`` when no CPLEX model is attached (collects the rows)."""

    def __init__(self):
        self.rows = []

    def add(self, lin_expr=None, rhs=None, senses=None, **kw):
        self.rows.extend(zip(lin_expr, rhs, senses))


class _Prob(object):
    def __init__(self):
        self.linear_constraints = _RowSink()


class B200CutSelection(object):
    # class constants of the reference (cut_select_qp.py:22-41)
    _THRES_MIN_OPT = 0
    _THRES_NEG_EIGVAL = -10 ** (-15)
    _BIG_M = 1000
    _CONVERGENCE_TOL = 10 ** (-3)
    _THRES_TRI_DENSE = 2
    _THRES_TRI_VIOL = 10 ** (-7)
    # the reference stops at 4e6 sub-problems because agg_list would not fit in RAM (cut_select_qp.py:35, 117,
    # 527); nothing is materialised here, so the wall is the index width of the kernels (C(250,5) < 2^44)
    _THRES_MAX_SUBS = 2 ** 44
    _SDP_CUTS_PER_ROUND_MAX = 5000
    _TRI_CUTS_PER_ROUND_MIN = 5000
    _TRI_CUTS_PER_ROUND_MAX = 10000
    _RANK_PREFIX = 5000          # length of the ranked prefix returned when the caller gives no sel_size
    _DEVICE = 0

    def __init__(self, *a, **kw):
        super(B200CutSelection, self).__init__(*a, **kw)
        for name, val in (("_dim", 0), ("_nb_vars", 0), ("_nb_lifted", 0), ("_Q", []), ("_Q_adj", []), ("_Q_arr", []),
                          ("_my_prob", None), ("_agg_list", []), ("_nns", None), ("_rank_list_tri", []), ("_idx_list_tri", [])):
            if not hasattr(self, name):
                setattr(self, name, val)
        self._blobs = {}
        self._last_vars_values = None
        self._tri_engine = None

    # -- standalone setup (what __parse_boxqp_into_cplex leaves behind, cut_select_qp.py:313-328) ------
    def set_instance(self, Q_arr, Q_adj, nb_vars, dim=None, my_prob=None):
        self._Q_arr = np.ascontiguousarray(Q_arr, dtype=np.float64)
        self._Q_adj = Q_adj
        self._nb_vars, self._nb_lifted = int(nb_vars), int(nb_vars) * (int(nb_vars) + 1) // 2
        if dim is not None:
            self._dim = dim
        self._my_prob = my_prob if my_prob is not None else _Prob()
        self._agg_list, self._tri_engine, self._dense_engine = [], None, None

    # -- engines ---------------------------------------------------------------------------------------
    _GUARD_LAM = 10 ** (-12)     # near-tie guard on eigenvalue scores (device vs LAPACK differ by <= ~1e-14)
    _GAP_TOL = 10 ** (-9)        # eigenvalue gap below which an eigenvector cut is taken from numpy eigh (not unique)
    # True: every selected cut row (<= 5000 per round) is taken from numpy eigh, i.e. bit-identical to the reference's row.
    # The device rows agree with those to ~1e-15, but an LP with alternative optima may answer even such a difference
    # with another vertex; set this when the LP trajectory has to follow the reference's run vertex for vertex.
    _CUT_ROWS_FROM_LAPACK = False
    _GUARD_OBJ_REL = 4 * 10 ** (-12)   # near-tie guard on optimality measures, times rho * max|Q_arr| (the scale of max_elem)

    def _guards(self):
        Q_arr = np.asarray(self._Q_arr, dtype=np.float64)
        scale = float(np.abs(Q_arr).max()) if Q_arr.size else 1.0
        return float(self._GUARD_LAM), max(10 ** (-12), self._GUARD_OBJ_REL * max(self._dim, 2) * max(scale, 1.0))

    def _new_engine(self):
        eng = _capi.Engine(self._DEVICE)
        g_lam, g_obj = self._guards()
        eng.set_params(thres_min_opt=float(self._THRES_MIN_OPT), thres_neg_eigval=float(self._THRES_NEG_EIGVAL),
                       big_m=float(self._BIG_M), thres_tri_viol=float(self._THRES_TRI_VIOL),
                       thres_tri_dense=int(self._THRES_TRI_DENSE), guard_lam=g_lam, guard_obj=g_obj)
        for d, blob in self._blobs.items():
            eng.set_weights(d, blob)
        eng.set_instance(self._nb_vars, np.asarray(self._Q_arr, dtype=np.float64))
        return eng

    def _engine_for(self, agg_list):
        """The device context that holds `agg_list` as its cover (the caller may re-point self._agg_list)."""
        if not isinstance(agg_list, cover.AggList):
            # a plain list in the reference's own tuple format (e.g. built by the reference's
            # CutSolverQCQP.__get_vertex_cover, cut_select_qcqp.py:319-333): adopt its index lists
            agg_list = self._adopt_agg_list(agg_list)
        if agg_list._engine is None:
            eng = self._new_engine()
            if agg_list.is_all:
                eng.set_cover_all(agg_list.dim)
            else:
                eng.set_cover_list(agg_list.dim, agg_list.idx, agg_list.offset)
            agg_list._engine = eng
        return agg_list._engine

    def _adopt_agg_list(self, plain):
        """AggList view of a plain list of reference tuples (setInds, Xarr_inds, Q_slice, max_elem); cached per list object."""
        cache = self.__dict__.setdefault("_adopted", {})
        hit = cache.get(id(plain))
        if hit is not None and hit[0] is plain and len(plain) == len(hit[1]):
            return hit[1]
        dim = max([self._dim] + [len(el[0]) for el in plain]) if len(plain) else max(self._dim, 2)
        idx = np.full((len(plain), dim), -1, dtype=np.int16)
        for i, el in enumerate(plain):
            idx[i, :len(el[0])] = el[0]
        agg = cover.AggList(self._nb_vars, dim, np.asarray(self._Q_arr, dtype=np.float64), idx=idx)
        if len(cache) > 8:
            cache.clear()
        cache[id(plain)] = (plain, agg)
        return agg

    # -- reference surface -----------------------------------------------------------------------------
    def _load_neural_nets(self):
        """NN_2D.._dimD weights from the packed .m constants (replaces the ctypes load of neural_nets/NNs.so).
        self._nns keeps the reference's shape: a list of (callable, input_array) per dimension 2.._dim."""
        self._nns = []
        for d in range(2, self._dim + 1):
            if d not in self._blobs:
                self._blobs[d] = nn_weights.load_packed(d)
            input_arr = np.zeros(d * (d + 3) // 2)
            self._nns.append((_NNFunc(self, d), input_arr))

    def _get_sdp_vertex_cover(self, dim, ch_ext=0):
        if ch_ext in (1, 2):
            raise NotImplementedError("chordal-extension covers (ch_ext 1, 2) need chompack and are out of scope")
        n = self._nb_vars
        self._agg_list = None
        Q_arr = np.asarray(self._Q_arr, dtype=np.float64)
        if dim == 3 and ch_ext not in (0, -1):
            # the reference's dim-3 branch knows ch_ext 0, 1, 2, -1 only (cut_select_qp.py:405-455): anything else leaves
            # idx_list empty
            agg = cover.AggList(n, dim, Q_arr, idx=np.zeros((0, dim), dtype=np.int16))
        elif dim == 3 and ch_ext == -1 and n >= 3:
            # P^E+_3: all triples (cut_select_qp.py:451-455); for dim 4 / 5 the reference ignores ch_ext and builds P^E
            agg = cover.AggList(n, dim, Q_arr, n_all=_capi.binom(n, dim))
        else:
            adj = _as_dense_adj(self._Q_adj, n)
            if n >= dim and _is_complete(adj):   # dense pattern: P^E_dim is all subsets in lex order -> nothing to store
                agg = cover.AggList(n, dim, Q_arr, n_all=_capi.binom(n, dim))
            else:
                # the nested clique loops (cut_select_qp.py:401-522) run on the device; the index tuples come back once
                # for the lazy agg_list view, the device context stays attached as the cover's engine
                eng = self._new_engine()
                eng.set_cover_pattern(dim, adj)
                agg = cover.AggList(n, dim, Q_arr, idx=eng.cover_rows())
                agg._engine = eng
        if len(agg) >= self._THRES_MAX_SUBS:
            return len(agg)
        self._agg_list = agg
        return len(agg)

    def _sel_eigcut_by_ordering_on_measure(self, strat, vars_values, cut_round, sel_size=0):
        if strat == 5:
            return self._random_order()
        if strat not in (1, 2, 3, 4, -1):
            raise ValueError("strat must be 1 (feasibility), 2 (optimality), 3 (exact SDP), 4 (combined), 5 (random) or -1 (figure 8)")
        if strat == -1:
            return self._figure_8(vars_values, cut_round, sel_size)
        agg = self._agg_view()
        eng = self._engine_for(agg)
        N = len(agg)
        sel_size = min(sel_size, N)                                   # cut_select_qp.py:550
        vars_values = np.ascontiguousarray(vars_values, dtype=np.float64)
        self._last_vars_values = vars_values
        n, nb_lifted = self._nb_vars, self._nb_lifted
        if strat == 4 and sel_size == 0:
            # the reference divides by sel_size, swallows the ZeroDivisionError and falls through (cut_select_qp.py:628-632)
            strat_eff, k = 2, min(N, self._RANK_PREFIX)
        else:
            strat_eff = strat
            k = sel_size if strat == 4 else min(N, max(sel_size, self._RANK_PREFIX))
        res = self._select_resolved(eng, agg, strat_eff, vars_values, k)
        out = RankList()
        out.n_total = N
        out.n_violated = int(res["counts"][1])
        out.degenerate, out.n_near_ties = int(res["degenerate"]), int(res["n_near_ties"])
        X_vals, x_vals = vars_values[:nb_lifted], vars_values[nb_lifted:]
        # the reference's tuples, built column-wise (numpy gathers + tolist) instead of entry by entry
        rows = self._set_rows(agg, res["idx"])
        sets, xinds, sizes, pts, Xs = self._entry_columns(agg, res["idx"], x_vals, X_vals, with_values=(strat_eff != 1), rows=rows)
        if strat_eff == 1:
            out.extend(zip(sets, res["score"].tolist(), xinds, sizes))                # cut_select_qp.py:649
            out._seal(rows.astype(np.int16), res["score"])
            return out
        out.extend(zip(res["idx"].tolist(), res["score"].tolist(), pts, Xs))          # cut_select_qp.py:599
        out._seal(rows.astype(np.int16), res["score"])
        if strat_eff == 4:
            return (int(res["new_strat"]), out)                                    # cut_select_qp.py:629-630
        return out

    _SHUFFLE_MAX_SUBS = 5 * 10 ** 7

    def _random_order(self):
        """strat 5 (cut_select_qp.py:634-637): ``np.random.shuffle(agg_list)`` in place, the shuffled cover is the ranking.
        numpy's legacy shuffle draws the same sequence for a list and for an index array of the same length, so the
        permutation is the reference's for the same ``np.random.seed``; like there the cover STAYS shuffled (agg_idx of later
        rounds are positions in the shuffled list).  The lazy cover becomes a list cover in the new order; its device
        context is dropped and rebuilt on the next scoring call."""
        agg = self._agg_list
        if not isinstance(agg, cover.AggList):
            np.random.shuffle(agg)                                    # a plain list of reference tuples: the reference's statement
            self.__dict__.get("_adopted", {}).pop(id(agg), None)
            return agg
        N = len(agg)
        if N > self._SHUFFLE_MAX_SUBS:
            raise ValueError("random selection materialises the shuffled cover; %d sub-problems are too many" % N)
        perm = np.arange(N, dtype=np.int64)
        np.random.shuffle(perm)
        agg.idx = self._set_rows(agg, perm + agg.offset).astype(np.int16)
        agg.n_all, agg._engine = None, None
        return agg

    _FIG8_MAX_SUBS = 2 * 10 ** 6

    def _figure_8(self, vars_values, cut_round, sel_size):
        """strat -1 (cut_select_qp.py:660-702): every sub-problem scored by the NN estimate AND by the exact SDP, both
        complete rankings, overlap of the two selections.  Returns (rank_list, share selected by both, std of the exact
        measures of the exact selection, this_round_cuts) like the reference."""
        agg = self._agg_view()
        eng = self._engine_for(agg)
        N = len(agg)
        if N > self._FIG8_MAX_SUBS:
            raise ValueError("figure-8 mode returns complete rankings; %d sub-problems are too many" % N)
        sel_size = min(sel_size, N)                                   # cut_select_qp.py:550
        vars_values = np.ascontiguousarray(vars_values, dtype=np.float64)
        self._last_vars_values = vars_values
        eng.score(vars_values, 2)
        _, est = eng.scores(lam=False)
        eng.score(None, 4)
        _, exact = eng.scores(lam=False)
        o_est = np.argsort(-est, kind="stable")                       # list.sort(key=itemgetter(1), reverse=True)
        o_ex = np.argsort(-exact, kind="stable")
        place_ex = np.empty(N, dtype=np.int64)
        place_ex[o_ex] = np.arange(N)
        nb_lifted = self._nb_lifted
        sets, xinds, sizes, pts, Xs = self._entry_columns(agg, o_est + agg.offset, vars_values[nb_lifted:], vars_values[:nb_lifted], True)
        rank_list = RankList(zip((o_est + agg.offset).tolist(), est[o_est].tolist(), pts, Xs))
        rank_list.n_total = N
        by_est = (np.arange(N) < sel_size).astype(int)
        by_ex = (place_ex[o_est] < sel_size).astype(int)
        this_round_cuts = [[cut_round, c, a, b, e, x] for c, a, b, e, x in
                           zip((o_est + agg.offset).tolist(), by_est.tolist(), by_ex.tolist(), est[o_est].tolist(), exact[o_est].tolist())]
        std_dev_exact = np.std(exact[o_ex[:sel_size]])
        return rank_list, int((by_est & by_ex).sum()) / sel_size, std_dev_exact, this_round_cuts

    def _entry_columns(self, agg, idx, x_vals, X_vals, with_values, rows=None):
        """Per selected candidate: set_inds list, Xarr_inds list and size (cut_select_qp.py:530-531; the feasibility format,
        with_values False) or the curr_pt / X_slice tuples (:573-574; the optimality formats, with_values True).  Grouped by
        subset size so that everything is array work; the columns the format does not use come back as lists of None."""
        if rows is None:
            rows = self._set_rows(agg, idx)
        m, n = rows.shape[0], self._nb_vars
        sets, xinds, pts, Xs = [None] * m, [None] * m, [None] * m, [None] * m
        sizes = (rows >= 0).sum(axis=1)
        for d in (np.unique(sizes) if m else ()):
            whole = bool((sizes == d).all())
            where = None if whole else np.nonzero(sizes == d)[0]
            sub = rows[:, :d] if whole else rows[where, :d]
            pt = neartie._pairs(int(d))
            a, b = sub[:, pt[:, 0]], sub[:, pt[:, 1]]
            xi = n * a - a * (a + 1) // 2 + b
            if with_values:      # tuples straight out of zip over the columns
                cols = [None, None, list(zip(*x_vals[sub].T.tolist())), list(zip(*X_vals[xi].T.tolist()))]
            else:
                cols = [sub.tolist(), xi.tolist(), None, None]
            if whole:
                sets, xinds, pts, Xs = [c if c is not None else [None] * m for c in cols]
            else:
                for dst, col in zip((sets, xinds, pts, Xs), cols):
                    if col is not None:
                        for j, p in enumerate(where.tolist()):
                            dst[p] = col[j]
        return sets, xinds, sizes.tolist(), pts, Xs

    def _select_resolved(self, eng, agg, strat, vars_values, k):
        """Device selection (winners + guard band) followed by the near-tie resolution of neartie.resolve.  If re-scoring
        drops entries of the strong list of the combined rule (classified differently by the reference's arithmetic), the
        selection is repeated deeper."""
        g_lam, g_obj = float(eng.params.guard_lam), float(eng.params.guard_obj)
        sel = ShardedSelector(eng, local=True)
        rescorer = neartie.Rescorer(self._nb_vars, self._Q_arr, vars_values, self._blobs,
                                    lambda idx: self._set_rows(agg, idx), thr_eig=float(self._THRES_NEG_EIGVAL),
                                    thr_opt=float(self._THRES_MIN_OPT), big_m=float(self._BIG_M))
        if strat == 3:
            # exact SDP measure: the reference's values come out of Mosek (1e-8-level solver noise), so there is no
            # reference arithmetic to re-score near ties with; they are counted and reported only
            raw = sel.select(3, vars_values, k)
            band_n = int(raw["guard"]["n_band"])
            near = int(neartie.tie_runs(raw["score"], g_obj).sum()) + band_n
            res = dict(raw, n_near_ties=near, degenerate=2 if near else 0)
            res.pop("band", None)
            return res
        k_try, vv = k, vars_values
        for _ in range(4):
            raw = sel.select(strat, vv, k_try)
            res = neartie.resolve(raw, k, rescorer, g_lam, g_obj)
            short = res.get("short", 0)
            if not short or raw["idx"].size < k_try:
                break
            k_try, vv = min(len(agg), k_try + short + 64), None       # the LP point is resident now
        if res.get("short", 0) and raw["idx"].size >= k_try:
            res["degenerate"] = 2
        return res

    def _set_rows(self, agg, idx):
        """(m, dim) int array of index tuples, -1 padded, for candidate indices idx."""
        idx = np.asarray(idx, dtype=np.int64)
        if agg.is_all:
            if idx.size == 0:
                return np.zeros((0, agg.dim), dtype=np.int64)
            return _capi.unrank(agg.n, agg.dim, idx - agg.offset).astype(np.int64)
        return agg.idx[idx - agg.offset].astype(np.int64)

    def _sets_of(self, agg, idx):
        if len(idx) == 0:
            return []
        rows = self._set_rows(agg, idx).tolist()
        return rows if agg.is_all else [[v for v in r if v >= 0] for r in rows]

    def _agg_view(self):
        """self._agg_list as an AggList (the caller may have re-pointed it to a plain list in the reference's format)."""
        agg = self._agg_list
        return agg if isinstance(agg, cover.AggList) else self._adopt_agg_list(agg)

    def _any_engine(self):
        """A device context of the current instance for calls that do not need a cover (cuts, single eigen-decompositions)."""
        if isinstance(self._agg_list, cover.AggList) or (isinstance(self._agg_list, list) and len(self._agg_list)):
            return self._engine_for(self._agg_list)
        eng = getattr(self, "_misc_engine", None)
        if eng is None or eng.n != self._nb_vars or getattr(self, "_misc_engine_Q", None) is not self._Q_arr:
            self._misc_engine, self._misc_engine_Q = self._new_engine(), self._Q_arr
        return self._misc_engine

    def _gen_eigcuts_selected(self, strat, sel_size, rank_list, strong_only=False, vars_values=None):
        opt_sel, feas_sel, rand_sel = (strat in [2, 3, 4, -1]), (strat == 1), (strat == 5)
        my_prob = self._my_prob
        sel_size = min(sel_size, len(rank_list))                                   # cut_select_qp.py:713
        if vars_values is None:
            vars_values = self._last_vars_values                                   # opt entries carry their own point
        packed = None
        sealed = rank_list._sealed_rows() if type(rank_list) is RankList else None
        if isinstance(rank_list, cover.AggList):
            # random selection hands the (shuffled) cover itself over (cut_select_qp.py:636-637, 728-731): its first rows
            if sel_size:
                packed = self._set_rows(rank_list, np.arange(sel_size, dtype=np.int64) + rank_list.offset).astype(np.int16)
        elif sealed is not None:
            # the list is the one _sel_eigcut_by_ordering_on_measure returned: its index tuples are at hand as an array
            rows, scores = sealed
            m = sel_size
            if opt_sel and strong_only:                                           # cut_select_qp.py:725-726
                weak = np.nonzero(scores[:sel_size] <= 0)[0]
                m = int(weak[0]) if weak.size else sel_size
            if m:
                packed = rows[:m]
        elif opt_sel:
            idxs = []
            for ix in range(sel_size):
                entry = rank_list[ix]
                if strong_only and entry[1] <= 0:                                  # cut_select_qp.py:725-726
                    break
                idxs.append(entry[0])
            if idxs:
                packed = self._set_rows(self._agg_view(), idxs).astype(np.int16)
        elif sel_size:
            sets = [rank_list[ix][0] for ix in range(sel_size)]
            dim = max(self._dim, max(len(s) for s in sets))
            packed = np.full((len(sets), dim), -1, dtype=np.int16)
            for i, s in enumerate(sets):
                packed[i, :len(s)] = s
        coeffs_sdp, rhs_sdp, senses_sdp = [], [], []
        if packed is not None and packed.shape[0]:
            eng = self._any_engine()
            csr = eng.gen_cuts_csr(packed.shape[1], packed, vars_values)   # violated cuts only (eigvals[0] < _THRES_NEG_EIGVAL)
            # rows whose eigenvector is not unique (repeated smallest eigenvalue: every LP vertex has them) or whose lam_min
            # is within the guard of the threshold: the reference's row is what numpy eigh returns -- recompute those
            thr = float(self._THRES_NEG_EIGVAL)
            fix = (csr["gap"] <= self._GAP_TOL) | (np.abs(csr["lam"] - thr) <= float(eng.params.guard_lam))
            if self._CUT_ROWS_FROM_LAPACK:
                fix[:] = True
            if fix.any():
                csr = neartie.fix_cut_rows(csr, packed, fix, self._nb_vars, vars_values, thr)
            if _add_rows_csr(my_prob, csr):
                return len(csr["rhs"])
            coeffs_sdp, rhs_sdp = _sparse_pairs(csr), csr["rhs"].tolist()
            senses_sdp = ["G"] * len(rhs_sdp)
        my_prob.linear_constraints.add(lin_expr=coeffs_sdp, rhs=rhs_sdp, senses=senses_sdp)
        return len(rhs_sdp)

    def _get_eigendecomp(self, dim_subpr, curr_pt, X_slice, ev_yes):
        vals, vecs = self._any_engine().eigendecomp(dim_subpr, curr_pt, X_slice, want_vecs=bool(ev_yes))
        return (vals, vecs) if ev_yes else vals

    def _dense_eigcuts(self, vars_values=None):
        """Strat 0 (cut_select_qp.py:757-786): one dense row per negative eigenvalue of the full [1 x^T; x X]."""
        eng = getattr(self, "_dense_engine", None)
        # the reference re-uses one solver object for many instances (generate_figs_tables.py:319-373): the context is
        # only valid for the instance it was created for
        if eng is None or eng.n != self._nb_vars or getattr(self, "_dense_engine_Q", None) is not self._Q_arr:
            self._dense_engine, self._dense_engine_Q = self._new_engine(), self._Q_arr
        d = self._dense_engine.dense_eigcuts(vars_values)
        # a repeated negative eigenvalue (LP vertices have them) leaves the individual eigenvectors -- hence the rows --
        # undetermined: take them from numpy eigh like the reference does; likewise an eigenvalue within the guard of the
        # threshold, or when rows bit-identical to the reference's are asked for
        ev, thr = d["eigvals"], float(self._THRES_NEG_EIGVAL)
        m_neg = int((ev[:self._nb_vars] < thr + self._GUARD_LAM).sum())
        near = m_neg and (np.diff(ev[:min(m_neg + 1, ev.size)]) <= self._GAP_TOL).any()
        if self._CUT_ROWS_FROM_LAPACK or near or (np.abs(ev - thr) <= self._GUARD_LAM).any():
            val, rhs, _ = neartie.lapack_dense_rows(self._nb_vars, np.asarray(vars_values, dtype=np.float64), thr)
            d = dict(d, val=val, rhs=rhs)
        nb, width = d["val"].shape
        csr = dict(rowptr=np.arange(nb + 1, dtype=np.int64) * width, ind=np.tile(d["ind"], nb), val=d["val"].ravel(), rhs=d["rhs"])
        if not _add_rows_csr(self._my_prob, csr):
            ind = d["ind"].tolist()
            self._my_prob.linear_constraints.add(lin_expr=[SparsePair(ind=ind, val=row.tolist()) for row in d["val"]],
                                                 rhs=d["rhs"].tolist(), senses=["G"] * nb)
        return nb

    # name-mangled exactly like the reference's private triangle methods (_CutSolver__...)
    def _tri_preprocess(self):
        n = self._nb_vars
        eng = self._new_engine()
        eng.set_tri_pattern(_as_dense_adj(self._Q_adj, n))
        self._tri_engine = eng
        self._rank_list_tri, self._idx_list_tri = None, None     # the per-triple lists are not materialised

    def _tri_separate(self, sel_size, vars_values):
        if self._tri_engine is None or self._tri_engine.n != self._nb_vars:
            self._tri_preprocess()
        n, nb_lifted, my_prob = self._nb_vars, self._nb_lifted, self._my_prob
        t = self._tri_engine.triangles(vars_values, self._TRI_CUTS_PER_ROUND_MAX)
        V = t["n_violated"]
        nb_tri_cuts = max(min(self._TRI_CUTS_PER_ROUND_MIN, int(np.floor(sel_size * V))),
                          min(self._TRI_CUTS_PER_ROUND_MAX, V))                    # cut_select_qp.py:843-844
        if nb_tri_cuts > V:
            raise IndexError("list index out of range")                            # the reference indexes past the list here
        csr = _capi.triangle_rows_csr(n, t["rank"][:nb_tri_cuts], t["type"][:nb_tri_cuts])      # cut_select_qp.py:846-860
        if _add_rows_csr(my_prob, csr):
            return nb_tri_cuts
        coeffs_tri, rhs_tri = _sparse_pairs(csr, as_int=True), csr["rhs"].astype(np.int64).tolist()
        senses_tri = ["G"] * nb_tri_cuts
        my_prob.linear_constraints.add(lin_expr=coeffs_tri, rhs=rhs_tri, senses=senses_tri)
        return nb_tri_cuts


def _add_rows_csr(my_prob, csr):
    """One-shot row emission: a sink that offers ``linear_constraints.add_rows_csr(rowptr, ind, val, rhs, senses)``
    (e.g. a thin CPXaddrows wrapper) receives the CSR arrays as they come from the C ABI."""
    add = getattr(my_prob.linear_constraints, "add_rows_csr", None)
    if add is None:
        return False
    add(csr["rowptr"], csr["ind"], csr["val"], csr["rhs"], "G" * len(csr["rhs"]))
    return True


def _sparse_pairs(csr, as_int=False):
    """CSR -> the reference's list of cplex.SparsePair (cut_select_qp.py:747, 849-858), sliced from two flat lists."""
    ind, ptr = csr["ind"].tolist(), csr["rowptr"].tolist()
    val = csr["val"].astype(np.int64).tolist() if as_int else csr["val"].tolist()
    return [SparsePair(ind=ind[a:b], val=val[a:b]) for a, b in zip(ptr[:-1], ptr[1:])]


class _NNFunc(object):
    """Callable with the signature of NNs.so's neural_net_dD (cut_select_qp.py:579-582), evaluated on the GPU."""

    def __init__(self, owner, d):
        self.owner, self.d, self.eng = owner, d, None
        self.restype = None

    def __call__(self, input_arr):
        if self.eng is None:
            self.eng = _capi.Engine(self.owner._DEVICE)
            self.eng.set_weights(self.d, self.owner._blobs[self.d])
        return float(self.eng.nn_eval(self.d, np.asarray(list(input_arr), dtype=np.float64))[0])


def _as_dense_adj(Q_adj, n):
    """numpy 0/1 pattern from a numpy array or a cvxopt spmatrix (cut_select_qp.py:323-326)."""
    if isinstance(Q_adj, np.ndarray):
        return (Q_adj != 0).astype(np.uint8)
    A = np.zeros((n, n), dtype=np.uint8)
    try:
        for i, j in zip(Q_adj.I, Q_adj.J):       # cvxopt.spmatrix
            A[int(i), int(j)] = 1
    except AttributeError:
        for i in range(n):
            for j in range(n):
                A[i, j] = 1 if Q_adj[i, j] else 0
    return A


def _is_complete(adj):
    n = adj.shape[0]
    off = (adj | adj.T).astype(bool)
    off[np.arange(n), np.arange(n)] = True
    return bool(off.all())


class CutSolver(B200CutSelection):
    """Standalone solver object exposing the reference's hot-path surface (no LP loop)."""

    def __init__(self):
        super(CutSolver, self).__init__()

    def __gen_dense_eigcuts(self, vars_values=None):   # -> _CutSolver__gen_dense_eigcuts
        return self._dense_eigcuts(vars_values)

    def __preprocess_triangle_ineq(self):          # -> _CutSolver__preprocess_triangle_ineq
        return self._tri_preprocess()

    def __separate_and_add_triangle(self, sel_size, vars_values):   # -> _CutSolver__separate_and_add_triangle
        return self._tri_separate(sel_size, vars_values)
