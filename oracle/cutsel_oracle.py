"""ORACLE -- test infrastructure, NOT product code.

CPU (numpy + a small C helper) restatement of the reference's per-round cut-selection path
(rb2309/SDPCutSel-via-NN, ``cut_select_qp.py`` / ``cut_select_qcqp.py``).  Every function cites the
reference lines it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product package
(``sdpcutsel-via-nn_b200/``) never does.

Pinning (see tests/test_oracle_golden.py): the functions below are checked against fixtures produced
by importing the UNMODIFIED reference in the authoring container (tests/golden/make_golden.py) and
against the reference's committed Fig-8 data; the NN helper (oracle/nn_oracle.c) is bit-identical
to ``neural_nets/NNs.so`` on 20,000 random inputs.

Third-party arithmetic: the reference's eigenvalues come from LAPACK ``dsyevd`` through
``numpy.linalg.eigvalsh/eigh(M, "U")`` (cut_select_qp.py:796-797); the oracle calls the same numpy
entry points (batched calls are bit-identical to per-matrix calls).
"""
import ctypes
import itertools
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_nn.so")

# class constants, cut_select_qp.py:22-41
THRES_MIN_OPT = 0
THRES_NEG_EIGVAL = -10 ** (-15)
BIG_M = 1000
THRES_TRI_DENSE = 2
THRES_TRI_VIOL = 10 ** (-7)
SDP_CUTS_PER_ROUND_MAX = 5000
TRI_CUTS_PER_ROUND_MIN = 5000
TRI_CUTS_PER_ROUND_MAX = 10000


def build_c_helper(force=False):
    """gcc build of oracle/nn_oracle.c -> oracle/_build/liboracle_nn.so (no FMA contraction)."""
    src = os.path.join(_HERE, "nn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


_lib = None


def _c():
    global _lib
    if _lib is None:
        _lib = ctypes.cdll.LoadLibrary(build_c_helper())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------------------------
# instance arrays and synthetic inputs
# ---------------------------------------------------------------------------------------------
def boxqp_arrays(Qf):
    """BoxQP reader semantics, cut_select_qp.py:313-326: Q = -Q_file; Q_arr = upper-tri row-major of Q with
    the diagonal halved; Q_adj[i,j] = 1 iff Q_file[i,j] != 0."""
    Qf = np.asarray(Qf, dtype=np.float64)
    n = Qf.shape[0]
    Qa = -Qf.copy()
    Qa[np.arange(n), np.arange(n)] /= 2.0
    Q_arr = Qa[np.triu_indices(n, k=0)]
    adj = (Qf != 0).astype(np.uint8)
    return Q_arr, adj


def synth_instance(n, density, seed=7):
    """symmetric integer Qf_ij in U{-50..50} kept with probability `density`."""
    rng = np.random.default_rng(seed)
    vals = rng.integers(-50, 51, size=(n, n))
    keep = rng.random((n, n)) < density
    U = np.triu(vals * keep)
    Qf = U + np.triu(U, 1).T
    return Qf.astype(np.float64)


def synth_point(n, seed=8):
    """non-degenerate point inside the McCormick box (cut_select_qp.py:357-373).
    Returns vars_values = [X upper-tri row-major | x] (cut_select_qp.py:547)."""
    rng = np.random.default_rng(seed)
    x = rng.random(n)
    iu = np.triu_indices(n)
    xi, xj = x[iu[0]], x[iu[1]]
    lo = np.maximum(0.0, xi + xj - 1.0)
    hi = np.minimum(xi, xj)
    X = lo + rng.random(lo.size) * (hi - lo)
    return np.concatenate([X, x])


def degenerate_point(n, Q_arr):
    """LP-vertex-like point: x = 0.5, X_ij in {0, 0.5} by sign of Q_arr."""
    X = np.where(Q_arr < 0, 0.5, 0.0)
    return np.concatenate([X, np.full(n, 0.5)])


# ---------------------------------------------------------------------------------------------
# vertex cover / candidate enumeration
# ---------------------------------------------------------------------------------------------
def cover_all(n, rho):
    """All rho-subsets in the order of the nested loops at cut_select_qp.py:451-455
    (= itertools.combinations = lexicographic). Returns int32 (N, rho)."""
    cnt = 1
    for j in range(rho):
        cnt = cnt * (n - j) // (j + 1)
    out = np.fromiter(itertools.chain.from_iterable(itertools.combinations(range(n), rho)),
                      dtype=np.int32, count=cnt * rho)
    return out.reshape(cnt, rho)


def cover_all_window(n, rho, r0, r1):
    """Rows [r0, r1) of cover_all(n, rho) without materialising the rest (lex unranking)."""
    out = np.empty((r1 - r0, rho), dtype=np.int32)
    c = lex_unrank(n, rho, r0)
    for i in range(r1 - r0):
        out[i] = c
        j = rho - 1
        while j >= 0 and c[j] == n - rho + j:
            j -= 1
        if j < 0:
            break
        c[j] += 1
        for t in range(j + 1, rho):
            c[t] = c[t - 1] + 1
    return out


def cover_all_block(n, rho, i1, i2):
    """The subsets of cover_all(n, rho) whose two leading indices are (i1, i2), in enumeration order (the inner loops of
    cut_select_qp.py:451-455 for fixed i1 < i2).  int32 (C(n-1-i2, rho-2), rho)."""
    c = comb(n - 1 - i2, rho - 2)
    tail = np.fromiter(itertools.chain.from_iterable(itertools.combinations(range(i2 + 1, n), rho - 2)),
                       dtype=np.int32, count=c * (rho - 2)).reshape(c, rho - 2)
    out = np.empty((c, rho), dtype=np.int32)
    out[:, 0], out[:, 1], out[:, 2:] = i1, i2, tail
    return out


def comb(n, k):
    if k < 0 or k > n:
        return 0
    r = 1
    for j in range(k):
        r = r * (n - j) // (j + 1)
    return r


def lex_unrank(n, rho, r):
    """rank -> ascending tuple, lexicographic order."""
    c, prev = [], -1
    for j in range(1, rho + 1):
        v = prev + 1
        while True:
            cnt = comb(n - 1 - v, rho - j)
            if r < cnt:
                break
            r -= cnt
            v += 1
        c.append(v)
        prev = v
    return c


def cover_pattern_E_loops(adj, dim):
    """Literal restatement of the nested loops at cut_select_qp.py:401-424 (dim 3), 457-483 (dim 4),
    485-522 (dim 5) for ch_ext = 0. Pure Python, small n only. Returns list of tuples."""
    n = adj.shape[0]
    A = adj
    out = []
    for i1 in range(n):
        for i2 in range(i1 + 1, n):
            if not A[i1, i2]:
                continue
            triple_flag = False
            for i3 in range(i2 + 1, n):
                if A[i1, i3] and A[i2, i3]:
                    triple_flag = True
                    if dim == 3:
                        out.append((i1, i2, i3))
                        continue
                    quad_flag = False
                    for i4 in range(i3 + 1, n):
                        if A[i1, i4] and A[i2, i4] and A[i3, i4]:
                            quad_flag = True
                            if dim == 4:
                                out.append((i1, i2, i3, i4))
                                continue
                            cinq_flag = False
                            for i5 in range(i4 + 1, n):
                                if A[i1, i5] and A[i2, i5] and A[i3, i5] and A[i4, i5]:
                                    out.append((i1, i2, i3, i4, i5))
                                    cinq_flag = True
                            if not cinq_flag:
                                for i5 in range(i4):
                                    if i5 in (i1, i2, i3):
                                        continue
                                    if A[i1, i5] and A[i2, i5] and A[i3, i5] and A[i4, i5]:
                                        cinq_flag = True
                                        break
                                if not cinq_flag:
                                    out.append((i1, i2, i3, i4))
                    if not quad_flag:
                        for i4 in range(i3):
                            if i4 in (i1, i2):
                                continue
                            if A[i1, i4] and A[i2, i4] and A[i3, i4]:
                                quad_flag = True
                                break
                        if not quad_flag:
                            out.append((i1, i2, i3))
            if not triple_flag:
                for i3 in range(i2):
                    if i3 == i1:
                        continue
                    if A[i1, i3] and A[i2, i3]:
                        triple_flag = True
                        break
                if not triple_flag:
                    out.append((i1, i2))
    return out


def cover_pattern_E(adj, dim):
    """Closed form of the same set and order: all dim-cliques of the off-diagonal graph plus
    the maximal cliques of size 2..dim-1 (empty common neighbourhood), sorted as Python tuples.
    Returns (idx int32 (N, dim) padded with -1, sizes int32 (N,))."""
    A = np.array(adj, dtype=bool)
    n = A.shape[0]
    A = A | A.T
    A[np.arange(n), np.arange(n)] = False
    tuples = []
    # cliques of size s as arrays, grown level by level
    level = np.argwhere(np.triu(A, 1))  # size-2 cliques, lex order
    s = 2
    while True:
        if level.size == 0:
            break
        common = np.ones((level.shape[0], n), dtype=bool)
        for c in range(s):
            common &= A[level[:, c]]
        if s == dim:
            tuples.extend(map(tuple, level.tolist()))
            break
        maximal = ~common.any(axis=1)
        tuples.extend(map(tuple, level[maximal].tolist()))
        # extend by a larger last vertex
        ext = common & (np.arange(n)[None, :] > level[:, -1][:, None])
        rows, cols = np.nonzero(ext)
        level = np.concatenate([level[rows], cols[:, None]], axis=1)
        s += 1
    tuples.sort()
    N = len(tuples)
    idx = np.full((N, dim), -1, dtype=np.int32)
    sizes = np.empty(N, dtype=np.int32)
    for i, t in enumerate(tuples):
        idx[i, :len(t)] = t
        sizes[i] = len(t)
    return idx, sizes


def xarr_inds(n, idx_row):
    """Flat upper-tri indices of a subset, cut_select_qp.py:530-531."""
    return [n * a - a * (a + 1) // 2 + b for a, b in itertools.combinations_with_replacement(idx_row, 2)]


def _pair_index_table(d):
    return np.array(list(itertools.combinations_with_replacement(range(d), 2)), dtype=np.int64)


def aggregate(Q_arr, n, idx):
    """Per-candidate aggregation for candidates of ONE size d (idx int (m, d)), cut_select_qp.py:529-539:
    Xarr_inds (m, t), Q_slice scaled (m, t), max_elem (m,)."""
    idx = np.asarray(idx, dtype=np.int64)
    d = idx.shape[1]
    pt = _pair_index_table(d)
    a, b = idx[:, pt[:, 0]], idx[:, pt[:, 1]]
    Xinds = n * a - a * (a + 1) // 2 + b
    Qraw = np.asarray(Q_arr, dtype=np.float64)[Xinds]
    max_elem = d * np.abs(Qraw).max(axis=1)          # lenInds * abs(max(Q_slice, key=abs))
    max_elem = np.where(max_elem == 0, 1.0, max_elem)  # += 1 if 0
    Qs = Qraw / max_elem[:, None]                    # np.divide(Q_slice, max_elem)
    return Xinds, Qs, max_elem


def split_vars(vars_values, n):
    nb_lifted = n * (n + 1) // 2
    v = np.asarray(vars_values, dtype=np.float64)
    return v[:nb_lifted], v[nb_lifted:]      # X_vals, x_vals  (cut_select_qp.py:547)


def eig_matrix(x_s, X_s):
    """M = [[1, x^T],[x, X]] with only the upper triangle filled, cut_select_qp.py:64-68, 792-794."""
    m, d = x_s.shape
    M = np.zeros((m, d + 1, d + 1))
    M[:, 0, 0] = 1.0
    M[:, 0, 1:] = x_s
    iu = np.triu_indices(d)
    M[:, 1 + iu[0], 1 + iu[1]] = X_s
    return M


def lam_min(x_s, X_s):
    """np.linalg.eigvalsh(M, 'U')[0], cut_select_qp.py:647, 796-797."""
    return np.linalg.eigvalsh(eig_matrix(x_s, X_s), "U")[:, 0]


def opt_measure(blob, x_s, X_s, Qs, max_elem):
    """cut_select_qp.py:573-582 through the C helper (left-to-right sum, no FMA, NN bit-equal to NNs.so)."""
    m, d = x_s.shape
    x_s, X_s, Qs = (np.ascontiguousarray(a, dtype=np.float64) for a in (x_s, X_s, Qs))
    max_elem = np.ascontiguousarray(max_elem, dtype=np.float64)
    blob = np.ascontiguousarray(blob, dtype=np.float64)
    out = np.empty(m)
    _c().opt_measure(_p(blob), ctypes.c_int(d), _p(x_s), _p(X_s), _p(Qs), _p(max_elem), ctypes.c_int64(m), _p(out))
    return out


def nn_eval(blob, inputs):
    inputs = np.ascontiguousarray(inputs, dtype=np.float64)
    blob = np.ascontiguousarray(blob, dtype=np.float64)
    out = np.empty(inputs.shape[0])
    _c().nn_oracle_eval(_p(blob), _p(inputs), ctypes.c_int64(inputs.shape[0]), _p(out))
    return out


def score_cover(Q_arr, n, idx, sizes, vars_values, blobs=None, want_lam=True, want_obj=True):
    """lam_min and/or obj_improve for every candidate of a (possibly mixed-size) cover, in cover order."""
    X_vals, x_vals = split_vars(vars_values, n)
    N = idx.shape[0]
    lam = np.full(N, np.nan)
    obj = np.full(N, np.nan)
    for d in np.unique(sizes):
        sel = np.nonzero(sizes == d)[0]
        sub = idx[sel, :d]
        Xinds, Qs, max_elem = aggregate(Q_arr, n, sub)
        x_s, X_s = x_vals[sub], X_vals[Xinds]
        if want_lam:
            lam[sel] = lam_min(x_s, X_s)
        if want_obj:
            obj[sel] = opt_measure(blobs[int(d)], x_s, X_s, Qs, max_elem)
    return lam, obj


# ---------------------------------------------------------------------------------------------
# selection rules
# ---------------------------------------------------------------------------------------------
def stable_desc(score):
    """list.sort(key=itemgetter(1), reverse=True): stable, ties keep ascending original index."""
    return np.argsort(-np.asarray(score, dtype=np.float64), kind="stable")


def select_feas(lam):
    """cut_select_qp.py:639-654: violated iff lam < -1e-15; score -lam; sort desc stable; keep violated.
    Returns (order, scores)."""
    viol = lam < THRES_NEG_EIGVAL
    score = np.where(viol, -lam, 0.0)
    order = stable_desc(score)[: int(viol.sum())]
    return order, score[order]


def select_opt(obj):
    """cut_select_qp.py:599-601. Returns (order, scores) over all candidates."""
    order = stable_desc(obj)
    return order, obj[order]


def select_comb_walk(obj, lam, sel_size):
    """Literal sequential walk of cut_select_qp.py:601-630 (lam must hold eigvalsh values for all candidates,
    the reference computes them lazily for the walked ones only). Returns (new_strat, order, scores)."""
    N = obj.shape[0]
    sel_size = min(sel_size, N)
    order = list(stable_desc(obj))
    score = [float(obj[i]) for i in order]
    strong = viol = 0
    for ix, i in enumerate(order):
        o = score[ix]
        if o > THRES_MIN_OPT and strong < sel_size:
            if lam[i] < THRES_NEG_EIGVAL:
                score[ix] = o + BIG_M
                strong += 1
                viol += 1
            else:
                score[ix] = o - BIG_M
        elif strong < sel_size:
            if lam[i] < THRES_NEG_EIGVAL:
                score[ix] = -float(lam[i])
                viol += 1
        else:
            break
    score = np.array(score)
    order = np.array(order)
    o2 = stable_desc(score)
    new_strat = 1 if strong / sel_size < viol / N else 4
    return new_strat, order[o2], score[o2]


def select_comb(obj, lam, sel_size):
    """Vectorised restatement of the same rule, for sizes where the walk is too slow."""
    N = obj.shape[0]
    sel_size = min(sel_size, N)
    order = stable_desc(obj)
    o, l = obj[order], lam[order]
    violated = l < THRES_NEG_EIGVAL
    pos = o > THRES_MIN_OPT
    strong_mask = pos & violated
    cs = np.cumsum(strong_mask)
    total_strong = int(cs[-1]) if N else 0
    if total_strong >= sel_size:
        pivot = int(np.searchsorted(cs, sel_size))      # position of the sel_size-th strong element
        walked = np.arange(N) <= pivot
        strong = sel_size
    else:
        walked = np.ones(N, dtype=bool)
        strong = total_strong
    score = o.copy()
    m1 = walked & pos & violated
    score[m1] = o[m1] + BIG_M
    m2 = walked & pos & ~violated
    score[m2] = o[m2] - BIG_M
    m3 = walked & ~pos & violated
    score[m3] = -l[m3]
    viol = int((walked & violated).sum())
    o2 = stable_desc(score)
    new_strat = 1 if strong / sel_size < viol / N else 4
    return new_strat, order[o2], score[o2]


# ---------------------------------------------------------------------------------------------
# eigenvector cuts and triangle cuts
# ---------------------------------------------------------------------------------------------
def gen_eigcut(n, idx_row, vars_values):
    """One eigcut, cut_select_qp.py:737-751. Returns None if not violated, else (ind, val, rhs, lam)."""
    X_vals, x_vals = split_vars(vars_values, n)
    d = len(idx_row)
    Xi = xarr_inds(n, idx_row)
    M = eig_matrix(x_vals[list(idx_row)][None, :], X_vals[Xi][None, :])[0]
    w, V = np.linalg.eigh(M, "U")
    if not (w[0] < THRES_NEG_EIGVAL):
        return None
    v = V.T[0]
    v = np.where(abs(v) <= -THRES_NEG_EIGVAL, 0, v)
    val = [v[i] * v[j] * 2 if i != j else v[i] * v[j] for i in range(d + 1) for j in range(max(i, 1), d + 1)]
    nb_lifted = n * (n + 1) // 2
    ind = [int(a) + nb_lifted for a in idx_row] + Xi
    return ind, np.array(val), -v[0] * v[0], w[0]


def triangles_pre(adj, n):
    """cut_select_qp.py:799-821: triples with >= 2 edges, lex order. Returns (triples int32 (T,3), density (T,))."""
    A = np.asarray(adj, dtype=np.float64)
    tri = cover_all(n, 3)
    dens = A[tri[:, 0], tri[:, 1]] + A[tri[:, 0], tri[:, 2]] + A[tri[:, 1], tri[:, 2]]
    keep = dens >= THRES_TRI_DENSE
    return tri[keep], dens[keep]


def triangles_sep(n, triples, dens, vars_values, sel_size_frac):
    """cut_select_qp.py:823-860. Returns arrays (triple_pos, type, violation) for the selected cuts, in order,
    where triple_pos indexes `triples`."""
    X_vals, x_vals = split_vars(vars_values, n)
    i1, i2, i3 = (triples[:, c].astype(np.int64) for c in range(3))
    ix = lambda a, b: n * a - a * (a + 1) // 2 + b
    X1, X2, X4 = X_vals[ix(i1, i2)], X_vals[ix(i1, i3)], X_vals[ix(i2, i3)]
    x1, x2, x3 = x_vals[i1], x_vals[i2], x_vals[i3]
    v = np.empty((triples.shape[0], 4))
    v[:, 0] = X1 + X2 - X4 - x1
    v[:, 1] = X1 - X2 + X4 - x2
    v[:, 2] = -X1 + X2 + X4 - x3
    v[:, 3] = -X1 - X2 - X4 + ((0 + x1) + x2 + x3) - 1
    flat_v = v.ravel()
    flat_d = np.repeat(dens, 4)
    keep = np.nonzero(flat_v >= THRES_TRI_VIOL)[0]
    # sort(key=itemgetter(2, 3), reverse=True) stable: density desc, violation desc, original order
    o = np.lexsort((-flat_v[keep], -flat_d[keep]))   # lexsort is stable, last key primary
    keep = keep[o]
    V = keep.size
    nb = max(min(TRI_CUTS_PER_ROUND_MIN, int(np.floor(sel_size_frac * V))), min(TRI_CUTS_PER_ROUND_MAX, V))
    keep = keep[:nb]
    return keep // 4, keep % 4, flat_v[keep]


def triangle_row(n, triple, typ):
    """Row coefficients, cut_select_qp.py:846-860. Returns (ind, val, rhs)."""
    nb_lifted = n * (n + 1) // 2
    Xi = xarr_inds(n, list(triple))
    coeffs = {0: [-1, -1, 1, 1], 1: [-1, 1, -1, 1], 2: [1, -1, -1, 1], 3: [1, 1, 1, -1, -1, -1]}
    if typ == 3:
        return [Xi[1], Xi[2], Xi[4]] + [int(t) + nb_lifted for t in triple], coeffs[3], -1
    return [Xi[1], Xi[2], Xi[4], int(triple[typ]) + nb_lifted], coeffs[typ], 0


def sel_size_rule(sel, N):
    """cut_select_qp.py:123-125."""
    return min(int(np.floor(sel * N)) if sel < 1 else min(sel, N), SDP_CUTS_PER_ROUND_MAX)


def dense_eigcuts(n, vars_values, thres=THRES_NEG_EIGVAL):
    """Dense eigenvalue cuts (strat 0), cut_select_qp.py:757-786: one eigh of the full [1 x^T; x X] of order n + 1;
    every eigenvalue among the n smallest that is below the threshold gives the row
    [2 v0 v_i (on x_i) | v_i v_j (2 if i != j) (on X_ij, upper-triangular row-major)] >= -v0^2.
    Returns (ind, val (ncuts, n + n(n+1)/2), rhs, eigvals)."""
    nb_lifted = n * (n + 1) // 2
    X_vals, x_vals = np.asarray(vars_values[:nb_lifted]), np.asarray(vars_values[nb_lifted:])
    mat = np.zeros((n + 1, n + 1))
    mat[0, 0] = 1
    mat[0, 1:] = x_vals
    iu = np.triu_indices(n)
    mat[iu[0] + 1, iu[1] + 1] = X_vals
    eigvals, evecs = np.linalg.eigh(mat, "U")
    ind = np.array([i + nb_lifted for i in range(n)] + list(range(nb_lifted)), dtype=np.int64)
    i1, i2 = np.triu_indices(n + 1)
    keep = i2 >= 1                                     # idx2 runs from max(idx1, 1)
    i1, i2 = i1[keep], i2[keep]
    vals, rhs = [], []
    for ix in range(n):
        if eigvals[ix] < thres:
            v = evecs[:, ix]
            vals.append(v[i1] * v[i2] * np.where(i1 != i2, 2.0, 1.0))
            rhs.append(-v[0] * v[0])
    return ind, np.array(vals).reshape(len(rhs), n + nb_lifted), np.array(rhs), eigvals


# ---------------------------------------------------------------------------------------------
# exact optimality measure (strat 3 / figure 8): the rho-dimensional SDP the reference solves with Mosek
# ---------------------------------------------------------------------------------------------
def sdp_exact_value(C_triu, x, mu_final=1e-13):
    """v(x, C) = min <C, X> s.t. [[X, x], [x^T, 1]] PSD, diag(X) <= x  -- the Mosek model of cut_select_qp.py:555-567
    (Z = [[X, x], [x^T, 1]] in the PSD cone, X.diag() <= x, objective sum(Q_sub o X) with Q_sub upper triangular, :592-593)
    and of utilities.py:40-47.  Mosek is not available here, so this is a restatement of the PROBLEM, not of Mosek's
    algorithm: the equivalent dual  x^T C x - min{sum y : y >= 0, D C D + Diag(y) PSD}, D = diag(sqrt(x (1 - x))), solved
    by a barrier method with backtracking line search (numpy, batched), duality gap 2 d mu_final.  Pinned against Mosek's
    own answers in data_figures/fig8_data.csv (tests/test_oracle_golden.py): 1051 sub-problems, max difference 4.3e-6 --
    Mosek's tolerance.  C_triu: (m, d(d+1)/2) upper triangle row-major, <C, X> = sum_{i<=j} C_ij X_ij; x: (m, d)."""
    x = np.asarray(x, dtype=np.float64)
    m, d = x.shape
    iu = np.triu_indices(d)
    C = np.zeros((m, d, d))
    C[:, iu[0], iu[1]] = np.asarray(C_triu, dtype=np.float64)
    C = (C + C.transpose(0, 2, 1)) / 2.0
    s = np.sqrt(np.maximum(x - x * x, 0.0))
    A = C * s[:, :, None] * s[:, None, :]
    const = np.einsum("mi,mij,mj->m", x, C, x)
    y = np.abs(A).sum(axis=2) + 1.0
    eye = np.eye(d)[None]
    mu = 1.0
    while mu > mu_final:
        mu = max(mu * 0.2, mu_final * 0.999)
        for _ in range(50):
            S = A + y[:, :, None] * eye
            Sinv = np.linalg.inv(S)
            g = 1.0 - mu * (np.einsum("mii->mi", Sinv) + 1.0 / y)
            H = mu * (Sinv * Sinv + (1.0 / y ** 2)[:, :, None] * eye)
            dy = -np.linalg.solve(H, g[:, :, None])[:, :, 0]
            dec = -(g * dy).sum(axis=1)
            t = np.ones(m)
            f0 = y.sum(axis=1) - mu * (np.linalg.slogdet(S)[1] + np.log(y).sum(axis=1))
            for _ls in range(60):
                yn = y + t[:, None] * dy
                ok = (yn > 0).all(axis=1)
                Sn = A + yn[:, :, None] * eye
                ok &= np.linalg.eigvalsh(Sn)[:, 0] > 0
                ld = np.linalg.slogdet(np.where(ok[:, None, None], Sn, eye))[1]
                fn = np.where(ok, yn.sum(axis=1) - mu * (ld + np.log(np.where(ok[:, None], yn, 1.0)).sum(axis=1)), np.inf)
                good = fn <= f0 - 0.25 * t * dec
                if good.all():
                    break
                t = np.where(good, t, t * 0.5)
            y = y + t[:, None] * dy
            if (dec / mu < 1e-3).all():
                break
    return const - y.sum(axis=1) + d * mu


def exact_measure(Q_arr, n, idx, sizes, vars_values):
    """obj_improve of strat 3 for every candidate (cut_select_qp.py:575 + 595): (v - <Q~, X>) * max_elem, cover order."""
    X_vals, x_vals = split_vars(vars_values, n)
    out = np.full(idx.shape[0], np.nan)
    for d in np.unique(sizes):
        sel = np.nonzero(sizes == d)[0]
        sub = idx[sel, :d]
        Xinds, Qs, max_elem = aggregate(Q_arr, n, sub)
        cur = np.zeros(sel.size)
        for k in range(Qs.shape[1]):
            cur = cur + Qs[:, k] * X_vals[Xinds][:, k]
        out[sel] = (-cur) * max_elem + sdp_exact_value(Qs, x_vals[sub]) * max_elem
    return out
