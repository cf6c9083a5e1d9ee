"""Near-tie resolution (hard part 1): make the selected prefix identical to the reference's even where
scores agree to rounding noise.

The reference ranks with ``list.sort`` on scores that come out of LAPACK ``dsyevd`` (``np.linalg.eigvalsh(M, "U")``,
cut_select_qp.py:647, 796-797) and of the MATLAB-Coder NN with libm ``exp`` (cut_select_qp.py:575-582).  On LP vertices
(x = 0.5, X in {0, 0.5}) thousands of candidates share one exact score and the reference's order inside such a class is
its own round-off.  The device scores agree with the reference's to ~1e-14 (lam) / ~1e-10 (obj), so the device order is
the reference's order except inside runs of candidates closer than that.  The device therefore returns the k winners
PLUS every candidate within a guard band of the k-th score (``sdpcs_last_band``); this module

  * finds the runs of consecutive entries closer than the guard (and the entries whose violated / positive
    classification lies inside the guard of a threshold),
  * re-scores only those with the reference's own arithmetic -- the same numpy call for eigenvalues, a restatement of
    the generated ``neural_net_dD`` function for the NN (k-ascending sums, no FMA, libm exp) --
  * re-sorts with the reference's stable rule (score desc, agg_idx asc) and cuts at k.

Entries outside runs keep their device scores.  Nothing here imports the oracle; this is part of the product's
selection semantics, O(near ties) work, zero on non-degenerate points (``degenerate`` = 0).
"""
import math

import numpy as np

from . import nn_weights

_TRIU = {}
_PAIRS = {}
_exp = np.frompyfunc(math.exp, 1, 1)          # libm exp, the function NNs.so calls (numpy's SIMD exp may round differently)


def _pairs(d):
    if d not in _PAIRS:
        _PAIRS[d] = np.array([(a, b) for a in range(d) for b in range(a, d)], dtype=np.int64)   # combinations_with_replacement
    return _PAIRS[d]


def lapack_lam_min(x_s, X_s):
    """np.linalg.eigvalsh(M, 'U')[0] of M = [[1, x^T], [x, X]] filled as cut_select_qp.py:64-68, 792-797 (upper triangle
    only; the lower one stays zero exactly as in the reference's preformed matrices).  Batched over rows."""
    x_s = np.asarray(x_s, dtype=np.float64)
    m, d = x_s.shape
    if m == 0:
        return np.zeros(0)
    M = np.zeros((m, d + 1, d + 1))
    M[:, 0, 0] = 1.0
    M[:, 0, 1:] = x_s
    if d not in _TRIU:
        _TRIU[d] = np.triu_indices(d)
    iu = _TRIU[d]
    M[:, 1 + iu[0], 1 + iu[1]] = X_s
    return np.linalg.eigvalsh(M, "U")[:, 0]


def nn_exact(blob, inputs):
    """neural_net_dD(input) in the arithmetic of the generated code (neural_net_3D.m:47-85): mapminmax, tansig layers
    a = 2 / (1 + exp(-2 n)) - 1 with n = b + W a (sum over k ascending, one rounding per product and per add), linear
    output layer, reverse mapminmax."""
    net = nn_weights.unpack_blob(blob)
    x = np.asarray(inputs, dtype=np.float64)
    a = (x - net["x_xoffset"][None, :]) * net["x_gain"][None, :] + -1.0
    L = len(net["W"])
    for l in range(L):
        W, b = net["W"][l], net["b"][l]
        s = np.zeros((a.shape[0], W.shape[0]))
        for k in range(W.shape[1]):
            s = s + W[None, :, k] * a[:, k, None]
        z = b[None, :] + s
        if l == L - 1:
            y = z[:, 0]
            return (y - -1.0) / net["y_gain"] + net["y_xoffset"]
        a = 2.0 / (1.0 + _exp(-2.0 * z).astype(np.float64)) - 1.0


class Rescorer(object):
    """Exact (reference-arithmetic) scores of selected candidates of one instance at one LP point."""

    def __init__(self, n, Q_arr, vars_values, blobs, sets_of, thr_eig=-1e-15, thr_opt=0.0, big_m=1000.0):
        self.n = int(n)
        self.Q_arr = np.asarray(Q_arr, dtype=np.float64)
        v = np.asarray(vars_values, dtype=np.float64)
        nl = self.n * (self.n + 1) // 2
        self.X_vals, self.x_vals = v[:nl], v[nl:]
        self.blobs = blobs                      # {d: blob}
        self.sets_of = sets_of                  # agg_idx array -> (m, rho) int array, -1 padded
        self.thr_eig, self.thr_opt, self.big_m = float(thr_eig), float(thr_opt), float(big_m)

    def _groups(self, idx):
        sets = np.asarray(self.sets_of(np.asarray(idx, dtype=np.int64)), dtype=np.int64)
        if sets.ndim != 2:
            sets = sets.reshape(len(idx), -1)
        sizes = (sets >= 0).sum(axis=1)
        for d in np.unique(sizes):
            sel = np.nonzero(sizes == d)[0]
            sub = sets[sel, :d]
            pt = _pairs(int(d))
            a, b = sub[:, pt[:, 0]], sub[:, pt[:, 1]]
            xinds = self.n * a - a * (a + 1) // 2 + b                      # cut_select_qp.py:531
            yield int(d), sel, sub, xinds

    def lam(self, idx):
        out = np.empty(len(idx))
        for d, sel, sub, xinds in self._groups(idx):
            out[sel] = lapack_lam_min(self.x_vals[sub], self.X_vals[xinds])
        return out

    def obj(self, idx):
        """cut_select_qp.py:575-582: -(sum Q~ X) * max_elem + NN([x | Q~]) * max_elem, Q~ = Q_slice / max_elem with
        max_elem = d * max|Q_slice| (1 if that is 0, :536-538); Python's left-to-right sum."""
        out = np.empty(len(idx))
        for d, sel, sub, xinds in self._groups(idx):
            Qraw = self.Q_arr[xinds]
            max_elem = d * np.abs(Qraw).max(axis=1)
            max_elem = np.where(max_elem == 0, 1.0, max_elem)
            Qs = Qraw / max_elem[:, None]
            Xs, xs = self.X_vals[xinds], self.x_vals[sub]
            s = np.zeros(len(sel))
            for k in range(Qs.shape[1]):
                s = s + Qs[:, k] * Xs[:, k]
            o = (-s) * max_elem
            o = o + nn_exact(self.blobs[d], np.concatenate([xs, Qs], axis=1)) * max_elem
            out[sel] = o
        return out


def tie_runs(score, delta):
    """Mask of the entries of a descending score list that have a neighbour closer than delta."""
    score = np.asarray(score, dtype=np.float64)
    m = score.size
    near = np.zeros(m, dtype=bool)
    if m > 1:
        close = np.abs(np.diff(score)) <= delta
        near[:-1] |= close
        near[1:] |= close
    return near


def _order(score, idx, obj2=None):
    """Reference order: score descending, stable = ties by the earlier order, i.e. (obj2 desc,) agg_idx asc."""
    if obj2 is None:
        return np.lexsort((idx, -score))
    return np.lexsort((idx, -obj2, -score))


def resolve(res, k, rescorer, guard_lam, guard_obj):
    """res: dict from distributed.ShardedSelector.select (winners + band + guard info).  Returns the final dict with the
    first <= k entries in the reference's order plus
        n_near_ties  entries that had to be re-scored,
        degenerate   0: no near tie touched the selection; 1: near ties re-scored and resolved; 2: not resolvable
                     (more near ties than the band holds, or a classification inside the guard that the walk of the
                     combined rule depends on) -- the order is then the device's own (score desc, agg_idx asc).
    """
    strat, path = res["strat"], res["path"]
    g = res["guard"]
    idx = np.concatenate([res["idx"], res["band"]["idx"]])
    score = np.concatenate([res["score"], res["band"]["score"]])
    lam = np.concatenate([res["lam"], res["band"]["lam"]])
    obj = np.concatenate([res["obj"], res["band"]["obj"]])
    out = dict(res)
    out.pop("band", None)
    thr_eig, thr_opt, big_m = rescorer.thr_eig, rescorer.thr_opt, rescorer.big_m
    if strat == 4 and path == 3:
        score = obj.copy()                 # strong-prefix path: the list is ordered by obj (winners carry obj + big_m)
    uses_lam = strat in (1, 4)
    delta = guard_lam if strat == 1 else guard_obj
    near = tie_runs(score, delta) if idx.size else np.zeros(0, dtype=bool)
    unc_l = (np.abs(lam - thr_eig) <= guard_lam) if (uses_lam and guard_lam > 0) else np.zeros(idx.size, dtype=bool)
    unc_o = (np.abs(obj - thr_opt) <= guard_obj) if (strat == 4 and guard_obj > 0) else np.zeros(idx.size, dtype=bool)
    open_band = bool(g["band_open"])
    # classification inside the guard somewhere in the cover but outside what came back: the counters / the walk may differ
    unc_elsewhere = (uses_lam and g["n_unc_lam"] > int(unc_l.sum())) or (strat == 4 and g["n_unc_obj"] > int(unc_o.sum()))
    touched = near | unc_l | unc_o
    n_touch = int(touched.sum())
    out["n_near_ties"] = n_touch
    if n_touch == 0 and not open_band:
        out["degenerate"] = 2 if (strat == 4 and path == 4 and unc_elsewhere) else 0
        if unc_elsewhere and strat == 1 and res["idx"].size < k:
            out["degenerate"] = 2          # the whole violated list was asked for and some classifications are uncertain
        for key in ("idx", "score", "lam", "obj"):
            out[key] = res[key][:k]
        return out
    lam_e, obj_e = lam.copy(), obj.copy()
    if strat == 1:
        need = touched
        lam_e[need] = rescorer.lam(idx[need])
        keep = lam_e < thr_eig
        s = -lam_e
        o = _order(s[keep], idx[keep])[:k]
        sel = np.nonzero(keep)[0][o]
        out.update(idx=idx[sel], score=s[sel], lam=lam_e[sel], obj=obj_e[sel])
        counts = np.array(res["counts"], dtype=np.int64)
        was = lam[unc_l] < thr_eig
        now = lam_e[unc_l] < thr_eig
        counts[1] += int(now.sum()) - int(was.sum())
        out["counts"] = counts
    elif strat == 2:
        obj_e[touched] = rescorer.obj(idx[touched])
        o = _order(obj_e, idx)[:k]
        out.update(idx=idx[o], score=obj_e[o], lam=lam_e[o], obj=obj_e[o])
    elif path == 3:
        # combined rule on the strong-prefix path: the list is the (relaxed) strong set by obj; winners get obj + big_m
        ro = near | unc_o
        obj_e[ro] = rescorer.obj(idx[ro])
        lam_e[unc_l] = rescorer.lam(idx[unc_l])
        keep = (lam_e < thr_eig) & (obj_e > thr_opt)
        o = _order(obj_e[keep], idx[keep])[:k]
        sel = np.nonzero(keep)[0][o]
        out.update(idx=idx[sel], score=obj_e[sel] + big_m, lam=lam_e[sel], obj=obj_e[sel])
        if sel.size < min(k, res["idx"].size):
            out["short"] = int(min(k, res["idx"].size) - sel.size)     # dropped after re-scoring: caller re-selects deeper
    else:
        # general path of the combined rule: re-scored measure of cut_select_qp.py:603-625 with the device's pivot
        pobj, pidx, all_walked = res["pivot"]
        lam_e[touched] = rescorer.lam(idx[touched])
        obj_e[touched] = rescorer.obj(idx[touched])
        walked = np.ones(idx.size, dtype=bool) if all_walked else ((obj_e > pobj) | ((obj_e == pobj) & (idx <= pidx)))
        pos, viol = obj_e > thr_opt, lam_e < thr_eig
        f = obj_e.copy()
        m1, m2, m3 = walked & pos & viol, walked & pos & ~viol, walked & ~pos & viol
        f[m1] = obj_e[m1] + big_m
        f[m2] = obj_e[m2] - big_m
        f[m3] = -lam_e[m3]
        f = np.where(touched, f, score)
        o = _order(f, idx, obj_e)[:k]
        out.update(idx=idx[o], score=f[o], lam=lam_e[o], obj=obj_e[o])
    # strong-prefix path: only candidates up to the pivot matter and all of those came back.  General path of the combined
    # rule: the walk (pivot, +-big_m) depends on every classification, also of candidates that did not come back
    unresolved = open_band or (strat == 4 and path == 4 and (unc_elsewhere or unc_l.any() or unc_o.any()))
    out["degenerate"] = 2 if unresolved else 1
    return out


def lapack_cut_rows(x_s, X_s, thr_eig):
    """Eigenvector cuts of cut_select_qp.py:737-751 in the reference's arithmetic, batched over subsets of one size d:
    w, V = eigh(M, 'U'); v = V[:, 0] with |v_i| <= 1e-15 zeroed; row = [2 v0 v_i | v_i v_j (x 2 if i != j)], rhs = -v0^2.
    Returns (val (m, d + d(d+1)/2), rhs (m,), lam_min (m,), violated (m,))."""
    x_s = np.asarray(x_s, dtype=np.float64)
    m, d = x_s.shape
    M = np.zeros((m, d + 1, d + 1))
    M[:, 0, 0] = 1.0
    M[:, 0, 1:] = x_s
    iu = np.triu_indices(d)
    M[:, 1 + iu[0], 1 + iu[1]] = X_s
    w, V = np.linalg.eigh(M, "U")
    v = V[:, :, 0]
    v = np.where(np.abs(v) <= -thr_eig, 0, v)
    i1 = np.array([a for a in range(d + 1) for b in range(max(a, 1), d + 1)])
    i2 = np.array([b for a in range(d + 1) for b in range(max(a, 1), d + 1)])
    val = v[:, i1] * v[:, i2] * np.where(i1 != i2, 2.0, 1.0)[None, :]
    return val, -v[:, 0] * v[:, 0], w[:, 0], w[:, 0] < thr_eig


def fix_cut_rows(csr, sets, fix, n, vars_values, thr_eig):
    """Replace the CSR rows flagged in `fix` (eigenvector not unique: eigenvalue gap ~ 0, or lam_min inside the guard of the
    violation threshold) by the rows the reference's numpy eigh produces; rows eigh does not find violated are dropped.
    sets: (m, rho) int array (-1 padded) indexed by csr['src']."""
    fix = np.asarray(fix, dtype=bool)
    v = np.asarray(vars_values, dtype=np.float64)
    nl = n * (n + 1) // 2
    X_vals, x_vals = v[:nl], v[nl:]
    rows = np.nonzero(fix)[0]
    sub = np.asarray(sets, dtype=np.int64)[csr["src"][rows]]
    sizes = (sub >= 0).sum(axis=1)
    val, rhs, lam = csr["val"].copy(), csr["rhs"].copy(), csr["lam"].copy()
    keep = np.ones(csr["rhs"].shape[0], dtype=bool)
    for d in np.unique(sizes):
        sel = np.nonzero(sizes == d)[0]
        s = sub[sel, :d]
        pt = _pairs(int(d))
        a, b = s[:, pt[:, 0]], s[:, pt[:, 1]]
        xinds = n * a - a * (a + 1) // 2 + b
        nv, nr, w0, viol = lapack_cut_rows(x_vals[s], X_vals[xinds], thr_eig)
        r = rows[sel]
        val[csr["rowptr"][r][:, None] + np.arange(nv.shape[1])[None, :]] = nv
        rhs[r], lam[r], keep[r] = nr, w0, viol
    out = dict(csr, val=val, rhs=rhs, lam=lam)
    if keep.all():
        return out
    lens = np.diff(csr["rowptr"])
    mask = np.repeat(keep, lens)
    out.update(rowptr=np.concatenate([[0], np.cumsum(lens[keep])]).astype(np.int64), ind=csr["ind"][mask], val=val[mask],
               rhs=rhs[keep], lam=lam[keep], src=csr["src"][keep], gap=csr["gap"][keep])
    return out


def lapack_dense_rows(n, vars_values, thr_eig):
    """Dense eigenvalue cuts (strat 0) in the reference's arithmetic (cut_select_qp.py:757-786): one eigh of the full
    [1 x^T; x X]; every eigenvalue among the n smallest below the threshold gives the row
    [2 v0 v_i | v_i v_j (x 2 if i != j)] >= -v0^2.  Returns (val (ncuts, n + n(n+1)/2), rhs (ncuts,), eigvals)."""
    v = np.asarray(vars_values, dtype=np.float64)
    nl = n * (n + 1) // 2
    mat = np.zeros((n + 1, n + 1))
    mat[0, 0] = 1
    mat[0, 1:] = v[nl:]
    iu = np.triu_indices(n)
    mat[iu[0] + 1, iu[1] + 1] = v[:nl]
    w, V = np.linalg.eigh(mat, "U")
    i1, i2 = np.triu_indices(n + 1)
    keep = i2 >= 1
    i1, i2 = i1[keep], i2[keep]
    cols = [ix for ix in range(n) if w[ix] < thr_eig]
    vals = np.array([V[i1, ix] * V[i2, ix] * np.where(i1 != i2, 2.0, 1.0) for ix in cols]).reshape(len(cols), n + nl)
    rhs = np.array([-V[0, ix] * V[0, ix] for ix in cols])
    return vals, rhs, w
