"""ctypes binding of libsdpcutsel.so (include/sdpcutsel.h). No torch types cross this boundary.

The library is the only compute path: if it is missing or no sm_100a device is present, construction
raises -- there is no CPU fallback.
"""
import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDPCS_LIB") or os.path.join(_PKG, "libsdpcutsel.so")   # SDPCS_LIB: alternative build (trace / experiment libraries)

c_i64, c_dbl, c_int, c_vp = ctypes.c_int64, ctypes.c_double, ctypes.c_int, ctypes.c_void_p
P = ctypes.POINTER

EXPORTS = [
    "sdpcs_default_params", "sdpcs_create", "sdpcs_destroy", "sdpcs_last_error", "sdpcs_set_stream", "sdpcs_set_params",
    "sdpcs_get_timings", "sdpcs_set_weights", "sdpcs_set_instance", "sdpcs_set_cover_all", "sdpcs_set_cover_list",
    "sdpcs_num_candidates", "sdpcs_score", "sdpcs_scores", "sdpcs_counts", "sdpcs_topk", "sdpcs_merge_topk",
    "sdpcs_select", "sdpcs_unrank", "sdpcs_binom", "sdpcs_gen_cuts", "sdpcs_eigendecomp", "sdpcs_set_tri_pattern",
    "sdpcs_triangles", "sdpcs_nn_eval", "sdpcs_nn_debug_layer", "sdpcs_fp64_peak",
    "sdpcs_set_cover_pattern", "sdpcs_get_cover_rows", "sdpcs_cover_restrict", "sdpcs_gen_cuts_csr", "sdpcs_triangle_rows_csr", "sdpcs_dense_eigcuts", "sdpcs_max_pos_nonviolated",
    "sdpcs_last_band", "sdpcs_topk_pack_dev", "sdpcs_merge_packed_dev", "sdpcs_cover_filter", "sdpcs_sdp_solve",
]

NN_TCGEN05, NN_DMMA, NN_SCREEN = 0, 1, 2


class Params(ctypes.Structure):
    _fields_ = [("thres_min_opt", c_dbl), ("thres_neg_eigval", c_dbl), ("big_m", c_dbl), ("thres_tri_viol", c_dbl),
                ("thres_tri_dense", ctypes.c_int32), ("jacobi_sweeps", ctypes.c_int32), ("nn_engine", ctypes.c_int32),
                ("nn_fused_prep", ctypes.c_int32), ("guard_lam", c_dbl), ("guard_obj", c_dbl), ("band_cap", c_i64), ("sdp_mu_final", c_dbl)]


class Timings(ctypes.Structure):
    _fields_ = [("score_ms", c_dbl), ("select_ms", c_dbl), ("h2d_ms", c_dbl), ("score_launches", c_i64),
                ("select_launches", c_i64), ("nn_fallbacks", c_i64), ("nn_ms", c_dbl)]


_lib = None


def load_library(path=None):
    """dlopen libsdpcutsel.so (built in-tree by build.py); raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError("libsdpcutsel.so not found at %s: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA library is the only compute path; there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    lib.sdpcs_last_error.restype = ctypes.c_char_p
    lib.sdpcs_last_error.argtypes = [c_vp]
    for name in EXPORTS:
        if name != "sdpcs_last_error":
            getattr(lib, name).restype = c_int
    _lib = lib
    return lib


class SdpcsError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(c_vp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Engine(object):
    """One GPU context of libsdpcutsel (one per process / per cover)."""

    def __init__(self, device=0):
        self._lib = load_library()
        self._ctx = c_vp()
        rc = self._lib.sdpcs_create(ctypes.byref(self._ctx), c_int(device))
        if rc != 0:
            msg = self._lib.sdpcs_last_error(None)
            self._ctx = None
            raise SdpcsError("sdpcs_create failed (%d): %s" % (rc, msg.decode() if msg else ""))
        self._params = None
        self.n = 0
        self.big_m = 1000.0            # sdpcs_default_params
        self.rho = 0
        self.device = device

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.sdpcs_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self._lib.sdpcs_last_error(self._ctx)
            raise SdpcsError("libsdpcutsel error %d: %s" % (rc, msg.decode() if msg else ""))

    # -- setup -------------------------------------------------------------------------------------
    def set_stream(self, stream_ptr):
        self._ck(self._lib.sdpcs_set_stream(self._ctx, c_vp(stream_ptr)))

    def set_params(self, **kw):
        """Update algorithmic / engine parameters (sdpcs_params); fields not named keep their current value."""
        if self._params is None:
            self._params = Params()
            self._lib.sdpcs_default_params(ctypes.byref(self._params))
        p = self._params
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError("sdpcs_params has no field %r" % k)
            setattr(p, k, v)
        self._ck(self._lib.sdpcs_set_params(self._ctx, ctypes.byref(p)))
        self.big_m = float(p.big_m)

    @property
    def params(self):
        if self._params is None:
            self.set_params()
        return self._params

    def set_weights(self, rho, blob):
        blob = _f64(blob)
        self._ck(self._lib.sdpcs_set_weights(self._ctx, c_int(rho), _ptr(blob), c_i64(blob.size)))

    def set_instance(self, n, Q_arr):
        Q_arr = _f64(Q_arr)
        if Q_arr.size != n * (n + 1) // 2:
            raise ValueError("Q_arr must hold n(n+1)/2 values")
        self._ck(self._lib.sdpcs_set_instance(self._ctx, c_int(n), _ptr(Q_arr)))
        self.n = n

    def set_cover_all(self, rho, rank_begin=0, rank_end=-1):
        self._ck(self._lib.sdpcs_set_cover_all(self._ctx, c_int(rho), c_i64(rank_begin), c_i64(rank_end)))
        self.rho = rho

    def set_cover_list(self, rho, idx, agg_offset=0):
        idx = np.ascontiguousarray(idx, dtype=np.int16).reshape(-1, rho)
        self._ck(self._lib.sdpcs_set_cover_list(self._ctx, c_int(rho), _ptr(idx), c_i64(idx.shape[0]), c_i64(agg_offset)))
        self.rho = rho

    def set_cover_pattern(self, rho, adj, agg_offset=0):
        """P^E_rho built on the device from the n x n sparsity pattern (cut_select_qp.py:401-522); returns N."""
        adj = np.ascontiguousarray(np.asarray(adj) != 0, dtype=np.uint8)
        if adj.shape != (self.n, self.n):
            raise SdpcsError("adjacency must be n x n")
        N = c_i64()
        self._ck(self._lib.sdpcs_set_cover_pattern(self._ctx, c_int(rho), _ptr(adj), c_i64(agg_offset), ctypes.byref(N)))
        self.rho = rho
        return N.value

    def cover_restrict(self, begin, end):
        """Keep the candidates [begin, end) of the current cover (this rank's shard); agg_idx values are unchanged."""
        self._ck(self._lib.sdpcs_cover_restrict(self._ctx, c_i64(begin), c_i64(end)))

    def cover_filter(self, other, keep_members=True):
        """Cover algebra on the device (sdpcs_cover_filter): keep the candidates that (do not) occur in `other`'s list cover."""
        N = c_i64()
        self._ck(self._lib.sdpcs_cover_filter(self._ctx, other._ctx, c_int(1 if keep_members else 0), ctypes.byref(N)))
        return N.value

    def cover_rows(self):
        """The current list cover as (N, rho) int16 rows padded with -1 (agg_list[i][0] of the reference)."""
        out = np.empty((self.num_candidates, self.rho), dtype=np.int16)
        self._ck(self._lib.sdpcs_get_cover_rows(self._ctx, _ptr(out), c_i64(out.shape[0])))
        return out

    @property
    def num_candidates(self):
        N = c_i64()
        self._ck(self._lib.sdpcs_num_candidates(self._ctx, ctypes.byref(N)))
        return N.value

    # -- scoring / selection -----------------------------------------------------------------------
    def _vars(self, vars_values, required=False):
        if vars_values is None:          # re-use the LP point resident on the device
            if required:
                raise ValueError("vars_values is required here")
            return None
        v = _f64(vars_values)
        if v.size != self.n * (self.n + 1) // 2 + self.n:
            raise ValueError("vars_values must hold n(n+1)/2 + n values")
        return v

    def score(self, vars_values, want=3):
        v = self._vars(vars_values)
        self._ck(self._lib.sdpcs_score(self._ctx, _ptr(v), c_int(want)))

    def scores(self, i0=0, i1=None, lam=True, obj=True):
        i1 = self.num_candidates if i1 is None else i1
        ol = np.empty(i1 - i0) if lam else None
        oo = np.empty(i1 - i0) if obj else None
        self._ck(self._lib.sdpcs_scores(self._ctx, c_i64(i0), c_i64(i1), _ptr(ol), _ptr(oo)))
        return ol, oo

    def counts(self):
        out = np.zeros(3, dtype=np.int64)
        self._ck(self._lib.sdpcs_counts(self._ctx, _ptr(out)))
        return out

    def topk(self, mode, k, pivot_obj=0.0, pivot_idx=0, all_walked=0):
        k = int(k)
        kk = max(k, 1)
        idx, sc, lam, obj = np.empty(kk, np.int64), np.empty(kk), np.empty(kk), np.empty(kk)
        n = c_i64()
        self._ck(self._lib.sdpcs_topk(self._ctx, c_int(mode), c_i64(k), c_dbl(pivot_obj), c_i64(pivot_idx), c_int(all_walked),
                                      _ptr(idx), _ptr(sc), _ptr(lam), _ptr(obj), ctypes.byref(n)))
        m = n.value
        return idx[:m], sc[:m], lam[:m], obj[:m]

    def last_band(self, cap=None):
        """Guard band of the last topk / select pass (sdpcs_last_band): the near ties of the k-th score that follow the
        winners, in selection order, plus the guard counters."""
        info = np.zeros(4, dtype=np.int64)
        n = c_i64()
        self._ck(self._lib.sdpcs_last_band(self._ctx, c_i64(0), None, None, None, None, ctypes.byref(n), _ptr(info)))
        m = int(info[0]) if cap is None else min(int(info[0]), int(cap))
        kk = max(m, 1)
        idx, sc, lam, obj = np.empty(kk, np.int64), np.empty(kk), np.empty(kk), np.empty(kk)
        if m > 0:
            self._ck(self._lib.sdpcs_last_band(self._ctx, c_i64(m), _ptr(idx), _ptr(sc), _ptr(lam), _ptr(obj), ctypes.byref(n), _ptr(info)))
            m = n.value
        return dict(idx=idx[:m], score=sc[:m], lam=lam[:m], obj=obj[:m], n_band=int(info[0]), band_open=int(info[1]),
                    n_unc_lam=int(info[2]), n_unc_obj=int(info[3]))

    def topk_pack_dev(self, mode, k, band_rows, block_ptr, pivot_obj=0.0, pivot_idx=0, all_walked=0):
        """sdpcs_topk whose result stays on the device, packed into the DEVICE buffer at block_ptr ((2 + k + band_rows) x 4 f64)."""
        self._ck(self._lib.sdpcs_topk_pack_dev(self._ctx, c_int(mode), c_i64(int(k)), c_dbl(pivot_obj), c_i64(int(pivot_idx)),
                                               c_int(all_walked), c_i64(int(band_rows)), c_vp(block_ptr)))

    def merge_packed_dev(self, gathered_ptr, world, rows_cap, k, use_obj2, delta, out_cap):
        """Merge `world` gathered blocks on the device; returns (idx, score, lam, obj, n_winners, n_band, hdr[8])."""
        cap = max(int(out_cap), 1)
        idx, sc, lam, obj = np.empty(cap, np.int64), np.empty(cap), np.empty(cap), np.empty(cap)
        n, nb, hdr = c_i64(), c_i64(), np.zeros(8)
        self._ck(self._lib.sdpcs_merge_packed_dev(self._ctx, c_vp(gathered_ptr), c_int(world), c_i64(int(rows_cap)), c_i64(int(k)),
                                                  c_int(1 if use_obj2 else 0), c_dbl(delta), c_i64(int(out_cap)), _ptr(idx), _ptr(sc),
                                                  _ptr(lam), _ptr(obj), ctypes.byref(n), ctypes.byref(nb), _ptr(hdr)))
        m = n.value + nb.value
        return idx[:m], sc[:m], lam[:m], obj[:m], n.value, nb.value, hdr

    def max_pos_nonviolated(self):
        out = c_dbl()
        self._ck(self._lib.sdpcs_max_pos_nonviolated(self._ctx, ctypes.byref(out)))
        return out.value

    def merge_topk(self, score, obj2, idx, k):
        score, idx = _f64(score), np.ascontiguousarray(idx, dtype=np.int64)
        obj2 = None if obj2 is None else _f64(obj2)
        perm = np.empty(max(min(k, score.size), 1), np.int64)
        n = c_i64()
        self._ck(self._lib.sdpcs_merge_topk(self._ctx, c_i64(score.size), _ptr(score), _ptr(obj2), _ptr(idx), c_i64(k),
                                            _ptr(perm), ctypes.byref(n)))
        return perm[:n.value]

    def select(self, strat, vars_values, k):
        """One-call selection with host buffers. Returns dict(idx, score, lam, obj, counts, new_strat)."""
        v = self._vars(vars_values)
        k = int(k)
        kk = max(k, 1)
        idx, sc, lam, obj = np.empty(kk, np.int64), np.empty(kk), np.empty(kk), np.empty(kk)
        n, ns = c_i64(), c_int()
        counts = np.zeros(3, dtype=np.int64)
        self._ck(self._lib.sdpcs_select(self._ctx, c_int(strat), _ptr(v), c_i64(k), _ptr(idx), _ptr(sc), _ptr(lam), _ptr(obj),
                                        ctypes.byref(n), _ptr(counts), ctypes.byref(ns)))
        m = n.value
        return dict(idx=idx[:m], score=sc[:m], lam=lam[:m], obj=obj[:m], counts=counts, new_strat=ns.value)

    def timings(self):
        t = Timings()
        self._ck(self._lib.sdpcs_get_timings(self._ctx, ctypes.byref(t)))
        return dict(score_ms=t.score_ms, select_ms=t.select_ms, h2d_ms=t.h2d_ms, score_launches=t.score_launches,
                    select_launches=t.select_launches, nn_fallbacks=t.nn_fallbacks, nn_ms=t.nn_ms)

    # -- cuts / eig / triangles / nn ---------------------------------------------------------------
    def gen_cuts(self, rho, sets, vars_values, with_gap=False):
        sets = np.ascontiguousarray(sets, dtype=np.int16).reshape(-1, rho)
        m = sets.shape[0]
        width = rho + rho * (rho + 1) // 2
        ind, val = np.empty((m, width), np.int64), np.empty((m, width))
        rhs, lam, viol, gap = np.empty(m), np.empty(m), np.empty(m, np.uint8), np.empty(m)
        v = self._vars(vars_values, required=True)
        self._ck(self._lib.sdpcs_gen_cuts(self._ctx, c_int(rho), _ptr(sets), c_i64(m), _ptr(v), _ptr(ind), _ptr(val), _ptr(rhs),
                                          _ptr(lam), _ptr(viol), _ptr(gap)))
        if with_gap:
            return ind, val, rhs, lam, viol.astype(bool), gap
        return ind, val, rhs, lam, viol.astype(bool)

    def gen_cuts_csr(self, rho, sets, vars_values):
        """Eigenvector cuts as CSR rows: dict(rowptr, ind, val, rhs, src, lam, gap); sense >= for every row.  Rows exist
        for lam_min < thres_neg_eigval + guard_lam; lam / gap let the caller settle the near-threshold and the
        non-unique-eigenvector cases with the reference's arithmetic."""
        sets = np.ascontiguousarray(sets, dtype=np.int16).reshape(-1, rho)
        m = sets.shape[0]
        width = rho + rho * (rho + 1) // 2
        rowptr, ind, val = np.zeros(m + 1, np.int64), np.empty(m * width, np.int64), np.empty(m * width)
        rhs, src, lam, gap, nrows = np.empty(m), np.empty(m, np.int64), np.empty(m), np.empty(m), c_i64()
        v = self._vars(vars_values, required=True)
        self._ck(self._lib.sdpcs_gen_cuts_csr(self._ctx, c_int(rho), _ptr(sets), c_i64(m), _ptr(v), _ptr(rowptr), _ptr(ind), _ptr(val),
                                              _ptr(rhs), _ptr(src), _ptr(lam), _ptr(gap), ctypes.byref(nrows)))
        r = nrows.value
        nnz = int(rowptr[r])
        return dict(rowptr=rowptr[:r + 1], ind=ind[:nnz], val=val[:nnz], rhs=rhs[:r], src=src[:r], lam=lam[:r], gap=gap[:r])

    def dense_eigcuts(self, vars_values):
        """Strat 0 (cut_select_qp.py:757-786): dict(eigvals (n+1,), ind (width,), val (ncuts, width), rhs (ncuts,))."""
        v = self._vars(vars_values, required=True)
        n = self.n
        nb_lifted = n * (n + 1) // 2
        width = n + nb_lifted
        eig, val, rhs, nc = np.empty(n + 1), np.empty((n, width)), np.empty(n), c_i64()
        self._ck(self._lib.sdpcs_dense_eigcuts(self._ctx, _ptr(v), c_i64(n), _ptr(eig), ctypes.byref(nc), _ptr(val), _ptr(rhs)))
        ind = np.concatenate([np.arange(nb_lifted, nb_lifted + n), np.arange(nb_lifted)]).astype(np.int64)
        return dict(eigvals=eig, ind=ind, val=val[:nc.value], rhs=rhs[:nc.value])

    def eigendecomp(self, d, curr_pt, X_slice, want_vecs=True):
        pt, Xs = _f64(curr_pt), _f64(X_slice)
        vals = np.empty(d + 1)
        vecs = np.empty((d + 1, d + 1)) if want_vecs else None
        self._ck(self._lib.sdpcs_eigendecomp(self._ctx, c_int(d), _ptr(pt), _ptr(Xs), _ptr(vals), _ptr(vecs)))
        return vals, vecs

    def set_tri_pattern(self, adj):
        a = None if adj is None else np.ascontiguousarray(adj, dtype=np.uint8)
        self._ck(self._lib.sdpcs_set_tri_pattern(self._ctx, _ptr(a)))

    def triangles(self, vars_values, kmax):
        v = self._vars(vars_values, required=True)
        kk = max(int(kmax), 1)
        rank, typ, viol, dens = np.empty(kk, np.int64), np.empty(kk, np.int8), np.empty(kk), np.empty(kk, np.int8)
        n, nv, nt = c_i64(), c_i64(), c_i64()
        self._ck(self._lib.sdpcs_triangles(self._ctx, _ptr(v), c_i64(int(kmax)), _ptr(rank), _ptr(typ), _ptr(viol), _ptr(dens),
                                           ctypes.byref(n), ctypes.byref(nv), ctypes.byref(nt)))
        m = n.value
        return dict(rank=rank[:m], type=typ[:m], viol=viol[:m], density=dens[:m], n_violated=nv.value, n_triples=nt.value)

    def nn_eval(self, rho, inputs):
        x = _f64(inputs).reshape(-1, rho * (rho + 3) // 2)
        out = np.empty(x.shape[0])
        self._ck(self._lib.sdpcs_nn_eval(self._ctx, c_int(rho), _ptr(x), c_i64(x.shape[0]), _ptr(out)))
        return out

    def sdp_solve(self, d, x, C_triu, with_iters=False):
        """Exact SDP values v(x, C) for m sub-problems of size d (sdpcs_sdp_solve): x (m, d), C_triu (m, d(d+1)/2) upper
        triangle row-major with <C, X> = sum_{i<=j} C_ij X_ij."""
        rows = np.ascontiguousarray(np.concatenate([_f64(x).reshape(-1, d), _f64(C_triu).reshape(-1, d * (d + 1) // 2)], axis=1))
        m = rows.shape[0]
        out, its = np.empty(m), np.empty(m, np.int32)
        self._ck(self._lib.sdpcs_sdp_solve(self._ctx, c_int(d), _ptr(rows), c_i64(m), _ptr(out), _ptr(its)))
        return (out, its) if with_iters else out

    def nn_debug_layer(self, rho, inputs, layer):
        """Test hook: scaled pre-activations of tansig layer `layer` from the tcgen05 engine, (m, 64)."""
        x = _f64(inputs).reshape(-1, rho * (rho + 3) // 2)
        out = np.empty((x.shape[0], 64))
        self._ck(self._lib.sdpcs_nn_debug_layer(self._ctx, c_int(rho), _ptr(x), c_i64(x.shape[0]), c_int(layer), _ptr(out)))
        return out

    def fp64_peak(self):
        a, b = c_dbl(), c_dbl()
        self._ck(self._lib.sdpcs_fp64_peak(self._ctx, ctypes.byref(a), ctypes.byref(b)))
        return dict(dfma_tflops=a.value, dmma_tflops=b.value)


def triangle_rows_csr(n, triple_rank, types):
    """CSR rows of the triangle inequalities (cut_select_qp.py:846-860): dict(rowptr, ind, val, rhs)."""
    lib = load_library()
    rank = np.ascontiguousarray(triple_rank, dtype=np.int64).ravel()
    typ = np.ascontiguousarray(types, dtype=np.int8).ravel()
    m = rank.size
    rowptr, ind, val, rhs = np.zeros(m + 1, np.int64), np.empty(6 * m, np.int64), np.empty(6 * m), np.empty(m)
    rc = lib.sdpcs_triangle_rows_csr(c_int(n), _ptr(rank), _ptr(typ), c_i64(m), _ptr(rowptr), _ptr(ind), _ptr(val), _ptr(rhs))
    if rc != 0:
        raise SdpcsError("sdpcs_triangle_rows_csr failed (%d)" % rc)
    nnz = int(rowptr[m])
    return dict(rowptr=rowptr, ind=ind[:nnz], val=val[:nnz], rhs=rhs)


def unrank(n, rho, ranks):
    lib = load_library()
    ranks = np.ascontiguousarray(ranks, dtype=np.int64).ravel()
    out = np.empty((ranks.size, rho), dtype=np.int32)
    rc = lib.sdpcs_unrank(c_int(n), c_int(rho), _ptr(ranks), c_i64(ranks.size), _ptr(out))
    if rc != 0:
        raise SdpcsError("sdpcs_unrank failed (%d)" % rc)
    return out


def binom(n, k):
    lib = load_library()
    out = c_i64()
    if lib.sdpcs_binom(c_int(n), c_int(k), ctypes.byref(out)) != 0:
        raise SdpcsError("sdpcs_binom: bad arguments")
    return out.value
