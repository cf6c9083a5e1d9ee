"""CPU tests of the near-tie resolution (neartie.py + the band plumbing of distributed.py).

The exact host arithmetic must be bit-identical to the reference's (NNs.so golden vectors, LAPACK call), and the
resolution must turn a device-like selection -- scores perturbed by a few ulps, ties broken by index -- into the
reference's own order on the golden LP vertices (tests/golden: cfg1_deg_*, fig8_*).  The device is emulated by a numpy
engine that restates the selection semantics of select_kernels.cuh (relaxed classification, guard band, counters)."""
import types

import numpy as np
import pytest

import sdpcutsel_via_nn_b200 as pkg
from conftest import inst_arrays
from oracle import cutsel_oracle as orc
from sdpcutsel_via_nn_b200 import neartie
from sdpcutsel_via_nn_b200.distributed import ShardedSelector

THR = -1e-15


class GuardedFakeEngine(object):
    """Numpy restatement of the device selection: modes 1..4 of sdpcs_topk with the guard of sdpcs_params."""
    big_m = 1000.0

    def __init__(self, lam, obj, base=0, guard_lam=1e-12, guard_obj=1e-9, band_cap=65536):
        self.lam, self.obj, self.base = lam, obj, base
        self.params = types.SimpleNamespace(guard_lam=guard_lam, guard_obj=guard_obj, band_cap=band_cap)
        self._counts = np.zeros(3, dtype=np.int64)
        self._band = None

    def score(self, vars_values, want):
        pass

    def counts(self):
        return self._counts

    def max_pos_nonviolated(self):
        m = (self.obj > 0) & ~(self.lam < THR)
        return float(self.obj[m].max()) if m.any() else -np.inf

    def topk(self, mode, k, pivot_obj=0.0, pivot_idx=0, all_walked=0):
        lam, obj, p = self.lam, self.obj, self.params
        idx = self.base + np.arange(lam.size)
        viol, pos = lam < THR, obj > 0
        viol_r, pos_r = lam < THR + p.guard_lam, obj > 0 - p.guard_obj
        self._counts = np.array([lam.size, viol.sum(), (viol & pos).sum()], dtype=np.int64)
        key2 = np.zeros_like(lam)
        if mode == 1:
            valid, key = viol_r, -lam
        elif mode == 2:
            valid, key = np.ones_like(viol), obj
        elif mode == 3:
            valid, key = viol_r & pos_r, obj
        else:
            walked = np.ones_like(viol) if all_walked else (obj > pivot_obj) | ((obj == pivot_obj) & (idx <= pivot_idx))
            key = obj.copy()
            m = walked & pos
            key[m & viol] = obj[m & viol] + 1000
            key[m & ~viol] = obj[m & ~viol] - 1000
            m = walked & ~pos & viol
            key[m] = -lam[m]
            valid, key2 = np.ones_like(viol), obj
        sel = np.nonzero(valid)[0]
        order = sel[np.lexsort((idx[sel], -key2[sel], -key[sel]))]
        win = order[:k]
        delta = p.guard_lam if mode == 1 else p.guard_obj
        rest = order[k:]
        band = rest[key[rest] >= key[win[-1]] - delta][:p.band_cap] if (win.size and rest.size) else rest[:0]
        n_band = int((key[rest] >= key[win[-1]] - delta).sum()) if (win.size and rest.size) else 0
        self._band = dict(idx=idx[band], score=key[band], lam=lam[band], obj=obj[band], n_band=n_band,
                          band_open=int(n_band > band.size),
                          n_unc_lam=int((np.abs(lam - THR) <= p.guard_lam).sum()) if mode != 2 else 0,
                          n_unc_obj=int((np.abs(obj) <= p.guard_obj).sum()) if mode != 1 else 0)
        return idx[win], key[win], lam[win], obj[win]

    def last_band(self, cap=None):
        b = dict(self._band)
        if cap is not None and b["idx"].size > cap:
            for key in ("idx", "score", "lam", "obj"):
                b[key] = b[key][:cap]
            b["band_open"] = 1
        return b

    def merge_topk(self, score, obj2, idx, k):
        obj2 = np.zeros_like(score) if obj2 is None else obj2
        return np.lexsort((idx, -obj2, -score))[:k]


def device_like(lam, obj, seed=0):
    """What the GPU returns: the same scores up to a few ulps, identical for identical inputs."""
    rng = np.random.default_rng(seed)
    return lam + rng.integers(-3, 4, lam.size) * 1.1e-16, obj + rng.integers(-3, 4, obj.size) * 2.0e-12


def run(strat, k, lam_o, obj_o, rescorer, **kw):
    lam_d, obj_d = device_like(lam_o, obj_o)
    eng = GuardedFakeEngine(lam_d, obj_d, **kw)
    sel = ShardedSelector(eng, local=True)
    k_try = k
    for _ in range(4):
        raw = sel.select(strat, None, k_try)
        res = neartie.resolve(raw, k, rescorer, eng.params.guard_lam, eng.params.guard_obj)
        if not res.get("short") or raw["idx"].size < k_try:
            break
        k_try += res["short"] + 64
    return res


@pytest.mark.parametrize("d", [2, 3, 4, 5])
def test_exact_nn_is_bit_identical_to_NNs_so(golden, blobs, d):
    assert np.array_equal(neartie.nn_exact(blobs[d], golden["nn%d_in" % d]), golden["nn%d_out" % d])


def test_rescorer_is_bit_identical_to_the_oracle(golden, blobs):
    n, Q_arr, adj = inst_arrays(golden, "spar040-030-1")
    idx, sizes = orc.cover_pattern_E(adj, 5)                       # mixed subset sizes 2..5
    vv = golden["mix_vars"]
    lam_o, obj_o = orc.score_cover(Q_arr, n, idx, sizes, vv, blobs)
    rs = neartie.Rescorer(n, Q_arr, vv, blobs, lambda i: idx[i])
    pick = np.random.default_rng(1).permutation(idx.shape[0])[:500]
    assert np.array_equal(rs.lam(pick), lam_o[pick]) and np.array_equal(rs.obj(pick), obj_o[pick])


def test_degenerate_vertex_order_is_the_references(golden, blobs):
    """cfg1 at x = 0.5, X in {0, 0.5}: 4060 violated triples in 12 tie classes of up to 2661 members whose order in the
    reference is LAPACK round-off (golden cfg1_deg_s1_sets).  Device-like scores + resolution reproduce it."""
    n, Q_arr, adj = inst_arrays(golden, "spar030-060-1")
    idx = orc.cover_all(n, 3)
    vd = orc.degenerate_point(n, Q_arr)
    lam_o, obj_o = orc.score_cover(Q_arr, n, idx, np.full(4060, 3), vd, blobs)
    rs = neartie.Rescorer(n, Q_arr, vd, blobs, lambda i: idx[i])
    want = [tuple(r) for r in golden["cfg1_deg_s1_sets"]]
    for k in (4060, 5000, 406, 100):
        res = run(1, k, lam_o, obj_o, rs)
        assert [tuple(r) for r in idx[res["idx"]]] == want[:k]
        assert np.abs(res["score"] - golden["cfg1_deg_s1_score"][:k]).max() < 1e-14   # singletons keep the device score
        assert res["degenerate"] == 1 and res["n_near_ties"] > 0
    # the device's own order is NOT the reference's here (that is why the guard exists)
    lam_d, _ = device_like(lam_o, obj_o)
    own = np.lexsort((np.arange(4060), lam_d))
    assert [tuple(r) for r in idx[own]] != want
    # a band that does not fit is reported, not silently cut
    res = run(1, 100, lam_o, obj_o, rs, band_cap=50)
    assert res["degenerate"] == 2


def test_fig8_lp_vertex(golden, blobs):
    """The reference's own round-1 LP vertex of spar020-100-1 (data_figures/fig8_data.csv): NN ranking (strat 2) has no near
    ties; 64 PSD-singular triples sit inside the guard of the violation threshold and are classified by LAPACK."""
    n, Q_arr, adj = inst_arrays(golden, "spar020-100-1")
    idx, sizes = orc.cover_pattern_E(adj, 3)
    vv = golden["fig8_vars"]
    lam_o, obj_o = orc.score_cover(Q_arr, n, idx, sizes, vv, blobs)
    rs = neartie.Rescorer(n, Q_arr, vv, blobs, lambda i: idx[i])
    N = idx.shape[0]
    res = run(2, N, lam_o, obj_o, rs)
    assert np.array_equal(res["idx"], golden["fig8_r1_cut_idx"]) and res["degenerate"] == 0
    order, score = orc.select_feas(lam_o)
    res = run(1, N, lam_o, obj_o, rs)
    assert np.array_equal(res["idx"], order) and np.abs(res["score"] - score).max() < 1e-14
    assert int(res["counts"][1]) == order.size and res["degenerate"] == 1
    for k in (105, 400):
        ns, order, score = orc.select_comb_walk(obj_o, lam_o, k)
        res = run(4, k, lam_o, obj_o, rs)
        if res["degenerate"] != 2:
            assert np.array_equal(res["idx"], order[:k]) and res["new_strat"] == ns
            assert np.abs(res["score"] - score[:k]).max() < 1e-9
        assert k != 105 or res["degenerate"] != 2


def test_non_degenerate_point_needs_no_rescoring(golden, blobs):
    n, Q_arr, adj = inst_arrays(golden, "spar030-060-1")
    idx = orc.cover_all(n, 3)
    vv = golden["cfg1_vars"]
    lam_o, obj_o = orc.score_cover(Q_arr, n, idx, np.full(4060, 3), vv, blobs)

    class Boom(object):
        thr_eig, thr_opt, big_m = -1e-15, 0.0, 1000.0

        def lam(self, i):
            raise AssertionError("no re-scoring expected")
        obj = lam

    for strat, k in ((1, 406), (2, 406), (4, 406)):
        res = run(strat, k, lam_o, obj_o, Boom())
        assert res["degenerate"] == 0 and res["n_near_ties"] == 0 and res["idx"].size == k
    assert np.array_equal(run(4, 406, lam_o, obj_o, Boom())["idx"], golden["cfg1_s4_idx"][:406])


def test_tie_runs():
    s = np.array([5.0, 4.0, 4.0 - 1e-13, 3.0, 2.0, 2.0, 2.0 - 5e-13, 1.0])
    assert neartie.tie_runs(s, 1e-12).tolist() == [False, True, True, False, True, True, True, False]
    assert neartie.tie_runs(s[:1], 1e-12).tolist() == [False] and neartie.tie_runs(s[:0], 1e-12).size == 0
