"""CPU tests: the oracle (oracle/cutsel_oracle.py + nn_oracle.c) against fixtures produced by the
unmodified reference (tests/golden/make_golden.py) and the reference's own Fig-8 data."""
import numpy as np
import pytest

from oracle import cutsel_oracle as orc
from conftest import inst_arrays


@pytest.mark.parametrize("d", [2, 3, 4, 5])
def test_nn_oracle_bit_equal_to_NNs_so(golden, blobs, d):
    y = orc.nn_eval(blobs[d], golden["nn%d_in" % d])
    assert np.array_equal(y, golden["nn%d_out" % d])          # bit-exact


@pytest.mark.parametrize("name", ["spar030-060-1", "spar040-030-1", "spar050-030-1"])
@pytest.mark.parametrize("dim", [3, 4, 5])
def test_cover_pattern_E_set_and_order(golden, name, dim):
    n, Q_arr, adj = inst_arrays(golden, name)
    want = golden["cover_%s_d%d" % (name.replace("-", "_"), dim)]
    idx, sizes = orc.cover_pattern_E(adj, dim)
    assert np.array_equal(idx, want)
    if n <= 40:
        loops = orc.cover_pattern_E_loops(adj, dim)
        assert loops == [tuple(int(v) for v in r[:s]) for r, s in zip(idx, sizes)]


@pytest.mark.parametrize("dim", [3, 4, 5])
def test_aggregation(golden, dim):
    n, Q_arr, adj = inst_arrays(golden, "spar030-060-1")
    idx, sizes = orc.cover_pattern_E(adj, dim)
    Qs_want = golden["agg_spar030_060_1_d%d_Qslice" % dim]
    me_want = golden["agg_spar030_060_1_d%d_maxelem" % dim]
    for d in np.unique(sizes):
        sel = sizes == d
        _, Qs, me = orc.aggregate(Q_arr, n, idx[sel, :d])
        t = d * (d + 1) // 2
        assert np.array_equal(Qs, Qs_want[sel, :t])
        assert np.array_equal(me, me_want[sel])


def _sets(idx):
    return [tuple(int(v) for v in r if v >= 0) for r in idx]


def test_cfg1_all_triples(golden, blobs):
    n, Q_arr, adj = inst_arrays(golden, "spar030-060-1")
    idx = orc.cover_all(n, 3)
    assert idx.shape[0] == 4060
    assert np.array_equal(orc.cover_all_window(n, 3, 1000, 1100), idx[1000:1100])
    sizes = np.full(idx.shape[0], 3)
    vv = golden["cfg1_vars"]
    assert np.array_equal(vv, orc.synth_point(n, seed=8))
    lam, obj = orc.score_cover(Q_arr, n, idx, sizes, vv, blobs)
    order, score = orc.select_feas(lam)
    assert np.array_equal(score, golden["cfg1_s1_score"])
    assert _sets(idx[order]) == _sets(golden["cfg1_s1_sets"])
    order, score = orc.select_opt(obj)
    assert np.array_equal(order, golden["cfg1_s2_idx"])
    assert np.array_equal(score, golden["cfg1_s2_score"])
    k = 406
    for fn in (orc.select_comb_walk, orc.select_comb):
        ns, order, score = fn(obj, lam, k)
        assert ns == int(golden["cfg1_s4_newstrat"])
        assert np.array_equal(order, golden["cfg1_s4_idx"])
        assert np.array_equal(score, golden["cfg1_s4_score"])
    # eigcuts of the feasibility selection
    order, _ = orc.select_feas(lam)
    cuts = [orc.gen_eigcut(n, idx[i], vv) for i in order[:k]]
    cuts = [c for c in cuts if c is not None]
    assert len(cuts) == golden["cfg1_s1_cut_rhs"].shape[0]
    for c, ind, val, rhs in zip(cuts, golden["cfg1_s1_cut_ind"], golden["cfg1_s1_cut_val"], golden["cfg1_s1_cut_rhs"]):
        assert c[0] == list(ind) and np.array_equal(c[1], val) and c[2] == rhs
    # degenerate vertex: scores reproduce bit-for-bit through the same LAPACK path
    lamd, _ = orc.score_cover(Q_arr, n, idx, sizes, orc.degenerate_point(n, Q_arr), want_obj=False)
    order, score = orc.select_feas(lamd)
    assert np.array_equal(score, golden["cfg1_deg_s1_score"])
    assert _sets(idx[order]) == _sets(golden["cfg1_deg_s1_sets"])


@pytest.mark.parametrize("dim", [3, 4, 5])
def test_mixed_size_cover_selection(golden, blobs, dim):
    n, Q_arr, adj = inst_arrays(golden, "spar040-030-1")
    idx, sizes = orc.cover_pattern_E(adj, dim)
    vv = golden["mix_vars"]
    lam, obj = orc.score_cover(Q_arr, n, idx, sizes, vv, blobs)
    N = idx.shape[0]
    k = max(1, min(int(np.floor(0.1 * N)), 5000))
    order, score = orc.select_feas(lam)
    assert np.array_equal(score, golden["mix_d%d_s1_score" % dim])
    assert _sets(idx[order]) == _sets(golden["mix_d%d_s1_sets" % dim])
    order, score = orc.select_opt(obj)
    assert np.array_equal(order, golden["mix_d%d_s2_idx" % dim])
    assert np.array_equal(score, golden["mix_d%d_s2_score" % dim])
    for fn in (orc.select_comb_walk, orc.select_comb):
        ns, order, score = fn(obj, lam, k)
        assert ns == int(golden["mix_d%d_s4_newstrat" % dim])
        assert np.array_equal(order, golden["mix_d%d_s4_idx" % dim])
        assert np.array_equal(score, golden["mix_d%d_s4_score" % dim])
    # cuts of the optimality selection (cut_select_qp.py:720-751)
    order, _ = orc.select_opt(obj)
    cuts = [orc.gen_eigcut(n, idx[i, :sizes[i]], vv) for i in order[:k]]
    cuts = [c for c in cuts if c is not None]
    want_rhs = golden["mix_d%d_s2_cut_rhs" % dim]
    assert len(cuts) == want_rhs.shape[0]
    for c, ind, val, rhs in zip(cuts, golden["mix_d%d_s2_cut_ind" % dim], golden["mix_d%d_s2_cut_val" % dim], want_rhs):
        m = len(c[0])
        assert c[0] == list(ind[:m]) and np.array_equal(c[1], val[:m]) and c[2] == rhs


def test_cfg2_spar125_pattern_E_and_triangles(golden, blobs):
    n, Q_arr, adj = inst_arrays(golden, "spar125-075-1")
    idx, sizes = orc.cover_pattern_E(adj, 3)
    assert idx.shape[0] == 133242                     # data_tables nb_subproblems column
    vv = golden["cfg2_vars"]
    lam, obj = orc.score_cover(Q_arr, n, idx, sizes, vv, blobs)
    k = 5000
    order, score = orc.select_feas(lam)
    assert order.shape[0] == int(golden["cfg2_s1_len"])
    assert np.array_equal(score[:k], golden["cfg2_s1_score"])
    assert _sets(idx[order[:k]]) == _sets(golden["cfg2_s1_sets"])
    order, score = orc.select_opt(obj)
    assert np.array_equal(order[:k], golden["cfg2_s2_idx"]) and np.array_equal(score[:k], golden["cfg2_s2_score"])
    ns, order, score = orc.select_comb(obj, lam, k)
    assert ns == int(golden["cfg2_s4_newstrat"])
    assert np.array_equal(order[:k], golden["cfg2_s4_idx"]) and np.array_equal(score[:k], golden["cfg2_s4_score"])
    # triangles
    tri, dens = orc.triangles_pre(adj, n)
    assert tri.shape[0] == int(golden["cfg2_tri_ntriples"])
    pos, typ, viol = orc.triangles_sep(n, tri, dens, vv, 0.1)
    assert pos.shape[0] == golden["cfg2_tri_rhs"].shape[0] == 10000
    for i in list(range(0, 10000, 97)) + [9999]:
        ind, val, rhs = orc.triangle_row(n, tri[pos[i]], int(typ[i]))
        m = len(ind)
        assert ind == list(golden["cfg2_tri_ind"][i][:m]) and val == list(golden["cfg2_tri_val"][i][:m])
        assert rhs == golden["cfg2_tri_rhs"][i]


def test_fig8_golden_rows(golden, blobs):
    """data_figures/fig8_data.csv round-1 rows: 1,051 NN_3D optimality measures in sorted order."""
    n, Q_arr, adj = inst_arrays(golden, "spar020-100-1")
    idx, sizes = orc.cover_pattern_E(adj, 3)
    assert idx.shape[0] == 1051
    _, obj = orc.score_cover(Q_arr, n, idx, sizes, golden["fig8_vars"], blobs, want_lam=False)
    order, score = orc.select_opt(obj)
    assert np.array_equal(order, golden["fig8_r1_cut_idx"])
    assert np.abs(score - golden["fig8_r1_estim"]).max() < 2e-13


@pytest.mark.parametrize("dim", [3, 4, 5])
def test_qcqp_caller_pattern(golden, blobs, dim):
    """cut_select_qcqp.py:63-98 on q_20_20_100_1: objective cover = all C(20,dim), constraint-only cover empty."""
    n = 20
    Q_arr = golden["qcqp_Q_arr"]
    idx, sizes = orc.cover_pattern_E(golden["qcqp_adj"], dim)
    idx_c, _ = orc.cover_pattern_E(golden["qcqp_adj_cons"], dim)
    N = idx.shape[0]
    assert [N, 0] == list(golden["qcqp_d%d_N" % dim]) and np.array_equal(idx, idx_c)
    assert np.array_equal(idx, orc.cover_all(n, dim))
    vv = golden["qcqp_vars"]
    lam, obj = orc.score_cover(Q_arr, n, idx, sizes, vv, blobs)
    k = max(1, min(int(np.floor(0.1 * N)), 5000))
    order, score = orc.select_feas(lam)
    assert np.array_equal(score[:k], golden["qcqp_d%d_s1_score" % dim])
    assert _sets(idx[order[:k]]) == _sets(golden["qcqp_d%d_s1_sets" % dim])
    ns, order, score = orc.select_comb(obj, lam, k)
    assert ns == int(golden["qcqp_d%d_s4_newstrat" % dim])
    assert np.array_equal(order[:k], golden["qcqp_d%d_s4_idx" % dim])
    assert np.array_equal(score[:k], golden["qcqp_d%d_s4_score" % dim])


def test_exact_sdp_oracle_reproduces_moseks_figure_8_values(golden):
    """oracle.sdp_exact_value / exact_measure (strat 3, cut_select_qp.py:555-598) against the reference's committed Mosek
    results of figure 8, round 1 (data_figures/fig8_data.csv -> tests/golden/fig8_exact.npz): 1,051 exact measures within
    Mosek's own tolerance, the exact selection (100 sub-problems) identical, the published round statistics reproduced."""
    import os
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fig8_exact.npz")) as z:
        f = {k: z[k] for k in z.files}
    n, Q_arr, adj = inst_arrays(golden, "spar020-100-1")
    idx, sizes = orc.cover_pattern_E(adj, 3)
    ex = orc.exact_measure(Q_arr, n, idx, sizes, golden["fig8_vars"])
    assert np.abs(ex[f["r1_cut_idx"]] - f["r1_exact"]).max() < 1e-5
    order = np.argsort(-ex, kind="stable")
    sel = np.zeros(idx.shape[0], dtype=int)
    sel[order[:100]] = 1
    assert np.array_equal(sel[f["r1_cut_idx"]], f["r1_sel_exact"])
    assert (sel[f["r1_cut_idx"]] & f["r1_sel_estim"]).sum() / 100 == f["summary"][0, 2]
    assert abs(np.std(ex[order[:100]]) - f["summary"][0, 3]) < 1e-5
