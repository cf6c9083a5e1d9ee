"""GPU tests of the drop-in Python surface (same method names / tuple formats as the reference's CutSolver),
checked against fixtures produced by the unmodified reference. They read like the calls the reference's own
cut_select_algo makes (cut_select_qp.py:114, 140-141, 169-185; cut_select_qcqp.py:50-98)."""
import numpy as np
import pytest

import sdpcutsel_via_nn_b200 as pkg
from conftest import inst_arrays
from oracle import cutsel_oracle as orc

pytestmark = pytest.mark.gpu


def rows_of(cs):
    rows = cs._my_prob.linear_constraints.rows
    return rows


def check_rows(rows, g_ind, g_val, g_rhs, tol=1e-9):
    assert len(rows) == g_rhs.shape[0]
    for (sp, rhs, sense), ind, val, r in zip(rows, g_ind, g_val, g_rhs):
        m = len(sp.ind)
        assert sense == "G" and sp.ind == [int(v) for v in ind[:m]]
        assert np.abs(np.array(sp.val) - val[:m]).max() < tol and abs(rhs - r) < tol


def test_cfg1_feasibility_round(golden):
    n, Q_arr, adj = inst_arrays(golden, "spar030-060-1")
    cs = pkg.CutSolver()
    cs.set_instance(Q_arr, adj, n, dim=3)
    cs._load_neural_nets()
    assert cs._get_sdp_vertex_cover(3, ch_ext=-1) == 4060
    vv = golden["cfg1_vars"]
    k = 406
    rl = cs._sel_eigcut_by_ordering_on_measure(1, vv, 1, sel_size=k)
    assert len(rl) == golden["cfg1_s1_score"].shape[0] and rl.n_violated == len(rl)
    for e, s, sc in zip(rl, golden["cfg1_s1_sets"], golden["cfg1_s1_score"]):
        assert e[0] == [int(v) for v in s] and abs(e[1] - sc) < 1e-12 and e[3] == 3
        assert e[2] == orc.xarr_inds(n, e[0])
    nb = cs._gen_eigcuts_selected(1, k, rl, vars_values=vv)
    assert nb == k
    check_rows(rows_of(cs), golden["cfg1_s1_cut_ind"], golden["cfg1_s1_cut_val"], golden["cfg1_s1_cut_rhs"])


@pytest.mark.parametrize("strat", [2, 4])
def test_cfg1_optimality_and_combined_round(golden, strat):
    n, Q_arr, adj = inst_arrays(golden, "spar030-060-1")
    cs = pkg.CutSolver()
    cs.set_instance(Q_arr, adj, n, dim=3)
    cs._load_neural_nets()
    cs._get_sdp_vertex_cover(3, ch_ext=-1)
    vv = golden["cfg1_vars"]
    k = 406
    out = cs._sel_eigcut_by_ordering_on_measure(strat, vv, 1, sel_size=k)
    if strat == 4:
        new_strat, rl = out
        assert new_strat == int(golden["cfg1_s4_newstrat"])
    else:
        rl = out
    X_vals, x_vals = orc.split_vars(vv, n)
    for e, i, sc in zip(rl, golden["cfg1_s%d_idx" % strat], golden["cfg1_s%d_score" % strat]):
        assert isinstance(e[0], int) and e[0] == int(i) and abs(e[1] - sc) < 1e-9
        s = cs._agg_list[e[0]][0]
        assert e[2] == tuple(x_vals[s]) and e[3] == tuple(X_vals[orc.xarr_inds(n, s)])
    nb = cs._gen_eigcuts_selected(strat, k, rl)                 # opt strategies pass no vars_values (cut_select_qp.py:181)
    check_rows(rows_of(cs), golden["cfg1_s%d_cut_ind" % strat], golden["cfg1_s%d_cut_val" % strat], golden["cfg1_s%d_cut_rhs" % strat])
    assert nb == golden["cfg1_s%d_cut_rhs" % strat].shape[0]
    # the reference's k = 0 quirk: ZeroDivisionError swallowed, bare list returned (cut_select_qp.py:628-632)
    if strat == 4:
        assert isinstance(cs._sel_eigcut_by_ordering_on_measure(4, vv, 1, sel_size=0), list)


@pytest.mark.parametrize("strat", [1, 2, 4])
def test_cut_rows_do_not_depend_on_the_rank_list_shortcut(golden, strat):
    """_gen_eigcuts_selected takes the index tuples from the arrays behind a RankList it produced itself; a list the caller
    copied, sliced or re-ordered goes through the entries (cut_select_qp.py:713-735) -- both give the same rows."""
    n, Q_arr, adj = inst_arrays(golden, "spar030-060-1")
    cs = pkg.CutSolver()
    cs.set_instance(Q_arr, adj, n, dim=3)
    cs._load_neural_nets()
    cs._get_sdp_vertex_cover(3, ch_ext=-1)
    vv = golden["cfg1_vars"]
    k = 406
    out = cs._sel_eigcut_by_ordering_on_measure(strat, vv, 1, sel_size=k)
    rl = out[1] if strat == 4 else out
    assert rl._sealed_rows() is not None

    def rows_for(lst, kk, **kw):
        cs._my_prob.linear_constraints.rows = []
        nb = cs._gen_eigcuts_selected(strat, kk, lst, vars_values=vv, **kw)
        return nb, [(sp.ind, sp.val, rhs) for sp, rhs, _ in rows_of(cs)]

    for kk in (k, 17, 0):
        assert rows_for(rl, kk) == rows_for(list(rl), kk)
    if strat != 1:
        assert rows_for(rl, k, strong_only=True) == rows_for(list(rl), k, strong_only=True)
    # a re-ordered list is no longer the sealed one: the entries decide
    rl[1], rl[2] = rl[2], rl[1]
    assert rl._sealed_rows() is None
    assert rows_for(rl, 5) == rows_for(list(rl), 5)
    del rl[3:]
    assert rl._sealed_rows() is None
    assert rows_for(rl, 5)[0] <= 3


def test_cfg2_rounds_and_triangles(golden):
    n, Q_arr, adj = inst_arrays(golden, "spar125-075-1")
    cs = pkg.CutSolver()
    cs.set_instance(Q_arr, adj, n, dim=3)
    cs._load_neural_nets()
    N = cs._get_sdp_vertex_cover(3)
    assert N == 133242
    vv = golden["cfg2_vars"]
    k = min(int(np.floor(0.1 * N)), cs._SDP_CUTS_PER_ROUND_MAX)
    for strat in (1, 2, 4):
        cs._my_prob.linear_constraints.rows = []
        out = cs._sel_eigcut_by_ordering_on_measure(strat, vv, 1, sel_size=k)
        rl = out[1] if strat == 4 else out
        if strat == 4:
            assert out[0] == int(golden["cfg2_s4_newstrat"])
        if strat == 1:
            assert rl.n_violated == int(golden["cfg2_s1_len"])
            assert [e[0] for e in rl[:k]] == [[int(v) for v in r if v >= 0] for r in golden["cfg2_s1_sets"]]
        else:
            assert [e[0] for e in rl[:k]] == [int(v) for v in golden["cfg2_s%d_idx" % strat]]
        assert np.abs(np.array([e[1] for e in rl[:k]]) - golden["cfg2_s%d_score" % strat]).max() < 1e-9
        nb = cs._gen_eigcuts_selected(strat, k, rl, vars_values=vv)
        assert nb == int(golden["cfg2_s%d_nbcuts" % strat])
        check_rows(rows_of(cs)[:64], golden["cfg2_s%d_cut_ind" % strat], golden["cfg2_s%d_cut_val" % strat], golden["cfg2_s%d_cut_rhs" % strat])
    # triangle inequalities through the reference's (name-mangled) private methods
    cs._my_prob.linear_constraints.rows = []
    cs._CutSolver__preprocess_triangle_ineq()
    nb = cs._CutSolver__separate_and_add_triangle(0.1, vv)
    assert nb == 10000
    rows = rows_of(cs)
    for (sp, rhs, sense), ind, val, r in zip(rows, golden["cfg2_tri_ind"], golden["cfg2_tri_val"], golden["cfg2_tri_rhs"]):
        m = len(sp.ind)
        assert sp.ind == [int(v) for v in ind[:m]] and sp.val == [int(v) for v in val[:m]] and rhs == int(r) and sense == "G"


@pytest.mark.parametrize("dim", [3, 4, 5])
def test_qcqp_round(golden, dim):
    n = 20
    cs = pkg.CutSolverQCQP()
    cs.set_instance(golden["qcqp_Q_arr"], golden["qcqp_adj"], n, dim=dim, Q_adj_cons=golden["qcqp_adj_cons"])
    cs._load_neural_nets()
    agg_list_cons = cs._CutSolverQCQP__get_vertex_cover(dim)
    agg_list = cs._agg_list[:]
    N = len(agg_list)
    assert [N, len(agg_list_cons)] == list(golden["qcqp_d%d_N" % dim])
    k = max(1, min(int(np.floor(0.1 * N)), 5000))
    vv = golden["qcqp_vars"]
    for strat in (1, 4):
        cs._my_prob.linear_constraints.rows = []
        new_strat, nb, nb_opt, rl = cs.select_and_cut_round(strat, vv, k, agg_list, agg_list_cons)
        if strat == 1:
            assert [e[0] for e in rl] == [[int(v) for v in r] for r in golden["qcqp_d%d_s1_sets" % dim]]
        else:
            assert new_strat == int(golden["qcqp_d%d_s4_newstrat" % dim])
            assert [e[0] for e in rl] == [int(v) for v in golden["qcqp_d%d_s4_idx" % dim]]
        assert np.abs(np.array([e[1] for e in rl]) - golden["qcqp_d%d_s%d_score" % (dim, strat)]).max() < 1e-9
        check_rows(rows_of(cs), golden["qcqp_d%d_s%d_cut_ind" % (dim, strat)], golden["qcqp_d%d_s%d_cut_val" % (dim, strat)],
                   golden["qcqp_d%d_s%d_cut_rhs" % (dim, strat)])


def test_mixed_cover_through_solver(golden):
    n, Q_arr, adj = inst_arrays(golden, "spar040-030-1")
    vv = golden["mix_vars"]
    for dim in (4, 5):
        cs = pkg.CutSolver()
        cs.set_instance(Q_arr, adj, n, dim=dim)
        cs._load_neural_nets()
        N = cs._get_sdp_vertex_cover(dim)
        k = max(1, min(int(np.floor(0.1 * N)), 5000))
        new_strat, rl = cs._sel_eigcut_by_ordering_on_measure(4, vv, 1, sel_size=k)
        assert new_strat == int(golden["mix_d%d_s4_newstrat" % dim])
        assert [e[0] for e in rl] == [int(v) for v in golden["mix_d%d_s4_idx" % dim][:k]]
        nb = cs._gen_eigcuts_selected(4, k, rl)
        check_rows(rows_of(cs), golden["mix_d%d_s4_cut_ind" % dim], golden["mix_d%d_s4_cut_val" % dim], golden["mix_d%d_s4_cut_rhs" % dim])


def test_eigendecomp_and_nn_callables(golden, blobs):
    n, Q_arr, adj = inst_arrays(golden, "spar030-060-1")
    cs = pkg.CutSolver()
    cs.set_instance(Q_arr, adj, n, dim=5)
    cs._load_neural_nets()
    cs._get_sdp_vertex_cover(5)
    rng = np.random.default_rng(3)
    for d in (2, 3, 4, 5):
        pt = rng.random(d)
        Xs = rng.random(d * (d + 1) // 2)
        M = orc.eig_matrix(pt[None, :], Xs[None, :])[0]
        w = cs._get_eigendecomp(d, pt, Xs, False)
        assert np.abs(w - np.linalg.eigvalsh(M, "U")).max() < 1e-13
        w2, V = cs._get_eigendecomp(d, pt, Xs, True)
        Mf = np.triu(M) + np.triu(M, 1).T
        assert np.abs(Mf @ V - V * w2[None, :]).max() < 1e-12 and np.abs(V.T @ V - np.eye(d + 1)).max() < 1e-12
        func, input_arr = cs._nns[d - 2]                      # call pattern of cut_select_qp.py:579-582
        x = golden["nn%d_in" % d][20]
        input_arr[:] = x
        assert abs(func(input_arr) - golden["nn%d_out" % d][20]) < 1e-11


class _CsrSink(object):
    """A cut sink that takes rows in one shot (the shape of a CPXaddrows wrapper)."""

    def __init__(self):
        self.calls = []

    def add_rows_csr(self, rowptr, ind, val, rhs, senses):
        self.calls.append((np.array(rowptr), np.array(ind), np.array(val), np.array(rhs), senses))

    def add(self, **kw):  # pragma: no cover - must not be used when add_rows_csr exists
        raise AssertionError("per-row path used although the sink takes CSR")


class _CsrProb(object):
    def __init__(self):
        self.linear_constraints = _CsrSink()


def test_one_shot_csr_row_emission(golden):
    """Same round as test_cfg2_rounds_and_triangles, but the cuts leave as CSR arrays: identical rows."""
    n, Q_arr, adj = inst_arrays(golden, "spar125-075-1")
    vv = golden["cfg2_vars"]
    ref, csr = pkg.CutSolver(), pkg.CutSolver()
    ref.set_instance(Q_arr, adj, n, dim=3)
    csr.set_instance(Q_arr, adj, n, dim=3, my_prob=_CsrProb())
    for cs in (ref, csr):
        cs._load_neural_nets()
        N = cs._get_sdp_vertex_cover(3)
    k = min(int(np.floor(0.1 * N)), ref._SDP_CUTS_PER_ROUND_MAX)
    for strat in (1, 2):
        ref._my_prob.linear_constraints.rows = []
        csr._my_prob.linear_constraints.calls = []
        nb = [cs._gen_eigcuts_selected(strat, k, cs._sel_eigcut_by_ordering_on_measure(strat, vv, 1, sel_size=k), vars_values=vv)
              for cs in (ref, csr)]
        assert nb[0] == nb[1] == int(golden["cfg2_s%d_nbcuts" % strat])
        (rowptr, ind, val, rhs, senses), = csr._my_prob.linear_constraints.calls
        rows = rows_of(ref)
        assert len(rows) == nb[0] == rhs.size == len(senses) and set(senses) == {"G"} and rowptr[0] == 0
        for r, (sp, rh, _) in enumerate(rows):
            a, b = rowptr[r], rowptr[r + 1]
            assert ind[a:b].tolist() == sp.ind and val[a:b].tolist() == sp.val and rhs[r] == rh
    ref._my_prob.linear_constraints.rows = []
    csr._my_prob.linear_constraints.calls = []
    for cs in (ref, csr):
        cs._CutSolver__preprocess_triangle_ineq()
        assert cs._CutSolver__separate_and_add_triangle(0.1, vv) == 10000
    (rowptr, ind, val, rhs, senses), = csr._my_prob.linear_constraints.calls
    for r, (sp, rh, _) in enumerate(rows_of(ref)):
        a, b = rowptr[r], rowptr[r + 1]
        assert ind[a:b].tolist() == sp.ind and val[a:b].tolist() == sp.val and rhs[r] == rh
