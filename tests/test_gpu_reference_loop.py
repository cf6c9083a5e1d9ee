"""The reference's OWN ``cut_select_algo`` (cut_select_qp.py:73-221, cut_select_qcqp.py:16-113) driving the GPU path.

``make_solvers`` puts the B200 selection in front of the unmodified reference classes (baseline/_ref, installed by
tools/install_reference.py); CPLEX is replaced by a HiGHS-backed stand-in (tests/refloader.py).  Every selection and
every cut round the loop performs on the GPU is replayed with the unmodified reference methods at the same LP point:
same ranked prefix (index for index), same rows.  The LP trajectories of the two complete runs are compared as well.
"""
import numpy as np
import pytest

import refloader

pytestmark = pytest.mark.gpu

REF = refloader.load_reference()
needs_ref = pytest.mark.skipif(REF is None, reason="no reference tree (baseline/_ref) on this host")


class Recorder(object):
    """Mixed into the GPU solver: logs every selection / cut generation the reference's loop asks for."""

    def _sel_eigcut_by_ordering_on_measure(self, strat, vars_values, cut_round, sel_size=0):
        out = super(Recorder, self)._sel_eigcut_by_ordering_on_measure(strat, vars_values, cut_round, sel_size=sel_size)
        self.__dict__.setdefault("log", []).append(dict(kind="sel", strat=strat, vv=np.array(vars_values), cut_round=cut_round,
                                                        sel_size=sel_size, out=out, agg=self._agg_list))
        return out

    def _gen_eigcuts_selected(self, strat, sel_size, rank_list, strong_only=False, vars_values=None):
        rows = self._my_prob.linear_constraints.rows
        n0 = len(rows)
        nb = super(Recorder, self)._gen_eigcuts_selected(strat, sel_size, rank_list, strong_only=strong_only, vars_values=vars_values)
        self.__dict__.setdefault("log", []).append(dict(kind="gen", strat=strat, sel_size=sel_size, nb=nb, rows=rows[n0:],
                                                        rank_list=rank_list, vv=None if vars_values is None else np.array(vars_values)))
        return nb


class Sink(object):
    class _LC(object):
        def __init__(self):
            self.rows = []

        def add(self, lin_expr=None, rhs=None, senses=None, **kw):
            self.rows.extend(zip(lin_expr, rhs, senses))

    def __init__(self):
        self.linear_constraints = Sink._LC()


def prefix_equal(strat, got, want, m, tol):
    got, want = list(got[:m]), list(want[:m])
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert list(g[0]) == list(w[0]) if strat == 1 else g[0] == w[0]
        assert abs(g[1] - w[1]) < tol


def rows_equal(got, want, vv, n):
    """Same rows.  Coefficients are compared element-wise (1e-9); where the smallest eigenvalue is repeated the mirror takes
    the row from numpy eigh like the reference, so they agree there too."""
    assert len(got) == len(want)
    for (sg, rg, eg), (sw, rw, ew) in zip(got, want):
        assert sg.ind == sw.ind and eg == ew == "G"
        assert np.abs(np.array(sg.val) - np.array(sw.val)).max() < 1e-9 and abs(rg - rw) < 1e-9


def run_pair(filename, dim, strat, rounds, lapack_rows=False, **kw):
    ref, refq, d = REF
    import sdpcutsel_via_nn_b200 as pkg
    GpuSolver, _ = pkg.make_solvers(ref, refq)

    class Logged(Recorder, GpuSolver):
        _CUT_ROWS_FROM_LAPACK = lapack_rows

    with refloader.in_reference_dir(d):
        a = ref.CutSolver()
        out_ref = a.cut_select_algo(filename, dim, 0.1, strat=strat, nb_rounds_cuts=rounds, **kw)
        b = Logged()
        out_gpu = b.cut_select_algo(filename, dim, 0.1, strat=strat, nb_rounds_cuts=rounds, **kw)
        # replay every GPU selection with the unmodified reference at the same LP point
        r = ref.CutSolver()
        r._dim = dim
        r._CutSolver__parse_boxqp_into_cplex(filename)
        r._load_neural_nets()
        r._get_sdp_vertex_cover(dim, ch_ext=kw.get("ch_ext", 0) or 0)
        agg_ref = r._agg_list
        n = r._nb_vars
        sel_events = [ev for ev in getattr(b, "log", []) if ev["kind"] == "sel"]
        gen_events = [ev for ev in getattr(b, "log", []) if ev["kind"] == "gen"]
        for ev, gen in zip(sel_events, gen_events):
            r._agg_list = agg_ref
            want = r._sel_eigcut_by_ordering_on_measure(ev["strat"], ev["vv"], ev["cut_round"], sel_size=ev["sel_size"])
            got = ev["out"]
            if ev["strat"] == 4:
                assert got[0] == want[0], "new_strat of the combined rule"
                got, want = got[1], want[1]
            m = gen["sel_size"]
            prefix_equal(ev["strat"], got, want, m, 1e-12 if ev["strat"] == 1 else 1e-9)
            r._my_prob = Sink()
            nb = r._gen_eigcuts_selected(gen["strat"], m, want, vars_values=ev["vv"])
            assert nb == gen["nb"]
            rows_equal(gen["rows"], r._my_prob.linear_constraints.rows, ev["vv"], n)
    return out_ref, out_gpu, len(sel_events)


@needs_ref
@pytest.mark.parametrize("filename", ["spar020-100-1", "spar030-060-1"])
@pytest.mark.parametrize("strat", [1, 2, 4])
def test_boxqp_loop_four_rounds(filename, strat):
    """Default mode: cut rows from the device (numpy eigh only where the eigenvector is not unique).  Every selection and
    every row of every round equals the reference's at the same LP point (checked inside run_pair); the first cut round
    gives the same bound to 1e-9.  Later rounds may sit on another vertex of a degenerate LP (the rows differ from
    LAPACK's by ~1e-15 and HiGHS answers that with an alternative optimum), so the trajectories are compared loosely."""
    out_ref, out_gpu, n_sel = run_pair(filename, 3, strat, 4)
    assert n_sel == 4
    objs_ref, objs_gpu = np.array(out_ref[0]), np.array(out_gpu[0])
    assert out_gpu[6] == out_ref[6] and out_gpu[4][:2] == out_ref[4][:2] and out_gpu[5] == out_ref[5]
    assert np.abs(objs_gpu[:2] - objs_ref[:2]).max() < 1e-9 * abs(objs_ref[0])
    assert np.abs(objs_gpu - objs_ref).max() < 5e-2 * abs(objs_ref[0])
    assert abs(np.array(out_gpu[4]).sum() - np.array(out_ref[4]).sum()) <= 0.1 * np.array(out_ref[4]).sum()


@needs_ref
@pytest.mark.parametrize("filename", ["spar020-100-1", "spar030-060-1"])
@pytest.mark.parametrize("strat", [1, 2, 4])
def test_boxqp_loop_follows_the_reference_vertex_for_vertex(filename, strat):
    """_CUT_ROWS_FROM_LAPACK: the <= 5000 selected rows per round come from numpy eigh -> bit-identical rows, and because the
    GPU selection is index-identical, the whole run (bounds and cut counts of every round) equals the reference's."""
    out_ref, out_gpu, n_sel = run_pair(filename, 3, strat, 4, lapack_rows=True)
    assert n_sel == 4
    assert out_gpu[4] == out_ref[4] and out_gpu[5] == out_ref[5] and out_gpu[6] == out_ref[6]      # nbs_sdp_cuts, nbs_tri_cuts, N
    assert np.abs(np.array(out_gpu[0]) - np.array(out_ref[0])).max() < 1e-9 * abs(out_ref[0][0])


@needs_ref
@pytest.mark.parametrize("filename", ["spar020-100-1", "spar030-060-1"])
def test_boxqp_loop_random_selection(filename):
    """strat 5 (Table 3's random baseline): np.random.shuffle of the cover, cuts from its first sel_size entries.  Same
    seed -> same permutation (in place, round after round) -> with rows from numpy eigh the whole run equals the reference's."""
    ref, refq, d = REF
    import sdpcutsel_via_nn_b200 as pkg
    GpuSolver0, _ = pkg.make_solvers(ref, refq)

    class GpuSolver(GpuSolver0):
        _CUT_ROWS_FROM_LAPACK = True

    with refloader.in_reference_dir(d):
        np.random.seed(7)
        out_ref = ref.CutSolver().cut_select_algo(filename, 3, 0.1, strat=5, nb_rounds_cuts=3)
        np.random.seed(7)
        out_gpu = GpuSolver().cut_select_algo(filename, 3, 0.1, strat=5, nb_rounds_cuts=3)
    assert out_gpu[4] == out_ref[4] and out_gpu[6] == out_ref[6] and sum(out_ref[4]) > 0
    assert np.abs(np.array(out_gpu[0]) - np.array(out_ref[0])).max() < 1e-9 * abs(out_ref[0][0])


@needs_ref
@pytest.mark.parametrize("lapack_rows", [False, True])
def test_boxqp_loop_dense_cuts_and_triangles(lapack_rows):
    """strat 0 (__gen_dense_eigcuts) + triangle separation through the name-mangled private methods, and feasibility
    selection + triangles together (M + tri + S^E_3).  With rows from numpy eigh the whole run equals the reference's;
    with device rows the first round does (same LP point) and the later ones stay close (degenerate LP, see above)."""
    ref, refq, d = REF
    import sdpcutsel_via_nn_b200 as pkg
    GpuSolver0, _ = pkg.make_solvers(ref, refq)

    class GpuSolver(GpuSolver0):
        _CUT_ROWS_FROM_LAPACK = lapack_rows

    for strat in (0, 1):
        with refloader.in_reference_dir(d):
            out_ref = ref.CutSolver().cut_select_algo("spar030-060-1", 3, 0.1, strat=strat, nb_rounds_cuts=3, triangle_on=True)
            out_gpu = GpuSolver().cut_select_algo("spar030-060-1", 3, 0.1, strat=strat, nb_rounds_cuts=3, triangle_on=True)
        objs_ref, objs_gpu = np.array(out_ref[0]), np.array(out_gpu[0])
        assert out_gpu[4][:2] == out_ref[4][:2] and out_gpu[5][:1] == out_ref[5][:1]        # round 1: same cuts
        assert abs(objs_gpu[1] - objs_ref[1]) < 1e-9 * abs(objs_ref[0])
        if lapack_rows:
            assert out_gpu[4] == out_ref[4] and out_gpu[5] == out_ref[5]                    # SDP / triangle cuts of every round
            assert np.abs(objs_gpu - objs_ref).max() < 1e-9 * abs(objs_ref[0])
        else:
            assert np.abs(objs_gpu - objs_ref).max() < 5e-2 * abs(objs_ref[0])


@needs_ref
@pytest.mark.parametrize("dim", [3, 4])
@pytest.mark.parametrize("strat", [1, 4])
def test_qcqp_loop(dim, strat):
    """CutSolverQCQP.cut_select_algo on q_20_20_100_1 (BASELINE configs[4]): the two-cover caller pattern, the _agg_list
    swaps and the format sniffing run unmodified; the covers come from the key-based algebra of the mix-in."""
    ref, refq, d = REF
    import sdpcutsel_via_nn_b200 as pkg
    _, GpuQ0 = pkg.make_solvers(ref, refq)

    class GpuQ(GpuQ0):
        _CUT_ROWS_FROM_LAPACK = True

    with refloader.in_reference_dir(d):
        out_ref = refq.CutSolverQCQP().cut_select_algo("q_20_20_100_1", dim, 0.1, strat=strat, nb_rounds_cuts=3)
        out_gpu = GpuQ().cut_select_algo("q_20_20_100_1", dim, 0.1, strat=strat, nb_rounds_cuts=3)
    assert out_gpu[1] == out_ref[1] and out_gpu[2] == out_ref[2]
    assert [int(v) for v in out_gpu[3]] == [int(v) for v in out_ref[3]]
    assert np.abs(np.array(out_gpu[0]) - np.array(out_ref[0])).max() < 1e-6 * max(1.0, abs(out_ref[0][0]))


@needs_ref
def test_subproblem_wall_is_lifted_where_the_loop_reads_it():
    """cut_select_algo checks CutSolver._THRES_MAX_SUBS on the reference's module-level class (cut_select_qp.py:117)."""
    ref, refq, d = REF
    import sdpcutsel_via_nn_b200 as pkg
    GpuSolver, _ = pkg.make_solvers(ref, refq)
    assert ref.CutSolver._THRES_MAX_SUBS == pkg.B200CutSelection._THRES_MAX_SUBS > 4 * 10 ** 6
    with refloader.in_reference_dir(d):
        out = GpuSolver().cut_select_algo("spar125-075-1", 3, 0.1, strat=1, nb_rounds_cuts=0)
    assert out[-1] == 133242
