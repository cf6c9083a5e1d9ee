"""CPU tests of the host-side logic of the product (no GPU): vertex-cover construction, the lazy agg_list,
synthetic inputs, weight blobs."""
import numpy as np
import pytest

import sdpcutsel_via_nn_b200 as pkg
from conftest import inst_arrays
from oracle import cutsel_oracle as orc


@pytest.mark.parametrize("name", ["spar030-060-1", "spar040-030-1", "spar050-030-1"])
@pytest.mark.parametrize("dim", [3, 4, 5])
def test_pattern_E_cover_matches_reference(golden, name, dim):
    n, Q_arr, adj = inst_arrays(golden, name)
    idx = pkg.cover.pattern_E(adj, dim)
    assert np.array_equal(idx, golden["cover_%s_d%d" % (name.replace("-", "_"), dim)])


def test_pattern_E_spar125(golden):
    n, Q_arr, adj = inst_arrays(golden, "spar125-075-1")
    idx = pkg.cover.pattern_E(adj, 3)
    assert idx.shape[0] == 133242                                   # data_tables nb_subproblems
    assert np.array_equal(idx, orc.cover_pattern_E(adj, 3)[0])


def test_agg_list_elements(golden):
    n, Q_arr, adj = inst_arrays(golden, "spar030-060-1")
    for dim in (3, 4, 5):
        agg = pkg.cover.AggList(n, dim, Q_arr, idx=pkg.cover.pattern_E(adj, dim))
        Qs = golden["agg_spar030_060_1_d%d_Qslice" % dim]
        me = golden["agg_spar030_060_1_d%d_maxelem" % dim]
        assert len(agg) == Qs.shape[0]
        for i in (0, 1, len(agg) // 2, len(agg) - 1):
            s, xi, q, m = agg[i]
            assert xi == orc.xarr_inds(n, s) and m == me[i]
            assert np.array_equal(np.array(q), Qs[i, :len(q)])
        assert agg[:] is agg and len(agg[2:5]) == 3
    allc = pkg.cover.AggList(n, 3, Q_arr, n_all=4060)
    assert len(allc) == 4060 and allc[4059][0] == [27, 28, 29] and allc[0][0] == [0, 1, 2]


def test_synthetic_inputs_match_oracle_recipe():
    Qf = pkg.synthetic.instance(40, 0.75, seed=7)
    assert np.array_equal(Qf, orc.synth_instance(40, 0.75, seed=7)) and np.array_equal(Qf, Qf.T)
    a, b = pkg.synthetic.boxqp_arrays(Qf), orc.boxqp_arrays(Qf)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(pkg.synthetic.lp_point(40, 8), orc.synth_point(40, 8))


@pytest.mark.parametrize("d", [2, 3, 4, 5])
def test_weight_blob_roundtrip(blobs, d):
    net = pkg.nn_weights.unpack_blob(blobs[d])
    assert np.array_equal(pkg.nn_weights.pack_blob(net), blobs[d])
    nin, h = d * (d + 3) // 2, (64 if d in (2, 5) else 50)
    assert net["W"][0].shape == (h, nin) and net["W"][-1].shape == (1, h) and len(net["W"]) == (5 if d == 5 else 4)


def test_solver_surface_names():
    cs = pkg.CutSolver()
    for name in ("_load_neural_nets", "_get_sdp_vertex_cover", "_sel_eigcut_by_ordering_on_measure", "_gen_eigcuts_selected",
                 "_get_eigendecomp", "_CutSolver__preprocess_triangle_ineq", "_CutSolver__separate_and_add_triangle"):
        assert callable(getattr(cs, name))
    assert hasattr(pkg.CutSolverQCQP(), "_CutSolverQCQP__get_vertex_cover")
    assert (cs._THRES_NEG_EIGVAL, cs._BIG_M, cs._SDP_CUTS_PER_ROUND_MAX, cs._TRI_CUTS_PER_ROUND_MAX) == (-1e-15, 1000, 5000, 10000)
    with pytest.raises(NotImplementedError):
        cs._get_sdp_vertex_cover(3, ch_ext=1)
    with pytest.raises(ValueError):
        cs._sel_eigcut_by_ordering_on_measure(6, None, 1)


def test_random_selection_shuffles_like_the_reference():
    """strat 5 (cut_select_qp.py:634-637): np.random.shuffle(agg_list) in place.  The lazy cover is permuted with the same
    draws (numpy's legacy shuffle of an index array of the same length), stays shuffled, and a plain list of reference
    tuples is shuffled by the very same statement.  Host logic only: no device call before a score is asked for."""
    import itertools
    n, dim = 9, 3
    Q_arr = np.arange(n * (n + 1) // 2, dtype=np.float64) - 7.0
    ref_list = [tuple(c) for c in itertools.combinations(range(n), dim)]
    np.random.seed(7)
    np.random.shuffle(ref_list)
    ref_twice = list(ref_list)
    np.random.shuffle(ref_twice)

    cs = pkg.CutSolver()
    cs.set_instance(Q_arr, np.ones((n, n), dtype=np.uint8), n, dim=dim)
    cs._agg_list = pkg.cover.AggList(n, dim, Q_arr, n_all=84)
    np.random.seed(7)
    out = cs._sel_eigcut_by_ordering_on_measure(5, None, 1)
    assert out is cs._agg_list and len(out) == 84 and not out.is_all
    assert [tuple(e[0]) for e in out] == ref_list
    s, xi, qs, me = out[5]
    assert xi == pkg.cover.xarr_inds(n, s) and me == dim * np.abs(Q_arr[xi]).max()          # cut_select_qp.py:530-538
    out2 = cs._sel_eigcut_by_ordering_on_measure(5, None, 2)                                   # the list stays shuffled
    assert [tuple(e[0]) for e in out2] == ref_twice
    # a plain list in the reference's own format (the QCQP caller re-points _agg_list to such lists)
    plain = [(list(c), [], (), 1.0) for c in itertools.combinations(range(n), dim)]
    cs._agg_list = plain
    np.random.seed(7)
    assert cs._sel_eigcut_by_ordering_on_measure(5, None, 1) is plain
    assert [tuple(e[0]) for e in plain] == ref_list


def test_triangle_rows_csr_match_the_reference_rows():
    """sdpcs_triangle_rows_csr (host utility of the C ABI) against the restated row assembly of
    cut_select_qp.py:846-860 for every type, incl. ragged row lengths (4 / 6 entries) and m = 0."""
    n = 23
    rng = np.random.default_rng(3)
    T = pkg._capi.binom(n, 3)
    ranks = np.concatenate([[0, T - 1], rng.integers(0, T, 200)])
    types = np.concatenate([[3, 0], rng.integers(0, 4, 200)]).astype(np.int8)
    csr = pkg._capi.triangle_rows_csr(n, ranks, types)
    triples = pkg._capi.unrank(n, 3, ranks)
    assert csr["rowptr"][0] == 0 and csr["rowptr"][-1] == csr["ind"].size == csr["val"].size
    for r in range(ranks.size):
        ind, val, rhs = orc.triangle_row(n, tuple(int(v) for v in triples[r]), int(types[r]))
        a, b = csr["rowptr"][r], csr["rowptr"][r + 1]
        assert csr["ind"][a:b].tolist() == list(ind) and csr["val"][a:b].tolist() == list(val) and csr["rhs"][r] == rhs
    empty = pkg._capi.triangle_rows_csr(n, [], [])
    assert empty["rowptr"].tolist() == [0] and empty["ind"].size == 0
    with pytest.raises(pkg._capi.SdpcsError):
        pkg._capi.triangle_rows_csr(n, [T], [0])
    with pytest.raises(pkg._capi.SdpcsError):
        pkg._capi.triangle_rows_csr(n, [0], [4])


def test_mixin_composition_with_the_reference_classes():
    """make_solvers in front of the unmodified reference (no device needed to check the wiring): the hot-path methods
    resolve to the GPU ones, the name-mangled private methods are overridden under the reference's class names, the
    QCQP class keeps the reference's cut_select_algo but not its O(N^2) cover algebra, and the 4e6 wall is lifted on the
    class cut_select_algo reads it from (cut_select_qp.py:117)."""
    import refloader
    loaded = refloader.load_reference()
    if loaded is None:
        pytest.skip("no reference tree")
    ref, refq, _ = loaded
    G, GQ = pkg.make_solvers(ref, refq)
    B = pkg.B200CutSelection
    for name in ("_load_neural_nets", "_get_sdp_vertex_cover", "_sel_eigcut_by_ordering_on_measure", "_gen_eigcuts_selected",
                 "_get_eigendecomp"):
        assert getattr(G, name) is getattr(B, name) and getattr(GQ, name) is getattr(B, name)
    assert G.cut_select_algo is ref.CutSolver.cut_select_algo and GQ.cut_select_algo is refq.CutSolverQCQP.cut_select_algo
    for name in ("_CutSolver__preprocess_triangle_ineq", "_CutSolver__separate_and_add_triangle", "_CutSolver__gen_dense_eigcuts"):
        assert name in G.__dict__
    assert "_CutSolverQCQP__get_vertex_cover" in GQ.__dict__
    assert [c.__name__ for c in GQ.__mro__[:5]] == ["CutSolverQCQP", "CutSolverQCQP", "_Mix", "B200CutSelection", "CutSolver"]
    assert ref.CutSolver._THRES_MAX_SUBS == B._THRES_MAX_SUBS == 2 ** 44
    g = G()
    assert g._Mat[0].shape == (3, 3) and g._blobs == {}          # both __init__ ran


def test_rank_list_seal_detects_any_change():
    """RankList._sealed_rows (cut_select_qp.py mirror): the arrays behind the entries are only handed out while the list
    holds exactly the entries it was returned with -- _gen_eigcuts_selected walks the entries otherwise."""
    RankList = pkg.cut_select_qp.RankList
    rows, scores = np.arange(12, dtype=np.int16).reshape(4, 3), np.array([4.0, 3.0, 2.0, 1.0])
    rl = RankList([([0, 1, 2], 4.0), ([3, 4, 5], 3.0), ([6, 7, 8], 2.0), ([9, 10, 11], 1.0)])
    assert rl._sealed_rows() is None                      # never sealed
    rl._seal(rows, scores)
    got = rl._sealed_rows()
    assert got is not None and got[0] is rows and got[1] is scores
    assert list(rl)._sealed_rows() is None if hasattr(list(rl), "_sealed_rows") else True   # a copy is a plain list
    for change in (lambda l: l.reverse(), lambda l: l.pop(), lambda l: l.append(([1], 0.0)),
                   lambda l: l.__setitem__(1, ([3, 4, 5], 3.0)), lambda l: l.sort(key=lambda e: e[1])):
        cp = RankList(rl)
        cp._seal(rows, scores)
        assert cp._sealed_rows() is not None
        change(cp)
        assert cp._sealed_rows() is None
    empty = RankList()
    empty._seal(np.zeros((0, 3), np.int16), np.zeros(0))
    assert empty._sealed_rows() is not None
