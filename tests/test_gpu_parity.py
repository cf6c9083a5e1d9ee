"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the
oracle on the same inputs and against the golden fixtures produced by the unmodified reference.

Tolerances (north_star): selected indices and order bit-exact; eigenvalues 1e-9 relative (we assert 1e-12
absolute, the matrices have O(1) entries); NN-based scores 1e-5 (we assert 1e-9 absolute)."""
import numpy as np
import pytest

from conftest import inst_arrays
from oracle import cutsel_oracle as orc

pytestmark = pytest.mark.gpu

LAM_TOL = 1e-12
OBJ_TOL = 1e-9


@pytest.fixture(scope="module")
def capi():
    import sdpcutsel_via_nn_b200 as pkg
    return pkg._capi


def make_engine(capi, blobs, n, Q_arr, rhos=(2, 3, 4, 5)):
    eng = capi.Engine(0)
    for d in rhos:
        eng.set_weights(d, blobs[d])
    eng.set_instance(n, Q_arr)
    return eng


def _sets(idx):
    return [tuple(int(v) for v in r if v >= 0) for r in idx]


@pytest.mark.parametrize("d", [2, 3, 4, 5])
def test_nn_eval_vs_NNs_so(capi, blobs, golden, d):
    eng = capi.Engine(0)
    eng.set_weights(d, blobs[d])
    y = eng.nn_eval(d, golden["nn%d_in" % d])
    assert np.abs(y - golden["nn%d_out" % d]).max() < 1e-11


def test_unrank_matches_itertools(capi):
    for n, rho in [(30, 3), (20, 4), (17, 5), (250, 2)]:
        want = orc.cover_all(n, rho)
        got = capi.unrank(n, rho, np.arange(want.shape[0]))
        assert np.array_equal(got, want)
        assert capi.binom(n, rho) == want.shape[0]
    assert capi.binom(250, 5) == 7817031300


def test_cfg1_all_triples(capi, blobs, golden):
    n, Q_arr, adj = inst_arrays(golden, "spar030-060-1")
    eng = make_engine(capi, blobs, n, Q_arr)
    eng.set_cover_all(3)
    assert eng.num_candidates == 4060
    vv = golden["cfg1_vars"]
    idx = orc.cover_all(n, 3)
    lam_o, obj_o = orc.score_cover(Q_arr, n, idx, np.full(4060, 3), vv, blobs)
    eng.score(vv, 3)
    lam, obj = eng.scores()
    assert np.abs(lam - lam_o).max() < LAM_TOL
    assert np.abs(obj - obj_o).max() < OBJ_TOL
    k = 406
    r = eng.select(1, vv, 5000)                      # strat 1: whole violated list
    assert _sets(idx[r["idx"]]) == _sets(golden["cfg1_s1_sets"])
    assert np.abs(r["score"] - golden["cfg1_s1_score"]).max() < LAM_TOL
    r = eng.select(2, vv, 4060)                      # strat 2: the complete ranking
    assert np.array_equal(r["idx"], golden["cfg1_s2_idx"])
    assert np.abs(r["score"] - golden["cfg1_s2_score"]).max() < OBJ_TOL
    r = eng.select(4, vv, k)
    assert r["new_strat"] == int(golden["cfg1_s4_newstrat"])
    assert np.array_equal(r["idx"], golden["cfg1_s4_idx"][:k])
    assert np.abs(r["score"] - golden["cfg1_s4_score"][:k]).max() < OBJ_TOL
    # combined rule where fewer strong cuts than k exist: the whole list is walked
    ns, order, score = orc.select_comb(obj_o, lam_o, 3000)
    r = eng.select(4, vv, 3000)
    assert r["new_strat"] == ns and np.array_equal(r["idx"], order[:3000])
    assert np.abs(r["score"] - score[:3000]).max() < OBJ_TOL


def test_cfg1_eigcuts(capi, blobs, golden):
    n, Q_arr, adj = inst_arrays(golden, "spar030-060-1")
    eng = make_engine(capi, blobs, n, Q_arr, rhos=())
    vv = golden["cfg1_vars"]
    sets = golden["cfg1_s1_sets"][:406]
    ind, val, rhs, lam, viol = eng.gen_cuts(3, sets, vv)
    assert viol.all()
    assert np.array_equal(ind, golden["cfg1_s1_cut_ind"])
    # eigenvector sign is arbitrary but the cut coefficients are products -> sign free
    assert np.abs(val - golden["cfg1_s1_cut_val"]).max() < 1e-9
    assert np.abs(rhs - golden["cfg1_s1_cut_rhs"]).max() < 1e-9
    assert np.abs(-lam - golden["cfg1_s1_score"][:406]).max() < LAM_TOL
    # eigendecomp entry point vs LAPACK on one subset
    X_vals, x_vals = orc.split_vars(vv, n)
    s = [int(v) for v in sets[0]]
    Xs = X_vals[orc.xarr_inds(n, s)]
    w, V = eng.eigendecomp(3, x_vals[s], Xs)
    M = orc.eig_matrix(x_vals[s][None, :], Xs[None, :])[0]
    w_ref = np.linalg.eigvalsh(M, "U")
    assert np.abs(w - w_ref).max() < 1e-13
    Mfull = np.triu(M) + np.triu(M, 1).T
    assert np.abs(Mfull @ V - V * w[None, :]).max() < 1e-12


@pytest.mark.parametrize("dim", [3, 4, 5])
def test_mixed_size_cover(capi, blobs, golden, dim):
    n, Q_arr, adj = inst_arrays(golden, "spar040-030-1")
    idx, sizes = orc.cover_pattern_E(adj, dim)
    eng = make_engine(capi, blobs, n, Q_arr)
    eng.set_cover_list(dim, idx)
    vv = golden["mix_vars"]
    N = idx.shape[0]
    k = max(1, min(int(np.floor(0.1 * N)), 5000))
    lam_o, obj_o = orc.score_cover(Q_arr, n, idx, sizes, vv, blobs)
    eng.score(vv, 3)
    lam, obj = eng.scores()
    assert np.abs(lam - lam_o).max() < LAM_TOL and np.abs(obj - obj_o).max() < OBJ_TOL
    r = eng.select(1, vv, N)
    assert _sets(idx[r["idx"]]) == _sets(golden["mix_d%d_s1_sets" % dim])
    r = eng.select(2, vv, N)
    assert np.array_equal(r["idx"], golden["mix_d%d_s2_idx" % dim])
    r = eng.select(4, vv, k)
    assert r["new_strat"] == int(golden["mix_d%d_s4_newstrat" % dim])
    assert np.array_equal(r["idx"], golden["mix_d%d_s4_idx" % dim][:k])
    assert np.abs(r["score"] - golden["mix_d%d_s4_score" % dim][:k]).max() < OBJ_TOL
    # cuts for the optimality selection, mixed subset sizes in one call
    top = golden["mix_d%d_s2_idx" % dim][:k]
    ind, val, rhs, lamc, viol = eng.gen_cuts(dim, idx[top], vv)
    want_ind, want_val, want_rhs = (golden["mix_d%d_s2_cut_%s" % (dim, s)] for s in ("ind", "val", "rhs"))
    assert int(viol.sum()) == want_rhs.shape[0]
    assert np.array_equal(ind[viol], want_ind)
    assert np.abs(val[viol] - want_val).max() < 1e-9 and np.abs(rhs[viol] - want_rhs).max() < 1e-9


def test_cfg2_spar125_pattern_E_and_triangles(capi, blobs, golden):
    n, Q_arr, adj = inst_arrays(golden, "spar125-075-1")
    idx, sizes = orc.cover_pattern_E(adj, 3)
    eng = make_engine(capi, blobs, n, Q_arr)
    eng.set_cover_list(3, idx)
    vv = golden["cfg2_vars"]
    k = 5000
    r = eng.select(1, vv, k)
    assert _sets(idx[r["idx"]]) == _sets(golden["cfg2_s1_sets"])
    assert np.abs(r["score"] - golden["cfg2_s1_score"]).max() < LAM_TOL
    assert int(r["counts"][1]) == int(golden["cfg2_s1_len"])          # number of violated subsets
    r = eng.select(2, vv, k)
    assert np.array_equal(r["idx"], golden["cfg2_s2_idx"])
    assert np.abs(r["score"] - golden["cfg2_s2_score"]).max() < OBJ_TOL
    r = eng.select(4, vv, k)
    assert r["new_strat"] == int(golden["cfg2_s4_newstrat"])
    assert np.array_equal(r["idx"], golden["cfg2_s4_idx"])
    assert np.abs(r["score"] - golden["cfg2_s4_score"]).max() < OBJ_TOL
    # triangles: bit-exact (adds only)
    eng.set_tri_pattern(adj)
    t = eng.triangles(vv, 10000)
    tri, dens = orc.triangles_pre(adj, n)
    pos, typ, viol = orc.triangles_sep(n, tri, dens, vv, 0.1)
    assert t["n_triples"] == tri.shape[0] == int(golden["cfg2_tri_ntriples"])
    assert t["rank"].shape[0] == 10000
    got_tri = capi.unrank(n, 3, t["rank"])
    assert np.array_equal(got_tri, tri[pos]) and np.array_equal(t["type"], typ)
    assert np.array_equal(t["viol"], viol)                              # bit-exact
    for i in (0, 1, 5000, 9999):
        ind, val, rhs = orc.triangle_row(n, got_tri[i], int(t["type"][i]))
        m = len(ind)
        assert ind == list(golden["cfg2_tri_ind"][i][:m]) and rhs == golden["cfg2_tri_rhs"][i]


@pytest.mark.parametrize("dim", [3, 4, 5])
def test_qcqp_all_subsets(capi, blobs, golden, dim):
    n = 20
    Q_arr = golden["qcqp_Q_arr"]
    eng = make_engine(capi, blobs, n, Q_arr)
    eng.set_cover_all(dim)
    N = eng.num_candidates
    assert N == int(golden["qcqp_d%d_N" % dim][0])
    vv = golden["qcqp_vars"]
    k = max(1, min(int(np.floor(0.1 * N)), 5000))
    idx = orc.cover_all(n, dim)
    r = eng.select(1, vv, k)
    assert _sets(idx[r["idx"]]) == _sets(golden["qcqp_d%d_s1_sets" % dim])
    assert np.abs(r["score"] - golden["qcqp_d%d_s1_score" % dim]).max() < LAM_TOL
    r = eng.select(4, vv, k)
    assert r["new_strat"] == int(golden["qcqp_d%d_s4_newstrat" % dim])
    assert np.array_equal(r["idx"], golden["qcqp_d%d_s4_idx" % dim])
    assert np.abs(r["score"] - golden["qcqp_d%d_s4_score" % dim]).max() < OBJ_TOL


@pytest.mark.parametrize("n,rho,density", [(40, 5, 0.75), (60, 4, 0.5), (125, 3, 0.75)])
def test_all_subsets_vs_oracle_medium(capi, blobs, n, rho, density):
    """Synthetic instances (SURVEY 8d recipe), hundreds of thousands of candidates, both scores + top-k."""
    Q_arr, adj = orc.boxqp_arrays(orc.synth_instance(n, density, seed=7))
    vv = orc.synth_point(n, seed=8)
    eng = make_engine(capi, blobs, n, Q_arr)
    eng.set_cover_all(rho)
    N = eng.num_candidates
    idx = orc.cover_all(n, rho)
    assert N == idx.shape[0]
    lam_o, obj_o = orc.score_cover(Q_arr, n, idx, np.full(N, rho), vv, blobs)
    eng.score(vv, 3)
    lam, obj = eng.scores()
    assert np.abs(lam - lam_o).max() < LAM_TOL and np.abs(obj - obj_o).max() < OBJ_TOL
    k = 5000
    order, score = orc.select_feas(lam_o)
    r = eng.select(1, vv, k)
    assert np.array_equal(r["idx"], order[:k])
    order, score = orc.select_opt(obj_o)
    r = eng.select(2, vv, k)
    assert np.array_equal(r["idx"], order[:k])
    ns, order, score = orc.select_comb(obj_o, lam_o, k)
    r = eng.select(4, vv, k)
    assert r["new_strat"] == ns and np.array_equal(r["idx"], order[:k])
    # sharded rank ranges reproduce the same scores (multi-GPU partitioning, one shard at a time)
    cut = N // 3 + 7
    eng.set_cover_all(rho, cut, N)
    eng.score(vv, 3)
    lam2, obj2 = eng.scores()
    assert np.array_equal(lam2, lam[cut:]) and np.array_equal(obj2, obj[cut:])


def _lapack_here_reproduces(golden_scores, live_scores):
    """The golden eigenvalue scores came out of LAPACK in the authoring container; OpenBLAS picks its kernels by CPU, so a
    different host may round differently.  Orders inside exact-tie classes are only comparable when the scores agree
    bit for bit; otherwise the oracle executed on this host is the reference ("the reference's call executed here")."""
    return golden_scores.shape == live_scores.shape and np.array_equal(golden_scores, live_scores)


def test_degenerate_vertex_is_the_references_order(capi, blobs, golden):
    """x = 0.5, X in {0, 0.5} (what round 1 of every BoxQP run looks like): 4060 violated triples in 12 classes of up to
    2661 exact ties; the reference's order inside a class is LAPACK round-off.  The device returns winners + guard band,
    the mirror re-scores the near-tie runs with the reference's arithmetic: the list equals the reference's own
    (golden cfg1_deg_s1_sets, generated by the unmodified reference)."""
    import sdpcutsel_via_nn_b200 as pkg
    n, Q_arr, adj = inst_arrays(golden, "spar030-060-1")
    idx = orc.cover_all(n, 3)
    vd = orc.degenerate_point(n, Q_arr)
    lam_o, _ = orc.score_cover(Q_arr, n, idx, np.full(4060, 3), vd, want_obj=False)
    order_o, score_o = orc.select_feas(lam_o)
    # raw device pass: scores within tolerance, own order (score desc, index asc), band and counters reported
    eng = make_engine(capi, blobs, n, Q_arr)
    eng.set_cover_all(3)
    eng.score(vd, 1)
    lam, _ = eng.scores(obj=False)
    assert np.abs(lam - lam_o).max() < LAM_TOL
    r = eng.select(1, vd, 100)
    viol = lam < -1e-15
    own = np.arange(4060)[viol][np.lexsort((np.arange(4060)[viol], lam[viol]))]
    assert np.array_equal(r["idx"], own[:100])
    band = eng.last_band()
    kth = -lam[own[99]]
    in_band = own[100:][-lam[own[100:]] >= kth - 1e-12]
    assert band["n_band"] == in_band.size > 0 and np.array_equal(band["idx"], in_band) and band["band_open"] == 0
    # the drop-in surface: identical to the reference's list
    cs = pkg.CutSolver()
    cs.set_instance(Q_arr, adj, n, dim=3)
    cs._get_sdp_vertex_cover(3, ch_ext=-1)
    for sel_size in (0, 406, 4060):
        rl = cs._sel_eigcut_by_ordering_on_measure(1, vd, 1, sel_size=sel_size)
        assert rl.degenerate == 1 and rl.n_near_ties > 4000 and rl.n_violated == order_o.size
        assert [e[0] for e in rl] == [[int(v) for v in r] for r in idx[order_o]]
        assert np.abs(np.array([e[1] for e in rl]) - score_o).max() < LAM_TOL
        if _lapack_here_reproduces(golden["cfg1_deg_s1_score"], score_o):
            assert [e[0] for e in rl] == [[int(v) for v in r] for r in golden["cfg1_deg_s1_sets"]]
    # a prefix shorter than the list: the near ties of the k-th place come from the band
    cs._RANK_PREFIX = 150
    rl = cs._sel_eigcut_by_ordering_on_measure(1, vd, 1, sel_size=100)
    assert [e[0] for e in rl] == [[int(v) for v in r] for r in idx[order_o[:150]]] and rl.degenerate == 1
    # a band that cannot hold the tie class is reported
    eng2 = cs._engine_for(cs._agg_list)
    eng2.set_params(band_cap=64)
    rl = cs._sel_eigcut_by_ordering_on_measure(1, vd, 1, sel_size=100)
    assert rl.degenerate == 2 and len(rl) == 150


def test_fig8_lp_vertex_through_the_dropin(capi, blobs, golden):
    """The reference's own round-1 LP vertex of spar020-100-1 and its committed ranking (data_figures/fig8_data.csv:7-1057,
    golden fig8_r1_cut_idx): optimality ranking identical in index and order; feasibility and combined rankings equal to
    the reference's rules on LAPACK eigenvalues (64 PSD-singular triples sit inside the guard of the -1e-15 threshold)."""
    import sdpcutsel_via_nn_b200 as pkg
    n, Q_arr, adj = inst_arrays(golden, "spar020-100-1")
    vv = golden["fig8_vars"]
    idx, sizes = orc.cover_pattern_E(adj, 3)
    lam_o, obj_o = orc.score_cover(Q_arr, n, idx, sizes, vv, blobs)
    cs = pkg.CutSolver()
    cs.set_instance(Q_arr, adj, n, dim=3)
    cs._load_neural_nets()
    N = cs._get_sdp_vertex_cover(3)
    assert N == 1051 == golden["fig8_r1_cut_idx"].size
    rl = cs._sel_eigcut_by_ordering_on_measure(2, vv, 1)
    assert [e[0] for e in rl] == golden["fig8_r1_cut_idx"].tolist()
    assert np.abs(np.array([e[1] for e in rl]) - golden["fig8_r1_estim"]).max() < 1e-9 and rl.degenerate == 0
    order, score = orc.select_feas(lam_o)
    rl = cs._sel_eigcut_by_ordering_on_measure(1, vv, 1)
    assert [e[0] for e in rl] == [[int(v) for v in r] for r in idx[order]] and rl.n_violated == order.size
    assert rl.degenerate == 1 and rl.n_near_ties >= 64
    k = 105
    ns, order, score = orc.select_comb_walk(obj_o, lam_o, k)
    new_strat, rl = cs._sel_eigcut_by_ordering_on_measure(4, vv, 1, sel_size=k)
    assert new_strat == ns and [e[0] for e in rl] == order[:k].tolist() and rl.degenerate in (0, 1)
    assert np.abs(np.array([e[1] for e in rl]) - score[:k]).max() < OBJ_TOL


def test_guard_is_silent_on_non_degenerate_points(capi, blobs, golden):
    n, Q_arr, adj = inst_arrays(golden, "spar125-075-1")
    import sdpcutsel_via_nn_b200 as pkg
    cs = pkg.CutSolver()
    cs.set_instance(Q_arr, adj, n, dim=3)
    cs._load_neural_nets()
    cs._get_sdp_vertex_cover(3)
    for strat in (1, 2, 4):
        out = cs._sel_eigcut_by_ordering_on_measure(strat, golden["cfg2_vars"], 1, sel_size=5000)
        rl = out[1] if strat == 4 else out
        assert rl.degenerate == 0 and rl.n_near_ties == 0


def test_errors_are_loud(capi, blobs):
    eng = capi.Engine(0)
    with pytest.raises(capi.SdpcsError):
        eng.set_cover_all(3)                       # no instance
    eng.set_instance(10, np.zeros(55))
    with pytest.raises(capi.SdpcsError):
        eng.set_weights(3, blobs[4])               # wrong net
    eng.set_cover_all(3)
    with pytest.raises(capi.SdpcsError):
        eng.score(np.zeros(65), 2)                 # NN weights missing
    with pytest.raises(capi.SdpcsError):
        eng.set_cover_list(3, np.array([[3, 2, 1]], dtype=np.int16))
    # empty / tiny covers
    eng.set_cover_list(3, np.zeros((0, 3), dtype=np.int16))
    assert eng.num_candidates == 0
    r = eng.select(1, np.zeros(65), 10)
    assert r["idx"].size == 0


@pytest.mark.parametrize("rho", [2, 3, 4, 5])
def test_eigen_methods_vs_lapack(capi, blobs, golden, rho):
    """Default eigen path (Householder tridiagonalisation + Laguerre) and the Jacobi path (jacobi_sweeps > 0)
    against numpy/LAPACK eigvalsh on random and on degenerate (x = 0.5, X in {0, 0.5}) LP points."""
    n = 26
    Q_arr, adj = orc.boxqp_arrays(orc.synth_instance(n, 0.6, seed=11))
    idx = orc.cover_all(n, rho)
    sizes = np.full(idx.shape[0], rho)
    for vv in (orc.synth_point(n, seed=3), orc.degenerate_point(n, Q_arr), np.zeros(n * (n + 1) // 2 + n),
               np.ones(n * (n + 1) // 2 + n)):
        lam_o, _ = orc.score_cover(Q_arr, n, idx, sizes, vv, want_obj=False)
        for sweeps in (0, 6):
            eng = capi.Engine(0)
            eng.set_params(jacobi_sweeps=sweeps)
            eng.set_instance(n, Q_arr)
            eng.set_cover_all(rho)
            eng.score(vv, 1)
            lam, _ = eng.scores(obj=False)
            assert np.abs(lam - lam_o).max() < 2e-14, (rho, sweeps)


def test_merge_topk_sorted_runs_and_unsorted_input(capi):
    """sdpcs_merge_topk: concatenated per-shard lists (sorted runs: host k-way merge) and arbitrary input (device rank
    sort) both give the first k of the order (score desc, obj2 desc, agg_idx asc), ties included."""
    rng = np.random.default_rng(12)
    eng = capi.Engine(0)
    for runs, per, with_obj2 in ((8, 5000, True), (2, 7, False), (1, 100, True), (300, 40, True)):
        score = np.round(rng.normal(size=runs * per), 1)                 # many exact ties
        obj2 = np.round(rng.normal(size=runs * per), 1)
        idx = rng.permutation(runs * per).astype(np.int64)
        rows = []
        for r in range(runs):                                             # sort each run the way a shard would
            sl = slice(r * per, (r + 1) * per)
            o = np.lexsort((idx[sl], -obj2[sl] if with_obj2 else np.zeros(per), -score[sl]))
            rows.append((score[sl][o], obj2[sl][o], idx[sl][o]))
        score, obj2, idx = (np.concatenate([x[i] for x in rows]) for i in range(3))
        for k in (1, per, runs * per, runs * per + 5):
            want = np.lexsort((idx, -obj2 if with_obj2 else np.zeros(idx.size), -score))[:k]
            got = eng.merge_topk(score, obj2 if with_obj2 else None, idx, k)
            assert np.array_equal(got, want)
    # unsorted input
    score, idx = rng.normal(size=3000), rng.permutation(3000).astype(np.int64)
    assert np.array_equal(eng.merge_topk(score, None, idx, 100), np.argsort(-score, kind="stable")[:100])


def test_combined_rule_shortcut_and_general_path(capi, blobs):
    """sdpcs_select strat 4: when the strong set has >= k elements and no non-violated candidate comes within 2 big_m of
    the pivot, the strong list of pass 1 is returned directly; with a tiny big_m the condition fails and the two-pass
    path runs.  Both must equal the oracle's literal walk (cut_select_qp.py:601-630), scores included."""
    n, rho = 26, 4
    Q_arr, adj = orc.boxqp_arrays(orc.synth_instance(n, 0.75, seed=31))
    vv = orc.synth_point(n, seed=32)
    idx = orc.cover_all(n, rho)
    lam_o, obj_o = orc.score_cover(Q_arr, n, idx, np.full(idx.shape[0], rho), vv, blobs)
    for k in (1, 50, 700, idx.shape[0]):
        ns, order, score = orc.select_comb_walk(obj_o, lam_o, k)
        eng = make_engine(capi, blobs, n, Q_arr, rhos=(rho,))
        eng.set_cover_all(rho)
        r = eng.select(4, vv, k)
        assert r["new_strat"] == ns and np.array_equal(r["idx"], order[:k])
        assert np.abs(r["score"] - score[:k]).max() < OBJ_TOL
        assert np.abs(r["lam"] - lam_o[r["idx"]]).max() < LAM_TOL and np.abs(r["obj"] - obj_o[r["idx"]]).max() < OBJ_TOL
    # A point between the random one and a PSD one: some candidates with obj > 0 are NOT violated.  With big_m = 1000 the
    # shortcut still applies (oracle walk), with a tiny big_m the non-violated ones overtake: general two-pass path,
    # compared with a restatement of the rule.
    nb = n * (n + 1) // 2
    x, iu = vv[nb:], np.triu_indices(n)
    Xr = np.zeros((n, n))
    Xr[iu] = vv[:nb]
    Xp = np.outer(x, x) + np.diag(np.full(n, 0.2))
    v2 = np.concatenate([(0.1 * Xp + 0.9 * (Xr + np.triu(Xr, 1).T))[iu], x])
    lam_o, obj_o = orc.score_cover(Q_arr, n, idx, np.full(idx.shape[0], rho), v2, blobs)
    strong_o = (obj_o > 0) & (lam_o < -1e-15)
    k = int(strong_o.sum())                       # the pivot is the weakest strong element
    assert k >= 5 and ((obj_o > 0) & ~(lam_o < -1e-15)).any()
    eng = make_engine(capi, blobs, n, Q_arr, rhos=(rho,))
    eng.set_cover_all(rho)
    ns, order, score = orc.select_comb_walk(obj_o, lam_o, k)
    r = eng.select(4, v2, k)
    assert r["new_strat"] == ns and np.array_equal(r["idx"], order[:k]) and np.abs(r["score"] - score[:k]).max() < OBJ_TOL
    eng.set_params(big_m=1e-3)
    r = eng.select(4, v2, k)
    lam, obj = eng.scores()
    o = np.argsort(-obj, kind="stable")
    viol = lam[o] < -1e-15
    strong = (obj[o] > 0) & viol
    pivot = np.nonzero(np.cumsum(strong) == k)[0][0]
    f = obj[o].copy()
    w = np.arange(o.size) <= pivot
    f[w & strong] += 1e-3
    f[w & (obj[o] > 0) & ~viol] -= 1e-3
    want = o[np.argsort(-f, kind="stable")][:k]
    assert np.array_equal(r["idx"], want)
    nonviol_pos = (obj > 0) & ~(lam < -1e-15)
    assert obj[nonviol_pos].max() - 1e-3 >= obj[o][pivot] + 1e-3          # the shortcut condition does not hold here ...
    assert not np.array_equal(np.sort(want), np.sort(o[np.nonzero(strong)[0][:k]]))   # ... and it would have been wrong
