"""GPU tests of the device-built pattern-E vertex cover (sdpcs_set_cover_pattern, cover_kernels.cuh) through the C ABI.

Bit-exact (set AND order) against: the covers recorded from the unmodified reference (tests/golden, three BoxQP
instances x rho = 3, 4, 5, and spar125-075-1 rho = 3 with its published count 133,242), the oracle's restatement of
the reference loops on random graphs of every density (empty, ragged, complete, n up to 250), and the host DFS of the
package.  Scores computed on a device-built cover equal those on the same cover shipped as a list."""
import numpy as np
import pytest

from conftest import inst_arrays
from oracle import cutsel_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    import sdpcutsel_via_nn_b200 as pkg
    return pkg._capi


def _engine(capi, n, Q_arr=None):
    eng = capi.Engine(0)
    eng.set_instance(n, np.zeros(n * (n + 1) // 2) if Q_arr is None else Q_arr)
    return eng


@pytest.mark.parametrize("name", ["spar030-060-1", "spar040-030-1", "spar050-030-1"])
@pytest.mark.parametrize("dim", [3, 4, 5])
def test_reference_covers(capi, golden, name, dim):
    n, Q_arr, adj = inst_arrays(golden, name)
    eng = _engine(capi, n, Q_arr)
    N = eng.set_cover_pattern(dim, adj)
    want = golden["cover_%s_d%d" % (name.replace("-", "_"), dim)]
    assert N == want.shape[0] == eng.num_candidates
    assert np.array_equal(eng.cover_rows(), want)


def test_spar125_rho3_count_and_order(capi, golden):
    n, Q_arr, adj = inst_arrays(golden, "spar125-075-1")
    eng = _engine(capi, n, Q_arr)
    assert eng.set_cover_pattern(3, adj) == 133242                   # data_tables nb_subproblems
    assert np.array_equal(eng.cover_rows(), orc.cover_pattern_E(adj, 3)[0])


def _random_adj(n, density, seed):
    rng = np.random.default_rng(seed)
    a = np.triu(rng.random((n, n)) < density, 1)
    a = a | a.T
    a[np.arange(n), np.arange(n)] = rng.random(n) < 0.5               # the diagonal must not matter
    return a.astype(np.uint8)


@pytest.mark.parametrize("n,density", [(12, 0.0), (12, 1.0), (17, 0.5), (40, 0.2), (64, 0.35), (65, 0.3), (130, 0.08), (250, 0.03)])
@pytest.mark.parametrize("dim", [2, 3, 4, 5])
def test_random_graphs_match_the_reference_loops(capi, n, density, dim):
    adj = _random_adj(n, density, seed=1000 * n + dim)
    eng = _engine(capi, n)
    N = eng.set_cover_pattern(dim, adj)
    rows = eng.cover_rows()
    if dim >= 3 and n <= 65:
        loops = orc.cover_pattern_E_loops(adj, dim)                   # the reference's nested loops, restated
        want = np.full((len(loops), dim), -1, dtype=np.int16)
        for i, t in enumerate(loops):
            want[i, :len(t)] = t
    else:
        import sdpcutsel_via_nn_b200 as pkg
        want = pkg.cover.pattern_E(adj, dim)
    assert N == want.shape[0]
    assert np.array_equal(rows, want.reshape(-1, dim))
    if density == 1.0:
        assert N == capi.binom(n, dim)
    if density == 0.0:
        assert N == 0


def test_one_sided_pattern_and_asymmetric_input(capi):
    """Q_adj of the BoxQP reader is symmetric, the QCQP one may be given as either triangle."""
    n = 30
    adj = _random_adj(n, 0.4, seed=5)
    eng = _engine(capi, n)
    eng.set_cover_pattern(4, adj)
    full = eng.cover_rows()
    eng.set_cover_pattern(4, np.triu(adj))
    assert np.array_equal(eng.cover_rows(), full)
    eng.set_cover_pattern(4, np.tril(adj))
    assert np.array_equal(eng.cover_rows(), full)


@pytest.mark.parametrize("dim", [4, 5])
def test_scores_on_device_built_cover(capi, blobs, golden, dim):
    n, Q_arr, adj = inst_arrays(golden, "spar040-030-1")
    vv = golden["mix_vars"]
    a = capi.Engine(0)
    b = capi.Engine(0)
    for eng in (a, b):
        for d in range(2, dim + 1):
            eng.set_weights(d, blobs[d])
        eng.set_instance(n, Q_arr)
    a.set_cover_pattern(dim, adj)
    idx = a.cover_rows()
    b.set_cover_list(dim, idx)
    a.score(vv, 3)
    b.score(vv, 3)
    la, oa = a.scores()
    lb, ob = b.scores()
    assert np.array_equal(la, lb) and np.array_equal(oa, ob)
    k = max(1, idx.shape[0] // 10)
    ra, rb = a.select(4, vv, k), b.select(4, vv, k)
    assert np.array_equal(ra["idx"], rb["idx"]) and np.array_equal(ra["idx"], golden["mix_d%d_s4_idx" % dim][:k])


def test_large_sparse_instance_rho5(capi):
    """n = 125 at 30 % density, rho = 5: ~10^5..10^6 candidates of mixed sizes; order checked by sortedness and the
    clique / maximality properties on a sample, count against the host DFS."""
    import sdpcutsel_via_nn_b200 as pkg
    n = 125
    adj = _random_adj(n, 0.3, seed=77)
    eng = _engine(capi, n)
    N = eng.set_cover_pattern(5, adj)
    rows = eng.cover_rows().astype(np.int64)
    want = pkg.cover.pattern_E(adj, 5)
    assert N == want.shape[0] and np.array_equal(rows, want)
    key = np.where(rows < 0, -1, rows)
    for i in range(0, N - 1, max(1, N // 5000)):                      # lexicographic order of the tuples
        a, b = tuple(x for x in key[i] if x >= 0), tuple(x for x in key[i + 1] if x >= 0)
        assert a < b


def test_errors(capi):
    eng = capi.Engine(0)
    with pytest.raises(capi.SdpcsError):
        eng.set_cover_pattern(3, np.zeros((4, 4)))                    # no instance yet
    eng.set_instance(6, np.zeros(21))
    with pytest.raises(capi.SdpcsError):
        eng.set_cover_pattern(6, np.ones((6, 6)))                     # rho out of range
    with pytest.raises(capi.SdpcsError):
        eng.set_cover_pattern(3, np.ones((5, 5)))                     # wrong shape
    eng.set_cover_all(3)
    with pytest.raises(capi.SdpcsError):
        eng.cover_rows()                                              # no list cover


@pytest.mark.parametrize("dim", [3, 5])
def test_cover_restrict_shards_a_list_cover(capi, blobs, golden, dim):
    """sdpcs_cover_restrict: every rank builds the pattern cover and keeps its contiguous shard.  Scores of
    a shard are the slice of the full scores, agg_idx values are unchanged, and the merge of the shard selections is
    the full selection."""
    n, Q_arr, adj = inst_arrays(golden, "spar040-030-1")
    vv = golden["mix_vars"]

    def engine():
        eng = capi.Engine(0)
        for d in range(2, dim + 1):
            eng.set_weights(d, blobs[d])
        eng.set_instance(n, Q_arr)
        return eng

    full = engine()
    N = full.set_cover_pattern(dim, adj)
    rows = full.cover_rows()
    full.score(vv, 3)
    lam, obj = full.scores()
    k = max(1, N // 10)
    want = full.select(2, vv, k)
    cuts = [0, N // 3 + 1, N // 3 + 1, 2 * N // 3, N]          # includes an empty shard
    parts = []
    for b, e in zip(cuts[:-1], cuts[1:]):
        eng = engine()
        eng.set_cover_pattern(dim, adj)
        eng.cover_restrict(b, e)
        assert eng.num_candidates == e - b
        assert np.array_equal(eng.cover_rows(), rows[b:e])
        if e > b:
            eng.score(vv, 3)
            l2, o2 = eng.scores()
            assert np.array_equal(l2, lam[b:e]) and np.array_equal(o2, obj[b:e])
            parts.append(eng.topk(2, k))
            with pytest.raises(capi.SdpcsError):
                eng.cover_restrict(0, 1)                              # a list cover is restricted once
    gidx = np.concatenate([p[0] for p in parts])
    gsc = np.concatenate([p[1] for p in parts])
    perm = full.merge_topk(gsc, None, gidx, k)
    assert np.array_equal(gidx[perm], want["idx"]) and np.array_equal(gsc[perm], want["score"])
    with pytest.raises(capi.SdpcsError):
        full.cover_restrict(5, N + 1)
    # all-subsets cover: the rank range moves
    allc = engine()
    allc.set_cover_all(dim)
    allc.score(vv, 2)
    o_all = allc.scores(lam=False)[1]
    allc.cover_restrict(100, 1000)
    allc.cover_restrict(50, 500)                                      # relative to the current range: ranks 150 .. 600
    allc.score(vv, 2)
    assert np.array_equal(allc.scores(lam=False)[1], o_all[150:600])


@pytest.mark.parametrize("dim", [3, 4, 5])
def test_two_cover_algebra_on_the_device(dim):
    """cut_select_qcqp.py:319-333 (P(E_m) intersected with / minus P(E_0)) on the device vs the host set algebra and vs a
    literal restatement of the reference's list scans, on sparse objective / constraint patterns (mixed clique sizes)."""
    import sdpcutsel_via_nn_b200 as pkg
    from sdpcutsel_via_nn_b200.cut_select_qcqp import two_pattern_covers
    rng = np.random.default_rng(40 + dim)
    n = 34
    obj = np.triu(rng.random((n, n)) < 0.35, 1)
    cons = obj | np.triu(rng.random((n, n)) < 0.25, 1)
    adj_obj, adj_cons = (obj | obj.T).astype(np.uint8), (cons | cons.T).astype(np.uint8)
    Q_arr = rng.integers(-9, 10, n * (n + 1) // 2).astype(np.float64)
    out = {}
    for dev in (True, False):
        cs = pkg.CutSolverQCQP()
        cs.set_instance(Q_arr, adj_obj, n, dim=dim, Q_adj_cons=adj_cons)
        diff = two_pattern_covers(cs, dim, device_algebra=dev)
        out[dev] = (cs._agg_list.idx.copy(), diff.idx.copy(), cs._agg_list, diff, cs)
    assert np.array_equal(out[True][0], out[False][0]) and np.array_equal(out[True][1], out[False][1])
    # the reference's own statement of it
    p0 = set(orc.cover_pattern_E_loops(adj_obj, dim))
    pm = orc.cover_pattern_E_loops(adj_cons, dim)
    want_int = [t for t in pm if t in p0]
    want_diff = [t for t in pm if t not in p0]
    assert out[True][2].keys() == want_int and out[True][3].keys() == want_diff
    assert len(want_int) > 0 and len(want_diff) > 0
    # the filtered covers are live device covers: selection on them works and agrees with the oracle
    cs = out[True][4]
    vv = orc.synth_point(n, seed=5)
    for agg, tuples in ((out[True][2], want_int), (out[True][3], want_diff)):
        cs._agg_list = agg
        rl = cs._sel_eigcut_by_ordering_on_measure(1, vv, 1)
        idx = np.full((len(tuples), dim), -1, dtype=np.int32)
        sizes = np.array([len(t) for t in tuples])
        for i, t in enumerate(tuples):
            idx[i, :len(t)] = t
        lam_o, _ = orc.score_cover(Q_arr, n, idx, sizes, vv, want_obj=False)
        order, score = orc.select_feas(lam_o)
        assert [tuple(e[0]) for e in rl] == [tuples[i] for i in order[:5000]]
