"""GPU tests of the tcgen05 int8-sliced NN engine (k_prep_i8 + k_mlp_i8) through the C ABI.

Checked against (a) oracle/nn_i8_model.py, the integer-for-integer CPU statement of the sliced arithmetic
(layer-0 pre-activations agree to the last bit or two, deeper layers to 1e-12), (b) the NNs.so-exact C oracle and
the NNs.so golden vectors of the reference (tolerance 1e-11 on the raw output; north_star asks 1e-5), and
(c) the FP64 DMMA engine on whole covers (same selection, scores within 1e-9)."""
import numpy as np
import pytest

from oracle import cutsel_oracle as orc
from oracle import nn_i8_model as m8

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    import sdpcutsel_via_nn_b200 as pkg
    return pkg._capi


def nn_inputs(rho, m, seed):
    rng = np.random.default_rng(seed)
    nin = rho * (rho + 3) // 2
    return np.concatenate([rng.uniform(0, 1, (m, rho)), rng.uniform(-1.0 / rho, 1.0 / rho, (m, nin - rho))], axis=1)


@pytest.mark.parametrize("rho", [2, 3, 4, 5])
def test_layers_match_the_integer_model(capi, blobs, rho):
    eng = capi.Engine(0)
    eng.set_weights(rho, blobs[rho])
    x = nn_inputs(rho, 700, seed=rho)
    nhid = int(blobs[rho][1]) - 1
    for layer in range(nhid):
        zg = eng.nn_debug_layer(rho, x, layer)
        _, zm = m8.forward(blobs[rho], x, layer)
        tol = 4e-15 * np.maximum(1.0, np.abs(zm)) if layer == 0 else 1e-12
        assert np.all(np.abs(zg - zm) <= tol), (rho, layer, np.abs(zg - zm).max())


@pytest.mark.parametrize("rho", [2, 3, 4, 5])
@pytest.mark.parametrize("m", [1, 127, 128, 129, 383, 40000])
def test_nn_eval_ragged_sizes(capi, blobs, rho, m):
    """tile tails, odd tile counts per CTA (one lane idle) and many CTAs."""
    eng = capi.Engine(0)
    eng.set_weights(rho, blobs[rho])
    x = nn_inputs(rho, m, seed=100 + m)
    want = orc.nn_eval(blobs[rho], x)
    eng.set_params(nn_engine=capi.NN_TCGEN05)
    y = eng.nn_eval(rho, x)
    assert np.abs(y - want).max() < 1e-11
    eng.set_params(nn_engine=capi.NN_DMMA)
    y2 = eng.nn_eval(rho, x)
    assert np.abs(y - y2).max() < 1e-11
    assert eng.timings()["nn_fallbacks"] == 0


@pytest.mark.parametrize("d", [2, 3, 4, 5])
def test_golden_NNs_so_vectors(capi, blobs, golden, d):
    eng = capi.Engine(0)
    eng.set_weights(d, blobs[d])
    eng.set_params(nn_engine=capi.NN_TCGEN05)
    y = eng.nn_eval(d, golden["nn%d_in" % d])
    assert np.abs(y - golden["nn%d_out" % d]).max() < 1e-11


def test_saturated_and_extreme_inputs(capi, blobs):
    """x = 0 / 1 exactly and Q~ = +-1/rho hit the ends of the fixed-point range (|p| = 1, tansig saturation)."""
    rho = 5
    eng = capi.Engine(0)
    eng.set_weights(rho, blobs[rho])
    nin = 20
    rows = [np.zeros(nin), np.ones(nin) / rho, -np.ones(nin) / rho]
    r = np.ones(nin) / rho
    r[:rho] = 1.0
    rows.append(r)
    r = -np.ones(nin) / rho
    r[:rho] = 0.0
    rows.append(r)
    x = np.array(rows)
    y = eng.nn_eval(rho, x)
    assert np.abs(y - orc.nn_eval(blobs[rho], x)).max() < 1e-11
    assert eng.timings()["nn_fallbacks"] == 0


def test_out_of_range_inputs_fall_back_to_dmma(capi, blobs):
    """mapminmax'ed inputs outside (-2, 2) cannot be sliced: the call re-scores with the FP64 DMMA engine."""
    rho = 3
    eng = capi.Engine(0)
    eng.set_weights(rho, blobs[rho])
    x = nn_inputs(rho, 300, seed=9)
    x[17, 1] = 40.0
    y = eng.nn_eval(rho, x)
    assert np.abs(y - orc.nn_eval(blobs[rho], x)).max() < 1e-10
    assert eng.timings()["nn_fallbacks"] == 1


def test_weights_outside_the_tansig_range_are_served_by_dmma(capi, blobs):
    """sdpcs_set_weights bounds the scaled pre-activations of every layer from the weights; the tcgen05 engines' tansig has no
    clamp, so a net that could reach |z| >= 960 (here: first-layer weights x 2000) is evaluated by the FP64 DMMA engine --
    silently, correctly and without counting a fall-back."""
    rho = 3
    blob = np.array(blobs[rho], dtype=np.float64).copy()
    n_in, h = int(blob[0]), int(blob[2])
    off = 3 + 2 * n_in
    blob[off:off + h * n_in] *= 2000.0                           # W_1
    eng = capi.Engine(0)
    eng.set_weights(rho, blob)
    x = nn_inputs(rho, 500, seed=21)
    y = eng.nn_eval(rho, x)
    assert np.abs(y - orc.nn_eval(blob, x)).max() < 1e-9
    assert eng.timings()["nn_fallbacks"] == 0
    eng.set_params(nn_engine=capi.NN_DMMA)
    assert np.array_equal(y, eng.nn_eval(rho, x))                # bit for bit the DMMA engine's answer


@pytest.mark.parametrize("rho", [3, 4, 5])
def test_cover_scores_both_engines(capi, blobs, rho):
    """whole all-subsets cover: tcgen05 and DMMA engines give the same combined selection and scores within 1e-9."""
    n = 30 if rho < 5 else 24
    Q_arr, adj = orc.boxqp_arrays(orc.synth_instance(n, 0.7, seed=21))
    vv = orc.synth_point(n, seed=5)
    out = {}
    for name, engine in (("i8", capi.NN_TCGEN05), ("dmma", capi.NN_DMMA)):
        eng = capi.Engine(0)
        eng.set_params(nn_engine=engine)
        eng.set_weights(rho, blobs[rho])
        eng.set_instance(n, Q_arr)
        eng.set_cover_all(rho)
        res = eng.select(4, vv, 400)
        lam, obj = eng.scores()
        out[name] = (res, lam, obj)
        assert eng.timings()["nn_fallbacks"] == 0
    idx = orc.cover_all(n, rho)
    lam_o, obj_o = orc.score_cover(Q_arr, n, idx, np.full(idx.shape[0], rho), vv, {rho: blobs[rho]})
    assert np.abs(out["i8"][2] - obj_o).max() < 1e-9
    assert np.abs(out["i8"][2] - out["dmma"][2]).max() < 1e-9
    assert np.array_equal(out["i8"][0]["idx"], out["dmma"][0]["idx"])
    ns, order, score = orc.select_comb(obj_o, lam_o, 400)
    assert np.array_equal(out["i8"][0]["idx"], order[:400])
    assert out["i8"][0]["new_strat"] == ns


def test_sharded_rank_ranges_are_consistent(capi, blobs):
    """a shard starting in the middle of the rank space (chunk offsets, unranking from rank_begin + c0)."""
    n, rho = 26, 5
    Q_arr, adj = orc.boxqp_arrays(orc.synth_instance(n, 0.75, seed=3))
    vv = orc.synth_point(n, seed=4)
    eng = capi.Engine(0)
    eng.set_weights(rho, blobs[rho])
    eng.set_instance(n, Q_arr)
    eng.set_cover_all(rho)
    eng.score(vv, 2)
    _, full = eng.scores(lam=False)
    N = full.size
    r0, r1 = N // 3 + 5, 2 * N // 3 + 77
    eng.set_cover_all(rho, r0, r1)
    eng.score(vv, 2)
    _, part = eng.scores(lam=False)
    assert np.array_equal(part, full[r0:r1])


@pytest.mark.parametrize("rho", [2, 3, 4, 5])
def test_fused_producer_is_bit_identical_to_the_staged_images(capi, blobs, golden, rho):
    """nn_fused_prep = 1: the producer warps of k_mlp_i8 unrank / gather / slice in shared memory; same arithmetic as
    k_prep_i8, so the scores are bit-identical -- all-subsets shards (ragged tile tails, mid-range start) and a
    mixed-size pattern-E list cover."""
    n = 31
    Q_arr, adj = orc.boxqp_arrays(orc.synth_instance(n, 0.7, seed=40 + rho))
    vv = orc.synth_point(n, seed=6)
    out = []
    for fused in (0, 1):
        eng = capi.Engine(0)
        eng.set_params(nn_fused_prep=fused)
        for d in range(2, rho + 1):
            eng.set_weights(d, blobs[d])
        eng.set_instance(n, Q_arr)
        res = []
        N = capi.binom(n, rho)
        for r0, r1 in ((0, N), (N // 3 + 5, 2 * N // 3 + 77), (7, 8)):
            eng.set_cover_all(rho, r0, r1)
            eng.score(vv, 2)
            res.append(eng.scores(lam=False)[1])
        assert eng.timings()["nn_fallbacks"] == 0
        out.append(res)
    for a, b in zip(*out):
        assert np.array_equal(a, b)
    if rho >= 3:
        n2, Q2, adj2 = __import__("conftest").inst_arrays(golden, "spar040-030-1")
        idx, sizes = orc.cover_pattern_E(adj2, rho)
        got = []
        for fused in (0, 1):
            eng = capi.Engine(0)
            eng.set_params(nn_fused_prep=fused)
            for d in range(2, rho + 1):
                eng.set_weights(d, blobs[d])
            eng.set_instance(n2, Q2)
            eng.set_cover_list(rho, idx)
            eng.score(golden["mix_vars"], 2)
            got.append(eng.scores(lam=False)[1])
        assert np.array_equal(got[0], got[1])


@pytest.mark.parametrize("rho", [2, 3, 4, 5])
@pytest.mark.parametrize("engine", ["exact", "screen"])
def test_eigenvalue_score_from_the_staging_kernel_is_bit_identical(capi, blobs, golden, rho, engine):
    """A call that wants both scores (want = 3, the combined rule) gets lam_min from k_prep_i8<.., FEAS> -- the arithmetic of
    k_score_feas on the point the staging kernel has gathered anyway.  nn_fused_prep = 2 keeps the two launches apart: lam
    and obj must agree bit for bit, on ragged all-subsets shards and on a mixed-size list cover, and with the scores of
    separate want = 1 / want = 2 calls."""
    n = 29
    Q_arr, adj = orc.boxqp_arrays(orc.synth_instance(n, 0.7, seed=50 + rho))
    vv = orc.synth_point(n, seed=9)
    N = capi.binom(n, rho)
    out = []
    for mode in (0, 2):
        eng = capi.Engine(0)
        eng.set_params(nn_fused_prep=mode, nn_engine=capi.NN_SCREEN if engine == "screen" else capi.NN_TCGEN05)
        for d in range(2, rho + 1):
            eng.set_weights(d, blobs[d])
        eng.set_instance(n, Q_arr)
        res = []
        for r0, r1 in ((0, N), (N // 3 + 5, 2 * N // 3 + 77), (7, 8)):
            eng.set_cover_all(rho, r0, r1)
            eng.score(vv, 3)
            launches = eng.timings()["score_launches"]
            res.append(eng.scores())
            if mode == 0:
                eng.score(vv, 1)
                lam1 = eng.scores(obj=False)[0]
                eng.score(vv, 2)
                obj2 = eng.scores(lam=False)[1]
                assert np.array_equal(lam1, res[-1][0]) and np.array_equal(obj2, res[-1][1])
        assert launches == (2 if mode == 0 else 3)
        if rho >= 3:
            n2, Q2, adj2 = __import__("conftest").inst_arrays(golden, "spar040-030-1")
            idx, sizes = orc.cover_pattern_E(adj2, rho)
            eng2 = capi.Engine(0)
            eng2.set_params(nn_fused_prep=mode, nn_engine=capi.NN_SCREEN if engine == "screen" else capi.NN_TCGEN05)
            for d in range(2, rho + 1):
                eng2.set_weights(d, blobs[d])
            eng2.set_instance(n2, Q2)
            eng2.set_cover_list(rho, idx)
            eng2.score(golden["mix_vars"], 3)
            res.append(eng2.scores())
            assert eng2.timings()["nn_fallbacks"] == 0
        assert eng.timings()["nn_fallbacks"] == 0
        out.append(res)
    for (la, oa), (lb, ob) in zip(*out):
        assert np.array_equal(la, lb) and np.array_equal(oa, ob)
        assert np.isfinite(la).all() and np.isfinite(oa).all()


@pytest.mark.parametrize("rho", [2, 3, 4, 5])
def test_screening_engine_matches_its_integer_model(capi, blobs, rho):
    """SDPCS_NN_SCREEN: the same pipeline with 4-digit operands and two TMEM accumulator stages.  Layer by layer against the
    integer model with ns = 4 (exact integer sums: layer 0 agrees to the last bits), outputs within 1e-5 of NNs.so's
    arithmetic (north_star's NN tolerance; observed ~3e-6), ragged sizes included."""
    eng = capi.Engine(0)
    eng.set_weights(rho, blobs[rho])
    eng.set_params(nn_engine=capi.NN_SCREEN)
    x = nn_inputs(rho, 700, seed=40 + rho)
    nhid = int(blobs[rho][1]) - 1
    for layer in range(nhid):
        zg = eng.nn_debug_layer(rho, x, layer)
        _, zm = m8.forward(blobs[rho], x, layer, ns=4)
        tol = 4e-15 * np.maximum(1.0, np.abs(zm)) if layer == 0 else 1e-6
        assert np.all(np.abs(zg - zm) <= tol), (rho, layer, np.abs(zg - zm).max())
    for m in (1, 129, 383, 40000):
        x = nn_inputs(rho, m, seed=300 + m)
        y = eng.nn_eval(rho, x)
        want = orc.nn_eval(blobs[rho], x)
        assert np.abs(y - want).max() < 1e-5
        ym, _ = m8.forward(blobs[rho], x[:2000], ns=4)
        assert np.abs(y[:2000] - ym).max() < 1e-6
    assert eng.timings()["nn_fallbacks"] == 0
