"""GPU tests of the dense eigenvalue cuts (strat 0, sdpcs_dense_eigcuts, dense_kernels.cuh) through the C ABI and the
drop-in method, against rows recorded from the unmodified reference (tests/golden/reference_dense_eigcuts.npz,
CutSolver.__gen_dense_eigcuts, cut_select_qp.py:757-786) and the oracle's restatement.

Tolerances: eigenvalues 1e-12 absolute (north_star: 1e-9 relative); coefficients of cuts from simple eigenvalues
1e-9 (products v_i v_j are invariant to the sign of v); every cut is also checked through what makes it a cut:
<v v^T, M> = lam < 0 at the LP point, to 1e-10."""
import os

import numpy as np
import pytest

import sdpcutsel_via_nn_b200 as pkg
from conftest import ROOT, inst_arrays
from oracle import cutsel_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dense_golden():
    with np.load(os.path.join(ROOT, "tests", "golden", "reference_dense_eigcuts.npz")) as z:
        return {k: z[k] for k in z.files}


def _violation(n, vv, ind, val, rhs):
    """lhs - rhs of a row at the LP point = v^T M v (negative: the point violates the cut)."""
    return float(np.dot(val, np.asarray(vv)[ind]) - rhs)


@pytest.mark.parametrize("name", ["spar030-060-1", "spar040-030-1", "spar125-075-1"])
def test_reference_rows(golden, dense_golden, name):
    n, Q_arr, adj = inst_arrays(golden, name)
    key = name.replace("-", "_")
    vv = dense_golden["dense_%s_vars" % key]
    eng = pkg._capi.Engine(0)
    eng.set_instance(n, Q_arr)
    d = eng.dense_eigcuts(vv)
    ind_o, val_o, rhs_o, w_o = orc.dense_eigcuts(n, vv)
    assert np.abs(d["eigvals"] - w_o).max() < 1e-12
    assert d["val"].shape[0] == int(dense_golden["dense_%s_nb" % key]) == rhs_o.size
    assert np.array_equal(d["ind"], dense_golden["dense_%s_ind" % key])
    assert np.abs(d["rhs"] - dense_golden["dense_%s_rhs" % key]).max() < 1e-9
    gv = dense_golden["dense_%s_val" % key]
    gaps = np.diff(w_o)
    for r in range(gv.shape[0]):
        if min(gaps[r - 1] if r else 1.0, gaps[r]) > 1e-6:           # simple eigenvalue: the row is unique
            assert np.abs(d["val"][r] - gv[r]).max() < 1e-9
    for r in range(d["val"].shape[0]):
        assert abs(_violation(n, vv, d["ind"], d["val"][r], d["rhs"][r]) - w_o[r]) < 1e-10
        assert w_o[r] < -1e-15


def test_psd_point_gives_no_cut_and_n250(golden):
    """X = x x^T makes [1 x^T; x X] PSD of rank one: eigenvalues {0 (n times), 1 + |x|^2}, no row; then the largest
    supported order (n = 250) on a random point, checked against numpy."""
    n = 250
    rng = np.random.default_rng(4)
    x = rng.uniform(0, 1, n)
    X = np.outer(x, x)
    vv = np.concatenate([X[np.triu_indices(n)], x])
    eng = pkg._capi.Engine(0)
    eng.set_instance(n, np.zeros(n * (n + 1) // 2))
    d = eng.dense_eigcuts(vv)
    assert np.abs(d["eigvals"][:-1]).max() < 1e-12 and abs(d["eigvals"][-1] - (1 + x @ x)) < 1e-10
    assert all(_violation(n, vv, d["ind"], v, r) > -1e-12 for v, r in zip(d["val"], d["rhs"]))   # rounding-level rows only
    vv = orc.synth_point(n, seed=3)
    d = eng.dense_eigcuts(vv)
    ind_o, val_o, rhs_o, w_o = orc.dense_eigcuts(n, vv)
    assert np.abs(d["eigvals"] - w_o).max() < 1e-11 and d["val"].shape[0] == rhs_o.size
    assert np.abs(d["rhs"] - rhs_o).max() < 1e-9
    for r in (0, 1, rhs_o.size // 2, rhs_o.size - 1):
        assert abs(_violation(n, vv, d["ind"], d["val"][r], d["rhs"][r]) - w_o[r]) < 1e-9


def test_repeated_eigenvalues_span_the_same_subspace():
    """x = 0, X = -I on two coordinates: eigenvalue -1 twice.  The two rows differ from LAPACK's, the cut cone is the same:
    the sum of the two v v^T is the projector on the eigenspace."""
    n = 6
    X = np.eye(n)
    X[1, 1] = X[4, 4] = -1.0
    vv = np.concatenate([X[np.triu_indices(n)], np.zeros(n)])
    eng = pkg._capi.Engine(0)
    eng.set_instance(n, np.zeros(n * (n + 1) // 2))
    d = eng.dense_eigcuts(vv)
    assert d["val"].shape[0] == 2 and np.abs(d["eigvals"][:2] + 1).max() < 1e-14
    tot = d["val"].sum(axis=0)
    nb = n * (n + 1) // 2
    want = np.zeros(n + nb)
    want[n + orc.xarr_inds(n, [1])[0]] = 1.0
    want[n + orc.xarr_inds(n, [4])[0]] = 1.0
    assert np.abs(tot - want).max() < 1e-12 and np.abs(d["rhs"]).max() < 1e-14


def test_drop_in_method_and_csr_sink(golden, dense_golden):
    n, Q_arr, adj = inst_arrays(golden, "spar030-060-1")
    vv = dense_golden["dense_spar030_060_1_vars"]
    cs = pkg.CutSolver()
    cs.set_instance(Q_arr, adj, n, dim=3)
    nb = cs._CutSolver__gen_dense_eigcuts(vars_values=vv)
    rows = cs._my_prob.linear_constraints.rows
    assert nb == len(rows) == int(dense_golden["dense_spar030_060_1_nb"])
    for (sp, rhs, sense), gval, grhs in zip(rows, dense_golden["dense_spar030_060_1_val"], dense_golden["dense_spar030_060_1_rhs"]):
        assert sense == "G" and sp.ind == dense_golden["dense_spar030_060_1_ind"].tolist()
        assert np.abs(np.array(sp.val) - gval).max() < 1e-9 and abs(rhs - grhs) < 1e-9
