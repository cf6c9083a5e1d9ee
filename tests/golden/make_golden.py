"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):   python tests/golden/make_golden.py
The reference is imported with stub cplex/mosek/cvxopt/chompack/lxml modules; its
hot-path methods (_get_sdp_vertex_cover, _sel_eigcut_by_ordering_on_measure, _gen_eigcuts_selected,
the two triangle methods) and neural_nets/NNs.so run unchanged.  Nothing here is imported by the product.
"""
import csv
import ctypes
import os
import sys
import types
import warnings

import numpy as np

REF = os.environ.get("SDPCS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference():
    cplex = types.ModuleType("cplex")

    class SparsePair:
        def __init__(self, ind=None, val=None):
            self.ind, self.val = list(ind), list(val)

    class _LC:
        def __init__(self):
            self.rows = []

        def add(self, lin_expr=None, rhs=None, senses=None, **k):
            self.rows.extend(zip(lin_expr, rhs, senses))

    class FakeCplex:
        def __init__(self):
            self.linear_constraints = _LC()

    cplex.SparsePair, cplex.Cplex = SparsePair, FakeCplex
    sys.modules["cplex"] = cplex
    mosek = types.ModuleType("mosek")
    fusion = types.ModuleType("mosek.fusion")
    for nm in ("Model", "Domain", "ObjectiveSense", "Expr"):
        setattr(fusion, nm, object)
    mosek.fusion = fusion
    sys.modules["mosek"], sys.modules["mosek.fusion"] = mosek, fusion
    cvxopt = types.ModuleType("cvxopt")
    cvxopt.spmatrix = cvxopt.amd = object
    sys.modules["cvxopt"] = cvxopt
    sys.modules["chompack"] = types.ModuleType("chompack")
    import xml.etree.ElementTree as ET
    sys.modules["lxml"] = types.ModuleType("lxml")
    sys.modules["lxml.etree"] = ET
    sys.path.insert(0, REF)
    os.chdir(REF)  # NNs.so path is cwd-relative (cut_select_qp.py:293)
    import cut_select_qp as ref
    warnings.resetwarnings()
    return ref, FakeCplex


def read_boxqp(name):
    with open(os.path.join(REF, "boxqp_instances", name + ".in")) as f:
        content = f.readlines()
    n = int(content[0].split()[0])
    c = np.array([int(v) for v in content[1].split()])
    Qf = np.array([[float(v) for v in row.split()] for row in content[2:2 + n]])
    return n, c, Qf


def make_solver(ref, FakeCplex, Qf, dim):
    """State exactly as __parse_boxqp_into_cplex leaves it (cut_select_qp.py:313-328), Q_adj as numpy."""
    n = Qf.shape[0]
    cs = ref.CutSolver()
    Q = -Qf.copy()
    Q_arr = Q.copy()
    for i in range(n):
        Q_arr[i, i] /= 2.0
    Q_arr = Q_arr[np.triu_indices(n, k=0)]
    cs._Q = Q / 2
    cs._Q_adj = (Qf != 0).astype(float)
    cs._Q_arr = Q_arr
    cs._nb_vars, cs._nb_lifted = n, n * (n + 1) // 2
    cs._dim = dim
    cs._my_prob = FakeCplex()
    cs._load_neural_nets()
    return cs


def cover_arrays(agg_list, dim):
    N = len(agg_list)
    idx = np.full((N, dim), -1, dtype=np.int16)
    for i, a in enumerate(agg_list):
        idx[i, :len(a[0])] = a[0]
    return idx


def rows_to_arrays(rows, width):
    m = len(rows)
    ind = np.full((m, width), -1, dtype=np.int32)
    val = np.zeros((m, width))
    rhs = np.zeros(m)
    for i, (sp, r, s) in enumerate(rows):
        ind[i, :len(sp.ind)] = sp.ind
        val[i, :len(sp.val)] = sp.val
        rhs[i] = r
    return ind, val, rhs


def mccormick_lp_point(Qf, c):
    """Round-0 LP of cut_select_algo (cut_select_qp.py:330-375) solved with HiGHS dual simplex."""
    from scipy.optimize import linprog
    from scipy.sparse import lil_matrix
    n = Qf.shape[0]
    nl = n * (n + 1) // 2
    Q = -Qf.copy()
    Qa = Q.copy()
    Qa[np.arange(n), np.arange(n)] /= 2
    obj = np.concatenate([Qa[np.triu_indices(n)], -c.astype(float)])
    rows, rhs = [], []
    for i in range(n):
        iXii, ixi = n * i - i * (i - 1) // 2, nl + i
        rows.append({iXii: 1, ixi: -1}); rhs.append(0)
        rows.append({iXii: -1, ixi: 2}); rhs.append(1)
        for j in range(i + 1, n):
            iXij, ixj = iXii + j - i, ixi + j - i
            if Qf[i, j] != 0:
                rows.append({iXij: -1, ixi: 1, ixj: 1}); rhs.append(1)
                rows.append({iXij: 1, ixi: -1}); rhs.append(0)
                rows.append({iXij: 1, ixj: -1}); rhs.append(0)
    A = lil_matrix((len(rows), nl + n))
    for r, d in enumerate(rows):
        for k, v in d.items():
            A[r, k] = v
    res = linprog(obj, A_ub=A.tocsr(), b_ub=np.array(rhs, float), bounds=(0, 1), method="highs-ds")
    assert res.status == 0
    return res.x


def main():
    from oracle import cutsel_oracle as orc  # only for the synthetic-point recipe (shared with tests)
    ref, FakeCplex = import_reference()
    out = {}

    # ---- NN known-answer vectors straight from NNs.so ------------------------------------------
    lib = ctypes.cdll.LoadLibrary(os.path.join(REF, "neural_nets", "NNs.so"))
    rng = np.random.default_rng(2024)
    for d in (2, 3, 4, 5):
        f = getattr(lib, "neural_net_%dD" % d)
        f.restype = ctypes.c_double
        nin = d * (d + 3) // 2
        X = np.concatenate([rng.uniform(0, 1, (256, d)), rng.uniform(-1.0 / d, 1.0 / d, (256, nin - d))], axis=1)
        X[:8] = rng.uniform(-3, 3, (8, nin))      # a few out-of-domain rows (saturated tansig)
        y = np.array([f((ctypes.c_double * nin)(*row)) for row in X])
        out["nn%d_in" % d], out["nn%d_out" % d] = X, y

    # ---- instances -------------------------------------------------------------------------------
    names = ["spar020-100-1", "spar030-060-1", "spar040-030-1", "spar050-030-1", "spar125-075-1"]
    inst = {}
    for nm in names:
        n, c, Qf = read_boxqp(nm)
        inst[nm] = (n, c, Qf)
        key = nm.replace("-", "_")
        out["inst_%s_Q" % key] = Qf.astype(np.int8)
        out["inst_%s_c" % key] = c.astype(np.int8)
        assert np.array_equal(out["inst_%s_Q" % key].astype(float), Qf)

    # ---- covers: pattern-E set/order and aggregation for three small instances, rho = 3,4,5 -------
    for nm in ["spar030-060-1", "spar040-030-1", "spar050-030-1"]:
        n, c, Qf = inst[nm]
        key = nm.replace("-", "_")
        for dim in (3, 4, 5):
            cs = make_solver(ref, FakeCplex, Qf, dim)
            N = cs._get_sdp_vertex_cover(dim)
            out["cover_%s_d%d" % (key, dim)] = cover_arrays(cs._agg_list, dim)
            if nm == "spar030-060-1":
                t = dim * (dim + 1) // 2
                qs = np.full((N, t), np.nan)
                me = np.empty(N)
                for i, a in enumerate(cs._agg_list):
                    qs[i, :len(a[2])] = a[2]
                    me[i] = a[3]
                out["agg_%s_d%d_Qslice" % (key, dim)] = qs
                out["agg_%s_d%d_maxelem" % (key, dim)] = me

    # ---- config 1: spar030-060-1, all C(30,3), strat 1, 10 % ---------------------------------------
    n, c, Qf = inst["spar030-060-1"]
    cs = make_solver(ref, FakeCplex, Qf, 3)
    N = cs._get_sdp_vertex_cover(3, ch_ext=-1)
    assert N == 4060
    vv = orc.synth_point(n, seed=8)
    out["cfg1_vars"] = vv
    k = min(int(np.floor(0.1 * N)), 5000)
    for strat in (1, 2, 4):
        cs._my_prob = FakeCplex()
        rl = cs._sel_eigcut_by_ordering_on_measure(strat, vv, 1, sel_size=k)
        if strat == 4:
            out["cfg1_s4_newstrat"] = np.array(rl[0])
            rl = rl[1]
        if strat == 1:
            out["cfg1_s1_sets"] = np.array([e[0] for e in rl], dtype=np.int16)
            out["cfg1_s1_score"] = np.array([e[1] for e in rl])
        else:
            out["cfg1_s%d_idx" % strat] = np.array([e[0] for e in rl], dtype=np.int32)
            out["cfg1_s%d_score" % strat] = np.array([e[1] for e in rl])
        nb = cs._gen_eigcuts_selected(strat, k, rl, vars_values=vv)
        ind, val, rhs = rows_to_arrays(cs._my_prob.linear_constraints.rows, 9)
        out["cfg1_s%d_cut_ind" % strat], out["cfg1_s%d_cut_val" % strat], out["cfg1_s%d_cut_rhs" % strat] = ind, val, rhs
        assert nb == len(rhs)
    # degenerate vertex point, scores only
    vd = orc.degenerate_point(n, cs._Q_arr)
    rl = cs._sel_eigcut_by_ordering_on_measure(1, vd, 1)
    out["cfg1_deg_s1_score"] = np.array([e[1] for e in rl])
    out["cfg1_deg_s1_sets"] = np.array([e[0] for e in rl], dtype=np.int16)

    # ---- pattern-E selection on mixed-size covers: spar040-030-1, rho 3,4,5, strat 1,2,4 -----------
    n, c, Qf = inst["spar040-030-1"]
    vv = orc.synth_point(n, seed=9)
    out["mix_vars"] = vv
    for dim in (3, 4, 5):
        cs = make_solver(ref, FakeCplex, Qf, dim)
        N = cs._get_sdp_vertex_cover(dim)
        k = max(1, min(int(np.floor(0.1 * N)), 5000))
        for strat in (1, 2, 4):
            cs._my_prob = FakeCplex()
            rl = cs._sel_eigcut_by_ordering_on_measure(strat, vv, 1, sel_size=k)
            if strat == 4:
                out["mix_d%d_s4_newstrat" % dim] = np.array(rl[0])
                rl = rl[1]
            if strat == 1:
                sets = np.full((len(rl), dim), -1, dtype=np.int16)
                for i, e in enumerate(rl):
                    sets[i, :len(e[0])] = e[0]
                out["mix_d%d_s1_sets" % dim] = sets
            else:
                out["mix_d%d_s%d_idx" % (dim, strat)] = np.array([e[0] for e in rl], dtype=np.int32)
            out["mix_d%d_s%d_score" % (dim, strat)] = np.array([e[1] for e in rl])
            nb = cs._gen_eigcuts_selected(strat, k, rl, vars_values=vv)
            ind, val, rhs = rows_to_arrays(cs._my_prob.linear_constraints.rows, dim + dim * (dim + 1) // 2)
            out["mix_d%d_s%d_cut_ind" % (dim, strat)] = ind
            out["mix_d%d_s%d_cut_val" % (dim, strat)] = val
            out["mix_d%d_s%d_cut_rhs" % (dim, strat)] = rhs

    # ---- config 2: spar125-075-1, P^E_3 (N = 133,242) + triangles, strat 1, 2, 4, k = 5000 ----------
    n, c, Qf = inst["spar125-075-1"]
    cs = make_solver(ref, FakeCplex, Qf, 3)
    N = cs._get_sdp_vertex_cover(3)
    assert N == 133242, N
    vv = orc.synth_point(n, seed=10)
    out["cfg2_vars"] = vv
    k = min(int(np.floor(0.1 * N)), 5000)
    for strat in (1, 2, 4):
        cs._my_prob = FakeCplex()
        rl = cs._sel_eigcut_by_ordering_on_measure(strat, vv, 1, sel_size=k)
        if strat == 4:
            out["cfg2_s4_newstrat"] = np.array(rl[0])
            rl = rl[1]
        out["cfg2_s%d_len" % strat] = np.array(len(rl))
        top = rl[:k]
        if strat == 1:
            out["cfg2_s1_sets"] = np.array([list(e[0]) + [-1] * (3 - len(e[0])) for e in top], dtype=np.int16)
        else:
            out["cfg2_s%d_idx" % strat] = np.array([e[0] for e in top], dtype=np.int32)
        out["cfg2_s%d_score" % strat] = np.array([e[1] for e in top])
        nb = cs._gen_eigcuts_selected(strat, k, rl, vars_values=vv)
        out["cfg2_s%d_nbcuts" % strat] = np.array(nb)
        ind, val, rhs = rows_to_arrays(cs._my_prob.linear_constraints.rows[:64], 9)
        out["cfg2_s%d_cut_ind" % strat], out["cfg2_s%d_cut_val" % strat], out["cfg2_s%d_cut_rhs" % strat] = ind, val, rhs
    cs._my_prob = FakeCplex()
    cs._CutSolver__preprocess_triangle_ineq()
    out["cfg2_tri_ntriples"] = np.array(len(cs._idx_list_tri))
    nb = cs._CutSolver__separate_and_add_triangle(0.1, vv)
    rows = cs._my_prob.linear_constraints.rows
    ind, val, rhs = rows_to_arrays(rows, 6)
    out["cfg2_tri_ind"], out["cfg2_tri_val"], out["cfg2_tri_rhs"] = ind.astype(np.int32), val.astype(np.int8), rhs.astype(np.int8)
    assert nb == len(rows)

    # ---- Fig 8 golden: round-1 rows of data_figures/fig8_data.csv + the LP point that reproduces them
    n, c, Qf = inst["spar020-100-1"]
    vlp = mccormick_lp_point(Qf, c)
    out["fig8_vars"] = vlp
    with open(os.path.join(REF, "data_figures", "fig8_data.csv")) as f:
        rows = list(csv.reader(f))
    start = [i for i, r in enumerate(rows) if r and r[0] == "cuts_round" and "cut_number" in r][0]
    r1 = [r for r in rows[start + 1:] if r[0] == "1"]
    out["fig8_r1_cut_idx"] = np.array([int(r[1]) for r in r1], dtype=np.int32)
    out["fig8_r1_estim"] = np.array([float(r[4]) for r in r1])
    cs = make_solver(ref, FakeCplex, Qf, 3)
    N = cs._get_sdp_vertex_cover(3)
    rl = cs._sel_eigcut_by_ordering_on_measure(2, vlp, 1)
    got_idx = np.array([e[0] for e in rl])
    got = np.array([e[1] for e in rl])
    print("fig8: N", N, "rows", len(r1), "order equal", np.array_equal(got_idx, out["fig8_r1_cut_idx"]),
          "max abs diff", np.abs(got - out["fig8_r1_estim"]).max())

    # ---- config 5: QCQP q_20_20_100_1 caller pattern (cut_select_qcqp.py:36-98) ----------------------
    import cut_select_qcqp as refq
    csq = refq.CutSolverQCQP()
    # the OSiL parser builds a CPLEX model; feed it a recording stub
    cplex = sys.modules["cplex"]

    class _Rec:
        def __getattr__(self, k):
            return _Rec()

        def __call__(self, *a, **k):
            return _Rec()

    class FakeCplexQ(FakeCplex):
        def __init__(self):
            super().__init__()
            self.objective = _Rec()
            self.variables = _Rec()
            self.parameters = _Rec()

        def set_results_stream(self, *a):
            pass

    cplex.Cplex = FakeCplexQ
    try:
        csq._dim = 3
        csq._CutSolverQCQP__parse_qcqp_osil_into_cplex("q_20_20_100_1")
        n = csq._nb_vars
        out["qcqp_Q_arr"] = np.array(csq._Q_arr, dtype=np.float64)
        out["qcqp_adj"] = np.array(csq._Q_adj, dtype=np.uint8)
        out["qcqp_adj_cons"] = np.array(csq._Q_adj_cons, dtype=np.uint8)
        vq = orc.synth_point(n, seed=11)
        out["qcqp_vars"] = vq
        csq._load_neural_nets.__func__  # noqa
        for dim in (3, 4, 5):
            csq._dim = dim
            csq._load_neural_nets()
            agg_cons = csq._CutSolverQCQP__get_vertex_cover(dim)
            agg = csq._agg_list[:]
            N = len(agg)
            out["qcqp_d%d_N" % dim] = np.array([N, len(agg_cons)])
            k = max(1, min(int(np.floor(0.1 * N)), 5000))
            for strat in (1, 4):
                csq._my_prob = FakeCplex()
                if strat == 4:
                    ns, rl_obj = csq._sel_eigcut_by_ordering_on_measure(4, vq, 1, sel_size=k)
                    out["qcqp_d%d_s4_newstrat" % dim] = np.array(ns)
                else:
                    rl_obj = csq._sel_eigcut_by_ordering_on_measure(1, vq, 1)
                csq._agg_list = agg_cons
                rl_cons = csq._sel_eigcut_by_ordering_on_measure(1, vq, 1)
                csq._agg_list = agg
                rl = (rl_obj + rl_cons)[0:k]
                if strat == 1:
                    out["qcqp_d%d_s1_sets" % dim] = np.array([e[0] for e in rl], dtype=np.int16)
                else:
                    out["qcqp_d%d_s4_idx" % dim] = np.array([e[0] for e in rl], dtype=np.int32)
                out["qcqp_d%d_s%d_score" % (dim, strat)] = np.array([e[1] for e in rl])
                nb = csq._gen_eigcuts_selected(strat, k, rl, vars_values=vq)
                ind, val, rhs = rows_to_arrays(csq._my_prob.linear_constraints.rows, dim + dim * (dim + 1) // 2)
                out["qcqp_d%d_s%d_cut_ind" % (dim, strat)] = ind
                out["qcqp_d%d_s%d_cut_val" % (dim, strat)] = val
                out["qcqp_d%d_s%d_cut_rhs" % (dim, strat)] = rhs
        print("qcqp ok: n", n, {d: out["qcqp_d%d_N" % d].tolist() for d in (3, 4, 5)})
    except Exception as e:  # parser needs lxml-specific behaviour
        print("QCQP golden skipped:", repr(e))
        raise

    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
