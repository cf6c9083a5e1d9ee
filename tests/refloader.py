"""Test infrastructure: import the UNMODIFIED reference (baseline/_ref, installed by tools/install_reference.py, or
/root/reference in the authoring container) with stand-ins for the solvers it imports at module level.

``cplex`` is replaced by a small LP model backed by scipy's HiGHS dual simplex -- enough of the CPLEX Python API for
``CutSolver.cut_select_algo`` (cut_select_qp.py:73-221) and ``CutSolverQCQP.cut_select_algo`` (cut_select_qcqp.py:16-113)
to run end to end: variables.add, linear_constraints.add, objective.set_sense, parameters.lpmethod, solve,
solution.get_objective_value / get_values.  mosek / cvxopt / chompack / lxml are stubs.
"""
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_dir():
    for d in (os.environ.get("SDPCS_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if d and os.path.exists(os.path.join(d, "cut_select_qp.py")) and os.path.exists(os.path.join(d, "neural_nets", "NNs.so")):
            return d
    return None


class SparsePair(object):
    def __init__(self, ind=None, val=None):
        self.ind, self.val = list(ind), list(val)


class _Obj(object):
    class sense(object):
        minimize, maximize = 1, -1

    def __init__(self):
        self._sense = 1

    def set_sense(self, s):
        self._sense = s


class _Vars(object):
    def __init__(self):
        self.obj, self.lb, self.ub, self.names = [], [], [], []

    def add(self, obj=None, lb=None, ub=None, names=None, **kw):
        self.obj += list(obj)
        self.lb += list(lb)
        self.ub += list(ub)
        self.names += list(names or [])


class _Rows(object):
    def __init__(self):
        self.rows = []                 # (SparsePair, rhs, sense) -- the format the golden generator reads
        self.batches = []              # number of rows of every add() call

    def add(self, lin_expr=None, rhs=None, senses=None, **kw):
        rows = list(zip(lin_expr, rhs, senses))
        self.rows.extend(rows)
        self.batches.append(len(rows))

    def add_rows(self, rows):
        self.rows.extend(rows)


class _LpMethod(object):
    class values(object):
        dual = 2

    def set(self, v):
        pass


class _Params(object):
    def __init__(self):
        self.lpmethod = _LpMethod()


class _Solution(object):
    def __init__(self):
        self.x, self.fun = None, None

    def get_objective_value(self):
        return float(self.fun)

    def get_values(self):
        return list(self.x)


class Cplex(object):
    """LP model; solve() = scipy.optimize.linprog(method='highs-ds') on everything added so far."""
    solves = 0

    def __init__(self):
        self.objective, self.variables, self.linear_constraints = _Obj(), _Vars(), _Rows()
        self.parameters, self.solution = _Params(), _Solution()
        self.points = []               # LP solutions in solve order

    def set_results_stream(self, *a):
        pass

    set_log_stream = set_warning_stream = set_error_stream = set_results_stream

    def solve(self):
        from scipy.optimize import linprog
        from scipy.sparse import csr_matrix
        nv = len(self.variables.obj)
        ub_r, ub_c, ub_v, ub_b, eq_r, eq_c, eq_v, eq_b = [], [], [], [], [], [], [], []
        for sp, rhs, sense in self.linear_constraints.rows:
            if sense == "E":
                r = len(eq_b)
                eq_r += [r] * len(sp.ind); eq_c += sp.ind; eq_v += sp.val; eq_b.append(rhs)
            else:
                sg = 1.0 if sense == "L" else -1.0
                r = len(ub_b)
                ub_r += [r] * len(sp.ind); ub_c += sp.ind; ub_v += [sg * v for v in sp.val]; ub_b.append(sg * rhs)
        A_ub = csr_matrix((ub_v, (ub_r, ub_c)), shape=(len(ub_b), nv)) if ub_b else None
        A_eq = csr_matrix((eq_v, (eq_r, eq_c)), shape=(len(eq_b), nv)) if eq_b else None
        res = linprog(np.array(self.variables.obj, dtype=float) * self.objective._sense, A_ub=A_ub,
                      b_ub=np.array(ub_b, dtype=float) if ub_b else None, A_eq=A_eq,
                      b_eq=np.array(eq_b, dtype=float) if eq_b else None,
                      bounds=list(zip(self.variables.lb, self.variables.ub)), method="highs-ds")
        if res.status != 0:
            raise RuntimeError("LP stand-in failed: " + res.message)
        Cplex.solves += 1
        self.solution.x, self.solution.fun = res.x, res.fun * self.objective._sense
        self.points.append(np.array(res.x))


_loaded = {}


def load_reference():
    """(cut_select_qp module, cut_select_qcqp module, reference dir) or None if no reference tree is around.
    NB: importing the reference makes warnings fatal (cut_select_qp.py:14); they are reset here."""
    d = reference_dir()
    if d is None:
        return None
    if d in _loaded:
        return _loaded[d]
    cplex = types.ModuleType("cplex")
    cplex.SparsePair, cplex.Cplex = SparsePair, Cplex
    sys.modules["cplex"] = cplex
    mosek, fusion = types.ModuleType("mosek"), types.ModuleType("mosek.fusion")
    for nm in ("Model", "Domain", "ObjectiveSense", "Expr"):
        setattr(fusion, nm, object)
    mosek.fusion = fusion
    sys.modules["mosek"], sys.modules["mosek.fusion"] = mosek, fusion
    cvxopt = types.ModuleType("cvxopt")

    def spmatrix(v, I, J, size):                      # cut_select_qp.py:326 -> a dense 0/1 array indexes the same way
        A = np.zeros(size)
        A[np.asarray(I, dtype=int), np.asarray(J, dtype=int)] = v
        return A

    cvxopt.spmatrix, cvxopt.amd = spmatrix, object
    sys.modules["cvxopt"] = cvxopt
    sys.modules["chompack"] = types.ModuleType("chompack")
    import xml.etree.ElementTree as ET
    sys.modules["lxml"] = types.ModuleType("lxml")
    sys.modules["lxml.etree"] = ET
    for name in ("cut_select_qp", "cut_select_qcqp"):
        sys.modules.pop(name, None)
    sys.path.insert(0, d)
    cwd = os.getcwd()
    try:
        import cut_select_qp as ref
        import cut_select_qcqp as refq
    finally:
        sys.path.remove(d)
        os.chdir(cwd)
    warnings.resetwarnings()
    _loaded[d] = (ref, refq, d)
    return _loaded[d]


class in_reference_dir(object):
    """The reference loads 'neural_nets/NNs.so' relative to the cwd (cut_select_qp.py:293)."""

    def __init__(self, d):
        self.d = d

    def __enter__(self):
        self.cwd = os.getcwd()
        os.chdir(self.d)

    def __exit__(self, *a):
        os.chdir(self.cwd)
