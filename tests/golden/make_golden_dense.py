"""Golden fixtures for the dense eigenvalue cuts (strat 0) from the UNMODIFIED reference:
CutSolver.__gen_dense_eigcuts (cut_select_qp.py:757-786).

Run in the authoring container only (needs /root/reference):   python tests/golden/make_golden_dense.py
Writes tests/golden/reference_dense_eigcuts.npz; uses the stubs of make_golden.py.  Nothing here is imported by the product.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

ROOT = mg.ROOT


def main():
    ref, FakeCplex = mg.import_reference()
    from oracle import cutsel_oracle as orc
    out = {}
    for name, seed in (("spar030-060-1", 8), ("spar040-030-1", 9), ("spar125-075-1", 10)):
        n, c, Qf = mg.read_boxqp(name)
        cs = mg.make_solver(ref, FakeCplex, Qf, 3)
        vv = orc.synth_point(n, seed=seed)
        cs._my_prob = FakeCplex()
        nb = cs._CutSolver__gen_dense_eigcuts(vars_values=vv)
        rows = cs._my_prob.linear_constraints.rows
        assert nb == len(rows) and all(s == "G" for _, _, s in rows)
        key = name.replace("-", "_")
        out["dense_%s_vars" % key] = vv
        out["dense_%s_ind" % key] = np.array(rows[0][0].ind, dtype=np.int64) if rows else np.zeros(0, np.int64)
        keep = rows if n <= 40 else rows[:8]              # full rows are n + n(n+1)/2 wide: keep 8 for the large instance
        out["dense_%s_val" % key] = np.array([sp.val for sp, _, _ in keep])
        out["dense_%s_rhs" % key] = np.array([r for _, r, _ in rows])
        out["dense_%s_nb" % key] = np.array(nb)
        print(name, "n", n, "cuts", nb)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_dense_eigcuts.npz"), **out)


if __name__ == "__main__":
    main()
