"""Golden data of the exact optimality measure (strat 3 / figure 8): the reference's OWN committed results
data_figures/fig8_data.csv -- per sub-problem the NN estimate and the exact SDP measure (Mosek), the two selections,
and the per-round summary (share selected by both, standard deviation of the exact selection), produced by
CutSolver._sel_eigcut_by_ordering_on_measure(strat=-1) (cut_select_qp.py:660-702) on spar020-100-1, dim 3, 10 %.
Only round 1 is reproducible without CPLEX (its LP point is the McCormick vertex stored as fig8_vars in
reference_golden.npz).  Run where /root/reference exists:   python tests/golden/make_golden_fig8.py
"""
import csv
import os

import numpy as np

REF = os.environ.get("SDPCS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    with open(os.path.join(REF, "data_figures", "fig8_data.csv")) as f:
        rows = list(csv.reader(f))
    start = [i for i, r in enumerate(rows) if r and r[0] == "cuts_round" and "cut_number" in r][0]
    summary = np.array([[float(v) for v in r] for r in rows[1:start]])
    r1 = [r for r in rows[start + 1:] if r[0] == "1"]
    out = dict(summary=summary,                                           # cuts_round, gap_closed, percent_same_sel, std_dev_exact_selection
               r1_cut_idx=np.array([int(r[1]) for r in r1], dtype=np.int32),
               r1_sel_estim=np.array([int(r[2]) for r in r1], dtype=np.int8),
               r1_sel_exact=np.array([int(r[3]) for r in r1], dtype=np.int8),
               r1_estim=np.array([float(r[4]) for r in r1]), r1_exact=np.array([float(r[5]) for r in r1]))
    path = os.path.join(HERE, "fig8_exact.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(r1), "round-1 rows;", "selected by estimate", int(out["r1_sel_estim"].sum()), "by exact", int(out["r1_sel_exact"].sum()))


if __name__ == "__main__":
    main()
