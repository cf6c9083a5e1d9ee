"""Small-N latency probe: one separation round of BASELINE configs[1] (spar125-075-1, P^E_3, N = 133,242 + triangles)
through the drop-in CutSolver surface, per phase, wall clock (the span the reference times as sep_times,
cut_select_qp.py:162-187).  Builder tool; bench.py --workload cfg2 is the driver-visible version."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdpcutsel_via_nn_b200 as pkg  # noqa: E402


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    Qf = g["inst_spar125_075_1_Q"].astype(np.float64)
    Q_arr, adj = pkg.synthetic.boxqp_arrays(Qf)
    n = Qf.shape[0]
    vv = g["cfg2_vars"]
    cs = pkg.CutSolver()
    cs.set_instance(Q_arr, adj, n, dim=3)
    cs._load_neural_nets()
    t = time.perf_counter()
    N = cs._get_sdp_vertex_cover(3)
    t_cover = time.perf_counter() - t
    k = min(int(np.floor(0.1 * N)), 5000)
    cs._CutSolver__preprocess_triangle_ineq()
    out = dict(N=N, k=k, cover_ms=t_cover * 1e3)
    for strat in (1, 2, 4):
        ph = {}
        for rep in range(6):
            cs._my_prob.linear_constraints.rows = []
            t0 = time.perf_counter()
            r = cs._sel_eigcut_by_ordering_on_measure(strat, vv, 1, sel_size=k)
            rl = r[1] if strat == 4 else r
            t1 = time.perf_counter()
            cs._gen_eigcuts_selected(strat, k, rl, vars_values=vv)
            t2 = time.perf_counter()
            cs._CutSolver__separate_and_add_triangle(0.1, vv)
            t3 = time.perf_counter()
            if rep:
                for nm, d in (("select", t1 - t0), ("gen_cuts", t2 - t1), ("triangles", t3 - t2), ("round", t3 - t0)):
                    ph.setdefault(nm, []).append(d * 1e3)
        out["strat%d_ms" % strat] = {nm: float(np.median(v)) for nm, v in ph.items()}
        eng = cs._engine_for(cs._agg_list)
        out["strat%d_device" % strat] = eng.timings()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
