"""Where the host time of a small round goes: cProfile over the cfg2 rounds of bench.py (_dropin_rounds) with the CSR
row sink, per strategy, sorted by own time.  Builder tool (needs a GPU)."""
import cProfile
import gc
import io
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sdpcutsel_via_nn_b200 as pkg  # noqa: E402
from oracle import cutsel_oracle as orc  # noqa: E402


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    Qf = g["inst_spar125_075_1_Q"].astype(np.float64)
    Q_arr, adj = pkg.synthetic.boxqp_arrays(Qf)
    n = Qf.shape[0]
    cs = pkg.CutSolver()
    cs.set_instance(Q_arr, adj, n, dim=3)
    cs._load_neural_nets()
    N = cs._get_sdp_vertex_cover(3)
    k = min(int(np.floor(0.1 * N)), 5000)
    cs._CutSolver__preprocess_triangle_ineq()
    points = [orc.synth_point(n, seed=10 + i) for i in range(20)]
    sink = bench._CsrSink()
    cs._my_prob.linear_constraints = sink
    gc.disable()
    for strat in (1, 2, 4):
        def rounds(timed=None):
            for vv in points:
                sink.blocks = []
                t0 = time.perf_counter()
                r = cs._sel_eigcut_by_ordering_on_measure(strat, vv, 1, sel_size=k)
                t1 = time.perf_counter()
                cs._gen_eigcuts_selected(strat, k, r[1] if strat == 4 else r, vars_values=vv)
                t2 = time.perf_counter()
                cs._CutSolver__separate_and_add_triangle(0.1, vv)
                t3 = time.perf_counter()
                if timed is not None:
                    timed.append((t1 - t0, t2 - t1, t3 - t2, t3 - t0))
        rounds()
        tm = []
        rounds(tm)
        med = np.median(np.array(tm), axis=0) * 1e3
        print("strat %d: select %.3f  gen_cuts %.3f  triangles %.3f  round %.3f ms (median of 20, CSR sink)" % (strat, *med))
        pr = cProfile.Profile()
        pr.enable()
        rounds()
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
        print("\n".join(l for l in s.getvalue().splitlines() if l.strip())[:6000])
        eng = cs._engine_for(cs._agg_list)
        print("device timings of the last select:", eng.timings())


if __name__ == "__main__":
    main()
