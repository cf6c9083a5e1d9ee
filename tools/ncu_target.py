#!/usr/bin/env python
"""Small deterministic launch sequence for ncu captures (one chunk of the headline workload, every hot kernel once or twice):

    cfg4 instance (n = 125, rho = 5), the first 33,554,432 candidates (= one staging chunk of 262,144 tiles):
      k_score_feas<5>                      eigenvalue scores
      k_prep_i8<5,7> + k_mlp_i8<4,0,7,0>   FP64-accurate NN engine (7 digits, one TMEM stage)
      k_prep_i8<5,4> + k_mlp_i8<4,0,4,0>   screening NN engine (4 digits, two TMEM stages)
      k_sel_hist / scan / collect / rank_sort / finish / gather    one combined selection (strong-prefix path)

    python tools/ncu_target.py            # plain run first (must exit 0), then the same line under ncu
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdpcutsel_via_nn_b200 as pkg  # noqa: E402


def main():
    n, rho, k = 125, 5, 5000
    N = 262144 * 128
    Q_arr, _ = pkg.synthetic.boxqp_arrays(pkg.synthetic.instance(n, 0.75, seed=7))
    vv = pkg.synthetic.lp_point(n, seed=8)
    eng = pkg._capi.Engine(0)
    eng.set_weights(rho, pkg.nn_weights.load_packed(rho))
    eng.set_instance(n, Q_arr)
    eng.set_cover_all(rho, 0, N)
    r = eng.select(4, vv, k)                               # feas + exact NN + selection
    eng.set_params(nn_engine=pkg._capi.NN_SCREEN)
    eng.score(None, 2)                                     # screening NN
    t = eng.timings()
    print("candidates %d, selected %d, new_strat %d, last score pass %.3f ms (nn %.3f ms)" % (N, r["idx"].size, r["new_strat"], t["score_ms"], t["nn_ms"]))


if __name__ == "__main__":
    main()
