#!/usr/bin/env python
"""Whole-cover parity at the headline size (needs a B200 and ~20 GB of host RAM; a few minutes on 16 host cores):
the oracle port scores ALL C(125,5) = 234,531,275 subsets of BASELINE.json configs[3] (numpy LAPACK eigvalsh + the
NNs.so-exact C network), applies the reference's selection rules, and the result is compared with the GPU path:
every score (max abs difference) and the selected lists of strat 1, 2 and 4 (indices and order, bit for bit).

    python tools/full_parity_cfg4.py [n rho]      -> one JSON line (also written to gpurun_out/full_parity_n<n>_rho<rho>.json)
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHM = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"


def job(args):
    n, rho, density, r0, r1, N = args
    from oracle import cutsel_oracle as orc
    import sdpcutsel_via_nn_b200 as pkg
    Q_arr, _ = pkg.synthetic.boxqp_arrays(pkg.synthetic.instance(n, density, seed=7))
    vv = pkg.synthetic.lp_point(n, seed=8)
    blobs = {rho: pkg.nn_weights.load_packed(rho)}
    idx = pkg._capi.unrank(n, rho, np.arange(r0, r1))
    lam, obj = orc.score_cover(Q_arr, n, idx, np.full(idx.shape[0], rho), vv, blobs)
    np.memmap(os.path.join(SHM, "sdpcs_lam.f64"), dtype=np.float64, mode="r+", shape=(N,))[r0:r1] = lam
    np.memmap(os.path.join(SHM, "sdpcs_obj.f64"), dtype=np.float64, mode="r+", shape=(N,))[r0:r1] = obj
    return r1 - r0


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 125
    rho = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    density, k = 0.75, 5000
    import sdpcutsel_via_nn_b200 as pkg
    from oracle import cutsel_oracle as orc
    N = pkg._capi.binom(n, rho)
    for nm in ("lam", "obj"):
        np.memmap(os.path.join(SHM, "sdpcs_%s.f64" % nm), dtype=np.float64, mode="w+", shape=(N,)).flush()
    step = 500000
    jobs = [(n, rho, density, r0, min(N, r0 + step), N) for r0 in range(0, N, step)]
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        done = sum(pool.imap_unordered(job, jobs, chunksize=1))
    t_score = time.perf_counter() - t0
    assert done == N
    lam_o = np.array(np.memmap(os.path.join(SHM, "sdpcs_lam.f64"), dtype=np.float64, mode="r", shape=(N,)))
    obj_o = np.array(np.memmap(os.path.join(SHM, "sdpcs_obj.f64"), dtype=np.float64, mode="r", shape=(N,)))
    for nm in ("lam", "obj"):
        os.remove(os.path.join(SHM, "sdpcs_%s.f64" % nm))
    # GPU
    Q_arr, _ = pkg.synthetic.boxqp_arrays(pkg.synthetic.instance(n, density, seed=7))
    vv = pkg.synthetic.lp_point(n, seed=8)
    out = dict(n=n, rho=rho, candidates=N, k=k, oracle_cores=cores, oracle_score_seconds=t_score)
    for name, engine in (("tcgen05", pkg._capi.NN_TCGEN05), ("dmma", pkg._capi.NN_DMMA)):
        eng = pkg._capi.Engine(0)
        eng.set_params(nn_engine=engine)
        eng.set_weights(rho, pkg.nn_weights.load_packed(rho))
        eng.set_instance(n, Q_arr)
        eng.set_cover_all(rho)
        r1 = eng.select(1, vv, k)
        r2 = eng.select(2, None, k)
        t0 = time.perf_counter()
        r4 = eng.select(4, None, k)          # scores both measures: lam and obj stay resident for the comparison below
        t_gpu = time.perf_counter() - t0
        lam, obj = eng.scores()
        res = dict(gpu_select_seconds=t_gpu, max_abs_dlam=float(np.abs(lam - lam_o).max()), max_abs_dobj=float(np.abs(obj - obj_o).max()))
        del lam, obj
        if name == "tcgen05":
            # selections of the oracle (only once): top-k of the stable descending orders
            t0 = time.perf_counter()
            viol = lam_o < orc.THRES_NEG_EIGVAL
            s1 = np.where(viol, -lam_o, -np.inf)
            cand = np.argpartition(-s1, k)[:4 * k] if N > 4 * k else np.arange(N)
            thr = np.sort(s1[cand])[::-1][k - 1]
            c = np.nonzero(s1 >= thr)[0]
            o1 = c[np.lexsort((c, -s1[c]))][:k]
            cand = np.argpartition(-obj_o, k)[:4 * k]
            thr = np.sort(obj_o[cand])[::-1][k - 1]
            c = np.nonzero(obj_o >= thr)[0]
            o2 = c[np.lexsort((c, -obj_o[c]))][:k]
            ns, order, score = orc.select_comb(obj_o, lam_o, k)
            o4, sc4 = order[:k].copy(), score[:k].copy()
            del order, score
            out["oracle_select_seconds"] = time.perf_counter() - t0
            out["oracle_new_strat"] = int(ns)
        res.update(strat1_identical=bool(np.array_equal(r1["idx"], o1)), strat2_identical=bool(np.array_equal(r2["idx"], o2)),
                   strat4_identical=bool(np.array_equal(r4["idx"], o4)), strat4_new_strat=int(r4["new_strat"]),
                   strat4_max_abs_dscore=float(np.abs(r4["score"] - sc4).max()),
                   min_gap_between_selected_scores=float(np.min(-np.diff(r2["score"]))))
        out[name] = res
    line = json.dumps(out)
    print(line)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "full_parity_n%d_rho%d.json" % (n, rho)), "w").write(line + "\n")


if __name__ == "__main__":
    main()
