"""Golden selections at BASELINE.json's full sizes, from the ORACLE alone (no GPU, no product code on the data path):

    configs[2]  n = 125, 75 %, rho = 4, all C(125,4) =   9,691,375 subsets
    configs[3]  n = 125, 75 %, rho = 5, all C(125,5) = 234,531,275 subsets

Every subset is scored with the oracle (numpy LAPACK eigvalsh + the NNs.so-exact C network, oracle/cutsel_oracle.py),
the reference's selection rules are applied, and the first 5000 entries of strat 1, 2 and 4 (indices, scores,
new_strat, counters) are written to tests/golden/fullsize_topk.npz.  tests/test_gpu_fullsize.py and bench.py compare
the GPU selection with these lists (bench.py prints `selection.matches_oracle_golden`).

    python tests/golden/make_golden_fullsize.py [rho ...]        # ~1 min for rho = 4, ~6 min for rho = 5 on 8 cores

Candidates are enumerated per (i1, i2) prefix with itertools.combinations -- the order of the reference's nested loops
(cut_select_qp.py:451-455) -- so that no unranking code of the product is involved.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
SHM = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
N_VARS, DENSITY, K = 125, 0.75, 5000


def comb(n, k):
    r = 1
    for j in range(k):
        r = r * (n - j) // (j + 1)
    return r if k >= 0 and n >= k else 0


def prefix_jobs(n, rho):
    """(i1, i2, rank of the first subset with that prefix, count) in enumeration order."""
    jobs, r = [], 0
    for i1 in range(n):
        for i2 in range(i1 + 1, n):
            c = comb(n - 1 - i2, rho - 2)
            if c:
                jobs.append((i1, i2, r, c))
            r += c
    assert r == comb(n, rho)
    return jobs


def job(args):
    n, rho, batch, N = args
    from oracle import cutsel_oracle as orc
    import sdpcutsel_via_nn_b200 as pkg          # weight blob + synthetic recipe only
    Q_arr, _ = orc.boxqp_arrays(orc.synth_instance(n, DENSITY, seed=7))
    vv = orc.synth_point(n, seed=8)
    blobs = {rho: pkg.nn_weights.load_packed(rho)}
    lam_m = np.memmap(os.path.join(SHM, "sdpcs_gold_lam.f64"), dtype=np.float64, mode="r+", shape=(N,))
    obj_m = np.memmap(os.path.join(SHM, "sdpcs_gold_obj.f64"), dtype=np.float64, mode="r+", shape=(N,))
    done = 0
    for i1, i2, r0, c in batch:
        idx = orc.cover_all_block(n, rho, i1, i2)
        lam, obj = orc.score_cover(Q_arr, n, idx, np.full(c, rho), vv, blobs)
        lam_m[r0:r0 + c], obj_m[r0:r0 + c] = lam, obj
        done += c
    return done


def topk_stable_desc(score, k):
    """First k of the stable descending order (ties by ascending index) without sorting everything."""
    N = score.size
    if N <= 8 * k:
        return np.argsort(-score, kind="stable")[:k]
    part = np.argpartition(-score, k)[:k]
    thr = score[part].min()
    c = np.nonzero(score >= thr)[0]
    return c[np.lexsort((c, -score[c]))][:k]


def run(rho, out):
    from oracle import cutsel_oracle as orc
    n = N_VARS
    N = comb(n, rho)
    for nm in ("lam", "obj"):
        np.memmap(os.path.join(SHM, "sdpcs_gold_%s.f64" % nm), dtype=np.float64, mode="w+", shape=(N,)).flush()
    jobs = prefix_jobs(n, rho)
    cores = os.cpu_count() or 1
    # batches of ~400k subsets, interleaved so that all workers finish together
    batches, cur, cnt = [], [], 0
    for j in jobs:
        cur.append(j); cnt += j[3]
        if cnt >= 400000:
            batches.append(cur); cur, cnt = [], 0
    if cur:
        batches.append(cur)
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        done = sum(pool.imap_unordered(job, [(n, rho, b, N) for b in batches], chunksize=1))
    assert done == N
    t_score = time.perf_counter() - t0
    lam = np.array(np.memmap(os.path.join(SHM, "sdpcs_gold_lam.f64"), dtype=np.float64, mode="r", shape=(N,)))
    obj = np.array(np.memmap(os.path.join(SHM, "sdpcs_gold_obj.f64"), dtype=np.float64, mode="r", shape=(N,)))
    for nm in ("lam", "obj"):
        os.remove(os.path.join(SHM, "sdpcs_gold_%s.f64" % nm))
    t0 = time.perf_counter()
    viol = lam < orc.THRES_NEG_EIGVAL
    s1 = np.where(viol, -lam, -np.inf)
    o1 = topk_stable_desc(s1, K)
    o1 = o1[np.isfinite(s1[o1])]
    o2 = topk_stable_desc(obj, K)
    # combined rule (cut_select_qp.py:603-630) restricted to what can reach the first K places: with >= K strong
    # candidates the list is the K strongest by obj, re-scored obj + BIG_M, provided no walked non-violated candidate
    # can overtake them (obj - BIG_M < pivot + BIG_M); asserted, else the full rule of the oracle is applied
    strong = viol & (obj > orc.THRES_MIN_OPT)
    n_strong = int(strong.sum())
    if n_strong >= K:
        s3 = np.where(strong, obj, -np.inf)
        o4 = topk_stable_desc(s3, K)
        pivot = obj[o4[-1]]
        walked_nonviol = (~viol) & (obj > orc.THRES_MIN_OPT) & (obj >= pivot)
        assert not walked_nonviol.any() or obj[walked_nonviol].max() - orc.BIG_M < pivot + orc.BIG_M
        sc4 = obj[o4] + orc.BIG_M
        new_strat = 1 if K / K < K / N else 4
        counts = [N, K, K]
    else:
        new_strat, order, score = orc.select_comb(obj, lam, K)
        o4, sc4 = order[:K].copy(), score[:K].copy()
        counts = [N, int(viol.sum()), n_strong]
    t_sel = time.perf_counter() - t0
    tag = "n%d_rho%d" % (n, rho)
    out[tag + "_s1_idx"], out[tag + "_s1_score"] = o1.astype(np.int64), s1[o1]
    out[tag + "_s2_idx"], out[tag + "_s2_score"] = o2.astype(np.int64), obj[o2]
    out[tag + "_s4_idx"], out[tag + "_s4_score"] = o4.astype(np.int64), sc4
    out[tag + "_s4_newstrat"] = np.array(new_strat)
    out[tag + "_s4_counts"] = np.array(counts, dtype=np.int64)
    out[tag + "_n_violated"] = np.array(int(viol.sum()))
    out[tag + "_n_strong"] = np.array(n_strong)
    # smallest gap between consecutive selected scores: far above the 1e-10 evaluation noise -> the order is unambiguous
    out[tag + "_min_gap"] = np.array([np.min(-np.diff(s1[o1])), np.min(-np.diff(obj[o2])), np.min(-np.diff(sc4))])
    # checksums of all scores (order-independent up to float addition order; for information)
    out[tag + "_sum_lam"], out[tag + "_sum_obj"] = np.array(lam.sum()), np.array(obj.sum())
    print("rho %d: N %d scored in %.0f s on %d cores, selection %.0f s, violated %d, strong %d, min gaps %s"
          % (rho, N, t_score, cores, t_sel, int(viol.sum()), n_strong, out[tag + "_min_gap"]))


def main():
    rhos = [int(a) for a in sys.argv[1:]] or [4, 5]
    path = os.path.join(HERE, "fullsize_topk.npz")
    out = {}
    if os.path.exists(path):
        with np.load(path) as z:
            out = {k: z[k] for k in z.files}
    for rho in rhos:
        run(rho, out)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
