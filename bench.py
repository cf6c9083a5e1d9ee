#!/usr/bin/env python
"""Headline benchmark: candidate cuts scored+selected per second (BASELINE.json metric).

Workload (BASELINE.json configs[3]): synthetic n=125, density 75 % BoxQP, rho=5, ALL C(125,5) = 234,531,275
subsets, eigenvalue + NN_5D scoring of every candidate, combined selection (strat 4) of k = 5000; the subset rank
space is sharded in contiguous ranges over the N GPUs (strong scaling), local top-k + all-gather + merge.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on all host cores
    python bench.py --workload cfg2           # BASELINE configs[1]: spar125-075-1 rounds through the drop-in CutSolver

One JSON line on stdout (rank 0). `value` = device-resident throughput; `e2e` = the same through the public call with
host buffers (H2D of the LP point, D2H of the selection and its near-tie band, near-tie resolution) inside the timed
region.  `selection.matches_oracle_golden` compares the selected indices with the list the ORACLE computed over the
whole cover (tests/golden/fullsize_topk.npz, made by tests/golden/make_golden_fullsize.py); `selection.idx_sha256`
lets runs at different GPU counts be compared.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate cuts scored+selected/sec"
W_FLOPS = {3: 11601, 4: 12471, 5: 28708}      # algorithmic FP64 flop / subset, eig + NN
WORKLOADS = {
    "cfg4": dict(n=125, rho=5, density=0.75, k=5000, name="synthetic n=125 d=75% BoxQP, rho=5, all C(125,5)=234,531,275 subsets, eig+NN_5D, strat 4, k=5000"),
    "cfg3": dict(n=125, rho=4, density=0.75, k=5000, name="synthetic n=125 d=75% BoxQP, rho=4, all C(125,4)=9,691,375 subsets, eig+NN_4D, strat 4, k=5000"),
    "small": dict(n=60, rho=5, density=0.75, k=5000, name="synthetic n=60 rho=5 (debug)"),
    "patternE5": dict(n=125, rho=5, density=0.75, k=5000, pattern=True,
                      name="synthetic n=125 d=75% BoxQP, rho=5, pattern-E cover P^E_5 (9,775,800 cliques of sizes 2..5, built on the device), eig+NN, strat 4, k=5000"),
    # drop-in workloads: whole separation rounds through the CutSolver surface (selection + cut rows + triangles)
    "cfg1": dict(dropin="spar030-060-1", rho=3, all_subsets=True, name="spar030-060-1 BoxQP, rho=3, all C(30,3)=4,060 subsets, strat 1, 10%"),
    "cfg2": dict(dropin="spar125-075-1", rho=3, triangles=True, name="spar125-075-1 BoxQP, rho=3, P^E_3 (133,242) + triangles, strat 1/2/4, k=5000, 20 LP points"),
    "cfg5": dict(dropin="qcqp", name="q_20_20_100_1 QCQP, rho=3..5 (1,140/4,845/15,504), strat 1 and 4 rounds"),
}


def comb(n, k):
    r = 1
    for j in range(k):
        r = r * (n - j) // (j + 1)
    return r


def config_of(wl, args):
    """The `config` object of the JSON line: identical for the GPU arm and the reference arm of one workload."""
    cfg = dict(workload=wl["name"], strat=args.strat)
    if "n" in wl:
        N = comb(wl["n"], wl["rho"])
        cfg.update(n=wl["n"], rho=wl["rho"], k=wl["k"], candidates=N, shard="contiguous lex-rank ranges",
                   l2_policy="each step streams %.1f GB of scores per GPU through HBM (> 126 MB L2)" % (N / max(args.gpus, 1) * 16 / 1e9))
    return cfg


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            f = [s.strip() for s in line.split(",")]
            if len(f) < 7 or not (t0 - 0.1 <= t <= t1 + 0.3):
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------------------------
# CPU legs: only oracle/ and numpy below this line -- the product package is NOT imported by them
# ---------------------------------------------------------------------------------------------------------------
def _packed_blob(rho):
    """NN_rhoD weight blob straight from the data file shipped with the package (no product code involved)."""
    with np.load(os.path.join(ROOT, "sdpcutsel-via-nn_b200", "weights", "neural_nets.npz")) as z:
        return np.array(z["nn%dD" % rho], dtype=np.float64)


def cpu_block_job(args):
    """Oracle port on the subsets with leading indices (i1, i2) of the given blocks: score (eig + NN) + combined selection.
    Returns (subsets, seconds)."""
    n, rho, density, blocks, k = args
    from oracle import cutsel_oracle as orc
    Q_arr, _ = orc.boxqp_arrays(orc.synth_instance(n, density, seed=7))
    vv = orc.synth_point(n, seed=8)
    blobs = {rho: _packed_blob(rho)}
    idx = np.concatenate([orc.cover_all_block(n, rho, i1, i2) for i1, i2 in blocks])    # input preparation, not timed
    t0 = time.perf_counter()
    lam, obj = orc.score_cover(Q_arr, n, idx, np.full(idx.shape[0], rho), vv, blobs)
    orc.select_comb(obj, lam, min(k, idx.shape[0]))
    return idx.shape[0], time.perf_counter() - t0


def sample_blocks(n, rho, target, parts):
    """`parts` lists of (i1, i2) prefix blocks of the enumeration, each holding about `target` subsets."""
    out, cur, cnt = [], [], 0
    for i1 in range(n):
        for i2 in range(i1 + 1, n):
            c = comb(n - 1 - i2, rho - 2)
            if not c:
                continue
            cur.append((i1, i2)); cnt += c
            if cnt >= target:
                out.append(cur); cur, cnt = [], 0
                if len(out) == parts:
                    return out
    if cur:
        out.append(cur)
    return out


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's CPU path as restated by the oracle (numpy LAPACK eigvalsh + NNs.so-exact C network,
    the reference's selection rule) on all host cores, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    N = comb(wl["n"], wl["rho"])
    per = min(200000, max(1, N // cores))
    jobs = [(wl["n"], wl["rho"], wl["density"], b, wl["k"]) for b in sample_blocks(wl["n"], wl["rho"], per, cores)]
    units = 0
    with mp.get_context("spawn").Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(cpu_block_job, jobs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            units += sum(u for u, _ in pool.map(cpu_block_job, jobs))
        dt = time.perf_counter() - t0
    val = units / dt
    sample = "%d blocks of about %d consecutive subsets per step (%d of the %d candidates), one process per core; the " \
             "selection rule runs per block, not over all %d candidates" % (len(jobs), per, units // max(args.steps, 1), N, N)
    out = dict(metric=METRIC, value=val, unit="subsets/s", n_gpus=args.gpus, steps=args.steps,
               warmup=args.warmup, ms_per_step=dt / args.steps * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
               dtype="f64", data="synthetic", impl="reference", config=config_of(wl, args),
               cpu_baseline=dict(value=val, unit="subsets/s", cores=cores, kind="port", sample=sample),
               e2e=dict(value=val, unit="subsets/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(out))


def cpu_baseline_port(wl):
    """Oracle port, 1 core, first prefix blocks of the same workload (~1 M subsets)."""
    n, rho = wl["n"], wl["rho"]
    blocks = sample_blocks(n, rho, min(comb(n, rho), 1000000), 1)[0]
    units, dt = cpu_block_job((n, rho, wl["density"], blocks, wl["k"]))
    return dict(value=units / dt, unit="subsets/s", cores=1, kind="port", host_cores=os.cpu_count(),
                sample="first %d subsets of the same workload (whole (i1,i2) prefix blocks), oracle port (numpy eigvalsh + C NN), %.1f s" % (units, dt))


def cpu_baseline_reference(wl, strat):
    """The UNMODIFIED reference loop (baseline/_ref, tools/install_reference.py), 1 core: _sel_eigcut_by_ordering_on_measure +
    _gen_eigcuts_selected, the span cut_select_algo times as sep_times (cut_select_qp.py:162-187), on the candidates of the
    same instance and LP point whose indices all lie among the first m variables (C(m, rho) ~ 1e5: the reference keeps
    ~1.3 KB per candidate in RAM and enumerates in pure Python, so the full cover is out of its reach)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refloader
    loaded = refloader.load_reference()
    if loaded is None:
        return dict(unavailable="no reference tree under baseline/_ref (run tools/install_reference.py where /root/reference exists)")
    ref, _, d = loaded
    from oracle import cutsel_oracle as orc
    n, rho = wl["n"], wl["rho"]
    m = rho
    while comb(m + 1, rho) <= 100000 and m + 1 <= n:
        m += 1
    Qf = orc.synth_instance(n, wl["density"], seed=7)[:m, :m]
    vv_full = orc.synth_point(n, seed=8)
    iu = np.triu_indices(n)
    Xf = np.zeros((n, n))
    Xf[iu] = vv_full[:n * (n + 1) // 2]
    vv = np.concatenate([Xf[:m, :m][np.triu_indices(m)], vv_full[n * (n + 1) // 2:][:m]])
    Q_arr, _ = orc.boxqp_arrays(Qf)
    with refloader.in_reference_dir(d):
        cs = ref.CutSolver()
        cs._Q, cs._Q_adj, cs._Q_arr = -Qf / 2, np.ones((m, m)), Q_arr          # complete pattern: P^E_rho = all subsets
        cs._nb_vars, cs._nb_lifted, cs._dim = m, m * (m + 1) // 2, rho
        cs._my_prob = refloader.Cplex()
        cs._load_neural_nets()
        N = cs._get_sdp_vertex_cover(rho)                                         # once per instance, not timed
        k = min(int(np.floor(0.1 * N)), 5000)
        t0 = time.perf_counter()
        out = cs._sel_eigcut_by_ordering_on_measure(strat, vv, 1, sel_size=k)
        rl = out[1] if strat == 4 else out
        t1 = time.perf_counter()
        nb = cs._gen_eigcuts_selected(strat, k, rl, vars_values=vv)
        t2 = time.perf_counter()
    return dict(value=N / (t2 - t0), unit="subsets/s", cores=1, kind="reference", host_cores=os.cpu_count(),
                select_s=t1 - t0, gen_cuts_s=t2 - t1, cuts=int(nb), k=k,
                sample="unmodified reference (baseline/_ref), all C(%d,%d) = %d subsets of the same instance and LP point restricted "
                       "to the first %d variables, strat %d, selection + cut generation (cut_select_qp.py:162-187)" % (m, rho, N, m, strat))


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def golden_check(wl, strat, idx):
    """Compare the selected indices with the oracle's whole-cover selection (tests/golden/fullsize_topk.npz)."""
    path = os.path.join(ROOT, "tests", "golden", "fullsize_topk.npz")
    key = "n%d_rho%d_s%d_idx" % (wl.get("n", 0), wl.get("rho", 0), strat)
    if wl.get("pattern") or not os.path.exists(path):
        return None
    with np.load(path) as z:
        if key not in z.files:
            return None
        want = z[key]
    return bool(np.array_equal(np.asarray(idx, dtype=np.int64), want[:len(idx)]) and len(idx) == min(len(want), wl["k"]))


def dropin_rounds(name, reps=3):
    """Separation rounds of a real instance through the drop-in CutSolver surface (what cut_select_algo calls per round,
    cut_select_qp.py:162-187): selection, cut rows, triangle rows.  Wall clock, host buffers, LP sink included.
    The cyclic garbage collector is paused while rounds are timed: a round allocates ~50,000 small Python objects (the
    reference's tuple / SparsePair formats) and a generation-2 collection landing inside one costs more than the round."""
    import gc
    gc.collect()
    gc.disable()
    try:
        return _dropin_rounds(name, reps)
    finally:
        gc.enable()


def _dropin_rounds(name, reps):
    import sdpcutsel_via_nn_b200 as pkg
    from oracle import cutsel_oracle as orc
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
    out = {}
    if name == "qcqp":
        n = 20
        for dim in (3, 4, 5):
            cs = pkg.CutSolverQCQP()
            cs.set_instance(g["qcqp_Q_arr"], g["qcqp_adj"], n, dim=dim, Q_adj_cons=g["qcqp_adj_cons"])
            cs._load_neural_nets()
            agg_cons = cs._CutSolverQCQP__get_vertex_cover(dim)
            agg = cs._agg_list[:]
            N = len(agg)
            k = max(1, min(int(np.floor(0.1 * N)), 5000))
            for strat in (1, 4):
                ts = []
                for rep in range(reps + 1):
                    vv = orc.synth_point(n, seed=11 + rep)
                    cs._my_prob.linear_constraints.rows = []
                    t0 = time.perf_counter()
                    cs.select_and_cut_round(strat, vv, k, agg, agg_cons)
                    ts.append(time.perf_counter() - t0)
                out["rho%d_strat%d" % (dim, strat)] = dict(candidates=N, k=k, round_ms=float(np.median(ts[1:])) * 1e3)
        return out
    Qf = g["inst_%s_Q" % name.replace("-", "_")].astype(np.float64)
    Q_arr, adj = pkg.synthetic.boxqp_arrays(Qf)
    n = Qf.shape[0]
    wl = WORKLOADS["cfg1" if name == "spar030-060-1" else "cfg2"]
    cs = pkg.CutSolver()
    cs.set_instance(Q_arr, adj, n, dim=wl["rho"])
    cs._load_neural_nets()
    t0 = time.perf_counter()
    N = cs._get_sdp_vertex_cover(wl["rho"], ch_ext=-1 if wl.get("all_subsets") else 0)
    out["cover_ms"] = (time.perf_counter() - t0) * 1e3
    k = min(int(np.floor(0.1 * N)), 5000)
    out.update(candidates=N, k=k)
    if wl.get("triangles"):
        t0 = time.perf_counter()
        cs._CutSolver__preprocess_triangle_ineq()
        out["triangle_preprocess_ms"] = (time.perf_counter() - t0) * 1e3
    points = [orc.synth_point(n, seed=10 + i) for i in range(20 if wl.get("triangles") else 5)]
    for strat in ((1, 2, 4) if wl.get("triangles") else (1,)):
        ph = dict(select=[], gen_cuts=[], triangles=[], round=[])
        for rep, vv in enumerate(points):
            cs._my_prob.linear_constraints.rows = []
            t0 = time.perf_counter()
            r = cs._sel_eigcut_by_ordering_on_measure(strat, vv, 1, sel_size=k)
            rl = r[1] if strat == 4 else r
            t1 = time.perf_counter()
            nb = cs._gen_eigcuts_selected(strat, k, rl, vars_values=vv)
            t2 = time.perf_counter()
            nt = cs._CutSolver__separate_and_add_triangle(0.1, vv) if wl.get("triangles") else 0
            t3 = time.perf_counter()
            if rep:
                for nm, dt in (("select", t1 - t0), ("gen_cuts", t2 - t1), ("triangles", t3 - t2), ("round", t3 - t0)):
                    ph[nm].append(dt * 1e3)
        out["strat%d" % strat] = dict(sdp_cuts=int(nb), tri_cuts=int(nt), degenerate=int(rl.degenerate),
                                      **{nm + "_ms": float(np.median(v)) for nm, v in ph.items()})
        out["strat%d" % strat]["subsets_per_s"] = N / (out["strat%d" % strat]["select_ms"] * 1e-3)
        # the same rounds with an LP sink that takes the rows as CSR arrays (linear_constraints.add_rows_csr: a thin CPXaddrows
        # wrapper): no per-row SparsePair objects -- what is left is the selection call and the C ABI
        sink, keep = _CsrSink(), cs._my_prob.linear_constraints
        cs._my_prob.linear_constraints = sink
        try:
            ts = []
            for rep, vv in enumerate(points):
                sink.blocks = []
                t0 = time.perf_counter()
                r = cs._sel_eigcut_by_ordering_on_measure(strat, vv, 1, sel_size=k)
                cs._gen_eigcuts_selected(strat, k, r[1] if strat == 4 else r, vars_values=vv)
                if wl.get("triangles"):
                    cs._CutSolver__separate_and_add_triangle(0.1, vv)
                if rep:
                    ts.append((time.perf_counter() - t0) * 1e3)
            out["strat%d" % strat]["round_csr_sink_ms"] = float(np.median(ts))
            out["strat%d" % strat]["rows_csr_sink"] = int(sum(len(b[3]) for b in sink.blocks))
        finally:
            cs._my_prob.linear_constraints = keep
    return out


class _CsrSink(object):
    """LP row sink with the one-shot CSR entry point the drop-in looks for (cut_select_qp._add_rows_csr)."""

    def __init__(self):
        self.blocks, self.rows = [], []

    def add_rows_csr(self, rowptr, ind, val, rhs, senses):
        self.blocks.append((rowptr, ind, val, rhs, senses))

    def add(self, lin_expr=None, rhs=None, senses=None, **kw):
        self.rows.extend(zip(lin_expr, rhs, senses))


PUBLISHED_SEP = {"cfg2": "reference separation time per round on spar125-075-1 (data_tables/data_all_boxqp_4rounds.csv:100): "
                         "feasibility 3.54-3.65 s, optimality 2.33-2.52 s, combined 2.50-5.43 s (unstated CPU, 1 thread)",
                 "cfg1": "reference separation time per round on spar030-060-1, P^E_3 N = 756 (data_all_boxqp_4rounds.csv:7): 0.022 s / 0.017 s"}


def run_dropin_workload(args, wl):
    """--workload cfg1 / cfg2 / cfg5: whole rounds through the Python drop-in; value = candidates scored+selected per second of
    the selection call (strat as given), the per-phase round times ride along."""
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    res = dropin_rounds(wl["dropin"])
    if wl["dropin"] == "qcqp":
        key = "rho5_strat%d" % (4 if args.strat == 4 else 1)
        val, ms = res[key]["candidates"] / (res[key]["round_ms"] * 1e-3), res[key]["round_ms"]
    else:
        key = "strat%d" % (args.strat if ("strat%d" % args.strat) in res else 1)
        val, ms = res[key]["subsets_per_s"], res[key]["select_ms"]
    out = dict(metric=METRIC, value=val, unit="subsets/s", n_gpus=1, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
               higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="reference instance, synthetic LP points",
               config=config_of(wl, args), rounds=res, published=PUBLISHED_SEP.get(args.workload),
               e2e=dict(value=val, unit="subsets/s", note="value is already end to end (Python call surface, host buffers)"))
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--strat", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-screened", action="store_true", help="skip the second measurement in screen-and-refine mode")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short cfg1 / cfg2 / cfg3 / cfg5 measurements appended at N=1")
    ap.add_argument("--nn-engine", default="tcgen05", choices=["tcgen05", "dmma"],
                    help="NN_rhoD evaluation: int8-sliced tcgen05 contraction (default) or FP64 DMMA")
    ap.add_argument("--fused-prep", action="store_true",
                    help="tcgen05 engine: build the layer-0 digit images in the MLP kernel's producer warps (no image in HBM)")
    ap.add_argument("--separate-eig", action="store_true",
                    help="tcgen05 engine: lam_min by its own kernel (k_score_feas) instead of inside the staging kernel (A/B)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if "n" not in wl:
            wl = WORKLOADS["cfg4"]
        return run_reference(args, wl, rank, world)
    if "dropin" in wl:
        return run_dropin_workload(args, wl) if rank == 0 else None

    import torch
    import torch.distributed as dist
    import sdpcutsel_via_nn_b200 as pkg
    from sdpcutsel_via_nn_b200 import neartie
    from sdpcutsel_via_nn_b200.distributed import ShardedSelector, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, rho, k = wl["n"], wl["rho"], wl["k"]
    N = comb(n, rho)
    r0, r1 = shard_range(N, world, rank)

    Q_arr, adj = pkg.synthetic.boxqp_arrays(pkg.synthetic.instance(n, wl["density"], seed=7))
    vv = pkg.synthetic.lp_point(n, seed=8)
    blobs = {d: pkg.nn_weights.load_packed(d) for d in range(2, rho + 1)}
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        eng = pkg._capi.Engine(local_rank)
        eng.set_stream(stream.cuda_stream)
        g_obj = max(1e-12, 4e-12 * rho * float(np.abs(Q_arr).max()))
        eng.set_params(nn_engine=pkg._capi.NN_DMMA if args.nn_engine == "dmma" else pkg._capi.NN_TCGEN05, nn_fused_prep=1 if args.fused_prep else 2 if args.separate_eig else 0,
                       guard_lam=1e-12, guard_obj=g_obj)
        eng.set_weights(rho, blobs[rho])
        eng.set_instance(n, Q_arr)
        cover_rows = None
        if wl.get("pattern"):
            for d in range(2, rho):
                eng.set_weights(d, blobs[d])
            N = eng.set_cover_pattern(rho, adj)                      # every rank builds the cover, then keeps its shard
            cover_rows = eng.cover_rows()
            r0, r1 = pkg.distributed.shard_cover(eng, world, rank)
        else:
            eng.set_cover_all(rho, r0, r1)
        sel = ShardedSelector(eng, device=dev if world > 1 else None)
        peak = eng.fp64_peak()

        def sets_of(idx):
            if cover_rows is not None:
                return cover_rows[np.asarray(idx, dtype=np.int64)]
            return pkg._capi.unrank(n, rho, np.asarray(idx, dtype=np.int64))

        rescorer = neartie.Rescorer(n, Q_arr, vv, blobs, sets_of)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # upload the LP point once; warm-up
        res = sel.select(args.strat, vv, k)
        for _ in range(max(args.warmup - 1, 0)):
            res = sel.select(args.strat, None, k)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            time.sleep(0.3)
        # ---- device-resident timed region ---------------------------------------------------------
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        score_ms, select_ms, nn_ms, launches, fallbacks = 0.0, 0.0, 0.0, 0, 0
        barrier()
        t_wall0 = time.time()
        e0.record(stream)
        for _ in range(args.steps):
            res = sel.select(args.strat, None, k)
            tm = eng.timings()
            score_ms += tm["score_ms"]; select_ms += tm["select_ms"]; nn_ms += tm["nn_ms"]; fallbacks = tm["nn_fallbacks"]
            launches += tm["score_launches"] + tm["select_launches"]
        e1.record(stream)
        barrier()
        t_wall1 = time.time()
        ms_dev = e0.elapsed_time(e1)
        # ---- end-to-end timed region: host buffers in, resolved host selection out ---------------
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for i in range(args.steps):
            raw = sel.select(args.strat, vv, k)
            res_e2e = neartie.resolve(raw, k, rescorer, 1e-12, g_obj)
        e3.record(stream)
        barrier()
        ms_e2e = e2.elapsed_time(e3)
        clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
        # cross-check at full size (not timed): the FP64 DMMA engine must select the identical list
        engines_agree = None
        if args.nn_engine == "tcgen05":
            eng.set_params(nn_engine=pkg._capi.NN_DMMA)
            res_dmma = sel.select(args.strat, None, k)
            eng.set_params(nn_engine=pkg._capi.NN_TCGEN05)
            engines_agree = bool(np.array_equal(res["idx"], res_dmma["idx"]))
        # ---- screen-and-refine mode (second measurement, not the headline): every candidate scored by the 4-digit
        # engine, the contenders for the k places re-evaluated by the FP64-accurate engine (ScreenedSelector) ----------
        scr = None
        if not args.no_screened and args.nn_engine == "tcgen05" and not wl.get("pattern"):
            from sdpcutsel_via_nn_b200.distributed import ScreenedSelector
            fine = pkg._capi.Engine(local_rank)
            fine.set_stream(stream.cuda_stream)
            fine.set_weights(rho, blobs[rho])
            fine.set_instance(n, Q_arr)
            s_guard = 1e-4 * rho * float(np.abs(Q_arr).max())
            ssel = ScreenedSelector(eng, fine, sets_of, rho, s_guard, g_obj, device=dev if world > 1 else None)
            rs = ssel.select(args.strat, vv, k)
            for _ in range(2):
                rs = ssel.select(args.strat, None, k)
            s_score = s_nn = s_select = 0.0
            barrier()
            e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e4.record(stream)
            for _ in range(args.steps):
                rs = ssel.select(args.strat, None, k)
                tm = eng.timings()
                s_score += tm["score_ms"]; s_nn += tm["nn_ms"]; s_select += tm["select_ms"]
            e5.record(stream)
            barrier()
            ms_s_dev = e4.elapsed_time(e5)
            e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e6.record(stream)
            for _ in range(args.steps):
                raw_s = ssel.select(args.strat, vv, k)
                rs_e2e = neartie.resolve(raw_s, k, rescorer, 1e-12, g_obj)
            e7.record(stream)
            barrier()
            ms_s_e2e = e6.elapsed_time(e7)
            ts = torch.tensor([ms_s_dev, ms_s_e2e, s_score / args.steps, s_nn / args.steps, s_select / args.steps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            ms_s_dev, ms_s_e2e, s_score, s_nn, s_select = (float(v) for v in ts.cpu())
            idx_s = np.asarray(rs_e2e["idx"], dtype=np.int64)
            scr = dict(ms_per_step=ms_s_dev / args.steps, value=N * args.steps / (ms_s_dev * 1e-3), unit="subsets/s",
                       e2e=dict(value=N * args.steps / (ms_s_e2e * 1e-3), ms_per_step=ms_s_e2e / args.steps),
                       screen_kernels_ms=s_score, nn_kernels_ms=s_nn, select_ms=s_select,
                       contenders=ssel.last.get("contenders"), max_screen_error=ssel.last.get("max_screen_error"),
                       screen_guard=s_guard, fallbacks=ssel.fallbacks,
                       matches_oracle_golden=golden_check(wl, args.strat, idx_s),
                       idx_sha256=hashlib.sha256(idx_s.tobytes()).hexdigest(), degenerate=int(rs_e2e["degenerate"]),
                       note="NOT the headline: tier 1 scores every candidate with k_mlp_i8<.., NS=4> (32-bit fixed point, 10 digit pairs, two "
                            "TMEM accumulator stages; NN output within ~3e-6 of the reference's, north_star asks 1e-5); tier 2 re-evaluates "
                            "the contenders (winners + everything within screen_guard of the k-th score) with the FP64-accurate engine; "
                            "tier 3 = neartie.resolve.  The selected list and its scores are the exact engine's.")
            eng.set_params(nn_engine=pkg._capi.NN_TCGEN05, guard_obj=g_obj)
        t = torch.tensor([ms_dev, ms_e2e, score_ms / args.steps, nn_ms / args.steps, select_ms / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e, score_ms_step, nn_ms_step, select_ms_step = (float(v) for v in t.cpu())

    if rank == 0:
        W = W_FLOPS[rho]
        n_local = r1 - r0
        achieved = n_local * W / (score_ms_step * 1e-3) * 1e-12
        traffic = None
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        bf16 = json.load(open(peaks_path)).get("bf16_tflops_sustained") if os.path.exists(peaks_path) else None
        bf16_src = "2 x bf16_tflops_sustained of MEASURED_PEAKS.json" if bf16 else "2 x 1400 (B200_PROFILING.md fallback)"
        bf16 = bf16 or 1400.0
        prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("score_kernel_dram_bytes_per_launch")
            except Exception:
                traffic = None
        if args.nn_engine == "tcgen05":
            # int8 tensor operations actually issued per candidate: 28 digit pairs x (32 + (NHID-1) x 64) x 64 MACs x 2
            nhid = 4 if rho == 5 else 3
            i8_ops = 28 * (32 + (nhid - 1) * 64) * 64 * 2
            i8_tops = n_local * i8_ops / (nn_ms_step * 1e-3) * 1e-12
            eig = "k_score_feas<%d> (FP64 tridiagonal + Laguerre)" % rho if (args.separate_eig or args.fused_prep or args.strat != 4) else \
                  "lam_min (FP64 tridiagonal + Laguerre) inside k_prep_i8<%d,7,FEAS>: nn_kernels_ms includes it" % rho
            kname = "k_mlp_i8<%d> (tcgen05.mma kind::i8, FP64-accurate 7x7-digit slicing of the FP64 MLP, TMEM accumulators) " \
                    "+ k_prep_i8<%d> + %s" % (nhid, rho, eig)
            extra = dict(nn_kernels_ms=nn_ms_step, int8_tensor=dict(achieved=i8_tops, peak=2 * bf16, unit="TOP/s", frac=i8_tops / (2 * bf16),
                                                                  ops_per_subset=i8_ops, peak_source=bf16_src),
                         note="achieved / peak / frac are FP64-EQUIVALENT: algorithmic FP64 flop of SURVEY 8(d) over the measured FP64 DMMA peak "
                              "(north_star's FP64 roofline), not a pipe utilisation; the contraction itself runs on the int8 tensor pipe: "
                              "int8_tensor.frac is the utilisation of the pipe the kernel actually uses")
        else:
            kname = "k_score_nn<%d,16> (DMMA.8x8x4 FP64 MLP) + k_score_feas<%d> (FP64 tridiagonal + Laguerre)" % (rho, rho)
            extra = dict(nn_kernels_ms=nn_ms_step)
        roofline = dict(bound="tensor", kernel=kname, achieved=achieved, peak=peak["dmma_tflops"], unit="TFLOP/s",
                        frac=achieved / peak["dmma_tflops"], frac_kind="fp64_equivalent", fp64_equivalent_frac=achieved / peak["dmma_tflops"],
                        traffic=traffic, flops_per_subset=W, subsets_per_launch=n_local,
                        kernel_ms=score_ms_step, select_ms=select_ms_step, nn_engine=args.nn_engine, nn_fallbacks=int(fallbacks),
                        peak_source="FP64 DMMA.8x8x4 micro-benchmark run in this process (MEASURED_PEAKS.json has no FP64 entry); DFMA peak %.1f" % peak["dfma_tflops"],
                        **extra)
        idx_final = np.asarray(res_e2e["idx"], dtype=np.int64)
        band_rows = int(raw["band"]["idx"].size)
        out = dict(
            metric=METRIC, value=N * args.steps / (ms_dev * 1e-3), unit="subsets/s",
            n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_dev / args.steps, higher_is_better=True,
            scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
            config=config_of(wl, args),
            e2e=dict(value=N * args.steps / (ms_e2e * 1e-3), unit="subsets/s", ms_per_step=ms_e2e / args.steps,
                     h2d_bytes_per_step=int(vv.size * 8), d2h_bytes_per_step=int((k + band_rows) * 32 + 4 * 8400)),
            gpu_launches=int(launches),
            roofline=roofline,
            clocks=clocks,
            selection=dict(n_selected=int(idx_final.size), new_strat=int(res_e2e["new_strat"]), counts=[int(v) for v in res_e2e["counts"]],
                           matches_oracle_golden=golden_check(wl, args.strat, idx_final),
                           idx_sha256=hashlib.sha256(idx_final.tobytes()).hexdigest(),
                           degenerate=int(res_e2e["degenerate"]), n_near_ties=int(res_e2e["n_near_ties"]),
                           guard=dict(raw["guard"], guard_lam=1e-12, guard_obj=g_obj),
                           e2e_matches_resident=bool(np.array_equal(res["idx"], res_e2e["idx"])),
                           tcgen05_and_dmma_engines_select_identically=engines_agree),
        )
        if scr is not None:
            out["screened"] = scr
        if sel.prof:
            out["host_phase_ms_per_select"] = {k: 1e3 * v / (args.warmup + 2 * args.steps + (1 if args.nn_engine == "tcgen05" else 0)) for k, v in sel.prof.items()}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_port(wl)
            try:
                out["cpu_baseline_reference"] = cpu_baseline_reference(wl, args.strat)
            except Exception as e:                      # the reference arm must never take the GPU line down
                out["cpu_baseline_reference"] = dict(unavailable=repr(e)[:200])
        if world == 1 and not args.no_other_configs and args.workload == "cfg4":
            other = {}
            for name, inst in (("cfg1", "spar030-060-1"), ("cfg2", "spar125-075-1"), ("cfg5", "qcqp")):
                try:
                    other[name] = dict(workload=WORKLOADS[name]["name"], published=PUBLISHED_SEP.get(name), **dropin_rounds(inst, reps=2))
                except Exception as e:
                    other[name] = dict(error=repr(e)[:200])
            try:
                other["cfg3"] = quick_throughput(pkg, "cfg3", local_rank, stream.cuda_stream)
            except Exception as e:
                other["cfg3"] = dict(error=repr(e)[:200])
            out["other_configs"] = other
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def quick_throughput(pkg, name, device, stream_ptr, steps=3):
    """Short device-resident + e2e throughput of another all-subsets workload (same path as the headline one)."""
    from sdpcutsel_via_nn_b200.distributed import ShardedSelector
    wl = WORKLOADS[name]
    n, rho, k = wl["n"], wl["rho"], wl["k"]
    Q_arr, _ = pkg.synthetic.boxqp_arrays(pkg.synthetic.instance(n, wl["density"], seed=7))
    vv = pkg.synthetic.lp_point(n, seed=8)
    eng = pkg._capi.Engine(device)
    eng.set_stream(stream_ptr)
    eng.set_weights(rho, pkg.nn_weights.load_packed(rho))
    eng.set_instance(n, Q_arr)
    eng.set_cover_all(rho)
    sel = ShardedSelector(eng, local=True)
    for _ in range(2):
        res = sel.select(4, vv, k)
    t0 = time.perf_counter()
    for _ in range(steps):
        res = sel.select(4, None, k)
    t1 = time.perf_counter()
    for _ in range(steps):
        res = sel.select(4, vv, k)
    t2 = time.perf_counter()
    N = comb(n, rho)
    tm = eng.timings()
    return dict(workload=wl["name"], candidates=N, ms_per_step=(t1 - t0) / steps * 1e3, subsets_per_s=N * steps / (t1 - t0),
                e2e_ms_per_step=(t2 - t1) / steps * 1e3, e2e_subsets_per_s=N * steps / (t2 - t1), score_ms=tm["score_ms"], select_ms=tm["select_ms"],
                fp64_equivalent_tflops=N * W_FLOPS[rho] / (tm["score_ms"] * 1e-3) * 1e-12,
                matches_oracle_golden=golden_check(wl, 4, res["idx"]), idx_sha256=hashlib.sha256(np.asarray(res["idx"], dtype=np.int64).tobytes()).hexdigest())


if __name__ == "__main__":
    main()
