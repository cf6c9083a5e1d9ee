#!/usr/bin/env python
"""Headline benchmark: candidate cuts scored+selected per second (BASELINE.json metric).

Workload (BASELINE.json configs[3]): synthetic n=125, density 75 % BoxQP, rho=5, ALL C(125,5) = 234,531,275
subsets, eigenvalue + NN_5D scoring of every candidate, combined selection (strat 4) of k = 5000; the subset rank
space is sharded in contiguous ranges over the N GPUs (strong scaling), local top-k + all-gather + merge.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the oracle port of the reference's CPU path on the host cores

One JSON line on stdout (rank 0). `value` = device-resident throughput; `e2e` = same through the C-ABI call with
host buffers (H2D of the LP point and D2H of the selection inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_FLOPS = {3: 11601, 4: 12471, 5: 28708}      # algorithmic FP64 flop / subset, eig + NN
WORKLOADS = {
    "cfg4": dict(n=125, rho=5, density=0.75, k=5000, name="synthetic n=125 d=75% BoxQP, rho=5, all C(125,5)=234,531,275 subsets, eig+NN_5D, strat 4, k=5000"),
    "cfg3": dict(n=125, rho=4, density=0.75, k=5000, name="synthetic n=125 d=75% BoxQP, rho=4, all C(125,4)=9,691,375 subsets, eig+NN_4D, strat 4, k=5000"),
    "small": dict(n=60, rho=5, density=0.75, k=5000, name="synthetic n=60 rho=5 (debug)"),
    "patternE5": dict(n=125, rho=5, density=0.75, k=5000, pattern=True,
                      name="synthetic n=125 d=75% BoxQP, rho=5, pattern-E cover P^E_5 (12.6 M cliques, built on the device), eig+NN, strat 4, k=5000"),
}


def comb(n, k):
    r = 1
    for j in range(k):
        r = r * (n - j) // (j + 1)
    return r


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


def cpu_window_job(args):
    """Oracle port on one window of ranks: score (eig + NN) + local combined selection. Returns seconds."""
    n, rho, density, r0, r1, k = args
    from oracle import cutsel_oracle as orc
    import sdpcutsel_via_nn_b200 as pkg
    Q_arr, _ = pkg.synthetic.boxqp_arrays(pkg.synthetic.instance(n, density, seed=7))
    vv = pkg.synthetic.lp_point(n, seed=8)
    blobs = {rho: pkg.nn_weights.load_packed(rho)}
    idx = pkg._capi.unrank(n, rho, np.arange(r0, r1))        # input preparation, not timed
    t0 = time.perf_counter()
    lam, obj = orc.score_cover(Q_arr, n, idx, np.full(idx.shape[0], rho), vv, blobs)
    orc.select_comb(obj, lam, min(k, idx.shape[0]))
    return time.perf_counter() - t0


def run_reference(args, wl, rank, world):
    """--impl reference: the oracle port (numpy LAPACK eigvalsh + C NN, the reference's arithmetic) on all host cores,
    each step a bounded sample of the same workload."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    N = comb(wl["n"], wl["rho"])
    per = min(200000, max(1, N // cores))
    jobs = [(wl["n"], wl["rho"], wl["density"], i * per, (i + 1) * per, wl["k"]) for i in range(cores)]
    with mp.get_context("spawn").Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(cpu_window_job, jobs[:cores])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(cpu_window_job, jobs)
        dt = time.perf_counter() - t0
    units = per * cores * args.steps
    val = units / dt
    sample = "%d windows x %d consecutive ranks per step (of %d), one process per core" % (cores, per, N)
    out = dict(metric="candidate cuts scored+selected/sec", value=val, unit="subsets/s", n_gpus=args.gpus, steps=args.steps,
               warmup=args.warmup, ms_per_step=dt / args.steps * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
               dtype="f64", data="synthetic", impl="reference", config=dict(workload=wl["name"]),
               cpu_baseline=dict(value=val, unit="subsets/s", cores=cores, kind="port", sample=sample),
               e2e=dict(value=val, unit="subsets/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
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
    ap.add_argument("--nn-engine", default="tcgen05", choices=["tcgen05", "dmma"],
                    help="NN_rhoD evaluation: int8-sliced tcgen05 contraction (default) or FP64 DMMA")
    ap.add_argument("--fused-prep", action="store_true",
                    help="tcgen05 engine: build the layer-0 digit images in the MLP kernel's producer warps (no image in HBM)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, wl, rank, world)

    import torch
    import torch.distributed as dist
    import sdpcutsel_via_nn_b200 as pkg
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
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        eng = pkg._capi.Engine(local_rank)
        eng.set_stream(stream.cuda_stream)
        eng.set_params(nn_engine=pkg._capi.NN_DMMA if args.nn_engine == "dmma" else pkg._capi.NN_TCGEN05, nn_fused_prep=int(args.fused_prep))
        eng.set_weights(rho, pkg.nn_weights.load_packed(rho))
        eng.set_instance(n, Q_arr)
        if wl.get("pattern"):
            for d in range(2, rho):
                eng.set_weights(d, pkg.nn_weights.load_packed(d))
            N = eng.set_cover_pattern(rho, adj)                      # every rank builds the cover, then keeps its shard
            r0, r1 = pkg.distributed.shard_cover(eng, world, rank)
        else:
            eng.set_cover_all(rho, r0, r1)
        sel = ShardedSelector(eng, device=dev if world > 1 else None)
        peak = eng.fp64_peak()

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
            launches += tm["score_launches"] + tm["select_launches"] * (2 if args.strat == 4 else 1)
        e1.record(stream)
        barrier()
        t_wall1 = time.time()
        ms_dev = e0.elapsed_time(e1)
        # ---- end-to-end timed region: host buffers in, host selection out -------------------------
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for i in range(args.steps):
            res_e2e = sel.select(args.strat, vv, k)
        e3.record(stream)
        barrier()
        ms_e2e = e2.elapsed_time(e3)
        clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
        # cross-check at full size (not timed): the FP64 DMMA engine must select the identical list
        engines_agree = None
        if args.nn_engine == "tcgen05":
            eng.set_params(nn_engine=pkg._capi.NN_DMMA)
            res_dmma = sel.select(args.strat, None, k)
            eng.set_params(nn_engine=pkg._capi.NN_TCGEN05, nn_fused_prep=int(args.fused_prep))
            engines_agree = bool(np.array_equal(res["idx"], res_dmma["idx"]))
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
            kname = "k_mlp_i8<%d> (tcgen05.mma kind::i8, error-free 7x7-digit slicing of the FP64 MLP, TMEM accumulators) " \
                    "+ k_prep_i8<%d> + k_score_feas<%d> (FP64 tridiagonal + Laguerre)" % (nhid, rho, rho)
            extra = dict(nn_kernels_ms=nn_ms_step, int8_tensor=dict(achieved=i8_tops, peak=2 * bf16, unit="TOP/s", frac=i8_tops / (2 * bf16),
                                                                  ops_per_subset=i8_ops, peak_source=bf16_src),
                         note="achieved/peak are FP64-equivalent: algorithmic FP64 flop of SURVEY 8(d) over the measured FP64 DMMA peak "
                              "(north_star's FP64 roofline); the contraction itself runs on the int8 tensor pipe, see int8_tensor")
        else:
            kname = "k_score_nn<%d,16> (DMMA.8x8x4 FP64 MLP) + k_score_feas<%d> (FP64 tridiagonal + Laguerre)" % (rho, rho)
            extra = dict(nn_kernels_ms=nn_ms_step)
        roofline = dict(bound="tensor", kernel=kname, achieved=achieved, peak=peak["dmma_tflops"], unit="TFLOP/s",
                        frac=achieved / peak["dmma_tflops"], traffic=traffic, flops_per_subset=W, subsets_per_launch=n_local,
                        kernel_ms=score_ms_step, select_ms=select_ms_step, nn_engine=args.nn_engine, nn_fallbacks=int(fallbacks),
                        peak_source="FP64 DMMA.8x8x4 micro-benchmark run in this process (MEASURED_PEAKS.json has no FP64 entry); DFMA peak %.1f" % peak["dfma_tflops"],
                        **extra)
        out = dict(
            metric="candidate cuts scored+selected/sec", value=N * args.steps / (ms_dev * 1e-3), unit="subsets/s",
            n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_dev / args.steps, higher_is_better=True,
            scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
            config=dict(workload=wl["name"], n=n, rho=rho, k=k, strat=args.strat, candidates=N, shard="contiguous lex-rank ranges",
                        l2_policy="each step streams %.1f GB of scores/keys through HBM (> 126 MB L2)" % (n_local * 24 / 1e9)),
            e2e=dict(value=N * args.steps / (ms_e2e * 1e-3), unit="subsets/s", ms_per_step=ms_e2e / args.steps,
                     h2d_bytes_per_step=int(vv.size * 8), d2h_bytes_per_step=int((2 if args.strat == 4 else 1) * (k * 32 + 8288))),
            gpu_launches=int(launches),
            roofline=roofline,
            clocks=clocks,
            selection=dict(n_selected=int(res["idx"].size), new_strat=int(res["new_strat"]), counts=[int(v) for v in res["counts"]],
                           e2e_matches_resident=bool(np.array_equal(res["idx"], res_e2e["idx"])),
                           tcgen05_and_dmma_engines_select_identically=engines_agree),
        )
        if sel.prof:
            out["host_phase_ms_per_select"] = {k: 1e3 * v / (args.warmup + 2 * args.steps + (1 if args.nn_engine == "tcgen05" else 0)) for k, v in sel.prof.items()}
        if world == 1 and not args.no_cpu_baseline:
            sample_n = min(N, 1000000)
            dt = cpu_window_job((n, rho, wl["density"], 0, sample_n, k))
            out["cpu_baseline"] = dict(value=sample_n / dt, unit="subsets/s", cores=1, kind="port",
                                       sample="first %d lex ranks of the same workload, oracle port (numpy eigvalsh + C NN), %.1f s" % (sample_n, dt),
                                       host_cores=os.cpu_count())
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
