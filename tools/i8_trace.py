#!/usr/bin/env python
"""Pipeline timeline of k_mlp_i8 (needs a B200).  `--build` compiles libsdpcutsel_trace.so (-DSDPCS_I8_TRACE) here,
without arguments the tool loads it, runs NN_5D on 148 * 48 tiles and prints per-step clock64 stamps of CTA 0:

  epilogue warps: t0 step start, t1 accumulators ready (B_FULL), t2 accumulators handed back (B_EMPTY), t3 step end
  MMA warp      : t0 step start, t1 operands ready (B_A0 / B_ACT), t2 accumulators free, t3 MMAs issued
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "sdpcutsel-via-nn_b200")
LIB = os.path.join(PKG, "libsdpcutsel_trace.so")


def build():
    from sdpcutsel_via_nn_b200 import build as b
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    lib = LIB if not extra else LIB.replace(".so", "_" + "_".join(a[2:].replace("=", "") for a in extra) + ".so")
    cmd = [b.NVCC] + b.FLAGS + ["-DSDPCS_I8_TRACE"] + extra + ["-o", lib, b.SRC]
    subprocess.run(cmd, check=True, capture_output=True)
    print(lib)


def main():
    if "--build" in sys.argv:
        return build()
    import sdpcutsel_via_nn_b200 as pkg
    libs = [a for a in sys.argv[1:] if a.endswith(".so")]
    lib = pkg._capi.load_library(libs[0] if libs else LIB)
    rho = 5
    eng = pkg._capi.Engine(0)
    blob = pkg.nn_weights.load_packed(rho)
    eng.set_weights(rho, blob)
    if "--screen" in sys.argv:
        eng.set_params(nn_engine=pkg._capi.NN_SCREEN)
    m = 148 * 48 * 128
    rng = np.random.default_rng(3)
    nin = rho * (rho + 3) // 2
    x = np.concatenate([rng.uniform(0, 1, (m, rho)), rng.uniform(-1.0 / rho, 1.0 / rho, (m, nin - rho))], axis=1)
    eng.nn_eval(rho, x)
    nw, ns = 20, 96
    buf = np.zeros(nw * ns * 4, dtype=np.int64)
    rc = lib.sdpcs_i8_trace_read(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(buf.size))
    assert rc == buf.size, rc
    t = buf.reshape(nw, ns, 4)
    base = t[:16, 0, 0].min()
    t = t - base
    ep = t[:16]
    mma = t[16]
    np.save(os.path.join(ROOT, "gpurun_out", "i8_trace_screen.npy" if "--screen" in sys.argv else "i8_trace.npy"), t)
    if "-q" not in sys.argv:
      print("step |  MMA: start  opnd_ok  acc_free  issued | EPI: full_min full_max | empty_min empty_max | end_min end_max | wait_full(avg) busy(avg)")
    for s in range(ns if "-q" not in sys.argv else 0):
        wf = (ep[:, s, 1] - ep[:, s, 0]).mean()
        busy = (ep[:, s, 3] - ep[:, s, 1]).mean()
        print("%4d | %9d %8d %9d %7d | %13d %8d | %9d %9d | %7d %7d | %8.0f %8.0f"
              % (s, mma[s, 0], mma[s, 1], mma[s, 2], mma[s, 3], ep[:, s, 1].min(), ep[:, s, 1].max(), ep[:, s, 2].min(),
                 ep[:, s, 2].max(), ep[:, s, 3].min(), ep[:, s, 3].max(), wf, busy))
    st = ep[:, 8:88]
    per = (st[:, -1, 0] - st[:, 0, 0]).mean() / (st.shape[1] - 1)
    print("cycles per step %.0f; epilogue wait for accumulators %.0f (%.1f %%); phase 1 (to hand-back) %.0f; rest %.0f"
          % (per, (st[:, :, 1] - st[:, :, 0]).mean(), 100 * (st[:, :, 1] - st[:, :, 0]).mean() / per,
             (st[:, :, 2] - st[:, :, 1]).mean(), (st[:, :, 3] - st[:, :, 2]).mean()))
    # MMA duration: from the issue of step s to the earliest epilogue warp that sees its accumulators
    dur = ep[:, 8:88, 1].min(axis=0) - mma[8:88, 2]
    print("MMA issue -> accumulators visible: mean %.0f  min %d  max %d cycles" % (dur.mean(), dur.min(), dur.max()))
    lag = mma[8:88, 2] - np.maximum(mma[8:88, 1], mma[8:88, 0])
    print("MMA warp waits for the accumulator hand-back (after operands are ready): mean %.0f cycles" % lag.mean())
    # by position in the schedule: step mod 8 = 2 * layer + lane for NHID = 4 (two tiles in flight, four tansig layers each)
    d = np.diff(ep[:, :, 0], axis=1)
    for r in range(8):
        sel = [s for s in range(8, 88) if s % 8 == r]
        print("step mod 8 = %d: step %5.0f (slowest warp %5.0f) | wait %5.0f  phase 1 %5.0f  phase 2 %5.0f | issuer: operands %5.0f  hand-back %5.0f  issue %5.0f"
              % (r, d[:, sel].mean(), d[:, sel].mean(axis=1).max(), (ep[:, sel, 1] - ep[:, sel, 0]).mean(), (ep[:, sel, 2] - ep[:, sel, 1]).mean(),
                 (ep[:, sel, 3] - ep[:, sel, 2]).mean(), (mma[sel, 1] - mma[sel, 0]).mean(), (mma[sel, 2] - mma[sel, 1]).mean(),
                 (mma[sel, 3] - mma[sel, 2]).mean()))
    sp = ep[:, 8:88, 0].max(axis=0) - ep[:, 8:88, 0].min(axis=0)
    print("spread of the step starts across the 16 warps: mean %.0f max %.0f" % (sp.mean(), sp.max()))
    for w in range(16):
        print("warp %2d (q %d cq %d): busy %.0f  wait %.0f" % (w, w & 3, w >> 2, (st[w, :, 3] - st[w, :, 1]).mean(), (st[w, :, 1] - st[w, :, 0]).mean()))


if __name__ == "__main__":
    main()
