#!/usr/bin/env python
"""Producer-warp timeline of the fused-input mode of k_mlp_i8 (trace library; needs a B200)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdpcutsel_via_nn_b200 as pkg
lib = pkg._capi.load_library(os.path.join(ROOT, "sdpcutsel-via-nn_b200", "libsdpcutsel_trace.so"))
n, rho = 60, 5
Q_arr, _ = pkg.synthetic.boxqp_arrays(pkg.synthetic.instance(n, 0.75, seed=7))
vv = pkg.synthetic.lp_point(n, seed=8)
eng = pkg._capi.Engine(0)
eng.set_params(nn_fused_prep=1)
eng.set_weights(rho, pkg.nn_weights.load_packed(rho))
eng.set_instance(n, Q_arr)
eng.set_cover_all(rho)
eng.score(vv, 2)
nw, ns = 20, 96
buf = np.zeros(nw * ns * 4, dtype=np.int64)
rc = lib.sdpcs_i8_trace_read(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(buf.size))
assert rc == buf.size, rc
t = buf.reshape(nw, ns, 4)
for w in (17, 18, 19):
    p = t[w]
    ok = p[:, 3] > 0
    work = (p[ok, 2] - p[ok, 1]); wait = (p[ok, 1] - p[ok, 0]); per = np.diff(p[ok, 0])
    print("producer warp %d: passes traced %d; row preparation %.0f cycles per 32-row pass (min %d max %d); wait for the free A buffer %.0f; period %.0f"
          % (w, ok.sum(), work.mean(), work.min(), work.max(), wait.mean(), per.mean()))
ep = t[:16]
st = ep[:, 8:88]
print("epilogue: cycles per step %.0f, wait for accumulators %.0f" % ((st[:, -1, 0] - st[:, 0, 0]).mean() / (st.shape[1] - 1), (st[:, :, 1] - st[:, :, 0]).mean()))
