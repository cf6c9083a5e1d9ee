#!/usr/bin/env python
"""Layer-by-layer check of the tcgen05 int8-sliced MLP against oracle/nn_i8_model.py (needs a B200).

    python tools/i8_debug.py [rho ...]
Prints, per net and per tansig layer, the largest deviation of the scaled pre-activations and where it sits,
then the end-to-end deviation of sdpcs_nn_eval (both engines) from the NNs.so-exact C oracle.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdpcutsel_via_nn_b200 as pkg  # noqa: E402
from oracle import cutsel_oracle as orc  # noqa: E402
from oracle import nn_i8_model as m8  # noqa: E402


def inputs(rho, m, seed=3):
    rng = np.random.default_rng(seed)
    nin = rho * (rho + 3) // 2
    return np.concatenate([rng.uniform(0, 1, (m, rho)), rng.uniform(-1.0 / rho, 1.0 / rho, (m, nin - rho))], axis=1)


def main():
    rhos = [int(a) for a in sys.argv[1:]] or [5, 3]
    eng = pkg._capi.Engine(0)
    for rho in rhos:
        blob = pkg.nn_weights.load_packed(rho)
        eng.set_weights(rho, blob)
        nhid = int(blob[1]) - 1
        for m in (128, 300, 1000):
            x = inputs(rho, m)
            for layer in range(nhid):
                try:
                    zg = eng.nn_debug_layer(rho, x, layer)
                except Exception as e:  # noqa: BLE001
                    print("rho=%d m=%d layer=%d: ERROR %s" % (rho, m, layer, e))
                    break
                _, zm = m8.forward(blob, x, layer)
                d = np.abs(zg - zm)
                i, j = np.unravel_index(np.argmax(d), d.shape)
                print("rho=%d m=%4d layer=%d: max |dz| %.3e at row %d neuron %d (gpu %.17g model %.17g); rows with any |dz|>1e-9: %d"
                      % (rho, m, layer, d.max(), i, j, zg[i, j], zm[i, j], int((d.max(axis=1) > 1e-9).sum())))
                if d.max() > 1e-9:
                    bad = np.argwhere(d > 1e-9)
                    print("   first bad entries (row, neuron):", bad[:12].tolist())
                    print("   gpu row %d:" % bad[0][0], np.array2string(zg[bad[0][0], :8], precision=6))
                    print("   mdl row %d:" % bad[0][0], np.array2string(zm[bad[0][0], :8], precision=6))
        for m in (1, 127, 129, 4097, 200000):
            x = inputs(rho, m, seed=11)
            yo = orc.nn_eval(blob, x)
            for name, engine in (("tcgen05", 0), ("dmma", 1)):
                eng.set_params(nn_engine=engine)
                t0 = time.perf_counter()
                try:
                    y = eng.nn_eval(rho, x)
                except Exception as e:  # noqa: BLE001
                    print("rho=%d m=%d %s: ERROR %s" % (rho, m, name, e))
                    continue
                dt = time.perf_counter() - t0
                print("rho=%d m=%6d %-7s: max |y - oracle| %.3e  (%.1f ms incl. copies, fallbacks %d)"
                      % (rho, m, name, np.abs(y - yo).max(), dt * 1e3, eng.timings()["nn_fallbacks"]))
            eng.set_params()


if __name__ == "__main__":
    main()
