#!/usr/bin/env python
"""Numerical model of the int8-sliced (Ozaki-style) NN_rhoD forward pass used by the tcgen05 MLP kernel.

Activations a in (-1, 1) and weights (per-neuron scale) are written as balanced base-256 digit strings
(int8 digits), the digit-pair products are accumulated exactly in int32 (tcgen05.mma kind::i8) per
"diagonal" (sum of digit positions), and the kept diagonals are recombined in integer / FP64 arithmetic.
This script measures the error of that scheme against an 80-bit long-double forward pass and compares
it with the error of the plain FP64 evaluation (the reference's arithmetic), for several digit counts.

    python tools/i8_mlp_model.py [rho] [n_samples]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdpcutsel_via_nn_b200 as pkg  # noqa: E402

LD = np.longdouble


def digits_balanced(v, ns):
    """v: int64 array (|v| < 2^(8 ns - 2)); returns ns int64 arrays of balanced digits in [-128, 127], most significant first."""
    c = 0
    for i in range(ns):
        c |= 0x80 << (8 * i)
    u = v + c
    out = []
    for i in range(ns):
        out.append(((u >> (8 * i)) & 0xFF) - 128)
    rec = sum(d << (8 * i) for i, d in enumerate(out))
    assert np.array_equal(rec, v)
    return out[::-1]


def quant_weights(W, nsw):
    """per-row scale: W[j, :] = scale[j] * wint[j, :] * 2^-(8 nsw - 2)"""
    mx = np.abs(W).max(axis=1)
    mx[mx == 0] = 1.0
    e = np.ceil(np.log2(mx))            # power-of-two row scale >= max|w|
    scale = 2.0 ** e
    wint = np.rint(W / scale[:, None] * 2.0 ** (8 * nsw - 2)).astype(np.int64)
    return scale, wint


def layer_i8(a, W, b, nsa, nsw, dmax, ea):
    """a: (m, k) doubles with |a| * 2^ea < 2^(8 nsa - 2). Returns z (m, h) computed by the sliced scheme."""
    vint = np.rint(a * 2.0 ** ea).astype(np.int64)
    ad = digits_balanced(vint, nsa)
    scale, wint = quant_weights(W, nsw)
    wd = digits_balanced(wint, nsw)
    nd = dmax + 1
    P = [np.zeros((a.shape[0], W.shape[0]), dtype=np.int64) for _ in range(nd)]
    for s in range(nsa):
        for t in range(nsw):
            if s + t <= dmax:
                P[s + t] += ad[s] @ wd[t].T
    for p in P:
        assert np.abs(p).max() < 2 ** 31
    # recombine: value = sum_d P_d 2^(8 (dmax - d)), split into two int64 halves, one FP64 rounding
    lo_n = min(4, nd)
    L = np.zeros_like(P[0])
    H = np.zeros_like(P[0])
    for d in range(nd):
        sh = 8 * (dmax - d)
        if dmax - d < lo_n:
            L += P[d] << sh
        else:
            H += P[d] << (sh - 32)
    assert np.abs(L).max() < 2 ** 51 and np.abs(H).max() < 2 ** 51
    val = H.astype(np.float64) * 2.0 ** 32 + L.astype(np.float64)
    # units: pair (s,t) has weight 2^(8 (nsa + nsw - 2 - s - t)); we dropped the common factor 2^(8 (nsa + nsw - 2 - dmax))
    unit = 2.0 ** (8 * (nsa + nsw - 2 - dmax)) * 2.0 ** (-ea) * 2.0 ** (-(8 * nsw - 2))
    return val * (scale * unit)[None, :] + b[None, :]


def tansig(z):
    return 2.0 / (1.0 + np.exp(-2.0 * z)) - 1.0


def forward_ref(net, x, dtype):
    a = ((x.astype(dtype) - net["x_xoffset"].astype(dtype)) * net["x_gain"].astype(dtype)) + dtype(-1.0)
    L = len(net["W"])
    for l in range(L - 1):
        z = a @ net["W"][l].astype(dtype).T + net["b"][l].astype(dtype)
        a = dtype(2.0) / (dtype(1.0) + np.exp(dtype(-2.0) * z)) - dtype(1.0)
    y = a @ net["W"][L - 1].astype(dtype).T + net["b"][L - 1].astype(dtype)
    return ((y[:, 0] - dtype(-1.0)) / dtype(net["y_gain"])) + dtype(net["y_xoffset"])


def forward_i8(net, x, nsa, nsw, dmax, nsa0=None):
    p = ((x - net["x_xoffset"]) * net["x_gain"]) + -1.0
    L = len(net["W"])
    a = p
    for l in range(L - 1):
        ns = (nsa0 or nsa) if l == 0 else nsa
        ea = 8 * ns - 3 if l == 0 else 8 * ns - 2          # inputs in (-2, 2), hidden activations in [-1, 1]
        z = layer_i8(a, net["W"][l], net["b"][l], ns, nsw, dmax + (ns - nsa), ea)
        a = tansig(z)
    y = a @ net["W"][L - 1].T + net["b"][L - 1]
    return ((y[:, 0] - -1.0) / net["y_gain"]) + net["y_xoffset"]


def main():
    rho = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    net = pkg.nn_weights.unpack_blob(pkg.nn_weights.load_packed(rho))
    rng = np.random.default_rng(5)
    nin = rho * (rho + 3) // 2
    x = np.concatenate([rng.uniform(0, 1, (m, rho)), rng.uniform(-1.0 / rho, 1.0 / rho, (m, nin - rho))], axis=1)
    y_true = forward_ref(net, x, LD)
    y64 = forward_ref(net, x, np.float64)
    e64 = np.abs(y64.astype(LD) - y_true).astype(np.float64)
    print("rho=%d  m=%d   row scales: %s" % (rho, m, [float(np.abs(W).max()) for W in net["W"]]))
    print("plain FP64 forward vs long double : max %.3e  rms %.3e" % (e64.max(), np.sqrt((e64 ** 2).mean())))
    for nsa, nsw, dmax in [(6, 6, 5), (6, 6, 6), (7, 6, 6), (7, 7, 6), (7, 7, 7), (6, 7, 6)]:
        y = forward_i8(net, x, nsa, nsw, dmax)
        e = np.abs(y.astype(LD) - y_true).astype(np.float64)
        d = np.abs(y - y64)
        npairs = sum(1 for s in range(nsa) for t in range(nsw) if s + t <= dmax)
        print("i8 nsa=%d nsw=%d diagonals=%d pairs=%2d : vs truth max %.3e rms %.3e | vs fp64 max %.3e"
              % (nsa, nsw, dmax + 1, npairs, e.max(), np.sqrt((e ** 2).mean()), d.max()))


if __name__ == "__main__":
    main()
