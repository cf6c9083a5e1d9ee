"""CPU statement of the int8-sliced NN_rhoD forward pass of the tcgen05 engine (TEST INFRASTRUCTURE ONLY).

Mirrors, integer for integer, what ``csrc/mlp_i8_kernels.cuh`` / ``pack_i8`` (capi.cu) compute, so that the GPU
kernel can be checked layer by layer (``sdpcs_nn_debug_layer``):

* weights: row ``j`` of a layer is scaled by ``2^e >= max|W[j,:]|``, ``wint = rint(W * 2^(54-e))``, 7 balanced
  base-256 digits (most significant first);
* layer inputs: ``v = rint(a * 2^51)`` (``2^50`` for the mapminmax'ed network inputs); the two's complement word
  ``8 v`` is cut into six unsigned low digits and one signed top digit (unsigned-A / signed-A MMAs on the GPU);
* digit-pair products on the diagonals ``s + t <= 6`` are summed exactly (int32 on the GPU), recombined as
  ``H * 2^32 + L`` with one FP64 rounding, ``z = fma(val, cs, bs)`` with ``cs = 2^(e-54-ea+48) * (-2 log2 e)``;
* ``tansig`` from the scaled pre-activation; the output layer and ``mapminmax`` reverse in FP64
  (neural_net_3D.m:47-85 in the reference).

Only ``tests/`` import this module; the product never does.
"""
import numpy as np

NS = 7
TANSIG_SCALE = -2.8853900817779268147


def digits(v, ns=NS):
    """int64 array -> ns balanced digits in [-128, 127], most significant first (sum d_s 256^(ns-1-s) == v)."""
    bias = sum(0x80 << (8 * i) for i in range(ns))
    u = v.astype(np.int64) + bias
    out = [((u >> (8 * b)) & 0xFF) - 128 for b in range(ns)]
    return out[::-1]


def digits_twos(v, ns=NS):
    """int64 array -> ns - 1 unsigned low digits [0, 255] and one signed top digit, most significant first."""
    u = v.astype(np.int64)
    out = [(u >> (8 * b)) & 0xFF for b in range(ns - 1)] + [u >> (8 * (ns - 1))]
    assert np.abs(out[-1]).max() <= 128
    return out[::-1]


def unpack_blob(blob):
    blob = np.asarray(blob, dtype=np.float64)
    n_in, L, h = int(blob[0]), int(blob[1]), int(blob[2])
    o = 3
    xo, xg = blob[o:o + n_in], blob[o + n_in:o + 2 * n_in]
    o += 2 * n_in
    Ws, bs = [], []
    for l in range(L):
        rows = 1 if l == L - 1 else h
        cols = n_in if l == 0 else h
        Ws.append(blob[o:o + rows * cols].reshape(rows, cols)); o += rows * cols
        bs.append(blob[o:o + rows]); o += rows
    return dict(W=Ws, b=bs, xo=xo, xg=xg, y_gain=float(blob[o]), y_xoff=float(blob[o + 1]))


def pack_layer(W, b, ea, K, ns=NS):
    """-> (digit slices [ns] of shape (64, K), cs[64], bs[64]); weights carry KW = 8 ns - 2 fractional bits of the row scale"""
    kw = 8 * ns - 2
    h, cols = W.shape
    Wp = np.zeros((64, K))
    Wp[:h, :cols] = W
    mx = np.abs(Wp).max(axis=1)
    e = np.where(mx > 0, np.frexp(np.where(mx > 0, mx, 1.0))[1], 0).astype(np.int64)
    wint = np.rint(np.ldexp(Wp, (kw - e)[:, None])).astype(np.int64)
    cs = np.ldexp(1.0, e - kw - ea + 8 * (ns - 1)) * TANSIG_SCALE
    bsv = np.zeros(64)
    bsv[:h] = TANSIG_SCALE * b
    return digits(wint, ns), cs, bsv


def layer_z(a, wd, cs, bs, scale_log2, ns=NS):
    """a: (m, K) layer inputs; returns the scaled pre-activations z (m, 64) exactly as the kernel forms them."""
    v = np.rint(a * 2.0 ** scale_log2).astype(np.int64) * 8
    ad = digits_twos(v, ns)
    H = np.zeros((a.shape[0], 64), dtype=np.int64)
    L = np.zeros((a.shape[0], 64), dtype=np.int64)
    for d in range(ns):                                   # kept diagonals s + t <= ns - 1
        P = np.zeros((a.shape[0], 64), dtype=np.int64)
        for s in range(d + 1):
            P += ad[s] @ wd[d - s].T
        assert np.abs(P).max() < 2 ** 31
        if ns == 7:
            if d >= 3:
                L += P << (8 * (6 - d))
            else:
                H += P << (8 * (2 - d))
        else:
            L += P << (8 * (ns - 1 - d))
    if ns == 7:
        val = H.astype(np.float64) * 4294967296.0 + L.astype(np.float64)      # one rounding, as fma(dh, 2^32, dl)
    else:
        val = L.astype(np.float64)                                            # exact (|sum| < 2^48)
    z = (val.astype(np.longdouble) * cs[None, :].astype(np.longdouble) + bs[None, :].astype(np.longdouble)).astype(np.float64)
    return z


def tansig_scaled(z):
    """2 / (1 + 2^z) - 1 with z = -2 log2(e) n, i.e. tanh(n)."""
    return (2.0 / (1.0 + np.exp2(z.astype(np.longdouble))) - 1.0).astype(np.float64)


def forward(blob, inputs, dbg_layer=None, ns=NS):
    """-> (y, z_dbg): network outputs for raw input rows, and the scaled pre-activations of layer dbg_layer.
    ns = 7: the FP64-accurate engine; ns = 4: the screening engine (activations carry 2^-27)."""
    ka = 8 * ns - 5
    net = unpack_blob(blob)
    x = np.asarray(inputs, dtype=np.float64)
    a = ((x - net["xo"]) * net["xg"]) + -1.0
    nhid = len(net["W"]) - 1
    zdbg = None
    for l in range(nhid):
        K = 32 if l == 0 else 64
        ap = np.zeros((a.shape[0], K))
        ap[:, :a.shape[1]] = a
        wd, cs, bs = pack_layer(net["W"][l], net["b"][l], ka + 2 if l == 0 else ka + 3, K, ns)
        z = layer_z(ap, wd, cs, bs, ka - 1 if l == 0 else ka, ns)
        if dbg_layer == l:
            zdbg = z
        a = tansig_scaled(z)
    h = net["W"][-1].shape[1]
    y = a[:, :h] @ net["W"][-1][0] + net["b"][-1][0]
    return ((y - -1.0) / net["y_gain"]) + net["y_xoff"], zdbg
