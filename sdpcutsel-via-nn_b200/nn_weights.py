"""NN_rhoD weight sources -> flat FP64 blob consumed by ``sdpcs_set_weights``.

Replaces the reference's MATLAB-Coder binary ``neural_nets/NNs.so`` (bound at
``cut_select_qp.py:284-303``): the weights are read from the genFunction text
``neural_nets/neural_net_{2..5}D.m`` (constants at ``neural_net_3D.m:9-32``,
``neural_net_5D.m:9-36``) or from the MATLAB training checkpoints
``training_checkpoint_neural_net_{2,4}D.mat`` and evaluated on the GPU.

Blob layout (all float64):
    [n_in, n_layers, n_hidden,
     x_xoffset(n_in), x_gain(n_in),
     W_1 (h x n_in, row-major), b_1 (h), W_2 (h x h), b_2 (h), ...,
     W_L (1 x h), b_L (1),
     y_gain, y_xoffset]
``n_layers`` counts the linear output layer; ``ymin = -1`` in every net.
"""
import os
import re
import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
PACKED_PATH = os.path.join(_PKG_DIR, "weights", "neural_nets.npz")

_ASSIGN = re.compile(r"^\s*([A-Za-z_][\w\.]*)\s*=\s*(\[.*?\]|[-+0-9.eE]+)\s*;", re.M | re.S)


def _parse_matlab_literal(txt):
    txt = txt.strip()
    if not txt.startswith("["):
        return np.array([[float(txt)]])
    rows = txt[1:-1].split(";")
    return np.array([[float(v) for v in r.replace(",", " ").split()] for r in rows if r.strip()])


def parse_m_file(path):
    """Parse a genFunction ``neural_net_dD.m`` into a dict of arrays."""
    with open(path) as f:
        src = f.read()
    src = src.split("% ===== SIMULATION")[0]
    vals = {m.group(1): _parse_matlab_literal(m.group(2)) for m in _ASSIGN.finditer(src)}
    Ws, bs = [vals["IW1_1"]], [vals["b1"].ravel()]
    layer = 2
    while "LW%d_%d" % (layer, layer - 1) in vals:
        Ws.append(vals["LW%d_%d" % (layer, layer - 1)])
        bs.append(vals["b%d" % layer].ravel())
        layer += 1
    assert vals["x1_step1.ymin"].item() == -1.0 and vals["y1_step1.ymin"].item() == -1.0
    return dict(W=Ws, b=bs,
                x_xoffset=vals["x1_step1.xoffset"].ravel(), x_gain=vals["x1_step1.gain"].ravel(),
                y_gain=vals["y1_step1.gain"].item(), y_xoffset=vals["y1_step1.xoffset"].item())


def load_mat_checkpoint(path):
    """Read ``checkpoint.net`` (IW, LW, b, mapminmax settings) from a MATLAB training checkpoint."""
    import scipy.io
    m = scipy.io.loadmat(path, squeeze_me=False, struct_as_record=False)
    net = m["checkpoint"][0, 0].net[0, 0]
    L = int(net.numLayers[0, 0])
    Ws = [np.asarray(net.IW[0, 0], dtype=np.float64)]
    for l in range(1, L):
        Ws.append(np.asarray(net.LW[l, l - 1], dtype=np.float64))
    bs = [np.asarray(net.b[l, 0], dtype=np.float64).ravel() for l in range(L)]
    xs = net.inputs[0, 0][0, 0].processSettings[0, 0][0, 0]
    ys = net.outputs[0, L - 1][0, 0].processSettings[0, 0][0, 0]
    return dict(W=Ws, b=bs,
                x_xoffset=np.asarray(xs.xoffset, float).ravel(), x_gain=np.asarray(xs.gain, float).ravel(),
                y_gain=float(ys.gain[0, 0]), y_xoffset=float(ys.xoffset[0, 0]))


def pack_blob(net):
    Ws, bs = net["W"], net["b"]
    n_in, h, L = Ws[0].shape[1], Ws[0].shape[0], len(Ws)
    assert Ws[-1].shape == (1, h) and all(W.shape == (h, h) for W in Ws[1:-1])
    parts = [np.array([n_in, L, h], dtype=np.float64), net["x_xoffset"], net["x_gain"]]
    for W, b in zip(Ws, bs):
        parts += [np.ascontiguousarray(W, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()]
    parts.append(np.array([net["y_gain"], net["y_xoffset"]], dtype=np.float64))
    return np.concatenate([np.asarray(p, dtype=np.float64).ravel() for p in parts])


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
    y_gain, y_xoffset = blob[o], blob[o + 1]
    assert o + 2 == blob.size, "malformed weight blob"
    return dict(W=Ws, b=bs, x_xoffset=xo, x_gain=xg, y_gain=float(y_gain), y_xoffset=float(y_xoffset))


def load_packed(rho, path=None):
    """Blob of NN_rhoD from the packed archive shipped with the package."""
    with np.load(path or PACKED_PATH) as z:
        return np.array(z["nn%dD" % rho], dtype=np.float64)


def pack_from_reference(nn_dir, out_path=PACKED_PATH):
    """One-time conversion: ``neural_nets/*.m`` -> ``weights/neural_nets.npz`` (run where the .m files exist)."""
    blobs = {"nn%dD" % d: pack_blob(parse_m_file(os.path.join(nn_dir, "neural_net_%dD.m" % d))) for d in (2, 3, 4, 5)}
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    np.savez_compressed(out_path, **blobs)
    return blobs


if __name__ == "__main__":
    import sys
    b = pack_from_reference(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/neural_nets")
    print({k: v.size for k, v in b.items()})
