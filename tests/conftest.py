import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def blobs():
    import sdpcutsel_via_nn_b200 as pkg
    return {d: pkg.nn_weights.load_packed(d) for d in (2, 3, 4, 5)}


def inst_arrays(golden, name):
    from oracle import cutsel_oracle as orc
    Qf = golden["inst_%s_Q" % name.replace("-", "_")].astype(np.float64)
    Q_arr, adj = orc.boxqp_arrays(Qf)
    return Qf.shape[0], Q_arr, adj
