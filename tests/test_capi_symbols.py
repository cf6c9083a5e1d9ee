"""CPU tests: the C-ABI library builds for sm_100a, loads, and exports every symbol include/sdpcutsel.h declares
(no compute calls here -- there is no GPU in this container)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import sdpcutsel_via_nn_b200 as pkg
    from sdpcutsel_via_nn_b200 import build
    build.build()
    return pkg._capi.load_library()


def test_header_symbols_exported(lib):
    import sdpcutsel_via_nn_b200 as pkg
    hdr = open(os.path.join(ROOT, "include", "sdpcutsel.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(sdpcs_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    for nm in names:
        assert hasattr(lib, nm), "missing export " + nm
    assert sorted(pkg._capi.EXPORTS) == names


def test_host_utilities_without_gpu(lib):
    import sdpcutsel_via_nn_b200 as pkg
    import itertools
    for n, rho in [(12, 3), (9, 4), (8, 5), (7, 2)]:
        want = np.array(list(itertools.combinations(range(n), rho)))
        assert np.array_equal(pkg._capi.unrank(n, rho, np.arange(want.shape[0])), want)
    assert pkg._capi.binom(125, 5) == 234531275 and pkg._capi.binom(125, 4) == 9691375
    with pytest.raises(pkg._capi.SdpcsError):
        pkg._capi.unrank(10, 3, [120])
    # triangle rows (cut_select_qp.py:846-860): every rank of small n, both ends of n = 250, all four inequality types
    for n in (3, 4, 9, 31, 250):
        T = pkg._capi.binom(n, 3)
        ranks = np.arange(T) if T < 5000 else np.concatenate([np.arange(3000), np.arange(T - 3000, T), np.arange(0, T, 997)])
        for t in range(4):
            csr = pkg._capi.triangle_rows_csr(n, ranks, np.full(ranks.size, t, np.int8))
            i1, i2, i3 = pkg._capi.unrank(n, 3, ranks).astype(np.int64).T
            ptr, L = csr["rowptr"][:-1], n * (n + 1) // 2
            assert np.array_equal(np.diff(csr["rowptr"]), np.full(ranks.size, 6 if t == 3 else 4))
            assert np.array_equal(csr["ind"][ptr], n * i1 - i1 * (i1 + 1) // 2 + i2)
            assert np.array_equal(csr["ind"][ptr + 1], n * i1 - i1 * (i1 + 1) // 2 + i3)
            assert np.array_equal(csr["ind"][ptr + 2], n * i2 - i2 * (i2 + 1) // 2 + i3)
            assert np.array_equal(csr["ind"][ptr + 3], (i1, i2, i3, i1)[t] + L)


def test_no_cpu_fallback_without_device(lib):
    """Without a CUDA device the product refuses to run (it must not silently compute on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sdpcutsel_via_nn_b200 as pkg
    with pytest.raises(pkg._capi.SdpcsError):
        pkg._capi.Engine(0)


def test_params_struct_layout(lib):
    import sdpcutsel_via_nn_b200 as pkg
    p = pkg._capi.Params()
    assert lib.sdpcs_default_params(ctypes.byref(p)) == 0
    assert (p.thres_min_opt, p.thres_neg_eigval, p.big_m, p.thres_tri_viol, p.thres_tri_dense) == (0.0, -1e-15, 1000.0, 1e-7, 2)
