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
