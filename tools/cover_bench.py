#!/usr/bin/env python
"""Times the device-built pattern-E cover (sdpcs_set_cover_pattern) against the package's host DFS (needs a B200).
    python tools/cover_bench.py [n density rho ...]   default: spar125-075-like graphs, rho = 3, 4, 5"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdpcutsel_via_nn_b200 as pkg  # noqa: E402


def main():
    cases = [(125, 0.75, 3), (125, 0.75, 4), (125, 0.75, 5), (250, 0.2, 5)]
    if len(sys.argv) > 3:
        a = sys.argv[1:]
        cases = [(int(a[i]), float(a[i + 1]), int(a[i + 2])) for i in range(0, len(a), 3)]
    eng = pkg._capi.Engine(0)
    for n, dens, rho in cases:
        rng = np.random.default_rng(7)
        adj = np.triu(rng.random((n, n)) < dens, 1)
        adj = (adj | adj.T).astype(np.uint8)
        eng.set_instance(n, np.zeros(n * (n + 1) // 2))
        eng.set_cover_pattern(rho, adj)                       # warm-up (allocations)
        t0 = time.perf_counter()
        N = eng.set_cover_pattern(rho, adj)
        t1 = time.perf_counter()
        rows = eng.cover_rows()
        t2 = time.perf_counter()
        line = "n=%d density=%.2f rho=%d: N=%d  device build %.1f ms (+ %.1f ms to download the %d x %d index rows)" % (
            n, dens, rho, N, (t1 - t0) * 1e3, (t2 - t1) * 1e3, N, rho)
        if N <= 4000000:
            t0 = time.perf_counter()
            want = pkg.cover.pattern_E(adj, rho)
            th = time.perf_counter() - t0
            line += "; host DFS %.2f s, identical: %s" % (th, bool(np.array_equal(rows, want)))
        print(line, flush=True)


if __name__ == "__main__":
    main()
