"""CPU test (gloo, world_size 2) of the multi-GPU host logic: contiguous rank-range shards, local top-k,
all-gather + merge, and the two-round combined strategy. The device engine is replaced by a numpy stand-in built
on the oracle, so this checks only the sharding / collective / merge plumbing of distributed.py."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle import cutsel_oracle as orc

N_VARS, RHO, K = 14, 3, 40


class FakeEngine(object):
    """Numpy stand-in with the Engine methods ShardedSelector uses (score / topk / counts / merge_topk)."""

    big_m = 1000.0

    def __init__(self, lam, obj, base):
        self.lam, self.obj, self.base = lam, obj, base
        self._counts = np.zeros(3, dtype=np.int64)

    def max_pos_nonviolated(self):
        m = (self.obj > 0) & ~(self.lam < -1e-15)
        return float(self.obj[m].max()) if m.any() else -np.inf

    def score(self, vars_values, want):
        pass

    def counts(self):
        return self._counts

    def topk(self, mode, k, pivot_obj=0.0, pivot_idx=0, all_walked=0):
        lam, obj = self.lam, self.obj
        idx = self.base + np.arange(lam.size)
        viol, pos = lam < -1e-15, obj > 0
        self._counts = np.array([lam.size, viol.sum(), (viol & pos).sum()], dtype=np.int64)
        if mode == 1:
            valid, key, key2 = viol, -lam, np.zeros_like(lam)
        elif mode == 2:
            valid, key, key2 = np.ones_like(viol), obj, np.zeros_like(lam)
        elif mode == 3:
            valid, key, key2 = viol & pos, obj, np.zeros_like(lam)
        else:
            walked = np.ones_like(viol) if all_walked else (obj > pivot_obj) | ((obj == pivot_obj) & (idx <= pivot_idx))
            key = obj.copy()
            m = walked & pos
            key[m & viol] = obj[m & viol] + 1000
            key[m & ~viol] = obj[m & ~viol] - 1000
            m = walked & ~pos & viol
            key[m] = -lam[m]
            valid, key2 = np.ones_like(viol), obj
        sel = np.nonzero(valid)[0]
        order = sel[np.lexsort((idx[sel], -key2[sel], -key[sel]))][:k]
        return idx[order], key[order], lam[order], obj[order]

    def merge_topk(self, score, obj2, idx, k):
        obj2 = np.zeros_like(score) if obj2 is None else obj2
        return np.lexsort((idx, -obj2, -score))[:k]


def _scores(huge=False):
    Q_arr, adj = orc.boxqp_arrays(orc.synth_instance(N_VARS, 0.8, seed=7))
    vv = orc.synth_point(N_VARS, seed=8)
    idx = orc.cover_all(N_VARS, RHO)
    lam = orc.lam_min(*(lambda X, x: (x[idx], X[np.array([orc.xarr_inds(N_VARS, r) for r in idx])]))(*orc.split_vars(vv, N_VARS)))
    rng = np.random.default_rng(5)
    obj = rng.normal(size=lam.size)          # any scores do: only the plumbing is under test
    obj[::7] = obj[3]                        # exact ties across shard boundaries
    if huge:                                 # a non-violated candidate whose obj - 1000 still beats the strong ones:
        obj[np.nonzero(~(lam < -1e-15))[0][5]] = 5000.0      # the shortcut of the combined rule must not be taken
    return lam, obj


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sdpcutsel_via_nn_b200 as pkg
    out = {}
    for huge in (False, True):
        lam, obj = _scores(huge)
        N = lam.size
        r0, r1 = pkg.distributed.shard_range(N, world, rank)
        sel = pkg.distributed.ShardedSelector(FakeEngine(lam[r0:r1], obj[r0:r1], r0))
        for strat, k in ((1, K), (2, K), (4, K), (4, 10 * K)):
            r = sel.select(strat, None, k)
            out[(strat, k, huge)] = (r["idx"].tolist(), r["score"].tolist(), int(r["new_strat"]), [int(v) for v in r["counts"]])
    ret[rank] = out
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_selection_matches_single_shard():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0] == ret[1]                                    # identical global selection on every rank
    for huge in (False, True):                                 # huge: the general two-pass path of the combined rule
        lam, obj = _scores(huge)
        order, score = orc.select_feas(lam)
        assert ret[0][(1, K, huge)][0] == order[:K].tolist() and ret[0][(1, K, huge)][1] == score[:K].tolist()
        order, score = orc.select_opt(obj)
        assert ret[0][(2, K, huge)][0] == order[:K].tolist()
        for k in (K, 10 * K):
            kk = min(k, lam.size)
            ns, order, score = orc.select_comb(obj, lam, kk)
            ns2, order2, score2 = orc.select_comb_walk(obj, lam, kk)
            assert ns == ns2 and np.array_equal(order, order2)
            got = ret[0][(4, k, huge)]
            assert got[0] == order[:kk].tolist() and got[1] == score[:kk].tolist() and got[2] == ns


def test_shard_ranges_partition():
    import sdpcutsel_via_nn_b200 as pkg
    for N in (0, 1, 7, 4060, 234531275):
        for w in (1, 2, 3, 8):
            edges = [pkg.distributed.shard_range(N, w, r) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert max(e[1] - e[0] for e in edges) - min(e[1] - e[0] for e in edges) <= 1
