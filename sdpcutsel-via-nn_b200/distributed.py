"""Multi-GPU selection: the candidate index space is cut into contiguous rank ranges, one per process/GPU
(torch.distributed, NCCL over NVLink on the B200 box, gloo in CPU tests). Each rank scores its shard and takes a
local top-k; one small all-gather (k x 4 doubles per rank) plus an identical merge on every rank gives the
global selection. The combined strategy needs the global pivot, hence two gather rounds and one
all-reduce of three counters.
"""
import os
import time

import numpy as np


def shard_range(N, world, rank):
    return rank * N // world, (rank + 1) * N // world


def shard_cover(engine, world, rank):
    """Restrict the cover `engine` holds (all-subsets, or a pattern-E / list cover set in full on every rank) to this rank's
    contiguous range of candidates (sdpcs_cover_restrict); returns (begin, end)."""
    b, e = shard_range(engine.num_candidates, world, rank)
    engine.cover_restrict(b, e)
    return b, e


class ShardedSelector(object):
    """engine: an _capi.Engine (or any object with score/topk/counts/merge_topk) whose cover is this rank's shard."""

    def __init__(self, engine, group=None, device=None):
        self.eng = engine
        self.group = group
        self.device = device
        try:
            import torch.distributed as dist
            self.dist = dist if dist.is_available() and dist.is_initialized() else None
        except ImportError:
            self.dist = None
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.prof = {} if os.environ.get("SDPCS_DIST_PROFILE") else None     # phase -> accumulated seconds (host clock)

    def _t(self, name, t0):
        if self.prof is not None:
            self.prof[name] = self.prof.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    # -- collectives on small host arrays ------------------------------------------------------------
    def _allgather(self, arr):
        """arr: float64 (m, c) with the same shape on every rank -> (world, m, c)."""
        if self.world == 1:
            return arr[None]
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if self.device is not None:
            t = t.to(self.device)
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t, group=self.group)
        return out.cpu().numpy().reshape((self.world,) + tuple(t.shape))

    def _allreduce_sum(self, arr):
        if self.world == 1:
            return arr
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if self.device is not None:
            t = t.to(self.device)
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def _gather_merge(self, k, idx, score, lam, obj, use_obj2, counts=None, extra=0.0):
        """All-gather of the local lists (k + 2 rows of 4 doubles per rank) and the identical merge on every rank.
        Row 0 carries the list length and, when given, the rank's three counters (they ride along instead of needing
        their own all-reduce), row 1 one more statistic (`extra`, reduced with max); returns (idx, score, lam, obj,
        summed counters or None, max of extra)."""
        m = idx.shape[0]
        pack = np.zeros((k + 2, 4))
        pack[0, 0] = m
        if counts is not None:
            pack[0, 1:4] = counts          # < 2^44: exact in float64
        pack[1, 0] = extra
        pack[2:m + 2, 0] = idx            # agg_idx < 2^44: exact in float64
        pack[2:m + 2, 1] = score
        pack[2:m + 2, 2] = lam
        pack[2:m + 2, 3] = obj
        allp = self._allgather(pack)
        tot = allp[:, 0, 1:4].sum(axis=0).astype(np.int64) if counts is not None else None
        ext = float(allp[:, 1, 0].max())
        rows = np.concatenate([allp[r, 2:int(allp[r, 0, 0]) + 2] for r in range(self.world)], axis=0)
        if rows.shape[0] == 0:
            z = np.zeros(0)
            return z.astype(np.int64), z, z, z, tot, ext
        gidx = rows[:, 0].astype(np.int64)
        if self.world == 1:
            perm = np.arange(min(k, rows.shape[0]))
        else:
            perm = self.eng.merge_topk(rows[:, 1], rows[:, 3] if use_obj2 else None, gidx, k)
        return gidx[perm], rows[perm, 1], rows[perm, 2], rows[perm, 3], tot, ext

    # -- public ------------------------------------------------------------------------------------
    def select(self, strat, vars_values, k, n_total=None):
        """Global selection over all shards. Returns dict(idx, score, lam, obj, counts, new_strat) -- identical
        on every rank. vars_values None = LP point already resident on the device."""
        eng = self.eng
        k = int(k)
        if strat not in (1, 2, 4):
            raise ValueError("strat must be 1, 2 or 4")
        t = time.perf_counter()
        eng.score(vars_values, 1 if strat == 1 else 2 if strat == 2 else 3)
        if strat != 4:
            idx, sc, lam, obj = eng.topk(strat, k)
            t = self._t("score+topk", t)
            gi, gs, gl, go, counts, _ = self._gather_merge(k, idx, sc, lam, obj, False, eng.counts())
            self._t("exchange+merge", t)
            return dict(idx=gi, score=gs, lam=gl, obj=go, counts=counts, new_strat=strat)
        idx, sc, lam, obj = eng.topk(3, k)
        t = self._t("score+topk1", t)
        # the pivot of the combined rule needs k <= N; N is only known after the exchange, so gather k rows and cut after
        mpn = eng.max_pos_nonviolated() if hasattr(eng, "max_pos_nonviolated") else np.inf
        si, ss, sl, so, counts, mpn = self._gather_merge(k, idx, sc, lam, obj, False, eng.counts(), mpn)
        N, n_viol, n_strong = (int(v) for v in counts)
        k = min(k, N)
        si, sl, so = si[:k], sl[:k], so[:k]
        t = self._t("gather+merge1", t)
        all_walked = n_strong < k or k == 0
        pobj, pidx = (0.0, 0) if all_walked else (float(so[k - 1]), int(si[k - 1]))
        big_m = float(getattr(eng, "big_m", 1000.0))
        strong = min(n_strong, k)
        viol_walked = n_viol if all_walked else k
        new_strat = 4
        if k > 0 and N > 0:
            new_strat = 1 if strong / k < viol_walked / N else 4
        counts = np.array([N, viol_walked, strong])
        if not all_walked and big_m > 0 and pobj < pobj + big_m and mpn - big_m < pobj + big_m:
            # the k strong elements up to the pivot are re-scored obj + big_m and nothing else can reach them
            # (combined_is_strong_prefix in capi.cu): the merged strong list is the answer, no second pass
            return dict(idx=si, score=so + big_m, lam=sl, obj=so, counts=counts, new_strat=new_strat)
        idx, sc, lam, obj = eng.topk(4, k, pobj, pidx, 1 if all_walked else 0)
        t = self._t("topk2", t)
        gi, gs, gl, go, _, _ = self._gather_merge(k, idx, sc, lam, obj, True)
        self._t("gather+merge2", t)
        return dict(idx=gi, score=gs, lam=gl, obj=go, counts=counts, new_strat=new_strat)
