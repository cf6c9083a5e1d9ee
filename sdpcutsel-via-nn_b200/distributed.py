"""Multi-GPU selection: the candidate index space is cut into contiguous rank ranges, one per process/GPU
(torch.distributed, NCCL over NVLink on the B200 box, gloo in CPU tests). Each rank scores its shard and takes a
local top-k; one small all-gather (k + B rows of 4 doubles per rank) plus an identical merge on every rank gives the
global selection. The combined strategy needs the global pivot, hence up to two gather rounds; the
counters ride in the header rows.

The selector returns the k winners AND their guard band (the candidates ranked after the k-th whose score is within the
near-tie guard of the k-th score, sdpcs_last_band) so that ``neartie.resolve`` can reproduce the reference's order where
scores agree to rounding noise.  With one process the band is whatever the device collected (up to band_cap entries);
across processes each rank contributes at most ``BAND_ROWS`` band rows to the exchange.
"""
import os
import time

import numpy as np

BAND_ROWS = 64          # band rows per rank in the exchange (a rank's band only matters when its k-th score is the global one)


def shard_range(N, world, rank):
    return rank * N // world, (rank + 1) * N // world


def shard_cover(engine, world, rank):
    """Restrict the cover `engine` holds (all-subsets, or a pattern-E / list cover set in full on every rank) to this rank's
    contiguous range of candidates (sdpcs_cover_restrict); returns (begin, end)."""
    b, e = shard_range(engine.num_candidates, world, rank)
    engine.cover_restrict(b, e)
    return b, e


_EMPTY_BAND = dict(n_band=0, band_open=0, n_unc_lam=0, n_unc_obj=0)


class ShardedSelector(object):
    """engine: an _capi.Engine (or any object with score/topk/counts/merge_topk) whose cover is this rank's shard."""

    def __init__(self, engine, group=None, device=None, local=False):
        self.eng = engine
        self.group = group
        self.device = device
        self.dist = None
        if not local:                      # local: one process owns the whole cover (the drop-in CutSolver)
            try:
                import torch.distributed as dist
                self.dist = dist if dist.is_available() and dist.is_initialized() else None
            except ImportError:
                self.dist = None
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.prof = {} if os.environ.get("SDPCS_DIST_PROFILE") else None     # phase -> accumulated seconds (host clock)
        # device-resident exchange: pack on the GPU, NCCL all-gather of device buffers, merge on the GPU (no host hop);
        # engines without the entry points (CPU stand-ins in the gloo tests) use the host-staged exchange below
        self.dev_exchange = (self.world > 1 and device is not None and getattr(device, "type", "cuda") == "cuda"
                             and hasattr(engine, "topk_pack_dev") and not os.environ.get("SDPCS_HOST_EXCHANGE"))
        self._bufs = {}
        self._xstream = None
        if self.dev_exchange:
            # engine kernels and the NCCL collective must be ordered on ONE stream: torch's current stream if it is a real
            # one, else (legacy default stream, handle 0 -- which the C ABI reads as "use the context's own stream") a
            # stream of our own.  Set once, before any scoring, so that no work is left behind on another stream.
            import torch
            cur = torch.cuda.current_stream(device)
            self._xstream = cur if cur.cuda_stream != 0 else torch.cuda.Stream(device=device)
            engine.set_stream(self._xstream.cuda_stream)

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

    def _topk_exchange(self, mode, k, pivot=(0.0, 0, 0), want_counts=True):
        """Local top-k of `mode` + exchange + merge.  Returns (winners (idx, score, lam, obj), band dict, summed counters,
        max over ranks of the largest positive non-violated obj)."""
        eng = self.eng
        if not self.dev_exchange:
            top = eng.topk(mode, k, *pivot)
            band = self._band()
            mpn = eng.max_pos_nonviolated() if (mode == 3 and hasattr(eng, "max_pos_nonviolated")) else (np.inf if mode == 3 else 0.0)
            return self._gather_merge(mode, k, top, band, mode == 4, eng.counts() if want_counts else None, mpn)
        import torch
        rows_cap = k + BAND_ROWS
        if rows_cap not in self._bufs:
            with torch.cuda.stream(self._xstream):
                self._bufs = {rows_cap: (torch.empty((2 + rows_cap) * 4, dtype=torch.float64, device=self.device),
                                         torch.empty(self.world * (2 + rows_cap) * 4, dtype=torch.float64, device=self.device))}
        send, recv = self._bufs[rows_cap]
        with torch.cuda.stream(self._xstream):                                   # the collective is ordered on the engine's stream
            eng.topk_pack_dev(mode, k, BAND_ROWS, send.data_ptr(), *pivot)
            self.dist.all_gather_into_tensor(recv, send, group=self.group)
            idx, sc, lam, obj, nwin, nband, hdr = eng.merge_packed_dev(recv.data_ptr(), self.world, rows_cap, k, mode == 4,
                                                                       self._guard_of(mode), k + self.world * BAND_ROWS)
        band = dict(idx=idx[nwin:], score=sc[nwin:], lam=lam[nwin:], obj=obj[nwin:], n_band=int(nband), band_open=int(hdr[7]),
                    n_unc_lam=int(hdr[5]), n_unc_obj=int(hdr[6]))
        counts = np.array([int(hdr[1]), int(hdr[2]), int(hdr[3])], dtype=np.int64) if want_counts else None
        return (idx[:nwin], sc[:nwin], lam[:nwin], obj[:nwin]), band, counts, float(hdr[4])

    def _band(self):
        """Guard band + counters of the engine's last top-k pass (engines without a guard report none)."""
        if not hasattr(self.eng, "last_band"):
            z = np.zeros(0)
            return dict(idx=z.astype(np.int64), score=z, lam=z, obj=z, **_EMPTY_BAND)
        return self.eng.last_band(None if self.world == 1 else BAND_ROWS)

    def _guard_of(self, mode):
        p = getattr(self.eng, "params", None)
        if p is None:
            return 0.0
        return float(p.guard_lam if mode == 1 else p.guard_obj)

    def _gather_merge(self, mode, k, top, band, use_obj2, counts=None, extra=0.0):
        """All-gather of the local lists and the identical merge on every rank.  `top` = (idx, score, lam, obj) of the
        local winners, `band` the local guard band.  Returns (winners, band dict, summed counters or None, max of extra).
        Row 0 of a rank's block carries the list lengths and, when given, the rank's three counters (they ride along
        instead of needing their own all-reduce); row 1 `extra` (reduced with max) and the guard counters."""
        idx, score, lam, obj = top
        if self.world == 1:
            return top, band, (np.asarray(counts, dtype=np.int64) if counts is not None else None), float(extra)
        m, mb = idx.shape[0], band["idx"].shape[0]
        pack = np.zeros((k + 2 + BAND_ROWS, 4))
        pack[0, 0] = m + mb
        if counts is not None:
            pack[0, 1:4] = counts          # < 2^44: exact in float64
        pack[1] = (extra, band["n_unc_lam"], band["n_unc_obj"], 1.0 if (band["band_open"] or band["n_band"] > mb) else 0.0)
        rows = pack[2:2 + m + mb]
        rows[:, 0] = np.concatenate([idx, band["idx"]])     # agg_idx < 2^44: exact in float64
        rows[:, 1] = np.concatenate([score, band["score"]])
        rows[:, 2] = np.concatenate([lam, band["lam"]])
        rows[:, 3] = np.concatenate([obj, band["obj"]])
        allp = self._allgather(pack)
        tot = allp[:, 0, 1:4].sum(axis=0).astype(np.int64) if counts is not None else None
        ext = float(allp[:, 1, 0].max())
        lens = [int(allp[r, 0, 0]) for r in range(self.world)]
        rows = np.concatenate([allp[r, 2:lens[r] + 2] for r in range(self.world)], axis=0)
        gb = dict(n_unc_lam=int(allp[:, 1, 1].sum()), n_unc_obj=int(allp[:, 1, 2].sum()), band_open=0, n_band=0)
        z = np.zeros(0)
        if rows.shape[0] == 0:
            gb.update(idx=z.astype(np.int64), score=z, lam=z, obj=z)
            return (z.astype(np.int64), z, z, z), gb, tot, ext
        gidx = rows[:, 0].astype(np.int64)
        perm = self.eng.merge_topk(rows[:, 1], rows[:, 3] if use_obj2 else None, gidx, rows.shape[0])
        kk = min(k, perm.size)
        win = perm[:kk]
        # global band: merged entries after the k-th within the guard of the k-th score; it is complete unless a rank whose
        # list ends inside the band had more near ties than it could send
        delta = self._guard_of(mode)
        rest = perm[kk:]
        if kk > 0 and rest.size:
            sk = rows[win[-1], 1]
            inb = rest[rows[rest, 1] >= sk - delta]
            for r in range(self.world):
                if lens[r] and allp[r, 1, 3] and allp[r, 2 + lens[r] - 1, 1] >= sk - delta:
                    gb["band_open"] = 1
        else:
            inb = rest[:0]
            gb["band_open"] = int(allp[:, 1, 3].max()) if kk < k else 0
        gb.update(idx=gidx[inb], score=rows[inb, 1], lam=rows[inb, 2], obj=rows[inb, 3], n_band=int(inb.size))
        return (gidx[win], rows[win, 1], rows[win, 2], rows[win, 3]), gb, tot, ext

    # -- public ------------------------------------------------------------------------------------
    def select(self, strat, vars_values, k, n_total=None):
        """Global selection over all shards. Returns dict(idx, score, lam, obj, counts, new_strat, band, guard, strat, path,
        pivot) -- identical on every rank. vars_values None = LP point already resident on the device.  `band` holds the
        near ties that follow the k winners, `guard` = dict(n_band, band_open, n_unc_lam, n_unc_obj), `path` = the top-k
        mode that produced the list (3: strong-prefix shortcut of the combined rule, scores already + big_m)."""
        eng = self.eng
        k = int(k)
        if strat not in (1, 2, 3, 4):
            raise ValueError("strat must be 1, 2, 3 or 4")
        t = time.perf_counter()
        eng.score(vars_values, {1: 1, 2: 2, 3: 4, 4: 3}[strat])      # want bits: lam, NN measure, exact SDP measure

        def pack(top, band, counts, new_strat, path, pivot=None):
            gi, gs, gl, go = top
            guard = {key: int(band[key]) for key in ("n_band", "band_open", "n_unc_lam", "n_unc_obj")}
            return dict(idx=gi, score=gs, lam=gl, obj=go, counts=counts, new_strat=new_strat, strat=strat, path=path, pivot=pivot,
                        band={key: band[key] for key in ("idx", "score", "lam", "obj")}, guard=guard)

        if strat != 4:
            mode = 2 if strat == 3 else strat                          # strat 3 ranks like strat 2, by the exact measure
            top, band, counts, _ = self._topk_exchange(mode, k)
            self._t("topk+exchange+merge", t)
            return pack(top, band, counts, strat, mode)
        # the pivot of the combined rule needs k <= N; N is only known after the exchange, so gather k rows and cut after
        (si, ss, sl, so), band, counts, mpn = self._topk_exchange(3, k)
        N, n_viol, n_strong = (int(v) for v in counts)
        k = min(k, N)
        si, sl, so = si[:k], sl[:k], so[:k]
        t = self._t("topk1+exchange+merge", t)
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
            return pack((si, so + big_m, sl, so), band, counts, new_strat, 3, (pobj, pidx, False))
        top, band2, _, _ = self._topk_exchange(4, k, (pobj, pidx, 1 if all_walked else 0), want_counts=False)
        band2["n_unc_lam"], band2["n_unc_obj"] = band["n_unc_lam"], band["n_unc_obj"]
        self._t("topk2+exchange+merge", t)
        return pack(top, band2, counts, new_strat, 4, (pobj, pidx, bool(all_walked)))


class ScreenedSelector(object):
    """Screen-and-refine selection for the strategies that rank by the NN measure (2 and 4).

    Tier 1 (screen): `engine` scores EVERY candidate of the shard(s) with the 4-digit tcgen05 engine (SDPCS_NN_SCREEN: NN
    outputs accurate to ~3e-6, ten digit pairs instead of 28, two TMEM stages) and selects with a guard band of
    `screen_guard` on the measure: the k winners plus every candidate within the guard of the k-th score come back
    (globally merged when sharded).  Tier 2 (refine): those contenders -- a few thousand of 2.3e8 -- become the list cover
    of `refine_engine` (FP64-accurate 7-digit engine, same instance and weights) and the SAME selection runs on them;
    its result carries the usual 1e-9 guard band for tier 3, the reference-arithmetic re-score of neartie.resolve.
    The selection is the exact engine's provided no candidate's screening error exceeds screen_guard / 2; the guard is
    ~40 x the largest error observed (2.6e-6 on the NN output), and every call checks the contenders' actual errors:
    if the largest exceeds screen_guard / 8, or the band did not fit, or the two tiers disagree on the path of the
    combined rule, the call falls back to one exact pass over everything (`fallbacks` counts them).
    sets_of: global agg_idx array -> (m, rho) index rows (-1 padded), to build the refine cover."""

    def __init__(self, engine, refine_engine, sets_of, rho, screen_guard, exact_guard, group=None, device=None, local=False):
        from . import _capi
        self._capi = _capi
        self.eng, self.fine_eng, self.sets_of, self.rho = engine, refine_engine, sets_of, int(rho)
        self.screen_guard, self.exact_guard = float(screen_guard), float(exact_guard)
        self.coarse = ShardedSelector(engine, group=group, device=device, local=local)
        self.fine = ShardedSelector(refine_engine, local=True)
        self.fallbacks = 0
        self.last = {}
        self._fine_has_point = False
        engine.set_params(nn_engine=_capi.NN_SCREEN, guard_obj=self.screen_guard)
        refine_engine.set_params(nn_engine=_capi.NN_TCGEN05, guard_obj=self.exact_guard)

    @property
    def prof(self):
        return self.coarse.prof

    def _exact_pass(self, strat, vars_values, k):
        self.fallbacks += 1
        self.eng.set_params(nn_engine=self._capi.NN_TCGEN05, guard_obj=self.exact_guard)
        try:
            return self.coarse.select(strat, vars_values, k)
        finally:
            self.eng.set_params(nn_engine=self._capi.NN_SCREEN, guard_obj=self.screen_guard)

    def select(self, strat, vars_values, k):
        if strat not in (2, 4):
            return self.coarse.select(strat, vars_values, k)         # no NN measure involved: nothing to screen
        raw = self.coarse.select(strat, vars_values, k)
        if raw["guard"]["band_open"]:
            return self._exact_pass(strat, None, k)
        cand = np.concatenate([raw["idx"], raw["band"]["idx"]])
        approx = np.concatenate([raw["obj"], raw["band"]["obj"]])
        order = np.argsort(cand, kind="stable")                       # ascending agg_idx: local position order == global tie-break
        cand, approx = cand[order], approx[order]
        self.last = dict(contenders=int(cand.size), screen_band=int(raw["band"]["idx"].size))
        if cand.size == 0:
            return raw
        self.fine_eng.set_cover_list(self.rho, self.sets_of(cand))
        vv_fine = vars_values if (vars_values is not None or self._fine_has_point) else None
        if vars_values is None and not self._fine_has_point:
            raise ValueError("the refine engine has no LP point yet: pass vars_values on the first call")
        res = self.fine.select(strat, vv_fine, k)
        self._fine_has_point = True
        _, exact = self.fine_eng.scores(lam=False)
        err = float(np.abs(exact - approx).max())
        self.last.update(max_screen_error=err)
        unsure_walk = strat == 4 and res["path"] != raw["path"]       # the tiers disagree on how the combined rule was walked
        if err > self.screen_guard / 8 or unsure_walk:
            return self._exact_pass(strat, None, k)
        # back to global candidate indices; counters and the strategy switch are global quantities of tier 1
        out = dict(res)
        out["idx"] = cand[res["idx"]]
        out["band"] = dict(res["band"], idx=cand[res["band"]["idx"]])
        if res.get("pivot") is not None:
            out["pivot"] = (res["pivot"][0], int(cand[res["pivot"][1]]) if cand.size else 0, res["pivot"][2])
        out["counts"], out["new_strat"] = raw["counts"], raw["new_strat"]
        g = dict(res["guard"])
        g["n_unc_lam"] = raw["guard"]["n_unc_lam"]
        out["guard"] = g
        return out
