// Multi-GPU exchange on the device: every rank packs its local winners (+ guard band) into rows of four
// doubles, the ranks all-gather the packed blocks over NVLink (NCCL, driven by the caller on the context's stream), and
// every rank merges the gathered sorted runs with the selection comparator (score desc, obj2 desc, agg_idx asc).
// Nothing but the final k (+ band) rows and one header goes back to the host.
//
// Block of one rank = (2 + rows_cap) x 4 doubles:
//   row 0: [len (rows that follow), N_local, n_violated, n_strong]
//   row 1: [extra (max obj among positive non-violated, -inf if none), n_unc_lam, n_unc_obj, open (1: the rank had more
//           near ties than it could send)]
//   rows 2..: [agg_idx, score, lam, obj], in selection order (winners, then band)
#pragma once
#include "device_math.cuh"

namespace sdpcs {

struct PackHdr {
    double len, n_local, n_violated, n_strong, extra, n_unc_lam, n_unc_obj, open;
};

__global__ void __launch_bounds__(256) k_pack_rows(PackHdr h, i64 rows, i64 rows_cap, const i64* s_idx, const double* o_score,
                                                   const double* o_lam, const double* o_obj, double* out)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        out[0] = h.len; out[1] = h.n_local; out[2] = h.n_violated; out[3] = h.n_strong;
        out[4] = h.extra; out[5] = h.n_unc_lam; out[6] = h.n_unc_obj; out[7] = h.open;
    }
    if (i < rows_cap) {
        double4 r = make_double4(0.0, 0.0, 0.0, 0.0);
        if (i < rows) r = make_double4((double)s_idx[i], o_score[i], o_lam[i], o_obj[i]);   // agg_idx < 2^44: exact in FP64
        reinterpret_cast<double4*>(out + 8)[i] = r;
    }
}

// a before b in selection order?
__device__ __forceinline__ bool row_before(u64 a1, u64 a2, double ai, u64 b1, u64 b2, double bi)
{
    return (a1 > b1) || (a1 == b1 && (a2 > b2 || (a2 == b2 && ai < bi)));
}

// One thread per gathered row: its place in the merged order = sum over the runs of the rows that come before it
// (binary search, the runs are sorted).  Rows with place < out_cap are written to out[place].
__global__ void __launch_bounds__(256) k_merge_rows(const double* gathered, int world, i64 rows_cap, int use_obj2, i64 out_cap, double* out)
{
    const i64 block = 8 + 4 * rows_cap;
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = (int)(t / rows_cap);
    const i64 i = t - (i64)r * rows_cap;
    if (r >= world) return;
    const double* mine = gathered + (i64)r * block;
    if (i >= (i64)mine[0]) return;
    const double4 me = reinterpret_cast<const double4*>(mine + 8)[i];
    const u64 m1 = enc_key(me.y), m2 = use_obj2 ? enc_key(me.w) : 0;
    i64 place = 0;
    for (int q = 0; q < world; ++q) {
        const double* run = gathered + (i64)q * block;
        const double4* rows = reinterpret_cast<const double4*>(run + 8);
        i64 lo = 0, hi = (i64)run[0];               // first position whose row does NOT come before me
        while (lo < hi) {
            const i64 mid = (lo + hi) >> 1;
            const double4 o = rows[mid];
            if (row_before(enc_key(o.y), use_obj2 ? enc_key(o.w) : 0, o.x, m1, m2, me.x)) lo = mid + 1; else hi = mid;
        }
        place += lo;
    }
    if (place < out_cap) reinterpret_cast<double4*>(out + 16)[place] = me;
}

// Header of the merged list (16 doubles): [n_rows, sum N, sum violated, sum strong, max extra, sum unc_lam, sum unc_obj,
// band_open, n_winners, n_band, 0...].  Winners = first min(k, total); band = rows after them within delta of the k-th score;
// the band is open when a rank that flagged more near ties ends its block inside the band.
__global__ void k_merge_finish(const double* gathered, int world, i64 rows_cap, i64 k, double delta, i64 out_cap, double* out)
{
    if (threadIdx.x || blockIdx.x) return;
    const i64 block = 8 + 4 * rows_cap;
    double tot = 0, sN = 0, sV = 0, sS = 0, ext = -INFINITY, ul = 0, uo = 0;
    for (int r = 0; r < world; ++r) {
        const double* h = gathered + (i64)r * block;
        tot += h[0]; sN += h[1]; sV += h[2]; sS += h[3]; ext = fmax(ext, h[4]); ul += h[5]; uo += h[6];
    }
    i64 total = (i64)tot;
    if (total > out_cap) total = out_cap;
    const i64 nwin = k < total ? k : total;
    const double4* rows = reinterpret_cast<const double4*>(out + 16);
    i64 nband = 0;
    double open = 0.0;
    if (nwin > 0 && delta >= 0.0) {
        const double sk = rows[nwin - 1].y;
        while (nwin + nband < total && rows[nwin + nband].y >= sk - delta) ++nband;
        for (int r = 0; r < world; ++r) {
            const double* h = gathered + (i64)r * block;
            const i64 len = (i64)h[0];
            if (h[7] != 0.0 && len > 0 && reinterpret_cast<const double4*>(h + 8)[len - 1].y >= sk - delta) open = 1.0;
        }
        if ((i64)tot > out_cap && nwin + nband == total) open = 1.0;
    } else if (nwin < k) {
        for (int r = 0; r < world; ++r) open = fmax(open, gathered[(i64)r * block + 7]);
    }
    out[0] = (double)total; out[1] = sN; out[2] = sV; out[3] = sS; out[4] = ext; out[5] = ul; out[6] = uo; out[7] = open;
    out[8] = (double)nwin; out[9] = (double)nband;
    for (int j = 10; j < 16; ++j) out[j] = 0.0;
}

}  // namespace sdpcs
