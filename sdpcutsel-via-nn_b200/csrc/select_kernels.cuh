// K5: deterministic top-k over resident scores, with a near-tie guard band.
// Replaces rank_list.sort(key=itemgetter(1), reverse=True) + prefix slicing (cut_select_qp.py:601, 625, 653-654,
// consumer bound :718) and the triangle sort (:841-844).  Python's stable descending sort == order by
// (score desc, original position asc); for the combined rule the "original position" of the second sort is the
// order of the first one, i.e. (obj desc, agg_idx asc).  Both are expressed as a 3-level key
//     (key1 desc, key2 desc, idx asc)        key = order-preserving u64 image of the FP64 score
//
// Keys are never materialised: every pass recomputes them from the resident lam / obj arrays (8 or 16 B read per
// candidate and pass, nothing written).  An MSD radix select (11-bit digits) narrows key1; as soon as the number of
// keys at or above the current prefix fits the output buffers (normally after 2-3 passes) everything above the
// prefix is collected and ordered by a rank-counting sort: the first k entries are the selection, the entries after
// them are the near ties of the k-th score (guard band, SURVEY 7 hard part 1).  Only if a tie class of identical
// key1 values is larger than the buffers does the select continue over key2 and idx restricted to that class
// (exact 3-level threshold) -- then the band is reported as open.
#pragma once
#include "device_math.cuh"

namespace sdpcs {

constexpr int SEL_BINS = 2048;
constexpr u64 IDX_TOP = (1ull << 44) - 1;   // level-2 key = IDX_TOP - idx (smaller idx first); idx < 2^44

struct SelState {
    u64 T[3];        // per-level thresholds of the k-th element (exact path)
    u64 prefix;      // partial prefix of the level being resolved
    i64 need;        // elements still to take from the current tie class
    i64 n_valid;     // # key1 != 0
    i64 k_eff;       // min(k, n_valid)
    i64 n_violated;  // counters of the first pass (strict thresholds)
    i64 n_strong;
    u64 max_pos_nonviol; // enc_key of the largest obj among candidates with obj > thres_min_opt that are NOT violated (0: none)
    i64 n_unc_lam;   // candidates with |lam - thres_neg_eigval| <= guard_lam (classification inside the guard band)
    i64 n_unc_obj;   // candidates with |obj - thres_min_opt| <= guard_obj
    u64 collect_lo;  // early exit: every key1 >= collect_lo is collected
    i64 n_collect;   // ... and this many there are
    i64 k_out;       // winners = min(k_eff, collected)
    i64 band_count;  // entries right after the winners whose key1 score lies within the guard band of the k-th score
    u64 band_key;    // enc_key(score_k - delta)
    int band_open;   // 1: near ties may exist that were not collected (band reaches below collect_lo / exact path)
    int level;
    int done;        // 0 running, 1 exact thresholds T[0..2], 2 early exit (collect_lo)
    unsigned out_count;
    unsigned pad;
    unsigned hist[SEL_BINS];
};

// Where the keys come from.  MODE 0: stored key1 array (triangles); MODE 1..4: computed from the resident scores
//   1  feasibility: violated only, key -lam                                  (cut_select_qp.py:639-654)
//   2  optimality:  all, key obj                                             (cut_select_qp.py:599-601)
//   3  strong set:  obj > thres_min_opt and violated, key obj                (cut_select_qp.py:607-612)
//   4  combined final: key1 = re-scored measure given the pivot, key2 = obj  (cut_select_qp.py:603-625)
// With a guard (guard_lam / guard_obj > 0) modes 1 and 3 classify with the RELAXED thresholds lam < thr + guard_lam,
// obj > thr - guard_obj, so that a candidate whose classification the reference's LAPACK / libm rounding could flip is
// never excluded on the device; the counters stay strict and n_unc_* say how many such candidates exist.
struct KeySrc {
    const u64* key1;
    const i64* idx;              // MODE 0: explicit idx or nullptr (idx = base + position)
    const double* lam;
    const double* obj;
    i64 N, base;
    double thr_eig, thr_opt, big_m, guard_lam, guard_obj;
    double pivot_obj; i64 pivot_idx; int all_walked;
    SelState* st;
};

struct SelCounters {
    i64 nvalid = 0, nv = 0, ns = 0, nul = 0, nuo = 0;
    u64 mx = 0;
};

template <int MODE, bool COUNT>
__device__ __forceinline__ void make_key(const KeySrc& s, i64 i, u64 a_bits, u64 b_bits, u64& k1, u64& k2, SelCounters& c)
{
    k2 = 0;
    if (MODE == 0) {
        k1 = a_bits;
    } else {
        const double lam = __longlong_as_double((long long)a_bits), obj = __longlong_as_double((long long)b_bits);
        const bool has_lam = (MODE != 2), has_obj = (MODE != 1);
        const bool viol = has_lam && lam < s.thr_eig, pos = has_obj && obj > s.thr_opt;
        const bool viol_r = has_lam && lam < s.thr_eig + s.guard_lam, pos_r = has_obj && obj > s.thr_opt - s.guard_obj;
        if (COUNT) {
            c.nv += viol; c.ns += (viol && pos);
            c.nul += has_lam && s.guard_lam > 0.0 && fabs(lam - s.thr_eig) <= s.guard_lam;
            c.nuo += has_obj && s.guard_obj > 0.0 && fabs(obj - s.thr_opt) <= s.guard_obj;
            if (pos && !viol) { const u64 e = enc_key(obj); c.mx = e > c.mx ? e : c.mx; }
        }
        if (MODE == 1) k1 = viol_r ? enc_key(-lam) : 0;
        else if (MODE == 2) k1 = enc_key(obj);
        else if (MODE == 3) k1 = (viol_r && pos_r) ? enc_key(obj) : 0;
        else {
            const i64 idx = s.base + i;
            const bool walked = s.all_walked || obj > s.pivot_obj || (obj == s.pivot_obj && idx <= s.pivot_idx);
            double f = obj;
            if (walked) {
                if (pos) f = viol ? obj + s.big_m : obj - s.big_m;   // cut_select_qp.py:611, 615
                else if (viol) f = -lam;                             // cut_select_qp.py:620
            }
            k1 = enc_key(f); k2 = enc_key(obj);
        }
    }
    if (COUNT) c.nvalid += (k1 != 0);
}

// Grid-stride walk over all candidates with 16-byte loads, 4 (one array) or 2 x 2 (two arrays) loads in flight per
// thread: an HBM-bound pass needs that much memory-level parallelism.  f(i, key1, key2).
template <int MODE, bool COUNT, typename F>
__device__ __forceinline__ void sel_foreach(const KeySrc& s, SelCounters& c, F&& f)
{
    constexpr bool LA = (MODE == 1 || MODE >= 3), LB = (MODE == 2 || MODE >= 3);
    constexpr int U = (MODE >= 3) ? 2 : 4;
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x, nthr = (i64)gridDim.x * blockDim.x;
    const i64 N2 = s.N & ~(i64)1;
    const ulonglong2 zero = make_ulonglong2(0, 0);
    for (i64 i = tid * 2; i < N2; i += nthr * 2 * U) {
        ulonglong2 A[U], B[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const i64 j = i + (i64)u * nthr * 2;
            const bool in = j < N2;
            A[u] = zero; B[u] = zero;
            if (MODE == 0 && in) A[u] = *reinterpret_cast<const ulonglong2*>(s.key1 + j);
            if (LA && in) A[u] = *reinterpret_cast<const ulonglong2*>(s.lam + j);
            if (LB && in) B[u] = *reinterpret_cast<const ulonglong2*>(s.obj + j);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const i64 j = i + (i64)u * nthr * 2;
            if (j >= N2) continue;
            u64 k1, k2;
            make_key<MODE, COUNT>(s, j, A[u].x, B[u].x, k1, k2, c);
            f(j, k1, k2);
            make_key<MODE, COUNT>(s, j + 1, A[u].y, B[u].y, k1, k2, c);
            f(j + 1, k1, k2);
        }
    }
    if ((s.N & 1) && tid == 0) {
        const i64 j = s.N - 1;
        u64 a = 0, b = 0, k1, k2;
        if (MODE == 0) a = s.key1[j];
        if (LA) a = (u64)__double_as_longlong(s.lam[j]);
        if (LB) b = (u64)__double_as_longlong(s.obj[j]);
        make_key<MODE, COUNT>(s, j, a, b, k1, k2, c);
        f(j, k1, k2);
    }
}

__global__ void k_sel_reset(SelState* st)
{
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) st->hist[i] = 0;
    if (threadIdx.x == 0) {
        st->T[0] = st->T[1] = st->T[2] = 0; st->prefix = 0; st->need = 0; st->n_valid = 0; st->k_eff = 0;
        st->n_violated = 0; st->n_strong = 0; st->max_pos_nonviol = 0; st->n_unc_lam = 0; st->n_unc_obj = 0;
        st->collect_lo = 0; st->n_collect = 0; st->k_out = 0; st->band_count = 0; st->band_key = 0; st->band_open = 0;
        st->level = 0; st->done = 0; st->out_count = 0;
    }
}

__device__ __forceinline__ u64 idx_key(const KeySrc& s, i64 i) { return IDX_TOP - (u64)(s.idx ? s.idx[i] : s.base + i); }

// one radix pass: histogram of digit (key >> shift) & (2^width - 1) over elements still matching
template <int MODE, bool COUNT>
__global__ void __launch_bounds__(512) k_sel_hist(KeySrc s, int level, int shift, int width)
{
    __shared__ unsigned sh[SEL_BINS];
    SelState* st = s.st;
    if (st->done || st->level != level) return;
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const u64 T0 = st->T[0], T1 = st->T[1], prefix = st->prefix;
    const int hs = shift + width;
    const unsigned mask = (1u << width) - 1;
    SelCounters c;
    // run-length aggregation per thread: the leading digits (sign + exponent of an FP64 score) put almost every key
    // into the same two or three bins, and one shared-memory atomic per key would serialise on them
    unsigned run_bin = 0, run_cnt = 0;
    auto take = [&](i64 i, u64 k1, u64 k2) {
        if (k1 == 0) return;
        if (level >= 1 && k1 != T0) return;
        if (level == 2 && k2 != T1) return;
        const u64 key = (level == 0) ? k1 : (level == 1) ? k2 : idx_key(s, i);
        if (hs < 64 && (key >> hs) != (prefix >> hs)) return;
        const unsigned bin = (unsigned)(key >> shift) & mask;
        if (bin == run_bin) ++run_cnt;
        else {
            if (run_cnt) atomicAdd(&sh[run_bin], run_cnt);
            run_bin = bin; run_cnt = 1;
        }
    };
    sel_foreach<MODE, COUNT>(s, c, take);
    if (run_cnt) atomicAdd(&sh[run_bin], run_cnt);
    __syncthreads();
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&st->hist[i], sh[i]);
    if (COUNT) {
        for (int o = 16; o; o >>= 1) {
            c.nvalid += __shfl_xor_sync(0xffffffffu, c.nvalid, o); c.nv += __shfl_xor_sync(0xffffffffu, c.nv, o);
            c.ns += __shfl_xor_sync(0xffffffffu, c.ns, o); c.nul += __shfl_xor_sync(0xffffffffu, c.nul, o);
            c.nuo += __shfl_xor_sync(0xffffffffu, c.nuo, o);
            const u64 m2 = __shfl_xor_sync(0xffffffffu, c.mx, o);
            c.mx = m2 > c.mx ? m2 : c.mx;
        }
        if ((threadIdx.x & 31) == 0) {
            if (c.nvalid) atomicAdd((u64*)&st->n_valid, (u64)c.nvalid);
            if (c.nv) atomicAdd((u64*)&st->n_violated, (u64)c.nv);
            if (c.ns) atomicAdd((u64*)&st->n_strong, (u64)c.ns);
            if (c.nul) atomicAdd((u64*)&st->n_unc_lam, (u64)c.nul);
            if (c.nuo) atomicAdd((u64*)&st->n_unc_obj, (u64)c.nuo);
            if (c.mx) atomicMax((unsigned long long*)&st->max_pos_nonviol, (unsigned long long)c.mx);
        }
    }
}

// digit decision after a pass (single block).  Level 0: as soon as the keys at or above the chosen prefix number at most
// cap_exit, the select stops (done = 2) and everything >= prefix is collected.
__global__ void __launch_bounds__(1024) k_sel_scan(SelState* st, int level, int shift, int width, int first_pass,
                                                   int last_pass, int next_level, i64 k, i64 cap_exit)
{
    __shared__ i64 suf[SEL_BINS + 1];
    if (st->done || st->level != level) return;
    const int nb = 1 << width;
    // suffix sums: suf[b] = sum_{b' >= b} hist[b']
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) suf[i] = (i < nb) ? (i64)st->hist[i] : 0;
    if (threadIdx.x == 0) suf[SEL_BINS] = 0;
    __syncthreads();
    for (int off = 1; off < SEL_BINS; off <<= 1) {
        i64 v0 = 0, v1 = 0;
        int i0 = threadIdx.x, i1 = threadIdx.x + blockDim.x;
        if (i0 + off < SEL_BINS) v0 = suf[i0 + off];
        if (i1 < SEL_BINS && i1 + off < SEL_BINS) v1 = suf[i1 + off];
        __syncthreads();
        suf[i0] += v0;
        if (i1 < SEL_BINS) suf[i1] += v1;
        __syncthreads();
    }
    __shared__ i64 need_s, keff_s;
    if (threadIdx.x == 0) {
        i64 need = st->need, ke = st->k_eff;
        if (first_pass && level == 0) {
            ke = k < st->n_valid ? k : st->n_valid;
            st->k_eff = ke;
            need = ke;
        }
        need_s = need; keff_s = ke;
    }
    __syncthreads();
    const i64 need = need_s, keff = keff_s;
    if (need <= 0) {   // nothing to select
        if (threadIdx.x == 0) { st->done = 1; st->T[0] = ~0ull; st->T[1] = ~0ull; st->T[2] = ~0ull; }
        for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) st->hist[i] = 0;
        return;
    }
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        i64 above = suf[b + 1], incl = suf[b];
        if (above < need && need <= incl) {
            u64 prefix = st->prefix | ((u64)b << shift);
            i64 need2 = need - above;
            i64 cnt = incl - above;
            st->prefix = prefix;
            st->need = need2;
            const i64 total = (keff - need) + incl;          // # key1 >= prefix (its low bits are zero)
            if (level == 0 && total <= cap_exit) {
                st->collect_lo = prefix; st->n_collect = total; st->done = 2;
            } else if (last_pass) {
                st->T[level] = prefix;
                if (need2 == cnt || next_level < 0) st->done = 1;   // whole tie class taken
                else { st->level = next_level; st->prefix = 0; }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) st->hist[i] = 0;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_sel_collect(KeySrc s, i64 cap, u64* out_k1, u64* out_k2, i64* out_idx)
{
    SelState* st = s.st;
    if (!st->done || st->k_eff <= 0) return;
    const bool early = st->done == 2;
    const u64 lo = st->collect_lo, T0 = st->T[0], T1 = st->T[1], T2 = st->T[2];
    SelCounters c;
    auto take = [&](i64 i, u64 k1, u64 k2) {
        if (k1 == 0) return;
        const i64 idx = (MODE == 0 && s.idx) ? s.idx[i] : s.base + i;
        if (early) {
            if (k1 < lo) return;
        } else {
            if (k1 < T0) return;
            if (k1 == T0) {
                if (k2 < T1) return;
                if (k2 == T1 && (IDX_TOP - (u64)idx) < T2) return;
            }
        }
        unsigned p = atomicAdd(&st->out_count, 1u);
        if ((i64)p < cap) { out_k1[p] = k1; out_k2[p] = k2; out_idx[p] = idx; }
    };
    sel_foreach<MODE, false>(s, c, take);
}

// rank-counting sort of the m <= cap collected entries by (k1 desc, k2 desc, idx asc).
// grid = (ceil(m_max / 256), S): block (bx, by) counts, for its 256 entries, the entries of segment by of the list that come
// before them; the S partial ranks of an entry meet in rank[] (atomicAdd), and the last of the S blocks of a column (ticket
// in blk_cnt[bx]) scatters the entries to their places.  rank[] and blk_cnt[] must be zero on entry and are zero again on
// exit.  S = 1 (rank == nullptr allowed): one block scans the whole list, no atomics.  The O(m^2) comparisons are spread
// over the whole GPU instead of m / 256 SMs: m = 10,000 took 0.32 ms with S = 1 (most of the device time of a small
// selection, and a fixed cost of every sharded step).
__global__ void __launch_bounds__(256) k_rank_sort(const SelState* st, i64 m_fixed, i64 cap, const u64* k1, const u64* k2,
                                                   const i64* idx, u64* s_k1, u64* s_k2, i64* s_idx, unsigned* rank, unsigned* blk_cnt)
{
    __shared__ u64 t1[256], t2[256];
    __shared__ i64 ti[256];
    __shared__ int s_last;
    if (st && !st->done) return;
    i64 m = st ? (i64)st->out_count : m_fixed;
    if (m > cap) m = cap;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if ((i64)blockIdx.x * blockDim.x >= m) return;
    const int S = (int)gridDim.y;
    const i64 tiles = (m + 255) / 256, per = (tiles + S - 1) / S;
    const i64 jb = (i64)blockIdx.y * per * 256, je = (jb + per * 256 < m) ? jb + per * 256 : m;
    u64 a1 = 0, a2 = 0; i64 ai = 0;
    if (i < m) { a1 = k1[i]; a2 = k2[i]; ai = idx[i]; }
    unsigned pos = 0;
    for (i64 j0 = jb; j0 < je; j0 += 256) {
        i64 j = j0 + threadIdx.x;
        __syncthreads();
        if (j < je) { t1[threadIdx.x] = k1[j]; t2[threadIdx.x] = k2[j]; ti[threadIdx.x] = idx[j]; }
        __syncthreads();
        int lim = (int)((je - j0) < 256 ? (je - j0) : 256);
#pragma unroll 4
        for (int q = 0; q < lim; ++q) {
            u64 b1 = t1[q], b2 = t2[q]; i64 bi = ti[q];
            bool before = (b1 > a1) || (b1 == a1 && (b2 > a2 || (b2 == a2 && bi < ai)));
            pos += before;
        }
    }
    if (S > 1) {
        if (i < m && pos) atomicAdd(rank + i, pos);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(blk_cnt + blockIdx.x, 1u) == (unsigned)(S - 1));
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        if (threadIdx.x == 0) blk_cnt[blockIdx.x] = 0;
        if (i < m) pos = atomicExch(rank + i, 0u);
    }
    if (i < m) { s_k1[pos] = a1; s_k2[pos] = a2; s_idx[pos] = ai; }
}

// After the sort: number of winners, and the guard band = entries after the k-th whose key1 score is >= score_k - delta.
// delta < 0: no guard.  band_open = 1 when near ties may exist that were not collected.
__global__ void k_sel_finish(SelState* st, const u64* s_k1, i64 cap, double delta)
{
    if (threadIdx.x || blockIdx.x) return;
    if (!st->done) return;
    i64 m = (i64)st->out_count;
    if (m > cap) m = cap;
    const i64 k_out = st->k_eff < m ? st->k_eff : m;
    st->k_out = k_out;
    st->band_count = 0; st->band_open = 0; st->band_key = 0;
    if (k_out <= 0 || delta < 0.0) return;
    const double sk = dec_key(s_k1[k_out - 1]);
    const u64 bk = enc_key(sk - delta);
    st->band_key = bk;
    i64 lo = k_out, hi = m;                     // first position in [k_out, m) with key < bk (sorted descending)
    while (lo < hi) {
        const i64 mid = (lo + hi) >> 1;
        if (s_k1[mid] >= bk) lo = mid + 1; else hi = mid;
    }
    st->band_count = lo - k_out;
    const bool more_valid = st->n_valid > m;
    if (st->done == 2) st->band_open = (more_valid && bk < st->collect_lo) ? 1 : 0;
    else st->band_open = more_valid ? 1 : 0;    // exact path: only the k winners were collected
}

// re-open an early-exit selection with a lower collection bound (the band reached below collect_lo)
__global__ void k_sel_lower(SelState* st)
{
    if (threadIdx.x || blockIdx.x) return;
    if (st->done == 2 && st->band_open) { st->collect_lo = st->band_key; st->out_count = 0; }
}

// final gather of the collected entries' scores
__global__ void k_sel_gather(const SelState* st, i64 cap, const u64* s_k1, const i64* s_idx, i64 base, const double* lam,
                             const double* obj, double* o_score, double* o_lam, double* o_obj)
{
    if (!st->done) return;
    i64 m = (i64)st->out_count;
    if (m > cap) m = cap;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) {
        i64 loc = s_idx[i] - base;
        o_score[i] = dec_key(s_k1[i]);
        o_lam[i] = lam ? lam[loc] : 0.0;
        o_obj[i] = obj ? obj[loc] : 0.0;
    }
}

// keys for merging lists gathered from several shards
__global__ void k_merge_keys(i64 m, const double* score, const double* obj2, u64* k1, u64* k2)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) { k1[i] = enc_key(score[i]); k2[i] = obj2 ? enc_key(obj2[i]) : 0; }
}

}  // namespace sdpcs
