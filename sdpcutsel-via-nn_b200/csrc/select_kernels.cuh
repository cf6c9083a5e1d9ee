// K5: deterministic top-k over resident scores.
// Replaces rank_list.sort(key=itemgetter(1), reverse=True) + prefix slicing (cut_select_qp.py:601, 625, 653-654,
// consumer bound :718) and the triangle sort (:841-844).  Python's stable descending sort == order by
// (score desc, original position asc); for the combined rule the "original position" of the second sort is the
// order of the first one, i.e. (obj desc, agg_idx asc).  Both are expressed as a 3-level key
//     (key1 desc, key2 desc, idx asc)        key = order-preserving u64 image of the FP64 score
// and resolved exactly by an MSD radix select (11-bit digits) over key1, then -- only if the k-th place falls
// inside a tie class -- over key2 and over idx restricted to that class.  Then the <= k winners are collected
// and ordered by a rank-counting sort. Everything runs on the device; no host round trip inside a selection.
#pragma once
#include "device_math.cuh"

namespace sdpcs {

constexpr int SEL_BINS = 2048;
constexpr u64 IDX_TOP = (1ull << 44) - 1;   // level-2 key = IDX_TOP - idx (smaller idx first); idx < 2^44

struct SelState {
    u64 T[3];        // per-level thresholds of the k-th element
    u64 prefix;      // partial prefix of the level being resolved
    i64 need;        // elements still to take from the current tie class
    i64 n_valid;     // # key1 != 0
    i64 k_eff;       // min(k, n_valid)
    i64 n_violated;  // counters filled by k_make_keys
    i64 n_strong;
    u64 max_pos_nonviol; // enc_key of the largest obj among candidates with obj > thres_min_opt that are NOT violated (0: none)
    int level;
    int done;
    unsigned out_count;
    unsigned pad;
    unsigned hist[SEL_BINS];
};

struct KeyArgs {
    const double* lam;
    const double* obj;
    i64 N;
    i64 base;            // agg_idx of local element 0
    int mode;            // 1 feas, 2 opt, 3 strong, 4 combined-final
    double thr_eig, thr_opt, big_m;
    double pivot_obj; i64 pivot_idx; int all_walked;
    u64* key1; u64* key2;
    SelState* st;
};

__global__ void __launch_bounds__(256) k_make_keys(KeyArgs a)
{
    i64 nv = 0, ns = 0;
    u64 mx = 0;
    auto make = [&](i64 i, double lam, double obj, u64& k1, u64& k2) {
        bool viol = a.lam && lam < a.thr_eig;
        bool pos = a.obj && obj > a.thr_opt;
        nv += viol; ns += (viol && pos);
        if (pos && !viol) { const u64 e = enc_key(obj); mx = e > mx ? e : mx; }
        k1 = 0; k2 = 0;
        switch (a.mode) {
        case 1: k1 = viol ? enc_key(-lam) : 0; break;
        case 2: k1 = enc_key(obj); break;
        case 3: k1 = (viol && pos) ? enc_key(obj) : 0; break;
        default: {
            i64 idx = a.base + i;
            bool walked = a.all_walked || obj > a.pivot_obj || (obj == a.pivot_obj && idx <= a.pivot_idx);
            double f = obj;
            if (walked) {
                if (pos) f = viol ? obj + a.big_m : obj - a.big_m;   // cut_select_qp.py:611, 615
                else if (viol) f = -lam;                             // cut_select_qp.py:620
            }
            k1 = enc_key(f); k2 = enc_key(obj);
        } break;
        }
    };
    // 16-byte loads / stores, two pairs in flight per thread (HBM-bound pass: 16 B read + 8..16 B written per candidate)
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x, nthr = (i64)gridDim.x * blockDim.x;
    const i64 N4 = a.N & ~(i64)3;
    for (i64 i = tid * 2; i < N4; i += nthr * 4) {
        double2 L[2], O[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const i64 j = i + u * nthr * 2;
            const bool in = j < N4;
            L[u] = (in && a.lam) ? *reinterpret_cast<const double2*>(a.lam + j) : make_double2(0.0, 0.0);
            O[u] = (in && a.obj) ? *reinterpret_cast<const double2*>(a.obj + j) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const i64 j = i + u * nthr * 2;
            if (j >= N4) continue;
            ulonglong2 K1, K2;
            make(j, L[u].x, O[u].x, K1.x, K2.x);
            make(j + 1, L[u].y, O[u].y, K1.y, K2.y);
            *reinterpret_cast<ulonglong2*>(a.key1 + j) = K1;
            if (a.key2) *reinterpret_cast<ulonglong2*>(a.key2 + j) = K2;
        }
    }
    for (i64 i = N4 + tid; i < a.N; i += nthr) {
        u64 k1, k2;
        make(i, a.lam ? a.lam[i] : 0.0, a.obj ? a.obj[i] : 0.0, k1, k2);
        a.key1[i] = k1;
        if (a.key2) a.key2[i] = k2;
    }
    // block reduce the two counters
    __shared__ i64 sh[2][8];
    for (int o = 16; o; o >>= 1) {
        nv += __shfl_xor_sync(0xffffffffu, nv, o); ns += __shfl_xor_sync(0xffffffffu, ns, o);
        const u64 m2 = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = m2 > mx ? m2 : mx;
    }
    if ((threadIdx.x & 31) == 0 && mx) atomicMax((unsigned long long*)&a.st->max_pos_nonviol, (unsigned long long)mx);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = nv; sh[1][threadIdx.x >> 5] = ns; }
    __syncthreads();
    if (threadIdx.x == 0) {
        i64 a0 = 0, a1 = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a0 += sh[0][w]; a1 += sh[1][w]; }
        if (a0) atomicAdd((u64*)&a.st->n_violated, (u64)a0);
        if (a1) atomicAdd((u64*)&a.st->n_strong, (u64)a1);
    }
}

__global__ void k_sel_reset(SelState* st)
{
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) st->hist[i] = 0;
    if (threadIdx.x == 0) {
        st->T[0] = st->T[1] = st->T[2] = 0; st->prefix = 0; st->need = 0; st->n_valid = 0; st->k_eff = 0;
        st->n_violated = 0; st->n_strong = 0; st->max_pos_nonviol = 0; st->level = 0; st->done = 0; st->out_count = 0;
    }
}

struct SelArgs {
    const u64* key1; const u64* key2; const i64* idx;  // idx == nullptr: idx = base + position
    i64 N; i64 base; SelState* st;
};

__device__ __forceinline__ u64 level_key(const SelArgs& a, i64 i, int level)
{
    if (level == 0) return a.key1[i];
    if (level == 1) return a.key2 ? a.key2[i] : 0;
    return IDX_TOP - (u64)(a.idx ? a.idx[i] : a.base + i);
}

// one radix pass: histogram of digit (key >> shift) & (2^width - 1) over elements still matching
__global__ void __launch_bounds__(512) k_sel_hist(SelArgs a, int level, int shift, int width, int count_valid)
{
    __shared__ unsigned sh[SEL_BINS];
    SelState* st = a.st;
    if (st->done || st->level != level) return;
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const u64 T0 = st->T[0], T1 = st->T[1], prefix = st->prefix;
    const int hs = shift + width;
    const unsigned mask = (1u << width) - 1;
    i64 nvalid = 0;
    // run-length aggregation per thread: the leading digits (sign + exponent of an FP64 score) put almost every key
    // into the same two or three bins, and one shared-memory atomic per key would serialise on them
    unsigned run_bin = 0, run_cnt = 0;
    auto take = [&](i64 i, u64 k1) {
        if (count_valid) nvalid += (k1 != 0);
        if (level >= 1 && k1 != T0) return;
        if (level == 2 && a.key2 && a.key2[i] != T1) return;
        u64 key = (level == 0) ? k1 : level_key(a, i, level);
        if (hs < 64 && (key >> hs) != (prefix >> hs)) return;
        const unsigned bin = (unsigned)(key >> shift) & mask;
        if (bin == run_bin) ++run_cnt;
        else {
            if (run_cnt) atomicAdd(&sh[run_bin], run_cnt);
            run_bin = bin; run_cnt = 1;
        }
    };
    // four 16-byte loads in flight per thread: one 8-byte load per thread and iteration leaves HBM latency-bound
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x, nthr = (i64)gridDim.x * blockDim.x;
    const i64 N8 = a.N & ~(i64)7;
    for (i64 i = tid * 2; i < N8; i += nthr * 8) {
        ulonglong2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const i64 j = i + u * nthr * 2;
            v[u] = (j < N8) ? *reinterpret_cast<const ulonglong2*>(a.key1 + j) : make_ulonglong2(0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const i64 j = i + u * nthr * 2;
            if (j < N8) { take(j, v[u].x); take(j + 1, v[u].y); }
        }
    }
    for (i64 i = N8 + tid; i < a.N; i += nthr) take(i, a.key1[i]);
    if (run_cnt) atomicAdd(&sh[run_bin], run_cnt);
    __syncthreads();
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&st->hist[i], sh[i]);
    if (count_valid) {
        for (int o = 16; o; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
        if ((threadIdx.x & 31) == 0 && nvalid) atomicAdd((u64*)&st->n_valid, (u64)nvalid);
    }
}

// digit decision after a pass (single block of SEL_BINS/2 threads)
__global__ void __launch_bounds__(1024) k_sel_scan(SelState* st, int level, int shift, int width, int first_pass,
                                                   int last_pass, int next_level, i64 k)
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
    __shared__ i64 need_s;
    if (threadIdx.x == 0) {
        i64 need = st->need;
        if (first_pass && level == 0) {
            i64 ke = k < st->n_valid ? k : st->n_valid;
            st->k_eff = ke;
            need = ke;
        }
        need_s = need;
    }
    __syncthreads();
    const i64 need = need_s;
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
            if (last_pass) {
                st->T[level] = prefix;
                if (need2 == cnt || next_level < 0) st->done = 1;   // whole tie class taken
                else { st->level = next_level; st->prefix = 0; }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) st->hist[i] = 0;
}

__global__ void __launch_bounds__(256) k_sel_collect(SelArgs a, i64 cap, u64* out_k1, u64* out_k2, i64* out_idx)
{
    SelState* st = a.st;
    if (st->k_eff <= 0) return;
    const u64 T0 = st->T[0], T1 = st->T[1], T2 = st->T[2];
    auto take = [&](i64 i, u64 k1) {
        if (k1 < T0 || k1 == 0) return;
        u64 k2 = a.key2 ? a.key2[i] : 0;
        i64 idx = a.idx ? a.idx[i] : a.base + i;
        if (k1 == T0) {
            if (k2 < T1) return;
            if (k2 == T1 && (IDX_TOP - (u64)idx) < T2) return;
        }
        unsigned p = atomicAdd(&st->out_count, 1u);
        if ((i64)p < cap) { out_k1[p] = k1; out_k2[p] = k2; out_idx[p] = idx; }
    };
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x, nthr = (i64)gridDim.x * blockDim.x;
    const i64 N8 = a.N & ~(i64)7;
    for (i64 i = tid * 2; i < N8; i += nthr * 8) {
        ulonglong2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const i64 j = i + u * nthr * 2;
            v[u] = (j < N8) ? *reinterpret_cast<const ulonglong2*>(a.key1 + j) : make_ulonglong2(0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const i64 j = i + u * nthr * 2;
            if (j < N8) { take(j, v[u].x); take(j + 1, v[u].y); }
        }
    }
    for (i64 i = N8 + tid; i < a.N; i += nthr) take(i, a.key1[i]);
}

// rank-counting sort of the m <= cap winners by (k1 desc, k2 desc, idx asc)
__global__ void __launch_bounds__(256) k_rank_sort(const SelState* st, i64 m_fixed, const u64* k1, const u64* k2,
                                                   const i64* idx, u64* s_k1, u64* s_k2, i64* s_idx)
{
    __shared__ u64 t1[256], t2[256];
    __shared__ i64 ti[256];
    const i64 m = st ? (i64)st->out_count : m_fixed;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if ((i64)blockIdx.x * blockDim.x >= m) return;
    u64 a1 = 0, a2 = 0; i64 ai = 0;
    if (i < m) { a1 = k1[i]; a2 = k2[i]; ai = idx[i]; }
    i64 pos = 0;
    for (i64 j0 = 0; j0 < m; j0 += 256) {
        i64 j = j0 + threadIdx.x;
        __syncthreads();
        if (j < m) { t1[threadIdx.x] = k1[j]; t2[threadIdx.x] = k2[j]; ti[threadIdx.x] = idx[j]; }
        __syncthreads();
        int lim = (int)((m - j0) < 256 ? (m - j0) : 256);
        for (int q = 0; q < lim; ++q) {
            u64 b1 = t1[q], b2 = t2[q]; i64 bi = ti[q];
            bool before = (b1 > a1) || (b1 == a1 && (b2 > a2 || (b2 == a2 && bi < ai)));
            pos += before;
        }
    }
    if (i < m) { s_k1[pos] = a1; s_k2[pos] = a2; s_idx[pos] = ai; }
}

// final gather of the winners' scores
__global__ void k_sel_gather(const SelState* st, const u64* s_k1, const i64* s_idx, i64 base, const double* lam,
                             const double* obj, double* o_score, double* o_lam, double* o_obj)
{
    const i64 m = (i64)st->out_count;
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
