// Semidefinite vertex cover P^E_rho on the sparsity pattern (cut_select_qp.py:377-522, ch_ext = 0), built on the device.
//
// The reference walks nested loops over the adjacency of Q and emits every rho-clique plus every smaller clique
// (size >= 2) that no clique one size larger contains, in the order of the loops = lexicographic order of the index
// tuples.  Here one thread owns one edge (i1, i2), i1 < i2, and walks the sub-tree of the cliques
// that start with it, on 256-bit adjacency rows (n <= 250):
//     grow(clique, common):  |clique| = rho            -> emit clique
//                            common = 0                -> emit clique (maximal, smaller than rho)
//                            else for v in common, v > last(clique), ascending:  grow(clique + v, common & adj[v])
// (a clique whose common neighbours are all smaller than its last vertex is contained in a larger clique that an
// earlier sub-tree emits or extends, so nothing is emitted for it: cut_select_qp.py:410-424, 462-483, 490-522).
// Two passes: count rows per edge and per size class, exclusive scan on the host (<= 31,125 edges), then emit
// straight into the per-size arrays the score kernels read (uint8 index rows + position in the overall order).
#pragma once
#include "device_math.cuh"

namespace sdpcs {

struct Mask256 {
    u64 w[4];
};

__device__ __forceinline__ Mask256 mask_and(const Mask256& a, const Mask256& b)
{
    Mask256 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.w[i] = a.w[i] & b.w[i];
    return r;
}
__device__ __forceinline__ bool mask_empty(const Mask256& a) { return (a.w[0] | a.w[1] | a.w[2] | a.w[3]) == 0; }

struct CoverArgs {
    int n, rho;
    const Mask256* adj;        // n rows; bit j of row i: Q_adj[i][j] != 0 (or [j][i]), i != j
    const int* edges;          // E pairs (i1, i2), lexicographic
    i64 E;
    // pass 1
    int* counts;               // [E][4]: rows of size 2, 3, 4, 5 emitted by the edge's sub-tree
    // pass 2
    const i64* pos_off;        // [E]: position of the edge's first row in the overall order
    const i64* cls_off;        // [E][4]: slot of the edge's first row of each size inside that size's array
    uint8_t* idx[6];           // [size]: Nd[size] x size vertex rows
    i64* pos[6];               // [size]: Nd[size] positions in the overall order
};

template <bool EMIT>
struct CoverSink {
    int cnt[4];
    i64 seq;                   // rows emitted so far by this thread (EMIT: next overall position)
    i64 slot[4];
    const CoverArgs* a;
    __device__ __forceinline__ void emit(const int* clique, int size)
    {
        if (EMIT) {
            const i64 s = slot[size - 2]++;
            uint8_t* dst = a->idx[size] + s * size;
            for (int t = 0; t < size; ++t) dst[t] = (uint8_t)clique[t];
            a->pos[size][s] = seq;
        } else {
            cnt[size - 2]++;
        }
        ++seq;
    }
};

template <int LEVEL, bool EMIT>
__device__ void cover_grow(const CoverArgs& a, int* clique, const Mask256& common, CoverSink<EMIT>& sink)
{
    // clique[0 .. LEVEL) is a clique, common = vertices adjacent to all of it
    if (LEVEL == a.rho) { sink.emit(clique, LEVEL); return; }
    if (mask_empty(common)) { sink.emit(clique, LEVEL); return; }
    if constexpr (LEVEL < 5) {
        const int first = clique[LEVEL - 1] + 1;
        for (int wi = first >> 6; wi < 4; ++wi) {
            u64 bits = common.w[wi];
            if (wi == (first >> 6)) bits &= ~0ull << (first & 63);
            while (bits) {
                const int v = wi * 64 + __ffsll((long long)bits) - 1;
                bits &= bits - 1;
                clique[LEVEL] = v;
                const Mask256 next = mask_and(common, a.adj[v]);
                cover_grow<LEVEL + 1, EMIT>(a, clique, next, sink);
            }
        }
    }
}

template <bool EMIT>
__global__ void __launch_bounds__(128) k_cover_pattern(CoverArgs a)
{
    const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.E) return;
    int clique[5];
    clique[0] = a.edges[2 * e];
    clique[1] = a.edges[2 * e + 1];
    CoverSink<EMIT> sink;
    sink.a = &a;
#pragma unroll
    for (int d = 0; d < 4; ++d) { sink.cnt[d] = 0; sink.slot[d] = EMIT ? a.cls_off[4 * e + d] : 0; }
    sink.seq = EMIT ? a.pos_off[e] : 0;
    const Mask256 common = mask_and(a.adj[clique[0]], a.adj[clique[1]]);
    cover_grow<2, EMIT>(a, clique, common, sink);
    if (!EMIT) {
#pragma unroll
        for (int d = 0; d < 4; ++d) a.counts[4 * e + d] = sink.cnt[d];
    }
}

// LIST cover -> padded int16 rows in the overall order (sdpcs_get_cover_rows)
__global__ void __launch_bounds__(256) k_cover_rows(const uint8_t* idx, const i64* pos, i64 Nd, int size, int rho, int16_t* out)
{
    for (i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x; s < Nd; s += (i64)gridDim.x * blockDim.x) {
        int16_t* dst = out + pos[s] * rho;
        for (int t = 0; t < rho; ++t) dst[t] = t < size ? (int16_t)idx[s * size + t] : (int16_t)-1;
    }
}

// positions are ascending inside a size class: [lo, hi) of the rows whose position lies in [b, e)
__global__ void k_pos_bounds(const i64* pos, i64 n, i64 b, i64 e, i64* out2)
{
    if (threadIdx.x || blockIdx.x) return;
    i64 lo = 0, hi = n;
    while (lo < hi) { const i64 m = (lo + hi) >> 1; if (pos[m] < b) lo = m + 1; else hi = m; }
    out2[0] = lo;
    hi = n;
    while (lo < hi) { const i64 m = (lo + hi) >> 1; if (pos[m] < e) lo = m + 1; else hi = m; }
    out2[1] = lo;
}

__global__ void __launch_bounds__(256) k_pos_shift(i64* pos, i64 n, i64 delta)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) pos[i] -= delta;
}

// ---------------------------------------------------------------------------------------------------
// Cover algebra on the device (cut_select_qcqp.py:319-333): the reference intersects / subtracts two vertex covers with
// `el in agg_list` list scans, O(N^2).  Rows of one size class are in lexicographic order in both covers, so membership
// is a binary search on the rows read as big-endian integers; survivors are compacted with exclusive scans.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 cover_row_key(const uint8_t* r, int size)
{
    u64 k = 0;
    for (int t = 0; t < size; ++t) k = (k << 8) | r[t];
    return k;
}

// keep_cls[s] = keep_all[pos[s]] = (row s of `mine` occurs in `other`) == keep_members
__global__ void __launch_bounds__(256) k_cover_member(const uint8_t* idx, const i64* pos, i64 Nd, int size, const uint8_t* oidx, i64 oNd,
                                                      int keep_members, int* keep_cls, int* keep_all)
{
    for (i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x; s < Nd; s += (i64)gridDim.x * blockDim.x) {
        const u64 key = cover_row_key(idx + s * size, size);
        i64 lo = 0, hi = oNd;
        while (lo < hi) {
            const i64 m = (lo + hi) >> 1;
            if (cover_row_key(oidx + m * size, size) < key) lo = m + 1; else hi = m;
        }
        const int member = (lo < oNd && cover_row_key(oidx + lo * size, size) == key) ? 1 : 0;
        const int keep = (member == (keep_members ? 1 : 0)) ? 1 : 0;
        keep_cls[s] = keep;
        keep_all[pos[s]] = keep;
    }
}

// exclusive scan of n ints into i64, three kernels: per-block scan (1024 elements per block) + block totals,
// scan of the totals by one block, add-back
constexpr int SCAN_TILE = 1024;
__global__ void __launch_bounds__(256) k_scan_block(const int* in, i64 n, i64* out, i64* sums)
{
    __shared__ i64 sh[256];
    const i64 base = (i64)blockIdx.x * SCAN_TILE + 4 * threadIdx.x;
    int v[4];
    i64 local = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j] = (base + j < n) ? in[base + j] : 0; local += v[j]; }
    sh[threadIdx.x] = local;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const i64 add = (threadIdx.x >= off) ? sh[threadIdx.x - off] : 0;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    i64 run = sh[threadIdx.x] - local;                    // exclusive prefix of this thread inside the block
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
    if (threadIdx.x == 255) sums[blockIdx.x] = sh[255];
}
__global__ void __launch_bounds__(1024) k_scan_sums(i64* sums, i64 nb, i64* total)
{
    __shared__ i64 sh[1024];
    __shared__ i64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (i64 c0 = 0; c0 < nb; c0 += 1024) {
        const i64 i = c0 + threadIdx.x;
        const i64 v = i < nb ? sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const i64 add = (threadIdx.x >= off) ? sh[threadIdx.x - off] : 0;
            __syncthreads();
            sh[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < nb) sums[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256) k_scan_add(i64* out, i64 n, const i64* sums)
{
    const i64 base = (i64)blockIdx.x * SCAN_TILE + 4 * threadIdx.x;
    const i64 add = sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (base + j < n) out[base + j] += add;
}

__global__ void __launch_bounds__(256) k_cover_compact(const uint8_t* idx, const i64* pos, i64 Nd, int size, const int* keep_cls,
                                                       const i64* cls_scan, const i64* all_scan, uint8_t* nidx, i64* npos)
{
    for (i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x; s < Nd; s += (i64)gridDim.x * blockDim.x) {
        if (!keep_cls[s]) continue;
        const i64 d = cls_scan[s];
        for (int t = 0; t < size; ++t) nidx[d * size + t] = idx[s * size + t];
        npos[d] = all_scan[pos[s]];
    }
}

}  // namespace sdpcs
