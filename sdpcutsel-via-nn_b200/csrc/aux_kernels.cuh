// K6 triangle-inequality scoring (cut_select_qp.py:799-844), K7 eigenvector-cut assembly (:737-751),
// single-matrix eigendecomposition (:788-797) and the FP64 peak probes used as roofline denominators.
#pragma once
#include "device_math.cuh"

namespace sdpcs {

// ---------------------------------------------------------------------------------------------------
// K6: one thread per triple (lex rank over C(n,3)); 4 keys per triple at position 4*rank + type.
// key = 0 if the triple has < thres_dense edges or the violation is < thres_viol, else
// (density == 3) << 63 | bits(violation)   -- violation > 0, so its bit pattern is monotone and < 2^63;
// ordering by key desc == sort(key=itemgetter(2, 3), reverse=True) (cut_select_qp.py:841).
// ---------------------------------------------------------------------------------------------------
struct TriArgs {
    int n; i64 T;                 // T = C(n,3)
    const double* X; const double* x;
    const uint8_t* adj;           // n x n, nullptr = dense
    int thres_dense; double thres_viol;
    u64* key;                     // 4*T
    unsigned long long* counters; // [0] = #triples kept by the density filter, [1] = #violated
};

__global__ void __launch_bounds__(256) k_tri_keys(TriArgs a)
{
    const int lane = threadIdx.x & 31;
    const i64 warps_total = (i64)gridDim.x * (blockDim.x >> 5);
    const i64 gw = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const i64 G = (a.T + 31) >> 5;
    const i64 g0 = gw * G / warps_total, g1 = (gw + 1) * G / warps_total;
    if (g0 >= g1) return;
    int c[3] = {0, 1, 2};
    if (g0 * 32 + lane < a.T) lex_unrank<3>(a.n, (u64)(g0 * 32 + lane), c);
    unsigned nkept = 0, nviol = 0;
    for (i64 g = g0; g < g1; ++g) {
        const i64 r = g * 32 + lane;
        const bool valid = r < a.T;
        if (valid) {
            const int i1 = c[0], i2 = c[1], i3 = c[2];
            int dens = 3;
            if (a.adj) dens = (a.adj[i1 * a.n + i2] != 0) + (a.adj[i1 * a.n + i3] != 0) + (a.adj[i2 * a.n + i3] != 0);
            u64 k[4] = {0, 0, 0, 0};
            if (dens >= a.thres_dense) {
                ++nkept;
                const double X1 = __ldg(a.X + tri_index(a.n, i1, i2)), X2 = __ldg(a.X + tri_index(a.n, i1, i3)),
                             X4 = __ldg(a.X + tri_index(a.n, i2, i3));
                const double x1 = __ldg(a.x + i1), x2 = __ldg(a.x + i2), x3 = __ldg(a.x + i3);
                double v[4];
                // same operation order as cut_select_qp.py:835-838 (adds only, no contraction possible)
                v[0] = __dsub_rn(__dsub_rn(__dadd_rn(X1, X2), X4), x1);
                v[1] = __dsub_rn(__dadd_rn(__dsub_rn(X1, X2), X4), x2);
                v[2] = __dsub_rn(__dadd_rn(__dadd_rn(-X1, X2), X4), x3);
                double sx = __dadd_rn(__dadd_rn(__dadd_rn(0.0, x1), x2), x3);
                v[3] = __dsub_rn(__dadd_rn(__dsub_rn(__dsub_rn(-X1, X2), X4), sx), 1.0);
                const u64 top = (dens >= 3) ? 0x8000000000000000ull : 0ull;
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (v[t] >= a.thres_viol) { k[t] = top | (u64)__double_as_longlong(v[t]); ++nviol; }
            }
            ulonglong2* dst = reinterpret_cast<ulonglong2*>(a.key + 4 * r);
            dst[0] = make_ulonglong2(k[0], k[1]);
            dst[1] = make_ulonglong2(k[2], k[3]);
        }
        if (g + 1 < g1) {
            if (!(valid && lex_advance<3>(a.n, c, 32))) { c[0] = 0; c[1] = 1; c[2] = 2; }
        }
    }
    for (int o = 16; o; o >>= 1) { nkept += __shfl_xor_sync(0xffffffffu, nkept, o); nviol += __shfl_xor_sync(0xffffffffu, nviol, o); }
    if (lane == 0) {
        if (nkept) atomicAdd(&a.counters[0], (unsigned long long)nkept);
        if (nviol) atomicAdd(&a.counters[1], (unsigned long long)nviol);
    }
}

__global__ void k_tri_unpack(i64 m, const u64* s_k1, const i64* s_idx, i64* o_rank, int8_t* o_type, double* o_viol,
                             int8_t* o_dens)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    u64 k = s_k1[i];
    o_rank[i] = s_idx[i] >> 2;
    o_type[i] = (int8_t)(s_idx[i] & 3);
    o_dens[i] = (k >> 63) ? 3 : 2;
    o_viol[i] = __longlong_as_double((long long)(k & 0x7fffffffffffffffull));
}

// ---------------------------------------------------------------------------------------------------
// K7: eigenvector cut of one subset (cut_select_qp.py:737-751). One thread per selected subset.
// ---------------------------------------------------------------------------------------------------
struct CutArgs {
    int n; int rho; i64 m;
    const int16_t* sets;   // m x rho, -1 padded
    const double* X; const double* x;
    double thr_eig; int sweeps;
    i64* ind; double* val; double* rhs; double* lam; uint8_t* violated;   // ind/val: m x width
    double* gap;           // second smallest eigenvalue minus lam_min: ~0 means the eigenvector is not unique
};

template <int D>
__device__ void gen_cut_one(const CutArgs& a, i64 i, const int16_t* s)
{
    constexpr int M = D + 1;
    const int width = a.rho + a.rho * (a.rho + 1) / 2;
    const i64 nb_lifted = (i64)a.n * (a.n + 1) / 2;
    int c[D];
#pragma unroll
    for (int t = 0; t < D; ++t) c[t] = s[t];
    double A[M][M], V[M][M];
    A[0][0] = 1.0;
#pragma unroll
    for (int p = 0; p < D; ++p) A[0][p + 1] = a.x[c[p]];
#pragma unroll
    for (int p = 0; p < D; ++p)
#pragma unroll
        for (int q = p; q < D; ++q) A[p + 1][q + 1] = a.X[tri_index(a.n, c[p], c[q])];
    jacobi_sweeps<M, true>(A, V, a.sweeps + 2);
    int best = 0;
    double lam = A[0][0];
#pragma unroll
    for (int p = 1; p < M; ++p)
        if (A[p][p] < lam) { lam = A[p][p]; best = p; }
    double second = 1e300;
#pragma unroll
    for (int p = 0; p < M; ++p)
        if (p != best) second = fmin(second, A[p][p]);
    a.gap[i] = second - lam;
    double v[M];
#pragma unroll
    for (int p = 0; p < M; ++p) {
        double e = 0.0;
#pragma unroll
        for (int q = 0; q < M; ++q) e = (q == best) ? V[p][q] : e;
        v[p] = (fabs(e) <= -a.thr_eig) ? 0.0 : e;   // np.where(abs(evect) <= 1e-15, 0, evect)
    }
    i64* ind = a.ind + i * width;
    double* val = a.val + i * width;
    int k = 0;
#pragma unroll
    for (int p = 0; p < D; ++p) { ind[k] = nb_lifted + c[p]; val[k] = v[0] * v[p + 1] * 2.0; ++k; }
#pragma unroll
    for (int p = 0; p < D; ++p)
#pragma unroll
        for (int q = p; q < D; ++q) {
            ind[k] = tri_index(a.n, c[p], c[q]);
            val[k] = (p != q) ? v[p + 1] * v[q + 1] * 2.0 : v[p + 1] * v[q + 1];
            ++k;
        }
    for (; k < width; ++k) { ind[k] = -1; val[k] = 0.0; }
    a.rhs[i] = -v[0] * v[0];
    a.lam[i] = lam;
    a.violated[i] = lam < a.thr_eig;
}

__global__ void __launch_bounds__(128) k_gen_cuts(CutArgs a)
{
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.m) return;
    const int16_t* s = a.sets + i * a.rho;
    int d = 0;
    for (int t = 0; t < a.rho; ++t) d += (s[t] >= 0);
    switch (d) {
    case 2: gen_cut_one<2>(a, i, s); break;
    case 3: gen_cut_one<3>(a, i, s); break;
    case 4: gen_cut_one<4>(a, i, s); break;
    case 5: gen_cut_one<5>(a, i, s); break;
    default: a.violated[i] = 0; a.lam[i] = 0; a.rhs[i] = 0; a.gap[i] = 0; break;
    }
}

// single [1 x^T; x X] eigendecomposition: vals (unsorted) and V (columns = eigenvectors), order d+1
template <int D>
__device__ void eig_one(const double* pt, const double* Xs, int sweeps, double* vals, double* vecs)
{
    constexpr int M = D + 1;
    double A[M][M], V[M][M];
    A[0][0] = 1.0;
    for (int p = 0; p < D; ++p) A[0][p + 1] = pt[p];
    int k = 0;
#pragma unroll
    for (int p = 0; p < D; ++p)
#pragma unroll
        for (int q = p; q < D; ++q) A[p + 1][q + 1] = Xs[k++];
    jacobi_sweeps<M, true>(A, V, sweeps + 2);
#pragma unroll
    for (int p = 0; p < M; ++p) {
        vals[p] = A[p][p];
#pragma unroll
        for (int q = 0; q < M; ++q) vecs[p * M + q] = V[p][q];
    }
}

__global__ void k_eig_one(int d, const double* pt, const double* Xs, int sweeps, double* vals, double* vecs)
{
    switch (d) {
    case 2: eig_one<2>(pt, Xs, sweeps, vals, vecs); break;
    case 3: eig_one<3>(pt, Xs, sweeps, vals, vecs); break;
    case 4: eig_one<4>(pt, Xs, sweeps, vals, vecs); break;
    case 5: eig_one<5>(pt, Xs, sweeps, vals, vecs); break;
    }
}

// ---------------------------------------------------------------------------------------------------
// FP64 peak probes (roofline denominators; MEASURED_PEAKS.json has no FP64 entry)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_peak_dfma(double* out, const double* in, int iters)
{
    double x = in[threadIdx.x & 31], y = in[32 + (threadIdx.x & 31)];
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fma(acc[j], x, y);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[j];
    if (s == 123.456) out[threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_peak_dmma(double* out, const double* in, int iters)
{
    double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    double c0[8], c1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { c0[j] = j; c1[j] = -j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(c0[j], c1[j], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c0[j] + c1[j];
    if (s == 123.456) out[threadIdx.x] = s;
}

}  // namespace sdpcs
