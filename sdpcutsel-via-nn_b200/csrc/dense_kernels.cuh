// Dense eigenvalue cuts (strat 0): CutSolver.__gen_dense_eigcuts, cut_select_qp.py:757-786.
// One symmetric eigen-decomposition of the full matrix [1 x^T; x X] of order m = n + 1 <= 256 per round, then one
// dense row per negative eigenvalue.  LAPACK dsyevd (numpy eigh) in the reference; here a parallel two-sided Jacobi
// iteration in one CTA: the m / 2 disjoint pairs of a round-robin step are rotated together (Brent-Luk ordering),
// rows, columns and the eigenvector matrix are updated by all 1024 threads, the matrix stays in L2 (<= 0.5 MB).
// Jacobi keeps the eigenvectors orthogonal to working precision also inside clusters of eigenvalues, which is what
// the cuts need (v v^T is taken eigenvector by eigenvector).
#pragma once
#include "device_math.cuh"

namespace sdpcs {

constexpr int DENSE_THREADS = 1024;
constexpr int DENSE_MAX_ORDER = 256;

struct DenseEigArgs {
    int n;                 // variables; matrix order m = n + 1, padded to the even mp
    int mp;
    const double* X;       // upper triangle, row-major (cut_select_qp.py:333-345)
    const double* x;
    double* A;             // mp x mp work matrix
    double* V;             // mp x mp eigenvectors (columns)
    double* eig;           // mp eigenvalues (diagonal of A at the end), unsorted
    int max_sweeps;
    int* sweeps_done;
};

__global__ void __launch_bounds__(DENSE_THREADS, 1) k_dense_jacobi(DenseEigArgs a)
{
    __shared__ double sc[DENSE_MAX_ORDER / 2], ss[DENSE_MAX_ORDER / 2];
    __shared__ int sp[DENSE_MAX_ORDER / 2], sq[DENSE_MAX_ORDER / 2];
    __shared__ double red[DENSE_THREADS / 32];
    __shared__ double s_off, s_tot;
    const int tid = threadIdx.x, mp = a.mp, m = a.n + 1, half = mp / 2;
    // fill: M[0][0] = 1, M[0][j] = x_{j-1}, M[i][j] = X_{i-1, j-1}; only the upper triangle is given (eigh "U")
    for (int e = tid; e < mp * mp; e += DENSE_THREADS) {
        const int i = e / mp, j = e % mp;
        double v = 0.0;
        if (i < m && j < m) {
            const int lo = i < j ? i : j, hi = i < j ? j : i;
            if (hi == 0) v = 1.0;
            else if (lo == 0) v = a.x[hi - 1];
            else v = a.X[tri_index(a.n, lo - 1, hi - 1)];
        }
        a.A[e] = v;
        a.V[e] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    int sweep = 0;
    for (; sweep < a.max_sweeps; ++sweep) {
        // off-diagonal norm: converged when it is at rounding level of the whole matrix
        double off = 0.0, tot = 0.0;
        for (int e = tid; e < mp * mp; e += DENSE_THREADS) {
            const double v = a.A[e];
            tot += v * v;
            if (e / mp != e % mp) off += v * v;
        }
        for (int pass = 0; pass < 2; ++pass) {
            double v = pass ? tot : off;
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0) red[tid >> 5] = v;
            __syncthreads();
            if (tid < 32) {
                double w = tid < DENSE_THREADS / 32 ? red[tid] : 0.0;
                for (int o = 16; o; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
                if (tid == 0) { if (pass) s_tot = w; else s_off = w; }
            }
            __syncthreads();
        }
        if (s_off <= 1e-30 * s_tot) break;      // off-diagonal norm below 1e-15 of the matrix norm
        for (int step = 0; step < mp - 1; ++step) {
            if (tid < half) {
                int p, q;
                if (tid == 0) { p = mp - 1; q = step; }
                else { p = (step + tid) % (mp - 1); q = (step - tid + (mp - 1)) % (mp - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                const double app = a.A[p * mp + p], aqq = a.A[q * mp + q], apq = a.A[p * mp + q];
                double c = 1.0, s = 0.0;
                if (apq != 0.0 && fabs(apq) > 1e-300) {
                    const double tau = (aqq - app) / (2.0 * apq);
                    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c = 1.0 / sqrt(1.0 + t * t);
                    s = t * c;
                }
                sc[tid] = c; ss[tid] = s; sp[tid] = p; sq[tid] = q;
            }
            __syncthreads();
            // rows p, q of every pair:  A <- J^T A
            for (int e = tid; e < half * mp; e += DENSE_THREADS) {
                const int k = e / mp, j = e % mp;
                const double c = sc[k], s = ss[k];
                if (s == 0.0) continue;
                double* rp = a.A + sp[k] * mp + j;
                double* rq = a.A + sq[k] * mp + j;
                const double vp = *rp, vq = *rq;
                *rp = c * vp - s * vq;
                *rq = s * vp + c * vq;
            }
            __syncthreads();
            // columns p, q of every pair:  A <- A J,  V <- V J
            for (int e = tid; e < half * mp; e += DENSE_THREADS) {
                const int i = e / half, k = e % half;
                const double c = sc[k], s = ss[k];
                if (s == 0.0) continue;
                const int p = sp[k], q = sq[k];
                double* row = a.A + i * mp;
                const double vp = row[p], vq = row[q];
                row[p] = c * vp - s * vq;
                row[q] = s * vp + c * vq;
                double* vrow = a.V + i * mp;
                const double up = vrow[p], uq = vrow[q];
                vrow[p] = c * up - s * uq;
                vrow[q] = s * up + c * uq;
            }
            __syncthreads();
            if (tid < half && ss[tid] != 0.0) {      // the rotated element is zero by construction
                a.A[sp[tid] * mp + sq[tid]] = 0.0;
                a.A[sq[tid] * mp + sp[tid]] = 0.0;
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < mp; i += DENSE_THREADS) a.eig[i] = a.A[i * mp + i];
    if (tid == 0) *a.sweeps_done = sweep;
}

// One dense row per selected eigenvector (cut_select_qp.py:774-781):
//   [2 v0 v_i, i = 1..n  |  v_i v_j (x 2 if i != j), 1 <= i <= j <= n, row-major]  >=  -v0^2
// grid = (rows of the triangle + 1, cuts); block row 0 writes the x part and the right-hand side.
__global__ void __launch_bounds__(256) k_dense_rows(int n, int mp, const double* V, const int* cols, double* val, double* rhs)
{
    const int cut = blockIdx.y, col = cols[cut];
    const i64 width = (i64)n + (i64)n * (n + 1) / 2;
    double* out = val + (i64)cut * width;
    const double v0 = V[col];
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = v0 * V[(i + 1) * mp + col] * 2.0;
        if (threadIdx.x == 0) rhs[cut] = -v0 * v0;
    } else {
        const int i = blockIdx.x - 1;                       // X row i (matrix index i + 1)
        const double vi = V[(i + 1) * mp + col];
        double* o = out + n + tri_index(n, i, i);
        for (int j = i + threadIdx.x; j < n; j += blockDim.x) {
            const double p = vi * V[(j + 1) * mp + col];
            o[j - i] = (j == i) ? p : p * 2.0;
        }
    }
}

}  // namespace sdpcs
