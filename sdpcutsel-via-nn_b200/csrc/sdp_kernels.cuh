// Exact optimality measure (strat 3 and figure 8 of the reference: Mosek per sub-problem, cut_select_qp.py:555-567,
// 584-598, 660-686; training-data sampler utilities.py:14-59) as a batched register-resident solver of the small SDP
//
//     v(x, C) = min <C, X>   s.t.  [[X, x], [x^T, 1]] >= 0 (PSD),  diag(X) <= x            (order d + 1 <= 6)
//
// one thread per sub-problem, FP64.  With Y = X - x x^T the constraint reads Y >= 0, Y_ii <= u_i = x_i (1 - x_i), and
// with W = D^-1 Y D^-1, D = diag(sqrt(u)), A = D C D:
//     v = x^T C x + min { <A, W> : W >= 0, W_ii <= 1 }  =  x^T C x - min { sum_i y_i : y >= 0, A + Diag(y) >= 0 }
// (Lagrangian dual; Slater holds on both sides).  The dual is a d-variable convex problem; it is solved by a
// path-following barrier method on  F_mu(y) = sum(y) / mu - log det(A + Diag(y)) - sum_i log y_i  with damped Newton
// steps t = 1 / (1 + lambda) (lambda = Newton decrement; F is self-concordant, so every iterate stays strictly
// feasible without a line search), mu shrinking by 10 once lambda < 1/2, down to mu_final: the returned value lies
// within d * mu_final of the optimum (duality gap 2 d mu).  ~50 Newton steps of ~500 flop.  Coordinates with u_i = 0
// (x_i in {0, 1}) decouple by themselves (row and column of A vanish).
// Mosek's own answers (data_figures/fig8_data.csv, 1051 sub-problems) are reproduced to 4e-6 -- its tolerance.
#pragma once
#include "score_kernels.cuh"

namespace sdpcs {

// Cholesky factor (lower, in place in the lower triangle of a) of an SPD matrix of order D; returns false if a pivot is
// not positive.  a is indexed a[i][j], j <= i.
template <int D>
__host__ __device__ __forceinline__ bool chol_lower(double (&a)[D][D])
{
    bool ok = true;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double s = a[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-a[j][k], a[j][k], s);
        ok = ok && (s > 0.0);
        const double r = 1.0 / sqrt(fmax(s, 1e-300));
        a[j][j] = s * r;                                   // sqrt(s)
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            double t = a[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma(-a[i][k], a[j][k], t);
            a[i][j] = t * r;
        }
    }
    return ok;
}

// inverse of the SPD matrix whose Cholesky factor L (lower) is given: out = L^-T L^-1 (full symmetric)
template <int D>
__host__ __device__ __forceinline__ void chol_inverse(const double (&L)[D][D], double (&out)[D][D])
{
    double Li[D][D];                                       // L^-1, lower
#pragma unroll
    for (int j = 0; j < D; ++j) {
        Li[j][j] = 1.0 / L[j][j];
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            double s = 0.0;
#pragma unroll
            for (int k = j; k < i; ++k) s = fma(L[i][k], Li[k][j], s);
            Li[i][j] = -s / L[i][i];
        }
    }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = i; k < D; ++k) s = fma(Li[k][i], Li[k][j], s);
            out[i][j] = s;
            out[j][i] = s;
        }
}

// solve H z = b for SPD H (destroyed); returns false if H is not numerically positive definite
template <int D>
__host__ __device__ __forceinline__ bool spd_solve(double (&H)[D][D], const double (&b)[D], double (&z)[D])
{
    const bool ok = chol_lower<D>(H);
    double w[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s = fma(-H[i][k], w[k], s);
        w[i] = s / H[i][i];
    }
#pragma unroll
    for (int i = D - 1; i >= 0; --i) {
        double s = w[i];
#pragma unroll
        for (int k = i + 1; k < D; ++k) s = fma(-H[k][i], z[k], s);
        z[i] = s / H[i][i];
    }
    return ok;
}

// v(x, C): C given as the upper triangle in row-major order with the off-diagonal entries carrying the full weight of
// the pair (i, j), i.e. <C, X> = sum_{i <= j} Cu_ij X_ij -- the Q_slice convention of cut_select_qp.py:530-538, 592-593.
template <int D>
__host__ __device__ __forceinline__ double sdp_value(const double (&x)[D], const double (&Cu)[D * (D + 1) / 2], double mu_final, int* iters)
{
    double A[D][D], y[D], sq[D];
    double cst = 0.0;
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < D; ++i) sq[i] = sqrt(fmax(x[i] - x[i] * x[i], 0.0));
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = i; j < D; ++j) {
                const double c = (i == j) ? Cu[k] : 0.5 * Cu[k];
                cst = fma((i == j) ? c : 2.0 * c, x[i] * x[j], cst);
                A[i][j] = A[j][i] = c * sq[i] * sq[j];
                ++k;
            }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = 1.0;
#pragma unroll
        for (int j = 0; j < D; ++j) s += fabs(A[i][j]);
        y[i] = s;                                          // strictly diagonally dominant start: A + Diag(y) > 0
    }
    double mu = 1.0;
    int it = 0;
#pragma unroll 1
    for (; it < 400; ++it) {
        double L[D][D], Si[D][D], H[D][D], g[D], dy[D];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) L[i][j] = A[i][j] + ((i == j) ? y[i] : 0.0);
        chol_lower<D>(L);
        chol_inverse<D>(L, Si);
        const double rmu = 1.0 / mu;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double ry = 1.0 / y[i];
            g[i] = -(rmu - (Si[i][i] + ry));               // right-hand side: -gradient of F_mu
#pragma unroll
            for (int j = 0; j <= i; ++j) H[i][j] = Si[i][j] * Si[i][j] + ((i == j) ? ry * ry : 0.0);
        }
        spd_solve<D>(H, g, dy);
        double l2 = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) l2 = fma(g[i], dy[i], l2);    // lambda^2 = g^T H^-1 g
        const double lam = sqrt(fmax(l2, 0.0));
        const double t = (lam > 0.25) ? 1.0 / (1.0 + lam) : 1.0;
#pragma unroll
        for (int i = 0; i < D; ++i) y[i] = fma(t, dy[i], y[i]);
        if (lam < 0.5) {
            if (mu <= mu_final) { ++it; break; }
            mu = fmax(0.1 * mu, mu_final);
        }
    }
    if (iters) *iters = it;
    double sy = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) sy += y[i];
    return cst - sy + (double)D * mu;                      // midpoint of [dual value, dual value + 2 D mu]
}

// exact optimality measure of every candidate: obj = max_elem * (v(x_rho, Q~_rho) - <Q~_rho, X_rho>), the value
// cut_select_qp.py:575 + 595 computes with Mosek.  Same work distribution as k_score_feas.
template <int D>
__global__ void __launch_bounds__(128) k_score_sdp(ScoreArgs a, double mu_final)
{
    constexpr int T = D * (D + 1) / 2;
    const int lane = threadIdx.x & 31;
    const i64 warps_total = (i64)gridDim.x * (blockDim.x >> 5);
    const i64 gw = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const i64 G = (a.N + 31) >> 5;
    const i64 g0 = gw * G / warps_total, g1 = (gw + 1) * G / warps_total;
    if (g0 >= g1) return;
    int c[D];
    const bool all_mode = (a.idx == nullptr);
    if (all_mode) {
        const i64 i0 = g0 * 32 + lane;
        if (i0 < a.N) lex_unrank<D>(a.n, (u64)(a.rank_begin + i0), c);
        else {
#pragma unroll
            for (int t = 0; t < D; ++t) c[t] = t;
        }
    }
    for (i64 g = g0; g < g1; ++g) {
        const i64 i = g * 32 + lane;
        const bool valid = i < a.N;
        if (!all_mode) load_list_indices<D>(a.idx, i, valid, c);
        double xs[D], Xs[T], Qs[T];
        gather_point<D>(a, c, xs, Xs);
        int k = 0;
        double mx = 0.0;
#pragma unroll
        for (int p = 0; p < D; ++p)
#pragma unroll
            for (int q = p; q < D; ++q) {
                const double v = __ldg(a.Q + tri_index(a.n, c[p], c[q]));
                Qs[k++] = v;
                mx = fmax(mx, fabs(v));
            }
        double max_elem = (double)D * mx;                  // cut_select_qp.py:536-538
        if (max_elem == 0.0) max_elem = 1.0;
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < T; ++q) {
            Qs[q] = __ddiv_rn(Qs[q], max_elem);
            s = __dadd_rn(s, __dmul_rn(Qs[q], Xs[q]));     // left-to-right, no FMA (cut_select_qp.py:575)
        }
        if (valid) {
            const double v = sdp_value<D>(xs, Qs, mu_final, nullptr);
            const double obj = __dadd_rn(__dmul_rn(-s, max_elem), __dmul_rn(v, max_elem));
            a.obj[a.pos ? __ldg(a.pos + i) : i] = obj;
        }
        if (all_mode && g + 1 < g1) {
            if (!(valid && lex_advance<D>(a.n, c, 32))) {
#pragma unroll
                for (int t = 0; t < D; ++t) c[t] = t;
            }
        }
    }
}

// raw batch: m sub-problems given as rows [x (D) | C upper triangle (D(D+1)/2)] -> v(x, C)   (sdpcs_sdp_solve)
template <int D>
__global__ void __launch_bounds__(128) k_sdp_solve(const double* in, i64 m, double mu_final, double* out, int* iters)
{
    constexpr int T = D * (D + 1) / 2;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double xs[D], Cu[T];
#pragma unroll
    for (int t = 0; t < D; ++t) xs[t] = in[i * (D + T) + t];
#pragma unroll
    for (int t = 0; t < T; ++t) Cu[t] = in[i * (D + T) + D + t];
    int it = 0;
    out[i] = sdp_value<D>(xs, Cu, mu_final, &it);
    if (iters) iters[i] = it;
}

}  // namespace sdpcs
