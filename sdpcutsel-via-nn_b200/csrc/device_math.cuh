// Device-side building blocks of the cut-selection path (sm_100a):
//   K1  lexicographic subset unranking / advancing (replaces the nested loops of cut_select_qp.py:451-455 and
//       the stored agg_list of :524-540 -- no index list is kept in HBM in all-subsets mode)
//   K3  register-resident cyclic-Jacobi symmetric eigensolver for orders 3..6 in FP64
//       (replaces np.linalg.eigvalsh/eigh(M, "U") at cut_select_qp.py:796-797)
//   tansig(n) = 2/(1+exp(-2n))-1 (neural_net_3D.m:76-78) in 13 FP64-pipe operations
//   DMMA.8x8x4 wrapper (mma.sync.m8n8k4.f64), the native FP64 tensor shape on sm_100a
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef long long i64;

namespace sdpcs {

// ---------------------------------------------------------------------------------------------------
// combinatorics
// ---------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u64 binom_small(int m, int j)
{
    // C(m, j) for j <= 5, m <= 250 (C(250,5) = 7.8e9; the products below stay < 2^63)
    if (m < j) return 0;
    u64 M = (u64)m;
    switch (j) {
    case 0: return 1;
    case 1: return M;
    case 2: return M * (M - 1) / 2;
    case 3: return M * (M - 1) * (M - 2) / 6;
    case 4: return M * (M - 1) * (M - 2) * (M - 3) / 24;
    default: return M * (M - 1) * (M - 2) * (M - 3) / 24 * (M - 4) / 5;
    }
}

// rank -> ascending tuple in lexicographic order (= itertools.combinations order). O(n + D) steps.
template <int D>
__host__ __device__ __forceinline__ void lex_unrank(int n, u64 r, int (&c)[D])
{
    int v = 0;
#pragma unroll
    for (int j = 1; j <= D; ++j) {
        while (true) {
            u64 cnt = binom_small(n - 1 - v, D - j);
            if (r < cnt) break;
            r -= cnt;
            ++v;
        }
        c[j - 1] = v;
        ++v;
    }
}

// Move `delta` positions forward in lexicographic order (delta small, e.g. 32). Returns false once the
// end of the enumeration is passed (c is then left at a valid but meaningless tuple).
template <int D>
__device__ __forceinline__ bool lex_advance(int n, int (&c)[D], int delta)
{
    int rem = delta;
    while (true) {
        int room = (n - 1) - c[D - 1];
        if (rem <= room) {
            c[D - 1] += rem;
            return true;
        }
        rem -= room + 1;
        bool done = false;
#pragma unroll
        for (int j = D - 2; j >= 0; --j) {
            if (!done && c[j] < n - D + j) {
                ++c[j];
#pragma unroll
                for (int t = j + 1; t < D; ++t) c[t] = c[t - 1] + 1;
                done = true;
            }
        }
        if (!done) {
#pragma unroll
            for (int t = 0; t < D; ++t) c[t] = t;
            return false;
        }
    }
}

// flat upper-triangular index of (i, j), i <= j: n*i - i(i+1)/2 + j   (cut_select_qp.py:531)
__host__ __device__ __forceinline__ int tri_index(int n, int i, int j) { return n * i - ((i * (i + 1)) >> 1) + j; }

// ---------------------------------------------------------------------------------------------------
// order-preserving double <-> u64 key (larger double -> larger key); key 0 is reserved for "excluded"
// ---------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u64 enc_key(double x)
{
    x = x + 0.0;     // -0.0 -> +0.0: Python's sort (the reference) compares them equal, the tie then goes to the index
    u64 b;
#ifdef __CUDA_ARCH__
    b = (u64)__double_as_longlong(x);
#else
    memcpy(&b, &x, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dec_key(u64 k)
{
    u64 b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------
// fast FP64 reciprocal / reciprocal square root: MUFU.RCP64H / MUFU.RSQ64H seed (~20 bits) + one cubic step.
// No subnormal / zero / inf slow path: callers guarantee a normal positive argument.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double b)
{
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(b));
    double e = fma(-b, y0, 1.0);
    return fma(y0, fma(e, e, e), y0);
}
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    double t = x * y0;
    double e = fma(-t, y0, 1.0);
    double p = fma(0.375, e, 0.5);
    return fma(y0 * e, p, y0);
}
// a / b with r = fast_rcp(b) shared by several quotients: one correction step on the residual.
__device__ __forceinline__ double div_by(double a, double b, double r)
{
    double q = a * r;
    return fma(fma(-q, b, a), r, q);
}

// ---------------------------------------------------------------------------------------------------
// K3: cyclic Jacobi, order M, fully unrolled so the matrix lives in registers.
// a[i][j] is used for i <= j only. On return the diagonal holds the eigenvalues (unsorted).
// Rotation (p,q) with b = a_pq, d = a_qq - a_pp:  t = sgn(d) 2b / (|d| + sqrt(d^2 + 4b^2)),
// c = 1/sqrt(1+t^2), s = t c -- one rsqrt + one rcp + one rsqrt, 39 FP64 operations per rotation at M = 6.
// ---------------------------------------------------------------------------------------------------
template <int M, bool VEC>
__device__ __forceinline__ void jacobi_sweeps(double (&a)[M][M], double (&v)[M][M], int sweeps)
{
    if (VEC) {
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < M; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
    }
#pragma unroll 1
    for (int sw = 0; sw < sweeps; ++sw) {
#pragma unroll
        for (int p = 0; p < M - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < M; ++q) {
                const double b = a[p][q];
                const double app = a[p][p], aqq = a[q][q];
                const double d = aqq - app;
                const double b2 = b + b;
                const double S = fma(d, d, b2 * b2);
                if (S > 1e-290 && b != 0.0) {
                    const double h = S * fast_rsqrt(S);
                    const double rd = fast_rcp(fabs(d) + h);
                    const double t = copysign(b2 * rd, (d < 0.0) ? -b2 : b2);
                    const double c = fast_rsqrt(fma(t, t, 1.0));
                    const double s = t * c;
                    a[p][p] = fma(-t, b, app);
                    a[q][q] = fma(t, b, aqq);
                    a[p][q] = 0.0;
#pragma unroll
                    for (int r = 0; r < M; ++r) {
                        if (r != p && r != q) {
                            double& arp = (r < p) ? a[r][p] : a[p][r];
                            double& arq = (r < q) ? a[r][q] : a[q][r];
                            const double x = arp, y = arq;
                            arp = fma(-s, y, c * x);
                            arq = fma(s, x, c * y);
                        }
                    }
                    if (VEC) {
#pragma unroll
                        for (int r = 0; r < M; ++r) {
                            const double x = v[r][p], y = v[r][q];
                            v[r][p] = fma(-s, y, c * x);
                            v[r][q] = fma(s, x, c * y);
                        }
                    }
                }
            }
        }
    }
}

__host__ __device__ __forceinline__ int default_sweeps(int m) { return m <= 4 ? 5 : 6; }

// lam_min of [[1, x^T],[x, X]] for a subset of size D (matrix order D+1). Xs is the subset's upper
// triangle in combinations_with_replacement order (cut_select_qp.py:530, 792-794).
// sweeps > 0: cyclic Jacobi with that many sweeps; sweeps == 0: tridiagonalisation + Laguerre (default).
template <int D>
__device__ __forceinline__ void fill_subset_matrix(const double (&xs)[D], const double (&Xs)[D * (D + 1) / 2],
                                                   double (&a)[D + 1][D + 1])
{
    a[0][0] = 1.0;
#pragma unroll
    for (int i = 0; i < D; ++i) a[0][i + 1] = xs[i];
    int k = 0;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = i; j < D; ++j) a[i + 1][j + 1] = Xs[k++];
}

template <int D>
__device__ __forceinline__ double lam_min_subset_jacobi(const double (&xs)[D], const double (&Xs)[D * (D + 1) / 2], int sweeps)
{
    constexpr int M = D + 1;
    double a[M][M], dummy[M][M];
    fill_subset_matrix<D>(xs, Xs, a);
    jacobi_sweeps<M, false>(a, dummy, sweeps);
    double lam = a[0][0];
#pragma unroll
    for (int i = 1; i < M; ++i) lam = fmin(lam, a[i][i]);
    return lam;
}

// ---------------------------------------------------------------------------------------------------
// K3 (default): lam_min only, without an eigen-decomposition.
//   1. Householder tridiagonalisation of the order-M matrix in registers (M-2 reflections, ~224 FP64 ops at M=6).
//   2. Smallest root of the characteristic polynomial p(l) = det(T - l I) by Laguerre's iteration started left of
//      the spectrum (Gershgorin bound): for a polynomial with only real roots the iterates increase monotonically
//      to lam_min with cubic convergence (4-9 evaluations on LP points). p, p', p'' come from the three-term
//      recurrence; the same recurrence gives the Sturm test "all leading minors of T - l I positive" (l < lam_min),
//      which maintains a bracket [lo, hi]: an iterate that is not positive definite (only possible through
//      rounding, next to the root) becomes hi and the iteration continues from the midpoint, so the result is
//      unconditionally bracketed; multiple smallest eigenvalues merely fall back to linear convergence.
//   Converged when the Laguerre step is below 3e-16 * ||T|| (or the bracket is). Deterministic per matrix:
//   a lane stops updating once converged, the warp leaves the loop when all lanes have.
// ~830 FP64 operations per matrix at M = 6 instead of ~3,500 for six Jacobi sweeps.
// ---------------------------------------------------------------------------------------------------
template <int M>
__device__ __forceinline__ double lam_min_tridiag_laguerre(double (&A)[M][M])
{
    double b[M - 1];
#pragma unroll
    for (int k = 0; k < M - 2; ++k) {
        const double x0 = A[k][k + 1];
        double sigma = 0.0;
#pragma unroll
        for (int j = k + 2; j < M; ++j) sigma = fma(A[k][j], A[k][j], sigma);
        double bk = x0;
        if (sigma > 1e-290) {
            const double S = fma(x0, x0, sigma);
            const double norm = S * fast_rsqrt(S);
            const double alpha = (x0 > 0.0) ? -norm : norm;
            const double beta = fast_rcp(fma(norm, fabs(x0), S));   // 2 / v^T v
            double v[M], pv[M];
            v[k + 1] = x0 - alpha;
#pragma unroll
            for (int j = k + 2; j < M; ++j) v[j] = A[k][j];
            double K = 0.0;
#pragma unroll
            for (int i = k + 1; i < M; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int j = k + 1; j < M; ++j) acc = fma((i <= j) ? A[i][j] : A[j][i], v[j], acc);
                pv[i] = beta * acc;
                K = fma(v[i], pv[i], K);
            }
            K *= 0.5 * beta;
#pragma unroll
            for (int i = k + 1; i < M; ++i) pv[i] = fma(-K, v[i], pv[i]);   // w
#pragma unroll
            for (int i = k + 1; i < M; ++i)
#pragma unroll
                for (int j = i; j < M; ++j) A[i][j] = fma(-v[i], pv[j], fma(-pv[i], v[j], A[i][j]));
            bk = alpha;
        }
        b[k] = bk;
    }
    b[M - 2] = A[M - 2][M - 1];
    double a[M], c[M - 1];
#pragma unroll
    for (int i = 0; i < M; ++i) a[i] = A[i][i];
    double g = 1e300, gh = -1e300, amin = 1e300;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double rad = 0.0;
        if (i > 0) rad += fabs(b[i - 1]);
        if (i < M - 1) rad += fabs(b[i]);
        g = fmin(g, a[i] - rad);
        gh = fmax(gh, a[i] + rad);
        amin = fmin(amin, a[i]);
    }
#pragma unroll
    for (int i = 0; i < M - 1; ++i) c[i] = b[i] * b[i];
    const double tol = fmax(3e-16 * fmax(fabs(g), fabs(gh)), 1e-300);
    const double n = (double)M;
    double lo = g - 1e-3 * (1.0 + fabs(g)), hi = amin, lam = lo, result = lo;
    bool active = true;
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
        if (!__any_sync(0xffffffffu, active)) break;
        double d = a[0] - lam;
        double pp2 = 1.0, pp = d, dp2 = 0.0, dp1 = -1.0, ddp2 = 0.0, ddp1 = 0.0;
        bool pd = __double2hiint(pp) > 0;
#pragma unroll
        for (int i = 1; i < M; ++i) {
            d = a[i] - lam;
            const double p = fma(d, pp, -(c[i - 1] * pp2));
            const double dp = fma(d, dp1, fma(-c[i - 1], dp2, -pp));
            const double ddp = fma(d, ddp1, fma(-c[i - 1], ddp2, -(dp1 + dp1)));
            pd = pd && (__double2hiint(p) > 0);
            pp2 = pp; pp = p; dp2 = dp1; dp1 = dp; ddp2 = ddp1; ddp1 = ddp;
        }
        const double r = fast_rcp(pp);
        const double S1 = -dp1 * r;
        const double S2 = fma(S1, S1, -(ddp1 * r));
        double disc = (n - 1.0) * fma(n, S2, -(S1 * S1));
        disc = fmax(disc, 0.0);
        const double sq = (disc > 1e-300) ? disc * fast_rsqrt(disc) : 0.0;
        const double step = n * fast_rcp(S1 + sq);
        if (active) {
            if (pd) lo = lam; else hi = lam;
            if (!(fabs(step) > tol)) {                 // also true for NaN (p == 0: lam is the root)
                result = (pd && step == step) ? lam + step : lam;
                active = false;
            } else if (!(hi - lo > tol)) {
                result = lo;
                active = false;
            } else {
                double nxt = pd ? lam + step : 0.5 * (lo + hi);
                if (!(nxt < hi) || !(nxt > lo)) nxt = 0.5 * (lo + hi);
                lam = nxt;
                result = lo;
            }
        }
    }
    return result;
}

template <int D>
__device__ __forceinline__ double lam_min_subset(const double (&xs)[D], const double (&Xs)[D * (D + 1) / 2], int sweeps)
{
    if (sweeps > 0) return lam_min_subset_jacobi<D>(xs, Xs, sweeps);   // warp-uniform
    double a[D + 1][D + 1];
    fill_subset_matrix<D>(xs, Xs, a);
    return lam_min_tridiag_laguerre<D + 1>(a);
}

// ---------------------------------------------------------------------------------------------------
// DMMA.8x8x4 : D(8x8) += A(8x4, row) * B(4x8, col), FP64.  Fragment ownership (lane = 4*g + t):
//   a = A[g][t], b = B[t][g], c0 = C[g][2t], c1 = C[g][2t+1]
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------------
// tansig from the pre-scaled pre-activation zs = -2*log2(e)*n:  2/(1 + 2^zs) - 1  (= tanh(n)).
// 2^w with w = -|zs|: 8-bit table (T[j] = 2^(j/256), shared memory) + degree-4 polynomial; the reciprocal of
// d = 1 + 2^w in (1, 2] starts from an FP32 MUFU.RCP seed (bit-cast, no FP64 conversion instructions) refined
// by one cubic step. 13 FP64-pipe operations; max abs error 3.6e-16 (same as the reference formula with libm exp).
// ---------------------------------------------------------------------------------------------------
#define SDPCS_TANSIG_SCALE (-2.8853900817779268147)  /* -2*log2(e) */
__device__ __forceinline__ double tansig_scaled(double zs, const double* __restrict__ T)
{
    const double A1 = 0.6931471805599453094, A2 = 0.2402265069591007123, A3 = 0.0555041086648215800,
                 A4 = 0.0096181291076284772;
    const double MAGIC = 26388279066624.0;  // 1.5 * 2^44: ulp = 2^-8
    int hi = __double2hiint(zs);
    int sgn = hi & 0x80000000;
    double w = __hiloint2double(hi | 0x80000000, __double2loint(zs));
    if ((hi & 0x7fffffff) > 0x408F4000) w = -1000.0;
    double kf = w + MAGIC;
    int i = __double2loint(kf);
    double s = w - (kf - MAGIC);
    double q = fma(A4, s, A3);
    q = fma(q, s, A2);
    q = fma(q, s, A1);
    double p = q * s;
    double Tj = T[i & 255];
    double t0 = fma(Tj, p, Tj);
    double t = __hiloint2double(__double2hiint(t0) + ((i >> 8) << 20), __double2loint(t0));
    double d = 1.0 + t;
    unsigned fb = ((unsigned)(__double2hiint(d) - 0x38000000) << 3) | ((unsigned)__double2loint(d) >> 29);
    float yf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"(__uint_as_float(fb)));
    unsigned yb = __float_as_uint(yf);
    double y0 = __hiloint2double((int)((yb >> 3) + 0x38000000u), (int)(yb << 29));
    double e = fma(-d, y0, 1.0);
    double e2 = fma(e, e, e);
    double y = fma(y0, e2, y0);
    double r = fma(2.0, y, -1.0);
    return __hiloint2double(__double2hiint(r) ^ (sgn ^ 0x80000000), __double2loint(r));
}
#endif  // __CUDACC__

}  // namespace sdpcs
