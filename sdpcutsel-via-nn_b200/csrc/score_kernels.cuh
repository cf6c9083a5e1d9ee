// Scoring kernels: one fused pass per candidate subset
//   unrank (K1) -> gather x_rho, X_rho, Q_rho (K2) -> lam_min by register Jacobi (K3) -> NN_rhoD by DMMA (K4)
// replacing the per-subset Python loop of _sel_eigcut_by_ordering_on_measure (cut_select_qp.py:569-599, 639-652).
//
// Work distribution: candidates are cut into groups of 32 consecutive local indices; every warp owns a
// contiguous range of groups and walks it with lex_advance, so in all-subsets mode a lane unranks once.
// Gathers go through the read-only L1/L2 path: x, X and Q_arr (<= 2*251 KB + 2 KB for n <= 250) are read from
// HBM once per kernel and then live in L1/L2; consecutive lanes hold consecutive last indices, so the loads of
// X[c_a, c_last] are coalesced and those of X[c_a, c_b] (a, b < last) are warp broadcasts.
#pragma once
#include "device_math.cuh"

namespace sdpcs {

struct ScoreArgs {
    int n;
    i64 N;              // candidates handled by this launch
    i64 rank_begin;     // ALL mode: lex rank of local candidate 0
    const uint8_t* idx; // LIST mode: N x D subset indices (this size class only); nullptr = ALL mode
    const i64* pos;     // LIST mode: output position of each candidate; nullptr = identity
    const double* X;    // n(n+1)/2
    const double* x;    // n
    const double* Q;    // n(n+1)/2
    const double* wfrag;// fragment-ordered NN weights (see NetCfg), nullptr for the feasibility kernel
    double* lam;        // outputs (may be nullptr)
    double* obj;
    int sweeps;
};

template <int D>
struct NetCfg {
    static constexpr int NIN = D * (D + 3) / 2;
    static constexpr int KIN = (NIN + 3) / 4 * 4;             // input width padded to k-steps of 4
    static constexpr int H = (D == 2 || D == 5) ? 64 : 50;    // neural_net_{2,5}D: 64, {3,4}D: 50
    static constexpr int HP = (H + 7) / 8 * 8;
    static constexpr int NT = HP / 8;                         // 8-wide n-tiles
    static constexpr int NHID = (D == 5) ? 4 : 3;
    static constexpr int KS0 = KIN / 4;
    static constexpr int STRIDE = (KIN % 16 == 4 || KIN % 16 == 12) ? KIN : KIN + 4; // conflict-free A loads
    // smem blob layout (doubles)
    // within a k-step the NT fragments are stored pairwise so one LDS.128 feeds two DMMAs:
    //   frag(nt, lane) = (nt/2)*64 + lane*2 + (nt&1) for nt < 2*(NT/2), else (NT-1)*32 + lane
    static constexpr int frag(int nt, int lane) { return nt < 2 * (NT / 2) ? (nt / 2) * 64 + lane * 2 + (nt & 1) : (NT - 1) * 32 + lane; }
    static constexpr int OFF_W0 = 0;                                   // [KS0][frag(NT, lane)]
    static constexpr int LAYER = 2 * NT * NT * 32;                     // doubles per hidden-layer weight block
    static constexpr int OFF_WH = OFF_W0 + KS0 * NT * 32;              // [NHID-1][2*NT][frag(NT, lane)]
    static constexpr int OFF_BIAS = OFF_WH + (NHID - 1) * LAYER;       // [NHID][HP]  (scaled)
    static constexpr int OFF_WOUT = OFF_BIAS + NHID * HP;              // [HP]
    static constexpr int OFF_XOFF = OFF_WOUT + HP;                     // [KIN]
    static constexpr int OFF_GAIN = OFF_XOFF + KIN;                    // [KIN]
    static constexpr int OFF_MISC = OFF_GAIN + KIN;                    // bout, y_gain, y_xoffset, pad
    static constexpr int OFF_TAB = OFF_MISC + 4;                       // [256] 2^(j/256)
    static constexpr int BLOB = OFF_TAB + 256;
};

// ---------------------------------------------------------------------------------------------------
// candidate indices for lane-local candidate i
// ---------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void load_list_indices(const uint8_t* idx, i64 i, bool valid, int (&c)[D])
{
#pragma unroll
    for (int t = 0; t < D; ++t) c[t] = valid ? (int)__ldg(idx + i * D + t) : t;
}

template <int D>
__device__ __forceinline__ void gather_point(const ScoreArgs& a, const int (&c)[D], double (&xs)[D],
                                             double (&Xs)[D * (D + 1) / 2])
{
#pragma unroll
    for (int i = 0; i < D; ++i) xs[i] = __ldg(a.x + c[i]);
    int k = 0;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = i; j < D; ++j) Xs[k++] = __ldg(a.X + tri_index(a.n, c[i], c[j]));
}

// ---------------------------------------------------------------------------------------------------
// feasibility-only kernel: lam_min for every candidate
// ---------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) k_score_feas(ScoreArgs a)
{
    const int lane = threadIdx.x & 31;
    const i64 warps_total = (i64)gridDim.x * (blockDim.x >> 5);
    const i64 gw = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const i64 G = (a.N + 31) >> 5;
    const i64 g0 = gw * G / warps_total, g1 = (gw + 1) * G / warps_total;
    if (g0 >= g1) return;
    int c[D];
    const bool all_mode = (a.idx == nullptr);
    if (all_mode) {
        i64 i0 = g0 * 32 + lane;
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
        double xs[D], Xs[D * (D + 1) / 2];
        gather_point<D>(a, c, xs, Xs);
        double lam = lam_min_subset<D>(xs, Xs, a.sweeps);
        if (valid) a.lam[a.pos ? __ldg(a.pos + i) : i] = lam;
        if (all_mode && g + 1 < g1) {
            if (!(valid && lex_advance<D>(a.n, c, 32))) {
#pragma unroll
                for (int t = 0; t < D; ++t) c[t] = t;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K4: NN_rhoD on 8 rows (subsets) per pass via DMMA.8x8x4, one rolled loop over the tansig layers.
//   rows  : 8 staging rows [8][STRIDE] of this pass holding the mapminmax'ed NN inputs (zero padded to KIN)
//   w     : the fragment-ordered weight blob in shared memory
// Layer l+1 consumes the C fragments of layer l directly as A fragments: the C fragment of n-tile nt holds
// columns 8nt+2t+{0,1} for lane (g,t), so k-step (nt,h) of the next layer is *defined* as the neuron set
// {8nt+2t+h : t=0..3} and the weight fragments are pre-permuted accordingly on the host. Activations never
// leave registers between layers; layer 0 takes its A fragments from the staging rows through the same
// registers. The weights are pre-scaled by -2*log2(e) so the accumulator is directly the tansig argument.
// Returns the NN output of row g, valid in all four lanes of the quad.
// ---------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ double mlp8(const double* __restrict__ w, const double* __restrict__ rows, int lane)
{
    using C = NetCfg<D>;
    const int g = lane >> 2, t = lane & 3;
    double acc[C::NT][2];
    double act[C::NT][2];
    const double* tab = w + C::OFF_TAB;
    {
        const double* r = rows + g * C::STRIDE + t;
#pragma unroll
        for (int kt = 0; kt < C::NT; ++kt)
#pragma unroll
            for (int h = 0; h < 2; ++h) act[kt][h] = (2 * kt + h < C::KS0) ? r[4 * (2 * kt + h)] : 0.0;
    }
#pragma unroll 1
    for (int l = 0; l < C::NHID; ++l) {
        const double2* bias = reinterpret_cast<const double2*>(w + C::OFF_BIAS + l * C::HP) + t;
#pragma unroll
        for (int nt = 0; nt < C::NT; ++nt) {
            double2 b = bias[4 * nt];
            acc[nt][0] = b.x; acc[nt][1] = b.y;
        }
        const double* wl = w + ((l == 0) ? C::OFF_W0 : C::OFF_WH + (l - 1) * C::LAYER);
        const int nks = (l == 0) ? C::KS0 : 2 * C::NT;
#pragma unroll
        for (int kt = 0; kt < C::NT; ++kt) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ks = 2 * kt + h;
                if (ks < nks) {
                    const double a = act[kt][h];
                    const double* wk = wl + ks * (C::NT * 32);
#pragma unroll
                    for (int np = 0; np < C::NT / 2; ++np) {
                        double2 b = reinterpret_cast<const double2*>(wk + np * 64)[lane];
                        dmma884(acc[2 * np][0], acc[2 * np][1], a, b.x);
                        dmma884(acc[2 * np + 1][0], acc[2 * np + 1][1], a, b.y);
                    }
                    if (C::NT & 1) {
                        double b = wk[(C::NT - 1) * 32 + lane];
                        dmma884(acc[C::NT - 1][0], acc[C::NT - 1][1], a, b);
                    }
                }
            }
        }
#pragma unroll
        for (int nt = 0; nt < C::NT; ++nt) {
            act[nt][0] = tansig_scaled(acc[nt][0], tab);
            act[nt][1] = tansig_scaled(acc[nt][1], tab);
        }
    }
    // linear output layer + mapminmax_reverse (neural_net_3D.m:61-65, 81-85)
    const double2* wout = reinterpret_cast<const double2*>(w + C::OFF_WOUT) + t;
    double y = 0.0;
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt) {
        double2 wo = wout[4 * nt];
        y = fma(wo.x, act[nt][0], y);
        y = fma(wo.y, act[nt][1], y);
    }
    y += __shfl_xor_sync(0xffffffffu, y, 1);
    y += __shfl_xor_sync(0xffffffffu, y, 2);
    const double bout = w[C::OFF_MISC], y_gain = w[C::OFF_MISC + 1], y_xoff = w[C::OFF_MISC + 2];
    return ((bout + y) - -1.0) / y_gain + y_xoff;
}

template <int D>
constexpr int score_nn_smem_doubles(int warps) { return NetCfg<D>::BLOB + warps * (32 * NetCfg<D>::STRIDE + 32); }

// ---------------------------------------------------------------------------------------------------
// optimality-measure kernel: obj = max_elem * (NN(x_rho, Q~_rho) - <Q~_rho, X_rho>) for every candidate
// ---------------------------------------------------------------------------------------------------
template <int D, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_score_nn(ScoreArgs a)
{
    using C = NetCfg<D>;
    constexpr int T = D * (D + 1) / 2;
    extern __shared__ __align__(16) double smem[];
    for (int i = threadIdx.x; i < C::BLOB; i += WARPS * 32) smem[i] = __ldg(a.wfrag + i);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* stage = smem + C::BLOB + warp * (32 * C::STRIDE + 32);
    double* ystage = stage + 32 * C::STRIDE;
#pragma unroll
    for (int k = C::NIN; k < C::STRIDE; ++k) stage[lane * C::STRIDE + k] = 0.0;

    const i64 warps_total = (i64)gridDim.x * WARPS;
    const i64 gw = (i64)blockIdx.x * WARPS + warp;
    const i64 G = (a.N + 31) >> 5;
    const i64 g0 = gw * G / warps_total, g1 = (gw + 1) * G / warps_total;
    if (g0 >= g1) return;
    int c[D];
    const bool all_mode = (a.idx == nullptr);
    if (all_mode) {
        i64 i0 = g0 * 32 + lane;
        if (i0 < a.N) lex_unrank<D>(a.n, (u64)(a.rank_begin + i0), c);
        else {
#pragma unroll
            for (int t = 0; t < D; ++t) c[t] = t;
        }
    }
    const double* xoff = smem + C::OFF_XOFF;
    const double* gain = smem + C::OFF_GAIN;
#pragma unroll 1
    for (i64 g = g0; g < g1; ++g) {
        const i64 i = g * 32 + lane;
        const bool valid = i < a.N;
        if (!all_mode) load_list_indices<D>(a.idx, i, valid, c);
        double dotq, max_elem;
        {
            double xs[D], Xs[T], Qs[T];
            gather_point<D>(a, c, xs, Xs);
            int k = 0;
            double mx = 0.0;
#pragma unroll
            for (int p = 0; p < D; ++p)
#pragma unroll
                for (int q = p; q < D; ++q) {
                    double v = __ldg(a.Q + tri_index(a.n, c[p], c[q]));
                    Qs[k++] = v;
                    mx = fmax(mx, fabs(v));
                }
            // max_elem = len * |max|, 1 if 0; Q~ = Q / max_elem   (cut_select_qp.py:536-538)
            max_elem = (double)D * mx;
            if (max_elem == 0.0) max_elem = 1.0;
            const bool tiny = max_elem < 1e-280;   // subnormal-range coefficients: take the IEEE division
            const double rme = fast_rcp(tiny ? 1.0 : max_elem);
            double s = 0.0;  // sum(map(mul, Q_slice, X_slice)) left to right, no FMA (cut_select_qp.py:575)
#pragma unroll
            for (int q = 0; q < T; ++q) {
                Qs[q] = tiny ? __ddiv_rn(Qs[q], max_elem) : div_by(Qs[q], max_elem, rme);
                s = __dadd_rn(s, __dmul_rn(Qs[q], Xs[q]));
            }
            dotq = s;
            // mapminmax_apply (neural_net_3D.m:69-73): (in - xoffset) * gain + ymin
            double* row = stage + lane * C::STRIDE;
#pragma unroll
            for (int q = 0; q < D; ++q)
                row[q] = __dadd_rn(__dmul_rn(__dsub_rn(xs[q], xoff[q]), gain[q]), -1.0);
#pragma unroll
            for (int q = 0; q < T; ++q)
                row[D + q] = __dadd_rn(__dmul_rn(__dsub_rn(Qs[q], xoff[D + q]), gain[D + q]), -1.0);
        }
        __syncwarp();
#pragma unroll 1
        for (int pass = 0; pass < 4; ++pass) {
            double y = mlp8<D>(smem, stage + 8 * pass * C::STRIDE, lane);
            if ((lane & 3) == 0) ystage[8 * pass + (lane >> 2)] = y;
        }
        __syncwarp();
        if (valid) {
            double nn = ystage[lane];
            // obj = -sum * max_elem + NN * max_elem   (cut_select_qp.py:575, 582)
            double obj = __dadd_rn(__dmul_rn(-dotq, max_elem), __dmul_rn(nn, max_elem));
            a.obj[a.pos ? __ldg(a.pos + i) : i] = obj;
        }
        __syncwarp();
        if (all_mode && g + 1 < g1) {
            if (!(valid && lex_advance<D>(a.n, c, 32))) {
#pragma unroll
                for (int t = 0; t < D; ++t) c[t] = t;
            }
        }
    }
}

// Plain batched NN forward pass (sdpcs_nn_eval): rows of raw inputs in global memory, same MLP core.
template <int D, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_nn_eval(const double* wfrag, const double* in, i64 m, double* out)
{
    using C = NetCfg<D>;
    extern __shared__ __align__(16) double smem[];
    for (int i = threadIdx.x; i < C::BLOB; i += WARPS * 32) smem[i] = __ldg(wfrag + i);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* stage = smem + C::BLOB + warp * (32 * C::STRIDE + 32);
    double* ystage = stage + 32 * C::STRIDE;
#pragma unroll
    for (int k = C::NIN; k < C::STRIDE; ++k) stage[lane * C::STRIDE + k] = 0.0;
    const double* xoff = smem + C::OFF_XOFF;
    const double* gain = smem + C::OFF_GAIN;
    const i64 G = (m + 31) >> 5;
    for (i64 g = (i64)blockIdx.x * WARPS + warp; g < G; g += (i64)gridDim.x * WARPS) {
        const i64 i = g * 32 + lane;
        const bool valid = i < m;
        double* row = stage + lane * C::STRIDE;
#pragma unroll
        for (int q = 0; q < C::NIN; ++q) {
            double v = valid ? in[i * C::NIN + q] : 0.0;
            row[q] = __dadd_rn(__dmul_rn(__dsub_rn(v, xoff[q]), gain[q]), -1.0);
        }
        __syncwarp();
#pragma unroll 1
        for (int pass = 0; pass < 4; ++pass) {
            double y = mlp8<D>(smem, stage + 8 * pass * C::STRIDE, lane);
            if ((lane & 3) == 0) ystage[8 * pass + (lane >> 2)] = y;
        }
        __syncwarp();
        if (valid) out[i] = ystage[lane];
        __syncwarp();
    }
}

}  // namespace sdpcs
