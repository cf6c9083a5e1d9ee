// libsdpcutsel: C ABI (include/sdpcutsel.h) over the sm_100a kernels. Host logic only: buffers, launches,
// pass sequencing. No CPU compute path exists -- every score, selection and cut is produced by a kernel.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sdpcutsel.h"
#include "aux_kernels.cuh"
#include "cover_kernels.cuh"
#include "dense_kernels.cuh"
#include "score_kernels.cuh"
#include "mlp_i8_kernels.cuh"
#include "select_kernels.cuh"
#include "exchange_kernels.cuh"
#include "sdp_kernels.cuh"

using namespace sdpcs;

static std::string g_create_error;

struct sdpcs_ctx {
    int device = 0;
    int sms = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;
    sdpcs_params params;
    // instance
    int n = 0;
    double* d_Q = nullptr;
    double* d_vars = nullptr;      // [X | x]
    double* h_vars = nullptr;      // pinned staging
    // weights (fragment-ordered), index = rho
    double* d_wfrag[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // weights as int8 digit images + FP64 parameter block for the tcgen05 MLP (mlp_i8_kernels.cuh)
    uint8_t* d_wi8[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint8_t* d_wi8s[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // 4-digit images of the screening engine
    bool i8_ok[6] = {false, false, false, false, false, false};   // pre-activations of the net stay inside the range of the tcgen05 engines
    uint8_t* d_tiles = nullptr;    // layer-0 digit images of one chunk of candidates
    i64 tiles_cap = 0;             // in bytes
    int* d_status = nullptr;       // device status word of the tcgen05 pipeline
    int* h_status = nullptr;       // pinned
    bool i8_used = false;
    bool fuse_feas = false;          // this scoring call: lam_min comes out of k_prep_i8<.., FEAS> (set by score_device)
    // cover
    int rho = 0, mode = 0;         // mode 0 none, 1 all-subsets, 2 list
    i64 N = 0, base = 0;           // base = agg_idx of local candidate 0 (rank_begin / agg_offset)
    uint8_t* d_idx[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    i64* d_pos[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    i64 Nd[6] = {0, 0, 0, 0, 0, 0};
    i64 lo[6] = {0, 0, 0, 0, 0, 0};    // list cover restricted to a shard (sdpcs_cover_restrict): first row of the view per size class
    bool restricted = false;
    // scores
    double *d_lam = nullptr, *d_obj = nullptr;
    i64 score_cap = 0;
    int have = 0;
    bool obj_exact = false;        // d_obj holds the exact SDP measure (strat 3) instead of the NN measure
    // selection scratch (keys are recomputed from lam / obj in every pass; d_key1 holds the triangle keys only)
    u64* d_key1 = nullptr;
    i64 key_cap = 0;
    SelState* d_state = nullptr;
    SelState* h_state = nullptr;   // pinned copy of the state after a selection
    u64 *d_c_k1 = nullptr, *d_c_k2 = nullptr, *d_s_k1 = nullptr, *d_s_k2 = nullptr;
    i64 *d_c_idx = nullptr, *d_s_idx = nullptr, *d_s_perm = nullptr;
    double *d_o_score = nullptr, *d_o_lam = nullptr, *d_o_obj = nullptr;
    i64 out_cap = 0, sel_cap = 0;  // allocated entries / entries the last selection could collect (k + band_cap)
    void* h_out = nullptr;         // pinned download staging: [idx | score | lam | obj], h_out_stride bytes each
    size_t h_out_bytes = 0, h_out_stride = 0;
    i64 h_out_rows = 0, h_out_k = 0;   // rows downloaded by the last selection (winners + guard band), winners among them
    int last_mode = 0;
    i64 last_counts[3] = {0, 0, 0};
    i64 last_guard[4] = {0, 0, 0, 0};  // band entries, band open, n_unc_lam, n_unc_obj of the last selection
    double last_max_pos_nonviol = -INFINITY;   // largest obj among (obj > thres_min_opt and not violated), last top-k pass
    // triangles
    uint8_t* d_adj = nullptr;
    bool have_adj = false;
    bool vars_resident = false;
    unsigned long long* d_tri_counters = nullptr;
    // generic scratch
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    // timing
    cudaEvent_t ev[8];
    bool ev_score = false, ev_select = false, ev_h2d = false, ev_nn = false;
    sdpcs_timings tm;

    int fail(int code, const std::string& m) { err = m; return code; }
};

#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return ctx->fail(SDPCS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
    } while (0)

template <typename T>
static int ensure_dev(sdpcs_ctx* ctx, T*& p, i64& cap, i64 want)
{
    if (want <= cap && p) return SDPCS_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    if ((size_t)want * sizeof(T) > free_b)
        return ctx->fail(SDPCS_ERR_NOMEM, "shard needs " + std::to_string((size_t)want * sizeof(T)) +
                                              " bytes, only " + std::to_string(free_b) + " free on device");
    CU(cudaMalloc(&p, std::max<i64>(want, 1) * sizeof(T)));
    cap = want;
    return SDPCS_OK;
}

static int ensure_scratch(sdpcs_ctx* ctx, size_t bytes)
{
    if (bytes <= ctx->scratch_bytes) return SDPCS_OK;
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
    CU(cudaMalloc(&ctx->d_scratch, bytes));
    ctx->scratch_bytes = bytes;
    return SDPCS_OK;
}

static int ensure_hout(sdpcs_ctx* ctx, size_t bytes)
{
    if (bytes <= ctx->h_out_bytes) return SDPCS_OK;
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    ctx->h_out = nullptr; ctx->h_out_bytes = 0;
    CU(cudaMallocHost(&ctx->h_out, bytes));
    ctx->h_out_bytes = bytes;
    return SDPCS_OK;
}

// ---------------------------------------------------------------------------------------------------
// NN weights: blob -> DMMA fragment order (see mlp8 in score_kernels.cuh)
// ---------------------------------------------------------------------------------------------------
template <int D>
static bool pack_fragments(const double* blob, i64 len, std::vector<double>& out, std::string& err)
{
    using C = NetCfg<D>;
    if (len < 3) { err = "weight blob too short"; return false; }
    const int n_in = (int)blob[0], L = (int)blob[1], h = (int)blob[2];
    const i64 expect = 3 + 2 * n_in + (i64)h * n_in + h + (i64)(L - 2) * (h * h + h) + h + 1 + 2;
    if (n_in != C::NIN || h != C::H || L != C::NHID + 1 || len != expect) {
        err = "weight blob does not describe NN_" + std::to_string(D) + "D (n_in=" + std::to_string(n_in) +
              " L=" + std::to_string(L) + " h=" + std::to_string(h) + " len=" + std::to_string(len) + ")";
        return false;
    }
    const double* xo = blob + 3;
    const double* xg = xo + n_in;
    const double* p = xg + n_in;
    std::vector<const double*> W(L), B(L);
    for (int l = 0; l < L; ++l) {
        int rows = (l == L - 1) ? 1 : h, cols = (l == 0) ? n_in : h;
        W[l] = p; p += (i64)rows * cols;
        B[l] = p; p += rows;
    }
    const double y_gain = p[0], y_xoff = p[1];
    const double s = SDPCS_TANSIG_SCALE;
    out.assign(C::BLOB, 0.0);
    for (int ks = 0; ks < C::KS0; ++ks)
        for (int nt = 0; nt < C::NT; ++nt)
            for (int lane = 0; lane < 32; ++lane) {
                int g = lane >> 2, t = lane & 3, nrn = 8 * nt + g, k = 4 * ks + t;
                out[C::OFF_W0 + ks * C::NT * 32 + C::frag(nt, lane)] = (nrn < h && k < n_in) ? s * W[0][nrn * n_in + k] : 0.0;
            }
    for (int l = 1; l < C::NHID; ++l)
        for (int kt = 0; kt < C::NT; ++kt)
            for (int hh = 0; hh < 2; ++hh)
                for (int nt = 0; nt < C::NT; ++nt)
                    for (int lane = 0; lane < 32; ++lane) {
                        int g = lane >> 2, t = lane & 3, nrn = 8 * nt + g, k = 8 * kt + 2 * t + hh, ks = 2 * kt + hh;
                        out[C::OFF_WH + (l - 1) * C::LAYER + ks * C::NT * 32 + C::frag(nt, lane)] =
                            (nrn < h && k < h) ? s * W[l][nrn * h + k] : 0.0;
                    }
    for (int l = 0; l < C::NHID; ++l)
        for (int j = 0; j < h; ++j) out[C::OFF_BIAS + l * C::HP + j] = s * B[l][j];
    for (int j = 0; j < h; ++j) out[C::OFF_WOUT + j] = W[L - 1][j];
    for (int j = 0; j < n_in; ++j) { out[C::OFF_XOFF + j] = xo[j]; out[C::OFF_GAIN + j] = xg[j]; }
    out[C::OFF_MISC + 0] = B[L - 1][0];
    out[C::OFF_MISC + 1] = y_gain;
    out[C::OFF_MISC + 2] = y_xoff;
    for (int j = 0; j < 256; ++j) out[C::OFF_TAB + j] = (double)exp2l((long double)j / 256.0L);
    return true;
}

// ---------------------------------------------------------------------------------------------------
// NN weights: blob -> int8 digit images (UMMA canonical K-major layout, [k chunk][slice][neuron][16 B]) + FP64
// parameter block for k_mlp_i8.
// Row j of a layer is scaled by 2^e >= max|W[j,:]|, rounded to 54 fractional bits and written as 7 balanced
// base-256 digits (slice 0 = most significant).  oracle/nn_i8_model.py states the same arithmetic.
// ---------------------------------------------------------------------------------------------------
// Returns false if a pre-activation could leave the range the kernel's tansig handles without a clamp (|z| < I8_Z_MAX):
// the FP64 DMMA engine then serves this net.
template <int D, int NS>
static bool pack_i8(const double* blob, std::vector<uint8_t>& out)
{
    using C = NetCfg<D>;
    using L = I8Smem<C::NHID, NS>;
    using G = I8Dig<NS>;
    const int n_in = C::NIN, h = C::H, NL = C::NHID + 1;
    const double* xo = blob + 3;
    const double* p = xo + 2 * n_in;
    std::vector<const double*> W(NL), B(NL);
    for (int l = 0; l < NL; ++l) {
        int rows = (l == NL - 1) ? 1 : h, cols = (l == 0) ? n_in : h;
        W[l] = p; p += (i64)rows * cols;
        B[l] = p; p += rows;
    }
    const double y_gain = p[0], y_xoff = p[1];
    out.assign(L::GLOBAL_BYTES, 0);
    double* par = reinterpret_cast<double*>(out.data() + L::W_TOTAL);
    unsigned long long bias = 0;                         // 0x80 in every digit: balanced digits
    for (int b = 0; b < NS; ++b) bias |= 0x80ull << (8 * b);
    double zmax = 0.0;
    for (int l = 0; l < C::NHID; ++l) {
        const int cols = (l == 0) ? n_in : h, K = (l == 0) ? I8_K0 : 64;
        const int ea = (l == 0) ? G::KA + 2 : G::KA + 3;  // digits carry 8 * rint(a * 2^(KA-1)) resp. 8 * rint(a * 2^KA)
        uint8_t* img = out.data() + (l == 0 ? 0 : G::W0_BYTES + (l - 1) * G::WH_BYTES);
        for (int j = 0; j < I8_N; ++j) {
            double mx = 0.0, sum = 0.0;
            if (j < h) for (int k = 0; k < cols; ++k) { mx = std::max(mx, std::fabs(W[l][j * cols + k])); sum += std::fabs(W[l][j * cols + k]); }
            if (j < h) zmax = std::max(zmax, std::fabs(SDPCS_TANSIG_SCALE) * (sum * (l == 0 ? 2.0 : 1.0) + std::fabs(B[l][j])));
            int e = 0;
            if (mx > 0.0) std::frexp(mx, &e);           // mx = f * 2^e, f in [0.5, 1): 2^e > = mx
            for (int k = 0; k < K; ++k) {
                const double w = (j < h && k < cols) ? W[l][j * cols + k] : 0.0;
                const long long wint = std::llrint(std::ldexp(w, G::KW - e));
                const unsigned long long u = (unsigned long long)(wint + (long long)bias);
                for (int b = 0; b < NS; ++b) {
                    const int digit = (int)((u >> (8 * b)) & 0xFF) - 128;
                    img[(k / 16) * (NS * I8_N * 16) + (NS - 1 - b) * (I8_N * 16) + j * 16 + (k % 16)] = (uint8_t)(int8_t)digit;
                }
            }
            // z = -2 log2(e) * (W a + b);  W a = 2^(e - KW) * 2^-ea * 256^(NS-1) * (sum of the kept digit-pair diagonals)
            par[L::P_CS + 2 * (l * 64 + j)] = std::ldexp(1.0, e - G::KW - ea + 8 * (NS - 1)) * SDPCS_TANSIG_SCALE;
            par[L::P_CS + 2 * (l * 64 + j) + 1] = (j < h) ? SDPCS_TANSIG_SCALE * B[l][j] : 0.0;
        }
    }
    for (int j = 0; j < h; ++j) par[L::P_WOUT + j] = W[NL - 1][j];
    par[L::P_MISC + 0] = B[NL - 1][0];
    par[L::P_MISC + 1] = y_gain;
    par[L::P_MISC + 2] = y_xoff;
    for (int j = 0; j < 256; ++j) par[L::P_TAB + j] = (double)exp2l((long double)j / 256.0L);
    return zmax < I8_Z_MAX;
}

// ---------------------------------------------------------------------------------------------------
// basic API
// ---------------------------------------------------------------------------------------------------
extern "C" int sdpcs_default_params(sdpcs_params* p)
{
    if (!p) return SDPCS_ERR_INVALID;
    p->thres_min_opt = 0.0;
    p->thres_neg_eigval = -1e-15;
    p->big_m = 1000.0;
    p->thres_tri_viol = 1e-7;
    p->thres_tri_dense = 2;
    p->jacobi_sweeps = 0;
    p->nn_engine = SDPCS_NN_TCGEN05;
    p->nn_fused_prep = 0;
    p->guard_lam = 1e-12;
    p->guard_obj = 1e-9;
    p->band_cap = 65536;
    p->sdp_mu_final = 1e-12;
    return SDPCS_OK;
}

extern "C" const char* sdpcs_last_error(const sdpcs_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int sdpcs_create(sdpcs_ctx** out, int device)
{
    if (!out) return SDPCS_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
        return SDPCS_ERR_CUDA;
    }
    if (device < 0 || device >= count) { g_create_error = "bad device ordinal"; return SDPCS_ERR_INVALID; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return SDPCS_ERR_CUDA; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return SDPCS_ERR_CUDA; }
    if (prop.major != 10) {
        g_create_error = "libsdpcutsel is built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor);
        return SDPCS_ERR_CUDA;
    }
    sdpcs_ctx* ctx = new sdpcs_ctx();
    ctx->device = device;
    ctx->sms = prop.multiProcessorCount;
    sdpcs_default_params(&ctx->params);
    memset(&ctx->tm, 0, sizeof(ctx->tm));
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&ctx->d_state, sizeof(SelState)) != cudaSuccess || cudaMallocHost(&ctx->h_state, sizeof(SelState)) != cudaSuccess ||
        cudaMalloc(&ctx->d_tri_counters, 2 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&ctx->d_status, sizeof(int)) != cudaSuccess || cudaMallocHost(&ctx->h_status, sizeof(int)) != cudaSuccess ||
        cudaMemset(ctx->d_status, 0, sizeof(int)) != cudaSuccess) {
        g_create_error = "context allocation failed";
        delete ctx;
        return SDPCS_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    *out = ctx;
    return SDPCS_OK;
}

extern "C" int sdpcs_destroy(sdpcs_ctx* ctx)
{
    if (!ctx) return SDPCS_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    void* ptrs[] = {ctx->d_Q, ctx->d_vars, ctx->d_lam, ctx->d_obj, ctx->d_key1, ctx->d_state, ctx->d_c_k1,
                    ctx->d_c_k2, ctx->d_s_k1, ctx->d_s_k2, ctx->d_c_idx, ctx->d_s_idx, ctx->d_s_perm, ctx->d_o_score,
                    ctx->d_o_lam, ctx->d_o_obj, ctx->d_adj, ctx->d_tri_counters, ctx->d_scratch, ctx->d_tiles, ctx->d_status};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int d = 0; d < 6; ++d) {
        if (ctx->d_wfrag[d]) cudaFree(ctx->d_wfrag[d]);
        if (ctx->d_wi8[d]) cudaFree(ctx->d_wi8[d]);
        if (ctx->d_wi8s[d]) cudaFree(ctx->d_wi8s[d]);
        if (ctx->d_idx[d]) cudaFree(ctx->d_idx[d]);
        if (ctx->d_pos[d]) cudaFree(ctx->d_pos[d]);
    }
    if (ctx->h_vars) cudaFreeHost(ctx->h_vars);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->h_state) cudaFreeHost(ctx->h_state);
    for (auto& ev : ctx->ev) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return SDPCS_OK;
}

extern "C" int sdpcs_set_stream(sdpcs_ctx* ctx, void* s)
{
    if (!ctx) return SDPCS_ERR_INVALID;
    cudaStream_t next = s ? (cudaStream_t)s : ctx->own_stream;
    if (next != ctx->stream) {
        CU(cudaSetDevice(ctx->device));
        CU(cudaStreamSynchronize(ctx->stream));      // nothing of this context is left running on the stream it leaves
    }
    ctx->stream = next;
    return SDPCS_OK;
}

extern "C" int sdpcs_set_params(sdpcs_ctx* ctx, const sdpcs_params* p)
{
    if (!ctx || !p) return SDPCS_ERR_INVALID;
    ctx->params = *p;
    return SDPCS_OK;
}

extern "C" int sdpcs_get_timings(const sdpcs_ctx* cctx, sdpcs_timings* t)
{
    sdpcs_ctx* ctx = const_cast<sdpcs_ctx*>(cctx);
    if (!ctx || !t) return SDPCS_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    float ms;
    if (ctx->ev_h2d) { CU(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5])); ctx->tm.h2d_ms = ms; }
    if (ctx->ev_score) { CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); ctx->tm.score_ms = ms; }
    ctx->tm.nn_ms = 0.0;
    if (ctx->ev_score && ctx->ev_nn) { CU(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7])); ctx->tm.nn_ms = ms; }
    if (ctx->ev_select) { CU(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3])); ctx->tm.select_ms = ms; }
    *t = ctx->tm;
    return SDPCS_OK;
}

extern "C" int sdpcs_set_weights(sdpcs_ctx* ctx, int rho, const double* blob, int64_t len)
{
    if (!ctx || !blob || rho < 2 || rho > 5) return ctx ? ctx->fail(SDPCS_ERR_INVALID, "rho must be 2..5") : SDPCS_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    std::vector<double> frag;
    std::string err;
    bool ok = false;
    switch (rho) {
    case 2: ok = pack_fragments<2>(blob, len, frag, err); break;
    case 3: ok = pack_fragments<3>(blob, len, frag, err); break;
    case 4: ok = pack_fragments<4>(blob, len, frag, err); break;
    case 5: ok = pack_fragments<5>(blob, len, frag, err); break;
    }
    if (!ok) return ctx->fail(SDPCS_ERR_INVALID, err);
    if (ctx->d_wfrag[rho]) { cudaFree(ctx->d_wfrag[rho]); ctx->d_wfrag[rho] = nullptr; }
    CU(cudaMalloc(&ctx->d_wfrag[rho], frag.size() * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->d_wfrag[rho], frag.data(), frag.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->i8_ok[rho] = true;
    for (int screen = 0; screen < 2; ++screen) {
        std::vector<uint8_t> img;
        bool in_range = false;
        switch (rho * 2 + screen) {
        case 4: in_range = pack_i8<2, I8_NS>(blob, img); break;
        case 5: in_range = pack_i8<2, I8_NS_SCREEN>(blob, img); break;
        case 6: in_range = pack_i8<3, I8_NS>(blob, img); break;
        case 7: in_range = pack_i8<3, I8_NS_SCREEN>(blob, img); break;
        case 8: in_range = pack_i8<4, I8_NS>(blob, img); break;
        case 9: in_range = pack_i8<4, I8_NS_SCREEN>(blob, img); break;
        case 10: in_range = pack_i8<5, I8_NS>(blob, img); break;
        default: in_range = pack_i8<5, I8_NS_SCREEN>(blob, img); break;
        }
        ctx->i8_ok[rho] = ctx->i8_ok[rho] && in_range;   // otherwise the FP64 DMMA engine serves this net
        uint8_t*& dst = screen ? ctx->d_wi8s[rho] : ctx->d_wi8[rho];
        if (dst) { cudaFree(dst); dst = nullptr; }
        CU(cudaMalloc(&dst, img.size()));
        CU(cudaMemcpy(dst, img.data(), img.size(), cudaMemcpyHostToDevice));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return SDPCS_OK;
}

extern "C" int sdpcs_set_instance(sdpcs_ctx* ctx, int n, const double* Q_arr)
{
    if (!ctx || !Q_arr || n < 2 || n > SDPCS_MAX_N) return ctx ? ctx->fail(SDPCS_ERR_INVALID, "need 2 <= n <= 250") : SDPCS_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    const size_t nl = (size_t)n * (n + 1) / 2;
    if (ctx->d_Q) cudaFree(ctx->d_Q);
    if (ctx->d_vars) cudaFree(ctx->d_vars);
    if (ctx->h_vars) cudaFreeHost(ctx->h_vars);
    ctx->d_Q = ctx->d_vars = ctx->h_vars = nullptr;
    CU(cudaMalloc(&ctx->d_Q, nl * sizeof(double)));
    CU(cudaMalloc(&ctx->d_vars, (nl + n) * sizeof(double)));
    CU(cudaMallocHost(&ctx->h_vars, (nl + n) * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->d_Q, Q_arr, nl * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->n = n;
    ctx->mode = 0; ctx->N = 0; ctx->have = 0;
    ctx->have_adj = false;
    ctx->vars_resident = false;
    return SDPCS_OK;
}

extern "C" int sdpcs_binom(int n, int k, int64_t* out)
{
    if (!out || n < 0 || n > SDPCS_MAX_N || k < 0 || k > SDPCS_MAX_RHO) return SDPCS_ERR_INVALID;
    *out = (int64_t)binom_small(n, k);
    return SDPCS_OK;
}

extern "C" int sdpcs_unrank(int n, int rho, const int64_t* ranks, int64_t m, int32_t* out_idx)
{
    if (!ranks || !out_idx || n < rho || n > SDPCS_MAX_N || rho < 1 || rho > SDPCS_MAX_RHO || m < 0) return SDPCS_ERR_INVALID;
    const i64 total = (i64)binom_small(n, rho);
    for (i64 i = 0; i < m; ++i) {
        if (ranks[i] < 0 || ranks[i] >= total) return SDPCS_ERR_INVALID;
        int c[5];
        switch (rho) {
        case 1: out_idx[i] = (int32_t)ranks[i]; continue;
        case 2: { int cc[2]; lex_unrank<2>(n, (u64)ranks[i], cc); c[0] = cc[0]; c[1] = cc[1]; } break;
        case 3: { int cc[3]; lex_unrank<3>(n, (u64)ranks[i], cc); for (int t = 0; t < 3; ++t) c[t] = cc[t]; } break;
        case 4: { int cc[4]; lex_unrank<4>(n, (u64)ranks[i], cc); for (int t = 0; t < 4; ++t) c[t] = cc[t]; } break;
        default: { int cc[5]; lex_unrank<5>(n, (u64)ranks[i], cc); for (int t = 0; t < 5; ++t) c[t] = cc[t]; } break;
        }
        for (int t = 0; t < rho; ++t) out_idx[i * rho + t] = c[t];
    }
    return SDPCS_OK;
}

static int alloc_scores(sdpcs_ctx* ctx, i64 N)
{
    i64 cap2 = ctx->score_cap;
    int rc = ensure_dev(ctx, ctx->d_lam, ctx->score_cap, N);
    if (rc) return rc;
    rc = ensure_dev(ctx, ctx->d_obj, cap2, N);
    if (rc) return rc;
    return SDPCS_OK;
}

extern "C" int sdpcs_set_cover_all(sdpcs_ctx* ctx, int rho, int64_t rank_begin, int64_t rank_end)
{
    if (!ctx) return SDPCS_ERR_INVALID;
    if (!ctx->n) return ctx->fail(SDPCS_ERR_STATE, "set_instance first");
    if (rho < 2 || rho > 5 || rho > ctx->n) return ctx->fail(SDPCS_ERR_INVALID, "rho must be 2..5");
    CU(cudaSetDevice(ctx->device));
    const i64 total = (i64)binom_small(ctx->n, rho);
    if (rank_end < 0) rank_end = total;
    if (rank_begin < 0 || rank_begin > rank_end || rank_end > total) return ctx->fail(SDPCS_ERR_INVALID, "bad rank range");
    ctx->mode = 0; ctx->have = 0;
    // both score arrays must fit
    if (ctx->d_lam && ctx->score_cap < rank_end - rank_begin) { cudaFree(ctx->d_lam); cudaFree(ctx->d_obj); ctx->d_lam = ctx->d_obj = nullptr; ctx->score_cap = 0; }
    int rc = alloc_scores(ctx, rank_end - rank_begin);
    if (rc) return rc;
    ctx->rho = rho; ctx->mode = 1; ctx->base = rank_begin; ctx->N = rank_end - rank_begin;
    return SDPCS_OK;
}

extern "C" int sdpcs_set_cover_list(sdpcs_ctx* ctx, int rho, const int16_t* idx, int64_t N, int64_t agg_offset)
{
    if (!ctx) return SDPCS_ERR_INVALID;
    if (!ctx->n) return ctx->fail(SDPCS_ERR_STATE, "set_instance first");
    if (rho < 2 || rho > 5 || N < 0 || (N > 0 && !idx)) return ctx->fail(SDPCS_ERR_INVALID, "bad cover list");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->mode = 0; ctx->have = 0;
    std::vector<uint8_t> packed[6];
    std::vector<i64> pos[6];
    for (i64 i = 0; i < N; ++i) {
        const int16_t* r = idx + i * rho;
        int d = 0;
        while (d < rho && r[d] >= 0) ++d;
        for (int t = d; t < rho; ++t) if (r[t] >= 0) return ctx->fail(SDPCS_ERR_INVALID, "cover row " + std::to_string(i) + ": -1 padding must trail");
        if (d < 2) return ctx->fail(SDPCS_ERR_INVALID, "cover row " + std::to_string(i) + " has fewer than 2 indices");
        for (int t = 0; t < d; ++t) {
            if (r[t] >= ctx->n || (t && r[t] <= r[t - 1])) return ctx->fail(SDPCS_ERR_INVALID, "cover row " + std::to_string(i) + " not strictly ascending in [0,n)");
            packed[d].push_back((uint8_t)r[t]);
        }
        pos[d].push_back(i);
    }
    for (int d = 2; d <= 5; ++d) {
        ctx->lo[d] = 0; ctx->restricted = false;
        if (ctx->d_idx[d]) { cudaFree(ctx->d_idx[d]); ctx->d_idx[d] = nullptr; }
        if (ctx->d_pos[d]) { cudaFree(ctx->d_pos[d]); ctx->d_pos[d] = nullptr; }
        ctx->Nd[d] = (i64)pos[d].size();
        if (!ctx->Nd[d]) continue;
        CU(cudaMalloc(&ctx->d_idx[d], packed[d].size()));
        CU(cudaMalloc(&ctx->d_pos[d], pos[d].size() * sizeof(i64)));
        CU(cudaMemcpy(ctx->d_idx[d], packed[d].data(), packed[d].size(), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(ctx->d_pos[d], pos[d].data(), pos[d].size() * sizeof(i64), cudaMemcpyHostToDevice));
    }
    if (ctx->d_lam && ctx->score_cap < N) { cudaFree(ctx->d_lam); cudaFree(ctx->d_obj); ctx->d_lam = ctx->d_obj = nullptr; ctx->score_cap = 0; }
    int rc = alloc_scores(ctx, N);
    if (rc) return rc;
    ctx->rho = rho; ctx->mode = 2; ctx->base = agg_offset; ctx->N = N;
    return SDPCS_OK;
}

// P^E_rho built on the device (cover_kernels.cuh); replaces the nested loops of cut_select_qp.py:401-522.
extern "C" int sdpcs_set_cover_pattern(sdpcs_ctx* ctx, int rho, const uint8_t* adj, int64_t agg_offset, int64_t* out_N)
{
    if (!ctx) return SDPCS_ERR_INVALID;
    if (!ctx->n) return ctx->fail(SDPCS_ERR_STATE, "set_instance first");
    if (rho < 2 || rho > 5 || !adj) return ctx->fail(SDPCS_ERR_INVALID, "bad pattern cover arguments");
    const int n = ctx->n;
    if (n > 256) return ctx->fail(SDPCS_ERR_INVALID, "pattern cover needs n <= 256");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->mode = 0; ctx->have = 0; ctx->N = 0;
    std::vector<Mask256> masks(n);
    std::vector<int> edges;
    for (int i = 0; i < n; ++i) {
        for (int w = 0; w < 4; ++w) masks[i].w[w] = 0;
        for (int j = 0; j < n; ++j)
            if (j != i && (adj[(size_t)i * n + j] || adj[(size_t)j * n + i])) masks[i].w[j >> 6] |= 1ull << (j & 63);
    }
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            if (masks[i].w[j >> 6] >> (j & 63) & 1) { edges.push_back(i); edges.push_back(j); }
    const i64 E = (i64)edges.size() / 2;
    for (int d = 2; d <= 5; ++d) {
        ctx->lo[d] = 0; ctx->restricted = false;
        if (ctx->d_idx[d]) { cudaFree(ctx->d_idx[d]); ctx->d_idx[d] = nullptr; }
        if (ctx->d_pos[d]) { cudaFree(ctx->d_pos[d]); ctx->d_pos[d] = nullptr; }
        ctx->Nd[d] = 0;
    }
    i64 N = 0;
    if (E > 0) {
        // scratch: masks | edges | counts | pos_off | cls_off
        const size_t o_edges = sizeof(Mask256) * n, o_counts = o_edges + sizeof(int) * 2 * E, o_pos = (o_counts + sizeof(int) * 4 * E + 7) & ~(size_t)7,
                     o_cls = o_pos + sizeof(i64) * E, total = o_cls + sizeof(i64) * 4 * E;
        int rc = ensure_scratch(ctx, total);
        if (rc) return rc;
        uint8_t* sc = static_cast<uint8_t*>(ctx->d_scratch);
        CU(cudaMemcpyAsync(sc, masks.data(), sizeof(Mask256) * n, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(sc + o_edges, edges.data(), sizeof(int) * 2 * E, cudaMemcpyHostToDevice, ctx->stream));
        CoverArgs a;
        memset(&a, 0, sizeof(a));
        a.n = n; a.rho = rho; a.adj = reinterpret_cast<const Mask256*>(sc); a.edges = reinterpret_cast<const int*>(sc + o_edges); a.E = E;
        a.counts = reinterpret_cast<int*>(sc + o_counts);
        const unsigned grid = (unsigned)((E + 127) / 128);
        k_cover_pattern<false><<<grid, 128, 0, ctx->stream>>>(a);
        CU(cudaGetLastError());
        std::vector<int> counts(4 * E);
        CU(cudaMemcpyAsync(counts.data(), a.counts, sizeof(int) * 4 * E, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        std::vector<i64> pos_off(E), cls_off(4 * E);
        i64 tot[4] = {0, 0, 0, 0};
        for (i64 e = 0; e < E; ++e) {
            pos_off[e] = N;
            for (int d = 0; d < 4; ++d) { cls_off[4 * e + d] = tot[d]; tot[d] += counts[4 * e + d]; N += counts[4 * e + d]; }
        }
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        size_t need = 0;
        for (int d = 2; d <= 5; ++d) need += (size_t)tot[d - 2] * (d + sizeof(i64));
        if (need + (size_t)N * 32 > free_b)
            return ctx->fail(SDPCS_ERR_NOMEM, "pattern cover of " + std::to_string(N) + " candidates does not fit on the device");
        for (int d = 2; d <= 5; ++d) {
            ctx->Nd[d] = tot[d - 2];
            if (!ctx->Nd[d]) continue;
            CU(cudaMalloc(&ctx->d_idx[d], (size_t)ctx->Nd[d] * d));
            CU(cudaMalloc(&ctx->d_pos[d], (size_t)ctx->Nd[d] * sizeof(i64)));
            a.idx[d] = ctx->d_idx[d]; a.pos[d] = ctx->d_pos[d];
        }
        CU(cudaMemcpyAsync(sc + o_pos, pos_off.data(), sizeof(i64) * E, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(sc + o_cls, cls_off.data(), sizeof(i64) * 4 * E, cudaMemcpyHostToDevice, ctx->stream));
        a.pos_off = reinterpret_cast<const i64*>(sc + o_pos); a.cls_off = reinterpret_cast<const i64*>(sc + o_cls);
        k_cover_pattern<true><<<grid, 128, 0, ctx->stream>>>(a);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(ctx->stream));   // the host vectors above go out of scope
    }
    if (ctx->d_lam && ctx->score_cap < N) { cudaFree(ctx->d_lam); cudaFree(ctx->d_obj); ctx->d_lam = ctx->d_obj = nullptr; ctx->score_cap = 0; }
    int rc = alloc_scores(ctx, N);
    if (rc) return rc;
    ctx->rho = rho; ctx->mode = 2; ctx->base = agg_offset; ctx->N = N;
    if (out_N) *out_N = N;
    return SDPCS_OK;
}

// The list cover as N x rho int16 rows (padded with -1), in candidate order: what the reference keeps as
// agg_list[i][0] (cut_select_qp.py:525-540).
extern "C" int sdpcs_get_cover_rows(sdpcs_ctx* ctx, int16_t* out_idx, int64_t cap_rows)
{
    if (!ctx) return SDPCS_ERR_INVALID;
    if (ctx->mode != 2) return ctx->fail(SDPCS_ERR_STATE, "no list cover set");
    if (cap_rows < ctx->N || (ctx->N > 0 && !out_idx)) return ctx->fail(SDPCS_ERR_INVALID, "output buffer too small");
    if (!ctx->N) return SDPCS_OK;
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)ctx->N * ctx->rho * sizeof(int16_t);
    int rc = ensure_scratch(ctx, bytes);
    if (rc) return rc;
    int16_t* d_rows = static_cast<int16_t*>(ctx->d_scratch);
    for (int d = 2; d <= ctx->rho; ++d) {
        if (!ctx->Nd[d]) continue;
        const unsigned grid = (unsigned)std::min<i64>((ctx->Nd[d] + 255) / 256, (i64)ctx->sms * 8);
        k_cover_rows<<<grid, 256, 0, ctx->stream>>>(ctx->d_idx[d] + ctx->lo[d] * d, ctx->d_pos[d] + ctx->lo[d], ctx->Nd[d], d, ctx->rho, d_rows);
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(out_idx, d_rows, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SDPCS_OK;
}

// Shard of the current cover: keep the candidates [begin, end) (local indices of the cover as set).  All-subsets covers
// just move their rank range; list covers keep their device arrays and narrow a view per size class (positions are
// ascending inside a class).  agg_idx of the survivors is unchanged.
extern "C" int sdpcs_cover_restrict(sdpcs_ctx* ctx, int64_t begin, int64_t end)
{
    if (!ctx) return SDPCS_ERR_INVALID;
    if (!ctx->mode) return ctx->fail(SDPCS_ERR_STATE, "cover not set");
    if (begin < 0 || end < begin || end > ctx->N) return ctx->fail(SDPCS_ERR_INVALID, "bad shard range");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have = 0;
    if (ctx->mode == 1) {
        ctx->base += begin; ctx->N = end - begin;
        return SDPCS_OK;
    }
    if (ctx->restricted) return ctx->fail(SDPCS_ERR_STATE, "list cover already restricted: set the cover again first");
    int rc = ensure_scratch(ctx, 8 * sizeof(i64));
    if (rc) return rc;
    i64* d_b = static_cast<i64*>(ctx->d_scratch);
    for (int d = 2; d <= 5; ++d)
        if (ctx->Nd[d]) k_pos_bounds<<<1, 1, 0, ctx->stream>>>(ctx->d_pos[d], ctx->Nd[d], begin, end, d_b + 2 * (d - 2));
    i64 b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CU(cudaMemcpyAsync(b, d_b, sizeof(b), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int d = 2; d <= 5; ++d) {
        if (!ctx->Nd[d]) continue;
        ctx->lo[d] = b[2 * (d - 2)];
        ctx->Nd[d] = b[2 * (d - 2) + 1] - b[2 * (d - 2)];
        if (ctx->Nd[d] && begin) {
            const unsigned grid = (unsigned)std::min<i64>((ctx->Nd[d] + 255) / 256, (i64)ctx->sms * 8);
            k_pos_shift<<<grid, 256, 0, ctx->stream>>>(ctx->d_pos[d] + ctx->lo[d], ctx->Nd[d], begin);
        }
    }
    CU(cudaGetLastError());
    ctx->restricted = true;
    ctx->base += begin; ctx->N = end - begin;
    return SDPCS_OK;
}

static int scan_exclusive(sdpcs_ctx* ctx, const int* d_in, i64 n, i64* d_out, i64* d_sums, i64* d_total)
{
    const i64 nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb > 0) k_scan_block<<<(unsigned)nb, 256, 0, ctx->stream>>>(d_in, n, d_out, d_sums);
    k_scan_sums<<<1, 1024, 0, ctx->stream>>>(d_sums, nb, d_total);
    if (nb > 0) k_scan_add<<<(unsigned)nb, 256, 0, ctx->stream>>>(d_out, n, d_sums);
    CU(cudaGetLastError());
    return SDPCS_OK;
}

// Cover algebra of the QCQP caller (cut_select_qcqp.py:319-333) on the device: keep the candidates of this context's list
// cover that also occur in `other`'s list cover (keep_members = 1: intersection) or that do not (0: difference), in
// their order; agg_idx is renumbered 0..N'-1 (+ agg_offset).
extern "C" int sdpcs_cover_filter(sdpcs_ctx* ctx, const sdpcs_ctx* other, int keep_members, int64_t* out_N)
{
    if (!ctx || !other) return SDPCS_ERR_INVALID;
    if (ctx->mode != 2 || other->mode != 2) return ctx->fail(SDPCS_ERR_STATE, "both contexts need a list cover (sdpcs_set_cover_pattern / _list)");
    if (ctx->restricted || other->restricted) return ctx->fail(SDPCS_ERR_STATE, "cover algebra needs unrestricted covers");
    if (ctx->device != other->device || ctx->n != other->n) return ctx->fail(SDPCS_ERR_INVALID, "covers of different instances / devices");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(other->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const i64 N = ctx->N;
    ctx->have = 0;
    if (N == 0) { if (out_N) *out_N = 0; return SDPCS_OK; }
    i64 maxNd = 0;
    for (int d = 2; d <= 5; ++d) maxNd = std::max(maxNd, ctx->Nd[d]);
    const i64 nb = (std::max(N, maxNd) + SCAN_TILE - 1) / SCAN_TILE + 1;
    // scratch: keep_all[N] | all_scan[N] | keep_cls[maxNd] | cls_scan[maxNd] | sums[nb] | totals[2]
    const size_t o_all = 0, o_ascan = (o_all + 4 * (size_t)N + 15) & ~(size_t)15, o_kcls = o_ascan + 8 * (size_t)N,
                 o_cscan = (o_kcls + 4 * (size_t)maxNd + 15) & ~(size_t)15, o_sums = o_cscan + 8 * (size_t)maxNd, o_tot = o_sums + 8 * (size_t)nb;
    int rc = ensure_scratch(ctx, o_tot + 16);
    if (rc) return rc;
    char* sc = static_cast<char*>(ctx->d_scratch);
    int* keep_all = reinterpret_cast<int*>(sc + o_all);
    i64* all_scan = reinterpret_cast<i64*>(sc + o_ascan);
    int* keep_cls = reinterpret_cast<int*>(sc + o_kcls);
    i64* cls_scan = reinterpret_cast<i64*>(sc + o_cscan);
    i64* sums = reinterpret_cast<i64*>(sc + o_sums);
    i64* totals = reinterpret_cast<i64*>(sc + o_tot);
    CU(cudaMemsetAsync(keep_all, 0, 4 * (size_t)N, ctx->stream));
    // pass 1: global keep flags (by position) from every size class
    for (int d = 2; d <= 5; ++d) {
        if (!ctx->Nd[d]) continue;
        const unsigned grid = (unsigned)std::min<i64>((ctx->Nd[d] + 255) / 256, (i64)ctx->sms * 8);
        k_cover_member<<<grid, 256, 0, ctx->stream>>>(ctx->d_idx[d], ctx->d_pos[d], ctx->Nd[d], d, other->d_idx[d], other->Nd[d], keep_members,
                                                      keep_cls, keep_all);
    }
    if ((rc = scan_exclusive(ctx, keep_all, N, all_scan, sums, totals))) return rc;
    i64 newN = 0;
    CU(cudaMemcpyAsync(&newN, totals, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    // pass 2: per class flags again (the scratch flag array is shared), class scan, compaction into fresh arrays
    for (int d = 2; d <= 5; ++d) {
        if (!ctx->Nd[d]) continue;
        const unsigned grid = (unsigned)std::min<i64>((ctx->Nd[d] + 255) / 256, (i64)ctx->sms * 8);
        k_cover_member<<<grid, 256, 0, ctx->stream>>>(ctx->d_idx[d], ctx->d_pos[d], ctx->Nd[d], d, other->d_idx[d], other->Nd[d], keep_members,
                                                      keep_cls, keep_all);
        if ((rc = scan_exclusive(ctx, keep_cls, ctx->Nd[d], cls_scan, sums, totals + 1))) return rc;
        i64 nd = 0;
        CU(cudaMemcpyAsync(&nd, totals + 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        uint8_t* nidx = nullptr;
        i64* npos = nullptr;
        if (nd > 0) {
            CU(cudaMalloc(&nidx, (size_t)nd * d));
            CU(cudaMalloc(&npos, (size_t)nd * sizeof(i64)));
            k_cover_compact<<<grid, 256, 0, ctx->stream>>>(ctx->d_idx[d], ctx->d_pos[d], ctx->Nd[d], d, keep_cls, cls_scan, all_scan, nidx, npos);
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(ctx->stream));
        }
        cudaFree(ctx->d_idx[d]); cudaFree(ctx->d_pos[d]);
        ctx->d_idx[d] = nidx; ctx->d_pos[d] = npos; ctx->Nd[d] = nd; ctx->lo[d] = 0;
    }
    ctx->N = newN;
    if (out_N) *out_N = newN;
    return SDPCS_OK;
}

extern "C" int sdpcs_num_candidates(const sdpcs_ctx* ctx, int64_t* N)
{
    if (!ctx || !N) return SDPCS_ERR_INVALID;
    *N = ctx->N;
    return SDPCS_OK;
}

// ---------------------------------------------------------------------------------------------------
// scoring
// ---------------------------------------------------------------------------------------------------
static int upload_vars(sdpcs_ctx* ctx, const double* vars_values)
{
    const size_t len = (size_t)ctx->n * (ctx->n + 1) / 2 + ctx->n;
    CU(cudaStreamSynchronize(ctx->stream));   // pinned staging buffer is reused
    memcpy(ctx->h_vars, vars_values, len * sizeof(double));
    CU(cudaEventRecord(ctx->ev[4], ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_vars, ctx->h_vars, len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaEventRecord(ctx->ev[5], ctx->stream));
    ctx->ev_h2d = true;
    ctx->vars_resident = true;
    return SDPCS_OK;
}

constexpr int NN_WARPS = 16;

constexpr i64 I8_CHUNK_TILES = 262144;  // 33.6 M candidates, 5.2 GB of digit images per chunk (7 chunks at cfg4 size)

static int ensure_tiles(sdpcs_ctx* ctx, i64 n_tiles, i64 tile_bytes = I8_TILE_BYTES)
{
    const i64 want = n_tiles * tile_bytes;
    if (want <= ctx->tiles_cap && ctx->d_tiles) return SDPCS_OK;
    if (ctx->d_tiles) { cudaFree(ctx->d_tiles); ctx->d_tiles = nullptr; ctx->tiles_cap = 0; }
    CU(cudaMalloc(&ctx->d_tiles, (size_t)want));
    ctx->tiles_cap = want;
    return SDPCS_OK;
}

template <int NHID, int D, int NS, bool DBG>
static int launch_mlp_i8_inst(sdpcs_ctx* ctx, const MlpI8Args& m)
{
    using L = I8Smem<NHID, NS>;
    auto kern = k_mlp_i8<NHID, D, NS, DBG>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    const unsigned grid = (unsigned)std::max<i64>(1, std::min<i64>(m.n_tiles, ctx->sms));
    kern<<<grid, I8_THREADS, L::TOTAL, ctx->stream>>>(m);
    CU(cudaGetLastError());
    return SDPCS_OK;
}

// the variant that can dump the pre-activations of a layer (sdpcs_nn_debug_layer) is a separate instance: the product
// kernel carries no debug branch
template <int NHID, int D, int NS = I8_NS>
static int launch_mlp_i8(sdpcs_ctx* ctx, const MlpI8Args& m)
{
    if constexpr (D == 0) {
        if (m.dbg_z) return launch_mlp_i8_inst<NHID, D, NS, true>(ctx, m);
    }
    return launch_mlp_i8_inst<NHID, D, NS, false>(ctx, m);
}

// optimality measure of the N candidates described by `a` through the tcgen05 int8-sliced MLP, layer-0 digit images
// staged through HBM in chunks by k_prep_i8 (default: faster today) --
template <int D, int NS, bool FEAS>
static int launch_nn_i8_staged(sdpcs_ctx* ctx, const ScoreArgs& a, i64 N)
{
    const uint8_t* wimg = (NS == I8_NS) ? ctx->d_wi8[D] : ctx->d_wi8s[D];
    if (!wimg) return ctx->fail(SDPCS_ERR_STATE, "NN_" + std::to_string(D) + "D weights not set");
    const i64 total_tiles = (N + I8_M - 1) / I8_M;
    int rc = ensure_tiles(ctx, std::min<i64>(total_tiles, I8_CHUNK_TILES), I8Dig<NS>::TILE_BYTES);
    if (rc) return rc;
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_prep_i8<D, NS, FEAS>, 256, 0));
    for (i64 c0 = 0; c0 < N; c0 += I8_CHUNK_TILES * I8_M) {
        const i64 rows = std::min<i64>(N - c0, I8_CHUNK_TILES * I8_M);
        const i64 nt = (rows + I8_M - 1) / I8_M;
        PrepI8Args pa;
        pa.s = a; pa.c0 = c0; pa.n_rows = rows; pa.tiles = ctx->d_tiles; pa.status = ctx->d_status;
        const i64 groups = nt * (I8_M / 32);
        const i64 pgrid = std::min<i64>((groups + 7) / 8, (i64)ctx->sms * std::max(occ, 1));
        k_prep_i8<D, NS, FEAS><<<(unsigned)std::max<i64>(pgrid, 1), 256, 0, ctx->stream>>>(pa);
        CU(cudaGetLastError());
        MlpI8Args m;
        m.wimg = wimg; m.tiles = ctx->d_tiles; m.n_tiles = nt; m.n_rows = rows; m.out_base = c0;
        m.pos = a.pos; m.obj = a.obj; m.dbg_z = nullptr; m.dbg_layer = -1; m.status = ctx->d_status;
        if ((rc = launch_mlp_i8<NetCfg<D>::NHID, 0, NS>(ctx, m))) return rc;
        ctx->tm.score_launches += 2;
    }
    ctx->i8_used = true;
    return SDPCS_OK;
}

// optimality measure of the N candidates described by `a` through the tcgen05 int8-sliced MLP: ONE persistent launch,
// unranking / gathering / slicing happen in the kernel's producer warp (nothing but the scores touches HBM)
template <int D>
static int launch_nn_i8(sdpcs_ctx* ctx, const ScoreArgs& a, i64 N)
{
    if (!ctx->d_wi8[D]) return ctx->fail(SDPCS_ERR_STATE, "NN_" + std::to_string(D) + "D weights not set");
    const bool feas = ctx->fuse_feas;       // the staging kernel also produces lam_min (score_device)
    if (ctx->params.nn_engine == SDPCS_NN_SCREEN)
        return feas ? launch_nn_i8_staged<D, I8_NS_SCREEN, true>(ctx, a, N) : launch_nn_i8_staged<D, I8_NS_SCREEN, false>(ctx, a, N);
    if (ctx->params.nn_fused_prep != 1)
        return feas ? launch_nn_i8_staged<D, I8_NS, true>(ctx, a, N) : launch_nn_i8_staged<D, I8_NS, false>(ctx, a, N);
    MlpI8Args m;
    m.wimg = ctx->d_wi8[D]; m.s = a; m.tiles = nullptr; m.n_tiles = (N + I8_M - 1) / I8_M; m.n_rows = N; m.out_base = 0;
    m.pos = a.pos; m.obj = a.obj; m.dbg_z = nullptr; m.dbg_layer = -1; m.status = ctx->d_status;
    int rc = launch_mlp_i8<NetCfg<D>::NHID, D>(ctx, m);
    if (rc) return rc;
    ctx->tm.score_launches += 1;
    ctx->i8_used = true;
    return SDPCS_OK;
}

template <int D>
static int launch_score(sdpcs_ctx* ctx, int want, const uint8_t* idx, const i64* pos, i64 N, i64 rank_begin)
{
    if (N <= 0) return SDPCS_OK;
    ScoreArgs a;
    a.n = ctx->n; a.N = N; a.rank_begin = rank_begin; a.idx = idx; a.pos = pos;
    a.X = ctx->d_vars; a.x = ctx->d_vars + (size_t)ctx->n * (ctx->n + 1) / 2; a.Q = ctx->d_Q;
    a.wfrag = ctx->d_wfrag[D]; a.lam = ctx->d_lam; a.obj = ctx->d_obj;
    a.sweeps = ctx->params.jacobi_sweeps > 0 ? ctx->params.jacobi_sweeps : 0;   // 0: tridiagonalisation + Laguerre
    const i64 groups = (N + 31) / 32;
    const bool i8_path = ctx->params.nn_engine != SDPCS_NN_DMMA && ctx->i8_ok[D] && a.wfrag;
    if ((want & 1) && !(ctx->fuse_feas && i8_path)) {   // K1+K2+K3: lam_min of every candidate (else: by the staging kernel of K4a)
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_score_feas<D>, 256, 0));
        i64 grid = std::min<i64>((groups + 7) / 8, (i64)ctx->sms * std::max(occ, 1));
        k_score_feas<D><<<(unsigned)std::max<i64>(grid, 1), 256, 0, ctx->stream>>>(a);
        CU(cudaGetLastError());
        ctx->tm.score_launches++;
    }
    if (want & 4) {   // K1+K2 + exact SDP optimality measure (strat 3): batched barrier solver, one sub-problem per thread
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_score_sdp<D>, 128, 0));
        i64 grid = std::min<i64>((groups + 3) / 4, (i64)ctx->sms * std::max(occ, 1));
        k_score_sdp<D><<<(unsigned)std::max<i64>(grid, 1), 128, 0, ctx->stream>>>(a, ctx->params.sdp_mu_final > 0 ? ctx->params.sdp_mu_final : 1e-12);
        CU(cudaGetLastError());
        ctx->tm.score_launches++;
        return SDPCS_OK;
    }
    if ((want & 2) && !a.wfrag) return ctx->fail(SDPCS_ERR_STATE, "NN_" + std::to_string(D) + "D weights not set");
    if ((want & 2) && i8_path) {
        // K1+K2 (k_prep_i8) + K4 on tcgen05 (k_mlp_i8), chunked through the layer-0 digit-image buffer
        int rc = launch_nn_i8<D>(ctx, a, N);
        if (rc) return rc;
    } else if (want & 2) {   // K1+K2+K4 with FP64 DMMA: optimality measure of every candidate
        const size_t smem = (size_t)score_nn_smem_doubles<D>(NN_WARPS) * sizeof(double);
        auto kern = k_score_nn<D, NN_WARPS>;
        int occ = 0;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NN_WARPS * 32, smem));
        i64 grid = std::min<i64>((groups + NN_WARPS - 1) / NN_WARPS, (i64)ctx->sms * std::max(occ, 1));
        kern<<<(unsigned)std::max<i64>(grid, 1), NN_WARPS * 32, smem, ctx->stream>>>(a);
        CU(cudaGetLastError());
        ctx->tm.score_launches++;
    }
    return SDPCS_OK;
}

static int launch_score_d(sdpcs_ctx* ctx, int d, int want, const uint8_t* idx, const i64* pos, i64 N, i64 rb)
{
    switch (d) {
    case 2: return launch_score<2>(ctx, want, idx, pos, N, rb);
    case 3: return launch_score<3>(ctx, want, idx, pos, N, rb);
    case 4: return launch_score<4>(ctx, want, idx, pos, N, rb);
    case 5: return launch_score<5>(ctx, want, idx, pos, N, rb);
    }
    return ctx->fail(SDPCS_ERR_INVALID, "bad subset size");
}

static int score_device(sdpcs_ctx* ctx, int want)
{
    if (!ctx->mode) return ctx->fail(SDPCS_ERR_STATE, "cover not set");
    if (!(want & 7) || (want & ~7) || ((want & 2) && (want & 4)))
        return ctx->fail(SDPCS_ERR_INVALID, "want: bit 0 (lam), and at most one of bit 1 (NN measure) / bit 2 (exact SDP measure)");
    ctx->tm.score_launches = 0;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    int rc = SDPCS_OK;
    ctx->ev_nn = false;
    // both scores wanted and the layer-0 images staged by k_prep_i8: that kernel computes lam_min as well (one unranking
    // and one gather per candidate, eigenvalue arithmetic under the image stores); nn_fused_prep = 2 keeps the launches apart
    ctx->fuse_feas = (want & 3) == 3 && ctx->params.jacobi_sweeps <= 0 && ctx->params.nn_fused_prep == 0 &&
                     ctx->params.nn_engine != SDPCS_NN_DMMA;
    for (int bit = 1; bit <= 4 && rc == SDPCS_OK; bit <<= 1) {   // all eigenvalue launches, then all NN (or exact SDP) launches
        if (!(want & bit)) continue;
        if (bit >= 2) CU(cudaEventRecord(ctx->ev[6], ctx->stream));
        if (ctx->mode == 1) rc = launch_score_d(ctx, ctx->rho, bit, nullptr, nullptr, ctx->N, ctx->base);
        else
            for (int d = 2; d <= ctx->rho && rc == SDPCS_OK; ++d)
                rc = launch_score_d(ctx, d, bit, ctx->d_idx[d] + ctx->lo[d] * d, ctx->d_pos[d] + ctx->lo[d], ctx->Nd[d], 0);
        if (bit >= 2 && rc == SDPCS_OK) { CU(cudaEventRecord(ctx->ev[7], ctx->stream)); ctx->ev_nn = true; }
    }
    if (rc) return rc;
    if (ctx->i8_used) {
        // the tcgen05 pipeline reports through a device status word: 2 = an NN input outside the fixed-point
        // range (-2, 2) (never for LP points inside the McCormick box) -> those scores are recomputed with the
        // FP64 DMMA kernel; 1 = a pipeline barrier timed out (a bug, reported loudly)
        ctx->i8_used = false;
        CU(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        const int st = *ctx->h_status;
        if (st != 0) {
            CU(cudaMemsetAsync(ctx->d_status, 0, sizeof(int), ctx->stream));
            if (st == 1) return ctx->fail(SDPCS_ERR_CUDA, "tcgen05 MLP pipeline timed out (k_mlp_i8)");
            const int keep = ctx->params.nn_engine;
            ctx->params.nn_engine = SDPCS_NN_DMMA;
            ctx->tm.nn_fallbacks++;
            if (ctx->mode == 1) rc = launch_score_d(ctx, ctx->rho, 2, nullptr, nullptr, ctx->N, ctx->base);
            else
                for (int d = 2; d <= ctx->rho && rc == SDPCS_OK; ++d)
                    rc = launch_score_d(ctx, d, 2, ctx->d_idx[d] + ctx->lo[d] * d, ctx->d_pos[d] + ctx->lo[d], ctx->Nd[d], 0);
            ctx->params.nn_engine = keep;
            if (rc) return rc;
        }
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    ctx->ev_score = true;
    ctx->have = (want & 1) | ((want & 6) ? 2 : 0);       // the exact SDP measure takes the place of the NN measure in d_obj
    ctx->obj_exact = (want & 4) != 0;
    return SDPCS_OK;
}

extern "C" int sdpcs_score(sdpcs_ctx* ctx, const double* vars_values, int want)
{
    if (!ctx) return SDPCS_ERR_INVALID;
    if (!ctx->mode) return ctx->fail(SDPCS_ERR_STATE, "cover not set");
    CU(cudaSetDevice(ctx->device));
    if (vars_values) {
        int rc = upload_vars(ctx, vars_values);
        if (rc) return rc;
    } else if (!ctx->vars_resident) return ctx->fail(SDPCS_ERR_STATE, "vars_values == NULL but no LP point is resident");
    return score_device(ctx, want);
}

extern "C" int sdpcs_scores(sdpcs_ctx* ctx, int64_t i0, int64_t i1, double* out_lam, double* out_obj)
{
    if (!ctx) return SDPCS_ERR_INVALID;
    if (i0 < 0 || i1 < i0 || i1 > ctx->N) return ctx->fail(SDPCS_ERR_INVALID, "bad range");
    if ((out_lam && !(ctx->have & 1)) || (out_obj && !(ctx->have & 2))) return ctx->fail(SDPCS_ERR_STATE, "requested scores not resident");
    CU(cudaSetDevice(ctx->device));
    if (out_lam) CU(cudaMemcpyAsync(out_lam, ctx->d_lam + i0, (i1 - i0) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_obj) CU(cudaMemcpyAsync(out_obj, ctx->d_obj + i0, (i1 - i0) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SDPCS_OK;
}

// ---------------------------------------------------------------------------------------------------
// selection
// ---------------------------------------------------------------------------------------------------
static int ensure_out(sdpcs_ctx* ctx, i64 k)
{
    if (k <= ctx->out_cap) return SDPCS_OK;
    void** ptrs[] = {(void**)&ctx->d_c_k1, (void**)&ctx->d_c_k2, (void**)&ctx->d_s_k1, (void**)&ctx->d_s_k2, (void**)&ctx->d_c_idx,
                     (void**)&ctx->d_s_idx, (void**)&ctx->d_s_perm, (void**)&ctx->d_o_score, (void**)&ctx->d_o_lam, (void**)&ctx->d_o_obj};
    for (void** p : ptrs) {
        if (*p) cudaFree(*p);
        *p = nullptr;
        CU(cudaMalloc(p, (size_t)k * 8));
    }
    CU(cudaMemsetAsync(ctx->d_s_perm, 0, (size_t)k * 8, ctx->stream));     // partial ranks / tickets of k_rank_sort: zero between launches
    ctx->out_cap = k;
    return SDPCS_OK;
}

// output capacity for a selection of k: the winners plus room for their near ties (params.band_cap)
static i64 sel_capacity(const sdpcs_ctx* ctx, i64 k) { return std::max<i64>(k, 1) + std::max<i64>(ctx->params.band_cap, 0); }
// the radix select stops as soon as this many keys or fewer remain at or above the prefix
static i64 sel_cap_exit(const sdpcs_ctx* ctx, i64 k) { return std::min<i64>(sel_capacity(ctx, k), k + std::max<i64>(k, 4096)); }

static double guard_delta(const sdpcs_ctx* ctx, int mode)
{
    if (mode == 0) return -1.0;
    const double d = (mode == 1) ? ctx->params.guard_lam : ctx->params.guard_obj;
    return d > 0.0 ? d : 0.0;           // 0: exact ties of the k-th score only
}

template <int MODE>
static void launch_hist_scan(sdpcs_ctx* ctx, const KeySrc& ks, unsigned grid, int level, int shift, int width, bool first,
                             bool last, int next_level, i64 k, i64 cap_exit)
{
    if (first) k_sel_hist<MODE, true><<<grid, 512, 0, ctx->stream>>>(ks, level, shift, width);
    else k_sel_hist<MODE, false><<<grid, 512, 0, ctx->stream>>>(ks, level, shift, width);
    k_sel_scan<<<1, 1024, 0, ctx->stream>>>(ctx->d_state, level, shift, width, first ? 1 : 0, last ? 1 : 0, next_level, k, cap_exit);
    ctx->tm.select_launches += 2;
}

template <int MODE>
static void launch_collect_sort(sdpcs_ctx* ctx, const KeySrc& ks, i64 cap, i64 cap_exit, double delta)
{
    const unsigned cgrid = (unsigned)std::max<i64>(1, std::min<i64>((ks.N + 511) / 512, (i64)ctx->sms * 8));
    k_sel_collect<MODE><<<cgrid, 256, 0, ctx->stream>>>(ks, cap, ctx->d_c_k1, ctx->d_c_k2, ctx->d_c_idx);
    // O(m^2) rank counting spread over ~4 blocks per SM: S segments of the list per column of 256 entries; the partial
    // ranks meet in d_s_perm (zero between launches: the kernel cleans up after itself)
    const unsigned nbx = (unsigned)((cap_exit + 255) / 256);
    const unsigned S = (unsigned)std::min<i64>(32, std::max<i64>(1, (4 * (i64)ctx->sms + nbx - 1) / nbx));
    unsigned* rank = reinterpret_cast<unsigned*>(ctx->d_s_perm);
    k_rank_sort<<<dim3(nbx, S), 256, 0, ctx->stream>>>(ctx->d_state, 0, cap, ctx->d_c_k1, ctx->d_c_k2, ctx->d_c_idx, ctx->d_s_k1,
                                                        ctx->d_s_k2, ctx->d_s_idx, rank, rank + ctx->out_cap);
    k_sel_finish<<<1, 32, 0, ctx->stream>>>(ctx->d_state, ctx->d_s_k1, cap, delta);
    ctx->tm.select_launches += 3;
}

// Radix select + collect + sort over the 3-level key of `ks`; results in d_s_k1/d_s_k2/d_s_idx, counts in the state
// (copied to ctx->h_state).  Fast path: three level-0 passes, then collect (the scan kernels stop the select as soon as
// the keys above the prefix fit); the remaining passes only run when a tie class is larger than the buffers.
template <int MODE>
static int run_select_t(sdpcs_ctx* ctx, KeySrc ks, i64 k)
{
    ks.st = ctx->d_state;
    const i64 cap = sel_capacity(ctx, k), cap_exit = sel_cap_exit(ctx, k);
    ctx->sel_cap = cap;
    const double delta = guard_delta(ctx, MODE);
    const unsigned grid = (unsigned)std::max<i64>(1, std::min<i64>((ks.N + 1023) / 1024, (i64)ctx->sms * 4));
    static const int shifts0[6] = {53, 42, 31, 20, 10, 0}, widths0[6] = {11, 11, 11, 11, 10, 10};
    const int next0 = (MODE == 4) ? 1 : 2;
    for (int p = 0; p < 3; ++p) launch_hist_scan<MODE>(ctx, ks, grid, 0, shifts0[p], widths0[p], p == 0, false, next0, k, cap_exit);
    launch_collect_sort<MODE>(ctx, ks, cap, cap_exit, delta);
    CU(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(SelState), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (!ctx->h_state->done) {
        for (int p = 3; p < 6; ++p) launch_hist_scan<MODE>(ctx, ks, grid, 0, shifts0[p], widths0[p], false, p == 5, next0, k, cap_exit);
        if (MODE == 4)
            for (int p = 0; p < 6; ++p) launch_hist_scan<MODE>(ctx, ks, grid, 1, shifts0[p], widths0[p], false, p == 5, 2, k, cap_exit);
        for (int p = 0; p < 4; ++p) launch_hist_scan<MODE>(ctx, ks, grid, 2, 33 - 11 * p, 11, false, p == 3, -1, k, cap_exit);
        launch_collect_sort<MODE>(ctx, ks, cap, cap_exit, delta);
        CU(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(SelState), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (ctx->h_state->done == 2 && ctx->h_state->band_open && delta >= 0.0) {
        // near ties of the k-th score reach below the collected prefix: collect again from the band key (one more pass);
        // if that overflows the buffers, fall back to the first collection and leave the band flagged open
        const SelState first = *ctx->h_state;
        k_sel_lower<<<1, 32, 0, ctx->stream>>>(ctx->d_state);
        launch_collect_sort<MODE>(ctx, ks, cap, cap, delta);
        CU(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(SelState), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if ((i64)ctx->h_state->out_count > cap) {
            SelState redo = first;
            redo.out_count = 0;
            CU(cudaMemcpyAsync(ctx->d_state, &redo, sizeof(SelState), cudaMemcpyHostToDevice, ctx->stream));
            launch_collect_sort<MODE>(ctx, ks, cap, cap_exit, delta);
            CU(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(SelState), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            ctx->h_state->band_open = 1;
        }
    }
    CU(cudaGetLastError());
    return SDPCS_OK;
}

static int topk_device(sdpcs_ctx* ctx, int mode, i64 k, double pivot_obj, i64 pivot_idx, int all_walked)
{
    if (mode < 1 || mode > 4) return ctx->fail(SDPCS_ERR_INVALID, "mode must be 1..4");
    const int need = (mode == 1) ? 1 : (mode == 2) ? 2 : 3;
    if ((ctx->have & need) != need) return ctx->fail(SDPCS_ERR_STATE, "scores needed by this mode are not resident; call sdpcs_score");
    if (k < 0) return ctx->fail(SDPCS_ERR_INVALID, "k < 0");
    k = std::min<i64>(k, ctx->N);
    int rc = ensure_out(ctx, sel_capacity(ctx, k));
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    ctx->tm.select_launches = 0;
    k_sel_reset<<<1, 256, 0, ctx->stream>>>(ctx->d_state);
    ctx->tm.select_launches++;
    KeySrc ks;
    memset(&ks, 0, sizeof(ks));
    ks.lam = (ctx->have & 1) ? ctx->d_lam : nullptr;
    ks.obj = (ctx->have & 2) ? ctx->d_obj : nullptr;
    ks.N = ctx->N; ks.base = ctx->base;
    ks.thr_eig = ctx->params.thres_neg_eigval; ks.thr_opt = ctx->params.thres_min_opt; ks.big_m = ctx->params.big_m;
    ks.guard_lam = std::max(ctx->params.guard_lam, 0.0); ks.guard_obj = std::max(ctx->params.guard_obj, 0.0);
    ks.pivot_obj = pivot_obj; ks.pivot_idx = pivot_idx; ks.all_walked = all_walked;
    // k == 0 still runs the first pass: it fills the counters (n_violated, n_strong, ...)
    switch (mode) {
    case 1: rc = run_select_t<1>(ctx, ks, k); break;
    case 2: rc = run_select_t<2>(ctx, ks, k); break;
    case 3: rc = run_select_t<3>(ctx, ks, k); break;
    default: rc = run_select_t<4>(ctx, ks, k); break;
    }
    if (rc) return rc;
    const i64 m = std::min<i64>((i64)ctx->h_state->out_count, ctx->sel_cap);
    if (m > 0) {
        k_sel_gather<<<(unsigned)std::max<i64>(1, (m + 255) / 256), 256, 0, ctx->stream>>>(
            ctx->d_state, ctx->sel_cap, ctx->d_s_k1, ctx->d_s_idx, ctx->base, ks.lam, ks.obj, ctx->d_o_score, ctx->d_o_lam, ctx->d_o_obj);
        ctx->tm.select_launches++;
    }
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    ctx->ev_select = true;
    ctx->last_mode = mode;
    CU(cudaGetLastError());
    return SDPCS_OK;
}

// download the winners and their guard band (the state is already on the host)
static int download_topk(sdpcs_ctx* ctx, i64 k, int64_t* out_idx, double* out_score, double* out_lam, double* out_obj, int64_t* out_n)
{
    const SelState* st = ctx->h_state;
    const i64 m = std::min<i64>(st->k_out, std::min<i64>(k, ctx->N));
    const i64 band = std::max<i64>(std::min<i64>(st->band_count, ctx->sel_cap - st->k_out), 0);
    const i64 tot = st->k_out + band;
    const size_t cb = (size_t)std::max<i64>(tot, 1) * 8;
    int rc = ensure_hout(ctx, 4 * cb);
    if (rc) return rc;
    char* h = (char*)ctx->h_out;
    if (tot > 0) {
        CU(cudaMemcpyAsync(h, ctx->d_s_idx, tot * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(h + cb, ctx->d_o_score, tot * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(h + 2 * cb, ctx->d_o_lam, tot * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(h + 3 * cb, ctx->d_o_obj, tot * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->h_out_rows = tot; ctx->h_out_stride = cb; ctx->h_out_k = st->k_out;
    ctx->last_counts[0] = ctx->N; ctx->last_counts[1] = st->n_violated; ctx->last_counts[2] = st->n_strong;
    ctx->last_max_pos_nonviol = st->max_pos_nonviol ? dec_key(st->max_pos_nonviol) : -INFINITY;
    ctx->last_guard[0] = band; ctx->last_guard[1] = st->band_open; ctx->last_guard[2] = st->n_unc_lam; ctx->last_guard[3] = st->n_unc_obj;
    if (out_n) *out_n = m;
    if (out_idx) memcpy(out_idx, h, m * 8);
    if (out_score) memcpy(out_score, h + cb, m * 8);
    if (out_lam) memcpy(out_lam, h + 2 * cb, m * 8);
    if (out_obj) memcpy(out_obj, h + 3 * cb, m * 8);
    return SDPCS_OK;
}

extern "C" int sdpcs_topk(sdpcs_ctx* ctx, int mode, int64_t k, double pivot_obj, int64_t pivot_idx, int all_walked,
                          int64_t* out_idx, double* out_score, double* out_lam, double* out_obj, int64_t* out_n)
{
    if (!ctx) return SDPCS_ERR_INVALID;
    if (!ctx->mode) return ctx->fail(SDPCS_ERR_STATE, "cover not set");
    CU(cudaSetDevice(ctx->device));
    int rc = topk_device(ctx, mode, k, pivot_obj, pivot_idx, all_walked);
    if (rc) return rc;
    return download_topk(ctx, k, out_idx, out_score, out_lam, out_obj, out_n);
}

// Guard band of the last sdpcs_topk / sdpcs_select pass: the candidates ranked right after the k winners whose primary
// score lies within the guard (guard_lam for mode 1, guard_obj otherwise) of the k-th score, in selection order.
extern "C" int sdpcs_last_band(sdpcs_ctx* ctx, int64_t cap, int64_t* out_idx, double* out_score, double* out_lam,
                               double* out_obj, int64_t* out_n, int64_t* out_info)
{
    if (!ctx || cap < 0 || !out_n) return SDPCS_ERR_INVALID;
    const i64 band = std::max<i64>(ctx->h_out_rows - ctx->h_out_k, 0);
    const i64 m = std::min<i64>(band, cap);
    const char* h = (const char*)ctx->h_out;
    const size_t cb = ctx->h_out_stride, off = (size_t)ctx->h_out_k * 8;
    *out_n = m;
    if (m > 0) {
        if (out_idx) memcpy(out_idx, h + off, m * 8);
        if (out_score) memcpy(out_score, h + cb + off, m * 8);
        if (out_lam) memcpy(out_lam, h + 2 * cb + off, m * 8);
        if (out_obj) memcpy(out_obj, h + 3 * cb + off, m * 8);
    }
    if (out_info) {
        out_info[0] = ctx->last_guard[0];                                   // near ties of the k-th score after the winners
        out_info[1] = ctx->last_guard[1] || band > m;                       // 1: more near ties exist than were returned
        out_info[2] = ctx->last_guard[2];                                   // |lam - thres_neg_eigval| <= guard_lam
        out_info[3] = ctx->last_guard[3];                                   // |obj - thres_min_opt| <= guard_obj
    }
    return SDPCS_OK;
}

extern "C" int sdpcs_counts(sdpcs_ctx* ctx, int64_t* out3)
{
    if (!ctx || !out3) return SDPCS_ERR_INVALID;
    out3[0] = ctx->last_counts[0]; out3[1] = ctx->last_counts[1]; out3[2] = ctx->last_counts[2];
    return SDPCS_OK;
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU exchange on the device (exchange_kernels.cuh)
// ---------------------------------------------------------------------------------------------------
extern "C" int sdpcs_topk_pack_dev(sdpcs_ctx* ctx, int mode, int64_t k, double pivot_obj, int64_t pivot_idx, int all_walked,
                                   int64_t band_rows, void* d_block)
{
    if (!ctx || !d_block || k < 0 || band_rows < 0) return SDPCS_ERR_INVALID;
    if (!ctx->mode) return ctx->fail(SDPCS_ERR_STATE, "cover not set");
    CU(cudaSetDevice(ctx->device));
    int rc = topk_device(ctx, mode, k, pivot_obj, pivot_idx, all_walked);
    if (rc) return rc;
    const SelState* st = ctx->h_state;
    const i64 rows_cap = k + band_rows;
    const i64 band = std::max<i64>(std::min<i64>(st->band_count, ctx->sel_cap - st->k_out), 0);
    const i64 rows = std::min<i64>(st->k_out + band, rows_cap);
    PackHdr h;
    h.len = (double)rows; h.n_local = (double)ctx->N; h.n_violated = (double)st->n_violated; h.n_strong = (double)st->n_strong;
    h.extra = st->max_pos_nonviol ? dec_key(st->max_pos_nonviol) : -INFINITY;
    h.n_unc_lam = (double)st->n_unc_lam; h.n_unc_obj = (double)st->n_unc_obj;
    h.open = (st->band_open || st->k_out + band > rows_cap) ? 1.0 : 0.0;
    ctx->last_counts[0] = ctx->N; ctx->last_counts[1] = st->n_violated; ctx->last_counts[2] = st->n_strong;
    ctx->last_max_pos_nonviol = h.extra;
    k_pack_rows<<<(unsigned)std::max<i64>(1, (rows_cap + 255) / 256), 256, 0, ctx->stream>>>(h, rows, rows_cap, ctx->d_s_idx, ctx->d_o_score,
                                                                                              ctx->d_o_lam, ctx->d_o_obj, (double*)d_block);
    CU(cudaGetLastError());
    ctx->tm.select_launches++;
    return SDPCS_OK;
}

extern "C" int sdpcs_merge_packed_dev(sdpcs_ctx* ctx, const void* d_gathered, int world, int64_t rows_cap, int64_t k, int use_obj2,
                                      double delta, int64_t out_cap, int64_t* out_idx, double* out_score, double* out_lam,
                                      double* out_obj, int64_t* out_n, int64_t* out_band, double* out_hdr)
{
    if (!ctx || !d_gathered || world < 1 || rows_cap < 0 || k < 0 || out_cap < 0 || !out_n || !out_band || !out_hdr) return SDPCS_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)(16 + 4 * std::max<i64>(out_cap, 1)) * 8;
    int rc = ensure_scratch(ctx, bytes);
    if (rc) return rc;
    if ((rc = ensure_hout(ctx, bytes))) return rc;
    double* d_out = (double*)ctx->d_scratch;
    const i64 threads = (i64)world * rows_cap;
    if (threads > 0)
        k_merge_rows<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>((const double*)d_gathered, world, rows_cap, use_obj2, out_cap, d_out);
    k_merge_finish<<<1, 32, 0, ctx->stream>>>((const double*)d_gathered, world, rows_cap, k, delta, out_cap, d_out);
    CU(cudaGetLastError());
    // header first (it says how many rows are worth copying): one small copy, then the rows
    double* h = (double*)ctx->h_out;
    CU(cudaMemcpyAsync(h, d_out, 16 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    const i64 guess = std::min<i64>(out_cap, k + 64);
    if (guess > 0) CU(cudaMemcpyAsync(h + 16, d_out + 16, (size_t)guess * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const i64 nwin = (i64)h[8], nband = (i64)h[9], tot = nwin + nband;
    if (tot > guess) {
        CU(cudaMemcpyAsync(h + 16 + 4 * guess, d_out + 16 + 4 * guess, (size_t)(tot - guess) * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    for (i64 i = 0; i < tot; ++i) {
        const double* r = h + 16 + 4 * i;
        if (out_idx) out_idx[i] = (int64_t)r[0];
        if (out_score) out_score[i] = r[1];
        if (out_lam) out_lam[i] = r[2];
        if (out_obj) out_obj[i] = r[3];
    }
    *out_n = nwin; *out_band = nband;
    for (int j = 0; j < 8; ++j) out_hdr[j] = h[j];
    ctx->tm.select_launches += 2;
    return SDPCS_OK;
}

// Combined rule (cut_select_qp.py:603-625), shortcut: when the strong set S = {obj > 0 and violated} has at least k
// elements, the walk stops at the k-th of them (the pivot).  The k strong elements up to the pivot are re-scored obj + big_m;
// every other candidate keeps obj <= pivot_obj (not walked) or gets obj - big_m (walked, not violated).  If even the
// largest such obj - big_m stays strictly below pivot_obj + big_m, the re-sorted list starts with exactly those k elements
// in (obj desc, idx asc) order -- the list pass 1 already produced -- and the second selection pass is not needed.
static bool combined_is_strong_prefix(double big_m, bool all_walked, i64 k, double pivot_obj, double max_pos_nonviol)
{
    if (all_walked || k <= 0 || !(big_m > 0.0)) return false;
    const double floor_top = pivot_obj + big_m;                       // smallest re-scored strong element
    if (!(pivot_obj < floor_top)) return false;                       // overflow / absorption: take the general path
    return max_pos_nonviol == -INFINITY || (max_pos_nonviol - big_m) < floor_top;
}

// largest obj among candidates with obj > thres_min_opt that are not violated, as seen by the last sdpcs_topk pass
// (-inf if there is none): lets a sharded caller apply the same shortcut on the merged lists
extern "C" int sdpcs_max_pos_nonviolated(sdpcs_ctx* ctx, double* out)
{
    if (!ctx || !out) return SDPCS_ERR_INVALID;
    *out = ctx->last_max_pos_nonviol;
    return SDPCS_OK;
}

extern "C" int sdpcs_select(sdpcs_ctx* ctx, int strat, const double* vars_values, int64_t k, int64_t* out_idx,
                            double* out_score, double* out_lam, double* out_obj, int64_t* out_n, int64_t* out_counts,
                            int* out_new_strat)
{
    if (!ctx) return SDPCS_ERR_INVALID;
    if (strat < 1 || strat > 4) return ctx->fail(SDPCS_ERR_INVALID, "strat must be 1, 2, 3 or 4");
    if (!ctx->mode) return ctx->fail(SDPCS_ERR_STATE, "cover not set");
    CU(cudaSetDevice(ctx->device));
    k = std::min<i64>(std::max<i64>(k, 0), ctx->N);       // sel_size = min(sel_size, len(agg_list)), cut_select_qp.py:550
    int rc = SDPCS_OK;
    if (vars_values) {
        if ((rc = upload_vars(ctx, vars_values))) return rc;
    } else if (!ctx->vars_resident) return ctx->fail(SDPCS_ERR_STATE, "vars_values == NULL but no LP point is resident");
    if ((rc = score_device(ctx, strat == 1 ? 1 : strat == 2 ? 2 : strat == 3 ? 4 : 3))) return rc;
    int new_strat = strat;
    i64 counts[3] = {ctx->N, 0, 0};
    if (strat != 4) {
        if ((rc = topk_device(ctx, strat == 3 ? 2 : strat, k, 0.0, 0, 0))) return rc;       // strat 3 ranks like strat 2, by the exact measure
        if ((rc = download_topk(ctx, k, out_idx, out_score, out_lam, out_obj, out_n))) return rc;
        counts[1] = ctx->last_counts[1];
    } else {
        // pass 1: the strong set S (obj > 0 and violated) by (obj desc, idx asc) -> pivot
        int64_t ns = 0;
        if ((rc = topk_device(ctx, 3, k, 0.0, 0, 0))) return rc;
        if ((rc = download_topk(ctx, k, nullptr, nullptr, nullptr, nullptr, &ns))) return rc;
        const int64_t* sidx = (const int64_t*)ctx->h_out;
        const double* sobj = (const double*)((const char*)ctx->h_out + 3 * ctx->h_out_stride);
        const i64 n_viol = ctx->last_counts[1], n_strong_total = ctx->last_counts[2];
        // (with a guard the list holds the relaxed strong set, a superset; the decisions below use the strict counters)
        const bool all_walked = n_strong_total < k || k == 0;
        const double pobj = all_walked ? 0.0 : sobj[k - 1];
        const i64 pidx = all_walked ? 0 : sidx[k - 1];
        float ms1 = 0, ms2 = 0;
        cudaEventElapsedTime(&ms1, ctx->ev[2], ctx->ev[3]);
        if (combined_is_strong_prefix(ctx->params.big_m, all_walked, k, pobj, ctx->last_max_pos_nonviol)) {
            // the k strong elements up to the pivot get +big_m and nothing else can reach them: the final list IS the
            // strong list of pass 1 in its own order (ties after the addition fall back to obj desc, idx asc); its guard
            // band (near ties of the pivot) stays available through sdpcs_last_band with scores = obj
            if (out_n) *out_n = ns;
            const char* h = (const char*)ctx->h_out;
            const size_t cb = ctx->h_out_stride;
            for (i64 i = 0; i < ns; ++i) {
                if (out_idx) out_idx[i] = sidx[i];
                if (out_score) out_score[i] = sobj[i] + ctx->params.big_m;
            }
            if (out_lam) memcpy(out_lam, h + 2 * cb, ns * 8);
            if (out_obj) memcpy(out_obj, h + 3 * cb, ns * 8);
        } else {
            // pass 2: final measure of cut_select_qp.py:603-625
            if ((rc = topk_device(ctx, 4, k, pobj, pidx, all_walked ? 1 : 0))) return rc;
            if ((rc = download_topk(ctx, k, out_idx, out_score, out_lam, out_obj, out_n))) return rc;
            cudaEventElapsedTime(&ms2, ctx->ev[2], ctx->ev[3]);
        }
        ctx->tm.select_ms = ms1 + ms2;
        ctx->ev_select = false;
        const i64 strong = std::min<i64>(n_strong_total, k);
        const i64 viol_walked = all_walked ? n_viol : k;
        counts[1] = viol_walked; counts[2] = strong;
        if (k > 0 && ctx->N > 0)
            new_strat = ((double)strong / (double)k < (double)viol_walked / (double)ctx->N) ? 1 : 4;   // cut_select_qp.py:629
    }
    if (out_counts) { out_counts[0] = counts[0]; out_counts[1] = counts[1]; out_counts[2] = counts[2]; }
    if (out_new_strat) *out_new_strat = new_strat;
    return SDPCS_OK;
}

extern "C" int sdpcs_merge_topk(sdpcs_ctx* ctx, int64_t m, const double* score, const double* obj2, const int64_t* idx,
                                int64_t k, int64_t* out_perm, int64_t* out_n)
{
    if (!ctx || m < 0 || (m && (!score || !idx)) || !out_perm || !out_n) return SDPCS_ERR_INVALID;
    *out_n = 0;
    if (m == 0 || k <= 0) return SDPCS_OK;
    const i64 take = std::min<i64>(k, m);
    // order: key1 desc, key2 desc, agg_idx asc (= the device selection order, select_kernels.cuh)
    auto before = [&](i64 a, i64 b) {
        const u64 a1 = enc_key(score[a]), b1 = enc_key(score[b]);
        if (a1 != b1) return a1 > b1;
        const u64 a2 = obj2 ? enc_key(obj2[a]) : 0, b2 = obj2 ? enc_key(obj2[b]) : 0;
        if (a2 != b2) return a2 > b2;
        return idx[a] < idx[b];
    };
    // The callers hand over the concatenation of per-shard lists, each already in selection order: find the sorted
    // runs and take the first k of their merge (k x runs comparisons on the host, no device round trip).
    std::vector<i64> run_begin{0};
    for (i64 i = 1; i < m; ++i)
        if (before(i, i - 1)) run_begin.push_back(i);
    const i64 runs = (i64)run_begin.size();
    if (runs <= 256) {
        std::vector<i64> head(run_begin), end(runs);
        for (i64 r = 0; r < runs; ++r) end[r] = (r + 1 < runs) ? run_begin[r + 1] : m;
        for (i64 o = 0; o < take; ++o) {
            i64 best = -1;
            for (i64 r = 0; r < runs; ++r)
                if (head[r] < end[r] && (best < 0 || before(head[r], head[best]))) best = r;
            out_perm[o] = head[best]++;
        }
        *out_n = take;
        return SDPCS_OK;
    }
    // unsorted input: rank-counting sort of all m entries on the device
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_scratch(ctx, (size_t)m * 8 * 9);
    if (rc) return rc;
    double* d_score = (double*)ctx->d_scratch;
    double* d_obj2 = d_score + m;
    i64* d_idx = (i64*)(d_obj2 + m);
    u64* d_k1 = (u64*)(d_idx + m);
    u64* d_k2 = d_k1 + m;
    u64* d_s1 = d_k2 + m;
    u64* d_s2 = d_s1 + m;
    i64* d_si = (i64*)(d_s2 + m);
    CU(cudaMemcpyAsync(d_score, score, m * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (obj2) CU(cudaMemcpyAsync(d_obj2, obj2, m * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_idx, idx, m * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_merge_keys<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(m, d_score, obj2 ? d_obj2 : nullptr, d_k1, d_k2);
    k_rank_sort<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(nullptr, m, m, d_k1, d_k2, d_idx, d_s1, d_s2, d_si, nullptr, nullptr);
    CU(cudaGetLastError());
    std::vector<i64> sorted_idx(m), in_idx(idx, idx + m);
    CU(cudaMemcpyAsync(sorted_idx.data(), d_si, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    // agg_idx values are unique across shards: map them back to input positions
    std::vector<std::pair<i64, i64>> where(m);
    for (i64 i = 0; i < m; ++i) where[i] = {in_idx[i], i};
    std::sort(where.begin(), where.end());
    for (i64 i = 0; i < take; ++i) {
        auto it = std::lower_bound(where.begin(), where.end(), std::make_pair(sorted_idx[i], (i64)-1));
        out_perm[i] = it->second;
    }
    *out_n = take;
    return SDPCS_OK;
}

// ---------------------------------------------------------------------------------------------------
// cuts, eigendecomposition, triangles, NN, peaks
// ---------------------------------------------------------------------------------------------------
extern "C" int sdpcs_gen_cuts(sdpcs_ctx* ctx, int rho, const int16_t* sets, int64_t m, const double* vars_values,
                              int64_t* out_ind, double* out_val, double* out_rhs, double* out_lam, uint8_t* out_violated,
                              double* out_gap)
{
    if (!ctx || m < 0 || rho < 2 || rho > 5) return SDPCS_ERR_INVALID;
    if (!ctx->n) return ctx->fail(SDPCS_ERR_STATE, "set_instance first");
    if (m == 0) return SDPCS_OK;
    if (!sets || !vars_values || !out_ind || !out_val || !out_rhs || !out_lam || !out_violated) return ctx->fail(SDPCS_ERR_INVALID, "null pointer");
    for (i64 i = 0; i < m * rho; ++i)
        if (sets[i] >= ctx->n) return ctx->fail(SDPCS_ERR_INVALID, "subset index out of range");
    CU(cudaSetDevice(ctx->device));
    int rc = upload_vars(ctx, vars_values);
    if (rc) return rc;
    const int width = rho + rho * (rho + 1) / 2;
    const size_t b_sets = ((size_t)m * rho * 2 + 255) / 256 * 256, b_w = (size_t)m * width * 8, b_m = (size_t)m * 8;
    if ((rc = ensure_scratch(ctx, b_sets + 2 * b_w + 4 * b_m))) return rc;
    char* s = (char*)ctx->d_scratch;
    CutArgs a;
    a.n = ctx->n; a.rho = rho; a.m = m;
    a.sets = (const int16_t*)s;
    a.ind = (i64*)(s + b_sets); a.val = (double*)(s + b_sets + b_w);
    a.rhs = (double*)(s + b_sets + 2 * b_w); a.lam = a.rhs + m; a.gap = a.lam + m; a.violated = (uint8_t*)(a.gap + m);
    a.X = ctx->d_vars; a.x = ctx->d_vars + (size_t)ctx->n * (ctx->n + 1) / 2;
    a.thr_eig = ctx->params.thres_neg_eigval;
    a.sweeps = ctx->params.jacobi_sweeps > 0 ? ctx->params.jacobi_sweeps : default_sweeps(rho + 1);
    CU(cudaMemcpyAsync(s, sets, (size_t)m * rho * 2, cudaMemcpyHostToDevice, ctx->stream));
    k_gen_cuts<<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>(a);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_ind, a.ind, b_w, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_val, a.val, b_w, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_rhs, a.rhs, b_m, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_lam, a.lam, b_m, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_gap) CU(cudaMemcpyAsync(out_gap, a.gap, b_m, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_violated, a.violated, (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SDPCS_OK;
}

// The violated eigenvector cuts of sdpcs_gen_cuts as CSR rows (the layout CPXaddrows takes: row starts, column
// indices, values, right-hand sides; sense >= for all), in one shot instead of one SparsePair object per cut
// (cut_select_qp.py:747-754).  out_src[r] = index into `sets` of emitted row r.  With guard_lam > 0 a subset whose lam_min is
// within the guard of the threshold is emitted too (out_lam tells; the caller decides with the reference's arithmetic).
extern "C" int sdpcs_gen_cuts_csr(sdpcs_ctx* ctx, int rho, const int16_t* sets, int64_t m, const double* vars_values,
                                  int64_t* out_rowptr, int64_t* out_ind, double* out_val, double* out_rhs, int64_t* out_src,
                                  double* out_lam, double* out_gap, int64_t* out_nrows)
{
    if (!ctx || m < 0 || rho < 2 || rho > 5 || !out_rowptr || !out_nrows) return SDPCS_ERR_INVALID;
    *out_nrows = 0;
    out_rowptr[0] = 0;
    if (m == 0) return SDPCS_OK;
    if (!out_ind || !out_val || !out_rhs) return ctx->fail(SDPCS_ERR_INVALID, "null pointer");
    const int width = rho + rho * (rho + 1) / 2;
    std::vector<int64_t> ind((size_t)m * width);
    std::vector<double> val((size_t)m * width), rhs(m), lam(m), gap(m);
    std::vector<uint8_t> viol(m);
    int rc = sdpcs_gen_cuts(ctx, rho, sets, m, vars_values, ind.data(), val.data(), rhs.data(), lam.data(), viol.data(), gap.data());
    if (rc) return rc;
    const double relaxed = ctx->params.thres_neg_eigval + std::max(ctx->params.guard_lam, 0.0);
    i64 rows = 0, nnz = 0;
    for (i64 i = 0; i < m; ++i) {
        if (!(lam[i] < relaxed)) continue;                        // eigvals[0] >= _THRES_NEG_EIGVAL: no cut (cut_select_qp.py:743)
        for (int t = 0; t < width; ++t) {
            if (ind[i * width + t] < 0) break;                    // rows of cliques smaller than rho are -1 padded
            out_ind[nnz] = ind[i * width + t];
            out_val[nnz] = val[i * width + t];
            ++nnz;
        }
        out_rhs[rows] = rhs[i];
        if (out_src) out_src[rows] = i;
        if (out_lam) out_lam[rows] = lam[i];
        if (out_gap) out_gap[rows] = gap[i];
        out_rowptr[++rows] = nnz;
    }
    *out_nrows = rows;
    return SDPCS_OK;
}

// Triangle-inequality rows (cut_select_qp.py:846-860) for (triple rank, type) pairs as CSR; host utility.
extern "C" int sdpcs_triangle_rows_csr(int n, const int64_t* triple_rank, const int8_t* type, int64_t m, int64_t* out_rowptr,
                                       int64_t* out_ind, double* out_val, double* out_rhs)
{
    if (n < 3 || n > 65535 || m < 0 || !out_rowptr) return SDPCS_ERR_INVALID;
    out_rowptr[0] = 0;
    if (m == 0) return SDPCS_OK;
    if (!triple_rank || !type || !out_ind || !out_val || !out_rhs) return SDPCS_ERR_INVALID;
    const i64 T = (i64)binom_small(n, 3), nb_lifted = (i64)n * (n + 1) / 2;
    static const double COEF[4][6] = {{-1, -1, 1, 1, 0, 0}, {-1, 1, -1, 1, 0, 0}, {1, -1, -1, 1, 0, 0}, {1, 1, 1, -1, -1, -1}};
    // rank -> triple in closed form (combinatorial number system of the REVERSE rank: T - 1 - rank = C(x,3) + C(y,2) + z with
    // x > y > z >= 0, triple = (n-1-x, n-1-y, n-1-z)): a cube root and a square root with integer correction instead of
    // lex_unrank's O(n) walk or binary searches (10,000 rows per round; branch mispredictions dominated those)
    i64 nnz = 0;
    for (i64 r = 0; r < m; ++r) {
        if (triple_rank[r] < 0 || triple_rank[r] >= T || type[r] < 0 || type[r] > 3) return SDPCS_ERR_INVALID;
        int c[3];
        {
            i64 rem = T - 1 - triple_rank[r];
            i64 x = (i64)std::cbrt(6.0 * (double)rem) + 1;                  // C(x,3) ~ x^3 / 6
            while (x * (x - 1) * (x - 2) / 6 > rem) --x;
            while ((x + 1) * x * (x - 1) / 6 <= rem) ++x;
            rem -= x * (x - 1) * (x - 2) / 6;
            i64 y = (i64)((1.0 + std::sqrt(1.0 + 8.0 * (double)rem)) * 0.5);  // C(y,2) ~ y^2 / 2
            while (y * (y - 1) / 2 > rem) --y;
            while ((y + 1) * y / 2 <= rem) ++y;
            const i64 z = rem - y * (y - 1) / 2;
            c[0] = (int)(n - 1 - x); c[1] = (int)(n - 1 - y); c[2] = (int)(n - 1 - z);
        }
        const i64 i1 = c[0], i2 = c[1], i3 = c[2];
        out_ind[nnz] = n * i1 - i1 * (i1 + 1) / 2 + i2;
        out_ind[nnz + 1] = n * i1 - i1 * (i1 + 1) / 2 + i3;
        out_ind[nnz + 2] = n * i2 - i2 * (i2 + 1) / 2 + i3;
        const int t = type[r];
        int w = 4;
        if (t == 3) {
            out_ind[nnz + 3] = i1 + nb_lifted; out_ind[nnz + 4] = i2 + nb_lifted; out_ind[nnz + 5] = i3 + nb_lifted;
            w = 6;
        } else {
            out_ind[nnz + 3] = c[t] + nb_lifted;
        }
        for (int q = 0; q < w; ++q) out_val[nnz + q] = COEF[t][q];
        out_rhs[r] = (t == 3) ? -1.0 : 0.0;
        nnz += w;
        out_rowptr[r + 1] = nnz;
    }
    return SDPCS_OK;
}

// Dense eigenvalue cuts, strat 0 (cut_select_qp.py:757-786): eigen-decomposition of the full [1 x^T; x X] on the
// device (dense_kernels.cuh), one dense row per eigenvalue below thres_neg_eigval among the n smallest.
extern "C" int sdpcs_dense_eigcuts(sdpcs_ctx* ctx, const double* vars_values, int64_t max_cuts, double* out_eigvals,
                                   int64_t* out_ncuts, double* out_val, double* out_rhs)
{
    if (!ctx || !vars_values || !out_ncuts || max_cuts < 0) return SDPCS_ERR_INVALID;
    *out_ncuts = 0;
    if (!ctx->n) return ctx->fail(SDPCS_ERR_STATE, "set_instance first");
    const int n = ctx->n, m = n + 1, mp = (m + 1) & ~1;
    if (mp > DENSE_MAX_ORDER) return ctx->fail(SDPCS_ERR_INVALID, "dense eigcuts need n <= 255");
    CU(cudaSetDevice(ctx->device));
    int rc = upload_vars(ctx, vars_values);
    if (rc) return rc;
    const i64 width = (i64)n + (i64)n * (n + 1) / 2;
    const size_t b_mat = sizeof(double) * mp * mp, b_eig = sizeof(double) * mp, b_cols = 1024 /* <= 256 ints */;
    const i64 cap = std::min<i64>(max_cuts, n);
    if ((rc = ensure_scratch(ctx, 2 * b_mat + b_eig + b_cols + 256 + sizeof(double) * (size_t)cap * (width + 1)))) return rc;
    char* sc = static_cast<char*>(ctx->d_scratch);
    DenseEigArgs a;
    a.n = n; a.mp = mp;
    a.X = ctx->d_vars; a.x = ctx->d_vars + (size_t)n * (n + 1) / 2;
    a.A = reinterpret_cast<double*>(sc); a.V = reinterpret_cast<double*>(sc + b_mat); a.eig = reinterpret_cast<double*>(sc + 2 * b_mat);
    int* d_cols = reinterpret_cast<int*>(sc + 2 * b_mat + b_eig);
    a.sweeps_done = d_cols + 255;
    a.max_sweeps = 30;
    double* d_val = reinterpret_cast<double*>(sc + 2 * b_mat + b_eig + b_cols + 256);
    double* d_rhs = d_val + (size_t)cap * width;
    k_dense_jacobi<<<1, DENSE_THREADS, 0, ctx->stream>>>(a);
    CU(cudaGetLastError());
    std::vector<double> eig(mp);
    int sweeps = 0;
    CU(cudaMemcpyAsync(eig.data(), a.eig, b_eig, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&sweeps, a.sweeps_done, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (sweeps >= a.max_sweeps) return ctx->fail(SDPCS_ERR_CUDA, "dense Jacobi eigensolver did not converge");
    // ascending order as LAPACK returns it; the padding row/column (mp > m) is an exact zero eigenvalue: drop one
    std::vector<int> order;
    for (int i = 0; i < m; ++i) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](int p, int q) { return eig[p] < eig[q]; });
    if (out_eigvals)
        for (int i = 0; i < m; ++i) out_eigvals[i] = eig[order[i]];
    std::vector<int> cols;
    for (int ix = 0; ix < n; ++ix)                                 // the n smallest (cut_select_qp.py:772-773)
        if (eig[order[ix]] < ctx->params.thres_neg_eigval) cols.push_back(order[ix]);
    const i64 ncuts = (i64)cols.size();
    *out_ncuts = ncuts;
    if (ncuts == 0) return SDPCS_OK;
    if (ncuts > max_cuts || !out_val || !out_rhs)
        return ctx->fail(SDPCS_ERR_INVALID, "dense eigcuts: " + std::to_string(ncuts) + " cuts, output capacity " + std::to_string(max_cuts));
    CU(cudaMemcpyAsync(d_cols, cols.data(), sizeof(int) * ncuts, cudaMemcpyHostToDevice, ctx->stream));
    k_dense_rows<<<dim3((unsigned)(n + 1), (unsigned)ncuts), 256, 0, ctx->stream>>>(n, mp, a.V, d_cols, d_val, d_rhs);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_val, d_val, sizeof(double) * (size_t)ncuts * width, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(out_rhs, d_rhs, sizeof(double) * ncuts, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SDPCS_OK;
}

extern "C" int sdpcs_eigendecomp(sdpcs_ctx* ctx, int d, const double* curr_pt, const double* X_slice, double* out_vals,
                                 double* out_vecs)
{
    if (!ctx || d < 2 || d > 5 || !curr_pt || !X_slice || !out_vals) return SDPCS_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    const int M = d + 1, t = d * (d + 1) / 2;
    int rc = ensure_scratch(ctx, 8 * (size_t)(d + t + M + M * M));
    if (rc) return rc;
    double* s = (double*)ctx->d_scratch;
    CU(cudaMemcpyAsync(s, curr_pt, d * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(s + d, X_slice, t * 8, cudaMemcpyHostToDevice, ctx->stream));
    const int sweeps = ctx->params.jacobi_sweeps > 0 ? ctx->params.jacobi_sweeps : default_sweeps(M);
    k_eig_one<<<1, 1, 0, ctx->stream>>>(d, s, s + d, sweeps, s + d + t, s + d + t + M);
    CU(cudaGetLastError());
    double vals[6], vecs[36];
    CU(cudaMemcpyAsync(vals, s + d + t, M * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(vecs, s + d + t + M, M * M * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    int order[6];
    for (int i = 0; i < M; ++i) order[i] = i;
    std::stable_sort(order, order + M, [&](int a, int b) { return vals[a] < vals[b]; });   // ascending, as LAPACK
    for (int j = 0; j < M; ++j) {
        out_vals[j] = vals[order[j]];
        if (out_vecs)
            for (int i = 0; i < M; ++i) out_vecs[i * M + j] = vecs[i * M + order[j]];
    }
    return SDPCS_OK;
}

extern "C" int sdpcs_set_tri_pattern(sdpcs_ctx* ctx, const uint8_t* adj)
{
    if (!ctx) return SDPCS_ERR_INVALID;
    if (!ctx->n) return ctx->fail(SDPCS_ERR_STATE, "set_instance first");
    CU(cudaSetDevice(ctx->device));
    if (ctx->d_adj) { cudaFree(ctx->d_adj); ctx->d_adj = nullptr; }
    ctx->have_adj = false;
    if (!adj) return SDPCS_OK;
    CU(cudaMalloc(&ctx->d_adj, (size_t)ctx->n * ctx->n));
    CU(cudaMemcpy(ctx->d_adj, adj, (size_t)ctx->n * ctx->n, cudaMemcpyHostToDevice));
    ctx->have_adj = true;
    return SDPCS_OK;
}

extern "C" int sdpcs_triangles(sdpcs_ctx* ctx, const double* vars_values, int64_t kmax, int64_t* out_triple_rank,
                               int8_t* out_type, double* out_viol, int8_t* out_density, int64_t* out_n,
                               int64_t* out_n_violated, int64_t* out_n_triples)
{
    if (!ctx || !vars_values || kmax < 0 || !out_n) return SDPCS_ERR_INVALID;
    if (!ctx->n) return ctx->fail(SDPCS_ERR_STATE, "set_instance first");
    if (ctx->n < 3) return ctx->fail(SDPCS_ERR_INVALID, "n < 3");
    CU(cudaSetDevice(ctx->device));
    int rc = upload_vars(ctx, vars_values);
    if (rc) return rc;
    const i64 T = (i64)binom_small(ctx->n, 3), NK = 4 * T;
    kmax = std::min<i64>(kmax, NK);
    if ((rc = ensure_dev(ctx, ctx->d_key1, ctx->key_cap, NK))) return rc;
    if ((rc = ensure_out(ctx, sel_capacity(ctx, kmax)))) return rc;
    CU(cudaMemsetAsync(ctx->d_tri_counters, 0, 2 * sizeof(unsigned long long), ctx->stream));
    TriArgs ta;
    ta.n = ctx->n; ta.T = T; ta.X = ctx->d_vars; ta.x = ctx->d_vars + (size_t)ctx->n * (ctx->n + 1) / 2;
    ta.adj = ctx->have_adj ? ctx->d_adj : nullptr;
    ta.thres_dense = ctx->params.thres_tri_dense; ta.thres_viol = ctx->params.thres_tri_viol;
    ta.key = ctx->d_key1; ta.counters = ctx->d_tri_counters;
    const i64 groups = (T + 31) / 32;
    k_tri_keys<<<(unsigned)std::max<i64>(1, std::min<i64>((groups + 7) / 8, (i64)ctx->sms * 8)), 256, 0, ctx->stream>>>(ta);
    k_sel_reset<<<1, 256, 0, ctx->stream>>>(ctx->d_state);
    CU(cudaGetLastError());
    i64 m = 0;
    unsigned long long counters[2] = {0, 0};
    if (kmax > 0) {
        KeySrc ks;
        memset(&ks, 0, sizeof(ks));
        ks.key1 = ctx->d_key1; ks.N = NK; ks.base = 0;
        if ((rc = run_select_t<0>(ctx, ks, kmax))) return rc;
        // unpack into the output scratch arrays (reuse d_c_* as typed outputs)
        i64* d_rank = ctx->d_c_idx; double* d_viol = ctx->d_o_score;
        int8_t* d_type = (int8_t*)ctx->d_c_k1; int8_t* d_dens = (int8_t*)ctx->d_c_k2;
        m = std::min<i64>(ctx->h_state->k_out, kmax);
        if (m > 0) {
            k_tri_unpack<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>(m, ctx->d_s_k1, ctx->d_s_idx, d_rank, d_type, d_viol, d_dens);
            CU(cudaGetLastError());
            if (out_triple_rank) CU(cudaMemcpyAsync(out_triple_rank, d_rank, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
            if (out_viol) CU(cudaMemcpyAsync(out_viol, d_viol, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
            if (out_type) CU(cudaMemcpyAsync(out_type, d_type, m, cudaMemcpyDeviceToHost, ctx->stream));
            if (out_density) CU(cudaMemcpyAsync(out_density, d_dens, m, cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    CU(cudaMemcpyAsync(counters, ctx->d_tri_counters, sizeof(counters), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *out_n = m;
    if (out_n_violated) *out_n_violated = (int64_t)counters[1];
    if (out_n_triples) *out_n_triples = (int64_t)counters[0];
    return SDPCS_OK;
}

static int read_i8_status(sdpcs_ctx* ctx, int* st)
{
    CU(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *st = *ctx->h_status;
    if (*st) CU(cudaMemsetAsync(ctx->d_status, 0, sizeof(int), ctx->stream));
    return SDPCS_OK;
}

// batched forward pass of raw input rows through the tcgen05 engine (sdpcs_nn_eval, sdpcs_nn_debug_layer)
template <int D, int NS>
static int launch_nn_i8_raw_ns(sdpcs_ctx* ctx, const double* d_in, i64 m, double* d_out, double* d_z, int dbg_layer)
{
    const uint8_t* wimg = (NS == I8_NS) ? ctx->d_wi8[D] : ctx->d_wi8s[D];
    for (i64 c0 = 0; c0 < m; c0 += I8_CHUNK_TILES * I8_M) {
        const i64 rows = std::min<i64>(m - c0, I8_CHUNK_TILES * I8_M);
        const i64 nt = (rows + I8_M - 1) / I8_M;
        int rc = ensure_tiles(ctx, std::min<i64>((m + I8_M - 1) / I8_M, I8_CHUNK_TILES), I8Dig<NS>::TILE_BYTES);
        if (rc) return rc;
        const unsigned pgrid = (unsigned)std::max<i64>(1, std::min<i64>((nt * I8_M + 255) / 256, (i64)ctx->sms * 8));
        k_prep_i8_raw<D, NS><<<pgrid, 256, 0, ctx->stream>>>(ctx->d_wfrag[D], d_in + c0 * NetCfg<D>::NIN, rows, ctx->d_tiles, ctx->d_status);
        CU(cudaGetLastError());
        MlpI8Args a;
        a.wimg = wimg; a.tiles = ctx->d_tiles; a.n_tiles = nt; a.n_rows = rows; a.out_base = c0;
        a.pos = nullptr; a.obj = d_out; a.dbg_z = d_z ? d_z + c0 * 64 : nullptr; a.dbg_layer = dbg_layer; a.status = ctx->d_status;
        if ((rc = launch_mlp_i8<NetCfg<D>::NHID, 0, NS>(ctx, a))) return rc;
    }
    return SDPCS_OK;
}

// batched forward pass of raw input rows through the tcgen05 engine (sdpcs_nn_eval, sdpcs_nn_debug_layer); the screening
// engine (4 digits) when params.nn_engine says so
template <int D>
static int launch_nn_i8_raw(sdpcs_ctx* ctx, const double* d_in, i64 m, double* d_out, double* d_z, int dbg_layer)
{
    if (ctx->params.nn_engine == SDPCS_NN_SCREEN) return launch_nn_i8_raw_ns<D, I8_NS_SCREEN>(ctx, d_in, m, d_out, d_z, dbg_layer);
    return launch_nn_i8_raw_ns<D, I8_NS>(ctx, d_in, m, d_out, d_z, dbg_layer);
}

template <int D>
static int launch_nn(sdpcs_ctx* ctx, const double* d_in, i64 m, double* d_out)
{
    const size_t smem = (size_t)score_nn_smem_doubles<D>(NN_WARPS) * sizeof(double);
    auto kern = k_nn_eval<D, NN_WARPS>;
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 groups = (m + 31) / 32;
    const unsigned grid = (unsigned)std::max<i64>(1, std::min<i64>((groups + NN_WARPS - 1) / NN_WARPS, ctx->sms));
    kern<<<grid, NN_WARPS * 32, smem, ctx->stream>>>(ctx->d_wfrag[D], d_in, m, d_out);
    CU(cudaGetLastError());
    return SDPCS_OK;
}

extern "C" int sdpcs_nn_eval(sdpcs_ctx* ctx, int rho, const double* inputs, int64_t m, double* out)
{
    if (!ctx || rho < 2 || rho > 5 || m < 0 || (m && (!inputs || !out))) return SDPCS_ERR_INVALID;
    if (!ctx->d_wfrag[rho]) return ctx->fail(SDPCS_ERR_STATE, "NN weights not set for this rho");
    if (m == 0) return SDPCS_OK;
    CU(cudaSetDevice(ctx->device));
    const int nin = rho * (rho + 3) / 2;
    int rc = ensure_scratch(ctx, (size_t)m * (nin + 1) * 8);
    if (rc) return rc;
    double* d_in = (double*)ctx->d_scratch;
    double* d_out = d_in + (size_t)m * nin;
    CU(cudaMemcpyAsync(d_in, inputs, (size_t)m * nin * 8, cudaMemcpyHostToDevice, ctx->stream));
    bool dmma = ctx->params.nn_engine == SDPCS_NN_DMMA || !ctx->i8_ok[rho];
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (dmma) {
            switch (rho) {
            case 2: rc = launch_nn<2>(ctx, d_in, m, d_out); break;
            case 3: rc = launch_nn<3>(ctx, d_in, m, d_out); break;
            case 4: rc = launch_nn<4>(ctx, d_in, m, d_out); break;
            default: rc = launch_nn<5>(ctx, d_in, m, d_out); break;
            }
            if (rc) return rc;
            break;
        }
        switch (rho) {
        case 2: rc = launch_nn_i8_raw<2>(ctx, d_in, m, d_out, nullptr, -1); break;
        case 3: rc = launch_nn_i8_raw<3>(ctx, d_in, m, d_out, nullptr, -1); break;
        case 4: rc = launch_nn_i8_raw<4>(ctx, d_in, m, d_out, nullptr, -1); break;
        default: rc = launch_nn_i8_raw<5>(ctx, d_in, m, d_out, nullptr, -1); break;
        }
        if (rc) return rc;
        int st = 0;
        if ((rc = read_i8_status(ctx, &st))) return rc;
        if (st == 0) break;
        if (st == 1) return ctx->fail(SDPCS_ERR_CUDA, "tcgen05 MLP pipeline timed out (k_mlp_i8)");
        dmma = true;                      // inputs outside the fixed-point range: FP64 DMMA engine
        ctx->tm.nn_fallbacks++;
    }
    CU(cudaMemcpyAsync(out, d_out, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SDPCS_OK;
}

extern "C" int sdpcs_nn_debug_layer(sdpcs_ctx* ctx, int rho, const double* inputs, int64_t m, int layer, double* out_z)
{
    if (!ctx || rho < 2 || rho > 5 || m <= 0 || !inputs || !out_z || layer < 0) return SDPCS_ERR_INVALID;
    if (!ctx->d_wi8[rho]) return ctx->fail(SDPCS_ERR_STATE, "NN weights not set for this rho");
    CU(cudaSetDevice(ctx->device));
    const int nin = rho * (rho + 3) / 2;
    const i64 rows = (m + I8_M - 1) / I8_M * I8_M;
    int rc = ensure_scratch(ctx, (size_t)m * (nin + 1) * 8 + (size_t)rows * 64 * 8);
    if (rc) return rc;
    double* d_in = (double*)ctx->d_scratch;
    double* d_out = d_in + (size_t)m * nin;
    double* d_z = d_out + m;
    CU(cudaMemcpyAsync(d_in, inputs, (size_t)m * nin * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(d_z, 0, (size_t)rows * 64 * 8, ctx->stream));
    switch (rho) {
    case 2: rc = launch_nn_i8_raw<2>(ctx, d_in, m, d_out, d_z, layer); break;
    case 3: rc = launch_nn_i8_raw<3>(ctx, d_in, m, d_out, d_z, layer); break;
    case 4: rc = launch_nn_i8_raw<4>(ctx, d_in, m, d_out, d_z, layer); break;
    default: rc = launch_nn_i8_raw<5>(ctx, d_in, m, d_out, d_z, layer); break;
    }
    if (rc) return rc;
    int st = 0;
    if ((rc = read_i8_status(ctx, &st))) return rc;
    if (st) return ctx->fail(SDPCS_ERR_CUDA, st == 1 ? "tcgen05 MLP pipeline timed out (k_mlp_i8)" : "NN input outside (-2, 2)");
    CU(cudaMemcpyAsync(out_z, d_z, (size_t)m * 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SDPCS_OK;
}

// Batched exact SDP values (strat 3 / figure 8 / training-data sampler): m sub-problems of size d given as rows
// [x (d) | C upper triangle row-major, off-diagonals with the weight of the pair (d(d+1)/2)] -> v(x, C) (sdp_kernels.cuh)
extern "C" int sdpcs_sdp_solve(sdpcs_ctx* ctx, int d, const double* in, int64_t m, double* out, int32_t* out_iters)
{
    if (!ctx || d < 2 || d > 5 || m < 0 || (m && (!in || !out))) return SDPCS_ERR_INVALID;
    if (m == 0) return SDPCS_OK;
    CU(cudaSetDevice(ctx->device));
    const int w = d + d * (d + 1) / 2;
    int rc = ensure_scratch(ctx, (size_t)m * (w + 1) * 8 + (size_t)m * 4);
    if (rc) return rc;
    double* d_in = (double*)ctx->d_scratch;
    double* d_out = d_in + (size_t)m * w;
    int* d_it = (int*)(d_out + m);
    CU(cudaMemcpyAsync(d_in, in, (size_t)m * w * 8, cudaMemcpyHostToDevice, ctx->stream));
    const double mu = ctx->params.sdp_mu_final > 0 ? ctx->params.sdp_mu_final : 1e-12;
    const unsigned grid = (unsigned)((m + 127) / 128);
    switch (d) {
    case 2: k_sdp_solve<2><<<grid, 128, 0, ctx->stream>>>(d_in, m, mu, d_out, d_it); break;
    case 3: k_sdp_solve<3><<<grid, 128, 0, ctx->stream>>>(d_in, m, mu, d_out, d_it); break;
    case 4: k_sdp_solve<4><<<grid, 128, 0, ctx->stream>>>(d_in, m, mu, d_out, d_it); break;
    default: k_sdp_solve<5><<<grid, 128, 0, ctx->stream>>>(d_in, m, mu, d_out, d_it); break;
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_out, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_iters) CU(cudaMemcpyAsync(out_iters, d_it, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SDPCS_OK;
}

extern "C" int sdpcs_fp64_peak(sdpcs_ctx* ctx, double* dfma_tflops, double* dmma_tflops)
{
    if (!ctx) return SDPCS_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_scratch(ctx, 8192 * 8);
    if (rc) return rc;
    double* d = (double*)ctx->d_scratch;
    std::vector<double> h(4096);
    for (int i = 0; i < 4096; ++i) h[i] = 1e-3 / (1.0 + i % 7);
    CU(cudaMemcpyAsync(d, h.data(), 4096 * 8, cudaMemcpyHostToDevice, ctx->stream));
    const int iters = 8192, grid = ctx->sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    double best[2] = {0, 0};
    for (int which = 0; which < 2; ++which)
        for (int rep = 0; rep < 5; ++rep) {
            CU(cudaEventRecord(e0, ctx->stream));
            if (which == 0) k_peak_dfma<<<grid, threads, 0, ctx->stream>>>(d + 4096, d, iters);
            else k_peak_dmma<<<grid, threads, 0, ctx->stream>>>(d + 4096, d, iters);
            CU(cudaEventRecord(e1, ctx->stream));
            CU(cudaEventSynchronize(e1));
            float ms;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            double fl = which == 0 ? 2.0 * 8 * iters * (double)grid * threads : 2.0 * 256 * 8 * iters * (double)grid * (threads / 32);
            if (rep) best[which] = std::max(best[which], fl / ms * 1e-9);
        }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (dfma_tflops) *dfma_tflops = best[0];
    if (dmma_tflops) *dmma_tflops = best[1];
    return SDPCS_OK;
}

#ifdef SDPCS_I8_TRACE
// trace build only (tools/i8_trace.py): clock64 stamps of CTA 0, [warp 0..17][step][4]
extern "C" int sdpcs_i8_trace_read(long long* out, int64_t cap)
{
    const size_t bytes = sizeof(long long) * (I8_EPI_WARPS + 4) * I8_TRACE_STEPS * 4;
    if ((size_t)cap * sizeof(long long) < bytes) return SDPCS_ERR_INVALID;
    if (cudaMemcpyFromSymbol(out, g_i8_trace, bytes) != cudaSuccess) return SDPCS_ERR_CUDA;
    return (I8_EPI_WARPS + 4) * I8_TRACE_STEPS * 4;
}
#endif
