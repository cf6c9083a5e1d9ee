// K4 on the 5th-generation tensor cores: NN_rhoD as an FP64-accurate int8-sliced contraction (tcgen05.mma kind::i8,
// accumulators in TMEM) instead of FP64 DMMA.  Replaces `nns[d-2][0](input_arr)` (cut_select_qp.py:579-582).
//
// Arithmetic (oracle/nn_i8_model.py is the bit-exact CPU statement of it):
//   * every layer input a_k (|a| <= 1 after tansig, |p| < 2 after mapminmax) is rounded to a fixed-point integer
//     v = rint(a * 2^51) (2^50 for layer 0), and the two's complement word 8v is cut into 7 base-256 digits: six
//     unsigned low digits and one signed top digit (no re-balancing arithmetic in the epilogue; the MMAs on the top
//     slice use a signed A operand, the others an unsigned one);
//   * every weight row W[j,:] is scaled by a power of two >= max|W[j,:]| and rounded to 54 fractional bits, written
//     as 7 balanced digits (host side, once per net);
//   * digit-pair products are exact in the int32 tensor-core accumulators; the 28 pairs (s,t) with s + t <= 6
//     are kept, grouped into 7 "diagonals" d = s + t (pairs on a diagonal share a TMEM accumulator);
//   * the diagonals are recombined exactly in two int64 words, converted with one FP64 rounding, scaled and
//     biased with one DFMA, then tansig (13 FP64 operations) as in the DMMA path.
//   The dropped pairs are below 2^-54 of the row scale per term: the result differs from an exactly rounded FP64
//   evaluation by about as much as two FP64 evaluations with different summation orders differ from each other
//   (tools/i8_mlp_model.py: 4.2e-14 rms vs 2.2e-14 rms for plain FP64 at rho = 5, raw network output).
//
// Pipeline per CTA (one per SM, persistent over a contiguous range of 128-candidate tiles, two tiles in flight):
//   warp 17      : TMA producer  -- cp.async.bulk of the layer-0 digit image of a tile (written by k_prep_i8)
//   warp 16      : MMA issuer    -- one thread issues the 28 digit-pair products of a layer as 10 tcgen05.mma per k
//                                   step (N up to 256: one A slice against up to four stacked weight slices) into 7
//                                   TMEM accumulators (one per diagonal, 64 columns each), one commit per layer
//   warps 0..15  : epilogue      -- tcgen05.ld a diagonal, int64 accumulate, DFMA + tansig, re-slice the activations
//                                   into the next layer's A operand in shared memory (UMMA canonical K-major layout)
//   While the epilogue warps work on tile X, the tensor core runs the next layer of tile Y.
#pragma once
#include <type_traits>
#include "score_kernels.cuh"

namespace sdpcs {

constexpr int I8_NS = 7;                                  // digits per operand of the FP64-accurate engine
constexpr int I8_NS_SCREEN = 4;                           // digits of the screening engine (32-bit fixed point)
constexpr int I8_M = 128;                                 // candidates per tile (UMMA M)
constexpr int I8_N = 64;                                  // neurons (UMMA N); 50-neuron nets are zero padded
constexpr int I8_K0 = 32;                                 // padded input width (UMMA K of layer 0)
constexpr int I8_SLOTS = 8;                               // barrier slots reserved for the TMEM accumulator stages
constexpr int I8_AUX_BYTES = I8_M * 16;                   // base[128], max_elem[128]

// Everything that depends on the number of base-256 digits NS per operand.  NS = 7: 56-bit words, 28 digit pairs
// (s + t <= 6) on 7 diagonals -- FP64-accurate.  NS = 4: 32-bit words, 10 pairs on 4 diagonals: activations carry 2^-27,
// the NN output ~1e-7 -- the SCREENING engine (its scores only decide which candidates the exact engine re-evaluates).
// The diagonals of one step take NS x 64 TMEM columns: two accumulator stages fit for NS = 4 (the MMAs of the next step
// never wait for the epilogue of this one), one for NS = 7.
__host__ __device__ constexpr double i8_pow2(int e) { return e <= 0 ? 1.0 : 2.0 * i8_pow2(e - 1); }

template <int NS>
struct I8Dig {
    static constexpr int ND = NS;                                      // diagonals kept (s + t <= NS - 1)
    static constexpr int A0_BYTES = NS * I8_M * I8_K0;                 // layer-0 A image of a tile
    static constexpr int AH_BYTES = NS * I8_M * 64;                    // hidden-layer A image of a tile
    // Staged tile in global memory (k_prep_i8 -> TMA): compact -- the 16 leading input digits of every row and slice,
    // then the (at most 4) remaining ones as one word, then aux; the zero padding of the K = 32 image is added in smem.
    static constexpr int G_MAIN = NS * I8_M * 16;
    static constexpr int G_TAIL = NS * I8_M * 4;
    static constexpr int TILE_BYTES = G_MAIN + G_TAIL + I8_AUX_BYTES;  // NS = 7: 19968 bytes per tile (156 B per candidate)
    static constexpr int W0_BYTES = NS * I8_N * I8_K0;
    static constexpr int WH_BYTES = NS * I8_N * 64;
    static constexpr int STAGES = (2 * NS * I8_N <= 512) ? 2 : 1;     // TMEM accumulator stages
    static constexpr int KA = 8 * NS - 5;                              // hidden activations: u = 8 rint(a 2^KA), |a| <= 1
    static constexpr int KW = 8 * NS - 2;                              // weights: rint(w 2^(KW - e)) with 2^e >= max|row|
    static constexpr double SCALE_H = i8_pow2(KA);                     // 2^51 (NS = 7), 2^27 (NS = 4)
    static constexpr double SCALE_0 = i8_pow2(KA - 1);                 // layer-0 inputs |p| < 2: one bit less
};
constexpr int I8_A0_BYTES = I8Dig<I8_NS>::A0_BYTES;       // 28672
constexpr int I8_AH_BYTES = I8Dig<I8_NS>::AH_BYTES;       // 57344
constexpr int I8_G_MAIN = I8Dig<I8_NS>::G_MAIN;           // 14336
constexpr int I8_G_TAIL = I8Dig<I8_NS>::G_TAIL;           // 3584
constexpr int I8_TILE_BYTES = I8Dig<I8_NS>::TILE_BYTES;   // 19968
constexpr int I8_W0_BYTES = I8Dig<I8_NS>::W0_BYTES;       // 14336
constexpr int I8_WH_BYTES = I8Dig<I8_NS>::WH_BYTES;       // 28672
constexpr int I8_EPI_WARPS = 16;
constexpr int I8_THREADS = (I8_EPI_WARPS + 4) * 32;       // 640: 16 epilogue warps + one warp group of helpers (MMA, TMA, 2 idle)
constexpr int I8_NBAR = 2 * I8_SLOTS + 8;                 // slot_full[8] slot_empty[8] a0_full[2] lane_free[2] act_ready[2] y_ready[2]

// instruction descriptor: D = S32, B = signed int8, A = unsigned int8 (bit 7 set: signed, for the top slice), both
// K-major, M = 128, N = n (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t i8_idesc(int n)
{
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(I8_M >> 4) << 24);
}

// fixed-point constants
#define I8_MAGIC52 6755399441055744.0                     /* 1.5 * 2^52 */
#define I8_MAGIC52_BITS 0x4338000000000000ll

template <int NHID, int NS = I8_NS>
struct I8Smem {
    using G = I8Dig<NS>;
    static constexpr int W_TOTAL = G::W0_BYTES + (NHID - 1) * G::WH_BYTES;
    static constexpr int OFF_A = W_TOTAL;                                  // 2 lanes x AH_BYTES
    static constexpr int OFF_AUX = OFF_A + 2 * G::AH_BYTES;                // [lane][buf][256] doubles
    static constexpr int OFF_PAR = OFF_AUX + 4 * I8_AUX_BYTES;
    // parameter block (doubles): (cs, bs)[NHID][64] interleaved pairs, wout[64] misc[4] tab[256]
    static constexpr int P_CS = 0, P_BS = NHID * 64, P_WOUT = 2 * NHID * 64, P_MISC = P_WOUT + 64, P_TAB = P_MISC + 4;
    static constexpr int PAR = P_TAB + 256;
    static constexpr int OFF_BAR = OFF_PAR + PAR * 8;
    static constexpr int OFF_XG = OFF_BAR + I8_NBAR * 8 + 16;               // mapminmax xoffset[24], gain[24] (fused producer)
    // partial sums of the output layer, [lane][tile parity][column quarter][row].  With ONE accumulator stage they live
    // in the lane's A buffer behind the layer-0 image (no warp can start the next step of that lane before every warp
    // has finished this one); with TWO stages a fast warp may already be writing the next tile's activations there
    // while the finalising warp still reads the sums, so they get their own 16 KB.
    static constexpr int OFF_Y = OFF_XG + 2 * 24 * 8;
    static constexpr int Y_BYTES = (G::STAGES == 2) ? 2 * 2 * 4 * I8_M * 8 : 0;
    static constexpr int TOTAL = OFF_Y + Y_BYTES;
    static constexpr int GLOBAL_BYTES = W_TOTAL + PAR * 8;                 // device blob: images then parameters
};

struct MlpI8Args {
    const uint8_t* wimg;     // weight digit images followed by the FP64 parameter block (I8Smem::GLOBAL_BYTES)
    ScoreArgs s;             // tiles == nullptr: instance, cover and LP point; the producer warp unranks, gathers and
                             // slices the layer-0 digit image of each tile straight into shared memory (K1 + K2)
    const uint8_t* tiles;    // raw-input mode (sdpcs_nn_eval): n_tiles x I8_TILE_BYTES written by k_prep_i8_raw, TMA loaded
    i64 n_tiles;
    i64 n_rows;              // valid candidates in this chunk
    i64 out_base;            // local candidate index of row 0 of the chunk
    const i64* pos;          // LIST mode: output position per local candidate, or nullptr
    double* obj;
    double* dbg_z;           // debug: scaled pre-activations of layer dbg_layer, [row][64]; nullptr in production
    int dbg_layer;
    int* status;             // device word: 0 ok, 1 pipeline time-out, 2 NN input outside (-2, 2)
};

// Optional pipeline trace (tools/i8_trace.py builds a second library with -DSDPCS_I8_TRACE): CTA 0 stamps clock64 at
// four points of steps 64 .. 64 + I8_TRACE_STEPS of every warp.  Compiled out of the product library.
#ifdef SDPCS_I8_TRACE
constexpr int I8_TRACE_STEPS = 96;
__device__ long long g_i8_trace[I8_EPI_WARPS + 4][I8_TRACE_STEPS][4];
#define I8_STAMP(STEP, K)                                                                                              \
    do {                                                                                                               \
        if (blockIdx.x == 0 && lane == 0 && (STEP) - 64u < (uint32_t)I8_TRACE_STEPS)                                   \
            g_i8_trace[warp][(STEP) - 64u][K] = clock64();                                                              \
    } while (0)
#else
#define I8_STAMP(STEP, K) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive by lane 0 of a converged warp as a predicated instruction (no branch: the step stays one basic block for ptxas)
__device__ __forceinline__ void mbar_arrive_lane0(uint32_t bar, int lane)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, 0;\n\t@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(bar), "r"(lane) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0;
}
// same with a suspend-time hint: the warp sleeps in hardware until the phase completes or `ns` nanoseconds have passed,
// so that a waiting helper warp takes (almost) no issue slots from the epilogue warps of its scheduler
__device__ __forceinline__ bool mbar_try_hint(uint32_t bar, uint32_t parity, uint32_t ns)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity), "r"(ns)
                 : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error ends as status = 1 and a clean kernel exit, never as a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag, int* status)
{
    if (mbar_try(bar, parity)) return true;
    const long long t0 = clock64();
    while (true) {
        if (mbar_try_hint(bar, parity, 20000u)) return true;
        if (*abort_flag) return false;
        if (clock64() - t0 > (1ll << 31)) {
            *abort_flag = 1;
            atomicCAS(status, 0, 1);
            return false;
        }
    }
}
__device__ __forceinline__ bool mbar_wait_relaxed(uint32_t bar, uint32_t parity, volatile int* abort_flag, int* status)
{
    return mbar_wait(bar, parity, abort_flag, status);
}
// one lane of a converged warp (elect.sync): the compiler then knows the branch is single-threaded and keeps
// tcgen05 / TMA operands in uniform registers instead of emitting a per-lane waterfall loop
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred)::"memory");
    return pred != 0;
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// named barrier among the four epilogue warps of one TMEM lane quarter (they share an SM sub-partition): keeps them
// in lock-step so that no warp is left to finish a step alone at single-warp issue rate
__device__ __forceinline__ void quarter_sync(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory"); }
// Named barriers for the hand-offs from the 16 epilogue warps to the MMA issuer (staged inputs): the epilogue warps arrive
// without waiting, the issuer warp blocks in the hardware barrier unit (like the idle helper warps that wait for the end of the
// kernel in bar.sync) instead of polling an mbarrier.
constexpr int I8_NB_EMPTY = 1, I8_NB_ACT = 3;             // named barrier ids: empty[stage 0..1], act_ready[lane 0..1]
constexpr int I8_NB_COUNT = (I8_EPI_WARPS + 1) * 32;     // 16 arriving warps + the issuer warp
__device__ __forceinline__ void nb_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(I8_NB_COUNT) : "memory"); }
__device__ __forceinline__ void nb_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(I8_NB_COUNT) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle: 8 x 16 B core matrices, LBO = step between the two
// 16-byte K chunks of one instruction, SBO = step between 8-row groups (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(taddr),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// tansig_scaled (device_math.cuh) on W independent arguments, written stage by stage so that the W dependency
// chains are interleaved in the instruction stream (the epilogue has only four warps per scheduler: the FP64
// latency has to be covered by instruction-level parallelism).
// |zs| < I8_Z_MAX is guaranteed by the host (pack_i8 bounds the pre-activations of every layer from the weights and
// refuses the engine otherwise), so -|zs| needs no clamp and enters the FP64 pipe through operand modifiers.
#define I8_Z_MAX 960.0
template <int W>
__device__ __forceinline__ void tansig_scaled_vec(const double (&zs)[W], double (&out)[W], const double* __restrict__ T)
{
    const double A1 = 0.6931471805599453094, A2 = 0.2402265069591007123, A3 = 0.0555041086648215800,
                 A4 = 0.0096181291076284772;
    const double MAGIC = 26388279066624.0;  // 1.5 * 2^44: ulp = 2^-8
    int idx[W];
    double s[W], q[W], t[W], d[W], y0[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const double kf = MAGIC - fabs(zs[i]);
        idx[i] = __double2loint(kf);
        s[i] = (MAGIC - kf) - fabs(zs[i]);            // -|z| - rint(-|z| 2^8) 2^-8
    }
#pragma unroll
    for (int i = 0; i < W; ++i) q[i] = fma(A4, s[i], A3);
#pragma unroll
    for (int i = 0; i < W; ++i) q[i] = fma(q[i], s[i], A2);
#pragma unroll
    for (int i = 0; i < W; ++i) q[i] = fma(q[i], s[i], A1);
#pragma unroll
    for (int i = 0; i < W; ++i) q[i] = q[i] * s[i];
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const double Tj = T[idx[i] & 255];
        const double t0 = fma(Tj, q[i], Tj);
        t[i] = __hiloint2double(__double2hiint(t0) + (int)((uint32_t)(idx[i] & ~255) << 12), __double2loint(t0));
    }
#pragma unroll
    for (int i = 0; i < W; ++i) {
        d[i] = 1.0 + t[i];                      // in (1, 2]
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0[i]) : "d"(d[i]));   // MUFU.RCP64H: ~20-bit seed, cubic step below
    }
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const double e = fma(-d[i], y0[i], 1.0);
        q[i] = fma(e, e, e);
    }
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const double y = fma(y0[i], q[i], y0[i]);
        const double r = fma(2.0, y, -1.0);      // in [0, 1): tanh(|n|); the result takes the sign of n = -zs / (2 log2 e)
        out[i] = __hiloint2double(__double2hiint(r) ^ (~__double2hiint(zs[i]) & (int)0x80000000), __double2loint(r));
    }
}

// ---------------------------------------------------------------------------------------------------
// fixed-point slicing
// ---------------------------------------------------------------------------------------------------
// u = 8 * rint(a * 2^51) for |a| <= 1 (scale = 2^51), or with scale = 2^50 for |a| < 2, as a two's complement word:
// bytes 0..5 are the six low digits (unsigned), byte 6 is the signed top digit.
__device__ __forceinline__ unsigned long long i8_quantize(double a, double scale)
{
    const double q = fma(a, scale, I8_MAGIC52);
    const long long v = __double_as_longlong(q) - I8_MAGIC52_BITS;
    return (unsigned long long)(v << 3);
}

// gather byte b of four 64-bit digit words into one 32-bit word (neuron j -> byte j), b = 0..6
template <int B>
__device__ __forceinline__ uint32_t i8_pack4(unsigned long long u0, unsigned long long u1, unsigned long long u2, unsigned long long u3)
{
    constexpr int b = B & 3;
    const uint32_t x0 = B < 4 ? (uint32_t)u0 : (uint32_t)(u0 >> 32), x1 = B < 4 ? (uint32_t)u1 : (uint32_t)(u1 >> 32);
    const uint32_t x2 = B < 4 ? (uint32_t)u2 : (uint32_t)(u2 >> 32), x3 = B < 4 ? (uint32_t)u3 : (uint32_t)(u3 >> 32);
    const uint32_t t01 = __byte_perm(x0, x1, ((4 + b) << 4) | b);
    const uint32_t t23 = __byte_perm(x2, x3, ((4 + b) << 4) | b);
    return __byte_perm(t01, t23, 0x5410);
}

// digit words of four fixed-point values, slice by slice: w[s] holds digit (NS - 1 - s) of u0..u3 (slice 0 = most
// significant).  Two 4 x 4 byte transposes (PRMT butterflies): 8 + 7 operations for 7 digits instead of 21.
template <int NS>
__device__ __forceinline__ void i8_pack_slices(unsigned long long u0, unsigned long long u1, unsigned long long u2, unsigned long long u3,
                                               uint32_t (&w)[NS])
{
    static_assert(NS >= 4 && NS <= 7, "digits 0..3 come from the low words, 4..NS-1 from the high words");
    {
        const uint32_t x0 = (uint32_t)u0, x1 = (uint32_t)u1, x2 = (uint32_t)u2, x3 = (uint32_t)u3;
        const uint32_t a01 = __byte_perm(x0, x1, 0x5140), b01 = __byte_perm(x0, x1, 0x7362);   // bytes (0,1) / (2,3) of x0, x1 interleaved
        const uint32_t a23 = __byte_perm(x2, x3, 0x5140), b23 = __byte_perm(x2, x3, 0x7362);
        w[NS - 1] = __byte_perm(a01, a23, 0x5410);
        w[NS - 2] = __byte_perm(a01, a23, 0x7632);
        w[NS - 3] = __byte_perm(b01, b23, 0x5410);
        w[NS - 4] = __byte_perm(b01, b23, 0x7632);
    }
    if constexpr (NS > 4) {
        const uint32_t x0 = (uint32_t)(u0 >> 32), x1 = (uint32_t)(u1 >> 32), x2 = (uint32_t)(u2 >> 32), x3 = (uint32_t)(u3 >> 32);
        const uint32_t a01 = __byte_perm(x0, x1, 0x5140), a23 = __byte_perm(x2, x3, 0x5140);
        w[NS - 5] = __byte_perm(a01, a23, 0x5410);
        if constexpr (NS > 5) w[NS - 6] = __byte_perm(a01, a23, 0x7632);
        if constexpr (NS > 6) {
            const uint32_t b01 = __byte_perm(x0, x1, 0x7362), b23 = __byte_perm(x2, x3, 0x7362);
            w[NS - 7] = __byte_perm(b01, b23, 0x5410);
        }
    }
}

// tansig of four pre-activations (as tansig_scaled_vec<4>) with independent "side work" woven into it, stage by stage:
// side(integral_constant<int, k>), k = 0..5, is called between the stages.  The four epilogue warps of an SM sub-partition
// run in lock-step (they wait for the same accumulators), so a stretch of pure FP64 work in the instruction stream
// saturates the FP64 pipe (one warp instruction every two clocks) while the integer pipes idle, and a stretch of pure
// integer work does the opposite: measured step time was the SUM of the two.  Woven, the integer work (re-slicing of the
// previous four activations, recombination of the next accumulators) issues in the shadow of the FP64 pipe.
template <int K> using i8_stage = std::integral_constant<int, K>;
template <class Side>
__device__ __forceinline__ void tansig4_woven(const double (&zs)[4], double (&out)[4], const double* __restrict__ T, Side&& side)
{
    const double A1 = 0.6931471805599453094, A2 = 0.2402265069591007123, A3 = 0.0555041086648215800,
                 A4 = 0.0096181291076284772;
    const double MAGIC = 26388279066624.0;  // 1.5 * 2^44: ulp = 2^-8
    int idx[4];
    double s[4], q[4], t[4], d[4], y0[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double kf = MAGIC - fabs(zs[i]);
        idx[i] = __double2loint(kf);
        s[i] = (MAGIC - kf) - fabs(zs[i]);
    }
    side(i8_stage<0>{});
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = fma(A4, s[i], A3);
    side(i8_stage<1>{});
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = fma(q[i], s[i], A2);
    side(i8_stage<2>{});
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = fma(q[i], s[i], A1);
    side(i8_stage<3>{});
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = q[i] * s[i];
    side(i8_stage<4>{});
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double Tj = T[idx[i] & 255];
        const double t0 = fma(Tj, q[i], Tj);
        t[i] = __hiloint2double(__double2hiint(t0) + (int)((uint32_t)(idx[i] & ~255) << 12), __double2loint(t0));
    }
    side(i8_stage<5>{});
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        d[i] = 1.0 + t[i];
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0[i]) : "d"(d[i]));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double e = fma(-d[i], y0[i], 1.0);
        q[i] = fma(e, e, e);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double y = fma(y0[i], q[i], y0[i]);
        const double r = fma(2.0, y, -1.0);
        out[i] = __hiloint2double(__double2hiint(r) ^ (~__double2hiint(zs[i]) & (int)0x80000000), __double2loint(r));
    }
}

// side work: digit slicing of four activations (i8_quantize + i8_pack_slices), cut into the six stages of tansig4_woven
template <int NS>
struct I8SliceSide {
    const double (&prev)[4];
    double scale;
    uint32_t (&w)[NS];
    double qd[4];
    uint32_t lo[4], hi[4], a01, a23, b01, b23;
    __device__ __forceinline__ I8SliceSide(const double (&p)[4], double sc, uint32_t (&ww)[NS]) : prev(p), scale(sc), w(ww) {}
    template <int K>
    __device__ __forceinline__ void operator()(i8_stage<K>)
    {
        if constexpr (K == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) qd[i] = fma(prev[i], scale, I8_MAGIC52);
        } else if constexpr (K == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {                                               // (bits - magic) << 3 as two words
                const uint32_t l32 = (uint32_t)__double2loint(qd[i]), h32 = (uint32_t)__double2hiint(qd[i]) - 0x43380000u;
                lo[i] = l32 << 3;
                hi[i] = __funnelshift_l(l32, h32, 3);
            }
        } else if constexpr (K == 2) {
            a01 = __byte_perm(lo[0], lo[1], 0x5140); b01 = __byte_perm(lo[0], lo[1], 0x7362);
            a23 = __byte_perm(lo[2], lo[3], 0x5140); b23 = __byte_perm(lo[2], lo[3], 0x7362);
        } else if constexpr (K == 3) {
            w[NS - 1] = __byte_perm(a01, a23, 0x5410);
            w[NS - 2] = __byte_perm(a01, a23, 0x7632);
            w[NS - 3] = __byte_perm(b01, b23, 0x5410);
            w[NS - 4] = __byte_perm(b01, b23, 0x7632);
        } else if constexpr (K == 4) {
            if constexpr (NS > 4) {
                a01 = __byte_perm(hi[0], hi[1], 0x5140);
                a23 = __byte_perm(hi[2], hi[3], 0x5140);
                if constexpr (NS > 6) {
                    b01 = __byte_perm(hi[0], hi[1], 0x7362);
                    b23 = __byte_perm(hi[2], hi[3], 0x7362);
                }
            }
        } else {
            if constexpr (NS > 4) w[NS - 5] = __byte_perm(a01, a23, 0x5410);
            if constexpr (NS > 5) w[NS - 6] = __byte_perm(a01, a23, 0x7632);
            if constexpr (NS > 6) w[NS - 7] = __byte_perm(b01, b23, 0x5410);
        }
    }
};
struct I8NoSide {
    template <int K>
    __device__ __forceinline__ void operator()(i8_stage<K>) {}
};

// exact recombination of the NS diagonals of one output: value = sum_d S_d 256^(NS-1-d) as exact integers, pre-biased so that
// the bit pattern is the double 1.5 * 2^52 + acc; one rounding; z = -2 log2(e) * (W a + b)
template <int NS>
__device__ __forceinline__ double i8_recombine(const uint32_t (&v)[NS][8], int j, double2 cb)
{
    double val;
    if constexpr (NS == 7) {
        long long accH = I8_MAGIC52_BITS, accL = I8_MAGIC52_BITS;
        accH = (long long)(int)v[0][j] * 65536ll + accH;
        accH = (long long)(int)v[1][j] * 256ll + accH;
        accH = (long long)(int)v[2][j] * 1ll + accH;
        accL = (long long)(int)v[3][j] * 16777216ll + accL;
        accL = (long long)(int)v[4][j] * 65536ll + accL;
        accL = (long long)(int)v[5][j] * 256ll + accL;
        accL = (long long)(int)v[6][j] * 1ll + accL;
        const double dl = __longlong_as_double(accL) - I8_MAGIC52;
        const double dh = __longlong_as_double(accH) - I8_MAGIC52;
        val = fma(dh, 4294967296.0, dl);
    } else {
        static_assert(NS == 7 || NS <= 4, "recombination written for 7 and for <= 4 digits");
        long long acc = I8_MAGIC52_BITS;           // |sum| < 2^(24 + 8 (NS - 1)) <= 2^48
#pragma unroll
        for (int dg = 0; dg < NS; ++dg) acc = (long long)(int)v[dg][j] * (1ll << (8 * (NS - 1 - dg))) + acc;
        val = __longlong_as_double(acc) - I8_MAGIC52;
    }
    return fma(val, cb.x, cb.y);
}
// side work: recombination of outputs 4..7 of the second batch of TMEM loads (z[12..15]), one per stage
template <int NS>
struct I8RecombineSide {
    const uint32_t (&v)[NS][8];
    const double2* csbs;
    double (&z)[16];
    __device__ __forceinline__ I8RecombineSide(const uint32_t (&vv)[NS][8], const double2* c, double (&zz)[16]) : v(vv), csbs(c), z(zz) {}
    template <int K>
    __device__ __forceinline__ void operator()(i8_stage<K>)
    {
        if constexpr (K < 4) z[12 + K] = i8_recombine<NS>(v, 4 + K, csbs[12 + K]);
    }
};

// The same recombination in two parts (NS = 7): the integer part -- the two pre-biased int64 words -- and the FP64 part.
// While the tensor core executes tcgen05.mma, FP64 instructions and shared-memory loads of the SM get ~8 % of their
// throughput (tools/tc_fp64_overlap.cu) but integer arithmetic runs at full speed: the epilogue keeps the stretch right
// after the accumulator hand-back (when the MMAs of the other tile start) free of FP64 work.  Same operations, same
// order as i8_recombine<7>: bit-identical results.
template <int NS>
__device__ __forceinline__ void i8_recombine_int7(const uint32_t (&v)[NS][8], int j, long long& accH, long long& accL)
{
    accH = I8_MAGIC52_BITS; accL = I8_MAGIC52_BITS;
    if constexpr (NS == 7) {      // (only called for NS = 7; a template so that the other instances of the kernel still compile)
        accH = (long long)(int)v[0][j] * 65536ll + accH;
        accH = (long long)(int)v[1][j] * 256ll + accH;
        accH = (long long)(int)v[2][j] * 1ll + accH;
        accL = (long long)(int)v[3][j] * 16777216ll + accL;
        accL = (long long)(int)v[4][j] * 65536ll + accL;
        accL = (long long)(int)v[5][j] * 256ll + accL;
        accL = (long long)(int)v[6][j] * 1ll + accL;
    }
}
__device__ __forceinline__ double i8_recombine_fin7(long long accH, long long accL, double2 cb)
{
    const double dl = __longlong_as_double(accL) - I8_MAGIC52;
    const double dh = __longlong_as_double(accH) - I8_MAGIC52;
    return fma(fma(dh, 4294967296.0, dl), cb.x, cb.y);
}
// side work: recombination of outputs JB .. JB + 3 of the batch in v into z[ZB ..], one per stage
template <int NS, int JB, int ZB, int NZ>
struct I8RecombineSideAt {
    const uint32_t (&v)[NS][8];
    const double2* cb;          // parameters of output JB
    double (&z)[NZ];
    __device__ __forceinline__ I8RecombineSideAt(const uint32_t (&vv)[NS][8], const double2* c, double (&zz)[NZ]) : v(vv), cb(c), z(zz) {}
    template <int K>
    __device__ __forceinline__ void operator()(i8_stage<K>)
    {
        if constexpr (K < 4) z[ZB + K] = i8_recombine<NS>(v, JB + K, cb[K]);
    }
};
// integer part of the digit slicing of four activations whose quantisation FMA (qd = fma(a, scale, 1.5 * 2^52)) is done:
// stages 1..5 of I8SliceSide
template <int NS>
__device__ __forceinline__ void i8_slice_from_qd(const double (&qd)[4], uint32_t (&w)[NS])
{
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t l32 = (uint32_t)__double2loint(qd[i]), h32 = (uint32_t)__double2hiint(qd[i]) - 0x43380000u;
        lo[i] = l32 << 3;
        hi[i] = __funnelshift_l(l32, h32, 3);
    }
    {
        const uint32_t a01 = __byte_perm(lo[0], lo[1], 0x5140), b01 = __byte_perm(lo[0], lo[1], 0x7362);
        const uint32_t a23 = __byte_perm(lo[2], lo[3], 0x5140), b23 = __byte_perm(lo[2], lo[3], 0x7362);
        w[NS - 1] = __byte_perm(a01, a23, 0x5410);
        w[NS - 2] = __byte_perm(a01, a23, 0x7632);
        w[NS - 3] = __byte_perm(b01, b23, 0x5410);
        w[NS - 4] = __byte_perm(b01, b23, 0x7632);
    }
    if constexpr (NS > 4) {
        const uint32_t a01 = __byte_perm(hi[0], hi[1], 0x5140), a23 = __byte_perm(hi[2], hi[3], 0x5140);
        w[NS - 5] = __byte_perm(a01, a23, 0x5410);
        if constexpr (NS > 5) w[NS - 6] = __byte_perm(a01, a23, 0x7632);
        if constexpr (NS > 6) {
            const uint32_t b01 = __byte_perm(hi[0], hi[1], 0x7362), b23 = __byte_perm(hi[2], hi[3], 0x7362);
            w[NS - 7] = __byte_perm(b01, b23, 0x5410);
        }
    }
}

// Write one row (candidate) of a layer-0 tile image: NIN mapminmax'ed inputs -> NS slices x 32 digit bytes, plus aux.
// Shared-memory image: [slice s (0 = most significant)][k chunk c = k / 16][row][k % 16]; staged tile: see I8Dig::TILE_BYTES.
template <int NIN, int NS>
__device__ __forceinline__ void i8_store_row(uint8_t* tile, int row, const double (&p)[NIN], double base, double max_elem,
                                             bool valid, int* status)
{
    using G = I8Dig<NS>;
    uint32_t w[5][NS];                           // words of 4 input digits: 16 inputs of the first k chunk + inputs 16..19
    bool bad = false;
#pragma unroll
    for (int g = 0; g < 5; ++g) {
        unsigned long long u[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = 4 * g + t;
            const double pk = (k < NIN && valid) ? p[k < NIN ? k : 0] : 0.0;
            if (k < NIN) bad |= !(fabs(pk) < 2.0);
            u[t] = i8_quantize(pk, G::SCALE_0);
        }
        i8_pack_slices<NS>(u[0], u[1], u[2], u[3], w[g]);
    }
    if (bad && valid) atomicCAS(status, 0, 2);
    static_assert(NIN <= 20, "the compact tile format keeps one word of the second k chunk");
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        *reinterpret_cast<uint4*>(tile + s * (I8_M * 16) + row * 16) = make_uint4(w[0][s], w[1][s], w[2][s], w[3][s]);
        *reinterpret_cast<uint32_t*>(tile + G::G_MAIN + s * (I8_M * 4) + row * 4) = w[4][s];     // inputs 16..19
    }
    double* aux = reinterpret_cast<double*>(tile + G::G_MAIN + G::G_TAIL);
    aux[row] = valid ? base : 0.0;
    aux[I8_M + row] = valid ? max_elem : 0.0;
}

// ---------------------------------------------------------------------------------------------------
// K1 + K2 + input preparation: unrank, gather x_rho / X_rho / Q_rho, max_elem, <Q~, X>, mapminmax, digit slicing.
// Same per-candidate arithmetic as k_score_nn (cut_select_qp.py:536-538, 573-575; neural_net_3D.m:69-73).
// Candidates [c0, c0 + n_rows) of the launch's local index space -> tiles 0.. of the prep buffer.
// ---------------------------------------------------------------------------------------------------
struct PrepI8Args {
    ScoreArgs s;             // instance, cover and LP point (wfrag = DMMA blob: xoffset / gain are read from it)
    i64 c0, n_rows;
    uint8_t* tiles;
    int* status;
};

// FEAS: the launch also produces lam_min of every candidate (K3, same arithmetic as k_score_feas) from the point it has
// gathered anyway -- when a call wants both scores the eigenvalue work (FP64 pipe) runs under the image stores (HBM) of
// the other warps instead of in a launch of its own with its own unranking and gathers.
// Two CTAs per SM (<= 128 registers) is the measured optimum for both variants: left to itself ptxas takes 154 registers for
// FEAS (one CTA per SM, +10 ms per cfg4 step), and three or four CTAs per SM (80 / 64 registers) spill.
template <int D, int NS, bool FEAS = false>
__global__ void __launch_bounds__(256, 2) k_prep_i8(PrepI8Args pa)
{
    using C = NetCfg<D>;
    constexpr int T = D * (D + 1) / 2;
    const ScoreArgs& a = pa.s;
    __shared__ double sxo[C::NIN], sgn[C::NIN];
    if (threadIdx.x < C::NIN) {
        sxo[threadIdx.x] = __ldg(a.wfrag + C::OFF_XOFF + threadIdx.x);
        sgn[threadIdx.x] = __ldg(a.wfrag + C::OFF_GAIN + threadIdx.x);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const i64 warps_total = (i64)gridDim.x * (blockDim.x >> 5);
    const i64 gw = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const i64 G = ((pa.n_rows + I8_M - 1) / I8_M) * (I8_M / 32);       // whole tiles: the tail rows are zero filled
    const i64 g0 = gw * G / warps_total, g1 = (gw + 1) * G / warps_total;
    if (g0 >= g1) return;
    int c[D];
    const bool all_mode = (a.idx == nullptr);
    if (all_mode) {
        const i64 i0 = g0 * 32 + lane;
        if (i0 < pa.n_rows) lex_unrank<D>(a.n, (u64)(a.rank_begin + pa.c0 + i0), c);
        else {
#pragma unroll
            for (int t = 0; t < D; ++t) c[t] = t;
        }
    }
#pragma unroll 1
    for (i64 g = g0; g < g1; ++g) {
        const i64 r = g * 32 + lane;                 // row inside the chunk
        const bool valid = r < pa.n_rows;
        if (!all_mode) load_list_indices<D>(a.idx, pa.c0 + r, valid, c);
        double xs[D], Xs[T], Qs[T], p[C::NIN];
        gather_point<D>(a, c, xs, Xs);
        int k = 0;
        double mx = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = i; j < D; ++j) {
                const double v = __ldg(a.Q + tri_index(a.n, c[i], c[j]));
                Qs[k++] = v;
                mx = fmax(mx, fabs(v));
            }
        double max_elem = (double)D * mx;
        if (max_elem == 0.0) max_elem = 1.0;
        const bool tiny = max_elem < 1e-280;
        const double rme = fast_rcp(tiny ? 1.0 : max_elem);
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < T; ++q) {
            Qs[q] = tiny ? __ddiv_rn(Qs[q], max_elem) : div_by(Qs[q], max_elem, rme);
            s = __dadd_rn(s, __dmul_rn(Qs[q], Xs[q]));
        }
#pragma unroll
        for (int q = 0; q < D; ++q) p[q] = __dadd_rn(__dmul_rn(__dsub_rn(xs[q], sxo[q]), sgn[q]), -1.0);
#pragma unroll
        for (int q = 0; q < T; ++q) p[D + q] = __dadd_rn(__dmul_rn(__dsub_rn(Qs[q], sxo[D + q]), sgn[D + q]), -1.0);
        uint8_t* tile = pa.tiles + (r / I8_M) * (i64)I8Dig<NS>::TILE_BYTES;
        i8_store_row<C::NIN, NS>(tile, (int)(r % I8_M), p, __dmul_rn(-s, max_elem), max_elem, valid, pa.status);
        if constexpr (FEAS) {
            const double lam = lam_min_subset<D>(xs, Xs, 0);
            if (valid) a.lam[a.pos ? __ldg(a.pos + pa.c0 + r) : pa.c0 + r] = lam;
        }
        if (all_mode && g + 1 < g1) {
            if (!(valid && lex_advance<D>(a.n, c, 32))) {
#pragma unroll
                for (int t = 0; t < D; ++t) c[t] = t;
            }
        }
    }
}

// raw NN inputs in global memory (sdpcs_nn_eval): rows of NIN doubles -> tile images; base = 0, max_elem = 1
template <int D, int NS>
__global__ void __launch_bounds__(256) k_prep_i8_raw(const double* wfrag, const double* in, i64 m, uint8_t* tiles, int* status)
{
    using C = NetCfg<D>;
    const i64 rows = (m + I8_M - 1) / I8_M * I8_M;
    for (i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (i64)gridDim.x * blockDim.x) {
        const bool valid = r < m;
        double p[C::NIN];
#pragma unroll
        for (int q = 0; q < C::NIN; ++q) {
            const double v = valid ? in[r * C::NIN + q] : 0.0;
            p[q] = __dadd_rn(__dmul_rn(__dsub_rn(v, __ldg(wfrag + C::OFF_XOFF + q)), __ldg(wfrag + C::OFF_GAIN + q)), -1.0);
        }
        i8_store_row<C::NIN, NS>(tiles + (r / I8_M) * (i64)I8Dig<NS>::TILE_BYTES, (int)(r % I8_M), p, 0.0, 1.0, valid, status);
    }
}

// ---------------------------------------------------------------------------------------------------
// K1 + K2 inside the MLP kernel: one row (candidate) of a layer-0 tile image, written straight into the lane's A
// buffer in shared memory by the producer warp.  Same per-candidate arithmetic, operation for operation, as
// k_score_nn / k_prep_i8 (cut_select_qp.py:536-538, 573-575; neural_net_3D.m:69-73), but streamed four inputs at a
// time so that it fits the 64 registers of the helper warp group.
// img: [slice s][k chunk (2)][128 rows][16 B]; aux: base[128], max_elem[128].
// ---------------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ void i8_prep_row_smem(const ScoreArgs& a, const int (&c)[D], bool valid, int row, uint8_t* img, double* aux,
                                                 const double* __restrict__ sxo, const double* __restrict__ sgn, int* status)
{
    using C = NetCfg<D>;
    // pass A: max |Q_slice| (the values are re-read in pass B: an L2 hit costs less than 30 live registers here)
    double mx = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = i; j < D; ++j) mx = fmax(mx, fabs(__ldg(a.Q + tri_index(a.n, c[i], c[j]))));
    double max_elem = (double)D * mx;
    if (max_elem == 0.0) max_elem = 1.0;
    const bool tiny = max_elem < 1e-280;
    const double rme = fast_rcp(tiny ? 1.0 : max_elem);
    // zero the 32 bytes of every slice of this row, then overwrite four input digits at a time
    uint8_t* rowp = img + row * 16;
#pragma unroll
    for (int sl = 0; sl < I8_NS; ++sl) {
        *reinterpret_cast<uint4*>(rowp + sl * (I8_M * I8_K0)) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(rowp + sl * (I8_M * I8_K0) + I8_M * 16) = make_uint4(0, 0, 0, 0);
    }
    bool bad = false;
    double sdot = 0.0;
    unsigned long long u[4];
    auto put = [&](int k, double v) {            // network input k (compile-time after unrolling): mapminmax, quantise, store
        double pk = __dadd_rn(__dmul_rn(__dsub_rn(v, sxo[k]), sgn[k]), -1.0);
        if (!valid) pk = 0.0;
        bad |= !(fabs(pk) < 2.0);
        u[k & 3] = i8_quantize(pk, 1125899906842624.0 /* 2^50 */);
        if ((k & 3) == 3 || k == C::NIN - 1) {
            if ((k & 3) != 3) {
                for (int t = (k & 3) + 1; t < 4; ++t) u[t] = 0;
            }
            const int g = k >> 2;
            uint8_t* dst = rowp + (g >> 2) * (I8_M * 16) + 4 * (g & 3);
            *reinterpret_cast<uint32_t*>(dst + 6 * (I8_M * I8_K0)) = i8_pack4<0>(u[0], u[1], u[2], u[3]);
            *reinterpret_cast<uint32_t*>(dst + 5 * (I8_M * I8_K0)) = i8_pack4<1>(u[0], u[1], u[2], u[3]);
            *reinterpret_cast<uint32_t*>(dst + 4 * (I8_M * I8_K0)) = i8_pack4<2>(u[0], u[1], u[2], u[3]);
            *reinterpret_cast<uint32_t*>(dst + 3 * (I8_M * I8_K0)) = i8_pack4<3>(u[0], u[1], u[2], u[3]);
            *reinterpret_cast<uint32_t*>(dst + 2 * (I8_M * I8_K0)) = i8_pack4<4>(u[0], u[1], u[2], u[3]);
            *reinterpret_cast<uint32_t*>(dst + 1 * (I8_M * I8_K0)) = i8_pack4<5>(u[0], u[1], u[2], u[3]);
            *reinterpret_cast<uint32_t*>(dst + 0 * (I8_M * I8_K0)) = i8_pack4<6>(u[0], u[1], u[2], u[3]);
        }
    };
#pragma unroll
    for (int k = 0; k < D; ++k) put(k, __ldg(a.x + c[k]));
    {
        int k = D;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = i; j < D; ++j) {
                const int t = tri_index(a.n, c[i], c[j]);
                const double Qv = __ldg(a.Q + t), Xv = __ldg(a.X + t);
                const double Qt = tiny ? __ddiv_rn(Qv, max_elem) : div_by(Qv, max_elem, rme);
                sdot = __dadd_rn(sdot, __dmul_rn(Qt, Xv));          // left-to-right, no FMA (cut_select_qp.py:575)
                put(k, Qt);
                ++k;
            }
    }
    if (bad && valid) atomicCAS(status, 0, 2);
    aux[row] = valid ? __dmul_rn(-sdot, max_elem) : 0.0;
    aux[I8_M + row] = valid ? max_elem : 0.0;
}

// ---------------------------------------------------------------------------------------------------
// the MLP
// ---------------------------------------------------------------------------------------------------
// D > 0: candidates come from the instance (the producer warp builds the layer-0 image in shared memory);
// D = 0: raw network inputs, layer-0 images prepared by k_prep_i8_raw and loaded by TMA (sdpcs_nn_eval).
template <int NHID, int D, int NS = I8_NS, bool DBG = false>
__global__ void __launch_bounds__(I8_THREADS, 1) k_mlp_i8(MlpI8Args a)
{
    using L = I8Smem<NHID, NS>;
    using G = I8Dig<NS>;
    constexpr int STAGES = G::STAGES;
    static_assert(D == 0 || NS == I8_NS, "the fused producer writes 7-digit images");
    extern __shared__ __align__(1024) uint8_t sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double* par = reinterpret_cast<double*>(sm + L::OFF_PAR);
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + L::OFF_BAR + I8_NBAR * 8);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(sm + L::OFF_BAR + I8_NBAR * 8 + 4);
    const uint32_t bar0 = smem_u32(sm + L::OFF_BAR);
    const uint32_t B_FULL = bar0, B_EMPTY = bar0 + 8 * I8_SLOTS, B_A0 = bar0 + 16 * I8_SLOTS, B_FREE = B_A0 + 16,
                   B_ACT = B_A0 + 32, B_Y = B_A0 + 48;

    const i64 t0 = a.n_tiles * blockIdx.x / gridDim.x, t1 = a.n_tiles * (blockIdx.x + 1) / gridDim.x;

    // weights + parameters -> shared memory (generic proxy), barriers, TMEM
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.wimg);
        uint4* dst = reinterpret_cast<uint4*>(sm);
        for (int i = tid; i < L::W_TOTAL / 16; i += I8_THREADS) dst[i] = __ldg(src + i);
        const double* ps = reinterpret_cast<const double*>(a.wimg + L::W_TOTAL);
        for (int i = tid; i < L::PAR; i += I8_THREADS) par[i] = __ldg(ps + i);
        if constexpr (D > 0) {
            double* xg = reinterpret_cast<double*>(sm + L::OFF_XG);
            if (tid < NetCfg<D>::NIN) {
                xg[tid] = __ldg(a.s.wfrag + NetCfg<D>::OFF_XOFF + tid);
                xg[24 + tid] = __ldg(a.s.wfrag + NetCfg<D>::OFF_GAIN + tid);
            }
        }
    }
    if (tid == 0) {
        for (int i = 0; i < I8_SLOTS; ++i) {
            mbar_init(B_FULL + 8 * i, 1);
            mbar_init(B_EMPTY + 8 * i, I8_EPI_WARPS);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(B_A0 + 8 * i, D > 0 ? I8_M / 32 : 2);   // fused: one arrival per 32-row pass; staged: expect_tx + tail
            mbar_init(B_FREE + 8 * i, 1);
            mbar_init(B_ACT + 8 * i, I8_EPI_WARPS);
            mbar_init(B_Y + 8 * i, I8_EPI_WARPS);
        }
        *abort_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == I8_EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const i64 npair = (t1 - t0 + 1) / 2;
    // register re-partitioning (per warp group): the helper warps keep 24 (64 when they gather and slice the inputs
    // themselves) registers, the epilogue warps take 112 (104): 20 x 96 >= 4 x 24 + 16 x 112 per thread -- an increase
    // beyond the pool released by the helpers would block for ever; with 227 KB of shared memory there is no L1 left for
    // spills, every spilled value is an L2 round trip (the instruction sits at the head of each role's branch so that
    // ptxas allocates each role within its own budget)
    if (warp > I8_EPI_WARPS + 1 && D == 0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");      // idle members of the helper warp group (TMA mode)
    } else if (warp >= I8_EPI_WARPS + 1) {
        if constexpr (D > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        // ===== producer: the layer-0 digit image of the next tile of each lane =====
        bool ok = true;
        if constexpr (D > 0) {
            // K1 + K2 fused: unrank (once, then advance by 32 per pass), gather x / X / Q through L2, mapminmax, slice
            const double* sxo = reinterpret_cast<const double*>(sm + L::OFF_XG);
            const double* sgn = sxo + 24;
            const bool all_mode = (a.s.idx == nullptr);
            // three producer warps: 32-row pass g = 4 (tile - t0) + pass goes to warp g mod 3
            constexpr int NPROD = 3;
            const int pw = warp - (I8_EPI_WARPS + 1);
            int c[D > 0 ? D : 1];
#pragma unroll
            for (int t = 0; t < D; ++t) c[t] = t;
            bool live = all_mode && t0 * I8_M + pw * 32 + lane < a.n_rows;
            if (live) lex_unrank<D>(a.s.n, (u64)(a.s.rank_begin + a.out_base + t0 * I8_M + pw * 32 + lane), c);
            const i64 npass = (t1 - t0) * (I8_M / 32);
            for (i64 g = pw; g < npass && ok; g += NPROD) {
                const i64 tile = t0 + (g >> 2);
                const int pass = (int)(g & 3);
                const int ln = (int)((tile - t0) & 1);
                const uint32_t cnt = (uint32_t)((tile - t0) >> 1);
                I8_STAMP((uint32_t)(g / NPROD), 0);
                ok = mbar_wait_relaxed(B_FREE + 8 * ln, (cnt & 1) ^ 1, abort_flag, a.status);
                ok = __all_sync(0xffffffffu, ok);
                if (!ok) break;
                I8_STAMP((uint32_t)(g / NPROD), 1);
                uint8_t* img = sm + L::OFF_A + ln * G::AH_BYTES;
                double* aux = reinterpret_cast<double*>(sm + L::OFF_AUX + (2 * ln + (cnt & 1)) * I8_AUX_BYTES);
                const i64 r = tile * I8_M + pass * 32 + lane;
                const bool valid = r < a.n_rows;
                if (!all_mode) load_list_indices<D>(a.s.idx, a.out_base + r, valid, c);
                i8_prep_row_smem<D>(a.s, c, valid, pass * 32 + lane, img, aux, sxo, sgn, a.status);
                if (all_mode && !(valid && lex_advance<D>(a.s.n, c, 32 * NPROD))) {
#pragma unroll
                    for (int t = 0; t < D; ++t) c[t] = t;
                }
                I8_STAMP((uint32_t)(g / NPROD), 2);
                fence_async_smem();          // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(B_A0 + 8 * ln);
                I8_STAMP((uint32_t)(g / NPROD), 3);
            }
        } else {
            // TMA (whole warp walks the tile list, one elected lane issues the bulk copies)
            for (i64 tile = t0; tile < t1 && ok; ++tile) {
                const int ln = (int)((tile - t0) & 1);
                const uint32_t cnt = (uint32_t)((tile - t0) >> 1);
                ok = mbar_wait_relaxed(B_FREE + 8 * ln, (cnt & 1) ^ 1, abort_flag, a.status);
                ok = __all_sync(0xffffffffu, ok);
                if (!ok) break;
                const uint8_t* src = a.tiles + tile * (i64)G::TILE_BYTES;
                uint8_t* img = sm + L::OFF_A + ln * G::AH_BYTES;
                if (elect_one()) {
                    mbar_expect_tx(B_A0 + 8 * ln, G::G_MAIN + I8_AUX_BYTES);
#pragma unroll
                    for (int sl = 0; sl < NS; ++sl)      // first k chunk of every slice
                        tma_load_1d(smem_u32(img + sl * (I8_M * I8_K0)), src + sl * (I8_M * 16), I8_M * 16, B_A0 + 8 * ln);
                    tma_load_1d(smem_u32(sm + L::OFF_AUX + (2 * ln + (cnt & 1)) * I8_AUX_BYTES), src + G::G_MAIN + G::G_TAIL, I8_AUX_BYTES,
                                B_A0 + 8 * ln);
                }
                __syncwarp();
                // second k chunk: one word per row and slice from the staged tile, 12 zero bytes of padding
                const uint32_t* tail = reinterpret_cast<const uint32_t*>(src + G::G_MAIN);
                for (int t = lane; t < NS * I8_M; t += 32)
                    *reinterpret_cast<uint4*>(img + (t >> 7) * (I8_M * I8_K0) + I8_M * 16 + (t & (I8_M - 1)) * 16) = make_uint4(__ldg(tail + t), 0, 0, 0);
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(B_A0 + 8 * ln);
            }
        }
    } else if (warp == I8_EPI_WARPS) {
        if constexpr (D > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        // ===== MMA issuer (whole warp walks the schedule and waits, one elected lane issues) =====
        // Accumulator stage st = step mod STAGES: diagonal d lives in TMEM columns [NS 64 st + 64 d, + 64).  A stage is
        // handed to the epilogue with one commit per step and handed back once every epilogue warp has read it.  With one
        // stage (NS = 7) the MMAs of step n+1 (other tile) start when step n has been read and run while the epilogue
        // warps finalize it; with two stages (NS = 4) they never wait for the epilogue of the step before.
        uint32_t step = 0, actc[2] = {0, 0};
        bool ok = true;
        for (i64 p = 0; p < npair && ok; ++p)
            for (int l = 0; l < NHID && ok; ++l)
                for (int ln = 0; ln < 2 && ok; ++ln) {
                    if (t0 + 2 * p + ln >= t1) continue;
                    const uint32_t st = step % STAGES, use = step / STAGES;
                    I8_STAMP(step, 0);
                    if (l == 0) ok = mbar_wait(B_A0 + 8 * ln, (uint32_t)p & 1, abort_flag, a.status);
                    else if constexpr (D == 0) nb_sync(I8_NB_ACT + ln);
                    else {
                        ok = mbar_wait(B_ACT + 8 * ln, actc[ln] & 1, abort_flag, a.status);
                        actc[ln]++;
                    }
                    I8_STAMP(step, 1);
                    if constexpr (D == 0) {
                        if (use > 0) nb_sync(I8_NB_EMPTY + st);          // the first use of a stage has nothing to wait for
                    } else if (ok) ok = mbar_wait(B_EMPTY + 8 * st, (use & 1) ^ 1, abort_flag, a.status);
                    ok = __all_sync(0xffffffffu, ok);
                    if (!ok) break;
                    I8_STAMP(step, 2);
                    tc_fence_after();
                    if (elect_one()) {
                        // A image: [slice s][k chunk][128 rows][16 B]  -> LBO (between the two k chunks of one MMA) 2048
                        // W image: [k chunk][slice t][64 rows][16 B]   -> LBO NS * 1024; the slices t = ta..tb of one k
                        //          chunk are 64 (tb - ta + 1) consecutive rows: ONE MMA of N = 64 (tb - ta + 1) forms
                        //          the products of A slice s with all of them and lands them in the adjacent
                        //          accumulators of the diagonals s + ta .. s + tb.  NS = 7: 10 MMAs per k step instead
                        //          of 28, and the A operand is read 10 times instead of 28; NS = 4: 4 MMAs for 10 pairs.
                        const uint32_t a_base = smem_u32(sm + L::OFF_A + ln * G::AH_BYTES);
                        const uint32_t w_base = smem_u32(sm + (l == 0 ? 0 : G::W0_BYTES + (l - 1) * G::WH_BYTES));
                        const uint32_t a_slice = (l == 0) ? I8_M * I8_K0 : I8_M * 64;
                        const uint32_t acc0 = tmem + st * (NS * I8_N);
                        const int ksteps = (l == 0) ? 1 : 2;
                        for (int kk = 0; kk < ksteps; ++kk) {
                            const uint64_t bd0 = umma_desc(w_base + kk * 2 * (NS * I8_N * 16), NS * I8_N * 16, 128);
#pragma unroll
                            for (int sd = 0; sd < NS; ++sd) {
                                const uint64_t ad = umma_desc(a_base + sd * a_slice + kk * 2 * (I8_M * 16), I8_M * 16, 128);
                                constexpr uint32_t SGN = 1u << 7;
                                const int nt = NS - sd;                    // weight slices t = 0 .. NS - 1 - sd
                                const int n0 = nt < 4 ? nt : 4;
                                umma_i8(acc0 + sd * I8_N, ad, bd0, i8_idesc(64 * n0) | (sd ? 0u : SGN), (sd | kk) > 0);
                                if (nt > 4)
                                    umma_i8(acc0 + (sd + 4) * I8_N, ad, bd0 + (uint64_t)((4 * I8_N * 16) >> 4),
                                            i8_idesc(64 * (nt - 4)) | (sd ? 0u : SGN), (sd | kk) > 0);
                            }
                        }
                        umma_commit(B_FULL + 8 * st);
                        if (l == NHID - 1) umma_commit(B_FREE + 8 * ln);
                    }
                    __syncwarp();
                    I8_STAMP(step, 3);
                    ++step;
                }
    } else {
        if constexpr (D > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
        // ===== epilogue warps =====
        const int q = warp & 3, cq = warp >> 2;
        const int row = q * 32 + lane;
        const double* tab = par + L::P_TAB;
        uint32_t step = 0;
        bool ok = true;
        for (i64 p = 0; p < npair && ok; ++p)
            for (int l = 0; l < NHID && ok; ++l)
                for (int ln = 0; ln < 2 && ok; ++ln) {
                    const i64 tile = t0 + 2 * p + ln;
                    if (tile >= t1) continue;
                    const uint32_t st = step % STAGES, use = step / STAGES;
                    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + st * (NS * I8_N) + cq * 16;
                    I8_STAMP(step, 0);
                    if (l == 0)   // acquire the TMA-written aux block of this tile (read in the last layer)
                        ok = mbar_wait(B_A0 + 8 * ln, (uint32_t)p & 1, abort_flag, a.status);
                    if (ok) ok = mbar_wait(B_FULL + 8 * st, use & 1, abort_flag, a.status);
                    ok = __all_sync(0xffffffffu, ok);
                    if (!ok) break;
                    I8_STAMP(step, 1);
                    tc_fence_after();
                    const double2* csbs = reinterpret_cast<const double2*>(par + L::P_CS) + l * 64 + cq * 16;   // (cs, bs) pairs
                    const double* wout = par + L::P_WOUT + cq * 16;
                    // Phase 1: read the NS diagonals (two batches of TMEM loads, 8 neurons each), recombine them exactly in int64 and
                    // reduce to the FP64 pre-activations.  The accumulator stage is handed back to the MMA issuer as soon as the
                    // second batch is in registers.  Phase 2: tansig, four neurons at a time, as a software pipeline: the tansig of
                    // group g is woven with the recombination of the last four outputs (g = 0) or the re-slicing of group g - 1
                    // (tansig4_woven).  From the hand-back to the stores the step is ONE basic block (no branch on the layer
                    // inside: hidden layers and the output layer are two instances of the body).
                    uint8_t* abuf = sm + L::OFF_A + ln * G::AH_BYTES + cq * (I8_M * 16) + row * 16;
                    constexpr bool WIDE_STORE = (NS <= 4);   // 7 digits: three groups of pending words would spill (measured: +6 % step time)
                    double part = 0.0;
                    // NS = 7 (one accumulator stage): the MMAs of the other tile start at the hand-back, and while they run the FP64
                    // pipe and the shared-memory loads of the SM all but stand still (tools/tc_fp64_overlap.cu).  The step is therefore
                    // ordered FP64 | hand-back | integer | FP64: the first eight outputs go through tansig BEFORE the hand-back,
                    // the stretch after it holds integer work only (slicing and stores of the first eight activations, integer
                    // recombination of the last eight accumulators), the FP64 work of the last eight outputs follows.
                    auto body7 = [&](auto last_tag) {
                        constexpr bool LAST = decltype(last_tag)::value;
                        double z[8], act0[4], act1[4];
                        uint32_t v[NS][8];
#pragma unroll
                        for (int dg = 0; dg < NS; ++dg) tmem_ld8_async(tbase + dg * I8_N, v[dg]);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 4; ++j) z[j] = i8_recombine<NS>(v, j, csbs[j]);
                        {
                            double zz[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) zz[i] = z[i];
                            tansig4_woven(zz, act0, tab, I8RecombineSideAt<NS, 4, 4, 8>(v, csbs + 4, z));
                        }
                        uint32_t keep[NS], w[NS];
                        double qd1[4];
                        if constexpr (!LAST) {
                            double zz[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) zz[i] = z[4 + i];
                            tansig4_woven(zz, act1, tab, I8SliceSide<NS>(act0, G::SCALE_H, keep));
#pragma unroll
                            for (int i = 0; i < 4; ++i) qd1[i] = fma(act1[i], G::SCALE_H, I8_MAGIC52);
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) part = fma(wout[i], act0[i], part);
                            double zz[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) zz[i] = z[4 + i];
                            tansig4_woven(zz, act1, tab, I8NoSide());
#pragma unroll
                            for (int i = 0; i < 4; ++i) part = fma(wout[4 + i], act1[i], part);
                        }
#pragma unroll
                        for (int dg = 0; dg < NS; ++dg) tmem_ld8_async(tbase + dg * I8_N + 8, v[dg]);
                        tmem_wait_ld();
                        tc_fence_before();
                        __syncwarp();
                        if constexpr (D == 0) nb_arrive(I8_NB_EMPTY + st);
                        else mbar_arrive_lane0(B_EMPTY + 8 * st, lane);
                        I8_STAMP(step, 2);
                        // ---- integer only: the tensor core is busy with the other tile
                        if constexpr (!LAST) {
                            i8_slice_from_qd<NS>(qd1, w);
#pragma unroll
                            for (int b = 0; b < NS; ++b) *reinterpret_cast<uint2*>(abuf + b * (I8_M * 64)) = make_uint2(keep[b], w[b]);
                        }
                        long long aH[8], aL[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) i8_recombine_int7(v, j, aH[j], aL[j]);
                        // ---- FP64 again (program order: the conversions need the integer words above)
#pragma unroll
                        for (int j = 0; j < 8; ++j) z[j] = i8_recombine_fin7(aH[j], aL[j], csbs[8 + j]);
                        {
                            double zz[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) zz[i] = z[i];
                            tansig4_woven(zz, act0, tab, I8NoSide());
                        }
                        if constexpr (!LAST) {
                            double zz[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) zz[i] = z[4 + i];
                            tansig4_woven(zz, act1, tab, I8SliceSide<NS>(act0, G::SCALE_H, keep));
                            unsigned long long u[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) u[i] = i8_quantize(act1[i], G::SCALE_H);
                            i8_pack_slices<NS>(u[0], u[1], u[2], u[3], w);
#pragma unroll
                            for (int b = 0; b < NS; ++b) *reinterpret_cast<uint2*>(abuf + b * (I8_M * 64) + 8) = make_uint2(keep[b], w[b]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) part = fma(wout[8 + i], act0[i], part);
                            double zz[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) zz[i] = z[4 + i];
                            tansig4_woven(zz, act1, tab, I8NoSide());
#pragma unroll
                            for (int i = 0; i < 4; ++i) part = fma(wout[12 + i], act1[i], part);
                        }
                    };
                    // NS = 4 (two accumulator stages: the MMAs are not tied to the hand-back) and the debug instance (which dumps all
                    // sixteen pre-activations of a layer): the order of round 2's first half -- same arithmetic
                    auto body = [&](auto last_tag) {
                        constexpr bool LAST = decltype(last_tag)::value;
                        double z[16];
                        uint32_t v[NS][8];
#pragma unroll
                        for (int dg = 0; dg < NS; ++dg) tmem_ld8_async(tbase + dg * I8_N, v[dg]);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 8; ++j) z[j] = i8_recombine<NS>(v, j, csbs[j]);
#pragma unroll
                        for (int dg = 0; dg < NS; ++dg) tmem_ld8_async(tbase + dg * I8_N + 8, v[dg]);
                        tmem_wait_ld();
                        tc_fence_before();
                        __syncwarp();
                        if constexpr (D == 0) nb_arrive(I8_NB_EMPTY + st);
                        else mbar_arrive_lane0(B_EMPTY + 8 * st, lane);
                        I8_STAMP(step, 2);
#pragma unroll
                        for (int j = 0; j < 4; ++j) z[8 + j] = i8_recombine<NS>(v, j, csbs[8 + j]);
                        double act[2][4];
                        {
                            double zz[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) zz[i] = z[i];
                            tansig4_woven(zz, act[0], tab, I8RecombineSide<NS>(v, csbs, z));
                        }
                        if constexpr (!LAST) {
                            // The warp owns 16 neurons = one 16-byte row of the K-major core matrices.  4 digits: ONE 16-byte store per
                            // slice (no bank conflict), the digit words wait in the registers the z values leave; 7 digits: 8-byte
                            // stores after every second group
                            uint32_t keep[WIDE_STORE ? 3 : 1][NS], w[NS];
#pragma unroll
                            for (int g = 1; g <= 4; ++g) {
                                if (g < 4) {
                                    double zz[4];
#pragma unroll
                                    for (int i = 0; i < 4; ++i) zz[i] = z[g * 4 + i];
                                    tansig4_woven(zz, act[g & 1], tab, I8SliceSide<NS>(act[(g - 1) & 1], G::SCALE_H, w));
                                } else {
                                    unsigned long long u[4];
#pragma unroll
                                    for (int i = 0; i < 4; ++i) u[i] = i8_quantize(act[1][i], G::SCALE_H);
                                    i8_pack_slices<NS>(u[0], u[1], u[2], u[3], w);
                                }
                                const int gp = g - 1;                                  // w: digit words of group gp, w[s] = slice s (0 = most significant)
                                if constexpr (WIDE_STORE) {
                                    if (gp < 3) {
#pragma unroll
                                        for (int b = 0; b < NS; ++b) keep[gp][b] = w[b];
                                    } else {
#pragma unroll
                                        for (int b = 0; b < NS; ++b)
                                            *reinterpret_cast<uint4*>(abuf + b * (I8_M * 64)) = make_uint4(keep[0][b], keep[1][b], keep[2][b], w[b]);
                                    }
                                } else {
                                    if ((gp & 1) == 0) {
#pragma unroll
                                        for (int b = 0; b < NS; ++b) keep[0][b] = w[b];
                                    } else {
#pragma unroll
                                        for (int b = 0; b < NS; ++b)
                                            *reinterpret_cast<uint2*>(abuf + b * (I8_M * 64) + (gp >> 1) * 8) = make_uint2(keep[0][b], w[b]);
                                    }
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i) part = fma(wout[i], act[0][i], part);
#pragma unroll
                            for (int g = 1; g < 4; ++g) {
                                double zz[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) zz[i] = z[g * 4 + i];
                                tansig4_woven(zz, act[1], tab, I8NoSide());
#pragma unroll
                                for (int i = 0; i < 4; ++i) part = fma(wout[g * 4 + i], act[1][i], part);
                            }
                        }
                        if constexpr (DBG) {
                            if (a.dbg_z && l == a.dbg_layer) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) a.dbg_z[(tile * I8_M + row) * 64 + cq * 16 + j] = z[j];
                            }
                        }
                    };
                    if constexpr (NS == 7 && !DBG) {
                        if (l < NHID - 1) body7(std::false_type{});
                        else body7(std::true_type{});
                    } else
                    {
                        if (l < NHID - 1) body(std::false_type{});
                        else body(std::true_type{});
                    }
                    if (l < NHID - 1) {
                        fence_async_smem();
                        __syncwarp();
                        if constexpr (D == 0) nb_arrive(I8_NB_ACT + ln);
                        else if (lane == 0) mbar_arrive(B_ACT + 8 * ln);
                    } else {
                        // linear output layer (neural_net_3D.m:61-65, 81-85): partial dot products per column quarter
                        double* yp = (STAGES == 2) ? reinterpret_cast<double*>(sm + L::OFF_Y) + (2 * ln + ((int)p & 1)) * (4 * I8_M)
                                                   : reinterpret_cast<double*>(sm + L::OFF_A + ln * G::AH_BYTES + G::A0_BYTES);
                        yp[cq * I8_M + row] = part;
                        __syncwarp();
                        if (lane == 0) mbar_arrive(B_Y + 8 * ln);
                        if (cq == ((2 * (int)p + ln) & 3)) {   // the finalising role rotates: no warp falls systematically behind
                            ok = mbar_wait(B_Y + 8 * ln, (uint32_t)p & 1, abort_flag, a.status);
                            ok = __all_sync(0xffffffffu, ok);
                            if (!ok) break;
                            const double y = ((yp[row] + yp[I8_M + row]) + yp[2 * I8_M + row]) + yp[3 * I8_M + row];
                            const double* aux = reinterpret_cast<const double*>(sm + L::OFF_AUX + (2 * ln + ((int)p & 1)) * I8_AUX_BYTES);
                            const double bout = par[L::P_MISC], y_gain = par[L::P_MISC + 1], y_xoff = par[L::P_MISC + 2];
                            const double nn = ((bout + y) - -1.0) / y_gain + y_xoff;
                            const double obj = __dadd_rn(aux[row], __dmul_rn(nn, aux[I8_M + row]));
                            const i64 r = tile * I8_M + row;
                            if (r < a.n_rows) {
                                const i64 gi = a.out_base + r;
                                a.obj[a.pos ? __ldg(a.pos + gi) : gi] = obj;
                            }
                        }
                    }
                    I8_STAMP(step, 3);
                    ++step;
                }
    }
    // a protocol time-out (status 1) cannot be waited out when the issuer is parked in a named barrier: end the grid instead
    if (D == 0 && *abort_flag) asm volatile("trap;");
    tc_fence_before();
    __syncthreads();
    if (warp == I8_EPI_WARPS)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace sdpcs
