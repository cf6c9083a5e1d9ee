// Micro-benchmark: does the FP64 pipe (DFMA) slow down while tcgen05.mma kind::i8 runs on the same SM (and vice versa)?
// One CTA per SM: warp 16 issues back-to-back M128 x N256 x K32 int8 MMAs on (zeroed) shared-memory operands, warps
// 0..NW-1 run register-only arithmetic loops: FP64 (DFMA chains), integer (IMAD / LOP3 chains) or a shared-memory LDS loop.
// Every role reports its own elapsed cycles, alone and together.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sdpcutsel-via-nn_b200/csrc -o tools/tc_fp64_overlap tools/tc_fp64_overlap.cu
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "score_kernels.cuh"
#include "mlp_i8_kernels.cuh"
using namespace sdpcs;

constexpr int SMEM = 96 * 1024;

__global__ void __launch_bounds__(544, 1) k_overlap(int n_mma, int math_mode, int n_math, int nw_math, long long* out, double* sink, int gap = 0)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < SMEM / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        mbar_init(b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const long long t0 = clock64();
    long long t1 = t0;
    if (warp == 16) {
        if (n_mma > 0) {
            uint32_t phase = 0;
            for (int it = 0; it < n_mma; ++it) {
                if (elect_one()) {
                    const uint64_t ad = umma_desc(smem_u32(sm), I8_M * 16, 128);                 // A: 128 rows x 32 B
                    const uint64_t bd = umma_desc(smem_u32(sm) + 16384, 256 * 16, 128);         // B: 256 rows x 32 B
#pragma unroll
                    for (int j = 0; j < 14; ++j) umma_i8(tmem + (j & 1) * 256, ad, bd, i8_idesc(256), 1u);   // = one hidden-layer step of k_mlp_i8 (1,792 columns x 2 k steps)
                    umma_commit(b);
                }
                __syncwarp();
                while (!mbar_try_hint(b, phase, 20000u)) { if (clock64() - t0 > 4000000000ll) break; }     // never hang the box
                phase ^= 1;
                if (gap > 0) { const long long tg = clock64(); while (clock64() - tg < gap) {} }              // idle tensor core between the steps
            }
            t1 = clock64();
        }
    } else if (warp < nw_math) {
        if (math_mode == 1) {            // FP64: 8 independent DFMA chains
            double x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = 1.0 + 1e-9 * (lane + i);
            const double a = 1.0000001, c = 1e-12;
            for (int it = 0; it < n_math; ++it) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, c);
            }
            double s = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += x[i];
            if (s == 12345.678) sink[threadIdx.x] = s;
        } else if (math_mode == 2) {     // integer: 8 independent IMAD chains
            unsigned x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = lane + i;
            for (int it = 0; it < n_math; ++it) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = x[i] * 1664525u + 1013904223u;
            }
            unsigned s = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) s ^= x[i];
            if (s == 0x12345u) sink[threadIdx.x] = s;
        } else if (math_mode == 3) {     // shared-memory loads: dependent 8-byte look-ups (random banks), 4 chains
            const double* T = reinterpret_cast<const double*>(sm + 65536);
            unsigned j[4] = {(unsigned)lane * 7u, (unsigned)lane * 13u + 1, (unsigned)lane * 29u + 2, (unsigned)lane * 31u + 3};
            double acc = 0;
            for (int it = 0; it < n_math; ++it) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double v = T[j[i] & 255];
                        j[i] = j[i] * 5u + (unsigned)__double2loint(v) + 1u;
                        acc += v;
                    }
            }
            if (acc == 12345.678) sink[threadIdx.x] = acc;
        } else if (math_mode == 4) {     // FP32 FFMA chains (control)
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = 1.0f + 1e-3f * (lane + i);
            for (int it = 0; it < n_math; ++it) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], 1.0001f, 1e-6f);
            }
            float s = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += x[i];
            if (s == 12345.678f) sink[threadIdx.x] = s;
        }
        t1 = clock64();
    }
    if (lane == 0) out[blockIdx.x * 17 + warp] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d_out; double* d_sink;
    cudaMalloc(&d_out, sizeof(long long) * sms * 17);
    cudaMalloc(&d_sink, sizeof(double) * 1024);
    cudaFuncSetAttribute(k_overlap, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    std::vector<long long> h(sms * 17);
    auto run = [&](int n_mma, int mode, int n_math, int nw, double& t_mma, double& t_math) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(d_out, 0, sizeof(long long) * sms * 17);
            k_overlap<<<sms, 544, SMEM>>>(n_mma, mode, n_math, nw, d_out, d_sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
        }
        cudaMemcpy(h.data(), d_out, sizeof(long long) * sms * 17, cudaMemcpyDeviceToHost);
        double a = 0, m = 0;
        for (int s = 0; s < sms; ++s) {
            a += (double)h[s * 17 + 16];
            double mx = 0;
            for (int w = 0; w < nw; ++w) mx = std::max(mx, (double)h[s * 17 + w]);
            m += mx;
        }
        t_mma = a / sms; t_math = m / sms;
    };
    const int n_mma = 400;                 // 400 "steps" of 14 N=256 MMAs: 1,792 cycles each at full rate
    const char* names[5] = {"", "FP64 DFMA", "integer IMAD", "shared-memory LDS.64 (random)", "FP32 FFMA"};
    double tm0, dummy;
    run(n_mma, 0, 0, 0, tm0, dummy);
    printf("MMA alone: %.0f cycles per step of 14 x (M128 N256 K32) int8 MMAs (floor 1,792)\n", tm0 / n_mma);
    for (int nw : {16, 4}) {
        for (int mode = 1; mode <= 4; ++mode) {
            // size the math loop to last about as long as the MMA stream
            int n_math = 2000;
            double t_alone, t_mma_with, t_with;
            run(0, mode, n_math, nw, dummy, t_alone);
            n_math = (int)(n_math * (tm0 / t_alone));
            run(0, mode, n_math, nw, dummy, t_alone);
            run(n_mma, mode, n_math, nw, t_mma_with, t_with);
            printf("%2d warps of %-30s: alone %9.0f cycles, with MMAs %9.0f (x%.3f) | MMA stream: alone %9.0f, with the math %9.0f (x%.3f)\n", nw, names[mode],
                   t_alone, t_with, t_with / t_alone, tm0, t_mma_with, t_mma_with / tm0);
        }
    }
    // duty cycle: the same MMA steps with idle gaps between them; FP64 work sized to the MMA stream alone (no gaps)
    printf("\nFP64 DFMA (16 warps) next to MMA steps separated by idle gaps (one step = 14 MMAs, ~2,000 cycles):\n");
    {
        int n_math = 2000;
        double t_alone, d2;
        run(0, 1, n_math, 16, d2, t_alone);
        n_math = (int)(n_math * (tm0 / t_alone));
        run(0, 1, n_math, 16, d2, t_alone);
        for (int gap : {0, 1000, 2000, 4000, 8000}) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaMemset(d_out, 0, sizeof(long long) * sms * 17);
                k_overlap<<<sms, 544, SMEM>>>(n_mma, 1, n_math, 16, d_out, d_sink, gap);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h.data(), d_out, sizeof(long long) * sms * 17, cudaMemcpyDeviceToHost);
            double a = 0, m = 0;
            for (int s = 0; s < sms; ++s) {
                a += (double)h[s * 17 + 16];
                double mx = 0;
                for (int w = 0; w < 16; ++w) mx = std::max(mx, (double)h[s * 17 + w]);
                m += mx;
            }
            a /= sms; m /= sms;
            const double duty = tm0 / a;      // share of the time the tensor core is busy
            printf("  gap %5d cycles: MMA stream %9.0f cycles (tensor core busy %.0f %% of it), FP64 work %9.0f cycles (alone %9.0f): FP64 throughput while both run = %.0f %% of alone\n",
                   gap, a, 100 * duty, m, t_alone, m <= a ? 100.0 * t_alone / m : 100.0 * (t_alone - (m - a)) / a);
        }
    }
    return 0;
}
