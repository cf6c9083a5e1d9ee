// Per-sub-partition issue rates of the instruction forms the tcgen05 MLP epilogue uses (sm_100a), 4 warps per SMSP, ILP 4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_rates tools/pipe_rates.cu && tools/pipe_rates
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2000
#define REP 16
template <int OP>
__global__ void k(double* out, long long* clk, double a, double b, int ia, int ib)
{
    double x[4];
    long long y[4];
    unsigned u[4];
    for (int i = 0; i < 4; ++i) { x[i] = threadIdx.x * 1e-3 + i; y[i] = threadIdx.x + i; u[i] = threadIdx.x * 77 + i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < REP; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (OP == 0) x[i] = fma(x[i], a, b);                                   // DFMA reg, reg, reg
                if (OP == 1) x[i] = fma(x[i], 0.999, 1e-3);                            // DFMA with constants
                if (OP == 2) x[i] = x[i] + 26388279066624.0;                           // DADD immediate
                if (OP == 3) x[i] = x[i] * a;                                          // DMUL
                if (OP == 4) y[i] = (long long)(int)u[i] * 65536ll + y[i];             // IMAD.WIDE accumulate
                if (OP == 5) u[i] = __byte_perm(u[i], ia, 0x5410 + r);                 // PRMT
                if (OP == 6) u[i] = (u[i] ^ ia) & (ib + r);                            // LOP3
                if (OP == 7) u[i] = u[i] * ia + ib;                                    // IMAD
                if (OP == 8) { double t; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(x[i])); x[i] = t; }   // MUFU.RCP64H
                if (OP == 9) y[i] = y[i] + (long long)ia;                              // 64-bit add (IADD3 + IADD3.X)
                if (OP == 10) u[i] = min(u[i] | 0x80000000u, 0xC08F4000u + r);         // LOP3 + VIMNMX
                if (OP == 11) y[i] = y[i] << 3;                                        // 64-bit shift
            }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 4; ++i) s += x[i] + (double)y[i] + u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}
template <int OP>
void run(const char* name)
{
    double* out; long long* clk;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&clk, 8);
    for (int warps : {4, 16}) {
        k<OP><<<1, warps * 32>>>(out, clk, 0.999, 1e-3, 3, 5);
        k<OP><<<1, warps * 32>>>(out, clk, 0.999, 1e-3, 3, 5);
        long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
        double n = (double)ITERS * REP * 4;
        printf("%-34s %2d warps/SMSP: %.3f source ops/clk/SMSP\n", name, warps / 4, n * (warps / 4) / c);
    }
    cudaFree(out); cudaFree(clk);
}
int main()
{
    run<0>("DFMA reg"); run<1>("DFMA const"); run<2>("DADD imm"); run<3>("DMUL"); run<4>("IMAD.WIDE acc (int64)");
    run<5>("PRMT"); run<6>("LOP3 x2"); run<7>("IMAD"); run<8>("MUFU.RCP64H"); run<9>("int64 add"); run<10>("LOP3+VIMNMX");
    run<11>("int64 shl");
    return 0;
}
