// Dependent-issue latency and per-SMSP throughput of the FP64 pipe on sm_100a (one warp, ILP 1..8; 1..4 warps per SMSP).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_latency tools/fp64_latency.cu && tools/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, long long* clk, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int ILP>
void run(int warps)
{
    double* out; long long* clk;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&clk, 8);
    const int iters = 2000;
    k<ILP><<<1, warps * 32>>>(out, clk, iters, 0.999, 1e-3);
    k<ILP><<<1, warps * 32>>>(out, clk, iters, 0.999, 1e-3);
    long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    double n = (double)iters * 16 * ILP;
    printf("warps/CTA %2d (per SMSP %d) ILP %d: %.2f clk per DFMA per warp, %.3f DFMA/clk/SMSP\n", warps, (warps + 3) / 4, ILP, c / n,
           n * ((warps + 3) / 4) / c);
    cudaFree(out); cudaFree(clk);
}
int main()
{
    for (int w : {1, 4, 8, 16}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
