// FP64 peak micro-benchmarks for B200 (sm_100a): DFMA pipe, DMMA.8x8x4 pipe, both at once,
// and broadcast LDS feeding DFMA. Output: one JSON object on stdout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peaks fp64_peaks.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               :"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}

template<int CH>
__global__ void __launch_bounds__(256) k_dfma(double* out, const double* in, int iters){
  double x = in[threadIdx.x & 31], y = in[32 + (threadIdx.x & 31)];
  double acc[CH];
  #pragma unroll
  for(int j=0;j<CH;j++) acc[j] = x + j;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<CH;j++) acc[j] = fma(acc[j], x, y);
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<CH;j++) s+=acc[j];
  if(s==123.456) out[threadIdx.x]=s;
}

template<int CH>
__global__ void __launch_bounds__(256) k_dmma(double* out, const double* in, int iters){
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c0[CH], c1[CH];
  #pragma unroll
  for(int j=0;j<CH;j++){ c0[j]=j; c1[j]=-j; }
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<CH;j++) dmma884(c0[j], c1[j], a, b);
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<CH;j++) s+=c0[j]+c1[j];
  if(s==123.456) out[threadIdx.x]=s;
}

// per iteration: CH DMMA (=8 warp-DFMA equivalents each) interleaved with NF DFMA per DMMA
template<int CH, int NF>
__global__ void __launch_bounds__(256) k_mix(double* out, const double* in, int iters){
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c0[CH], c1[CH], f[CH*NF+1];
  #pragma unroll
  for(int j=0;j<CH;j++){ c0[j]=j; c1[j]=-j; }
  #pragma unroll
  for(int j=0;j<CH*NF;j++) f[j]=a+j;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<CH;j++){
      dmma884(c0[j], c1[j], a, b);
      #pragma unroll
      for(int q=0;q<NF;q++) f[j*NF+q] = fma(f[j*NF+q], a, b);
    }
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<CH;j++) s+=c0[j]+c1[j];
  #pragma unroll
  for(int j=0;j<CH*NF;j++) s+=f[j];
  if(s==123.456) out[threadIdx.x]=s;
}

// DFMA fed by broadcast LDS.128: W weights per load pair used for S "subsets" per thread
template<int S>
__global__ void __launch_bounds__(256) k_lds_dfma(double* out, const double* in, int iters){
  __shared__ double2 w[1024];
  for(int i=threadIdx.x;i<1024;i+=blockDim.x) w[i]=make_double2(in[i&63], in[(i+7)&63]);
  __syncthreads();
  double x[S]; double acc[S][4];
  #pragma unroll
  for(int s=0;s<S;s++){ x[s]=in[(threadIdx.x+s)&63]; for(int j=0;j<4;j++) acc[s][j]=j; }
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int k=0;k<16;k+=2){
      double2 w0 = w[(i*16+k)&1023], w1 = w[(i*16+k+1)&1023];
      #pragma unroll
      for(int s=0;s<S;s++){
        acc[s][0]=fma(w0.x,x[s],acc[s][0]); acc[s][1]=fma(w0.y,x[s],acc[s][1]);
        acc[s][2]=fma(w1.x,x[s],acc[s][2]); acc[s][3]=fma(w1.y,x[s],acc[s][3]);
      }
    }
  }
  double s=0;
  #pragma unroll
  for(int q=0;q<S;q++) for(int j=0;j<4;j++) s+=acc[q][j];
  if(s==123.456) out[threadIdx.x]=s;
}

template<typename F>
double time_ms(F launch, int reps=5){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  std::vector<float> t;
  for(int r=0;r<reps;r++){
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); t.push_back(ms);
  }
  std::sort(t.begin(),t.end());
  return t[t.size()/2];
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount;
  double *in,*out; CK(cudaMalloc(&in,4096*8)); CK(cudaMalloc(&out,4096*8));
  std::vector<double> h(4096); for(int i=0;i<4096;i++) h[i]=1.0/(1.0+i%7)*1e-3;
  CK(cudaMemcpy(in,h.data(),4096*8,cudaMemcpyHostToDevice));
  const int iters=4096;
  printf("{\"gpu\":\"%s\",\"sms\":%d", p.name, sms);
  for(int bps : {1,2,4,8}){
    int grid=sms*bps; const int T=256;
    double ms=time_ms([&]{k_dfma<8><<<grid,T>>>(out,in,iters);});
    double fl=2.0*8*iters*(double)grid*T;
    printf(",\"dfma_tflops_bps%d\":%.3f",bps,fl/ms*1e-9);
  }
  for(int bps : {1,2,4,8}){
    int grid=sms*bps; const int T=256;
    double ms=time_ms([&]{k_dmma<8><<<grid,T>>>(out,in,iters);});
    double fl=2.0*256*8*iters*(double)grid*(T/32);
    printf(",\"dmma_tflops_bps%d\":%.3f",bps,fl/ms*1e-9);
  }
  { // latency-ish: 1 chain
    int grid=sms; const int T=32;
    double ms=time_ms([&]{k_dmma<1><<<grid,T>>>(out,in,iters);});
    printf(",\"dmma_dep_ns_per_op\":%.3f",ms*1e6/iters);
    ms=time_ms([&]{k_dfma<1><<<grid,T>>>(out,in,iters);});
    printf(",\"dfma_dep_ns_per_op\":%.3f",ms*1e6/iters);
  }
  {
    int grid=sms*4; const int T=256;
    double ms=time_ms([&]{k_mix<4,2><<<grid,T>>>(out,in,iters);});
    double fl=(2.0*256*4 + 2.0*32*8)*iters*(double)grid*(T/32);
    printf(",\"mix_dmma4_dfma8_tflops\":%.3f",fl/ms*1e-9);
    ms=time_ms([&]{k_mix<4,4><<<grid,T>>>(out,in,iters);});
    fl=(2.0*256*4 + 2.0*32*16)*iters*(double)grid*(T/32);
    printf(",\"mix_dmma4_dfma16_tflops\":%.3f",fl/ms*1e-9);
    ms=time_ms([&]{k_mix<4,8><<<grid,T>>>(out,in,iters);});
    fl=(2.0*256*4 + 2.0*32*32)*iters*(double)grid*(T/32);
    printf(",\"mix_dmma4_dfma32_tflops\":%.3f",fl/ms*1e-9);
  }
  {
    int grid=sms*4; const int T=256;
    double ms=time_ms([&]{k_lds_dfma<1><<<grid,T>>>(out,in,iters);});
    double fl=2.0*32*iters*(double)grid*T;
    printf(",\"lds128_dfma_s1_tflops\":%.3f",fl/ms*1e-9);
    ms=time_ms([&]{k_lds_dfma<2><<<grid,T>>>(out,in,iters);});
    fl=2.0*64*iters*(double)grid*T;
    printf(",\"lds128_dfma_s2_tflops\":%.3f",fl/ms*1e-9);
    ms=time_ms([&]{k_lds_dfma<4><<<grid,T>>>(out,in,iters);});
    fl=2.0*128*iters*(double)grid*T;
    printf(",\"lds128_dfma_s4_tflops\":%.3f",fl/ms*1e-9);
  }
  printf("}\n");
  return 0;
}
