#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ unsigned ex2_bf16x2(unsigned x){unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;":"=r"(y):"r"(x)); return y;}
__device__ __forceinline__ unsigned ex2_f16x2(unsigned x){unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;":"=r"(y):"r"(x)); return y;}
__device__ __forceinline__ float ex2(float x){float y; asm volatile("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
template<int MODE> __global__ void probe(unsigned* out, long long* cyc, int iters){
  unsigned v[32];
  #pragma unroll
  for(int i=0;i<32;++i) v[i]= (MODE==0)? __float_as_uint(-0.001f*(threadIdx.x+i)) : 0xbc00bc00u + threadIdx.x + i;
  __syncthreads();
  long long t0=clock64();
  for(int it=0;it<iters;++it){
    #pragma unroll
    for(int i=0;i<32;++i){
      if(MODE==0) v[i]=__float_as_uint(ex2(__uint_as_float(v[i])));
      else if(MODE==1) v[i]=ex2_bf16x2(v[i]);
      else v[i]=ex2_f16x2(v[i]);
    }
  }
  long long t1=clock64();
  unsigned s=0;
  #pragma unroll
  for(int i=0;i<32;++i) s^=v[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if(threadIdx.x==0&&blockIdx.x==0) cyc[0]=t1-t0;
}
__global__ void max3(float* o, const float* a){ float x=a[threadIdx.x], y=a[threadIdx.x+32], z=a[threadIdx.x+64]; float r; asm("max.f32 %0, %1, %2, %3;":"=f"(r):"f"(x),"f"(y),"f"(z)); o[threadIdx.x]=r; }
int main(){
  unsigned* out; long long* cyc; cudaMalloc(&out,1<<22); cudaMalloc(&cyc,8);
  const int iters=1000;
  for(int mode=0;mode<3;++mode) for(int warps=4;warps<=32;warps*=2){
    for(int rep=0;rep<2;++rep){
      if(mode==0) probe<0><<<148,warps*32>>>(out,cyc,iters); else if(mode==1) probe<1><<<148,warps*32>>>(out,cyc,iters); else probe<2><<<148,warps*32>>>(out,cyc,iters);
    }
    cudaDeviceSynchronize(); long long h; cudaMemcpy(&h,cyc,8,cudaMemcpyDeviceToHost);
    double per_instr=(double)h/(iters*32.0); double per_smsp=per_instr/(warps/4.0);
    printf("mode %d (%s) warps/SM %2d: %.2f cyc/op/warp, %.2f cyc per warp-op per SMSP, %.1f lane-ops/clk/SM (x2 elements for packed)\n",mode,mode==0?"f32":mode==1?"bf16x2":"f16x2",warps,per_instr,per_smsp,32.0*4/per_smsp);
  }
  printf("%s\n",cudaGetErrorString(cudaGetLastError()));
}
