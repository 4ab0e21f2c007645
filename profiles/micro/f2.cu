// microbenchmark: FFMA vs FFMA2 vs FADD vs FADD2 throughput per SM (sm_100a)
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ void add2(float &ax, float &ay, float bx, float by){ asm volatile("{.reg .b64 ra, rb; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; add.rn.f32x2 ra, ra, rb; mov.b64 {%0,%1}, ra;}" : "+f"(ax), "+f"(ay) : "f"(bx), "f"(by)); }
__device__ __forceinline__ void fma2(float &ax, float &ay, float bx, float by, float cx, float cy){ asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; mov.b64 rc, {%4,%5}; fma.rn.f32x2 ra, ra, rb, rc; mov.b64 {%0,%1}, ra;}" : "+f"(ax), "+f"(ay) : "f"(bx), "f"(by), "f"(cx), "f"(cy)); }
template<int MODE> __global__ void __launch_bounds__(512) k(float *out, int iters, float b, float c){
  float x[16], y[16];
  for (int i=0;i<16;++i){ x[i]=threadIdx.x*0.001f+i; y[i]=x[i]+0.5f; }
  for (int it=0; it<iters; ++it){
#pragma unroll
    for (int i=0;i<16;++i){
      if (MODE==0){ x[i]=fmaf(x[i],b,c); y[i]=fmaf(y[i],b,c);}            // 2 FFMA
      if (MODE==1){ fma2(x[i],y[i],b,b,c,c);}                              // 1 FFMA2
      if (MODE==2){ x[i]=x[i]+b; y[i]=y[i]+c;}                            // 2 FADD
      if (MODE==3){ add2(x[i],y[i],b,c);}                                  // 1 FADD2
    }
  }
  float s=0; for (int i=0;i<16;++i) s+=x[i]+y[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  float *d; cudaMalloc(&d, 148*8*512*4);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters=20000;
  const char* names[4]={"FFMA x2","FFMA2","FADD x2","FADD2"};
  for (int m=0;m<4;++m){
    for (int rep=0;rep<2;++rep){
      cudaEventRecord(e0);
      if(m==0) k<0><<<148*2,512>>>(d,iters,1.0001f,0.5f);
      if(m==1) k<1><<<148*2,512>>>(d,iters,1.0001f,0.5f);
      if(m==2) k<2><<<148*2,512>>>(d,iters,1.0001f,0.5f);
      if(m==3) k<3><<<148*2,512>>>(d,iters,1.0001f,0.5f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms,e0,e1);
      double lane_ops = 148.0*2*512*(double)iters*32;  // scalar ops (each of fma/add on one float)
      if(rep) printf("%s: %.3f ms, %.1f G scalar-ops/s (%.2f per clk per SM at 1.965 GHz)\n", names[m], ms, lane_ops/ms/1e6, lane_ops/(ms*1e-3)/148/1.965e9);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
