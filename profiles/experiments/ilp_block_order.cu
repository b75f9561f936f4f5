#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
// ILP experiment: same R x 4 block, source order = column-major (VAR 0) or 4-column wavefront (VAR 1).
template<int VAR, int R>
__device__ __forceinline__ void block(int (&h)[R], int& dg, const int (&xc)[R], const int (&up_in)[4], unsigned yword, int (&outv)[4]) {
  if (VAR==0) {
#pragma unroll
    for (int c=0;c<4;++c){ unsigned sel=__funnelshift_l(0u,1u,yword>>(8*c)); int diag=dg; int up=up_in[c]; dg=up;
#pragma unroll
      for (int r=0;r<R;++r){ const int left=h[r]; int t=__dp4a(xc[r],(int)sel,diag); int v=min(min(up,left),t); diag=left; up=v; h[r]=v; }
      outv[c]=up; }
  } else {
    int diag[4], up[4]; unsigned sel[4];
#pragma unroll
    for (int c=0;c<4;++c){ sel[c]=__funnelshift_l(0u,1u,yword>>(8*c)); up[c]=up_in[c]; diag[c]=(c==0)?dg:up_in[c-1]; }
    dg = up_in[3];
#pragma unroll
    for (int k=0;k<R+3;++k){
#pragma unroll
      for (int c=0;c<4;++c){ const int r=k-c; if (r>=0&&r<R){ const int left=h[r]; int t=__dp4a(xc[r],(int)sel[c],diag[c]); int v=min(min(up[c],left),t); diag[c]=left; up[c]=v; h[r]=v; } }
    }
#pragma unroll
    for (int c=0;c<4;++c) outv[c]=up[c];
  }
}
template<int VAR, int R>
__global__ void __launch_bounds__(128) k(int* out, const unsigned* yws, int n) {
  int h[R]; int xc[R]; int dg = 0; int outv[4]={0,0,0,0};
  for (int r=0;r<R;++r){h[r]=0; xc[r]=out[(r+threadIdx.x)&1023];}
  for (int q=0;q<n;++q){
    int up_in[4];
    for (int c=0;c<4;++c) up_in[c]=__shfl_up_sync(0xffffffffu,outv[c],1);
    block<VAR,R>(h,dg,xc,up_in,yws[q&1023],outv);
  }
  int s=0; for (int r=0;r<R;++r) s+=h[r];
  out[(blockIdx.x*128+threadIdx.x)&1023]=s;
}
template<int VAR,int R> void run(const char* name, int blocks_per_sm){
  int *out; unsigned* y; cudaMalloc(&out,4096); cudaMalloc(&y,4096); cudaMemset(out,1,4096); cudaMemset(y,8,4096);
  int n=20000; int grid=148*blocks_per_sm;
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<VAR,R><<<grid,128>>>(out,y,100); cudaDeviceSynchronize();
  cudaEventRecord(a); k<VAR,R><<<grid,128>>>(out,y,n); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  double cells=(double)grid*128*n*R*4;
  int nb=0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb,k<VAR,R>,128,0);
  cudaFuncAttributes at; cudaFuncGetAttributes(&at,k<VAR,R>);
  printf("%s R=%d blocks/SM=%d (max %d, regs %d): %.1f GCUPS\n",name,R,blocks_per_sm,nb,at.numRegs,cells/ms/1e6);
}
int main(){
  run<0,32>("colmajor",4); run<1,32>("wavefront",4);
  run<0,32>("colmajor",2); run<1,32>("wavefront",2);
  run<0,32>("colmajor",1); run<1,32>("wavefront",1);
  run<0,16>("colmajor",6); run<1,16>("wavefront",6);
  run<0,8>("colmajor",1); run<1,8>("wavefront",1);
  return 0;
}
