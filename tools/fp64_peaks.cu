// Microbenchmark: FP64 denominators for the roofline (MEASURED_PEAKS.json has no FP64 entry).
//   dmma  : mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) issue rate, register operands only
//   dfma  : plain DFMA issue rate
//   dgemm : cublasDgemm square and the skinny N x 256 . 256 x 16 shape of the dense E-step
//   hbm   : streaming read of a buffer >> L2
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int CHAINS>
__global__ void dmma_kernel(double* out, int iters, double seed){
  double c[CHAINS][2];
  #pragma unroll
  for(int j=0;j<CHAINS;j++){c[j][0]=0;c[j][1]=0;}
  double a=seed*threadIdx.x, b=seed+threadIdx.x;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<CHAINS;j++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};":"+d"(c[j][0]),"+d"(c[j][1]):"d"(a),"d"(b));
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<CHAINS;j++) s+=c[j][0]+c[j][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int CHAINS>
__global__ void dfma_kernel(double* out, int iters, double seed){
  double c[CHAINS];
  #pragma unroll
  for(int j=0;j<CHAINS;j++) c[j]=j;
  double a=seed*threadIdx.x, b=seed+threadIdx.x;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<CHAINS;j++) c[j]=fma(a,c[j],b);
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<CHAINS;j++) s+=c[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void read_kernel(const double2* __restrict__ p, size_t n, double* out){
  double s=0;
  for(size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x;i<n;i+=(size_t)gridDim.x*blockDim.x){ double2 v=p[i]; s+=v.x+v.y; }
  if(s==123.456) out[0]=s;
}
static float timeit(void(*f)(void*), void* ctx, int reps){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(ctx); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<reps;r++){ cudaEventRecord(e0); f(ctx); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best) best=ms; }
  return best;
}
struct Ctx{ double* out; int warps; int ctas; int iters; int kind; };
static void run_dmma(void* p){ Ctx* c=(Ctx*)p; 
  if(c->kind==0) dmma_kernel<8><<<c->ctas,c->warps*32>>>(c->out,c->iters,1e-9);
  else if(c->kind==1) dfma_kernel<16><<<c->ctas,c->warps*32>>>(c->out,c->iters,1e-9);
  else if(c->kind==2) dmma_kernel<2><<<c->ctas,c->warps*32>>>(c->out,c->iters,1e-9);
  else dmma_kernel<1><<<c->ctas,c->warps*32>>>(c->out,c->iters,1e-9);
}
int main(){
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop,0));
  int sms=prop.multiProcessorCount; printf("device %s sms %d\n",prop.name,sms);
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*4*1024));
  int iters=20000;
  for(int kind=0;kind<4;kind++) for(int warps: {1,4,8,16,32}){
    Ctx c{out,warps,sms*((kind==0||kind==1)?1:1),iters,kind};
    float ms=timeit(run_dmma,&c,5);
    double chains = kind==0?8:(kind==1?16:(kind==2?2:1));
    double flop = (kind==1? 2.0*32 : 512.0)*chains*iters*(double)warps*c.ctas;
    printf("%s chains=%d warps/SM=%d : %.3f ms  %.2f TFLOP/s  (%.2f clk/instr/SM @1.9GHz)\n", kind==1?"dfma":"dmma",(int)chains,warps,ms,flop/ms*1e-9, ms*1e-3*1.9e9/(chains*iters*warps));
  }
  // HBM read
  { size_t bytes=(size_t)8<<30; double2* buf; CK(cudaMalloc(&buf,bytes)); CK(cudaMemset(buf,0,bytes));
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best=1e30f;
    for(int r=0;r<6;r++){ cudaEventRecord(e0); read_kernel<<<sms*8,512>>>(buf,bytes/16,out); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    printf("hbm read 8GiB: %.3f ms %.1f GB/s\n",best,bytes/best*1e-6); cudaFree(buf); }
  // cublas
  cublasHandle_t h; cublasCreate(&h);
  { int n=8192; double *A,*B,*C; CK(cudaMalloc(&A,8.0*n*n)); CK(cudaMalloc(&B,8.0*n*n)); CK(cudaMalloc(&C,8.0*n*n)); cudaMemset(A,0,8.0*n*n); cudaMemset(B,0,8.0*n*n);
    double al=1,be=0; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best=1e30f;
    for(int r=0;r<5;r++){ cudaEventRecord(e0); cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,n,n,n,&al,A,n,B,n,&be,C,n); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    printf("cublasDgemm %d^3: %.3f ms %.2f TFLOP/s\n",n,best,2.0*n*n*n/best*1e-9);
    // sustained: 3 seconds
    cudaEventRecord(e0); int cnt=0; for(int r=0;r<80;r++){ cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,n,n,n,&al,A,n,B,n,&be,C,n); cnt++; } cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1);
    printf("cublasDgemm %d^3 sustained x%d: %.3f ms/each %.2f TFLOP/s\n",n,cnt,ms/cnt,2.0*n*n*n*cnt/ms*1e-9);
    cudaFree(A);cudaFree(B);cudaFree(C); }
  { size_t N=1<<20; int D=256,M=16; double *X,*W,*T; CK(cudaMalloc(&X,8.0*N*D)); CK(cudaMalloc(&W,8.0*D*M)); CK(cudaMalloc(&T,8.0*N*M)); cudaMemset(X,0,8.0*N*D); cudaMemset(W,0,8.0*D*M);
    double al=1,be=0; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best=1e30f;
    // row-major X (N x D) . W (D x M) == col-major W^T(M x D) . X^T (D x N) -> T^T (M x N)
    for(int r=0;r<5;r++){ cudaEventRecord(e0); cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,M,(int)N,D,&al,W,M,X,D,&be,T,M); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    printf("cublasDgemm skinny X(2^20x256).W(256x16): %.3f ms %.2f TFLOP/s %.1f GB/s\n",best,2.0*N*D*M/best*1e-9, 8.0*N*D/best*1e-6);
    best=1e30f; double* XE; CK(cudaMalloc(&XE,8.0*D*M));
    // X^T (D x N) . T (N x M) -> D x M ; col-major: T^T (M x N) . X (N x D as col-major D x N transposed) 
    for(int r=0;r<5;r++){ cudaEventRecord(e0); cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_T,M,D,(int)N,&al,T,M,X,D,&be,XE,M); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
    printf("cublasDgemm skinny X^T.T (256x2^20 . 2^20x16): %.3f ms %.2f TFLOP/s %.1f GB/s\n",best,2.0*N*D*M/best*1e-9, 8.0*N*D/best*1e-6);
  }
  return 0;
}
