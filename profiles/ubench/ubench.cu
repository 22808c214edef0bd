// micro-benchmarks for FP64 on B200: latencies and throughputs that decide the dense-window kernel design
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e){printf("err %s line %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)
__global__ void k_lat_dfma(double* out, long long* cyc, int n){ double a=out[0],b=out[1],c=out[2]; long long t0=clock64();
  for(int i=0;i<n;++i){ a=fma(a,b,c); a=fma(a,b,c); a=fma(a,b,c); a=fma(a,b,c);} long long t1=clock64(); out[3]=a; cyc[0]=t1-t0; }
__global__ void k_lat_rcp(double* out, long long* cyc, int n){ double a=out[0]; long long t0=clock64();
  for(int i=0;i<n;++i){ a=__drcp_rn(a); a=__drcp_rn(a); a=__drcp_rn(a); a=__drcp_rn(a);} long long t1=clock64(); out[3]=a; cyc[0]=t1-t0; }
__global__ void k_lat_div(double* out, long long* cyc, int n){ double a=out[0],b=out[1]; long long t0=clock64();
  for(int i=0;i<n;++i){ a=b/a; a=b/a; a=b/a; a=b/a;} long long t1=clock64(); out[3]=a; cyc[0]=t1-t0; }
__global__ void k_lat_shfl(double* out, long long* cyc, int n){ double a=out[threadIdx.x&3]; long long t0=clock64();
  for(int i=0;i<n;++i){ a=__shfl_sync(0xffffffffu,a,(threadIdx.x+1)&31); a=__shfl_sync(0xffffffffu,a,3); a=__shfl_sync(0xffffffffu,a,(threadIdx.x+5)&31); a=__shfl_sync(0xffffffffu,a,7);} long long t1=clock64(); out[3]=a; cyc[0]=t1-t0; }
__global__ void k_lat_lds(double* out, long long* cyc, int n){ __shared__ double s[64]; s[threadIdx.x&63]=(double)((threadIdx.x*7+1)&31); __syncthreads(); int j=threadIdx.x&31; long long t0=clock64();
  for(int i=0;i<n;++i){ j=(int)s[j]; j=(int)s[j]; j=(int)s[j]; j=(int)s[j]; } long long t1=clock64(); out[3]=j; cyc[0]=t1-t0; }
__global__ void k_lat_bar(double* out, long long* cyc, int n){ long long t0=clock64();
  for(int i=0;i<n;++i){ __syncthreads(); __syncthreads(); __syncthreads(); __syncthreads(); } long long t1=clock64(); if(threadIdx.x==0) cyc[0]=t1-t0; }
// DFMA throughput: ILP independent chains per thread
template<int ILP> __global__ void k_thr_dfma(double* out, int n){ double a[ILP]; double b=out[1],c=out[2];
  for(int u=0;u<ILP;++u) a[u]=out[0]+u+threadIdx.x;
  for(int i=0;i<n;++i){
#pragma unroll
    for(int u=0;u<ILP;++u) a[u]=fma(a[u],b,c); }
  double s=0; for(int u=0;u<ILP;++u) s+=a[u]; if(s==123.456) out[3]=s; }
// 8x8 outer-product micro-tile from registers only (no memory): what the update kernel's inner loop could reach
__global__ void k_thr_outer(double* out, int n){ double acc[8][8]; double av[8], bv[8];
  for(int u=0;u<8;++u){ av[u]=out[0]+u+threadIdx.x; bv[u]=out[1]+u; for(int v=0;v<8;++v) acc[u][v]=0; }
  for(int i=0;i<n;++i){
#pragma unroll
    for(int u=0;u<8;++u)
#pragma unroll
      for(int v=0;v<8;++v) acc[u][v]=fma(av[u],bv[v],acc[u][v]);
    av[i&7]+=1.0; }
  double s=0; for(int u=0;u<8;++u) for(int v=0;v<8;++v) s+=acc[u][v]; if(s==123.456) out[3]=s; }
// DMMA m8n8k4 throughput: NACC independent accumulator tiles per warp
template<int NACC> __global__ void k_thr_dmma(double* out, int n){ double c0[NACC], c1[NACC]; double a=out[0]+threadIdx.x, b=out[1]+threadIdx.x;
  for(int u=0;u<NACC;++u){c0[u]=0;c1[u]=0;}
  for(int i=0;i<n;++i){
#pragma unroll
    for(int u=0;u<NACC;++u) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0[u]), "+d"(c1[u]) : "d"(a), "d"(b)); }
  double s=0; for(int u=0;u<NACC;++u) s+=c0[u]+c1[u]; if(s==123.456) out[3]=s; }
// mixed: DMMA and DFMA interleaved (are they separate pipes?)
__global__ void k_thr_mixed(double* out, int n){ double c0[8], c1[8], f[16]; double a=out[0]+threadIdx.x, b=out[1]+threadIdx.x, cc=out[2];
  for(int u=0;u<8;++u){c0[u]=0;c1[u]=0;} for(int u=0;u<16;++u) f[u]=u;
  for(int i=0;i<n;++i){
#pragma unroll
    for(int u=0;u<8;++u){ asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0[u]), "+d"(c1[u]) : "d"(a), "d"(b));
      f[2*u]=fma(f[2*u],b,cc); f[2*u+1]=fma(f[2*u+1],b,cc);} }
  double s=0; for(int u=0;u<8;++u) s+=c0[u]+c1[u]; for(int u=0;u<16;++u) s+=f[u]; if(s==123.456) out[3]=s; }
// LDS.128 patterns of the update kernel: 16 distinct 16B addresses (A operand) + 2 distinct (B operand) per warp
__global__ void k_thr_lds(double* out, int n, int mode){ extern __shared__ double sm[]; for(int i=threadIdx.x;i<4096;i+=blockDim.x) sm[i]=i; __syncthreads();
  int tx=threadIdx.x&15, ty=threadIdx.x>>4; double s=0;
  for(int i=0;i<n;++i){ int c=i&15;
#pragma unroll
    for(int u=0;u<4;++u){ const double2 x=*reinterpret_cast<const double2*>(sm + c*128 + (mode==1? 2*ty : 2*tx) + 32*u); s+=x.x+x.y; }
  }
  if(s==123.456) out[3]=s; }
int main(){ double* d; long long* c; CK(cudaMalloc(&d,64)); CK(cudaMalloc(&c,64)); double h[4]={1.000001,0.999999,1e-9,0}; CK(cudaMemcpy(d,h,32,cudaMemcpyHostToDevice));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount; double ghz=p.clockRate*1e-6; printf("%s sms %d clock %.3f GHz\n",p.name,sms,ghz);
  long long hc; const int n=4096;
#define LAT(k,thr,label) k<<<1,thr>>>(d,c,n); CK(cudaDeviceSynchronize()); k<<<1,thr>>>(d,c,n); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&hc,c,8,cudaMemcpyDeviceToHost)); printf("latency %-28s %.1f cycles\n",label,(double)hc/(4.0*n));
  LAT(k_lat_dfma,32,"dependent DFMA") LAT(k_lat_rcp,32,"__drcp_rn") LAT(k_lat_div,32,"double divide") LAT(k_lat_shfl,32,"shfl double") LAT(k_lat_lds,32,"LDS + cvt (pointer chase)") LAT(k_lat_bar,256,"__syncthreads 256 thr") LAT(k_lat_bar,32,"__syncthreads 32 thr")
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
#define THR(launch,flops,label) launch; CK(cudaDeviceSynchronize()); cudaEventRecord(e0); launch; cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms,e0,e1); printf("throughput %-44s %.2f TFLOP/s\n",label,(flops)/(ms*1e-3)/1e12);
  const int it=20000;
  for(int w=1;w<=16;w*=2){ char lab[96]; snprintf(lab,96,"DFMA ILP8, %2d warps/SM",w); THR((k_thr_dfma<8><<<sms,32*w>>>(d,it)),2.0*8*it*32*w*sms,lab) }
  for(int w=4;w<=16;w*=2){ char lab[96]; snprintf(lab,96,"DFMA ILP2, %2d warps/SM",w); THR((k_thr_dfma<2><<<sms,32*w>>>(d,it)),2.0*2*it*32*w*sms,lab) }
  for(int w=4;w<=16;w*=2){ char lab[96]; snprintf(lab,96,"8x8 outer product regs, %2d warps/SM",w); THR((k_thr_outer<<<sms,32*w>>>(d,it/8)),2.0*64*(it/8)*32*w*sms,lab) }
  for(int w=1;w<=16;w*=2){ char lab[96]; snprintf(lab,96,"DMMA m8n8k4 8 acc, %2d warps/SM",w); THR((k_thr_dmma<8><<<sms,32*w>>>(d,it)),2.0*256*8*it*w*sms,lab) }
  for(int w=4;w<=16;w*=2){ char lab[96]; snprintf(lab,96,"DMMA m8n8k4 2 acc, %2d warps/SM",w); THR((k_thr_dmma<2><<<sms,32*w>>>(d,it)),2.0*256*2*it*w*sms,lab) }
  for(int w=4;w<=16;w*=2){ char lab[96]; snprintf(lab,96,"mixed 8 DMMA + 16 DFMA, %2d warps/SM",w); THR((k_thr_mixed<<<sms,32*w>>>(d,it)),2.0*(256*8+16*32)*it*w*sms,lab) }
  for(int mode=0;mode<2;++mode){ k_thr_lds<<<sms,256,32768>>>(d,it,mode); CK(cudaDeviceSynchronize()); cudaEventRecord(e0); k_thr_lds<<<sms,256,32768>>>(d,it,mode); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms,e0,e1);
    double cyc=ms*1e-3*ghz*1e9; printf("LDS.128 pattern %s: %.2f cycles per warp-instruction (8 warps/SM)\n", mode?"B (2 distinct addr)":"A (16 distinct 16B)", cyc/(4.0*it*8)); }
  return 0; }
