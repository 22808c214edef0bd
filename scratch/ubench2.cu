// single-warp calibration: how fast does ONE warp run a 32-column register LDL^T (the k_panel_diag inner loop)?
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e){printf("err %s line %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)
__device__ __forceinline__ double rcp_fast(double d){ double x; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e=fma(-d,x,1.0); e=fma(e,e,e); x=fma(x,e,x); e=fma(-d,x,1.0); return fma(x,e,x); }
template<int MODE> __global__ void k_ldl(const double* in, double* out, long long* cyc, int reps)
{
  extern __shared__ __align__(16) double sm[];
  double* colbuf = sm; const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    long long total = 0;
    for (int rep = 0; rep < reps; ++rep) {
      double ar[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) ar[j] = in[lane * 32 + j] + (j == lane ? 64.0 : 0.0);
      __syncwarp();
      long long t0 = clock64();
      double d, inv, nxt;
      if (MODE == 0) {            // shared-memory broadcast, pipelined (as in the kernel)
        colbuf[lane] = ar[0]; __syncwarp(); d = colbuf[0]; nxt = colbuf[1]; inv = rcp_fast(d);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const double* cb = colbuf + (c % 3) * 64;
          const double arc = ar[c]; const double lr = arc * inv;
          if (c + 1 < 32) {
            ar[c + 1] = fma(-lr, nxt, ar[c + 1]);
            double* cn = colbuf + ((c + 1) % 3) * 64; cn[lane] = ar[c + 1]; __syncwarp();
            d = cn[c + 1]; nxt = (c + 2 < 32) ? cn[c + 2] : 0.0; inv = rcp_fast(d);
#pragma unroll
            for (int j = c + 2; j < 32; ++j) ar[j] = fma(-lr, cb[j], ar[j]);
          }
          if (lane > c) ar[c] = lr;
        }
      } else if (MODE == 1) {     // no pipelining, no bulk: only the dependent chain (fma, store, sync, load, rcp, mul)
        colbuf[lane] = ar[0]; __syncwarp(); d = colbuf[0]; nxt = colbuf[1]; inv = rcp_fast(d);
#pragma unroll
        for (int c = 0; c < 31; ++c) {
          const double lr = ar[c] * inv;
          ar[c + 1] = fma(-lr, nxt, ar[c + 1]);
          double* cn = colbuf + ((c + 1) % 3) * 64; cn[lane] = ar[c + 1]; __syncwarp();
          d = cn[c + 1]; nxt = cn[(c + 2) & 31]; inv = rcp_fast(d);
        }
      } else if (MODE == 2) {     // bulk only: 496 fma + LDS.64 broadcast, no chain
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const double* cb = colbuf + (c % 3) * 64; const double lr = ar[c];
#pragma unroll
          for (int j = c + 2; j < 32; ++j) ar[j] = fma(-lr, cb[j], ar[j]);
        }
      } else {                    // chain by shuffle
        d = __shfl_sync(0xffffffffu, ar[0], 0); nxt = __shfl_sync(0xffffffffu, ar[0], 1); inv = rcp_fast(d);
#pragma unroll
        for (int c = 0; c < 31; ++c) {
          const double lr = ar[c] * inv;
          ar[c + 1] = fma(-lr, nxt, ar[c + 1]);
          d = __shfl_sync(0xffffffffu, ar[c + 1], c + 1); nxt = __shfl_sync(0xffffffffu, ar[c + 1], (c + 2) & 31); inv = rcp_fast(d);
        }
      }
      long long t1 = clock64(); total += t1 - t0; if (lane == 0 && rep < 4) cyc[1 + rep] = t1 - t0;
      double s = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) s += ar[j];
      out[lane] = s + d + inv;
    }
    if (lane == 0) cyc[0] = total / reps;
    __syncthreads();
  } else {
    __syncthreads();
  }
}
int main(){ double *in,*out; long long* c; CK(cudaMalloc(&in,8192)); CK(cudaMalloc(&out,256)); CK(cudaMalloc(&c,64));
  double h[1024]; for(int i=0;i<1024;++i) h[i]=((i*7919)%1000)/1000.0; CK(cudaMemcpy(in,h,8192,cudaMemcpyHostToDevice));
  long long hc;
#define RUN(MODE,thr,label) k_ldl<MODE><<<1,thr,4096>>>(in,out,c,8); CK(cudaDeviceSynchronize()); { long long hh[5]; CK(cudaMemcpy(hh,c,40,cudaMemcpyDeviceToHost)); hc=hh[0]; printf("%-70s %lld cycles per 32-column block (%.0f per column); passes 1-4: %lld %lld %lld %lld\n",label,hc,hc/32.0,hh[1],hh[2],hh[3],hh[4]); }
  RUN(0,32,"pipelined smem LDL, CTA = 1 warp") RUN(0,256,"pipelined smem LDL, CTA = 8 warps (7 at a barrier)")
  RUN(1,32,"dependent chain only via smem (fma,sts,sync,lds,rcp,mul), 1 warp") RUN(3,32,"dependent chain only via shuffle, 1 warp")
  RUN(2,32,"bulk only (465 LDS.64 broadcast + fma), 1 warp") RUN(2,256,"bulk only, CTA = 8 warps (7 at a barrier)")
  return 0; }
