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
      } else if (MODE >= 4) {     // MODE 0 + the kernel's per-column bookkeeping: bit 0 term magnitude + pivot test, bit 1 pivot stores, bit 2 parked column
        constexpr int X = MODE - 4;
        double* sd = sm + 256; double* sinv = sm + 320; int* skeep = reinterpret_cast<int*>(sm + 384); double* park = sm + 512;
        double wmr = fabs(ar[lane & 31]) ; bool bad = false; double magc = 0.0;
        colbuf[lane] = ar[0]; colbuf[32 + lane] = wmr; __syncwarp(); d = colbuf[0]; magc = colbuf[32]; nxt = colbuf[1]; inv = rcp_fast(d);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const double* cb = colbuf + (c % 3) * 64;
          const double arc = ar[c];
          if (X & 1) bad = bad || (fabs(d) <= 2.2e-16 * magc);
          const double lr = arc * inv;
          const bool mine = lane == c, below = lane > c;
          if (X & 2) { if (mine) { sd[c] = d; sinv[c] = inv; skeep[c] = 1; } }
          if (X & 4) { if (below) park[lane * 33 + c] = arc; }
          if (X & 1) { const double term = fabs(lr * arc); wmr = (below && term > wmr) ? term : wmr; }
          if (c + 1 < 32) {
            ar[c + 1] = fma(-lr, nxt, ar[c + 1]);
            double* cn = colbuf + ((c + 1) % 3) * 64; cn[lane] = ar[c + 1]; if (X & 1) cn[32 + lane] = wmr; __syncwarp();
            d = cn[c + 1]; if (X & 1) magc = cn[32 + c + 1]; nxt = (c + 2 < 32) ? cn[c + 2] : 0.0; inv = rcp_fast(d);
            if (c & 1) {
              if (c + 2 < 32) ar[c + 2] = fma(-lr, cb[c + 2], ar[c + 2]);
#pragma unroll
              for (int j = c + 3; j < 32; j += 2) { const double2 v = *reinterpret_cast<const double2*>(cb + j); ar[j] = fma(-lr, v.x, ar[j]); ar[j + 1] = fma(-lr, v.y, ar[j + 1]); }
            } else {
#pragma unroll
              for (int j = c + 2; j < 32; j += 2) { const double2 v = *reinterpret_cast<const double2*>(cb + j); ar[j] = fma(-lr, v.x, ar[j]); ar[j + 1] = fma(-lr, v.y, ar[j + 1]); }
            }
          }
        }
        if (bad) ar[0] += wmr;
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
#define RUN(MODE,thr,label) k_ldl<MODE><<<1,thr,16384>>>(in,out,c,8); CK(cudaDeviceSynchronize()); { long long hh[5]; CK(cudaMemcpy(hh,c,40,cudaMemcpyDeviceToHost)); hc=hh[0]; printf("%-70s %lld cycles per 32-column block (%.0f per column); passes 1-4: %lld %lld %lld %lld\n",label,hc,hc/32.0,hh[1],hh[2],hh[3],hh[4]); }
  RUN(0,32,"pipelined smem LDL, CTA = 1 warp") RUN(0,256,"pipelined smem LDL, CTA = 8 warps (7 at a barrier)")
  RUN(1,32,"dependent chain only via smem (fma,sts,sync,lds,rcp,mul), 1 warp") RUN(3,32,"dependent chain only via shuffle, 1 warp")
  RUN(4,32,"kernel-style loop, LDS.128 bulk, no bookkeeping") RUN(5,32,"  + term magnitude / pivot test") RUN(6,32,"  + pivot stores") RUN(8,32,"  + parked column") RUN(11,32,"  + all three")
  return 0; }
