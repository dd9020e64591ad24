// Latency probes for the FP64 pipe on B200 (dependent chains, one warp / several warps per SM).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../mrlda_b200/csrc/device_math.cuh"
using namespace mrlda;

template <int MODE>
__global__ void k(double *out, long long *cyc, int iters, double x0) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.000001, z = 1e-9;
  double a1 = x + 1, a2 = x + 2, a3 = x + 3;
  double b0 = x + 4, b1 = x + 5, b2 = x + 6, b3 = x + 7, c0 = x + 8, c1 = x + 9, c2 = x + 10, c3 = x + 11;
  double d0 = x + 12, d1 = x + 13, d2 = x + 14, d3 = x + 15;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) { x = fma(x, y, z); }                                   // 1 chain
    if (MODE == 1) { x = fma(x, y, z); a1 = fma(a1, y, z); }               // 2 chains
    if (MODE == 2) { x = fma(x, y, z); a1 = fma(a1, y, z); a2 = fma(a2, y, z); a3 = fma(a3, y, z); }
    if (MODE == 3) { x = exp_digamma(x) + 3.0; }
    if (MODE == 4) { x = rcp_fast(x) + 1.5; }
    if (MODE == 5) { x = exp(x) * 1e-3 + 0.5; }
    if (MODE == 6) { x = x + y; }                                           // DADD chain
    if (MODE == 7) { x = __shfl_xor_sync(0xffffffffu, x, 1) + y; }          // SHFL(2x32) + DADD
    if (MODE == 8) { x = digamma_fast(x) + 5.0; }
    if (MODE == 9) {  // 16 independent chains: the issue rate of one warp's DFMA stream
      x = fma(x, y, z); a1 = fma(a1, y, z); a2 = fma(a2, y, z); a3 = fma(a3, y, z);
      b0 = fma(b0, y, z); b1 = fma(b1, y, z); b2 = fma(b2, y, z); b3 = fma(b3, y, z);
      c0 = fma(c0, y, z); c1 = fma(c1, y, z); c2 = fma(c2, y, z); c3 = fma(c3, y, z);
      d0 = fma(d0, y, z); d1 = fma(d1, y, z); d2 = fma(d2, y, z); d3 = fma(d3, y, z);
    }
    if (MODE == 10) { const double in[1] = {x}; double o[1]; exp_digamma_n<1>(in, o); x = o[0] + 3.0; }
    if (MODE == 11) { const double in[2] = {x, a1}; double o[2]; exp_digamma_n<2>(in, o); x = o[0] + 3.0; a1 = o[1] + 2.0; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + a1 + a2 + a3 + b0 + b1 + b2 + b3 + c0 + c1 + c2 + c3 + d0 + d1 + d2 + d3;
}

template <int MODE>
void run(const char *name, int threads, int iters, int per_iter) {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  k<MODE><<<148, threads>>>(out, cyc, iters, 1.5);
  k<MODE><<<148, threads>>>(out, cyc, iters, 1.5);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s threads/SM=%4d  %.1f cycles per iteration (%d op(s)/iter)\n", name, threads,
         (double)h / iters, per_iter);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int t : {32, 128, 256, 512}) {
    run<0>("DFMA 1 chain", t, 4096, 1);
    run<1>("DFMA 2 chains", t, 4096, 2);
    run<2>("DFMA 4 chains", t, 4096, 4);
    run<9>("DFMA 16 chains", t, 4096, 16);
  }
  for (int t : {32, 256}) {
    run<6>("DADD chain", t, 4096, 1);
    run<7>("SHFL.64 + DADD", t, 4096, 1);
    run<4>("rcp_fast + DADD", t, 2048, 1);
    run<5>("exp() + DFMA", t, 2048, 1);
    run<8>("digamma_fast + DADD", t, 1024, 1);
    run<3>("exp_digamma + DADD", t, 1024, 1);
    run<10>("exp_digamma_n<1> + DADD", t, 1024, 1);
    run<11>("exp_digamma_n<2> + DADD", t, 1024, 2);
  }
  return 0;
}
