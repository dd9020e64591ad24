// Latency probes for the FP64 pipe on B200 (dependent chains, one warp / several warps per SM).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../mrlda_b200/csrc/device_math.cuh"
using namespace mrlda;

template <int MODE>
__global__ void k(double *out, long long *cyc, int iters, double x0) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.000001, z = 1e-9;
  double a1 = x + 1, a2 = x + 2, a3 = x + 3;
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
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + a1 + a2 + a3;
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
  }
  for (int t : {32, 256}) {
    run<6>("DADD chain", t, 4096, 1);
    run<7>("SHFL.64 + DADD", t, 4096, 1);
    run<4>("rcp_fast + DADD", t, 2048, 1);
    run<5>("exp() + DFMA", t, 2048, 1);
    run<8>("digamma_fast + DADD", t, 1024, 1);
    run<3>("exp_digamma + DADD", t, 1024, 1);
  }
  return 0;
}
