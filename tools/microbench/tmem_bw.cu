// TMEM as thread-private storage: tcgen05.ld bandwidth into registers, 8 warps per SM, compared
// with the same volume from shared memory (LDS.128, conflict-free lane-interleaved layout).
#include <cstdio>
#include <cuda_runtime.h>
#include "tmem_ops.cuh"

__global__ void __launch_bounds__(256, 1) tmem_kernel(double *out, long long *cyc, int iters) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = base_s;
  // my private region: lane quarter (warp % 4), 256 columns chosen by warp / 4
  const uint32_t mine = base + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 256u;
  uint32_t r[32];
  for (int c = 0; c < 8; c++) {
    for (int i = 0; i < 32; i += 2) {
      double v = 1.0 + threadIdx.x * 1e-3 + (c * 32 + i) * 1e-6;
      r[i] = __double2loint(v);
      r[i + 1] = __double2hiint(v);
    }
    tmem_st32(mine + c * 32, r);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;");
  __syncthreads();
  double acc0 = 0, acc1 = 0;
  uint32_t ia = 0, ib = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
      tmem_ld32(mine + c * 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        ia += r[i];
        ib ^= r[i + 1];
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * 256 + threadIdx.x] = acc0 + acc1 + ia + ib;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}

__global__ void __launch_bounds__(256, 1) smem_kernel(double *out, long long *cyc, int iters) {
  extern __shared__ double2 sm[];  // 256 threads x 64 double2 = 256 KB?  no: 128 doubles/thread = 1 KB -> 256 KB too big; use 96 doubles
  const int tid = threadIdx.x;
  constexpr int U = 40;  // 40 double2 = 640 B per thread -> 160 KB
  for (int i = 0; i < U; i++) sm[i * 256 + tid] = make_double2(1.0 + tid * 1e-3, i * 1e-6);
  __syncthreads();
  double acc0 = 0, acc1 = 0;
  uint32_t ia = 0, ib = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < U; i++) {
      unsigned a = (unsigned)__cvta_generic_to_shared(&sm[i * 256 + tid]);
      uint32_t x0, x1, x2, x3;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(a));
      ia += x0 + x2;
      ib ^= x1 ^ x3;
    }
  }
  long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * 256 + tid] = acc0 + acc1 + ia + ib;
}

int main() {
  double *out;
  long long *cyc, h;
  cudaMalloc(&out, 148 * 256 * 8);
  cudaMalloc(&cyc, 8);
  const int iters = 200;
  for (int rep = 0; rep < 2; rep++) {
    tmem_kernel<<<148, 256>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("TMEM  ld: %s  %.1f cycles per 256 KB (1 KB/thread, 8 warps) -> %.1f B/cycle/SM\n",
           cudaGetErrorString(e), (double)h / iters, 262144.0 * iters / h);
  }
  cudaFuncSetAttribute(smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 256 * 16);
  for (int rep = 0; rep < 2; rep++) {
    smem_kernel<<<148, 256, 40 * 256 * 16>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("SMEM LDS: %s  %.1f cycles per 160 KB (640 B/thread, 8 warps) -> %.1f B/cycle/SM\n",
           cudaGetErrorString(e), (double)h / iters, 163840.0 * iters / h);
  }
  return 0;
}
