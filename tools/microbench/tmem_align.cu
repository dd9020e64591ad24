// Does tcgen05.ld/st 32x32b accept column offsets that are not multiples of the vector width?
// Writes a 48-word-stride layout (x32 at +0, x16 at +32 of every 48-column slot) and reads it back.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../mrlda_b200/csrc/tmem_ops.cuh"
using namespace mrlda;

__global__ void __launch_bounds__(256, 1) k(int *bad) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&base_s, 512);
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t mine = base_s + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 256u;
  uint32_t a[32], b[16];
  for (int s = 0; s < 5; s++) {
    for (int i = 0; i < 32; i++) a[i] = threadIdx.x * 1000 + s * 100 + i;
    for (int i = 0; i < 16; i++) b[i] = threadIdx.x * 1000 + s * 100 + 32 + i;
    TmemVec<32>::st(mine + s * 48, a);
    TmemVec<16>::st(mine + s * 48 + 32, b);
  }
  tmem_wait_st();
  int nbad = 0;
  for (int s = 0; s < 5; s++) {
    TmemVec<32>::ld(mine + s * 48, a);
    TmemVec<16>::ld(mine + s * 48 + 32, b);
    tmem_wait_ld();
    for (int i = 0; i < 32; i++) nbad += a[i] != threadIdx.x * 1000 + s * 100 + i;
    for (int i = 0; i < 16; i++) nbad += b[i] != threadIdx.x * 1000 + s * 100 + 32 + i;
  }
  if (nbad) atomicAdd(bad, nbad);
  tmem_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(base_s, 512);
}
int main() {
  int *bad, h = -1;
  cudaMalloc(&bad, 4);
  cudaMemset(bad, 0, 4);
  k<<<148, 256>>>(bad);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost);
  printf("48-column stride round trip: %s, mismatches = %d\n", cudaGetErrorString(e), h);
  return 0;
}
