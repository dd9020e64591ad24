// TMA tile::gather4 of fp64 rows: does cp.async.bulk.tensor.2d...tile::gather4 fetch four K-contiguous
// rows of the term-major E matrix (picked by term id) into shared memory as a [4][COLS] tile, and how
// fast is it per SM next to the coalesced LDG row read of rowload.cu (55 B/clk/SM)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gather4 gather4.cu && ./gather4
// Prints, for box = {COLS, 1} and {COLS, 4}: whether the tile verifies, then cycles per 128-row document
// (32 gathers in flight on one mbarrier phase) with every SM busy.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

constexpr int COLS = 200, ROWS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ CUtensorMap tm, const int *__restrict__ rows,
                                            int ndocs, const double *__restrict__ E, int *bad, long long *cyc,
                                            int issuers) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  double *tile = reinterpret_cast<double *>(smem_raw);
  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(issuers));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  unsigned phase = 0;
  int nbad = 0;
  const long long t0 = clock64();
  for (int d = 0; d < ndocs; d++) {
    const int *rl = rows + ((size_t)blockIdx.x * ndocs + d) * ROWS;
    // `issuers` threads (one per warp) share the 32 gathers of a document
    if ((tid & 31) == 0 && (tid >> 5) < issuers) {
      const int w = tid >> 5, per = (ROWS / 4) / issuers;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)),
                   "r"(per * 4 * COLS * 8));
      for (int g = w * per; g < (w + 1) * per; g++) {
        const int r0 = rl[4 * g], r1 = rl[4 * g + 1], r2 = rl[4 * g + 2], r3 = rl[4 * g + 3];
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, "
            "%3, %4, %5, %6}], [%7];" ::"r"(smem_u32(tile + (size_t)g * 4 * COLS)),
            "l"(&tm), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(&mbar))
            : "memory");
      }
    }
    // everybody waits for the phase
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
          : "=r"(done)
          : "r"(smem_u32(&mbar)), "r"(phase)
          : "memory");
    }
    phase ^= 1;
    if (bad && d == 0) {
      for (int i = tid; i < ROWS * COLS; i += blockDim.x) {
        const int r = i / COLS, c = i % COLS;
        if (tile[i] != E[(size_t)rl[r] * COLS + c]) nbad++;
      }
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (bad && nbad) atomicAdd(bad, nbad);
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int V = 100000, NDOCS = 64;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  std::vector<double> hE((size_t)V * COLS);
  for (size_t i = 0; i < hE.size(); i++) hE[i] = (double)(i % 1000003) * 0.5;
  std::vector<int> hrows((size_t)sms * NDOCS * ROWS);
  srand(7);
  for (auto &r : hrows) r = rand() % V;
  double *E;
  int *rows, *bad;
  long long *cyc;
  cudaMalloc(&E, hE.size() * 8);
  cudaMalloc(&rows, hrows.size() * 4);
  cudaMalloc(&bad, 4);
  cudaMalloc(&cyc, sms * 8);
  cudaMemcpy(E, hE.data(), hE.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(rows, hrows.data(), hrows.size() * 4, cudaMemcpyHostToDevice);

  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres) != cudaSuccess ||
      !encode) {
    printf("cuTensorMapEncodeTiled not available\n");
    return 1;
  }
  const size_t smem = (size_t)ROWS * COLS * 8;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (unsigned boxrows : {1u, 4u}) {
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)COLS, (cuuint64_t)V};
    const cuuint64_t gstride[1] = {(cuuint64_t)COLS * 8};
    const cuuint32_t box[2] = {(cuuint32_t)COLS, boxrows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, E, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box {%d, %u}: encode rc=%d\n", COLS, boxrows, (int)rc);
    if (rc != CUDA_SUCCESS) continue;
    for (int issuers : {1, 8}) {
      cudaMemset(bad, 0, 4);
      k<<<sms, 256, smem>>>(tm, rows, NDOCS, E, bad, cyc, issuers);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("  issuers %d: kernel failed: %s\n", issuers, cudaGetErrorString(e));
        return 1;
      }
      int hbad = -1;
      cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost);
      k<<<sms, 256, smem>>>(tm, rows, NDOCS, E, nullptr, cyc, issuers);
      cudaDeviceSynchronize();
      std::vector<long long> hc(sms);
      cudaMemcpy(hc.data(), cyc, sms * 8, cudaMemcpyDeviceToHost);
      double mean = 0;
      for (auto c : hc) mean += (double)c / sms;
      printf("  issuers %d: mismatches %d; %.0f cycles per %d-row document, %.1f B/clk/SM (all %d SMs busy)\n", issuers,
             hbad, mean / NDOCS, ROWS, (double)ROWS * COLS * 8 / (mean / NDOCS), sms);
    }
  }
  return 0;
}
