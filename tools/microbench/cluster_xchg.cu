// cluster_xchg.cu — the warp-to-warp norm exchange of the cluster warp-group E-step kernel in
// isolation: warp w of every CTA of a cluster sends one double per lane to warp w of every peer CTA
// (st.async into the peer's shared memory, completing on the peer warp's mbarrier), waits for the
// peers' values and sums the CL values in rank order.  Checks the sums, reports cycles per exchange,
// and whether a cluster size that is not a power of two (5: K = 1000 is 5 x 200) launches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_xchg cluster_xchg.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, int rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async(uint32_t raddr, double v, uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr),
               "l"(__double_as_longlong(v)), "r"(rmbar)
               : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long *m, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(m)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(unsigned long long *m, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(m)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *m, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(m)), "r"(parity)
        : "memory");
}

template <int CL>
__global__ void __launch_bounds__(256, 1) xchg_kernel(int iters, int live, double *out, long long *cycles, int *errors) {
  __shared__ double buf[2][CL][8][32];
  __shared__ unsigned long long mbar[2][8];
  const int crank = (int)cg::this_cluster().block_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    mbar_init(&mbar[0][warp], 1);
    mbar_init(&mbar[1][warp], 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  cg::this_cluster().sync();
  double acc = 0.0;
  int bad = 0;
  const bool sender = (lane & 7) < live;  // only `live` lanes of every 8 carry a value
  const uint32_t bytes = (uint32_t)((CL - 1) * 4 * live * sizeof(double));
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const int par = it & 1;
    const double v = (double)(it * 1000 + crank * 100 + warp * 10) + lane * 0.001 + acc * 1e-30;
    if (lane == 0) mbar_expect(&mbar[par][warp], bytes);
    __syncwarp();
    if (sender) {
      buf[par][crank][warp][lane] = v;
      const uint32_t la = smem_u32(&buf[par][crank][warp][lane]), lm = smem_u32(&mbar[par][warp]);
#pragma unroll
      for (int d = 1; d < CL; d++) {
        const int r = crank + d < CL ? crank + d : crank + d - CL;
        st_async(mapa(la, r), v, mapa(lm, r));
      }
    }
    mbar_wait(&mbar[par][warp], (uint32_t)((it >> 1) & 1));
    double s = 0.0, expect = 0.0;
    if (sender) {
#pragma unroll
      for (int r = 0; r < CL; r++) {
        s += buf[par][r][warp][lane];
        expect += (double)(it * 1000 + r * 100 + warp * 10) + lane * 0.001;
      }
      if (fabs(s - expect) > 1e-6) bad++;
    }
    acc += s;
  }
  const long long t1 = clock64();
  cg::this_cluster().sync();
  if (bad) atomicAdd(errors, bad);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int CL>
static void run(int iters, int live) {
  double *out;
  long long *cyc;
  int *err;
  cudaMalloc(&out, 148 * 256 * sizeof(double));
  cudaMalloc(&cyc, sizeof(long long));
  cudaMalloc(&err, sizeof(int));
  cudaMemset(err, 0, sizeof(int));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL * (148 / CL));
  cfg.blockDim = dim3(256);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int maxc = -1;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, xchg_kernel<CL>, &cfg);
  printf("CL=%d: max active clusters %d (%s)\n", CL, maxc, cudaGetErrorString(e));
  e = cudaLaunchKernelEx(&cfg, xchg_kernel<CL>, iters, live, out, cyc, err);
  cudaError_t e2 = cudaDeviceSynchronize();
  long long hc = 0;
  int he = -1;
  cudaMemcpy(&hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);
  cudaMemcpy(&he, err, sizeof he, cudaMemcpyDeviceToHost);
  printf("CL=%d live=%d: launch %s, sync %s, %.1f cycles per exchange, %d wrong sums\n", CL, live,
         cudaGetErrorString(e), cudaGetErrorString(e2), (double)hc / iters, he);
  cudaFree(out);
  cudaFree(cyc);
  cudaFree(err);
}

int main() {
  run<2>(20000, 8);
  run<4>(20000, 8);
  run<4>(20000, 4);
  run<5>(20000, 8);
  run<5>(20000, 4);
  run<8>(20000, 4);
  return 0;
}
