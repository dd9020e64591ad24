// How fast can one SM fetch the E rows of a document?  (K=200: 208 doubles = 1664 B per row, rows
// picked at random from a 166 MB matrix, 128 rows per document.)  Compares the access patterns the
// E-step kernels use or could use:
//   A  lanes = rows, each lane reads its 26-column slice with 13 LDG.128 (the register-stationary load)
//   B  one warp per row, lanes across columns, LDG.64, 8 rows in flight (the statistics scatter's re-read)
//   C  one warp per row, LDG.128 (26 lanes x 4 chunks)
//   D  TMA bulk copies (cp.async.bulk) of whole rows into shared memory, 32 rows per mbarrier phase
// Reports cycles per 128-row document with every SM busy and with a single CTA.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int ROW = 208, ROWS = 128, NW = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const double *__restrict__ E, const int *__restrict__ rows,
                                            int ndocs, double *out, long long *cyc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double acc = 0.0;
  if (MODE == 3 && tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  unsigned phase = 0;
  long long t0 = clock64();
  for (int d = 0; d < ndocs; d++) {
    const int *rl = rows + ((size_t)blockIdx.x * ndocs + d) * ROWS;
    if (MODE == 0) {
      for (int q = 0; q < ROWS / 32; q++) {
        const double2 *src = reinterpret_cast<const double2 *>(E + (size_t)__ldg(rl + 32 * q + lane) * ROW + warp * 26);
        double2 v[13];
#pragma unroll
        for (int i = 0; i < 13; i++) v[i] = __ldg(src + i);
#pragma unroll
        for (int i = 0; i < 13; i++) acc += v[i].x + v[i].y;
      }
    } else if (MODE == 4) {
      // A2: as A, but 6 x 256-bit + 1 x 128-bit loads per lane (the slice starts 32-byte aligned for
      // even warps, 16 bytes off for odd ones)
      for (int q = 0; q < ROWS / 32; q++) {
        const double *src = E + (size_t)__ldg(rl + 32 * q + lane) * ROW + warp * 26;
        double v[26];
        const int head = (warp & 1) ? 2 : 0;  // doubles before the first 32-byte boundary
        if (head) {
          asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v[24]), "=d"(v[25]) : "l"(src));
        } else {
          asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v[24]), "=d"(v[25]) : "l"(src + 24));
        }
#pragma unroll
        for (int i = 0; i < 6; i++)
          asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                       : "=d"(v[4 * i]), "=d"(v[4 * i + 1]), "=d"(v[4 * i + 2]), "=d"(v[4 * i + 3])
                       : "l"(src + head + 4 * i));
#pragma unroll
        for (int i = 0; i < 26; i++) acc += v[i];
      }
    } else if (MODE == 1) {
      for (int r0 = warp; r0 < ROWS; r0 += 8 * NW) {
        double v[8][7];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const double *e = E + (size_t)__ldg(rl + r0 + j * NW) * ROW;
#pragma unroll
          for (int i = 0; i < 7; i++) v[j][i] = (lane + 32 * i < ROW) ? __ldg(e + lane + 32 * i) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
          for (int i = 0; i < 7; i++) acc += v[j][i];
      }
    } else if (MODE == 2) {
      for (int r0 = warp; r0 < ROWS; r0 += 8 * NW) {
        double2 v[8][4];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const double2 *e = reinterpret_cast<const double2 *>(E + (size_t)__ldg(rl + r0 + j * NW) * ROW);
#pragma unroll
          for (int i = 0; i < 4; i++) v[j][i] = (lane + 32 * i < ROW / 2) ? __ldg(e + lane + 32 * i) : make_double2(0, 0);
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
          for (int i = 0; i < 4; i++) acc += v[j][i].x + v[j][i].y;
      }
    } else {
      // 4 batches of 32 rows; padded row pitch 1680 B keeps the later slice reads conflict-free
      constexpr int PITCH = 1680;
      for (int q = 0; q < ROWS / 32; q++) {
        if (warp == 0) {
          if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(32 * ROW * 8));
          __syncwarp();
          const double *src = E + (size_t)__ldg(rl + 32 * q + lane) * ROW;
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  smem_u32(smem_raw + lane * PITCH)),
              "l"(src), "r"(ROW * 8), "r"(smem_u32(&mbar))
              : "memory");
        }
        unsigned done = 0;
        while (!done)
          asm volatile(
              "{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
              : "=r"(done)
              : "r"(smem_u32(&mbar)), "r"(phase)
              : "memory");
        phase ^= 1;
        // every thread reads its slice (row = lane, 26 columns of warp) like the tile load would
        const double2 *s = reinterpret_cast<const double2 *>(smem_raw + lane * PITCH) + warp * 13;
#pragma unroll
        for (int i = 0; i < 13; i++) acc += s[i].x + s[i].y;
        __syncthreads();  // the staging buffer is reused by the next batch
      }
    }
  }
  long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  out[(size_t)blockIdx.x * 256 + tid] = acc;
}

template <int MODE>
void run(const char *name, const double *E, const int *rows, int grid, int ndocs, size_t smem) {
  double *out;
  long long *cyc;
  cudaMalloc(&out, 148 * 256 * 8);
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; rep++) k<MODE><<<grid, 256, smem>>>(E, rows, ndocs, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
  double s = 0;
  for (long long c : h) s += (double)c;
  printf("%-34s grid=%3d  %8.0f cycles per document  (%.1f B/clk/SM)  %s\n", name, grid, s / grid / ndocs,
         (double)ROWS * ROW * 8 / (s / grid / ndocs), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

// Scatter patterns: add a 200-double row into a V x 200 matrix at random rows, 128 rows per document.
//   E  one warp per row, lanes across columns, REDG.64 (the statistics scatter's atomics)
//   F  rows staged in shared memory, one TMA bulk reduce (cp.reduce.async.bulk .add.f64) per row
template <int MODE>
__global__ void __launch_bounds__(256, 1) ks(double *S, const int *__restrict__ rows, int ndocs, long long *cyc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int PITCH = 1680;
  for (int i = tid; i < 32 * PITCH / 8; i += 256) reinterpret_cast<double *>(smem_raw)[i] = 1.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  long long t0 = clock64();
  for (int d = 0; d < ndocs; d++) {
    const int *rl = rows + ((size_t)blockIdx.x * ndocs + d) * ROWS;
    if (MODE == 0) {
      for (int r = warp; r < ROWS; r += NW) {
        double *dst = S + (size_t)__ldg(rl + r) * 200;
#pragma unroll
        for (int i = 0; i < 7; i++)
          if (lane + 32 * i < 200) atomicAdd(dst + lane + 32 * i, 1.0);
      }
    } else {
      if (warp == 0) {
        for (int q = 0; q < ROWS / 32; q++) {
          double *dst = S + (size_t)__ldg(rl + 32 * q + lane) * 200;
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst),
                       "r"(smem_u32(smem_raw + lane * PITCH)), "r"(1600)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
    __syncthreads();
  }
  if (MODE == 1 && warp == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void runs(const char *name, double *S, const int *rows, int grid, int ndocs, size_t smem) {
  long long *cyc;
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(ks<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; rep++) ks<MODE><<<grid, 256, smem>>>(S, rows, ndocs, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
  double s = 0;
  for (long long c : h) s += (double)c;
  printf("%-34s grid=%3d  %8.0f cycles per document  (%.1f B/clk/SM)  %s\n", name, grid, s / grid / ndocs,
         (double)ROWS * 1600 / (s / grid / ndocs), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(cyc);
}

int main() {
  const int V = 100000, ndocs = 64;
  double *E;
  cudaMalloc(&E, (size_t)V * ROW * 8);
  cudaMemset(E, 0, (size_t)V * ROW * 8);
  for (int hot = 0; hot < 2; hot++) {
    // hot = 1: rows from a 12k-row (20 MB) subset, L2 resident; hot = 0: uniform over 166 MB
    std::vector<int> h((size_t)148 * ndocs * ROWS);
    srand(1);
    for (auto &x : h) x = (int)(((long long)rand() * 7919 + rand()) % (hot ? 12000 : V));
    int *rows;
    cudaMalloc(&rows, h.size() * 4);
    cudaMemcpy(rows, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    printf("== rows drawn from %s\n", hot ? "a 20 MB subset (L2 hits)" : "the whole 166 MB matrix");
    for (int grid : {148, 1}) {
      const size_t big = 200 * 1024;  // same carve-out as the E-step kernel: L1 is small
      run<0>("A lanes=rows, 13 x LDG.128", E, rows, grid, ndocs, big);
      run<4>("A2 lanes=rows, 6 x LDG.256 + 1", E, rows, grid, ndocs, big);
      run<1>("B warp per row, LDG.64 x7, 8 rows", E, rows, grid, ndocs, big);
      run<2>("C warp per row, LDG.128 x4, 8 rows", E, rows, grid, ndocs, big);
      run<3>("D TMA bulk rows, 32 per phase", E, rows, grid, ndocs, big);
      runs<0>("E scatter: warp per row, REDG.64", E, rows, grid, ndocs, big);
      runs<1>("F scatter: TMA bulk reduce per row", E, rows, grid, ndocs, big);
    }
    cudaFree(rows);
  }
  return 0;
}
