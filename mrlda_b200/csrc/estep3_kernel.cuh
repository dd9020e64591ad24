// estep3_kernel.cuh — K1, warp-group variant: several documents in flight per SM.
//
// Same arithmetic as estep_kernel.cuh / estep2_kernel.cuh (linear-space fixed point, 99 sweeps, exact
// log-space redo of underflowing rows; reference: DocumentMapper.java:175-350, updatePhi :402-429).
// What changes is who owns what:
//
//   * A CTA is NW warps on one SM (8, 12 or 16: 255, 168 or 128 registers per thread).  A document
//     is owned by a GROUP of G consecutive warps (G divides NW, a launch parameter), so NW/G
//     documents are in flight per SM and one document's serial phases (S exchange, exp(digamma),
//     barriers) overlap another document's DFMA phases on the same FP64 pipes.  Groups
//     synchronise on named barriers (bar.sync id, 32*G) and pull documents from the work queue
//     independently of each other.
//   * Inside a group ROWS are split across warps: warp k owns a contiguous range of row batches.
//     A batch is RPB = 32/L rows; L lanes share a row, lane j of a row owns the KPL columns
//     {2(j + L i), 2(j + L i) + 1}, i < KPL/2 (same column interleave as estep_kernel.cuh, so row
//     loads and statistics atomics touch 16-byte units that are contiguous across the L lanes)
//     and, when KPL is odd, the single column 2 L (KPL/2) + j: K = 200 is 8 lanes x 25 columns
//     with no padding column, and a batch is 50 Tensor Memory words, so that five fit a warp's
//     256 columns.
//     Every warp therefore holds COMPLETE rows: norm_w needs only an L-lane butterfly (3 shuffle
//     levels at K = 200) and r_w = n_w / norm_w is used by the same lanes right away — no CTA
//     barrier and no shared-memory exchange per row.  S_c accumulates per lane over the warp's rows.
//   * A warp's batches live, in this order, in REGISTERS (the first R batches), in TENSOR MEMORY
//     (the next T batches: thread-private tcgen05.st / tcgen05.ld, no MMA; the NW/4 warps that
//     share a lane quarter split the 512 columns), and in SHARED MEMORY (the last M batches).  A
//     shared-memory batch is the RPB x KS row-major tile that one TMA tile::gather4 copy
//     (UTMALDG.2D.GATHER4: four rows of the term-major E matrix picked by term id) writes: lane 0
//     of the warp issues the copies, they complete on the warp's mbarrier while all 32 lanes load
//     the register and Tensor Memory batches; the L lanes of a row read 16 L contiguous bytes per
//     LDS.128, so the row-major tile is bank-conflict free.
//   * Per sweep and warp.  Pass A: theta in registers; partial dot products of every batch (TMEM
//     batches are fetched in two parts so that one tcgen05.ld is always in flight behind the
//     DFMAs).  Then the L-lane butterflies and the reciprocals of ALL batches as independent,
//     interleaved chains.  Pass B: theta is dead, S takes its registers; S += r E over every batch
//     (TMEM and shared-memory batches are read a second time).  The warp then publishes its
//     RPB x K partial sums to shared memory as they are (no shuffle reduction), group barrier,
//     the group's threads own one or two columns each: sum the RPB*G partials,
//     gamma_c = alpha_c + theta_c S_c, theta_c = exp(digamma(gamma_c)), group barrier.
//   * The 98 sweeps that are not the last one run in a loop body specialised on the number of
//     TMEM and shared-memory batches the warp holds (chosen once per document), so the loop is
//     straight-line code: with two to four warps per scheduler and in-order issue, every branch,
//     jump table and dependent load in the sweep shows up in the run time.  The last sweep
//     (likelihood terms, statistics scatter) is generic code.
//   * Padding rows (a partially filled last batch, unused register batches) hold a copy of the
//     document's last row with n_w = 0: their norm is a real norm and their r is 0, so the sweep
//     needs no validity test per row; one integer minimum over the norms' high words per sweep
//     detects underflow.
//   * The statistics scatter of the last sweep is issued straight from the tile (r theta E), one
//     RED.64 per element (DocumentMapper.java:315-329).
#pragma once
#include "estep_kernel.cuh"
#include "tmem_ops.cuh"
#include "estep2_kernel.cuh"  // tmem_ld_words / tmem_st_words
#include "estep3_xchg.cuh"    // cluster variant (estep3c_kernel.cuh): norm exchange between CTAs

namespace mrlda {

template <int L, int KPL, int R, int NW>
struct Estep3Shape {
  static constexpr int RPB = 32 / L;
  static constexpr int KS = L * KPL;
  static constexpr int HP = KPL / 2;                   // column pairs of a lane ...
  static constexpr int ODD = KPL & 1;                  // ... and the single column of an odd KPL
  static constexpr int TW = 2 * KPL;                   // 32-bit TMEM columns per batch and lane
  static constexpr int WQ = NW / 4;                    // warps that share a TMEM lane quarter
  static constexpr int TWORDS = (512 / WQ) / 4 * 4;    // TMEM columns of one warp
  static constexpr int T = TWORDS / TW;                // TMEM batches per warp
  static constexpr int MMAX = 3;                       // upper bound of the shared-memory batches per warp
  static constexpr int NBMAX = R + T + MMAX;
  static constexpr int CA = 16;                        // columns of a TMEM batch fetched by the first tcgen05.ld
  static constexpr int CB = KPL - CA;                  // ... and by the second
  static constexpr int RA = R > 0 ? R : 1;             // extent of the register-batch array
  static constexpr int THREADS = NW * 32;
  static_assert(KPL > CA && KPL <= 32, "KPL in (16, 32]");
  static_assert(L == 1 || L == 2 || L == 4 || L == 8 || L == 16 || L == 32, "L must divide 32");
  static_assert(NW == 8 || NW == 12 || NW == 16, "NW in {8, 12, 16}");
};

// the split of a warp's NA batches into the two halves of a sweep: the first one ends with or after
// the register batches (their dot products are computed together)
template <int R, int NA>
struct Estep3Halves {
  static constexpr int HALFUP = (NA + 1) / 2;
  static constexpr int N1 = NA <= R ? NA : (HALFUP > R ? HALFUP : R);
};

// shared-memory carve-up (offsets in doubles) for M shared-memory batches per warp, groups of G warps
template <int L, int KPL, int R, int NW>
struct Estep3Smem {
  using S = Estep3Shape<L, KPL, R, NW>;
  size_t tile, spart, theta, gam, gextra, alpha, ssacc, cnt, red, meta, mbar, total;
  __host__ __device__ Estep3Smem(int M, int G) {
    const int NG = NW / G;
    size_t o = 0;
    tile = o;   o += (size_t)NW * M * 32 * KPL;
    spart = o;  o += (size_t)NW * S::KS;      // every warp's column sums of S: [warp][column]
    theta = o;  o += (size_t)NG * S::KS;
    gam = o;    o += (size_t)NG * S::KS;
    gextra = o; o += (size_t)NG * S::KS;
    alpha = o;  o += S::KS;
    ssacc = o;  o += S::KS;
    cnt = o;    o += (size_t)NW * S::NBMAX * S::RPB;  // n_w of every batch slot (0: padding)
    red = o;    o += NW * 4;
    meta = o;   o += 10;  // 20 ints: [0] TMEM base, [1 + group] document
    mbar = o;   o += 2 * NW;  // per warp: the mbarrier its TMA row gathers complete on, and its phase parity
    total = o * sizeof(double);
  }
};

__device__ __forceinline__ void estep3_group_sync(const int id, const int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ double estep3_w2d(const uint32_t lo, const uint32_t hi) {
  return __hiloint2double((int)hi, (int)lo);
}

// A lane's KPL values of a K-contiguous row (global or shared memory): HP 16-byte pairs and, for an
// odd KPL, one more double.
template <int L, int KPL, bool GLOBAL>
__device__ __forceinline__ void estep3_load_cols(const double *row, const int j, double (&v)[KPL]) {
  constexpr int HP = KPL / 2;
  const double2 *src = reinterpret_cast<const double2 *>(row) + j;
#pragma unroll
  for (int i = 0; i < HP; i++) {
    const double2 t2 = GLOBAL ? __ldg(src + L * i) : src[L * i];
    v[2 * i] = t2.x;
    v[2 * i + 1] = t2.y;
  }
  if constexpr (KPL & 1) v[KPL - 1] = GLOBAL ? __ldg(row + 2 * L * HP + j) : row[2 * L * HP + j];
}
// A shared-memory batch is the [RPB][KS] row-major tile a TMA gather writes; lane (g, j) reads its
// KPL columns of row g exactly as it reads a row of E or theta.
template <int L, int KPL>
__device__ __forceinline__ void estep3_tile_ld(const double *batch, const int g, const int j, double (&v)[KPL]) {
  estep3_load_cols<L, KPL, false>(batch + (size_t)g * (L * KPL), j, v);
}

// mbarrier + TMA tile::gather4 (sm_100a): four rows of the 2-D tensor behind `tmap` (the term-major E
// matrix, box = KS columns x 1 row), picked by their row indices, land as a [4][KS] tile at dst.
__device__ __forceinline__ uint32_t estep3_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void estep3_mbar_init(unsigned long long *mbar, const int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(estep3_smem_u32(mbar)), "r"(count));
}
__device__ __forceinline__ void estep3_mbar_expect(unsigned long long *mbar, const uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(estep3_smem_u32(mbar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void estep3_mbar_wait(unsigned long long *mbar, const uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(estep3_smem_u32(mbar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void estep3_tma_gather4(double *dst, const void *tmap, unsigned long long *mbar,
                                                   const int r0, const int r1, const int r2, const int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, "
      "%4, %5, %6}], [%7];" ::"r"(estep3_smem_u32(dst)),
      "l"(tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(estep3_smem_u32(mbar))
      : "memory");
}

// n_w phi_wk = r_w theta_k E_wk into the K x V statistics (DocumentMapper.java:315-329)
template <int L, int KPL>
__device__ __forceinline__ void estep3_scatter(double *srow, const double *theta_g, const double r,
                                               const double (&ev)[KPL], const int j, const int K) {
  constexpr int HP = KPL / 2;
  const double2 *theta2 = reinterpret_cast<const double2 *>(theta_g);
#pragma unroll
  for (int i = 0; i < HP; i++) {
    const double2 tv = theta2[j + L * i];
    const int col = 2 * (j + L * i);
    if (col < K) atomicAdd(srow + col, r * tv.x * ev[2 * i]);
    if (col + 1 < K) atomicAdd(srow + col + 1, r * tv.y * ev[2 * i + 1]);
  }
  if constexpr (KPL & 1) {
    const int col = 2 * L * HP + j;
    if (col < K) atomicAdd(srow + col, r * theta_g[col] * ev[KPL - 1]);
  }
}

// Exact log-space redo (updatePhi as written, DocumentMapper.java:402-429) of this warp's rows whose
// norm underflowed in the linear-space form.  Returns the lane's share of likelihoodPhi on the last
// sweep.  Rare (no row of the BASELINE configurations takes it): out of line.
template <int L>
static __device__ __noinline__ double estep3_slow_rows(const int32_t *term_id, const int32_t *count,
                                                       const double *logbeta, double *stats,
                                                       int *slow_rows, const double *gam_g,
                                                       double *gextra_g, const int K, const int nb,
                                                       const unsigned badbits, const int lane,
                                                       const bool last) {
  constexpr int RPB = 32 / L;
  double lik = 0.0;
  for (int b = 0; b < nb; b++) {
    unsigned m = __ballot_sync(0xffffffffu, (badbits >> b) & 1u);
    while (m) {
      const int gi = (__ffs(m) - 1) / L;
      m &= ~((L == 32 ? 0xffffffffu : ((1u << L) - 1u)) << (gi * L));
      const int z = b * RPB + gi;
      const int term = __ldg(term_id + z);
      const double nw_ = (double)__ldg(count + z);
      const double *lb = logbeta + (size_t)(term - 1) * K;
      double mx = -INFINITY;
      for (int k = lane; k < K; k += 32) mx = fmax(mx, lb[k] + digamma_fast(gam_g[k]));
      mx = warp_max(mx);
      double sum = 0.0;
      for (int k = lane; k < K; k += 32) sum += exp(lb[k] + digamma_fast(gam_g[k]) - mx);
      sum = warp_sum(sum);
      const double lsum = log(sum);
      for (int k = lane; k < K; k += 32) {
        const double lphi = lb[k] + digamma_fast(gam_g[k]) - mx - lsum;
        const double c = nw_ * exp(lphi);
        if (c > 0.0) {
          atomicAdd(gextra_g + k, c);  // other warps of the group may add to the same topic
          if (last) {
            lik += c * (lb[k] - lphi);  // DocumentMapper.java:421
            if (stats) atomicAdd(stats + (size_t)(term - 1) * K + k, c);
          }
        }
      }
      if (lane == 0) atomicAdd(slow_rows, 1);
    }
  }
  return lik;
}

// Exact log-space redo of this warp's underflowed rows (estep3_slow_rows with the topics of the
// other CTAs read through distributed shared memory; contributions go to this CTA's topics only).
// Every CTA of the cluster takes the same decisions (identical norms), so the same warp of every CTA
// is here; they meet on the warp's slow-path mbarrier before anybody's gamma can change.
template <int L, int KS, int CL>
static __device__ __noinline__ double estep3c_slow_rows(const int32_t *term_id, const int32_t *count,
                                                        const double *logbeta, double *stats, int *slow_rows,
                                                        double *gam_g, double *gextra_g, const int K,
                                                        const int nb, const unsigned badbits, const int lane,
                                                        const bool last, const int crank, const uint32_t xbar_u32,
                                                        unsigned long long *xbar_w) {
  namespace cg = cooperative_groups;
  constexpr int RPB = 32 / L;
  auto gamma_of = [&](int k) -> double { return cg::this_cluster().map_shared_rank(gam_g, k / KS)[k % KS]; };
  double lik = 0.0;
  for (int b = 0; b < nb; b++) {
    unsigned m = __ballot_sync(0xffffffffu, (badbits >> b) & 1u);
    while (m) {
      const int gi = (__ffs(m) - 1) / L;
      m &= ~((L == 32 ? 0xffffffffu : ((1u << L) - 1u)) << (gi * L));
      const int z = b * RPB + gi;
      const int term = __ldg(term_id + z);
      const double nw_ = (double)__ldg(count + z);
      const double *lb = logbeta + (size_t)(term - 1) * K;
      double mx = -INFINITY;
      for (int k = lane; k < K; k += 32) mx = fmax(mx, lb[k] + digamma_fast(gamma_of(k)));
      mx = warp_max(mx);
      double sum = 0.0;
      for (int k = lane; k < K; k += 32) sum += exp(lb[k] + digamma_fast(gamma_of(k)) - mx);
      sum = warp_sum(sum);
      const double lsum = log(sum);
      for (int k = lane; k < K; k += 32) {
        if (k / KS != crank) continue;
        const double lphi = lb[k] + digamma_fast(gamma_of(k)) - mx - lsum;
        const double c = nw_ * exp(lphi);
        if (c > 0.0) {
          atomicAdd(gextra_g + (k - crank * KS), c);
          if (last) {
            lik += c * (lb[k] - lphi);  // DocumentMapper.java:421
            if (stats) atomicAdd(stats + (size_t)(term - 1) * K + k, c);
          }
        }
      }
      if (lane == 0 && crank == 0) atomicAdd(slow_rows, 1);
    }
  }
  // rendezvous of the CL warps that read each other's gamma
  __syncwarp();
  const uint32_t sb = xbar_u32 + 4u * (uint32_t)sizeof(unsigned long long);
  if (lane == 0) {
#pragma unroll
    for (int d = 1; d < CL; d++) {
      const int r = crank + d < CL ? crank + d : crank + d - CL;
      estep3c_arrive_remote(estep3c_mapa(sb, r));
    }
  }
  const uint32_t parity = (uint32_t)xbar_w[5];
  estep3c_mbar_wait(sb, parity);
  __syncwarp();
  if (lane == 0) xbar_w[5] = parity ^ 1u;
  __syncwarp();
  return lik;
}

// Everything a warp needs to walk its batches (built once per kernel).
template <int L, int KPL, int R, int NW>
struct Estep3Warp {
  uint32_t tbase;          // this warp's TMEM slots
  const double *tile_w;    // this warp's shared-memory batches
  const double *theta_g;   // the group's theta
  const double *cnt_w;     // n_w of this warp's batch slots (0 for padding rows)
  int g, j;                // row of the batch / column group this lane works on
};

// Norms and r = n_w / norm_w of the batches [START, START + NSV) of a warp (NSV <= L).  The L lanes
// of a row hold L partial sums for each batch.  A reduce-scatter (lane bit b keeps the odd or the
// even half and sends the other one) leaves lane j with the complete norm of batch START + j after
// NSV (1 - 1/L) exchanges instead of NSV log2(L); every lane then needs ONE reciprocal, and the NSV
// values of r come back with one shuffle each.  mycnt is n_w of the (batch, row) this lane
// completes (0 for padding rows and for lanes j >= NSV).
template <int L, int NA, int START, int NSV, int RT, int NTM, int T>
__device__ __forceinline__ void estep3_norm_exchange(const double (&pr)[NA], double (&rr)[NA], const double mycnt,
                                                     const int j, const int lane, unsigned &badbits) {
  static_assert(NSV >= 1 && NSV <= L && START + NSV <= NA, "one batch per lane of a row");
  double q[L];
#pragma unroll
  for (int i = 0; i < L; i++) q[i] = i < NSV ? pr[START + (i < NSV ? i : 0)] : 1.0;
#pragma unroll
  for (int o = 1; o < L; o <<= 1) {
    const bool up = (j & o) != 0;
#pragma unroll
    for (int m = 0; m < L / (2 * o); m++) {
      if (2 * m * o < NSV) {  // entry 2m still carries live batches (compile-time after unrolling)
        const double send = up ? q[2 * m] : q[2 * m + 1];
        const double keep = up ? q[2 * m + 1] : q[2 * m];
        q[m] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
  }
  const double nrm = q[0];  // norm of batch START + j, row g (lanes j >= NSV hold a dummy)
  const bool live = mycnt > 0.0;
  const bool ok = nrm >= MRLDA_NORM_TINY;
  const double r = (live && ok) ? mycnt * rcp_fast(nrm) : 0.0;
  const int base = lane & ~(L - 1);
#pragma unroll
  for (int i = 0; i < NSV; i++) rr[START + i] = __shfl_sync(0xffffffffu, r, base + i);
  const unsigned bm = __ballot_sync(0xffffffffu, live && !ok);
  if (bm) {  // some norm underflowed (rare): those rows contribute through the exact log-space redo
    const unsigned mine = bm >> base;
#pragma unroll
    for (int i = 0; i < NSV; i++)
      if ((mine >> i) & 1u) badbits |= 1u << (START + i < RT + NTM ? START + i : START + i - NTM + T);
  }
}

// One of the sweeps 1..I-1 over the batches of this warp: NTM TMEM and NSM shared-memory batches
// in use (compile time), so the whole sweep is straight-line code.  Returns S in s, the underflow
// flags in badbits.  The batches are handled as two halves, [0, N1) and [N1, NA): the norm exchange
// of the first half (shuffles, one reciprocal: a long dependent chain) is in flight behind the dot
// products of the second half, and the exchange of the second half behind pass B of the first.
//
// CL > 1 (estep3c_kernel.cuh: the topics are split across a cluster of CL CTAs and this CTA's dot
// products are partial): the L-lane reduce-scatter of a half leaves lane j with the CTA's partial
// norm, which goes to the same lane of the same warp in every peer CTA (estep3c_send); both halves
// are sent as soon as they are known, the warp waits once on its exchange mbarrier and sums the CL
// partials in rank order.  xc is the warp's exchange counter: buffer and mbarrier xc & 1, mbarrier
// phase (xc >> 1) & 1.
template <int L, int KPL, int R, int NW, int NTM, int NSM, int CL = 1>
__device__ __forceinline__ void estep3_sweep_fast(const Estep3Warp<L, KPL, R, NW> &w,
                                                  const double (&e)[Estep3Shape<L, KPL, R, NW>::RA][KPL],
                                                  const double mycnt1, const double mycnt2, const int lane,
                                                  double (&s)[KPL], unsigned &badbits, long long (&tm)[8],
                                                  const Estep3cX &x, const int xc) {
  using S = Estep3Shape<L, KPL, R, NW>;
  constexpr int T = S::T, TW = S::TW, CA = S::CA, CB = S::CB;
  constexpr int NA = R + NTM + NSM;  // batches in use; batch i < R + NTM sits in slot i, the others in slot i - NTM + T
  constexpr int N1 = Estep3Halves<R, NA>::N1, N2 = NA - N1;
  static_assert(NA > 0 && N1 <= L && N2 <= L, "a warp holds between one and 2 L batches");
  double pr[NA];  // partial dot products -> norms
  double rr[NA];  // r_w = n_w / norm_w
  const int j = w.j;
  uint32_t wa[2 * CA], wb[2 * CB];
  badbits = 0;
  const int par = xc & 1;
  // first half of the norms: complete them (CL = 1) or send this CTA's partials (CL > 1)
  auto first_half = [&]() {
    if constexpr (CL == 1) {
      estep3_norm_exchange<L, NA, 0, N1, R, NTM, T>(pr, rr, mycnt1, j, lane, badbits);
    } else {
      const double p1 = estep3c_norm_rs<L, NA, 0, N1>(pr, j);
      if (j < N1) estep3c_send<CL>(x, par, lane, p1);
    }
  };
  if constexpr (CL > 1) {
    if (lane == 0)
      estep3c_mbar_expect(x.xbar_u32 + (uint32_t)(par * sizeof(unsigned long long)),
                          (uint32_t)((CL - 1) * S::RPB * NA * sizeof(double)));
  }
  // ---- pass A: theta in registers; partial dot products of every batch ----
  {
    double th[KPL];
    estep3_load_cols<L, KPL, false>(w.theta_g, j, th);
    MRLDA_TICK_DEP(1, th[KPL - 1]);
    if constexpr (NTM > 0) tmem_ld_words<2 * CA>(w.tbase, wa);  // in flight behind the register batches
    if constexpr (R > 0) {
      double acc[S::RA][4];
#pragma unroll
      for (int q = 0; q < R; q++) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0;
#pragma unroll
      for (int c = 0; c < KPL; c++) {
#pragma unroll
        for (int q = 0; q < R; q++) acc[q][c & 3] = fma(th[c], e[q][c], acc[q][c & 3]);
      }
#pragma unroll
      for (int q = 0; q < R; q++) pr[q] = (acc[q][0] + acc[q][1]) + (acc[q][2] + acc[q][3]);
      if constexpr (N1 == R) first_half();
    }
#pragma unroll
    for (int t = 0; t < NTM; t++) {
      tmem_wait_ld();
      tmem_ld_words<2 * CB>(w.tbase + (uint32_t)(t * TW + 2 * CA), wb);
      double ac[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int c = 0; c < CA; c++) ac[c & 3] = fma(th[c], estep3_w2d(wa[2 * c], wa[2 * c + 1]), ac[c & 3]);
      tmem_wait_ld();
      if (t + 1 < NTM) tmem_ld_words<2 * CA>(w.tbase + (uint32_t)((t + 1) * TW), wa);
#pragma unroll
      for (int c = 0; c < CB; c++)
        ac[c & 3] = fma(th[CA + c], estep3_w2d(wb[2 * c], wb[2 * c + 1]), ac[c & 3]);
      pr[R + t] = (ac[0] + ac[1]) + (ac[2] + ac[3]);
      if (R + t == N1 - 1 && N1 > R) first_half();
    }
#pragma unroll
    for (int m = 0; m < NSM; m++) {
      double ev[KPL];
      estep3_tile_ld<L, KPL>(w.tile_w + (size_t)m * 32 * KPL, w.g, j, ev);
      double ac[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int c = 0; c < KPL; c++) ac[c & 3] = fma(th[c], ev[c], ac[c & 3]);
      pr[R + NTM + m] = (ac[0] + ac[1]) + (ac[2] + ac[3]);
      if (R + NTM + m == N1 - 1 && N1 > R) first_half();
    }
  }
  MRLDA_TICK_DEP(2, pr[NA - 1]);
  if constexpr (NTM > 0) tmem_ld_words<2 * CA>(w.tbase, wa);  // pass B's first load, behind the chains
  if constexpr (CL == 1) {
    if constexpr (N2 > 0)
      estep3_norm_exchange<L, NA, N1, (N2 > 0 ? N2 : 1), R, NTM, T>(pr, rr, mycnt2, j, lane, badbits);
  } else {
    if constexpr (N2 > 0) {
      const double p2 = estep3c_norm_rs<L, NA, N1, (N2 > 0 ? N2 : 1)>(pr, j);
      if (j < N2) estep3c_send<CL>(x, par, 32 + lane, p2);
    }
    // the peers' partial norms of both halves
    estep3c_mbar_wait(x.xbar_u32 + (uint32_t)(par * sizeof(unsigned long long)), (uint32_t)((xc >> 1) & 1));
    const double g1 = j < N1 ? estep3c_total<CL>(x, par, lane) : 1.0;
    double g2 = 1.0;
    if constexpr (N2 > 0) g2 = j < N2 ? estep3c_total<CL>(x, par, 32 + lane) : 1.0;
    estep3c_norm_finish<L, NA, 0, N1, R, NTM, T>(g1, rr, mycnt1, lane, badbits);
    if constexpr (N2 > 0) estep3c_norm_finish<L, NA, N1, (N2 > 0 ? N2 : 1), R, NTM, T>(g2, rr, mycnt2, lane, badbits);
  }
  MRLDA_TICK_DEP(3, rr[0]);
  // ---- pass B: S_c += r_w E_wc ----
#pragma unroll
  for (int c = 0; c < KPL; c++) s[c] = 0.0;
#pragma unroll
  for (int q = 0; q < R; q++) {
#pragma unroll
    for (int c = 0; c < KPL; c++) s[c] = fma(rr[q], e[q][c], s[c]);
  }
#pragma unroll
  for (int t = 0; t < NTM; t++) {
    const double r = rr[R + t];
    tmem_wait_ld();
    tmem_ld_words<2 * CB>(w.tbase + (uint32_t)(t * TW + 2 * CA), wb);
#pragma unroll
    for (int c = 0; c < CA; c++) s[c] = fma(r, estep3_w2d(wa[2 * c], wa[2 * c + 1]), s[c]);
    tmem_wait_ld();
    if (t + 1 < NTM) tmem_ld_words<2 * CA>(w.tbase + (uint32_t)((t + 1) * TW), wa);
#pragma unroll
    for (int c = 0; c < CB; c++) s[CA + c] = fma(r, estep3_w2d(wb[2 * c], wb[2 * c + 1]), s[CA + c]);
  }
#pragma unroll
  for (int m = 0; m < NSM; m++) {
    double ev[KPL];
    estep3_tile_ld<L, KPL>(w.tile_w + (size_t)m * 32 * KPL, w.g, j, ev);
    const double r = rr[R + NTM + m];
#pragma unroll
    for (int c = 0; c < KPL; c++) s[c] = fma(r, ev[c], s[c]);
  }
}

// The last sweep (generic over the batch count): one fused pass per batch with the likelihood terms
// (DocumentMapper.java:421) and the statistics scatter (:315-329).
template <int L, int KPL>
__device__ __forceinline__ void estep3_last_batch(const EstepArgs &a, const double *theta_g, const double *cnt_w,
                                                  const double (&ev)[KPL], const int slot, const int64_t zrow0,
                                                  const int g, const int j, double (&s)[KPL], double &lik,
                                                  unsigned &badbits) {
  constexpr int RPB = 32 / L, HP = KPL / 2;
  const double2 *theta2 = reinterpret_cast<const double2 *>(theta_g);
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int i = 0; i < HP; i++) {
    const double2 tv = theta2[j + L * i];
    a0 = fma(tv.x, ev[2 * i], a0);
    a1 = fma(tv.y, ev[2 * i + 1], a1);
  }
  if constexpr (KPL & 1) a0 = fma(theta_g[2 * L * HP + j], ev[KPL - 1], a0);
  double p = a0 + a1;
#pragma unroll
  for (int o = 1; o < L; o <<= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
  const double cn = cnt_w[slot * RPB + g];
  const bool ok = p >= MRLDA_NORM_TINY;
  if (cn > 0.0 && !ok) badbits |= 1u << slot;
  const double r = (cn > 0.0 && ok) ? cn * rcp_fast(ok ? p : 1.0) : 0.0;
#pragma unroll
  for (int c = 0; c < KPL; c++) s[c] = fma(r, ev[c], s[c]);
  if (r > 0.0) {
    const int term = __ldg(a.term_id + zrow0 + slot * RPB + g);
    if (j == 0) lik = fma(cn, log(p) + __ldg(a.rowmax + term - 1), lik);
    if (a.learning) estep3_scatter<L, KPL>(a.stats + (size_t)(term - 1) * a.K, theta_g, r, ev, j, a.K);
  }
}
template <int L, int KPL, int R, int NW>
__device__ __forceinline__ void estep3_sweep_last(const EstepArgs &a, const Estep3Warp<L, KPL, R, NW> &w,
                                                  const double (&e)[Estep3Shape<L, KPL, R, NW>::RA][KPL],
                                                  double (&s)[KPL], double &lik, unsigned &badbits,
                                                  const int nb, const int64_t zrow0) {
  using S = Estep3Shape<L, KPL, R, NW>;
  constexpr int T = S::T, TW = S::TW;
  badbits = 0;
#pragma unroll
  for (int c = 0; c < KPL; c++) s[c] = 0.0;
#pragma unroll
  for (int q = 0; q < R; q++)
    estep3_last_batch<L, KPL>(a, w.theta_g, w.cnt_w, e[q], q, zrow0, w.g, w.j, s, lik, badbits);
#pragma unroll 1
  for (int t = 0; R + t < nb && t < T; t++) {
    uint32_t w32[TW];
    tmem_ld_words<TW>(w.tbase + (uint32_t)t * TW, w32);
    tmem_wait_ld();
    double ev[KPL];
#pragma unroll
    for (int c = 0; c < KPL; c++) ev[c] = estep3_w2d(w32[2 * c], w32[2 * c + 1]);
    estep3_last_batch<L, KPL>(a, w.theta_g, w.cnt_w, ev, R + t, zrow0, w.g, w.j, s, lik, badbits);
  }
#pragma unroll 1
  for (int m = 0; R + T + m < nb; m++) {
    double ev[KPL];
    estep3_tile_ld<L, KPL>(w.tile_w + (size_t)m * 32 * KPL, w.g, w.j, ev);
    estep3_last_batch<L, KPL>(a, w.theta_g, w.cnt_w, ev, R + T + m, zrow0, w.g, w.j, s, lik, badbits);
  }
}

// S exchange, first half.  After pass B lane (g, j) holds the partial sums of its 26 columns over
// its own rows.  A two-level reduce-scatter over the warp's four row groups (lane bits 4 and 3)
// leaves every lane with the WARP's sums of 7 or 6 of those columns: level one keeps the x or the y
// element of each column pair, level two the lower or the upper half of the pairs.  20 exchanges
// instead of 26 stores per lane, and the column owners then read G values per column, not 4 G.
template <int L, int KPL>
__device__ __forceinline__ void estep3_publish_s(const double (&s)[KPL], double *wtot_w, const int lane) {
  static_assert(L == 8, "the row-group reduce-scatter is written for four row groups per warp");
  constexpr int HP = KPL / 2, H2 = HP / 2, LO = HP - H2;
  const bool up1 = (lane & 16) != 0, up2 = (lane & 8) != 0;
  if constexpr (KPL & 1) {  // the single column of an odd KPL: plain butterfly over the row groups
    double t = s[KPL - 1];
    t += __shfl_xor_sync(0xffffffffu, t, 16);
    t += __shfl_xor_sync(0xffffffffu, t, 8);
    if (lane < L) wtot_w[2 * L * HP + lane] = t;
  }
  double v[HP];
#pragma unroll
  for (int m = 0; m < HP; m++) {
    const double send = up1 ? s[2 * m] : s[2 * m + 1];
    const double keep = up1 ? s[2 * m + 1] : s[2 * m];
    v[m] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  double f[LO];
#pragma unroll
  for (int m = 0; m < H2; m++) {
    const double send = up2 ? v[m] : v[LO + m];
    const double keep = up2 ? v[LO + m] : v[m];
    f[m] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  if constexpr (LO > H2) f[H2] = v[H2] + __shfl_xor_sync(0xffffffffu, v[H2], 8);  // lower half only
  // value m of this lane is column 2 (j + L (m + (up2 ? LO : 0))) + (up1 ? 1 : 0)
  double *dst = wtot_w + 2 * (lane & (L - 1)) + (up1 ? 1 : 0) + (up2 ? 2 * L * LO : 0);
#pragma unroll
  for (int m = 0; m < H2; m++) dst[2 * L * m] = f[m];
  if constexpr (LO > H2) {
    if (!up2) dst[2 * L * H2] = f[H2];
  }
}

// S exchange, second half: the sum of the G warp sums of column col
template <int KS, int G>
__device__ __forceinline__ double estep3_colsum(const double *wtot, const int w0, const int col) {
  const double *p = wtot + (size_t)w0 * KS + col;
  double v[G];
#pragma unroll
  for (int w2 = 0; w2 < G; w2++) v[w2] = p[(size_t)w2 * KS];
#pragma unroll
  for (int n = G / 2; n >= 1; n >>= 1) {
#pragma unroll
    for (int x = 0; x < n; x++) v[x] += v[x + n];
  }
  return v[0];
}

// Per-group state in shared memory and the launch geometry of the group this thread belongs to.
struct Estep3Group {
  double *spart, *theta_g, *gam_g, *gextra_g, *alpha_s;
  int w0, gt, bar_id, K;
};

// column owners after a sweep that is not the last: gamma_c = alpha_c + theta_c S_c
// (DocumentMapper.java:232-234) and the next theta_c = exp(digamma(gamma_c)).  A thread owns the
// columns gt, gt + 32 G, ...; NW_ of them reach into [0, K) for this warp (warp-uniform).  They are
// evaluated as interleaved chains (exp_digamma_n).
template <int L, int KPL, int G, int NW_>
__device__ __forceinline__ void estep3_update_gamma_n(const Estep3Group &q) {
  constexpr int GT = 32 * G, KS = L * KPL;
  double gv[NW_], tv[NW_];
#pragma unroll
  for (int x = 0; x < NW_; x++) {
    const int col = q.gt + GT * x;
    const int cc = col < q.K ? col : q.K - 1;
    const double sk = estep3_colsum<KS, G>(q.spart, q.w0, cc);
    gv[x] = fma(q.theta_g[cc], sk, q.alpha_s[cc]) + q.gextra_g[cc];
  }
  exp_digamma_n<NW_>(gv, tv);
#pragma unroll
  for (int x = 0; x < NW_; x++) {
    const int col = q.gt + GT * x;
    if (col < q.K) {
      q.gextra_g[col] = 0.0;
      q.gam_g[col] = gv[x];
      q.theta_g[col] = tv[x];
    }
  }
}
template <int L, int KPL, int G, int NC_>
__device__ __forceinline__ void estep3_update_gamma_pick(const Estep3Group &q, const int ncols) {
  if constexpr (NC_ <= 1) {
    estep3_update_gamma_n<L, KPL, G, 1>(q);
  } else {
    if (ncols >= NC_)
      estep3_update_gamma_n<L, KPL, G, NC_>(q);
    else
      estep3_update_gamma_pick<L, KPL, G, NC_ - 1>(q, ncols);
  }
}
template <int L, int KPL, int G>
__device__ __forceinline__ void estep3_update_gamma(const Estep3Group &q) {
  constexpr int GT = 32 * G, KS = L * KPL, NC = (KS + GT - 1) / GT;
  const int wbase = q.gt & ~31;
  if (wbase >= q.K) return;  // a warp whose columns all lie beyond K
  const int ncols = (q.K - wbase + GT - 1) / GT;  // columns wbase + GT x < K
  estep3_update_gamma_pick<L, KPL, G, NC>(q, ncols);
}

// sweeps 1..I-1 of one document for a warp that holds NTM TMEM and NSM shared-memory batches.
// CL > 1: the sweep counter is the warp's exchange counter (kept in xbar_w[2] between documents).
template <int L, int KPL, int R, int NW, int G, int NTM, int NSM, int CL = 1>
__device__ __forceinline__ void estep3_run_sweeps(const EstepArgs &a, const Estep3Warp<L, KPL, R, NW> &w,
                                                  const Estep3Group &q,
                                                  const double (&e)[Estep3Shape<L, KPL, R, NW>::RA][KPL],
                                                  double *wtot_w, const int nb, const int64_t zrow0,
                                                  const int lane, const Estep3cX &x, unsigned long long *xbar_w) {
  using S = Estep3Shape<L, KPL, R, NW>;
  constexpr int GT = 32 * G, NA = R + NTM + NSM;
  // n_w of the (batch, row) whose norm this lane completes in the two reduce-scatters: batches j and
  // N1 + j, row g
  constexpr int N1 = Estep3Halves<R, NA>::N1;
  const int b2 = N1 + w.j;
  const int slot1 = w.j < R + NTM ? w.j : w.j - NTM + S::T;
  const int slot2 = b2 < R + NTM ? b2 : b2 - NTM + S::T;
  const double mycnt1 = w.j < N1 ? w.cnt_w[slot1 * S::RPB + w.g] : 0.0;
  const double mycnt2 = b2 < NA ? w.cnt_w[slot2 * S::RPB + w.g] : 0.0;
  int xc0 = 0;
  if constexpr (CL > 1) xc0 = (int)xbar_w[2];
  for (int sweep = 0; sweep < a.sweeps - 1; sweep++) {
    double s[KPL];
    unsigned badbits;
    long long tm[8];
    (void)tm;
    MRLDA_TICK(0);
    estep3_sweep_fast<L, KPL, R, NW, NTM, NSM, CL>(w, e, mycnt1, mycnt2, lane, s, badbits, tm, x, xc0 + sweep);
    MRLDA_TICK_DEP(4, s[KPL - 1]);
    if (__any_sync(0xffffffffu, badbits != 0)) {
      if constexpr (CL == 1)
        (void)estep3_slow_rows<L>(a.term_id + zrow0, a.count + zrow0, a.logbeta, nullptr, a.slow_rows, q.gam_g,
                                  q.gextra_g, q.K, nb, badbits, lane, false);
      else
        (void)estep3c_slow_rows<L, S::KS, CL>(a.term_id + zrow0, a.count + zrow0, a.logbeta, nullptr, a.slow_rows,
                                              q.gam_g, q.gextra_g, a.K, nb, badbits, lane, false, x.crank,
                                              x.xbar_u32, xbar_w);
    }
    estep3_publish_s<L, KPL>(s, wtot_w, lane);
    estep3_group_sync(q.bar_id, GT);  // B1: warp sums of S (and slow-path contributions) complete
    MRLDA_TICK(5);
    estep3_update_gamma<L, KPL, G>(q);
    MRLDA_TICK(6);
    estep3_group_sync(q.bar_id, GT);  // B2: theta published
    MRLDA_TICK(7);
#ifdef MRLDA_PHASE_TIMING
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      for (int i = 0; i < 7; i++) a.phase_cycles[i] += tm[i + 1] - tm[i];
      a.phase_cycles[7] += 1;
      a.phase_cycles[8] += NA;
    }
#endif
  }
  if constexpr (CL > 1) {
    __syncwarp();
    if (lane == 0) xbar_w[2] = (unsigned long long)(xc0 + a.sweeps - 1);
    __syncwarp();
  }
}

// picks the specialisation for this warp's batch count (once per document)
template <int L, int KPL, int R, int NW, int G, int NB_, int CL = 1>
__device__ __forceinline__ void estep3_dispatch(const EstepArgs &a, const Estep3Warp<L, KPL, R, NW> &w,
                                                const Estep3Group &q,
                                                const double (&e)[Estep3Shape<L, KPL, R, NW>::RA][KPL],
                                                double *wtot_w, const int nb, const int64_t zrow0,
                                                const int lane, const Estep3cX &x, unsigned long long *xbar_w) {
  using S = Estep3Shape<L, KPL, R, NW>;
  constexpr int NTM = NB_ > R ? (NB_ - R < S::T ? NB_ - R : S::T) : 0;
  constexpr int NSM = NB_ > R + S::T ? NB_ - R - S::T : 0;
  if constexpr (NB_ >= S::NBMAX) {
    estep3_run_sweeps<L, KPL, R, NW, G, NTM, NSM, CL>(a, w, q, e, wtot_w, nb, zrow0, lane, x, xbar_w);
  } else {
    if (nb <= NB_)
      estep3_run_sweeps<L, KPL, R, NW, G, NTM, NSM, CL>(a, w, q, e, wtot_w, nb, zrow0, lane, x, xbar_w);
    else
      estep3_dispatch<L, KPL, R, NW, G, NB_ + 1, CL>(a, w, q, e, wtot_w, nb, zrow0, lane, x, xbar_w);
  }
}

template <int L, int KPL, int R, int NW, int G>
__global__ void __launch_bounds__(NW * 32, 1) estep3_kernel(const EstepArgs a) {
  using S = Estep3Shape<L, KPL, R, NW>;
  constexpr int RPB = S::RPB, KS = S::KS, TW = S::TW, T = S::T, NBMAX = S::NBMAX;
  constexpr int THREADS = S::THREADS;
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int K = a.K;
  const int M = a.smem_batches;
  constexpr int GT = 32 * G;
  const int grp = warp / G, gw = warp % G;
  const Estep3Smem<L, KPL, R, NW> lay(M, G);

  Estep3Warp<L, KPL, R, NW> w;
  w.g = lane / L;
  w.j = lane % L;
  double *tile_w = smem + lay.tile + (size_t)warp * M * 32 * KPL;
  w.tile_w = tile_w;
  double *cnt_w = smem + lay.cnt + (size_t)warp * (NBMAX * RPB);
  w.cnt_w = cnt_w;
  Estep3Group q;
  q.spart = smem + lay.spart;
  q.theta_g = smem + lay.theta + (size_t)grp * KS;
  q.gam_g = smem + lay.gam + (size_t)grp * KS;
  q.gextra_g = smem + lay.gextra + (size_t)grp * KS;
  q.alpha_s = smem + lay.alpha;
  q.w0 = grp * G;
  q.gt = gw * 32 + lane;
  q.bar_id = 1 + grp;
  q.K = K;
  w.theta_g = q.theta_g;
  double *wtot_w = q.spart + (size_t)warp * KS;
  double *ssacc_s = smem + lay.ssacc;
  double *red_s = smem + lay.red;  // [warp][4]
  int *meta_s = reinterpret_cast<int *>(smem + lay.meta);  // [0] TMEM base, [1 + grp] document
  // per warp, in shared memory so that nothing of it stays in registers across the sweep loops:
  // [2 warp] the mbarrier its TMA row gathers complete on, [2 warp + 1] the parity of its next phase
  if (lane == 0) {
    unsigned long long *mb = reinterpret_cast<unsigned long long *>(smem + lay.mbar) + 2 * warp;
    estep3_mbar_init(mb, 1);
    mb[1] = 0ull;
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");

  for (int k = tid; k < KS; k += THREADS) {
    q.alpha_s[k] = k < K ? a.alpha[k] : 1.0;
    ssacc_s[k] = 0.0;
  }
  for (int k = tid; k < (NW / G) * KS; k += THREADS) {
    smem[lay.theta + k] = 0.0;
    smem[lay.gam + k] = 1.0;
    smem[lay.gextra + k] = 0.0;
  }
  // TMEM: the CTA owns the SM's 512 columns; thread-private slots (tmem_ops.cuh)
  if (warp == 0) tmem_alloc(reinterpret_cast<uint32_t *>(meta_s), 512);
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  w.tbase = (uint32_t)meta_s[0] + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * (uint32_t)S::TWORDS;
  const double lik_alpha = *a.lik_alpha;

  for (;;) {
    estep3_group_sync(q.bar_id, GT);  // everybody is done with the previous document's shared state
    if (q.gt == 0) meta_s[1 + grp] = a.doc_begin + atomicAdd(a.work_counter, 1);
    estep3_group_sync(q.bar_id, GT);
    const int qi = meta_s[1 + grp];
    if (qi >= a.doc_end) break;
    const int d = a.doc_order[qi];
    const int64_t b = a.row_ptr[d];
    const int n = (int)(a.row_ptr[d + 1] - b);
    if (n <= 0) continue;  // DocumentMapper.java:197-201: null content, nothing is updated
    // row batches of the document, dealt to the group's warps in contiguous, balanced ranges
    const int nbt = (n + RPB - 1) / RPB;
    const int nb_lo = nbt / G, nb_rem = nbt % G;
    const int nb = nb_lo + (gw < nb_rem ? 1 : 0);
    const int bstart = gw * nb_lo + (gw < nb_rem ? gw : nb_rem);
    const int64_t zrow0 = b + (int64_t)bstart * RPB;

    // ---- load my batches: complete E rows, L lanes per row (compulsory traffic).  Padding rows
    // are a copy of the document's last row with n_w = 0. ----
    for (int x = lane; x < NBMAX * RPB; x += 32) cnt_w[x] = 0.0;
    __syncwarp();
    // shared-memory batches first: lane 0 hands them to the TMA unit (one gather4 per batch: the
    // four rows are picked by term id), the copies run while the warp loads its other batches
    const int nsm = nb > R + T ? nb - (R + T) : 0;  // warp-uniform
    unsigned long long *mbar_w = reinterpret_cast<unsigned long long *>(smem + lay.mbar) + 2 * warp;
    if (nsm > 0 && lane == 0) {
      static_assert(RPB == 4, "tile::gather4 fetches the four rows of a batch");
      // the previous document's reads of this tile (generic proxy) are ordered before the copies
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      estep3_mbar_expect(mbar_w, (uint32_t)(nsm * RPB * KS * sizeof(double)));
      for (int m = 0; m < nsm; m++) {
        const int row0 = (bstart + R + T + m) * RPB;
        int id[4];
#pragma unroll
        for (int x = 0; x < 4; x++) {
          const int row = row0 + (x < RPB ? x : RPB - 1);
          id[x] = __ldg(a.term_id + b + (row < n ? row : n - 1)) - 1;
        }
        estep3_tma_gather4(tile_w + (size_t)m * 32 * KPL, a.tmap_E, mbar_w, id[0], id[1], id[2], id[3]);
      }
    }
    double e[S::RA][KPL];
#pragma unroll
    for (int bi = 0; bi < NBMAX; bi++) {
      if (bi < R || bi < nb || bi == 0) {  // warp-uniform (slot 0 always holds a row: a warp without
                                           // batches walks one padding batch)
        const int row = (bstart + bi) * RPB + w.g;
        const bool valid = bi < nb && row < n;
        const int rowc = valid ? row : n - 1;
        if (valid && w.j == 0) cnt_w[bi * RPB + w.g] = (double)__ldg(a.count + b + row);
        if (bi < R + T) {
          double v[KPL];
          estep3_load_cols<L, KPL, true>(a.E + (size_t)(__ldg(a.term_id + b + rowc) - 1) * a.e_stride, w.j, v);
          if (bi < R) {
#pragma unroll
            for (int c = 0; c < KPL; c++) e[bi < R ? bi : 0][c] = v[c];
          } else {
            uint32_t w32[TW];
#pragma unroll
            for (int c = 0; c < KPL; c++) {
              w32[2 * c] = (uint32_t)__double2loint(v[c]);
              w32[2 * c + 1] = (uint32_t)__double2hiint(v[c]);
            }
            tmem_st_words<TW>(w.tbase + (uint32_t)(bi - R) * TW, w32);
          }
        }
      }
    }
    tmem_wait_st();
    // gamma start (DocumentMapper.java:184-193) and the first theta, by the column owners
    {
      const double share = (double)((float)a.doc_tokens[d] / (float)K);
      for (int col = q.gt; col < K; col += GT) {
        const double gam = a.use_stored_gamma ? a.gamma[(size_t)d * K + col] : q.alpha_s[col] + share;
        q.gam_g[col] = gam;
        q.theta_g[col] = exp_digamma(gam);
      }
    }
    if (nsm > 0) {  // the TMA copies of this warp's shared-memory batches have landed
      const uint32_t parity = (uint32_t)mbar_w[1];
      estep3_mbar_wait(mbar_w, parity);
      __syncwarp();
      if (lane == 0) mbar_w[1] = parity ^ 1u;
    }
    estep3_group_sync(q.bar_id, GT);

    // a warp without rows (short document, more warps than batches) still takes part in the barriers
    // ---- sweeps 1..I-1: loop body specialised on this warp's batch count ----
    estep3_dispatch<L, KPL, R, NW, G, (R > 1 ? R : 1)>(a, w, q, e, wtot_w, nb, zrow0, lane, Estep3cX{}, nullptr);

    // ---- last sweep, document epilogue (DocumentMapper.java:244-259,342-346) ----
    {
      double s[KPL];
      double lik = 0.0;
      unsigned badbits;
      estep3_sweep_last<L, KPL, R, NW>(a, w, e, s, lik, badbits, nb, zrow0);
      if (__any_sync(0xffffffffu, badbits != 0))
        lik += estep3_slow_rows<L>(a.term_id + zrow0, a.count + zrow0, a.logbeta, a.learning ? a.stats : nullptr,
                                   a.slow_rows, q.gam_g, q.gextra_g, K, nb, badbits, lane, true);
      estep3_publish_s<L, KPL>(s, wtot_w, lane);
      lik = warp_sum(lik);
      if (lane == 0) red_s[warp * 4 + 3] = lik;
      estep3_group_sync(q.bar_id, GT);
      double sg = 0.0, lg = 0.0, lk = 0.0;
      for (int col = q.gt; col < K; col += GT) {
        const double sk = estep3_colsum<KS, G>(q.spart, q.w0, col);
        const double add = q.theta_g[col] * sk;
        const double g_old = q.gam_g[col];
        const double g_new = q.alpha_s[col] + add + q.gextra_g[col];
        q.gextra_g[col] = 0.0;
        if (add != 0.0) lk -= digamma_fast(g_old) * add;
        q.gam_g[col] = g_new;
        sg += g_new;
        lg += lgamma(g_new);
        a.gamma[(size_t)d * K + col] = g_new;
      }
      sg = warp_sum(sg);
      lg = warp_sum(lg);
      lk = warp_sum(lk);
      if (lane == 0) {
        red_s[warp * 4 + 0] = sg;
        red_s[warp * 4 + 1] = lg;
        red_s[warp * 4 + 2] = lk;
      }
      estep3_group_sync(q.bar_id, GT);
      double tg = 0.0, tl = 0.0, tk = 0.0;
      for (int w2 = 0; w2 < G; w2++) {
        const double *rp = red_s + (q.w0 + w2) * 4;
        tg += rp[0];
        tl += rp[1];
        tk += rp[2] + rp[3];
      }
      const double dsg = digamma_literal(tg);
      for (int col = q.gt; col < K; col += GT) atomicAdd(ssacc_s + col, digamma_literal(q.gam_g[col]) - dsg);
      if (q.gt == 0) {
        const double ll = lik_alpha + (tl - lgamma(tg)) + tk;
        const long long c = (long long)(-ll * 1e6);  // DocumentMapper.java:253-254
        atomicAdd(a.ll_counter, (unsigned long long)c);
      }
    }
  }
  // DocumentMapper.close (DocumentMapper.java:354-361)
  tmem_fence_before_sync();
  __syncthreads();
  for (int k = tid; k < K; k += THREADS) {
    const double v = ssacc_s[k];
    if (v != 0.0) atomicAdd(a.alpha_ss + k, v);
  }
  if (warp == 0) tmem_dealloc((uint32_t)meta_s[0], 512);
}

struct Estep3Variant {
  int L, KPL, R, T, NW, MMAX;
  cudaError_t (*launch)(const EstepArgs &, int grid, int G, int M, size_t smem, cudaStream_t);
  size_t (*smem_bytes)(int M, int G);
  cudaError_t (*prepare)(size_t smem, int G);
};

template <int L, int KPL, int R, int NW>
static size_t variant3_smem(int M, int G) {
  return Estep3Smem<L, KPL, R, NW>(M, G).total;
}
template <int L, int KPL, int R, int NW>
static cudaError_t variant3_prepare(size_t smem, int G) {
  const int sz = (int)smem;
  switch (G) {
    case 8: return cudaFuncSetAttribute(estep3_kernel<L, KPL, R, NW, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, sz);
    case 4: return cudaFuncSetAttribute(estep3_kernel<L, KPL, R, NW, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, sz);
    case 2: return cudaFuncSetAttribute(estep3_kernel<L, KPL, R, NW, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sz);
    case 1: return cudaFuncSetAttribute(estep3_kernel<L, KPL, R, NW, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sz);
  }
  return cudaErrorInvalidValue;
}
template <int L, int KPL, int R, int NW>
static cudaError_t variant3_launch(const EstepArgs &args, int grid, int G, int M, size_t smem, cudaStream_t st) {
  EstepArgs a = args;
  a.group_warps = G;
  a.smem_batches = M;
  switch (G) {
    case 8: estep3_kernel<L, KPL, R, NW, 8><<<grid, NW * 32, smem, st>>>(a); break;
    case 4: estep3_kernel<L, KPL, R, NW, 4><<<grid, NW * 32, smem, st>>>(a); break;
    case 2: estep3_kernel<L, KPL, R, NW, 2><<<grid, NW * 32, smem, st>>>(a); break;
    case 1: estep3_kernel<L, KPL, R, NW, 1><<<grid, NW * 32, smem, st>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

#define MRLDA_VARIANT3(L_, KPL_, R_, NW_)                                                          \
  {                                                                                                \
    L_, KPL_, R_, Estep3Shape<L_, KPL_, R_, NW_>::T, NW_, Estep3Shape<L_, KPL_, R_, NW_>::MMAX,     \
        variant3_launch<L_, KPL_, R_, NW_>, variant3_smem<L_, KPL_, R_, NW_>,                      \
        variant3_prepare<L_, KPL_, R_, NW_>                                                        \
  }

extern const Estep3Variant g_estep3_variants_L8[];
extern const int g_estep3_variants_L8_n;
extern const Estep3Variant g_estep3_variants_L8o[];
extern const int g_estep3_variants_L8o_n;
extern const Estep3Variant g_estep3_variants_L8n17[];
extern const int g_estep3_variants_L8n17_n;
extern const Estep3Variant g_estep3_variants_L8n20[];
extern const int g_estep3_variants_L8n20_n;
extern const Estep3Variant g_estep3_variants_L8n22[];
extern const int g_estep3_variants_L8n22_n;

}  // namespace mrlda
