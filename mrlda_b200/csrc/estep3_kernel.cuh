// estep3_kernel.cuh — K1, warp-group variant: several documents in flight per SM.
//
// Same arithmetic as estep_kernel.cuh / estep2_kernel.cuh (linear-space fixed point, 99 sweeps, exact
// log-space redo of underflowing rows; reference: DocumentMapper.java:175-350, updatePhi :402-429).
// What changes is who owns what:
//
//   * A CTA is 8 warps on one SM.  A document is owned by a GROUP of G consecutive warps
//     (G in {1,2,4,8}, a launch parameter), so 8/G documents are in flight per SM and one
//     document's serial phases (S exchange, exp(digamma), barriers) overlap another document's
//     DFMA phases on the same FP64 pipes.  Groups synchronise on named barriers (bar.sync id, 32*G)
//     and pull documents from the work queue independently of each other.
//   * Inside a group ROWS are split across warps: warp k owns a contiguous range of row batches.
//     A batch is RPB = 32/L rows; L lanes share a row, lane j of a row owns the KPL columns
//     {2(j + L i), 2(j + L i) + 1}, i < KPL/2 (same column interleave as estep_kernel.cuh, so row
//     loads and statistics atomics touch 16-byte units that are contiguous across the L lanes).
//     Every warp therefore holds COMPLETE rows: norm_w needs only an L-lane butterfly (3 shuffle
//     levels at K = 200) and r_w = n_w / norm_w is used by the same lanes right away — no CTA
//     barrier and no shared-memory exchange per row.  S_c accumulates per lane over the warp's rows.
//   * A warp's batches live, in this order, in REGISTERS (the first R batches), in TENSOR MEMORY
//     (the next T = 256 / (2 KPL) batches: thread-private tcgen05.st / tcgen05.ld, no MMA; warps w
//     and w+4 share a lane quarter and split the 512 columns), and in a thread-private,
//     lane-interleaved region of SHARED MEMORY (the last M batches).
//   * Per sweep and warp.  Pass A: theta in registers; partial dot products of every batch (TMEM
//     batches are fetched in two parts so that one tcgen05.ld is always in flight behind the
//     DFMAs).  Then the L-lane butterflies and the reciprocals of ALL batches as independent,
//     interleaved chains.  Pass B: theta is dead, S takes its registers; S += r E over every batch
//     (TMEM and shared-memory batches are read a second time).  The warp then publishes its
//     RPB x K partial sums to shared memory as they are (no shuffle reduction), group barrier,
//     the group's threads own one or two columns each: sum the RPB*G partials,
//     gamma_c = alpha_c + theta_c S_c, theta_c = exp(digamma(gamma_c)), group barrier.
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

namespace mrlda {

template <int L, int KPL, int R>
struct Estep3Shape {
  static constexpr int RPB = 32 / L;
  static constexpr int KS = L * KPL;
  static constexpr int HALF = KPL / 2;
  static constexpr int TW = 2 * KPL;   // 32-bit TMEM columns per batch and lane
  static constexpr int T = 256 / TW;   // TMEM batches per warp
  static constexpr int MMAX = 4;       // upper bound of the shared-memory batches per warp
  static constexpr int NBMAX = R + T + MMAX;
  static constexpr int CA = 16;        // columns of a TMEM batch fetched by the first tcgen05.ld
  static constexpr int CB = KPL - CA;  // ... and by the second
  static_assert(KPL % 2 == 0 && KPL > CA && KPL <= 32, "KPL must be even, in (16, 32]");
  static_assert(L == 1 || L == 2 || L == 4 || L == 8 || L == 16 || L == 32, "L must divide 32");
  static_assert(T <= 4 && MMAX == 4, "the batch-count dispatch below is written for T, MMAX <= 4");
};

// shared-memory carve-up (offsets in doubles) for M shared-memory batches per warp, groups of G warps
template <int L, int KPL, int R>
struct Estep3Smem {
  using S = Estep3Shape<L, KPL, R>;
  size_t tile, spart, theta, gam, gextra, alpha, ssacc, cnt, red, meta, total;
  int capw;  // rows one warp can hold
  __host__ __device__ Estep3Smem(int M, int G) {
    const int NG = 8 / G;
    capw = (R + S::T + M) * S::RPB;
    size_t o = 0;
    tile = o;   o += (size_t)8 * M * 32 * KPL;
    spart = o;  o += (size_t)8 * 32 * KPL;    // every lane's S partials: [warp][i][lane] double2
    theta = o;  o += (size_t)NG * S::KS;
    gam = o;    o += (size_t)NG * S::KS;
    gextra = o; o += (size_t)NG * S::KS;
    alpha = o;  o += S::KS;
    ssacc = o;  o += S::KS;
    cnt = o;    o += (size_t)8 * S::NBMAX * S::RPB;  // n_w of every batch slot (0: padding)
    red = o;    o += 8 * 4;
    meta = o;   o += 8;  // 16 ints
    total = o * sizeof(double);
  }
};

__device__ __forceinline__ void estep3_group_sync(const int id, const int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ double estep3_w2d(const uint32_t lo, const uint32_t hi) {
  return __hiloint2double((int)hi, (int)lo);
}

// n_w phi_wk = r_w theta_k E_wk into the K x V statistics (DocumentMapper.java:315-329)
template <int L, int KPL>
__device__ __forceinline__ void estep3_scatter(double *srow, const double2 *theta2, const double r,
                                               const double (&ev)[KPL], const int j, const int K) {
#pragma unroll
  for (int i = 0; i < KPL / 2; i++) {
    const double2 tv = theta2[j + L * i];
    const int col = 2 * (j + L * i);
    if (col < K) atomicAdd(srow + col, r * tv.x * ev[2 * i]);
    if (col + 1 < K) atomicAdd(srow + col + 1, r * tv.y * ev[2 * i + 1]);
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

// ---- pass A pieces: partial dot products ----
// NT TMEM batches, software pipelined: part A (CA columns) and part B (CB columns) of a batch are
// separate loads, and the next load is issued before the DFMAs of the part that just arrived.
template <int L, int KPL, int R, int NT>
__device__ __forceinline__ void estep3_dots_tmem(const double (&th)[KPL], const uint32_t tbase,
                                                 double (&pr)[Estep3Shape<L, KPL, R>::NBMAX]) {
  using S = Estep3Shape<L, KPL, R>;
  constexpr int CA = S::CA, CB = S::CB, TW = S::TW;
  if constexpr (NT > 0) {
    uint32_t wa[2 * CA], wb[2 * CB];
    tmem_ld_words<2 * CA>(tbase, wa);
#pragma unroll
    for (int t = 0; t < NT; t++) {
      tmem_wait_ld();
      tmem_ld_words<2 * CB>(tbase + (uint32_t)(t * TW + 2 * CA), wb);
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int c = 0; c < CA; c += 4) {
        a0 = fma(th[c], estep3_w2d(wa[2 * c], wa[2 * c + 1]), a0);
        a1 = fma(th[c + 1], estep3_w2d(wa[2 * c + 2], wa[2 * c + 3]), a1);
        a2 = fma(th[c + 2], estep3_w2d(wa[2 * c + 4], wa[2 * c + 5]), a2);
        a3 = fma(th[c + 3], estep3_w2d(wa[2 * c + 6], wa[2 * c + 7]), a3);
      }
      tmem_wait_ld();
      if (t + 1 < NT) tmem_ld_words<2 * CA>(tbase + (uint32_t)((t + 1) * TW), wa);
#pragma unroll
      for (int c = 0; c < CB; c += 2) {
        if (c & 2) {
          a2 = fma(th[CA + c], estep3_w2d(wb[2 * c], wb[2 * c + 1]), a2);
          a3 = fma(th[CA + c + 1], estep3_w2d(wb[2 * c + 2], wb[2 * c + 3]), a3);
        } else {
          a0 = fma(th[CA + c], estep3_w2d(wb[2 * c], wb[2 * c + 1]), a0);
          a1 = fma(th[CA + c + 1], estep3_w2d(wb[2 * c + 2], wb[2 * c + 3]), a1);
        }
      }
      pr[R + t] = (a0 + a1) + (a2 + a3);
    }
  }
}
template <int L, int KPL, int R, int NM>
__device__ __forceinline__ void estep3_dots_smem(const double (&th)[KPL], const double2 *tile2,
                                                 double (&pr)[Estep3Shape<L, KPL, R>::NBMAX]) {
  using S = Estep3Shape<L, KPL, R>;
  constexpr int HALF = S::HALF, T = S::T;
#pragma unroll
  for (int m = 0; m < NM; m++) {
    const double2 *src = tile2 + (size_t)m * HALF * 32;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int i = 0; i < HALF; i++) {
      const double2 v = src[(size_t)i * 32];
      if (i & 1) {
        a2 = fma(th[2 * i], v.x, a2);
        a3 = fma(th[2 * i + 1], v.y, a3);
      } else {
        a0 = fma(th[2 * i], v.x, a0);
        a1 = fma(th[2 * i + 1], v.y, a1);
      }
    }
    pr[R + T + m] = (a0 + a1) + (a2 + a3);
  }
}

// ---- norms -> r for the first NA batch slots: every step runs over all slots before the next ----
template <int L, int KPL, int R, int NA>
__device__ __forceinline__ void estep3_norms_to_r(double (&pr)[Estep3Shape<L, KPL, R>::NBMAX],
                                                  double (&rr)[Estep3Shape<L, KPL, R>::NBMAX],
                                                  const double *cnt_w, const int g, int &minhi) {
  constexpr int RPB = 32 / L;
#pragma unroll
  for (int o = 1; o < L; o <<= 1) {
#pragma unroll
    for (int bi = 0; bi < NA; bi++) pr[bi] += __shfl_xor_sync(0xffffffffu, pr[bi], o);
  }
  double y[NA], t1[NA];
#pragma unroll
  for (int bi = 0; bi < NA; bi++) {
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y[bi]) : "d"(pr[bi]));
    minhi = min(minhi, __double2hiint(pr[bi]));
  }
#pragma unroll
  for (int bi = 0; bi < NA; bi++) t1[bi] = fma(-pr[bi], y[bi], 1.0);
#pragma unroll
  for (int bi = 0; bi < NA; bi++) y[bi] = fma(y[bi], t1[bi], y[bi]);
#pragma unroll
  for (int bi = 0; bi < NA; bi++) t1[bi] = fma(-pr[bi], y[bi], 1.0);
#pragma unroll
  for (int bi = 0; bi < NA; bi++) y[bi] = fma(y[bi], t1[bi], y[bi]);
#pragma unroll
  for (int bi = 0; bi < NA; bi++) rr[bi] = cnt_w[bi * RPB + g] * y[bi];
}

// ---- pass B pieces: S_c += r_w E_wc ----
template <int L, int KPL, int R, int NT, bool LAST>
__device__ __forceinline__ void estep3_sacc_tmem(const EstepArgs &a, const uint32_t tbase,
                                                 const double (&rr)[Estep3Shape<L, KPL, R>::NBMAX],
                                                 double (&s)[KPL], const double2 *theta2,
                                                 const int64_t zrow0, const int g, const int j) {
  using S = Estep3Shape<L, KPL, R>;
  constexpr int CA = S::CA, CB = S::CB, TW = S::TW, RPB = S::RPB;
  if constexpr (NT > 0 && !LAST) {
    uint32_t wa[2 * CA], wb[2 * CB];
    tmem_ld_words<2 * CA>(tbase, wa);
#pragma unroll
    for (int t = 0; t < NT; t++) {
      const double r = rr[R + t];
      tmem_wait_ld();
      tmem_ld_words<2 * CB>(tbase + (uint32_t)(t * TW + 2 * CA), wb);
#pragma unroll
      for (int c = 0; c < CA; c++) s[c] = fma(r, estep3_w2d(wa[2 * c], wa[2 * c + 1]), s[c]);
      tmem_wait_ld();
      if (t + 1 < NT) tmem_ld_words<2 * CA>(tbase + (uint32_t)((t + 1) * TW), wa);
#pragma unroll
      for (int c = 0; c < CB; c++) s[CA + c] = fma(r, estep3_w2d(wb[2 * c], wb[2 * c + 1]), s[CA + c]);
    }
  } else if constexpr (NT > 0) {
#pragma unroll
    for (int t = 0; t < NT; t++) {
      uint32_t w32[TW];
      tmem_ld_words<TW>(tbase + (uint32_t)t * TW, w32);
      tmem_wait_ld();
      double ev[KPL];
#pragma unroll
      for (int c = 0; c < KPL; c++) ev[c] = estep3_w2d(w32[2 * c], w32[2 * c + 1]);
#pragma unroll
      for (int c = 0; c < KPL; c++) s[c] = fma(rr[R + t], ev[c], s[c]);
      if (a.learning && rr[R + t] > 0.0)
        estep3_scatter<L, KPL>(a.stats + (size_t)(__ldg(a.term_id + zrow0 + (R + t) * RPB + g) - 1) * a.K,
                               theta2, rr[R + t], ev, j, a.K);
    }
  }
}
template <int L, int KPL, int R, int NM, bool LAST>
__device__ __forceinline__ void estep3_sacc_smem(const EstepArgs &a, const double2 *tile2,
                                                 const double (&rr)[Estep3Shape<L, KPL, R>::NBMAX],
                                                 double (&s)[KPL], const double2 *theta2,
                                                 const int64_t zrow0, const int g, const int j) {
  using S = Estep3Shape<L, KPL, R>;
  constexpr int HALF = S::HALF, T = S::T, RPB = S::RPB;
#pragma unroll
  for (int m = 0; m < NM; m++) {
    const double2 *src = tile2 + (size_t)m * HALF * 32;
    const double r = rr[R + T + m];
    double ev[KPL];
#pragma unroll
    for (int i = 0; i < HALF; i++) {
      const double2 v = src[(size_t)i * 32];
      ev[2 * i] = v.x;
      ev[2 * i + 1] = v.y;
    }
#pragma unroll
    for (int c = 0; c < KPL; c++) s[c] = fma(r, ev[c], s[c]);
    if (LAST && a.learning && r > 0.0)
      estep3_scatter<L, KPL>(a.stats + (size_t)(__ldg(a.term_id + zrow0 + (R + T + m) * RPB + g) - 1) * a.K,
                             theta2, r, ev, j, a.K);
  }
}

#define MRLDA_E3_DISPATCH(n_, CALL)            \
  switch (n_) {                                \
    case 1: { CALL(1); } break;                \
    case 2: { CALL(2); } break;                \
    case 3: { CALL(3); } break;                \
    case 4: { CALL(4); } break;                \
    default: break;                            \
  }

// One sweep over the batches of this warp.  e: register batches; tbase: this warp's TMEM slots;
// tile2: this warp's shared-memory batches (+ lane); theta2: the group's theta as double2 pairs;
// cnt_w: n_w of this warp's batch slots (0 for padding rows); zrow0: CSR offset of this warp's
// first row; ntm / nsm: TMEM / shared-memory batches in use.
template <int L, int KPL, int R, bool LAST>
__device__ __forceinline__ void estep3_sweep(const EstepArgs &a, const double (&e)[R][KPL],
                                             const uint32_t tbase, const double2 *tile2,
                                             const double2 *theta2, const double *cnt_w,
                                             double (&s)[KPL], double &lik, unsigned &badbits,
                                             const int ntm, const int nsm, const int64_t zrow0,
                                             const int g, const int j) {
  using S = Estep3Shape<L, KPL, R>;
  constexpr int RPB = S::RPB, HALF = S::HALF, T = S::T, NB = S::NBMAX;
  double pr[NB];  // partial dot products -> norms
  double rr[NB];  // r_w = n_w / norm_w
#pragma unroll
  for (int bi = 0; bi < NB; bi++) {  // slots that are not in use
    pr[bi] = 1.0;
    rr[bi] = 0.0;
  }
  // ---- pass A: theta in registers; partial dot products of every batch ----
  {
    double th[KPL];
#pragma unroll
    for (int i = 0; i < HALF; i++) {
      const double2 v = theta2[j + L * i];
      th[2 * i] = v.x;
      th[2 * i + 1] = v.y;
    }
    {
      double acc[R][4];
#pragma unroll
      for (int q = 0; q < R; q++) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0;
#pragma unroll
      for (int c = 0; c < KPL; c += 4) {
#pragma unroll
        for (int q = 0; q < R; q++) {
          acc[q][0] = fma(th[c], e[q][c], acc[q][0]);
          acc[q][1] = fma(th[c + 1], e[q][c + 1], acc[q][1]);
          if (c + 2 < KPL) {
            acc[q][2] = fma(th[c + 2], e[q][c + 2], acc[q][2]);
            acc[q][3] = fma(th[c + 3], e[q][c + 3], acc[q][3]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < R; q++) pr[q] = (acc[q][0] + acc[q][1]) + (acc[q][2] + acc[q][3]);
    }
#define MRLDA_E3_CALL(N) estep3_dots_tmem<L, KPL, R, N>(th, tbase, pr)
    MRLDA_E3_DISPATCH(ntm, MRLDA_E3_CALL)
#undef MRLDA_E3_CALL
#define MRLDA_E3_CALL(N) estep3_dots_smem<L, KPL, R, N>(th, tile2, pr)
    MRLDA_E3_DISPATCH(nsm, MRLDA_E3_CALL)
#undef MRLDA_E3_CALL
  }
  // ---- norms (L-lane butterflies), r = n_w / norm_w ----
  int minhi = 0x7fffffff;
  if (nsm > 0)
    estep3_norms_to_r<L, KPL, R, NB>(pr, rr, cnt_w, g, minhi);
  else if (ntm > 2)
    estep3_norms_to_r<L, KPL, R, R + T>(pr, rr, cnt_w, g, minhi);
  else
    estep3_norms_to_r<L, KPL, R, R + 2>(pr, rr, cnt_w, g, minhi);
  badbits = 0;
  if (__any_sync(0xffffffffu, minhi < __double2hiint(MRLDA_NORM_TINY))) {
    // some norm underflowed (rare): those rows contribute through the exact log-space redo
#pragma unroll
    for (int bi = 0; bi < NB; bi++) {
      if (!(pr[bi] >= MRLDA_NORM_TINY)) {
        if (cnt_w[bi * RPB + g] > 0.0) badbits |= 1u << bi;
        rr[bi] = 0.0;
      }
    }
  }
  if (LAST && j == 0) {  // likelihoodPhi, first part (DocumentMapper.java:421 summed over topics)
#pragma unroll
    for (int bi = 0; bi < NB; bi++) {
      if (rr[bi] > 0.0)
        lik = fma(cnt_w[bi * RPB + g], log(pr[bi]) + __ldg(a.rowmax + __ldg(a.term_id + zrow0 + bi * RPB + g) - 1), lik);
    }
  }
  // ---- pass B: S_c += r_w E_wc ----
#pragma unroll
  for (int c = 0; c < KPL; c++) s[c] = 0.0;
#pragma unroll
  for (int c = 0; c < KPL; c++) {
#pragma unroll
    for (int q = 0; q < R; q++) s[c] = fma(rr[q], e[q][c], s[c]);
  }
  if (LAST && a.learning) {
#pragma unroll
    for (int q = 0; q < R; q++)
      if (rr[q] > 0.0)
        estep3_scatter<L, KPL>(a.stats + (size_t)(__ldg(a.term_id + zrow0 + q * RPB + g) - 1) * a.K, theta2,
                               rr[q], e[q], j, a.K);
  }
#define MRLDA_E3_CALL(N) estep3_sacc_tmem<L, KPL, R, N, LAST>(a, tbase, rr, s, theta2, zrow0, g, j)
  MRLDA_E3_DISPATCH(ntm, MRLDA_E3_CALL)
#undef MRLDA_E3_CALL
#define MRLDA_E3_CALL(N) estep3_sacc_smem<L, KPL, R, N, LAST>(a, tile2, rr, s, theta2, zrow0, g, j)
  MRLDA_E3_DISPATCH(nsm, MRLDA_E3_CALL)
#undef MRLDA_E3_CALL
}

// sum of the RPB * G partial sums of column col that the group's warps published
template <int L, int KPL>
__device__ __forceinline__ double estep3_colsum(const double *spart, const int w0, const int G, const int col) {
  constexpr int RPB = 32 / L, HALF = KPL / 2;
  const int i = (col >> 1) / L, jj = (col >> 1) % L;
  const double *p = spart + ((size_t)(w0 * HALF + i) * 32 + jj) * 2 + (col & 1);
  double s0 = 0.0, s1 = 0.0;
  for (int w2 = 0; w2 < G; w2++) {
#pragma unroll
    for (int gg = 0; gg < RPB; gg++) {
      const double v = p[(size_t)w2 * HALF * 64 + gg * L * 2];
      if (gg & 1) s1 += v;
      else s0 += v;
    }
  }
  return s0 + s1;
}

template <int L, int KPL, int R>
__global__ void __launch_bounds__(256, 1) estep3_kernel(const EstepArgs a) {
  using S = Estep3Shape<L, KPL, R>;
  constexpr int RPB = S::RPB, KS = S::KS, HALF = S::HALF, TW = S::TW, T = S::T, NBMAX = S::NBMAX;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane / L, j = lane % L;
  const int K = a.K;
  const int G = a.group_warps, M = a.smem_batches;
  const int grp = warp / G, gw = warp % G, gt = gw * 32 + lane, GT = 32 * G;
  const int bar_id = 1 + grp;
  const Estep3Smem<L, KPL, R> lay(M, G);

  double2 *tile2 = reinterpret_cast<double2 *>(smem + lay.tile) + (size_t)warp * M * HALF * 32 + lane;
  double *spart = smem + lay.spart;
  double2 *spart_w = reinterpret_cast<double2 *>(spart) + (size_t)warp * HALF * 32 + lane;
  double *theta_g = smem + lay.theta + (size_t)grp * KS;
  double *gam_g = smem + lay.gam + (size_t)grp * KS;
  double *gextra_g = smem + lay.gextra + (size_t)grp * KS;
  double *alpha_s = smem + lay.alpha;
  double *ssacc_s = smem + lay.ssacc;
  double *cnt_w = smem + lay.cnt + (size_t)warp * (NBMAX * RPB);
  double *red_s = smem + lay.red;  // [warp][4]
  int *meta_s = reinterpret_cast<int *>(smem + lay.meta);  // [0] TMEM base, [1 + grp] document
  const double2 *theta2 = reinterpret_cast<const double2 *>(theta_g);

  for (int k = tid; k < KS; k += 256) {
    alpha_s[k] = k < K ? a.alpha[k] : 1.0;
    ssacc_s[k] = 0.0;
  }
  for (int k = tid; k < (8 / G) * KS; k += 256) {
    smem[lay.theta + k] = 0.0;
    smem[lay.gam + k] = 1.0;
    smem[lay.gextra + k] = 0.0;
  }
  // TMEM: the CTA owns the SM's 512 columns; thread-private slots (tmem_ops.cuh)
  if (warp == 0) tmem_alloc(reinterpret_cast<uint32_t *>(meta_s), 512);
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t tbase = (uint32_t)meta_s[0] + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 256u;
  const double lik_alpha = *a.lik_alpha;

  for (;;) {
    estep3_group_sync(bar_id, GT);  // everybody is done with the previous document's shared state
    if (gt == 0) meta_s[1 + grp] = a.doc_begin + atomicAdd(a.work_counter, 1);
    estep3_group_sync(bar_id, GT);
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
    const int ntm = nb > R ? (nb - R < T ? nb - R : T) : 0;
    const int nsm = nb > R + T ? nb - R - T : 0;

    // ---- load my batches: complete E rows, L lanes per row (compulsory traffic).  Padding rows
    // are a copy of the document's last row with n_w = 0. ----
    for (int x = lane; x < NBMAX * RPB; x += 32) cnt_w[x] = 0.0;
    __syncwarp();
    double e[R][KPL];
#pragma unroll
    for (int bi = 0; bi < NBMAX; bi++) {
      if (bi < R || bi < nb) {  // warp-uniform
        const int row = (bstart + bi) * RPB + g;
        const bool valid = bi < nb && row < n;
        const int rowc = valid ? row : n - 1;
        double v[KPL];
        const double2 *src = reinterpret_cast<const double2 *>(
            a.E + (size_t)(__ldg(a.term_id + b + rowc) - 1) * a.e_stride) + j;
#pragma unroll
        for (int i = 0; i < HALF; i++) {
          const double2 t2 = __ldg(src + L * i);
          v[2 * i] = t2.x;
          v[2 * i + 1] = t2.y;
        }
        if (valid && j == 0) cnt_w[bi * RPB + g] = (double)__ldg(a.count + b + row);
        if (bi < R) {
#pragma unroll
          for (int c = 0; c < KPL; c++) e[bi < R ? bi : 0][c] = v[c];
        } else if (bi < R + T) {
          uint32_t w32[TW];
#pragma unroll
          for (int c = 0; c < KPL; c++) {
            w32[2 * c] = (uint32_t)__double2loint(v[c]);
            w32[2 * c + 1] = (uint32_t)__double2hiint(v[c]);
          }
          tmem_st_words<TW>(tbase + (uint32_t)(bi - R) * TW, w32);
        } else {
          double2 *dst = tile2 + (size_t)(bi - R - T) * HALF * 32;
#pragma unroll
          for (int i = 0; i < HALF; i++) dst[(size_t)i * 32] = make_double2(v[2 * i], v[2 * i + 1]);
        }
      }
    }
    tmem_wait_st();
    // gamma start (DocumentMapper.java:184-193) and the first theta, by the column owners
    {
      const double share = (double)((float)a.doc_tokens[d] / (float)K);
      for (int col = gt; col < K; col += GT) {
        const double gam = a.use_stored_gamma ? a.gamma[(size_t)d * K + col] : alpha_s[col] + share;
        gam_g[col] = gam;
        theta_g[col] = exp_digamma(gam);
      }
    }
    estep3_group_sync(bar_id, GT);

    double lik = 0.0;
    for (int sweep = 0; sweep < a.sweeps; sweep++) {
      const bool last = (sweep == a.sweeps - 1);
      double s[KPL];
      unsigned badbits;
      if (!last)
        estep3_sweep<L, KPL, R, false>(a, e, tbase, tile2, theta2, cnt_w, s, lik, badbits, ntm, nsm, zrow0, g, j);
      else
        estep3_sweep<L, KPL, R, true>(a, e, tbase, tile2, theta2, cnt_w, s, lik, badbits, ntm, nsm, zrow0, g, j);
      if (__any_sync(0xffffffffu, badbits != 0))
        lik += estep3_slow_rows<L>(a.term_id + zrow0, a.count + zrow0, a.logbeta, a.learning ? a.stats : nullptr,
                                   a.slow_rows, gam_g, gextra_g, K, nb, badbits, lane, last);
      // S partials of every lane as they are; the column owners sum them
#pragma unroll
      for (int i = 0; i < HALF; i++) spart_w[(size_t)i * 32] = make_double2(s[2 * i], s[2 * i + 1]);
      if (last) {
        lik = warp_sum(lik);
        if (lane == 0) red_s[warp * 4 + 3] = lik;
      }
      estep3_group_sync(bar_id, GT);  // B1: S partials (and slow-path contributions) complete
      if (!last) {
        // column owners: gamma_c = alpha_c + theta_c S_c (DocumentMapper.java:232-234), next theta
        for (int col = gt; col < K; col += 2 * GT) {
          const int col2 = col + GT;
          const bool two = col2 < K;
          const double s0 = estep3_colsum<L, KPL>(spart, grp * G, G, col);
          const double s1 = estep3_colsum<L, KPL>(spart, grp * G, G, two ? col2 : col);
          const double g0 = fma(theta_g[col], s0, alpha_s[col]) + gextra_g[col];
          const double g1 = two ? fma(theta_g[col2], s1, alpha_s[col2]) + gextra_g[col2] : 1.0;
          gextra_g[col] = 0.0;
          gam_g[col] = g0;
          const double t0 = exp_digamma(g0);
          const double t1 = exp_digamma(g1);
          theta_g[col] = t0;
          if (two) {
            gextra_g[col2] = 0.0;
            gam_g[col2] = g1;
            theta_g[col2] = t1;
          }
        }
        estep3_group_sync(bar_id, GT);  // B2: theta published
      } else {
        // ---- document epilogue (DocumentMapper.java:244-259,342-346) ----
        double sg = 0.0, lg = 0.0, lk = 0.0;
        for (int col = gt; col < K; col += GT) {
          const double sk = estep3_colsum<L, KPL>(spart, grp * G, G, col);
          const double add = theta_g[col] * sk;
          const double g_old = gam_g[col];
          const double g_new = alpha_s[col] + add + gextra_g[col];
          gextra_g[col] = 0.0;
          if (add != 0.0) lk -= digamma_fast(g_old) * add;
          gam_g[col] = g_new;
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
        estep3_group_sync(bar_id, GT);
        double tg = 0.0, tl = 0.0, tk = 0.0;
        for (int w2 = 0; w2 < G; w2++) {
          const double *rp = red_s + (grp * G + w2) * 4;
          tg += rp[0];
          tl += rp[1];
          tk += rp[2] + rp[3];
        }
        const double dsg = digamma_literal(tg);
        for (int col = gt; col < K; col += GT) atomicAdd(ssacc_s + col, digamma_literal(gam_g[col]) - dsg);
        if (gt == 0) {
          const double ll = lik_alpha + (tl - lgamma(tg)) + tk;
          const long long c = (long long)(-ll * 1e6);  // DocumentMapper.java:253-254
          atomicAdd(a.ll_counter, (unsigned long long)c);
        }
      }
    }
  }
  // DocumentMapper.close (DocumentMapper.java:354-361)
  tmem_fence_before_sync();
  __syncthreads();
  for (int k = tid; k < K; k += 256) {
    const double v = ssacc_s[k];
    if (v != 0.0) atomicAdd(a.alpha_ss + k, v);
  }
  if (warp == 0) tmem_dealloc((uint32_t)meta_s[0], 512);
}

struct Estep3Variant {
  int L, KPL, R, T;
  cudaError_t (*launch)(const EstepArgs &, int grid, int G, int M, size_t smem, cudaStream_t);
  size_t (*smem_bytes)(int M, int G);
  cudaError_t (*prepare)(size_t smem);
};

template <int L, int KPL, int R>
static size_t variant3_smem(int M, int G) {
  return Estep3Smem<L, KPL, R>(M, G).total;
}
template <int L, int KPL, int R>
static cudaError_t variant3_prepare(size_t smem) {
  return cudaFuncSetAttribute(estep3_kernel<L, KPL, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <int L, int KPL, int R>
static cudaError_t variant3_launch(const EstepArgs &args, int grid, int G, int M, size_t smem, cudaStream_t st) {
  EstepArgs a = args;
  a.group_warps = G;
  a.smem_batches = M;
  estep3_kernel<L, KPL, R><<<grid, 256, smem, st>>>(a);
  return cudaGetLastError();
}

#define MRLDA_VARIANT3(L_, KPL_, R_)                                                       \
  {                                                                                        \
    L_, KPL_, R_, Estep3Shape<L_, KPL_, R_>::T, variant3_launch<L_, KPL_, R_>,             \
        variant3_smem<L_, KPL_, R_>, variant3_prepare<L_, KPL_, R_>                        \
  }

extern const Estep3Variant g_estep3_variants_L8[];
extern const int g_estep3_variants_L8_n;

}  // namespace mrlda
