// estep3c_kernel.cuh — K1, warp-group E-step for more topics than one SM's row slice holds
// (208 < K <= 1040; cfg 5: K = 1000 = 5 x 200): the warp-group kernel of estep3_kernel.cuh with the
// TOPICS split across a thread-block cluster.
//
// Same arithmetic as the other E-step kernels (linear-space fixed point, 99 sweeps, exact log-space
// redo of underflowing rows; reference: DocumentMapper.java:175-350, updatePhi :402-429).  Mapping:
//
//   * A cluster is CL CTAs (one per SM); CTA r owns the KS = L * KPL columns [r KS, (r + 1) KS) of
//     every row.  Inside a CTA everything is as in estep3_kernel.cuh: a document is owned by a
//     group of G consecutive warps, rows are split across the group's warps, a warp's batches sit
//     in registers, Tensor Memory and TMA-gathered shared memory, S is summed inside the CTA, and
//     the group's threads own the CTA's columns (gamma, theta).  Group g of every CTA of the
//     cluster works on the same document, warp w of every CTA on the same rows of it.
//   * The only per-sweep traffic between CTAs is the row norms: norm_w = sum_k theta_k E_wk needs
//     the partial dot products of all CL column slices.  After the L-lane reduce-scatter of a
//     half-sweep, the lane that holds the partial norm of (batch, row) sends it to the same lane of
//     the same warp in every peer CTA: st.async into the peer's shared memory, completing on that
//     warp's mbarrier.  Both halves of a sweep are sent as soon as they are known; the warp waits
//     once, after the second half, and sums the CL partials in rank order, so every CTA of the
//     cluster gets bit-identical norms (and takes identical underflow decisions).  No cluster
//     barrier and no CTA barrier is involved: warps of different groups and different rows drift
//     freely, which is what lets several documents per SM overlap.  Buffers and mbarriers are used
//     in turn by the parity of a per-warp exchange counter (a peer can be at most one exchange
//     ahead of the data this warp has not read yet).
//   * Per document: the leader CTA's group leader pulls the document from the launch's work queue
//     and writes it into every CTA's shared memory; one cross-CTA group barrier (an mbarrier per
//     group in every CTA, one remote arrival per CTA) publishes it and doubles as "everybody is done
//     with the previous document".  The epilogue's three sums (sum gamma, sum lgamma(gamma), the
//     likelihood terms) take the same route.
//   * The last sweep computes all partial norms first, exchanges them once, then runs the fused
//     likelihood / statistics pass; rank 0 adds the row terms of the likelihood.
//
// The sweep loop itself is estep3_kernel.cuh's (estep3_sweep_fast / estep3_run_sweeps /
// estep3_dispatch instantiated with CL > 1); the exchange primitives are in estep3_xchg.cuh.  This
// file holds what differs per document: the shared-memory layout, the document loop with its
// cross-CTA group barrier, the last sweep and the launch glue.
#pragma once
#include <cooperative_groups.h>

#include "estep3_kernel.cuh"

namespace mrlda {

// shared-memory carve-up (offsets in doubles): the single-CTA layout plus the exchange state
template <int L, int KPL, int R, int NW, int CL>
struct Estep3cSmem {
  using S = Estep3Shape<L, KPL, R, NW>;
  static constexpr int XSLOTS = ESTEP3_XSLOTS;  // doubles per (parity, rank) of a warp's exchange buffer
  static_assert(2 * 32 <= XSLOTS && S::NBMAX * S::RPB <= XSLOTS, "exchange buffer too small");
  Estep3Smem<L, KPL, R, NW> base;
  size_t xbuf, xbar, gbar, gsum, total;
  __host__ __device__ Estep3cSmem(int M, int G) : base(M, G) {
    const int NG = NW / G;
    size_t o = base.total / sizeof(double);
    xbuf = o;  o += (size_t)NW * 2 * CL * XSLOTS;  // [warp][parity][rank][XSLOTS]
    xbar = o;  o += (size_t)NW * 8;   // per warp: [0..1] exchange mbarriers, [2] exchange counter,
                                      // [3] group-barrier parity, [4] slow-path mbarrier, [5] its parity
    gbar = o;  o += NG;               // per group: cross-CTA group barrier
    gsum = o;  o += (size_t)NG * CL * 4;  // per group and rank: epilogue sums
    total = o * sizeof(double);
  }
};

// TMA gather of four rows, columns [c0, c0 + box) of each
__device__ __forceinline__ void estep3c_tma_gather4(double *dst, const void *tmap, unsigned long long *mbar,
                                                    const int c0, const int r0, const int r1, const int r2,
                                                    const int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, "
      "%4, %5, %6}], [%7];" ::"r"(estep3_smem_u32(dst)),
      "l"(tmap), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(estep3_smem_u32(mbar))
      : "memory");
}

// last sweep, first pass: this CTA's partial norm of the rows of one batch (all L lanes of a row
// hold it)
template <int L, int KPL>
__device__ __forceinline__ double estep3c_last_dot(const double *theta_g, const double (&ev)[KPL], const int j) {
  constexpr int HP = KPL / 2;
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
  return p;
}
// last sweep, second pass: S, likelihood terms (DocumentMapper.java:421; rank 0 adds the row terms)
// and the statistics scatter of this CTA's columns (:315-329), given the complete norm p
template <int L, int KPL>
__device__ __forceinline__ void estep3c_last_apply(const EstepArgs &a, const double *theta_g, const double *cnt_w,
                                                   const double (&ev)[KPL], const double p, const int slot,
                                                   const int64_t zrow0, const int g, const int j,
                                                   double (&s)[KPL], double &lik, unsigned &badbits,
                                                   const int col0, const int kloc, const bool rank0) {
  constexpr int RPB = 32 / L;
  const double cn = cnt_w[slot * RPB + g];
  const bool ok = p >= MRLDA_NORM_TINY;
  if (cn > 0.0 && !ok) badbits |= 1u << slot;
  const double r = (cn > 0.0 && ok) ? cn * rcp_fast(ok ? p : 1.0) : 0.0;
#pragma unroll
  for (int c = 0; c < KPL; c++) s[c] = fma(r, ev[c], s[c]);
  if (r > 0.0) {
    const int term = __ldg(a.term_id + zrow0 + slot * RPB + g);
    if (j == 0 && rank0) lik = fma(cn, log(p) + __ldg(a.rowmax + term - 1), lik);
    if (a.learning)
      estep3_scatter<L, KPL>(a.stats + (size_t)(term - 1) * a.K + col0, theta_g, r, ev, j, kloc);
  }
}

template <int L, int KPL, int R, int NW, int CL>
__device__ __forceinline__ void estep3c_sweep_last(const EstepArgs &a, const Estep3Warp<L, KPL, R, NW> &w,
                                                   const Estep3cX &x,
                                                   const double (&e)[Estep3Shape<L, KPL, R, NW>::RA][KPL],
                                                   double (&s)[KPL], double &lik, unsigned &badbits, const int nb,
                                                   const int64_t zrow0, const int xc, const int lane,
                                                   const int col0, const int kloc) {
  using S = Estep3Shape<L, KPL, R, NW>;
  constexpr int XS = Estep3cSmem<L, KPL, R, NW, CL>::XSLOTS;
  constexpr int T = S::T, TW = S::TW, RPB = S::RPB;
  const int par = xc & 1;
  const int nwalk = nb > R ? nb : R;  // batch slots walked: the register batches always are
  if (lane == 0)
    estep3c_mbar_expect(x.xbar_u32 + (uint32_t)(par * sizeof(unsigned long long)),
                        (uint32_t)((CL - 1) * RPB * nwalk * sizeof(double)));
  // pass 1: partial norms of every batch, sent as they are known
#pragma unroll
  for (int q = 0; q < R; q++) {
    const double p = estep3c_last_dot<L, KPL>(w.theta_g, e[q], w.j);
    if (w.j == 0) estep3c_send<CL, XS>(x, par, q * RPB + w.g, p);
  }
#pragma unroll 1
  for (int t = 0; R + t < nb && t < T; t++) {
    uint32_t w32[TW];
    tmem_ld_words<TW>(w.tbase + (uint32_t)t * TW, w32);
    tmem_wait_ld();
    double ev[KPL];
#pragma unroll
    for (int c = 0; c < KPL; c++) ev[c] = estep3_w2d(w32[2 * c], w32[2 * c + 1]);
    const double p = estep3c_last_dot<L, KPL>(w.theta_g, ev, w.j);
    if (w.j == 0) estep3c_send<CL, XS>(x, par, (R + t) * RPB + w.g, p);
  }
#pragma unroll 1
  for (int m = 0; R + T + m < nb; m++) {
    double ev[KPL];
    estep3_tile_ld<L, KPL>(w.tile_w + (size_t)m * 32 * KPL, w.g, w.j, ev);
    const double p = estep3c_last_dot<L, KPL>(w.theta_g, ev, w.j);
    if (w.j == 0) estep3c_send<CL, XS>(x, par, (R + T + m) * RPB + w.g, p);
  }
  estep3c_mbar_wait(x.xbar_u32 + (uint32_t)(par * sizeof(unsigned long long)), (uint32_t)((xc >> 1) & 1));
  __syncwarp();  // the lanes j == 0 wrote this CTA's own partials
  // pass 2
  badbits = 0;
  const bool rank0 = x.crank == 0;
#pragma unroll
  for (int c = 0; c < KPL; c++) s[c] = 0.0;
#pragma unroll
  for (int q = 0; q < R; q++)
    estep3c_last_apply<L, KPL>(a, w.theta_g, w.cnt_w, e[q], estep3c_total<CL, XS>(x, par, q * RPB + w.g), q, zrow0,
                               w.g, w.j, s, lik, badbits, col0, kloc, rank0);
#pragma unroll 1
  for (int t = 0; R + t < nb && t < T; t++) {
    uint32_t w32[TW];
    tmem_ld_words<TW>(w.tbase + (uint32_t)t * TW, w32);
    tmem_wait_ld();
    double ev[KPL];
#pragma unroll
    for (int c = 0; c < KPL; c++) ev[c] = estep3_w2d(w32[2 * c], w32[2 * c + 1]);
    estep3c_last_apply<L, KPL>(a, w.theta_g, w.cnt_w, ev, estep3c_total<CL, XS>(x, par, (R + t) * RPB + w.g), R + t,
                               zrow0, w.g, w.j, s, lik, badbits, col0, kloc, rank0);
  }
#pragma unroll 1
  for (int m = 0; R + T + m < nb; m++) {
    double ev[KPL];
    estep3_tile_ld<L, KPL>(w.tile_w + (size_t)m * 32 * KPL, w.g, w.j, ev);
    estep3c_last_apply<L, KPL>(a, w.theta_g, w.cnt_w, ev, estep3c_total<CL, XS>(x, par, (R + T + m) * RPB + w.g),
                               R + T + m, zrow0, w.g, w.j, s, lik, badbits, col0, kloc, rank0);
  }
}

// Cross-CTA barrier of one warp group: a CTA-local group barrier, one remote arrival per CTA on the
// group's mbarrier in every CTA of the cluster, then everybody waits on its own CTA's mbarrier.
// What the arriving thread wrote into the peers' shared memory before arriving is visible after it.
template <int CL>
__device__ __forceinline__ void estep3c_group_arrive(const uint32_t gbar_u32) {
#pragma unroll
  for (int r = 0; r < CL; r++) estep3c_arrive_remote(estep3c_mapa(gbar_u32, r));
}
__device__ __forceinline__ void estep3c_group_wait(const uint32_t gbar_u32, unsigned long long *xbar_w,
                                                   const int lane) {
  const uint32_t parity = (uint32_t)xbar_w[3];
  estep3c_mbar_wait(gbar_u32, parity);
  __syncwarp();
  if (lane == 0) xbar_w[3] = parity ^ 1u;
  __syncwarp();
}

template <int L, int KPL, int R, int NW, int G, int CL>
__global__ void __launch_bounds__(NW * 32, 1) estep3c_kernel(const EstepArgs a) {
  namespace cg = cooperative_groups;
  using S = Estep3Shape<L, KPL, R, NW>;
  constexpr int RPB = S::RPB, KS = S::KS, TW = S::TW, T = S::T, NBMAX = S::NBMAX;
  constexpr int THREADS = S::THREADS;
  constexpr int XS = Estep3cSmem<L, KPL, R, NW, CL>::XSLOTS;
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int crank = (int)cg::this_cluster().block_rank();
  const int col0 = crank * KS;                                        // first topic of this CTA
  const int K = a.K;
  const int kloc = K - col0 < KS ? (K - col0 > 0 ? K - col0 : 0) : KS;  // topics of this CTA
  const int M = a.smem_batches;
  constexpr int GT = 32 * G;
  const int grp = warp / G, gw = warp % G;
  const Estep3cSmem<L, KPL, R, NW, CL> layc(M, G);
  const Estep3Smem<L, KPL, R, NW> &lay = layc.base;

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
  q.K = kloc;
  w.theta_g = q.theta_g;
  double *wtot_w = q.spart + (size_t)warp * KS;
  double *ssacc_s = smem + lay.ssacc;
  double *red_s = smem + lay.red;  // [warp][4]
  int *meta_s = reinterpret_cast<int *>(smem + lay.meta);  // [0] TMEM base, [1 + grp] document
  unsigned long long *xbar_w = reinterpret_cast<unsigned long long *>(smem + layc.xbar) + 8 * warp;
  unsigned long long *gbar = reinterpret_cast<unsigned long long *>(smem + layc.gbar) + grp;
  double *gsum_g = smem + layc.gsum + (size_t)grp * CL * 4;
  Estep3cX x;
  x.xbuf_w = smem + layc.xbuf + (size_t)warp * 2 * CL * XS;
  x.xbuf_u32 = estep3_smem_u32(x.xbuf_w);
  x.xbar_u32 = estep3_smem_u32(xbar_w);
  x.crank = crank;
  const uint32_t gbar_u32 = estep3_smem_u32(gbar);
  if (lane == 0) {
    unsigned long long *mb = reinterpret_cast<unsigned long long *>(smem + lay.mbar) + 2 * warp;
    estep3_mbar_init(mb, 1);  // TMA row gathers
    mb[1] = 0ull;
    estep3_mbar_init(xbar_w, 1);      // exchange, parity 0: the warp's own expect_tx arrival
    estep3_mbar_init(xbar_w + 1, 1);  // exchange, parity 1
    xbar_w[2] = 0ull;
    xbar_w[3] = 0ull;
    estep3_mbar_init(xbar_w + 4, CL - 1);  // slow-path rendezvous: one arrival per peer
    xbar_w[5] = 0ull;
    if (gw == 0) estep3_mbar_init(gbar, CL);  // group barrier: one arrival per CTA
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");

  for (int k = tid; k < KS; k += THREADS) {
    q.alpha_s[k] = k < kloc ? a.alpha[col0 + k] : 1.0;
    ssacc_s[k] = 0.0;
  }
  for (int k = tid; k < (NW / G) * KS; k += THREADS) {
    smem[lay.theta + k] = 0.0;
    smem[lay.gam + k] = 1.0;
    smem[lay.gextra + k] = 0.0;
  }
  if (warp == 0) tmem_alloc(reinterpret_cast<uint32_t *>(meta_s), 512);
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  w.tbase = (uint32_t)meta_s[0] + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * (uint32_t)S::TWORDS;
  const double lik_alpha = *a.lik_alpha;
  // every CTA's mbarriers exist before anybody sends
  cg::this_cluster().sync();

  for (;;) {
    // everybody, in every CTA, is done with the previous document; the leader CTA's group leader
    // pulls the next one for the whole cluster
    estep3_group_sync(q.bar_id, GT);
    if (q.gt == 0) {
      if (crank == 0) {
        const int v = a.doc_begin + atomicAdd(a.work_counter, 1);
        const uint32_t la = estep3_smem_u32(meta_s + 1 + grp);
#pragma unroll
        for (int r = 0; r < CL; r++) estep3c_st_remote_i32(estep3c_mapa(la, r), v);
      }
      estep3c_group_arrive<CL>(gbar_u32);
    }
    estep3c_group_wait(gbar_u32, xbar_w, lane);
    const int qi = meta_s[1 + grp];
    if (qi >= a.doc_end) break;
    const int d = a.doc_order[qi];
    const int64_t b = a.row_ptr[d];
    const int n = (int)(a.row_ptr[d + 1] - b);
    if (n <= 0) continue;  // DocumentMapper.java:197-201: null content, nothing is updated
    const int nbt = (n + RPB - 1) / RPB;
    const int nb_lo = nbt / G, nb_rem = nbt % G;
    const int nb = nb_lo + (gw < nb_rem ? 1 : 0);
    const int bstart = gw * nb_lo + (gw < nb_rem ? gw : nb_rem);
    const int64_t zrow0 = b + (int64_t)bstart * RPB;

    // ---- load my batches: this CTA's column slice of complete E rows ----
    for (int x2 = lane; x2 < NBMAX * RPB; x2 += 32) cnt_w[x2] = 0.0;
    __syncwarp();
    const int nsm = nb > R + T ? nb - (R + T) : 0;  // warp-uniform
    unsigned long long *mbar_w = reinterpret_cast<unsigned long long *>(smem + lay.mbar) + 2 * warp;
    if (nsm > 0 && lane == 0) {
      static_assert(RPB == 4, "tile::gather4 fetches the four rows of a batch");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      estep3_mbar_expect(mbar_w, (uint32_t)(nsm * RPB * KS * sizeof(double)));
      for (int m = 0; m < nsm; m++) {
        const int row0 = (bstart + R + T + m) * RPB;
        int id[4];
#pragma unroll
        for (int x2 = 0; x2 < 4; x2++) {
          const int row = row0 + (x2 < RPB ? x2 : RPB - 1);
          id[x2] = __ldg(a.term_id + b + (row < n ? row : n - 1)) - 1;
        }
        estep3c_tma_gather4(tile_w + (size_t)m * 32 * KPL, a.tmap_E, mbar_w, col0, id[0], id[1], id[2], id[3]);
      }
    }
    double e[S::RA][KPL];
#pragma unroll
    for (int bi = 0; bi < NBMAX; bi++) {
      if (bi < R || bi < nb || bi == 0) {
        const int row = (bstart + bi) * RPB + w.g;
        const bool valid = bi < nb && row < n;
        const int rowc = valid ? row : n - 1;
        if (valid && w.j == 0) cnt_w[bi * RPB + w.g] = (double)__ldg(a.count + b + row);
        if (bi < R + T) {
          double v[KPL];
          estep3_load_cols<L, KPL, true>(a.E + (size_t)(__ldg(a.term_id + b + rowc) - 1) * a.e_stride + col0, w.j, v);
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
      for (int col = q.gt; col < kloc; col += GT) {
        const double gam = a.use_stored_gamma ? a.gamma[(size_t)d * K + col0 + col] : q.alpha_s[col] + share;
        q.gam_g[col] = gam;
        q.theta_g[col] = exp_digamma(gam);
      }
    }
    if (nsm > 0) {
      const uint32_t parity = (uint32_t)mbar_w[1];
      estep3_mbar_wait(mbar_w, parity);
      __syncwarp();
      if (lane == 0) mbar_w[1] = parity ^ 1u;
    }
    estep3_group_sync(q.bar_id, GT);

    // ---- sweeps 1..I-1: loop body specialised on this warp's batch count ----
    estep3_dispatch<L, KPL, R, NW, G, (R > 1 ? R : 1), CL>(a, w, q, e, wtot_w, nb, zrow0, lane, x, xbar_w);

    // ---- last sweep, document epilogue (DocumentMapper.java:244-259,342-346) ----
    {
      double s[KPL];
      double lik = 0.0;
      unsigned badbits;
      const int xc = (int)xbar_w[2];
      estep3c_sweep_last<L, KPL, R, NW, CL>(a, w, x, e, s, lik, badbits, nb, zrow0, xc, lane, col0, kloc);
      __syncwarp();
      if (lane == 0) xbar_w[2] = (unsigned long long)(xc + 1);
      __syncwarp();
      if (__any_sync(0xffffffffu, badbits != 0))
        lik += estep3c_slow_rows<L, KS, CL>(a.term_id + zrow0, a.count + zrow0, a.logbeta,
                                            a.learning ? a.stats : nullptr, a.slow_rows, q.gam_g, q.gextra_g, K, nb,
                                            badbits, lane, true, crank, x.xbar_u32, xbar_w);
      estep3_publish_s<L, KPL>(s, wtot_w, lane);
      lik = warp_sum(lik);
      if (lane == 0) red_s[warp * 4 + 3] = lik;
      estep3_group_sync(q.bar_id, GT);
      double sg = 0.0, lg = 0.0, lk = 0.0;
      for (int col = q.gt; col < kloc; col += GT) {
        const double sk = estep3_colsum<KS, G>(q.spart, q.w0, col);
        const double add = q.theta_g[col] * sk;
        const double g_old = q.gam_g[col];
        const double g_new = q.alpha_s[col] + add + q.gextra_g[col];
        q.gextra_g[col] = 0.0;
        if (add != 0.0) lk -= digamma_fast(g_old) * add;
        q.gam_g[col] = g_new;
        sg += g_new;
        lg += lgamma(g_new);
        a.gamma[(size_t)d * K + col0 + col] = g_new;
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
      if (q.gt == 0) {  // this CTA's sums to every CTA of the cluster
        double tg = 0.0, tl = 0.0, tk = 0.0;
        for (int w2 = 0; w2 < G; w2++) {
          const double *rp = red_s + (q.w0 + w2) * 4;
          tg += rp[0];
          tl += rp[1];
          tk += rp[2] + rp[3];
        }
        const uint32_t la = estep3_smem_u32(gsum_g + crank * 4);
#pragma unroll
        for (int r = 0; r < CL; r++) {
          const uint32_t ra = estep3c_mapa(la, r);
          estep3c_st_remote_f64(ra, tg);
          estep3c_st_remote_f64(ra + 8, tl);
          estep3c_st_remote_f64(ra + 16, tk);
        }
        estep3c_group_arrive<CL>(gbar_u32);
      }
      estep3c_group_wait(gbar_u32, xbar_w, lane);
      double tg = 0.0, tl = 0.0, tk = 0.0;
#pragma unroll
      for (int r = 0; r < CL; r++) {  // rank order: the same bits in every CTA
        tg += gsum_g[r * 4 + 0];
        tl += gsum_g[r * 4 + 1];
        tk += gsum_g[r * 4 + 2];
      }
      const double dsg = digamma_literal(tg);
      for (int col = q.gt; col < kloc; col += GT) atomicAdd(ssacc_s + col, digamma_literal(q.gam_g[col]) - dsg);
      if (q.gt == 0 && crank == 0) {
        const double ll = lik_alpha + (tl - lgamma(tg)) + tk;
        const long long c = (long long)(-ll * 1e6);  // DocumentMapper.java:253-254
        atomicAdd(a.ll_counter, (unsigned long long)c);
      }
    }
  }
  // DocumentMapper.close (DocumentMapper.java:354-361)
  tmem_fence_before_sync();
  __syncthreads();
  for (int k = tid; k < kloc; k += THREADS) {
    const double v = ssacc_s[k];
    if (v != 0.0) atomicAdd(a.alpha_ss + col0 + k, v);
  }
  if (warp == 0) tmem_dealloc((uint32_t)meta_s[0], 512);
  cg::this_cluster().sync();  // nobody exits while a peer may still reach into its shared memory
}

template <int L, int KPL, int R, int NW, int CL>
static size_t variant3c_smem(int M, int G) {
  return Estep3cSmem<L, KPL, R, NW, CL>(M, G).total;
}
template <int L, int KPL, int R, int NW, int CL>
static cudaError_t variant3c_prepare(size_t smem, int G) {
  const int sz = (int)smem;
  switch (G) {
    case 8: return cudaFuncSetAttribute(estep3c_kernel<L, KPL, R, NW, 8, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, sz);
    case 4: return cudaFuncSetAttribute(estep3c_kernel<L, KPL, R, NW, 4, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, sz);
    case 2: return cudaFuncSetAttribute(estep3c_kernel<L, KPL, R, NW, 2, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, sz);
    case 1: return cudaFuncSetAttribute(estep3c_kernel<L, KPL, R, NW, 1, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, sz);
  }
  return cudaErrorInvalidValue;
}
template <int L, int KPL, int R, int NW, int G, int CL>
static cudaError_t variant3c_launch_g(const EstepArgs &a, int clusters, size_t smem, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent clusters pulling documents from the launch's work queue: no more than the GPU can
  // co-schedule
  cfg.gridDim = dim3((unsigned)CL);
  int max_clusters = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, estep3c_kernel<L, KPL, R, NW, G, CL>, &cfg);
  if (e != cudaSuccess) return e;
  if (max_clusters < 1) return cudaErrorLaunchOutOfResources;
  if (getenv("MRLDA_DEBUG"))
    fprintf(stderr, "estep3c<%d,%d,%d,%d,G=%d,CL=%d>: %d clusters wanted, smem %zu, max active clusters %d\n", L, KPL,
            R, NW, G, CL, clusters, smem, max_clusters);
  if (clusters > max_clusters) clusters = max_clusters;
  cfg.gridDim = dim3((unsigned)(clusters * CL));
  return cudaLaunchKernelEx(&cfg, estep3c_kernel<L, KPL, R, NW, G, CL>, a);
}
// `grid` counts clusters
template <int L, int KPL, int R, int NW, int CL>
static cudaError_t variant3c_launch(const EstepArgs &args, int grid, int G, int M, size_t smem, cudaStream_t st) {
  EstepArgs a = args;
  a.group_warps = G;
  a.smem_batches = M;
  switch (G) {
    case 8: return variant3c_launch_g<L, KPL, R, NW, 8, CL>(a, grid, smem, st);
    case 4: return variant3c_launch_g<L, KPL, R, NW, 4, CL>(a, grid, smem, st);
    case 2: return variant3c_launch_g<L, KPL, R, NW, 2, CL>(a, grid, smem, st);
    case 1: return variant3c_launch_g<L, KPL, R, NW, 1, CL>(a, grid, smem, st);
  }
  return cudaErrorInvalidValue;
}

// A cluster variant of the warp-group kernel: the Estep3Variant fields plus the cluster size; `grid`
// of launch() counts clusters.
struct Estep3cVariant {
  Estep3Variant v;
  int CL;
};
#define MRLDA_VARIANT3C(L_, KPL_, R_, NW_, CL_)                                                       \
  {                                                                                                   \
    {L_, KPL_, R_, Estep3Shape<L_, KPL_, R_, NW_>::T, NW_, Estep3Shape<L_, KPL_, R_, NW_>::MMAX,       \
     variant3c_launch<L_, KPL_, R_, NW_, CL_>, variant3c_smem<L_, KPL_, R_, NW_, CL_>,                \
     variant3c_prepare<L_, KPL_, R_, NW_, CL_>},                                                      \
        CL_                                                                                           \
  }

extern const Estep3cVariant g_estep3c_variants_C2[];
extern const int g_estep3c_variants_C2_n;
extern const Estep3cVariant g_estep3c_variants_C2o[];
extern const int g_estep3c_variants_C2o_n;
extern const Estep3cVariant g_estep3c_variants_C3[];
extern const int g_estep3c_variants_C3_n;
extern const Estep3cVariant g_estep3c_variants_C3o[];
extern const int g_estep3c_variants_C3o_n;
extern const Estep3cVariant g_estep3c_variants_C4[];
extern const int g_estep3c_variants_C4_n;
extern const Estep3cVariant g_estep3c_variants_C4o[];
extern const int g_estep3c_variants_C4o_n;
extern const Estep3cVariant g_estep3c_variants_C5[];
extern const int g_estep3c_variants_C5_n;
extern const Estep3cVariant g_estep3c_variants_C5o[];
extern const int g_estep3c_variants_C5o_n;

}  // namespace mrlda
