// estep2_kernel.cuh — K1, register-stationary variant: the E-step hot kernel for documents whose
// n_d x K tile fits on one SM (registers + shared memory).
//
// Same arithmetic as estep_kernel.cuh (linear-space fixed point, 99 sweeps, exact log-space redo of
// underflowing rows; reference: DocumentMapper.java:175-350, updatePhi :402-429); different mapping.
//
//   * Topics are split across the NW warps of the CTA: warp w owns the CPW columns
//     [w*CPW, (w+1)*CPW).  Rows (distinct terms of the document) are split across lanes: lane l owns
//     rows l, l+32, l+64, ... ("slots").  Thread (w,l) therefore owns the E elements
//     (row 32q+l, columns of warp w) for every slot q, and nobody else ever touches them.
//   * The first RREG slots live in REGISTERS for all 99 sweeps (the register file, 256 KB, is the
//     largest on-chip memory of the SM); the next slots live in Tensor Memory (thread-private
//     tcgen05.st / tcgen05.ld, no MMA), the rest in a thread-private region of shared memory.
//   * Per sweep: (1) the lane that owns a column computes theta_c = exp(digamma(gamma_c)) and the
//     warp broadcasts its CPW thetas; (2) every thread forms the partial dot products of its rows
//     over its warp's columns and publishes them; CTA barrier; (3a) one thread per row sums the NW
//     partials (norm_w) and publishes r_w = n_w / norm_w; CTA barrier; (3b) every thread
//     accumulates S_c += r_w E_wc over its rows; (4) the per-lane S vectors of a warp are
//     reduce-scattered in registers (shuffles) so that S_c lands in the lane that owns column c,
//     which updates gamma_c.  gamma, alpha and the alpha sufficient
//     statistics stay in the owner lane's registers.
//   * CL > 1: a thread-block cluster of CL CTAs (one per SM) owns the document; the topic split
//     continues across the CTAs (CTA r, warp w owns columns [(r*NW+w)*CPW, +CPW)), so K up to
//     CL*NW*CPW = 1024 fits on chip.  The only cross-CTA traffic is the per-row partial dot products,
//     read through distributed shared memory after one cluster barrier per sweep, and the document
//     index the leader CTA pulls from the work queue.
//   * WPS = warps per SM the variant is compiled for: 8 (255 registers per thread, most of the tile
//     in registers) or 16 (128 registers per thread, more of the tile in shared memory, twice the
//     warps to hide the FP64 / shared-memory latencies of the serial phases).
//   * Code that runs once per document or almost never — the statistics scatter, the epilogue, the
//     log-space slow path — is kept out of line (__noinline__): the sweep loop sits at the register
//     limit and ptxas' allocation for it flips (±9 % run time) with anything inlined next to it.
//   * Work is perfectly balanced across warps (every warp walks all rows); the only idle lanes are
//     those of the last, partially filled slot.
#pragma once
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>

#include "estep_kernel.cuh"
#include "tmem_ops.cuh"

#ifdef MRLDA_PHASE_TIMING
#define MRLDA_TICK(i)                    \
  do {                                   \
    asm volatile("" ::: "memory");       \
    tm[i] = clock64();                   \
  } while (0)
// tick that issues only after the value is available (a scoreboard dependency on its register)
#define MRLDA_TICK_DEP(i, v)                                          \
  do {                                                                \
    asm volatile("{.reg .b64 t_; mov.b64 t_, %0;}" ::"d"(v) : "memory"); \
    tm[i] = clock64();                                                \
  } while (0)
#else
#define MRLDA_TICK(i)
#define MRLDA_TICK_DEP(i, v)
#endif

namespace mrlda {

template <int NW, int CPW, int RREG, int WPS>
struct Estep2Shape {
  static constexpr int THREADS = NW * 32;
  static constexpr int CTAS_PER_SM = WPS / NW;
  static constexpr int KS = NW * CPW;
  static constexpr int HALF = CPW / 2;
  static_assert(CPW % 2 == 0 && CPW <= 32, "CPW must be even and at most 32");
  static_assert(NW == 1 || NW == 2 || NW == 4 || NW == 8 || NW == 16, "NW in {1,2,4,8,16}");
  static_assert((WPS == 8 || WPS == 16) && NW <= WPS, "WPS in {8,16}");
};

// TMEM slot stride in 32-bit columns: one tcgen05.ld.32x32b.xS fetches a slot (2*CPW words used)
__host__ __device__ constexpr int estep2_tmem_stride(int cpw) {
  return 2 * cpw > 32 ? 64 : (2 * cpw > 16 ? 32 : (2 * cpw > 8 ? 16 : 8));
}
// TMEM columns a CTA allocates (all co-resident CTAs together own the 512 columns of the SM) and
// the slots each thread gets out of them (warps w and w+4 share a lane quarter and split columns)
__host__ __device__ constexpr int estep2_tmem_cols(int nw, int wps) { return 512 / (wps / nw); }
// Columns of a slot that live in TMEM.  tcgen05.ld accepts any column offset (measured:
// tools/microbench/tmem_align.cu), so slots are packed without padding; when dropping up to two
// columns of the slice makes one more slot fit into the thread's TMEM columns, those "tail" columns
// go to a small thread-private shared-memory array instead (K=200: 24 of 26 columns = 48 words per
// slot, 5 slots in 256 words instead of 4).
__host__ __device__ constexpr int estep2_tmem_words_per_thread(int nw, int wps) {
  return estep2_tmem_cols(nw, wps) / (nw > 4 ? nw / 4 : 1);
}
__host__ __device__ constexpr int estep2_tmem_slot_cols(int nw, int cpw, int wps) {
  const int avail = estep2_tmem_words_per_thread(nw, wps);
  int best = cpw;
  for (int tc = cpw; tc >= cpw - 2 && tc >= 4; tc -= 2)  // multiples of 4 words only below
    if ((2 * tc) % 8 == 0 && avail / (2 * tc) > avail / (2 * best)) best = tc;
  return best;
}
__host__ __device__ constexpr int estep2_tmem_slots(int nw, int cpw, int wps) {
  return estep2_tmem_words_per_thread(nw, wps) / (2 * estep2_tmem_slot_cols(nw, cpw, wps));
}
// load / store W consecutive words of my TMEM lane as power-of-two vectors
template <int W, int OFF = 0>
__device__ __forceinline__ void tmem_ld_words(uint32_t addr, uint32_t *w) {
  if constexpr (W >= 64) {
    TmemVec<64>::ld(addr + OFF, *reinterpret_cast<uint32_t(*)[64]>(w + OFF));
    tmem_ld_words<W - 64, OFF + 64>(addr, w);
  } else if constexpr (W >= 32) {
    TmemVec<32>::ld(addr + OFF, *reinterpret_cast<uint32_t(*)[32]>(w + OFF));
    tmem_ld_words<W - 32, OFF + 32>(addr, w);
  } else if constexpr (W >= 16) {
    TmemVec<16>::ld(addr + OFF, *reinterpret_cast<uint32_t(*)[16]>(w + OFF));
    tmem_ld_words<W - 16, OFF + 16>(addr, w);
  } else if constexpr (W >= 8) {
    TmemVec<8>::ld(addr + OFF, *reinterpret_cast<uint32_t(*)[8]>(w + OFF));
    tmem_ld_words<W - 8, OFF + 8>(addr, w);
  } else if constexpr (W >= 4) {
    TmemVec<4>::ld(addr + OFF, *reinterpret_cast<uint32_t(*)[4]>(w + OFF));
    tmem_ld_words<W - 4, OFF + 4>(addr, w);
  } else if constexpr (W >= 2) {
    TmemVec<2>::ld(addr + OFF, *reinterpret_cast<uint32_t(*)[2]>(w + OFF));
    tmem_ld_words<W - 2, OFF + 2>(addr, w);
  } else {
    static_assert(W == 0, "TMEM slots hold doubles: an even number of words");
  }
}
template <int W, int OFF = 0>
__device__ __forceinline__ void tmem_st_words(uint32_t addr, const uint32_t *w) {
  if constexpr (W >= 64) {
    TmemVec<64>::st(addr + OFF, *reinterpret_cast<const uint32_t(*)[64]>(w + OFF));
    tmem_st_words<W - 64, OFF + 64>(addr, w);
  } else if constexpr (W >= 32) {
    TmemVec<32>::st(addr + OFF, *reinterpret_cast<const uint32_t(*)[32]>(w + OFF));
    tmem_st_words<W - 32, OFF + 32>(addr, w);
  } else if constexpr (W >= 16) {
    TmemVec<16>::st(addr + OFF, *reinterpret_cast<const uint32_t(*)[16]>(w + OFF));
    tmem_st_words<W - 16, OFF + 16>(addr, w);
  } else if constexpr (W >= 8) {
    TmemVec<8>::st(addr + OFF, *reinterpret_cast<const uint32_t(*)[8]>(w + OFF));
    tmem_st_words<W - 8, OFF + 8>(addr, w);
  } else if constexpr (W >= 4) {
    TmemVec<4>::st(addr + OFF, *reinterpret_cast<const uint32_t(*)[4]>(w + OFF));
    tmem_st_words<W - 4, OFF + 4>(addr, w);
  } else if constexpr (W >= 2) {
    TmemVec<2>::st(addr + OFF, *reinterpret_cast<const uint32_t(*)[2]>(w + OFF));
    tmem_st_words<W - 2, OFF + 2>(addr, w);
  } else {
    static_assert(W == 0, "TMEM slots hold doubles: an even number of words");
  }
}

// register-resident slots: what fits in 255 registers next to the working set (theta or S: 2*CPW
// registers, partial sums, addresses) without spilling E into local memory (-Xptxas -v)
__host__ __device__ constexpr int estep2_rreg(int cpw, int wps = 8) {
  int r = wps == 8 ? (255 - 75 - 2 * cpw) / (2 * cpw) : (128 - 44 - 2 * cpw) / (2 * cpw);
  return r < 1 ? 1 : (r > 8 ? 8 : r);
}

// shared memory bytes: rsm slots of thread-private E, double-buffered partials, small arrays
template <int NW, int CPW>
__host__ __device__ constexpr size_t estep2_smem_bytes(int rsm, int cap_rows) {
  return sizeof(double) * ((size_t)rsm * NW * 32 * CPW   // E slots
                           // tail columns of the TMEM slots (all variants here are WPS = 8)
                           + (size_t)estep2_tmem_slots(NW, CPW, 8) * NW * 32 *
                                 (CPW - estep2_tmem_slot_cols(NW, CPW, 8))
                           + 2 * (size_t)NW * cap_rows   // partial dots, two buffers
                           + 2 * (size_t)cap_rows        // counts, r
                           + 3 * (size_t)NW * CPW        // theta, gamma, gextra
                           + 96);                        // reductions + doc meta
}

// One thread's CPW-column slice of an E row -> registers.  Lanes hold different rows, so every
// warp-wide load touches 32 distinct lines and the load phase is bound by the number of load
// instructions (tools/microbench/rowload.cu: 13 x LDG.128 per slice 10 B/clk/SM, 6 x LDG.256 + 1 x
// LDG.128 16 B/clk/SM).  mode 0: 128-bit loads only (row stride not a multiple of 32 bytes);
// mode 1: the slice starts on a 32-byte boundary; mode 2: it starts 16 bytes past one.
__device__ __forceinline__ void ldg256(const double *p, double &a, double &b, double &c, double &d) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
template <int CPW>
__device__ __forceinline__ void estep2_load_slice(const double *src, const int mode, double (&v)[CPW]) {
  if (mode == 0) {
#pragma unroll
    for (int i = 0; i < CPW / 2; i++) {
      const double2 t = __ldg(reinterpret_cast<const double2 *>(src) + i);
      v[2 * i] = t.x;
      v[2 * i + 1] = t.y;
    }
  } else if (mode == 1) {
#pragma unroll
    for (int c = 0; c + 4 <= CPW; c += 4) ldg256(src + c, v[c], v[c + 1], v[c + 2], v[c + 3]);
    if constexpr (CPW % 4 != 0) {
      const double2 t = __ldg(reinterpret_cast<const double2 *>(src + CPW - 2));
      v[CPW - 2] = t.x;
      v[CPW - 1] = t.y;
    }
  } else {
    const double2 h = __ldg(reinterpret_cast<const double2 *>(src));
    v[0] = h.x;
    v[1] = h.y;
#pragma unroll
    for (int c = 2; c + 4 <= CPW; c += 4) ldg256(src + c, v[c], v[c + 1], v[c + 2], v[c + 3]);
    if constexpr ((CPW - 2) % 4 != 0) {
      const double2 t = __ldg(reinterpret_cast<const double2 *>(src + CPW - 2));
      v[CPW - 2] = t.x;
      v[CPW - 1] = t.y;
    }
  }
}

// Statistics scatter of the last sweep: n_w phi_wk = r_w theta_k E_wk (DocumentMapper.java:315-329).
// One warp per row, lanes across the CTA's columns, so that both the re-read of the E row (L2) and
// the atomics are contiguous; the tile's own layout (lanes = rows) would spread every warp-wide
// atomic over 32 rows.  Two rows and all their loads are in flight at a time.  Rows redone in log
// space (r < 0) are scattered by the slow path.  Not inlined: it runs once per document and must
// not take part in the register allocation of the sweep loop.
static __device__ __noinline__ void estep2_scatter_stats(const int32_t *term_id, const double *E,
                                                         const int e_stride, double *stats, const int K,
                                                         const double *r_s, const double *theta_s,
                                                         const int n, const int col0, const int KS,
                                                         const int warp, const int nw, const int lane) {
  for (int row = warp; row < n; row += 2 * nw) {
    const int row2 = row + nw;
    const double r0 = r_s[row], r1 = row2 < n ? r_s[row2] : 0.0;
    const bool ok0 = r0 > 0.0, ok1 = r1 > 0.0;
    const size_t t0 = ok0 ? (size_t)(__ldg(term_id + row) - 1) : 0;
    const size_t t1 = ok1 ? (size_t)(__ldg(term_id + row2) - 1) : 0;
    const double *e0 = E + t0 * e_stride + col0, *e1 = E + t1 * e_stride + col0;
    double v0[8], v1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int k = lane + 32 * i;
      const bool in = k < KS && col0 + k < K;
      v0[i] = ok0 && in ? __ldg(e0 + k) : 0.0;
      v1[i] = ok1 && in ? __ldg(e1 + k) : 0.0;
    }
    double *s0 = stats + t0 * K + col0, *s1 = stats + t1 * K + col0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int k = lane + 32 * i;
      if (k < KS && col0 + k < K) {
        const double th = theta_s[k];
        if (ok0) atomicAdd(s0 + k, r0 * th * v0[i]);
        if (ok1) atomicAdd(s1 + k, r1 * th * v1[i]);
      }
    }
  }
}

// Exact log-space redo (updatePhi as written, DocumentMapper.java:402-429) of the rows whose norm
// underflowed in the linear-space form, by one warp.  Returns the warp lane's share of
// likelihoodPhi on the last sweep.  Rare (no row of the BASELINE configurations takes it), so it is
// kept out of line and out of the sweep loop's register allocation.
template <int CL, int KS>
static __device__ __noinline__ double estep2_slow_rows(const int32_t *term_id, const int32_t *count,
                                                       const double *logbeta, double *stats,
                                                       int *slow_rows, double *gam_s, double *gextra_s,
                                                       const int K, const int nslots,
                                                       const unsigned badbits, const int lane,
                                                       const int crank, const bool last) {
  namespace cg = cooperative_groups;
  auto gamma_of = [&](int k) -> double {  // gamma_k lives in the CTA that owns topic k
    if constexpr (CL > 1)
      return cg::this_cluster().map_shared_rank(gam_s, k / KS)[k % KS];
    else
      return gam_s[k];
  };
  double lik = 0.0;
  for (int q = 0; q < nslots; q++) {
    unsigned m = __ballot_sync(0xffffffffu, (badbits >> q) & 1u);
    while (m) {
      const int l2 = __ffs(m) - 1;
      m &= m - 1;
      const int z = 32 * q + l2;
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
        if (k / KS != crank) continue;  // contributions go to the CTA that owns topic k
        double lphi = lb[k] + digamma_fast(gamma_of(k)) - mx - lsum;
        double c = nw_ * exp(lphi);
        if (c > 0.0) {
          gextra_s[k - crank * KS] += c;  // one warp, distinct k per lane: no race
          if (last) {
            lik += c * (lb[k] - lphi);  // DocumentMapper.java:421
            if (stats) atomicAdd(stats + (size_t)(term - 1) * K + k, c);
          }
        }
      }
      if (lane == 0 && crank == 0) atomicAdd(slow_rows, 1);
    }
  }
  return lik;
}

// Document epilogue (DocumentMapper.java:244-259,342-346): final gamma, the document's ELBO term
// into the job counter, and the owner lane's alpha sufficient statistic (returned).  Runs once per
// document and is kept out of line, away from the sweep loop's register allocation.
template <int NW, int CL>
static __device__ __noinline__ double estep2_epilogue(const bool owner, const double theta_c, const double S_c,
                                                      const double alpha_c, const double gam, double lik,
                                                      double *gextra, double *gamma_out, double *red_s,
                                                      const double lik_alpha, unsigned long long *ll_counter,
                                                      const int warp, const int lane, const int crank) {
  namespace cg = cooperative_groups;
  double g_new = 1.0, lg = 0.0;
  if (owner) {
    const double add = theta_c * S_c;
    g_new = alpha_c + add;
    if (gextra) {
      g_new += *gextra;
      *gextra = 0.0;
    }
    if (add != 0.0) lik -= digamma_fast(gam) * add;
    lg = lgamma(g_new);
    *gamma_out = g_new;
  }
  double sg = warp_sum(owner ? g_new : 0.0);
  lg = warp_sum(lg);
  lik = warp_sum(lik);
  if (lane == 0) {
    red_s[warp] = sg;
    red_s[16 + warp] = lg;
    red_s[32 + warp] = lik;
  }
  if constexpr (CL > 1)
    cg::this_cluster().sync();
  else
    __syncthreads();
  double tg = 0.0, tl = 0.0, tk = 0.0;
#pragma unroll
  for (int r = 0; r < CL; r++) {
    const double *rp = red_s;
    if constexpr (CL > 1) rp = cg::this_cluster().map_shared_rank(red_s, r);
#pragma unroll
    for (int w2 = 0; w2 < NW; w2++) {
      tg += rp[w2];
      tl += rp[16 + w2];
      tk += rp[32 + w2];
    }
  }
  if constexpr (CL > 1)
    cg::this_cluster().sync();  // red_s is rewritten by the next document
  else
    __syncthreads();
  if (warp == 0 && lane == 0 && crank == 0) {
    double ll = lik_alpha + (tl - lgamma(tg)) + tk;
    long long c = (long long)(-ll * 1e6);  // DocumentMapper.java:253-254
    atomicAdd(ll_counter, (unsigned long long)c);
  }
  return owner ? digamma_literal(g_new) - digamma_literal(tg) : 0.0;
}

template <int NW, int CPW, int RREG, int WPS, int CL, bool LAST>
__device__ __forceinline__ void estep2_sweep(
    const EstepArgs &a, double (&e)[RREG][CPW], const double *es, double *part, const double *cnt_s,
    double *r_s, int *flag_s, const double *theta_w, double (&s)[CPW], double &lik,
    unsigned &badbits, const int64_t b, const int n, const int nslots, const int cap_rows,
    const int warp, const int lane, long long *tm, const uint32_t tbase, const int rt,
    const int crank, const double *tail_s) {
  (void)tm;
  namespace cg = cooperative_groups;
  constexpr int HALF = CPW / 2;
  constexpr int TC = estep2_tmem_slot_cols(NW, CPW, WPS), TW = 2 * TC, T2 = CPW - TC;
  // slot q lives in registers (q < RREG), in TMEM (q < RREG + rt) or in shared memory
  const int nt_end = nslots < RREG + rt ? nslots : RREG + rt;
  // (2) partial dot products of my rows over my warp's columns.  Register slots are processed
  // unconditionally (an unused slot holds zeros) so that the chains of different slots interleave;
  // theta is read just in time (warp-wide broadcast loads) to keep registers for the tile.
  double *pw = part + (size_t)warp * cap_rows + lane;
  {
    double acc[RREG][2];
#pragma unroll
    for (int q = 0; q < RREG; q++) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll
    for (int i = 0; i < HALF; i++) {
      const double2 tv = *reinterpret_cast<const double2 *>(theta_w + 2 * i);
#pragma unroll
      for (int q = 0; q < RREG; q++) {
        acc[q][0] = fma(tv.x, e[q][2 * i], acc[q][0]);
        acc[q][1] = fma(tv.y, e[q][2 * i + 1], acc[q][1]);
      }
    }
#pragma unroll
    for (int q = 0; q < RREG; q++) pw[32 * q] = acc[q][0] + acc[q][1];
  }
  if (nslots > RREG) {  // slots outside the register file: theta held in registers for all of them
    double th[CPW];
#pragma unroll
    for (int i = 0; i < HALF; i++) {
      const double2 tv = *reinterpret_cast<const double2 *>(theta_w + 2 * i);
      th[2 * i] = tv.x;
      th[2 * i + 1] = tv.y;
    }
    // TMEM slots (a second load buffer for software pipelining does not fit the register file)
    for (int q = RREG; q < nt_end; q++) {
      uint32_t w32[TW];
      tmem_ld_words<TW>(tbase + (uint32_t)(q - RREG) * TW, w32);
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      if constexpr (T2 > 0) {  // tail columns of the slot: thread-private shared memory
        const double *tp = tail_s + ((size_t)((q - RREG) * NW + warp) * 32 + lane) * T2;
#pragma unroll
        for (int c = 0; c < T2; c += 2) {
          const double2 v = *reinterpret_cast<const double2 *>(tp + c);
          a2 = fma(th[TC + c], v.x, a2);
          a3 = fma(th[TC + c + 1], v.y, a3);
        }
      }
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < TC; c += 4) {
        a0 = fma(th[c], __hiloint2double((int)w32[2 * c + 1], (int)w32[2 * c]), a0);
        a1 = fma(th[c + 1], __hiloint2double((int)w32[2 * c + 3], (int)w32[2 * c + 2]), a1);
        if (c + 2 < TC) {
          a2 = fma(th[c + 2], __hiloint2double((int)w32[2 * c + 5], (int)w32[2 * c + 4]), a2);
          a3 = fma(th[c + 3], __hiloint2double((int)w32[2 * c + 7], (int)w32[2 * c + 6]), a3);
        }
      }
      pw[32 * q] = (a0 + a1) + (a2 + a3);
    }
    for (int q = RREG + rt; q < nslots; q++) {  // shared-memory slots
      const double2 *src = reinterpret_cast<const double2 *>(es) +
                           ((size_t)((q - RREG - rt) * NW + warp) * HALF) * 32 + lane;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int i = 0; i < HALF; i++) {
        const double2 v = src[(size_t)i * 32];
        a0 = fma(th[2 * i], v.x, a0);
        a1 = fma(th[2 * i + 1], v.y, a1);
      }
      pw[32 * q] = a0 + a1;
    }
  }
  MRLDA_TICK(2);
  const int nrows = 32 * (nslots > RREG ? nslots : RREG);
  if constexpr (CL > 1) {
    // distributed shared memory is slow (~20 B/clk): fold the NW partials of every row locally
    // first, so that each CTA reads one value per row and peer instead of NW
    __syncthreads();
    for (int row = warp * 32 + lane; row < nrows; row += NW * 32) {
      double *pp = part + row;
      double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
#pragma unroll
      for (int w2 = 0; w2 < NW; w2++) {
        const double v = pp[(size_t)w2 * cap_rows];
        if ((w2 & 3) == 0) n0 += v;
        else if ((w2 & 3) == 1) n1 += v;
        else if ((w2 & 3) == 2) n2 += v;
        else n3 += v;
      }
      pp[0] = (n0 + n1) + (n2 + n3);  // only this thread touches the partials of this row
    }
    cg::this_cluster().sync();  // B1 across the cluster: every CTA's row partials published
  } else {
    __syncthreads();  // B1: partial dots published
  }
  MRLDA_TICK(3);

  // (3a) one thread per row: norm_w = sum of the partials, r_w = n_w / norm_w; a row whose norm
  // underflowed is flagged with r = -1 and redone exactly in log space by the caller
  {
    for (int row = warp * 32 + lane; row < nrows; row += NW * 32) {
      double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
      if constexpr (CL > 1) {
#pragma unroll
        for (int r = 0; r < CL; r++) {  // fixed order: every CTA of the cluster gets the same bits
          const double v = cg::this_cluster().map_shared_rank(part, r)[row];
          if (r & 1) n1 += v;
          else n0 += v;
        }
      } else {
        const double *pp = part + row;
#pragma unroll
        for (int w2 = 0; w2 < NW; w2++) {
          const double v = pp[(size_t)w2 * cap_rows];
          if ((w2 & 3) == 0) n0 += v;
          else if ((w2 & 3) == 1) n1 += v;
          else if ((w2 & 3) == 2) n2 += v;
          else n3 += v;
        }
      }
      const double norm = (n0 + n1) + (n2 + n3);
      const bool valid = row < n;
      const bool bad = valid && !(norm >= MRLDA_NORM_TINY);
      const bool ok = valid && !bad;
      const double cn = cnt_s[row];
      r_s[row] = bad ? -1.0 : (ok ? cn : 0.0) * rcp_fast(ok ? norm : 1.0);
      if (bad) *flag_s = 1;
      if (LAST && ok && crank == 0)  // likelihoodPhi, first part (DocumentMapper.java:421 summed over topics)
        lik = fma(cn, log(norm) + __ldg(a.rowmax + __ldg(a.term_id + b + row) - 1), lik);
    }
  }
  MRLDA_TICK(4);
  __syncthreads();  // B2: r published
  MRLDA_TICK(5);

  // (3b) S_c += r_w E_wc over my rows
#pragma unroll
  for (int c = 0; c < CPW; c++) s[c] = 0.0;
  badbits = 0;
  {
    double r[RREG];
#pragma unroll
    for (int q = 0; q < RREG; q++) {
      r[q] = r_s[32 * q + lane];
      if (r[q] < 0.0) {
        badbits |= 1u << q;
        r[q] = 0.0;
      }
    }
#pragma unroll
    for (int c = 0; c < CPW; c++) {
#pragma unroll
      for (int q = 0; q < RREG; q++) s[c] = fma(r[q], e[q][c], s[c]);
    }
  }
  // TMEM slots
  for (int q = RREG; q < nt_end; q++) {
    uint32_t w32[TW];
    tmem_ld_words<TW>(tbase + (uint32_t)(q - RREG) * TW, w32);
    double r = r_s[32 * q + lane];
    if (r < 0.0) {
      badbits |= 1u << q;
      r = 0.0;
    }
    double tmp[CPW];
    if constexpr (T2 > 0) {
      const double *tp = tail_s + ((size_t)((q - RREG) * NW + warp) * 32 + lane) * T2;
#pragma unroll
      for (int c = 0; c < T2; c += 2) {
        const double2 v = *reinterpret_cast<const double2 *>(tp + c);
        tmp[TC + c] = v.x;
        tmp[TC + c + 1] = v.y;
      }
    }
    tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < TC; c++) tmp[c] = __hiloint2double((int)w32[2 * c + 1], (int)w32[2 * c]);
#pragma unroll
    for (int c = 0; c < CPW; c++) s[c] = fma(r, tmp[c], s[c]);
  }
  for (int q = RREG + rt; q < nslots; q++) {  // shared-memory slots
    double r = r_s[32 * q + lane];
    if (r < 0.0) {
      badbits |= 1u << q;
      r = 0.0;
    }
    const double2 *src = reinterpret_cast<const double2 *>(es) +
                         ((size_t)((q - RREG - rt) * NW + warp) * HALF) * 32 + lane;
#pragma unroll
    for (int i = 0; i < HALF; i++) {
      double2 v = src[(size_t)i * 32];
      s[2 * i] = fma(r, v.x, s[2 * i]);
      s[2 * i + 1] = fma(r, v.y, s[2 * i + 1]);
    }
  }
}

template <int NW, int CPW, int RREG, int WPS, int CL>
__global__ void __launch_bounds__(NW * 32, WPS / NW) estep2_kernel(const EstepArgs a) {
  namespace cg = cooperative_groups;
  constexpr int HALF = CPW / 2, KS = NW * CPW;  // KS: columns of this CTA
  int crank = 0;
  if constexpr (CL > 1) crank = (int)cg::this_cluster().block_rank();
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int K = a.K;
  constexpr int TC = estep2_tmem_slot_cols(NW, CPW, WPS), TW = 2 * TC, T2 = CPW - TC;
  const int rt = a.tmem_slots;              // slots per thread in TMEM (0: TMEM not used)
  const int cap_rows = a.cap_rows;          // 32 * (RREG + rt + rsm)
  const int rsm = cap_rows / 32 - RREG - rt;

  double *es = smem;                                        // rsm x NW x HALF x 32 double2
  double *tail_s = es + (size_t)rsm * NW * 32 * CPW;        // TMEM slots' tail columns
  double *part_s = tail_s + (size_t)estep2_tmem_slots(NW, CPW, 8) * NW * 32 * T2;  // 2 x NW x cap_rows
  double *cnt_s = part_s + 2 * (size_t)NW * cap_rows;       // cap_rows
  double *r_s = cnt_s + cap_rows;                           // cap_rows
  double *theta_s = r_s + cap_rows;                         // KS
  double *gam_s = theta_s + KS;                             // KS
  double *gextra_s = gam_s + KS;                            // KS
  double *red_s = gextra_s + KS;                            // 96: 3 x 16 sums, then doc meta
  int *meta_s = reinterpret_cast<int *>(red_s + 64);        // [0] document, [1] slow-path flag, [2] TMEM base

  // TMEM: this CTA's share of the SM's 512 columns, thread-private slots (tmem_ops.cuh)
  uint32_t tbase = 0;
  if (rt > 0) {
    if (warp == 0) tmem_alloc(reinterpret_cast<uint32_t *>(meta_s + 2), estep2_tmem_cols(NW, WPS));
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t cols_per_thread = estep2_tmem_words_per_thread(NW, WPS);
    tbase = (uint32_t)meta_s[2] + (((uint32_t)(warp & 3) * 32u) << 16) +
            (uint32_t)(warp >> 2) * cols_per_thread;
  }

  // Lane c owns column c of its warp (gamma_c, alpha_c, theta_c live in its registers).  The
  // register reduce-scatter leaves column rs_col in this lane; rs_src is the lane that ends up
  // with the column this lane owns.
  const int mycol = lane < CPW ? lane : -1;
  int rs_src = 0;
  {
    double dummy[CPW];
#pragma unroll
    for (int c = 0; c < CPW; c++) dummy[c] = 0.0;
    int base = 0, cnt = CPW;
    reduce_scatter<CPW, 16, 1, CPW>(dummy, lane, base, cnt);
    const int rs_col = cnt > 0 ? base : -1;
    for (int l2 = 0; l2 < 32; l2++) {
      int c2 = __shfl_sync(0xffffffffu, rs_col, l2);
      if (c2 == lane) rs_src = l2;
    }
  }
  const int gwarp = crank * NW + warp;
  const int ldmode = (a.e_stride & 3) ? 0 : (((gwarp * CPW) & 3) ? 2 : 1);  // see estep2_load_slice
  const int lcol = warp * CPW + mycol;   // column index inside this CTA
  const int col = gwarp * CPW + mycol;   // topic
  const bool owner = mycol >= 0 && col < K;
  const double alpha_c = owner ? a.alpha[col] : 1.0;
  double ss_acc = 0.0;  // my column's alpha sufficient statistic over all my documents
  const double lik_alpha = *a.lik_alpha;
  for (int k = tid; k < KS; k += NW * 32) {
    theta_s[k] = 0.0;
    gextra_s[k] = 0.0;
    gam_s[k] = 1.0;
  }
  if (tid == 0) meta_s[1] = 0;
  double *theta_w = theta_s + warp * CPW;
  // no CTA may write into a peer's shared memory before that peer has started
  if constexpr (CL > 1) cg::this_cluster().sync();

  int round = 0;
  for (;;) {
    __syncthreads();
    int qi;
    if constexpr (CL == 1) {
      if (tid == 0) meta_s[0] = a.doc_begin + atomicAdd(a.work_counter, 1);
      __syncthreads();
      qi = meta_s[0];
    } else {
      // all CTAs of a cluster must agree on the document: the leader CTA pulls it from the launch's
      // work queue and writes it into every CTA's shared memory.  Two slots, used in turn: the
      // leader can only write slot r again after the cluster barrier of round r + 1, which every
      // thread reaches after it has read slot r.
      if (crank == 0 && tid == 0) {
        const int v = a.doc_begin + atomicAdd(a.work_counter, 1);
#pragma unroll
        for (int r = 0; r < CL; r++) cg::this_cluster().map_shared_rank(meta_s, r)[4 + (round & 1)] = v;
      }
      cg::this_cluster().sync();
      qi = meta_s[4 + (round & 1)];
      round++;
    }
    if (qi >= a.doc_end) break;
#ifdef MRLDA_PHASE_TIMING
    const long long t_doc0 = clock64();
    long long t_doc1 = 0, t_doc2 = 0, t_doc3 = 0;
#endif
    const int d = a.doc_order[qi];
    const int64_t b = a.row_ptr[d];
    const int n = (int)(a.row_ptr[d + 1] - b);
    if (n <= 0) continue;  // DocumentMapper.java:197-201: null content, nothing is updated
    const int nslots = (n + 31) >> 5;

    // ---- load my elements: E rows of my slots, my warp's column slice (compulsory traffic) ----
    double e[RREG][CPW];
#pragma unroll
    for (int q = 0; q < RREG; q++) {
      const int row = 32 * q + lane;
      if (row < n) {
        estep2_load_slice<CPW>(a.E + (size_t)(__ldg(a.term_id + b + row) - 1) * a.e_stride + gwarp * CPW,
                               ldmode, e[q]);
      } else {
#pragma unroll
        for (int c = 0; c < CPW; c++) e[q][c] = 0.0;
      }
    }
    const int nt_end = nslots < RREG + rt ? nslots : RREG + rt;
    for (int q = RREG; q < nt_end; q++) {  // TMEM slots (+ their tail columns in shared memory)
      const int row = 32 * q + lane;
      uint32_t w32[TW];
#pragma unroll
      for (int i = 0; i < TW; i++) w32[i] = 0u;
      double *tp = tail_s + ((size_t)((q - RREG) * NW + warp) * 32 + lane) * (T2 > 0 ? T2 : 1);
      if constexpr (T2 > 0) {
#pragma unroll
        for (int c = 0; c < T2; c++) tp[c] = 0.0;
      }
      if (row < n) {
        double v[CPW];
        estep2_load_slice<CPW>(a.E + (size_t)(__ldg(a.term_id + b + row) - 1) * a.e_stride + gwarp * CPW,
                               ldmode, v);
#pragma unroll
        for (int c = 0; c < TC; c++) {
          w32[2 * c] = (uint32_t)__double2loint(v[c]);
          w32[2 * c + 1] = (uint32_t)__double2hiint(v[c]);
        }
        if constexpr (T2 > 0) {
#pragma unroll
          for (int c = 0; c < T2; c += 2) *reinterpret_cast<double2 *>(tp + c) = make_double2(v[TC + c], v[TC + c + 1]);
        }
      }
      tmem_st_words<TW>(tbase + (uint32_t)(q - RREG) * TW, w32);
    }
    if (nt_end > RREG) tmem_wait_st();
    for (int q = RREG + rt; q < nslots; q++) {  // shared-memory slots
      const int row = 32 * q + lane;
      double2 *dst = reinterpret_cast<double2 *>(es) +
                     ((size_t)((q - RREG - rt) * NW + warp) * HALF) * 32 + lane;
      if (row < n) {
        const double2 *src = reinterpret_cast<const double2 *>(
            a.E + (size_t)(__ldg(a.term_id + b + row) - 1) * a.e_stride + gwarp * CPW);
#pragma unroll
        for (int i = 0; i < HALF; i++) dst[(size_t)i * 32] = __ldg(src + i);
      } else {
#pragma unroll
        for (int i = 0; i < HALF; i++) dst[(size_t)i * 32] = make_double2(0.0, 0.0);
      }
    }
    {
      const int fill = 32 * (nslots > RREG ? nslots : RREG);  // register slots are always walked
      for (int r = tid; r < fill; r += NW * 32) cnt_s[r] = r < n ? (double)a.count[b + r] : 0.0;
    }
    // gamma start (DocumentMapper.java:184-193)
    double gam = 1.0;
    if (owner) {
      if (a.use_stored_gamma)
        gam = a.gamma[(size_t)d * K + col];
      else
        gam = alpha_c + (double)((float)a.doc_tokens[d] / (float)K);
    }
    __syncthreads();
#ifdef MRLDA_PHASE_TIMING
    t_doc1 = clock64();
#endif

    double lik = 0.0, theta_c = 0.0;
    double s[CPW];
    unsigned badbits = 0;
    for (int sweep = 0; sweep < a.sweeps; sweep++) {
      const bool last = (sweep == a.sweeps - 1);
#ifdef MRLDA_PHASE_TIMING
      if (last) t_doc2 = clock64();
#endif
      long long tm[10];
      (void)tm;
      MRLDA_TICK_DEP(0, gam);
      // (1) theta_c = exp(digamma(gamma_c)) in the owner lane, broadcast within the warp
      theta_c = owner ? exp_digamma(gam) : 0.0;
      MRLDA_TICK_DEP(8, theta_c);
      if (mycol >= 0) theta_w[mycol] = theta_c;
      __syncwarp();
      MRLDA_TICK(1);
      double *part = part_s + (size_t)(sweep & 1) * NW * cap_rows;
      if (!last)
        estep2_sweep<NW, CPW, RREG, WPS, CL, false>(a, e, es, part, cnt_s, r_s, meta_s + 1, theta_w, s, lik,
                                           badbits, b, n, nslots, cap_rows, warp, lane, tm, tbase, rt, crank, tail_s);
      else
        estep2_sweep<NW, CPW, RREG, WPS, CL, true>(a, e, es, part, cnt_s, r_s, meta_s + 1, theta_w, s, lik,
                                          badbits, b, n, nslots, cap_rows, warp, lane, tm, tbase, rt, crank, tail_s);
      MRLDA_TICK_DEP(6, s[CPW - 1]);
      // (4) S_c to the owner lane: reduce-scatter in registers (shuffles)
      double S_c;
      {
        int base = 0, cnt = CPW;
        reduce_scatter<CPW, 16, 1, CPW>(s, lane, base, cnt);
        S_c = __shfl_sync(0xffffffffu, s[0], rs_src);
      }
      MRLDA_TICK_DEP(7, S_c);
#ifdef MRLDA_PHASE_TIMING
      if (blockIdx.x == 0 && tid == 0 && !last) {
        for (int i = 0; i < 7; i++) a.phase_cycles[i] += tm[i + 1] - tm[i];
        a.phase_cycles[7] += 1;
        a.phase_cycles[8] += nslots;
        a.phase_cycles[9] += tm[8] - tm[0];  // exp(digamma) alone
      }
#endif
      // rows whose norm underflowed: exact log-space redo by warp 0 (CTA-uniform branch: every
      // warp computed the same norms)
      const bool anybad = meta_s[1] != 0;  // written before B2 by the row threads: CTA-uniform
      if (anybad) {
        auto csync = [&]() {
          if constexpr (CL > 1)
            cg::this_cluster().sync();
          else
            __syncthreads();
        };
        if (mycol >= 0) gam_s[lcol] = owner ? gam : 1.0;
        csync();
        if (tid == 0) meta_s[1] = 0;
        if (warp == 0)
          lik += estep2_slow_rows<CL, KS>(a.term_id + b, a.count + b, a.logbeta, a.learning ? a.stats : nullptr,
                                          a.slow_rows, gam_s, gextra_s, K, nslots, badbits, lane, crank, last);
        csync();
      }
      if (!last) {
        if (owner) {
          double gnew = fma(theta_c, S_c, alpha_c);  // DocumentMapper.java:232-234
          if (anybad) {
            gnew += gextra_s[lcol];
            gextra_s[lcol] = 0.0;
          }
          gam = gnew;
        }
      } else {
        // ---- document epilogue (DocumentMapper.java:244-259,342-346) ----
        ss_acc += estep2_epilogue<NW, CL>(owner, theta_c, S_c, alpha_c, gam, lik, anybad ? gextra_s + lcol : nullptr,
                                          owner ? a.gamma + (size_t)d * K + col : nullptr, red_s, lik_alpha,
                                          a.ll_counter, warp, lane, crank);
#ifdef MRLDA_PHASE_TIMING
        t_doc3 = clock64();
#endif
      }
    }
    // r and theta of the last sweep are still in shared memory (nothing rewrites them before the
    // barrier at the top of the document loop)
    if (a.learning)
      estep2_scatter_stats(a.term_id + b, a.E, a.e_stride, a.stats, K, r_s, theta_s, n, crank * KS, KS,
                           warp, NW, lane);
#ifdef MRLDA_PHASE_TIMING
    if (blockIdx.x == 0 && tid == 0) {
      a.phase_cycles[10] += t_doc1 - t_doc0;     // document load
      a.phase_cycles[11] += t_doc2 - t_doc1;     // sweeps 1..I-1
      a.phase_cycles[12] += t_doc3 - t_doc2;     // last sweep + epilogue
      a.phase_cycles[14] += clock64() - t_doc3;  // statistics scatter (warp 0's share)
      a.phase_cycles[13] += 1;
    }
#endif
  }
  // DocumentMapper.close (DocumentMapper.java:354-361)
  if (owner && ss_acc != 0.0) atomicAdd(a.alpha_ss + col, ss_acc);
  if constexpr (CL > 1) cg::this_cluster().sync();  // nobody may exit while its shared memory is read
  if (rt > 0) {
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc((uint32_t)meta_s[2], estep2_tmem_cols(NW, WPS));
  }
}

struct Estep2Variant {
  int NW, CPW, RREG, WPS, CL;
  cudaError_t (*launch)(const EstepArgs &, int grid, int cap_rows, size_t smem, cudaStream_t);
  size_t (*smem_bytes)(int rsm, int cap_rows);
  cudaError_t (*prepare)(size_t smem);
};

template <int NW, int CPW>
static size_t variant2_smem(int rsm, int cap_rows) {
  return estep2_smem_bytes<NW, CPW>(rsm, cap_rows);
}
template <int NW, int CPW, int RREG, int WPS, int CL>
static cudaError_t variant2_prepare(size_t smem) {
  return cudaFuncSetAttribute(estep2_kernel<NW, CPW, RREG, WPS, CL>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <int NW, int CPW, int RREG, int WPS, int CL>
static cudaError_t variant2_launch(const EstepArgs &args, int grid, int cap_rows, size_t smem,
                                   cudaStream_t st) {
  EstepArgs a = args;
  a.cap_rows = cap_rows;
  if constexpr (CL == 1) {
    estep2_kernel<NW, CPW, RREG, WPS, CL><<<grid, NW * 32, smem, st>>>(a);
    return cudaGetLastError();
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
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
    // persistent clusters pulling documents from the launch's work queue: no more clusters than the
    // GPU can co-schedule (cluster placement leaves some SMs unused: 132 of 148 CTAs at cluster
    // size 4)
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, estep2_kernel<NW, CPW, RREG, WPS, CL>, &cfg);
    if (e != cudaSuccess) return e;
    if (max_clusters < 1) return cudaErrorLaunchOutOfResources;
    if (getenv("MRLDA_DEBUG"))
      fprintf(stderr, "estep2<%d,%d,%d,%d,%d>: grid %d CTAs, smem %zu, cap %d rows, max active clusters %d\n",
              NW, CPW, RREG, WPS, CL, grid, smem, cap_rows, max_clusters);
    if (grid / CL > max_clusters) cfg.gridDim = dim3((unsigned)(max_clusters * CL));
    return cudaLaunchKernelEx(&cfg, estep2_kernel<NW, CPW, RREG, WPS, CL>, a);
  }
}

#define MRLDA_VARIANT2R(NW_, CPW_, RREG_, WPS_)                                                     \
  {                                                                                                 \
    NW_, CPW_, RREG_, WPS_, 1, variant2_launch<NW_, CPW_, RREG_, WPS_, 1>, variant2_smem<NW_, CPW_>, \
        variant2_prepare<NW_, CPW_, RREG_, WPS_, 1>                                                 \
  }
#define MRLDA_VARIANT2(NW_, CPW_, WPS_) MRLDA_VARIANT2R(NW_, CPW_, estep2_rreg(CPW_, WPS_), WPS_)
// cluster variants: CL CTAs of NW warps x 255 registers per document, one CTA per SM.  (4-warp CTAs
// at twice the cluster size put two documents on every SM but hold the same number of documents
// on the GPU: measured equal, and not instantiated.)
#define MRLDA_VARIANT2CN(NW_, CPW_, CL_)                                                   \
  {                                                                                        \
    NW_, CPW_, estep2_rreg(CPW_, 8), 8, CL_,                                               \
        variant2_launch<NW_, CPW_, estep2_rreg(CPW_, 8), 8, CL_>, variant2_smem<NW_, CPW_>, \
        variant2_prepare<NW_, CPW_, estep2_rreg(CPW_, 8), 8, CL_>                          \
  }
#define MRLDA_VARIANT2C(CPW_, CL_) MRLDA_VARIANT2CN(8, CPW_, CL_)

extern const Estep2Variant g_estep2_variants_W1[];
extern const int g_estep2_variants_W1_n;
extern const Estep2Variant g_estep2_variants_W2[];
extern const int g_estep2_variants_W2_n;
extern const Estep2Variant g_estep2_variants_W4[];
extern const int g_estep2_variants_W4_n;
extern const Estep2Variant g_estep2_variants_W8[];
extern const int g_estep2_variants_W8_n;
extern const Estep2Variant g_estep2_variants_C2[];
extern const int g_estep2_variants_C2_n;
extern const Estep2Variant g_estep2_variants_C4[];
extern const int g_estep2_variants_C4_n;
// the wide cluster class: twice the cluster size at half the column slice (longer documents)
extern const Estep2Variant g_estep2_variants_X4[];
extern const int g_estep2_variants_X4_n;
extern const Estep2Variant g_estep2_variants_X8[];
extern const int g_estep2_variants_X8_n;

}  // namespace mrlda
