// estep_kernel.cuh — K1: the per-document variational E-step with the fused statistics scatter.
//
// Replaces DocumentMapper.map (DocumentMapper.java:175-350) + updatePhi (:402-429) and the phi
// emission that feeds Hadoop's shuffle (:315-329), for all documents of this rank's shard.
//
// Arithmetic.  The reference iterates in log space; this kernel uses the algebraically identical
// linear-space form (hard part 2).  With E_wk = exp(logbeta_wk - m_w),
// m_w = max_k logbeta_wk (both produced once per EM iteration by the finalize kernel) and
// theta_k = exp(digamma(gamma_k)):
//     norm_w  = sum_k theta_k E_wk                     (phi_wk = theta_k E_wk / norm_w)
//     gamma_k = alpha_k + theta_k * sum_w (n_w / norm_w) E_wk
// repeated sweep_cap-1 = 99 times (DocumentMapper.java:204,236,242 — no convergence exit), and on
// the last sweep
//     likelihoodPhi = sum_w n_w (log norm_w + m_w) - sum_k digamma(gamma_k) (gamma'_k - alpha_k)
//     stats[w][k]  += n_w phi_wk                        (what the mapper emits as log(n_w phi_wk))
// A (doc,term) row whose norm underflows (< 1e-200: every topic with weight on the term has
// theta == 0) is redone exactly in log space from logbeta, so the result never depends on the
// linear-space dynamic range.
//
// Mapping.  One CTA owns one document at a time (persistent CTAs pull documents, longest first,
// from an atomic queue).  The document's n_d x K tile of E rows is gathered once into shared memory
// and stays there for all 99 sweeps; rows beyond the tile capacity are streamed from L2 each sweep.
// A warp processes RPB = 32/L rows at a time, L lanes per row; lane j of a row group owns the
// KPL = KS/L columns {2(j + L i), 2(j + L i) + 1}, i < KPL/2, so one LDS.128 per lane per 16 bytes
// is bank-conflict free (row stride KS = L*KPL doubles, KPL = 2*odd for L < 8).  theta, the row
// chunk and the S accumulators live in registers: every E element is read from shared memory once
// per sweep and used by two DFMAs (norm and S).  Per sweep there are two CTA barriers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_math.cuh"

namespace mrlda {

struct EstepArgs {
  // corpus (device)
  const int64_t *row_ptr;     // D+1
  const int32_t *term_id;     // Z, 1-based
  const int32_t *count;       // Z
  const int32_t *doc_order;   // D, documents sorted by nnz descending
  const int32_t *doc_tokens;  // D, sum of counts
  int64_t n_docs;
  int doc_begin, doc_end;     // range of doc_order this launch works on
  // model (device)
  const double *E;        // V x e_stride, columns >= K are zero
  int e_stride;           // row stride of E in doubles (>= the kernel's KS, even)
  const double *rowmax;   // V
  const double *logbeta;  // V x K
  const double *alpha;    // K
  const double *lik_alpha;  // scalar: lngamma(sum alpha) - sum lngamma(alpha)
  // state / outputs (device)
  double *gamma;                  // D x K
  double *stats;                  // V x K, may be NULL when !learning
  double *alpha_ss;               // K
  unsigned long long *ll_counter; // int64 two's complement accumulated with wrapping adds
  int *work_counter;
  int *slow_rows;
  long long *phase_cycles;  // 16 counters, written only when built with -DMRLDA_PHASE_TIMING
  int K;
  int sweeps;
  int use_stored_gamma;
  int learning;
  int cap_rows;  // tile capacity in rows (multiple of RPB)
  int tmem_slots;  // register-stationary kernel: slots per thread kept in Tensor Memory
  int group_warps;   // warp-group kernel (estep3): warps per document
  int smem_batches;  // warp-group kernel (estep3): shared-memory batches per warp
  // deterministic statistics mode (estep_kernel<L, KPL, true> only): fixed-point accumulators, three
  // 64-bit limbs per cell (fx_add below); NULL otherwise
  const void *tmap_E;  // warp-group kernel (estep3): CUtensorMap of E in global memory (TMA row gathers)
  unsigned long long *stats_fx;     // V x K x 3
  unsigned long long *alpha_ss_fx;  // K x 3
};

// Order-independent accumulation for the deterministic statistics mode.  x is cut, exactly, into the
// integer floor(x 2^16) and a 64-bit fraction (units of 2^-80), and the three pieces
// {fraction low 32 bits, fraction high 32 bits, integer part} are added to three 64-bit limbs with
// integer atomics.  Integer addition is associative, so the limbs do not depend on the order in
// which documents, CTAs or ranks arrive; a limb takes 2^31 contributions before it can overflow
// (a cell receives at most one per document), and the limbs of several ranks add up limb by limb
// (ncclInt64).  Range |x| < 2^47, resolution 2^-80 = 8.3e-25 absolute.
__device__ __forceinline__ void fx_add(unsigned long long *cell, const double x) {
  const double xs = x * 65536.0;  // exact
  const double fl = floor(xs);
  const long long hi = __double2ll_rd(xs);
  const unsigned long long lo = __double2ull_rd((xs - fl) * 18446744073709551616.0);  // exact, < 2^64
  if (lo & 0xffffffffull) atomicAdd(cell, lo & 0xffffffffull);
  if (lo >> 32) atomicAdd(cell + 1, lo >> 32);
  if (hi) atomicAdd(cell + 2, (unsigned long long)hi);
}
// the value of a cell: (limb2 2^64 + limb1 2^32 + limb0) 2^-80, one rounding per part
__host__ __device__ __forceinline__ double fx_value(const unsigned long long *cell) {
  const unsigned long long l0 = cell[0], l1 = cell[1];
  const long long l2 = (long long)cell[2];
  const unsigned long long lo = l0 + (l1 << 32);
  const long long hi = l2 + (long long)(l1 >> 32) + (lo < l0 ? 1 : 0);
  return (double)hi * (1.0 / 65536.0) + (double)lo * (1.0 / 65536.0 / 18446744073709551616.0);
}

#define MRLDA_NORM_TINY 1e-200

template <int L, int KPL>
struct EstepShape {
  static constexpr int RPB = 32 / L;
  static constexpr int KS = L * KPL;
  static constexpr int HALF = KPL / 2;
  static constexpr int MAX_THREADS = KPL <= 22 ? 384 : 256;  // 168 vs 255 registers per thread
  // row blocks in flight per warp: as many as the register budget allows (3*KPL + U*KPL doubles)
  static constexpr int UNROLL = (KPL <= 14) ? 2 : ((KPL >= 24 && KPL <= 26) ? 2 : 1);
  static_assert(KPL % 2 == 0, "KPL must be even (16-byte units)");
  static_assert(L >= 8 || (KPL / 2) % 2 == 1, "KPL must be 2*odd for L < 8 (bank-conflict-free rows)");
};

// shared memory bytes for a CTA of nw warps and a tile of cap rows
template <int L, int KPL>
__host__ __device__ constexpr size_t estep_smem_bytes(int nw, int cap) {
  return sizeof(double) * ((size_t)cap * (L * KPL)  // tile
                           + (size_t)cap            // counts
                           + 4 * (size_t)(L * KPL)  // theta, gamma, gextra, ss_acc
                           + (size_t)nw * (L * KPL) // per-warp S partials
                           + 64);                   // reduction scratch + doc meta
}

// reduce-scatter of v[0..N) across the row groups of a warp (lanes that differ in bits >= L)
template <int N, int OFF, int L, int KPL>
__device__ __forceinline__ void reduce_scatter(double (&v)[KPL], int lane, int &base, int &cnt) {
  if constexpr (OFF >= L && OFF >= 1 && L < 32) {
    constexpr int H = (N + 1) / 2;
    const bool upper = (lane & OFF) != 0;
#pragma unroll
    for (int i = 0; i < H; i++) {
      double lo = v[i];
      double hi = (i + H < N) ? v[i + H] : 0.0;
      double send = upper ? lo : hi;
      double keep = upper ? hi : lo;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
    if (upper) {
      base += H;
      cnt = cnt > H ? cnt - H : 0;
    } else {
      cnt = cnt < H ? cnt : H;
    }
    if constexpr (OFF / 2 >= L) reduce_scatter<H, OFF / 2, L, KPL>(v, lane, base, cnt);
  }
}

template <int N, int OFF, int L>
struct RsFinal {
  static constexpr int value = (OFF >= L && L < 32) ? RsFinal<(N + 1) / 2, OFF / 2, L>::value : N;
};
template <int N, int L>
struct RsFinal<N, 0, L> {
  static constexpr int value = N;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One sweep's pass over the rows owned by this warp: U row blocks (RPB rows each) are processed
// together so that their load -> DFMA -> shuffle -> reciprocal chains overlap.  LAST adds the
// likelihood terms and the statistics scatter (DocumentMapper.java:421,315-329).
template <int L, int KPL, int U, bool LAST, bool DET>
__device__ __forceinline__ void sweep_rows(const EstepArgs &a, const double *tile,
                                           const double *cnt_s, const double *gam_s,
                                           double *gextra_s, const double (&th)[KPL],
                                           double (&s)[KPL], double &lik, const int64_t b,
                                           const int n, const int cap, const int nblocks,
                                           const int warp, const int NW, const int lane) {
  constexpr int RPB = 32 / L, KS = L * KPL, HALF = KPL / 2;
  const int g = lane / L, j = lane % L;
  const int K = a.K;
  for (int rb0 = warp; rb0 < nblocks; rb0 += U * NW) {
    double e[U][KPL];
    double cw[U], dot[U];
    bool valid[U];
    int rbu[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      int rb = rb0 + u * NW;
      const bool active = rb < nblocks;  // warp-uniform
      rb = active ? rb : nblocks - 1;
      rbu[u] = rb;
      const int w = rb * RPB + g;
      valid[u] = active && w < n;
      const int wc = w < n ? w : n - 1;
      if (rb * RPB < cap) {  // warp-uniform: rows resident in the tile
        const double *rowp = tile + (size_t)wc * KS + 2 * j;
#pragma unroll
        for (int i = 0; i < HALF; i++) {
          double2 v = *reinterpret_cast<const double2 *>(rowp + 2 * L * i);
          e[u][2 * i] = v.x;
          e[u][2 * i + 1] = v.y;
        }
        cw[u] = valid[u] ? cnt_s[wc] : 0.0;
      } else {  // long document: stream the row from L2
        const int64_t z = b + wc;
        const double *rowp = a.E + (size_t)(__ldg(a.term_id + z) - 1) * a.e_stride + 2 * j;
#pragma unroll
        for (int i = 0; i < HALF; i++) {
          double2 v = __ldg(reinterpret_cast<const double2 *>(rowp + 2 * L * i));
          e[u][2 * i] = v.x;
          e[u][2 * i + 1] = v.y;
        }
        cw[u] = valid[u] ? (double)__ldg(a.count + z) : 0.0;
      }
    }
    // norm_w = sum_k theta_k E_wk   (updatePhi's normalizeFactor, DocumentMapper.java:408-415)
#pragma unroll
    for (int u = 0; u < U; u++) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int i = 0; i < KPL; i += 4) {
        a0 = fma(th[i], e[u][i], a0);
        if (i + 1 < KPL) a1 = fma(th[i + 1], e[u][i + 1], a1);
        if (i + 2 < KPL) a2 = fma(th[i + 2], e[u][i + 2], a2);
        if (i + 3 < KPL) a3 = fma(th[i + 3], e[u][i + 3], a3);
      }
      dot[u] = (a0 + a1) + (a2 + a3);
    }
#pragma unroll
    for (int o = 1; o < L; o <<= 1) {
#pragma unroll
      for (int u = 0; u < U; u++) dot[u] += __shfl_xor_sync(0xffffffffu, dot[u], o);
    }
    bool bad[U];
    double r[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      bad[u] = valid[u] && !(dot[u] >= MRLDA_NORM_TINY);
      r[u] = (valid[u] && !bad[u]) ? cw[u] * rcp_fast(dot[u]) : 0.0;
    }
    // S_k += (n_w / norm_w) E_wk   (updateLogGamma fold, DocumentMapper.java:424-425)
#pragma unroll
    for (int u = 0; u < U; u++) {
#pragma unroll
      for (int i = 0; i < KPL; i++) s[i] = fma(r[u], e[u][i], s[i]);
    }

    if (LAST) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int w = rbu[u] * RPB + g;
        const int term = __ldg(a.term_id + b + (w < n ? w : n - 1));
        if (valid[u] && !bad[u] && j == 0)
          lik = fma(cw[u], log(dot[u]) + __ldg(a.rowmax + term - 1), lik);
        if (a.learning && valid[u] && !bad[u]) {
          double *srow = a.stats + (size_t)(term - 1) * K;
#pragma unroll
          for (int i = 0; i < KPL; i++) {
            int col = 2 * (j + L * (i >> 1)) + (i & 1);
            if (col < K) {
              if constexpr (DET)
                fx_add(a.stats_fx + 3 * ((size_t)(term - 1) * K + col), r[u] * th[i] * e[u][i]);
              else
                atomicAdd(srow + col, r[u] * th[i] * e[u][i]);
            }
          }
        }
      }
    }

#pragma unroll
    for (int u = 0; u < U; u++) {
      unsigned badmask = __ballot_sync(0xffffffffu, bad[u]);
      while (badmask) {  // exact log-space redo of the rows whose norm underflowed (rare)
        const int gi = (__ffs(badmask) - 1) / L;
        badmask &= ~((L == 32 ? 0xffffffffu : ((1u << L) - 1u)) << (gi * L));
        const int64_t z = b + rbu[u] * RPB + gi;
        const int term = __ldg(a.term_id + z);
        const double nw_ = (double)__ldg(a.count + z);
        const double *lb = a.logbeta + (size_t)(term - 1) * K;
        double mx = -INFINITY;
        for (int k = lane; k < K; k += 32) mx = fmax(mx, lb[k] + digamma_fast(gam_s[k]));
        mx = warp_max(mx);
        double sum = 0.0;
        for (int k = lane; k < K; k += 32) sum += exp(lb[k] + digamma_fast(gam_s[k]) - mx);
        sum = warp_sum(sum);
        const double lsum = log(sum);
        for (int k = lane; k < K; k += 32) {
          double lphi = lb[k] + digamma_fast(gam_s[k]) - mx - lsum;
          double c = nw_ * exp(lphi);
          if (c > 0.0) {
            atomicAdd(gextra_s + k, c);
            if (LAST) {
              lik += c * (lb[k] - lphi);  // DocumentMapper.java:421
              if (a.learning) {
                if constexpr (DET)
                  fx_add(a.stats_fx + 3 * ((size_t)(term - 1) * K + k), c);
                else
                  atomicAdd(a.stats + (size_t)(term - 1) * K + k, c);
              }
            }
          }
        }
        if (lane == 0) atomicAdd(a.slow_rows, 1);
      }
    }
  }
}

// DET = true is the deterministic statistics mode: the phi-weighted counts and the alpha sufficient
// statistics go to the fixed-point accumulators (fx_add), so the statistics are bit-identical from
// run to run and for any number of ranks.  (Rows on the exact log-space slow path still add their
// gamma contribution through shared-memory atomics in arrival order; no BASELINE shape takes it.)
template <int L, int KPL, bool DET = false>
__global__ void __launch_bounds__(EstepShape<L, KPL>::MAX_THREADS, 1)
estep_kernel(const EstepArgs a) {
  using S = EstepShape<L, KPL>;
  constexpr int RPB = S::RPB, KS = S::KS, HALF = S::HALF, UNROLL = S::UNROLL;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, NW = T >> 5;
  const int j = lane % L;
  const int K = a.K;
  const int cap = a.cap_rows;

  double *tile = smem;                     // cap x KS
  double *cnt_s = tile + (size_t)cap * KS; // cap
  double *theta_s = cnt_s + cap;           // KS
  double *gam_s = theta_s + KS;            // KS
  double *gextra_s = gam_s + KS;           // KS
  double *ssacc_s = gextra_s + KS;         // KS
  double *spart = ssacc_s + KS;            // NW x KS
  double *red_s = spart + (size_t)NW * KS; // 64
  int *meta_s = reinterpret_cast<int *>(red_s + 48);

  for (int k = tid; k < KS; k += T) {
    theta_s[k] = 0.0;
    gextra_s[k] = 0.0;
    ssacc_s[k] = 0.0;
    gam_s[k] = 1.0;
  }
  const double lik_alpha = *a.lik_alpha;

  for (;;) {
    __syncthreads();
    if (tid == 0) meta_s[0] = a.doc_begin + atomicAdd(a.work_counter, 1);
    __syncthreads();
    const int qi = meta_s[0];
    if (qi >= a.doc_end) break;
    const int d = a.doc_order[qi];
    const int64_t b = a.row_ptr[d];
    const int n = (int)(a.row_ptr[d + 1] - b);
    if (n <= 0) continue;  // DocumentMapper.java:197-201: null content, nothing is updated
    const int nt = n < cap ? n : cap;
    const int nblocks = (n + RPB - 1) / RPB;

    // ---- gather the document's E rows into the tile (the compulsory K*8 bytes per nnz) ----
    {
      constexpr int UNITS = KS / 2;
      const int total = nt * UNITS;
      for (int u = tid; u < total; u += T) {
        int r = u / UNITS, c = u - r * UNITS;
        const double *src = a.E + (size_t)(a.term_id[b + r] - 1) * a.e_stride + 2 * c;
        unsigned dst = (unsigned)__cvta_generic_to_shared(tile + (size_t)r * KS + 2 * c);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
      }
      for (int r = tid; r < nt; r += T) cnt_s[r] = (double)a.count[b + r];
      // gamma start (DocumentMapper.java:184-193)
      if (a.use_stored_gamma) {
        for (int k = tid; k < K; k += T) gam_s[k] = a.gamma[(size_t)d * K + k];
      } else {
        float share = (float)a.doc_tokens[d] / (float)K;
        for (int k = tid; k < K; k += T) gam_s[k] = a.alpha[k] + (double)share;
      }
      asm volatile("cp.async.commit_group;");
      asm volatile("cp.async.wait_group 0;");
    }

    double lik = 0.0;  // per-thread partial of likelihoodPhi (last sweep only)

    for (int sweep = 0; sweep < a.sweeps; sweep++) {
      const bool last = (sweep == a.sweeps - 1);
      // ---- theta_k = exp(digamma(gamma_k))  (DocumentMapper.java:208-211) ----
      for (int k = tid; k < K; k += T) theta_s[k] = exp_digamma(gam_s[k]);
      __syncthreads();  // B1: theta (and, first sweep, the tile) visible

      double th[KPL], s[KPL];
#pragma unroll
      for (int i = 0; i < HALF; i++) {
        double2 v = *reinterpret_cast<const double2 *>(theta_s + 2 * (j + L * i));
        th[2 * i] = v.x;
        th[2 * i + 1] = v.y;
        s[2 * i] = 0.0;
        s[2 * i + 1] = 0.0;
      }

      if (!last)
        sweep_rows<L, KPL, UNROLL, false, DET>(a, tile, cnt_s, gam_s, gextra_s, th, s, lik, b, n, cap, nblocks,
                                          warp, NW, lane);
      else
        sweep_rows<L, KPL, 1, true, DET>(a, tile, cnt_s, gam_s, gextra_s, th, s, lik, b, n, cap, nblocks,
                                    warp, NW, lane);

      // ---- S partials: reduce over the row groups of the warp, publish per warp ----
      {
        int base = 0, cnt = KPL;
        reduce_scatter<KPL, 16, L, KPL>(s, lane, base, cnt);
        constexpr int NF = RsFinal<KPL, 16, L>::value;
        double *sp = spart + (size_t)warp * KS;
#pragma unroll
        for (int i = 0; i < NF; i++) {
          if (i < cnt) {
            int idx = base + i;
            sp[2 * (j + L * (idx >> 1)) + (idx & 1)] = s[i];
          }
        }
      }
      __syncthreads();  // B2: S partials (and slow-path contributions) complete

      if (!last) {
        for (int k = tid; k < K; k += T) {
          double sk = 0.0;
          for (int w2 = 0; w2 < NW; w2++) sk += spart[(size_t)w2 * KS + k];
          gam_s[k] = a.alpha[k] + theta_s[k] * sk + gextra_s[k];  // DocumentMapper.java:232-234
          gextra_s[k] = 0.0;
        }
      }
    }

    // ---- document epilogue: ELBO, alpha statistics, gamma write-back (DocumentMapper.java:244-259,342-346)
    double sum_gamma = 0.0, sum_lg = 0.0;
    for (int k = tid; k < K; k += T) {
      double sk = 0.0;
      for (int w2 = 0; w2 < NW; w2++) sk += spart[(size_t)w2 * KS + k];
      const double add = theta_s[k] * sk;
      const double g_old = gam_s[k];
      const double g_new = a.alpha[k] + add + gextra_s[k];
      gextra_s[k] = 0.0;
      if (add != 0.0) lik -= digamma_fast(g_old) * add;
      gam_s[k] = g_new;
      sum_gamma += g_new;
      sum_lg += lgamma(g_new);
      a.gamma[(size_t)d * K + k] = g_new;
    }
    sum_gamma = warp_sum(sum_gamma);
    sum_lg = warp_sum(sum_lg);
    lik = warp_sum(lik);
    if (lane == 0) {
      red_s[warp] = sum_gamma;
      red_s[16 + warp] = sum_lg;
      red_s[32 + warp] = lik;
    }
    __syncthreads();
    double tg = 0.0, tl = 0.0, tk = 0.0;
    for (int w2 = 0; w2 < NW; w2++) {
      tg += red_s[w2];
      tl += red_s[16 + w2];
      tk += red_s[32 + w2];
    }
    const double dsg = digamma_literal(tg);
    for (int k = tid; k < K; k += T) {
      if constexpr (DET)
        fx_add(a.alpha_ss_fx + 3 * k, digamma_literal(gam_s[k]) - dsg);
      else
        ssacc_s[k] += digamma_literal(gam_s[k]) - dsg;
    }
    if (tid == 0) {
      double ll = lik_alpha + (tl - lgamma(tg)) + tk;
      long long c = (long long)(-ll * 1e6);  // DocumentMapper.java:253-254, truncation toward zero
      atomicAdd(a.ll_counter, (unsigned long long)c);
    }
  }

  // flush the CTA's alpha sufficient statistics (DocumentMapper.close, DocumentMapper.java:354-361)
  for (int k = tid; k < K; k += T) {
    double v = ssacc_s[k];
    if (v != 0.0) atomicAdd(a.alpha_ss + k, v);
  }
}

// host-side variant table -------------------------------------------------------------------------
struct EstepVariant {
  int L, KPL, max_threads;
  cudaError_t (*launch)(const EstepArgs &, int grid, int threads, int cap_rows, cudaStream_t);
  size_t (*smem_bytes)(int nw, int cap);
  cudaError_t (*prepare)(size_t smem);
  // the deterministic statistics mode (estep_kernel<L, KPL, true>)
  cudaError_t (*launch_det)(const EstepArgs &, int grid, int threads, int cap_rows, cudaStream_t);
  cudaError_t (*prepare_det)(size_t smem);
};

template <int L, int KPL>
static size_t variant_smem(int nw, int cap) {
  return estep_smem_bytes<L, KPL>(nw, cap);
}
template <int L, int KPL, bool DET>
static cudaError_t variant_prepare(size_t smem) {
  return cudaFuncSetAttribute(estep_kernel<L, KPL, DET>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)smem);
}
template <int L, int KPL, bool DET>
static cudaError_t variant_launch(const EstepArgs &args, int grid, int threads, int cap_rows,
                                  cudaStream_t st) {
  EstepArgs a = args;
  a.cap_rows = cap_rows;
  size_t smem = estep_smem_bytes<L, KPL>(threads / 32, cap_rows);
  estep_kernel<L, KPL, DET><<<grid, threads, smem, st>>>(a);
  return cudaGetLastError();
}

#define MRLDA_VARIANT(L_, KPL_)                                                       \
  {                                                                                   \
    L_, KPL_, EstepShape<L_, KPL_>::MAX_THREADS, variant_launch<L_, KPL_, false>,     \
        variant_smem<L_, KPL_>, variant_prepare<L_, KPL_, false>,                     \
        variant_launch<L_, KPL_, true>, variant_prepare<L_, KPL_, true>               \
  }

// each estep_inst_*.cu defines one of these tables
extern const EstepVariant g_estep_variants_L1[];
extern const int g_estep_variants_L1_n;
extern const EstepVariant g_estep_variants_L2[];
extern const int g_estep_variants_L2_n;
extern const EstepVariant g_estep_variants_L4[];
extern const int g_estep_variants_L4_n;
extern const EstepVariant g_estep_variants_L8[];
extern const int g_estep_variants_L8_n;
extern const EstepVariant g_estep_variants_L16[];
extern const int g_estep_variants_L16_n;
extern const EstepVariant g_estep_variants_L32[];
extern const int g_estep_variants_L32_n;

}  // namespace mrlda
