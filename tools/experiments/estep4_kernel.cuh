// estep4_kernel.cuh — K1, paired variant: TWO documents per CTA in lockstep, software-pipelined.
//
// Same arithmetic as the other E-step kernels (linear-space fixed point, 99 sweeps, exact log-space
// redo of underflowing rows; reference: DocumentMapper.java:175-350, updatePhi :402-429) and the
// same building blocks as estep3_kernel.cuh (complete rows per warp, L lanes per row, batches in
// registers / Tensor Memory / shared memory, reduce-scatter norms, warp sums of S).  What changes:
//
//   * A sweep of ONE document is a strictly serial chain: dot products -> norms -> S -> exchange ->
//     barrier -> gamma, exp(digamma) -> barrier -> theta load.  With one document per warp group
//     about half of a sweep is spent in the phases after "S", where the FP64 pipe has nothing to do
//     (profiles/r2_phase_*.log).  Every reference document runs the same number of sweeps
//     (DocumentMapper.java:242 has no convergence exit), so two documents A and B can share all
//     eight warps of a CTA and alternate: while the warps run the dot products and the S pass of
//     A, the column owners evaluate gamma / exp(digamma) of B as independent instructions in the
//     same straight-line code, and vice versa.  One CTA barrier per document and sweep instead of
//     two, and no exposed exp(digamma) latency.
//   * Slots.  A warp has NS batch slots: register batch 0, TLO Tensor Memory batches, M
//     shared-memory batches, the other T - TLO Tensor Memory batches, register batch 1.  Document
//     A fills them from the bottom, B from the top, so both use their fast slots first and the
//     slow shared-memory slots in the middle go to whoever needs them.  A's batches are dealt to
//     the warps with the extra ones at the low warps, B's with the extra ones at the high warps:
//     the sum per warp is balanced to one batch.  The host pairs the documents (longest with the
//     longest that still fits, mrlda_capi.cu) and a document that found no partner runs alone.
//   * The loop body of a half sweep is specialised on the number of batches the warp holds of
//     that document (compile time: straight-line code), chosen once per pair.
#pragma once
#include <type_traits>

#include "estep3_kernel.cuh"

namespace mrlda {

// tcgen05.ld whose destination registers are read-write operands: a load under a warp-uniform branch
// then updates its buffer in place, instead of defining new registers that have to be copied into
// the buffer's registers where the branches join (32 moves per skipped-or-not load).
template <int N>
struct TmemVecRW;
template <>
struct TmemVecRW<2> {
  static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[2]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];"
                 : "+r"(r[0]), "+r"(r[1])
                 : "r"(taddr));
  }
};
template <>
struct TmemVecRW<4> {
  static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3])
                 : "r"(taddr));
  }
};
template <>
struct TmemVecRW<8> {
  static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 : "r"(taddr));
  }
};
template <>
struct TmemVecRW<16> {
  static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 : "r"(taddr));
  }
};
template <>
struct TmemVecRW<32> {
  static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 : "r"(taddr));
  }
};
template <int W, int OFF = 0>
__device__ __forceinline__ void tmem_ldrw_words(uint32_t addr, uint32_t *w) {
  if constexpr (W >= 32) {
    TmemVecRW<32>::ld(addr + OFF, *reinterpret_cast<uint32_t(*)[32]>(w + OFF));
    tmem_ldrw_words<W - 32, OFF + 32>(addr, w);
  } else if constexpr (W >= 16) {
    TmemVecRW<16>::ld(addr + OFF, *reinterpret_cast<uint32_t(*)[16]>(w + OFF));
    tmem_ldrw_words<W - 16, OFF + 16>(addr, w);
  } else if constexpr (W >= 8) {
    TmemVecRW<8>::ld(addr + OFF, *reinterpret_cast<uint32_t(*)[8]>(w + OFF));
    tmem_ldrw_words<W - 8, OFF + 8>(addr, w);
  } else if constexpr (W >= 4) {
    TmemVecRW<4>::ld(addr + OFF, *reinterpret_cast<uint32_t(*)[4]>(w + OFF));
    tmem_ldrw_words<W - 4, OFF + 4>(addr, w);
  } else if constexpr (W >= 2) {
    TmemVecRW<2>::ld(addr + OFF, *reinterpret_cast<uint32_t(*)[2]>(w + OFF));
    tmem_ldrw_words<W - 2, OFF + 2>(addr, w);
  } else {
    static_assert(W == 0, "TMEM slots hold doubles: an even number of words");
  }
}


template <int L, int KPL, int NW>
struct Estep4Shape {
  static constexpr int RPB = 32 / L;
  static constexpr int KS = L * KPL;
  static constexpr int TW = 2 * KPL;
  static constexpr int WQ = NW / 4;
  static constexpr int TWORDS = 512 / WQ;
  static constexpr int T = TWORDS / TW;     // Tensor Memory batches per warp
  static constexpr int M = 3;               // shared-memory batches per warp
  static constexpr int NS = 2 + T + M;      // slots per warp
  static constexpr int TLO = (T + 1) / 2;   // Tensor Memory batches on A's side of the shared-memory slots
  static constexpr int CA = 16, CB = KPL - CA;
  static constexpr int THREADS = NW * 32;
  static_assert(L == 8 && NW == 8, "written for 8 lanes per row and 8 warps");
  static_assert(KPL > CA && KPL <= 32, "KPL in (16, 32]");
  // storage of a slot: 0 register batch, 1 Tensor Memory, 2 shared memory; and its index there
  static __host__ __device__ constexpr int kind(int slot) {
    return (slot == 0 || slot == NS - 1) ? 0 : (slot <= TLO ? 1 : (slot <= TLO + M ? 2 : 1));
  }
  static __host__ __device__ constexpr int index(int slot) {
    return slot == 0 ? 0 : (slot == NS - 1 ? 1 : (slot <= TLO ? slot - 1 : (slot <= TLO + M ? slot - 1 - TLO : slot - 1 - M)));
  }
};

template <int L, int KPL, int NW>
struct Estep4Smem {
  using S = Estep4Shape<L, KPL, NW>;
  size_t tile, wtot, theta, gam, gextra, alpha, ssacc, cnt, red, meta, total;
  __host__ __device__ Estep4Smem() {
    size_t o = 0;
    tile = o;   o += (size_t)NW * S::M * 32 * KPL;
    wtot = o;   o += (size_t)2 * NW * S::KS;       // [doc][warp][column]
    theta = o;  o += 2 * S::KS;
    gam = o;    o += 2 * S::KS;
    gextra = o; o += 2 * S::KS;
    alpha = o;  o += S::KS;
    ssacc = o;  o += S::KS;
    cnt = o;    o += (size_t)2 * NW * S::NS * S::RPB;  // [doc][warp][batch][row]: n_w (0: padding)
    red = o;    o += NW * 4;
    meta = o;   o += 8;
    total = o * sizeof(double);
  }
};

template <int I, int N, class F>
__device__ __forceinline__ void estep4_for(F &&f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    estep4_for<I + 1, N>(f);
  }
}

// per-warp view of one document of the pair
struct Estep4Doc {
  double *theta, *gam, *gextra, *wtot;  // this document's shared state (wtot: [warp][column])
  double *cnt_w;                         // n_w of this warp's batches, batch order
  int64_t zrow0;                         // first CSR entry of this warp's rows
  int d, n, nb;                          // document, its rows, this warp's batches (>= 1: padding batch)
  int64_t b;                             // first CSR entry of the document
};

struct Estep4Warp {
  uint32_t tbase;
  const double *tile_w;
  int g, j, lane, warp, tid, K;
  const double *alpha_s;
};

// The column owner's update of the OTHER document (thread tid owns column tid): load() issues the
// loads, compute() is register arithmetic that the scheduler spreads over this document's S pass,
// store() writes gamma and the next theta.  Loads and arithmetic are unconditional (threads beyond
// K work on column K - 1, and a half sweep that has nothing to update works on stale values): only
// the stores are predicated, so that the half sweep stays one basic block.
template <int KS, int NW>
struct Estep4Owner {
  double part[NW], th, al, gx, g, t;
  __device__ __forceinline__ void load(const Estep4Doc &o, const Estep4Warp &w) {
    const int cc = w.tid < w.K ? w.tid : w.K - 1;
#pragma unroll
    for (int x = 0; x < NW; x++) part[x] = o.wtot[(size_t)x * KS + cc];
    th = o.theta[cc];
    al = w.alpha_s[cc];
    gx = o.gextra[cc];
  }
  __device__ __forceinline__ void compute() {
#pragma unroll
    for (int n = NW / 2; n >= 1; n >>= 1) {
#pragma unroll
      for (int x = 0; x < n; x++) part[x] += part[x + n];
    }
    g = fma(th, part[0], al) + gx;  // DocumentMapper.java:232-234
    const double in[1] = {g};
    double out[1];
    exp_digamma_n<1>(in, out);
    t = out[0];
  }
  __device__ __forceinline__ void store(const Estep4Doc &o, const Estep4Warp &w, const bool upd) {
    if (upd && w.tid < w.K) {
      o.gextra[w.tid] = 0.0;
      o.gam[w.tid] = g;
      o.theta[w.tid] = t;
    }
  }
};

// order of a document's batches over the slots FIRST, FIRST + STEP, ...
template <class S, int FIRST, int STEP>
struct Estep4Seq {
  // position of the first Tensor Memory batch at or after batch i (NS: none)
  static __host__ __device__ constexpr int next_tm(int i) {
    for (; i < S::NS; i++)
      if (S::kind(FIRST + STEP * i) == 1) return i;
    return S::NS;
  }
};

// the column owners' update of document o: gamma and the next theta from the published warp sums
template <int L, int KPL, int NW>
__device__ __forceinline__ void estep4_update(const Estep4Warp &w, const Estep4Doc &o, const bool upd) {
  Estep4Owner<L * KPL, NW> own;
  own.load(o, w);
  own.compute();
  own.store(o, w, upd);
}

// One half sweep (sweeps 1..I-1): the nb batches of document x held by this warp, in the slots
// FIRST, FIRST + STEP, ..., and the column owners' update of the other document o.
//   * ONE body per document for every batch count: the code of batch i is skipped when i >= nb
//     (warp-uniform branch).  A body per count, as in estep3_kernel.cuh, makes eight warps cycle
//     through ~90 KB of straight-line code per sweep and the kernel waits for instruction fetches
//     (measured: 1.3 "no instruction" stall cycles per issued instruction).
//   * The update of o is a ~60-instruction dependent chain.  The warps w < 4 run it before the
//     passes and the warps w >= 4 after them: each scheduler hosts warps w and w + 4, so one of its
//     two warps streams DFMAs while the other one waits on the chain.
// Ends with x's warp sums of S published; the caller places the CTA barrier.
template <int L, int KPL, int NW, int FIRST, int STEP>
__device__ __forceinline__ void estep4_half(const EstepArgs &a, const Estep4Warp &w, const Estep4Doc &x,
                                            const Estep4Doc &o, const bool upd, const double (&e)[2][KPL],
                                            const double mycnt1, const double mycnt2) {
  using S = Estep4Shape<L, KPL, NW>;
  using Q = Estep4Seq<S, FIRST, STEP>;
  constexpr int TW = S::TW, CA = S::CA, CB = S::CB, NS = S::NS;
  constexpr int N1 = (NS + 1) / 2, N2 = NS - N1;
  static_assert(N1 <= L && N2 >= 1, "two reduce-scatters of at most L batches");
  auto taddr = [&](int i) { return w.tbase + (uint32_t)(S::index(FIRST + STEP * i) * TW); };
  const int nb = x.nb;
  const bool chain_first = w.warp < NW / 2;
  // pr: partial dot products of the batches, replaced in place by r_w = n_w / norm_w (1 resp. 0 in
  // the slots the warp does not use)
  double pr[NS], s[KPL];
  double(&rr)[NS] = pr;
  uint32_t wa[2 * CA], wb[2 * CB];  // staging of the Tensor Memory batches (read-write asm operands)
  unsigned badbits = 0;
  const int j = w.j, lane = w.lane;
#pragma unroll
  for (int i = 0; i < NS; i++) pr[i] = 1.0;
#pragma unroll
  for (int i = 0; i < 2 * CA; i++) wa[i] = 0u;
#pragma unroll
  for (int i = 0; i < 2 * CB; i++) wb[i] = 0u;
  if (chain_first) estep4_update<L, KPL, NW>(w, o, upd);
  // ---- pass A: theta in registers; partial dot products of every batch ----
  {
    double th[KPL];
    estep3_load_cols<L, KPL, false>(x.theta, j, th);
    constexpr int t0 = Q::next_tm(0);
    if (t0 < nb) tmem_ldrw_words<2 * CA>(taddr(t0), wa);
    estep4_for<0, NS>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      constexpr int slot = FIRST + STEP * i, kd = S::kind(slot), ix = S::index(slot);
      if (i == 0 || i < nb) {
        double ac[4] = {0.0, 0.0, 0.0, 0.0};
        if constexpr (kd == 0) {
#pragma unroll
          for (int c = 0; c < KPL; c++) ac[c & 3] = fma(th[c], e[ix][c], ac[c & 3]);
        } else if constexpr (kd == 1) {
          tmem_wait_ld();
          tmem_ldrw_words<2 * CB>(taddr(i) + 2 * CA, wb);
#pragma unroll
          for (int c = 0; c < CA; c++) ac[c & 3] = fma(th[c], estep3_w2d(wa[2 * c], wa[2 * c + 1]), ac[c & 3]);
          tmem_wait_ld();
          constexpr int tn = Q::next_tm(i + 1);
          if (tn < nb) tmem_ldrw_words<2 * CA>(taddr(tn), wa);
#pragma unroll
          for (int c = 0; c < CB; c++)
            ac[c & 3] = fma(th[CA + c], estep3_w2d(wb[2 * c], wb[2 * c + 1]), ac[c & 3]);
        } else {
          double ev[KPL];
          estep3_tile_ld<KPL>(w.tile_w + (size_t)ix * 32 * KPL, lane, ev);
#pragma unroll
          for (int c = 0; c < KPL; c++) ac[c & 3] = fma(th[c], ev[c], ac[c & 3]);
        }
        pr[i] = (ac[0] + ac[1]) + (ac[2] + ac[3]);
      }
      if constexpr (i == N1 - 1)
        estep3_norm_exchange<L, NS, 0, N1, NS, 0, 0>(pr, rr, mycnt1, j, lane, badbits);
    });
  }
  {
    constexpr int t0 = Q::next_tm(0);
    if (t0 < nb) tmem_ldrw_words<2 * CA>(taddr(t0), wa);  // pass B's first load, behind the chains
  }
  if (nb > N1) {
    estep3_norm_exchange<L, NS, N1, N2, NS, 0, 0>(pr, rr, mycnt2, j, lane, badbits);
  } else {
#pragma unroll
    for (int i = N1; i < NS; i++) rr[i] = 0.0;
  }
  // ---- pass B: S_c += r_w E_wc ----
#pragma unroll
  for (int c = 0; c < KPL; c++) s[c] = 0.0;
  estep4_for<0, NS>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr int slot = FIRST + STEP * i, kd = S::kind(slot), ix = S::index(slot);
    if (i == 0 || i < nb) {
      const double r = rr[i];
      if constexpr (kd == 0) {
#pragma unroll
        for (int c = 0; c < KPL; c++) s[c] = fma(r, e[ix][c], s[c]);
      } else if constexpr (kd == 1) {
        tmem_wait_ld();
        tmem_ldrw_words<2 * CB>(taddr(i) + 2 * CA, wb);
#pragma unroll
        for (int c = 0; c < CA; c++) s[c] = fma(r, estep3_w2d(wa[2 * c], wa[2 * c + 1]), s[c]);
        tmem_wait_ld();
        constexpr int tn = Q::next_tm(i + 1);
        if (tn < nb) tmem_ldrw_words<2 * CA>(taddr(tn), wa);
#pragma unroll
        for (int c = 0; c < CB; c++) s[CA + c] = fma(r, estep3_w2d(wb[2 * c], wb[2 * c + 1]), s[CA + c]);
      } else {
        double ev[KPL];
        estep3_tile_ld<KPL>(w.tile_w + (size_t)ix * 32 * KPL, lane, ev);
#pragma unroll
        for (int c = 0; c < KPL; c++) s[c] = fma(r, ev[c], s[c]);
      }
    }
  });
  if (!chain_first) estep4_update<L, KPL, NW>(w, o, upd);
  if (__any_sync(0xffffffffu, badbits != 0))
    (void)estep3_slow_rows<L>(a.term_id + x.zrow0, a.count + x.zrow0, a.logbeta, nullptr, a.slow_rows, x.gam,
                              x.gextra, w.K, nb, badbits, lane, false);
  estep3_publish_s<L, KPL>(s, x.wtot + (size_t)w.warp * S::KS, lane);
}

// n_w of the (batch, row) whose norm this lane completes in the two reduce-scatters of a half:
// batches j and N1 + j of the document, row g (0 in the slots the warp does not use)
template <int L, int KPL, int NW>
__device__ __forceinline__ void estep4_mycnt(const Estep4Doc &x, const int g, const int j, double &c1, double &c2) {
  using S = Estep4Shape<L, KPL, NW>;
  constexpr int RPB = S::RPB, NS = S::NS, N1 = (NS + 1) / 2;
  c1 = j < N1 ? x.cnt_w[j * RPB + g] : 0.0;
  c2 = N1 + j < NS ? x.cnt_w[(N1 + j) * RPB + g] : 0.0;
}

// Sweeps 1..I-1 of a pair (or of A alone).  Half sweeps alternate: A, B, A, B, ...; every half
// sweep but the first carries the column owners' update of the other document, and B's last
// update is drained after the loop.  Without a partner the odd steps are A's update alone.
template <int L, int KPL, int NW>
__device__ __forceinline__ void estep4_run(const EstepArgs &a, const Estep4Warp &w, const Estep4Doc &A,
                                           const Estep4Doc &B, const bool hasB, const double (&e)[2][KPL]) {
  using S = Estep4Shape<L, KPL, NW>;
  double a1, a2, b1 = 0.0, b2 = 0.0;
  estep4_mycnt<L, KPL, NW>(A, w.g, w.j, a1, a2);
  if (hasB) estep4_mycnt<L, KPL, NW>(B, w.g, w.j, b1, b2);
  const int nsw = a.sweeps - 1;
  for (int it = 0; it < 2 * nsw; it++) {
    if ((it & 1) == 0)
      estep4_half<L, KPL, NW, 0, 1>(a, w, A, B, hasB && it > 0, e, a1, a2);
    else if (hasB)
      estep4_half<L, KPL, NW, S::NS - 1, -1>(a, w, B, A, true, e, b1, b2);
    else
      estep4_update<L, KPL, NW>(w, A, true);
    __syncthreads();
  }
  if (hasB && nsw > 0) {
    estep4_update<L, KPL, NW>(w, B, true);
    __syncthreads();
  }
}

// Rows of one document into this warp's slots (complete E rows, L lanes per row: compulsory
// traffic).  Padding rows are a copy of the document's last row with n_w = 0.
template <int L, int KPL, int NW, int FIRST, int STEP>
__device__ __forceinline__ void estep4_load_doc(const EstepArgs &a, const Estep4Warp &w, const Estep4Doc &x,
                                                const int bstart, double (&e)[2][KPL], double *tile_w) {
  using S = Estep4Shape<L, KPL, NW>;
  constexpr int RPB = S::RPB, TW = S::TW, NS = S::NS;
  for (int q = w.lane; q < NS * RPB; q += 32) x.cnt_w[q] = 0.0;
  __syncwarp();
  estep4_for<0, NS>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr int slot = FIRST + STEP * i, kd = S::kind(slot), ix = S::index(slot);
    if (i < x.nb) {  // warp-uniform (nb >= 1: a warp without rows walks one padding batch)
      const int row = (bstart + i) * RPB + w.g;
      const bool valid = row < x.n;
      const int rowc = row < x.n ? row : x.n - 1;
      double v[KPL];
      estep3_load_cols<L, KPL, true>(a.E + (size_t)(__ldg(a.term_id + x.b + rowc) - 1) * a.e_stride, w.j, v);
      if (valid && w.j == 0) x.cnt_w[i * RPB + w.g] = (double)__ldg(a.count + x.b + row);
      if constexpr (kd == 0) {
#pragma unroll
        for (int c = 0; c < KPL; c++) e[ix][c] = v[c];
      } else if constexpr (kd == 1) {
        uint32_t w32[TW];
#pragma unroll
        for (int c = 0; c < KPL; c++) {
          w32[2 * c] = (uint32_t)__double2loint(v[c]);
          w32[2 * c + 1] = (uint32_t)__double2hiint(v[c]);
        }
        tmem_st_words<TW>(w.tbase + (uint32_t)ix * TW, w32);
      } else {
        estep3_tile_st<KPL>(tile_w + (size_t)ix * 32 * KPL, w.lane, v);
      }
    }
  });
}

// The last sweep and the document epilogue of one document of the pair (generic over the batch
// count): likelihood terms (DocumentMapper.java:421), statistics scatter (:315-329), gamma output,
// ELBO counter and alpha statistics (:244-259,342-346).  All warps of the CTA take part.
template <int L, int KPL, int NW, int FIRST, int STEP>
__device__ __forceinline__ void estep4_finish_doc(const EstepArgs &a, const Estep4Warp &w, const Estep4Doc &x,
                                               const double (&e)[2][KPL], double *red_s, double *ssacc_s,
                                               const double lik_alpha) {
  using S = Estep4Shape<L, KPL, NW>;
  constexpr int TW = S::TW, KS = S::KS;
  const int K = w.K, lane = w.lane;
  double s[KPL];
  double lik = 0.0;
  unsigned badbits = 0;
#pragma unroll
  for (int c = 0; c < KPL; c++) s[c] = 0.0;
#pragma unroll 1
  for (int i = 0; i < x.nb; i++) {
    const int slot = FIRST + STEP * i, kd = S::kind(slot), ix = S::index(slot);
    double ev[KPL];
    if (kd == 0) {
#pragma unroll
      for (int c = 0; c < KPL; c++) ev[c] = ix == 0 ? e[0][c] : e[1][c];
    } else if (kd == 1) {
      uint32_t w32[TW];
      tmem_ld_words<TW>(w.tbase + (uint32_t)ix * TW, w32);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < KPL; c++) ev[c] = estep3_w2d(w32[2 * c], w32[2 * c + 1]);
    } else {
      estep3_tile_ld<KPL>(w.tile_w + (size_t)ix * 32 * KPL, lane, ev);
    }
    estep3_last_batch<L, KPL>(a, x.theta, x.cnt_w, ev, i, x.zrow0, w.g, w.j, s, lik, badbits);
  }
  if (__any_sync(0xffffffffu, badbits != 0))
    lik += estep3_slow_rows<L>(a.term_id + x.zrow0, a.count + x.zrow0, a.logbeta, a.learning ? a.stats : nullptr,
                               a.slow_rows, x.gam, x.gextra, K, x.nb, badbits, lane, true);
  estep3_publish_s<L, KPL>(s, x.wtot + (size_t)w.warp * KS, lane);
  lik = warp_sum(lik);
  if (lane == 0) red_s[w.warp * 4 + 3] = lik;
  __syncthreads();
  double sg = 0.0, lg = 0.0, lk = 0.0;
  for (int col = w.tid; col < K; col += S::THREADS) {
    const double sk = estep3_colsum<KS, NW>(x.wtot, 0, col);
    const double add = x.theta[col] * sk;
    const double g_old = x.gam[col];
    const double g_new = w.alpha_s[col] + add + x.gextra[col];
    x.gextra[col] = 0.0;
    if (add != 0.0) lk -= digamma_fast(g_old) * add;
    x.gam[col] = g_new;
    sg += g_new;
    lg += lgamma(g_new);
    a.gamma[(size_t)x.d * K + col] = g_new;
  }
  sg = warp_sum(sg);
  lg = warp_sum(lg);
  lk = warp_sum(lk);
  if (lane == 0) {
    red_s[w.warp * 4 + 0] = sg;
    red_s[w.warp * 4 + 1] = lg;
    red_s[w.warp * 4 + 2] = lk;
  }
  __syncthreads();
  double tg = 0.0, tl = 0.0, tk = 0.0;
  for (int w2 = 0; w2 < NW; w2++) {
    const double *rp = red_s + w2 * 4;
    tg += rp[0];
    tl += rp[1];
    tk += rp[2] + rp[3];
  }
  const double dsg = digamma_literal(tg);
  for (int col = w.tid; col < K; col += S::THREADS) ssacc_s[col] += digamma_literal(x.gam[col]) - dsg;
  if (w.tid == 0) {
    const double ll = lik_alpha + (tl - lgamma(tg)) + tk;
    const long long c = (long long)(-ll * 1e6);  // DocumentMapper.java:253-254
    atomicAdd(a.ll_counter, (unsigned long long)c);
  }
  __syncthreads();  // red_s and the document's shared state are reused
}

template <int L, int KPL, int NW>
__global__ void __launch_bounds__(NW * 32, 1) estep4_kernel(const EstepArgs a) {
  using S = Estep4Shape<L, KPL, NW>;
  constexpr int RPB = S::RPB, KS = S::KS, NS = S::NS, THREADS = S::THREADS;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int K = a.K;
  const Estep4Smem<L, KPL, NW> lay;

  Estep4Warp w;
  w.g = lane / L;
  w.j = lane % L;
  w.lane = lane;
  w.warp = warp;
  w.tid = tid;
  w.K = K;
  double *tile_w = smem + lay.tile + (size_t)warp * S::M * 32 * KPL;
  w.tile_w = tile_w;
  double *alpha_s = smem + lay.alpha;
  w.alpha_s = alpha_s;
  double *ssacc_s = smem + lay.ssacc;
  double *red_s = smem + lay.red;
  int *meta_s = reinterpret_cast<int *>(smem + lay.meta);  // [0] TMEM base, [1] work item
  Estep4Doc X[2];
#pragma unroll
  for (int p = 0; p < 2; p++) {
    X[p].theta = smem + lay.theta + (size_t)p * KS;
    X[p].gam = smem + lay.gam + (size_t)p * KS;
    X[p].gextra = smem + lay.gextra + (size_t)p * KS;
    X[p].wtot = smem + lay.wtot + (size_t)p * NW * KS;
    X[p].cnt_w = smem + lay.cnt + ((size_t)p * NW + warp) * (NS * RPB);
  }
  for (int k = tid; k < KS; k += THREADS) {
    alpha_s[k] = k < K ? a.alpha[k] : 1.0;
    ssacc_s[k] = 0.0;
  }
  for (int k = tid; k < 2 * KS; k += THREADS) {
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

  for (;;) {
    __syncthreads();  // everybody is done with the previous pair's shared state
    if (tid == 0) meta_s[1] = a.doc_begin + atomicAdd(a.work_counter, 1);
    __syncthreads();
    const int qi = meta_s[1];
    if (qi >= a.doc_end) break;
    const int dA = a.pairs[2 * qi], dB = a.pairs[2 * qi + 1];
    const bool hasB = dB >= 0;
    int bstart[2];
#pragma unroll
    for (int p = 0; p < 2; p++) {
      Estep4Doc &x = X[p];
      x.d = p == 0 ? dA : dB;
      x.nb = 0;
      x.n = 0;
      x.b = 0;
      x.zrow0 = 0;
      bstart[p] = 0;
      if (p == 1 && !hasB) continue;
      x.b = a.row_ptr[x.d];
      x.n = (int)(a.row_ptr[x.d + 1] - x.b);
      // row batches of the document, dealt to the warps in contiguous, balanced ranges: A's extra
      // batches go to the low warps, B's to the high warps
      const int nbt = (x.n + RPB - 1) / RPB;
      const int lo = nbt / NW, rem = nbt % NW;
      const int wv = p == 0 ? warp : NW - 1 - warp;  // position in the dealing order
      const int mine = lo + (wv < rem ? 1 : 0);
      bstart[p] = wv * lo + (wv < rem ? wv : rem);
      x.zrow0 = x.b + (int64_t)bstart[p] * RPB;
      x.nb = mine > 0 ? mine : 1;
    }
    double e[2][KPL];
#pragma unroll
    for (int c = 0; c < KPL; c++) e[1][c] = 0.0;
    estep4_load_doc<L, KPL, NW, 0, 1>(a, w, X[0], bstart[0], e, tile_w);
    if (hasB) estep4_load_doc<L, KPL, NW, NS - 1, -1>(a, w, X[1], bstart[1], e, tile_w);
    tmem_wait_st();
    // gamma start (DocumentMapper.java:184-193) and the first theta, by the column owners
#pragma unroll
    for (int p = 0; p < 2; p++) {
      if (p == 1 && !hasB) continue;
      const Estep4Doc &x = X[p];
      const double share = (double)((float)a.doc_tokens[x.d] / (float)K);
      for (int col = tid; col < K; col += THREADS) {
        const double gam = a.use_stored_gamma ? a.gamma[(size_t)x.d * K + col] : alpha_s[col] + share;
        x.gam[col] = gam;
        x.theta[col] = exp_digamma(gam);
      }
    }
    __syncthreads();

    // ---- sweeps 1..I-1: loop bodies specialised on this warp's batch counts ----
    estep4_run<L, KPL, NW>(a, w, X[0], X[1], hasB, e);

    // ---- last sweep and document epilogue, one document after the other ----
    estep4_finish_doc<L, KPL, NW, 0, 1>(a, w, X[0], e, red_s, ssacc_s, lik_alpha);
    if (hasB) estep4_finish_doc<L, KPL, NW, NS - 1, -1>(a, w, X[1], e, red_s, ssacc_s, lik_alpha);
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

struct Estep4Variant {
  int L, KPL, NW, NS;  // NS: batch slots per warp (a pair fits when its batches per warp do)
  cudaError_t (*launch)(const EstepArgs &, int grid, cudaStream_t);
  size_t smem_bytes;
  cudaError_t (*prepare)();
};

template <int L, int KPL, int NW>
static cudaError_t variant4_prepare() {
  return cudaFuncSetAttribute(estep4_kernel<L, KPL, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)Estep4Smem<L, KPL, NW>().total);
}
template <int L, int KPL, int NW>
static cudaError_t variant4_launch(const EstepArgs &args, int grid, cudaStream_t st) {
  estep4_kernel<L, KPL, NW><<<grid, NW * 32, Estep4Smem<L, KPL, NW>().total, st>>>(args);
  return cudaGetLastError();
}

#define MRLDA_VARIANT4(L_, KPL_, NW_)                                                               \
  {                                                                                                 \
    L_, KPL_, NW_, Estep4Shape<L_, KPL_, NW_>::NS, variant4_launch<L_, KPL_, NW_>,                  \
        Estep4Smem<L_, KPL_, NW_>().total, variant4_prepare<L_, KPL_, NW_>                          \
  }

extern const Estep4Variant g_estep4_variants[];
extern const int g_estep4_variants_n;

}  // namespace mrlda
