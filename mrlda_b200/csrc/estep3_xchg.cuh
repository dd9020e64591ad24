// estep3_xchg.cuh — the cross-CTA row-norm exchange of the cluster warp-group E-step kernel
// (estep3c_kernel.cuh): warp w of every CTA of a thread-block cluster sends its partial norms to warp
// w of every peer CTA with st.async, completing on the peer warp's mbarrier; the receiver sums the
// CL partials in rank order.  Included by estep3_kernel.cuh, whose sweep is written once for both
// the single-CTA kernel (CL = 1: none of this is instantiated) and the cluster kernel.
#pragma once
#include <stdint.h>

#include "device_math.cuh"
#include "estep_kernel.cuh"

namespace mrlda {

// doubles per (parity, rank) of a warp's exchange buffer: two halves of a sweep x 32 lanes, or the
// 40 (batch slot, row) norms of the last sweep
constexpr int ESTEP3_XSLOTS = 64;

__device__ __forceinline__ uint32_t estep3c_mapa(const uint32_t addr, const int rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// one double into a peer's shared memory, completing 8 bytes on the peer's mbarrier
__device__ __forceinline__ void estep3c_st_async(const uint32_t raddr, const double v, const uint32_t rmbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(raddr),
               "l"(__double_as_longlong(v)), "r"(rmbar)
               : "memory");
}
__device__ __forceinline__ void estep3c_st_remote_i32(const uint32_t raddr, const int v) {
  asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(raddr), "r"(v) : "memory");
}
__device__ __forceinline__ void estep3c_st_remote_f64(const uint32_t raddr, const double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(raddr), "d"(v) : "memory");
}
__device__ __forceinline__ void estep3c_arrive_remote(const uint32_t rmbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rmbar) : "memory");
}
__device__ __forceinline__ void estep3c_mbar_wait(const uint32_t mbar, const uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, "
        "p;\n}"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void estep3c_mbar_expect(const uint32_t mbar, const uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// What a warp needs for the norm exchange (built once per kernel).
struct Estep3cX {
  const double *xbuf_w;  // this warp's exchange buffer [parity][rank][XSLOTS]
  uint32_t xbuf_u32;     // its shared-memory address
  uint32_t xbar_u32;     // shared-memory address of this warp's two exchange mbarriers
  int crank;
};

// this lane's value to slot `slot` of the exchange buffer of parity `par`: own copy and every peer's
template <int CL, int XS = ESTEP3_XSLOTS>
__device__ __forceinline__ void estep3c_send(const Estep3cX &x, const int par, const int slot, const double v) {
  const uint32_t la = x.xbuf_u32 + (uint32_t)(((par * CL + x.crank) * XS + slot) * sizeof(double));
  const uint32_t lm = x.xbar_u32 + (uint32_t)(par * sizeof(unsigned long long));
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(la), "d"(v) : "memory");
#pragma unroll
  for (int d = 1; d < CL; d++) {
    const int r = x.crank + d < CL ? x.crank + d : x.crank + d - CL;
    estep3c_st_async(estep3c_mapa(la, r), v, estep3c_mapa(lm, r));
  }
}
// sum over the ranks, in rank order (every CTA gets the same bits)
template <int CL, int XS = ESTEP3_XSLOTS>
__device__ __forceinline__ double estep3c_total(const Estep3cX &x, const int par, const int slot) {
  const double *p = x.xbuf_w + (size_t)par * CL * XS + slot;
  double v[CL];
#pragma unroll
  for (int r = 0; r < CL; r++) v[r] = p[(size_t)r * XS];
  double t = v[0];
#pragma unroll
  for (int r = 1; r < CL; r++) t += v[r];
  return t;
}

// L-lane reduce-scatter of the partial dot products of the batches [START, START + NSV): lane j ends
// up with this CTA's partial norm of batch START + j, row g (estep3_norm_exchange, first part)
template <int L, int NA, int START, int NSV>
__device__ __forceinline__ double estep3c_norm_rs(const double (&pr)[NA], const int j) {
  static_assert(NSV >= 1 && NSV <= L && START + NSV <= NA, "one batch per lane of a row");
  double q[L];
#pragma unroll
  for (int i = 0; i < L; i++) q[i] = i < NSV ? pr[START + (i < NSV ? i : 0)] : 1.0;
#pragma unroll
  for (int o = 1; o < L; o <<= 1) {
    const bool up = (j & o) != 0;
#pragma unroll
    for (int m = 0; m < L / (2 * o); m++) {
      if (2 * m * o < NSV) {
        const double send = up ? q[2 * m] : q[2 * m + 1];
        const double keep = up ? q[2 * m + 1] : q[2 * m];
        q[m] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
  }
  return q[0];
}
// r = n_w / norm_w from the complete norm this lane holds, back to the lanes of the row
// (estep3_norm_exchange, second part)
template <int L, int NA, int START, int NSV, int RT, int NTM, int T>
__device__ __forceinline__ void estep3c_norm_finish(const double nrm, double (&rr)[NA], const double mycnt,
                                                    const int lane, unsigned &badbits) {
  const bool live = mycnt > 0.0;
  const bool ok = nrm >= MRLDA_NORM_TINY;
  const double r = (live && ok) ? mycnt * rcp_fast(nrm) : 0.0;
  const int base = lane & ~(L - 1);
#pragma unroll
  for (int i = 0; i < NSV; i++) rr[START + i] = __shfl_sync(0xffffffffu, r, base + i);
  const unsigned bm = __ballot_sync(0xffffffffu, live && !ok);
  if (bm) {
    const unsigned mine = bm >> base;
#pragma unroll
    for (int i = 0; i < NSV; i++)
      if ((mine >> i) & 1u) badbits |= 1u << (START + i < RT + NTM ? START + i : START + i - NTM + T);
  }
}

}  // namespace mrlda
