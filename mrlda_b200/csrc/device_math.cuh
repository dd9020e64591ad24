// device_math.cuh — special functions used by the Mr.LDA kernels (fp64).
//
// digamma / trigamma follow the lda-c style series that cloud9's Gamma class uses; the reference's
// known-answer tests pin them (src/test/java/cc/mrlda/VariationalInferenceTest.java:27-62).  Call
// sites replaced: DocumentMapper.java:209,256,258; TermReducer.java:173,195,213,233;
// VariationalInference.java:434,438,449,597-603.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace mrlda {

// Literal operation order (no FMA contraction) — used where the (x+6)-6 cancellation is visible in
// the result: the beta finalize evaluates digamma(1e-12 + s) ~ -1e12 and must round like Java does.
__device__ __forceinline__ double digamma_literal(double x) {
  double x6 = __dadd_rn(x, 6.0);
  double p = __ddiv_rn(1.0, __dmul_rn(x6, x6));
  p = __dmul_rn(
      __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(0.004166666666667, p),
                                                         -0.003968253986254), p),
                                    0.008333333333333), p),
                -0.083333333333333), p);
  double r = __dadd_rn(p, log(x6));
  r = __dadd_rn(r, -__ddiv_rn(0.5, x6));
  r = __dadd_rn(r, -__ddiv_rn(1.0, __dadd_rn(x6, -1.0)));
  r = __dadd_rn(r, -__ddiv_rn(1.0, __dadd_rn(x6, -2.0)));
  r = __dadd_rn(r, -__ddiv_rn(1.0, __dadd_rn(x6, -3.0)));
  r = __dadd_rn(r, -__ddiv_rn(1.0, __dadd_rn(x6, -4.0)));
  r = __dadd_rn(r, -__ddiv_rn(1.0, __dadd_rn(x6, -5.0)));
  r = __dadd_rn(r, -__ddiv_rn(1.0, __dadd_rn(x6, -6.0)));
  return r;
}

__device__ __forceinline__ double trigamma_literal(double x) {
  double x6 = __dadd_rn(x, 6.0);
  double p = __ddiv_rn(1.0, __dmul_rn(x6, x6));
  double q = __dadd_rn(__dmul_rn(0.075757575757576, p), -0.033333333333333);
  q = __dadd_rn(__dmul_rn(q, p), 0.0238095238095238);
  q = __dadd_rn(__dmul_rn(q, p), -0.033333333333333);
  q = __dadd_rn(__dmul_rn(q, p), 0.166666666666667);
  q = __dadd_rn(__dmul_rn(q, p), 1.0);
  p = __dadd_rn(__ddiv_rn(q, x6), __dmul_rn(0.5, p));
#pragma unroll
  for (int i = 0; i < 6; i++) {
    x6 = __dadd_rn(x6, -1.0);
    p = __dadd_rn(__ddiv_rn(1.0, __dmul_rn(x6, x6)), p);
  }
  return p;
}

// 1/x to ~1 ulp for normal positive x: MUFU.RCP64H seed + two Newton steps.
__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}

// exp(u) for u <= ~1 with a short dependency chain: Cody-Waite reduction, degree-12 Taylor
// polynomial in Estrin form (|r| <= ln2/2: truncation 2e-16), exponent built by integer add.
// Returns 0 for u < -708 (the result would be subnormal; callers treat such weights as 0).
__device__ __forceinline__ double exp_short(double u) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: rint() in the low mantissa bits
  double t = fma(u, 1.4426950408889634074, magic);
  const int n = __double2loint(t);
  const double nf = t - magic;
  double r = fma(nf, -6.93147180369123816490e-01, u);
  r = fma(nf, -1.90821492927058770002e-10, r);
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double a0 = 1.0 + r;
  const double a1 = fma(r, 1.0 / 6.0, 0.5);
  const double a2 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  const double a3 = fma(r, 1.0 / 5040.0, 1.0 / 720.0);
  const double a4 = fma(r, 1.0 / 362880.0, 1.0 / 40320.0);
  const double a5 = fma(r, 1.0 / 39916800.0, 1.0 / 3628800.0);
  const double b0 = fma(a1, r2, a0);
  const double b1 = fma(a3, r2, a2);
  const double b2 = fma(a5, r2, a4);
  const double d0 = fma(b1, r4, b0);
  const double d1 = fma(1.0 / 479001600.0, r4, b2);
  const double p = fma(d1, r8, d0);
  const double scale = __hiloint2double((n + 1023) << 20, 0);
  return u < -708.0 ? 0.0 : p * scale;
}

// The same digamma series restructured for the E-step's inner fixed point
// (DocumentMapper.java:208-211 evaluates it K times per sweep, 99 sweeps per document):
//   psi(x) = P(1/x6^2) + log(x6) - 0.5/x6 - sum_{j=0..5} 1/(t+j),  x6 = x+6, t = x6-6
// with the six reciprocals folded into one rational Q'(t)/Q(t), Q(t) = t(t+1)...(t+5), and a single
// reciprocal 1/(x6*Q).  With y = t^2 + 5t:  Q = y (y+4) (y+6),  Q' = (2t+5)(3y^2 + 20y + 24), which
// keeps the dependency chain short.  t keeps the reference's (x+6)-6 rounding.  Returns
// u = psi(x) - log(x6) and x6, so that exp(psi(x)) = x6 * exp(u) needs no logarithm.  The difference
// to the literal form is a few ulp of the largest term (no cancellation: every factor is positive).
__device__ __forceinline__ double digamma_minus_log(double x, double &x6_out) {
  const double x6 = x + 6.0;
  const double t = x6 - 6.0;
  const double y = fma(t, t, 5.0 * t);
  const double q = y * ((y + 4.0) * (y + 6.0));
  const double dq = fma(2.0, t, 5.0) * fma(fma(3.0, y, 20.0), y, 24.0);
  const double inv = rcp_fast(x6 * q);  // 1/(x6*Q)
  const double rx6 = q * inv;           // 1/x6
  const double ratio = (dq * x6) * inv; // Q'/Q
  double p = rx6 * rx6;
  p = fma(fma(fma(0.004166666666667, p, -0.003968253986254), p, 0.008333333333333), p,
          -0.083333333333333) * p;
  x6_out = x6;
  return fma(-0.5, rx6, p - ratio);
}

// psi(x) by the restructured series.
__device__ __forceinline__ double digamma_fast(double x) {
  double x6;
  double u = digamma_minus_log(x, x6);
  return u + log(x6);
}

// exp(psi(x)): the weight theta_k of the linear-space E-step.  Underflows to 0 for x << 1.
__device__ __forceinline__ double exp_digamma(double x) {
  double x6;
  double u = digamma_minus_log(x, x6);
  return x6 * exp_short(u);
}

// 1/x to < 1 ulp from the MUFU.RCP64H seed with one cubic step, y (1 + e + e^2), e = 1 - x y: three
// dependent operations instead of the four of two Newton steps (seed error ~2^-22 -> 2^-66).
__device__ __forceinline__ double rcp_cubic(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  return fma(y, fma(e, e, e), y);
}

// N independent evaluations of exp(psi(x)) written level by level, so that the chains are issued
// interleaved, and with the shortest dependency chain the series allows: the inner fixed point
// waits for these values with nothing else to run (one document's warps all stand at the same
// barrier).  Same series and the same (x+6)-6 rounding as digamma_minus_log; 1/x6 and 1/Q take
// separate reciprocals (two short chains instead of one long one) and exp(u) x6 is assembled as
// p(r) * (2^n x6).
template <int N>
__device__ __forceinline__ void exp_digamma_n(const double (&x)[N], double (&out)[N]) {
  double x6[N], t[N], rx6[N], y[N], q[N], dq[N], invq[N], a[N], u[N];
#pragma unroll
  for (int i = 0; i < N; i++) x6[i] = x[i] + 6.0;
#pragma unroll
  for (int i = 0; i < N; i++) t[i] = x6[i] - 6.0;
#pragma unroll
  for (int i = 0; i < N; i++) rx6[i] = rcp_cubic(x6[i]);
#pragma unroll
  for (int i = 0; i < N; i++) y[i] = t[i] * (t[i] + 5.0);
#pragma unroll
  for (int i = 0; i < N; i++) q[i] = (y[i] * (y[i] + 4.0)) * (y[i] + 6.0);
#pragma unroll
  for (int i = 0; i < N; i++) invq[i] = rcp_cubic(q[i]);
#pragma unroll
  for (int i = 0; i < N; i++) dq[i] = fma(2.0, t[i], 5.0) * fma(fma(3.0, y[i], 20.0), y[i], 24.0);
#pragma unroll
  for (int i = 0; i < N; i++) {
    const double p = rx6[i] * rx6[i], p2 = p * p;
    const double lo = fma(0.008333333333333, p, -0.083333333333333);
    const double hi = fma(0.004166666666667, p, -0.003968253986254);
    a[i] = fma(fma(hi, p2, lo), p, -0.5 * rx6[i]);
  }
#pragma unroll
  for (int i = 0; i < N; i++) u[i] = fma(-dq[i], invq[i], a[i]);
  const double magic = 6755399441055744.0;  // 1.5 * 2^52
#pragma unroll
  for (int i = 0; i < N; i++) {
    const double tt = fma(u[i], 1.4426950408889634074, magic);
    const int n = __double2loint(tt);
    const double nf = tt - magic;
    double r = fma(nf, -6.93147180369123816490e-01, u[i]);
    r = fma(nf, -1.90821492927058770002e-10, r);
    const double sx = __hiloint2double((n + 1023) << 20, 0) * x6[i];  // 2^n x6
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = 1.0 + r;
    const double a1 = fma(r, 1.0 / 6.0, 0.5);
    const double a2 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    const double a3 = fma(r, 1.0 / 5040.0, 1.0 / 720.0);
    const double a4 = fma(r, 1.0 / 362880.0, 1.0 / 40320.0);
    const double a5 = fma(r, 1.0 / 39916800.0, 1.0 / 3628800.0);
    const double b0 = fma(a1, r2, a0);
    const double b1 = fma(a3, r2, a2);
    const double b2 = fma(a5, r2, a4);
    const double d0 = fma(b1, r4, b0);
    const double d1 = fma(1.0 / 479001600.0, r4, b2);
    const double pz = fma(d1, r8, d0);
    out[i] = u[i] < -708.0 ? 0.0 : pz * sx;
  }
}

// LogMath.add without a drop threshold (see oracle/mrlda_oracle.c header for the pinning status).
__device__ __forceinline__ double logadd(double a, double b) {
  double hi = fmax(a, b), lo = fmin(a, b);
  if (hi == -INFINITY) return hi;
  return hi + log1p(exp(lo - hi));
}

}  // namespace mrlda
