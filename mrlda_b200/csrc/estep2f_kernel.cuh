// estep2f_kernel.cuh — optional fp32 mode of the register/TMEM-stationary E-step (K1a).
//
// Same mapping as estep2_kernel.cuh (topics over warps, rows over lanes, thread-private slots in
// registers then Tensor Memory), with the tile E, theta, the partial dot products and the S
// accumulators in single precision: half the storage per slot (twice the rows in registers and
// TMEM, no shared-memory slots needed), FFMA instead of DFMA, one shuffle per exchanged value.
// Everything that is O(K) or O(rows) per sweep stays in double: gamma, digamma/exp, the row norms
// (sum of the NW float partials) and reciprocals, the likelihood, the alpha statistics, and the
// K x V statistics (float contributions are added into the fp64 matrix, so the finalize and alpha
// kernels are shared with the fp64 mode).  BASELINE.json's tolerance for this mode is 1e-4 relative
// on the per-iteration ELBO and logbeta (tests/test_gpu_parity.py::test_fp32_mode).
// Rows whose norm falls below 1e-30 take the exact fp64 log-space path (float range).
// Reference path: DocumentMapper.java:175-350, updatePhi :402-429.
#pragma once
#include "estep2_kernel.cuh"

namespace mrlda {

#define MRLDA_NORM_TINY_F 1e-30

__host__ __device__ constexpr int estep2f_tmem_stride(int cpw) {  // one 32-bit word per value
  return cpw > 16 ? 32 : (cpw > 8 ? 16 : 8);
}
__host__ __device__ constexpr int estep2f_tmem_slots(int nw, int cpw) {
  return estep2_tmem_cols(nw, 8) / (nw > 4 ? nw / 4 : 1) / estep2f_tmem_stride(cpw);
}
__host__ __device__ constexpr int estep2f_rreg(int cpw) {
  int r = (255 - 60 - 2 * cpw - estep2f_tmem_stride(cpw)) / cpw;
  return r < 1 ? 1 : (r > 8 ? 8 : r);
}
template <int NW, int CPW>
__host__ __device__ constexpr size_t estep2f_smem_bytes(int cap_rows) {
  return sizeof(float) * (2 * (size_t)NW * cap_rows  // partial dots, two buffers
                          + 2 * (size_t)cap_rows     // counts, r
                          + (size_t)NW * CPW)        // theta
         + sizeof(double) * (2 * (size_t)NW * CPW + 96);  // gamma, gextra, reductions + meta
}

template <int N, int OFF, int CPW>
__device__ __forceinline__ void reduce_scatter_f(float (&v)[CPW], int lane, int &base, int &cnt) {
  if constexpr (OFF >= 1) {
    constexpr int H = (N + 1) / 2;
    const bool upper = (lane & OFF) != 0;
#pragma unroll
    for (int i = 0; i < H; i++) {
      float lo = v[i];
      float hi = (i + H < N) ? v[i + H] : 0.0f;
      float send = upper ? lo : hi;
      float keep = upper ? hi : lo;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
    if (upper) {
      base += H;
      cnt = cnt > H ? cnt - H : 0;
    } else {
      cnt = cnt < H ? cnt : H;
    }
    if constexpr (OFF / 2 >= 1) reduce_scatter_f<H, OFF / 2, CPW>(v, lane, base, cnt);
  }
}

// fp32 form of estep2_scatter_stats (estep2_kernel.cuh): n_w phi_wk = r_w theta_k E_wk of the last
// sweep, evaluated in float exactly as the tile would (float r, float theta, E rounded to float), one
// warp per row with lanes across the topics so that the E re-read and the atomics are contiguous.
static __device__ __noinline__ void estep2f_scatter_stats(const int32_t *term_id, const double *E,
                                                          const int e_stride, double *stats, const int K,
                                                          const float *r_s, const float *theta_s,
                                                          const int n, const int KS, const int warp,
                                                          const int nw, const int lane) {
  for (int row = warp; row < n; row += 2 * nw) {
    const int row2 = row + nw;
    const float r0 = r_s[row], r1 = row2 < n ? r_s[row2] : 0.0f;
    const bool ok0 = r0 > 0.0f, ok1 = r1 > 0.0f;
    const size_t t0 = ok0 ? (size_t)(__ldg(term_id + row) - 1) : 0;
    const size_t t1 = ok1 ? (size_t)(__ldg(term_id + row2) - 1) : 0;
    const double *e0 = E + t0 * e_stride, *e1 = E + t1 * e_stride;
    float v0[8], v1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int k = lane + 32 * i;
      const bool in = k < KS && k < K;
      v0[i] = ok0 && in ? (float)__ldg(e0 + k) : 0.0f;
      v1[i] = ok1 && in ? (float)__ldg(e1 + k) : 0.0f;
    }
    double *s0 = stats + t0 * K, *s1 = stats + t1 * K;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int k = lane + 32 * i;
      if (k < KS && k < K) {
        const float th = theta_s[k];
        if (ok0) atomicAdd(s0 + k, (double)(r0 * th * v0[i]));
        if (ok1) atomicAdd(s1 + k, (double)(r1 * th * v1[i]));
      }
    }
  }
}

template <int NW, int CPW, int RREG>
__global__ void __launch_bounds__(NW * 32, 8 / NW) estep2f_kernel(const EstepArgs a) {
  constexpr int KS = NW * CPW, TS = estep2f_tmem_stride(CPW);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int K = a.K;
  const int rt = a.tmem_slots;
  const int cap_rows = a.cap_rows;  // 32 * (RREG + rt)

  double *gam_s = reinterpret_cast<double *>(smem_raw);       // KS
  double *gextra_s = gam_s + KS;                              // KS
  double *red_s = gextra_s + KS;                              // 96
  int *meta_s = reinterpret_cast<int *>(red_s + 64);
  float *part_s = reinterpret_cast<float *>(red_s + 96);      // 2 x NW x cap_rows
  float *cnt_s = part_s + 2 * (size_t)NW * cap_rows;          // cap_rows
  float *r_s = cnt_s + cap_rows;                              // cap_rows
  float *theta_s = r_s + cap_rows;                            // KS

  uint32_t tbase = 0;
  if (rt > 0) {
    if (warp == 0) tmem_alloc(reinterpret_cast<uint32_t *>(meta_s + 2), estep2_tmem_cols(NW, 8));
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t cols_per_thread = estep2_tmem_cols(NW, 8) / (NW > 4 ? NW / 4 : 1);
    tbase = (uint32_t)meta_s[2] + (((uint32_t)(warp & 3) * 32u) << 16) +
            (uint32_t)(warp >> 2) * cols_per_thread;
  }

  const int mycol = lane < CPW ? lane : -1;
  int rs_src = 0;
  {
    float dummy[CPW];
#pragma unroll
    for (int c = 0; c < CPW; c++) dummy[c] = 0.0f;
    int base = 0, cnt = CPW;
    reduce_scatter_f<CPW, 16, CPW>(dummy, lane, base, cnt);
    const int rs_col = cnt > 0 ? base : -1;
    for (int l2 = 0; l2 < 32; l2++) {
      int c2 = __shfl_sync(0xffffffffu, rs_col, l2);
      if (c2 == lane) rs_src = l2;
    }
  }
  const int lcol = warp * CPW + mycol;
  const int col = lcol;
  const bool owner = mycol >= 0 && col < K;
  const double alpha_c = owner ? a.alpha[col] : 1.0;
  double ss_acc = 0.0;
  const double lik_alpha = *a.lik_alpha;
  for (int k = tid; k < KS; k += NW * 32) {
    theta_s[k] = 0.0f;
    gextra_s[k] = 0.0;
    gam_s[k] = 1.0;
  }
  if (tid == 0) meta_s[1] = 0;
  float *theta_w = theta_s + warp * CPW;
  const int ldmode = (a.e_stride & 3) ? 0 : (((warp * CPW) & 3) ? 2 : 1);  // see estep2_load_slice

  for (;;) {
    __syncthreads();
    if (tid == 0) meta_s[0] = a.doc_begin + atomicAdd(a.work_counter, 1);
    __syncthreads();
    const int qi = meta_s[0];
    if (qi >= a.doc_end) break;
    const int d = a.doc_order[qi];
    const int64_t b = a.row_ptr[d];
    const int n = (int)(a.row_ptr[d + 1] - b);
    if (n <= 0) continue;
    const int nslots = (n + 31) >> 5;
    const int nrows = 32 * (nslots > RREG ? nslots : RREG);

    // ---- load my elements (converted to float once per document) ----
    float e[RREG][CPW];
#pragma unroll
    for (int q = 0; q < RREG; q++) {
      const int row = 32 * q + lane;
      if (row < n) {
        double v[CPW];
        estep2_load_slice<CPW>(a.E + (size_t)(__ldg(a.term_id + b + row) - 1) * a.e_stride + warp * CPW,
                               ldmode, v);
#pragma unroll
        for (int c = 0; c < CPW; c++) e[q][c] = (float)v[c];
      } else {
#pragma unroll
        for (int c = 0; c < CPW; c++) e[q][c] = 0.0f;
      }
    }
    for (int q = RREG; q < nslots; q++) {  // TMEM slots
      const int row = 32 * q + lane;
      uint32_t w32[TS];
#pragma unroll
      for (int i = 0; i < TS; i++) w32[i] = 0u;
      if (row < n) {
        double v[CPW];
        estep2_load_slice<CPW>(a.E + (size_t)(__ldg(a.term_id + b + row) - 1) * a.e_stride + warp * CPW,
                               ldmode, v);
#pragma unroll
        for (int c = 0; c < CPW; c++) w32[c] = __float_as_uint((float)v[c]);
      }
      TmemVec<TS>::st(tbase + (uint32_t)(q - RREG) * TS, w32);
    }
    if (nslots > RREG) tmem_wait_st();
    for (int r = tid; r < nrows; r += NW * 32) cnt_s[r] = r < n ? (float)a.count[b + r] : 0.0f;
    double gam = 1.0;
    if (owner) {
      if (a.use_stored_gamma)
        gam = a.gamma[(size_t)d * K + col];
      else
        gam = alpha_c + (double)((float)a.doc_tokens[d] / (float)K);
    }
    __syncthreads();

    double lik = 0.0, theta_c = 0.0;
    for (int sweep = 0; sweep < a.sweeps; sweep++) {
      const bool last = (sweep == a.sweeps - 1);
      // (1) theta in double by the owner lane, broadcast as float
      theta_c = owner ? exp_digamma(gam) : 0.0;
      if (mycol >= 0) theta_w[mycol] = (float)theta_c;
      __syncwarp();
      float th[CPW];
#pragma unroll
      for (int c = 0; c < CPW; c += 2) {
        const float2 tv = *reinterpret_cast<const float2 *>(theta_w + c);
        th[c] = tv.x;
        th[c + 1] = tv.y;
      }
      // (2) partial dot products
      float *part = part_s + (size_t)(sweep & 1) * NW * cap_rows;
      float *pw = part + (size_t)warp * cap_rows + lane;
      {
        float acc[RREG][2];
#pragma unroll
        for (int q = 0; q < RREG; q++) acc[q][0] = acc[q][1] = 0.0f;
#pragma unroll
        for (int c = 0; c < CPW; c += 2) {
#pragma unroll
          for (int q = 0; q < RREG; q++) {
            acc[q][0] = fmaf(th[c], e[q][c], acc[q][0]);
            acc[q][1] = fmaf(th[c + 1], e[q][c + 1], acc[q][1]);
          }
        }
#pragma unroll
        for (int q = 0; q < RREG; q++) pw[32 * q] = acc[q][0] + acc[q][1];
      }
      for (int q = RREG; q < nslots; q++) {
        uint32_t w32[TS];
        TmemVec<TS>::ld(tbase + (uint32_t)(q - RREG) * TS, w32);
        tmem_wait_ld();
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
        for (int c = 0; c < CPW; c += 2) {
          a0 = fmaf(th[c], __uint_as_float(w32[c]), a0);
          a1 = fmaf(th[c + 1], __uint_as_float(w32[c + 1]), a1);
        }
        pw[32 * q] = a0 + a1;
      }
      __syncthreads();  // B1

      // (3a) one thread per row: norm in double, r = n_w / norm as float
      for (int row = warp * 32 + lane; row < nrows; row += NW * 32) {
        const float *pp = part + row;
        double n0 = 0.0, n1 = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < NW; w2++) {
          if (w2 & 1)
            n1 += (double)pp[(size_t)w2 * cap_rows];
          else
            n0 += (double)pp[(size_t)w2 * cap_rows];
        }
        const double norm = n0 + n1;
        const bool valid = row < n;
        const bool bad = valid && !(norm >= MRLDA_NORM_TINY_F);
        const bool ok = valid && !bad;
        const double cn = (double)cnt_s[row];
        r_s[row] = bad ? -1.0f : (float)((ok ? cn : 0.0) * rcp_fast(ok ? norm : 1.0));
        if (bad) meta_s[1] = 1;
        if (last && ok)
          lik = fma(cn, log(norm) + __ldg(a.rowmax + __ldg(a.term_id + b + row) - 1), lik);
      }
      __syncthreads();  // B2

      // (3b) S_c += r_w E_wc
      float s[CPW];
#pragma unroll
      for (int c = 0; c < CPW; c++) s[c] = 0.0f;
      unsigned badbits = 0;
      {
        float r[RREG];
#pragma unroll
        for (int q = 0; q < RREG; q++) {
          r[q] = r_s[32 * q + lane];
          if (r[q] < 0.0f) {
            badbits |= 1u << q;
            r[q] = 0.0f;
          }
        }
#pragma unroll
        for (int c = 0; c < CPW; c++) {
#pragma unroll
          for (int q = 0; q < RREG; q++) s[c] = fmaf(r[q], e[q][c], s[c]);
        }
      }
      for (int q = RREG; q < nslots; q++) {
        float r = r_s[32 * q + lane];
        if (r < 0.0f) {
          badbits |= 1u << q;
          r = 0.0f;
        }
        uint32_t w32[TS];
        TmemVec<TS>::ld(tbase + (uint32_t)(q - RREG) * TS, w32);
        tmem_wait_ld();
        float tmp[CPW];
#pragma unroll
        for (int c = 0; c < CPW; c++) {
          tmp[c] = __uint_as_float(w32[c]);
          s[c] = fmaf(r, tmp[c], s[c]);
        }
      }
      // (4) S_c to the owner lane
      float S_f;
      {
        int base = 0, cnt = CPW;
        reduce_scatter_f<CPW, 16, CPW>(s, lane, base, cnt);
        S_f = __shfl_sync(0xffffffffu, s[0], rs_src);
      }
      const double S_c = (double)S_f;

      const bool anybad = meta_s[1] != 0;
      if (anybad) {  // exact fp64 log-space redo of the rows whose norm left the float range
        if (mycol >= 0) gam_s[lcol] = owner ? gam : 1.0;
        __syncthreads();
        if (tid == 0) meta_s[1] = 0;
        if (warp == 0) {
          for (int q = 0; q < nslots; q++) {
            unsigned m = __ballot_sync(0xffffffffu, (badbits >> q) & 1u);
            while (m) {
              const int l2 = __ffs(m) - 1;
              m &= m - 1;
              const int64_t z = b + 32 * q + l2;
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
                  gextra_s[k] += c;
                  if (last) {
                    lik += c * (lb[k] - lphi);
                    if (a.learning) atomicAdd(a.stats + (size_t)(term - 1) * K + k, c);
                  }
                }
              }
              if (lane == 0) atomicAdd(a.slow_rows, 1);
            }
          }
        }
        __syncthreads();
      }
      if (!last) {
        if (owner) {
          double gnew = fma(theta_c, S_c, alpha_c);
          if (anybad) {
            gnew += gextra_s[lcol];
            gextra_s[lcol] = 0.0;
          }
          gam = gnew;
        }
      } else {
        double g_new = 1.0, lg = 0.0;
        if (owner) {
          const double add = theta_c * S_c;
          g_new = alpha_c + add;
          if (anybad) {
            g_new += gextra_s[lcol];
            gextra_s[lcol] = 0.0;
          }
          if (add != 0.0) lik -= digamma_fast(gam) * add;
          lg = lgamma(g_new);
          a.gamma[(size_t)d * K + col] = g_new;
        }
        double sg = warp_sum(owner ? g_new : 0.0);
        lg = warp_sum(lg);
        lik = warp_sum(lik);
        if (lane == 0) {
          red_s[warp] = sg;
          red_s[16 + warp] = lg;
          red_s[32 + warp] = lik;
        }
        __syncthreads();
        double tg = 0.0, tl = 0.0, tk = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < NW; w2++) {
          tg += red_s[w2];
          tl += red_s[16 + w2];
          tk += red_s[32 + w2];
        }
        __syncthreads();
        if (owner) ss_acc += digamma_literal(g_new) - digamma_literal(tg);
        if (tid == 0) {
          double ll = lik_alpha + (tl - lgamma(tg)) + tk;
          long long c = (long long)(-ll * 1e6);
          atomicAdd(a.ll_counter, (unsigned long long)c);
        }
      }
    }
    // statistics scatter of the last sweep, one warp per row (see estep2_scatter_stats); r and theta
    // stay in shared memory until the barrier at the top of the document loop
    if (a.learning)
      estep2f_scatter_stats(a.term_id + b, a.E, a.e_stride, a.stats, K, r_s, theta_s, n, KS, warp, NW, lane);
  }
  if (owner && ss_acc != 0.0) atomicAdd(a.alpha_ss + col, ss_acc);
  if (rt > 0) {
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc((uint32_t)meta_s[2], estep2_tmem_cols(NW, 8));
  }
}

struct Estep2fVariant {
  int NW, CPW, RREG;
  cudaError_t (*launch)(const EstepArgs &, int grid, int cap_rows, size_t smem, cudaStream_t);
  size_t (*smem_bytes)(int cap_rows);
  cudaError_t (*prepare)(size_t smem);
};
template <int NW, int CPW>
static size_t variant2f_smem(int cap_rows) {
  return estep2f_smem_bytes<NW, CPW>(cap_rows);
}
template <int NW, int CPW>
static cudaError_t variant2f_prepare(size_t smem) {
  return cudaFuncSetAttribute(estep2f_kernel<NW, CPW, estep2f_rreg(CPW)>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <int NW, int CPW>
static cudaError_t variant2f_launch(const EstepArgs &args, int grid, int cap_rows, size_t smem,
                                    cudaStream_t st) {
  EstepArgs a = args;
  a.cap_rows = cap_rows;
  estep2f_kernel<NW, CPW, estep2f_rreg(CPW)><<<grid, NW * 32, smem, st>>>(a);
  return cudaGetLastError();
}
#define MRLDA_VARIANT2F(NW_, CPW_)                                                           \
  {                                                                                          \
    NW_, CPW_, estep2f_rreg(CPW_), variant2f_launch<NW_, CPW_>, variant2f_smem<NW_, CPW_>,   \
        variant2f_prepare<NW_, CPW_>                                                         \
  }
extern const Estep2fVariant g_estep2f_variants_W1[];
extern const int g_estep2f_variants_W1_n;
extern const Estep2fVariant g_estep2f_variants_W2[];
extern const int g_estep2f_variants_W2_n;
extern const Estep2fVariant g_estep2f_variants_W4[];
extern const int g_estep2f_variants_W4_n;
extern const Estep2fVariant g_estep2f_variants_W8[];
extern const int g_estep2f_variants_W8_n;

}  // namespace mrlda
