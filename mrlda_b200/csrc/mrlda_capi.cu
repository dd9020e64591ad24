// mrlda_capi.cu — context, model kernels (K2 beta finalize, K3 alpha update) and the C ABI of
// libmrlda_b200.so declared in include/mrlda_b200.h.
//
// Reference code replaced here:
//   TermCombiner.java:19-35, TermReducer.java:134-238   -> finalize_* kernels (K2)
//   DocumentMapper.importBeta (DocumentMapper.java:475-536) -> import_beta_kernel
//   DocumentMapper.retrieveBeta (:446-463)               -> init_beta_kernel (seeded, not Math.random)
//   VariationalInference.updateVectorAlpha / updateScalarAlpha (VariationalInference.java:409-511,
//   573-625, wrapper :295-311)                           -> alpha_*_kernel (K3)
//   Hadoop shuffle + Counters (VariationalInference.java:242-243,260-268) -> one NCCL all-reduce
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mrlda_b200.h"
#include "estep2f_kernel.cuh"
#include <cuda.h>  // CUtensorMap types only; the encoder is looked up through the runtime (no -lcuda)

#include "estep3_kernel.cuh"
#include "estep3c_kernel.cuh"

using namespace mrlda;


// ------------------------------------------------------------------------------------------------
// small device kernels
// ------------------------------------------------------------------------------------------------

// likelihoodAlpha = lngamma(sum alpha) - sum lngamma(alpha_k)  (DocumentMapper.java:121-126)
__global__ void prep_kernel(const double *alpha, int K, double *lik_alpha) {
  __shared__ double sa[32], sl[32];
  double a = 0.0, l = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    a += alpha[k];
    l += lgamma(alpha[k]);
  }
  a = warp_sum(a);
  l = warp_sum(l);
  if ((threadIdx.x & 31) == 0) {
    sa[threadIdx.x >> 5] = a;
    sl[threadIdx.x >> 5] = l;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tl = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
      ta += sa[w];
      tl += sl[w];
    }
    *lik_alpha = lgamma(ta) - tl;
  }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double u01(uint64_t h) {  // (0,1)
  return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// retrieveBeta's initialisation log(2*U1/V + U2) (DocumentMapper.java:456) for rows with
// only_absent ? present[v]==0 : all.
__global__ void init_beta_kernel(double *logbeta, const uint8_t *present, int only_absent, int V,
                                 int K, uint64_t seed) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t n = (int64_t)V * K;
  if (i >= n) return;
  int v = (int)(i / K);
  if (only_absent && present[v]) return;
  uint64_t h1 = splitmix64(seed ^ splitmix64(2 * (uint64_t)i));
  uint64_t h2 = splitmix64(seed ^ splitmix64(2 * (uint64_t)i + 1));
  logbeta[i] = log(2.0 * u01(h1) / (double)V + u01(h2));
}

// importBeta: logbeta = value - (double)normaliserFloat for present rows (DocumentMapper.java:511)
__global__ void import_beta_kernel(double *logbeta, const double *dl, const float *nrm,
                                   const uint8_t *present, int V, int K) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t n = (int64_t)V * K;
  if (i >= n) return;
  int v = (int)(i / K), k = (int)(i - (int64_t)v * K);
  if (present[v]) logbeta[i] = dl[i] - (double)nrm[k];
}

// m_v = max_k logbeta_vk; E_vk = exp(logbeta_vk - m_v), zero in the KS-K padding.  One warp per term.
__global__ void expand_rows_kernel(const double *logbeta, double *E, double *rowmax, int V, int K,
                                   int KS) {
  int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= V) return;
  const double *lb = logbeta + (size_t)warp * K;
  double mx = -INFINITY;
  for (int k = lane; k < K; k += 32) mx = fmax(mx, lb[k]);
  mx = warp_max(mx);
  double *e = E + (size_t)warp * KS;
  for (int k = lane; k < KS; k += 32) e[k] = k < K ? exp(lb[k] - mx) : 0.0;
  if (lane == 0) rowmax[warp] = mx;
}

// K2a: per-slab partial column sums of lambda = eta + stats over the terms seen in the corpus
// (the logNormalizeFactor fold of TermReducer.java:189,212, in linear space), and present[v].
// A term was emitted by some mapper iff sum_k stats[v][k] = n_v > 0.
struct EtaSpec {  // eta per (term, topic) cell: uniform, or the informed prior's two levels
  double eta, hi, lo;
  const uint32_t *seed_bits;  // V*K bits, term-major; NULL = uniform eta
  __device__ __forceinline__ double at(int v, int k, int K) const {
    if (!seed_bits) return eta;
    const size_t i = (size_t)v * K + k;
    return (seed_bits[i >> 5] >> (i & 31)) & 1u ? hi : lo;
  }
};

__global__ void finalize_colsum_kernel(const double *stats, int V, int K, EtaSpec eta, int slab,
                                       double *partial, uint8_t *present) {
  extern __shared__ double rs[];  // slab row sums
  int v0 = blockIdx.x * slab, v1 = min(V, v0 + slab);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int v = v0 + warp; v < v1; v += nw) {
    const double *row = stats + (size_t)v * K;
    double s = 0.0;
    for (int k = lane; k < K; k += 32) s += row[k];
    s = warp_sum(s);
    if (lane == 0) {
      rs[v - v0] = s;
      present[v] = s > 0.0 ? 1 : 0;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double acc = 0.0;
    for (int v = v0; v < v1; v++)
      if (rs[v - v0] > 0.0) acc += eta.at(v, k, K) + stats[(size_t)v * K + k];
    partial[(size_t)blockIdx.x * K + k] = acc;
  }
}

// K2b: normaliser_k = (float) digamma(sum_v lambda_kv)  (TermReducer.java:173,233); fixed order.
__global__ void finalize_norm_kernel(const double *partial, int nslabs, int K, float *nrm,
                                     double *colsum) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double s = 0.0;
  for (int b = 0; b < nslabs; b++) s += partial[(size_t)b * K + k];
  colsum[k] = s;
  nrm[k] = s > 0.0 ? (float)digamma_literal(s) : 0.0f;
}

// K2c: per term row: digamma(lambda) (the beta file value, TermReducer.java:195,213), the
// importBeta subtraction (DocumentMapper.java:511), row max and E.  One warp per term.
__global__ void finalize_rows_kernel(const double *stats, const uint8_t *present, const float *nrm,
                                     EtaSpec eta, int V, int K, int KS, double *dl, double *logbeta,
                                     double *E, double *rowmax) {
  int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= V) return;
  if (!present[warp]) return;  // no row in the beta file; keeps its previous logbeta / E
  const double *row = stats + (size_t)warp * K;
  double *dlr = dl + (size_t)warp * K, *lbr = logbeta + (size_t)warp * K;
  double mx = -INFINITY;
  for (int k = lane; k < K; k += 32) {
    double lam = eta.at(warp, k, K) + row[k];
    double v = digamma_literal(lam);
    dlr[k] = v;
    double lb = v - (double)nrm[k];
    lbr[k] = lb;
    mx = fmax(mx, lb);
  }
  mx = warp_max(mx);
  __syncwarp();
  double *e = E + (size_t)warp * KS;
  for (int k = lane; k < KS; k += 32) e[k] = k < K ? exp(lbr[k] - mx) : 0.0;
  if (lane == 0) rowmax[warp] = mx;
}

// K3 vector: VariationalInference.updateVectorAlpha (VariationalInference.java:409-511) with the
// Java aliasing semantics: the first Newton pass writes a fresh vector, later passes update in place,
// a singular step (alpha_i <= step) at index i leaves entries [0,i) updated once the arrays alias.
// Single CTA; thread i owns topic i (+ blockDim strides).  io: alpha in/out.
__global__ void alpha_vector_kernel(double *alpha, const double *ss, int K, int n_docs,
                                    double *work /* 3K: grad, hess, fresh */) {
  __shared__ double red[2][32];
  __shared__ int sh_first, sh_flag;
  double *grad = work, *hess = work + K, *fresh = work + 2 * K;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, NW = T >> 5;
  const double D = (double)n_docs;
  auto block_sum2 = [&](double a, double b, double &oa, double &ob) {
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();
    if (lane == 0) {
      red[0][warp] = a;
      red[1][warp] = b;
    }
    __syncthreads();
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < NW; w++) {
      ta += red[0][w];
      tb += red[1][w];
    }
    oa = ta;
    ob = tb;
  };
  double *cur = alpha;   // "alphaVector"
  bool aliased = false;  // alphaVectorUpdate == alphaVector ?
  int decay = 0, iter = 0;
  double alpha_sum, dummy;
  {
    double a = 0.0;
    for (int i = tid; i < K; i += T) a += cur[i];
    block_sum2(a, 0.0, alpha_sum, dummy);
  }
  bool keep_going = true;
  while (keep_going) {
    double g_h = 0.0, one_h = 0.0;
    int bad = 0;
    const double dsum = digamma_literal(alpha_sum);
    for (int i = tid; i < K; i += T) {
      double g = D * (dsum - digamma_literal(cur[i])) + ss[i];
      double h = -D * trigamma_literal(cur[i]);
      grad[i] = g;
      hess[i] = h;
      if (isinf(g)) bad = 1;
      g_h += g / h;
      one_h += 1.0 / h;
    }
    if (tid == 0) sh_flag = 0;
    double sum_g_h, sum_1_h;
    block_sum2(g_h, one_h, sum_g_h, sum_1_h);
    if (bad) atomicOr(&sh_flag, 1);
    __syncthreads();
    if (sh_flag) break;  // ArithmeticException path (VariationalInference.java:440-443,505-508)
    const double z = D * trigamma_literal(alpha_sum);
    const double c = sum_g_h / (1.0 / z + sum_1_h);

    double *upd = aliased ? cur : fresh;
    for (;;) {
      const double scale = pow((double)0.8f, (double)decay);
      if (tid == 0) sh_first = K;
      __syncthreads();
      for (int i = tid; i < K; i += T) {
        double step = scale * (grad[i] - c) / hess[i];
        if (cur[i] <= step) atomicMin(&sh_first, i);
      }
      __syncthreads();
      const int first = sh_first;
      // entries before the first singular index are written (sequential loop with break)
      for (int i = tid; i < first; i += T) {
        double step = scale * (grad[i] - c) / hess[i];
        upd[i] = cur[i] - step;
      }
      __syncthreads();
      if (first < K) {
        decay++;
        upd = cur;  // VariationalInference.java:471
        aliased = true;
        if (decay > 10) break;
      } else {
        break;
      }
    }
    double a = 0.0;
    int kg = 0;
    for (int i = tid; i < K; i += T) {
      a += upd[i];
      if (fabs((upd[i] - cur[i]) / cur[i]) >= (double)0.000001f) kg = 1;
    }
    if (tid == 0) sh_flag = 0;
    block_sum2(a, 0.0, alpha_sum, dummy);
    if (kg) atomicOr(&sh_flag, 1);
    __syncthreads();
    keep_going = sh_flag != 0;
    __syncthreads();
    if (iter >= 1000) keep_going = false;
    if (decay > 10) break;
    iter++;
    if (!aliased) {  // alphaVector = alphaVectorUpdate (VariationalInference.java:500)
      cur = fresh;
      aliased = true;
    }
  }
  __syncthreads();
  if (cur != alpha)
    for (int i = tid; i < K; i += T) alpha[i] = cur[i];
}

// K3 scalar: updateScalarAlpha (VariationalInference.java:573-625) behind the symmetric wrapper
// (:295-307): alpha_init = mean(alpha), ss = sum_k alpha_ss_k.  One thread (O(1) work per step).
__device__ double scalar_alpha_newton(int K, int n_docs, double alpha_init, double ss) {
  int iter = 0;
  double a = alpha_init;
  const double Kd = (double)K, D = (double)n_docs;
  for (;;) {
    iter++;
    if (isnan(a) || isinf(a)) {
      alpha_init *= 10.0;
      a = alpha_init;
    }
    double sum = a * Kd;
    double g = D * (Kd * digamma_literal(sum) - Kd * digamma_literal(a)) + ss;
    double h = D * (Kd * Kd * trigamma_literal(sum) - Kd * trigamma_literal(a));
    a = exp(log(a) - g / (h * a + g));
    if (fabs(g) < (double)0.000001f) break;
    if (iter > 1000) break;
  }
  return a;
}
__global__ void alpha_scalar_kernel(double *alpha, const double *ss, int K, int n_docs) {
  __shared__ double res;
  if (threadIdx.x == 0) {
    double tot = 0.0, old = 0.0;
    for (int i = 0; i < K; i++) {
      tot += ss[i];
      old += alpha[i];
    }
    old /= (double)K;
    res = scalar_alpha_newton(K, n_docs, old, tot);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K; i += blockDim.x) alpha[i] = res;
}
__global__ void alpha_scalar_direct_kernel(int K, int n_docs, double alpha_init, double ss,
                                           double *out) {
  *out = scalar_alpha_newton(K, n_docs, alpha_init, ss);
}

__global__ void eval_special_kernel(int which, int64_t n, const double *x, double *y) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = x[i];
  if (which == 4) {  // the interleaved short-chain form the warp-group kernel uses
    const double in[2] = {v, v + 1.0};
    double out[2];
    exp_digamma_n<2>(in, out);
    y[i] = out[0];
    return;
  }
  y[i] = which == 0 ? digamma_literal(v)
                    : which == 1 ? trigamma_literal(v) : which == 2 ? lgamma(v) : exp_digamma(v);
}


// DFMA throughput probe: the E-step is bound by the FP64 pipe, not HBM, so
// bench.py reports its flop rate against this measured peak next to the HBM roofline.
__global__ void fp64_peak_kernel(double *out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) out[0] = s;
}

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen (no link-time dependency; the Java/ctypes host decides whether to shard)
// ------------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId_t;
enum { kNcclSum = 0, kNcclInt64 = 4, kNcclFloat64 = 8 };
struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_t, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::string err;
  bool load() {
    if (handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    for (const char *n : names) {
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) {
      err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
      return false;
    }
#define LOADSYM(field, name)                                        \
  *(void **)(&field) = dlsym(handle, name);                         \
  if (!field) {                                                     \
    err = std::string("missing NCCL symbol ") + name;               \
    return false;                                                   \
  }
    LOADSYM(GetUniqueId, "ncclGetUniqueId");
    LOADSYM(CommInitRank, "ncclCommInitRank");
    LOADSYM(AllReduce, "ncclAllReduce");
    LOADSYM(GroupStart, "ncclGroupStart");
    LOADSYM(GroupEnd, "ncclGroupEnd");
    LOADSYM(CommDestroy, "ncclCommDestroy");
    LOADSYM(GetErrorString, "ncclGetErrorString");
#undef LOADSYM
    return true;
  }
};
NcclApi g_nccl;
std::string g_global_error;
}  // namespace

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct mrlda_ctx {
  mrlda_config cfg;
  int K = 0, V = 0, KS = 0;
  const EstepVariant *variant = nullptr;    // general kernel (any document length)
  const Estep2Variant *variant2 = nullptr;  // register-stationary kernel (documents that fit one SM)
  const Estep2fVariant *variant2f = nullptr;  // fp32-mode counterpart of variant2 (same NW, CPW)
  // second cluster class beside variant2 when it is a cluster variant (K > 256): twice the cluster
  // size at half the column slice, for documents longer than variant2 holds
  const Estep2Variant *variant2_wide = nullptr;
  const Estep3Variant *variant3 = nullptr;  // warp-group kernel (several documents in flight per SM)
  // warp-group kernel with the topics split across a cluster (K > 208); takes the place of variant3
  const Estep3cVariant *variant3c = nullptr;
  int ES = 0;                               // row stride of E in doubles
  std::vector<int32_t> h_nnz_sorted;        // nnz per document in doc_order (descending)
  int num_sms = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // general E-step kernel, concurrent with the register-stationary one
  cudaEvent_t ev[8] = {};
  std::string err;
  int64_t launches = 0;

  // corpus
  int64_t D = 0, Z = 0, n_tokens = 0;
  size_t cap_D = 0, cap_Z = 0;  // capacity of the corpus buffers
  int64_t *d_row_ptr = nullptr;
  int32_t *d_term = nullptr, *d_count = nullptr, *d_order = nullptr, *d_tokens = nullptr;
  double *d_gamma = nullptr;
  bool gamma_valid = false;
  int p90_rows = 0, max_rows = 0;

  // model
  double *d_alpha = nullptr, *d_lik_alpha = nullptr;
  double *d_logbeta = nullptr, *d_E = nullptr, *d_rowmax = nullptr, *d_dl = nullptr;
  double *d_stats = nullptr;
  float *d_nrm = nullptr;
  uint8_t *d_present = nullptr;
  bool alpha_set = false, beta_set = false, have_dl = false;
  uint32_t *d_seed_bits = nullptr;  // informed prior (NULL: uniform eta)
  void *d_tmap_E = nullptr;  // CUtensorMap of E (TMA row gathers of the warp-group kernel)
  unsigned long long *d_stats_fx = nullptr;     // deterministic mode: V x K x 3 fixed-point limbs
  unsigned long long *d_alpha_ss_fx = nullptr;  // deterministic mode: K x 3

  // per-iteration
  double *d_alpha_ss = nullptr;   // K
  long long *d_counters = nullptr; // [0] ll counter, [1] n_docs, [2] n_tokens, [3] slow rows
  int *d_work = nullptr;           // [0] work queue, [1] slow rows, [2] estep2f, [4..9] estep3 classes, [10..15] estep2 classes
  long long *d_phase = nullptr;    // phase cycle counters (MRLDA_PHASE_TIMING builds)
  double *d_partial = nullptr, *d_colsum = nullptr, *d_alpha_work = nullptr;
  int nslabs = 0, slab = 0;
  int64_t total_docs = 0, total_tokens = 0, terms_seen = 0;

  ncclComm_t comm = nullptr;
};

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      set_err(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return MRLDA_ECUDA;                                                                \
    }                                                                                    \
  } while (0)

static void set_err(mrlda_ctx *ctx, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  g_global_error = buf;
}

static const EstepVariant *select_variant(int K) {
  const EstepVariant *tabs[] = {g_estep_variants_L1, g_estep_variants_L2, g_estep_variants_L4,
                                g_estep_variants_L8, g_estep_variants_L16, g_estep_variants_L32};
  const int ns[] = {g_estep_variants_L1_n, g_estep_variants_L2_n, g_estep_variants_L4_n,
                    g_estep_variants_L8_n, g_estep_variants_L16_n, g_estep_variants_L32_n};
  const char *force = getenv("MRLDA_ESTEP_VARIANT");  // "L,KPL" — profiling aid
  int fl = 0, fk = 0;
  if (force && sscanf(force, "%d,%d", &fl, &fk) != 2) fl = fk = 0;
  const EstepVariant *best = nullptr;
  for (int t = 0; t < 6; t++)
    for (int i = 0; i < ns[t]; i++) {
      const EstepVariant *v = &tabs[t][i];
      int ks = v->L * v->KPL;
      if (ks < K) continue;
      if (fl && v->L == fl && v->KPL == fk) return v;
      if (!best || ks < best->L * best->KPL ||
          (ks == best->L * best->KPL && v->KPL < best->KPL))
        best = v;
    }
  return best;
}

static const Estep2Variant *select_variant2(int K) {
  const Estep2Variant *tabs[] = {g_estep2_variants_W1, g_estep2_variants_W2, g_estep2_variants_W4,
                                 g_estep2_variants_W8, g_estep2_variants_C2, g_estep2_variants_C4};
  const int ns[] = {g_estep2_variants_W1_n, g_estep2_variants_W2_n, g_estep2_variants_W4_n,
                    g_estep2_variants_W8_n, g_estep2_variants_C2_n, g_estep2_variants_C4_n};
  const char *force = getenv("MRLDA_ESTEP2_VARIANT");  // "NW,CPW[,RREG]" — profiling aid
  int fw = 0, fc = 0, fr = 0;
  if (force && sscanf(force, "%d,%d,%d", &fw, &fc, &fr) < 2) fw = fc = fr = 0;
  const int wps = 8;       // warps per SM: the 16-warp class (128 registers) measured slower and is gone
  const int max_cpw = 26;
  // fewest warps per CTA whose column slice stays <= max_cpw doubles, then the narrowest slice;
  // beyond 8 warps x 32 columns the topic split continues across a cluster of 2 or 4 CTAs
  const Estep2Variant *best = nullptr;
  for (int t = 0; t < 6; t++) {
    for (int i = 0; i < ns[t]; i++) {
      const Estep2Variant *v = &tabs[t][i];
      if (v->CL * v->NW * v->CPW < K) continue;
      if (fw) {
        if (v->CL == 1 && v->NW == fw && v->CPW == fc &&
            (fr ? v->RREG == fr : v->RREG == estep2_rreg(v->CPW, v->WPS)) && v->WPS == wps)
          return v;
        continue;
      }
      if (v->WPS != wps || v->RREG != estep2_rreg(v->CPW, v->WPS)) continue;
      if (!best || v->CPW < best->CPW) best = v;
    }
    if (best) {
      if (best->CPW <= max_cpw || t >= 3) return best;
      best = nullptr;  // slice too wide at this warp count: try more warps
    }
  }
  return best;
}

// The wide cluster class of a cluster variant (see mrlda_ctx::variant2_wide).
static const Estep2Variant *select_cluster_wide(const Estep2Variant *base, int K) {
  if (!base || base->CL < 2 || base->CL > 4) return nullptr;
  const Estep2Variant *tab = base->CL == 2 ? g_estep2_variants_X4 : g_estep2_variants_X8;
  const int n = base->CL == 2 ? g_estep2_variants_X4_n : g_estep2_variants_X8_n;
  const Estep2Variant *best = nullptr;
  for (int i = 0; i < n; i++) {
    const Estep2Variant *v = &tab[i];
    // same padded topic count as the base variant: E has one row stride
    if (v->CL * v->NW * v->CPW < K || v->CL * v->NW * v->CPW > base->CL * base->NW * base->CPW) continue;
    if (!best || v->CPW < best->CPW) best = v;
  }
  return best;
}

// Warp-group kernel: narrowest row (L * KPL >= K).  MRLDA_ESTEP_KERNEL=2 (register-stationary
// kernel of round 1) or =1 (general kernel only) switch it off for A/B runs.
static const Estep3Variant *select_variant3(int K) {
  const char *fk = getenv("MRLDA_ESTEP_KERNEL");
  if (fk && (atoi(fk) == 1 || atoi(fk) == 2)) return nullptr;
  // up to 104 topics the register-stationary kernel splits its narrower column slices over fewer
  // warps (several CTAs per SM); a row of the warp-group kernel is 200 or 208 columns wide, so fewer
  // topics would mostly be padding (MRLDA_ESTEP3_MIN_TOPICS overrides the bound: profiling aid)
  const char *fm = getenv("MRLDA_ESTEP3_MIN_TOPICS");
  if (K < (fm ? atoi(fm) : 105)) return nullptr;
  const int want_nw = 8;  // 12- and 16-warp CTAs (no register batches) measured 1.5x / 1.8x slower
  const Estep3Variant *tabs[] = {g_estep3_variants_L8, g_estep3_variants_L8o, g_estep3_variants_L8n17,
                                 g_estep3_variants_L8n20, g_estep3_variants_L8n22};
  const int ns[] = {g_estep3_variants_L8_n, g_estep3_variants_L8o_n, g_estep3_variants_L8n17_n,
                    g_estep3_variants_L8n20_n, g_estep3_variants_L8n22_n};
  const Estep3Variant *best = nullptr;
  for (size_t t = 0; t < sizeof(ns) / sizeof(ns[0]); t++)
    for (int i = 0; i < ns[t]; i++) {
      const Estep3Variant *v = &tabs[t][i];
      if (v->L * v->KPL < K || v->NW != want_nw) continue;
      if (!best || v->L * v->KPL < best->L * best->KPL) best = v;
    }
  return best;
}

// Cluster warp-group kernel: fewest padded topics (CL * L * KPL >= K), then the smallest cluster.
// MRLDA_ESTEP_KERNEL=1 / 2 switch it off like the single-CTA one.
static const Estep3cVariant *select_variant3c(int K) {
  const char *fk = getenv("MRLDA_ESTEP_KERNEL");
  if (fk && (atoi(fk) == 1 || atoi(fk) == 2)) return nullptr;
  // up to 256 topics the register-stationary kernel still holds a document on one SM and beats a
  // 2-CTA cluster of warp groups (K = 224: 73 against 83 ms, K = 256: 67 against 84 ms per 40 000
  // documents; from K = 320 on the cluster kernel wins: 84 against 120 ms; tools/kcross_on_box.sh)
  if (K <= 256) return nullptr;
  const Estep3cVariant *tabs[] = {g_estep3c_variants_C2, g_estep3c_variants_C2o, g_estep3c_variants_C3, g_estep3c_variants_C3o, g_estep3c_variants_C4, g_estep3c_variants_C4o, g_estep3c_variants_C5, g_estep3c_variants_C5o};
  const int ns[] = {g_estep3c_variants_C2_n, g_estep3c_variants_C2o_n, g_estep3c_variants_C3_n, g_estep3c_variants_C3o_n, g_estep3c_variants_C4_n, g_estep3c_variants_C4o_n, g_estep3c_variants_C5_n, g_estep3c_variants_C5o_n};
  const Estep3cVariant *best = nullptr;
  for (size_t t = 0; t < sizeof(ns) / sizeof(ns[0]); t++)
    for (int i = 0; i < ns[t]; i++) {
      const Estep3cVariant *v = &tabs[t][i];
      const int ks = v->CL * v->v.L * v->v.KPL;
      // every CTA of the cluster must own at least one topic
      if (ks < K || (v->CL - 1) * v->v.L * v->v.KPL >= K) continue;
      if (!best || ks < best->CL * best->v.L * best->v.KPL ||
          (ks == best->CL * best->v.L * best->v.KPL && v->CL < best->CL))
        best = v;
    }
  return best;
}

static const Estep2fVariant *select_variant2f(const Estep2Variant *v2) {
  if (!v2 || v2->CL != 1 || v2->WPS != 8) return nullptr;
  const Estep2fVariant *tabs[] = {g_estep2f_variants_W1, g_estep2f_variants_W2, g_estep2f_variants_W4,
                                  g_estep2f_variants_W8};
  const int ns[] = {g_estep2f_variants_W1_n, g_estep2f_variants_W2_n, g_estep2f_variants_W4_n,
                    g_estep2f_variants_W8_n};
  for (int t = 0; t < 4; t++)
    for (int i = 0; i < ns[t]; i++)
      if (tabs[t][i].NW == v2->NW && tabs[t][i].CPW == v2->CPW) return &tabs[t][i];
  return nullptr;
}

extern "C" {

const char *mrlda_version(void) { return "mrlda_b200 0.1 (sm_100a, fp64 + optional fp32 E-step)"; }

void mrlda_default_config(mrlda_config *cfg) {
  memset(cfg, 0, sizeof *cfg);
  cfg->sweep_cap = 100;  // Settings.java:54
  cfg->learning = 1;     // Settings.java:46
  cfg->world_size = 1;
  cfg->update_alpha = 1;
  cfg->eta = 1e-12;  // Settings.java:58
  cfg->seed = 0x4d724c4441ull;
}

const char *mrlda_last_error(const mrlda_ctx *ctx) {
  return ctx ? ctx->err.c_str() : g_global_error.c_str();
}

int64_t mrlda_kernel_launches(const mrlda_ctx *ctx) { return ctx ? ctx->launches : 0; }

int mrlda_describe_kernels(int n_topics, char *buf, size_t len) {
  if (!buf || len == 0 || n_topics <= 0) return MRLDA_EINVAL;
  if (n_topics > MRLDA_MAX_TOPICS) return MRLDA_EUNSUPPORTED;
  const int K = n_topics;
  const size_t sm_total = 233472, optin = 232448;  // B200: 228 KB per SM, 227 KB per CTA
  std::string out;
  char tmp[256];
  const Estep3Variant *v3 = select_variant3(K);
  const Estep3cVariant *v3c = v3 ? nullptr : select_variant3c(K);
  const Estep2Variant *v2 = select_variant2(K);
  const Estep2Variant *v2w = select_cluster_wide(v2, K);
  int held = 0;  // rows the classes so far hold
  if (v3 || v3c) {
    const Estep3Variant *v = v3 ? v3 : &v3c->v;
    const int cl = v3 ? 1 : v3c->CL;
    const size_t per = std::min(sm_total - 1024, optin) & ~(size_t)15;
    int M = v->MMAX;
    while (M > 0 && v->smem_bytes(M, v->NW) > per) M--;
    held = v->NW * (v->R + v->T + M) * (32 / v->L);
    snprintf(tmp, sizeof tmp, "%s<%d,%d,%d,%d,G%s> x %d CTA%s per document, G = 1/2/4/8 warps, <= %d rows",
             v3 ? "estep3_kernel" : "estep3c_kernel", v->L, v->KPL, v->R, v->NW, v3 ? "" : ",CL", cl,
             cl > 1 ? "s" : "", held);
    out += tmp;
  }
  const Estep2Variant *cand[2] = {v2, (v3c && v2w) ? v2w : nullptr};
  if (!v3 && !v3c && v2w) cand[1] = v2w;
  if (v3) cand[0] = cand[1] = nullptr;       // single-CTA warp groups: the general kernel takes the rest
  if (v3c) cand[0] = nullptr;                // cluster warp groups: only a class that holds more rows
  for (int k = 0; k < 2; k++) {
    const Estep2Variant *v = cand[k];
    if (!v) continue;
    const int c2 = v->WPS / v->NW;
    const size_t per = std::min((sm_total - (size_t)c2 * 1024) / c2, optin) & ~(size_t)15;
    const int rt = estep2_tmem_slots(v->NW, v->CPW, v->WPS);
    int rsm = 0;
    while (v->smem_bytes(rsm + 1, 32 * (v->RREG + rt + rsm + 1)) <= per) rsm++;
    const int cap = 32 * (v->RREG + rt + rsm);
    if (cap <= held) continue;
    snprintf(tmp, sizeof tmp, "%sestep2_kernel<%d,%d,%d,%d,%d> x %d CTA%s per document, <= %d rows",
             out.empty() ? "" : " | ", v->NW, v->CPW, v->RREG, v->WPS, v->CL, v->CL, v->CL > 1 ? "s" : "", cap);
    out += tmp;
    held = cap;
  }
  const EstepVariant *vg = select_variant(K);
  if (!vg) return MRLDA_EUNSUPPORTED;
  snprintf(tmp, sizeof tmp, "%sestep_kernel<%d,%d> longer documents and deterministic mode", out.empty() ? "" : " | ",
           vg->L, vg->KPL);
  out += tmp;
  snprintf(buf, len, "%s", out.c_str());
  return (int)std::min(out.size(), len - 1);
}

void mrlda_destroy(mrlda_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
  void *ptrs[] = {ctx->d_row_ptr, ctx->d_term,   ctx->d_count,    ctx->d_order,  ctx->d_tokens,
                  ctx->d_gamma,   ctx->d_alpha,  ctx->d_lik_alpha, ctx->d_logbeta, ctx->d_E,
                  ctx->d_rowmax,  ctx->d_dl,     ctx->d_stats,    ctx->d_nrm,    ctx->d_present,
                  ctx->d_alpha_ss, ctx->d_counters, ctx->d_work,  ctx->d_partial, ctx->d_colsum,
                  ctx->d_alpha_work, ctx->d_phase, ctx->d_seed_bits, ctx->d_stats_fx,
                  ctx->d_alpha_ss_fx, ctx->d_tmap_E};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  for (auto &e : ctx->ev)
    if (e) cudaEventDestroy(e);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int mrlda_create(const mrlda_config *cfg, mrlda_ctx **out) {
  if (!cfg || !out) {
    set_err(nullptr, "mrlda_create: null argument");
    return MRLDA_EINVAL;
  }
  *out = nullptr;
  if (cfg->n_topics <= 0 || cfg->n_terms <= 0 || cfg->sweep_cap < 2 || cfg->world_size < 1 ||
      cfg->rank < 0 || cfg->rank >= cfg->world_size || !(cfg->eta > 0.0)) {
    set_err(nullptr, "mrlda_create: invalid config (topics=%d terms=%d sweep_cap=%d rank=%d/%d eta=%g)",
            cfg->n_topics, cfg->n_terms, cfg->sweep_cap, cfg->rank, cfg->world_size, cfg->eta);
    return MRLDA_EINVAL;
  }
  if (cfg->n_topics > MRLDA_MAX_TOPICS) {
    set_err(nullptr, "mrlda_create: %d topics exceed the compiled kernel range (max %d)",
            cfg->n_topics, MRLDA_MAX_TOPICS);
    return MRLDA_EUNSUPPORTED;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    set_err(nullptr, "mrlda_create: no CUDA device (%s); there is no CPU fallback",
            cudaGetErrorString(e));
    return MRLDA_ECUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) {
    set_err(nullptr, "mrlda_create: device %d out of range (%d devices)", cfg->device, ndev);
    return MRLDA_EINVAL;
  }
  mrlda_ctx *ctx = new mrlda_ctx();
  ctx->cfg = *cfg;
  ctx->K = cfg->n_topics;
  ctx->V = cfg->n_terms;
  ctx->variant = select_variant(ctx->K);
  if (!ctx->variant) {
    set_err(nullptr, "mrlda_create: no E-step kernel variant for %d topics", ctx->K);
    delete ctx;
    return MRLDA_EUNSUPPORTED;
  }
  ctx->KS = ctx->variant->L * ctx->variant->KPL;
  ctx->variant2 = select_variant2(ctx->K);
  ctx->variant2f = cfg->fp32_mode ? select_variant2f(ctx->variant2) : nullptr;
  if (!cfg->fp32_mode) ctx->variant2_wide = select_cluster_wide(ctx->variant2, ctx->K);
  ctx->ES = ctx->KS;
  if (ctx->variant2)
    ctx->ES = std::max(ctx->ES, ctx->variant2->CL * ctx->variant2->NW * ctx->variant2->CPW);
  ctx->variant3 = cfg->fp32_mode ? nullptr : select_variant3(ctx->K);
  if (ctx->variant3) ctx->ES = std::max(ctx->ES, ctx->variant3->L * ctx->variant3->KPL);
  if (!ctx->variant3 && !cfg->fp32_mode) ctx->variant3c = select_variant3c(ctx->K);
  if (ctx->variant3c)
    ctx->ES = std::max(ctx->ES, ctx->variant3c->CL * ctx->variant3c->v.L * ctx->variant3c->v.KPL);
  auto fail = [&](int code) {
    g_global_error = ctx->err;
    mrlda_destroy(ctx);
    return code;
  };
#define CKC(call)                                                                            \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      set_err(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return fail(e_ == cudaErrorMemoryAllocation ? MRLDA_ENOMEM : MRLDA_ECUDA);             \
    }                                                                                        \
  } while (0)
  CKC(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CKC(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10) {
    set_err(ctx, "mrlda_create: device %d is sm_%d%d; this library is built for sm_100a only",
            cfg->device, prop.major, prop.minor);
    return fail(MRLDA_ECUDA);
  }
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
  for (auto &ev : ctx->ev) CKC(cudaEventCreate(&ev));
  const size_t K = ctx->K, V = ctx->V, KS = ctx->ES;
  CKC(cudaMalloc(&ctx->d_alpha, K * sizeof(double)));
  CKC(cudaMalloc(&ctx->d_lik_alpha, sizeof(double)));
  CKC(cudaMalloc(&ctx->d_logbeta, V * K * sizeof(double)));
  CKC(cudaMalloc(&ctx->d_E, V * KS * sizeof(double)));
  CKC(cudaMalloc(&ctx->d_rowmax, V * sizeof(double)));
  if (ctx->variant3 || ctx->variant3c) {
    // Tensor map of E for the warp-group kernel's TMA tile::gather4 copies: a 2-D fp64 tensor
    // {columns, terms}, box = one row of the kernel's KS columns (the gather picks four rows).
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CKC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    // (the cluster kernel's CTA r gathers the columns [r ks3, (r + 1) ks3) of the rows)
    const Estep3Variant *v3 = ctx->variant3 ? ctx->variant3 : &ctx->variant3c->v;
    const cuuint32_t ks3 = (cuuint32_t)(v3->L * v3->KPL);
    alignas(64) CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)ks3 * (ctx->variant3 ? 1 : ctx->variant3c->CL), (cuuint64_t)V};
    const cuuint64_t gstride[1] = {(cuuint64_t)KS * sizeof(double)};
    const cuuint32_t box[2] = {ks3, 1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult trc = encode ? encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, ctx->d_E, gdim, gstride, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
                                : CUDA_ERROR_NOT_SUPPORTED;
    if (trc != CUDA_SUCCESS) {
      set_err(ctx, "mrlda_create: cuTensorMapEncodeTiled failed (%d)", (int)trc);
      return fail(MRLDA_ECUDA);
    }
    CKC(cudaMalloc(&ctx->d_tmap_E, sizeof tm));
    CKC(cudaMemcpy(ctx->d_tmap_E, &tm, sizeof tm, cudaMemcpyHostToDevice));
  }
  CKC(cudaMalloc(&ctx->d_dl, V * K * sizeof(double)));
  CKC(cudaMalloc(&ctx->d_stats, V * K * sizeof(double)));
  CKC(cudaMalloc(&ctx->d_nrm, K * sizeof(float)));
  CKC(cudaMalloc(&ctx->d_present, V));
  CKC(cudaMalloc(&ctx->d_alpha_ss, K * sizeof(double)));
  CKC(cudaMalloc(&ctx->d_counters, 4 * sizeof(long long)));
  CKC(cudaMalloc(&ctx->d_work, 16 * sizeof(int)));
  CKC(cudaMalloc(&ctx->d_phase, 16 * sizeof(long long)));
  CKC(cudaMemset(ctx->d_phase, 0, 16 * sizeof(long long)));
  ctx->slab = 256;
  ctx->nslabs = (int)((V + ctx->slab - 1) / ctx->slab);
  CKC(cudaMalloc(&ctx->d_partial, (size_t)ctx->nslabs * K * sizeof(double)));
  CKC(cudaMalloc(&ctx->d_colsum, K * sizeof(double)));
  CKC(cudaMalloc(&ctx->d_alpha_work, 3 * K * sizeof(double)));
  CKC(cudaMemsetAsync(ctx->d_stats, 0, V * K * sizeof(double), ctx->stream));
  CKC(cudaMemsetAsync(ctx->d_present, 0, V, ctx->stream));
  CKC(cudaMemsetAsync(ctx->d_dl, 0, V * K * sizeof(double), ctx->stream));
  CKC(cudaMemsetAsync(ctx->d_nrm, 0, K * sizeof(float), ctx->stream));
  CKC(cudaMemsetAsync(ctx->d_alpha_ss, 0, K * sizeof(double), ctx->stream));
  CKC(cudaStreamSynchronize(ctx->stream));
#undef CKC
  *out = ctx;
  return MRLDA_OK;
}

// deterministic mode: fixed-point limbs (fx_add, estep_kernel.cuh) -> the fp64 statistics the finalize reads
__global__ void fx_to_double_kernel(const unsigned long long *fx, double *out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = fx_value(fx + 3 * i);
}

static int refresh_rows(mrlda_ctx *ctx) {
  int threads = 256;
  int64_t blocks = ((int64_t)ctx->V * 32 + threads - 1) / threads;
  expand_rows_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(ctx->d_logbeta, ctx->d_E,
                                                                     ctx->d_rowmax, ctx->V, ctx->K,
                                                                     ctx->ES);
  ctx->launches++;
  CK(cudaGetLastError());
  return MRLDA_OK;
}

int mrlda_set_corpus_csr(mrlda_ctx *ctx, int64_t n_docs, const int64_t *row_ptr,
                         const int32_t *term_id, const int32_t *count, const int32_t *doc_id,
                         const double *gamma0) {
  (void)doc_id;  // keys are carried by the host side; the kernels address documents by position
  if (!ctx) return MRLDA_EINVAL;
  if (n_docs < 0 || n_docs > 0x7fffffff || !row_ptr || (n_docs > 0 && row_ptr[0] != 0)) {
    set_err(ctx, "mrlda_set_corpus_csr: bad row_ptr / n_docs");
    return MRLDA_EINVAL;
  }
  const int64_t D = n_docs, Z = row_ptr[D];
  if (Z < 0 || (Z > 0 && (!term_id || !count))) {
    set_err(ctx, "mrlda_set_corpus_csr: null term_id/count");
    return MRLDA_EINVAL;
  }
  CK(cudaSetDevice(ctx->cfg.device));
  // device buffers: reuse when the new corpus fits (the end-to-end path sets the corpus every call)
  const size_t K = ctx->K;
  if ((size_t)D > ctx->cap_D || (size_t)Z > ctx->cap_Z) {
    void *olds[] = {ctx->d_row_ptr, ctx->d_term, ctx->d_count, ctx->d_order, ctx->d_tokens, ctx->d_gamma};
    for (void *p : olds)
      if (p) cudaFree(p);
    ctx->d_row_ptr = nullptr;
    ctx->d_term = ctx->d_count = ctx->d_order = ctx->d_tokens = nullptr;
    ctx->d_gamma = nullptr;
    ctx->cap_D = ctx->cap_Z = 0;
    ctx->D = 0;
    CK(cudaMalloc(&ctx->d_row_ptr, (size_t)(D + 1) * sizeof(int64_t)));
    CK(cudaMalloc(&ctx->d_term, std::max<size_t>(1, (size_t)Z) * sizeof(int32_t)));
    CK(cudaMalloc(&ctx->d_count, std::max<size_t>(1, (size_t)Z) * sizeof(int32_t)));
    CK(cudaMalloc(&ctx->d_order, std::max<size_t>(1, (size_t)D) * sizeof(int32_t)));
    CK(cudaMalloc(&ctx->d_tokens, std::max<size_t>(1, (size_t)D) * sizeof(int32_t)));
    CK(cudaMalloc(&ctx->d_gamma, std::max<size_t>(1, (size_t)D * K) * sizeof(double)));
    ctx->cap_D = (size_t)D;
    ctx->cap_Z = (size_t)Z;
  }
  // start the bulk copies first; the host-side checks below overlap with them
  CK(cudaMemcpyAsync(ctx->d_row_ptr, row_ptr, (size_t)(D + 1) * sizeof(int64_t),
                     cudaMemcpyHostToDevice, ctx->stream));
  if (Z) {
    CK(cudaMemcpyAsync(ctx->d_term, term_id, (size_t)Z * sizeof(int32_t), cudaMemcpyHostToDevice,
                       ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_count, count, (size_t)Z * sizeof(int32_t), cudaMemcpyHostToDevice,
                       ctx->stream));
  }
  if (gamma0 && D)
    CK(cudaMemcpyAsync(ctx->d_gamma, gamma0, (size_t)D * K * sizeof(double), cudaMemcpyHostToDevice,
                       ctx->stream));
  else  // documents without content never get a gamma (DocumentMapper.java:197-201): keep them 0
    CK(cudaMemsetAsync(ctx->d_gamma, 0, std::max<size_t>(1, (size_t)D * K) * sizeof(double), ctx->stream));

  // validation + per-document token counts, on all host cores
  std::vector<int32_t> tokens((size_t)D), order((size_t)D), nnz((size_t)D);
  const int nthr = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
  std::vector<int64_t> part_tokens((size_t)nthr, 0), bad_doc((size_t)nthr, -1);
  std::vector<int> bad_kind((size_t)nthr, 0);
  const int V = ctx->V;
  auto work = [&](int t) {
    const int64_t d0 = D * t / nthr, d1 = D * (t + 1) / nthr;
    int64_t tot = 0;
    for (int64_t d = d0; d < d1; d++) {
      const int64_t b = row_ptr[d], e = row_ptr[d + 1];
      if (e < b || e > Z) {
        bad_doc[t] = d;
        bad_kind[t] = 1;
        return;
      }
      int64_t tk = 0;
      int flag = 0;
      for (int64_t z = b; z < e; z++) {
        flag |= (term_id[z] < 1) | (term_id[z] > V) | ((count[z] <= 0) << 1);
        tk += count[z];
      }
      if (flag || tk > 0x7fffffff) {
        bad_doc[t] = d;
        bad_kind[t] = (flag & 1) ? 2 : ((flag & 2) ? 3 : 4);
        return;
      }
      tokens[d] = (int32_t)tk;
      nnz[d] = (int32_t)(e - b);
      tot += tk;
    }
    part_tokens[t] = tot;
  };
  if (D * 4 + Z < (1 << 16) || nthr == 1) {
    for (int t = 0; t < nthr; t++) work(t);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthr; t++) pool.emplace_back(work, t);
    for (auto &th : pool) th.join();
  }
  int64_t total_tokens = 0;
  for (int t = 0; t < nthr; t++) {
    if (bad_doc[t] >= 0) {
      static const char *why[] = {"", "row_ptr not monotone", "term id outside [1, n_terms]",
                                  "non-positive count", "more than 2^31 tokens"};
      cudaStreamSynchronize(ctx->stream);
      ctx->D = 0;
      set_err(ctx, "mrlda_set_corpus_csr: %s in document %lld", why[bad_kind[t]], (long long)bad_doc[t]);
      return MRLDA_EINVAL;
    }
    total_tokens += part_tokens[t];
  }
  // documents by distinct-term count, longest first (stable): counting sort
  {
    int32_t mx = 0;
    for (int64_t d = 0; d < D; d++) mx = std::max(mx, nnz[d]);
    std::vector<int64_t> start((size_t)mx + 2, 0);
    for (int64_t d = 0; d < D; d++) start[(size_t)(mx - nnz[d]) + 1]++;
    for (size_t i = 1; i < start.size(); i++) start[i] += start[i - 1];
    for (int64_t d = 0; d < D; d++) order[(size_t)start[(size_t)(mx - nnz[d])]++] = (int32_t)d;
  }
  ctx->max_rows = D ? nnz[order[0]] : 0;
  ctx->p90_rows = D ? nnz[order[(size_t)(D / 10)]] : 0;
  ctx->h_nnz_sorted.resize((size_t)D);
  for (int64_t i = 0; i < D; i++) ctx->h_nnz_sorted[i] = nnz[order[i]];

  if (D) {
    CK(cudaMemcpyAsync(ctx->d_order, order.data(), (size_t)D * sizeof(int32_t),
                       cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_tokens, tokens.data(), (size_t)D * sizeof(int32_t),
                       cudaMemcpyHostToDevice, ctx->stream));
  }
  ctx->gamma_valid = gamma0 && D;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->D = D;
  ctx->Z = Z;
  ctx->n_tokens = total_tokens;
  return MRLDA_OK;
}

int mrlda_set_alpha(mrlda_ctx *ctx, const double *alpha) {
  if (!ctx || !alpha) return MRLDA_EINVAL;
  for (int k = 0; k < ctx->K; k++)
    if (!(alpha[k] > 0.0) || !(alpha[k] < 1e300)) {
      set_err(ctx, "mrlda_set_alpha: alpha[%d] = %g is not a positive finite number", k, alpha[k]);
      return MRLDA_EINVAL;
    }
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemcpyAsync(ctx->d_alpha, alpha, (size_t)ctx->K * sizeof(double), cudaMemcpyHostToDevice,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->alpha_set = true;
  return MRLDA_OK;
}

int mrlda_get_alpha(mrlda_ctx *ctx, double *alpha) {
  if (!ctx || !alpha) return MRLDA_EINVAL;
  if (!ctx->alpha_set) {
    set_err(ctx, "mrlda_get_alpha: alpha not set");
    return MRLDA_ESTATE;
  }
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemcpyAsync(alpha, ctx->d_alpha, (size_t)ctx->K * sizeof(double), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MRLDA_OK;
}

int mrlda_init_beta_random(mrlda_ctx *ctx) {
  if (!ctx) return MRLDA_EINVAL;
  CK(cudaSetDevice(ctx->cfg.device));
  int64_t n = (int64_t)ctx->V * ctx->K;
  init_beta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
      ctx->d_logbeta, ctx->d_present, 0, ctx->V, ctx->K, ctx->cfg.seed);
  ctx->launches++;
  CK(cudaGetLastError());
  int rc = refresh_rows(ctx);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->beta_set = true;
  ctx->have_dl = false;
  return MRLDA_OK;
}

int mrlda_set_logbeta(mrlda_ctx *ctx, const double *logbeta) {
  if (!ctx || !logbeta) return MRLDA_EINVAL;
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemcpyAsync(ctx->d_logbeta, logbeta, (size_t)ctx->V * ctx->K * sizeof(double),
                     cudaMemcpyHostToDevice, ctx->stream));
  int rc = refresh_rows(ctx);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->beta_set = true;
  ctx->have_dl = false;
  return MRLDA_OK;
}

int mrlda_get_logbeta(mrlda_ctx *ctx, double *logbeta) {
  if (!ctx || !logbeta) return MRLDA_EINVAL;
  if (!ctx->beta_set) {
    set_err(ctx, "mrlda_get_logbeta: beta not set");
    return MRLDA_ESTATE;
  }
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemcpyAsync(logbeta, ctx->d_logbeta, (size_t)ctx->V * ctx->K * sizeof(double),
                     cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MRLDA_OK;
}

int mrlda_set_beta(mrlda_ctx *ctx, const double *digamma_lambda, const float *normaliser,
                   const uint8_t *present) {
  if (!ctx || !digamma_lambda || !normaliser || !present) return MRLDA_EINVAL;
  CK(cudaSetDevice(ctx->cfg.device));
  const size_t n = (size_t)ctx->V * ctx->K;
  CK(cudaMemcpyAsync(ctx->d_dl, digamma_lambda, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_nrm, normaliser, (size_t)ctx->K * sizeof(float), cudaMemcpyHostToDevice,
                     ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_present, present, (size_t)ctx->V, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  // rows absent from the file: retrieveBeta's lazy initialisation (DocumentMapper.java:450-460)
  init_beta_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_logbeta, ctx->d_present, 1, ctx->V,
                                                    ctx->K, ctx->cfg.seed);
  import_beta_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_logbeta, ctx->d_dl, ctx->d_nrm,
                                                      ctx->d_present, ctx->V, ctx->K);
  ctx->launches += 2;
  CK(cudaGetLastError());
  int rc = refresh_rows(ctx);
  if (rc) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->beta_set = true;
  ctx->have_dl = true;
  return MRLDA_OK;
}

int mrlda_get_beta(mrlda_ctx *ctx, double *digamma_lambda, float *normaliser, uint8_t *present) {
  if (!ctx || !digamma_lambda || !normaliser || !present) return MRLDA_EINVAL;
  if (!ctx->have_dl) {
    set_err(ctx, "mrlda_get_beta: no beta in file form (run a learning iteration or mrlda_set_beta)");
    return MRLDA_ESTATE;
  }
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemcpyAsync(digamma_lambda, ctx->d_dl, (size_t)ctx->V * ctx->K * sizeof(double),
                     cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(normaliser, ctx->d_nrm, (size_t)ctx->K * sizeof(float), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CK(cudaMemcpyAsync(present, ctx->d_present, (size_t)ctx->V, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MRLDA_OK;
}

int mrlda_set_informed_prior(mrlda_ctx *ctx, int64_t n, const int32_t *topic, const int32_t *term) {
  if (!ctx || n < 0 || (n > 0 && (!topic || !term))) return MRLDA_EINVAL;
  CK(cudaSetDevice(ctx->cfg.device));
  if (n == 0) {
    if (ctx->d_seed_bits) cudaFree(ctx->d_seed_bits);
    ctx->d_seed_bits = nullptr;
    return MRLDA_OK;
  }
  const size_t words = ((size_t)ctx->V * ctx->K + 31) / 32;
  std::vector<uint32_t> bits(words, 0u);
  for (int64_t i = 0; i < n; i++) {
    if (topic[i] < 1 || topic[i] > ctx->K || term[i] < 1 || term[i] > ctx->V) {
      set_err(ctx, "mrlda_set_informed_prior: (topic %d, term %d) out of range", topic[i], term[i]);
      return MRLDA_EINVAL;
    }
    const size_t at = (size_t)(term[i] - 1) * ctx->K + (size_t)(topic[i] - 1);
    bits[at >> 5] |= 1u << (at & 31);
  }
  if (!ctx->d_seed_bits) CK(cudaMalloc(&ctx->d_seed_bits, words * sizeof(uint32_t)));
  CK(cudaMemcpyAsync(ctx->d_seed_bits, bits.data(), words * sizeof(uint32_t), cudaMemcpyHostToDevice,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MRLDA_OK;
}

int mrlda_get_gamma(mrlda_ctx *ctx, double *gamma) {
  if (!ctx || !gamma) return MRLDA_EINVAL;
  if (!ctx->d_gamma) {
    set_err(ctx, "mrlda_get_gamma: corpus not set");
    return MRLDA_ESTATE;
  }
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemcpyAsync(gamma, ctx->d_gamma, (size_t)ctx->D * ctx->K * sizeof(double),
                     cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MRLDA_OK;
}

int mrlda_get_alpha_ss(mrlda_ctx *ctx, double *alpha_ss) {
  if (!ctx || !alpha_ss) return MRLDA_EINVAL;
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemcpyAsync(alpha_ss, ctx->d_alpha_ss, (size_t)ctx->K * sizeof(double),
                     cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MRLDA_OK;
}

int mrlda_get_stats(mrlda_ctx *ctx, double *stats) {
  if (!ctx || !stats) return MRLDA_EINVAL;
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemcpyAsync(stats, ctx->d_stats, (size_t)ctx->V * ctx->K * sizeof(double),
                     cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return MRLDA_OK;
}

// Launch geometry of K1 for the resident corpus: CTAs per SM, threads per CTA, tile rows.
static void plan_estep(mrlda_ctx *ctx, int *ctas_per_sm, int *threads, int *cap_rows, size_t *smem) {
  const EstepVariant *v = ctx->variant;
  const int rpb = 32 / v->L;
  const int gran = rpb < 2 ? 2 : rpb;  // tile rows: multiple of RPB and even (16-byte alignment)
  const size_t sm_total = 233472;      // 228 KB per SM, 1 KB reserved per resident CTA
  const int want = std::max(ctx->p90_rows, gran);
  int best_c = 1;
  const char *fc = getenv("MRLDA_ESTEP_CTAS_PER_SM");
  // deterministic mode: one CTA per SM always, so that the warps per CTA (and with them the order
  // in which a document's rows are summed) do not depend on the shard's length distribution
  const int forced = ctx->cfg.deterministic ? 1 : (fc ? atoi(fc) : 0);
  for (int c = 8; c >= 1; c--) {
    if (forced && c != forced) continue;
    int t = (v->max_threads / c) & ~31;
    if (t < 64 && c != 1 && !forced) continue;
    if (t < 32) continue;
    size_t per = std::min((sm_total - (size_t)c * 1024) / c, ctx->smem_optin) & ~(size_t)15;
    size_t fixed = v->smem_bytes(t / 32, 0);
    if (per <= fixed) continue;
    int cap = (int)((per - fixed) / ((size_t)ctx->KS * 8 + 8));
    cap = cap / gran * gran;
    if (cap < gran) continue;
    if (cap >= want || c == 1 || forced) {
      best_c = c;
      break;
    }
  }
  int c = best_c;
  int t = std::max(32, (v->max_threads / c) & ~31);
  size_t per = std::min((sm_total - (size_t)c * 1024) / c, ctx->smem_optin) & ~(size_t)15;
  size_t fixed = v->smem_bytes(t / 32, 0);
  int cap = per > fixed ? (int)((per - fixed) / ((size_t)ctx->KS * 8 + 8)) : 0;
  cap = cap / gran * gran;
  int need = (std::max(ctx->max_rows, 1) + gran - 1) / gran * gran;
  if (cap > need) cap = need;  // do not reserve more than the longest document needs
  if (cap < gran) cap = gran;
  *ctas_per_sm = c;
  *threads = t;
  *cap_rows = cap;
  *smem = v->smem_bytes(t / 32, cap);
}

static int exchange(mrlda_ctx *ctx) {
  if (ctx->cfg.world_size <= 1) return MRLDA_OK;
  if (!ctx->comm) {
    set_err(ctx, "world_size %d but mrlda_comm_init was not called", ctx->cfg.world_size);
    return MRLDA_ESTATE;
  }
  int rc = g_nccl.GroupStart();
  if (ctx->cfg.deterministic) {
    // integer limbs add up exactly, whatever order the ring visits the ranks in
    if (!rc && ctx->cfg.learning)
      rc = g_nccl.AllReduce(ctx->d_stats_fx, ctx->d_stats_fx, (size_t)ctx->V * ctx->K * 3, kNcclInt64,
                            kNcclSum, ctx->comm, ctx->stream);
    if (!rc)
      rc = g_nccl.AllReduce(ctx->d_alpha_ss_fx, ctx->d_alpha_ss_fx, (size_t)ctx->K * 3, kNcclInt64,
                            kNcclSum, ctx->comm, ctx->stream);
  } else {
    if (!rc && ctx->cfg.learning)
      rc = g_nccl.AllReduce(ctx->d_stats, ctx->d_stats, (size_t)ctx->V * ctx->K, kNcclFloat64,
                            kNcclSum, ctx->comm, ctx->stream);
    if (!rc)
      rc = g_nccl.AllReduce(ctx->d_alpha_ss, ctx->d_alpha_ss, (size_t)ctx->K, kNcclFloat64, kNcclSum,
                            ctx->comm, ctx->stream);
  }
  if (!rc)
    rc = g_nccl.AllReduce(ctx->d_counters, ctx->d_counters, 4, kNcclInt64, kNcclSum, ctx->comm,
                          ctx->stream);
  int rc2 = g_nccl.GroupEnd();
  if (rc || rc2) {
    set_err(ctx, "NCCL all-reduce failed: %s", g_nccl.GetErrorString(rc ? rc : rc2));
    return MRLDA_ENCCL;
  }
  return MRLDA_OK;
}

int mrlda_e_step(mrlda_ctx *ctx, mrlda_iter_result *out) {
  if (!ctx) return MRLDA_EINVAL;
  if (!ctx->d_row_ptr || !ctx->alpha_set || !ctx->beta_set) {
    set_err(ctx, "mrlda_e_step: %s not set", !ctx->d_row_ptr ? "corpus" : !ctx->alpha_set ? "alpha" : "beta");
    return MRLDA_ESTATE;
  }
  if (ctx->D <= 0 && ctx->cfg.world_size <= 1) {
    // a job without input records leaves no alpha statistics file and the reference's driver fails
    // on importAlpha (VariationalInference.java:288-289); fail here instead of producing NaNs.
    // In a multi-rank job a rank with an empty shard launches no kernel but still takes part in the
    // all-reduce below (otherwise the other ranks would wait for it forever); the job fails, on
    // every rank alike, only when the global document count is 0.
    set_err(ctx, "mrlda_e_step: the corpus has no documents");
    return MRLDA_ESTATE;
  }
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = ctx->stream;
  const int K = ctx->K;
  int launches0 = (int)ctx->launches;
  CK(cudaEventRecord(ctx->ev[0], st));
  prep_kernel<<<1, 256, 0, st>>>(ctx->d_alpha, K, ctx->d_lik_alpha);
  ctx->launches++;
  if (ctx->cfg.learning) CK(cudaMemsetAsync(ctx->d_stats, 0, (size_t)ctx->V * K * sizeof(double), st));
  CK(cudaMemsetAsync(ctx->d_alpha_ss, 0, (size_t)K * sizeof(double), st));
  long long hc[4] = {0, (long long)ctx->D, (long long)ctx->n_tokens, 0};
  CK(cudaMemcpyAsync(ctx->d_counters, hc, sizeof hc, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(ctx->d_work, 0, 16 * sizeof(int), st));
  const bool det = ctx->cfg.deterministic != 0;
  if (det) {
    if (!ctx->d_stats_fx) {
      CK(cudaMalloc(&ctx->d_stats_fx, (size_t)ctx->V * K * 3 * sizeof(unsigned long long)));
      CK(cudaMalloc(&ctx->d_alpha_ss_fx, (size_t)K * 3 * sizeof(unsigned long long)));
    }
    if (ctx->cfg.learning)
      CK(cudaMemsetAsync(ctx->d_stats_fx, 0, (size_t)ctx->V * K * 3 * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(ctx->d_alpha_ss_fx, 0, (size_t)K * 3 * sizeof(unsigned long long), st));
  }

  EstepArgs a;
  a.row_ptr = ctx->d_row_ptr;
  a.term_id = ctx->d_term;
  a.count = ctx->d_count;
  a.doc_order = ctx->d_order;
  a.doc_tokens = ctx->d_tokens;
  a.n_docs = ctx->D;
  a.E = ctx->d_E;
  a.e_stride = ctx->ES;
  a.rowmax = ctx->d_rowmax;
  a.logbeta = ctx->d_logbeta;
  a.alpha = ctx->d_alpha;
  a.lik_alpha = ctx->d_lik_alpha;
  a.gamma = ctx->d_gamma;
  a.stats = ctx->cfg.learning ? ctx->d_stats : nullptr;
  a.alpha_ss = ctx->d_alpha_ss;
  a.ll_counter = reinterpret_cast<unsigned long long *>(ctx->d_counters);
  a.slow_rows = ctx->d_work + 1;
  a.phase_cycles = ctx->d_phase;
  a.tmap_E = ctx->d_tmap_E;
  a.stats_fx = det ? ctx->d_stats_fx : nullptr;
  a.alpha_ss_fx = det ? ctx->d_alpha_ss_fx : nullptr;
  a.tmem_slots = 0;
  a.group_warps = 0;
  a.smem_batches = 0;
  a.K = K;
  a.sweeps = ctx->cfg.sweep_cap - 1;
  a.use_stored_gamma = (ctx->gamma_valid && !ctx->cfg.random_start) ? 1 : 0;
  a.learning = ctx->cfg.learning ? 1 : 0;

  // Documents are sorted by distinct-term count, longest first.  Those that fit one SM's registers +
  // shared memory go to the register-stationary kernel; the longer head of the order goes to the
  // general kernel, which streams the rows beyond its shared-memory tile from L2.
  int n_long = (int)ctx->D;
  int cap2 = 0, grid2 = 0;
  size_t smem2 = 0;
  enum { MAXCLS2 = 2 };
  struct Class2 {
    const Estep2Variant *v;
    int rt, cap, grid, begin, end;
    size_t smem;
  } cls2[MAXCLS2];
  int ncls2 = 0;
  const char *fk = getenv("MRLDA_ESTEP_KERNEL");  // "1": general kernel only (test / profiling aid)
  // the deterministic statistics mode exists in the general kernel only
  const bool use2 = ctx->variant2 && !(fk && atoi(fk) == 1) && !det;
  const bool use2f = use2 && ctx->variant2f != nullptr;
  if (use2f && ctx->D > 0) {
    // optional fp32 mode: registers then Tensor Memory only; longer documents take the fp64 general kernel
    const Estep2fVariant *vf = ctx->variant2f;
    const int c2 = 8 / vf->NW;
    const size_t per = std::min(((size_t)233472 - (size_t)c2 * 1024) / c2, ctx->smem_optin) & ~(size_t)15;
    const int need_slots = (ctx->max_rows + 31) / 32;
    int rt = std::max(0, std::min(estep2f_tmem_slots(vf->NW, vf->CPW), need_slots - vf->RREG));
    cap2 = 32 * (vf->RREG + rt);
    a.tmem_slots = rt;
    if (vf->smem_bytes(cap2) <= per) {
      n_long = (int)(std::upper_bound(ctx->h_nnz_sorted.begin(), ctx->h_nnz_sorted.end(), cap2,
                                      [](int cap, int32_t nn) { return cap >= nn; }) -
                     ctx->h_nnz_sorted.begin());
      grid2 = (int)std::min<int64_t>((int64_t)c2 * ctx->num_sms, ctx->D - n_long);
      const size_t min_dyn = (size_t)233472 / (c2 + 1) - 1024 + 16;
      smem2 = std::max(vf->smem_bytes(cap2), std::min(min_dyn, per));
      if (grid2 > 0) CK(vf->prepare(smem2));
    }
  } else if (use2 && ctx->D > 0) {
    // classes, longest documents first: the wide cluster variant (cluster variants only), then the
    // base variant; each takes the documents it holds that the next one does not
    const char *fcls = getenv("MRLDA_ESTEP2_CLASSES");  // "0": base variant only (profiling aid)
    const bool wide = !(fcls && atoi(fcls) == 0);
    const Estep2Variant *cand[MAXCLS2] = {wide ? ctx->variant2_wide : nullptr, ctx->variant2};
    const char *et = getenv("MRLDA_ESTEP2_TMEM");  // "0" disables TMEM slots (profiling aid)
    for (int k = 0; k < MAXCLS2; k++) {
      const Estep2Variant *v2 = cand[k];
      if (!v2) continue;
      Class2 &c = cls2[ncls2];
      c.v = v2;
      const int c2 = v2->WPS / v2->NW;
      const size_t per = std::min(((size_t)233472 - (size_t)c2 * 1024) / c2, ctx->smem_optin) & ~(size_t)15;
      // slot hierarchy per thread: RREG in registers, then Tensor Memory, then shared memory
      int rt = (et && atoi(et) == 0) ? 0 : estep2_tmem_slots(v2->NW, v2->CPW, v2->WPS);
      int rsm = 0;
      while (v2->smem_bytes(rsm + 1, 32 * (v2->RREG + rt + rsm + 1)) <= per) rsm++;
      if (v2->smem_bytes(rsm, 32 * (v2->RREG + rt + rsm)) > per) continue;
      const int capc = 32 * (v2->RREG + rt + rsm);
      // first document (descending nnz order) that fits this class
      c.begin = (int)(std::upper_bound(ctx->h_nnz_sorted.begin(), ctx->h_nnz_sorted.end(), capc,
                                       [](int cap, int32_t nn) { return cap >= nn; }) -
                      ctx->h_nnz_sorted.begin());
      if (ncls2 > 0) c.begin = std::max(c.begin, cls2[ncls2 - 1].begin);
      ncls2++;
    }
    for (int k = 0; k < ncls2; k++) {
      Class2 &c = cls2[k];
      const Estep2Variant *v2 = c.v;
      c.end = k + 1 < ncls2 ? cls2[k + 1].begin : (int)ctx->D;
      const int c2 = v2->WPS / v2->NW;
      const size_t per = std::min(((size_t)233472 - (size_t)c2 * 1024) / c2, ctx->smem_optin) & ~(size_t)15;
      // do not reserve more than the class's longest document needs
      const int need_slots = c.end > c.begin ? (ctx->h_nnz_sorted[c.begin] + 31) / 32 : 1;
      int rt = (et && atoi(et) == 0) ? 0 : estep2_tmem_slots(v2->NW, v2->CPW, v2->WPS);
      rt = std::max(0, std::min(rt, need_slots - v2->RREG));
      int rsm = 0;
      while (v2->smem_bytes(rsm + 1, 32 * (v2->RREG + rt + rsm + 1)) <= per) rsm++;
      rsm = std::max(0, std::min(rsm, need_slots - v2->RREG - rt));
      c.rt = rt;
      c.cap = 32 * (v2->RREG + rt + rsm);
      // a cluster variant runs CL CTAs per document; the grid is a whole number of clusters
      c.grid = v2->CL * (int)std::min<int64_t>((int64_t)c2 * ctx->num_sms / v2->CL, c.end - c.begin);
      // never let more than c2 CTAs share an SM: they would over-subscribe its 512 TMEM columns
      // and the extra CTA's tcgen05.alloc would wait forever
      const size_t min_dyn = (size_t)233472 / (c2 + 1) - 1024 + 16;
      c.smem = std::max(v2->smem_bytes(rsm, c.cap), std::min(min_dyn, per));
      if (c.grid > 0) CK(v2->prepare(c.smem));
    }
    if (ncls2 > 0) {
      n_long = cls2[0].begin;
      for (int k = 0; k < ncls2; k++) grid2 += cls2[k].grid;
    }
  }
  // Warp-group kernel (estep3): documents are classed by the number of warps they need; class G
  // (G warps per document, 8/G documents in flight per SM) takes the documents of at most
  // G * (R + T + M_G) * RPB rows that do not fit class G/2.  One launch per class, longest first.
  const Estep3Variant *v3 = (use2f || !use2) ? nullptr
                                             : (ctx->variant3 ? ctx->variant3
                                                              : (ctx->variant3c ? &ctx->variant3c->v : nullptr));
  const int cl3 = (v3 && !ctx->variant3) ? ctx->variant3c->CL : 1;  // CTAs per document class launch unit
  enum { MAXCLS = 6 };
  int ncls = 0, cls_G[MAXCLS], cls_begin[MAXCLS + 1], cls_M[MAXCLS];
  size_t cls_smem[MAXCLS];
  if (v3 && ctx->D > 0) {
    const size_t per = std::min((size_t)233472 - 1024, ctx->smem_optin) & ~(size_t)15;
    const int rpb = 32 / v3->L;
    const char *fg = getenv("MRLDA_ESTEP3_GROUPS");  // bit mask over the group sizes, largest first (profiling aid)
    const int gmask = fg ? atoi(fg) : ~0;
    // group sizes: the divisors of NW, largest first (named barriers 1..15: at most 15 groups)
    for (int G = v3->NW, bit = 0; G >= 1 && ncls < MAXCLS; G--) {
      if (v3->NW % G != 0 || v3->NW / G > 15) continue;
      if (ncls == 0 || ((gmask >> bit) & 1)) cls_G[ncls++] = G;
      bit++;
    }
    bool ok3 = true;
    for (int ci = 0; ci < ncls; ci++) {
      const int G = cls_G[ci];
      int M = v3->MMAX;
      while (M > 0 && v3->smem_bytes(M, G) > per) M--;
      if (v3->smem_bytes(M, G) > per) ok3 = false;
      cls_M[ci] = M;
      cls_smem[ci] = v3->smem_bytes(M, G);
      const int capG = G * (v3->R + v3->T + M) * rpb;
      // first document (in descending nnz order) that fits class ci
      cls_begin[ci] = (int)(std::upper_bound(ctx->h_nnz_sorted.begin(), ctx->h_nnz_sorted.end(), capG,
                                             [](int cap, int32_t nn) { return cap >= nn; }) -
                            ctx->h_nnz_sorted.begin());
      if (ci > 0) cls_begin[ci] = std::max(cls_begin[ci], cls_begin[ci - 1]);
    }
    cls_begin[ncls] = (int)ctx->D;
    if (ok3 && cl3 == 1) {
      n_long = cls_begin[0];
      grid2 = 0;
      ncls2 = 0;
    } else if (ok3) {
      // cluster warp-group kernel: the documents it does not hold stay with the cluster classes of
      // the register-stationary kernel (those that hold more rows) and the general kernel
      grid2 = 0;
      for (int k = 0; k < ncls2; k++) {
        Class2 &c2 = cls2[k];
        c2.begin = std::min(c2.begin, cls_begin[0]);
        c2.end = std::min(c2.end, cls_begin[0]);
        const int per_sm = c2.v->WPS / c2.v->NW;
        c2.grid = c2.v->CL * (int)std::min<int64_t>((int64_t)per_sm * ctx->num_sms / c2.v->CL, c2.end - c2.begin);
        grid2 += c2.grid;
      }
      n_long = ncls2 > 0 ? cls2[0].begin : cls_begin[0];
    } else {
      v3 = nullptr;
    }
  }
  int c = 1, t = 32, cap = 2;
  size_t smem = 0;
  if (n_long > 0) {
    plan_estep(ctx, &c, &t, &cap, &smem);
    CK(det ? ctx->variant->prepare_det(smem) : ctx->variant->prepare(smem));
  }
  CK(cudaEventRecord(ctx->ev[1], st));
  if (v3) {
    // the general kernel (documents longer than one SM holds) runs beside the class launches
    const bool fork = n_long > 0 && n_long < (int)ctx->D && !getenv("MRLDA_ESTEP_SERIAL");
    if (n_long > 0) {
      cudaStream_t sp = fork ? ctx->stream2 : st;
      if (fork) CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev[1], 0));
      a.doc_begin = 0;
      a.doc_end = n_long;
      a.work_counter = ctx->d_work;
      a.cap_rows = cap;
      int grid = (int)std::min<int64_t>((int64_t)c * ctx->num_sms, n_long);
      CK(ctx->variant->launch(a, grid, t, cap, sp));
      ctx->launches++;
      if (fork) CK(cudaEventRecord(ctx->ev[7], ctx->stream2));
    }
    // cluster kernel: the register-stationary cluster classes that hold longer documents go first
    for (int k = 0; k < ncls2 && cl3 > 1; k++) {
      const Class2 &cl = cls2[k];
      if (cl.grid <= 0) continue;
      a.doc_begin = cl.begin;
      a.doc_end = cl.end;
      a.work_counter = ctx->d_work + 10 + k;
      a.tmem_slots = cl.rt;
      CK(cl.v->launch(a, cl.grid, cl.cap, cl.smem, st));
      ctx->launches++;
    }
    for (int ci = 0; ci < ncls; ci++) {
      const int nd = cls_begin[ci + 1] - cls_begin[ci];
      if (nd <= 0) continue;
      const int G = cls_G[ci], per_cta = v3->NW / G;
      a.doc_begin = cls_begin[ci];
      a.doc_end = cls_begin[ci + 1];
      a.work_counter = ctx->d_work + 4 + ci;
      CK(v3->prepare(cls_smem[ci], G));
      // launch units: CTAs, or clusters of cl3 CTAs (cluster kernel)
      const int grid = (int)std::min<int64_t>(ctx->num_sms / cl3, ((int64_t)nd + per_cta - 1) / per_cta);
      CK(v3->launch(a, grid, G, cls_M[ci], cls_smem[ci], st));
      ctx->launches++;
    }
    if (fork) CK(cudaStreamWaitEvent(st, ctx->ev[7], 0));
  }
  // The two kernels work on disjoint documents and only meet in atomics, so they run on two
  // streams.  A cluster launch cannot use every SM (132 of 148 at cluster size 4): it goes first and
  // the general kernel's persistent CTAs fill the SMs it leaves free, then the rest of the GPU once
  // it drains.  Otherwise the general kernel (the longest documents) goes first and the
  // register-stationary CTAs take over each SM as it runs out of long documents.
  const bool overlap = !v3 && n_long > 0 && grid2 > 0 && !getenv("MRLDA_ESTEP_SERIAL");
  const bool cluster_first = overlap && !use2f && ctx->variant2->CL > 1;
  // the kernel that should start first stays on the main stream (nothing to wait for); the other
  // one goes to the second stream behind the fork event
  if (overlap) CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev[1], 0));
  for (int pass = 0; pass < 2 && !v3; pass++) {
    const bool long_pass = (pass == 0) != cluster_first;
    cudaStream_t sp = (pass == 1 && overlap) ? ctx->stream2 : st;
    if (long_pass && n_long > 0) {
      a.doc_begin = 0;
      a.doc_end = n_long;
      a.work_counter = ctx->d_work;
      a.cap_rows = cap;
      int grid = (int)std::min<int64_t>((int64_t)c * ctx->num_sms, n_long);
      CK(det ? ctx->variant->launch_det(a, grid, t, cap, sp) : ctx->variant->launch(a, grid, t, cap, sp));
      ctx->launches++;
    }
    if (!long_pass && grid2 > 0 && use2f) {
      a.doc_begin = n_long;
      a.doc_end = (int)ctx->D;
      a.work_counter = ctx->d_work + 2;
      CK(ctx->variant2f->launch(a, grid2, cap2, smem2, sp));
      ctx->launches++;
    }
    for (int k = 0; k < ncls2 && !long_pass && !use2f; k++) {
      const Class2 &cl = cls2[k];
      if (cl.grid <= 0) continue;
      a.doc_begin = cl.begin;
      a.doc_end = cl.end;
      a.work_counter = ctx->d_work + 10 + k;
      a.tmem_slots = cl.rt;
      cudaError_t e2 = cl.v->launch(a, cl.grid, cl.cap, cl.smem, sp);
      if (e2 == cudaErrorLaunchOutOfResources && cl.v->CL > 1) {
        // no cluster of this size can be scheduled on this device: the general kernel (still CUDA,
        // never the host) takes these documents as well
        cudaGetLastError();
        plan_estep(ctx, &c, &t, &cap, &smem);
        CK(ctx->variant->prepare(smem));
        a.cap_rows = cap;
        a.work_counter = ctx->d_work + 13 + k;
        int grid = (int)std::min<int64_t>((int64_t)c * ctx->num_sms, cl.end - cl.begin);
        CK(ctx->variant->launch(a, grid, t, cap, sp));
      } else {
        CK(e2);
      }
      ctx->launches++;
    }
  }
  if (overlap) {
    CK(cudaEventRecord(ctx->ev[7], ctx->stream2));
    CK(cudaStreamWaitEvent(st, ctx->ev[7], 0));
  }
  CK(cudaEventRecord(ctx->ev[2], st));
  int rc = exchange(ctx);
  if (rc) return rc;
  if (det) {  // fixed-point limbs (summed over the ranks) -> the fp64 statistics K2 / K3 read
    if (ctx->cfg.learning) {
      const int64_t n = (int64_t)ctx->V * K;
      fx_to_double_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(
          ctx->d_stats_fx, ctx->d_stats, n);
      ctx->launches++;
    }
    fx_to_double_kernel<<<1, 256, 0, st>>>(ctx->d_alpha_ss_fx, ctx->d_alpha_ss, K);
    ctx->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(ctx->ev[3], st));
  long long hres[4];
  int hwork[4];
  CK(cudaMemcpyAsync(hres, ctx->d_counters, sizeof hres, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hwork, ctx->d_work, sizeof hwork, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  // the reference keeps the gamma it wrote as the next iteration's input unless -randomstart in
  // training mode suppresses the gamma output (DocumentMapper.java:342, VI.java:232,359)
  ctx->gamma_valid = true;
  ctx->total_docs = hres[1];
  ctx->total_tokens = hres[2];
  if (hres[1] <= 0) {
    set_err(ctx, "mrlda_e_step: the corpus has no documents on any rank");
    return MRLDA_ESTATE;
  }
  if (out) {
    float ms = 0.f;
    memset(out, 0, sizeof *out);
    out->ll_counter = hres[0];
    out->log_likelihood = -(double)hres[0] * 1.0 / 1e6;  // VariationalInference.java:261-262
    out->n_docs = hres[1];
    out->n_tokens = hres[2];
    cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
    out->e_step_ms = ms;
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
    out->allreduce_ms = ctx->cfg.world_size > 1 ? ms : 0.0;
    out->estep_launches = (int)ctx->launches - launches0;
    out->total_launches = out->estep_launches;
    out->slow_path_rows = hwork[1];
  }
  return MRLDA_OK;
}

int mrlda_m_step(mrlda_ctx *ctx, mrlda_iter_result *out) {
  if (!ctx) return MRLDA_EINVAL;
  if (!ctx->cfg.learning) {
    set_err(ctx, "mrlda_m_step: context is in test mode (learning == 0)");
    return MRLDA_ESTATE;
  }
  CK(cudaSetDevice(ctx->cfg.device));
  cudaStream_t st = ctx->stream;
  const int K = ctx->K, V = ctx->V;
  int launches0 = (int)ctx->launches;
  CK(cudaEventRecord(ctx->ev[4], st));
  EtaSpec eta;
  eta.eta = ctx->cfg.eta;
  // InformedPrior.java:43-44: the log etas are floats; the reducer exponentiates the promoted value
  eta.hi = exp((double)(float)log(1000.0));
  eta.lo = exp((double)(float)log(0.001));
  eta.seed_bits = ctx->d_seed_bits;
  finalize_colsum_kernel<<<ctx->nslabs, 256, ctx->slab * sizeof(double), st>>>(
      ctx->d_stats, V, K, eta, ctx->slab, ctx->d_partial, ctx->d_present);
  finalize_norm_kernel<<<(K + 127) / 128, 128, 0, st>>>(ctx->d_partial, ctx->nslabs, K, ctx->d_nrm,
                                                        ctx->d_colsum);
  {
    int threads = 256;
    int64_t blocks = ((int64_t)V * 32 + threads - 1) / threads;
    finalize_rows_kernel<<<(unsigned)blocks, threads, 0, st>>>(ctx->d_stats, ctx->d_present,
                                                                ctx->d_nrm, eta, V, K,
                                                                ctx->ES, ctx->d_dl, ctx->d_logbeta,
                                                                ctx->d_E, ctx->d_rowmax);
  }
  ctx->launches += 3;
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev[5], st));
  if (ctx->cfg.update_alpha) {
    if (ctx->total_docs > 0x7fffffff) {
      set_err(ctx, "mrlda_m_step: document count exceeds int (VariationalInference.java:265 casts to int)");
      return MRLDA_EINVAL;
    }
    if (ctx->cfg.symmetric_alpha)
      alpha_scalar_kernel<<<1, 256, 0, st>>>(ctx->d_alpha, ctx->d_alpha_ss, K, (int)ctx->total_docs);
    else
      alpha_vector_kernel<<<1, 256, 0, st>>>(ctx->d_alpha, ctx->d_alpha_ss, K, (int)ctx->total_docs,
                                             ctx->d_alpha_work);
    ctx->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(ctx->ev[6], st));
  std::vector<uint8_t> pres((size_t)V);
  CK(cudaMemcpyAsync(pres.data(), ctx->d_present, (size_t)V, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  int64_t seen = 0;
  for (int v = 0; v < V; v++) seen += pres[v];
  ctx->terms_seen = seen;
  ctx->have_dl = true;
  if (out) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
    out->m_step_ms = ms;
    cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]);
    out->alpha_ms = ms;
    out->n_terms_seen = seen;  // TOTAL_TERMS / K (VariationalInference.java:267)
    out->total_launches += (int)ctx->launches - launches0;
  }
  return MRLDA_OK;
}

int mrlda_iterate(mrlda_ctx *ctx, mrlda_iter_result *out) {
  mrlda_iter_result tmp;
  mrlda_iter_result *r = out ? out : &tmp;
  int rc = mrlda_e_step(ctx, r);
  if (rc) return rc;
  if (ctx->cfg.learning) rc = mrlda_m_step(ctx, r);
  return rc;
}

int mrlda_update_vector_alpha(mrlda_ctx *ctx, int32_t n_topics, int32_t n_docs,
                              const double *alpha_in, const double *alpha_ss, double *alpha_out) {
  if (!ctx || !alpha_in || !alpha_ss || !alpha_out || n_topics <= 0) return MRLDA_EINVAL;
  CK(cudaSetDevice(ctx->cfg.device));
  double *buf = nullptr;
  const size_t K = (size_t)n_topics;
  CK(cudaMalloc(&buf, 5 * K * sizeof(double)));
  cudaMemcpyAsync(buf, alpha_in, K * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(buf + K, alpha_ss, K * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  alpha_vector_kernel<<<1, 256, 0, ctx->stream>>>(buf, buf + K, n_topics, n_docs, buf + 2 * K);
  ctx->launches++;
  cudaMemcpyAsync(alpha_out, buf, K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(buf);
  CK(e);
  return MRLDA_OK;
}

int mrlda_update_scalar_alpha(mrlda_ctx *ctx, int32_t n_topics, int32_t n_docs, double alpha_init,
                              double alpha_ss_total, double *alpha_out) {
  if (!ctx || !alpha_out || n_topics <= 0) return MRLDA_EINVAL;
  CK(cudaSetDevice(ctx->cfg.device));
  double *buf = nullptr;
  CK(cudaMalloc(&buf, sizeof(double)));
  alpha_scalar_direct_kernel<<<1, 1, 0, ctx->stream>>>(n_topics, n_docs, alpha_init, alpha_ss_total, buf);
  ctx->launches++;
  cudaMemcpyAsync(alpha_out, buf, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(buf);
  CK(e);
  return MRLDA_OK;
}

int mrlda_eval_special(mrlda_ctx *ctx, int32_t which, int64_t n, const double *x, double *y) {
  if (!ctx || !x || !y || n < 0 || which < 0 || which > 4) return MRLDA_EINVAL;
  if (n == 0) return MRLDA_OK;
  CK(cudaSetDevice(ctx->cfg.device));
  double *buf = nullptr;
  CK(cudaMalloc(&buf, 2 * (size_t)n * sizeof(double)));
  cudaMemcpyAsync(buf, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  eval_special_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(which, n, buf, buf + n);
  ctx->launches++;
  cudaMemcpyAsync(y, buf + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(buf);
  CK(e);
  return MRLDA_OK;
}

int mrlda_measure_fp64_peak(mrlda_ctx *ctx, double *tflops) {
  if (!ctx || !tflops) return MRLDA_EINVAL;
  CK(cudaSetDevice(ctx->cfg.device));
  double *buf = nullptr;
  CK(cudaMalloc(&buf, sizeof(double)));
  const int iters = 1 << 16, threads = 256, grid = ctx->num_sms * 8;
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    fp64_peak_kernel<<<grid, threads, 0, ctx->stream>>>(buf, iters, 1.0 + rep);
    ctx->launches++;
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    double fl = 2.0 * 8.0 * (double)iters * threads * (double)grid;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaFree(buf);
  *tflops = best;
  return MRLDA_OK;
}

int mrlda_debug_phase_cycles(mrlda_ctx *ctx, int64_t out[16]) {
  if (!ctx || !out) return MRLDA_EINVAL;
  CK(cudaSetDevice(ctx->cfg.device));
  CK(cudaMemcpy(out, ctx->d_phase, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  CK(cudaMemset(ctx->d_phase, 0, 16 * sizeof(long long)));
  return MRLDA_OK;
}

int mrlda_nccl_unique_id(uint8_t id[MRLDA_NCCL_ID_BYTES]) {
  if (!id) return MRLDA_EINVAL;
  if (!g_nccl.load()) {
    set_err(nullptr, "%s", g_nccl.err.c_str());
    return MRLDA_ENCCL;
  }
  ncclUniqueId_t u;
  int rc = g_nccl.GetUniqueId(&u);
  if (rc) {
    set_err(nullptr, "ncclGetUniqueId: %s", g_nccl.GetErrorString(rc));
    return MRLDA_ENCCL;
  }
  memcpy(id, u.internal, MRLDA_NCCL_ID_BYTES);
  return MRLDA_OK;
}

int mrlda_comm_init(mrlda_ctx *ctx, const uint8_t id[MRLDA_NCCL_ID_BYTES]) {
  if (!ctx || !id) return MRLDA_EINVAL;
  if (ctx->cfg.world_size <= 1) return MRLDA_OK;
  if (!g_nccl.load()) {
    set_err(ctx, "%s", g_nccl.err.c_str());
    return MRLDA_ENCCL;
  }
  CK(cudaSetDevice(ctx->cfg.device));
  ncclUniqueId_t u;
  memcpy(u.internal, id, MRLDA_NCCL_ID_BYTES);
  int rc = g_nccl.CommInitRank(&ctx->comm, ctx->cfg.world_size, u, ctx->cfg.rank);
  if (rc) {
    set_err(ctx, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc));
    ctx->comm = nullptr;
    return MRLDA_ENCCL;
  }
  return MRLDA_OK;
}

}  // extern "C"
