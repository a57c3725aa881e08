// PosteriorModel.log_posterior_and_ingredients as one kernel (reference posterior_model.py:69-99).
//
// One thread per variant: log priors (posterior_model_priors.py:121-139), the five spectra log-likelihoods
// (posterior_model_spectra.py:78-124: somatic_spectrum.py:74-98, artifact_spectra.py:45-55,
// normal_artifact_spectrum.py:38-58, germline :18-58), the matched-normal table, and their sum with the cached artifact
// logit.  The reference builds ~25 [B] or [B, K, 100] temporaries per call; here a variant's 5 x 100 binomial terms of
// the somatic mixture live in registers and nothing but the requested [B, 5] tables touches HBM.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/permutect_b200.h"
#include "pmt_host.h"

namespace post {

constexpr int N_CALLS = 5, SOMATIC = 0, ARTIFACT = 1, SEQ_ERROR = 2, GERMLINE = 3, NORMAL_ARTIFACT = 4;   // utils/enums.py Call
constexpr int N_TYPES = 5, N_DEPTH_BINS = 3;
constexpr int N_INTERP = 100;   // torch.arange(0.001, 0.999, 0.01), stats_utils.py:183

struct Online {   // streaming logsumexp
  float m = -INFINITY, s = 0.f;
  __device__ __forceinline__ void add(float v) {
    if (v > m) { s = s * expf(m - v) + 1.f; m = v; }
    else s += expf(v - m);
  }
  __device__ __forceinline__ float value() const { return m + logf(s); }
};

__device__ __forceinline__ float comb_term(float n, float k) { return lgammaf(n + 1.f) - lgammaf(n - k + 1.f) - lgammaf(k + 1.f); }
// stats_utils.py:29-41
__device__ __forceinline__ float beta_binomial(float n, float k, float comb, float a, float b) {
  return comb + lgammaf(k + a) + lgammaf(n - k + b) + lgammaf(a + b) - lgammaf(n + a + b) - lgammaf(a) - lgammaf(b);
}
__device__ __forceinline__ int depth_bin(float d) { return (d >= 10.f ? 1 : 0) + (d >= 20.f ? 1 : 0); }   // artifact_spectra.py:17-24

// posterior_model_spectra.py:18-58
__device__ __forceinline__ float germline(float af, float maf, float alt, float depth, float comb, float het_beta) {
  const float het = 2.f * af * (1.f - af), hom = af * af;
  const float het_prop = het / (het + hom), hom_prop = 1.f - het_prop;
  const float ref = depth - alt;
  float minor, major;
  if (het_beta < 0.f) {
    const float lm = logf(maf), l1m = logf(1.f - maf);
    minor = comb + alt * lm + ref * l1m;
    major = comb + ref * lm + alt * l1m;
  } else {
    minor = major = beta_binomial(depth, alt, comb, het_beta, het_beta);
  }
  const float half = logf(het_prop / 2.f);
  const float hom_ll = logf(hom_prop) + beta_binomial(depth, alt, comb, 98.f, 2.f);
  const float a = half + minor, b = half + major, c = hom_ll;
  const float m = fmaxf(a, fmaxf(b, c));
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

__device__ __forceinline__ float load_float(const void* p, int kind, long long i) {
  return kind == PMT_F16 ? __half2float(reinterpret_cast<const __half*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}

__global__ void __launch_bounds__(128)
log_posteriors_kernel(const PmtPosteriorDesc D, const float* __restrict__ P, const int16_t* __restrict__ ints, long long int_stride,
                      const void* __restrict__ floats, int float_kind, long long float_stride, int n_variants,
                      PmtPosteriorOutputs out) {
  __shared__ float sp[PMT_POSTERIOR_MAX_COMPONENTS * 2 + 4 + 4 * N_DEPTH_BINS * N_TYPES + 2 * N_TYPES + N_TYPES * N_CALLS];
  const int K = D.n_components;
  const int n_small = 2 * K + 4 + 4 * N_DEPTH_BINS * N_TYPES + 2 * N_TYPES + N_TYPES * N_CALLS;
  for (int i = threadIdx.x; i < n_small; i += blockDim.x) sp[i] = P[i];
  __syncthreads();
  const float* cf_k = sp;
  const float* logw_k = cf_k + K;
  const float log_bg = logw_k[K], log_non_bg = logw_k[K + 1], bg_alpha = logw_k[K + 2], bg_beta = logw_k[K + 3];
  const float* art_alpha = logw_k + K + 4;
  const float* art_beta = art_alpha + N_DEPTH_BINS * N_TYPES;
  const float* na_alpha = art_beta + N_DEPTH_BINS * N_TYPES;
  const float* na_beta = na_alpha + N_DEPTH_BINS * N_TYPES;
  const float* na_mult = na_beta + N_DEPTH_BINS * N_TYPES;
  const float* na_conc = na_mult + N_TYPES;
  const float* log_priors_vc = na_conc + N_TYPES;
  const float* snv_rrra = P + n_small;   // [5][5][5][5], global (L1 / L2)

  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_variants) return;
  const int16_t* ir = ints + (long long)v * int_stride;
  const long long fo = (long long)v * float_stride;
  int vt = ir[3];
  vt = vt < 0 ? 0 : (vt >= N_TYPES ? N_TYPES - 1 : vt);
  const float depth = (float)ir[5], alt = (float)ir[6], ndepth = (float)ir[7], nalt = (float)ir[8];
  const float seq_err = load_float(floats, float_kind, fo + 0), nseq_err = load_float(floats, float_kind, fo + 1);
  const float af = load_float(floats, float_kind, fo + 2), maf = load_float(floats, float_kind, fo + 3);
  const float nmaf = load_float(floats, float_kind, fo + 4), logit = load_float(floats, float_kind, fo + 5);

  // ---- priors ----
  float pri[N_CALLS];
#pragma unroll
  for (int c = 0; c < N_CALLS; ++c) pri[c] = log_priors_vc[vt * N_CALLS + c];
  pri[SEQ_ERROR] = 0.f;
  pri[GERMLINE] = D.no_germline_mode ? -9999.f : logf(1.f - (1.f - af) * (1.f - af));
  if (D.use_context_dependent_snv_priors && vt == 0) {
    const int L = D.hap_len, c = (L - 1) / 2;
    const int16_t* hap = ir + D.hap_start;
    const int i0 = min(max((int)hap[c - 1], 0), 4), i1 = min(max((int)hap[c], 0), 4), i2 = min(max((int)hap[c + 1], 0), 4),
              i3 = min(max((int)hap[c + L], 0), 4);
    pri[SOMATIC] = __ldg(snv_rrra + ((i0 * 5 + i1) * 5 + i2) * 5 + i3);
  }
  {
    float m = pri[0];
#pragma unroll
    for (int c = 1; c < N_CALLS; ++c) m = fmaxf(m, pri[c]);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < N_CALLS; ++c) s += expf(pri[c] - m);
    const float lse = m + logf(s);
#pragma unroll
    for (int c = 0; c < N_CALLS; ++c) pri[c] -= lse;
  }

  // ---- spectra ----
  float spec[N_CALLS], norm[N_CALLS];
  const float comb = comb_term(depth, alt), ncomb = comb_term(ndepth, nalt);
  {
    const float mafc = fminf(maf, 0.49f);                        // somatic_spectrum.py:78
    Online mix;
    for (int k = 0; k < K; ++k) {
      const float x1 = mafc * cf_k[k], x2 = (1.f - mafc) * cf_k[k];
      Online u;
      for (int j = 0; j < N_INTERP; ++j) {
        const float t = 0.001f + 0.01f * (float)j;
        const float p = x2 * t + x1 * (1.f - t);
        u.add(comb + alt * logf(p) + (depth - alt) * logf(1.f - p));
      }
      mix.add(logw_k[k] + u.value() - logf((float)N_INTERP));
    }
    const float a = log_non_bg + mix.value(), b = log_bg + beta_binomial(depth, alt, comb, bg_alpha, bg_beta);
    const float m = fmaxf(a, b);
    spec[SOMATIC] = m + logf(expf(a - m) + expf(b - m));
  }
  const int db = depth_bin(depth), ndb = depth_bin(ndepth);
  spec[ARTIFACT] = beta_binomial(depth, alt, comb, art_alpha[db * N_TYPES + vt], art_beta[db * N_TYPES + vt]);
  const float na_normal = beta_binomial(ndepth, nalt, ncomb, na_alpha[ndb * N_TYPES + vt], na_beta[ndb * N_TYPES + vt]);
  {
    const float conc = na_conc[vt];
    const float a_b = 0.001f + (nalt / (ndepth + 0.001f)) * na_mult[vt] * conc;
    const float b_b = fmaxf(conc - a_b, 0.001f);
    spec[NORMAL_ARTIFACT] = beta_binomial(depth, alt, comb, a_b, b_b);
  }
  spec[SEQ_ERROR] = seq_err;
  spec[GERMLINE] = germline(af, maf, alt, depth, comb, D.het_beta);

  norm[SOMATIC] = nseq_err; norm[ARTIFACT] = nseq_err; norm[SEQ_ERROR] = nseq_err;
  norm[NORMAL_ARTIFACT] = nalt < 1.f ? -9999.f : na_normal;
  norm[GERMLINE] = germline(af, nmaf, nalt, ndepth, ncomb, D.het_beta);

  float post[N_CALLS];
#pragma unroll
  for (int c = 0; c < N_CALLS; ++c) post[c] = pri[c] + spec[c] + norm[c];
  post[ARTIFACT] += logit;
  post[NORMAL_ARTIFACT] += logit;
  if (logit < 0.f) post[ARTIFACT] = -9999.f;                     // posterior_model.py:90-93

  const long long o = (long long)v * N_CALLS;
#pragma unroll
  for (int c = 0; c < N_CALLS; ++c) {
    if (out.log_priors_bc) out.log_priors_bc[o + c] = pri[c];
    if (out.spectra_log_lks_bc) out.spectra_log_lks_bc[o + c] = spec[c];
    if (out.normal_log_lks_bc) out.normal_log_lks_bc[o + c] = norm[c];
    if (out.log_posteriors_bc) out.log_posteriors_bc[o + c] = post[c];
  }
  if (out.posterior_probabilities_bc) {
    float m = post[0];
#pragma unroll
    for (int c = 1; c < N_CALLS; ++c) m = fmaxf(m, post[c]);
    float e[N_CALLS], s = 0.f;
#pragma unroll
    for (int c = 0; c < N_CALLS; ++c) { e[c] = expf(post[c] - m); s += e[c]; }
#pragma unroll
    for (int c = 0; c < N_CALLS; ++c) out.posterior_probabilities_bc[o + c] = e[c] / s;
  }
}

}  // namespace post

extern "C" int pmt_posterior_param_count(int32_t n_components) {
  return 2 * n_components + 4 + 4 * post::N_DEPTH_BINS * post::N_TYPES + 2 * post::N_TYPES + post::N_TYPES * post::N_CALLS + 625;
}

extern "C" int pmt_posterior_log_posteriors(const PmtPosteriorDesc* desc, const float* params, const int16_t* int_array,
                                            int64_t int_stride, const void* float_array, int32_t float_kind, int64_t float_stride,
                                            int32_t n_variants, const PmtPosteriorOutputs* out, void* stream) {
  PMT_CHECK(desc && params && int_array && float_array && out, "pmt_posterior_log_posteriors: null argument");
  PMT_CHECK(desc->n_components >= 1 && desc->n_components <= PMT_POSTERIOR_MAX_COMPONENTS, "somatic spectrum components %d outside 1..%d",
            desc->n_components, PMT_POSTERIOR_MAX_COMPONENTS);
  PMT_CHECK(float_kind == PMT_F16 || float_kind == PMT_F32, "float_kind must be PMT_F16 or PMT_F32");
  PMT_CHECK(!desc->use_context_dependent_snv_priors || desc->hap_len >= 3, "context-dependent SNV priors need haplotypes of >= 3 bases");
  if (n_variants <= 0) return 0;
  const int threads = 128, blocks = (n_variants + threads - 1) / threads;
  post::log_posteriors_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      *desc, params, int_array, int_stride, float_array, float_kind, float_stride, n_variants, *out);
  cudaError_t e = cudaGetLastError();
  PMT_CHECK(e == cudaSuccess, "pmt_posterior_log_posteriors launch failed: %s", cudaGetErrorString(e));
  return 0;
}
