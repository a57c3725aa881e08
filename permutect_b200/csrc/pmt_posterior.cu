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

struct OnlineD {   // streaming logsumexp that also carries sum_j softmax_j * q_j (q_j: a derivative attached to term j)
  float m = -INFINITY, s = 0.f, d = 0.f;
  __device__ __forceinline__ void add(float v, float q) {
    if (v > m) { const float sc = expf(m - v); s = s * sc + 1.f; d = d * sc + q; m = v; }
    else { const float e = expf(v - m); s += e; d += e * q; }
  }
  __device__ __forceinline__ float value() const { return m + logf(s); }
  __device__ __forceinline__ float mean_q() const { return d / s; }
};

// digamma for x > 0: recurrence up to x >= 6, then the asymptotic series
__device__ __forceinline__ float digammaf(float x) {
  float r = 0.f;
  while (x < 6.f) { r -= 1.f / x; x += 1.f; }
  const float i = 1.f / x, i2 = i * i;
  return r + logf(x) - 0.5f * i - i2 * (1.f / 12.f - i2 * (1.f / 120.f - i2 * (1.f / 252.f)));
}
// d beta_binomial / d alpha, d beta (stats_utils.py:29-41)
__device__ __forceinline__ void beta_binomial_grads(float n, float k, float a, float b, float& ga, float& gb) {
  const float common = digammaf(a + b) - digammaf(n + a + b);
  ga = digammaf(k + a) + common - digammaf(a);
  gb = digammaf(n - k + b) + common - digammaf(b);
}

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

constexpr int THREADS = 128, WARPS = THREADS / 32;
constexpr int MAXK = PMT_POSTERIOR_MAX_COMPONENTS;
// slots of the fitting pass, per CTA then summed over CTAs in order: loss, d cf_k, d log_weights_k, d artifact alpha / beta
// [3][5], d normal-artifact alpha / beta [3][5], d mean_multiplier_v, d concentration_v, E-step totals [type][call]
__host__ __device__ constexpr int slot_count(int K) { return 1 + 2 * K + 4 * N_DEPTH_BINS * N_TYPES + 2 * N_TYPES + N_TYPES * N_CALLS; }

struct FitArgs {
  float* partials;                 // [n_blocks][slot_count]
  float* somatic_snv_totals_rrra;  // [5][5][5][5] += posterior(SOMATIC) of SNVs (posterior_model_priors.py:111-119), may be null
  float* snv_context_totals_rrra;  // [5][5][5][5] += 1 per SNV, may be null
};

// Deterministic CTA sum of one slot: warp shuffle tree, lane 0 of each warp owns wsum[warp][slot]
__device__ __forceinline__ void slot_add(float* wsum, int n_slots, int slot, float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) wsum[(threadIdx.x >> 5) * n_slots + slot] += v;
}

template <bool FIT>
__global__ void __launch_bounds__(THREADS)
log_posteriors_kernel(const PmtPosteriorDesc D, const float* __restrict__ P, const int16_t* __restrict__ ints, long long int_stride,
                      const void* __restrict__ floats, int float_kind, long long float_stride, int n_variants,
                      PmtPosteriorOutputs out, FitArgs fit) {
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

  extern __shared__ float wsum[];   // FIT: [WARPS][slot_count]
  const int n_slots = slot_count(K);
  if (FIT) {
    for (int i = threadIdx.x; i < WARPS * n_slots; i += blockDim.x) wsum[i] = 0.f;
    __syncthreads();
  }
  const int v_raw = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = v_raw < n_variants;
  if (!FIT && !valid) return;
  const int v = valid ? v_raw : n_variants - 1;     // FIT: out-of-range threads take part in the reductions with weight 0
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
  float mix_term[FIT ? MAXK : 1], mix_dcf[FIT ? MAXK : 1], non_bg_share = 0.f;
  const float comb = comb_term(depth, alt), ncomb = comb_term(ndepth, nalt);
  {
    const float mafc = fminf(maf, 0.49f);                        // somatic_spectrum.py:78
    Online mix;
    for (int k = 0; k < K; ++k) {
      const float x1 = mafc * cf_k[k], x2 = (1.f - mafc) * cf_k[k];
      OnlineD u;
      for (int j = 0; j < N_INTERP; ++j) {
        const float t = 0.001f + 0.01f * (float)j;
        const float p = x2 * t + x1 * (1.f - t);
        // d/d cf_k of this term: (alt / p - (depth - alt) / (1 - p)) * p / cf_k
        u.add(comb + alt * logf(p) + (depth - alt) * logf(1.f - p), FIT ? (alt - (depth - alt) * p / (1.f - p)) / cf_k[k] : 0.f);
      }
      const float term = logw_k[k] + u.value() - logf((float)N_INTERP);
      mix.add(term);
      if (FIT) { mix_term[k] = term; mix_dcf[k] = u.mean_q(); }
    }
    const float mixv = mix.value();
    const float a = log_non_bg + mixv, b = log_bg + beta_binomial(depth, alt, comb, bg_alpha, bg_beta);
    const float m = fmaxf(a, b);
    spec[SOMATIC] = m + logf(expf(a - m) + expf(b - m));
    if (FIT) {
      non_bg_share = expf(a - spec[SOMATIC]);
      for (int k = 0; k < K; ++k) mix_term[k] = expf(mix_term[k] - mixv);     // responsibilities of the components
    }
  }
  const int db = depth_bin(depth), ndb = depth_bin(ndepth);
  spec[ARTIFACT] = beta_binomial(depth, alt, comb, art_alpha[db * N_TYPES + vt], art_beta[db * N_TYPES + vt]);
  const float na_normal = beta_binomial(ndepth, nalt, ncomb, na_alpha[ndb * N_TYPES + vt], na_beta[ndb * N_TYPES + vt]);
  const float conc = na_conc[vt], naf = nalt / (ndepth + 0.001f);
  const float a_b = 0.001f + naf * na_mult[vt] * conc;
  const float b_b = fmaxf(conc - a_b, 0.001f);
  spec[NORMAL_ARTIFACT] = beta_binomial(depth, alt, comb, a_b, b_b);
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
  if (FIT) {
    // ---- E step + gradient of -mean log evidence (posterior_model.py:139-151) ----
    float pm = post[0];
#pragma unroll
    for (int c = 1; c < N_CALLS; ++c) pm = fmaxf(pm, post[c]);
    float r[N_CALLS], rs = 0.f;
#pragma unroll
    for (int c = 0; c < N_CALLS; ++c) { r[c] = expf(post[c] - pm); rs += r[c]; }
    const float w = valid ? 1.f : 0.f;
#pragma unroll
    for (int c = 0; c < N_CALLS; ++c) r[c] = w * r[c] / rs;
    int slot = 0;
    slot_add(wsum, n_slots, slot++, w * (pm + logf(rs)));                                  // log evidence
    const float ws = r[SOMATIC] * non_bg_share;
    for (int k = 0; k < K; ++k) slot_add(wsum, n_slots, slot + k, ws * mix_term[k] * mix_dcf[k]);
    slot += K;
    for (int k = 0; k < K; ++k) slot_add(wsum, n_slots, slot + k, ws * mix_term[k]);
    slot += K;
    float ga, gb;
    beta_binomial_grads(depth, alt, art_alpha[db * N_TYPES + vt], art_beta[db * N_TYPES + vt], ga, gb);
    const float wa = logit < 0.f ? 0.f : r[ARTIFACT];                                       // the -9999 branch is a constant
    const int cell = db * N_TYPES + vt, ncell = ndb * N_TYPES + vt;
    for (int c = 0; c < N_DEPTH_BINS * N_TYPES; ++c) slot_add(wsum, n_slots, slot + c, c == cell ? wa * ga : 0.f);
    slot += N_DEPTH_BINS * N_TYPES;
    for (int c = 0; c < N_DEPTH_BINS * N_TYPES; ++c) slot_add(wsum, n_slots, slot + c, c == cell ? wa * gb : 0.f);
    slot += N_DEPTH_BINS * N_TYPES;
    const float wn = r[NORMAL_ARTIFACT], wn2 = nalt < 1.f ? 0.f : wn;
    beta_binomial_grads(ndepth, nalt, na_alpha[ncell], na_beta[ncell], ga, gb);
    for (int c = 0; c < N_DEPTH_BINS * N_TYPES; ++c) slot_add(wsum, n_slots, slot + c, c == ncell ? wn2 * ga : 0.f);
    slot += N_DEPTH_BINS * N_TYPES;
    for (int c = 0; c < N_DEPTH_BINS * N_TYPES; ++c) slot_add(wsum, n_slots, slot + c, c == ncell ? wn2 * gb : 0.f);
    slot += N_DEPTH_BINS * N_TYPES;
    beta_binomial_grads(depth, alt, a_b, b_b, ga, gb);
    const float unclamped = (conc - a_b) > 0.001f ? 1.f : 0.f, mult = na_mult[vt];
    const float d_mult = ga * naf * conc - gb * unclamped * naf * conc;
    const float d_conc = ga * naf * mult + gb * unclamped * (1.f - naf * mult);
    for (int t = 0; t < N_TYPES; ++t) slot_add(wsum, n_slots, slot + t, t == vt ? wn * d_mult : 0.f);
    slot += N_TYPES;
    for (int t = 0; t < N_TYPES; ++t) slot_add(wsum, n_slots, slot + t, t == vt ? wn * d_conc : 0.f);
    slot += N_TYPES;
    for (int t = 0; t < N_TYPES; ++t)
#pragma unroll
      for (int c = 0; c < N_CALLS; ++c) slot_add(wsum, n_slots, slot + t * N_CALLS + c, t == vt ? r[c] : 0.f);
    if (valid && vt == 0 && (fit.somatic_snv_totals_rrra || fit.snv_context_totals_rrra)) {
      const int L = D.hap_len, c = (L - 1) / 2;
      const int16_t* hap = ir + D.hap_start;
      const int i0 = min(max((int)hap[c - 1], 0), 4), i1 = min(max((int)hap[c], 0), 4), i2 = min(max((int)hap[c + 1], 0), 4),
                i3 = min(max((int)hap[c + L], 0), 4);
      const int at = ((i0 * 5 + i1) * 5 + i2) * 5 + i3;
      if (fit.somatic_snv_totals_rrra) atomicAdd(fit.somatic_snv_totals_rrra + at, r[SOMATIC]);
      if (fit.snv_context_totals_rrra) atomicAdd(fit.snv_context_totals_rrra + at, 1.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_slots; i += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int wq = 0; wq < WARPS; ++wq) t += wsum[wq * n_slots + i];
      fit.partials[(long long)blockIdx.x * n_slots + i] = t;
    }
    if (!valid) return;
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

// Sums the per-CTA slots in CTA order (bitwise reproducible) and applies -1/B to the loss and gradient slots.
__global__ void fit_reduce_kernel(const float* __restrict__ partials, int n_blocks, int n_slots, int n_grad_slots, float inv_b,
                                  float* __restrict__ loss_out, float* __restrict__ grads_out, float* __restrict__ totals_tc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  float s = 0.f;
  for (int b = 0; b < n_blocks; ++b) s += partials[(long long)b * n_slots + i];
  if (i == 0) { if (loss_out) *loss_out = -s * inv_b; }
  else if (i <= n_grad_slots) { if (grads_out) grads_out[i - 1] = -s * inv_b; }
  else if (totals_tc) totals_tc[i - 1 - n_grad_slots] += s;
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
  post::log_posteriors_kernel<false><<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      *desc, params, int_array, int_stride, float_array, float_kind, float_stride, n_variants, *out, post::FitArgs{});
  cudaError_t e = cudaGetLastError();
  PMT_CHECK(e == cudaSuccess, "pmt_posterior_log_posteriors launch failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" size_t pmt_posterior_fit_workspace_size(int32_t n_variants, int32_t n_components) {
  const size_t blocks = (size_t)(n_variants + post::THREADS - 1) / post::THREADS;
  return blocks * post::slot_count(n_components) * sizeof(float) + 256;
}

extern "C" int pmt_posterior_fit_step(const PmtPosteriorDesc* desc, const float* params, const int16_t* int_array,
                                      int64_t int_stride, const void* float_array, int32_t float_kind, int64_t float_stride,
                                      int32_t n_variants, float* loss_out, float* grads_out, float* posterior_totals_tc,
                                      float* somatic_snv_totals_rrra, float* snv_context_totals_rrra, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  PMT_CHECK(desc && params && int_array && float_array && workspace, "pmt_posterior_fit_step: null argument");
  PMT_CHECK(desc->n_components >= 1 && desc->n_components <= PMT_POSTERIOR_MAX_COMPONENTS, "somatic spectrum components %d outside 1..%d",
            desc->n_components, PMT_POSTERIOR_MAX_COMPONENTS);
  PMT_CHECK(float_kind == PMT_F16 || float_kind == PMT_F32, "float_kind must be PMT_F16 or PMT_F32");
  PMT_CHECK(n_variants > 0, "empty batch");
  PMT_CHECK(workspace_bytes >= pmt_posterior_fit_workspace_size(n_variants, desc->n_components), "workspace too small");
  PMT_CHECK((!somatic_snv_totals_rrra && !snv_context_totals_rrra) || desc->hap_len >= 3, "SNV context totals need haplotypes of >= 3 bases");
  PMT_CHECK(!desc->use_context_dependent_snv_priors || desc->hap_len >= 3, "context-dependent SNV priors need haplotypes of >= 3 bases");
  const int K = desc->n_components, n_slots = post::slot_count(K), blocks = (n_variants + post::THREADS - 1) / post::THREADS;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  post::FitArgs fit{reinterpret_cast<float*>(workspace), somatic_snv_totals_rrra, snv_context_totals_rrra};
  PmtPosteriorOutputs none{};
  const size_t smem = (size_t)post::WARPS * n_slots * sizeof(float);
  post::log_posteriors_kernel<true><<<blocks, post::THREADS, smem, st>>>(*desc, params, int_array, int_stride, float_array, float_kind,
                                                                         float_stride, n_variants, none, fit);
  const int n_grad = 2 * K + 4 * post::N_DEPTH_BINS * post::N_TYPES + 2 * post::N_TYPES;
  post::fit_reduce_kernel<<<(n_slots + 127) / 128, 128, 0, st>>>(fit.partials, blocks, n_slots, n_grad, 1.f / (float)n_variants, loss_out,
                                                                grads_out, posterior_totals_tc);
  cudaError_t e = cudaGetLastError();
  PMT_CHECK(e == cudaSuccess, "pmt_posterior_fit_step launch failed: %s", cudaGetErrorString(e));
  return 0;
}
