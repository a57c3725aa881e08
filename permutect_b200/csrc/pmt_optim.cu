// Flat optimiser step: misc_utils.backpropagate's clip_grad_norm_(max_norm) + AdamW.step (misc_utils.py:125-129,
// model_training.py:68-72) over ONE flat fp32 parameter buffer, two launches, no host synchronisation.
//   1. grad_sqnorm_kernel: fixed-order partial sums of g^2 (one per CTA)
//   2. adamw_kernel: every CTA re-adds the partials in the same order (total norm), derives the clip coefficient
//      min(1, max_norm / (norm + 1e-6)) (torch.nn.utils.clip_grad_norm_) and applies decoupled weight decay, moment
//      updates and the bias-corrected step exactly as torch.optim.AdamW does (amsgrad off, maximize off).
// Entries whose mask is 0 (parameters without a gradient, e.g. frozen during a calibration epoch) are left untouched,
// as torch skips parameters whose .grad is None.
#include "pmt_host.h"

namespace pmt {
namespace optim {

constexpr int NORM_CTAS = 64;
constexpr int NT = 256;

__global__ void __launch_bounds__(NT) grad_sqnorm_kernel(const float* __restrict__ g, const float* __restrict__ mask, long long n,
                                                          float* __restrict__ partials) {
  __shared__ float red[NT];
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long lo = (long long)blockIdx.x * per, hi = lo + per < n ? lo + per : n;
  float s = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += NT) {
    const float v = (mask && mask[i] == 0.f) ? 0.f : g[i];
    s = fmaf(v, v, s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = NT / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
}

struct AdamArgs {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  const float* mask;
  const float* partials;
  float* total_norm_out;
  long long n;
  int* step_count;   // per-entry number of updates taken (torch keeps one step counter per parameter)
  float lr, beta1, beta2, eps, weight_decay, max_norm;
};

__global__ void __launch_bounds__(NT) adamw_kernel(const AdamArgs A) {
  float total = 0.f;
  for (int c = 0; c < NORM_CTAS; ++c) total += A.partials[c];     // same order in every CTA
  const float norm = sqrtf(total);
  float coef = 1.f;
  if (A.max_norm > 0.f) coef = fminf(A.max_norm / (norm + 1e-6f), 1.f);
  if (blockIdx.x == 0 && threadIdx.x == 0 && A.total_norm_out) *A.total_norm_out = norm;
  const double lb1 = log((double)A.beta1), lb2 = log((double)A.beta2);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < A.n; i += (long long)gridDim.x * NT) {
    if (A.mask && A.mask[i] == 0.f) continue;
    const float g = A.grad[i] * coef;
    float p = A.param[i];
    p *= 1.f - A.lr * A.weight_decay;
    const float m = A.beta1 * A.exp_avg[i] + (1.f - A.beta1) * g;              // exp_avg.lerp_(grad, 1 - beta1)
    const float v = A.beta2 * A.exp_avg_sq[i] + (1.f - A.beta2) * g * g;
    A.exp_avg[i] = m;
    A.exp_avg_sq[i] = v;
    const int t = ++A.step_count[i];
    const float bias1 = (float)(-expm1((double)t * lb1));                 // 1 - beta1^t
    const float bias2_sqrt = (float)sqrt(-expm1((double)t * lb2));        // sqrt(1 - beta2^t)
    const float denom = sqrtf(v) / bias2_sqrt + A.eps;
    A.param[i] = p - (A.lr / bias1) * (m / denom);
  }
}

}  // namespace optim
}  // namespace pmt

using namespace pmt::optim;

extern "C" size_t pmt_adamw_workspace_size(void) { return NORM_CTAS * sizeof(float) + 256; }

extern "C" int pmt_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t* step_count,
                              const float* mask, int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm,
                              float* total_norm_out, void* workspace, size_t workspace_bytes, void* stream) {
  PMT_CHECK(params && grads && exp_avg && exp_avg_sq && step_count && n > 0, "pmt_adamw_step: bad arguments");
  PMT_CHECK(workspace && workspace_bytes >= pmt_adamw_workspace_size(), "pmt_adamw_step: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  grad_sqnorm_kernel<<<NORM_CTAS, NT, 0, st>>>(grads, mask, n, partials);
  AdamArgs A;
  A.param = params; A.grad = grads; A.exp_avg = exp_avg; A.exp_avg_sq = exp_avg_sq; A.mask = mask; A.partials = partials;
  A.total_norm_out = total_norm_out; A.n = n;
  A.lr = lr; A.beta1 = beta1; A.beta2 = beta2; A.eps = eps; A.weight_decay = weight_decay; A.max_norm = max_norm;
  A.step_count = step_count;
  int grid = (int)((n + NT - 1) / NT);
  if (grid > 148 * 8) grid = 148 * 8;
  adamw_kernel<<<grid, NT, 0, st>>>(A);
  cudaError_t e = cudaGetLastError();
  PMT_CHECK(e == cudaSuccess, "pmt_adamw_step launch failed: %s", cudaGetErrorString(e));
  return 0;
}
