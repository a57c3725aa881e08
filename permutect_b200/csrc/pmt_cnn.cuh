// Register-tiled 1-D convolution over a group of variants held feature(channel)-major in shared memory,
// shared by the haplotype-CNN forward and backward kernels (dna_sequence_convolution.py:57-111).
#pragma once
#include "pmt_device.cuh"

namespace pmt {

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == PMT_ACT_SELU) return selu(x);
  if (act == PMT_ACT_LEAKY_RELU) return x > 0.f ? x : 0.01f * x;
  return x;
}

template <int KS>
__device__ __forceinline__ void conv_units(const PmtCnnOp& op, const float* __restrict__ in, int in_ld, int lp_in,
                                           float* __restrict__ out, int out_ld, int lp_out, const float* __restrict__ img,
                                           const float* __restrict__ wflat, int vt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (op.out_ch + 7) / 8;
  const int q_per_var = lp_out / 4;
  const int n_q = vt * q_per_var;  // 4-row groups
  const int q_blocks = (n_q + 31) / 32;
  for (int unit = warp; unit < q_blocks * G; unit += NWARPS) {
    const int g = unit % G, q = (unit / G) * 32 + lane;
    if (q >= n_q) continue;
    const int v = q / q_per_var, p0 = (q % q_per_var) * 4;
    float acc[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = g * 8 + j;
      const float b = (co < op.out_ch && op.b_off >= 0) ? __ldg(wflat + op.b_off + co) : 0.f;
      acc[0][j] = b; acc[1][j] = b; acc[2][j] = b; acc[3][j] = b;
    }
    const float* xp = in + v * lp_in + p0;
    const float* wp = img + g * GROUP_STRIDE;
    const int wstride = G * GROUP_STRIDE;
    for (int ci = 0; ci < op.in_ch; ++ci) {
      float xw[12];
      const float4 a = *reinterpret_cast<const float4*>(xp + ci * in_ld);
      xw[0] = a.x; xw[1] = a.y; xw[2] = a.z; xw[3] = a.w;
      if (KS > 1) {
        const float4 b = *reinterpret_cast<const float4*>(xp + ci * in_ld + 4);
        xw[4] = b.x; xw[5] = b.y; xw[6] = b.z; xw[7] = b.w;
      }
      if (KS > 5) {
        const float4 c = *reinterpret_cast<const float4*>(xp + ci * in_ld + 8);
        xw[8] = c.x; xw[9] = c.y; xw[10] = c.z; xw[11] = c.w;
      }
#pragma unroll
      for (int t = 0; t < KS; ++t) {
        const float* wr = wp + (ci * KS + t) * wstride;
        const float4 w0 = *reinterpret_cast<const float4*>(wr);
        const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] = fmaf(xw[t], w[j], acc[0][j]);
          acc[1][j] = fmaf(xw[t + 1], w[j], acc[1][j]);
          acc[2][j] = fmaf(xw[t + 2], w[j], acc[2][j]);
          acc[3][j] = fmaf(xw[t + 3], w[j], acc[3][j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = g * 8 + j;
      if (co < op.out_ch) {
        float4 o = make_float4(apply_act(acc[0][j], op.act), apply_act(acc[1][j], op.act),
                               apply_act(acc[2][j], op.act), apply_act(acc[3][j], op.act));
        *reinterpret_cast<float4*>(out + co * out_ld + v * lp_out + p0) = o;
      }
    }
  }
}


#define PMT_CONV_DISPATCH(KSV, ...)                   \
  switch (KSV) {                                      \
    case 1: conv_units<1>(__VA_ARGS__); break;        \
    case 2: conv_units<2>(__VA_ARGS__); break;        \
    case 3: conv_units<3>(__VA_ARGS__); break;        \
    case 4: conv_units<4>(__VA_ARGS__); break;        \
    case 5: conv_units<5>(__VA_ARGS__); break;        \
    case 6: conv_units<6>(__VA_ARGS__); break;        \
    case 7: conv_units<7>(__VA_ARGS__); break;        \
    case 8: conv_units<8>(__VA_ARGS__); break;        \
    default: conv_units<9>(__VA_ARGS__); break;       \
  }

}  // namespace pmt
