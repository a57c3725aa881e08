// Tensor-core (tcgen05 / TMEM) forward of the haplotype CNN (batch.py:115-130 one-hot + dna_sequence_convolution.py:57-111).
//
// A 1-D convolution over G variants is one GEMM whose rows are the FLAT positions j = v * L_in + p of the group:
//     out[j][co] = sum_t sum_ci in[j + t][ci] * W[co][ci][t]
// i.e. ksize accumulating MMAs whose A operand is the SAME activation buffer shifted by t rows.  Activations live in
// shared memory position-major in the no-swizzle K-major canonical layout "[16-byte channel chunk][row][4 floats]"
// (8 planes of PLANE_ROWS rows): core matrices of 8 rows x 16 B are contiguous, 8-row groups are 128 B apart
// (SBO) and K-adjacent core matrices one plane apart (LBO), so a shift by t rows is the start address + 16 t bytes.
// Rows j whose window straddles two variants compute garbage that nobody reads.
//
//   warps 0-7  epilogue: thread = output row (TMEM lane) of a 128-row chunk; even chunks -> warps 0-3, odd -> warps 4-7.
//              bias, max-pool, SELU, hi/lo split, compaction to the next layer's row stride, IN PLACE (a compacted row
//              index never exceeds the source row index, and chunks complete in order).
//   warp  8    MMA issuer: per layer, all chunks back to back into separate TMEM accumulators, one commit per chunk.
//
// The first conv reads the one-hot input as an im2col row [onehot(p) | onehot(p+1) | ...] (exact in TF32, no lo pass);
// a MaxPool(kernel 2, stride 1) that follows it is fused by computing the conv at p and p+1 side by side in N.
// MaxPool(2, 2) elsewhere is a lane shuffle (row strides are even).  Flatten + Linear is a conv whose kernel covers
// the whole remaining length.  SELU scales are folded into the consuming weights.  Precision modes as pmt_tc.cu.
#include <cstring>

#include "pmt_host.h"
#include "pmt_tc_ptx.cuh"

namespace pmt {
namespace cnntc {

using namespace pmt::tc;

constexpr int EPI_WARPS = 8;
constexpr int THREADS = 32 * (EPI_WARPS + 1);
constexpr int MMA_WARP = EPI_WARPS;
constexpr int MAX_LAYERS = 10;
constexpr int MAX_CHUNKS = 4;
constexpr int CHUNK_COLS = 64;
constexpr int PLANE_ROWS = 320;
constexpr int PLANE_BYTES = PLANE_ROWS * 16;
constexpr int BUF_BYTES = 8 * PLANE_BYTES;   // one activation buffer: 32 channels
constexpr int C0 = 10;                       // one-hot channels: 2 haplotypes x 5 codes
constexpr int MAX_CODES = 2048;              // staged haplotype codes per group (G * 2L)

struct Layer {
  int first;      // im2col'd one-hot conv
  int taps;       // shifted-window convs: kernel size; first: number of input positions per row (ksize + dup)
  int ksteps;     // first: k-steps of 8 columns
  int N;          // MMA N
  int dup, pool2; // fused MaxPool(2,1) after the first conv / MaxPool(2,2)
  int L_in, L_out, L_pool, L_next;
  int act, to_global, out_ch;
  int img_off, img_bytes;
  // packing
  int op, in_ch, ksize, flat_len, scale_in, is_linear;
};

struct Plan {
  int n_layers, G, L0, image_bytes;
  int n_chunks[MAX_LAYERS];
  Layer layer[MAX_LAYERS];
};

__device__ __forceinline__ uint64_t desc_ns(unsigned addr, unsigned lbo_bytes) {
  // K-major, no swizzle: LBO = distance between K-adjacent core matrices, SBO = 128 B between 8-row groups
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(lbo_bytes >> 4) << 16) | (8ull << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ss(unsigned tmem_d, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ unsigned make_idesc(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// shifted-window conv: TAPS x 4 k-steps; all descriptor offsets are compile-time constants (16-byte units)
template <int TAPS, int N, int PASSES>
__device__ __forceinline__ void issue_shifted(unsigned d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, unsigned idesc) {
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t ao = (uint64_t)(t + 2 * ks * (PLANE_BYTES / 16));
      const uint64_t bo = (uint64_t)(t * 8 * N + 2 * ks * N);
      mma_ss(d, a_hi + ao, b_hi + bo, idesc, (t | ks) != 0 ? 1u : 0u);
      if (PASSES == 3) {
        mma_ss(d, a_lo + ao, b_hi + bo, idesc, 1u);
        mma_ss(d, a_hi + ao, b_lo + bo, idesc, 1u);
      }
    }
  }
}
template <int N, int PASSES>
__device__ __forceinline__ void issue_shifted_n(int taps, unsigned d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo) {
  const unsigned idesc = make_idesc(N);
  switch (taps) {
    case 1: issue_shifted<1, N, PASSES>(d, a_hi, a_lo, b_hi, b_lo, idesc); break;
    case 2: issue_shifted<2, N, PASSES>(d, a_hi, a_lo, b_hi, b_lo, idesc); break;
    case 3: issue_shifted<3, N, PASSES>(d, a_hi, a_lo, b_hi, b_lo, idesc); break;
    case 4: issue_shifted<4, N, PASSES>(d, a_hi, a_lo, b_hi, b_lo, idesc); break;
    case 5: issue_shifted<5, N, PASSES>(d, a_hi, a_lo, b_hi, b_lo, idesc); break;
    case 6: issue_shifted<6, N, PASSES>(d, a_hi, a_lo, b_hi, b_lo, idesc); break;
    case 7: issue_shifted<7, N, PASSES>(d, a_hi, a_lo, b_hi, b_lo, idesc); break;
    default: issue_shifted<8, N, PASSES>(d, a_hi, a_lo, b_hi, b_lo, idesc); break;
  }
}
// first conv: one-hot im2col rows (exact: no A lo pass), K = 8 * KS columns over consecutive planes
template <int N, int PASSES>
__device__ __forceinline__ void issue_first(int ksteps, unsigned d, uint64_t a, uint64_t b_hi, uint64_t b_lo) {
  const unsigned idesc = make_idesc(N);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    if (ks < ksteps) {
      const uint64_t ao = (uint64_t)(2 * ks * (PLANE_BYTES / 16)), bo = (uint64_t)(2 * ks * N);
      mma_ss(d, a + ao, b_hi + bo, idesc, ks != 0 ? 1u : 0u);
      if (PASSES == 3) mma_ss(d, a + ao, b_lo + bo, idesc, 1u);
    }
  }
}

struct Bars {
  unsigned long long in_bar, acc_bar[MAX_CHUNKS];
  unsigned tmem_base;
  int pad_;
};

template <int PASSES>
__global__ void __launch_bounds__(THREADS, 1)
hap_cnn_tc_kernel(const __grid_constant__ Plan TP, const unsigned char* __restrict__ image, const float* __restrict__ wflat,
                  const __grid_constant__ PmtModelDesc D, const void* __restrict__ haps, int hap_kind, long long hap_stride,
                  int n_variants, float* __restrict__ info_seq) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const unsigned act = smem_addr(p); p += 2 * BUF_BYTES;                 // hi buffer, then lo buffer (planes 8..15)
  unsigned char* img_s = p; p += TP.image_bytes;
  float* bias_s = reinterpret_cast<float*>(p); p += MAX_LAYERS * 32 * sizeof(float);
  signed char* codes = reinterpret_cast<signed char*>(p); p += MAX_CODES;
  Bars* S = reinterpret_cast<Bars*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;
  const int n_layers = TP.n_layers, G = TP.G, L0 = TP.L0;

  if (tid == 0) {
    mbar_init(smem_addr(&S->in_bar), EPI_WARPS);
    for (int c = 0; c < MAX_CHUNKS; ++c) mbar_init(smem_addr(&S->acc_bar[c]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // resident weight images (no-swizzle K-major planes, built by pack_cnn_tc_kernel) and biases
  for (int i = tid; i < TP.image_bytes / 16; i += THREADS)
    reinterpret_cast<uint4*>(img_s)[i] = __ldg(reinterpret_cast<const uint4*>(image) + i);
  for (int i = tid; i < n_layers * 32; i += THREADS) {
    const Layer& Ly = TP.layer[i / 32];
    const int n = i % 32, b_off = D.cnn_ops[Ly.op].b_off;
    bias_s[i] = (n < Ly.out_ch && b_off >= 0) ? __ldg(wflat + b_off + n) : 0.f;
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&S->tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = __shfl_sync(0xffffffffu, S->tmem_base, 0);
  const int n_groups = (n_variants + G - 1) / G;
  const unsigned img_a = smem_addr(img_s);

  if (warp == MMA_WARP) {
    // ===================================== MMA issuer =====================================
    unsigned in_par = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
      for (int l = 0; l < n_layers; ++l) {
        const Layer& Ly = TP.layer[l];
        const int nc = TP.n_chunks[l], N = Ly.N, taps = Ly.taps, first = Ly.first, ksteps = Ly.ksteps;
        const unsigned b_addr = img_a + Ly.img_off;
        const uint64_t b_hi = desc_ns(b_addr, N * 16), b_lo = desc_ns(b_addr + Ly.img_bytes, N * 16);
        mbar_wait(smem_addr(&S->in_bar), in_par);
        in_par ^= 1;
        tc_fence_after();
        if (elect_one()) {
          for (int c = 0; c < nc; ++c) {
            const unsigned d = tmem_base + c * CHUNK_COLS;
            const uint64_t a_hi = desc_ns(act + c * 128 * 16, PLANE_BYTES), a_lo = desc_ns(act + BUF_BYTES + c * 128 * 16, PLANE_BYTES);
            if (first) {
              if (N == 64) issue_first<64, PASSES>(ksteps, d, a_hi, b_hi, b_lo);
              else issue_first<32, PASSES>(ksteps, d, a_hi, b_hi, b_lo);
            } else {
              issue_shifted_n<32, PASSES>(taps, d, a_hi, a_lo, b_hi, b_lo);
            }
            mma_commit(smem_addr(&S->acc_bar[c]));
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================== epilogue warps =====================================
    const int grp = warp >> 2, quarter = warp & 3;
    const int row_c = quarter * 32 + lane;     // row inside a chunk = TMEM lane
    const unsigned trow = tmem_base + ((unsigned)(quarter * 32) << 16);
    unsigned acc_par = 0;                      // bit c: parity of acc_bar[c]
    unsigned in_par = 0;
    bool arrived = false;
    // a warp may only arrive for the next phase of in_bar once the previous phase has completed (a warp without a
    // chunk in some layer would otherwise arrive twice in one phase)
    auto signal_input_ready = [&]() {
      if (arrived) { mbar_wait(smem_addr(&S->in_bar), in_par); in_par ^= 1; }
      arrived = true;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_addr(&S->in_bar));
    };
    const int L2 = 2 * L0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
      const int v0 = g * G, nv = min(G, n_variants - v0);
      // ---- haplotype codes of the group, then the im2col rows of the first conv (batch.py:115-130) ----
      for (int idx = tid; idx < nv * L2; idx += 32 * EPI_WARPS) {
        const int v = idx / L2, hp = idx - v * L2;
        const long long off = (long long)(v0 + v) * hap_stride + hp;
        const int code = hap_kind == PMT_I64 ? (int)__ldg(reinterpret_cast<const long long*>(haps) + off)
                                             : (int)__ldg(reinterpret_cast<const short*>(haps) + off);
        codes[idx] = (signed char)((code >= 0 && code < 5) ? code : -1);
      }
      named_barrier(1, 32 * EPI_WARPS);
      {
        const Layer& Ly = TP.layer[0];
        const int rows = TP.n_chunks[0] * 128 < PLANE_ROWS ? TP.n_chunks[0] * 128 : PLANE_ROWS;
        const int planes = 2 * Ly.ksteps;
        for (int j = tid; j < rows; j += 32 * EPI_WARPS) {
          const unsigned rbase = act + j * 16;
          for (int pl = 0; pl < planes; ++pl) sts128(rbase + pl * PLANE_BYTES, make_float4(0.f, 0.f, 0.f, 0.f));
          const int v = j / L0, pos0 = j - v * L0;
          if (v < nv) {
            for (int tt = 0; tt < Ly.taps; ++tt) {
              const int pos = pos0 + tt;
              if (pos < L0) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int c = codes[v * L2 + h * L0 + pos];
                  if (c >= 0) {
                    const int k = tt * C0 + 2 * c + h;
                    sts_f32(rbase + (k >> 2) * PLANE_BYTES + (k & 3) * 4, 1.f);
                  }
                }
              }
            }
          }
        }
      }
      signal_input_ready();

      for (int l = 0; l < n_layers; ++l) {
        const Layer& Ly = TP.layer[l];
        const int nc = TP.n_chunks[l];
        const int L_in = Ly.L_in, L_next = Ly.L_next;
        const unsigned bias_a = smem_addr(bias_s + l * 32);
        for (int c = grp; c < nc; c += 2) {
          mbar_wait(smem_addr(&S->acc_bar[c]), (acc_par >> c) & 1u);
          acc_par ^= 1u << c;
          tc_fence_after();
          float x[32];
          {
            unsigned r[32];
            tmem_ld32(trow + c * CHUNK_COLS, r);
            if (Ly.dup) {
              unsigned r2[32];
              tmem_ld32(trow + c * CHUNK_COLS + 32, r2);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) x[i] = fmaxf(__uint_as_float(r[i]), __uint_as_float(r2[i]));
            } else {
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(r[i]);
            }
          }
          if (Ly.pool2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) x[i] = fmaxf(x[i], __shfl_xor_sync(0xffffffffu, x[i], 1));
          }
          const int j = c * 128 + row_c;
          const int v = j / L_in, pos = j - v * L_in;
          int pp = pos;
          bool valid = v < nv;
          if (Ly.pool2) { valid = valid && !(pos & 1); pp = pos >> 1; }
          valid = valid && pp < Ly.L_pool;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = lds128(bias_a + i * 4);
            x[i] += b.x; x[i + 1] += b.y; x[i + 2] += b.z; x[i + 3] += b.w;
          }
          if (Ly.to_global) {
            if (valid) {
              float* dst = info_seq + (long long)(v0 + v) * (D.d_info + D.d_seq) + D.d_info;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < Ly.out_ch) dst[i] = Ly.act == PMT_ACT_SELU ? SELU_SCALE * selu_u(x[i]) : x[i];
            }
          } else {
            if (Ly.act == PMT_ACT_SELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) x[i] = selu_u(x[i]);
            }
            if (valid) {
              const unsigned dst = act + (v * L_next + pp) * 16;
#pragma unroll
              for (int i = 0; i < 32; i += 4) sts128(dst + (i >> 2) * PLANE_BYTES, make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]));
              if (PASSES == 3) {
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] -= __uint_as_float(__float_as_uint(x[i]) & 0xFFFFE000u);
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                  sts128(dst + BUF_BYTES + (i >> 2) * PLANE_BYTES, make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]));
              }
            }
          }
        }
        if (l + 1 < n_layers) {
          tc_fence_before();
          signal_input_ready();
        }
      }
      // the next group's im2col overwrites rows the last layer's MMAs have finished reading (their commit was
      // observed by the epilogue above); all eight warps must be past that point
      tc_fence_before();
      named_barrier(1, 32 * EPI_WARPS);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
}

// ------------------------------------------------------------------------------------------------
// Weight images, B operand [N][K] in the no-swizzle K-major layout [tap][16-byte K chunk][n][4 floats];
// hi = TF32-rounded, lo = remainder (at + img_bytes).
// ------------------------------------------------------------------------------------------------
__device__ float cnn_weight(const PmtModelDesc& D, const Layer& Ly, const float* __restrict__ w, int tap, int n, int k) {
  const PmtCnnOp& op = D.cnn_ops[Ly.op];
  if (Ly.first) {   // k = tt * C0 + ch over the im2col row; columns [32, 64) are the conv one position later
    const int tt = k / C0, ch = k - tt * C0;
    const int co = n & 31, t = tt - (n >> 5);
    if (co >= op.out_ch || t < 0 || t >= op.ksize || tt >= Ly.taps) return 0.f;
    return w[op.w_off + (co * op.in_ch + ch) * op.ksize + t];
  }
  if (n >= op.out_ch || k >= Ly.in_ch) return 0.f;
  const float s = Ly.scale_in ? SELU_SCALE : 1.f;
  if (Ly.is_linear) return s * w[op.w_off + n * op.in_ch + (Ly.flat_len > 1 ? k * Ly.flat_len + tap : k)];   // flatten is channel-major
  return s * w[op.w_off + (n * op.in_ch + k) * op.ksize + tap];
}

__global__ void pack_cnn_tc_kernel(const __grid_constant__ PmtModelDesc D, const __grid_constant__ Plan TP, const float* __restrict__ w,
                                   unsigned char* __restrict__ image) {
  const Layer& Ly = TP.layer[blockIdx.x];
  const int N = Ly.N;
  const int taps = Ly.first ? 1 : Ly.taps, chunks = Ly.first ? 2 * Ly.ksteps : 8;
  for (int idx = threadIdx.x; idx < taps * chunks * N * 4; idx += blockDim.x) {
    const int e = idx & 3, n = (idx >> 2) % N, ch = (idx >> 2) / N % chunks, tap = (idx >> 2) / N / chunks;
    const float v = cnn_weight(D, Ly, w, tap, n, ch * 4 + e);
    unsigned hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
    const float hi = __uint_as_float(hb);
    const size_t off = (size_t)Ly.img_off + (size_t)idx * 4;
    *reinterpret_cast<float*>(image + off) = hi;
    *reinterpret_cast<float*>(image + off + Ly.img_bytes) = v - hi;
  }
}

}  // namespace cnntc
}  // namespace pmt

// ================================================================================================
// host side
// ================================================================================================
using namespace pmt;
using namespace pmt::cnntc;

// Builds the layer program; returns false when the CNN is outside this kernel's envelope (the FP32 SIMT kernel
// hap_cnn_kernel then runs instead).
static bool build_cnn_tc_plan(const pmt::Plan& P, cnntc::Plan* out) {
  cnntc::Plan& T = *out;
  memset(&T, 0, sizeof(T));
  const PmtModelDesc& d = P.d;
  T.L0 = d.hap_len;
  if (d.n_cnn_ops < 2 || d.hap_len < 1 || d.d_seq > 32) return false;
  int i = 0, L = d.hap_len, ch = C0;
  bool prev_selu = false;
  while (i < d.n_cnn_ops) {
    const PmtCnnOp& op = d.cnn_ops[i];
    if (T.n_layers >= MAX_LAYERS) return false;
    Layer& Ly = T.layer[T.n_layers];
    memset(&Ly, 0, sizeof(Ly));
    Ly.op = i;
    Ly.act = op.act;
    if (op.act != PMT_ACT_NONE && op.act != PMT_ACT_SELU) return false;
    if (op.kind == PMT_CNN_CONV) {
      if (op.stride != 1 || op.in_ch != ch || op.out_ch > 32 || op.in_len != L || op.ksize < 1 || op.ksize > 8 || op.out_len != L - op.ksize + 1 ||
          op.out_len < 1)
        return false;
      Ly.first = T.n_layers == 0;
      if (!Ly.first && op.in_ch > 32) return false;
      Ly.L_in = L; Ly.L_out = op.out_len; Ly.L_pool = op.out_len;
      Ly.ksize = op.ksize; Ly.in_ch = op.in_ch; Ly.out_ch = op.out_ch; Ly.scale_in = prev_selu;
      // pools that follow (a monotone activation commutes with max-pooling and was folded into the conv's act)
      int j = i + 1;
      for (; j < d.n_cnn_ops && d.cnn_ops[j].kind == PMT_CNN_POOL; ++j) {
        const PmtCnnOp& pl = d.cnn_ops[j];
        if (pl.ksize == 1 && pl.stride == 1) continue;
        if (Ly.dup || Ly.pool2 || pl.ksize != 2) return false;
        if (pl.stride == 1 && Ly.first) { Ly.dup = 1; Ly.L_pool = Ly.L_out - 1; }
        else if (pl.stride == 2 && (L % 2 == 0)) { Ly.pool2 = 1; Ly.L_pool = Ly.L_out / 2; }
        else return false;
        if (pl.out_len != Ly.L_pool || Ly.L_pool < 1) return false;
      }
      Ly.taps = Ly.first ? op.ksize + Ly.dup : op.ksize;
      if (Ly.first) {
        Ly.ksteps = (Ly.taps * C0 + 7) / 8;
        if (Ly.ksteps > 8) return false;
        Ly.N = Ly.dup ? 64 : 32;
      } else {
        Ly.N = 32;
      }
      Ly.L_next = Ly.L_pool;
      L = Ly.L_pool; ch = op.out_ch;
      prev_selu = op.act == PMT_ACT_SELU;
      i = j;
    } else if (op.kind == PMT_CNN_LINEAR) {
      if (T.n_layers == 0 || op.out_ch > 32) return false;
      Ly.is_linear = 1;
      Ly.flat_len = L;                       // first linear consumes the flattened [ch][L] map as a conv of kernel L
      if (L > 8 || op.in_ch != ch * L || ch > 32) return false;
      Ly.taps = L; Ly.ksize = L; Ly.in_ch = ch; Ly.out_ch = op.out_ch; Ly.scale_in = prev_selu;
      Ly.L_in = L; Ly.L_out = 1; Ly.L_pool = 1; Ly.L_next = 1; Ly.N = 32;
      L = 1; ch = op.out_ch;
      prev_selu = op.act == PMT_ACT_SELU;
      ++i;
    } else {
      return false;   // a pool before any conv
    }
    ++T.n_layers;
  }
  if (T.n_layers < 2 || !T.layer[T.n_layers - 1].is_linear || T.layer[T.n_layers - 1].out_ch != d.d_seq) return false;
  T.layer[T.n_layers - 1].to_global = 1;
  // images
  int bytes = 0;
  for (int l = 0; l < T.n_layers; ++l) {
    Layer& Ly = T.layer[l];
    Ly.img_off = bytes;
    Ly.img_bytes = (Ly.first ? 2 * Ly.ksteps : Ly.taps * 8) * Ly.N * 16;
    bytes += 2 * Ly.img_bytes;
  }
  T.image_bytes = bytes;
  // variants per group: every layer's rows must fit MAX_CHUNKS chunks and the planes; pick the G with the fewest MMA
  // cycles per variant
  const int fixed = 2 * BUF_BYTES + bytes + MAX_LAYERS * 32 * 4 + MAX_CODES + (int)sizeof(Bars) + 1024 + 64;
  if (fixed > 227 * 1024) return false;
  double best = 1e30;
  int best_g = 0;
  for (int G = 1; G <= 128; ++G) {
    if (G * d.hap_len > PLANE_ROWS || G * 2 * d.hap_len > MAX_CODES) break;
    double cost = 0;
    bool ok = true;
    for (int l = 0; l < T.n_layers; ++l) {
      const Layer& Ly = T.layer[l];
      const int chunks = (G * Ly.L_in + 127) / 128;
      if (chunks > MAX_CHUNKS) ok = false;
      cost += chunks * (Ly.first ? Ly.ksteps : Ly.taps * 4) * (Ly.N / 2) + 600.0;   // + per-layer hand-off
    }
    if (ok && cost / G < best) { best = cost / G; best_g = G; }
  }
  if (best_g == 0) return false;
  T.G = best_g;
  for (int l = 0; l < T.n_layers; ++l) T.n_chunks[l] = (T.G * T.layer[l].L_in + 127) / 128;
  return true;
}

bool pmt_cnn_tc_supported(const pmt::Plan& P) {
  cnntc::Plan T;
  return build_cnn_tc_plan(P, &T);
}

size_t pmt_cnn_tc_image_bytes(const pmt::Plan& P) {
  cnntc::Plan T;
  if (!build_cnn_tc_plan(P, &T)) return 0;
  return (size_t)T.image_bytes + 256;
}

template <int PASSES>
static void launch_cnn_tc(const cnntc::Plan& T, const unsigned char* image, const float* weights, const PmtModelDesc& D,
                          const PmtBatch* batch, float* info_seq, int grid, cudaStream_t st) {
  const size_t smem = 2 * BUF_BYTES + T.image_bytes + MAX_LAYERS * 32 * sizeof(float) + MAX_CODES + sizeof(Bars) + 1024 + 64;
  cudaFuncSetAttribute(hap_cnn_tc_kernel<PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  hap_cnn_tc_kernel<PASSES><<<grid, THREADS, smem, st>>>(T, image, weights, D, batch->haplotypes, batch->hap_kind, batch->hap_stride,
                                                         batch->n_variants, info_seq);
}

// `image` is a 16-byte aligned device buffer of pmt_cnn_tc_image_bytes(P) bytes.
int pmt_launch_cnn_tc(const pmt::Plan& P, const float* weights, const PmtBatch* batch, float* info_seq, unsigned char* image,
                      int n_sm, int mode, cudaStream_t st) {
  cnntc::Plan T;
  PMT_CHECK(build_cnn_tc_plan(P, &T), "haplotype CNN outside the tensor-core envelope");
  pack_cnn_tc_kernel<<<T.n_layers, 256, 0, st>>>(P.d, T, weights, image);
  const int n_groups = (batch->n_variants + T.G - 1) / T.G;
  const int grid = n_groups < n_sm ? n_groups : n_sm;
  if (mode == PMT_PRECISION_TF32) launch_cnn_tc<1>(T, image, weights, P.d, batch, info_seq, grid, st);
  else launch_cnn_tc<3>(T, image, weights, P.d, batch, info_seq, grid, st);
  return 0;
}
