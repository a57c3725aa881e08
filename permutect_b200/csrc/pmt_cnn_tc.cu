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

#include "pmt_cnn_tc.cuh"

namespace pmt {
namespace cnntc {

using namespace pmt::tc;

// Shared-memory matrix descriptor, K-major, no swizzle.  Low word: start address >> 4 | (LBO >> 4) << 16 with LBO = the
// distance between K-adjacent core matrices; high word (constant): SBO = 128 B between 8-row groups, descriptor version 1.
constexpr unsigned DESC_HI = 8u | (1u << 14);
__device__ __forceinline__ unsigned desc_lo(unsigned addr, unsigned lbo_bytes) { return ((addr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16); }
template <unsigned IDESC>
__device__ __forceinline__ void mma_ss(unsigned tmem_d, unsigned a_lo32, unsigned b_lo32, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %3, 0;\n\t"
      "mov.b64 da, {%1, %4};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo32), "r"(b_lo32), "r"(accumulate), "n"(DESC_HI), "n"(IDESC)
      : "memory");
}
__host__ __device__ constexpr unsigned make_idesc(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// Split-precision scheme (PASSES == 3): the weight image stacks [W_hi ; W_lo] along N, so ONE MMA of width 2N computes
// A_hi.W_hi (columns [0, N)) and A_hi.W_lo (columns [N, 2N)) from a single read of the A rows -- operand fetch from shared
// memory, not the tensor pipe, bounds these small-N MMAs -- and a second MMA of width N adds A_lo.W_hi to columns [0, N).
// The epilogue sums the two column ranges.  PASSES == 1 uses the W_hi rows only.
//
// shifted-window conv: taps x 4 k-steps.  The tap loop is a warp-uniform runtime loop (no jump table: an indirect
// branch would push every descriptor out of the uniform registers); inside a tap all offsets are compile-time constants.
template <int PASSES>
__device__ __forceinline__ void issue_shifted(int taps, unsigned d, unsigned a_hi, unsigned a_lo, unsigned b) {
  constexpr int R = 64;   // image rows per plane: 32 hi + 32 lo
#pragma unroll 1
  for (int t = 0; t < taps; ++t) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const unsigned ao = (unsigned)(2 * ks * (PLANE_BYTES / 16));
      const unsigned bo = (unsigned)(2 * ks * R);
      const unsigned acc = (ks == 0) ? (t > 0 ? 1u : 0u) : 1u;
      if (PASSES == 3) {
        mma_ss<make_idesc(64)>(d, a_hi + ao, b + bo, acc);
        mma_ss<make_idesc(32)>(d, a_lo + ao, b + bo, 1u);
      } else {
        mma_ss<make_idesc(32)>(d, a_hi + ao, b + bo, acc);
      }
    }
    a_hi += 1; a_lo += 1; b += 8 * R;
  }
}
// first conv: one-hot im2col rows (exact: no A lo pass), K = 8 * KS columns over consecutive planes
template <int N, int PASSES>
__device__ __forceinline__ void issue_first(int ksteps, unsigned d, unsigned a, unsigned b) {
  constexpr int R = 2 * N;
  constexpr unsigned idesc = make_idesc(PASSES == 3 ? 2 * N : N);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    if (ks < ksteps) {
      const unsigned ao = (unsigned)(2 * ks * (PLANE_BYTES / 16)), bo = (unsigned)(2 * ks * R);
      mma_ss<idesc>(d, a + ao, b + bo, ks != 0 ? 1u : 0u);
    }
  }
}

struct ItemDev {   // per-item MMA operands, computed once per CTA
  unsigned a_hi, a_lo, b, d;
  int first, count, N, need;
};

struct EpiItem {   // per-item epilogue constants (three 16-byte shared loads)
  int flags, lo_col, tmem_col, bias_off;
  int inv_L, L_in, L_pool, L_next;
  int row0, out_ch, save_a, save_bits;   // SAVE: float offsets inside a group's block of the save buffer (-1: none)
};
constexpr int F_DUP = 1, F_POOL2 = 2, F_GLOBAL = 4, F_SELU = 8;

struct Bars {
  EpiItem epi[MAX_ITEMS];
  unsigned long long done_bar[MAX_ITEMS + 1];   // [0]: im2col rows written; [1 + i]: epilogue of item i finished (16 warps)
  unsigned long long acc_bar[MAX_ITEMS];        // accumulator of item i complete (tcgen05.commit)
  ItemDev item[MAX_ITEMS];
  unsigned tmem_base;
  int pad_;
};

template <int PASSES, bool TRACE, bool SAVE = false>
__global__ void __launch_bounds__(THREADS, 1)
hap_cnn_tc_kernel(const __grid_constant__ Plan TP, const unsigned char* __restrict__ image, const float* __restrict__ wflat,
                  const __grid_constant__ PmtModelDesc D, const void* __restrict__ haps, int hap_kind, long long hap_stride,
                  int n_variants, float* __restrict__ info_seq, long long* __restrict__ trace,
                  const __grid_constant__ SaveLayout SL, float* __restrict__ save) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const unsigned act = smem_addr(p); p += 2 * BUF_BYTES;                 // hi buffer, then lo buffer (planes 8..15)
  unsigned char* img_s = p; p += TP.image_bytes;
  float* bias_s = reinterpret_cast<float*>(p); p += MAX_LAYERS * 32 * sizeof(float);
  Bars* S = reinterpret_cast<Bars*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;
  const int n_layers = TP.n_layers, n_items = TP.n_items, G = TP.G, L0 = TP.L0;
  // optional cycle trace of CTA 0 (profiles/trace_cnn.py): entries of (event id, clock) per recording thread
  int tr_n = 0;
  const bool tr_on = trace != nullptr && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 5 || warp == MMA_WARP);
  long long* tr = trace + (warp == MMA_WARP ? 2 : (warp == 5 ? 1 : 0)) * 1024;
  auto TR = [&](int id) { if (TRACE && tr_on && tr_n < 510) { tr[2 * tr_n] = id; tr[2 * tr_n + 1] = clock64(); ++tr_n; } };

  if (tid == 0) {
    for (int i = 0; i <= n_items; ++i) mbar_init(smem_addr(&S->done_bar[i]), EPI_WARPS);
    for (int i = 0; i < n_items; ++i) mbar_init(smem_addr(&S->acc_bar[i]), 1);
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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&S->tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = __shfl_sync(0xffffffffu, S->tmem_base, 0);
  const int n_groups = (n_variants + G - 1) / G;
  if (tid < n_items) {
    const Item& I = TP.item[tid];
    const Layer& Ly = TP.layer[I.layer];
    ItemDev& o = S->item[tid];
    o.a_hi = desc_lo(act + I.chunk * 128 * 16, PLANE_BYTES);
    o.a_lo = desc_lo(act + BUF_BYTES + I.chunk * 128 * 16, PLANE_BYTES);
    o.b = desc_lo(smem_addr(img_s) + Ly.img_off, 2 * Ly.N * 16);
    o.d = tmem_base + I.chunk * CHUNK_COLS;
    o.first = Ly.first; o.count = Ly.first ? Ly.ksteps : Ly.taps; o.N = Ly.N; o.need = I.need;
    EpiItem& e = S->epi[tid];
    e.flags = (Ly.dup ? F_DUP : 0) | (Ly.pool2 ? F_POOL2 : 0) | (Ly.to_global ? F_GLOBAL : 0) | (Ly.act == PMT_ACT_SELU ? F_SELU : 0);
    e.lo_col = Ly.N; e.tmem_col = I.chunk * CHUNK_COLS; e.bias_off = I.layer * 32 * (int)sizeof(float);
    e.inv_L = Ly.inv_L; e.L_in = Ly.L_in; e.L_pool = Ly.L_pool; e.L_next = Ly.L_next;
    e.row0 = I.chunk * 128; e.out_ch = Ly.out_ch;
    e.save_a = SAVE ? SL.a_off[I.layer] : -1; e.save_bits = SAVE ? SL.bits_off[I.layer] : -1;
  }
  __syncthreads();

  if (warp == MMA_WARP) {
    // ===================================== MMA issuer =====================================
    // Items (layer, chunk) are issued in order; item i waits only for the epilogue items that produce the input rows it
    // reads (need), so a layer starts while the previous layer's last chunks are still in their epilogue.
    unsigned gpar = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, gpar ^= 1) {
      int waited = 0;   // done_bar[0 .. waited) observed in this group
      for (int it = 0; it < n_items; ++it) {
        const ItemDev I = S->item[it];
        TR(100 + it);
        for (; waited <= I.need; ++waited) mbar_wait(smem_addr(&S->done_bar[waited]), gpar);
        tc_fence_after();
        TR(130 + it);
        if (elect_one()) {
          if (I.first) {
            if (I.N == 64) issue_first<64, PASSES>(I.count, I.d, I.a_hi, I.b);
            else issue_first<32, PASSES>(I.count, I.d, I.a_hi, I.b);
          } else {
            issue_shifted<PASSES>(I.count, I.d, I.a_hi, I.a_lo, I.b);
          }
          mma_commit(smem_addr(&S->acc_bar[it]));
        }
        __syncwarp();
        TR(160 + it);
      }
    }
  } else {
    // ===================================== epilogue warps =====================================
    // every item is processed by all 16 warps: warp = (column quarter, lane quarter); a thread owns 8 channels of one row
    const int quarter = warp & 3, cq = warp >> 2;
    const int row_c = quarter * 32 + lane;     // row inside a chunk = TMEM lane
    const unsigned trow = tmem_base + ((unsigned)(quarter * 32) << 16) + cq * 8;
    const unsigned bias_base = smem_addr(bias_s) + cq * 8 * (int)sizeof(float);
    const Layer& L0y = TP.layer[0];
    const int rows0 = TP.n_chunks[0] * 128 < PLANE_ROWS ? TP.n_chunks[0] * 128 : PLANE_ROWS;
    const int rpw = (rows0 + EPI_WARPS - 1) / EPI_WARPS;            // im2col rows per warp
    const int r_lo = warp * rpw, r_hi = min(rows0, r_lo + rpw);
    const int planes0 = 2 * L0y.ksteps, taps0 = L0y.taps;
    // im2col scatter tasks of this lane: (row, input position, haplotype) -> one 1.0 of the one-hot row.  The mapping is
    // the same for every group; the codes of the NEXT group are fetched one group ahead and stay in registers.
    unsigned t_row[MAX_TASKS];   // shared address of the row (plane 0), 0 = no task
    int t_col[MAX_TASKS], t_var[MAX_TASKS], t_off[MAX_TASKS];
#pragma unroll
    for (int k = 0; k < MAX_TASKS; ++k) {
      const int task = lane + 32 * k;
      const int h = task & 1, rt = task >> 1;
      const int tt = rt % taps0, j = r_lo + rt / taps0;
      const int v = j / L0, pos = j - v * L0 + tt;
      const bool ok = j < r_hi && pos < L0;
      t_row[k] = ok ? act + j * 16 : 0u;
      t_col[k] = tt * C0 + h;
      t_var[k] = ok ? v : (1 << 30);
      t_off[k] = ok ? (int)(v * hap_stride) + h * L0 + pos : 0;
    }
    int codes[MAX_TASKS];
    auto fetch_codes = [&](int g) {
      const long long base = (long long)g * G * hap_stride;
      const int nv = g < n_groups ? min(G, n_variants - g * G) : 0;
      if (hap_kind == PMT_I64) {
#pragma unroll
        for (int k = 0; k < MAX_TASKS; ++k) codes[k] = t_var[k] < nv ? (int)__ldg(reinterpret_cast<const long long*>(haps) + base + t_off[k]) : -1;
      } else {
#pragma unroll
        for (int k = 0; k < MAX_TASKS; ++k) codes[k] = t_var[k] < nv ? (int)__ldg(reinterpret_cast<const short*>(haps) + base + t_off[k]) : -1;
      }
    };
    fetch_codes(blockIdx.x);
    unsigned gpar = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, gpar ^= 1) {
      const int v0 = g * G, nv = min(G, n_variants - v0);
      TR(1);
      // ---- im2col rows of the first conv (batch.py:115-130): zero this warp's rows, then scatter the ones ----
      for (int j = r_lo + lane; j < r_hi; j += 32) {
        const unsigned ra = act + j * 16;
#pragma unroll
        for (int pl = 0; pl < 16; ++pl)
          if (pl < planes0) sts128(ra + pl * PLANE_BYTES, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      __syncwarp();
      TR(3);
#pragma unroll
      for (int k = 0; k < MAX_TASKS; ++k) {
        if ((unsigned)codes[k] < 5u) {
          const int col = t_col[k] + 2 * codes[k];
          sts_f32(t_row[k] + (col >> 2) * PLANE_BYTES + (col & 3) * 4, 1.f);
        }
      }
      TR(4);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_addr(&S->done_bar[0]));
      TR(2);
      fetch_codes(g + gridDim.x);

      for (int it = 0; it < n_items; ++it) {
        const unsigned ea = smem_addr(&S->epi[it]);
        const float4 e0f = lds128(ea), e1f = lds128(ea + 16), e2f = lds128(ea + 32);
        const int flags = __float_as_int(e0f.x), N = __float_as_int(e0f.y);
        const unsigned tcol = trow + __float_as_int(e0f.z);
        const unsigned bias_a = bias_base + __float_as_int(e0f.w);
        const int inv_L = __float_as_int(e1f.x), L_in = __float_as_int(e1f.y), L_pool = __float_as_int(e1f.z), L_next = __float_as_int(e1f.w);
        const int out_ch = __float_as_int(e2f.y);
        const int save_a = __float_as_int(e2f.z), save_bits = __float_as_int(e2f.w);
        unsigned win = 0;   // SAVE: bit i = the second element of the pooling window won for channel 8 cq + i
        const int j = __float_as_int(e2f.x) + row_c;
        const int v = (j * inv_L) >> 16, pos = j - v * L_in;
        int pp = pos;
        bool valid = v < nv;
        if (flags & F_POOL2) { valid = valid && !(pos & 1); pp = pos >> 1; }
        valid = valid && pp < L_pool;
        const float4 b0 = lds128(bias_a), b1 = lds128(bias_a + 16);
        TR(200 + it);
        mbar_wait(smem_addr(&S->acc_bar[it]), gpar);
        tc_fence_after();
        TR(230 + it);
        float x[8];
        {
          unsigned r[8], rl[8];
          tmem_ld8(tcol, r);
          if (PASSES == 3) tmem_ld8(tcol + N, rl);
          if (flags & F_DUP) {
            unsigned r2[8], rl2[8];
            tmem_ld8(tcol + 32, r2);
            if (PASSES == 3) tmem_ld8(tcol + N + 32, rl2);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float u0 = PASSES == 3 ? __uint_as_float(r[i]) + __uint_as_float(rl[i]) : __uint_as_float(r[i]);
              const float u1 = PASSES == 3 ? __uint_as_float(r2[i]) + __uint_as_float(rl2[i]) : __uint_as_float(r2[i]);
              x[i] = fmaxf(u0, u1);
              if (SAVE) win |= (u1 > u0 ? 1u : 0u) << i;
            }
          } else {
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = PASSES == 3 ? __uint_as_float(r[i]) + __uint_as_float(rl[i]) : __uint_as_float(r[i]);
          }
        }
        TR(500 + it);
        tc_fence_before();   // the accumulator has been read: a later item may overwrite it once this warp has arrived
        if (flags & F_POOL2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float o = __shfl_xor_sync(0xffffffffu, x[i], 1);
            if (SAVE) win |= (o > x[i] ? 1u : 0u) << i;   // read on the even row of the pair: o is the odd row's value
            x[i] = fmaxf(x[i], o);
          }
        }
        x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w; x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
        if (flags & F_GLOBAL) {
          if (valid) {
            float* dst = info_seq + (long long)(v0 + v) * (D.d_info + D.d_seq) + D.d_info + cq * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (cq * 8 + i < out_ch) dst[i] = (flags & F_SELU) ? SELU_SCALE * selu_u(x[i]) : x[i];
          }
        } else {
          if (flags & F_SELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = selu_u(x[i]);
          }
          if (SAVE && valid && save_a >= 0) {   // training: a_l (and the window bits) for the backward (pmt_cnn_bwd.cu)
            const int cols = G * L_next, col = v * L_next + pp;
            float* sg = save + (size_t)g * SL.group_floats;
            float* dstg = sg + save_a + (cq * 8) * cols + col;
#pragma unroll
            for (int i = 0; i < 8; ++i) {   // rounded to TF32: the backward's MMAs read them as they lie
              unsigned rb;
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(rb) : "f"(x[i]));
              dstg[i * cols] = __uint_as_float(rb);
            }
            if (save_bits >= 0) reinterpret_cast<unsigned char*>(sg + save_bits)[cq * cols + col] = (unsigned char)win;
          }
          if (valid) {
            const unsigned dst = act + (2 * cq) * PLANE_BYTES + (v * L_next + pp) * 16;
            sts128(dst, make_float4(x[0], x[1], x[2], x[3]));
            sts128(dst + PLANE_BYTES, make_float4(x[4], x[5], x[6], x[7]));
            if (PASSES == 3) {
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] -= __uint_as_float(__float_as_uint(x[i]) & 0xFFFFE000u);
              sts128(dst + BUF_BYTES, make_float4(x[0], x[1], x[2], x[3]));
              sts128(dst + BUF_BYTES + PLANE_BYTES, make_float4(x[4], x[5], x[6], x[7]));
            }
          }
          TR(530 + it);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          TR(560 + it);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_addr(&S->done_bar[1 + it]));
        TR(260 + it);
      }
    }
  }
  if (TRACE && tr_on) tr[1022] = tr_n;
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// ------------------------------------------------------------------------------------------------
// Weight images, B operand in the no-swizzle K-major layout [tap][16-byte K chunk][row][4 floats]; a plane holds
// 2N rows: rows [0, N) = TF32-rounded W, rows [N, 2N) = the remainder W - W_hi.
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
  const int N = Ly.N, R = 2 * N;
  const int taps = Ly.first ? 1 : Ly.taps, chunks = Ly.first ? 2 * Ly.ksteps : 8;
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < taps * chunks * R * 4; idx += gridDim.y * blockDim.x) {
    const int e = idx & 3, r = (idx >> 2) % R, ch = (idx >> 2) / R % chunks, tap = (idx >> 2) / R / chunks;
    const float v = cnn_weight(D, Ly, w, tap, r % N, ch * 4 + e);
    unsigned hb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
    const float hi = __uint_as_float(hb);
    *reinterpret_cast<float*>(image + (size_t)Ly.img_off + (size_t)idx * 4) = r < N ? hi : v - hi;
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
bool pmt_build_cnn_tc_plan(const pmt::Plan& P, cnntc::Plan* out) {
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
    Ly.img_bytes = (Ly.first ? 2 * Ly.ksteps : Ly.taps * 8) * 2 * Ly.N * 16;   // hi and lo rows interleaved per plane
    bytes += Ly.img_bytes;
    Ly.inv_L = 65536 / Ly.L_in + 1;
  }
  T.image_bytes = bytes;
  // variants per group: every layer's rows must fit MAX_CHUNKS chunks and the planes; pick the G with the fewest MMA
  // cycles per variant
  const int fixed = 2 * BUF_BYTES + bytes + MAX_LAYERS * 32 * 4 + (int)sizeof(Bars) + 1024 + 64;
  if (fixed > 227 * 1024) return false;
  double best = 1e30;
  int best_g = 0;
  for (int G = 1; G <= 128; ++G) {
    if (G * d.hap_len > PLANE_ROWS) break;
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
  {   // the im2col scatter keeps its one-hot entries in MAX_TASKS registers per lane
    const int rows0 = T.n_chunks[0] * 128 < PLANE_ROWS ? T.n_chunks[0] * 128 : PLANE_ROWS;
    const int rpw = (rows0 + EPI_WARPS - 1) / EPI_WARPS;
    if ((rpw * T.layer[0].taps * 2 + 31) / 32 > MAX_TASKS) return false;
  }
  // work items in issue order, with the epilogue item each one depends on
  int first_item[MAX_LAYERS];
  for (int l = 0; l < T.n_layers; ++l) {
    first_item[l] = T.n_items;
    const Layer& Ly = T.layer[l];
    for (int c = 0; c < T.n_chunks[l]; ++c) {
      if (T.n_items >= MAX_ITEMS) return false;
      Item& I = T.item[T.n_items];
      I.layer = l; I.chunk = c; I.need = 0;
      if (l > 0) {
        const Layer& Pv = T.layer[l - 1];
        const int taps = Ly.taps, s = Pv.pool2 ? 2 : 1;
        int last_row = 128 * c + 128 + taps - 2;                    // last input row the chunk's MMAs read
        if (last_row > T.G * Ly.L_in - 1) last_row = T.G * Ly.L_in - 1;
        const int v = last_row / Ly.L_in, pq = last_row % Ly.L_in;  // written by the previous layer's row (v, pq * s)
        int src_chunk = (v * Pv.L_in + pq * s) / 128;
        if (src_chunk < c && c < T.n_chunks[l - 1]) src_chunk = c;   // the accumulator columns of (l-1, c) are reused
        if (src_chunk > T.n_chunks[l - 1] - 1) src_chunk = T.n_chunks[l - 1] - 1;
        I.need = 1 + first_item[l - 1] + src_chunk;
        if (T.n_items > 0 && I.need < T.item[T.n_items - 1].need) I.need = T.item[T.n_items - 1].need;
      }
      ++T.n_items;
    }
  }
  return true;
}

bool pmt_cnn_tc_supported(const pmt::Plan& P) {
  cnntc::Plan T;
  return pmt_build_cnn_tc_plan(P, &T);
}

size_t pmt_cnn_tc_image_bytes(const pmt::Plan& P) {
  cnntc::Plan T;
  if (!pmt_build_cnn_tc_plan(P, &T)) return 0;
  return (size_t)T.image_bytes + 256;
}

template <int PASSES, bool TRACE>
static int launch_cnn_tc(const cnntc::Plan& T, const unsigned char* image, const float* weights, const PmtModelDesc& D,
                          const PmtBatch* batch, float* info_seq, int grid, long long* trace, cudaStream_t st) {
  const size_t smem = 2 * BUF_BYTES + T.image_bytes + MAX_LAYERS * 32 * sizeof(float) + sizeof(Bars) + 1024 + 64;
  SaveLayout SL;
  memset(&SL, 0, sizeof(SL));
  PMT_CUDA(cudaFuncSetAttribute(hap_cnn_tc_kernel<PASSES, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hap_cnn_tc_kernel<PASSES, TRACE><<<grid, THREADS, smem, st>>>(T, image, weights, D, batch->haplotypes, batch->hap_kind, batch->hap_stride,
                                                         batch->n_variants, info_seq, trace, SL, nullptr);
  return 0;
}

void pmt_cnn_save_layout(const cnntc::Plan& T, SaveLayout* out) {
  SaveLayout& S = *out;
  memset(&S, 0, sizeof(S));
  S.G = T.G; S.n_layers = T.n_layers;
  int off = 0;
  for (int l = 0; l < MAX_LAYERS; ++l) { S.a_off[l] = -1; S.bits_off[l] = -1; }
  for (int l = 0; l < T.n_layers; ++l) {
    const Layer& Ly = T.layer[l];
    if (Ly.to_global) continue;
    const int cols = T.G * Ly.L_next;
    S.a_off[l] = off; off += 32 * cols;
    if (Ly.dup || Ly.pool2) { S.bits_off[l] = off; off += cols; }   // [4][cols] bytes
  }
  S.group_floats = (off + 3) & ~3;
}

// Training recompute: always the split-precision mode (the saved activations are the fp32-parity ones).
int pmt_launch_cnn_tc_save(const pmt::Plan& P, const cnntc::Plan& T, const float* weights, const PmtBatch* batch, int v_first, int n,
                           float* info_seq, const unsigned char* image, float* save, int n_sm, cudaStream_t st) {
  const size_t smem = 2 * BUF_BYTES + T.image_bytes + MAX_LAYERS * 32 * sizeof(float) + sizeof(Bars) + 1024 + 64;
  SaveLayout SL;
  pmt_cnn_save_layout(T, &SL);
  const int n_groups = (n + T.G - 1) / T.G;
  const int grid = n_groups < n_sm ? n_groups : n_sm;
  const size_t esz = batch->hap_kind == PMT_I64 ? 8 : 2;
  const void* haps = reinterpret_cast<const unsigned char*>(batch->haplotypes) + (size_t)v_first * batch->hap_stride * esz;
  float* out = info_seq + (size_t)v_first * (P.d.d_info + P.d.d_seq);
  PMT_CUDA(cudaFuncSetAttribute(hap_cnn_tc_kernel<3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hap_cnn_tc_kernel<3, false, true><<<grid, THREADS, smem, st>>>(T, image, weights, P.d, haps, batch->hap_kind, batch->hap_stride, n, out,
                                                                  nullptr, SL, save);
  return 0;
}

int pmt_pack_cnn_tc_images(const pmt::Plan& P, const cnntc::Plan& T, const float* weights, unsigned char* image, cudaStream_t st) {
  pack_cnn_tc_kernel<<<dim3(T.n_layers, 16), 256, 0, st>>>(P.d, T, weights, image);
  return 0;
}

// `image` is a 16-byte aligned device buffer of pmt_cnn_tc_image_bytes(P) bytes.
static long long* g_cnn_trace = nullptr;
// Measurement hook: device buffer of 3 x 1024 int64 that CTA 0 of the next tensor-core CNN launches fills with
// (event, clock64) pairs; NULL disarms.
extern "C" int pmt_set_cnn_trace(long long* device_buffer) { g_cnn_trace = device_buffer; return 0; }

int pmt_launch_cnn_tc(const pmt::Plan& P, const float* weights, const PmtBatch* batch, float* info_seq, unsigned char* image,
                      bool reuse_image, int n_sm, int mode, cudaStream_t st) {
  cnntc::Plan T;
  PMT_CHECK(pmt_build_cnn_tc_plan(P, &T), "haplotype CNN outside the tensor-core envelope");
  if (!reuse_image) pack_cnn_tc_kernel<<<dim3(T.n_layers, 16), 256, 0, st>>>(P.d, T, weights, image);
  const int n_groups = (batch->n_variants + T.G - 1) / T.G;
  const int grid = n_groups < n_sm ? n_groups : n_sm;
  if (g_cnn_trace) {
    if (mode == PMT_PRECISION_TF32) { if (launch_cnn_tc<1, true>(T, image, weights, P.d, batch, info_seq, grid, g_cnn_trace, st)) return 1; }
    else { if (launch_cnn_tc<3, true>(T, image, weights, P.d, batch, info_seq, grid, g_cnn_trace, st)) return 1; }
  } else {
    if (mode == PMT_PRECISION_TF32) { if (launch_cnn_tc<1, false>(T, image, weights, P.d, batch, info_seq, grid, nullptr, st)) return 1; }
    else { if (launch_cnn_tc<3, false>(T, image, weights, P.d, batch, info_seq, grid, nullptr, st)) return 1; }
  }
  return 0;
}
