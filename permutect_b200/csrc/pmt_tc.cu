// Tensor-core (tcgen05 / TMEM) forward of the read path for sm_100a.
//
// One CTA = one tile of 128 reads = the 128 lanes of TMEM.  Warps 0-3 are EPILOGUE warps: thread r owns
// read r of the tile for the whole network -- it keeps the read's residual stream in registers, writes the
// row of the next A operand (K-major, 128B-swizzled) to shared memory, and after the MMA reads its own
// accumulator row back with tcgen05.ld, so bias / SELU / residual / LayerNorm / gating / clustering head are
// all register-resident row-local work.  Warp 4 is the CONTROL warp: one lane streams each layer's
// pre-swizzled weight image with cp.async.bulk (mbarrier complete_tx) and issues the tcgen05.mma chain
// (M = 128, N = padded layer width, kind::tf32), committing to an mbarrier the epilogue waits on.
// The only cross-read coupling -- the per-variant mean fields of the gated blocks and the final set sums --
// goes through a small shared exchange buffer between the four epilogue warps.
//
// Precision modes (pmt_set_precision): TF32 (one MMA per k-step) and TF32x3 (hi/lo split of both operands,
// three MMAs: Ahi.Bhi + Alo.Bhi + Ahi.Blo, ~2^-21 relative error, the fp32-parity mode on tensor cores).
// Layers with separate ref / alt weights (proj1 / proj2 of the gated block) are computed for both weight
// sets side by side in N; each row keeps the half that matches its read type.
#include <cstring>

#include "pmt_host.h"
#include "pmt_tile.cuh"

namespace pmt {
namespace tc {

constexpr int EPI_THREADS = 128;
constexpr int THREADS = 160;
constexpr int A_KB_BYTES = 128 * 128;   // one 32-element K block of the A operand: 128 rows x 128 B
constexpr int TMEM_COLS = 128;
constexpr int MAX_TC_OPS = 64;

struct TcOp {
  int K, N;          // padded: K multiple of 8, N multiple of 16 (dual ops: N = 2 * Np)
  int Np;            // padded width of one weight set
  int n_out;         // real output width of one weight set
  int dual;
  int img_off;       // byte offset of the hi image in the TC image buffer (1024-aligned); lo image follows
  int img_bytes;     // bytes of ONE image (hi); the staged size is img_bytes * (passes == 3 ? 2 : 1)
  int b_off, b_alt_off;
  int k_real;        // un-padded reduction length (row length of the weight matrix)
  int w_off, w_alt_off;
  int a_exact;       // the A operand is exactly representable in TF32 (decoded reads): skip the Alo.Bhi MMA
};

struct TcPlan {
  int n_ops, read0, blk0, red0;
  int image_bytes;   // total bytes of the TC image buffer
  int stage_bytes;   // largest staged op image (both parts)
  TcOp op[MAX_TC_OPS];
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ uint64_t smem_desc(unsigned addr) {
  // K-major, SWIZZLE_128B: 8-row groups 1024 B apart, descriptor version 1 (sm_100)
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_tf32(unsigned tmem_d, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float* v) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tf32_round(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Shared state of one CTA
struct Shared {
  unsigned long long bar_a, bar_w, bar_d;   // operand ready (128 arrivals) / weights landed (tx) / MMA done (commit)
  unsigned tmem_base;
  int req;                                  // op requested by the epilogue (-1 = finished)
  int warp_cnt[4];
  int nv, v0, ref_pad, rows;
  int rowvar[TILE];
  long long rowidx[TILE];
  int ref_start[TILE], ref_cnt[TILE], alt_start[TILE], alt_cnt[TILE];
};

struct TcArgs {
  const float* wflat;
  const unsigned char* image;   // TC weight images (global)
  PmtBatch batch;
  PmtOutputs out;
  int n_claims, claim;
};

// Writes this row's vector v[0..n) (zero beyond n, up to KP) as the A operand row: hi part to a0, lo part to a1.
template <int KP, int PASSES>
__device__ __forceinline__ void write_a_row(const float* v, int n, int kchunks, int row, unsigned a0, unsigned a1) {
#pragma unroll
  for (int c = 0; c < KP / 4; ++c) {
    if (c < kchunks) {
      float4 hi, lo;
      float e[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) e[j] = (c * 4 + j < n) ? v[c * 4 + j] : 0.f;
      if (PASSES == 3) {
        hi.x = tf32_round(e[0]); hi.y = tf32_round(e[1]); hi.z = tf32_round(e[2]); hi.w = tf32_round(e[3]);
        lo.x = e[0] - hi.x; lo.y = e[1] - hi.y; lo.z = e[2] - hi.z; lo.w = e[3] - hi.w;
      } else {
        hi = make_float4(e[0], e[1], e[2], e[3]);
      }
      const unsigned off = (c >> 3) * A_KB_BYTES + row * 128 + ((((unsigned)c & 7u) ^ ((unsigned)row & 7u)) << 4);
      sts128(a0 + off, hi);
      if (PASSES == 3) sts128(a1 + off, lo);
    }
  }
}

template <int PASSES>
struct Epi {   // per-thread state of an epilogue thread
  Shared* S;
  unsigned a0, a1, bar_a, bar_d, tmem_row;   // tmem_row: TMEM address of this thread's lane, column 0
  unsigned phase_d;
  int row;
  const float* W;

  // publish the A operand for `op`, let the control warp run the MMA, wait for the accumulator
  __device__ __forceinline__ void run_mma(int op) {
    fence_async_proxy();
    tc_fence_before();
    if (row == 0) S->req = op;
    mbar_arrive(bar_a);
    mbar_wait(bar_d, phase_d);
    phase_d ^= 1;
    tc_fence_after();
  }
  // out[0..n_out) = acc[col0 .. col0 + n_out) for up to 64 columns
  __device__ __forceinline__ void load_cols(int col0, int n_pad, float* out) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c * 16 < n_pad) tmem_ld16(tmem_row + col0 + c * 16, out + c * 16);
  }
};

template <int PASSES>
__device__ __forceinline__ void control_loop(const TcPlan& TP, Shared* S, const unsigned char* image, unsigned wbuf,
                                             unsigned a0, unsigned a1) {
  const unsigned bar_a = smem_addr(&S->bar_a), bar_w = smem_addr(&S->bar_w), bar_d = smem_addr(&S->bar_d);
  unsigned phase_a = 0, phase_w = 0, phase_d = 0;
  int inflight = 0;
  {
    const TcOp& o = TP.op[0];
    const unsigned bytes = o.img_bytes * (PASSES == 3 ? 2 : 1);
    mbar_expect_tx(bar_w, bytes);
    bulk_g2s(wbuf, image + o.img_off, bytes, bar_w);
  }
  const unsigned tmem_d = S->tmem_base;
  for (;;) {
    mbar_wait(bar_a, phase_a);
    phase_a ^= 1;
    const int op = S->req;
    if (op < 0) break;
    if (inflight != op) {   // not the predicted op: drain the wrong prefetch, fetch the right one
      mbar_wait(bar_w, phase_w);
      phase_w ^= 1;
      const TcOp& o = TP.op[op];
      const unsigned bytes = o.img_bytes * (PASSES == 3 ? 2 : 1);
      mbar_expect_tx(bar_w, bytes);
      bulk_g2s(wbuf, image + o.img_off, bytes, bar_w);
      inflight = op;
    }
    mbar_wait(bar_w, phase_w);
    phase_w ^= 1;
    tc_fence_after();
    const TcOp& o = TP.op[op];
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(o.N >> 3) << 17) | ((128u >> 4) << 24);
    const int n_ks = o.K >> 3;
    const unsigned b_kb_bytes = o.N * 128;
    for (int ks = 0; ks < n_ks; ++ks) {
      const unsigned aoff = (ks >> 2) * A_KB_BYTES + (ks & 3) * 32;
      const unsigned boff = (ks >> 2) * b_kb_bytes + (ks & 3) * 32;
      mma_tf32(tmem_d, smem_desc(a0 + aoff), smem_desc(wbuf + boff), idesc, ks > 0 ? 1u : 0u);
      if (PASSES == 3) {
        if (!o.a_exact) mma_tf32(tmem_d, smem_desc(a1 + aoff), smem_desc(wbuf + boff), idesc, 1u);
        mma_tf32(tmem_d, smem_desc(a0 + aoff), smem_desc(wbuf + o.img_bytes + boff), idesc, 1u);
      }
    }
    mma_commit(bar_d);
    mbar_wait(bar_d, phase_d);   // the weight slot is free once the MMAs have completed
    phase_d ^= 1;
    const int next = (op + 1 == TP.n_ops) ? 0 : op + 1;
    const TcOp& no = TP.op[next];
    const unsigned nbytes = no.img_bytes * (PASSES == 3 ? 2 : 1);
    mbar_expect_tx(bar_w, nbytes);
    bulk_g2s(wbuf, image + no.img_off, nbytes, bar_w);
    inflight = next;
  }
  mbar_wait(bar_w, phase_w);   // never leave with a bulk copy in flight
}

// Greedy tile construction by the 128 epilogue threads (same packing as build_tile in pmt_tile.cuh).
__device__ __forceinline__ int build_tile_128(const PmtBatch& batch, int v_cur, int v_end, long long total_ref, Shared* S) {
  const int tid = threadIdx.x;
  const long long r_base = __ldg(batch.ref_off + v_cur), a_base = __ldg(batch.alt_off + v_cur);
  int fits = 0;
  if (v_cur + tid + 1 <= v_end) {
    const long long nr = __ldg(batch.ref_off + v_cur + tid + 1) - r_base;
    const long long na = __ldg(batch.alt_off + v_cur + tid + 1) - a_base;
    fits = (((nr + 3) & ~3LL) + na <= TILE) ? 1 : 0;
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, fits);
  epi_barrier();   // previous tile's readers of S are done
  if ((tid & 31) == 0) S->warp_cnt[tid >> 5] = __popc(ballot);
  S->rowvar[tid] = -1;
  S->rowidx[tid] = -1;
  epi_barrier();
  const int nv = S->warp_cnt[0] + S->warp_cnt[1] + S->warp_cnt[2] + S->warp_cnt[3];
  if (nv == 0) return 0;
  const long long nr_tot = __ldg(batch.ref_off + v_cur + nv) - r_base;
  const long long na_tot = __ldg(batch.alt_off + v_cur + nv) - a_base;
  const int ref_pad = (int)((nr_tot + 3) & ~3LL);
  if (tid < nv) {
    const long long r0 = __ldg(batch.ref_off + v_cur + tid), r1 = __ldg(batch.ref_off + v_cur + tid + 1);
    const long long a0 = __ldg(batch.alt_off + v_cur + tid), a1 = __ldg(batch.alt_off + v_cur + tid + 1);
    const int rs = (int)(r0 - r_base), rc = (int)(r1 - r0), as = ref_pad + (int)(a0 - a_base), ac = (int)(a1 - a0);
    S->ref_start[tid] = rs; S->ref_cnt[tid] = rc; S->alt_start[tid] = as; S->alt_cnt[tid] = ac;
    for (int i = 0; i < rc; ++i) { S->rowvar[rs + i] = tid; S->rowidx[rs + i] = r0 + i; }
    for (int i = 0; i < ac; ++i) { S->rowvar[as + i] = tid; S->rowidx[as + i] = total_ref + a0 + i; }
  }
  if (tid == 0) { S->nv = nv; S->v0 = v_cur; S->ref_pad = ref_pad; S->rows = ref_pad + (int)na_tot; }
  epi_barrier();
  return nv;
}

// Supported shape envelope of this kernel (register-resident rows): checked on the host.
constexpr int MAXW = 64;    // widest layer / d_model
constexpr int MAXH = 16;    // d_ffn / 2
constexpr int MAXE = 16;    // final feature dimension
constexpr int MAXK = 6;     // artifact clusters

template <int PASSES>
__global__ void __launch_bounds__(THREADS, 1)
reads_forward_tc_kernel(const __grid_constant__ Plan P, const __grid_constant__ TcPlan TP, const __grid_constant__ TcArgs A) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const PmtModelDesc& D = P.d;
  // carve: [A hi 32 KB][A lo 32 KB (3-pass)][weights stage][exchange floats][sums][Shared]
  unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const unsigned a0 = smem_addr(p); p += 2 * A_KB_BYTES;
  unsigned a1 = a0;
  if (PASSES == 3) { a1 = smem_addr(p); p += 2 * A_KB_BYTES; }
  const unsigned wbuf = smem_addr(p); p += (PASSES == 3 ? TP.stage_bytes : TP.stage_bytes / 2);
  float* xch = reinterpret_cast<float*>(p); p += 32 * TILE * sizeof(float);    // [32][TILE] per-row values to be summed
  float* sums = reinterpret_cast<float*>(p); p += TILE * 2 * MAXH * sizeof(float);  // [nv][2][MAXH] mean fields
  HeadConst* HC = reinterpret_cast<HeadConst*>(p); p += sizeof(HeadConst);
  Shared* S = reinterpret_cast<Shared*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));

  const int tid = threadIdx.x, warp = tid >> 5;
  const float* W = A.wflat;
  if (tid == 0) {
    mbar_init(smem_addr(&S->bar_a), EPI_THREADS);
    mbar_init(smem_addr(&S->bar_w), 1);
    mbar_init(smem_addr(&S->bar_d), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    head_constants(D, W, HC);
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&S->tmem_base)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 4) {
    if ((tid & 31) == 0) control_loop<PASSES>(TP, S, A.image, wbuf, a0, a1);
    __syncwarp();
  } else {
    // ===================================== epilogue: thread = row =====================================
    const int row = tid;
    const int E = D.d_feat, K = D.n_clusters, Dm = D.d_model, H = D.d_ffn / 2, DR = D.d_read, F = D.n_read_features;
    const int B = A.batch.n_variants;
    Epi<PASSES> ep;
    ep.S = S; ep.a0 = a0; ep.a1 = a1; ep.bar_a = smem_addr(&S->bar_a); ep.bar_d = smem_addr(&S->bar_d);
    ep.tmem_row = S->tmem_base + ((unsigned)(warp * 32) << 16);
    ep.phase_d = 0; ep.row = row; ep.W = W;
    const long long total_ref = __ldg(A.batch.ref_off + B);
    float x[MAXW];   // residual stream of this read
    float t[MAXW];   // temporary

    for (int c = blockIdx.x; c < A.n_claims; c += gridDim.x) {
      const long long cv0 = (long long)c * A.claim;
      const int cv1 = (int)min((long long)B, cv0 + A.claim);
      int v_cur = (int)cv0;
      while (v_cur < cv1) {
        const int nv = build_tile_128(A.batch, v_cur, cv1, total_ref, S);
        if (nv == 0) { v_cur += 1; continue; }   // longer than a tile: reads_forward_long_kernel
        v_cur += nv;
        const int ref_pad = S->ref_pad;
        const bool is_alt = row >= ref_pad;
        const int my_var = S->rowvar[row];
        const long long my_idx = S->rowidx[row];

        // ---- decode (batch.py:51-56): bits and wrapped quantised floats are exact in TF32 ----
        {
#pragma unroll
          for (int i = 0; i < MAXW; ++i) t[i] = 0.f;
          if (my_idx >= 0) {
            const long long src = A.batch.read_indices ? __ldg(A.batch.read_indices + my_idx) : my_idx;
            if (A.batch.reads_kind == PMT_READS_U8) {
              const int rb = D.read_row_bytes;
              const unsigned char* rp = reinterpret_cast<const unsigned char*>(A.batch.reads) + src * rb;
#pragma unroll
              for (int b = 0; b < 7; ++b) {
                const unsigned byte = __ldg(rp + b);
#pragma unroll
                for (int bit = 0; bit < 8; ++bit) t[b * 8 + bit] = (float)((byte >> (7 - bit)) & 1u);
              }
#pragma unroll
              for (int b = 7; b < 15; ++b)
                if (b < rb) t[56 + b - 7] = (float)((__ldg(rp + b) + 128u) & 255u) * 0.03125f;
            } else {
#pragma unroll
              for (int f = 0; f < MAXW; ++f)
                if (f < F)
                  t[f] = A.batch.reads_kind == PMT_READS_F16
                             ? __half2float(reinterpret_cast<const __half*>(A.batch.reads)[src * F + f])
                             : reinterpret_cast<const float*>(A.batch.reads)[src * F + f];
            }
          }
        }
        // ---- MLP programs (mlp.py): a tiny interpreter over the op list; vectors stay in registers ----
        auto run_program = [&](const PmtLinearOp* ops, int n_ops, int tc0, float* cur /* in/out */, float* tmp) {
          // cur holds the program input; on return cur holds the output
          for (int i = 0; i < n_ops; ++i) {
            const PmtLinearOp& lop = ops[i];
            const TcOp& o = TP.op[tc0 + i];
            if (lop.flags & PMT_OP_SKIP_BEGIN) {
              // residual block: cur + alpha * g(cur); layers i .. j1
              int j1 = i;
              while (!(ops[j1].flags & PMT_OP_SKIP_END)) ++j1;
#pragma unroll
              for (int k = 0; k < MAXW; ++k) tmp[k] = selu(cur[k]);
              for (int j = i; j <= j1; ++j) {
                const TcOp& oj = TP.op[tc0 + j];
                write_a_row<MAXW, PASSES>(tmp, ops[j].in_dim, oj.K >> 2, row, a0, a1);
                ep.run_mma(tc0 + j);
                ep.load_cols(0, oj.N, tmp);
                if (j < j1) {
#pragma unroll
                  for (int k = 0; k < MAXW; ++k)
                    tmp[k] = k < ops[j].out_dim ? selu(tmp[k] + __ldg(W + ops[j].b_off + min(k, ops[j].out_dim - 1))) : 0.f;
                }
              }
              const float alpha = __ldg(W + ops[j1].alpha_off);
#pragma unroll
              for (int k = 0; k < MAXW; ++k)
                if (k < ops[j1].out_dim) cur[k] = fmaf(alpha, tmp[k] + __ldg(W + ops[j1].b_off + k), cur[k]);
              i = j1;
            } else {
              write_a_row<MAXW, PASSES>(cur, lop.in_dim, o.K >> 2, row, a0, a1);
              ep.run_mma(tc0 + i);
              ep.load_cols(0, o.N, cur);
#pragma unroll
              for (int k = 0; k < MAXW; ++k) {
                if (k < lop.out_dim) {
                  const float v = cur[k] + __ldg(W + lop.b_off + k);
                  cur[k] = (lop.flags & PMT_OP_POST_SELU) ? selu(v) : v;
                } else {
                  cur[k] = 0.f;
                }
              }
            }
          }
        };
        run_program(D.read_ops, D.n_read_ops, TP.read0, t, x);
        // ---- concat (artifact_model.py:246-251) ----
        {
          const int w = D.d_info + D.d_seq;
          const float* src = my_var >= 0 ? A.out.info_seq_be + (long long)(S->v0 + my_var) * w : nullptr;
#pragma unroll
          for (int k = 0; k < MAXW; ++k) {
            if (k < DR) x[k] = t[k];
            else if (k < Dm) x[k] = src ? __ldg(src + (k - DR)) : 0.f;
            else x[k] = 0.f;
          }
        }
        // ---- gated blocks (gated_mlp.py:177-251) ----
        for (int blk = 0; blk < D.n_blocks; ++blk) {
          const PmtBlockOffsets& BO = D.blocks[blk];
          const int op1 = TP.blk0 + 2 * blk, op2 = op1 + 1;
          {
            float mean = 0.f;
#pragma unroll
            for (int k = 0; k < MAXW; ++k) if (k < Dm) mean += x[k];
            mean /= Dm;
            float var = 0.f;
#pragma unroll
            for (int k = 0; k < MAXW; ++k) if (k < Dm) { const float d = x[k] - mean; var = fmaf(d, d, var); }
            const float rstd = rsqrtf(var / Dm + LN_EPS);
#pragma unroll
            for (int k = 0; k < MAXW; ++k) t[k] = k < Dm ? (x[k] - mean) * rstd * __ldg(W + BO.ln_w + k) + __ldg(W + BO.ln_b + k) : 0.f;
          }
          write_a_row<MAXW, PASSES>(t, Dm, TP.op[op1].K >> 2, row, a0, a1);
          ep.run_mma(op1);
          float z[2 * MAXH];
          {
            const TcOp& o = TP.op[op1];
            float zr[2 * MAXH], za[2 * MAXH];
            ep.load_cols(0, 2 * MAXH, zr);
            ep.load_cols(o.Np, 2 * MAXH, za);
#pragma unroll
            for (int k = 0; k < 2 * MAXH; ++k) {
              const float v = (is_alt ? za[k] : zr[k]) + (k < 2 * H ? __ldg(W + (is_alt ? BO.p1_alt_b : BO.p1_ref_b) + k) : 0.f);
              z[k] = k < 2 * H ? selu(v) : 0.f;
            }
          }
          float z2n[MAXH];
          {
            float mean = 0.f;
#pragma unroll
            for (int k = 0; k < MAXH; ++k) if (k < H) mean += z[H + k];
            mean /= H;
            float var = 0.f;
#pragma unroll
            for (int k = 0; k < MAXH; ++k) if (k < H) { const float d = z[H + k] - mean; var = fmaf(d, d, var); }
            const float rstd = rsqrtf(var / H + LN_EPS);
#pragma unroll
            for (int k = 0; k < MAXH; ++k)
              if (k < H) {
                z2n[k] = (z[H + k] - mean) * rstd * __ldg(W + BO.ln2_w + k) + __ldg(W + BO.ln2_b + k);
                xch[k * TILE + row] = z2n[k];
              }
          }
          epi_barrier();
          {  // per-variant mean fields (gated_mlp.py:236-239)
            const float regw = __ldg(W + BO.reg_weight) + 0.25f;
            for (int idx = tid; idx < S->nv * 2 * H; idx += EPI_THREADS) {
              const int j = idx / (2 * H), s = (idx / H) & 1, f = idx % H;
              const int start = s ? S->alt_start[j] : S->ref_start[j], cnt = s ? S->alt_cnt[j] : S->ref_cnt[j];
              float sum = 0.f;
              for (int i = 0; i < cnt; ++i) sum += xch[f * TILE + start + i];
              sums[(j * 2 + s) * MAXH + f] = s == 0 ? (sum + regw * __ldg(W + BO.regularizer + f)) / ((float)cnt + regw)
                                                  : sum / ((float)cnt + 1e-4f);
            }
          }
          epi_barrier();
          {
            const float alpha = __ldg(W + (is_alt ? BO.alpha_alt : BO.alpha_ref));
            const float beta = __ldg(W + (is_alt ? BO.beta_alt : BO.beta_ref));
            const float gamma = __ldg(W + BO.gamma);
#pragma unroll
            for (int k = 0; k < MAXH; ++k) {
              float gate = 0.f;
              if (k < H) {
                gate = z2n[k] * alpha + 1.f;
                if (my_var >= 0) {
                  const float m_ref = sums[(my_var * 2 + 0) * MAXH + k];
                  gate = is_alt ? gate + beta * sums[(my_var * 2 + 1) * MAXH + k] + gamma * m_ref : gate + beta * m_ref;
                }
              }
              t[k] = k < H ? z[k] * gate : 0.f;
            }
          }
          write_a_row<MAXH, PASSES>(t, H, TP.op[op2].K >> 2, row, a0, a1);
          ep.run_mma(op2);
          {
            const TcOp& o = TP.op[op2];
            const int boff = is_alt ? BO.p2_alt_b : BO.p2_ref_b;
#pragma unroll
            for (int c = 0; c < MAXW / 16; ++c) {
              if (c * 16 < Dm) {
                float yr[16], ya[16];
                tmem_ld16(ep.tmem_row + c * 16, yr);
                tmem_ld16(ep.tmem_row + o.Np + c * 16, ya);
#pragma unroll
                for (int k = 0; k < 16; ++k)
                  if (c * 16 + k < Dm) x[c * 16 + k] += (is_alt ? ya[k] : yr[k]) + __ldg(W + boff + c * 16 + k);
              }
            }
          }
        }
        // ---- reducer (artifact_model.py:258-259) ----
        run_program(D.red_ops, D.n_red_ops, TP.red0, x, t);
        // ---- rotation + clustering head in registers (euclidean_transformation.py:19-20; feature_clustering.py:82-119) ----
        float f[MAXE];
#pragma unroll
        for (int i = 0; i < MAXE; ++i) {
          float a = 0.f;
          if (i < E) {
#pragma unroll
            for (int j = 0; j < MAXE; ++j)
              if (j < E) a = fmaf(__ldg(W + D.rotation + i * E + j), x[j] + __ldg(W + D.translation + j), a);
          }
          f[i] = a;
        }
        epi_barrier();   // mean-field readers of xch / sums are done
#pragma unroll
        for (int i = 0; i < MAXE; ++i) if (i < E) xch[i * TILE + row] = f[i];
        if (is_alt && my_var >= 0) {
          float q = 0.f, q2 = 0.f;
#pragma unroll
          for (int e = 0; e < MAXE; ++e)
            if (e < E) {
              const float a = f[e] / HC->sigma[e], b = f[e] / (2.f * HC->sigma[e]);
              q = fmaf(a, a, q); q2 = fmaf(b, b, q2);
            }
          xch[(MAXE + 0) * TILE + row] = HC->c_non - q / 2.f;
          xch[(MAXE + 1) * TILE + row] = HC->c_out - q2 / 2.f;
          for (int k = 0; k < K; ++k) {
            const float* u = W + D.unit_ke + k * E;
            float pr = 0.f;
#pragma unroll
            for (int e = 0; e < MAXE; ++e) if (e < E) pr = fmaf(f[e], __ldg(u + e), pr);
            float o2 = 0.f;
#pragma unroll
            for (int e = 0; e < MAXE; ++e) if (e < E) { const float d = f[e] - pr * __ldg(u + e); o2 = fmaf(d, d, o2); }
            const float dist = sqrtf(o2);
            const float ll_orth = HC->c_orth[k] - (dist * dist) / HC->two_tau2[k];
            const float ll_par = HC->log_half_lambda[k] + logerfc((HC->shift[k] - pr) / HC->sqrt2_sigma[k]) +
                                 HC->half_lambda[k] * (HC->two_mu_plus[k] - 2.f * pr);
            xch[(MAXE + 2 + k) * TILE + row] = ll_orth + ll_par;
          }
        }
        if (A.out.final_re && my_idx >= 0) {
#pragma unroll
          for (int e = 0; e < MAXE; ++e) if (e < E) A.out.final_re[my_idx * E + e] = f[e];
        }
        epi_barrier();
        // ---- per-variant sums and outputs ----
        for (int idx = tid; idx < S->nv * 2 * E; idx += EPI_THREADS) {
          const int j = idx / (2 * E), s = (idx / E) & 1, e = idx % E;
          const int start = s ? S->alt_start[j] : S->ref_start[j], cnt = s ? S->alt_cnt[j] : S->ref_cnt[j];
          float sum = 0.f;
          for (int i = 0; i < cnt; ++i) sum += xch[e * TILE + start + i];
          const long long v = S->v0 + j;
          float* dst = s ? A.out.alt_means_be : A.out.ref_means_be;
          if (dst) dst[v * E + e] = sum / ((float)cnt + 1e-4f);
        }
        if (tid < S->nv) {
          const int j = tid;
          const long long v = S->v0 + j;
          float ll[MAXK + 2];
#pragma unroll
          for (int k = 0; k < MAXK + 2; ++k) {
            float sum = 0.f;
            if (k < K + 2)
              for (int i = 0; i < S->alt_cnt[j]; ++i) sum += xch[(MAXE + k) * TILE + S->alt_start[j] + i];
            ll[k] = sum;
          }
          float art_max = -INFINITY;
#pragma unroll
          for (int k = 0; k < MAXK; ++k) if (k < K) { ll[2 + k] += HC->logw[k]; art_max = fmaxf(art_max, ll[2 + k]); }
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < MAXK; ++k) if (k < K) s += expf(ll[2 + k] - art_max);
          const float art = art_max + logf(s);
          if (A.out.logits_bk) {
#pragma unroll
            for (int k = 0; k < MAXK + 2; ++k) if (k < K + 2) A.out.logits_bk[v * (K + 2) + k] = ll[k];
          }
          if (A.out.logits_b) A.out.logits_b[v] = 20.f * tanhf((art - ll[0]) / 20.f);
          if (A.out.outlier_logits_b) A.out.outlier_logits_b[v] = ll[1] - logsumexp2(ll[0], art);
        }
      }
    }
    // tell the control warp we are done
    tc_fence_before();
    if (row == 0) S->req = -1;
    mbar_arrive(ep.bar_a);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(S->tmem_base), "r"(TMEM_COLS));
  }
}

// Packs one TC op: hi and lo images, K-major, 128B swizzle, K blocks of 32 elements, N rows per block.
__global__ void pack_tc_kernel(const __grid_constant__ TcPlan TP, const float* __restrict__ w, unsigned char* __restrict__ image) {
  const TcOp& o = TP.op[blockIdx.x];
  const int n_kb = (o.K * 4 + 127) / 128;
  for (int idx = threadIdx.x; idx < n_kb * o.N * 32; idx += blockDim.x) {
    const int kb = idx / (o.N * 32), rem = idx % (o.N * 32), n = rem / 32, kk = rem % 32;
    const int k = kb * 32 + kk;
    const int set = n / o.Np, nn = n % o.Np;
    float v = 0.f;
    if (nn < o.n_out && k < o.k_real && (set == 0 || o.dual)) v = w[(set ? o.w_alt_off : o.w_off) + nn * o.k_real + k];
    const float hi = tf32_round(v);
    const float lo = tf32_round(v - hi);
    const unsigned L = (unsigned)n * 128u + (unsigned)kk * 4u;
    const unsigned phys = L ^ (((L >> 7) & 7u) << 4);
    const size_t off = (size_t)o.img_off + (size_t)kb * o.N * 128 + phys;
    *reinterpret_cast<float*>(image + off) = hi;
    *reinterpret_cast<float*>(image + off + o.img_bytes) = lo;
  }
}

}  // namespace tc
}  // namespace pmt

// ================================================================================================
// host side
// ================================================================================================
using namespace pmt;
using namespace pmt::tc;

static int g_precision = PMT_PRECISION_FP32;
extern "C" int pmt_set_precision(int mode) {
  PMT_CHECK(mode == PMT_PRECISION_FP32 || mode == PMT_PRECISION_TF32X3 || mode == PMT_PRECISION_TF32,
            "unknown precision mode %d", mode);
  g_precision = mode;
  return 0;
}
extern "C" int pmt_get_precision(void) { return g_precision; }
int pmt_precision_mode() { return g_precision; }

static int pad_to(int v, int m) { return (v + m - 1) / m * m; }

bool pmt_tc_supported(const Plan& P) {
  const PmtModelDesc& d = P.d;
  if (d.d_model > MAXW || d.n_read_features > MAXW || d.d_ffn / 2 > MAXH || d.d_feat > MAXE || d.n_clusters > MAXK ||
      d.read_row_bytes > 15)
    return false;
  for (int i = 0; i < d.n_read_ops; ++i) if (d.read_ops[i].in_dim > MAXW || d.read_ops[i].out_dim > MAXW) return false;
  for (int i = 0; i < d.n_red_ops; ++i) if (d.red_ops[i].in_dim > MAXW || d.red_ops[i].out_dim > MAXW) return false;
  return 2 + d.n_read_ops + 2 * d.n_blocks + d.n_red_ops <= MAX_TC_OPS;
}

static void add_tc_op(TcPlan& T, int k_real, int n_out, int w, int b, int w_alt, int b_alt) {
  TcOp& o = T.op[T.n_ops++];
  memset(&o, 0, sizeof(o));
  o.k_real = k_real; o.n_out = n_out;
  o.K = pad_to(k_real, 8);
  o.Np = pad_to(n_out, 16);
  o.dual = w_alt >= 0;
  o.N = o.dual ? 2 * o.Np : o.Np;
  o.w_off = w; o.b_off = b; o.w_alt_off = w_alt; o.b_alt_off = b_alt;
  T.image_bytes = pad_to(T.image_bytes, 1024);
  o.img_off = T.image_bytes;
  o.img_bytes = ((o.K * 4 + 127) / 128) * o.N * 128;
  T.image_bytes += 2 * o.img_bytes;
  if (2 * o.img_bytes > T.stage_bytes) T.stage_bytes = 2 * o.img_bytes;
}

void pmt_tc_plan(const Plan& P, const PmtBatch* batch, TcPlan* out) {
  TcPlan& T = *out;
  memset(&T, 0, sizeof(T));
  const PmtModelDesc& d = P.d;
  T.read0 = T.n_ops;
  for (int i = 0; i < d.n_read_ops; ++i) add_tc_op(T, d.read_ops[i].in_dim, d.read_ops[i].out_dim, d.read_ops[i].w_off, d.read_ops[i].b_off, -1, -1);
  T.op[T.read0].a_exact = batch && batch->reads_kind == PMT_READS_U8;
  T.blk0 = T.n_ops;
  for (int b = 0; b < d.n_blocks; ++b) {
    const PmtBlockOffsets& o = d.blocks[b];
    add_tc_op(T, d.d_model, d.d_ffn, o.p1_ref_w, o.p1_ref_b, o.p1_alt_w, o.p1_alt_b);
    add_tc_op(T, d.d_ffn / 2, d.d_model, o.p2_ref_w, o.p2_ref_b, o.p2_alt_w, o.p2_alt_b);
  }
  T.red0 = T.n_ops;
  for (int i = 0; i < d.n_red_ops; ++i) add_tc_op(T, d.red_ops[i].in_dim, d.red_ops[i].out_dim, d.red_ops[i].w_off, d.red_ops[i].b_off, -1, -1);
  T.image_bytes = pad_to(T.image_bytes, 1024);
}

size_t pmt_tc_image_bytes(const Plan& P) {
  TcPlan T;
  pmt_tc_plan(P, nullptr, &T);
  return (size_t)T.image_bytes + 2048;
}

template <int PASSES>
static int launch_tc(const Plan& P, const TcPlan& T, const TcArgs& A, int grid, cudaStream_t st) {
  const size_t smem = (size_t)(PASSES == 3 ? 4 : 2) * A_KB_BYTES + (PASSES == 3 ? T.stage_bytes : T.stage_bytes / 2) +
                      32 * TILE * sizeof(float) + TILE * 2 * MAXH * sizeof(float) + sizeof(HeadConst) + sizeof(Shared) + 1024 + 64;
  PMT_CHECK(smem <= 227 * 1024, "tensor-core forward needs %zu bytes of shared memory", smem);
  cudaFuncSetAttribute(reads_forward_tc_kernel<PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  reads_forward_tc_kernel<PASSES><<<grid, THREADS, smem, st>>>(P, T, A);
  return 0;
}

// Launches the tensor-core read kernel.  `tc_image` is a device buffer of pmt_tc_image_bytes(P) bytes.
int pmt_launch_reads_tc(const Plan& P, const float* weights, const PmtBatch* batch, const PmtOutputs* out,
                        unsigned char* tc_image, int n_sm, int mode, cudaStream_t st) {
  TcPlan T;
  pmt_tc_plan(P, batch, &T);
  unsigned char* image = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_image) + 1023) & ~uintptr_t(1023));
  pack_tc_kernel<<<T.n_ops, 256, 0, st>>>(T, weights, image);
  TcArgs A;
  A.wflat = weights; A.image = image; A.batch = *batch; A.out = *out;
  const double avg = (double)(batch->n_rows > 0 ? batch->n_rows : 16LL * batch->n_variants) / batch->n_variants;
  int claim = (int)(8.0 * TILE / (avg + 1.0));
  const int ctas_per_sm = mode == PMT_PRECISION_TF32 ? 2 : 1;
  if (claim > batch->n_variants / (2 * n_sm * ctas_per_sm)) claim = batch->n_variants / (2 * n_sm * ctas_per_sm);
  if (claim < 1) claim = 1;
  if (claim > 512) claim = 512;
  A.claim = claim;
  A.n_claims = (batch->n_variants + claim - 1) / claim;
  int grid = n_sm * ctas_per_sm;
  if (grid > A.n_claims) grid = A.n_claims;
  pmt_profile_begin(st);
  const int rc = mode == PMT_PRECISION_TF32 ? launch_tc<1>(P, T, A, grid, st) : launch_tc<3>(P, T, A, grid, st);
  pmt_profile_end(st);
  return rc;
}
