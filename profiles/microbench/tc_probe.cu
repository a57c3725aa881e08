// Hardware probe for the design of the tcgen05 read kernel (sm_100a).  Stand-alone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_probe tc_probe.cu && ./tc_probe
// Answers: (1) is kind::tf32 with the A operand in TMEM (written thread-per-row with tcgen05.st 32x32b) correct,
// (2) cycles per tcgen05.mma for small N, A from shared memory vs TMEM, (3) tcgen05.ld / tcgen05.st throughput,
// (4) packed fp32x2 arithmetic throughput, (5) the epilogue <-> MMA hand-off round trip.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_addr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t smem_desc(unsigned addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ unsigned make_idesc(int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void mma_ss(unsigned d, uint64_t a, uint64_t b, unsigned idesc, unsigned acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(unsigned d, unsigned a_tmem, uint64_t b, unsigned idesc, unsigned acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

#define LD32(taddr, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
    "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr))
#define ST32(taddr, r) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};" \
  :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), \
    "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr) : "memory")
#define WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")

struct Sh {
  unsigned long long bar_a, bar_d;
  unsigned tmem_base;
};

// ------------------------------------------------------------------------------------------------
// (1) correctness: D[128 x N] = A[128 x K] * B[N x K]^T, A in TMEM (ts=1) or in shared memory (ts=0)
// B image: K-major, 128B swizzle, K blocks of 32 elements (host packed).  A smem: same layout with 128 rows.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160, 1) probe_correct(const float* __restrict__ A, const float* __restrict__ Bimg, float* __restrict__ D,
                                                       int N, int K, int ts, int a_col) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* a_s = reinterpret_cast<float*>(p); p += 2 * 128 * 128;           // up to K = 64
  float* b_s = reinterpret_cast<float*>(p); p += 2 * 128 * 128;           // up to N = 128, K = 64
  Sh* S = reinterpret_cast<Sh*>(p);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(smem_addr(&S->bar_a), 128); mbar_init(smem_addr(&S->bar_d), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&S->tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  const int n_kb = (K + 31) / 32;
  for (int i = tid; i < n_kb * N * 32; i += blockDim.x) b_s[i] = Bimg[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tm = S->tmem_base;
  if (warp < 4) {
    const int row = tid;
    const unsigned trow = tm + ((unsigned)(warp * 32) << 16);
    if (ts) {
      unsigned r[32];
      for (int c0 = 0; c0 < K; c0 += 32) {
        for (int j = 0; j < 32; ++j) r[j] = (c0 + j < K) ? __float_as_uint(A[row * K + c0 + j]) : 0u;
        ST32(trow + a_col + c0, r);
      }
      WAIT_ST();
    } else {
      for (int c = 0; c < K / 4; ++c) {
        const float4 v = make_float4(A[row * K + c * 4], A[row * K + c * 4 + 1], A[row * K + c * 4 + 2], A[row * K + c * 4 + 3]);
        const unsigned off = (c >> 3) * (128 * 128) + row * 128 + ((((unsigned)c & 7u) ^ ((unsigned)row & 7u)) << 4);
        *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(a_s) + off) = v;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    mbar_arrive(smem_addr(&S->bar_a));
    mbar_wait(smem_addr(&S->bar_d), 0);
    tc_fence_after();
    unsigned r[32];
    for (int c0 = 0; c0 < N; c0 += 32) {
      LD32(trow + c0, r);
      WAIT_LD();
      for (int j = 0; j < 32; ++j) if (c0 + j < N) D[row * N + c0 + j] = __uint_as_float(r[j]);
    }
  } else if (tid == 128) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_wait(smem_addr(&S->bar_a), 0);
    tc_fence_after();
    const unsigned idesc = make_idesc(N);
    for (int ks = 0; ks < K / 8; ++ks) {
      const unsigned boff = (ks >> 2) * (N * 128) + (ks & 3) * 32;
      if (ts) mma_ts(tm, tm + a_col + ks * 8, smem_desc(smem_addr(b_s) + boff), idesc, ks > 0);
      else mma_ss(tm, smem_desc(smem_addr(a_s) + (ks >> 2) * (128 * 128) + (ks & 3) * 32), smem_desc(smem_addr(b_s) + boff), idesc, ks > 0);
    }
    mma_commit(smem_addr(&S->bar_d));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

// ------------------------------------------------------------------------------------------------
// (2) MMA rate: chain of n MMAs into one accumulator (operands uninitialised -- timing only)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160, 1) probe_mma_rate(long long* out, int N, int n, int ts) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* a_s = reinterpret_cast<float*>(p); p += 2 * 128 * 128;
  float* b_s = reinterpret_cast<float*>(p); p += 2 * 128 * 128;
  Sh* S = reinterpret_cast<Sh*>(p);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 128 * 32; i += blockDim.x) { a_s[i] = 0.f; b_s[i] = 0.f; }
  if (tid == 0) { mbar_init(smem_addr(&S->bar_d), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&S->tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tm = S->tmem_base;
  if (tid == 128) {
    const unsigned idesc = make_idesc(N);
    unsigned phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < n; ++i) {
        const int ks = i & 7;
        const unsigned boff = (ks >> 2) * (N * 128) + (ks & 3) * 32;
        if (ts) mma_ts(tm, tm + 256 + ks * 8, smem_desc(smem_addr(b_s) + boff), idesc, i > 0);
        else mma_ss(tm, smem_desc(smem_addr(a_s) + (ks >> 2) * (128 * 128) + (ks & 3) * 32), smem_desc(smem_addr(b_s) + boff), idesc, i > 0);
      }
      const long long t1 = clock64();
      mma_commit(smem_addr(&S->bar_d));
      mbar_wait(smem_addr(&S->bar_d), phase);
      phase ^= 1;
      const long long t2 = clock64();
      out[rep * 2] = t1 - t0;
      out[rep * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

// ------------------------------------------------------------------------------------------------
// (3) tcgen05.ld / st throughput: nw warps (4 or 8), each `iters` x32 transfers
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(288, 1) probe_ldst(long long* out, unsigned* sink, int nw, int iters, int batch) {
  __shared__ Sh S;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&S.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tm = S.tmem_base;
  const unsigned trow = tm + ((unsigned)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  unsigned r[32];
  for (int j = 0; j < 32; ++j) r[j] = tid + j;
  unsigned acc = 0;
  if (warp < nw) {
    for (int c = 0; c < 256; c += 32) ST32(trow + c, r);
    WAIT_ST();
  }
  __syncthreads();
  long long t0 = clock64();
  if (warp < nw) {
    for (int i = 0; i < iters; i += batch) {
      for (int b = 0; b < batch; ++b) { LD32(trow + ((i + b) & 7) * 32, r); }
      WAIT_LD();
      acc += r[0] + r[31];
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (warp < nw) {
    for (int i = 0; i < iters; i += batch) {
      for (int b = 0; b < batch; ++b) { r[0] = acc + i; ST32(trow + ((i + b) & 7) * 32, r); }
      WAIT_ST();
    }
  }
  __syncthreads();
  long long t2 = clock64();
  if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t1; }
  sink[tid] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

// ------------------------------------------------------------------------------------------------
// (4) fp32x2 packed arithmetic vs scalar; MUFU.EX2
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) probe_f32x2(long long* out, float* sink, int iters) {
  const int tid = threadIdx.x;
  float a[16];
  for (int j = 0; j < 16; ++j) a[j] = tid * 0.001f + j;
  const float m = 1.0001f, c = 0.5f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], m, c);
  }
  __syncthreads();
  long long t1 = clock64();
  unsigned long long pk[8];
  for (int j = 0; j < 8; ++j) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pk[j]) : "f"(a[2 * j]), "f"(a[2 * j + 1]));
  unsigned long long mm, cc;
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
  __syncthreads();
  long long t2 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pk[j]) : "l"(mm), "l"(cc));
  }
  __syncthreads();
  long long t3 = clock64();
  float e[8];
  for (int j = 0; j < 8; ++j) e[j] = tid * 1e-3f + j * 0.1f;
  __syncthreads();
  long long t4 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[j]));
  }
  __syncthreads();
  long long t5 = clock64();
  if (tid == 0) { out[0] = t1 - t0; out[1] = t3 - t2; out[2] = t5 - t4; }
  float s = 0.f;
  for (int j = 0; j < 16; ++j) s += a[j];
  for (int j = 0; j < 8; ++j) { float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(pk[j])); s += lo + hi + e[j]; }
  sink[tid] = s;
}

// ------------------------------------------------------------------------------------------------
// (5) hand-off round trip: 128 epilogue threads arrive -> MMA thread issues one MMA chain of `nm` -> commit -> epilogue ld
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160, 1) probe_roundtrip(long long* out, unsigned* sink, int N, int nm, int iters, int ts) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* a_s = reinterpret_cast<float*>(p); p += 2 * 128 * 128;
  float* b_s = reinterpret_cast<float*>(p); p += 2 * 128 * 128;
  Sh* S = reinterpret_cast<Sh*>(p);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 128 * 32; i += blockDim.x) { a_s[i] = 0.f; b_s[i] = 0.f; }
  if (tid == 0) { mbar_init(smem_addr(&S->bar_a), 128); mbar_init(smem_addr(&S->bar_d), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&S->tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tm = S->tmem_base;
  const unsigned bar_a = smem_addr(&S->bar_a), bar_d = smem_addr(&S->bar_d);
  const long long t0 = clock64();
  if (warp < 4) {
    const unsigned trow = tm + ((unsigned)(warp * 32) << 16);
    unsigned r[32];
    unsigned acc = 0, ph = 0;
    for (int j = 0; j < 32; ++j) r[j] = tid;
    for (int it = 0; it < iters; ++it) {
      ST32(trow + 256, r);
      WAIT_ST();
      tc_fence_before();
      mbar_arrive(bar_a);
      mbar_wait(bar_d, ph);
      ph ^= 1;
      tc_fence_after();
      LD32(trow, r);
      WAIT_LD();
      acc += r[3];
    }
    sink[tid] = acc;
  } else if (tid == 128) {
    const unsigned idesc = make_idesc(N);
    unsigned ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(bar_a, ph);
      ph ^= 1;
      tc_fence_after();
      for (int i = 0; i < nm; ++i) {
        const int ks = i & 7;
        const unsigned boff = (ks >> 2) * (N * 128) + (ks & 3) * 32;
        if (ts) mma_ts(tm, tm + 256 + ks * 8, smem_desc(smem_addr(b_s) + boff), idesc, i > 0);
        else mma_ss(tm, smem_desc(smem_addr(a_s) + (ks >> 2) * (128 * 128) + (ks & 3) * 32), smem_desc(smem_addr(b_s) + boff), idesc, i > 0);
      }
      mma_commit(bar_d);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) out[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}


// (2b) MMA rate with `nacc` independent accumulators (round-robin), kind tf32 or bf16.  The issue loop is WARP-UNIFORM
// (whole warp runs it, elect.sync picks the issuing lane) so that descriptors stay in uniform registers.
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma_ss_bf16(unsigned d, uint64_t a, uint64_t b, unsigned idesc, unsigned acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_bf16(unsigned d, unsigned a_tmem, uint64_t b, unsigned idesc, unsigned acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int TS, int BF16>
__global__ void __launch_bounds__(160, 1) probe_mma_rate2(long long* out, int N, int n, int nacc) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* a_s = reinterpret_cast<float*>(p); p += 2 * 128 * 128;
  float* b_s = reinterpret_cast<float*>(p); p += 2 * 256 * 128;
  Sh* S = reinterpret_cast<Sh*>(p);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int i = tid; i < 2 * 128 * 32; i += blockDim.x) a_s[i] = 0.f;
  for (int i = tid; i < 2 * 256 * 32; i += blockDim.x) b_s[i] = 0.f;
  if (tid == 0) { mbar_init(smem_addr(&S->bar_d), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&S->tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tm = S->tmem_base;
  if (warp == 4) {
    const unsigned idesc = BF16 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24)) : make_idesc(N);
    const unsigned acc_stride = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    const unsigned a_base = smem_addr(a_s), b_base = smem_addr(b_s);
    unsigned phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      int acc_i = 0;
      for (int i = 0; i < n; ++i) {
        const int ks = i & 7;
        const unsigned boff = (ks >> 2) * (N * 128) + (ks & 3) * 32;
        const unsigned d = tm + acc_i * acc_stride;
        const unsigned a_t = tm + 448 + ks * 8;
        const uint64_t ad = smem_desc(a_base + (ks >> 2) * (128 * 128) + (ks & 3) * 32), bd = smem_desc(b_base + boff);
        if (elect_one()) {
          if (BF16) { if (TS) mma_ts_bf16(d, a_t, bd, idesc, i >= nacc); else mma_ss_bf16(d, ad, bd, idesc, i >= nacc); }
          else { if (TS) mma_ts(d, a_t, bd, idesc, i >= nacc); else mma_ss(d, ad, bd, idesc, i >= nacc); }
        }
        acc_i = (acc_i + 1 == nacc) ? 0 : acc_i + 1;
      }
      __syncwarp();
      const long long t1 = clock64();
      if (elect_one()) mma_commit(smem_addr(&S->bar_d));
      __syncwarp();
      mbar_wait(smem_addr(&S->bar_d), phase);
      phase ^= 1;
      const long long t2 = clock64();
      if ((tid & 31) == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

template <int TS, int BF16>
static void run_rate2(long long* d_out) {
  const size_t smem2 = 2 * 128 * 128 + 2 * 256 * 128 + 1024 + 256;
  CK(cudaFuncSetAttribute(probe_mma_rate2<TS, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  for (int N : {16, 32, 48, 64, 128, 256})
    for (int nacc : {1, 2, 4}) {
      if (nacc * (N <= 64 ? 64 : (N <= 128 ? 128 : 256)) > 448) continue;
      long long t32[6], t160[6];
      probe_mma_rate2<TS, BF16><<<1, 160, smem2>>>(d_out, N, 32, nacc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(t32, d_out, sizeof(t32), cudaMemcpyDeviceToHost));
      probe_mma_rate2<TS, BF16><<<1, 160, smem2>>>(d_out, N, 160, nacc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(t160, d_out, sizeof(t160), cudaMemcpyDeviceToHost));
      printf("mma2 %s M=128 N=%3d A=%s nacc=%d: %.1f cyc/mma, issue %.1f cyc/mma (chain32 %lld issue %lld, chain160 %lld)\n", BF16 ? "bf16" : "tf32", N, TS ? "TMEM" : "smem", nacc,
             (t160[5] - t32[5]) / 128.0, (t160[4] - t32[4]) / 128.0, t32[5], t32[4], t160[5]);
    }
}

// (2c) fully unrolled x3 chain: 8 k-steps x (Ahi.Bhi, Alo.Bhi, Ahi.Blo), A in TMEM, constant descriptor offsets
template <int N, int KS>
__global__ void __launch_bounds__(160, 1) probe_mma_rate3(long long* out, int reps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* b_s = reinterpret_cast<float*>(p); p += 4 * 128 * 128;
  Sh* S = reinterpret_cast<Sh*>(p);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int i = tid; i < 4 * 128 * 32; i += blockDim.x) b_s[i] = 0.f;
  if (tid == 0) { mbar_init(smem_addr(&S->bar_d), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&S->tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tm = __shfl_sync(0xffffffffu, S->tmem_base, 0);
  if (warp == 4) {
    constexpr unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t bhi = smem_desc(smem_addr(b_s)), blo = smem_desc(smem_addr(b_s) + 2 * N * 128);
    const unsigned ahi = tm + 256, alo = tm + 320;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const unsigned d = tm + (r & 1) * 64;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint64_t off = (uint64_t)(((ks >> 2) * (N * 128) + (ks & 3) * 32) >> 4);
          mma_ts(d, ahi + ks * 8, bhi + off, idesc, ks > 0);
          mma_ts(d, alo + ks * 8, bhi + off, idesc, 1);
          mma_ts(d, ahi + ks * 8, blo + off, idesc, 1);
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one()) mma_commit(smem_addr(&S->bar_d));
    __syncwarp();
    mbar_wait(smem_addr(&S->bar_d), 0);
    const long long t2 = clock64();
    if ((tid & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}
template <int N, int KS>
static void run_rate3(long long* d_out) {
  const size_t smem3 = 4 * 128 * 128 + 1024 + 256;
  CK(cudaFuncSetAttribute(probe_mma_rate3<N, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
  long long a[2], b[2];
  probe_mma_rate3<N, KS><<<1, 160, smem3>>>(d_out, 4); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(a, d_out, sizeof(a), cudaMemcpyDeviceToHost));
  probe_mma_rate3<N, KS><<<1, 160, smem3>>>(d_out, 20); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(b, d_out, sizeof(b), cudaMemcpyDeviceToHost));
  printf("mma3 unrolled x3 chain N=%3d ksteps=%d: %.1f cyc/mma complete, %.1f cyc/mma issue; one layer (%d mma) = %.0f cyc\n", N, KS,
         (b[1] - a[1]) / (16.0 * 3 * KS), (b[0] - a[0]) / (16.0 * 3 * KS), 3 * KS, (b[1] - a[1]) / 16.0);
}

static float tf32_exact(int v) { return (float)v / 8.0f; }

int main() {
  const size_t smem = 4 * 128 * 128 + 1024 + 256;
  CK(cudaFuncSetAttribute(probe_correct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(probe_mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(probe_roundtrip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long* d_out; unsigned* d_sink; float* d_fsink;
  CK(cudaMalloc(&d_out, 64 * sizeof(long long))); CK(cudaMalloc(&d_sink, 1024 * sizeof(unsigned))); CK(cudaMalloc(&d_fsink, 1024 * sizeof(float)));
  long long h[64];

  // ---- (1) correctness ----
  const int shapes[][2] = {{48, 24}, {32, 64}, {64, 64}, {16, 64}, {24, 16}, {40, 32}, {128, 16}};
  for (auto& sh : shapes) {
    const int N = sh[0], K = sh[1];
    for (int ts = 0; ts <= 1; ++ts) {
      std::vector<float> A(128 * K), B(N * K), Bimg(((K + 31) / 32) * N * 32, 0.f), D(128 * N, -777.f);
      srand(1234 + N + K);
      for (auto& v : A) v = tf32_exact(rand() % 33 - 16);
      for (auto& v : B) v = tf32_exact(rand() % 33 - 16);
      for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
        const int kb = k / 32, kk = k % 32;
        const unsigned L = (unsigned)n * 128u + (unsigned)kk * 4u;
        const unsigned phys = L ^ (((L >> 7) & 7u) << 4);
        Bimg[((size_t)kb * N * 128 + phys) / 4] = B[n * K + k];
      }
      float *dA, *dB, *dD;
      CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, Bimg.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
      CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dB, Bimg.data(), Bimg.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
      probe_correct<<<1, 160, smem>>>(dA, dB, dD, N, K, ts, 256);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("correct N=%d K=%d ts=%d: CUDA error %s\n", N, K, ts, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0; int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
        const double err = fabs(ref - D[m * N + n]); if (err > maxerr) maxerr = err; if (err > 1e-3) ++bad;
      }
      printf("correct N=%3d K=%2d A=%s: max|err|=%.3g bad=%d/%d\n", N, K, ts ? "TMEM" : "smem", maxerr, bad, 128 * N);
      cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
  }
  // precision probe: does the tensor core truncate or round fp32 inputs to tf32?  A = 1 + 2^-11 + 2^-13, B = 1
  {
    const int N = 16, K = 8;
    std::vector<float> A(128 * K, 0.f), B(N * K, 0.f), Bimg(N * 32, 0.f), D(128 * N);
    for (int m = 0; m < 128; ++m) A[m * K] = 1.f + ldexpf(1.f, -11) + ldexpf(1.f, -13) * (m & 1) + ldexpf(1.f, -10) * ((m >> 1) & 1);
    for (int n = 0; n < N; ++n) { B[n * K] = 1.f; const unsigned L = n * 128u; Bimg[(L ^ (((L >> 7) & 7u) << 4)) / 4] = 1.f; }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, Bimg.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bimg.data(), Bimg.size() * 4, cudaMemcpyHostToDevice));
    for (int ts = 0; ts <= 1; ++ts) {
      probe_correct<<<1, 160, smem>>>(dA, dB, dD, N, K, ts, 256);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      printf("rounding probe A=%s: ", ts ? "TMEM" : "smem");
      for (int m = 0; m < 4; ++m) printf("in=1+%.6g out=1+%.6g | ", (double)A[m * K] - 1.0, (double)D[m * N] - 1.0);
      printf("\n");
    }
  }

  // ---- (2) MMA rate ----
  for (int ts = 0; ts <= 1; ++ts)
    for (int N : {16, 24, 32, 48, 64, 96, 128}) {
      long long t32[6], t160[6];
      probe_mma_rate<<<1, 160, smem>>>(d_out, N, 32, ts); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(t32, d_out, sizeof(t32), cudaMemcpyDeviceToHost));
      probe_mma_rate<<<1, 160, smem>>>(d_out, N, 160, ts); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(t160, d_out, sizeof(t160), cudaMemcpyDeviceToHost));
      printf("mma tf32 M=128 N=%3d A=%s: issue %.1f cyc/mma, complete %.1f cyc/mma (chain32 total %lld, chain160 total %lld)\n", N, ts ? "TMEM" : "smem",
             (t160[4] - t32[4]) / 128.0, (t160[5] - t32[5]) / 128.0, t32[5], t160[5]);
    }


  // ---- (2b) MMA rate: independent accumulators / kinds, warp-uniform issue ----
  run_rate2<0, 0>(d_out); run_rate2<1, 0>(d_out); run_rate2<0, 1>(d_out); run_rate2<1, 1>(d_out);

  run_rate3<16, 8>(d_out); run_rate3<32, 4>(d_out); run_rate3<32, 8>(d_out); run_rate3<48, 8>(d_out); run_rate3<64, 3>(d_out); run_rate3<64, 8>(d_out); run_rate3<128, 8>(d_out);

  // ---- (3) tcgen05.ld / st ----
  for (int nw : {4, 8})
    for (int batch : {1, 2, 4}) {
      probe_ldst<<<1, 288>>>(d_out, d_sink, nw, 256, batch); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, d_out, 2 * sizeof(long long), cudaMemcpyDeviceToHost));
      const double bytes = (double)nw * 256 * 32 * 32 * 4;
      printf("tmem nw=%d batch=%d: ld %.1f cyc per x32 per warp (%.0f B/cyc/SM), st %.1f cyc per x32 per warp (%.0f B/cyc/SM)\n", nw, batch,
             h[0] / 256.0, bytes / h[0], h[1] / 256.0, bytes / h[1]);
    }

  // ---- (4) fp32x2 ----
  probe_f32x2<<<1, 512>>>(d_out, d_fsink, 256); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, d_out, 3 * sizeof(long long), cudaMemcpyDeviceToHost));
  printf("16 warps: scalar FFMA %.2f cyc/warp-instr/SMSP (%.0f flop/cyc/SM), fma.f32x2 %.2f cyc/warp-instr/SMSP (%.0f flop/cyc/SM), ex2 %.2f cyc/warp-instr/SMSP\n",
         h[0] / (256.0 * 16 * 4), 2.0 * 512 * 16 * 256 / h[0], h[1] / (256.0 * 8 * 4), 4.0 * 512 * 8 * 256 / h[1], h[2] / (256.0 * 8 * 4));

  // ---- (5) round trip ----
  for (int ts = 0; ts <= 1; ++ts)
    for (int nm : {1, 8, 24}) {
      probe_roundtrip<<<1, 160, smem>>>(d_out, d_sink, 64, nm, 200, ts); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, d_out, sizeof(long long), cudaMemcpyDeviceToHost));
      printf("round trip A=%s N=64 chain=%2d: %.0f cyc per iteration (st x32 + arrive + mma + commit + ld x32)\n", ts ? "TMEM" : "smem", nm, h[0] / 200.0);
    }
  printf("done\n");
  return 0;
}
